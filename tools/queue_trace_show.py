import json,sys
j=json.load(open('gpurun_out/queue_trace_n%s.json'%sys.argv[1]))
T=j['by_rank_device_start_end_and_host_sync_begin_end_launched']
for r,(tl,ht) in enumerate(T):
    print('rank',r)
    for i in range(int(sys.argv[2])): print('   step',i,'dev',tl[i],'host',ht[i])
