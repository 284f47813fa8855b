#!/bin/bash
# Queue depth / host pacing sweep of bench.py's headline loop (run under gpurun):  NG=8 bash tools/sweep_queue.sh "40 9 1" "40 8 1"
# each argument: "<steps> <plans per GPU> <pace 0|1>"
NG=${NG:-1}
for cfg in "$@"; do set -- $cfg
if [ "$NG" = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"; fi
timeout 200 $L bench.py --gpus $NG --steps $1 --warmup 3 --no-extras --no-cpu --queue $2 --pace $3 2>gpurun_out/bs.err | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('gpus',j['n_gpus'],'steps',j['steps'],'NQ',j['queue']['plans_in_flight'],'pace','$3','value',round(j['value'],1),'ms',round(j['ms_per_step'],3),'e2e',round(j['e2e']['value'],1),'ok',j['golden_ok'],'lat',round(j['latency']['fill_latency_ms'],3))
print([round(x,1) for x in j['step_ms'][:16]])
" || tail -5 gpurun_out/bs.err
done
