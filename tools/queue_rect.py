"""Queue mode on ONE GPU with a table the width of one rank's share of an 8-way strip group (49 strips x 100,000 rows):
how much do fills of different plans slow one another down when they share the SMs?  (run under gpurun)
usage: python tools/queue_rect.py [A] [B] [steps] [hx_spb] [NQ,NQ,...]"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import nw_b200 as nwb

A = int(sys.argv[1]) if len(sys.argv) > 1 else 12544
B = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
K = int(sys.argv[3]) if len(sys.argv) > 3 else 40
SPB = int(sys.argv[4]) if len(sys.argv) > 4 else 0      # nwb_tune hx_spb: adjacent strips per block (0 = 3)
NQS = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else [1, 2, 4, 8, 9, 12]
nwb.tune("hx_spb", SPB)
t, s = nwb.generate_pair(0x5EED0030, A, B)
# one fill alone, the kernel's own time (events around prep + fill inside the library): one-fill launch vs queue mode
for flags, name in ((0, "cooperative launch, strips dealt out cyclically"), (nwb.QUEUE, f"queue mode, hx_spb={SPB}")):
    nwb.tune("hx_spb", SPB if flags else 0)      # the override is read at every run: none for the one-fill launch
    pl = nwb.Plan(A, B, flags)
    pl.upload(t, s)
    ks = []
    for _ in range(4):
        pl.run(1, 1, 1)
        pl.summary()
        ks.append(round(pl.kernel_ms(), 3))
    print(f"A={A} B={B} one fill alone, {name}: kernel_ms {ks}", flush=True)
    pl.close()
nwb.tune("hx_spb", SPB)
for NQ in NQS:
    plans = [nwb.Plan(A, B, nwb.QUEUE) for _ in range(NQ)]
    streams = [torch.cuda.Stream() for _ in range(NQ)]
    for pl in plans:
        pl.upload(t, s)
    for i in range(NQ):
        plans[i].run(1, 1, 1, streams[i].cuda_stream)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(streams[0])
    for i in range(K):
        q = i % NQ
        ev[i][0].record(streams[q])
        plans[q].run(1, 1, 1, streams[q].cuda_stream)
        ev[i][1].record(streams[q])
    for q in range(1, NQ):
        streams[0].wait_stream(streams[q])
    b.record(streams[0])
    torch.cuda.synchronize()
    tot = a.elapsed_time(b)
    spans = [x.elapsed_time(y) for x, y in ev]
    print(f"A={A} B={B} spb={SPB} NQ={NQ:2d}  {tot / K:7.3f} ms per table  {A * B * K / tot / 1e6:8.1f} GCUPS   spans first {[round(x, 2) for x in spans[:NQ + 2]]} last {[round(x, 2) for x in spans[-3:]]}", flush=True)
    for pl in plans:
        pl.close()
