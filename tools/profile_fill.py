#!/usr/bin/env python
"""Run a few fills of one synthetic pair through the C ABI and print the device
time of the fill kernel -- the command profiled under ncu (profiles/) and used
for quick A/B experiments.   python tools/profile_fill.py --a 100000 --b 100000
--tune key=value[,key=value] passes nwb_tune() overrides (include/nwb.h), e.g. pk_hx=0, count_mode=2."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nw_b200 as nwb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--a", type=int, default=100000)
ap.add_argument("--b", type=int, default=100000)
ap.add_argument("--seed", type=lambda x: int(x, 0), default=0x5EED0030)
ap.add_argument("--alphabet", default="dna")
ap.add_argument("--m", type=int, default=1)
ap.add_argument("--k", type=int, default=1)
ap.add_argument("--d", type=int, default=1)
ap.add_argument("--flags", type=lambda x: int(x, 0), default=nwb.NO_BRANCH_COUNT, help="include/nwb.h flag bits; WANT_COUNT = 0x2")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--batch", type=int, default=0, help="run a batch of this many pairs (config 4 generator) instead")
ap.add_argument("--tune", default="")
args = ap.parse_args()

for kv in filter(None, args.tune.split(",")):
    k_, v_ = kv.split("=")
    nwb.tune(k_, int(v_))
alpha = nwb.DNA if args.alphabet == "dna" else nwb.PROTEIN
if args.batch:
    import numpy as np
    n = args.batch
    a, b = min(args.a, 4096), min(args.b, 4096)
    tcat = nwb.generate(0x5EED4000, a, alpha, count=n, seed_stride=2)
    scat = nwb.generate(0x5EED4001, b, alpha, count=n, seed_stride=2)
    toff = np.arange(n + 1, dtype=np.int64) * a
    soff = np.arange(n + 1, dtype=np.int64) * b
    bt = nwb.Batch.from_arrays(bytes(tcat), toff, bytes(scat), soff, args.m, args.k, args.d, args.flags)
    for r in range(args.reps):
        bt.run()
        bt.fetch()
        ms = bt.kernel_ms()
        print(f"rep {r}: {bt.kernel_name()} batch {n} x {a}x{b} kernel_ms={ms:.3f} GCUPS={n * a * b / ms / 1e6:.1f} "
              f"score0={bt.opt_score(0)} branches0={bt.branch_count(0)}", flush=True)
    bt.close()
    sys.exit(0)
t, s = nwb.generate_pair(args.seed, args.a, args.b, alpha)
plan = nwb.Plan(args.a, args.b, args.flags)
plan.upload(t, s)
for r in range(args.reps):
    plan.run(args.m, args.k, args.d)
    sm = plan.summary()
    ms = plan.kernel_ms()
    print(f"rep {r}: kind={sm.kernel_kind} kernel_ms={ms:.3f} GCUPS={args.a * args.b / ms / 1e6:.1f} "
          f"score={sm.opt_score} branches={sm.branch_count} count={sm.count} count_path={plan.count_path()} "
          f"rows={sm.count_rows}", flush=True)
plan.close()
