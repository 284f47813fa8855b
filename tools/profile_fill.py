#!/usr/bin/env python
"""Run a few fills of one synthetic pair through the C ABI and print the device
time of the fill kernel -- the command profiled under ncu (profiles/) and used
for quick A/B experiments.   python tools/profile_fill.py --a 100000 --b 100000
Environment knobs read by libnwb.so (diagnostics only): NWB_PK_K=1|2|4 forces the
packed strip width, NWB_DEBUG_NOWAIT=1 skips the inter-strip waits (wrong
results; isolates compute from synchronisation)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nw_b200 as nwb  # noqa: E402
import oracle  # noqa: E402

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
args = ap.parse_args()

alpha = oracle.DNA if args.alphabet == "dna" else oracle.PROTEIN
if args.batch:
    import numpy as np
    n = args.batch
    a, b = min(args.a, 4096), min(args.b, 4096)
    tcat = bytearray()
    scat = bytearray()
    for p in range(n):
        tt, ss = oracle.generate_pair(0x5EED4000 + 2 * p, a, b, alpha)
        tcat += tt
        scat += ss
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
t, s = oracle.generate_pair(args.seed, args.a, args.b, alpha)
plan = nwb.Plan(args.a, args.b, args.flags)
plan.upload(t, s)
for r in range(args.reps):
    plan.run(args.m, args.k, args.d)
    sm = plan.summary()
    ms = plan.kernel_ms()
    print(f"rep {r}: kind={sm.kernel_kind} kernel_ms={ms:.3f} GCUPS={args.a * args.b / ms / 1e6:.1f} "
          f"score={sm.opt_score} branches={sm.branch_count} count={sm.count}", flush=True)
plan.close()
