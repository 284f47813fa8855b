#!/usr/bin/env python
"""Sweep the packed kernel's diagnostic knobs (strip width K, rows per step R,
warps per block) over a few table sizes and print the fill-kernel time of each
combination.   python tools/sweep_pk.py [--sizes 100000,30000,10000] [--count]
The knobs are read by libnwb.so at every nwb_plan_run (NWB_PK_K / NWB_PK_R /
NWB_PK_WARPS); results are checked against the first combination's score."""
import argparse
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nw_b200 as nwb  # noqa: E402
import oracle  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="100000,30000,10000")
ap.add_argument("--ks", default="4,2,1")
ap.add_argument("--rs", default="2,1")
ap.add_argument("--warps", default="4,8")
ap.add_argument("--count", action="store_true", help="with the fused 64-bit count (K=4 only)")
ap.add_argument("--branches", action="store_true", help="with the fused branch counter")
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()

flags = 0 if args.branches else nwb.NO_BRANCH_COUNT
if args.count:
    flags |= nwb.WANT_COUNT
for n in [int(x) for x in args.sizes.split(",")]:
    protein = n == 30000
    alpha = oracle.PROTEIN if protein else oracle.DNA
    seed = 0x5EED0005 if protein else (0x5EED0030 if n == 100000 else 0x5EED0002)
    m, k, d = (2, 1, 2) if protein else (1, 1, 1)
    t, s = oracle.generate_pair(seed, n, n, alpha)
    plan = nwb.Plan(n, n, flags)
    plan.upload(t, s)
    ref = None
    for K, R, W in itertools.product([int(x) for x in args.ks.split(",")], [int(x) for x in args.rs.split(",")],
                                     [int(x) for x in args.warps.split(",")]):
        if args.count and K != 4:
            continue
        nwb.tune("pk_k", K)
        nwb.tune("pk_r", R)
        nwb.tune("pk_warps", W)
        best = 1e9
        for _ in range(args.reps):
            plan.run(m, k, d)
            sm = plan.summary()
            best = min(best, plan.kernel_ms())
        key = (sm.opt_score, sm.branch_count, sm.count)
        if ref is None:
            ref = key
        print(f"n={n} K={K} R={R} warps={W}: {best:8.3f} ms  {n * n / best / 1e6:8.1f} GCUPS  "
              f"score={sm.opt_score} {'OK' if key == ref else 'MISMATCH ' + str(key) + ' vs ' + str(ref)}", flush=True)
    plan.close()
