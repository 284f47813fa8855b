#!/usr/bin/env python
"""Rectangular tables A x B: what one GPU's share of a column-strip split looks like when every strip of
the share is resident at once.  Prints the fill-kernel time per (kernel, K, R, warps) for each A.
    python tools/sweep_rect.py [--tops 12500,25000] [--side 100000]
From t(A) at two widths: per-strip hop = dt / dstrips, sweep of one strip = intercept."""
import argparse
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nw_b200 as nwb  # noqa: E402
import oracle  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tops", default="12500,25000")
ap.add_argument("--side", type=int, default=100000)
ap.add_argument("--ks", default="4,2,1")
ap.add_argument("--rs", default="2,1")
ap.add_argument("--warps", default="4,8")
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()

B = args.side
for A in [int(x) for x in args.tops.split(",")]:
    t, s = oracle.generate_pair(0x5EED0030, max(A, B), max(A, B), oracle.DNA)
    t, s = t[:A], s[:B]
    plan = nwb.Plan(A, B, nwb.NO_BRANCH_COUNT)
    plan.upload(t, s)
    ref = None
    combos = [("hx", 4, 2, 0)] + [("pk", K, R, W) for K, R, W in itertools.product(
        [int(x) for x in args.ks.split(",")], [int(x) for x in args.rs.split(",")], [int(x) for x in args.warps.split(",")])]
    for kind, K, R, W in combos:
        nwb.tune("pk_hx", 1 if kind == "hx" else 0)
        nwb.tune("pk_k", K)
        nwb.tune("pk_r", R)
        nwb.tune("pk_warps", W)
        best = 1e9
        for _ in range(args.reps):
            plan.run(1, 1, 1)
            sm = plan.summary()
            best = min(best, plan.kernel_ms())
        if ref is None:
            ref = sm.opt_score
        print(f"A={A} B={B} {kind} K={K} R={R} warps={W}: {best:8.3f} ms  {A * B / best / 1e6:8.1f} GCUPS  "
              f"strips={-(-A // (64 * K))} score={sm.opt_score} {'OK' if sm.opt_score == ref else 'MISMATCH vs ' + str(ref)}",
              flush=True)
    plan.close()
