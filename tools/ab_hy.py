#!/usr/bin/env python
"""A/B of the sweeping warps' lane geometry: nwb_fill_hy.cuh (three rows of skew per lane, NWB_PK_HY=1) against the
default nwb_fill_hx.cuh (four rows); needs the experiments build (make -C needleman-wunsch_b200/csrc exp) on BASELINE configs 3, 5, 2, with and without the count behind -s.
    python tools/ab_hy.py [--reps 5]
Every run is checked against tests/golden/golden_big.json (score, branch counter)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nw_b200 as nwb  # noqa: E402
nwb.use_experiments_build()
import oracle  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--sizes", default="100000,30000,10000")
args = ap.parse_args()
golden = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_big.json")))
gold = {g["top_len"]: g for g in golden}

for n in [int(x) for x in args.sizes.split(",")]:
    protein = n == 30000
    alpha = oracle.PROTEIN if protein else oracle.DNA
    seed = 0x5EED0005 if protein else (0x5EED0030 if n == 100000 else 0x5EED0002)
    m, k, d = (2, 1, 2) if protein else (1, 1, 1)
    t, s = oracle.generate_pair(seed, n, n, alpha)
    for flags, what in ((0, "fill"), (nwb.WANT_COUNT, "fill+count")):
        res = {}
        for hy in ("0", "1"):
            nwb.tune("pk_hy", int(hy))
            plan = nwb.Plan(n, n, flags)
            plan.upload(t, s)
            ms = []
            for _ in range(args.reps):
                plan.run(m, k, d)
                sm = plan.summary()
                ms.append(plan.kernel_ms())
            res[hy] = (min(ms), sm.opt_score, sm.branch_count, sm.count, plan.kernel_name())
            plan.close()
        same = res["0"][1:4] == res["1"][1:4] and res["1"][1:3] == (gold[n]["final_score"], gold[n]["branch_count"])
        print(f"n={n} {what:10s} hx {res['0'][0]:8.3f} ms  hy {res['1'][0]:8.3f} ms  ({res['1'][4]})  "
              f"x{res['0'][0] / res['1'][0]:.3f}  hy {n * n / res['1'][0] / 1e6:8.1f} GCUPS  "
              f"score={res['1'][1]} branches={res['1'][2]} {'GOLDEN OK' if same else 'MISMATCH ' + str(res)}", flush=True)
