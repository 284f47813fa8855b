#!/usr/bin/env python
"""Device time of the single-pair fill (and fill + count) of BASELINE configs 2, 5, 3 with a given build of the
library:   python tools/time_fill.py [path/to/libnwb.so] [--count] [--sizes 10000,30000,100000]
Every run is checked against tests/golden/golden_big.json (score, branch count, whole-table digest)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nw_b200 as nwb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("lib", nargs="?", default=None)
ap.add_argument("--count", action="store_true")
ap.add_argument("--general", action="store_true", help="NWB_FORCE_GENERAL: the int32 kernel")
ap.add_argument("--sizes", default="10000,30000,100000")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--tune", default="", help="comma-separated key=value overrides (nwb_tune)")
args = ap.parse_args()
if args.lib:
    import importlib
    mod = importlib.import_module("needleman-wunsch_b200")
    mod.LIB_PATH = os.path.abspath(args.lib)
for kv in filter(None, args.tune.split(",")):
    k, v = kv.split("=")
    nwb.tune(k, int(v))
gold = {g["name"]: g for g in json.load(open(os.path.join(ROOT, "tests", "golden", "golden_big.json")))}
CFG = {10000: ("config2_dna_10k", 0x5EED0002, nwb.DNA, (1, 1, 1)), 30000: ("config5_protein_30k", 0x5EED0005, nwb.PROTEIN, (2, 1, 2)),
       100000: ("config3_dna_100k", 0x5EED0030, nwb.DNA, (1, 1, 1))}
for n in [int(x) for x in args.sizes.split(",")]:
    name, seed, alpha, mkd = CFG[n]
    g = gold[name]
    t, s = nwb.generate_pair(seed, n, n, alpha)
    plan = nwb.Plan(n, n, (nwb.WANT_COUNT if args.count else 0) | (nwb.FORCE_GENERAL if args.general else 0))
    plan.upload(t, s)
    ms = []
    for _ in range(args.reps):
        plan.run(*mkd)
        sm = plan.summary()
        ms.append(plan.kernel_ms())
    ok = (sm.opt_score, sm.branch_count) == (g["final_score"], g["branch_count"]) and plan.arrow_digest() == int(g["arrow_digest"], 16)
    print(f"{os.path.basename(args.lib or 'libnwb.so'):18s} {name:20s} {plan.kernel_name()} {'+count(' + plan.count_path() + ')' if args.count else ''} "
          f"min {min(ms):8.3f} ms  med {sorted(ms)[len(ms) // 2]:8.3f}  {n * n / min(ms) / 1e6:8.1f} GCUPS  {'GOLDEN OK' if ok else 'MISMATCH'}",
          flush=True)
    plan.close()
