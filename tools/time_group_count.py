#!/usr/bin/env python
"""`-q -s` on a strip group (nwb_fill_on over several GPUs of one process): wall time of the blocking call with the
sparse count on the last rank (default) against the group's dense sweep (nwb_tune count_mode = 2).
    python tools/time_group_count.py [--n 100000] [--gpus 2]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nw_b200 as nwb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100000)
ap.add_argument("--gpus", type=int, default=2)
args = ap.parse_args()
t, s = nwb.generate_pair(0x5EED0030, args.n, args.n)
for world in sorted({1, args.gpus}):
    for name, knobs, flags in (("fill", {}, 0), ("fill + count (sparse on the last rank)", {}, nwb.WANT_COUNT),
                               ("fill + count (dense sweep)", dict(count_mode=2), nwb.WANT_COUNT)):
        with nwb.tuned(**knobs):
            best = 1e9
            for _ in range(4):
                t0 = time.perf_counter()
                tab = nwb.fill(t, s, 1, 1, 1, flags, num_gpus=world)
                dt = (time.perf_counter() - t0) * 1e3
                best = min(best, dt)
                path = tab.summary().count_path
                res = (tab.opt_score, tab.count)
                tab.close()
        print(f"{world} GPU(s)  {name:42s} {best:8.2f} ms per call  count_path={path}  {res}", flush=True)
