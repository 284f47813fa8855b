#!/usr/bin/env python
"""Where a one-shot nwb_fill() call spends its host time (what the CLI pays per process): CUDA context creation, the first
call (module load, allocations, zeroing) and a second call of the same shape, for the BASELINE squares.
    python tools/host_phases.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
t0 = time.perf_counter()
import nw_b200 as nwb  # noqa: E402
import oracle  # noqa: E402
t1 = time.perf_counter()
n = nwb.device_count()
t2 = time.perf_counter()
print(f"import {t1 - t0:.3f} s, device_count() = {n}: {t2 - t1:.3f} s", flush=True)
for size, alpha, seed, mkd in ((10000, oracle.DNA, 0x5EED0002, (1, 1, 1)), (30000, oracle.PROTEIN, 0x5EED0005, (2, 1, 2)),
                               (100000, oracle.DNA, 0x5EED0030, (1, 1, 1))):
    t, s = oracle.generate_pair(seed, size, size, alpha)
    for flags, what in ((0, "-q"), (nwb.WANT_COUNT, "-q -s")):
        for rep in range(2):
            a = time.perf_counter()
            tab = nwb.fill(t, s, *mkd, flags)
            b = time.perf_counter()
            score = tab.opt_score
            tab.close() if hasattr(tab, "close") else None
            c = time.perf_counter()
            print(f"n={size} {what:6s} call {rep}: fill {b - a:.3f} s, free {c - b:.3f} s, kernel {tab.kernel_ms if hasattr(tab, 'kernel_ms') else -1} score={score}", flush=True)
