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
t1 = time.perf_counter()
n = nwb.device_count()
t2 = time.perf_counter()
print(f"import {t1 - t0:.3f} s, device_count() = {n}: {t2 - t1:.3f} s", flush=True)
for size, alpha, seed, mkd in ((10000, nwb.DNA, 0x5EED0002, (1, 1, 1)), (30000, nwb.PROTEIN, 0x5EED0005, (2, 1, 2)),
                               (100000, nwb.DNA, 0x5EED0030, (1, 1, 1))):
    t, s = nwb.generate_pair(seed, size, size, alpha)
    for flags, what in ((0, "-q"), (nwb.WANT_COUNT, "-q -s")):
        for rep in range(4):   # call 0 creates the cached workspace; calls 1.. reuse it (nwb_fill's plan cache)
            a = time.perf_counter()
            tab = nwb.fill(t, s, *mkd, flags)
            b = time.perf_counter()
            score = tab.opt_score
            tab.close() if hasattr(tab, "close") else None
            c = time.perf_counter()
            print(f"n={size} {what:6s} call {rep}: nwb_fill {1e3 * (b - a):8.3f} ms, nwb_free {1e3 * (c - b):6.3f} ms, kernel {tab.kernel_ms:7.3f} ms "
                  f"-> host overhead {1e3 * (b - a) - tab.kernel_ms:6.3f} ms  score={score}", flush=True)
    nwb.cache_clear()
