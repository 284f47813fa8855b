#!/usr/bin/env python
"""What the flush warps and the strip-to-strip waits cost the sweeping warps of nwb_fill_hx_kernel: the 100k x 100k fill
with the experiments build's debug bits (make -C needleman-wunsch_b200/csrc exp; results are WRONG by design, only the
time is of interest).  bit 1: strips do not wait for their left neighbour; bit 2: the flush warps do nothing."""
import sys, os
sys.path.insert(0, '/root/repo')
import importlib
mod = importlib.import_module("needleman-wunsch_b200")
mod.LIB_PATH = os.path.abspath("needleman-wunsch_b200/libnwb_exp.so")
import nw_b200 as nwb
t, s = nwb.generate_pair(0x5EED0030, 100000, 100000)
for bits in (0, 2, 1, 3):
    nwb.tune("debug_nowait", bits)
    plan = nwb.Plan(100000, 100000, 0)
    plan.upload(t, s)
    ms = []
    for _ in range(4):
        plan.run(1, 1, 1)
        try:
            plan.summary()
        except Exception as e:
            pass
        ms.append(plan.kernel_ms())
    print("debug bits", bits, "(1: no waits between strips, 2: flush warps do nothing)", "min %.3f ms" % min(ms), flush=True)
    plan.close()
