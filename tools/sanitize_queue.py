#!/usr/bin/env python
"""Queue mode (NWB_QUEUE) for `compute-sanitizer --tool memcheck` (closed on this round's GPU pool: the same cases run as
tests/test_gpu_parity.py::test_queue_mode_small_tables, the kernel source under ASan through tools/emu_asan.sh): small tables forced onto nwb_fill_hx_kernel<.., QUEUE>
(nwb_tune pk_hx = 1), two plans taking different fills round robin, 1 / 2 / 3 strips per block, with and without the
count behind the fill; every fill against the oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nw_b200 as nwb  # noqa: E402
import oracle  # noqa: E402

nwb.tune("pk_hx", 1)
for spb in (0, 1, 2):
    nwb.tune("hx_spb", spb)
    for (a, b, flags) in ((1500, 300, 0), (700, 260, nwb.WANT_COUNT), (2100, 64, 0), (7, 7, 0)):
        pairs = [oracle.generate_pair(0x5EED0F80 + 2 * i + a, a, b) for i in range(2)]
        want = []
        for t, s in pairs:
            o = oracle.fill(t, s, 2, 1, 2)
            want.append((o.final_score, o.branch_count, o.arrow_digest, o.count if flags else 0))
        plans = [nwb.Plan(a, b, flags | nwb.QUEUE) for _ in range(2)]
        took = [None, None]
        for step, q in enumerate([0, 1, 1, 0, 0, None, None]):
            pl = plans[step % 2]
            if took[step % 2] is not None:
                sm = pl.summary()
                assert pl.kernel_name() == "nwb_fill_hx_kernel", pl.kernel_name()
                assert (sm.opt_score, sm.branch_count, pl.arrow_digest(), sm.count) == want[took[step % 2]], (spb, a, b, step)
            if q is not None:
                pl.upload(*pairs[q])
                pl.run(2, 1, 2)
            took[step % 2] = q
        for pl in plans:
            pl.close()
nwb.tune_reset()
print("sanitize_queue ok")
