#!/usr/bin/env python
"""A small set of fills through the C ABI for `compute-sanitizer --tool memcheck|racecheck`:
every fill kernel, with and without the count (sweep and fused), several strips, the batch kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nw_b200 as nwb  # noqa: E402
import oracle  # noqa: E402

for a, b in ((7, 7), (300, 100), (700, 260), (1500, 64)):
    t, s = oracle.generate_pair(0x5EED0F00 + a, a, b)
    o = oracle.fill(t, s, 1, 1, 1)
    for flags in (0, nwb.WANT_COUNT, nwb.WANT_COUNT | nwb.FORCE_GENERAL | nwb.TRACK_ABS | nwb.WANT_SCORES):
        tab = nwb.fill(t, s, 1, 1, 1, flags | nwb.WANT_ARROWS_HOST)
        assert (tab.opt_score, tab.branch_count) == (o.final_score, o.branch_count), (a, b, flags)
        if flags & nwb.WANT_COUNT:
            assert tab.count == o.count
        tab.close()
# the sweeping + flush warp kernel (nwb_tune pk_hx = 1 forces it at small sizes) and the count sweep after it;
# 1500 x 300: 6 strips on one block's three sweeping warps
nwb.tune("pk_hx", 1)
for mode in (0, 2, 3):   # count: sparse backward sweep (default) / dense sweep after / trailing the fill
    nwb.tune("count_mode", mode)
    for a, b in ((7, 7), (300, 101), (700, 260), (1500, 300)):
        t, s = oracle.generate_pair(0x5EED0F40 + a, a, b)
        o = oracle.fill(t, s, 2, 1, 2)
        for flags in (0, nwb.WANT_COUNT):
            tab = nwb.fill(t, s, 2, 1, 2, flags | nwb.WANT_ARROWS_HOST)
            assert (tab.opt_score, tab.branch_count) == (o.final_score, o.branch_count), (a, b, flags, mode)
            if flags & nwb.WANT_COUNT:
                assert tab.count == o.count
            tab.close()
nwb.tune_reset()
tops, sides = zip(*(oracle.generate_pair(0x5EED4000 + 2 * p, 256 if p % 3 else 300, 256 if p % 2 else 100) for p in range(40)))
bt = nwb.Batch(list(tops), list(sides), 1, 1, 1, nwb.WANT_ARROWS_HOST)
bt.run()
bt.fetch()
for p in range(40):
    o = oracle.fill(tops[p], sides[p], 1, 1, 1)
    assert (bt.opt_score(p), bt.branch_count(p)) == (o.final_score, o.branch_count), p
bt.close()
# the bit-parallel batch kernel (one thread per pair; its left-over list through nwb_batch_pk_kernel) and the per-lane
# sparse count (its left-over list through nwb_batch_count_kernel): uniform and ragged groups, unaligned offsets,
# top strings with five and six letters, empty strings
import random  # noqa: E402
rng = random.Random(3)
shapes = [(256, 256)] * 40 + [(200, 90), (256, 31), (17, 130), (1, 1), (64, 64), (0, 3), (5, 0), (255, 77), (129, 300)] * 4
tops = [bytes(rng.choice(b"ACGTNR" if i % 11 == 0 else (b"ACGTN" if i % 11 == 1 else b"ACGT")) for _ in range(a)) for i, (a, b) in enumerate(shapes)]
sides = [bytes(rng.choice(b"ACGTX") for _ in range(b)) for a, b in shapes]
for flags in (nwb.WANT_ARROWS_HOST, nwb.WANT_COUNT):
    with nwb.tuned(batch_bp=1, batch_lcount=1):
        bt = nwb.Batch(tops, sides, 1, 1, 1, flags)
    assert bt.kernel_name() == "nwb_batch_bp_kernel"
    bt.run()
    bt.fetch()
    for p in range(0, len(shapes), 3):
        o = oracle.fill(tops[p], sides[p], 1, 1, 1)
        assert (bt.opt_score(p), bt.branch_count(p)) == (o.final_score, o.branch_count), p
        if flags & nwb.WANT_COUNT:
            assert bt.count(p) == o.count, p
    bt.close()
print("sanitize_run ok")
