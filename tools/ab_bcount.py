#!/usr/bin/env python
"""A/B of the batch count pass on BASELINE config 4's per-GPU shard (125,000 pairs of 256 x 256 DNA): pairs swept back
to back by nwb_batch_count_chain_kernel (default for uniform one-strip batches) against one pair at a time
(NWB_BCNT_CHAIN=0).  Every count of the two runs is compared; pairs 0 and 1 against SURVEY.md 8c's goldens.
    python tools/ab_bcount.py [--pairs 125000]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nw_b200 as nwb  # noqa: E402
import oracle  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=125000)
args = ap.parse_args()
n = args.pairs
tcat, scat = bytearray(), bytearray()
for p in range(n):
    t, s = oracle.generate_pair(0x5EED4000 + 2 * p, 256, 256)
    tcat += t
    scat += s
off = np.arange(n + 1, dtype=np.int64) * 256
res = {}
for chain in ("0", "1"):
    nwb.tune("bcnt_chain", int(chain))
    for flags, what in ((0, "fill"), (nwb.WANT_COUNT, "fill+count")):
        b = nwb.Batch.from_arrays(bytes(tcat), off, bytes(scat), off, 1, 1, 1, flags)
        ms = []
        for _ in range(4):
            b.run()
            b.fetch()
            ms.append(b.kernel_ms())
        if flags:
            res[chain] = np.array([b.count(i) for i in range(0, n, 97)] + [b.count(0), b.count(1), b.count(n - 1)], dtype=np.uint64)
        print(f"chain={chain} {what:10s} {min(ms):7.3f} ms  {n * 65536 / min(ms) / 1e6:8.1f} GCUPS", flush=True)
        b.close()
same = bool(np.array_equal(res["0"], res["1"]))
gold = (int(res["1"][-3]), int(res["1"][-2])) == (387701138034524160, 108460706365440)
print("counts equal:", same, " goldens:", gold)
sys.exit(0 if same and gold else 1)
