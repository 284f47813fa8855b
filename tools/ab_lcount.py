#!/usr/bin/env python
"""A/B of the batch count pass on BASELINE config 4's per-GPU shard (125,000 pairs of 256 x 256 DNA, 1/1/1): the
per-lane sparse backward sweep (nwb_batch_lcount_kernel, the dense kernel behind it for the pairs it gives up on)
against the dense pass alone (nwb_batch_count_chain_kernel).  Every pair's count is checked through the batch
digest against tests/golden/golden_big.json.
    python tools/ab_lcount.py [--pairs 125000] [--warps 0,8,12]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nw_b200 as nwb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=125000)
ap.add_argument("--warps", default="0")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--lib", default=None, help="another build of libnwb.so (variant A/B on the same box)")
args = ap.parse_args()
if args.lib:
    nwb.LIB_PATH = os.path.abspath(args.lib)
n = args.pairs
tcat = nwb.generate(0x5EED4000, 256, nwb.DNA, count=n, seed_stride=2)
scat = nwb.generate(0x5EED4001, 256, nwb.DNA, count=n, seed_stride=2)
off = np.arange(n + 1, dtype=np.int64) * 256
want = None
if n == 125000:
    with open(os.path.join(ROOT, "tests", "golden", "golden_big.json")) as f:
        g4 = {c["name"]: c for c in json.load(f)}["config4_batch_1M"]["shard_digests"][0]
    want = tuple(int(g4[k], 16) for k in ("arrow", "score", "branch", "count"))
ok = True
ref = None
fill_ms = None
runs = [("fill only", dict(), 0), ("dense count", dict(batch_lcount=0), nwb.WANT_COUNT)] + \
       [(f"lane count w={w}", dict(batch_lcount=1, lc_warps=int(w)), nwb.WANT_COUNT) for w in args.warps.split(",")]
for name, knobs, flags in runs:
    with nwb.tuned(**knobs):
        b = nwb.Batch.from_arrays(tcat, off, scat, off, 1, 1, 1, flags)
        ms = []
        for _ in range(args.reps):
            b.run()
            b.fetch()
            ms.append(b.kernel_ms())
        dg = b.digest(0)
        b.close()
    if not flags:
        fill_ms = min(ms)
        print(f"{name:18s} {min(ms):7.3f} ms", flush=True)
        continue
    if ref is None:
        ref = dg
    good = dg == ref and (want is None or dg == want)
    ok = ok and good
    print(f"{name:18s} {min(ms):7.3f} ms fill + count  (count pass {min(ms) - fill_ms:6.3f} ms)  "
          f"{n * 65536 / min(ms) / 1e6:8.1f} GCUPS  digests {'ok' if good else 'MISMATCH'}", flush=True)
sys.exit(0 if ok else 1)
