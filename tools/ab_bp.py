#!/usr/bin/env python
"""A/B of the batch fill on BASELINE config 4's per-GPU shard (125,000 pairs of 256 x 256 DNA, 1/1/1): the
bit-parallel kernel (nwb_batch_bp_kernel: one thread per pair, a row per addition) against the packed-difference
kernel (nwb_batch_cx_kernel: two pairs per warp, swept back to back).  Every pair of both runs is checked through
the batch digests (arrow tables, scores, branch counters) against tests/golden/golden_big.json.
    python tools/ab_bp.py [--pairs 125000] [--warps 0,8,12,14,16]"""
import argparse
import json
import time
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nw_b200 as nwb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=125000)
ap.add_argument("--warps", default="0")
ap.add_argument("--a", type=int, default=256, help="top string length")
ap.add_argument("--b", type=int, default=256, help="side string length")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--lib", default=None, help="another build of libnwb.so (variant A/B on the same box)")
args = ap.parse_args()
if args.lib:
    nwb.LIB_PATH = os.path.abspath(args.lib)
n = args.pairs
A, B = args.a, args.b
tcat = nwb.generate(0x5EED4000, A, nwb.DNA, count=n, seed_stride=2)
scat = nwb.generate(0x5EED4001, B, nwb.DNA, count=n, seed_stride=2)
off = np.arange(n + 1, dtype=np.int64) * A
soff = np.arange(n + 1, dtype=np.int64) * B
want = None
if n == 125000 and (A, B) == (256, 256):
    with open(os.path.join(ROOT, "tests", "golden", "golden_big.json")) as f:
        g4 = {c["name"]: c for c in json.load(f)}["config4_batch_1M"]["shard_digests"][0]
    want = tuple(int(g4[k], 16) for k in ("arrow", "score", "branch"))
tpin, spin = nwb.PinnedBuffer(tcat), nwb.PinnedBuffer(scat)
ok = True
ref = None
runs = [("cx", dict(batch_bp=0))] + [(f"bp warps={w}", dict(batch_bp=1, bp_warps=int(w))) for w in args.warps.split(",")] + \
       [("bp table anywhere", dict(batch_bp=1, bp_aligned=0))]
for name, knobs in runs:
    with nwb.tuned(**knobs):
        b = nwb.Batch.from_arrays(tcat, off, scat, soff, 1, 1, 1, 0)
        ms = []
        for _ in range(args.reps):
            b.run()
            b.fetch()
            ms.append(b.kernel_ms())
        dg = b.digest(0)[:3]
        kname = b.kernel_name()
        # end to end from page-locked host buffers: chunked H2D overlapped with the kernels + D2H of the results
        e2e = []
        for _ in range(args.reps):
            t0 = time.perf_counter()
            b.refill(tpin, spin)
            b.fetch()
            e2e.append((time.perf_counter() - t0) * 1e3)
        dg2 = b.digest(0)[:3]
        b.close()
        if dg2 != dg:
            dg = None
    if ref is None:
        ref = dg
    good = dg == ref and (want is None or dg == want)
    ok = ok and good
    print(f"{name:14s} {kname:22s} {min(ms):7.3f} ms  {n * A * B / min(ms) / 1e6:8.1f} GCUPS  "
          f"{n * (128 if kname != 'nwb_batch_bp_kernel' else (32 if A <= 64 else (64 if A <= 128 else 128))) * B / min(ms) / 1e6:7.1f} GB/s written  e2e {min(e2e):6.3f} ms  digests {'ok' if good else 'MISMATCH'}", flush=True)
sys.exit(0 if ok else 1)
