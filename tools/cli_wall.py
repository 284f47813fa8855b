#!/usr/bin/env python
"""Whole-process wall time of the drop-in CLI against the unmodified reference CLI (CPU baseline (ii)):
`needleman-wunsch -q -f <file> m k d` on BASELINE config 2 (10k x 10k DNA) for both, and the configurations the
reference cannot run (-s at this size is an exponential enumeration; 30k and 100k squares need 122 GB / 1.36 TB of
tables) for this repository's CLI only.  Process start, CUDA context creation, input parsing, H2D/D2H and output are
all inside the measured time.      python tools/cli_wall.py [--no-ref] [--reps 3]
Test/measurement infrastructure: the reference binary is oracle/_ref/needleman-wunsch (built by oracle/Makefile)."""
import argparse
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nw_b200 as oracle  # noqa: E402  (the package's own SURVEY 8d generator: generate_pair, DNA, PROTEIN)

OURS = os.path.join(ROOT, "needleman-wunsch_b200", "host", "needleman-wunsch")
REF = os.path.join(ROOT, "oracle", "_ref", "needleman-wunsch")

ap = argparse.ArgumentParser()
ap.add_argument("--no-ref", action="store_true")
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()


def wall(cmd, reps):
    best, out = 1e9, None
    for _ in range(reps):
        t0 = time.perf_counter()
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        best = min(best, time.perf_counter() - t0)
        out = (r.returncode, r.stdout, r.stderr)
    return best, out


cases = [("config2 dna 10k", 0x5EED0002, 10000, oracle.DNA, ("1", "1", "1")),
         ("config5 protein 30k", 0x5EED0005, 30000, oracle.PROTEIN, ("2", "1", "2")),
         ("config3 dna 100k", 0x5EED0030, 100000, oracle.DNA, ("1", "1", "1"))]
print(f"host cores: {os.cpu_count()}")
with tempfile.TemporaryDirectory() as tmp:
    for name, seed, n, alpha, mkd in cases:
        t, s = oracle.generate_pair(seed, n, n, alpha)
        path = os.path.join(tmp, f"pair_{n}.txt")
        with open(path, "wb") as f:
            f.write(t + b"\n" + s + b"\n")
        cells = n * n
        for flags in (["-q"], ["-q", "-s"]):
            w, out = wall([OURS] + flags + ["-f", path, *mkd], args.reps)
            print(f"{name:20s} ours {' '.join(flags):6s}: {w:7.3f} s  {cells / w / 1e9:8.2f} GCUPS whole process  rc={out[0]} "
                  f"stderr={out[2].decode().strip()!r}", flush=True)
        if n == 10000 and not args.no_ref and os.path.exists(REF):
            for extra in ([], ["-p", "2"], ["-p", str(os.cpu_count())]):
                w, out = wall([REF, "-q"] + extra + ["-f", path, *mkd], 1)
                print(f"{name:20s} reference -q {' '.join(extra):6s}: {w:7.3f} s  {cells / w / 1e9:8.4f} GCUPS whole process  rc={out[0]}",
                      flush=True)
