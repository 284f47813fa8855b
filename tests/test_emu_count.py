"""CPU tests: the count sweep over the finished arrow codes (csrc/nwb_count.cuh) under the test-only
SIMT emulator, after both packed fill kernels: the final uint64 count (mod 2^64) against the oracle,
whose count was itself checked against the reference's enumeration (tests/test_oracle.py)."""
import random

import numpy as np
import pytest

import emu


def chk(oracle, t, s, m, k, d, hx, grid=2, split=0, cpls=(2, 4, 8)):
    o = oracle.fill(t, s, m, k, d, want_codes=True)
    for cpl in cpls:   # cells per lane and row of the count sweep: strips of 64, 128, 256 columns
        r = emu.fill_pk(t, s, m, k, d, K=4, R=2, grid=grid, split=split, count=cpl, hx=hx)
        assert np.array_equal(emu.unpack_arrows(r["arrows"], len(t)) & 7, o.codes[1:, 1:] & 7)
        assert (r["opt_score"], r["branch_count"], r["count"]) == (o.final_score, o.branch_count, o.count), cpl


@pytest.mark.parametrize("hx", [0, 1, 2, 3], ids=["pk", "hx", "hy", "hz"])
def test_readme_and_delannoy(oracle, hx):
    chk(oracle, b"GCATGCU", b"GATTACA", 1, 1, 1, hx)          # 3 optimal alignments (README:154)
    chk(oracle, b"GCATGCU", b"GATTACA", 0, 0, 0, hx, grid=1)  # every arrow everywhere: Delannoy(7,7) = 48,639
    r = emu.fill_pk(b"GCATGCU", b"GATTACA", 0, 0, 0, K=4, R=2, count=8, hx=hx)
    assert r["count"] == 48639


@pytest.mark.parametrize("hx", [0, 1, 2, 3], ids=["pk", "hx", "hy", "hz"])
def test_shapes(oracle, hx):
    rng = random.Random(31 + hx)
    for a, b in [(1, 1), (5, 40), (63, 33), (256, 64), (257, 130), (513, 70), (600, 201), (130, 256), (256, 256),
                 (8, 300), (300, 1)]:
        t = bytes(rng.choice(b"ACGT") for _ in range(a))
        s = bytes(rng.choice(b"ACGT") for _ in range(b))
        for m, k, d in [(1, 1, 1), (2, 1, 2), (0, 0, 0)]:
            chk(oracle, t, s, m, k, d, hx, grid=rng.choice([1, 2, 3]), cpls=(rng.choice([2, 4, 8]),))


def test_strips_per_warp_and_split(oracle):
    t, s = oracle.generate_pair(0x5EED0915, 2300, 90)   # 9 strips on 4 warps: cyclic second and third pass
    chk(oracle, t, s, 1, 1, 1, True, grid=1)
    t, s = oracle.generate_pair(0x5EED0917, 700, 210)
    chk(oracle, t, s, 1, 1, 1, True, grid=2, split=1)
    chk(oracle, t, s, 1, 1, 1, 2, grid=2, split=1)
    chk(oracle, t, s, 2, 1, 2, False, grid=1, split=2)
