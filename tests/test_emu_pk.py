"""CPU tests: the REAL packed 16x2 kernel source (csrc/nwb_fill_pk.cuh) and the
branch-count pass executed under the test-only SIMT emulator against the oracle
(arrows, optimal score from the bottom-row difference sum, branch counter),
for every strip width K, several grids and the 2-GPU split."""
import random

import numpy as np
import pytest

import emu

SCHEMES = [(1, 1, 1), (2, 1, 2), (0, 0, 0), (1, 2, 3), (5, 4, 3), (1, 0, 0), (3, -1, 0)]


def check(oracle, t, s, m, k, d, K, grid=2, split=0, R=None):
    if R is None:
        for rr in (1, 2):
            check(oracle, t, s, m, k, d, K, grid, split, rr)
        return
    r = emu.fill_pk(t, s, m, k, d, K=K, R=R, grid=grid, split=split)
    o = oracle.fill(t, s, m, k, d, want_codes=True)
    assert np.array_equal(emu.unpack_arrows(r["arrows"], len(t)) & 7, o.codes[1:, 1:] & 7)
    assert r["opt_score"] == o.final_score
    assert r["branch_count"] == o.branch_count


def test_supported_schemes():
    for m, k, d in SCHEMES:
        assert emu.pk_supported(m, k, d)
    # mismatch diagonal negative / mismatch better than match / too large: general kernel
    assert not emu.pk_supported(1, 3, 1)
    assert not emu.pk_supported(-1, 3, -2)
    assert not emu.pk_supported(-3, 1, 2)
    assert not emu.pk_supported(100, 100, 1)


@pytest.mark.parametrize("K", [4, 2, 1])
def test_readme(oracle, K):
    check(oracle, b"GCATGCU", b"GATTACA", 1, 1, 1, K)
    check(oracle, b"GCATGCU", b"GATTACA", 0, 0, 0, K)


@pytest.mark.parametrize("a,b", [(1, 1), (5, 40), (40, 5), (63, 33), (64, 64), (65, 65), (255, 33), (256, 64),
                                 (257, 130), (513, 70), (600, 200), (200, 128), (130, 256), (70, 257), (90, 700)])
def test_shapes(oracle, a, b):
    rng = random.Random(a * 7919 + b)
    for alpha in (b"ACGT", bytes(range(1, 256))):
        t = bytes(rng.choice(alpha) for _ in range(a))
        s = bytes(rng.choice(alpha) for _ in range(b))
        for m, k, d in rng.sample(SCHEMES, 3):
            for K in (4, 2, 1):
                check(oracle, t, s, m, k, d, K, grid=rng.choice([1, 2, 3]))


def test_two_gpu_split(oracle):
    t, s = oracle.generate_pair(0x5EED0910, 700, 210)
    check(oracle, t, s, 1, 1, 1, 4, grid=2, split=1)
    check(oracle, t, s, 1, 1, 1, 2, grid=1, split=2)
    check(oracle, t, s, 2, 1, 2, 1, grid=2, split=5)


def test_fused_count(oracle):
    """NWB_WANT_COUNT in the packed kernel: the final uint64 count (mod 2^64) against the oracle on
    shapes with one and several strips, both row-group sizes, and across the 2-GPU split."""
    rng = random.Random(9)

    def chk(t, s, m, k, d, R, grid=2, split=0):
        r = emu.fill_pk(t, s, m, k, d, K=4, R=R, grid=grid, split=split, count=True)
        o = oracle.fill(t, s, m, k, d, want_codes=True)
        assert np.array_equal(emu.unpack_arrows(r["arrows"], len(t)) & 7, o.codes[1:, 1:] & 7)
        assert (r["opt_score"], r["branch_count"], r["count"]) == (o.final_score, o.branch_count, o.count)

    chk(b"GCATGCU", b"GATTACA", 1, 1, 1, 1)
    chk(b"GCATGCU", b"GATTACA", 0, 0, 0, 2)
    for a, b in [(1, 1), (5, 40), (63, 33), (256, 64), (257, 130), (513, 70), (600, 200), (130, 256), (256, 256)]:
        t = bytes(rng.choice(b"ACGT") for _ in range(a))
        s = bytes(rng.choice(b"ACGT") for _ in range(b))
        for m, k, d in [(1, 1, 1), (2, 1, 2)]:
            for R in (1, 2):
                chk(t, s, m, k, d, R, grid=rng.choice([1, 2, 3]))
    t, s = oracle.generate_pair(0x5EED0910, 700, 210)
    chk(t, s, 1, 1, 1, 2, grid=2, split=1)
    chk(t, s, 1, 1, 1, 1, grid=1, split=2)
