"""CPU tests: the REAL general kernel source (csrc/nwb_fill_i32.cuh) executed
under the test-only SIMT emulator (tests/emu) against the oracle.  This checks
indexing, the strip hand-off protocol and the 2-GPU split logic without a GPU;
the GPU parity tests proper are in test_gpu_parity.py."""
import random

import numpy as np
import pytest

import emu

ALL = 1 | 2 | 8 | 0x20


def check(oracle, t, s, m, k, d, grid=2, split=0, flags=ALL):
    r = emu.fill_i32(t, s, m, k, d, flags=flags, grid=grid, split=split)
    o = oracle.fill(t, s, m, k, d, want_scores=True, want_codes=True, want_counts=True)
    a = len(t)
    assert np.array_equal(emu.unpack_arrows(r["arrows"], a) & 7, o.codes[1:, 1:] & 7)
    assert r["opt_score"] == o.final_score
    assert r["branch_count"] == o.branch_count
    if flags & 8:
        assert r["greatest_abs"] == o.greatest_abs
    if flags & 2:
        assert r["count"] == o.count
    if flags & 1:
        assert np.array_equal(r["scores"][:, :a], o.scores[1:, 1:])
    if flags & 0x20:
        assert np.array_equal(r["cntmat"][:, :a], o.counts[1:, 1:])


def test_readme(oracle):
    check(oracle, b"GCATGCU", b"GATTACA", 1, 1, 1)
    check(oracle, b"GCATGCU", b"GATTACA", 0, 0, 0)


@pytest.mark.parametrize("a,b", [(1, 1), (5, 40), (40, 5), (255, 33), (256, 64), (257, 65), (300, 100),
                                 (513, 70), (600, 130)])
def test_shapes(oracle, a, b):
    rng = random.Random(a * 1000 + b)
    t = bytes(rng.choice(b"ACGT") for _ in range(a))
    s = bytes(rng.choice(b"ACGT") for _ in range(b))
    for m, k, d in [(1, 1, 1), (2, 1, 2), (0, 0, 0), (-1, 3, -2)]:
        check(oracle, t, s, m, k, d, grid=rng.choice([1, 2, 3]))


def test_flag_subsets_and_split(oracle):
    t, s = oracle.generate_pair(0x5EED0900, 700, 150)
    for flags in (0, 1, 2, 8, 2 | 8, 1 | 2):
        check(oracle, t, s, 1, 1, 1, flags=flags)
    check(oracle, t, s, 1, 1, 1, grid=2, split=1)
    check(oracle, t, s, 2, 1, 2, grid=1, split=2, flags=2)


def test_split_with_several_strips_per_rank(oracle):
    """Both halves of an emulated 2-GPU split own two or more strips (each rank has its own progress words and
    boundary arrays, as every nwb_plan has): A = 1281 is 6 strips."""
    t, s = oracle.generate_pair(0x5EED0902, 1281, 40)
    for split in (2, 3, 4):
        check(oracle, t, s, 1, 1, 1, grid=2, split=split, flags=2 | 8)
    t, s = oracle.generate_pair(0x5EED0904, 1025, 70)
    check(oracle, t, s, 2, 1, 2, grid=3, split=3, flags=2)
    t, s = oracle.generate_pair(0x5EED0906, 1024, 33)
    check(oracle, t, s, 1, 1, 1, grid=1, split=2, flags=2)
