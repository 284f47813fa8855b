"""CPU tests: the sweeping + flush warp variant of the packed kernel
(csrc/nwb_fill_hx.cuh) executed under the test-only SIMT emulator against the
oracle: arrows, optimal score (bottom-row difference sum), branch counter; single
and multiple strips, more strips than sweeping warps (several strips per warp, so
the ring and its flags run across strip changes), tall tables (ring wrap-around)
and the 2-GPU split."""
import random

import numpy as np
import pytest

import emu

SCHEMES = [(1, 1, 1), (2, 1, 2), (0, 0, 0), (1, 2, 3), (1, 0, 0), (3, -1, 0), (3, 1, 2), (1, 1, 3)]


VARIANTS = (1, 2, 3, 4, 5, 6)   # 1 = nwb_fill_hx.cuh, 2 = nwb_fill_hy.cuh (experiment), 3 = nwb_fill_hz.cuh (packing warps, two strips per block),
                                # 4 / 5 / 6 = nwb_fill_hx.cuh in queue mode (ticketed blocks, 1 / 2 / 3 adjacent strips each)


def check(oracle, t, s, m, k, d, grid=2, split=0):
    assert emu.hx_supported(m, k, d)
    o = oracle.fill(t, s, m, k, d, want_codes=True)
    for hx in VARIANTS:
        r = emu.fill_pk(t, s, m, k, d, K=4, R=2, grid=grid, split=split, hx=hx)
        assert np.array_equal(emu.unpack_arrows(r["arrows"], len(t)) & 7, o.codes[1:, 1:] & 7), hx
        assert r["opt_score"] == o.final_score, hx
        assert r["branch_count"] == o.branch_count, hx


def test_supported():
    assert emu.hx_supported(1, 1, 1) and emu.hx_supported(2, 1, 2) and emu.hx_supported(0, 0, 0)
    assert emu.hx_supported(1, 1, 3)        # 2d + m = 7
    assert not emu.hx_supported(2, 1, 3)    # 2d + m = 8: nibbles overflow, plain packed kernel
    assert not emu.hx_supported(5, 4, 3)


def test_readme(oracle):
    check(oracle, b"GCATGCU", b"GATTACA", 1, 1, 1)
    check(oracle, b"GCATGCU", b"GATTACA", 0, 0, 0, grid=1)


@pytest.mark.parametrize("a,b", [(1, 1), (5, 40), (40, 5), (63, 33), (64, 64), (255, 33), (256, 64), (257, 130),
                                 (513, 70), (600, 201), (130, 256), (70, 257), (90, 700), (300, 1)])
def test_shapes(oracle, a, b):
    rng = random.Random(a * 7919 + b)
    for alpha in (b"ACGT", bytes(range(1, 256))):
        t = bytes(rng.choice(alpha) for _ in range(a))
        s = bytes(rng.choice(alpha) for _ in range(b))
        for m, k, d in rng.sample(SCHEMES, 2):
            check(oracle, t, s, m, k, d, grid=rng.choice([1, 2, 3]))


def test_strips_per_warp_and_split(oracle):
    # 8 strips on one block of 3 sweeping warps: warps take 3, 3 and 2 strips one after the other
    t, s = oracle.generate_pair(0x5EED0911, 1900, 150)
    check(oracle, t, s, 1, 1, 1, grid=1)
    check(oracle, t, s, 2, 1, 2, grid=2, split=3)
    t, s = oracle.generate_pair(0x5EED0913, 700, 210)
    check(oracle, t, s, 1, 1, 1, grid=2, split=1)
    check(oracle, t, s, 1, 1, 1, grid=1, split=2)
