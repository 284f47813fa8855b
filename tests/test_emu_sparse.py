"""The sparse backward count (csrc/nwb_count_sparse.cuh) and the digest kernels (csrc/nwb_digest.cuh), run
under the SIMT emulator on the CPU and compared with the oracle's dense forward DP / digests."""
import random

import numpy as np
import pytest

import oracle
import emu


def _packed(top, side, m, k, d, pitch=None, garbage=False):
    a = len(top)
    pitch = pitch or max(128, ((a + 255) // 256) * 128)
    o = oracle.fill(top, side, m, k, d, want_packed=True, pitch=pitch)
    pk = o.packed.copy()
    if garbage:
        # bit 3 of every nibble and everything beyond column A is unspecified in the ABI: fill it with junk
        pk |= 0x88
        nb = (a + 1) // 2
        pk[:, nb:] = 0xFF
        if a & 1:
            pk[:, nb - 1] |= 0xF0
    return o, pk


CASES = [
    # (seed, A, B, alphabet, (m,k,d))
    (0x5EED4000, 256, 256, oracle.DNA, (1, 1, 1)),      # config 4 pair 0: 387,701,138,034,524,160
    (0x5EED4002, 256, 256, oracle.DNA, (1, 1, 1)),
    (0x5EED0C00, 1000, 700, oracle.DNA, (1, 1, 1)),
    (0x5EED0C02, 700, 1300, oracle.DNA, (1, 1, 1)),
    (0x5EED0005, 900, 800, oracle.PROTEIN, (2, 1, 2)),
    (0x5EED0300, 600, 400, oracle.PROTEIN, (5, 4, 3)),
    (0x5EED0200, 300, 500, oracle.DNA, (-1, 3, -2)),
    (0x5EED0111, 1, 1, oracle.DNA, (1, 1, 1)),
    (0x5EED0112, 9, 1, oracle.DNA, (1, 1, 1)),
    (0x5EED0113, 1, 40, oracle.DNA, (1, 1, 1)),
    (0x5EED0114, 257, 33, oracle.DNA, (1, 1, 1)),
    (0x5EED0115, 8, 8, oracle.DNA, (1, 1, 1)),
]


@pytest.mark.parametrize("seed,a,b,alpha,mkd", CASES)
def test_sparse_count_matches_dense_dp(seed, a, b, alpha, mkd):
    top, side = oracle.generate_pair(seed, a, b, alpha)
    o, pk = _packed(top, side, *mkd, garbage=True)
    for mode in (1, 2):
        rm = emu.sparse_count(pk, a, mode)
        assert rm["state"] == 2 or rm["count"] == o.count, mode
    r = emu.sparse_count(pk, a)
    if mkd == (-1, 3, -2):
        # gaps are rewarded: the optimal alignments spread over the whole 300-column table, wider than the
        # window -- the sweep must say so (the dense sweep then runs), never return a wrong count
        assert r["state"] == 2
        return
    assert r["state"] == 1
    assert r["count"] == o.count


def test_sparse_count_unique_alignment_walks_every_row():
    top, _ = oracle.generate_pair(0x5EED0777, 1500, 1)
    o, pk = _packed(top, top, 1, 1, 1)
    assert o.count == 1
    r = emu.sparse_count(pk, len(top))
    assert (r["state"], r["count"], r["rows"]) == (1, 1, 1500)


def test_sparse_count_near_identical_strings_nonzero():
    # a mutated copy: few branch points, so the count stays non-zero mod 2^64 over thousands of rows
    top, _ = oracle.generate_pair(0x5EED0778, 3000, 1)
    side = bytearray(top)
    rng = np.random.default_rng(7)
    for pos in rng.choice(len(side), 40, replace=False):
        side[pos] = ord("ACGT"[(("ACGT".index(chr(side[pos]))) + 1) % 4])
    del side[1000:1003]
    side[2000:2000] = b"GATTACA"
    o, pk = _packed(top, bytes(side), 1, 1, 1)
    assert o.count != 0
    r = emu.sparse_count(pk, len(top))
    assert (r["state"], r["count"]) == (1, o.count)
    assert r["rows"] == len(side)


def test_sparse_count_long_left_run_crosses_window():
    # top = side + 400 extra characters: the optimal paths end with a run of 400 LEFT arrows in row B,
    # wider than the 256-column window -> the sweep must give up cleanly (the dense sweep takes over)
    side, _ = oracle.generate_pair(0x5EED0779, 300, 1)
    top = side + b"T" * 400
    o, pk = _packed(top, side, 1, 1, 1)
    r = emu.sparse_count(pk, len(top))
    assert r["state"] == 2


def test_sparse_count_gap_inside_window():
    side, _ = oracle.generate_pair(0x5EED077A, 300, 1)
    top = side[:150] + b"T" * 100 + side[150:]
    o, pk = _packed(top, side, 1, 1, 1)
    r = emu.sparse_count(pk, len(top))
    assert (r["state"], r["count"]) == (1, o.count)


def test_sparse_count_all_ties_bails_or_matches():
    # m = k = d = 0: every cell has all three arrows, the live band is the whole table (Delannoy numbers)
    top, side = oracle.generate_pair(0x5EED077B, 200, 180)
    o, pk = _packed(top, side, 0, 0, 0)
    r = emu.sparse_count(pk, len(top))
    assert r["state"] == 1 and r["count"] == o.count   # 200 columns fit the 256-column window
    top, side = oracle.generate_pair(0x5EED077C, 400, 180)
    o, pk = _packed(top, side, 0, 0, 0)
    r = emu.sparse_count(pk, len(top))
    assert r["state"] == 2 or r["count"] == o.count


def test_sparse_count_tall_up_run():
    top, _ = oracle.generate_pair(0x5EED077D, 200, 1)
    side = top[:100] + b"A" * 700 + top[100:]
    o, pk = _packed(top, side, 1, 1, 1)
    r = emu.sparse_count(pk, len(top))
    assert (r["state"], r["count"]) == (1, o.count)


@pytest.mark.parametrize("a,b", [(777, 301), (256, 256), (1, 5), (1025, 70), (8, 3)])
def test_arrow_digest_kernel(a, b):
    top, side = oracle.generate_pair(0x5EED0C10 + a, a, b)
    o, pk = _packed(top, side, 1, 1, 1, garbage=True)
    assert emu.arrow_digest(pk, a) == o.arrow_digest
    # a strip group's shares add up
    nw = (a + 7) // 8
    cut = nw // 3
    total = (emu.arrow_digest(pk, a, 0, cut) + emu.arrow_digest(pk, a, cut, nw)) % 2**64
    assert total == o.arrow_digest


def test_dense_count_digests():
    top, side = oracle.generate_pair(0x5EED0C20, 700, 300)
    o = oracle.fill(top, side, 1, 1, 1)
    r = emu.fill_pk(top, side, 1, 1, 1, K=4, R=2, count=8, hx=True, grid=2)
    assert r["count"] == o.count
    assert (r["dig_row"], r["dig_col"]) == (o.lastrow_count_digest, o.lastcol_count_digest)
    r = emu.fill_pk(top, side, 1, 1, 1, K=4, R=2, count=8, hx=True, grid=2, split=2)
    assert (r["count"], r["dig_row"], r["dig_col"]) == (o.count, o.lastrow_count_digest, o.lastcol_count_digest)


def test_sparse_count_random_sweep():
    rng = np.random.default_rng(20261019)
    done = 0
    for it in range(60):
        a, b = int(rng.integers(1, 420)), int(rng.integers(1, 420))
        alpha = oracle.DNA if it % 3 else "AC"
        mkd = [(1, 1, 1), (2, 1, 2), (1, 2, 1), (3, 1, 2), (0, 1, 1), (1, 0, 1)][it % 6]
        top, side = oracle.generate_pair(0x5EED9000 + 2 * it, a, b, alpha)
        o, pk = _packed(top, side, *mkd, garbage=bool(it & 1))
        # as the product (64-column window, then 256), 256-column window only, 64-column window only
        for mode in (0, 1, 2):
            r = emu.sparse_count(pk, a, mode)
            assert r["state"] in (1, 2)
            if r["state"] == 1:
                assert r["count"] == o.count, (it, a, b, mkd, mode)
                done += mode == 0
    assert done >= 40


def test_sparse_count_on_the_last_rank_of_a_strip_group(oracle):
    """min_col > 1: the arrow columns left of min_col live on another GPU (here: overwritten with garbage).  The sweep
    either finishes inside its own columns with the right count or gives up -- it never reads the foreign columns into
    a result."""
    rng = random.Random(77)
    done = gave_up = 0
    for (a, b, seed) in ((1500, 1500, 1), (1500, 1500, 2), (1300, 1400, 3), (900, 2000, 4), (600, 600, 5)):
        t, s = oracle.generate_pair(0x5EED7000 + seed, a, b)
        o = oracle.fill(t, s, 1, 1, 1, want_packed=True, pitch=oracle.packed_pitch(a))
        full = emu.sparse_count(o.packed, a)
        assert full["state"] == 1 and full["count"] == o.count
        for min_col in (257, 513, 1025):
            if min_col > a:
                continue
            tab = o.packed.copy()
            tab[:, :(min_col - 1) // 2] = rng.choice([0x77, 0x22, 0x11, 0x55])     # not this rank's columns
            r = emu.sparse_count(tab, a, min_col=min_col)
            if r["state"] == 1:
                assert r["count"] == o.count, (a, b, min_col)
                done += 1
            else:
                assert r["state"] == 2
                gave_up += 1
    assert done >= 3 and gave_up >= 3
