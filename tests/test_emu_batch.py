"""CPU test: the REAL batch kernel source (csrc/nwb_batch.cuh, one warp per pair)
under the SIMT emulator against the oracle -- ragged lengths, pairs wider than
one 256-column strip, empty strings, more pairs than warps."""
import random

import numpy as np

import emu


def test_batch_ragged(oracle):
    rng = random.Random(11)
    lens = [(256, 256), (1, 1), (255, 257), (300, 40), (17, 130), (0, 5), (700, 90), (64, 64), (256, 1),
            (33, 33), (257, 31), (100, 300), (5, 0)] + [(rng.randint(1, 280), rng.randint(1, 200)) for _ in range(12)]
    tops = [bytes(rng.choice(b"ACGT") for _ in range(a)) for a, _ in lens]
    sides = [bytes(rng.choice(b"ACGT") for _ in range(b)) for _, b in lens]
    for (m, k, d), grid in (((1, 1, 1), 1), ((2, 1, 2), 2)):
        r = emu.fill_batch(tops, sides, m, k, d, grid=grid, bx=0)
        for i, (t, s) in enumerate(zip(tops, sides)):
            o = oracle.fill(t, s, m, k, d, want_codes=True)
            assert r["scores"][i] == o.final_score, (i, len(t), len(s))
            assert r["branches"][i] == o.branch_count, (i, len(t), len(s))
            if len(t) and len(s):
                assert np.array_equal(emu.unpack_arrows(r["tables"][i], len(t)) & 7, o.codes[1:, 1:] & 7), i


def test_batch_config4_goldens(oracle):
    # config 4 sample pairs (seeds 0x5EED4000 + 2p)
    tops, sides = [], []
    for p in (0, 1, 999999):
        t, s = oracle.generate_pair(0x5EED4000 + 2 * p, 256, 256)
        tops.append(t)
        sides.append(s)
    r = emu.fill_batch(tops, sides, 1, 1, 1, grid=1, bx=0)
    assert list(r["scores"]) == [19, 29, 19]
    assert list(r["branches"]) == [23713, 22912, 22090]


# ---- two pairs per warp (csrc/nwb_batch_bx.cuh) ---------------------------------------------------------

def _check_batch(oracle, tops, sides, m, k, d, *, grid, bx):
    r = emu.fill_batch(tops, sides, m, k, d, grid=grid, bx=bx)
    assert r["bx"] == bool(bx)
    for i, (t, s) in enumerate(zip(tops, sides)):
        o = oracle.fill(t, s, m, k, d, want_codes=True)
        assert r["scores"][i] == o.final_score, (i, len(t), len(s))
        assert r["branches"][i] == o.branch_count, (i, len(t), len(s))
        if len(t) and len(s):
            assert np.array_equal(emu.unpack_arrows(r["tables"][i], len(t)) & 7, o.codes[1:, 1:] & 7), (i, len(t), len(s))
    return r


def test_bx_config4_goldens(oracle):
    tops, sides = [], []
    for p in (0, 1, 999999):   # an odd number of pairs: the last warp sweeps one pair alone
        t, s = oracle.generate_pair(0x5EED4000 + 2 * p, 256, 256)
        tops.append(t)
        sides.append(s)
    r = _check_batch(oracle, tops, sides, 1, 1, 1, grid=1, bx=1)
    assert list(r["scores"]) == [19, 29, 19]
    assert list(r["branches"]) == [23713, 22912, 22090]


def test_bx_ragged_pairs_share_a_warp(oracle):
    # partners of different shapes (the shorter one finishes early), empty strings, single cells, tall and
    # flat tables, 31/32/33 and 63/64/65 rows (block and ring-slot boundaries), more pair-pairs than warps
    rng = random.Random(23)
    lens = [(256, 256), (1, 1), (255, 257), (3, 40), (17, 130), (0, 5), (200, 90), (64, 64), (256, 1), (33, 33),
            (256, 31), (100, 300), (5, 0), (8, 32), (9, 33), (249, 63), (250, 64), (7, 65), (1, 200), (256, 2),
            (0, 0), (31, 31), (130, 95), (96, 128)] + [(rng.randint(1, 256), rng.randint(1, 150)) for _ in range(9)]
    for alpha in (b"ACGT", bytes(range(1, 256))):
        tops = [bytes(rng.choice(alpha) for _ in range(a)) for a, _ in lens]
        sides = [bytes(rng.choice(alpha) for _ in range(b)) for _, b in lens]
        for (m, k, d), grid in (((1, 1, 1), 1), ((2, 1, 2), 2), ((0, 0, 0), 1), ((1, 1, 3), 1), ((3, -1, 0), 2)):
            _check_batch(oracle, tops, sides, m, k, d, grid=grid, bx=1)


def test_bx_is_the_default_when_it_applies(oracle):
    t, s = oracle.generate_pair(0x5EED0B01, 200, 70)
    assert emu.fill_batch([t, t], [s, s], 1, 1, 1)["bx"]                 # short top strings, nibble differences
    assert not emu.fill_batch([t + t, t], [s, s], 1, 1, 1)["bx"]         # a 400-column pair: strips, one pair per warp
    assert not emu.fill_batch([t, t], [s, s], 2, 1, 3)["bx"]             # 2d + m = 8 does not fit a nibble
    # the two kernels agree wherever both apply
    a = emu.fill_batch([t, s], [s, t], 2, 1, 2, bx=1)
    b = emu.fill_batch([t, s], [s, t], 2, 1, 2, bx=0)
    assert list(a["scores"]) == list(b["scores"]) and list(a["branches"]) == list(b["branches"])
    for x, y, top in zip(a["tables"], b["tables"], (t, s)):
        assert np.array_equal(emu.unpack_arrows(x, len(top)) & 7, emu.unpack_arrows(y, len(top)) & 7)


# ---- uniform shapes: pairs swept back to back (nwb_batch_cx_kernel) --------------------------------------

import pytest  # noqa: E402


@pytest.mark.parametrize("a,b,n,grid", [(256, 256, 3, 1), (64, 64, 1, 1), (256, 96, 2, 1), (40, 64, 61, 1), (100, 96, 50, 2), (8, 128, 27, 1), (1, 64, 26, 1),
                                        (255, 64, 49, 1), (9, 160, 30, 1)])
def test_cx_uniform_chains(oracle, a, b, n, grid):
    # several pairs of pairs per warp (12 warps per block): resets, side double-buffering, aligned flush groups,
    # odd batches (the last pair is swept with a copy of itself as partner)
    rng = random.Random(a * 1000 + b + n)
    for alpha, (m, k, d) in ((b"ACGT", (1, 1, 1)), (bytes(range(1, 256)), (2, 1, 2)), (b"AC", (0, 0, 0))):
        tops = [bytes(rng.choice(alpha) for _ in range(a)) for _ in range(n)]
        sides = [bytes(rng.choice(alpha) for _ in range(b)) for _ in range(n)]
        r = emu.fill_batch(tops, sides, m, k, d, grid=grid, bx=2)
        assert r["kernel"] == "cx"
        for i, (t, s) in enumerate(zip(tops, sides)):
            o = oracle.fill(t, s, m, k, d, want_codes=True)
            assert r["scores"][i] == o.final_score, i
            assert r["branches"][i] == o.branch_count, i
            assert np.array_equal(emu.unpack_arrows(r["tables"][i], a) & 7, o.codes[1:, 1:] & 7), i


def test_cx_config4_goldens_and_default(oracle):
    tops, sides = [], []
    for p in (0, 1, 999999):
        t, s = oracle.generate_pair(0x5EED4000 + 2 * p, 256, 256)
        tops.append(t)
        sides.append(s)
    r = emu.fill_batch(tops, sides, 1, 1, 1, grid=1)   # the library's own choice
    assert r["kernel"] == "cx"
    assert list(r["scores"]) == [19, 29, 19]
    assert list(r["branches"]) == [23713, 22912, 22090]
    # not uniform, or rows not a multiple of 32: the general two-pairs-per-warp kernel
    assert emu.fill_batch(tops, [sides[0], sides[1], sides[2][:200]], 1, 1, 1)["kernel"] == "bx"
    assert emu.fill_batch([t[:100] for t in tops], [s[:70] for s in sides], 1, 1, 1)["kernel"] == "bx"


# ---- the count behind -s for a batch (csrc/nwb_batch_count.cuh) -----------------------------------------

def test_batch_count_pass(oracle):
    # after each of the three fill kernels; ragged shapes, pairs wider than one strip (boundary counts through
    # the per-warp scratch line), empty strings, counts that wrap 2^64 (0/0/0: Delannoy numbers)
    rng = random.Random(41)
    lens = [(256, 256), (1, 1), (255, 257), (300, 40), (17, 130), (0, 5), (700, 90), (64, 64), (256, 1), (33, 33),
            (257, 31), (100, 300), (5, 0), (513, 70)] + [(rng.randint(1, 280), rng.randint(1, 150)) for _ in range(12)]
    tops = [bytes(rng.choice(b"ACGT") for _ in range(a)) for a, _ in lens]
    sides = [bytes(rng.choice(b"ACGT") for _ in range(b)) for _, b in lens]
    for (m, k, d), grid in (((1, 1, 1), 1), ((0, 0, 0), 2), ((2, 1, 2), 1)):
        r = emu.fill_batch(tops, sides, m, k, d, grid=grid, count=True)
        assert r["kernel"] == "pk"
        for i, (t, s) in enumerate(zip(tops, sides)):
            assert int(r["counts"][i]) == oracle.fill(t, s, m, k, d).count, (i, len(t), len(s))
    short = [i for i, (a, _) in enumerate(lens) if a <= 256]
    r = emu.fill_batch([tops[i] for i in short], [sides[i] for i in short], 1, 1, 1, grid=1, count=True)
    assert r["kernel"] == "bx"
    for q, i in enumerate(short):
        assert int(r["counts"][q]) == oracle.fill(tops[i], sides[i], 1, 1, 1).count, i


def test_batch_count_config4_goldens(oracle):
    # the three sample pairs of config 4
    tops, sides = [], []
    for p in (0, 1, 999999):
        t, s = oracle.generate_pair(0x5EED4000 + 2 * p, 256, 256)
        tops.append(t)
        sides.append(s)
    r = emu.fill_batch(tops, sides, 1, 1, 1, grid=1, count=True)
    assert r["kernel"] == "cx"
    assert [int(c) for c in r["counts"]] == [387701138034524160, 108460706365440, 4971798065203200]


def test_batch_count_chained_uniform(oracle):
    # uniform one-strip batches: nwb_batch_count_chain_kernel sweeps a warp's run of pairs back to back (lanes reset at
    # table boundaries); more pairs than warps (runs of 3 and 2 pairs on 16 warps), B not a multiple of 32, A < 256,
    # counts that wrap 2^64, and a single run longer than the 64-row ring several times over
    rng = random.Random(43)
    for (a, b, n, mkd, grid) in ((100, 68, 40, (1, 1, 1), 1), (256, 64, 37, (0, 0, 0), 1), (9, 128, 70, (2, 1, 2), 2),
                                 (256, 256, 19, (1, 1, 1), 1)):
        tops = [bytes(rng.choice(b"ACGT") for _ in range(a)) for _ in range(n)]
        sides = [bytes(rng.choice(b"ACGT") for _ in range(b)) for _ in range(n)]
        r = emu.fill_batch(tops, sides, *mkd, grid=grid, count=True)
        for i, (t, s) in enumerate(zip(tops, sides)):
            assert int(r["counts"][i]) == oracle.fill(t, s, *mkd).count, (a, b, i)


def test_batch_general_int32_engine(oracle):
    """csrc/nwb_batch_i32.cuh: schemes outside the packed range, the score matrix and the |score| maximum, one warp
    per pair; pairs wider than one strip, empty strings, more pairs than warps."""
    rng = random.Random(83)
    lens = [(256, 40), (1, 1), (257, 33), (0, 5), (5, 0), (600, 70), (64, 64), (33, 100), (300, 9)] + \
           [(rng.randint(1, 300), rng.randint(1, 90)) for _ in range(20)]
    tops = [bytes(rng.choice(b"ACGT") for _ in range(a)) for a, _ in lens]
    sides = [bytes(rng.choice(b"ACGT") for _ in range(b)) for _, b in lens]
    for (m, k, d), ws in (((1, 3, 1), False), ((5, 4, 3), True), ((-1, 3, -2), False), ((1, 1, 1), True)):
        r = emu.fill_batch_i32(tops, sides, m, k, d, grid=1, want_scores=ws)
        for i, (t, s) in enumerate(zip(tops, sides)):
            o = oracle.fill(t, s, m, k, d, want_codes=True, want_scores=True)
            assert (r["scores"][i], r["branches"][i], r["abs"][i]) == (o.final_score, o.branch_count, o.greatest_abs), (i, m, k, d)
            if len(t) and len(s):
                assert np.array_equal(emu.unpack_arrows(r["tables"][i], len(t)) & 7, o.codes[1:, 1:] & 7), i
                if ws:
                    assert np.array_equal(r["score_rows"][i], o.scores[1:, 1:]), i
