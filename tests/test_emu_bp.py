"""CPU tests of the bit-parallel batch kernel: the formulation itself (tools/bp_proto.py: Python integers as
row vectors) and the REAL kernel source (csrc/nwb_batch_bp.cuh, one thread per pair) under the SIMT emulator,
both against the oracle -- config 4's golden pairs, ragged shapes, empty strings, top strings with more than
five letters (handed to nwb_batch_pk_kernel through the fallback list), every scheme with 2d + m <= 3."""
import os
import random
import sys

import numpy as np

import emu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import bp_proto  # noqa: E402


def _rand(rng, alphabet, n):
    return bytes(rng.choice(alphabet) for _ in range(n))


def _check(oracle, tops, sides, m, k, d, *, grid=1, warps=2, expect_fallback=None):
    r = emu.fill_batch_bp(tops, sides, m, k, d, grid=grid, warps=warps)
    assert r is not None
    if expect_fallback is not None:
        assert r["n_fallback"] == expect_fallback
    amax = max(len(t) for t in tops)
    for i, (t, s) in enumerate(zip(tops, sides)):
        o = oracle.fill(t, s, m, k, d, want_codes=True)
        assert r["scores"][i] == o.final_score, (i, len(t), len(s))
        assert r["branches"][i] == o.branch_count, (i, len(t), len(s))
        if len(t) and len(s):
            got = emu.unpack_arrows(r["tables"][i], len(t)) & 7
            assert np.array_equal(got, o.codes[1:, 1:] & 7), (i, len(t), len(s))
            # cells beyond the top string, as far as the row vectors reach: no arrows at all (pairs this kernel computed itself)
            width = 64 if amax <= 64 else (128 if amax <= 128 else 256)
            if len(t) < width and len(set(t)) <= 5:
                assert not (emu.unpack_arrows(r["tables"][i], width)[:, len(t):] & 7).any(), i
    return r


def test_formulation_matches_the_oracle(oracle):
    rng = random.Random(3)
    for trial in range(60):
        alpha = rng.choice([b"ACGT", b"ACGT", b"ARNDCQEGHILKMFPSTWYV", b"AB", b"A"])
        t, s = _rand(rng, alpha, rng.randint(1, 90)), _rand(rng, alpha, rng.randint(1, 90))
        while True:
            d, m = rng.randint(0, 3), rng.randint(0, 3)
            k = rng.randint(-2, 2 * d)
            if 0 <= 2 * d - k <= 2 * d + m <= 7:
                break
        o = oracle.fill(t, s, m, k, d, want_codes=True)
        rows, score = bp_proto.bp_fill(t, s, m, k, d)
        assert score == o.final_score
        assert np.array_equal(np.array(rows, dtype=np.uint8), o.codes[1:, 1:] & 7), (trial, m, k, d)


def test_bp_config4_goldens(oracle):
    tops, sides = [], []
    for p in (0, 1, 999999):
        t, s = oracle.generate_pair(0x5EED4000 + 2 * p, 256, 256)
        tops.append(t)
        sides.append(s)
    r = _check(oracle, tops, sides, 1, 1, 1, expect_fallback=0)
    assert list(r["scores"]) == [19, 29, 19]
    assert list(r["branches"]) == [23713, 22912, 22090]


def test_bp_ragged_and_more_pairs_than_lanes(oracle):
    rng = random.Random(17)
    lens = [(256, 256), (1, 1), (255, 257), (17, 130), (0, 5), (64, 64), (256, 1), (33, 33), (100, 300), (5, 0),
            (32, 7), (31, 9), (225, 3)] + [(rng.randint(1, 256), rng.randint(1, 150)) for _ in range(60)]
    tops = [_rand(rng, b"ACGT", a) for a, _ in lens]
    sides = [_rand(rng, b"ACGT", b) for _, b in lens]
    _check(oracle, tops, sides, 1, 1, 1, grid=1, warps=2, expect_fallback=0)   # 73 pairs: 3 groups on 2 warps
    _check(oracle, tops[:40], sides[:40], 1, 1, 1, grid=2, warps=1, expect_fallback=0)


def test_bp_unaligned_strings(oracle):
    # odd lengths make every later string start at an unaligned offset (byte loads instead of word loads)
    rng = random.Random(5)
    lens = [(rng.randint(1, 200) | 1, rng.randint(1, 80) | 1) for _ in range(20)]
    tops = [_rand(rng, b"ACGT", a) for a, _ in lens]
    sides = [_rand(rng, b"ACGT", b) for _, b in lens]
    _check(oracle, tops, sides, 1, 1, 1, expect_fallback=0)


def test_bp_foreign_letters_and_fallback(oracle):
    rng = random.Random(23)
    # side strings may hold any letter (no match vector: never matches); top strings with more than five distinct
    # letters go to nwb_batch_pk_kernel through the list
    tops = [_rand(rng, b"ACGT", 120), _rand(rng, b"ACGTNR", 200), _rand(rng, b"AC", 77), _rand(rng, b"ARNDCQEGHILKMFPSTWYV", 256),
            b"A" * 256, _rand(rng, b"\x00\xff\x80", 90), _rand(rng, b"ACGTUN", 64), _rand(rng, b"ACGTN", 150)]
    sides = [_rand(rng, b"ACGTNX", 100), _rand(rng, b"ACGTN", 90), _rand(rng, b"ACGT", 50), _rand(rng, b"ARNDCQEGHILKMFPSTWYV", 70),
             b"A" * 100, _rand(rng, b"\x00\xff\x80\x7f", 60), _rand(rng, b"ACGTUN", 64), _rand(rng, b"ACGTNX", 99)]
    r = _check(oracle, tops, sides, 1, 1, 1)
    assert r["n_fallback"] == sum(1 for t in tops if len(set(t)) > 5)
    assert r["n_fallback"] >= 2


def test_bp_low_byte_values(oracle):
    # letters 0x00 / 0x01 / 0x02 ... must not be mistaken for the patterns of letter slots that are not in use yet
    rng = random.Random(59)
    tops, sides = [], []
    for alpha in (b"\x00\x01", b"\x01", b"\x00", b"\x01\x02\x03\x04", b"\x00\x01\x02\x03\x04", b"\x01\x00\xff"):
        for n in (5, 64, 200, 256):
            tops.append(_rand(rng, alpha, n))
            sides.append(_rand(rng, alpha + b"\x01\x00", 70))
    _check(oracle, tops, sides, 1, 1, 1, expect_fallback=0)


def test_bp_every_small_scheme(oracle):
    rng = random.Random(29)
    tops = [_rand(rng, b"ACGT", a) for a in (256, 40, 130, 1, 77)]
    sides = [_rand(rng, b"ACGT", b) for b in (60, 90, 33, 12, 1)]
    seen = set()
    for d in range(0, 2):
        for m in range(0, 4):
            for k in range(-3, 3):
                M, N = 2 * d + m, 2 * d - k
                if not (1 <= M <= 3 and 0 <= N <= M):
                    continue
                seen.add((M, N))
                _check(oracle, tops, sides, m, k, d)
    assert len(seen) == 9   # every instantiated (M, N)


def test_bp_narrow_tables(oracle):
    """Row vectors of 64 and 128 bits (top strings of at most 64 / 128 letters), as the library picks them and forced
    wider; uniform groups (the instantiation without masks) and ragged ones."""
    rng = random.Random(31)
    for amax, n in ((64, 40), (128, 36), (33, 34), (100, 35)):
        lens = [(amax, 50)] * 33 + [(rng.randint(1, amax), rng.randint(1, 90)) for _ in range(n - 33)] if n > 33 else []
        lens = lens or [(amax, 50)] * n
        tops = [_rand(rng, b"ACGT", a) for a, _ in lens]
        sides = [_rand(rng, b"ACGT", b) for _, b in lens]
        for words in (0, 8) if amax > 64 else (0, 4):
            r = emu.fill_batch_bp(tops, sides, 1, 1, 1, grid=1, warps=2, words=words)
            assert r is not None and r["n_fallback"] == 0
            for i, (t, s) in enumerate(zip(tops, sides)):
                o = oracle.fill(t, s, 1, 1, 1, want_codes=True)
                assert (r["scores"][i], r["branches"][i]) == (o.final_score, o.branch_count), (amax, words, i)
                assert np.array_equal(emu.unpack_arrows(r["tables"][i], len(t)) & 7, o.codes[1:, 1:] & 7), (amax, words, i)
    # other schemes on the narrow instantiations
    tops = [_rand(rng, b"ACGT", a) for a in (64, 40, 13, 1, 57)]
    sides = [_rand(rng, b"ACGT", b) for b in (60, 90, 33, 12, 1)]
    for m, k, d in ((0, 0, 1), (1, -1, 0), (3, 0, 0), (0, 1, 1), (1, 2, 1)):
        _check(oracle, tops, sides, m, k, d)
    assert emu.fill_batch_bp([b"A" * 100], [b"A" * 10], 1, 1, 1, words=2) is None   # does not fit 64 bits


def test_bp_refuses_what_it_cannot_do():
    assert emu.fill_batch_bp([b"A" * 300], [b"A" * 10], 1, 1, 1) is None     # wider than one strip
    assert emu.fill_batch_bp([b"ACGT"], [b"ACGT"], 2, 1, 2) is None          # 2d + m = 6
