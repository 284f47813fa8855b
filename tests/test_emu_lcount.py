"""CPU tests of the per-lane sparse backward count of a batch: the sweep itself (tools/lcount_proto.py) and the REAL
kernel source (csrc/nwb_batch_lcount.cuh, one thread per pair) under the SIMT emulator, against the oracle --
config 4's golden counts, ragged shapes, empty strings, pairs wider than a strip, all-ties tables and long gap runs
(which make the sweep give up: the dense kernel behind it must deliver the count), garbage beyond the top string."""
import os
import random
import sys

import numpy as np

import emu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import lcount_proto  # noqa: E402


def _rand(rng, alphabet, n):
    return bytes(rng.choice(alphabet) for _ in range(n))


def _tables(oracle, tops, sides, m, k, d, rng=None):
    tabs, counts = [], []
    for t, s in zip(tops, sides):
        pitch = max(1, (len(t) + 255) // 256) * 128
        o = oracle.fill(t, s, m, k, d, want_packed=True, pitch=pitch)
        tab = o.packed.copy() if len(s) else np.zeros((0, pitch), np.uint8)
        if rng is not None and tab.size:
            # what the fill kernels may leave behind: bit 3 of every nibble and whole nibbles beyond column A
            tab |= np.frombuffer(bytes(rng.getrandbits(8) & 0x88 for _ in range(tab.size)), np.uint8).reshape(tab.shape)
            a = len(t)
            if a & 1:
                tab[:, a // 2] |= 0x70
            tab[:, (a + 1) // 2:] = 0x77
        tabs.append(tab)
        counts.append(o.count)
    return tabs, counts


def test_sweep_matches_the_oracle(oracle):
    rng = random.Random(3)
    done = 0
    for trial in range(80):
        a, b = rng.randint(1, 120), rng.randint(1, 120)
        if trial < 30:
            b = max(1, a + rng.randint(-10, 10))
        alpha = rng.choice([b"ACGT", b"ACGT", b"AC", b"ARNDCQEGHILKMFPSTWYV"])
        t, s = _rand(rng, alpha, a), _rand(rng, alpha, b)
        m, k, d = rng.choice([(1, 1, 1), (2, 1, 2), (1, 0, 1)])
        o = oracle.fill(t, s, m, k, d, want_codes=True)
        c, gave_up = lcount_proto.lcount(o.codes, a, b)
        if not gave_up:
            assert c == o.count, (trial, a, b)
            done += 1
    assert done >= 30


def test_lcount_config4_goldens(oracle):
    tops, sides = [], []
    for p in (0, 1, 999999, 5, 6, 7):
        t, s = oracle.generate_pair(0x5EED4000 + 2 * p, 256, 256)
        tops.append(t)
        sides.append(s)
    tabs, want = _tables(oracle, tops, sides, 1, 1, 1)
    counts, nfb = emu.batch_lcount(tops, sides, tabs)
    assert list(counts[:3]) == [387701138034524160, 108460706365440, 4971798065203200]
    assert [int(c) for c in counts] == want
    assert nfb <= 2


def test_lcount_ragged_and_giving_up(oracle):
    rng = random.Random(41)
    lens = [(256, 256), (1, 1), (255, 257), (17, 130), (0, 5), (64, 64), (256, 1), (33, 33), (100, 300), (5, 0),
            (300, 40), (700, 90), (1, 200), (40, 40), (41, 39)] + [(rng.randint(1, 300), rng.randint(1, 200)) for _ in range(50)] + \
           [(n, n + rng.randint(-6, 6)) for n in (50, 80, 120, 200, 256, 290) for _ in range(3)]
    tops = [_rand(rng, b"ACGT", a) for a, _ in lens]
    sides = [_rand(rng, b"ACGT", max(0, b)) for _, b in lens]
    for (m, k, d), garbage in (((1, 1, 1), None), ((2, 1, 2), rng), ((0, 0, 0), None)):
        tabs, want = _tables(oracle, tops, sides, m, k, d, garbage)
        counts, nfb = emu.batch_lcount(tops, sides, tabs, grid=2, warps=2)
        assert [int(c) for c in counts] == want, (m, k, d)
        if (m, k, d) == (0, 0, 0):
            assert nfb > len(lens) // 2      # every cell ties three ways: the band is the whole table
        else:
            assert 0 < nfb < len(lens)
