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
        r = emu.fill_batch(tops, sides, m, k, d, grid=grid)
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
    r = emu.fill_batch(tops, sides, 1, 1, 1, grid=1)
    assert list(r["scores"]) == [19, 29, 19]
    assert list(r["branches"]) == [23713, 22912, 22090]
