"""GPU parity tests proper: the CUDA path, called through the C ABI
(libnwb.so via ctypes), against the CPU oracle and the committed goldens.
Bit-exact: integer scores, arrow sets, counts mod 2^64, branch counts."""
import json
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def golden(name):
    with open(os.path.join(HERE, "golden", name)) as f:
        return json.load(f)


def case_strings(oracle, c):
    if "seed" in c:
        alpha = oracle.DNA if c["alphabet"] == "dna" else oracle.PROTEIN
        return oracle.generate_pair(c["seed"], c["top_len"], c["side_len"], alpha)
    return c["top"].encode(), c["side"].encode()


def check_arrows(oracle, nwb, tab, t, s, m, k, d):
    """Whole nibble table against the oracle's, byte for byte over the valid region."""
    a, b = len(t), len(s)
    if a == 0 or b == 0:
        return
    rows = tab.arrow_rows()
    o = oracle.fill(t, s, m, k, d, want_packed=True, pitch=tab.pitch)
    nb = (a + 1) // 2
    got, exp = rows[:, :nb].copy(), o.packed[:, :nb].copy()
    got &= 0x77
    if a & 1:
        got[:, nb - 1] &= 0x07
    assert np.array_equal(got, exp)
    return o


def pk_supported(m, k, d):
    """The packed 16x2 kernel's precondition (csrc/nwb_fill_pk.cuh nwb_pk_supported)."""
    return 0 <= 2 * d - k <= 2 * d + m <= 4000 and m + k <= 128


def full_check(oracle, nwb, t, s, m, k, d, extra_flags=0, num_gpus=1):
    """Count-fused general kernel AND (without count) the automatically selected
    kernel -- the packed 16x2 one whenever the scheme allows it."""
    o = oracle.fill(t, s, m, k, d)
    flags = nwb.WANT_COUNT | nwb.WANT_ARROWS_HOST | extra_flags
    tab = nwb.fill(t, s, m, k, d, flags, num_gpus=num_gpus)
    assert tab.opt_score == o.final_score
    assert tab.branch_count == o.branch_count
    assert tab.count == o.count
    if extra_flags & nwb.TRACK_ABS:
        assert tab.greatest_abs == o.greatest_abs
    if len(t) and len(s) and not (extra_flags & (nwb.FORCE_GENERAL | nwb.TRACK_ABS)):
        assert tab.kernel_kind == (nwb.KIND_PK if pk_supported(m, k, d) else nwb.KIND_I32)
    check_arrows(oracle, nwb, tab, t, s, m, k, d)
    if not (extra_flags & (nwb.FORCE_GENERAL | nwb.TRACK_ABS)):
        tab2 = nwb.fill(t, s, m, k, d, nwb.WANT_ARROWS_HOST, num_gpus=num_gpus)
        if len(t) and len(s):
            assert tab2.kernel_kind == (nwb.KIND_PK if pk_supported(m, k, d) else nwb.KIND_I32)
        assert tab2.opt_score == o.final_score
        assert tab2.branch_count == o.branch_count
        check_arrows(oracle, nwb, tab2, t, s, m, k, d)
    return tab


@pytest.mark.parametrize("force", [True, False], ids=["general", "auto"])
def test_readme_example_every_cell(oracle, nwb, force):
    """BASELINE config 1: GCATGCU/GATTACA 1 1 1 -s -t, every score, arrow set, count."""
    ff = nwb.FORCE_GENERAL if force else 0
    tab = nwb.fill(b"GCATGCU", b"GATTACA", 1, 1, 1,
                   nwb.WANT_SCORES | nwb.WANT_COUNT | nwb.WANT_ARROWS_HOST | nwb.TRACK_ABS | nwb.WANT_COUNT_MATRIX | ff)
    o = oracle.fill("GCATGCU", "GATTACA", 1, 1, 1, want_scores=True, want_codes=True, want_counts=True)
    for j in range(8):
        for i in range(8):
            assert tab.score(i, j) == o.scores[j, i]
            assert tab.arrows(i, j) == o.codes[j, i]
            assert tab.count_at(i, j) == o.counts[j, i]
    assert (tab.opt_score, tab.count, tab.branch_count, tab.greatest_abs) == (0, 3, 12, 5)
    tab0 = nwb.fill(b"GCATGCU", b"GATTACA", 0, 0, 0, nwb.WANT_COUNT | ff)
    assert tab0.count == 48639 and tab0.branch_count == 49


@pytest.mark.parametrize("force", [True, False], ids=["general", "auto"])
@pytest.mark.parametrize("case", golden("golden.json"), ids=lambda c: c["name"])
def test_goldens(oracle, nwb, case, force):
    t, s = case_strings(oracle, case)
    ff = nwb.FORCE_GENERAL if force else 0
    tab = full_check(oracle, nwb, t, s, case["m"], case["k"], case["d"], ff)
    tab = full_check(oracle, nwb, t, s, case["m"], case["k"], case["d"], nwb.TRACK_ABS | ff)
    assert tab.opt_score == case["final_score"]
    assert tab.branch_count == case["branch_count"]
    assert tab.greatest_abs == case["greatest_abs"]
    assert tab.count == case["count_u64"]


@pytest.mark.parametrize("force", [True, False], ids=["general", "auto"])
def test_edge_shapes(oracle, nwb, force):
    """Lengths around lane / strip / chunk edges, ragged, arbitrary bytes."""
    ff = nwb.FORCE_GENERAL if force else 0
    rng = random.Random(7)
    sizes = [1, 2, 31, 32, 33, 63, 64, 65, 127, 128, 129, 191, 192, 193, 255, 256, 257, 511, 513, 769]
    for _ in range(40):
        a, b = rng.choice(sizes), rng.choice(sizes)
        alpha = rng.choice([b"ACGT", b"AB", bytes(range(1, 256)), b"ARNDCQEGHILKMFPSTWYV"])
        t = bytes(rng.choice(alpha) for _ in range(a))
        s = bytes(rng.choice(alpha) for _ in range(b))
        m, k, d = rng.choice([(1, 1, 1), (2, 1, 2), (0, 0, 0), (1, 2, 3), (5, 4, 3), (-1, 3, -2), (3, -1, 0), (1, 1, 0)])
        full_check(oracle, nwb, t, s, m, k, d, ff)


def test_scores_and_count_matrix(oracle, nwb):
    """Every score cell and every intermediate count (SURVEY hard part 6: the
    final count is 0 mod 2^64 on the big configs, so intermediates must match)."""
    for seed, a, b, (m, k, d) in [(0x5EED0A00, 700, 900, (1, 1, 1)), (0x5EED0A02, 1300, 517, (2, 1, 2))]:
        t, s = oracle.generate_pair(seed, a, b)
        tab = nwb.fill(t, s, m, k, d, nwb.WANT_SCORES | nwb.WANT_COUNT_MATRIX | nwb.TRACK_ABS)
        o = oracle.fill(t, s, m, k, d, want_scores=True, want_counts=True)
        assert np.array_equal(tab.score_rows(), o.scores[1:, 1:])
        assert np.array_equal(tab.count_rows(), o.counts[1:, 1:])
        assert tab.greatest_abs == o.greatest_abs and tab.count == o.count


def test_empty_top_string(oracle, nwb):
    # leading-whitespace input gives an empty top string
    tab = nwb.fill(b"", b"ACG", 1, 1, 1, nwb.WANT_COUNT | nwb.WANT_ARROWS_HOST | nwb.WANT_SCORES)
    assert (tab.opt_score, tab.count, tab.branch_count) == (-3, 1, 0)
    assert [tab.arrows(0, j) for j in range(4)] == [0, nwb.UP, nwb.UP, nwb.UP]
    assert [tab.score(0, j) for j in range(4)] == [0, -1, -2, -3]


@pytest.mark.parametrize("force", [True, False], ids=["general", "auto"])
def test_config2_dna_10k(oracle, nwb, force):
    """BASELINE config 2 (10k x 10k DNA, 1 1 1, -q -s) vs the reference's goldens."""
    t, s = oracle.generate_pair(0x5EED0002, 10000, 10000)
    ff = nwb.FORCE_GENERAL if force else 0
    tab = nwb.fill(t, s, 1, 1, 1, nwb.WANT_COUNT | nwb.WANT_ARROWS_HOST | ff)
    assert tab.kernel_kind == (nwb.KIND_I32 if force else nwb.KIND_PK)
    assert (tab.opt_score, tab.branch_count, tab.count) == (1056, 34377799, 0)
    check_arrows(oracle, nwb, tab, t, s, 1, 1, 1)


def test_config5_protein_30k_summary(oracle, nwb):
    """BASELINE config 5 (30k x 30k protein, 2 1 2, -q -s): summary vs golden_big.json."""
    g = [c for c in golden("golden_big.json") if c["name"] == "config5_protein_30k"]
    if not g:
        pytest.skip("golden_big.json has no config 5 entry")
    g = g[0]
    t, s = oracle.generate_pair(0x5EED0005, 30000, 30000, oracle.PROTEIN)
    tab = nwb.fill(t, s, 2, 1, 2, nwb.WANT_COUNT)
    assert (tab.opt_score, tab.branch_count, tab.count) == (g["final_score"], g["branch_count"], g["count_u64"])
    assert tab.kernel_kind == nwb.KIND_PK
    tab = nwb.fill(t, s, 2, 1, 2, nwb.WANT_COUNT | nwb.FORCE_GENERAL)
    assert tab.kernel_kind == nwb.KIND_I32
    assert (tab.opt_score, tab.branch_count, tab.count) == (g["final_score"], g["branch_count"], g["count_u64"])


def test_config3_dna_100k_properties(oracle, nwb):
    """BASELINE config 3 (100k x 100k DNA) at full size, by size-independent
    properties: (1) the first R rows of the table equal the table of
    (top, side[:R]) -- checked against the oracle for R = 1500, which crosses
    every strip hand-off; (2) swapping the strings transposes the table, so
    score, count and branch count are equal; (3) golden summary if recorded."""
    t, s = oracle.generate_pair(0x5EED0030, 100000, 100000)
    g = [c for c in golden("golden_big.json") if c["name"] == "config3_dna_100k"]
    R = 1500
    o = None
    for flags, kind in ((0, nwb.KIND_PK), (nwb.WANT_COUNT, nwb.KIND_PK), (nwb.WANT_COUNT | nwb.FORCE_GENERAL, nwb.KIND_I32)):
        plan = nwb.Plan(100000, 100000, flags)
        plan.upload(t, s)
        plan.run(1, 1, 1)
        sm = plan.summary()
        assert sm.kernel_kind == kind
        rows = plan.download_arrows(0, R)
        if o is None or o.packed.shape[1] != rows.shape[1]:
            o = oracle.fill(t, s[:R], 1, 1, 1, want_packed=True, pitch=rows.shape[1])
        assert np.array_equal(rows[:, :50000] & 0x77, o.packed[:, :50000])
        plan.upload(s, t)
        plan.run(1, 1, 1)
        sm2 = plan.summary()
        assert (sm.opt_score, sm.count, sm.branch_count) == (sm2.opt_score, sm2.count, sm2.branch_count)
        if g:
            assert (sm.opt_score, sm.branch_count) == (g[0]["final_score"], g[0]["branch_count"])
            if flags & nwb.WANT_COUNT:
                assert sm.count == g[0]["count_u64"]
        plan.close()


def test_two_gpu_strips_match_one(oracle, nwb):
    if nwb.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    t, s = oracle.generate_pair(0x5EED0B00, 5000, 3000)
    full_check(oracle, nwb, t, s, 1, 1, 1, num_gpus=2)
    full_check(oracle, nwb, t, s, 2, 1, 2, nwb.FORCE_GENERAL, num_gpus=2)


def test_queue_mode_on_one_gpu(oracle, nwb):
    """NWB_QUEUE: several plans on one GPU take a queue of different fills round robin, each on its own stream, so that
    consecutive fills overlap (ticketed blocks sweeping adjacent strips, ordinary launches).  Every fill is compared
    with the one-call fill of the same pair (score, branch count, digest of the whole arrow table); one shape has more
    blocks than the GPU has SMs (157 blocks of three strips: the late blocks start when earlier ones have left), one
    has a ragged last block, one runs the count behind the fill; nwb_tune hx_spb = 1, 2 puts fewer strips in a block."""
    for (A, B, mkd, flags, spb) in ((120400, 4200, (1, 1, 1), 0, 0), (30000, 20000, (2, 1, 2), 0, 0), (5100, 9000, (1, 1, 1), nwb.WANT_COUNT, 0),
                                    (20000, 6000, (1, 1, 1), 0, 1), (20000, 6000, (1, 1, 2), 0, 2)):
        pairs = [oracle.generate_pair(0x5EED0D00 + 16 * i, A, B) for i in range(3)]
        want = []
        for (t, s) in pairs:
            tab = nwb.fill(t, s, *mkd, flags | nwb.WANT_DIGEST)
            want.append((tab.opt_score, tab.branch_count, tab.arrow_digest(), tab.count))
            tab.close()
        nwb.cache_clear()
        with nwb.tuned(hx_spb=spb):
            plans = [nwb.Plan(A, B, flags | nwb.QUEUE) for _ in range(3)]
            queue = [0, 1, 2, 2, 0, 1, 1, 0, 2, 0, 1]
            which = [None] * len(plans)
            for step, q in enumerate(queue + [None] * len(plans)):
                pl = plans[step % len(plans)]
                if which[step % len(plans)] is not None:        # the fill this plan took three steps ago
                    sm = pl.summary()
                    assert sm.kernel_kind == 1 and pl.kernel_name() == "nwb_fill_hx_kernel"
                    got = (sm.opt_score, sm.branch_count, pl.arrow_digest(), sm.count)
                    assert got == want[which[step % len(plans)]], (A, B, mkd, step)
                    which[step % len(plans)] = None
                if q is not None:
                    pl.upload(*pairs[q])
                    pl.run(*mkd)
                    which[step % len(plans)] = q
            for pl in plans:
                pl.close()


def test_queue_mode_small_tables(oracle, nwb):
    """Queue mode on small tables forced onto the hx kernel (nwb_tune pk_hx = 1): 1 / 2 / 3 adjacent strips per block,
    ragged last blocks, a table smaller than one strip, the count behind the fill; two plans, every fill against the
    ORACLE (score, branch count, digest of every arrow set, count)."""
    with nwb.tuned(pk_hx=1):
        for spb in (0, 1, 2):
            nwb.tune("hx_spb", spb)
            for (a, b, flags) in ((1500, 300, 0), (700, 260, nwb.WANT_COUNT), (2100, 64, 0), (7, 7, 0)):
                pairs = [oracle.generate_pair(0x5EED0F80 + 2 * i + a, a, b) for i in range(2)]
                want = []
                for t, s in pairs:
                    o = oracle.fill(t, s, 2, 1, 2)
                    want.append((o.final_score, o.branch_count, o.arrow_digest, o.count if flags else 0))
                plans = [nwb.Plan(a, b, flags | nwb.QUEUE) for _ in range(2)]
                took = [None, None]
                for step, q in enumerate([0, 1, 1, 0, 0, None, None]):
                    pl = plans[step % 2]
                    if took[step % 2] is not None:
                        sm = pl.summary()
                        assert pl.kernel_name() == "nwb_fill_hx_kernel", pl.kernel_name()
                        assert (sm.opt_score, sm.branch_count, pl.arrow_digest(), sm.count) == want[took[step % 2]], (spb, a, b, step)
                    if q is not None:
                        pl.upload(*pairs[q])
                        pl.run(2, 1, 2)
                    took[step % 2] = q
                for pl in plans:
                    pl.close()


def _rank_share(pl):
    sm = pl.summary()                           # waits for this rank's last fill only
    b, e = pl.strip_range()
    return (sm.partial_r, sm.branch_count, pl.arrow_digest() if e > b else 0, sm.count, sm.opt_score, sm.kernel_kind, e > b)


def _group_result(nwb, shares, A, B, d):
    """(score, branch count, table digest, count) of a strip group from its ranks' shares (what bench.py all-reduces).
    Trailing ranks can be empty (fewer strips than ranks): the count and, for the general kernel, the score are the last
    non-empty rank's."""
    last = [x for x in shares if x[6]][-1]
    score = nwb.strip_group_score(sum(x[0] for x in shares), A, B, d) if shares[0][5] == 1 else last[4]
    return (score, sum(x[1] for x in shares) & 0xFFFFFFFF, sum(x[2] for x in shares) & 0xFFFFFFFFFFFFFFFF, last[3])


def test_pipelined_strip_group(oracle, nwb):
    """nwb_plan_run_pipelined: a strip group works through a queue of DIFFERENT fills with no barrier and no inbox reset
    between them (double-buffered inboxes + acknowledgement word).  Rank g collects its share of fill e and launches
    fill e + 1 before rank g + 1 has been looked at, so neighbouring ranks are on different fills; every fill of the
    queue is checked (score, branch count, digest of every rank's arrow columns, count) against the one-GPU fill of the
    same pair, whose results the other tests pin to the oracle.  Then one pair is filled 12 times back to back with
    nothing on the host between the launches, and the last fill is checked."""
    ndev = nwb.device_count()
    if ndev < 2:
        pytest.skip("needs 2 GPUs")
    for world in [w for w in (2, 4, 8) if w <= ndev]:
        for (A, B, mkd, flags) in ((10000, 10000, (1, 1, 1), 0), (6000, 9000, (2, 1, 2), 0), (3000, 2500, (5, 4, 3), 0),
                                   (5000, 6000, (1, 1, 1), nwb.WANT_COUNT), (3000, 2500, (1, 3, 1), nwb.FORCE_GENERAL),
                                   (200, 5000, (1, 1, 1), 0)):   # one strip: every rank but the first is empty
            pairs = [oracle.generate_pair(0x5EED0C00 + 16 * i, A, B) for i in range(3)]
            want = []
            for (t, s) in pairs:
                tab = nwb.fill(t, s, *mkd, flags | nwb.WANT_DIGEST)
                want.append((tab.opt_score, tab.branch_count, tab.arrow_digest(), tab.count))
                tab.close()
            plans = [nwb.Plan(A, B, flags, device=g, strip_rank=g, strip_world=world) for g in range(world)]
            for g in range(world - 1):
                plans[g].attach_right(plans[g + 1])
            queue = [0, 1, 2, 1, 0, 2, 2]
            for step in range(len(queue) + 1):
                shares = []
                for pl in plans:
                    if step > 0:
                        shares.append(_rank_share(pl))
                    if step < len(queue):
                        pl.upload(*pairs[queue[step]])
                        pl.run_pipelined(*mkd)
                if step > 0:
                    assert _group_result(nwb, shares, A, B, mkd[2]) == want[queue[step - 1]], (world, A, B, mkd, step)
            for pl in plans:
                pl.upload(*pairs[1])
            for _ in range(12):
                for pl in plans:
                    pl.run_pipelined(*mkd)
            assert _group_result(nwb, [_rank_share(pl) for pl in plans], A, B, mkd[2]) == want[1], (world, A, B, mkd, "back to back")
            for pl in plans:
                pl.close()
    nwb.cache_clear()


def test_strip_group_full_size_digests(oracle, nwb):
    """Column strips over 2 (and 4, 8 when present) GPUs of one process at full size: EVERY arrow set (the ranks' table
    digests add up to the oracle's), score, branch count and the count (the sparse backward sweep on the rank that owns
    column A -- on these inputs the live cells die long before the band reaches that rank's first column -- and, with
    nwb_tune count_mode = 2, the dense sweep handed from GPU to GPU)."""
    ndev = nwb.device_count()
    if ndev < 2:
        pytest.skip("needs 2 GPUs")
    gold = {c["name"]: c for c in golden("golden_big.json")}
    for name, seed, alpha, mkd in (("config2_dna_10k", 0x5EED0002, "dna", (1, 1, 1)), ("config5_protein_30k", 0x5EED0005, "protein", (2, 1, 2)),
                                   ("config3_dna_100k", 0x5EED0030, "dna", (1, 1, 1))):
        g = gold[name]
        t, s = oracle.generate_pair(seed, g["top_len"], g["side_len"], oracle.DNA if alpha == "dna" else oracle.PROTEIN)
        for world in [w for w in (2, 4, 8) if w <= ndev]:
            tab = nwb.fill(t, s, *mkd, nwb.WANT_DIGEST | nwb.WANT_COUNT, num_gpus=world)
            assert (tab.opt_score, tab.branch_count, tab.count) == (g["final_score"], g["branch_count"], g["count_u64"]), (name, world)
            assert tab.arrow_digest() == int(g["arrow_digest"], 16), (name, world)
            assert tab.summary().count_path == nwb.COUNT_SPARSE, (name, world, tab.summary().count_path)
            tab.close()
            if world == 2:
                with nwb.tuned(count_mode=2):
                    tab = nwb.fill(t, s, *mkd, nwb.WANT_COUNT, num_gpus=world)
                assert (tab.opt_score, tab.branch_count, tab.count) == (g["final_score"], g["branch_count"], g["count_u64"]), (name, world)
                assert tab.summary().count_path == nwb.COUNT_DENSE
                tab.close()
            nwb.cache_clear()


def test_strip_group_count_falls_back_to_the_dense_sweep(oracle, nwb):
    """A string against a mutated copy of itself on 2 GPUs: the count is not 0 mod 2^64, the live band runs through all
    30,000 rows and leaves the last rank's columns, so the sparse sweep on that rank gives up and the group's dense
    sweep delivers the count."""
    if nwb.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    g = [c for c in golden("golden_big.json") if c["name"] == "mutated_dna_30k"][0]
    (t, s), (m, k, d) = _mutated("mutated_dna_30k")
    tab = nwb.fill(t, s, m, k, d, nwb.WANT_COUNT | nwb.WANT_DIGEST, num_gpus=2)
    assert (tab.opt_score, tab.branch_count, tab.count) == (g["final_score"], g["branch_count"], g["count_u64"])
    assert tab.arrow_digest() == int(g["arrow_digest"], 16)
    assert tab.summary().count_path == nwb.COUNT_SPARSE_BAILED
    tab.close()
    nwb.cache_clear()


# ---- batch of independent pairs (BASELINE config 4) ---------------------------------
def test_batch_ragged(oracle, nwb):
    rng = random.Random(11)
    lens = [(256, 256), (1, 1), (255, 257), (300, 40), (17, 130), (0, 5), (700, 90), (64, 64), (256, 1),
            (33, 33), (257, 31), (100, 300), (5, 0), (1500, 700)] + \
           [(rng.randint(1, 600), rng.randint(1, 500)) for _ in range(300)]
    tops = [bytes(rng.choice(b"ACGT") for _ in range(a)) for a, _ in lens]
    sides = [bytes(rng.choice(b"ACGT") for _ in range(b)) for _, b in lens]
    for m, k, d in ((1, 1, 1), (2, 1, 2)):
        bt = nwb.Batch(tops, sides, m, k, d, nwb.WANT_ARROWS_HOST)
        bt.run()
        bt.fetch()
        for i, (t, s) in enumerate(zip(tops, sides)):
            o = oracle.fill(t, s, m, k, d, want_packed=True, pitch=max(1, (len(t) + 255) // 256) * 128)
            assert bt.opt_score(i) == o.final_score, (i, len(t), len(s))
            assert bt.branch_count(i) == o.branch_count, (i, len(t), len(s))
            if len(t) and len(s):
                nb = (len(t) + 1) // 2
                got = bt.arrow_rows(i)[:, :nb] & 0x77
                if len(t) & 1:
                    got[:, nb - 1] &= 0x07
                assert np.array_equal(got, o.packed[:, :nb]), i
        bt.close()


def test_batch_general_schemes_scores_and_abs(oracle, nwb):
    """Schemes outside the packed kernels' range ((1,3,1), (5,4,3), negative penalties), NWB_FORCE_GENERAL,
    NWB_WANT_SCORES and NWB_TRACK_ABS through nwb_fill_batch: the int32 batch engine (csrc/nwb_batch_i32.cuh)
    against the oracle -- every arrow set, score matrix, |score| maximum, branch count and alignment count."""
    rng = random.Random(71)
    lens = [(256, 256), (1, 1), (255, 257), (300, 40), (17, 130), (0, 5), (700, 90), (64, 64), (256, 1), (33, 33),
            (257, 31), (5, 0), (1100, 300)] + [(rng.randint(1, 600), rng.randint(1, 400)) for _ in range(120)]
    tops = [bytes(rng.choice(b"ACGT") for _ in range(a)) for a, _ in lens]
    sides = [bytes(rng.choice(b"ACGT") for _ in range(b)) for _, b in lens]
    for (m, k, d), flags in (((1, 3, 1), 0), ((1, 5, 2), nwb.WANT_COUNT), ((-1, 3, -2), nwb.WANT_COUNT),
                             ((70, 60, 50), nwb.WANT_COUNT), ((1, 1, 1), nwb.FORCE_GENERAL | nwb.WANT_COUNT),
                             ((2, 1, 2), nwb.WANT_SCORES | nwb.TRACK_ABS | nwb.WANT_COUNT)):
        assert (flags & (nwb.FORCE_GENERAL | nwb.WANT_SCORES)) or not pk_supported(m, k, d)
        bt = nwb.Batch(tops, sides, m, k, d, nwb.WANT_ARROWS_HOST | flags)
        assert bt.kernel_name() == "nwb_batch_i32_kernel"
        bt.run()
        bt.fetch()
        for i in list(range(13)) + rng.sample(range(13, len(lens)), 40):
            t, s = tops[i], sides[i]
            o = oracle.fill(t, s, m, k, d, want_packed=True, want_scores=True, pitch=max(1, (len(t) + 255) // 256) * 128)
            assert bt.opt_score(i) == o.final_score, (i, len(t), len(s), m, k, d)
            assert bt.branch_count(i) == o.branch_count, i
            if flags & nwb.WANT_COUNT:
                assert bt.count(i) == o.count, i
            if flags & nwb.TRACK_ABS:
                assert bt.greatest_abs(i) == o.greatest_abs, i
            if len(t) and len(s):
                nb = (len(t) + 1) // 2
                got = bt.arrow_rows(i)[:, :nb] & 0x77
                if len(t) & 1:
                    got[:, nb - 1] &= 0x07
                assert np.array_equal(got, o.packed[:, :nb]), i
                if flags & nwb.WANT_SCORES:
                    assert np.array_equal(bt.score_rows(i), o.scores[1:, 1:]), i
        bt.close()
    with pytest.raises(nwb.NwbError):
        nwb.Batch(tops[:2], sides[:2], 1, 1, 1, nwb.WANT_COUNT_MATRIX)


def test_batch_config4_shard(oracle, nwb):
    """One GPU's shard of config 4 at reduced count (20,000 pairs of 256 x 256, generator seeds
    0x5EED4000 + 2p): the three SURVEY goldens and a random sample against the oracle."""
    n = 20000
    idx = list(range(n - 1)) + [999999]
    tops, sides = zip(*(oracle.generate_pair(0x5EED4000 + 2 * p, 256, 256) for p in idx))
    bt = nwb.Batch(list(tops), list(sides), 1, 1, 1, nwb.WANT_ARROWS_HOST)
    assert bt.kernel_name() == "nwb_batch_bp_kernel"     # the default for 256-column DNA 1/1/1 pairs
    bt.run()
    bt.fetch()
    assert (bt.opt_score(0), bt.branch_count(0)) == (19, 23713)
    assert (bt.opt_score(1), bt.branch_count(1)) == (29, 22912)
    assert (bt.opt_score(n - 1), bt.branch_count(n - 1)) == (19, 22090)
    rng = random.Random(5)
    for i in rng.sample(range(n), 60):
        o = oracle.fill(tops[i], sides[i], 1, 1, 1, want_packed=True, pitch=128)
        assert bt.opt_score(i) == o.final_score and bt.branch_count(i) == o.branch_count
        assert np.array_equal(bt.arrow_rows(i) & 0x77, o.packed)
    bt.close()


def _batch_check(oracle, nwb, bt, tops, sides, m, k, d, sample):
    for i in sample:
        t, s = tops[i], sides[i]
        o = oracle.fill(t, s, m, k, d, want_packed=True, pitch=max(1, (len(t) + 255) // 256) * 128)
        assert bt.opt_score(i) == o.final_score, (i, len(t), len(s))
        assert bt.branch_count(i) == o.branch_count, (i, len(t), len(s))
        if len(t) and len(s):
            nb = (len(t) + 1) // 2
            got = bt.arrow_rows(i)[:, :nb] & 0x77
            if len(t) & 1:
                got[:, nb - 1] &= 0x07
            assert np.array_equal(got, o.packed[:, :nb]), (i, len(t), len(s))


def test_batch_two_pairs_per_warp(oracle, nwb, monkeypatch):
    """csrc/nwb_batch_bx.cuh: top strings of at most 256 characters, ragged partners sharing a warp,
    empty strings, an odd number of pairs, more pair-pairs than warps; and the same batch through the
    one-pair-per-warp kernel (nwb_tune batch_bx = 0)."""
    rng = random.Random(29)
    lens = [(256, 256), (1, 1), (255, 257), (3, 40), (17, 130), (0, 5), (200, 90), (64, 64), (256, 1), (33, 33),
            (256, 31), (100, 300), (5, 0), (8, 32), (9, 33), (249, 63), (250, 64), (7, 65), (1, 200), (256, 2),
            (0, 0), (31, 31), (130, 95), (96, 128), (256, 1000)] + \
           [(rng.randint(1, 256), rng.randint(1, 400)) for _ in range(4000)]
    for alpha, schemes in ((b"ACGT", ((1, 1, 1), (0, 0, 0))), (bytes(range(1, 256)), ((2, 1, 2), (1, 1, 3)))):
        tops = [bytes(rng.choice(alpha) for _ in range(a)) for a, _ in lens]
        sides = [bytes(rng.choice(alpha) for _ in range(b)) for _, b in lens]
        sample = list(range(25)) + rng.sample(range(25, len(lens)), 120) + [len(lens) - 1]
        for m, k, d in schemes:
            with nwb.tuned(batch_bp=0):
                bt = nwb.Batch(tops, sides, m, k, d, nwb.WANT_ARROWS_HOST)
            assert bt.kernel_name() == "nwb_batch_bx_kernel"
            bt.run()
            bt.fetch()
            _batch_check(oracle, nwb, bt, tops, sides, m, k, d, sample)
            scores = [bt.opt_score(i) for i in range(len(lens))]
            branches = [bt.branch_count(i) for i in range(len(lens))]
            bt.close()
            with nwb.tuned(batch_bx=0, batch_bp=0):
                b0 = nwb.Batch(tops, sides, m, k, d, 0)
            assert b0.kernel_name() == "nwb_batch_pk_kernel"
            b0.run()
            b0.fetch()
            assert scores == [b0.opt_score(i) for i in range(len(lens))]
            assert branches == [b0.branch_count(i) for i in range(len(lens))]
            b0.close()
    # batches of one pair (a warp sweeps it alone), uniform and not
    for t1, s1, name in ((tops[0], sides[0], "nwb_batch_cx_kernel"), (tops[6], sides[6], "nwb_batch_bx_kernel")):
        bt = nwb.Batch([t1], [s1], 2, 1, 2, nwb.WANT_ARROWS_HOST | nwb.WANT_COUNT)
        assert bt.kernel_name() == name
        bt.run()
        bt.fetch()
        _batch_check(oracle, nwb, bt, [t1], [s1], 2, 1, 2, [0])
        assert bt.count(0) == oracle.fill(t1, s1, 2, 1, 2).count
        bt.close()
    # 2d + m = 8 does not fit a nibble: the one-pair-per-warp kernel takes the batch
    bt = nwb.Batch(tops[:4], sides[:4], 2, 1, 3, 0)
    assert bt.kernel_name() == "nwb_batch_pk_kernel"
    bt.close()


def test_batch_uniform_pairs_back_to_back(oracle, nwb, monkeypatch):
    """csrc/nwb_batch_bx.cuh nwb_batch_cx_kernel: uniform shapes, a warp sweeps its pairs of pairs back to back
    (long chains: more pairs than 2 x 12 x 148 warps), odd batches, narrow and short tables; the same batches
    through the drained two-pairs-per-warp kernel (nwb_tune batch_cx = 0) must agree."""
    rng = random.Random(31)
    for a, b, n, alpha, (m, k, d) in ((256, 256, 12001, b"ACGT", (1, 1, 1)), (100, 96, 9000, bytes(range(1, 256)), (2, 1, 2)),
                                      (7, 64, 20001, b"AC", (0, 0, 0)), (255, 160, 5000, b"ACGT", (1, 1, 3))):
        tops = [bytes(rng.choice(alpha) for _ in range(a)) for _ in range(n)]
        sides = [bytes(rng.choice(alpha) for _ in range(b)) for _ in range(n)]
        with nwb.tuned(batch_bp=0):
            bt = nwb.Batch(tops, sides, m, k, d, nwb.WANT_ARROWS_HOST)
        assert bt.kernel_name() == "nwb_batch_cx_kernel"
        bt.run()
        bt.fetch()
        sample = [0, 1, 2, n - 3, n - 2, n - 1] + rng.sample(range(n), 60)
        _batch_check(oracle, nwb, bt, tops, sides, m, k, d, sample)
        scores = [bt.opt_score(i) for i in range(n)]
        branches = [bt.branch_count(i) for i in range(n)]
        tabs = [bt.arrow_rows(i).copy() for i in sample]
        bt.close()
        with nwb.tuned(batch_cx=0, batch_bp=0):
            b0 = nwb.Batch(tops, sides, m, k, d, nwb.WANT_ARROWS_HOST)
        assert b0.kernel_name() == "nwb_batch_bx_kernel"
        b0.run()
        b0.fetch()
        assert scores == [b0.opt_score(i) for i in range(n)]
        assert branches == [b0.branch_count(i) for i in range(n)]
        nb = (a + 1) // 2
        for i, tab in zip(sample, tabs):
            assert np.array_equal(tab[:, :nb] & 0x77, b0.arrow_rows(i)[:, :nb] & 0x77)
        b0.close()


def test_batch_bit_parallel(oracle, nwb):
    """csrc/nwb_batch_bp.cuh nwb_batch_bp_kernel: one thread per pair, rows as bit-vectors (2d + m <= 3, top strings of
    at most 256 characters).  Config 4's goldens, ragged shapes, empty strings, unaligned offsets, more groups of 32
    pairs than warps, side letters the top string does not have, top strings with more than five letters (worked off
    by nwb_batch_pk_kernel from the kernel's list), every instantiated (2d + m, 2d - k); the same batches through
    the packed-difference kernels (nwb_tune batch_bp = 0) must agree on every pair."""
    rng = random.Random(37)
    n = 9000
    idx = list(range(n - 1)) + [999999]
    tops, sides = zip(*(oracle.generate_pair(0x5EED4000 + 2 * p, 256, 256) for p in idx))
    with nwb.tuned(batch_bp=1):
        bt = nwb.Batch(list(tops), list(sides), 1, 1, 1, nwb.WANT_ARROWS_HOST | nwb.WANT_COUNT)
    assert bt.kernel_name() == "nwb_batch_bp_kernel"
    bt.run()
    bt.fetch()
    assert (bt.opt_score(0), bt.branch_count(0), bt.count(0)) == (19, 23713, 387701138034524160)
    assert (bt.opt_score(1), bt.branch_count(1), bt.count(1)) == (29, 22912, 108460706365440)
    assert (bt.opt_score(n - 1), bt.branch_count(n - 1), bt.count(n - 1)) == (19, 22090, 4971798065203200)
    _batch_check(oracle, nwb, bt, tops, sides, 1, 1, 1, rng.sample(range(n), 50))
    dg = bt.digest(0)
    bt.close()
    with nwb.tuned(batch_bp=0):
        b0 = nwb.Batch(list(tops), list(sides), 1, 1, 1, nwb.WANT_COUNT)
    assert b0.kernel_name() == "nwb_batch_cx_kernel"
    b0.run()
    b0.fetch()
    assert b0.digest(0) == dg          # every pair: arrow tables, scores, branch counters, counts
    b0.close()
    lens = [(256, 256), (1, 1), (255, 257), (3, 40), (17, 130), (0, 5), (200, 90), (64, 64), (256, 1), (33, 33),
            (256, 31), (100, 300), (5, 0), (8, 32), (9, 33), (249, 63), (250, 64), (7, 65), (1, 200), (256, 2),
            (0, 0), (31, 31), (130, 95), (96, 128), (256, 1000)] + \
           [(rng.randint(1, 256), rng.randint(1, 400)) for _ in range(6000)]
    alphas = [b"ACGT"] * 6 + [b"ACGTN", b"ACGTNR", b"AC", bytes(range(1, 256)), b"ARNDCQEGHILKMFPSTWYV"]
    tops = [bytes(rng.choice(rng.choice(alphas)) for _ in range(a)) for a, _ in lens]
    sides = [bytes(rng.choice(b"ACGTNX") for _ in range(b)) for _, b in lens]
    sample = list(range(25)) + rng.sample(range(25, len(lens)), 150) + [len(lens) - 1]
    sample += [i for i in range(25, 400) if len(set(tops[i])) > 5][:20] + [i for i in range(25, 400) if len(set(tops[i])) == 5][:10]
    for m, k, d in ((1, 1, 1), (0, 0, 1), (1, -1, 0), (3, 0, 0), (1, 2, 1), (0, 1, 1), (2, -1, 0)):
        with nwb.tuned(batch_bp=1):
            bt = nwb.Batch(tops, sides, m, k, d, nwb.WANT_ARROWS_HOST)
        assert bt.kernel_name() == "nwb_batch_bp_kernel", (m, k, d)
        bt.run()
        bt.fetch()
        _batch_check(oracle, nwb, bt, tops, sides, m, k, d, sample if (m, k, d) == (1, 1, 1) else sample[:60])
        dg = bt.digest(0)
        bt.close()
        with nwb.tuned(batch_bp=0):
            b0 = nwb.Batch(tops, sides, m, k, d, 0)
        assert b0.kernel_name() != "nwb_batch_bp_kernel"
        b0.run()
        b0.fetch()
        assert b0.digest(0)[:3] == dg[:3], (m, k, d)
        b0.close()
    # refill from host buffers in chunks (every chunk has its own left-over list)
    n = 40000
    tcat = nwb.generate(0x5EED4000, 256, nwb.DNA, count=n, seed_stride=2)
    scat = nwb.generate(0x5EED4001, 256, nwb.DNA, count=n, seed_stride=2)
    tmix = bytearray(tcat)
    for p in range(0, n, 997):      # some pairs with six letters in the top string
        tmix[p * 256:p * 256 + 6] = b"ACGTNR"
    off = np.arange(n + 1, dtype=np.int64) * 256
    res = {}
    for knob in (1, 0):
        with nwb.tuned(batch_bp=knob):
            bt = nwb.Batch.from_arrays(tcat, off, scat, off, 1, 1, 1, 0)
        bt.refill(bytes(tmix), scat)
        bt.fetch()
        res[knob] = (bt.digest(0)[:3], bt.opt_score(997), bt.branch_count(n - 1))
        bt.close()
    assert res[1] == res[0]
    # narrow tables: row vectors of 64 and 128 bits (uniform groups and ragged ones)
    for amax, n in ((64, 3000), (128, 3000), (100, 2500), (40, 2000)):
        lens = [(amax, 70)] * 1500 + [(rng.randint(1, amax), rng.randint(1, 200)) for _ in range(n - 1500)]
        tops = [bytes(rng.choice(b"ACGT") for _ in range(a)) for a, _ in lens]
        sides = [bytes(rng.choice(b"ACGT") for _ in range(b)) for _, b in lens]
        with nwb.tuned(batch_bp=1):
            bt = nwb.Batch(tops, sides, 1, 1, 1, nwb.WANT_ARROWS_HOST | nwb.WANT_COUNT)
        assert bt.kernel_name() == "nwb_batch_bp_kernel"
        bt.run()
        bt.fetch()
        _batch_check(oracle, nwb, bt, tops, sides, 1, 1, 1, [0, 1, 1499, 1500, n - 1] + rng.sample(range(n), 40))
        dg = bt.digest(0)
        bt.close()
        with nwb.tuned(batch_bp=0):
            b0 = nwb.Batch(tops, sides, 1, 1, 1, nwb.WANT_COUNT)
        b0.run()
        b0.fetch()
        assert b0.digest(0) == dg, amax
        b0.close()
    # not for this kernel: 2d + m > 3, strings wider than a strip
    with nwb.tuned(batch_bp=1):
        bt = nwb.Batch(tops[:4], sides[:4], 2, 1, 2, 0)
        assert bt.kernel_name() != "nwb_batch_bp_kernel"
        bt.close()
        bt = nwb.Batch([b"A" * 300], [b"C" * 10], 1, 1, 1, 0)
        assert bt.kernel_name() != "nwb_batch_bp_kernel"
        bt.close()



def test_batch_lane_count(oracle, nwb):
    """csrc/nwb_batch_lcount.cuh nwb_batch_lcount_kernel: the count behind -s for a batch, one thread per pair sweeping
    backwards over the cells on optimal paths; pairs whose band does not fit the 40-column window are counted by the
    dense kernel from the list.  Config 4's golden counts, ragged shapes, pairs wider than a strip, empty strings,
    schemes where nearly every pair gives up (0/0/0: every cell ties), protein 2/1/2; the dense pass alone
    (nwb_tune batch_lcount = 0) must agree on every pair, and refill in chunks too."""
    rng = random.Random(47)
    n = 6000
    idx = list(range(n - 1)) + [999999]
    tops, sides = zip(*(oracle.generate_pair(0x5EED4000 + 2 * p, 256, 256) for p in idx))
    res = {}
    for knob in (1, 0):
        with nwb.tuned(batch_lcount=knob):
            bt = nwb.Batch(list(tops), list(sides), 1, 1, 1, nwb.WANT_COUNT)
            bt.run()
            bt.fetch()
            res[knob] = (bt.digest(0), [bt.count(i) for i in (0, 1, n - 1)], bt.launches())
            if knob:
                for i in rng.sample(range(n), 40):
                    assert bt.count(i) == oracle.fill(tops[i], sides[i], 1, 1, 1).count, i
            bt.close()
    assert res[1][1] == [387701138034524160, 108460706365440, 4971798065203200]
    assert res[1][:2] == res[0][:2]
    lens = [(256, 256), (1, 1), (255, 257), (300, 40), (17, 130), (0, 5), (700, 90), (64, 64), (256, 1), (33, 33),
            (257, 31), (100, 300), (5, 0), (1500, 700), (40, 41), (41, 40)] + \
           [(rng.randint(1, 600), rng.randint(1, 500)) for _ in range(300)] + \
           [(a, max(1, a + rng.randint(-8, 8))) for a in (30, 60, 100, 200, 256, 300, 500) for _ in range(20)]
    for alpha, schemes in ((b"ACGT", ((1, 1, 1), (0, 0, 0))), (b"ARNDCQEGHILKMFPSTWYV", ((2, 1, 2), (5, 4, 3)))):
        tops = [bytes(rng.choice(alpha) for _ in range(a)) for a, _ in lens]
        sides = [bytes(rng.choice(alpha) for _ in range(b)) for _, b in lens]
        for m, k, d in schemes:
            got = {}
            for knob in (1, 0):
                with nwb.tuned(batch_lcount=knob):
                    bt = nwb.Batch(tops, sides, m, k, d, nwb.WANT_COUNT)
                    bt.run()
                    bt.fetch()
                    got[knob] = [bt.count(i) for i in range(len(lens))]
                    bt.close()
            assert got[1] == got[0], (m, k, d)
            for i in list(range(16)) + rng.sample(range(16, len(lens)), 25):
                assert got[1][i] == oracle.fill(tops[i], sides[i], m, k, d).count, (m, k, d, i, lens[i])
    # from host buffers in chunks: every chunk has its own left-over list
    n = 40000
    tcat = nwb.generate(0x5EED4000, 256, nwb.DNA, count=n, seed_stride=2)
    scat = nwb.generate(0x5EED4001, 256, nwb.DNA, count=n, seed_stride=2)
    off = np.arange(n + 1, dtype=np.int64) * 256
    res = {}
    for knob in (1, 0):
        with nwb.tuned(batch_lcount=knob):
            bt = nwb.Batch.from_arrays(tcat, off, scat, off, 1, 1, 1, nwb.WANT_COUNT)
            bt.refill(tcat, scat)
            bt.fetch()
            res[knob] = bt.digest(0)
            bt.close()
    assert res[1] == res[0]



def test_batch_count(oracle, nwb):
    """NWB_WANT_COUNT on the batch path (csrc/nwb_batch_count.cuh): SURVEY 8c's config 4 counts, ragged shapes,
    pairs wider than one strip, empty strings, wrap-around mod 2^64 (0/0/0)."""
    n = 3000
    idx = list(range(n - 1)) + [999999]
    tops, sides = zip(*(oracle.generate_pair(0x5EED4000 + 2 * p, 256, 256) for p in idx))
    bt = nwb.Batch(list(tops), list(sides), 1, 1, 1, nwb.WANT_COUNT)
    assert bt.kernel_name() == "nwb_batch_bp_kernel"
    bt.run()
    bt.fetch()
    assert [bt.count(i) for i in (0, 1, n - 1)] == [387701138034524160, 108460706365440, 4971798065203200]
    rng = random.Random(43)
    for i in rng.sample(range(n), 40):
        assert bt.count(i) == oracle.fill(tops[i], sides[i], 1, 1, 1).count, i
    bt.close()
    lens = [(256, 256), (1, 1), (255, 257), (300, 40), (17, 130), (0, 5), (700, 90), (64, 64), (256, 1), (33, 33),
            (257, 31), (100, 300), (5, 0), (1500, 700)] + [(rng.randint(1, 600), rng.randint(1, 500)) for _ in range(400)]
    tops = [bytes(rng.choice(b"ACGT") for _ in range(a)) for a, _ in lens]
    sides = [bytes(rng.choice(b"ACGT") for _ in range(b)) for _, b in lens]
    for m, k, d in ((1, 1, 1), (0, 0, 0)):
        bt = nwb.Batch(tops, sides, m, k, d, nwb.WANT_COUNT)
        bt.run()
        bt.fetch()
        for i in list(range(14)) + rng.sample(range(14, len(lens)), 60):
            o = oracle.fill(tops[i], sides[i], m, k, d)
            assert (bt.count(i), bt.opt_score(i), bt.branch_count(i)) == (o.count, o.final_score, o.branch_count), i
        bt.close()


def test_batch_count_chained_runs(oracle, nwb):
    """Uniform one-strip batches: nwb_batch_count_chain_kernel sweeps a warp's run of pairs back to back.  Runs of 5-6
    pairs per warp (12,000 pairs on 2,368 warps), A < 256, B a multiple of 4 but not of 32; every count against the
    one-pair-at-a-time kernel (nwb_tune bcnt_chain = 0) and a sample against the oracle, also mod 2^64 (0/0/0)."""
    rng = random.Random(47)
    n = 12000
    tops = [bytes(rng.choice(b"ACGT") for _ in range(100)) for _ in range(n)]
    sides = [bytes(rng.choice(b"ACGT") for _ in range(68)) for _ in range(n)]
    for m, k, d in ((1, 1, 1), (0, 0, 0)):
        got = {}
        for chain in ("1", "0"):
            with nwb.tuned(bcnt_chain=int(chain)):
                bt = nwb.Batch(tops, sides, m, k, d, nwb.WANT_COUNT)
                bt.run()
                bt.fetch()
                got[chain] = [bt.count(i) for i in range(n)]
                bt.close()
        assert got["1"] == got["0"]
        for i in [0, 1, 5, 6, n - 1] + rng.sample(range(n), 25):
            assert got["1"][i] == oracle.fill(tops[i], sides[i], m, k, d).count, i


def test_count_prefix_property(oracle, nwb):
    """Intermediate counts of the packed count kernel through the public ABI: the count of cell
    (i, j) equals the final count of the sub-problem (top[:i], side[:j]).  The final count of the
    big configs is 0 mod 2^64 (SURVEY hard part 6), so the check samples cells whose counts are not."""
    t, s = oracle.generate_pair(0x5EED0D00, 900, 700)
    o = oracle.fill(t, s, 1, 1, 1, want_counts=True)
    rng = random.Random(17)
    cells = [(900, 700), (256, 256), (257, 300), (512, 64), (513, 699)] + \
            [(rng.randint(1, 900), rng.randint(1, 700)) for _ in range(20)]
    nonzero = 0
    for i, j in cells:
        tab = nwb.fill(t[:i], s[:j], 1, 1, 1, nwb.WANT_COUNT)
        assert tab.kernel_kind == nwb.KIND_PK
        assert tab.count == int(o.counts[j, i]), (i, j)
        nonzero += int(o.counts[j, i]) != 0
    assert nonzero >= 10


def test_extreme_aspect_ratios_and_multi_pass(oracle, nwb):
    """Wide-and-short / tall-and-narrow tables, and a table with more strips than persistent warps
    (160,000 columns = 625 strips > 592 warps: the cyclic second pass of the strip assignment)."""
    for seed, a, b, mkd in [(0x5EED0E00, 20000, 7, (1, 1, 1)), (0x5EED0E02, 5, 20000, (1, 1, 1)),
                            (0x5EED0E04, 70000, 300, (2, 1, 2)), (0x5EED0E06, 160000, 200, (1, 1, 1)),
                            (0x5EED0E08, 300, 70000, (1, 1, 1))]:
        t, s = oracle.generate_pair(seed, a, b)
        o = oracle.fill(t, s, *mkd)
        for flags in (nwb.WANT_ARROWS_HOST, nwb.WANT_ARROWS_HOST | nwb.WANT_COUNT,
                      nwb.WANT_ARROWS_HOST | nwb.WANT_COUNT | nwb.FORCE_GENERAL):
            tab = nwb.fill(t, s, *mkd, flags)
            assert (tab.opt_score, tab.branch_count) == (o.final_score, o.branch_count), (a, b, flags)
            if flags & nwb.WANT_COUNT:
                assert tab.count == o.count, (a, b, flags)
            check_arrows(oracle, nwb, tab, t, s, *mkd)


@pytest.fixture
def force_hx(nwb):
    """nwb_tune pk_hx = 1: the sweeping + flush warp kernel (csrc/nwb_fill_hx.cuh) at every size the scheme allows,
    not only for tall tables."""
    nwb.tune("pk_hx", 1)
    yield "hx"
    nwb.tune_reset()


def test_hx_variant_edge_shapes(oracle, nwb, force_hx):
    rng = random.Random(23)
    sizes = [1, 2, 31, 33, 64, 65, 127, 129, 255, 256, 257, 511, 513, 769, 1025, 2049]
    for _ in range(40):
        a, b = rng.choice(sizes), rng.choice(sizes)
        alpha = rng.choice([b"ACGT", b"AB", bytes(range(1, 256)), b"ARNDCQEGHILKMFPSTWYV"])
        t = bytes(rng.choice(alpha) for _ in range(a))
        s = bytes(rng.choice(alpha) for _ in range(b))
        m, k, d = rng.choice([(1, 1, 1), (2, 1, 2), (0, 0, 0), (1, 2, 3), (3, -1, 0), (1, 1, 0), (1, 1, 3), (3, 1, 2)])
        tab = nwb.fill(t, s, m, k, d, nwb.WANT_ARROWS_HOST)
        o = check_arrows(oracle, nwb, tab, t, s, m, k, d)
        assert tab.kernel_kind == (nwb.KIND_PK if pk_supported(m, k, d) else nwb.KIND_I32)
        assert (tab.opt_score, tab.branch_count) == (o.final_score, o.branch_count), (a, b, m, k, d)
        # -s, default path (sparse backward sweep, dense sweep behind it) and the dense sweep trailing the hx
        # fill on a second stream (count_mode 3: rows published by the flush warps)
        tabc = nwb.fill(t, s, m, k, d, nwb.WANT_COUNT)
        assert (tabc.opt_score, tabc.branch_count, tabc.count) == (o.final_score, o.branch_count, o.count), (a, b, m, k, d)
        nwb.tune("count_mode", 3)
        tabc = nwb.fill(t, s, m, k, d, nwb.WANT_COUNT)
        nwb.tune("count_mode", 0)
        assert (tabc.opt_score, tabc.branch_count, tabc.count) == (o.final_score, o.branch_count, o.count), (a, b, m, k, d)


def test_hx_variant_big_and_multi_pass(oracle, nwb, force_hx):
    """Whole tables at 10k x 10k (config 2) and 3k x 3k protein, extreme aspect ratios, and more strips
    than sweeping warps (160,000 columns = 625 strips > 444: several strips per warp, ring across strips)."""
    t, s = oracle.generate_pair(0x5EED0002, 10000, 10000)
    plan = nwb.Plan(10000, 10000, 0)
    plan.upload(t, s)
    plan.run(1, 1, 1)
    assert plan.kernel_name() == "nwb_fill_hx_kernel"
    plan.close()
    tab = nwb.fill(t, s, 1, 1, 1, nwb.WANT_ARROWS_HOST)
    assert (tab.opt_score, tab.branch_count) == (1056, 34377799)
    check_arrows(oracle, nwb, tab, t, s, 1, 1, 1)
    for seed, a, b, mkd, alpha in [(0x5EED0005, 3000, 3000, (2, 1, 2), "protein"), (0x5EED0E00, 20000, 7, (1, 1, 1), "dna"),
                                   (0x5EED0E02, 5, 20000, (1, 1, 1), "dna"), (0x5EED0E06, 160000, 200, (1, 1, 1), "dna"),
                                   (0x5EED0E08, 300, 70001, (1, 1, 1), "dna")]:
        t, s = oracle.generate_pair(seed, a, b, oracle.PROTEIN if alpha == "protein" else oracle.DNA)
        tab = nwb.fill(t, s, *mkd, nwb.WANT_ARROWS_HOST)
        o = check_arrows(oracle, nwb, tab, t, s, *mkd)
        assert (tab.opt_score, tab.branch_count) == (o.final_score, o.branch_count), (a, b)


def test_plain_packed_kernel_without_hx(oracle, nwb):
    """nwb_tune pk_hx = 0 keeps the one-warp-per-strip packed kernel (the path for 2d + m > 7) covered at 10k x 10k."""
    with nwb.tuned(pk_hx=0):
        t, s = oracle.generate_pair(0x5EED0002, 10000, 10000)
        tab = nwb.fill(t, s, 1, 1, 1, nwb.WANT_ARROWS_HOST)
        assert (tab.opt_score, tab.branch_count) == (1056, 34377799)
        check_arrows(oracle, nwb, tab, t, s, 1, 1, 1)
    # a scheme whose differences do not fit a nibble (2d + m = 11) at a size where hx would otherwise run
    t, s = oracle.generate_pair(0x5EED0F00, 3000, 5000)
    tab = nwb.fill(t, s, 5, 4, 3, nwb.WANT_ARROWS_HOST)
    o = check_arrows(oracle, nwb, tab, t, s, 5, 4, 3)
    assert (tab.opt_score, tab.branch_count) == (o.final_score, o.branch_count)


def test_fused_count_kernel_still_matches(oracle, nwb):
    """nwb_tune count_mode = 1: the count fused into nwb_fill_pk_kernel: final counts of sub-problems whose count
    is not 0 mod 2^64, and config 2."""
    with nwb.tuned(count_mode=1):
        t, s = oracle.generate_pair(0x5EED0D00, 900, 700)
        o = oracle.fill(t, s, 1, 1, 1, want_counts=True)
        for i, j in [(900, 700), (256, 256), (257, 300), (513, 699), (100, 650)]:
            tab = nwb.fill(t[:i], s[:j], 1, 1, 1, nwb.WANT_COUNT)
            assert tab.summary().count_path == nwb.COUNT_FUSED
            assert tab.count == int(o.counts[j, i]), (i, j)
        t, s = oracle.generate_pair(0x5EED0002, 10000, 10000)
        tab = nwb.fill(t, s, 1, 1, 1, nwb.WANT_COUNT)
        assert (tab.opt_score, tab.branch_count, tab.count) == (1056, 34377799, 0)


def test_dense_count_sweep_after_and_trailing_the_fill(oracle, nwb):
    """nwb_tune count_mode = 2 / 3: the dense forward count sweep (csrc/nwb_count.cuh) launched after the hx fill,
    or trailing it on a second stream: same counts as the oracle, on sub-problems whose count is not 0 mod 2^64."""
    t, s = oracle.generate_pair(0x5EED0D10, 900, 4500)
    o = oracle.fill(t, s, 1, 1, 1, want_counts=True)
    for mode in (2, 3):
        with nwb.tuned(count_mode=mode):
            for i, j in [(900, 4500), (256, 4100), (513, 4499)]:
                tab = nwb.fill(t[:i], s[:j], 1, 1, 1, nwb.WANT_COUNT)
                assert tab.summary().count_path == nwb.COUNT_DENSE
                assert tab.count == int(o.counts[j, i]), (mode, i, j)


# ---- the sparse backward count (csrc/nwb_count_sparse.cuh), the default behind -s ---------------------------
def test_sparse_count_is_the_default_and_matches(oracle, nwb):
    rng = random.Random(61)
    seen_paths = set()
    for it in range(40):
        a, b = rng.randint(1, 3000), rng.randint(1, 3000)
        alpha = rng.choice([oracle.DNA, oracle.PROTEIN, "AC"])
        m, k, d = rng.choice([(1, 1, 1), (2, 1, 2), (0, 0, 0), (1, 2, 3), (5, 4, 3), (1, 1, 0)])
        t, s = oracle.generate_pair(0x5EEDA000 + 2 * it, a, b, alpha)
        o = oracle.fill(t, s, m, k, d)
        for ff in (0, nwb.FORCE_GENERAL):
            tab = nwb.fill(t, s, m, k, d, nwb.WANT_COUNT | ff)
            sm = tab.summary()
            assert sm.count_path in (nwb.COUNT_SPARSE, nwb.COUNT_SPARSE_BAILED), (it, sm.count_path)
            seen_paths.add(sm.count_path)
            assert (tab.count, tab.opt_score, tab.branch_count) == (o.count, o.final_score, o.branch_count), (it, a, b, m, k, d)
    # (0,0,0) tables wider than the 256-column window make the sparse sweep give up: the dense sweep behind it ran
    assert seen_paths == {nwb.COUNT_SPARSE, nwb.COUNT_SPARSE_BAILED}


def _mutated(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_big", os.path.join(HERE, "golden", "make_golden_big.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for nm, seed, n, subs, indels, mkd in mod.MUTATED:
        if nm == name:
            return mod.mutated_pair(seed, n, subs, indels), mkd
    raise KeyError(name)


@pytest.mark.parametrize("name", ["mutated_dna_30k", "mutated_dna_100k"])
def test_sparse_count_nonzero_at_full_size(oracle, nwb, name):
    """A string against a mutated copy of itself: the number of optimal alignments is NOT 0 mod 2^64, so the
    sparse sweep must carry live values through every one of the 30,000 / 100,000 rows; the same count from the
    dense sweep (every intermediate count of the last row and column checked through their digests)."""
    g = [c for c in golden("golden_big.json") if c["name"] == name][0]
    (t, s), (m, k, d) = _mutated(name)
    assert (len(t), len(s)) == (g["top_len"], g["side_len"])
    assert g["count_u64"] != 0
    plan = nwb.Plan(len(t), len(s), nwb.WANT_COUNT)
    plan.upload(t, s)
    plan.run(m, k, d)
    sm = plan.summary()
    assert sm.count_path == nwb.COUNT_SPARSE and sm.count_rows == len(s)
    assert (sm.opt_score, sm.branch_count, sm.count) == (g["final_score"], g["branch_count"], g["count_u64"])
    assert plan.arrow_digest() == int(g["arrow_digest"], 16)
    plan.close()
    plan = nwb.Plan(len(t), len(s), nwb.WANT_COUNT_DIGEST)
    plan.upload(t, s)
    plan.run(m, k, d)
    sm = plan.summary()
    assert sm.count_path == nwb.COUNT_DENSE
    assert sm.count == g["count_u64"]
    assert (sm.lastrow_count_digest, sm.lastcol_count_digest) == (int(g["lastrow_count_digest"], 16), int(g["lastcol_count_digest"], 16))
    plan.close()


# ---- full-size parity through on-device digests (csrc/nwb_digest.cuh) ---------------------------------------
@pytest.mark.parametrize("name,alpha,mkd", [("config2_dna_10k", "dna", (1, 1, 1)), ("config5_protein_30k", "protein", (2, 1, 2)),
                                            ("config3_dna_100k", "dna", (1, 1, 1))])
@pytest.mark.parametrize("kernel", ["hx", "pk", "i32"])
def test_full_size_digests(oracle, nwb, name, alpha, mkd, kernel):
    """EVERY arrow set of configs 2, 5 and 3 at full size (digest of the whole nibble table, computed on the
    device, against the oracle's), the summary, the default (sparse) count, and every count of the dense sweep's
    last row and last column (non-zero intermediate values: the final count of these configs is 0 mod 2^64)."""
    g = [c for c in golden("golden_big.json") if c["name"] == name][0]
    a, b = g["top_len"], g["side_len"]
    if kernel != "hx" and a > 30000:
        pytest.skip("the slower kernels are covered at 10k and 30k")
    t, s = oracle.generate_pair(g["seed"], a, b, oracle.DNA if alpha == "dna" else oracle.PROTEIN)
    ff = nwb.FORCE_GENERAL if kernel == "i32" else 0
    want_kernel = {"hx": "nwb_fill_hx_kernel", "pk": "nwb_fill_pk_kernel", "i32": "nwb_fill_i32_kernel"}[kernel]
    with nwb.tuned(pk_hx=0 if kernel == "pk" else -1):
        plan = nwb.Plan(a, b, nwb.WANT_COUNT | ff)
        plan.upload(t, s)
        plan.run(*mkd)
        sm = plan.summary()
        assert plan.kernel_name() == want_kernel
        assert (sm.opt_score, sm.branch_count, sm.count) == (g["final_score"], g["branch_count"], g["count_u64"])
        assert sm.count_path == nwb.COUNT_SPARSE
        assert plan.arrow_digest() == int(g["arrow_digest"], 16)
        plan.close()
        if kernel == "hx":
            plan = nwb.Plan(a, b, nwb.WANT_COUNT_DIGEST)
            plan.upload(t, s)
            plan.run(*mkd)
            sm = plan.summary()
            assert sm.count_path == nwb.COUNT_DENSE and sm.count == g["count_u64"]
            assert sm.lastrow_count_digest == int(g["lastrow_count_digest"], 16)
            assert sm.lastcol_count_digest == int(g["lastcol_count_digest"], 16)
            plan.close()


def test_table_digest_through_nwb_fill(oracle, nwb):
    t, s = oracle.generate_pair(0x5EED0C30, 3001, 1777)
    o = oracle.fill(t, s, 1, 1, 1)
    for ff in (0, nwb.FORCE_GENERAL):
        tab = nwb.fill(t, s, 1, 1, 1, nwb.WANT_DIGEST | ff)
        assert tab.arrow_digest() == o.arrow_digest
    tab = nwb.fill(t, s, 1, 1, 1, 0)
    with pytest.raises(nwb.NwbError):
        tab.arrow_digest()


def test_config4_full_shard_digest(oracle, nwb):
    """ALL 125,000 pairs of one GPU's shard of config 4 (every arrow set, score, branch count and alignment count,
    through the batch digests computed on the device) against the oracle-generated per-shard digests."""
    g = [c for c in golden("golden_big.json") if c["name"] == "config4_batch_1M"][0]
    per = g["pairs_per_shard"]
    for shard in (0, 5):
        first = shard * per
        tcat, scat = bytearray(), bytearray()
        for p in range(first, first + per):
            tt, ss = oracle.generate_pair(0x5EED4000 + 2 * p, 256, 256)
            tcat += tt
            scat += ss
        off = np.arange(per + 1, dtype=np.int64) * 256
        bt = nwb.Batch.from_arrays(bytes(tcat), off, bytes(scat), off, 1, 1, 1, nwb.WANT_COUNT)
        bt.run()
        dg = bt.digest(first)
        bt.close()
        e = g["shard_digests"][shard]
        assert dg == tuple(int(e[k], 16) for k in ("arrow", "score", "branch", "count")), shard


def test_watchdog_turns_a_lost_boundary_stream_into_an_error(oracle, nwb):
    """nwb_tune inject_fault = 1 makes every strip keep its boundary stream to itself, so the next strip waits for
    ever -- the device watchdog must end the wait, and the fill must FAIL (NWB_ERR_CUDA), for every kernel family;
    the library stays usable afterwards."""
    t, s = oracle.generate_pair(0x5EED0F50, 3000, 5000)
    o = oracle.fill(t, s, 1, 1, 1)
    for pk_hx, ff, flags in ((-1, 0, 0), (0, 0, 0), (-1, nwb.FORCE_GENERAL, 0), (-1, 0, nwb.WANT_COUNT_DIGEST)):
        nwb.tune("inject_fault", 1)
        nwb.tune("watchdog_ms", 250)
        nwb.tune("pk_hx", pk_hx)
        try:
            with pytest.raises(nwb.NwbError) as e:
                nwb.fill(t, s, 1, 1, 1, ff | flags)
            assert e.value.code == -3, e.value
            assert "watchdog" in str(e.value)
        finally:
            nwb.tune_reset()
        tab = nwb.fill(t, s, 1, 1, 1, ff | nwb.WANT_COUNT)
        assert (tab.opt_score, tab.branch_count, tab.count) == (o.final_score, o.branch_count, o.count)


def test_batch_refill_chunked_pipeline(oracle, nwb):
    """nwb_batch_refill(): new strings for the same shapes from host buffers, uploaded chunk by chunk while the
    previous chunk's kernels run (several chunks: > 32768 pairs).  Results must equal a fresh batch of the new
    strings: uniform shapes (back-to-back kernel), ragged shapes (two pairs per warp, odd count), with the count."""
    rng = random.Random(91)
    for shapes in ([(64, 64)] * 40000, [(rng.randint(1, 80), rng.randint(1, 70)) for _ in range(40001)]):
        n = len(shapes)
        def strings(seed):
            r = random.Random(seed)
            return ([bytes(r.choice(b"ACGT") for _ in range(a)) for a, _ in shapes],
                    [bytes(r.choice(b"ACGT") for _ in range(b)) for _, b in shapes])
        t0, s0 = strings(1)
        t1, s1 = strings(2)
        bt = nwb.Batch(t0, s0, 1, 1, 1, nwb.WANT_COUNT)
        bt.run()
        bt.fetch()
        bt.refill(b"".join(t1), b"".join(s1))
        bt.fetch()
        fresh = nwb.Batch(t1, s1, 1, 1, 1, nwb.WANT_COUNT)
        fresh.run()
        fresh.fetch()
        assert bt.digest(0) == fresh.digest(0)
        for i in [0, 1, n // 2, n - 2, n - 1] + rng.sample(range(n), 30):
            assert (bt.opt_score(i), bt.branch_count(i), bt.count(i)) == (fresh.opt_score(i), fresh.branch_count(i), fresh.count(i)), i
            o = oracle.fill(t1[i], s1[i], 1, 1, 1)
            assert (bt.opt_score(i), bt.branch_count(i), bt.count(i)) == (o.final_score, o.branch_count, o.count), i
        bt.close()
        fresh.close()
