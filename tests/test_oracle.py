"""CPU tests: pin the oracle (oracle/nw_oracle.c) against the reference's known
answers (README:117-173), the committed golden vectors (tests/golden/golden.json,
generated from the compiled reference) and -- when oracle/_ref is built -- the
reference itself."""
import json
import os
import random
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))

# Full README table (score:diag,left,up bits), rows j=0..7 -- README:157-173 as
# transcribed in SURVEY.md 8c from the compiled reference.
README_TABLE = """
  0:000  -1:010  -2:010  -3:010  -4:010  -5:010  -6:010  -7:010
 -1:001   1:100   0:010  -1:010  -2:010  -3:110  -4:010  -5:010
 -2:001   0:001   0:100   1:100   0:010  -1:010  -2:010  -3:010
 -3:001  -1:001  -1:101   0:001   2:100   1:010   0:010  -1:010
 -4:001  -2:001  -2:101  -1:001   1:101   1:100   0:110  -1:110
 -5:001  -3:001  -3:101  -1:100   0:001   0:101   0:100  -1:110
 -6:001  -4:001  -2:100  -2:001  -1:001  -1:101   1:100   0:010
 -7:001  -5:001  -3:001  -1:100  -2:011  -2:101   0:001   0:100
"""


def golden_cases():
    with open(os.path.join(HERE, "golden", "golden.json")) as f:
        return json.load(f)


def case_strings(oracle, c):
    if "seed" in c:
        alpha = oracle.DNA if c["alphabet"] == "dna" else oracle.PROTEIN
        return oracle.generate_pair(c["seed"], c["top_len"], c["side_len"], alpha)
    return c["top"].encode(), c["side"].encode()


def test_generator_self_checks(oracle):
    # SURVEY.md 8c "Generator self-checks (first 16 chars)"
    assert oracle.generate(0x5EED0002, 16) == b"CTTTCCTTAGCAGTTA"
    assert oracle.generate(0x5EED0003, 16) == b"AAAGCCCCCACACACA"
    assert oracle.generate(0x5EED0030, 16) == b"CATACCGTAGTTCCAG"
    assert oracle.generate(0x5EED0031, 16) == b"TTCACAGCTTCCCAAA"
    assert oracle.generate(0x5EED0005, 16, oracle.PROTEIN) == b"LILPNRYEGDTAGAKD"
    assert oracle.generate(0x5EED0006, 16, oracle.PROTEIN) == b"PDGAGPEFMCEANPWA"
    assert oracle.generate(0x5EED4000, 16) == b"TGTCTCATGCTTGAAC"
    assert oracle.generate(0x5EED4001, 16) == b"TCATCACCTCACTCAA"


def test_readme_table(oracle):
    r = oracle.fill("GCATGCU", "GATTACA", 1, 1, 1, want_scores=True, want_codes=True)
    rows = [ln.split() for ln in README_TABLE.strip().splitlines()]
    for j, row in enumerate(rows):
        for i, cell in enumerate(row):
            s, bits = cell.split(":")
            diag, left, up = (int(b) for b in bits)
            assert r.scores[j, i] == int(s), (i, j)
            code = int(r.codes[j, i])
            assert (code & 1, (code >> 1) & 1, (code >> 2) & 1) == (diag, left, up), (i, j)
    assert (r.final_score, r.branch_count, r.greatest_abs, r.count) == (0, 12, 5, 3)


def test_readme_alignments(oracle):
    # README:117-149
    r = oracle.fill("GT", "GT", 1, 1, 1, want_codes=True)
    assert oracle.enumerate_alignments("GT", "GT", r.codes) == [(b"GT", b"GT")]
    r = oracle.fill("GT", "GA", 1, 1, 1, want_codes=True)
    assert oracle.enumerate_alignments("GT", "GA", r.codes) == [(b"GT", b"GA")]
    r = oracle.fill("GAT", "GTA", 1, 1, 1, want_codes=True)
    assert oracle.enumerate_alignments("GAT", "GTA", r.codes) == [(b"G-AT", b"GTA-"), (b"GAT-", b"G-TA")]
    assert r.final_score == 0 and r.count == 2
    r = oracle.fill("GCATGCU", "GATTACA", 0, 0, 0)
    assert r.count == 48639 and r.branch_count == 49  # Delannoy(7,7)


def test_survey_goldens(oracle):
    # SURVEY.md 8c rows recorded from the unmodified reference
    t, s = oracle.generate_pair(0x5EED0002, 1000, 1000)
    r = oracle.fill(t, s, 1, 1, 1)
    assert (r.final_score, r.branch_count, r.greatest_abs, r.count) == (79, 353401, 998, 0)
    t, s = oracle.generate_pair(0x5EED0005, 3000, 3000, oracle.PROTEIN)
    r = oracle.fill(t, s, 2, 1, 2)
    assert (r.final_score, r.branch_count, r.greatest_abs, r.count) == (-1665, 5234118, 5996, 0)
    for seed, exp in ((0x5EED4000, (19, 23713, 254, 387701138034524160)),
                      (0x5EED4002, (29, 22912, 254, 108460706365440)),
                      (0x5EED4000 + 1999998, (19, 22090, 254, 4971798065203200))):
        t, s = oracle.generate_pair(seed, 256, 256)
        r = oracle.fill(t, s, 1, 1, 1)
        assert (r.final_score, r.branch_count, r.greatest_abs, r.count) == exp
    t, s = oracle.generate_pair(0x5EED0002, 10000, 10000)
    r = oracle.fill(t, s, 1, 1, 1)
    assert (r.final_score, r.branch_count, r.greatest_abs, r.count) == (1056, 34377799, 9998, 0)


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_golden_json(oracle, case):
    t, s = case_strings(oracle, case)
    assert t[:16].decode() == case["first16_top"] and s[:16].decode() == case["first16_side"]
    r = oracle.fill(t, s, case["m"], case["k"], case["d"])
    assert r.final_score == case["final_score"]
    assert r.branch_count == case["branch_count"]
    assert r.greatest_abs == case["greatest_abs"]
    assert f"{r.table_hash:016x}" == case["table_hash"]
    assert f"{r.arrow_hash:016x}" == case["arrow_hash"]
    assert r.count == case["count_u64"]
    assert f"{r.count_hash:016x}" == case["count_hash"]
    if "reference_solution_count" in case:
        assert (r.count & 0xFFFFFFFF) == case["reference_solution_count"]


def test_packed_layout_matches_codes(oracle):
    t, s = oracle.generate_pair(0x5EED0777, 77, 45)
    r = oracle.fill(t, s, 1, 1, 1, want_codes=True, want_packed=True)
    for j in range(1, 46):
        for i in range(1, 78):
            byte = int(r.packed[j - 1, (i - 1) // 2])
            nib = (byte >> 4) if (i - 1) & 1 else (byte & 0xF)
            assert nib == (int(r.codes[j, i]) & 7)


# ---- against the compiled reference (build container, or prebuilt oracle/_ref) ----
def _need_ref(oracle):
    if not oracle.have_reference():
        pytest.skip("oracle/_ref not built (no /root/reference here)")


def test_reference_struct_sizes(oracle):
    _need_ref(oracle)
    assert oracle.ref().nwref_sizeof_score_cell() == 104
    assert oracle.ref().nwref_sizeof_walk_cell() == 32


def test_oracle_vs_reference_tables(oracle):
    _need_ref(oracle)
    rng = random.Random(1234)
    n_enum = 0
    for _ in range(200):
        a = rng.choice([0, 1, 2, 3, 5, 8, 13, 31, 32, 33, 64, 100])
        b = rng.choice([1, 2, 3, 5, 8, 13, 31, 32, 33, 64, 100])
        alpha = rng.choice([b"ACGT", b"AB", b"ARNDCQEGHILKMFPSTWYV", b"A"])
        t = bytes(rng.choice(alpha) for _ in range(a))
        s = bytes(rng.choice(alpha) for _ in range(b))
        m, k, d = (rng.choice([0, 1, 2, 3, -1, -2, 5, 10]) for _ in range(3))
        o = oracle.fill(t, s, m, k, d, want_scores=True, want_codes=True)
        small = a * b <= 36 and (m, k, d) != (0, 0, 0)
        r = oracle.reference_fill(t, s, m, k, d, want_scores=True, want_codes=True, enumerate_count=small)
        assert np.array_equal(o.scores, r.scores) and np.array_equal(o.codes, r.codes), (t, s, m, k, d)
        assert (o.final_score, o.branch_count, o.greatest_abs, o.table_hash) == \
            (r.final_score, r.branch_count, r.greatest_abs, r.table_hash)
        if small:
            assert (o.count & 0xFFFFFFFF) == r.count
            n_enum += 1
    assert n_enum > 20


def test_oracle_vs_reference_threads_and_size(oracle):
    _need_ref(oracle)
    t, s = oracle.generate_pair(0x5EED0002, 1500, 1200)
    o = oracle.fill(t, s, 1, 1, 1)
    for threads in (1, 2, 3):
        r = oracle.reference_fill(t, s, 1, 1, 1, threads=threads)
        assert (o.final_score, o.branch_count, o.greatest_abs, o.table_hash) == \
            (r.final_score, r.branch_count, r.greatest_abs, r.table_hash)


def test_enumeration_order_vs_reference_cli(oracle):
    _need_ref(oracle)
    cli = oracle.reference_cli()
    rng = random.Random(99)
    for _ in range(25):
        a, b = rng.randint(1, 9), rng.randint(1, 9)
        t = bytes(rng.choice(b"ACGT") for _ in range(a))
        s = bytes(rng.choice(b"ACGT") for _ in range(b))
        m, k, d = rng.choice([(1, 1, 1), (2, 1, 2), (1, 0, 1)])
        out = subprocess.run([cli, str(m), str(k), str(d)], input=t + b" " + s + b"\n",
                             capture_output=True, check=True).stdout
        lines = out.split(b"\n")
        ref_pairs = [(lines[i], lines[i + 1]) for i in range(0, len(lines) - 2, 3)]
        o = oracle.fill(t, s, m, k, d, want_codes=True)
        assert oracle.enumerate_alignments(t, s, o.codes) == ref_pairs, (t, s, m, k, d)


def test_reference_cli_readme_example(oracle):
    _need_ref(oracle)
    p = subprocess.run([oracle.reference_cli(), "-q", "-s", "1", "1", "1"], input=b"GCATGCU GATTACA\n",
                       capture_output=True, check=True)
    assert p.stderr == b"3 optimal alignments\nOptimal score is 0\n"  # README:153-155, on stderr
    assert p.stdout == b""


def test_package_generator_matches_the_oracle_generator(oracle):
    """bench.py and the tools make their inputs with the product package's own SplitMix64 generator
    (needleman-wunsch_b200.generate, numpy); it must produce exactly the oracle's strings."""
    import nw_b200 as nwb
    for seed, n, alpha in [(0x5EED0002, 1000, nwb.DNA), (0x5EED0003, 999, nwb.DNA), (0x5EED0005, 777, nwb.PROTEIN),
                           (0x5EED0030, 4096, nwb.DNA), (2**64 - 5, 100, "AB"), (0, 1, nwb.DNA)]:
        assert nwb.generate(seed, n, alpha) == oracle.generate(seed, n, alpha)
    assert nwb.generate(0x5EED0002, 16) == b"CTTTCCTTAGCAGTTA" and nwb.generate(0x5EED0006, 16, nwb.PROTEIN) == b"PDGAGPEFMCEANPWA"
    many = nwb.generate(0x5EED4000, 256, nwb.DNA, count=50, seed_stride=2)
    for p in (0, 1, 7, 49):
        assert many[256 * p:256 * (p + 1)] == oracle.generate(0x5EED4000 + 2 * p, 256)
    assert nwb.generate_pair(0x5EED0030, 33, 17) == oracle.generate_pair(0x5EED0030, 33, 17)
