"""The C99 host shell (needleman-wunsch_b200/host/needleman-wunsch) against the
reference's own CLI (oracle/_ref/needleman-wunsch, the unmodified sources compiled
by oracle/Makefile): stdout, stderr and exit code byte for byte.  Paths that end
before the fill (usage, operand and input errors) run without a GPU; everything
that computes is marked gpu."""
import itertools
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "needleman-wunsch_b200", "host", "needleman-wunsch")


def run(exe, args, stdin=b""):
    # same argv[0] for both programs: getopt and the error messages print it
    p = subprocess.run(["needleman-wunsch"] + list(args), executable=exe, input=stdin, capture_output=True)
    return p.returncode, p.stdout, p.stderr


@pytest.fixture(scope="module")
def cli(oracle, nwb):
    ref = oracle.reference_cli()
    if ref is None:
        pytest.skip("oracle/_ref/needleman-wunsch not built (no /root/reference here)")
    if not os.path.exists(nwb.LIB_PATH):
        nwb.build()
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "needleman-wunsch_b200", "host")], check=True)
    return ref


def same(ref, args, stdin=b""):
    a = run(ref, args, stdin)
    b = run(OURS, args, stdin)
    assert a == b, (args, stdin, a, b)
    return a


def test_usage_and_operand_errors(cli, tmp_path):
    for args in (["-h"], ["-x", "1", "1", "1"], [], ["1"], ["1", "1"], ["1", "1", "1", "1"], ["-q", "1", "1"],
                 ["-p", "1", "1", "1", "1"], ["-p", "0", "1", "1", "1"], ["-p", "abc", "1", "1", "1"],
                 ["-f"], ["-f", str(tmp_path / "missing.txt"), "1", "1", "1"]):
        rc, out, err = same(cli, args, b"GT GA\n")
        assert rc == 1 and out == b""


def test_input_errors(cli, tmp_path):
    for stdin in (b"", b"ACG", b"ACG ", b"ACG\n", b"   ", b"\n\n"):
        rc, out, err = same(cli, ["1", "1", "1"], stdin)
        assert rc == 1 and b"EOF too early" in err
    f = tmp_path / "one.txt"
    f.write_bytes(b"ACGT")
    rc, _, err = same(cli, ["-f", str(f), "1", "1", "1"], b"GT GA\n")
    assert rc == 1


FLAG_SETS = [[]] + [list(c) for n in (1, 2, 3) for c in itertools.combinations(["-c", "-l", "-q", "-s", "-t", "-u"], n)] + \
    [["-c", "-l", "-s", "-t", "-u"], ["-q", "-l", "-s", "-t"], ["-c", "-q", "-s", "-t", "-u"]]


@pytest.mark.gpu
def test_readme_examples(cli):
    # README:117-173
    same(cli, ["1", "1", "1"], b"GT GT\n")
    same(cli, ["1", "1", "1"], b"GT GA\n")
    rc, out, err = same(cli, ["-s", "-l", "1", "1", "1"], b"GAT GTA\n")
    assert out.startswith(b"G-AT\nGTA-\n") and b"2 optimal alignments" in err
    rc, out, err = same(cli, ["-q", "-s", "-t", "1", "1", "1"], b"GCATGCU GATTACA\n")
    assert err == b"3 optimal alignments\nOptimal score is 0\n"
    rc, out, err = same(cli, ["-q", "-s", "0", "0", "0"], b"GCATGCU GATTACA\n")
    assert err.startswith(b"48639 optimal alignments")


@pytest.mark.gpu
def test_flag_combinations(cli, tmp_path):
    rng = random.Random(2024)
    inputs = [b"GCATGCU GATTACA\n", b" ACG", b"A A", b"ACGT\tTGCA  trailing text is ignored\n", b"AAAA AAAA"]
    for _ in range(6):
        a, b = rng.randint(1, 7), rng.randint(1, 7)
        inputs.append(bytes(rng.choice(b"ACGT") for _ in range(a)) + b"\n" + bytes(rng.choice(b"ACGT") for _ in range(b)))
    schemes = [("1", "1", "1"), ("2", "1", "2"), ("0", "0", "0"), ("-1", "3", "-2"), ("5", "4", "3"), ("1", "3", "1")]
    n = 0
    for stdin in inputs:
        for flags in rng.sample(FLAG_SETS, 3):   # every run of ours pays a CUDA context: keep the GPU suite short
            mkd = rng.choice(schemes)
            same(cli, flags + list(mkd), stdin)
            n += 1
    # -f reads the file and ignores stdin; -p N (N > 1) is accepted
    f = tmp_path / "pair.txt"
    f.write_bytes(b"GATTACA\nGCATGCU\nmore\n")
    same(cli, ["-f", str(f), "-s", "-t", "1", "1", "1"], b"GT GA\n")
    same(cli, ["-p", "2", "-s", "1", "1", "1"], b"GCATGCU GATTACA\n")
    assert n >= 30


@pytest.mark.gpu
def test_summary_beyond_32_bits(cli, oracle):
    """-q -s on a 256 x 256 pair: the count exceeds 2^32; the reference prints the low 32 bits with %d
    (it would need to enumerate 3.9e17 alignments; ours reads the fused DP).  SURVEY.md 8c goldens."""
    t, s = oracle.generate_pair(0x5EED4000, 256, 256)
    rc, out, err = run(OURS, ["-q", "-s", "1", "1", "1"], t + b" " + s + b"\n")
    assert (rc, out) == (0, b"")
    assert err == b"-2087714816 optimal alignments\nOptimal score is 19\n"


@pytest.mark.gpu
def test_p_splits_the_table_over_gpus(cli, oracle, nwb):
    """-p N: the reference's worker count (needleman-wunsch.c:738-742) selects min(N, devices) GPUs here, the table
    split into column strips (nwb_fill_on).  Same text as the reference; with one GPU the flag changes nothing."""
    rng = random.Random(77)
    top = bytearray(rng.choice(b"ACGT") for _ in range(700))          # 3 strips of 256 columns
    side = bytearray(top)
    for pos in (90, 333, 610):
        side[pos] = ord("A") if side[pos] != ord("A") else ord("C")    # substitutions: few optimal alignments
    stdin = bytes(top) + b"\n" + bytes(side) + b"\n"
    for flags in (["-l"], ["-q", "-l"], []):
        a = same(cli, ["-p", "2"] + flags + ["1", "1", "1"], stdin)
        assert a[0] == 0
    same(cli, ["-p", "8", "-s", "-t", "1", "1", "1"], b"GCATGCU GATTACA\n")   # -t keeps the fill on one device
    same(cli, ["-p", "8", "-l", "2", "1", "2"], b"GCATGCU GATTACA\n")         # more GPUs asked for than strips
    # count and score beyond what the reference can enumerate: against the one-GPU run and the oracle
    t, s = oracle.generate_pair(0x5EED0B10, 3000, 2000)
    one = run(OURS, ["-q", "-s", "1", "1", "1"], t + b" " + s + b"\n")
    two = run(OURS, ["-p", "2", "-q", "-s", "1", "1", "1"], t + b" " + s + b"\n")
    assert one == two and one[0] == 0
    assert b"Optimal score is %d\n" % oracle.fill(t, s, 1, 1, 1).final_score in one[2]


# ---- batch front-end: needleman-wunsch-batch vs the reference CLI looped per pair -----

BATCH = os.path.join(ROOT, "needleman-wunsch_b200", "host", "needleman-wunsch-batch")


def run_batch(args, stdin=b""):
    p = subprocess.run(["needleman-wunsch-batch"] + list(args), executable=BATCH, input=stdin, capture_output=True)
    return p.returncode, p.stdout, p.stderr


def test_batch_cli_usage_and_input_errors(cli):
    for args in (["-h"], ["-x", "1", "1", "1"], [], ["1", "1"], ["1", "1", "1", "1"]):
        rc, out, err = run_batch(args, b"GT GA\n")
        assert rc == 1 and out == b"" and b"usage: needleman-wunsch-batch" in err
    for stdin in (b"", b"ACG", b"ACG GT\nAC", b"  \n"):
        rc, out, err = run_batch(["1", "1", "1"], stdin)
        assert rc == 1 and out == b"" and b"got EOF too early when reading input strings" in err


@pytest.mark.gpu
def test_batch_cli_matches_the_reference_looped_per_pair(cli):
    rng = random.Random(77)
    pairs = [("GCATGCU", "GATTACA"), ("GAT", "GTA"), ("GT", "GT"), ("A", "C")] + \
            [("".join(rng.choice("ACGT") for _ in range(rng.randint(1, 11))),
              "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 11)))) for _ in range(40)]
    text = "\n".join(f"{a} {b}" for a, b in pairs).encode() + b"\n"
    # any m k d the reference takes (needleman-wunsch.c:783-785): the last three run the int32 batch engine
    # (negative operands need "--": both front-ends parse their options with getopt, needleman-wunsch.c:723)
    for mkd in (("1", "1", "1"), ("2", "1", "2"), ("0", "0", "0"), ("1", "3", "1"), ("5", "4", "3"), ("--", "-1", "3", "-2")):
        for flags in ([], ["-s"]):
            rc, out, err = run_batch(flags + list(mkd), text)
            assert rc == 0 and out == b""
            want = b""
            for a, b in pairs:
                r = run(cli, ["-q"] + flags + list(mkd), f"{a} {b}\n".encode())
                assert r[0] == 0
                # without -s the reference prints nothing under -q; the batch front-end always prints the score
                want += r[2] if flags else b"Optimal score is %d\n" % int(run(cli, ["-q", "-s"] + list(mkd), f"{a} {b}\n".encode())[2].split()[-1])
            assert err == want, (mkd, flags)
