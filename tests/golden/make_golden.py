"""Generate tests/golden/golden.json from the UNMODIFIED reference compiled into
oracle/_ref (oracle/Makefile).  Run in the build container, where
/root/reference exists:   python tests/golden/make_golden.py

Every number is read out of the reference's own tables by oracle/ref_harness.c
(final score, branch count, greatest_abs under tflag, table/arrow hashes) or is
the reference's own enumeration count (only where it terminates quickly).
`count_u64` for the larger cases cannot come from the reference (its
enumeration is exponential): it is the oracle's path-count DP, recorded so that
regressions are caught, and marked "count_source": "oracle-dp".
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

CASES = [
    # (name, top spec, side spec, m, k, d, enumerate_with_reference)
    ("readme_gt_gt", "GT", "GT", 1, 1, 1, True),
    ("readme_gt_ga", "GT", "GA", 1, 1, 1, True),
    ("readme_gat_gta", "GAT", "GTA", 1, 1, 1, True),
    ("readme_gcatgcu_gattaca", "GCATGCU", "GATTACA", 1, 1, 1, True),
    ("gcatgcu_gattaca_000", "GCATGCU", "GATTACA", 0, 0, 0, True),
    ("empty_top", "", "ACG", 1, 1, 1, True),
    ("dna_0x5EED0002_1000", ("dna", 0x5EED0002, 1000, 1000), None, 1, 1, 1, False),
    ("dna_0x5EED0002_2000", ("dna", 0x5EED0002, 2000, 2000), None, 1, 1, 1, False),
    ("protein_0x5EED0005_3000", ("protein", 0x5EED0005, 3000, 3000), None, 2, 1, 2, False),
    ("cfg4_pair0", ("dna", 0x5EED4000, 256, 256), None, 1, 1, 1, False),
    ("cfg4_pair1", ("dna", 0x5EED4002, 256, 256), None, 1, 1, 1, False),
    ("cfg4_pair999999", ("dna", 0x5EED4000 + 1999998, 256, 256), None, 1, 1, 1, False),
    ("dna_ragged_777x1301", ("dna", 0x5EED0100, 777, 1301), None, 1, 1, 1, False),
    ("dna_neg_params", ("dna", 0x5EED0200, 300, 500), None, -1, 3, -2, False),
    ("protein_blosumish", ("protein", 0x5EED0300, 600, 400), None, 5, 4, 3, False),
    ("dna_small_enum", ("dna", 0x5EED0400, 12, 11), None, 1, 1, 1, True),
]


def strings(top_spec, side_spec):
    if isinstance(top_spec, tuple):
        kind, seed, a, b = top_spec
        alpha = oracle.DNA if kind == "dna" else oracle.PROTEIN
        t, s = oracle.generate_pair(seed, a, b, alpha)
        return t, s, {"alphabet": kind, "seed": seed, "top_len": a, "side_len": b}
    return top_spec.encode(), side_spec.encode(), {"top": top_spec, "side": side_spec}


def main():
    oracle.build(with_reference=True)
    out = []
    for name, ts, ss, m, k, d, enum in CASES:
        t, s, spec = strings(ts, ss)
        r = oracle.reference_fill(t, s, m, k, d, tflag=True, enumerate_count=enum)
        o = oracle.fill(t, s, m, k, d)
        assert (o.final_score, o.branch_count, o.greatest_abs, o.table_hash, o.arrow_hash) == \
            (r.final_score, r.branch_count, r.greatest_abs, r.table_hash, r.arrow_hash), name
        rec = dict(name=name, m=m, k=k, d=d, **spec,
                   final_score=r.final_score, branch_count=r.branch_count, greatest_abs=r.greatest_abs,
                   table_hash=f"{r.table_hash:016x}", arrow_hash=f"{r.arrow_hash:016x}",
                   first16_top=t[:16].decode(), first16_side=s[:16].decode())
        if enum:
            assert (o.count & 0xFFFFFFFF) == r.count, name
            rec["reference_solution_count"] = r.count
            rec["count_u64"] = o.count
            rec["count_source"] = "reference-enumeration"
        else:
            rec["count_u64"] = o.count
            rec["count_source"] = "oracle-dp"
        rec["count_hash"] = f"{o.count_hash:016x}"
        out.append(rec)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, len(out), "cases")


if __name__ == "__main__":
    main()
