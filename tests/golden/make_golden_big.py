"""Golden summary values for the two BASELINE configs the compiled reference
cannot hold in memory (config 3: 100k x 100k needs 1.36 TB; config 5: 30k x 30k
needs 122 GB).  Source: the CPU oracle (oracle/nw_oracle.c), which is pinned
bit-exactly against the reference at every size the reference fits
(tests/test_oracle.py).  Takes a few minutes:  python tests/golden/make_golden_big.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

CASES = [
    ("config2_dna_10k", "dna", 0x5EED0002, 10000, 10000, 1, 1, 1),
    ("config5_protein_30k", "protein", 0x5EED0005, 30000, 30000, 2, 1, 2),
    ("config3_dna_100k", "dna", 0x5EED0030, 100000, 100000, 1, 1, 1),
]

out = []
for name, kind, seed, a, b, m, k, d in CASES:
    alpha = oracle.DNA if kind == "dna" else oracle.PROTEIN
    t, s = oracle.generate_pair(seed, a, b, alpha)
    t0 = time.time()
    r = oracle.fill(t, s, m, k, d)
    out.append(dict(name=name, alphabet=kind, seed=seed, top_len=a, side_len=b, m=m, k=k, d=d,
                    final_score=r.final_score, branch_count=r.branch_count, greatest_abs=r.greatest_abs,
                    table_hash=f"{r.table_hash:016x}", arrow_hash=f"{r.arrow_hash:016x}",
                    count_u64=r.count, count_hash=f"{r.count_hash:016x}",
                    lastrow_count_hash=f"{r.lastrow_count_hash:016x}",
                    lastcol_count_hash=f"{r.lastcol_count_hash:016x}",
                    source="oracle", oracle_seconds=round(time.time() - t0, 1)))
    print(out[-1], flush=True)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_big.json"), "w") as f:
        json.dump(out, f, indent=1)
