"""Golden values for the BASELINE configs the compiled reference cannot hold in memory (config 3: 100k x 100k
needs 1.36 TB; config 5: 30k x 30k needs 122 GB; config 4: one million pairs), plus two near-identical pairs
whose optimal-alignment count does NOT vanish mod 2^64.

Source: the CPU oracle (oracle/nw_oracle.c), which is pinned bit-exactly against the reference at every size the
reference fits (tests/test_oracle.py), including its order-independent digests (oracle/nw_oracle.h: arrow_digest
is also computed by oracle/ref_harness.c over the reference's own walk table).  The digests are what the GPU
computes on device at full size (csrc/nwb_digest.cuh).

    python tests/golden/make_golden_big.py            # ~10 minutes on 8 cores
"""
import json
import os
import sys
import time
from multiprocessing import Pool

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

M64 = (1 << 64) - 1
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    ("config2_dna_10k", "dna", 0x5EED0002, 10000, 10000, 1, 1, 1),
    ("config5_protein_30k", "protein", 0x5EED0005, 30000, 30000, 2, 1, 2),
    ("config3_dna_100k", "dna", 0x5EED0030, 100000, 100000, 1, 1, 1),
]


def mix64(pos: int, x: int) -> int:
    """oracle/nw_oracle.h nwo_mix64."""
    z = ((pos + 1) * 0x9E3779B97F4A7C15 + x) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def mutated_pair(seed: int, n: int, subs: int, indels: int, alphabet: str = oracle.DNA):
    """A string and a mutated copy of it (SplitMix64-driven, so reproducible anywhere): `subs` substitutions and
    `indels` short insertions/deletions.  Few branch points: the alignment count stays non-zero mod 2^64."""
    top = oracle.generate(seed, n, alphabet)
    side = bytearray(top)
    r = oracle.generate(seed + 7, 4 * (subs + indels) + 8, "0123456789ABCDEF")
    vals = [int(chr(c), 16) for c in r]
    state = seed

    def rnd(mod):
        nonlocal state
        state = (state * 6364136223846793005 + 1442695040888963407) & M64
        return (state >> 33) % mod

    for i in range(subs):
        pos = rnd(len(side))
        side[pos] = ord(alphabet[(alphabet.index(chr(side[pos])) + 1 + vals[i] % (len(alphabet) - 1)) % len(alphabet)])
    for i in range(indels):
        pos = rnd(len(side) - 8)
        if i & 1:
            del side[pos:pos + 1 + vals[subs + i] % 3]
        else:
            side[pos:pos] = bytes(ord(alphabet[v % len(alphabet)]) for v in vals[subs + i:subs + i + 1 + vals[subs + i] % 3])
    return top, bytes(side)


MUTATED = [
    # name, seed, n, substitutions, indels, (m, k, d)
    ("mutated_dna_30k", 0x5EED0A30, 30000, 300, 40, (1, 1, 1)),
    ("mutated_dna_100k", 0x5EED0A31, 100000, 500, 60, (1, 1, 1)),
]


def shard_digests(args):
    first, n = args
    d = [0, 0, 0, 0]
    for p in range(first, first + n):
        t, s = oracle.generate_pair(0x5EED4000 + 2 * p, 256, 256)
        o = oracle.fill(t, s, 1, 1, 1)
        d[0] = (d[0] + mix64(p, o.arrow_digest)) & M64
        d[1] = (d[1] + mix64(p, o.final_score & M64)) & M64
        d[2] = (d[2] + mix64(p, o.branch_count)) & M64
        d[3] = (d[3] + mix64(p, o.count)) & M64
    return first, n, d


def record(name, t, s, m, k, d, **extra):
    t0 = time.time()
    r = oracle.fill(t, s, m, k, d)
    rec = dict(name=name, top_len=len(t), side_len=len(s), m=m, k=k, d=d,
               final_score=r.final_score, branch_count=r.branch_count, greatest_abs=r.greatest_abs,
               table_hash=f"{r.table_hash:016x}", arrow_hash=f"{r.arrow_hash:016x}",
               count_u64=r.count, count_hash=f"{r.count_hash:016x}",
               lastrow_count_hash=f"{r.lastrow_count_hash:016x}", lastcol_count_hash=f"{r.lastcol_count_hash:016x}",
               arrow_digest=f"{r.arrow_digest:016x}", lastrow_count_digest=f"{r.lastrow_count_digest:016x}",
               lastcol_count_digest=f"{r.lastcol_count_digest:016x}",
               source="oracle", oracle_seconds=round(time.time() - t0, 1), **extra)
    print(rec, flush=True)
    return rec


def main():
    out = []
    for name, kind, seed, a, b, m, k, d in CASES:
        alpha = oracle.DNA if kind == "dna" else oracle.PROTEIN
        t, s = oracle.generate_pair(seed, a, b, alpha)
        out.append(record(name, t, s, m, k, d, alphabet=kind, seed=seed))
    for name, seed, n, subs, indels, (m, k, d) in MUTATED:
        t, s = mutated_pair(seed, n, subs, indels)
        out.append(record(name, t, s, m, k, d, alphabet="dna", seed=seed, mutated=dict(subs=subs, indels=indels)))
    # config 4: one million 256 x 256 DNA pairs, pair p seeded 0x5EED4000 + 2p; digests per shard of 125,000 pairs
    # (= one GPU's share on 8 GPUs) in chunks, combined by addition
    per, shards, chunk = 125_000, 8, 5_000
    jobs = [(sh * per + c, chunk) for sh in range(shards) for c in range(0, per, chunk)]
    t0 = time.time()
    with Pool(os.cpu_count() or 1) as pool:
        res = pool.map(shard_digests, jobs, chunksize=1)
    sh_dig = [[0, 0, 0, 0] for _ in range(shards)]
    for first, n, d in res:
        sh = first // per
        for i in range(4):
            sh_dig[sh][i] = (sh_dig[sh][i] + d[i]) & M64
    out.append(dict(name="config4_batch_1M", pairs=per * shards, pairs_per_shard=per, top_len=256, side_len=256,
                    m=1, k=1, d=1, seed_rule="pair p: top 0x5EED4000 + 2p, side 0x5EED4000 + 2p + 1",
                    digest_rule="sum over pairs p of mix64(p, x_p); x = arrow digest, score (sign-extended), branch count, count",
                    shard_digests=[dict(shard=i, first_pair=i * per,
                                        arrow=f"{d[0]:016x}", score=f"{d[1]:016x}", branch=f"{d[2]:016x}", count=f"{d[3]:016x}")
                                   for i, d in enumerate(sh_dig)],
                    source="oracle", oracle_seconds=round(time.time() - t0, 1)))
    print(out[-1], flush=True)
    with open(os.path.join(HERE, "golden_big.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
