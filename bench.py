#!/usr/bin/env python
"""bench.py -- GCUPS of the Needleman-Wunsch score-table fill on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json): config 3, one pair of 100,000 x 100,000 synthetic DNA
(SplitMix64 seeds 0x5EED0030/31), m=k=d=1 -- the configuration the target
("the 100k x 100k fill at >= 50% of the INT/DPX issue roofline") is quoted on.
A step = one fill of the whole table (scores in registers, every arrow set
written as 4-bit codes).  The K steps are a QUEUE of fills: a few plans per GPU
(workspaces with their own 5 GB arrow table and stream) take them round robin
and consecutive fills overlap on the device (NWB_QUEUE).  At
N > 1 EVERY table is cut into column strips across the ranks (boundary columns
stream rank-to-rank through CUDA-IPC peer memory over NVLink; rank r starts fill
e + 1 while the ranks to its right are still on fill e, nwb_plan_run_pipelined);
total work fixed -> "scaling": "strong".

value   = interior cells of K tables / device time from the first launch to the
          last completion, strings resident in HBM (CUDA events, max over ranks;
          barrier + synchronize on both sides, none between the steps);
e2e     = the same through the C ABI with HOST strings, wall clock: per step
          nwb_plan_upload (H2D) + nwb_plan_run[_pipelined], and nwb_plan_summary
          (D2H) of every step's result;
latency = what ONE fill takes from launch to completion (nothing else on the
          GPUs), and the one-call path nwb_fill() that a CLI run takes.

Every result is checked against tests/golden/golden_big.json -- score, branch
count and the digest of the WHOLE arrow table, computed on the device and summed
over the ranks -- and the process exits non-zero on a mismatch ("golden_ok").
`extras` holds the other BASELINE configs (2, 4, 5 and 3 with the -s count), each
with its own device time, host-buffer e2e time, golden check and CPU baseline.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GCUPS (giga cell updates/sec), NW score-table fill"
A = B = 100_000
SEED = 0x5EED0030
M_, K_, D_ = 1, 1, 1
OPS_PER_CELL = 10          # SURVEY.md 8d convention: score + arrows
SM_MAX_MHZ_FALLBACK = 1965.0
CPU_SAMPLE_N = 8000        # the reference arm and cpu_baseline time the same 8000 x 8000 prefix of config 3
M64 = (1 << 64) - 1


def queue_depth(top_len: int, world: int, sms: int) -> int:
    """Plans (fills in flight) per GPU for the queue of fills: as many fills as fit on the SMs at once -- a fill's share of a
    rank is ceil(strips of the rank / 3) blocks of nwb_fill_hx_kernel, one block per SM -- plus one whose blocks move in as
    others leave (tools/sweep_queue.sh, profiles/r02_queue_sweep.txt: more plans change nothing)."""
    n_strips = (top_len + 255) // 256
    nloc = -(-n_strips // world)
    blocks = max(1, -(-nloc // 3))
    return min(12, max(3, sms // blocks + 1))


def workload_config(n_gpus: int) -> dict:
    return {"workload": "config3: single pair 100000x100000 synthetic DNA (seeds 0x5EED0030/31), m=1 k=1 d=1, "
                        "fill + 4-bit arrow table, no count",
            "top_len": A, "side_len": B, "m": M_, "k": K_, "d": D_, "cells_per_step": A * B,
            "parallelism": (f"column strips x{n_gpus} over peer memory, consecutive fills pipelined across the ranks"
                            if n_gpus > 1 else "single GPU"),
            "l2": "each step writes a 5.0 GB arrow table (>> 126 MB L2); no explicit flush needed"}


def goldens() -> dict:
    with open(os.path.join(ROOT, "tests", "golden", "golden_big.json")) as f:
        return {c["name"]: c for c in json.load(f)}


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- CPU baselines (the checker, timed)
def _thread_candidates(ncpu: int) -> list[int]:
    return sorted({1, 2, min(4, ncpu), ncpu})


def cpu_config3_sample(oracle, nwb, steps: int = 1, warmup: int = 0) -> dict:
    """The reference's compute_table_scores() (oracle/_ref: the unmodified sources compiled by oracle/Makefile) on
    the CPU_SAMPLE_N^2 prefix of config 3 -- the full table would need 1.36 TB of its 136 B/cell tables."""
    n = CPU_SAMPLE_N
    t, s = nwb.generate_pair(SEED, n, n)
    ncpu = os.cpu_count() or 1
    if not oracle.have_reference():
        t0 = time.perf_counter()
        oracle.fill(t, s, M_, K_, D_)
        dt = time.perf_counter() - t0
        return {"value": n * n / dt / 1e9, "unit": "GCUPS", "cores": 1, "kind": "port", "host_cores": ncpu, "seconds": dt,
                "sample": f"{n}x{n} prefix of config 3, oracle/nw_oracle.c (scalar C port)"}
    # the reference's -p filler stops scaling at 2 threads (README:72-75): probe, keep the best
    probe = {T: oracle.reference_fill(t, s, M_, K_, D_, threads=T, tflag=False).fill_seconds for T in _thread_candidates(ncpu)}
    best = min(probe, key=probe.get)
    for _ in range(warmup):
        oracle.reference_fill(t, s, M_, K_, D_, threads=best, tflag=False)
    times = [probe[best]] if steps <= 1 else [oracle.reference_fill(t, s, M_, K_, D_, threads=best, tflag=False).fill_seconds
                                              for _ in range(steps)]
    sec = sum(times) / len(times)
    return {"value": n * n / sec / 1e9, "unit": "GCUPS", "cores": best, "kind": "reference", "host_cores": ncpu, "seconds": sec,
            "sample": f"{n}x{n} prefix of config 3 (the reference needs 136 B/cell: 1.36 TB for the full table), "
                      f"compute_table_scores() only, unmodified reference (oracle/_ref); seconds by -p threads: "
                      f"{ {k: round(v, 2) for k, v in probe.items()} }"}


def _ref_pairs_worker(args):
    first, n = args
    import nw_b200 as nwb
    import oracle
    tot = fill = 0.0
    for p in range(first, first + n):
        t, s = nwb.generate_pair(0x5EED4000 + 2 * p, 256, 256)
        r = oracle.reference_fill(t, s, M_, K_, D_, threads=1, tflag=False)
        tot += r.total_seconds
        fill += r.fill_seconds
    return tot, fill


def cpu_other_configs(oracle, nwb) -> dict:
    """config 2 at full size through the reference, config 4 as the reference's
    alloc/init/compute/free loop per pair (1 core and all cores), config 5 on a bounded prefix."""
    out = {}
    ncpu = os.cpu_count() or 1
    if not oracle.have_reference():
        return {"unavailable": "oracle/_ref is not built on this box"}
    # config 2: 10,000 x 10,000 at FULL size (13.6 GB of the reference's tables)
    t, s = nwb.generate_pair(0x5EED0002, 10000, 10000)
    r = oracle.reference_fill(t, s, 1, 1, 1, threads=1, tflag=False)
    out["config2_dna_10k"] = {"gcups_fill": 1e8 / r.fill_seconds / 1e9, "fill_seconds": r.fill_seconds,
                              "gcups_alloc_init_fill_free": 1e8 / r.total_seconds / 1e9, "total_seconds": r.total_seconds,
                              "cores": 1, "kind": "reference", "sample": "the full 10000x10000 table, -p 1 (serial path); "
                              "total = alloc_computation + init_computation + compute_table_scores + free_computation",
                              "golden_ok": bool((r.final_score, r.branch_count) == (1056, 34377799))}
    # config 4: the reference looped per pair
    n1 = 1000
    tot1, fill1 = _ref_pairs_worker((0, n1))
    from multiprocessing import get_context
    per = 640
    with get_context("fork").Pool(ncpu) as pool:
        t0 = time.perf_counter()
        pool.map(_ref_pairs_worker, [(2000 + i * per, per) for i in range(ncpu)])
        wall = time.perf_counter() - t0
    out["config4_batch"] = {"one_core": {"pairs": n1, "seconds": tot1, "gcups": n1 * 65536 / tot1 / 1e9,
                                         "fill_only_gcups": n1 * 65536 / fill1 / 1e9},
                            "all_cores": {"processes": ncpu, "pairs": ncpu * per, "wall_seconds": wall,
                                          "gcups": ncpu * per * 65536 / wall / 1e9},
                            "kind": "reference", "extrapolated_1M_pairs_seconds_all_cores": 1e6 / (ncpu * per / wall),
                            "sample": f"alloc/init/compute_table_scores/free per 256x256 pair (computation.c:51-214): {n1} pairs on "
                                      f"one core, {ncpu * per} pairs on {ncpu} processes"}
    # config 5: protein 2/1/2 on a bounded prefix (30,000^2 needs 122 GB and ~150 s)
    n5 = 8000
    t, s = nwb.generate_pair(0x5EED0005, n5, n5, nwb.PROTEIN)
    r = oracle.reference_fill(t, s, 2, 1, 2, threads=1, tflag=False)
    out["config5_protein_30k"] = {"gcups_fill": n5 * n5 / r.fill_seconds / 1e9, "fill_seconds": r.fill_seconds, "cores": 1,
                                  "kind": "reference", "sample": f"{n5}x{n5} prefix (the full 30000x30000 table needs 122 GB of "
                                  "the reference's tables and ~150 s)"}
    return out


# --------------------------------------------------------------------------- reference arm
def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import nw_b200 as nwb
    import oracle
    base = cpu_config3_sample(oracle, nwb, steps=max(1, args.steps), warmup=args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "GCUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["seconds"] * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------- our arm: the other configs
def run_extras(nwb, torch, dist, world, rank, local, barrier, gold) -> tuple[dict | None, bool]:
    out = {}
    ok_all = True
    hbm_peak = 6650.0
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm_peak = float(json.load(f).get("hbm_gbs", hbm_peak))
    except OSError:
        pass
    # ---- config 4: batch of 256 x 256 DNA pairs, pair p seeded 0x5EED4000 + 2p, sharded by rank (no communication)
    per = 125_000
    first, cnt = nwb.batch_partition(per * world, rank, world)
    assert cnt == per
    import numpy as np
    tcat = nwb.generate(0x5EED4000 + 2 * first, 256, nwb.DNA, count=per, seed_stride=2)
    scat = nwb.generate(0x5EED4000 + 2 * first + 1, 256, nwb.DNA, count=per, seed_stride=2)
    off = np.arange(per + 1, dtype=np.int64) * 256
    tpin, spin = nwb.PinnedBuffer(tcat), nwb.PinnedBuffer(scat)   # e2e inputs: page-locked host memory (nwb_host_alloc)
    g4 = gold["config4_batch_1M"]["shard_digests"][first // per]
    st = torch.cuda.current_stream().cuda_stream
    res4 = {}
    for name, flags in (("config4_batch", 0), ("config4_batch_with_count", nwb.WANT_COUNT)):
        bt = nwb.Batch.from_arrays(tcat, off, scat, off, M_, K_, D_, flags, device=local)
        for _ in range(2):
            bt.run()
        torch.cuda.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            bt.run(st)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / reps
        dg = bt.digest(first)
        want = tuple(int(g4[k], 16) for k in ("arrow", "score", "branch", "count"))
        ok = (dg[:3] == want[:3]) and (not flags or dg[3] == want[3])
        # end to end from HOST buffers: chunked H2D overlapped with the kernels, D2H of every pair's results
        barrier()
        t0 = time.perf_counter()
        e2e_reps = 3
        for _ in range(e2e_reps):
            bt.refill(tpin, spin)
            bt.fetch()
        e2e_ms = (time.perf_counter() - t0) / e2e_reps * 1e3
        t0 = time.perf_counter()
        for _ in range(e2e_reps):
            bt.refill(tcat, scat)
            bt.fetch()
        e2e_pageable_ms = (time.perf_counter() - t0) / e2e_reps * 1e3
        dg2 = bt.digest(first)
        ok = ok and dg2 == dg
        kname = bt.kernel_name()
        bt.close()
        vals = torch.tensor([ms, e2e_ms, 0.0 if ok else 1.0, e2e_pageable_ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        ms, e2e_ms, bad, e2e_pageable_ms = (float(x) for x in vals.tolist())
        ok_all = ok_all and bad == 0.0
        res4[name] = {"pairs_total": per * world, "pairs_per_gpu": per, "ms_per_pass": ms,
                      "gcups_total": per * world * 65536 / (ms * 1e-3) / 1e9, "kernel": kname,
                      "e2e_ms": e2e_ms, "e2e_gcups_total": per * world * 65536 / (e2e_ms * 1e-3) / 1e9,
                      "e2e_ms_from_pageable_host_memory": e2e_pageable_ms,
                      "e2e_h2d_bytes_per_gpu": 2 * per * 256, "e2e_d2h_bytes_per_gpu": per * (8 + (8 if flags else 0)),
                      "e2e_call": "nwb_batch_refill (host strings in page-locked memory from nwb_host_alloc, H2D in 7 chunks overlapped "
                                  "with the kernels, neighbouring chunks on two compute streams) + nwb_batch_fetch (D2H of every pair's "
                                  "score, branch count, count)",
                      "roofline_hbm": {"bound": "hbm", "unit": "GB/s", "peak": hbm_peak,
                                       "achieved": per * (32768 + 512 + 8) / (ms * 1e-3) / 1e9,
                                       "frac": per * (32768 + 512 + 8) / (ms * 1e-3) / 1e9 / hbm_peak,
                                       "algorithmic_bytes_per_pair": 32768 + 512 + 8,
                                       "note": "per pair: 32 KB of 4-bit arrow codes written + 512 B of strings read + score and "
                                               "branch counter; fill pass only (the count pass re-reads the codes)"} if not flags else None,
                      "golden_ok": bad == 0.0,
                      "golden": "every pair of every shard: arrow / score / branch / count digests vs golden_big.json config4_batch_1M",
                      "scaling": "weak (125,000 pairs per GPU; 8 GPUs = the 1M-pair config)"}
    out.update(res4)
    tpin.close()
    spin.close()
    if rank != 0:
        return None, ok_all

    def one(name, gname, seed, n, alpha, mkd):
        g = gold[gname]
        t, s = nwb.generate_pair(seed, n, n, alpha)
        rec = {}
        for label, flags in (("fill", 0), ("q_s", nwb.WANT_COUNT)):
            plan = nwb.Plan(n, n, flags, device=local)
            plan.upload(t, s)
            best = None
            for _ in range(3):
                plan.run(*mkd)
                sm = plan.summary()
                k = plan.kernel_ms()
                best = k if best is None else min(best, k)
            dig = plan.arrow_digest()
            ok = (sm.opt_score, sm.branch_count) == (g["final_score"], g["branch_count"]) and dig == int(g["arrow_digest"], 16)
            if flags:
                ok = ok and sm.count == g["count_u64"]
            rec[label] = {"kernel_ms": best, "gcups": n * n / (best * 1e-3) / 1e9, "kernel": plan.kernel_name(),
                          "golden_ok": bool(ok)}
            if flags:
                rec[label]["count_path"] = plan.count_path() + (" (gave up -> dense)" if sm.count_path == 4 else "")
                rec[label]["count_rows_visited"] = sm.count_rows
                rec[label]["count_u64"] = sm.count
            plan.close()
            # end to end through nwb_fill(): host strings in, summary out
            nwb.fill(t, s, *mkd, flags, device=local).close()
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                tab = nwb.fill(t, s, *mkd, flags, device=local)
                okf = (tab.opt_score, tab.branch_count) == (g["final_score"], g["branch_count"])
                tab.close()
            e2e = (time.perf_counter() - t0) / reps
            rec[label]["e2e_ms"] = e2e * 1e3
            rec[label]["e2e_gcups"] = n * n / e2e / 1e9
            rec[label]["golden_ok"] = bool(rec[label]["golden_ok"] and okf)
        # a QUEUE of such fills (NWB_QUEUE: plans take the fills round robin and consecutive fills overlap on the GPU; DESIGN.md
        # 5.2): device-resident, wall clock from the first launch to the last completion, launches included
        if n < A:
            blocks = -(-(-(-n // 256)) // 3)
            nq = min(12, max(2, torch.cuda.get_device_properties(local).multi_processor_count // blocks + 1))
            for label, flags in (("fill", 0), ("q_s", nwb.WANT_COUNT)):
                plans = [nwb.Plan(n, n, flags | nwb.QUEUE, device=local) for _ in range(nq)]
                for pl in plans:
                    pl.upload(t, s)
                    pl.run(*mkd)
                torch.cuda.synchronize()
                kq = 6 * nq
                t0 = time.perf_counter()
                for i in range(kq):
                    plans[i % nq].run(*mkd)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                okq = True
                for pl in plans:
                    sm = pl.summary()
                    okq = okq and (sm.opt_score, sm.branch_count) == (g["final_score"], g["branch_count"]) \
                        and pl.arrow_digest() == int(g["arrow_digest"], 16) and (not flags or sm.count == g["count_u64"])
                    pl.close()
                rec[label]["queue"] = {"plans": nq, "fills": kq, "ms_per_table": dt / kq * 1e3, "gcups": n * n * kq / dt / 1e9,
                                       "golden_ok": bool(okq)}
                rec[label]["golden_ok"] = bool(rec[label]["golden_ok"] and okq)
        # every count of the dense sweep's last row and column (the final count is 0 mod 2^64 on these inputs)
        plan = nwb.Plan(n, n, nwb.WANT_COUNT_DIGEST, device=local)
        plan.upload(t, s)
        plan.run(*mkd)
        sm = plan.summary()
        okd = (sm.lastrow_count_digest, sm.lastcol_count_digest, sm.count) == \
              (int(g["lastrow_count_digest"], 16), int(g["lastcol_count_digest"], 16), g["count_u64"])
        rec["dense_count_sweep"] = {"kernel_ms_fill_plus_dense_count": plan.kernel_ms(), "golden_ok": bool(okd),
                                    "golden": "digests of cnt(i,B) for all i and cnt(A,j) for all j vs the oracle"}
        plan.close()
        rec["ratio_q_s_over_fill"] = rec["q_s"]["kernel_ms"] / rec["fill"]["kernel_ms"]
        out[name] = rec
        return all(v.get("golden_ok", True) for v in rec.values() if isinstance(v, dict))

    ok_all = one("config2_dna_10k", "config2_dna_10k", 0x5EED0002, 10000, nwb.DNA, (1, 1, 1)) and ok_all
    ok_all = one("config5_protein_30k", "config5_protein_30k", 0x5EED0005, 30000, nwb.PROTEIN, (2, 1, 2)) and ok_all
    ok_all = one("config3_dna_100k", "config3_dna_100k", SEED, A, nwb.DNA, (1, 1, 1)) and ok_all
    nwb.cache_clear()
    return out, ok_all


def run_ours(args) -> None:
    # keep stdout clean for the ONE JSON line: libraries (NCCL prints its version there) go to stderr
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import nw_b200 as nwb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available() or nwb.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the fill has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gold = goldens()
    g3 = gold["config3_dna_100k"]
    t, s = nwb.generate_pair(SEED, A, B)  # the package's own SURVEY 8d generator (checked against the oracle's in tests/)
    # One step = one fill of the whole 100k x 100k table.  The K steps are taken as a QUEUE of fills: NQ plans (device
    # workspaces, each with its own arrow table and stream) take the steps round robin, so consecutive fills overlap -- on one
    # GPU the next fill's blocks move onto SMs as this fill's leave them (NWB_QUEUE: ticketed blocks, no co-residency needed);
    # in a strip group (world > 1, every table is cut into column strips over ALL the GPUs) rank r also starts fill e + 1
    # while the ranks to its right are still on fill e (nwb_plan_run_pipelined: double-buffered inboxes, no barrier or
    # inbox reset between steps).  The timed region is bracketed by barriers as the contract says; the time ONE fill takes
    # from launch to completion is measured separately (`latency`).
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    NQ = args.queue if args.queue > 0 else queue_depth(A, world, sms)
    flags = nwb.QUEUE
    plans = [nwb.Plan(A, B, flags, device=local, strip_rank=rank, strip_world=world) for _ in range(NQ)]
    if world > 1:
        blobs = [None] * world
        dist.all_gather_object(blobs, [pl.ipc_export() for pl in plans])
        if rank + 1 < world:
            for q, pl in enumerate(plans):
                pl.ipc_attach_right(blobs[rank + 1][q])
    for pl in plans:
        pl.upload(t, s)
    # non-default torch streams: the C ABI treats a NULL stream as "the plan's own", and torch.cuda.Event wants a torch stream
    tstreams = [torch.cuda.Stream(device=local) for _ in range(NQ)]
    torch.cuda.set_stream(tstreams[0])
    nstep = [0]

    def step():
        q = nstep[0] % NQ
        nstep[0] += 1
        if world > 1:
            plans[q].run_pipelined(M_, K_, D_, tstreams[q].cuda_stream)
        else:
            plans[q].run(M_, K_, D_, tstreams[q].cuda_stream)
        return q

    for _ in range(NQ):              # set-up, not a step: every plan's first fill (first touch of its 5 GB table, module load)
        step()
    barrier()
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    # ---- device-resident timing: K steps, CUDA events on the launching streams
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev_start, ev_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev_start.record(tstreams[0])                     # every stream is idle here
    host_t, hbase = [], time.perf_counter()
    for i in range(args.steps):
        q = nstep[0] % NQ
        h0 = time.perf_counter()
        if args.pace and i >= NQ:
            tstreams[q].synchronize()                # host pacing: at most NQ fills launched ahead (what the e2e loop does by reading results)
        h1 = time.perf_counter()
        ev[i][0].record(tstreams[q])
        step()
        ev[i][1].record(tstreams[q])
        host_t.append((round((h0 - hbase) * 1e3, 3), round((h1 - hbase) * 1e3, 3), round((time.perf_counter() - hbase) * 1e3, 3)))
    for q in range(1, NQ):
        tstreams[0].wait_stream(tstreams[q])
    ev_end.record(tstreams[0])                       # ... and the last of the K fills is done here
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = ev_start.elapsed_time(ev_end)
    if args.trace:
        # diagnostics: when each fill started / ended on each rank, in ms after the start of the timed region
        tl = [(round(ev_start.elapsed_time(a), 3), round(ev_start.elapsed_time(b), 3)) for a, b in ev]
        if world > 1:
            allt = [None] * world
            dist.all_gather_object(allt, [tl, host_t])
        else:
            allt = [[tl, host_t]]
        if rank == 0:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            json.dump({"n_gpus": world, "plans": NQ, "steps": args.steps, "by_rank_device_start_end_and_host_sync_begin_end_launched": allt},
                      open(os.path.join(ROOT, "gpurun_out", f"queue_trace_n{world}.json"), "w"))
    if world > 1:
        tt = torch.tensor([total_ms], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    # latency of one fill, nothing else on the GPUs: every rank starts together, the last rank's end counts
    lat = []
    for _ in range(3):
        barrier()
        q = nstep[0] % NQ
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(tstreams[q])
        step()
        b.record(tstreams[q])
        barrier()
        lat.append(a.elapsed_time(b))
    fill_latency_ms = sum(lat) / len(lat)
    if world > 1:
        tt = torch.tensor([fill_latency_ms], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        fill_latency_ms = float(tt.item())
    kernel_name = plans[0].kernel_name()
    launches_per_step = sum(pl.launches() for pl in plans) // nstep[0]
    # every plan's table (this rank's share of it), digested on the device
    shares = []
    for pl in plans:
        sm = pl.summary()
        shares.append((sm.partial_r, sm.branch_count, pl.arrow_digest(), sm.kernel_kind, sm.opt_score))

    # ---- end to end with HOST strings: per step the H2D copy of the two strings, the fill, and the D2H read of the step's
    # summary block -- which is read when the plan comes round again (NQ steps later), i.e. the host keeps NQ fills in flight
    barrier()
    in_flight = [False] * NQ
    e2e_shares = set()

    def collect(q):
        sm = plans[q].summary()          # D2H of that step's result (waits for this rank's share of that fill only)
        e2e_shares.add((sm.partial_r, sm.branch_count, sm.opt_score))
        in_flight[q] = False

    t0 = time.perf_counter()
    for _ in range(args.steps):
        q = nstep[0] % NQ
        if in_flight[q]:
            collect(q)
        plans[q].upload(t, s)            # H2D of this step's inputs
        step()
        in_flight[q] = True
    for q in range(NQ):
        if in_flight[q]:
            collect(q)
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_steps_agree = len(e2e_shares) == 1 and (shares[0][0], shares[0][1], shares[0][4]) in e2e_shares
    e2e_call = (f"per step: nwb_plan_upload + nwb_plan_run{'_pipelined' if world > 1 else ''} on plan (step mod {NQ}), nwb_plan_summary "
                f"of that plan's previous fill when it comes round again (the host keeps {NQ} fills in flight)")
    if world > 1:
        tt = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    # the one-call path a CLI run takes: nwb_fill(), blocking, host strings in, table handle out (no queue)
    blocking = None
    if world == 1:
        for pl in plans[1:]:
            pl.close()
        plans = plans[:1]
        nwb.fill(t, s, M_, K_, D_, 0, device=local).close()          # first call creates the cached workspace
        t1 = time.perf_counter()
        for _ in range(args.steps):
            tab = nwb.fill(t, s, M_, K_, D_, 0, device=local)        # H2D + fill + D2H summary, one blocking call
            e2e_score = tab.opt_score
            blocking_kernel_ms = tab.kernel_ms
            tab.close()
        blocking_s = time.perf_counter() - t1
        nwb.cache_clear()
        blocking = {"call": "nwb_fill() (host strings -> table handle; workspace cached between calls), one fill at a time",
                    "ms_per_fill": blocking_s / args.steps * 1e3, "gcups": A * B * args.steps / blocking_s / 1e9,
                    "kernel_ms": blocking_kernel_ms, "score_ok": e2e_score == g3["final_score"]}
    clocks = sampler.stop() if rank == 0 else None
    # golden check of EVERY plan's table: a strip group's score is the sum of the ranks' bottom-row shares (include/nwb.h
    # nwb_summary); branch counts and digests add up (the digest mod 2^64 as two 32-bit halves: no int64 overflow)
    golden_ok = True
    want = (g3["final_score"], g3["branch_count"], int(g3["arrow_digest"], 16))
    for (partial_r, branches, dig, kind, score) in shares:
        if world > 1:
            tt = torch.tensor([partial_r, branches, dig & 0xFFFFFFFF, dig >> 32], device="cuda", dtype=torch.int64)
            dist.all_reduce(tt, op=dist.ReduceOp.SUM)
            if kind == 1:
                score = nwb.strip_group_score(int(tt[0].item()), A, B, D_)
            branches = int(tt[1].item()) & 0xFFFFFFFF
            dig = (int(tt[2].item()) + (int(tt[3].item()) << 32)) & M64
        golden_ok = golden_ok and (score, branches, dig) == want
        opt_score, branch_total, digest, kernel_kind = score, branches, dig, kind
    tt = torch.tensor([1 if e2e_steps_agree else 0], device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MIN)
    golden_ok = golden_ok and int(tt.item()) == 1 and (blocking is None or blocking["score_ok"])
    for pl in plans:
        pl.close()

    extras, extras_ok = (None, True)
    if not args.no_extras:
        extras, extras_ok = run_extras(nwb, torch, dist, world, rank, local, barrier, gold)

    if rank == 0:
        cells = A * B
        ms_per_step = total_ms / args.steps
        gcups = cells / (ms_per_step * 1e-3) / 1e9
        e2e_gcups = cells * args.steps / e2e_s / 1e9
        # INT/DPX issue roofline: P = N_SM * R * f
        r_per_clk, r_gops = nwb.measure_int_issue(1, local)
        r16, _ = nwb.measure_int_issue(2, local)
        rmix, _ = nwb.measure_int_issue(3, local)
        props = torch.cuda.get_device_properties(local)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
        f_mhz = float(peaks.get("sm_max_mhz", SM_MAX_MHZ_FALLBACK))
        peak_ops = props.multi_processor_count * r_per_clk * f_mhz * 1e6          # thread-results/s at max clock
        k_ms = ms_per_step  # the fills of a queue overlap: one table's share of the device time, not the span of one launch
        achieved_ops = cells * OPS_PER_CELL / (k_ms * 1e-3)
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        alg_bytes = cells * 0.5 + A + B
        # dram bytes of the fill kernel: NOT measured in this run (ncu cannot run inside a timed bench); taken from the
        # committed ncu --set full capture of the same kernel and command
        traffic, traffic_src = None, None
        for cand in ("r02_fill_hx_queue_ncu_summary.json", "r02_fill_hx_ncu_summary.json", "r01_fill_hx_ncu_summary.json"):
            tpath = os.path.join(ROOT, "profiles", cand)
            if kernel_name == "nwb_fill_hx_kernel" and world == 1 and os.path.exists(tpath):
                try:
                    traffic = json.load(open(tpath)).get("dram_bytes_total")
                    traffic_src = f"profiles/{cand} (ncu --set full, one launch of the same kernel, per launch = per table; not measured in this run)"
                except Exception:
                    traffic = None
                if traffic is not None:
                    break
        # a ceiling that cannot be exceeded: the SMs' issue slots (4 schedulers per SM, one warp instruction per clock each);
        # warp instructions per table from the committed ncu capture of the kernel (not measured in this run)
        issue = None
        try:
            wi = json.load(open(os.path.join(ROOT, "profiles", "r02_fill_hx_queue_ncu_summary.json"))).get("smsp__inst_executed.sum")
            if wi and kernel_name == "nwb_fill_hx_kernel":
                peak_slots = props.multi_processor_count * 4 * f_mhz * 1e6 * world
                issue = {"warp_instructions_per_table": wi, "achieved_per_s": wi / (ms_per_step * 1e-3), "peak_per_s": peak_slots,
                         "frac": wi / (ms_per_step * 1e-3) / peak_slots,
                         "frac_one_fill_alone": wi / (fill_latency_ms * 1e-3) / peak_slots,
                         "source": "smsp__inst_executed.sum of profiles/r02_fill_hx_queue_ncu_summary.json (sweeping + flush warps, "
                                   "flag polls included) / device time per table; peak = SMs x 4 schedulers x sm_max_mhz x GPUs"}
        except Exception:
            issue = None
        line = {
            "metric": METRIC, "value": gcups, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u16x2" if kernel_kind == 1 else "int32",
            "data": "synthetic", "config": workload_config(world),
            "roofline": {"bound": "int_issue", "achieved": achieved_ops / 1e12, "peak": peak_ops / 1e12,
                         "unit": "Tops/s (algorithmic INT ops, 10 per cell)", "frac": achieved_ops / peak_ops,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": f"measured on this GPU: VIMNMX3 {r_per_clk:.1f} thread-results/clk/SM "
                                        f"(VIMNMX3.U16x2 {r16:.1f} instr/clk/SM, VIMNMX3+IMAD {rmix:.1f}) x "
                                        f"{props.multi_processor_count} SMs x {f_mhz:.0f} MHz "
                                        f"({'MEASURED_PEAKS.json' if peaks else 'fallback'} sm_max_mhz)",
                         "kernel": kernel_name,
                         "kernel_ms": k_ms,
                         "kernel_ms_is": "device time per table over the timed region (the launches of consecutive fills overlap); one "
                                         "fill alone: latency.fill_latency_ms, frac_one_fill_alone",
                         "frac_one_fill_alone": cells * OPS_PER_CELL / (fill_latency_ms * 1e-3) / peak_ops,
                         "note": "10 ops per cell is SURVEY 8d's counting convention, not a ceiling: one DPX instruction updates two cells; "
                                 "issue_slots is the ceiling that cannot be exceeded",
                         "issue_slots": issue},
            "roofline_hbm": {"bound": "hbm", "achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak, "traffic": traffic,
                             "algorithmic_bytes": alg_bytes,
                             "note": "secondary: 0.5 B/cell arrow write-back + strings; not the binding roofline"},
            "e2e": {"value": e2e_gcups, "unit": "GCUPS", "h2d_bytes_per_step": A + B, "d2h_bytes_per_step": 72,
                    "call": e2e_call},
            "gpu_launches": args.steps * world * launches_per_step,
            "launches_per_step": {"per_rank": launches_per_step,
                                  "kernels": ("nwb_pk_prep_side_kernel, " + kernel_name + ", nwb_pk_stream_sum_kernel"
                                              + (", nwb_inbox_gate_kernel, nwb_inbox_ack_kernel" if world > 1 else ""))
                                  if kernel_kind == 1 else kernel_name},
            "clocks": clocks,
            "golden_ok": bool(golden_ok and extras_ok),
            "result": {"opt_score": opt_score, "branch_count": branch_total, "arrow_digest": f"{digest:016x}",
                       "kernel_kind": kernel_kind, "headline_golden_ok": bool(golden_ok),
                       "golden": "tests/golden/golden_big.json config3_dna_100k: score, branch count and the digest of the whole "
                                 "arrow table (every cell, summed over the ranks)"},
            "step_ms": [round(x, 3) for x in step_ms],
        }
        line["queue"] = {
            "plans_in_flight": NQ,
            "what": f"the K timed steps are a queue of fills taken round robin by {NQ} plans (workspaces with their own arrow table and "
                    "stream) per GPU: consecutive fills overlap on the device (NWB_QUEUE: ticketed blocks sweeping adjacent strips, the "
                    "next fill's blocks move onto SMs as this one's leave)"
                    + ("; every table is cut into column strips over all the GPUs and rank r starts fill e + 1 while the ranks to its "
                       "right are still on fill e (nwb_plan_run_pipelined)" if world > 1 else "")
                    + "; every plan's first fill runs once as set-up before the W warm-up steps; no barrier between steps, barriers on both sides of the timed region; value = K tables / (first launch .. last "
                      "fill done, max over ranks); every plan's table is checked against the goldens",
            "step_ms_is": "rank 0: launch-to-completion span of each fill's kernels (they overlap)",
        }
        line["latency"] = {
            "fill_latency_ms": fill_latency_ms,
            "fill_latency_gcups": cells / (fill_latency_ms * 1e-3) / 1e9,
            "how": "one fill in queue mode with nothing else on the GPUs, barrier on both sides, max over ranks, mean of 3",
            "blocking_call": blocking,
        }
        if extras is not None:
            line["extras"] = extras
        if world == 1 and not args.no_cpu:
            # the CPU legs run in a process of their own (no CUDA context: they fork worker processes)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-legs"], capture_output=True, text=True)
            try:
                legs = json.loads(r.stdout.strip().splitlines()[-1])
                line["cpu_baseline"] = legs["cpu_baseline"]
                line["cpu_baselines_other_configs"] = legs["others"]
            except Exception as exc:  # never lose the GPU line over the baseline
                line["cpu_baseline"] = {"unavailable": f"cpu legs failed: {exc}; stderr tail: {r.stderr[-300:]}"}
        sys.stdout.flush()
        os.write(saved_stdout, (json.dumps(line) + "\n").encode())
    ok = torch.tensor([1 if (golden_ok and extras_ok) else 0], device="cuda")
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        dist.barrier()
        dist.destroy_process_group()
    if int(ok.item()) != 1:
        sys.stderr.write("bench.py: GOLDEN MISMATCH -- see golden_ok / extras[*].golden_ok in the JSON line\n")
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--queue", type=int, default=0, help="plans (fills in flight) per GPU; 0 = as many as fill the SMs")
    ap.add_argument("--pace", type=int, default=1, help="1: the host waits for a plan's previous fill before it launches the next one on it")
    ap.add_argument("--trace", action="store_true", help="dump each fill's start / end per rank to gpurun_out/queue_trace_nN.json")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary configs (2, 4, 5)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline legs")
    ap.add_argument("--cpu-legs", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_legs:
        import nw_b200 as nwb
        import oracle
        print(json.dumps({"cpu_baseline": cpu_config3_sample(oracle, nwb), "others": cpu_other_configs(oracle, nwb)}))
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
