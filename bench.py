#!/usr/bin/env python
"""bench.py -- GCUPS of the Needleman-Wunsch score-table fill on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json): config 3, one pair of 100,000 x 100,000 synthetic DNA
(SplitMix64 seeds 0x5EED0030/31), m=k=d=1 -- the configuration the target
("the 100k x 100k fill at >= 50% of the INT/DPX issue roofline") is quoted on.
A step = one fill of the whole table (scores in registers, every arrow set
written as 4-bit codes).  At N > 1 the table is cut into column strips across
the ranks (boundary columns stream rank-to-rank through CUDA-IPC peer memory
over NVLink), total work fixed -> "scaling": "strong".

value  = interior cells / device time, strings resident in HBM (CUDA events on
         the launching stream, max over ranks);
e2e    = the same through the public C-ABI plan calls with HOST strings:
         upload (H2D) + fill + summary (D2H) inside the timed region.
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


def workload_config(n_gpus: int) -> dict:
    return {"workload": "config3: single pair 100000x100000 synthetic DNA (seeds 0x5EED0030/31), m=1 k=1 d=1, "
                        "fill + 4-bit arrow table, no count",
            "top_len": A, "side_len": B, "m": M_, "k": K_, "d": D_, "cells_per_step": A * B,
            "parallelism": f"column strips x{n_gpus}" if n_gpus > 1 else "single GPU",
            "l2": "each step writes a 5.0 GB arrow table (>> 126 MB L2); no explicit flush needed"}


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


# --------------------------------------------------------------------------- reference arm
def run_reference(args) -> None:
    """The reference's own CPU fill (oracle/_ref, the unmodified sources compiled
    by oracle/Makefile), timed on this box's host cores on a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    n = 3000  # 9e6 cells per step: ~1.5 s at the reference's ~6 MCUPS, 1.2 GB of its 136 B/cell tables
    t, s = oracle.generate_pair(SEED, n, n)
    kind = "reference" if oracle.have_reference() else "port"
    ncpu = os.cpu_count() or 1

    def one(threads):
        if kind == "reference":
            r = oracle.reference_fill(t, s, M_, K_, D_, threads=threads, tflag=False)
            return r.fill_seconds
        t0 = time.perf_counter()
        oracle.fill(t, s, M_, K_, D_)
        return time.perf_counter() - t0

    # the reference's -p filler stops scaling at 2 threads (README:72-75); probe and keep the best
    cands = [1] if kind == "port" else sorted({1, 2, min(4, ncpu), ncpu})
    probe = {T: one(T) for T in cands}
    best_t = min(probe, key=probe.get)
    for _ in range(args.warmup):
        one(best_t)
    times = [one(best_t) for _ in range(args.steps)]
    sec = sum(times) / len(times)
    gcups = n * n / sec / 1e9
    cfg = workload_config(args.gpus)
    sample = (f"{n}x{n} prefix of the workload's strings per step (the reference needs 136 B/cell: the full "
              f"100k x 100k table would take 1.36 TB); fill only (compute_table_scores), threads probed {probe}")
    line = {"impl": "reference", "metric": METRIC, "value": gcups, "unit": "GCUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": best_t, "kind": kind, "sample": sample,
                             "host_cores": ncpu},
            "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------- our arm
def cpu_baseline_leg() -> dict:
    import oracle
    n = 8000  # 6.4e7 cells, 8.7 GB of the reference's 136 B/cell tables, ~3 s per run on the B200 host
    t, s = oracle.generate_pair(SEED, n, n)
    ncpu = os.cpu_count() or 1
    if oracle.have_reference():
        res = {}
        for T in sorted({1, 2, min(4, ncpu), ncpu}):
            res[T] = oracle.reference_fill(t, s, M_, K_, D_, threads=T, tflag=False).fill_seconds
        best = min(res, key=res.get)
        return {"value": n * n / res[best] / 1e9, "unit": "GCUPS", "cores": best, "kind": "reference",
                "host_cores": ncpu,
                "sample": f"{n}x{n} prefix of the workload, compute_table_scores() only, unmodified reference "
                          f"(oracle/_ref), seconds by -p threads: { {k: round(v, 2) for k, v in res.items()} }"}
    t0 = time.perf_counter()
    oracle.fill(t, s, M_, K_, D_)
    dt = time.perf_counter() - t0
    return {"value": n * n / dt / 1e9, "unit": "GCUPS", "cores": 1, "kind": "port", "host_cores": ncpu,
            "sample": f"{n}x{n} prefix of the workload, oracle/nw_oracle.c (scalar C)"}


def run_extras(nwb, oracle, torch, dist, world, rank, local, barrier) -> dict | None:
    """Secondary measurements (not the headline): the other BASELINE configs.
    config 4 runs on every rank (its own shard of 125,000 pairs = 1M / 8, no communication);
    configs 2, 5 and config 3 with the fused count run on rank 0 only."""
    out = {}
    # ---- config 4: batch of 256 x 256 DNA pairs, pair p seeded 0x5EED4000 + 2p, sharded by rank
    per = 125_000
    first, cnt = nwb.batch_partition(per * world, rank, world)   # include/nwb.h: contiguous pair ranges, no communication
    assert cnt == per
    tcat, scat = bytearray(), bytearray()
    for p in range(first, first + per):
        tt, ss = oracle.generate_pair(0x5EED4000 + 2 * p, 256, 256)
        tcat += tt
        scat += ss
    import numpy as np
    off = np.arange(per + 1, dtype=np.int64) * 256
    bt = nwb.Batch.from_arrays(bytes(tcat), off, bytes(scat), off, M_, K_, D_, 0, device=local)
    for _ in range(2):
        bt.run()
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream().cuda_stream
    reps = 3
    e0.record()
    for _ in range(reps):
        bt.run(st)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / reps
    if world > 1:
        tt = torch.tensor([ms], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    bt.fetch()
    ok4 = True
    if rank == 0:
        ok4 = (bt.opt_score(0), bt.branch_count(0), bt.opt_score(1), bt.branch_count(1)) == (19, 23713, 29, 22912)
    kms = bt.kernel_ms()
    kname4 = bt.kernel_name()
    bt.close()
    # the same shard with the optimal-alignment count of every pair (second pass over the arrow codes)
    bc = nwb.Batch.from_arrays(bytes(tcat), off, bytes(scat), off, M_, K_, D_, nwb.WANT_COUNT, device=local)
    for _ in range(2):
        bc.run()
    torch.cuda.synchronize()
    bc.run(st)
    torch.cuda.synchronize()
    cms = bc.kernel_ms()
    bc.fetch()
    okc = True
    if rank == 0:
        okc = (bc.count(0), bc.count(1)) == (387701138034524160, 108460706365440)
    bc.close()
    if world > 1:
        tt = torch.tensor([cms], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        cms = float(tt.item())
    out["config4_batch_with_count"] = {"pairs_total": per * world, "ms_per_pass": cms,
                                       "gcups_total": per * world * 65536 / (cms * 1e-3) / 1e9,
                                       "kernels": "fill + nwb_batch_count_chain_kernel", "golden_counts_ok": bool(okc)}
    out["config4_batch"] = {"pairs_total": per * world, "pairs_per_gpu": per, "ms_per_pass": ms,
                            "gcups_total": per * world * 65536 / (ms * 1e-3) / 1e9, "fill_kernel_ms": kms, "kernel": kname4,
                            "includes": "fill + per-pair branch counter, strings and 4.1 GB of arrow tables resident",
                            "golden_pairs_ok": bool(ok4), "scaling": "weak (125,000 pairs per GPU; 8 GPUs = the 1M-pair config)"}
    if rank != 0:
        return None

    def one(name, seed, n, alpha, mkd, flags, golden):
        t, s = oracle.generate_pair(seed, n, n, alpha)
        plan = nwb.Plan(n, n, flags, device=local)
        plan.upload(t, s)
        best = None
        for _ in range(3):
            plan.run(*mkd)
            sm = plan.summary()
            k = plan.kernel_ms()
            best = k if best is None else min(best, k)
        got = (sm.opt_score, sm.branch_count, sm.count)
        out[name] = {"fill_kernel_ms": best, "gcups": n * n / (best * 1e-3) / 1e9, "kernel": plan.kernel_name(),
                     "result": {"opt_score": got[0], "branch_count": got[1], "count_u64": got[2]},
                     "golden_ok": bool(got[:len(golden)] == golden)}
        plan.close()

    one("config2_dna_10k_fill", 0x5EED0002, 10000, oracle.DNA, (1, 1, 1), 0, (1056, 34377799))
    one("config2_dna_10k_q_s_with_count", 0x5EED0002, 10000, oracle.DNA, (1, 1, 1), nwb.WANT_COUNT, (1056, 34377799, 0))
    one("config5_protein_30k_fill", 0x5EED0005, 30000, oracle.PROTEIN, (2, 1, 2), 0, (-16401, 520440751))
    one("config5_protein_30k_q_s_with_count", 0x5EED0005, 30000, oracle.PROTEIN, (2, 1, 2), nwb.WANT_COUNT,
        (-16401, 520440751, 0))
    one("config3_dna_100k_with_count", SEED, A, oracle.DNA, (1, 1, 1), nwb.WANT_COUNT, (11389, 3439940792, 0))
    return out


def run_ours(args) -> None:
    # keep stdout clean for the ONE JSON line: libraries (NCCL prints its version there) go to stderr
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import nw_b200 as nwb
    import oracle

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

    t, s = oracle.generate_pair(SEED, A, B)  # input generation only (SURVEY 8d generator)
    flags = 0
    plan = nwb.Plan(A, B, flags, device=local, strip_rank=rank, strip_world=world)
    if world > 1:
        blobs = [None] * world
        dist.all_gather_object(blobs, plan.ipc_export())
        if rank + 1 < world:
            plan.ipc_attach_right(blobs[rank + 1])
    plan.upload(t, s)
    # a non-default torch stream: the C ABI treats a NULL stream as "the plan's own",
    # and torch.cuda.Event only sees torch's current stream
    tstream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    def step():
        if world > 1:
            plan.reset_inbox(stream)
            barrier()
        plan.run(M_, K_, D_, stream)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    # ---- device-resident timing: K steps, CUDA events on the launching stream
    kernel_ms = []
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        if world > 1:
            plan.reset_inbox(stream)
            barrier()
        ev[i][0].record()
        plan.run(M_, K_, D_, stream)
        ev[i][1].record()
        if world > 1:
            torch.cuda.synchronize()
        kernel_ms.append(None)
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(step_ms)
    if world > 1:
        tt = torch.tensor([total_ms], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    last_kernel_ms = plan.kernel_ms()
    kernel_name = plan.kernel_name()
    summ = plan.summary()
    launches_per_step = plan.launches() // (max(args.warmup, 3) + args.steps)

    # ---- end to end through the public plan API with host strings
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        plan.upload(t, s)            # H2D of this step's inputs
        if world > 1:
            plan.reset_inbox(stream)
            barrier()
        plan.run(M_, K_, D_, stream)
        summ = plan.summary()        # D2H of the step's result (waits for the fill)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    clocks = sampler.stop() if rank == 0 else None
    opt_score = summ.opt_score
    branch_total = summ.branch_count
    if world > 1:
        # a strip group's score is the sum of the ranks' bottom-row shares (include/nwb.h nwb_summary)
        tt = torch.tensor([summ.partial_r, summ.branch_count], device="cuda", dtype=torch.int64)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        if summ.kernel_kind == 1:
            opt_score = nwb.strip_group_score(int(tt[0].item()), A, B, D_)
        branch_total = int(tt[1].item()) & 0xFFFFFFFF

    extras = run_extras(nwb, oracle, torch, dist, world, rank, local, barrier) if not args.no_extras else None

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
        k_ms = last_kernel_ms if world == 1 else ms_per_step
        achieved_ops = cells * OPS_PER_CELL / (k_ms * 1e-3)
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        alg_bytes = cells * 0.5 + A + B
        # dram__bytes_read.sum + dram__bytes_write.sum of the fill kernel from the committed ncu capture
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_fill_hx_ncu_summary.json" if kernel_name == "nwb_fill_hx_kernel"
                             else "r01_fill_pk_ncu_summary.json")
        if os.path.exists(tpath) and world == 1:
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_total")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": gcups, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u16x2" if summ.kernel_kind == 1 else "int32",
            "data": "synthetic", "config": workload_config(world),
            "roofline": {"bound": "int_issue", "achieved": achieved_ops / 1e12, "peak": peak_ops / 1e12,
                         "unit": "Tops/s (algorithmic INT ops, 10 per cell)", "frac": achieved_ops / peak_ops,
                         "traffic": traffic,
                         "peak_source": f"measured on this GPU: VIMNMX3 {r_per_clk:.1f} thread-results/clk/SM "
                                        f"(VIMNMX3.U16x2 {r16:.1f} instr/clk/SM, VIMNMX3+IMAD {rmix:.1f}) x "
                                        f"{props.multi_processor_count} SMs x {f_mhz:.0f} MHz "
                                        f"({'MEASURED_PEAKS.json' if peaks else 'fallback'} sm_max_mhz)",
                         "kernel": kernel_name,
                         "kernel_ms": k_ms},
            "roofline_hbm": {"bound": "hbm", "achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak, "traffic": traffic,
                             "algorithmic_bytes": alg_bytes,
                             "note": "secondary: 0.5 B/cell arrow write-back + strings; not the binding roofline"},
            "e2e": {"value": e2e_gcups, "unit": "GCUPS", "h2d_bytes_per_step": A + B, "d2h_bytes_per_step": 32},
            "gpu_launches": args.steps * world * launches_per_step,
            "launches_per_step": {"per_rank": launches_per_step,
                                  "kernels": ("nwb_pk_prep_side_kernel, " + kernel_name + ", nwb_pk_stream_sum_kernel")
                                  if summ.kernel_kind == 1 else kernel_name},
            "clocks": clocks,
            "result": {"opt_score": opt_score, "branch_count": branch_total, "kernel_kind": summ.kernel_kind,
                       "golden": "tests/golden/golden_big.json config3_dna_100k: score 11389, branches 3439940792"},
            "step_ms": [round(x, 3) for x in step_ms],
        }
        if extras is not None:
            line["extras"] = extras
        if world == 1:
            line["cpu_baseline"] = cpu_baseline_leg()
        sys.stdout.flush()
        os.write(saved_stdout, (json.dumps(line) + "\n").encode())
    plan.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary configs (2, 4, 5)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
