"""CPU tests of the N > 1 paths with one process per rank over torch.distributed (gloo, world_size 2 and 3).

What runs on the GPU box as `torchrun ... bench.py --gpus N` is, per rank: ask the C ABI for this rank's share
(column strips of one long pair / a contiguous range of a batch), run the kernels on that share, and combine
the ranks' 32-byte summaries (score = sum of partial_r - d*(A+B), branch counts add).  The only data that
crosses ranks on the data path is the boundary stream of a rank's last strip, written straight into the right
neighbour's inbox (peer HBM over NVLink on the box).  Here each rank is a CPU process, the kernels are the
real sources under the test-only SIMT emulator, and the inbox travels as a gloo point-to-point message; no
collective touches the data path.  The partitions come from libnwb.so's host-only helpers,
i.e. the same code nwb_plan_create() uses."""
import os
import socket
import traceback

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, args, errq):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        for p in (root, os.path.join(root, "tests")):
            if p not in sys.path:
                sys.path.insert(0, p)
        fn(rank, world, *args)
        dist.barrier()
        dist.destroy_process_group()
    except Exception:  # noqa: BLE001 -- report to the parent and fail there
        errq.put((rank, traceback.format_exc()))
        raise


def _run(world, fn, *args):
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn, args, errq)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    errs = []
    while not errq.empty():
        errs.append(errq.get())
    for p in procs:
        if p.is_alive():
            p.kill()
            errs.append((-1, "worker timed out"))
    assert not errs, "\n".join(f"rank {r}:\n{t}" for r, t in errs)
    assert all(p.exitcode == 0 for p in procs)


# ----------------------------------------------------------------------------- column strips of one pair
def _strips_rank(rank, world, seed, a, b, mkd, hx):
    import emu
    import nw_b200 as nwb
    import oracle
    m, k, d = mkd
    t, s = oracle.generate_pair(seed, a, b)
    # the boundary stream of my left neighbour's last strip: a point-to-point message, not a collective
    inbox = None
    c0, c1 = nwb.strip_partition(a, rank, world)
    has_strips = c1 > c0
    left = [r for r in range(rank) if nwb.strip_partition(a, r, world)[1] > nwb.strip_partition(a, r, world)[0]]
    if has_strips and left:
        buf = torch.zeros(emu.bpitch_pk(a, b), dtype=torch.int32)
        dist.recv(buf, src=left[-1])
        inbox = buf.numpy().view(np.uint32)
    r = emu.fill_pk_rank(t, s, m, k, d, rank=rank, world=world, inbox=inbox, hx=hx)
    assert (min(r["strip_begin"] * 256, a), min(r["strip_end"] * 256, a)) == (c0, c1)  # ABI helper == kernel-side rule
    right = [q for q in range(rank + 1, world) if nwb.strip_partition(a, q, world)[1] > nwb.strip_partition(a, q, world)[0]]
    if has_strips and right:
        dist.send(torch.from_numpy(r["outbox"].view(np.int32).copy()), dst=right[0])
    # combine the summaries the way bench.py / nwb_fill_on() do
    tot = torch.tensor([r["partial_r"], r["branch_count"]], dtype=torch.int64)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    score = nwb.strip_group_score(int(tot[0]), a, b, d)
    branches = int(tot[1]) & 0xFFFFFFFF
    # each rank checks its own columns of the arrow table against the oracle
    o = oracle.fill(t, s, m, k, d, want_codes=True)
    mine = emu.unpack_arrows(r["arrows"], a)[:, c0:c1] & 7
    assert np.array_equal(mine, o.codes[1:, 1 + c0:1 + c1] & 7), f"rank {rank}: arrows differ in columns [{c0},{c1})"
    assert score == o.final_score, (score, o.final_score)
    assert branches == o.branch_count, (branches, o.branch_count)
    # the ranks' column ranges tile [0, A) in order
    spans = [None] * world
    dist.all_gather_object(spans, (c0, c1))
    assert spans[0][0] == 0 and spans[-1][1] == a
    assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


@pytest.mark.parametrize("hx", [2, 1, 0], ids=["hy", "hx", "pk"])
def test_column_strips_two_ranks(hx):
    # 5 strips: ranks own 3 + 2; DNA 1/1/1
    _run(2, _strips_rank, 0x5EED0A01, 1200, 90, (1, 1, 1), hx)


def test_column_strips_two_ranks_protein_scheme():
    _run(2, _strips_rank, 0x5EED0A05, 700, 140, (2, 1, 2), 2)


def test_column_strips_three_ranks_uneven():
    # 4 strips over 3 ranks: 2 + 2 + 0 -- the last rank owns nothing and must not disturb the sums
    _run(3, _strips_rank, 0x5EED0A03, 1000, 70, (1, 1, 1), 2)


# ----------------------------------------------------------------------------- a queue of fills through a strip group
def _pipelined_rank(rank, world, seeds, a, b, mkd):
    """The protocol of nwb_plan_run_pipelined() with one process per rank: fill e streams into copy e & 1 of the right
    neighbour's inbox; the neighbour clears a copy when its own fill e is done and acknowledges e + 1; before fill e + 2
    a rank waits for that acknowledgement.  No barrier between the fills: rank 0 is on fill e + 1 while rank 1 sweeps
    fill e.  The sweeps are the queue-mode hx kernel (ticketed blocks of adjacent strips) under the emulator."""
    import emu
    import nw_b200 as nwb
    import oracle
    m, k, d = mkd
    c0, c1 = nwb.strip_partition(a, rank, world)
    bp = emu.bpitch_pk(a, b)
    copies = [torch.zeros(bp, dtype=torch.int32), torch.zeros(bp, dtype=torch.int32)]
    acks, mine_sum, want = [], [], []
    acked = 0          # fills the right neighbour has acknowledged
    for e, seed in enumerate(seeds):
        t, s = oracle.generate_pair(seed, a, b)
        inbox = None
        if rank > 0:
            assert int(copies[e & 1].abs().sum()) == 0, "the inbox copy was not cleared after fill e - 2"
            dist.recv(copies[e & 1], src=rank - 1, tag=e)
            inbox = copies[e & 1].numpy().view(np.uint32)
        if rank + 1 < world:
            while acked < e - 1:          # copy e & 1 of the neighbour's inbox carried fill e - 2
                ack = torch.zeros(1, dtype=torch.int64)
                dist.recv(ack, src=rank + 1, tag=1000 + acked)
                assert int(ack) == acked + 1
                acked += 1
        r = emu.fill_pk_rank(t, s, m, k, d, rank=rank, world=world, inbox=inbox, hx=6)
        if rank + 1 < world:
            dist.send(torch.from_numpy(r["outbox"].view(np.int32).copy()), dst=rank + 1, tag=e)
        if rank > 0:
            copies[e & 1].zero_()
            acks.append(dist.isend(torch.tensor([e + 1], dtype=torch.int64), dst=rank - 1, tag=1000 + e))
        o = oracle.fill(t, s, m, k, d, want_codes=True)
        mine = emu.unpack_arrows(r["arrows"], a)[:, c0:c1] & 7
        assert np.array_equal(mine, o.codes[1:, 1 + c0:1 + c1] & 7), f"rank {rank}, fill {e}: arrows differ"
        mine_sum.append((r["partial_r"], r["branch_count"]))
        want.append((o.final_score, o.branch_count))
    if rank + 1 < world:                  # drain the acknowledgements still in flight
        while acked < len(seeds):
            ack = torch.zeros(1, dtype=torch.int64)
            dist.recv(ack, src=rank + 1, tag=1000 + acked)
            acked += 1
    for w in acks:
        w.wait()
    # the summaries are combined at the end (a collective of the CHECK; nothing synchronises the ranks between the fills)
    shares = [None] * world
    dist.all_gather_object(shares, mine_sum)
    for e in range(len(seeds)):
        assert nwb.strip_group_score(sum(x[e][0] for x in shares), a, b, d) == want[e][0], e
        assert sum(x[e][1] for x in shares) & 0xFFFFFFFF == want[e][1], e


def test_pipelined_queue_of_fills_two_ranks():
    # 5 strips (3 + 2), four different fills back to back
    _run(2, _pipelined_rank, [0x5EED0A21, 0x5EED0A23, 0x5EED0A25, 0x5EED0A27], 1200, 90, (1, 1, 1))


def test_pipelined_queue_of_fills_three_ranks():
    # 7 strips (3 + 3 + 1), protein scheme
    _run(3, _pipelined_rank, [0x5EED0A31, 0x5EED0A33, 0x5EED0A35], 1700, 70, (2, 1, 2))


# ----------------------------------------------------------------------------- batch of pairs, no communication
def _batch_rank(rank, world, n_pairs):
    import emu
    import nw_b200 as nwb
    import oracle
    first, count = nwb.batch_partition(n_pairs, rank, world)
    tops, sides = [], []
    for p in range(first, first + count):
        a = 40 + (p * 37) % 90
        b = 30 + (p * 53) % 80
        t, s = oracle.generate_pair(0x5EED4000 + 2 * p, a, b)
        tops.append(t)
        sides.append(s)
    r = emu.fill_batch(tops, sides, 1, 1, 1, grid=1) if count else dict(scores=[], branches=[], tables=[])
    for i in range(count):
        o = oracle.fill(tops[i], sides[i], 1, 1, 1, want_codes=True)
        assert r["scores"][i] == o.final_score
        assert r["branches"][i] == o.branch_count
        assert np.array_equal(emu.unpack_arrows(r["tables"][i], len(tops[i])) & 7, o.codes[1:, 1:] & 7)
    # shards are disjoint, contiguous and cover the batch; nothing but this bookkeeping crosses ranks
    shards = [None] * world
    dist.all_gather_object(shards, (first, count))
    assert shards[0][0] == 0 and sum(c for _, c in shards) == n_pairs
    assert all(shards[i][0] + shards[i][1] == shards[i + 1][0] for i in range(world - 1))
    assert max(c for _, c in shards) - min(c for _, c in shards) <= 1


def test_batch_shards_two_ranks():
    _run(2, _batch_rank, 7)


def test_partition_helpers_are_host_only():
    import nw_b200 as nwb
    assert nwb.strip_partition(100_000, 0, 8) == (0, 12544)       # 391 strips: 49 per rank
    assert nwb.strip_partition(100_000, 7, 8) == (87808, 100_000)
    assert nwb.strip_partition(0, 0, 1) == (0, 0)
    assert nwb.strip_partition(300, 1, 4) == (256, 300)
    assert nwb.strip_partition(300, 3, 4) == (300, 300)           # nothing left for the last ranks
    assert [nwb.batch_partition(1_000_000, r, 8) for r in (0, 7)] == [(0, 125_000), (875_000, 125_000)]
    assert [nwb.batch_partition(10, r, 4) for r in range(4)] == [(0, 3), (3, 3), (6, 2), (8, 2)]
    assert nwb.strip_group_score(211_389, 100_000, 100_000, 1) == 11_389
    with pytest.raises(nwb.NwbError):
        nwb.strip_partition(10, 2, 2)
