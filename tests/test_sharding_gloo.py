"""CPU tier: the N>1 host logic over the gloo backend, world_size = 2 (no GPU needed)."""
import os
import socket
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from racinglmpc_b200 import sharding


def test_shard_ranges_partition_the_batch():
    for B in (1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, Tmax = 6, 40
        lo, hi = sharding.shard_range(B, rank, world)
        rng = np.random.default_rng(100)                      # same stream on both ranks: global data, local slices
        Ts = rng.integers(5, Tmax, size=B)
        laps_x = [rng.normal(size=(T, 6)) for T in Ts]
        laps_u = [rng.normal(size=(T, 2)) for T in Ts]
        rows, lens = sharding.pack_laps(laps_x[lo:hi], laps_u[lo:hi], Tmax)
        rows_all, lens_all = sharding.allgather_laps(torch.from_numpy(rows), torch.from_numpy(lens))
        ref_rows, ref_lens = sharding.pack_laps(laps_x, laps_u, Tmax)
        ok = bool(np.array_equal(rows_all.numpy(), ref_rows) and np.array_equal(lens_all.numpy(), ref_lens))
        best = sharding.pooled_fastest(lens_all, 4)
        ok = ok and np.array_equal(best, np.argsort(Ts, kind="stable")[:4])
        m = sharding.max_over_ranks(10.0 + rank, torch.device("cpu"))
        ok = ok and (m == 10.0 + world - 1)
        # the pooled exchange of benchmarks/rollout_mc.py: 9-column packed laps (x | u | Qfun) + lap times; every rank derives the
        # same hand-out (the k globally fastest laps an instance does not own), ties to the lower global index
        rows9 = torch.from_numpy(np.concatenate([rows, np.zeros(rows.shape[:2] + (1,))], axis=2))
        times = torch.tensor([5, 3, 3] if rank == 0 else [3, 9, 2], dtype=torch.int32)
        r9_all, _ = sharding.allgather_laps(rows9, torch.from_numpy(lens))
        t_all = sharding.allgather_vec(times)
        ok = ok and tuple(r9_all.shape) == (B, Tmax, 9) and t_all.tolist() == [5, 3, 3, 3, 9, 2]
        best = [int(i) for i in sharding.pooled_fastest(t_all, 3)]
        ok = ok and best == [5, 1, 2]                           # lap times 2, 3, 3 -> global instances 5, 1, 2 (tie: lower index)
        mine = [[g for g in best if g != lo + b][:2] for b in range(hi - lo)]
        ok = ok and (mine == ([[5, 1], [5, 2], [5, 1]] if rank == 0 else [[5, 1], [5, 1], [1, 2]]))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_allgather_of_finished_laps_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
