#!/usr/bin/env python
"""BASELINE.json configs[3]: LMPC Monte-Carlo rollouts sharded over the GPUs of one node, with the once-per-lap NCCL
all-gather of finished laps (pooled-safe-set exchange).

Every instance is an independent LMPC controller + vehicle (its own PID lap as initial safe set, Philox process noise),
advanced entirely on the device: K1 regression -> K2 selection -> QP -> shift -> addPoint -> dynModel per step.
Lap ends are handled per instance (device-side lap hand-over, host only keeps the lap-time lists).

  --mode independent   reference semantics: no collective anywhere (replicas).
  --mode pooled        once every instance of a rank is `--ship-after` steps into LMPC lap r+1, the rank packs lap r of each
                       instance (with the rows LMPC.addPoint has appended past the finish line so far: the selection window
                       of PC.py:492-495 plus the N-step look-ahead of the terminal guess needs about 20 of them) on the device,
                       all ranks all-gather the packed laps and
                       lap times over NCCL, rank them (stable argsort of LapTime, PC.py:395) and every instance receives
                       the `--share` globally fastest laps it does not own as additional stored laps (safe set + model).

  python benchmarks/rollout_mc.py --batch 8192 --laps 3            (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 benchmarks/rollout_mc.py --batch 8192
Prints one JSON line (rank 0): closed-loop steps/s over all GPUs (device events, max over ranks), lap-time statistics per
LMPC lap, all-gather time and bytes."""
import argparse
import json
import os
import sys
import time
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from racinglmpc_b200 import workloads, sharding, reference_params as rp      # noqa: E402
from racinglmpc_b200.controller import BatchedController                      # noqa: E402

FIRST_LMPC_LAP = 4      # laps 0..3 are the PID seed laps (main.py:102-110)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8192, help="instances per GPU")
    ap.add_argument("--laps", type=int, default=3, help="LMPC laps every instance has to finish")
    ap.add_argument("--max-steps", type=int, default=0, help="0 = 300 per lap")
    ap.add_argument("--mode", choices=["independent", "pooled"], default="pooled")
    ap.add_argument("--share", type=int, default=2, help="pooled mode: globally fastest laps handed to every instance per exchange")
    ap.add_argument("--tpad", type=int, default=288, help="rows per exchanged lap (lap + addPoint overrun)")
    ap.add_argument("--ship-after", type=int, default=40, help="pooled mode: steps into the next lap before a lap is shipped")
    ap.add_argument("--seed-laps", choices=["golden", "device"], default="device",
                    help="golden: every instance is seeded with the reference's seed-0 PID lap (host upload); device: every "
                         "instance drives its OWN 1000-step PID lap on the device (main.py:65-66, Philox noise) and is seeded from it")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, N = args.batch, 12
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))
    xP, uP = g["pid_x"], g["pid_u"]
    numSS_it, numSS_Points, _, _, Qts, par = rp.lmpc_params(N)
    c = BatchedController(par, B, workloads.track_seg_table(), rp.TRACK_LENGTH, trToUse=4, numSS_Points=numSS_Points,
                          numSS_it=numSS_it, QterminalSlack=Qts, device=local, Tmax=1280, ss_cap=7, model_cap=5)
    t0 = time.perf_counter()
    x0 = np.tile(np.array([0.5, 0, 0, 0, 0, 0.0]), (B, 1))
    pid_s = 0.0
    if args.seed_laps == "golden":
        for b in range(B):                   # main.py:102-110: four copies of the PID lap seed both stores
            for _ in range(4):
                c.model_add_trajectory(b, xP, uP)
            for _ in range(4):
                c.add_trajectory(b, xP, uP)
        c.set_state(xLin=np.tile(xP[1:N + 2], (B, 1, 1)), uLin=np.tile(uP[1:N + 1], (B, 1, 1)),
                    zt=np.tile(np.array([0.0, 0, 0, 0, 10.0, 0]), (B, 1)), OldInput=np.zeros((B, 2)),
                    timeStep=np.zeros(B, np.int32), has_pred=np.zeros(B, np.int32))
        c.enable_rollout(Tcl=512)
    else:
        c.enable_rollout(Tcl=1024)
        c.rollout_set_state(x0, x0)
        tp = time.perf_counter()
        for _ in range(1000):                # main.py:65-66: the PID lap is the full 100 s simulation
            c.rollout_pid_step(0.8, seed=4321 + rank)
        _, n = c.rollout_done()
        c.rollout_seed_from_record(n, copies=4)
        c.sync()
        pid_s = time.perf_counter() - tp
    setup_s = time.perf_counter() - t0
    c.rollout_set_state(x0, x0)
    stream = torch.cuda.ExternalStream(c.stream, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    laps_done = np.zeros(B, np.int64)                   # LMPC laps finished per instance
    lap_len = [[] for _ in range(args.laps)]            # lap_len[r] = lengths of everybody's r-th LMPC lap
    since_lap = np.zeros(B, np.int64)
    host_s, xchg_s, xchg_bytes, n_xchg, took_total = 0.0, 0.0, 0, 0, 0
    next_round = 1                                      # exchange round r ships LMPC lap r early in lap r+1
    rows = torch.zeros(B, args.tpad, 9, dtype=torch.float64, device=dev)
    lens = torch.zeros(B, dtype=torch.int32, device=dev)
    max_steps = args.max_steps or 300 * args.laps
    steps = 0
    l0 = c.kernel_launches
    e0.record(stream)
    while steps < max_steps and laps_done.min() < args.laps:
        c.rollout_step(seed=1234 + rank)
        steps += 1
        since_lap += 1
        done, n = c.rollout_done()            # 2 x 4 B per instance back to the host: the only per-step traffic
        if done.any():
            th = time.perf_counter()
            fin = c.rollout_finish_laps(done, n)
            for b in fin:
                if laps_done[b] < args.laps:
                    lap_len[laps_done[b]].append(int(n[b]))
            laps_done[fin] += 1
            since_lap[fin] = 0
            host_s += time.perf_counter() - th
        if args.mode == "pooled" and next_round < args.laps and laps_done.min() >= next_round and \
                since_lap[laps_done == next_round].min(initial=10 ** 9) >= args.ship_after:
            # ---- once-per-lap exchange: pack on the device, all-gather over NCCL, rank, hand out -----------------------
            tx = time.perf_counter()
            # lap numbers shift by the laps an instance imported earlier (they sit before its own latest lap)
            own = np.array([c.own_lap_number(b, FIRST_LMPC_LAP + next_round - 1) for b in range(B)])
            c.export_laps(own, args.tpad, rows, lens)
            times = torch.tensor([c.LapTime[b][own[b]] for b in range(B)], dtype=torch.int32, device=dev)
            c.sync()
            rows_all, lens_all = sharding.allgather_laps(rows, lens)
            times_all = sharding.allgather_vec(times)
            best = [int(i) for i in sharding.pooled_fastest(times_all, args.share + 1)]
            times_np = times_all.cpu().numpy()
            for j in range(args.share):                     # pass j hands every instance its j-th foreign lap
                src = np.full(B, -1, np.int32)
                lt = np.zeros(B, np.int64)
                for b in range(B):
                    cand = [gi for gi in best if gi != rank * B + b]
                    if j < len(cand):
                        src[b], lt[b] = cand[j], times_np[cand[j]]
                took_total += len(c.import_laps(src, lt, args.tpad, rows_all, lens_all))
            n_xchg += 1
            xchg_bytes = int(rows.numel() * 8 + lens.numel() * 4 + times.numel() * 4)
            torch.cuda.synchronize()
            xchg_s += time.perf_counter() - tx
            next_round += 1
    e1.record(stream)
    c.sync()
    launches = c.kernel_launches - l0
    ms = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
    steps_all = steps
    if world > 1:
        t = torch.tensor([steps], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        steps_all = int(t.item())
    else:
        steps_all = steps
    st = c.rollout_state()
    flags_or, unsolved = c.rollout_health()
    stats = []
    for r in range(args.laps):
        a = np.array(lap_len[r]) if lap_len[r] else np.array([0])
        stats.append({"lap": r + 1, "n": int(len(lap_len[r])), "mean": float(a.mean()), "min": int(a.min()), "max": int(a.max())})
    if rank == 0:
        print(json.dumps({"benchmark": "configs[3] LMPC Monte-Carlo rollouts", "mode": args.mode, "n_gpus": world, "batch_per_gpu": B,
                          "closed_loop_steps_rank0": steps, "ms_total": ms,
                          "controller_steps_per_s": B * steps_all / (ms * 1e-3),
                          "kernel_launches_rank0": int(launches), "lap_stats_rank0": stats,
                          "host_lap_bookkeeping_s": host_s, "setup_s": setup_s, "seed_laps": args.seed_laps, "device_pid_laps_s": pid_s,
                          "exchanges": n_xchg, "exchange_s_total": xchg_s, "allgather_bytes_per_rank": xchg_bytes,
                          "laps_handed_out_rank0": took_total, "share": args.share, "ship_after": args.ship_after,
                          "instances_with_flags_rank0": int((flags_or != 0).sum()), "flag_bits_rank0": int(np.bitwise_or.reduce(flags_or)),
                          "unsolved_steps_rank0": int(unsolved.sum()), "s_mean": float(st["x"][:, 4].mean())}))
    c.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
