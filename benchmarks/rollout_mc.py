#!/usr/bin/env python
"""BASELINE.json configs[3]: LMPC Monte-Carlo rollouts sharded over the GPUs of one node, with the once-per-lap NCCL
all-gather of the finished laps (pooled-safe-set exchange).

Every instance is an independent LMPC controller + vehicle (seeded PID laps as initial safe set, Philox process noise),
advanced entirely on the device: K1 regression -> K2 selection -> QP -> shift -> addPoint -> dynModel per step.
Lap ends are handled per instance (device-side lap hand-over); after `--steps` closed-loop steps the laps driven so far
are packed on the device and all-gathered over NCCL, and every rank ranks the pooled laps by lap time.

  python benchmarks/rollout_mc.py --batch 8192 --steps 260            (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 benchmarks/rollout_mc.py --batch 8192
Prints one JSON line (rank 0): closed-loop steps/s over all GPUs (device events, max over ranks), laps finished,
lap-time statistics, all-gather time and bytes."""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8192, help="instances per GPU")
    ap.add_argument("--steps", type=int, default=260)
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
                          numSS_it=numSS_it, QterminalSlack=Qts, device=local, Tmax=1280, ss_cap=6, model_cap=5)
    t0 = time.perf_counter()
    for b in range(B):                       # main.py:102-110: four copies of the PID lap seed both stores
        for _ in range(4):
            c.model_add_trajectory(b, xP, uP)
        for _ in range(4):
            c.add_trajectory(b, xP, uP)
    c.set_state(xLin=np.tile(xP[1:N + 2], (B, 1, 1)), uLin=np.tile(uP[1:N + 1], (B, 1, 1)),
                zt=np.tile(np.array([0.0, 0, 0, 0, 10.0, 0]), (B, 1)), OldInput=np.zeros((B, 2)),
                timeStep=np.zeros(B, np.int32), has_pred=np.zeros(B, np.int32))
    setup_s = time.perf_counter() - t0
    c.enable_rollout(Tcl=512)
    x0 = np.tile(np.array([0.5, 0, 0, 0, 0, 0.0]), (B, 1))
    c.rollout_set_state(x0, x0)
    stream = torch.cuda.ExternalStream(c.stream, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    lap_times, host_s = [], 0.0
    l0 = c.kernel_launches
    e0.record(stream)
    for k in range(args.steps):
        c.rollout_step(seed=1234 + rank)
        done, n = c.rollout_done()            # 2 x 4 B per instance back to the host: the only per-step traffic
        if done.any():
            th = time.perf_counter()
            fin = c.rollout_finish_laps(done, n)
            lap_times += [int(n[b]) for b in fin]
            host_s += time.perf_counter() - th
    e1.record(stream)
    c.sync()
    launches = c.kernel_launches - l0
    ms = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
    # ---- once-per-lap exchange of the laps in flight / finished: pack on the device, all-gather over NCCL ----
    Tpad = 256
    rows = torch.zeros(B, Tpad, 8, dtype=torch.float64, device=dev)
    lens = torch.zeros(B, dtype=torch.int32, device=dev)
    c.rollout_export_laps(Tpad, rows, lens)
    c.sync()
    torch.cuda.synchronize()
    ta = time.perf_counter()
    rows_all, lens_all = sharding.allgather_laps(rows, lens) if world > 1 else (rows, lens)
    torch.cuda.synchronize()
    ag_s = time.perf_counter() - ta
    best = sharding.pooled_fastest(lens_all, 4)
    st = c.rollout_state()
    if rank == 0:
        lt = np.array(lap_times) if lap_times else np.array([0])
        print(json.dumps({"benchmark": "configs[3] LMPC Monte-Carlo rollouts", "n_gpus": world, "batch_per_gpu": B,
                          "closed_loop_steps": args.steps, "ms_total": ms, "controller_steps_per_s": B * world * args.steps / (ms * 1e-3),
                          "kernel_launches": int(launches), "laps_finished_rank0": int(len(lap_times)),
                          "lap_len_mean": float(lt.mean()), "lap_len_min": int(lt.min()), "lap_len_max": int(lt.max()),
                          "host_lap_bookkeeping_s": host_s, "setup_s": setup_s,
                          "allgather_ms": ag_s * 1e3, "allgather_bytes_per_rank": int(rows.numel() * 8 + lens.numel() * 4),
                          "pooled_fastest_lens": [int(lens_all[i]) for i in best], "s_mean": float(st["x"][:, 4].mean())}))
    c.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
