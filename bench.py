#!/usr/bin/env python
"""bench.py — BASELINE.json metric: QP solves/sec (N=12, nx=6, nu=2).

Workload = BASELINE.json configs[1]: batch=4096 LTV-MPC QPs, N=12, per-instance fixed (A_k,B_k,C_k)
(126 variables / 174 constraint rows each in the reference's OSQP form), built by
racinglmpc_b200/workloads.py from the committed fixture.  One "step" = one pass of the hot
path (assemble-free QP solve + unpack, one kernel launch) over the whole batch.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Under torchrun (N > 1) every rank owns its own 4096-QP batch on its own GPU (weak scaling, no data-path
collective: instances are independent — SURVEY §8e); timing is CUDA events on the launching stream, max
over ranks.  `--impl reference` times the CPU oracle (oracle/osqp_port.c: the OSQP algorithm with the
reference's settings, cold start + polish, PC.py:259-283) on the host cores instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "QP solves/sec (N=12, nx=6, nu=2)"
UNIT = "solves/s"
BATCH = 4096
HORIZON = 12
# algorithmic HBM bytes per solve (per-instance A/B/C):
#   in  x0 6 + OldInput 2 + ABC 648 = 656 f64 ; out xPred 78 + uPred 24 + status/iters/3 resid ~4 = 106 f64
ALGO_BYTES_PER_SOLVE = (656 + 106) * 8
# fp64 flops of one interior-point iteration of this QP (counted from ftocp_pdip.cuh, N=12, M=0):
# Riccati factor 12*(2*(288+216+126+72+36+48)) + 4 vector sweeps 12*2*(2*48+30) + elementwise ~4k
FLOPS_PER_ITER = 12 * 2 * (288 + 216 + 126 + 72 + 36 + 48) + 4 * 12 * 2 * 126 + 4000


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU oracle arm (cpu_baseline leg and --impl reference).  The ONLY place bench.py touches oracle/.
# ----------------------------------------------------------------------------------------------
def oracle_problem_set(nsample, seed=1):
    """Assemble `nsample` of the workload's QPs in the reference's OSQP form (Python, untimed)."""
    from oracle import ftocp, osqp_port
    from racinglmpc_b200 import workloads
    x0, uold, abc = workloads.ltv_mpc_batch(nsample, N=HORIZON, seed=seed)
    par = ftocp.mpc_params(6, 2, HORIZON, 0.8)[1]
    par.timeVarying = True
    F, bb = ftocp.build_ineq(par)
    Ps, qs, As, ls, us = [], [], [], [], []
    for b in range(nsample):
        A = abc[b][:, 0:36].reshape(HORIZON, 6, 6)
        Bm = abc[b][:, 36:48].reshape(HORIZON, 6, 2)
        C = abc[b][:, 48:54]
        H, q = ftocp.build_cost(par, uold[b])
        G, E, L = ftocp.build_eq(par, list(A), list(Bm), list(C))
        P, q, Am, l, u = ftocp.osqp_form(H, q, F, bb, G, E @ x0[b] + L)
        Ps.append(P); qs.append(q); As.append(Am); ls.append(l); us.append(u)
    maskP = np.any(np.array([p != 0 for p in Ps]), axis=0)
    maskA = np.any(np.array([a != 0 for a in As]), axis=0)
    patP, patA = osqp_port.csc_pattern(maskP, maskA)
    Px = np.stack([osqp_port.gather_values(p, *patP) for p in Ps])
    Ax = np.stack([osqp_port.gather_values(a, *patA) for a in As])
    return patP, patA, Px, np.stack(qs), Ax, np.stack(ls), np.stack(us)


def oracle_time(prob, nthreads, repeats=1):
    from oracle import osqp_port
    patP, patA, Px, q, Ax, l, u = prob
    best = None
    infos = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        _, infos, _ = osqp_port.solve_batch(patP, patA, Px, q, Ax, l, u, nthreads=nthreads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    solved = sum(1 for i in infos if i["status"] == 1)
    return q.shape[0] / best, best, solved, float(np.mean([i["iters"] for i in infos]))


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def _affinity_and_quota():
    try:
        a = len(os.sched_getaffinity(0))
    except AttributeError:
        a = os.cpu_count() or 1
    q = None
    try:                                                    # cgroup v2, then v1
        qs, ps = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if qs != "max":
            q = max(1, int(np.ceil(int(qs) / int(ps))))
    except (OSError, ValueError):
        try:
            qv = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            pv = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if qv > 0:
                q = max(1, int(np.ceil(qv / pv)))
        except (OSError, ValueError):
            pass
    return a, q


def best_oracle_threads(prob):
    """The CPU arm gets its best configuration, found by measuring: the logical CPUs of the affinity mask, half and a quarter of
    them (one thread per physical core / per two), and the container's CPU quota and half of it when one is set.  Deliberately
    NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to every rank, which would silently make the reference arm
    single-core -- and a fixed guess is wrong either way (measured on gpurun boxes: 128 SMT threads 16 k solves/s, 64 threads
    150-175 k, a 32-CPU quota box 44 k at 16).  LMPC_BENCH_THREADS pins the count."""
    if os.environ.get("LMPC_BENCH_THREADS"):
        return max(1, int(os.environ["LMPC_BENCH_THREADS"]))
    a, q = _affinity_and_quota()
    cand = {a, max(1, a // 2), max(1, a // 4)}
    if q:
        cand |= {min(a, q), max(1, min(a, q) // 2)}
    best, best_v = 1, -1.0
    for n in sorted(cand, reverse=True):
        oracle_time(prob, n)
        v = max(oracle_time(prob, n)[0], oracle_time(prob, n)[0])
        if v > best_v:
            best, best_v = n, v
    return best


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from oracle import osqp_port
    nsample = 1024
    prob = oracle_problem_set(nsample)
    cores = best_oracle_threads(prob)
    for _ in range(args.warmup):
        oracle_time(prob, cores)
    t0 = time.perf_counter()
    solved = 0
    for _ in range(args.steps):
        _, _, s, it = oracle_time(prob, cores)
        solved += s
    dt = time.perf_counter() - t0
    val = nsample * args.steps / dt
    sample = "each step = first %d of the 4096 QPs (OSQP-algorithm C port, reference settings eps 1e-3 + polish, cold)" % nsample
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1] batch=4096 LTV-MPC QPs N=12 (bounded sample of %d per step)" % nsample,
                   "cpu": cpu_model(), "solved_fraction": solved / (nsample * args.steps), "mean_admm_iters": it},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from racinglmpc_b200 import BatchedFTOCP, workloads, reference_params as rp

    rank, world, local = dist_env()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, N = BATCH, HORIZON

    x0, uold, abc = workloads.ltv_mpc_batch(B, N=N, seed=1 + rank)
    solver = BatchedFTOCP(rp.mpc_params(N), batch=B, device=local)
    stream = torch.cuda.ExternalStream(solver.stream, device=dev)

    # ---- device-resident inputs / outputs (the `value` leg) ----
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_x0, d_u, d_abc = t(x0), t(uold), t(abc)
    d_xP = torch.zeros(B, N + 1, 6, dtype=torch.float64, device=dev)
    d_uP = torch.zeros(B, N, 2, dtype=torch.float64, device=dev)
    d_st = torch.zeros(B, dtype=torch.int32, device=dev)
    d_it = torch.zeros(B, dtype=torch.int32, device=dev)
    d_rs = torch.zeros(B, 3, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    torch.cuda.synchronize()

    def step_dev():
        solver.solve_dev(d_x0, d_u, d_abc, N * 54, 54, d_xP, d_uP, d_st, d_it, d_rs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()            # runs through warm-up and the device-timed region (under load)
    for _ in range(max(args.warmup, 3)):
        step_dev()
    solver.sync()
    t_w = time.perf_counter()      # extra untimed warm-up: >= 0.4 s of load so that clocks settle and get sampled
    while time.perf_counter() - t_w < 0.4:
        for _ in range(8):
            step_dev()
        solver.sync()
    launches0 = solver.kernel_launches
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with torch.cuda.stream(stream):
        for i in range(args.steps):
            flush.zero_()                      # L2 flush between timed iterations (same stream, untimed)
            ev[i][0].record(stream)
            step_dev()
            ev[i][1].record(stream)
    solver.sync()
    barrier()
    launches = solver.kernel_launches - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    status = d_st.cpu().numpy()
    iters = d_it.cpu().numpy()
    resid = d_rs.cpu().numpy()
    ok_frac = float(np.mean(status == 1))
    # nvidia-smi polling takes driver locks that stall cudaMemcpyAsync enqueues: the sampler covers the device-timed region
    # only and is stopped before the host-timed leg (measured: 2.6-3.4 M/s with it running, 3.9 M/s without)
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end leg: host (pinned) buffers through the public API, copies inside the timed region ----
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    h_x0, h_u, h_abc = pin(x0), pin(uold), pin(abc)
    # The public host API, used the way a caller streams batch after batch: two batches in flight (solve_async on buffer
    # sets 0/1, wait before a set is reused), so the H2D copy of step i+1 overlaps the solve of step i.  Every step's inputs
    # are copied from pinned host memory and every step's results are copied back to pinned host memory inside the timed region.
    outs = [{k: torch.from_numpy(v).pin_memory().numpy() for k, v in solver.alloc_outputs(False).items()} for _ in range(2)]
    for i in range(max(args.warmup, 3) + 20):   # untimed; long enough to bring the clocks back up after the pinning pause
        slot = i & 1
        if i >= 2:
            solver.wait(slot)
        solver.solve_async(slot, h_x0, h_u, h_abc, outs[slot])
    solver.wait(0); solver.wait(1)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        slot = i & 1
        if i >= 2:
            solver.wait(slot)                  # results of step i-2 are on the host before its buffer set is reused
        solver.solve_async(slot, h_x0, h_u, h_abc, outs[slot])
    solver.wait(0); solver.wait(1)
    torch.cuda.synchronize()
    out = outs[(args.steps - 1) & 1]
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    h2d = int(h_x0.nbytes + h_u.nbytes + h_abc.nbytes)
    d2h = int(sum(out[k].nbytes for k in ("xPred", "uPred", "slack", "status", "iters", "resid")))

    if rank == 0:
        value = B * world * args.steps / (total_ms * 1e-3)
        ms_per_step = total_ms / args.steps
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        kern_ms = float(np.mean(step_ms))                 # one kernel per step: launch duration == step duration
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["ftocp_kernel<12,0,2,4>"]["dram_bytes_per_launch"]
        except Exception:
            pass
        achieved = B * ALGO_BYTES_PER_SOLVE / (kern_ms * 1e-3) / 1e9
        gflops = float(np.sum(iters + 1)) * FLOPS_PER_ITER / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: batch=4096 LTV-MPC QPs N=12 nx=6 nu=2, per-instance fixed A/B/C (126 vars / 174 rows in OSQP form)",
                       "batch_per_gpu": B, "horizon": N, "l2": "flushed between timed steps (256 MiB memset on the same stream)",
                       "e2e_mode": "public host API with two batches in flight (solve_async/wait, double-buffered device inputs), pinned host buffers; every step copies its inputs H2D and its results D2H inside the timed region",
                       "tolerance": "r_prim,r_dual <= 1e-9, gap <= 1e-11 (unscaled inf-norm)",
                       "solved_fraction": ok_frac, "ipm_iters_mean": float(iters.mean()), "ipm_iters_max": int(iters.max()),
                       "max_resid": float(resid.max())},
            "clocks": clocks,
            "e2e": {"value": B * world * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "ftocp_kernel<12,0,2,4>",
                         "note": "latency-bound fp64 kernel by construction (SURVEY §8d): HBM fraction is tiny; see fp64_gflops",
                         "fp64_gflops": gflops, "algorithmic_bytes_per_solve": ALGO_BYTES_PER_SOLVE},
        }
        if not args.no_cpu_baseline and world == 1:       # the CPU leg is measured at N = 1 only (rank 0 has the host to itself)
            nsample = 1024
            prob = oracle_problem_set(nsample)
            cores = best_oracle_threads(prob)
            oracle_time(prob, cores)
            v, dt, solved, it = oracle_time(prob, cores, repeats=3)
            v1, _, _, _ = oracle_time((prob[0], prob[1], prob[2][:128], prob[3][:128], prob[4][:128], prob[5][:128], prob[6][:128]), 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "first %d of the 4096 QPs, OSQP-algorithm C port with the reference's settings "
                                              "(eps 1e-3, polish, cold start per QP), best of 3; single-core %.0f solves/s; cpu %s"
                                              % (nsample, v1, cpu_model())}
        print(json.dumps(line))
    solver.close()
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
# configs[2]: full LMPC steps (K1 k-NN regression -> K2 safe-set selection -> QP -> shift), per-instance 5-lap safe sets
# ----------------------------------------------------------------------------------------------
def run_lmpc_steps(args):
    import torch
    import torch.distributed as dist
    from racinglmpc_b200 import workloads, reference_params as rp
    from racinglmpc_b200.controller import BatchedController
    rank, world, local = dist_env()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, N = args.batch, HORIZON
    seg = workloads.track_seg_table()      # Map.PointAndTangent[:,3:6] from the reference-pinned fixture
    data = workloads.lmpc_batch(B, seed=2 + rank)
    numSS_it, numSS_Points, _, _, Qts, par = rp.lmpc_params(N)
    c = BatchedController(par, B, seg, rp.TRACK_LENGTH, trToUse=5, numSS_Points=numSS_Points, numSS_it=numSS_it,
                          QterminalSlack=Qts, device=local, Tmax=1280, ss_cap=5, model_cap=5)
    workloads.restore_lmpc_batch(c, data)
    stream = torch.cuda.ExternalStream(c.stream, device=dev)
    d_x0 = torch.from_numpy(data["x0"]).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    state = dict(xLin=data["xLin"], uLin=data["uLin"], zt=data["zt"], OldInput=data["OldInput"],
                 timeStep=data["t"].astype(np.int32), has_pred=np.ones(B, np.int32), xPred=data["xPred"])

    def reset():
        c.set_state(**state)      # every timed step solves the same controller states (untimed host upload)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        reset(); c.step_dev(d_x0); c.sync()
    l0 = c.kernel_launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for i in range(args.steps):
        reset()
        with torch.cuda.stream(stream):
            flush.zero_()
            ev[i][0].record(stream)
            c.step_dev(d_x0)
            ev[i][1].record(stream)
        c.sync()
    launches = c.kernel_launches - l0
    torch.cuda.synchronize()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    tot = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    tot = float(tot.item())
    clocks = sampler.stop() if rank == 0 else None      # not during the host-timed leg (see run_gpu)
    # e2e: host x0 in, results out through the public API
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    h_x0 = pin(data["x0"])
    out = {k: pin(v) for k, v in c.alloc_step_outputs().items()}
    reset(); c.step(h_x0, out=out, want_ss=False)
    t_e2e = 0.0
    for _ in range(args.steps):
        reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        c.step(h_x0, out=out, want_ss=False)
        t_e2e += time.perf_counter() - t0
    ok = float(np.mean((out["status"] == 1) & (out["flags"] == 0)))
    if rank == 0:
        rows_model = sum(l[0].shape[0] for l in data["model_laps"][0])
        rows_ss = sum(l[0].shape[0] for l in data["ss_laps"][0])
        algo = rows_model * 64 + rows_ss * 48 + (656 + 106) * 8 + 5760
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        ms = tot / args.steps
        ach = B * algo / (ms * 1e-3) / 1e9
        line = {"metric": METRIC, "value": B * world * args.steps / (tot * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": "configs[2]: batch=%d full LMPC steps (k-NN LTV regression over a 5-lap store, 4-lap sampled safe set, "
                                       "180 vars / 229 rows QP), N=12" % B, "batch_per_gpu": B,
                           "l2": "flushed between timed steps (256 MiB memset on the same stream)", "solved_fraction": ok,
                           "ipm_iters_mean": float(out["iters"].mean()), "ipm_iters_max": int(out["iters"].max()),
                           "kernels_per_step": "knn_ltv_regress, ss_select, ftocp_kernel<12,48>, shift_state"},
                "clocks": clocks,
                "e2e": {"value": B * world * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h_x0.nbytes),
                        "d2h_bytes_per_step": int(sum(out[k].nbytes for k in ("xPred", "uPred", "lambd", "zt", "zt_u", "status", "iters", "resid", "flags")))},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                             "kernel": "whole step (4 kernels)", "algorithmic_bytes_per_solve": algo}}
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = lmpc_cpu_baseline(data)
        print(json.dumps(line))
    c.close()
    if world > 1:
        dist.destroy_process_group()


def lmpc_cpu_baseline(data, nsample=24):
    """The reference's own way: one Python controller, one QP at a time (PC.py:110-137), restated by the oracle and
    solved by the OSQP-algorithm port with the reference's settings."""
    from oracle import ftocp, ltv_model, osqp_port
    from oracle.track import TrackTable
    trk = TrackTable()
    numSS_it, numSS_Points, _, _, Qts, par = ftocp.lmpc_params(trk, HORIZON)
    par.timeVarying = True
    ctrls = []
    for b in range(nsample):
        pm = ltv_model.LocalLTVModel(6, 2, trk, 5)
        pm.xStored = [lx for lx, _ in data["model_laps"][b]]
        pm.uStored = [lu for _, lu in data["model_laps"][b]]
        pm.lapTime = [lx.shape[0] for lx in pm.xStored]
        lm = ftocp.OracleLMPC(numSS_Points, numSS_it, Qts, par, pm, qp=osqp_port.reference_qp)
        lm.SS = [s[0] for s in data["ss_laps"][b]]; lm.uSS = [s[1] for s in data["ss_laps"][b]]
        lm.Qfun = [s[2] for s in data["ss_laps"][b]]; lm.LapTime = list(data["lap_times"])
        lm.it, lm.timeStep = 4, int(data["t"][b])
        lm.zt, lm.xLin, lm.uLin = data["zt"][b].copy(), data["xLin"][b].copy(), data["uLin"][b].copy()
        lm.OldInput, lm.xPred = data["OldInput"][b].copy(), data["xPred"][b].copy()
        ctrls.append(lm)
    t0 = time.perf_counter()
    for b, lm in enumerate(ctrls):
        lm.solve(data["x0"][b])
    dt = time.perf_counter() - t0
    return {"value": nsample / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d controller steps, oracle restatement of LMPC.solve in Python (as the reference: single thread) + "
                      "OSQP-algorithm C port, reference settings; cpu %s" % (nsample, cpu_model())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle leg (profiling runs)")
    ap.add_argument("--config", type=int, default=1, choices=[1, 2],
                    help="index into BASELINE.json configs: 1 = batch=4096 LTV-MPC QPs (default, the headline), "
                         "2 = batch=4096 full LMPC steps with k-NN regression over a 5-lap safe set")
    ap.add_argument("--batch", type=int, default=BATCH)
    args = ap.parse_args()
    if args.config == 2:
        run_lmpc_steps(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
