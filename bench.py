#!/usr/bin/env python
"""bench.py — DIAL-MPC sampling core on B200.

One "step" = one MPC planning step of BASELINE.json configs[1] (unitree_go2_seq_jump,
Nsample=2048 per GPU, Hsample=25, Hnode=5, Ndiffuse=4): shift + Ndiffuse x reverse_once
(sample -> spline -> batched full-order rollout -> reward -> softmax update).
Metric: sample-steps/s = Ndiffuse * Nsample_total * Hsample / seconds per step.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # CPU restatement (oracle) on the host cores
"""
from __future__ import annotations

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

WORKLOAD = dict(env_name="unitree_go2_seq_jump", Nsample=2048, Hsample=25, Hnode=5, Ndiffuse=4,
                temp_sample=0.05, horizon_diffuse_factor=0.9, traj_diffuse_factor=0.5)
ENV_CFG = dict(pose_target_sequence=[[0, 0, 0.27], [0.4, 0, 0.27], [0.8, 0, 0.27], [1.2, 0, 0.27], [1.6, 0, 0.27]],
               yaw_target_sequence=[0.0] * 5)
METRIC = "sample-steps/s (Nsample*Hsample/wall-s) per MPC reverse_once, Go2"


# --------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the fp64 NumPy oracle (CPU restatement, NOT reference JAX)
# --------------------------------------------------------------------------------------------
def _oracle_worker(args):
    seed, nrows, n_calls = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle.envs_oracle import make_env
    from oracle.planner_oracle import PlannerOracle
    env = make_env(WORKLOAD["env_name"], ENV_CFG)
    s = env.reset()
    for _ in range(10):
        s, _, _ = env.step(s, np.zeros((1, env.nu)))
    pl = PlannerOracle(env, nrows, WORKLOAD["Hsample"], WORKLOAD["Hnode"], WORKLOAD["temp_sample"],
                       WORKLOAD["horizon_diffuse_factor"], WORKLOAD["traj_diffuse_factor"])
    rng = np.random.default_rng(seed)
    Y = np.zeros((WORKLOAD["Hnode"] + 1, env.nu))
    t0 = time.perf_counter()
    for i in range(n_calls):
        eps = rng.standard_normal((nrows, WORKLOAD["Hnode"] + 1, env.nu))
        Y, _ = pl.reverse_once(s, eps, Y, pl.sigma_control * WORKLOAD["traj_diffuse_factor"] ** (i % WORKLOAD["Ndiffuse"]))
    return time.perf_counter() - t0


def oracle_throughput(rows_per_proc: int, n_calls: int, procs: int):
    """sample-steps/s of the oracle's reverse_once on `procs` host processes, each rolling
    `rows_per_proc` samples of the bench workload `n_calls` times."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    if procs == 1:
        inner = [_oracle_worker((0, rows_per_proc, n_calls))]
    else:
        with ctx.Pool(procs) as pool:
            inner = pool.map(_oracle_worker, [(i, rows_per_proc, n_calls) for i in range(procs)])
    wall = max(inner)   # slowest worker's time inside reverse_once (model load / reset excluded)
    units = procs * rows_per_proc * WORKLOAD["Hsample"] * n_calls
    return units / wall, wall


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    rows = 128
    for _ in range(args.warmup and 1):
        oracle_throughput(rows, 1, cores)
    vals, t_all = [], 0.0
    for _ in range(args.steps):
        v, wall = oracle_throughput(rows, 1, cores)
        vals.append(v)
        t_all += wall
    value = float(np.sum([cores * rows * WORKLOAD["Hsample"]] * args.steps) / t_all)
    sample = (f"{cores} procs x {rows} samples x (Hsample+1)={WORKLOAD['Hsample'] + 1} env steps per step "
              f"(one reverse_once of the bench workload at Nsample={cores * rows}); fp64 NumPy oracle port, "
              "CPU restatement, not reference JAX")
    line = dict(impl="reference", metric=METRIC, value=value, unit="sample-steps/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * t_all / args.steps,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload="unitree_go2_seq_jump Nsample=2048 Hsample=25 Hnode=5 Ndiffuse=4 (bounded sample)",
                            **{k: WORKLOAD[k] for k in ("Hsample", "Hnode")}),
                cpu_baseline=dict(value=value, unit="sample-steps/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit="sample-steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampled every 20 ms from before the warm-up; stop(t0, t1) keeps the samples that
    fall inside the timed region (falls back to the nearest ones if the region is shorter than
    the sampling period)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [t.strip() for t in line.split(",")]))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.05)
        self.proc.terminate()
        ok = [(t, r) for t, r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        inside = [r for t, r in ok if t0 <= t <= t1 + 0.03]
        where = "timed region"
        if not inside and ok:
            mid = 0.5 * (t0 + t1)
            inside = [r for _, r in sorted(ok, key=lambda tr: abs(tr[0] - mid))[:3]]
            where = "nearest samples (region shorter than the sampling period)"
        sm = [float(r[0]) for r in inside]
        mx = [float(r[1]) for r in inside if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in inside for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(sm), window=where)


# --------------------------------------------------------------------------------------------
# own arm
# --------------------------------------------------------------------------------------------
def run_own(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as graft
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        graft.build()
    if world > 1:
        dist.barrier()
    import dial_mpc_b200.envs as E
    from dial_mpc_b200 import random as drandom
    from dial_mpc_b200.core.dial_config import DialConfig
    from dial_mpc_b200.core.dial_core import MBDPI

    W = WORKLOAD
    Ntotal = W["Nsample"] * world          # weak scaling: 2048 samples per GPU
    cfg = DialConfig(env_name=W["env_name"], Nsample=Ntotal, Hsample=W["Hsample"], Hnode=W["Hnode"],
                     Ndiffuse=W["Ndiffuse"], temp_sample=W["temp_sample"],
                     horizon_diffuse_factor=W["horizon_diffuse_factor"], traj_diffuse_factor=W["traj_diffuse_factor"])
    ecfg = E.UnitreeGo2SeqJumpEnvConfig(**{k: np.array(v) for k, v in ENV_CFG.items()})
    env = E.get_environment(cfg.env_name, config=ecfg)
    mb = MBDPI(cfg, env, rank=rank, world_size=world)
    dev = mb.device
    # synthetic state: reset, then 10 env steps with zero action so contacts are settled
    state = env.reset(drandom.PRNGKey(0))
    for _ in range(10):
        state = env.step(state, torch.zeros(mb.nu, device=dev))
    factors = mb.schedule(cfg.Ndiffuse)
    rng = drandom.PRNGKey(cfg.seed)
    Y = torch.zeros(cfg.Hnode + 1, mb.nu, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def mpc_step(Y, rng):
        Y = mb.shift(Y)
        rng, Y, info = mb.reverse_scan(state, rng, Y, factors)
        return Y, rng, info

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        Y, rng, info = mpc_step(Y, rng)
    sync_all()
    launches0 = mb.plan.launches
    evs = []
    sync_all()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()                                   # L2 flush, outside the timed events
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        Y, rng, info = mpc_step(Y, rng)
        e1.record()
        evs.append((e0, e1))
    sync_all()
    t_wall = time.perf_counter() - t_wall0
    launches = mb.plan.launches - launches0
    clocks = sampler.stop(t_wall0, t_wall0 + t_wall) if rank == 0 else None
    t_dev = sum(a.elapsed_time(b) for a, b in evs) / 1e3
    if world > 1:
        t = torch.tensor([t_dev], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev = float(t.item())
    units_per_step = cfg.Ndiffuse * Ntotal * cfg.Hsample
    value = units_per_step * args.steps / t_dev

    # ---- e2e: public API with HOST buffers (pinned), H2D of the state + plan, D2H of the plan ----
    ps = state.pipeline_state
    h_q, h_v, h_w = (t.cpu().pin_memory() for t in (ps.qpos, ps.qvel, ps.qacc_warmstart))
    h_Y = torch.zeros(cfg.Hnode + 1, mb.nu).pin_memory()
    h_out = torch.empty(cfg.Hnode + 1, mb.nu).pin_memory()
    from dial_mpc_b200.envs.base_env import PipelineState, State
    h2d = sum(t.numel() * 4 for t in (h_q, h_v, h_w, h_Y))
    d2h = h_out.numel() * 4 + 4

    def e2e_step(rng):
        d_state = State(PipelineState(h_q.to(dev, non_blocking=True), h_v.to(dev, non_blocking=True),
                                      h_w.to(dev, non_blocking=True)), None, 0.0, 0.0, {}, dict(state.info))
        Yd = mb.shift(h_Y.to(dev, non_blocking=True))
        rng, Yd, info = mb.reverse_scan(d_state, rng, Yd, factors)
        h_out.copy_(Yd, non_blocking=True)
        r = float(info["rews"][-1])                      # D2H read of the plan's reward (syncs)
        h_Y.copy_(h_out)
        return rng, r

    for _ in range(3):
        rng, _ = e2e_step(rng)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rng, _ = e2e_step(rng)
    sync_all()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_value = units_per_step * args.steps / t_e2e

    # ---- roofline of the dominant kernel (rollout_kernel), timed alone with CUDA events ----------
    m = env.sys
    Y0 = torch.zeros(cfg.Hnode + 1, mb.nu, device=dev)
    key = drandom.split(rng)[1]
    for _ in range(3):
        mb.plan.reverse_rollout(state, None, key, Y0, mb.sigma_control, mb._rews_local)
    torch.cuda.synchronize()
    reps = 20
    ks, ke = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ks.record()
    for _ in range(reps):
        mb.plan.reverse_rollout(state, None, key, Y0, mb.sigma_control, mb._rews_local)
    ke.record()
    torch.cuda.synchronize()
    t_kernel = ks.elapsed_time(ke) / 1e3 / reps
    rows, H = mb.Nlocal + 1, cfg.Hsample + 1
    per_rowstep = 4 * (m.nq + m.nv + 3 * (m.nbody - 1))            # q, qd, x.pos written once per env step
    alg_bytes = rows * H * per_rowstep + rows * 4 + (m.nq + 2 * m.nv) * 4
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / t_kernel / 1e9
    traffic, fp32 = None, None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "rollout_traffic.json")))
        traffic = prof["dram_bytes_per_launch"]
        # compute-side view (the binding resource): flop per physics step counted by ncu (committed
        # capture) x physics steps per second measured live, against the nominal fp32 peak
        fpp = prof["fp32"]["flop_per_physics_step"]
        sm_mhz = float(peaks.get("sm_max_mhz", 1965.0))
        peak_tf = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        ach_tf = fpp * rows * H / t_kernel / 1e12
        fp32 = dict(flop_per_physics_step=fpp, achieved_tflops=ach_tf, peak_tflops=peak_tf, frac=ach_tf / peak_tf,
                    peak_source="nominal 148 SM x 128 FFMA/clk x 2 x sm_max_mhz", flop_source="ncu capture profiles/rollout_traffic.json")
    except Exception:
        pass
    roofline = dict(bound="hbm", kernel="rollout_kernel", achieved=achieved, peak=peak, unit="GB/s",
                    frac=achieved / peak, traffic=traffic, peak_source="measured" if peaks else "fallback",
                    kernel_ms=t_kernel * 1e3, algorithmic_bytes_per_launch=alg_bytes,
                    kernel_share_of_step=cfg.Ndiffuse * t_kernel / (t_dev / args.steps),
                    note=("the path is fp32-issue/latency bound (~140 flop/B, SURVEY.md 8d): the HBM fraction is "
                          "reported as required but cannot approach 1; see DESIGN.md"),
                    physics_steps_per_s=rows * H / t_kernel, fp32=fp32)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- CPU baseline (oracle port) on a bounded sample, rank 0 / N=1 only -------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v1, wall1 = oracle_throughput(256, 1, 1)
        vall, wall = oracle_throughput(256, 1, cores)
        cpu = dict(value=vall, unit="sample-steps/s", cores=cores, kind="port",
                   sample=(f"{cores} procs x 256 samples x {W['Hsample'] + 1} env steps (one reverse_once, same env/state); "
                           f"single-core: {v1:.1f} sample-steps/s on 256 samples; fp64 NumPy oracle — CPU restatement, "
                           "not reference JAX"),
                   single_core_value=v1, wall_s=wall + wall1)
    line = dict(metric=METRIC, value=value, unit="sample-steps/s", n_gpus=world, steps=args.steps,
                warmup=max(args.warmup, 3), ms_per_step=1e3 * t_dev / args.steps, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload="unitree_go2_seq_jump (BASELINE configs[1])", Nsample_per_gpu=W["Nsample"],
                            Nsample_total=Ntotal, Hsample=W["Hsample"], Hnode=W["Hnode"], Ndiffuse=W["Ndiffuse"],
                            step="shift + Ndiffuse x reverse_once (rollout, allgather, update, bars)",
                            rng="in-kernel Threefry-2x32", l2="256 MiB memset between steps (outside the timed events)",
                            parallelism=f"samples sharded over {world} GPU(s), 1 allgather(rews) + 1 allreduce(bars) per reverse_once"),
                clocks=clocks, e2e=dict(value=e2e_value, unit="sample-steps/s", h2d_bytes_per_step=h2d,
                                        d2h_bytes_per_step=d2h, ms_per_step=1e3 * t_e2e / args.steps),
                gpu_launches=int(launches), wall_s_timed_region=t_wall, roofline=roofline)
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
