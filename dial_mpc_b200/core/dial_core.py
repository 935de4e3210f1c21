"""DIAL-MPC planner on the B200 sampling core.

Same class and method surface as the reference ``MBDPI`` (dial_mpc/core/dial_core.py:51-172)
and the same synchronous MPC loop as its ``main()`` (:175-268).  The shift-and-anneal loop
stays in Python; ``reverse_once`` is two kernel stages (rollout, update) reached through the
C ABI, with one NCCL allgather of the per-sample rewards between them when the samples are
sharded over several GPUs (one process per GPU).
"""
from __future__ import annotations

import argparse
import importlib
import os
import sys
import time
from typing import Any, Dict, Optional

import numpy as np
import torch
import yaml

import dial_mpc_b200.envs as dial_envs
from dial_mpc_b200 import random as drandom
from dial_mpc_b200.core.dial_config import DialConfig
from dial_mpc_b200.plan import Plan
from dial_mpc_b200.utils.io_utils import get_example_path, load_dataclass_from_dict
from dial_mpc_b200.utils.spline import interp_matrix


def rollout_us(step_env, state, us):
    """Reference semantics of ``rollout_us`` (dial_core.py:36-42) for a *single* action
    sequence, expressed with the env's own ``step`` (used by tests / custom callers)."""
    rews, pipeline_states = [], []
    for u in us:
        state = step_env(state, u)
        rews.append(state.reward)
        pipeline_states.append(state.pipeline_state)
    return torch.stack([torch.as_tensor(r) for r in rews]), pipeline_states


def softmax_update(weights, Y0s, sigma, mu_0t):
    """``softmax_update`` (dial_core.py:45-48) on torch tensors (API parity; the planner's
    own update runs in ``ybar_kernel``)."""
    return torch.einsum("n,nij->ij", weights, Y0s), sigma


class MBDPI:
    def __init__(self, args: DialConfig, env, rank: int = 0, world_size: int = 1, process_group=None,
                 compute_bars: bool = True, plan_factory=None):
        self.args = args
        self.env = env
        self.nu = env.action_size
        if args.update_method != "mppi":
            raise KeyError(args.update_method)
        self.update_fn = softmax_update
        self.rank, self.world_size, self.pg = rank, world_size, process_group
        self.compute_bars = compute_bars
        if args.Nsample % world_size != 0:
            raise ValueError("Nsample must be divisible by the number of ranks")
        self.Nlocal = args.Nsample // world_size

        sigma_control = args.horizon_diffuse_factor ** np.arange(args.Hnode + 1)[::-1]
        self.sigma_control_np = (sigma_control * args.sigma_scale).astype(np.float64)

        # node to u (dial_core.py:73-77; ctrl_dt is hard-coded to 0.02 there)
        self.ctrl_dt = 0.02
        self.step_us_np = np.linspace(0, self.ctrl_dt * args.Hsample, args.Hsample + 1)
        self.step_nodes_np = np.linspace(0, self.ctrl_dt * args.Hsample, args.Hnode + 1)
        self.node_dt = self.ctrl_dt * (args.Hsample) / (args.Hnode)
        self.M_n2u_np = interp_matrix(self.step_nodes_np, self.step_us_np)
        self.M_u2n_np = interp_matrix(self.step_us_np, self.step_nodes_np)

        desc = env.plan_desc(Nsample=self.Nlocal, Ntotal=args.Nsample, shard_offset=rank * self.Nlocal,
                             Hsample=args.Hsample, Hnode=args.Hnode, temp_sample=args.temp_sample,
                             M_n2u=self.M_n2u_np)
        # plan_factory exists for the CPU test harness (tests/emul); the product path is Plan
        self.plan = (plan_factory or Plan)(env, desc)
        dev = self.plan.device
        self.device = dev
        f = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32), device=dev)
        self.sigma_control = f(self.sigma_control_np)
        self.step_us, self.step_nodes = f(self.step_us_np), f(self.step_nodes_np)
        self.M_n2u, self.M_u2n = f(self.M_n2u_np), f(self.M_u2n_np)
        Hs1 = args.Hsample + 1
        P = np.zeros((Hs1, Hs1))
        P[np.arange(Hs1 - 1), np.arange(1, Hs1)] = 1.0  # roll(-1) with the last row zeroed
        self.M_shift = f(self.M_u2n_np @ P @ self.M_n2u_np)
        # persistent buffers
        N, Nl = args.Nsample, self.Nlocal
        self._rews_local = torch.empty(Nl + 1, dtype=torch.float32, device=dev)
        self._rews_all = torch.empty(N + 1, dtype=torch.float32, device=dev)
        self._weights = torch.empty(N + 1, dtype=torch.float32, device=dev)
        m = env.sys
        self._bars = torch.empty(Hs1 * (m.nq + m.nv + 3 * (m.nbody - 1)), dtype=torch.float32, device=dev)
        # the info-only bars (+ their allreduce) run on a side stream and overlap the next rollout
        self._side = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._bar_events = []
        # sharded plans: rewards (and bars) are exchanged through NVLink peer memory by the kernels
        # themselves (dial_exchange_*).  DIAL_EXCHANGE=nccl keeps the host-issued NCCL collectives;
        # they are also the fallback when CUDA IPC cannot map the peers (different nodes / no P2P).
        self.xch = False
        self.xch_error = None
        if world_size > 1 and dev.type == "cuda" and hasattr(self.plan, "exchange_setup") \
                and os.environ.get("DIAL_EXCHANGE", "p2p") != "nccl":
            try:
                self.plan.exchange_setup(rank, world_size, process_group)
                self.xch = True
            except Exception as e:  # noqa: BLE001  (kept: reported through exchange_name / bench line)
                self.xch_error = str(e)

    # -- spline maps (dial_core.py:82-101) -------------------------------------------------------
    def node2u(self, nodes):
        return self.M_n2u @ self._t(nodes)

    def u2node(self, us):
        return self.M_u2n @ self._t(us)

    def node2u_vmap(self, Y):      # (horizon, node)
        return self.M_n2u @ self._t(Y)

    def u2node_vmap(self, u):
        return self.M_u2n @ self._t(u)

    def node2u_vvmap(self, Ys):    # (batch, horizon, node)
        return torch.einsum("tk,bka->bta", self.M_n2u, self._t(Ys))

    def u2node_vvmap(self, us):
        return torch.einsum("kt,bta->bka", self.M_u2n, self._t(us))

    def _t(self, x):
        return self.plan.f32(x)

    # -- batched rollout (dial_core.py:80-81) ------------------------------------------------------
    def rollout_us_vmap(self, state, us):
        """-> (rewss [B,H], (q [B,H,nq], qd [B,H,nv], x_pos [B,H,nbody-1,3]))."""
        rewss, q, qd, x = self.plan.rollout(state, us)
        return rewss, (q, qd, x)

    # -- the hot path (dial_core.py:103-145) ----------------------------------------------------------
    def reverse_once(self, state, rng, Ybar_i, noise_scale, eps=None, _sync_bars=True):
        """One annealing iteration.  ``eps`` (optional, [Nsample,Hnode+1,nu]) injects the noise;
        otherwise it is drawn in-kernel from the Threefry stream keyed by ``split(rng)[1]``."""
        rng, Y0s_rng = drandom.split(rng)
        Ybar_i = self._t(Ybar_i)
        noise_scale = self._t(noise_scale)
        if eps is not None:
            eps = self.plan.f32(eps, (self.args.Nsample, self.args.Hnode + 1, self.nu))
        key = None if eps is not None else Y0s_rng
        N, Nl = self.args.Nsample, self.Nlocal
        if self._side is not None and len(self._bar_events) >= 2:
            # the trajectory buffer about to be overwritten was read by the bars two iterations ago
            torch.cuda.current_stream().wait_event(self._bar_events.pop(0))
        Ybar = torch.empty_like(Ybar_i)
        weights = torch.empty_like(self._weights)
        if self.world_size == 1:
            rews = torch.empty_like(self._rews_local)      # fresh output (functional API), written by the kernel
            self.plan.reverse_rollout(state, eps, key, Ybar_i, noise_scale, rews)
            self.plan.reverse_update(eps, key, Ybar_i, noise_scale, rews, Ybar, weights)
        elif self.xch:
            # the rollout epilogue stores the rewards into every rank's mailbox; the weights kernel waits
            rews = torch.empty_like(self._rews_all)
            self.plan.reverse_rollout(state, eps, key, Ybar_i, noise_scale, self._rews_local)
            self.plan.reverse_update(eps, key, Ybar_i, noise_scale, None, Ybar, weights, rews_gathered=rews)
        else:
            import torch.distributed as dist
            self.plan.reverse_rollout(state, eps, key, Ybar_i, noise_scale, self._rews_local)
            dist.all_gather_into_tensor(self._rews_all[:N], self._rews_local[:Nl], group=self.pg)
            self._rews_all[N:].copy_(self._rews_local[Nl:])
            rews = self._rews_all.clone()
            self.plan.reverse_update(eps, key, Ybar_i, noise_scale, rews, Ybar, weights)
        info: Dict[str, Any] = {"rews": rews, "new_noise_scale": noise_scale, "weights": weights}
        if self.compute_bars:
            m = self.env.sys
            Hs1 = self.args.Hsample + 1
            n1, n2 = Hs1 * m.nq, Hs1 * m.nv
            bars = torch.empty_like(self._bars)
            qbar, qdbar, xbar = bars[:n1], bars[n1:n1 + n2], bars[n1 + n2:]
            main = torch.cuda.current_stream() if self._side is not None else None
            if self._side is not None:
                self._side.wait_stream(main)
                ctx = torch.cuda.stream(self._side)
            else:
                import contextlib
                ctx = contextlib.nullcontext()
            with ctx:
                self.plan.reverse_trajbar(weights, self.rank, qbar, qdbar, xbar)   # exchange on: already the all-rank sum
                if self.world_size > 1 and not self.xch:
                    import torch.distributed as dist
                    dist.all_reduce(bars, group=self.pg)
                if self._side is not None:
                    ev = torch.cuda.Event()
                    ev.record(self._side)
                    self._bar_events.append(ev)
                    for t in (bars, weights):
                        t.record_stream(self._side)
            if self._side is not None and _sync_bars:
                main.wait_stream(self._side)
            info["qbar"] = qbar.view(Hs1, m.nq)
            info["qdbar"] = qdbar.view(Hs1, m.nv)
            info["xbar"] = xbar.view(Hs1, m.nbody - 1, 3)
        return rng, Ybar, info

    @property
    def exchange_name(self) -> str:
        if self.world_size == 1:
            return "no exchange (single GPU)"
        return ("one peer-memory exchange fused into the rollout epilogue / weights prologue (NVLink stores + flags)"
                if self.xch else "one NCCL allgather" + (f" (peer exchange unavailable: {self.xch_error})" if self.xch_error else ""))

    def phase_times(self, state, key, Ybar_i, noise_scale, reps: int = 10) -> Dict[str, float]:
        """Device time (microseconds, CUDA events on the current stream) of the stages of one
        ``reverse_once``: rollout | rewards exchange | weights + Ybar | bars (+ their allreduce).
        Measurement aid for bench.py; the stages run back to back on one stream here."""
        Ybar_i, noise_scale = self._t(Ybar_i), self._t(noise_scale)
        N, Nl = self.args.Nsample, self.Nlocal
        m = self.env.sys
        Hs1 = self.args.Hsample + 1
        n1, n2 = Hs1 * m.nq, Hs1 * m.nv
        bars = torch.empty_like(self._bars)
        Ybar = torch.empty_like(Ybar_i)
        acc = {"rollout": 0.0, "exchange": 0.0, "update": 0.0, "bars": 0.0}
        for rep in range(reps + 2):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            ev[0].record()
            self.plan.reverse_rollout(state, None, key, Ybar_i, noise_scale, self._rews_local)
            ev[1].record()
            if self.world_size > 1 and not self.xch:
                import torch.distributed as dist
                dist.all_gather_into_tensor(self._rews_all[:N], self._rews_local[:Nl], group=self.pg)
                self._rews_all[N:].copy_(self._rews_local[Nl:])
                rews_all = self._rews_all
            else:
                rews_all = None if self.xch else self._rews_local   # exchange: the wait is inside the update stage
            ev[2].record()
            self.plan.reverse_update(None, key, Ybar_i, noise_scale, rews_all, Ybar, self._weights)
            ev[3].record()
            self.plan.reverse_trajbar(self._weights, self.rank, bars[:n1], bars[n1:n1 + n2], bars[n1 + n2:])
            if self.world_size > 1 and not self.xch:
                dist.all_reduce(bars, group=self.pg)
            ev[4].record()
            torch.cuda.synchronize()
            if rep >= 2:
                for i, k in enumerate(acc):
                    acc[k] += ev[i].elapsed_time(ev[i + 1]) * 1e3 / reps
        return acc

    def reverse_scan(self, state, rng, Y0, factors):
        """``lax.scan(reverse_scan, (rng, Y0, state), factors)`` of dial_core.py:177-180,262-264."""
        info = None
        n = factors.shape[0]
        for i in range(n):
            rng, Y0, info = self.reverse_once(state, rng, Y0, factors[i], _sync_bars=(i == n - 1))
        return rng, Y0, info

    def schedule(self, n_diffuse: int) -> torch.Tensor:
        """``sigma_control * traj_diffuse_factor ** arange(n_diffuse)[:, None]`` (dial_core.py:259-261)."""
        f = self.args.traj_diffuse_factor ** torch.arange(n_diffuse, device=self.device, dtype=torch.float32)
        return self.sigma_control[None, :] * f[:, None]

    # -- shift (dial_core.py:160-172) -------------------------------------------------------------------
    def shift(self, Y):
        return self.M_shift @ self._t(Y)

    def shift_Y_from_u(self, u, n_step):
        u = self._t(u)
        u = torch.roll(u, -n_step, dims=0)
        u[-n_step:] = 0.0
        return self.u2node_vmap(u)


class DeviceLoop:
    """Device-resident synchronous MPC loop: the reference's per-step sequence
    (dial_core.py:242-268) — ``state = step_env(state, Y0[0]); Y0 = shift(Y0); Y0 = reverse_scan(...)``
    — replayed as ONE CUDA graph per control step (C ABI ``dial_mpc_bind`` / ``dial_mpc_step``).
    State, step counters, rng and control knots live in device tensors owned by this object; the
    host only launches the graph and reads back what it needs (e.g. ``action``).  Equals the
    eager ``env.step`` + ``MBDPI.shift`` + ``MBDPI.reverse_scan`` sequence (tests/test_gpu_parity.py).
    Sharded plans replay the same graph on every rank (one process per GPU): the rewards cross
    the GPUs inside the kernels (peer-memory exchange), so no host collective sits in the step."""

    def __init__(self, mbdpi: "MBDPI", state, rng, Y0=None, n_diffuse_max: Optional[int] = None,
                 compute_bars: bool = True, noise=None):
        """``noise`` [>= n_diffuse_max, Hnode+1]: annealing schedule, default ``mbdpi.schedule`` (the
        deploy planner passes its own, dial_plan.py:199-209)."""
        if mbdpi.world_size != 1 and not mbdpi.xch:
            raise RuntimeError("DeviceLoop on a sharded plan needs the peer-memory exchange (dial_exchange_*); "
                               f"it is off: {mbdpi.xch_error or 'DIAL_EXCHANGE=nccl'}")
        self.mbdpi, self.plan = mbdpi, mbdpi.plan
        a, pl, dev = mbdpi.args, mbdpi.plan, mbdpi.device
        nmax = int(n_diffuse_max or max(a.Ndiffuse, a.Ndiffuse_init))
        ps = state.pipeline_state
        f, e = pl.f32, pl.empty
        Hs1 = a.Hsample + 1
        m = mbdpi.env.sys
        key = np.ascontiguousarray(rng, dtype=np.uint32).view(np.int32)
        self.info0 = dict(state.info)
        self.buf = dict(
            qpos=f(ps.qpos).clone(), qvel=f(ps.qvel).clone(), qacc_warmstart=f(ps.qacc_warmstart).clone(),
            counters=torch.tensor([int(state.info.get("step", 0)), int(state.info.get("contact_stage", 0))],
                                  dtype=torch.int32, device=dev),
            rng=torch.as_tensor(key.copy(), device=dev),
            Y=(torch.zeros(a.Hnode + 1, mbdpi.nu, device=dev) if Y0 is None else f(Y0).clone()),
            ctrl=torch.zeros(mbdpi.nu, device=dev), reward=torch.zeros(1, device=dev),
            rews=torch.zeros(mbdpi.Nlocal + 1, device=dev),
            rews_all=(torch.zeros(a.Nsample + 1, device=dev) if mbdpi.world_size > 1 else None),
            qbar=e(Hs1, m.nq) if compute_bars else None, qdbar=e(Hs1, m.nv) if compute_bars else None,
            xbar=e(Hs1, m.nbody - 1, 3) if compute_bars else None,
            noise=(mbdpi.schedule(nmax) if noise is None else f(noise)).contiguous())
        assert tuple(self.buf["noise"].shape) == (nmax, a.Hnode + 1) or self.buf["noise"].shape[0] >= nmax
        self.n_diffuse_max = nmax
        # randomize_tasks: host mirror of info["step"] / info["rng"] (the env's key chain), from which
        # the one-step random command the horizon may reach is computed ahead (BaseEnv.command_override)
        self._rand = bool(state.info.get("randomize_target", False))
        self._env_info = {"randomize_target": self._rand, "step": int(state.info.get("step", 0)),
                          "rng": np.asarray(state.info.get("rng", np.zeros(2)), dtype=np.uint32).copy()}
        if self._rand and hasattr(mbdpi.env, "stage_tables"):
            # seq-jump: the jump sequence drawn at reset is constant afterwards; one upload at bind time
            pl.set_stages(mbdpi.env.stage_tables(state.info))
        pl.mpc_bind(self.buf, mbdpi.M_shift.cpu().numpy())

    def step(self, n_diffuse: Optional[int] = None, env_step=True) -> None:
        """One control step (asynchronous on the current stream).  env_step: True = env step + shift
        + plan (the reference's main loop), False = plan only, 2 = shift + plan (state untouched)."""
        n = self.mbdpi.args.Ndiffuse if n_diffuse is None else int(n_diffuse)
        if n > self.n_diffuse_max:
            raise ValueError("n_diffuse exceeds the bound noise schedule")
        stepping = env_step is True or env_step == 1
        if self._rand:
            # the env step (if any) runs at info["step"], the rollouts cover the Hsample+1 steps after it
            self.plan.set_command(self.mbdpi.env.command_override(
                self._env_info, self.mbdpi.args.Hsample + (2 if stepping else 1)))
        self.plan.mpc_step(n, env_step)
        if stepping:
            self._env_info["step"] += 1
            if self._rand:
                from dial_mpc_b200 import random as drandom
                self._env_info["rng"] = drandom.split(self._env_info["rng"])[0]

    def rng_host(self) -> np.ndarray:
        """The planner rng after the steps launched so far (synchronises)."""
        return self.buf["rng"].cpu().numpy().view(np.uint32).copy()

    def set_state(self, qpos, qvel, qacc_warmstart=None, step: Optional[int] = None) -> None:
        """Overwrite the planning state (deploy: the state comes from the robot / simulator).  ``step``
        sets ``info["step"]`` only, like the reference's ``update_mjx_state`` (dial_plan.py:149-155):
        a seq-jump ``contact_stage`` is whatever the bound state carries."""
        self.buf["qpos"].copy_(self.plan.f32(qpos))
        self.buf["qvel"].copy_(self.plan.f32(qvel))
        if qacc_warmstart is not None:
            self.buf["qacc_warmstart"].copy_(self.plan.f32(qacc_warmstart))
        if step is not None:
            self.buf["counters"][0] = int(step)
            self._env_info["step"] = int(step)

    @property
    def action(self) -> torch.Tensor:
        """``Y0[0]``: the action the next env step applies (device view)."""
        return self.buf["Y"][0]

    @property
    def Y(self) -> torch.Tensor:
        return self.buf["Y"]

    @property
    def reward(self) -> torch.Tensor:
        return self.buf["reward"][0]

    def info(self) -> Dict[str, Any]:
        b = self.buf
        d = {"rews": b["rews_all"] if b["rews_all"] is not None else b["rews"]}
        if b["qbar"] is not None:
            d.update(qbar=b["qbar"], qdbar=b["qdbar"], xbar=b["xbar"])
        return d

    def state(self):
        """Materialise the env ``State`` (synchronises: reads the counters)."""
        from dial_mpc_b200.envs.base_env import PipelineState, State
        b = self.buf
        c = b["counters"].cpu().numpy()
        info = dict(self.info0)
        info["step"] = int(c[0])
        if "contact_stage" in info:
            info["contact_stage"] = int(c[1])
        info["rng"] = b["rng"].cpu().numpy().view(np.uint32).copy()
        ps = PipelineState(b["qpos"].clone(), b["qvel"].clone(), b["qacc_warmstart"].clone(), b["ctrl"].clone())
        return State(ps, None, b["reward"][0].clone(), 0.0, {}, info)


def save_run(output_dir, rollout, infos, timestamp=None):
    """End-of-run dumps of the reference (dial_core.py:305-323): ``*_states.npy`` rows
    ``[i, qpos, qvel, ctrl]`` and ``*_predictions.npy`` = per control step the ``xbar`` of the LAST
    diffusion iteration, shape [n_steps, Hsample+1, nbody-1, 3] (``infos[i]["xbar"][-1]`` there:
    ``lax.scan`` stacks the iterations, ``[-1]`` picks the last one, not the last horizon step)."""
    os.makedirs(output_dir, exist_ok=True)
    timestamp = timestamp or time.strftime("%Y%m%d-%H%M%S")
    states = torch.stack([torch.as_tensor(r) for r in rollout]).cpu().numpy()
    preds = torch.stack([torch.as_tensor(x) for x in infos]).cpu().numpy()
    assert preds.ndim == 4 and preds.shape[-1] == 3, preds.shape
    np.save(os.path.join(output_dir, f"{timestamp}_states"), states)
    np.save(os.path.join(output_dir, f"{timestamp}_predictions"), preds)
    return states, preds


def main():
    """Synchronous MPC loop — dial_core.py:175-268 without the rendering / flask tail."""
    parser = argparse.ArgumentParser()
    g = parser.add_mutually_exclusive_group(required=True)
    g.add_argument("--config", type=str, default=None)
    g.add_argument("--example", type=str, default=None)
    g.add_argument("--list-examples", action="store_true")
    parser.add_argument("--custom-env", type=str, default=None, help="Custom environment to import dynamically")
    parser.add_argument("--n-steps", type=int, default=None)
    parser.add_argument("--eager", action="store_true",
                        help="per-call launches (env.step / reverse_scan) instead of the CUDA-graph loop")
    args = parser.parse_args()
    from dial_mpc_b200.examples import examples
    if args.list_examples:
        print("Examples:")
        for example in examples:
            print(f"  {example}")
        return
    if args.custom_env is not None:
        sys.path.append(os.getcwd())
        importlib.import_module(args.custom_env)
    if args.example is not None:
        config_dict = yaml.safe_load(open(get_example_path(args.example + ".yaml")))
    else:
        config_dict = yaml.safe_load(open(args.config))
    dial_config = load_dataclass_from_dict(DialConfig, config_dict)
    rng = drandom.PRNGKey(seed=dial_config.seed)
    env_config_type = dial_envs.get_config(dial_config.env_name)
    env_config = load_dataclass_from_dict(env_config_type, config_dict, convert_list_to_array=True)
    env = dial_envs.get_environment(dial_config.env_name, config=env_config)
    mbdpi = MBDPI(dial_config, env)
    rng, rng_reset = drandom.split(rng)
    state = env.reset(rng_reset)
    Y0 = torch.zeros(dial_config.Hnode + 1, mbdpi.nu, device=mbdpi.device)
    rng_exp, rng = drandom.split(rng)
    Nstep = args.n_steps or dial_config.n_steps
    rews, rollout, infos = [], [], []
    if mbdpi.world_size == 1 and not args.eager:
        # one CUDA graph per control step; the host launches it and logs
        loop = DeviceLoop(mbdpi, state, rng, Y0)
        b = loop.buf
        t0, tlast = time.time(), -1
        for t in range(Nstep):
            loop.step(dial_config.Ndiffuse_init if t == 0 else dial_config.Ndiffuse)
            rollout.append(torch.cat([torch.tensor([float(t)], device=mbdpi.device), b["qpos"], b["qvel"], b["ctrl"]]))
            rews.append(b["reward"][0].clone())
            infos.append(b["xbar"].clone())   # = infos[i]["xbar"][-1] of the reference: last diffusion iteration, full horizon
            if t % 10 == 0:
                r = float(rews[-1])  # synchronises: the rate below is whole control steps per second
                print(f"step {t}: rew={r:.3e} freq={(t - tlast) / (time.time() - t0):.1f} Hz")
                t0, tlast = time.time(), t
    else:
        for t in range(Nstep):
            state = env.step(state, Y0[0])
            ps = state.pipeline_state
            rollout.append(torch.cat([torch.tensor([float(t)], device=mbdpi.device), ps.qpos, ps.qvel, ps.ctrl]))
            rews.append(state.reward)
            Y0 = mbdpi.shift(Y0)
            n_diffuse = dial_config.Ndiffuse_init if t == 0 else dial_config.Ndiffuse
            t0 = time.time()
            rng, Y0, info = mbdpi.reverse_scan(state, rng, Y0, mbdpi.schedule(n_diffuse))
            torch.cuda.synchronize()
            freq = 1 / (time.time() - t0)
            infos.append(info["xbar"])
            if t % 10 == 0:
                print(f"step {t}: rew={float(state.reward):.3e} freq={freq:.1f} Hz")
    rew = torch.stack([torch.as_tensor(r) for r in rews]).mean()
    print(f"mean reward = {float(rew):.2e}")
    save_run(dial_config.output_dir, rollout, infos)


if __name__ == "__main__":
    main()
