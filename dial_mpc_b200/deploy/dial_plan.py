"""Asynchronous planner process on the B200 sampling core.

Mirror of the reference ``MBDPublisher`` (dial_mpc/deploy/dial_plan.py:64-229): attaches to the
POSIX shared-memory segments created by the simulator / robot bridge (``time_shm, state_shm,
acts_shm, refs_shm, plan_time_shm, tau_shm``; float32, sizes over-allocated x32 exactly like
the reference, dial_plan.py:92-134 / dial_sim.py:84-123), shifts the plan by the elapsed time
with the node spline, runs the annealed ``reverse_once`` scan and publishes joint targets,
torques and reference positions.  Reference quirks kept on purpose: the
deploy schedule has no ``sigma_control`` factor (:199-209), the first call runs
``Ndiffuse_init`` and then ``Ndiffuse`` iterations (:195-212), the planner state is built
without ``mjx.forward`` (qacc_warmstart = 0, :45-61,141-155), and ``update_mjx_state`` refreshes
only ``info["step"]`` (:149-155; a seq-jump ``contact_stage`` is advanced by the rollouts).

One planning cycle is ONE CUDA-graph launch (``DeviceLoop(env_step=False)`` -> C ABI
``dial_mpc_step``): the host writes the state and the time-shifted knots into the bound device
buffers (one small H2D each), launches, and reads back the knots and the ``xbar`` reference.
"""
from __future__ import annotations

import argparse
import importlib
import os
import sys
import time
from multiprocessing import shared_memory

import numpy as np
import torch
import yaml

import dial_mpc_b200.envs as dial_envs
from dial_mpc_b200 import random as drandom
from dial_mpc_b200.core.dial_config import DialConfig
from dial_mpc_b200.core.dial_core import DeviceLoop, MBDPI
from dial_mpc_b200.envs.base_env import PipelineState, State
from dial_mpc_b200.utils.io_utils import get_example_path, load_dataclass_from_dict
from dial_mpc_b200.utils.spline import interp_matrix


def shm_layout(n_acts: int, nx: int, nu: int):
    """name -> (shape, byte size) of the six segments (dial_plan.py:92-134)."""
    return {
        "acts_shm": ((n_acts, nu), n_acts * nu * 32),
        "refs_shm": ((n_acts, nu, 3), n_acts * nu * 3 * 32),
        "plan_time_shm": ((1,), 32),
        "time_shm": ((1,), 32),
        "state_shm": ((nx,), nx * 32),
        "tau_shm": ((n_acts, nu), n_acts * nu * 32),
    }


class MBDPublisher:
    def __init__(self, env, env_config, dial_config: DialConfig, create_shm: bool = False):
        self.dial_config, self.env, self.env_config = dial_config, env, env_config
        self.mbdpi = MBDPI(self.dial_config, self.env)
        self.rng = drandom.PRNGKey(seed=self.dial_config.seed)
        dev = self.mbdpi.device
        self.Y = torch.zeros(self.dial_config.Hnode + 1, self.mbdpi.nu, device=dev)
        self._loop = None      # DeviceLoop, created with the first state
        self.ctrl_dt = env_config.dt
        self.timer_period = env_config.dt
        self.n_acts = self.dial_config.Hsample + 1
        self.nq, self.nv = env.sys.nq, env.sys.nv
        self.nx = self.nq + self.nv
        self.nu = env.sys.nu
        self.default_q = env.sys.keyframe("home")
        self.default_u = np.asarray(env.sys.model.keyframes["home"].get("ctrl", np.zeros(self.nu)), dtype=np.float32)
        self._shm = {}
        for name, (shape, size) in shm_layout(self.n_acts, self.nx, self.nu).items():
            seg = shared_memory.SharedMemory(name=name, create=create_shm, size=size)
            self._shm[name] = seg
            setattr(self, name.replace("_shm", "_shared"), np.ndarray(shape, dtype=np.float32, buffer=seg.buf))
        self.acts_shared[:] = self.default_u
        self.refs_shared[:] = 1.0
        self.plan_time_shared[0] = -0.02
        self.time_shared[0] = 0.0
        self.state_shared[: self.default_q.shape[0]] = self.default_q
        self._first_time = True
        self._state = None
        self._last_plan_time = None

    def close(self, unlink: bool = False):
        for seg in self._shm.values():
            seg.close()
            if unlink:
                seg.unlink()

    # ---- plan shift by an arbitrary elapsed time (dial_plan.py:136-139) ---------------------------
    def shift(self, Y: torch.Tensor, shift_time: float) -> torch.Tensor:
        nodes = self.mbdpi.step_nodes_np
        M = interp_matrix(nodes, nodes + shift_time)       # extrapolates with the last polynomial piece
        return torch.as_tensor(M.astype(np.float32), device=Y.device) @ Y

    # ---- planner state from shared memory (dial_plan.py:141-155) -----------------------------------
    def init_mjx_state(self, q, qd, t) -> State:
        dev = self.mbdpi.device
        f = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32), device=dev)
        base = self.env.reset(drandom.PRNGKey(0))
        ps = PipelineState(f(q), f(qd), torch.zeros(self.nv, device=dev))   # no mjx.forward: warm-start 0
        return State(ps, None, 0.0, 0.0, {}, dict(base.info))

    def update_mjx_state(self, state: State, q, qd, t) -> State:
        """dial_plan.py:149-155: new qpos / qvel, ``info["step"] = int(t / ctrl_dt)``, nothing else."""
        dev = self.mbdpi.device
        f = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32), device=dev)
        ps = PipelineState(f(q), f(qd), state.pipeline_state.qacc_warmstart)
        info = state.info
        info["step"] = int(t / self.ctrl_dt)
        return State(ps, None, state.reward, state.done, state.metrics, info)

    def deploy_factors(self, n: int) -> torch.Tensor:
        """``traj_diffuse_factor ** arange(n)[:, None]`` (dial_plan.py:199-209: no sigma_control),
        broadcast over the Hnode+1 knots."""
        cfg, dev = self.dial_config, self.mbdpi.device
        return ((cfg.traj_diffuse_factor ** torch.arange(n, device=dev, dtype=torch.float32))[:, None]
                * torch.ones(cfg.Hnode + 1, device=dev)[None, :]).contiguous()

    # ---- one planning cycle (body of the reference's `while True`, dial_plan.py:172-229) ------------
    def plan_once(self) -> dict:
        t0 = time.time()
        cfg = self.dial_config
        if self._state is None:
            self._last_plan_time = float(self.time_shared[0])
            self._state = self.init_mjx_state(self.state_shared[: self.nq].copy(), self.state_shared[self.nq:].copy(),
                                              self._last_plan_time)
        plan_time = float(self.time_shared[0])
        state = self._state = self.update_mjx_state(self._state, self.state_shared[: self.nq].copy(),
                                                    self.state_shared[self.nq:].copy(), plan_time)
        shift_time = plan_time - self._last_plan_time
        if shift_time > self.ctrl_dt + 1e-3:
            print(f"[WARN] sim overtime {(shift_time - self.ctrl_dt) * 1000:.1f} ms")
        if shift_time > self.ctrl_dt * self.n_acts:
            print(f"[WARN] long time unplanned {shift_time * 1000:.1f} ms, reset control")
            self.Y = self.Y * 0.0
        else:
            self.Y = self.shift(self.Y, shift_time)
        if self._loop is None:
            nmax = max(cfg.Ndiffuse, cfg.Ndiffuse_init)
            self._loop = DeviceLoop(self.mbdpi, state, self.rng, self.Y, n_diffuse_max=nmax, noise=self.deploy_factors(nmax))
        loop = self._loop
        ps = state.pipeline_state
        loop.set_state(ps.qpos, ps.qvel, ps.qacc_warmstart, step=state.info["step"])
        loop.buf["Y"].copy_(self.Y)
        if self._first_time:
            self._first_time = False
            loop.step(cfg.Ndiffuse_init, env_step=False)
        loop.step(cfg.Ndiffuse, env_step=False)
        self.Y = loop.Y.clone()
        info = loop.info()
        x_targets = info["xbar"][:, 1:, :3]                      # [Hs+1, nbody-2, 3]
        us = self.mbdpi.node2u_vmap(self.Y).cpu().numpy()         # [Hs+1, nu]   (D2H: synchronises)
        self.rng = loop.rng_host()
        joint_targets = np.stack([self.env.act2joint(u) for u in us])
        taus = np.stack([self.env.act2tau(u, state.pipeline_state) for u in us])
        self.acts_shared[: joint_targets.shape[0], :] = joint_targets
        self.tau_shared[: taus.shape[0], :] = taus
        self.plan_time_shared[0] = plan_time
        xt = x_targets.cpu().numpy()
        nref = min(self.refs_shared.shape[1], xt.shape[1])
        self.refs_shared[:, :nref, :] = xt[: self.refs_shared.shape[0], :nref, :]
        self._last_plan_time = plan_time
        dt = time.time() - t0
        if dt > self.ctrl_dt:
            print(f"[WARN] real overtime {dt * 1000:.1f} ms")
        return dict(plan_time=plan_time, wall_s=dt, rews=info["rews"])

    def main_loop(self):
        while True:
            self.plan_once()


def main(args=None):
    parser = argparse.ArgumentParser()
    group = parser.add_mutually_exclusive_group(required=True)
    group.add_argument("--config", type=str, default=None, help="Path to config file")
    group.add_argument("--example", type=str, default=None, help="Example to run")
    parser.add_argument("--custom-env", type=str, default=None, help="Custom environment to import dynamically")
    args = parser.parse_args(args)
    if args.custom_env is not None:
        sys.path.append(os.getcwd())
        importlib.import_module(args.custom_env)
    path = get_example_path(args.example + ".yaml") if args.example else args.config
    config_dict = yaml.safe_load(open(path, "r"))
    dial_config = load_dataclass_from_dict(DialConfig, config_dict)
    env_config_type = dial_envs.get_config(dial_config.env_name)
    env_config = load_dataclass_from_dict(env_config_type, config_dict, convert_list_to_array=True)
    env = dial_envs.get_environment(dial_config.env_name, config=env_config)
    pub = MBDPublisher(env, env_config, dial_config)
    try:
        pub.main_loop()
    except KeyboardInterrupt:
        pass
    finally:
        pub.close()


if __name__ == "__main__":
    main()
