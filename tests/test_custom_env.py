"""Custom-environment path (reference README.md:223-312 "Writing Custom
Environment", ``--custom-env`` dial_mpc/core/dial_core.py:202-204): a user model compiled from
MJCF + a user reward written as a CUDA device function, fused into a dedicated library build.

CPU: the reward + device code run in the warp emulator against the oracle with the same
reward written in NumPy; the nvcc cross-compile of the custom library is checked for its
exports.  GPU: the real custom build against the oracle."""
import importlib
import os
import sys
import tempfile

import numpy as np
import pytest

from tests.conftest import ROOT

EX = os.path.join(ROOT, "dial_mpc_b200", "examples", "custom_env")


def _reward_np(c):
    """NumPy twin of examples/custom_env/quadpod_reward.cuh (the oracle's side of the check)."""
    u = c["user"]
    t = c["step"].astype(np.float64) * c["dt"]
    vx_tar = np.minimum(u[0] * t / u[1], u[0])
    v, w = c["xd_vel"][:, 1], c["xd_ang"][:, 1]
    r_vel = -((v[:, 0] - vx_tar) ** 2 + v[:, 1] ** 2)
    r_yaw = -w[:, 2] ** 2
    dz = c["xpos"][:, 1, 2] - u[2]
    q = c["xquat"][:, 1]
    upz = 1.0 - 2.0 * (q[:, 1] ** 2 + q[:, 2] ** 2)
    r_up = -(1.0 - upz)
    r_energy = -np.sum((c["ctrl"] / u[5]) ** 2, -1)
    r_feet = -np.sum(c["contact_dist"] ** 2, -1)
    r_head = -(c["site_xpos"][:, 0, 2] - (u[2] + 0.02)) ** 2
    return r_vel + 0.1 * r_yaw - 10.0 * dz * dz + 0.5 * r_up + u[3] * r_energy + u[4] * r_feet + r_head


@pytest.fixture(scope="module")
def pair():
    if EX not in sys.path:
        sys.path.insert(0, EX)
    qe = importlib.import_module("quadpod_env")
    import dial_mpc_b200.envs as E
    from oracle.envs_oracle import CustomRewardOracle
    cfg = qe.QuadpodEnvConfig()
    env = E.get_environment("quadpod_walk", config=cfg)
    tmp = tempfile.NamedTemporaryFile(suffix=".json", delete=False)
    tmp.close()
    env.sys.model.save(tmp.name)
    o = CustomRewardOracle(tmp.name, _reward_np, user=env.user_params(), joint_range=env.joint_range,
                           kp=cfg.kp, kd=cfg.kd, dt=cfg.dt, timestep=cfg.timestep)
    yield env, o
    os.unlink(tmp.name)


def test_custom_env_descriptor_and_registry(pair):
    env, o = pair
    from dial_mpc_b200 import _capi
    d = env.plan_desc(Nsample=8, Hsample=10, Hnode=4)
    assert d.env_id == _capi.ENV_IDS["custom"] == _capi.DEFINES.get("DIAL_ENV_CUSTOM", 5)
    assert d.n_user == 6 and abs(d.user[0] - 0.5) < 1e-7 and d.user[5] == 12.0
    assert env.action_size == 8 and env.sys.nv == 14
    np.testing.assert_allclose(env._init_q, o.init_q)
    import dial_mpc_b200.envs as E
    assert E.get_config("quadpod_walk").__name__ == "QuadpodEnvConfig"


def test_custom_reward_in_emulator_matches_oracle(pair):
    from tests.emul import emul
    env, o = pair
    s = o.reset()
    s.step[:] = 20
    rng = np.random.default_rng(3)
    H = 12
    us = np.clip(rng.normal(size=(3, H, env.action_size)) * 0.6, -1, 1)
    rew, q, qd, x = o.rollout(s, us)
    out = emul.rollout(env, env.plan_desc(), s.qpos[0], s.qvel[0], s.qacc_warmstart[0], us=us, step0=20)
    assert np.abs(out["q"] - q).max() < 1e-4
    assert np.abs(out["qd"] - qd).max() < 5e-3
    assert np.abs(out["xpos"] - x).max() < 1e-4
    assert np.abs(out["rewss"] - rew).max() < 1e-3 * (1 + np.abs(rew).max())
    assert np.abs(rew).max() > 1e-3   # the reward is not trivially zero


def test_stock_library_rejects_custom_env(pair, built):
    """The product path fails loudly: the stock build has no reward for DIAL_ENV_CUSTOM."""
    import ctypes as C
    from dial_mpc_b200 import _capi
    env, _ = pair
    lib = _capi.lib()
    assert lib.dial_custom_reward_id() == b""
    md = _capi.fill_model_desc(env.sys.model)
    assert lib.dial_solver_variant(md) == 1     # 4 hanging 2-dof legs fit the star<3,6> instantiation
    h = lib.dial_plan_create(C.byref(md), C.byref(env.plan_desc()))
    assert not h and b"DIAL_ENV_CUSTOM" in lib.dial_last_error()


def test_custom_library_builds_and_exports(pair, built):
    """nvcc cross-compiles the custom build without a GPU; it carries the same C ABI."""
    from dial_mpc_b200 import _capi, custom
    env, _ = pair
    path = env.library_path
    assert os.path.exists(path) and path.startswith(custom.CACHE_DIR)
    lib = _capi.lib(path)
    assert lib.dial_custom_reward_id().decode() == custom.reward_id(env.reward_source, 1)
    for sym in _capi.EXPORTS:
        assert hasattr(lib, sym)
    assert custom.build_library(env.reward_source, model=env.sys.model) == path   # cached


@pytest.mark.gpu
def test_gpu_custom_env_matches_oracle(pair, built):
    import torch
    from dial_mpc_b200.core.dial_config import DialConfig
    from dial_mpc_b200.core.dial_core import MBDPI
    from dial_mpc_b200 import random as drandom
    from oracle.planner_oracle import PlannerOracle, jax_normal_legacy
    env, o = pair
    # rollout_us_vmap through the custom library
    s = o.reset()
    state = env.reset(drandom.PRNGKey(0))
    np.testing.assert_allclose(state.pipeline_state.qpos.cpu().numpy(), s.qpos[0], atol=1e-6)
    rng = np.random.default_rng(5)
    N, Hs, Hn = 64, 12, 4
    args = DialConfig(env_name="quadpod_walk", Nsample=N, Hsample=Hs, Hnode=Hn, Ndiffuse=1, temp_sample=0.05,
                      horizon_diffuse_factor=0.9, traj_diffuse_factor=0.5)
    mb = MBDPI(args, env)
    us = np.clip(rng.normal(size=(5, Hs + 1, env.action_size)) * 0.6, -1, 1).astype(np.float32)
    rewss, ps = mb.rollout_us_vmap(state, us)
    rew, q, qd, x = o.rollout(s, us.astype(np.float64))
    assert np.abs(rewss.cpu().numpy() - rew).max() < 1e-3 * (1 + np.abs(rew).max())
    # one reverse_once with injected-equivalent native RNG
    pl = PlannerOracle(o, N, Hs, Hn, 0.05, 0.9, 0.5)
    key = drandom.PRNGKey(7)
    _, k2 = drandom.split(key)
    eps = jax_normal_legacy(tuple(int(v) for v in k2), (N, Hn + 1, env.action_size))
    Ybar = np.zeros((Hn + 1, env.action_size))
    Yo, info_o = pl.reverse_once(s, eps, Ybar, pl.sigma_control)
    _, Y, info = mb.reverse_once(state, key, torch.zeros(Hn + 1, env.action_size, device="cuda"),
                                 torch.as_tensor(pl.sigma_control, dtype=torch.float32, device="cuda"))
    r = info["rews"].cpu().numpy()
    ok = np.abs(r - info_o["rews"]) < 2e-3 * (1 + np.abs(info_o["rews"]))
    assert ok.mean() > 0.97, (ok.mean(), np.abs(r - info_o["rews"]).max())
    assert np.abs(Y.cpu().numpy() - Yo).max() < 5e-3


# ---------------------------------------------------------------------------------------------
# a custom env on the DENSE solver path (elliptic cones) with its own dof count: the stock library
# instantiates the dense solver for nv = 22, the custom build for the model at hand (-DDIAL_DENSE_NV)
# ---------------------------------------------------------------------------------------------
def _pincher_reward_np(c):
    """NumPy twin of examples/custom_env/pincher_reward.cuh."""
    u = c["user"]
    w = c["xd_ang"][:, 1]
    r_spin = -((w[:, 2] - u[1]) ** 2 + w[:, 0] ** 2 + w[:, 1] ** 2)
    p = c["xpos"][:, 1]
    r_pos = -(p[:, 0] ** 2 + p[:, 1] ** 2 + (p[:, 2] - u[0]) ** 2)
    r_joint = -np.sum((c["qpos"][:, 7:11] - u[3:7]) ** 2, -1)
    dmin = np.minimum(1.0, c["contact_dist"].min(-1))
    return 0.05 * r_spin + 50.0 * r_pos + u[2] * r_joint - 2.0 * np.maximum(dmin, 0.0)


@pytest.fixture(scope="module")
def pincher():
    if EX not in sys.path:
        sys.path.insert(0, EX)
    pe = importlib.import_module("pincher_env")
    import dial_mpc_b200.envs as E
    from oracle.envs_oracle import CustomRewardOracle
    cfg = pe.PincherEnvConfig()
    env = E.get_environment("pincher_spin", config=cfg)
    tmp = tempfile.NamedTemporaryFile(suffix=".json", delete=False)
    tmp.close()
    env.sys.model.save(tmp.name)
    o = CustomRewardOracle(tmp.name, _pincher_reward_np, user=env.user_params(), joint_range=env.joint_range,
                           dt=cfg.dt, timestep=cfg.timestep, leg_control="position")
    yield env, o
    os.unlink(tmp.name)


def _pincher_actions(rng, B, H):
    """Random finger targets biased towards closing, so that the tips reach the ball (contacts in all
    cone zones: sticking on the floor, sliding under the tips) within a few env steps."""
    return np.clip(rng.normal(size=(B, H, 4)) * 0.4 - 0.1, -1, 1)


def test_dense_custom_env_dimensions_and_variant(pincher):
    from dial_mpc_b200 import _capi, custom
    env, o = pincher
    m = env.sys
    assert (m.nq, m.nv, m.nu, m.nbody) == (11, 10, 4, 7) and env._n_frames == 4
    assert custom.dense_nv(env.sys.model) == 10 and custom.solver_variant(env.sys.model) == 3
    assert "_v3n10_" in custom.library_path(env.reward_source, 3, 10)
    assert "_v3_" in custom.library_path(env.reward_source, 3, 22)          # the stock instantiation keeps its name
    d = env.plan_desc(Nsample=8, Hsample=10, Hnode=4)
    assert d.env_id == _capi.ENV_IDS["custom"] and d.n_user == 7 and d.n_frames == 4 and d.leg_control_torque == 0
    np.testing.assert_allclose(env._init_q, o.init_q)
    assert o.m.elliptic and o.m.ncon == 8


def test_dense_custom_env_in_emulator_matches_oracle(pincher):
    """The device code instantiated for nv = 10 (g++ -DDIAL_DENSE_NV=10) + the custom reward against the
    oracle: 4 substeps per env step, elliptic cones, fingertips closing on the ball."""
    from tests.emul import emul
    env, o = pincher
    s = o.reset()
    rng = np.random.default_rng(7)
    us = _pincher_actions(rng, 3, 6)
    rew, q, qd, x = o.rollout(s, us)
    out = emul.rollout(env, env.plan_desc(), s.qpos[0], s.qvel[0], s.qacc_warmstart[0], us=us,
                       defines=("DIAL_DENSE_NV=10",))
    # ball position and finger joints tightly; the ball's quaternion loosely: MJX's line search can stall on a
    # fresh pinch and the unconverged forces then spin the 50 g ball up to 1e3 rad/s, which
    # amplifies fp32 rounding in its orientation (the oracle and the kernel stall at the same substeps)
    sel = [0, 1, 2, 7, 8, 9, 10]
    assert np.abs(out["q"][..., sel] - q[..., sel]).max() < 5e-4
    assert np.abs(out["q"][..., 3:7] - q[..., 3:7]).max() < 5e-3
    assert np.abs(out["xpos"] - x).max() < 5e-4
    assert np.abs(out["rewss"] - rew).max() < 2e-3 * (1 + np.abs(rew).max())
    assert np.abs(rew).max() > 1e-3
    # the fingers did reach the ball: it moved at some point of some rollout
    assert np.abs(q[:, :, :3] - s.qpos[0][:3]).max() > 2e-3


def test_dense_custom_library_builds_for_its_own_nv(pincher, built):
    """nvcc cross-compiles the dense solver for nv = 10 into the custom build; the stock library (nv = 22)
    refuses the model with a message that says so."""
    import ctypes as C
    from dial_mpc_b200 import _capi, custom
    env, _ = pincher
    md = _capi.fill_model_desc(env.sys.model)
    stock = _capi.lib()
    assert stock.dial_solver_variant(md) < 0 and b"nv = 22" in stock.dial_last_error()
    path = env.library_path
    assert os.path.exists(path) and "_v3n10_" in path
    lib = _capi.lib(path)
    assert lib.dial_solver_variant(md) == 3
    assert lib.dial_custom_reward_id().decode() == custom.reward_id(env.reward_source, 3, 10)
    for sym in _capi.EXPORTS:
        assert hasattr(lib, sym)


def _pincher_single_step_setup(n_traj, n_step, seed):
    """(env, oracle) with ONE physics substep per env step and mid-rollout states of the oracle: fingers
    closing on the ball, contacts in every cone zone, including the substeps on which the solver stalls."""
    if EX not in sys.path:
        sys.path.insert(0, EX)
    pe = importlib.import_module("pincher_env")
    import dial_mpc_b200.envs as E
    from oracle.envs_oracle import CustomRewardOracle
    cfg = pe.PincherEnvConfig(dt=0.005, timestep=0.005)
    env = E.get_environment("pincher_spin", config=cfg)
    assert env._n_frames == 1
    tmp = tempfile.NamedTemporaryFile(suffix=".json", delete=False)
    tmp.close()
    env.sys.model.save(tmp.name)
    o = CustomRewardOracle(tmp.name, _pincher_reward_np, user=env.user_params(), joint_range=env.joint_range,
                           dt=cfg.dt, timestep=cfg.timestep, leg_control="position")
    os.unlink(tmp.name)
    rng = np.random.default_rng(seed)
    s = o.reset().tile(n_traj)
    Q, V, W, A = [], [], [], []
    a = _pincher_actions(rng, n_traj, 1)[:, 0]
    for t in range(n_step):
        if t % 4 == 0:
            a = _pincher_actions(rng, n_traj, 1)[:, 0]              # a new finger target every 20 ms
        Q.append(s.qpos.copy()); V.append(s.qvel.copy()); W.append(s.qacc_warmstart.copy()); A.append(a.copy())
        s, _, _ = o.step(s, a)
    Q, V, W, A = (np.concatenate(x, 0) for x in (Q, V, W, A))
    from oracle.envs_oracle import OState
    st = OState(Q, V, W, np.zeros(len(Q), dtype=np.int64), np.zeros(len(Q), dtype=np.int64))
    ns, _, aux = o.step(st, A)
    active = (aux["data"].con_dist < 0).any(-1)
    return env, o, (Q, V, W, A), ns, active


def _single_step_report(eq, ev, ea, active):
    eq, ev, ea = np.array(eq), np.array(ev), np.array(ea)
    return dict(states=len(eq), in_contact=int(active.sum()), qpos_err_max=float(eq.max()), qvel_relerr_max=float(ev.max()),
                qacc_relerr_max=float(ea.max()), qacc_within_1e4=float((ea.max(1) <= 1e-4).mean()))


# One mjx.step from identical (qpos, qvel, ctrl, qacc_warmstart): positions, velocities (relative to 1 + |v|) and
# the solver output qacc (relative to 1 + |qacc|).  Calibrated on the emulator (the same fp32 device code; 48 / 96
# states): qpos 5e-5, qvel 8e-4 on every state; qacc within 1e-4 on 82-96 % of the states and 1-8 % off on the
# rest — the substeps on which MJX's line search stalls: whether the last bracket point "improved"
# is decided by a cost difference at rounding level, fp32 and fp64 then return different iterates of an
# unconverged solve (the integrated velocities still agree: dt * qacc is small against them).  On the B200 (fast-math
# division / sqrt / sincos) the worst velocity error is 1.2e-2, positions 5e-5: the bounds are those of the Allegro
# single-step test (tests/test_gpu_at_size.py), the qacc fractions sit a little below the emulator's.
_PINCHER_TOL = dict(q=5e-4, v=2e-2, a_max=0.5, frac_1e4=0.5, frac_5e3=0.7)


def _check_single_steps(rep, ea):
    t = _PINCHER_TOL
    ea = np.array(ea).max(1)
    rep = dict(rep, qacc_within_5e3=float((ea <= 5e-3).mean()))
    try:
        import json
        os.makedirs(os.path.join(ROOT, "gpurun_out", "parity"), exist_ok=True)
        json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "parity", "single_step_pincher.json"), "w"), indent=1)
    except OSError:
        pass
    assert rep["qpos_err_max"] < t["q"] and rep["qvel_relerr_max"] < t["v"], rep
    assert rep["qacc_relerr_max"] < t["a_max"] and rep["qacc_within_1e4"] >= t["frac_1e4"] and rep["qacc_within_5e3"] >= t["frac_5e3"], rep


def test_dense_custom_env_single_steps_in_emulator():
    from tests.emul import emul
    env, o, (Q, V, W, A), ns, active = _pincher_single_step_setup(4, 12, 5)
    assert active.sum() >= 10                                          # contacts are exercised
    eq, ev, ea = [], [], []
    for i in range(len(Q)):
        out = emul.rollout(env, env.plan_desc(), Q[i], V[i], W[i], us=A[i][None, None, :], defines=("DIAL_DENSE_NV=10",))
        eq.append(np.abs(out["qpos_out"] - ns.qpos[i]))
        ev.append(np.abs(out["qvel_out"] - ns.qvel[i]) / (1 + np.abs(ns.qvel[i])))
        ea.append(np.abs(out["warm_out"] - ns.qacc_warmstart[i]) / (1 + np.abs(ns.qacc_warmstart[i])))
    _check_single_steps(_single_step_report(eq, ev, ea, active), ea)


@pytest.mark.gpu
def test_gpu_dense_custom_env_matches_oracle(pincher, built):
    """The custom build (dense solver instantiated for nv = 10 + the pincher reward) on the GPU: single physics
    steps from >= 100 oracle states (rollouts themselves are chaotic on this scene: MJX's line search stalls on
    fresh pinches and the unconverged forces kick the ball, DESIGN.md 2 — the kernels stall on the same substeps,
    which is what stepping from identical states shows), the first env step of a batch of rollouts, and the
    planner end to end."""
    import torch
    from dial_mpc_b200.core.dial_config import DialConfig
    from dial_mpc_b200.core.dial_core import MBDPI
    from dial_mpc_b200 import random as drandom
    from dial_mpc_b200.envs.base_env import PipelineState, State
    env1, o1, (Q, V, W, A), ns, active = _pincher_single_step_setup(8, 16, 11)
    assert len(Q) >= 100 and active.sum() >= 30
    plan = env1._get_plan()
    eq, ev, ea = [], [], []
    for i in range(len(Q)):
        st = State(PipelineState(plan.f32(Q[i]), plan.f32(V[i]), plan.f32(W[i])), None, 0.0, 0.0, {}, {"step": 0})
        ps, _ = plan.env_step(st, A[i])
        eq.append(np.abs(ps.qpos.cpu().numpy() - ns.qpos[i]))
        ev.append(np.abs(ps.qvel.cpu().numpy() - ns.qvel[i]) / (1 + np.abs(ns.qvel[i])))
        ea.append(np.abs(ps.qacc_warmstart.cpu().numpy() - ns.qacc_warmstart[i]) / (1 + np.abs(ns.qacc_warmstart[i])))
    _check_single_steps(_single_step_report(eq, ev, ea, active), ea)
    # rollouts through the 4-substep env: the first env step of every row (before anything can amplify)
    env, o = pincher
    s = o.reset()
    state = env.reset(drandom.PRNGKey(0))
    np.testing.assert_allclose(state.pipeline_state.qpos.cpu().numpy(), s.qpos[0], atol=1e-6)
    N, Hs, Hn = 64, 8, 4
    args = DialConfig(env_name="pincher_spin", Nsample=N, Hsample=Hs, Hnode=Hn, Ndiffuse=1, temp_sample=0.05,
                      horizon_diffuse_factor=1.0, traj_diffuse_factor=0.5)
    mb = MBDPI(args, env)
    us = _pincher_actions(np.random.default_rng(9), 12, Hs + 1).astype(np.float32)
    rewss, ps = mb.rollout_us_vmap(state, us)
    rew, q, qd, x = o.rollout(s, us.astype(np.float64))
    rg = rewss.cpu().numpy()
    assert np.isfinite(rg).all()
    assert (np.abs(rg[:, 0] - rew[:, 0]) < 2e-3 * (1 + np.abs(rew[:, 0]))).all()
    # the planner runs on this build (weights / Ybar finite, mean row last)
    _, Y, info = mb.reverse_once(state, drandom.PRNGKey(3), torch.zeros(Hn + 1, 4, device="cuda"), mb.sigma_control)
    assert torch.isfinite(Y).all() and torch.isfinite(info["rews"]).all() and info["rews"].shape == (N + 1,)


def test_custom_build_with_robust_line_search(pincher, built):
    """`build_defines = ("DIAL_ROBUST_LS",)` on a custom env: a separate library (its own cache key) with the
    narrowing-bracket line search compiled in; the default build is untouched."""
    from dial_mpc_b200 import _capi, custom
    env, _ = pincher
    default = env.library_path
    path = custom.build_library(env.reward_source, model=env.sys.model, defines=("DIAL_ROBUST_LS",))
    assert path != default and "_v3n10x_" in path and os.path.exists(path) and os.path.exists(default)
    lib = _capi.lib(path)
    assert lib.dial_custom_reward_id().decode() == custom.reward_id(env.reward_source, 3, 10, ("DIAL_ROBUST_LS",))
    assert lib.dial_solver_variant(_capi.fill_model_desc(env.sys.model)) == 3
    assert type(env).build_defines == ()
