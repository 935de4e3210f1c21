"""CPU tests of the KERNEL LOGIC: csrc/dial_device.cuh compiled by g++ and executed by the
lock-step warp emulator (tests/emul) against the fp64 oracle.  The emulator is test
infrastructure only; the GPU tests (test_gpu_parity.py) check the real CUDA build."""
import numpy as np
import pytest

from tests.conftest import make_pair
from tests.emul import emul


@pytest.mark.parametrize("name,H", [("unitree_go2_walk", 12), ("unitree_go2_seq_jump", 12), ("unitree_h1_walk", 10),
                                    ("unitree_h1_loco", 10)])
def test_emulated_rollout_matches_oracle(name, H):
    env, o = make_pair(name)
    s = o.reset()
    s.step[:] = 55          # exercise the ramp / gait phase / second jump stage
    s.stage[:] = 1 if "jump" in name else 0
    rng = np.random.default_rng(1)
    us = np.clip(rng.normal(size=(3, H, env.action_size)) * 0.6, -1, 1)
    rew, q, qd, x = o.rollout(s, us)
    out = emul.rollout(env, env.plan_desc(), s.qpos[0], s.qvel[0], s.qacc_warmstart[0], us=us, step0=55,
                       stage0=int(s.stage[0]))
    # fp32 tolerances: positions 1e-4, velocities 5e-3, rewards 1e-3 relative
    assert np.abs(out["q"] - q).max() < 1e-4
    assert np.abs(out["qd"] - qd).max() < 5e-3
    assert np.abs(out["xpos"] - x).max() < 1e-4
    assert np.abs(out["rewss"] - rew).max() < 1e-3 * (1 + np.abs(rew).max())


def test_emulated_pipeline_init_and_env_step():
    env, o = make_pair("unitree_go2_walk")
    s = o.reset()
    out = emul.rollout(env, env.plan_desc(), o.init_q, np.zeros(18), np.zeros(18), mode=2, nrows=1, H=1)
    assert np.abs(out["qpos_out"] - s.qpos[0]).max() < 1e-6
    assert np.abs(out["warm_out"] - s.qacc_warmstart[0]).max() < 2e-3
    a = np.random.default_rng(0).uniform(-1, 1, (1, 1, 12))
    ns, r, aux = o.step(s, a[:, 0])
    out = emul.rollout(env, env.plan_desc(), s.qpos[0], s.qvel[0], s.qacc_warmstart[0], us=a)
    assert np.abs(out["qpos_out"] - ns.qpos[0]).max() < 1e-5
    assert np.abs(out["ctrl_out"] - aux["ctrl"][0]).max() < 1e-4
    assert abs(out["rewss"][0, 0] - r[0]) < 1e-4


def test_emulated_planner_rows_and_native_rng():
    """mode 1: Y0s construction (pin node 0, mean row, clip), spline, shard offsets, and the
    in-kernel Threefry/erfinv sampler against the oracle's restatement of jax.random.normal."""
    from oracle.planner_oracle import PlannerOracle, jax_normal_legacy
    env, o = make_pair("unitree_go2_walk")
    N, Hs, Hn = 6, 8, 4
    pl = PlannerOracle(o, N, Hs, Hn, 0.05, 0.9, 0.5)
    s = o.reset()
    key = (123, 456)
    eps = jax_normal_legacy(key, (N, Hn + 1, 12))
    Ybar = np.clip(np.random.default_rng(0).standard_normal((Hn + 1, 12)) * 0.5, -1.5, 1.5)
    Yo, info = pl.reverse_once(s, eps, Ybar, pl.sigma_control)
    from dial_mpc_b200.utils.spline import interp_matrix
    # rank 1 of 2: rows 3..5 + mean row, native RNG
    desc = env.plan_desc(Nsample=3, Ntotal=N, shard_offset=3, Hsample=Hs, Hnode=Hn, temp_sample=0.05,
                         M_n2u=interp_matrix(pl.step_nodes, pl.step_us))
    out = emul.rollout(env, desc, s.qpos[0], s.qvel[0], s.qacc_warmstart[0], Ybar=Ybar, noise=pl.sigma_control,
                       key=key, mode=1, nrows=4, H=Hs + 1)
    ref = np.concatenate([info["rews"][3:6], info["rews"][-1:]])
    assert np.abs(out["rews"] - ref).max() < 5e-4
    # injected eps gives the same rows
    out2 = emul.rollout(env, desc, s.qpos[0], s.qvel[0], s.qacc_warmstart[0], Ybar=Ybar, noise=pl.sigma_control,
                        eps=eps, mode=1, nrows=4, H=Hs + 1)
    assert np.abs(out2["rews"] - out["rews"]).max() < 1e-4


def test_emulated_planner_extreme_shapes():
    """Capacity edges of the planner rows: one sample (+ the mean row), the longest horizon and
    the most knots the fixed-size plan allows (DIAL_MAXH = 64 steps, DIAL_MAXNODE = 8 knots)."""
    from dial_mpc_b200 import _capi
    from dial_mpc_b200.utils.spline import interp_matrix
    from oracle.planner_oracle import PlannerOracle, jax_normal_legacy
    env, o = make_pair("unitree_go2_walk")
    N, Hs, Hn = 1, _capi.DEFINES["DIAL_MAXH"] - 1, _capi.DEFINES["DIAL_MAXNODE"] - 1
    pl = PlannerOracle(o, N, Hs, Hn, 0.05, 0.9, 0.5)
    s = o.reset()
    key = (7, 9)
    eps = jax_normal_legacy(key, (N, Hn + 1, 12))
    Ybar = np.zeros((Hn + 1, 12))
    Yo, info = pl.reverse_once(s, eps, Ybar, pl.sigma_control * 0.3)
    desc = env.plan_desc(Nsample=N, Hsample=Hs, Hnode=Hn, temp_sample=0.05,
                         M_n2u=interp_matrix(pl.step_nodes, pl.step_us))
    out = emul.rollout(env, desc, s.qpos[0], s.qvel[0], s.qacc_warmstart[0], Ybar=Ybar,
                       noise=pl.sigma_control * 0.3, key=key, mode=1, nrows=N + 1, H=Hs + 1)
    # 64 steps of contact dynamics: fp32 vs fp64 drift is visible but bounded
    assert np.abs(out["rews"] - info["rews"]).max() < 2e-2 * (1 + np.abs(info["rews"]).max())
    assert np.isfinite(out["q"]).all() and out["q"].shape == (2, Hs + 1, 19)
    with pytest.raises(ValueError):
        env.plan_desc(Nsample=1, Hsample=Hs + 1, Hnode=Hn)
    with pytest.raises(ValueError):
        env.plan_desc(Nsample=1, Hsample=Hs, Hnode=Hn + 1)


@pytest.mark.parametrize("name", ["unitree_go2_walk", "unitree_h1_walk", "unitree_h1_loco"])
def test_emulated_random_states_match_oracle(name):
    """States away from the nominal trajectory: tilted base, base height in and out of contact,
    joints pushed beyond their limits (limit rows active), joint rates up to tens of rad/s."""
    from oracle.envs_oracle import OState
    env, o = make_pair(name)
    s0 = o.reset()
    nv, nu = o.m.nv, o.m.nu
    rng = np.random.default_rng(11)
    for _ in range(4):
        q = s0.qpos[0].copy()
        q[2] += rng.uniform(-0.06, 0.15)
        ax = rng.normal(size=3)
        ax /= np.linalg.norm(ax)
        ang = rng.uniform(0, 0.5)
        q[3:7] = [np.cos(ang / 2), *(np.sin(ang / 2) * ax)]
        lo, hi = o.physical_joint_range[:, 0], o.physical_joint_range[:, 1]
        q[7:7 + nu] = np.clip(q[7:7 + nu] + rng.normal(size=nu) * 0.4, lo - 0.05, hi + 0.05)
        v = rng.normal(size=nv) * np.r_[np.ones(3) * 0.5, np.ones(3), np.ones(nv - 6) * 3.0]
        st = OState(q[None], v[None], np.zeros((1, nv)), np.array([17]), np.array([0]))
        us = np.clip(rng.normal(size=(1, 3, nu)), -1, 1)
        rew, qq, qd, x = o.rollout(st, us)
        out = emul.rollout(env, env.plan_desc(), q, v, np.zeros(nv), us=us, step0=17)
        assert np.abs(out["q"] - qq).max() < 5e-5
        assert np.abs(out["qd"] - qd).max() < 2e-3 * (1 + np.abs(qd).max() / 10)
        assert np.abs(out["rewss"] - rew).max() < 1e-4 * (1 + np.abs(rew).max())


@pytest.mark.parametrize("name", ["unitree_go2_walk", "unitree_h1_loco"])
def test_emulated_command_override_matches_oracle(name):
    """randomize_tasks: the env step whose info["step"] equals cmd_step uses the random command
    (unitree_go2_env.py:141-163); the others keep the default one."""
    env, o = make_pair(name)
    s = o.reset()
    s.step[:] = 497
    rng = np.random.default_rng(4)
    H = 6
    us = np.clip(rng.normal(size=(2, H, env.action_size)) * 0.5, -1, 1)
    vel, ang = np.array([-1.2, 0.4, 0.0]), np.array([0.0, 0.0, 1.1])
    base, *_ = o.rollout(s, us)
    o.cmd_override = (500, vel, ang)
    rew, q, qd, x = o.rollout(s, us)
    o.cmd_override = None
    assert np.abs(rew[:, 3] - base[:, 3]).min() > 1e-2 and np.abs(np.delete(rew - base, 3, axis=1)).max() == 0
    d = env.plan_desc()
    d.cmd_step = 500
    d.cmd_vel[:], d.cmd_ang[:] = list(vel), list(ang)
    out = emul.rollout(env, d, s.qpos[0], s.qvel[0], s.qacc_warmstart[0], us=us, step0=497)
    assert np.abs(out["rewss"] - rew).max() < 1e-3 * (1 + np.abs(rew).max())
    assert np.abs(out["q"] - q).max() < 1e-4
