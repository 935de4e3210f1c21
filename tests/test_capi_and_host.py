"""CPU tests: the C-ABI library loads and exports every symbol the header declares (no compute
calls without a GPU), struct layouts agree, and the host-side logic mirrors the reference API."""
import ctypes as C
import dataclasses
import re

import numpy as np
import pytest
import yaml

from dial_mpc_b200 import _capi
from tests.conftest import ENV_CASES


def test_library_exports_every_declared_symbol(built):
    lib = _capi.lib()
    text = open(_capi.HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(dial_\w+)\s*\(", text))
    assert declared == set(_capi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.dial_abi_version() == _capi.DEFINES["DIAL_ABI_VERSION"]
    for i, t in enumerate((_capi.dial_model_desc, _capi.dial_plan_desc, _capi.dial_state)):
        assert lib.dial_sizeof(i) == C.sizeof(t)


def test_key_split_matches_oracle(built):
    from dial_mpc_b200 import random as drandom
    from oracle.planner_oracle import jax_split_legacy
    for seed in (0, 1, 12345):
        key = drandom.PRNGKey(seed)
        a, b = drandom.split(key)
        ref = jax_split_legacy((int(key[0]), int(key[1])))
        assert tuple(a) == tuple(ref[0]) and tuple(b) == tuple(ref[1])


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_capi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _capi.lib()


def test_config_dataclasses_match_reference_fields():
    from dial_mpc_b200.config.base_env_config import BaseEnvConfig
    from dial_mpc_b200.core.dial_config import DialConfig
    assert [f.name for f in dataclasses.fields(DialConfig)] == [
        "seed", "output_dir", "n_steps", "env_name", "Nsample", "Hsample", "Hnode", "Ndiffuse",
        "Ndiffuse_init", "temp_sample", "horizon_diffuse_factor", "traj_diffuse_factor",
        "update_method", "sigma_scale"]
    d = DialConfig()
    assert (d.Nsample, d.Hsample, d.Hnode, d.Ndiffuse, d.Ndiffuse_init, d.temp_sample) == (2048, 16, 4, 2, 10, 0.06)
    assert [f.name for f in dataclasses.fields(BaseEnvConfig)] == [
        "task_name", "randomize_tasks", "kp", "kd", "debug", "dt", "timestep", "backend",
        "leg_control", "action_scale"]


def test_registry_and_yaml_loading():
    import dial_mpc_b200.envs as E
    from dial_mpc_b200.core.dial_config import DialConfig
    from dial_mpc_b200.examples import examples
    from dial_mpc_b200.utils.io_utils import get_example_path, load_dataclass_from_dict
    for ex in examples:
        cfg = yaml.safe_load(open(get_example_path(ex + ".yaml")))
        dc = load_dataclass_from_dict(DialConfig, cfg)
        ec = load_dataclass_from_dict(E.get_config(dc.env_name), cfg, convert_list_to_array=True)
        env = E.get_environment(dc.env_name, config=ec)
        assert env.action_size == env.sys.nu and abs(env.dt - 0.02) < 1e-12
        d = env.plan_desc(Nsample=8, Hsample=dc.Hsample, Hnode=dc.Hnode)
        assert d.env_id == _capi.ENV_IDS[dc.env_name] and d.n_frames == int(round(cfg["dt"] / cfg["timestep"]))
    E.register_config("custom", E.UnitreeGo2EnvConfig)
    assert E.get_config("custom") is E.UnitreeGo2EnvConfig

    class MyEnv(E.UnitreeGo2Env):
        pass
    E.register_environment("custom", MyEnv)
    assert isinstance(E.get_environment("custom", config=E.UnitreeGo2EnvConfig()), MyEnv)


def test_act2joint_act2tau_match_oracle():
    from tests.conftest import make_pair
    for name in ("unitree_go2_walk", "unitree_h1_walk"):
        env, o = make_pair(name)
        rng = np.random.default_rng(0)
        a = rng.uniform(-1.2, 1.2, env.action_size)
        assert np.abs(env.act2joint(a) - o.act2joint(a)).max() < 1e-12
        s = o.reset()

        class PS:
            qpos, qvel = s.qpos[0], s.qvel[0] + 0.1
        assert np.abs(env.act2tau(a, PS) - o.act2tau(a[None], s.qpos, s.qvel + 0.1)[0]).max() < 1e-9
    env, _ = make_pair("unitree_go2_walk")
    assert np.isinf(env.joint_torque_range).all()       # Go2 motors have no ctrlrange
    d = env.plan_desc()
    assert np.isfinite(np.ctypeslib.as_array(d.joint_torque_range)).all()


def test_seq_jump_targets_match_oracle():
    from tests.conftest import make_pair
    env, o = make_pair("unitree_go2_seq_jump")
    assert np.abs(env._contact_targets - o.contact_targets).max() < 1e-12
    assert env._contact_targets.shape == (5, 4, 3)
    yaw = np.array([0.0, 0.3, -0.2])
    pos = np.array([[0, 0, 0.27], [0.4, 0.1, 0.27], [0.8, 0, 0.27]])
    t, r, p, y = env.generate_jumping_sequence(pos, yaw, 0.1)
    c, s = np.cos(0.3), np.sin(0.3)
    assert np.allclose(t[1, 0, :2], pos[1, :2] + np.array([c * 0.2 + s * 0.135, s * 0.2 - c * 0.135]))


def test_mbdpi_host_math_without_gpu():
    """MBDPI's schedule / shift / spline maps on CPU through the test harness plan."""
    import torch
    from dial_mpc_b200.core.dial_config import DialConfig
    from dial_mpc_b200.core.dial_core import MBDPI
    from oracle.planner_oracle import PlannerOracle
    from tests.conftest import make_pair
    from tests.emul.emul import EmulPlan
    env, o = make_pair("unitree_go2_walk")
    cfg = DialConfig(env_name="unitree_go2_walk", Nsample=4, Hsample=16, Hnode=4, temp_sample=0.05)
    mb = MBDPI(cfg, env, plan_factory=EmulPlan)
    po = PlannerOracle(o, 4, 16, 4, 0.05, 0.9, 0.5)
    assert np.abs(mb.sigma_control.numpy() - po.sigma_control).max() < 1e-6
    assert np.abs(mb.schedule(3).numpy() - po.schedule(3)).max() < 1e-6
    Y = np.random.default_rng(0).standard_normal((5, 12))
    assert np.abs(mb.shift(Y).numpy() - po.shift(Y)).max() < 1e-5
    assert np.abs(mb.node2u_vmap(Y).numpy() - po.node2u(Y)).max() < 1e-5
    u = np.random.default_rng(1).standard_normal((3, 17, 12))
    assert np.abs(mb.u2node_vvmap(u).numpy() - po.u2node(u)).max() < 1e-5
    assert mb.ctrl_dt == 0.02 and abs(mb.node_dt - 0.08) < 1e-12 and mb.nu == 12
    sh = mb.shift_Y_from_u(torch.ones(17, 12), 2)
    assert sh.shape == (5, 12) and abs(float(sh[-1, 0])) < 1e-5


def test_deploy_shm_layout_and_time_shift():
    """deploy/dial_plan.py: segment sizes follow the reference (x32 over-allocation),
    and the plan time-shift is the node spline evaluated at shifted node times."""
    from dial_mpc_b200.deploy.dial_plan import shm_layout
    from dial_mpc_b200.utils.spline import interp_matrix
    lay = shm_layout(n_acts=17, nx=37, nu=12)
    assert lay["acts_shm"] == ((17, 12), 17 * 12 * 32) and lay["refs_shm"] == ((17, 12, 3), 17 * 12 * 3 * 32)
    assert lay["state_shm"] == ((37,), 37 * 32) and lay["time_shm"][1] == 32 and lay["plan_time_shm"][1] == 32
    nodes = np.linspace(0, 0.32, 5)
    M = interp_matrix(nodes, nodes + 0.08)               # one node period: node k+1 moves onto node k
    assert np.abs(M[:4] - np.eye(5)[1:]).max() < 1e-12
    y = np.random.default_rng(0).standard_normal(5)
    from scipy.interpolate import InterpolatedUnivariateSpline
    for dt in (0.013, 0.02, 0.055):
        assert np.abs(interp_matrix(nodes, nodes + dt) @ y - InterpolatedUnivariateSpline(nodes, y, k=2)(nodes + dt)).max() < 1e-12


def test_end_of_run_dumps_have_the_reference_layout(tmp_path):
    """dial_core.py:305-323: `*_states.npy` rows [i, qpos, qvel, ctrl]; `*_predictions.npy` holds per
    control step the xbar of the LAST diffusion iteration over the whole horizon
    ([n_steps, Hsample+1, nbody-1, 3]; `infos[i]["xbar"][-1]` indexes the scan's iteration axis)."""
    import numpy as np
    import torch
    from dial_mpc_b200.core.dial_core import save_run
    n_steps, Hs, nq, nv, nu, nb = 7, 16, 19, 18, 12, 14
    rollout = [torch.cat([torch.tensor([float(t)]), torch.zeros(nq), torch.ones(nv), torch.full((nu,), 2.0)]) for t in range(n_steps)]
    infos = [torch.full((Hs + 1, nb - 1, 3), float(t)) for t in range(n_steps)]
    states, preds = save_run(str(tmp_path), rollout, infos, timestamp="t")
    assert states.shape == (n_steps, 1 + nq + nv + nu) and preds.shape == (n_steps, Hs + 1, nb - 1, 3)
    assert np.load(tmp_path / "t_states.npy").shape == states.shape
    assert np.load(tmp_path / "t_predictions.npy").shape == preds.shape
    assert (states[:, 0] == np.arange(n_steps)).all() and (preds[3] == 3.0).all()


def test_host_random_pieces_for_sample_command():
    """dial_mpc_b200.random: Threefry-2x32 against the Random123 known answers, split_n(key, 2)
    against the C library's split, uniform1 in range and reproducible."""
    from dial_mpc_b200 import random as R
    kat = [((0, 0), (0, 0), (0x6B200159, 0x99BA4EFE)),
           ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
           ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]
    for key, ctr, out in kat:
        y0, y1 = R.threefry2x32(key, [ctr[0]], [ctr[1]])
        assert (int(y0[0]), int(y1[0])) == out
    k = R.PRNGKey(11)
    a, b = R.split(k)
    s2 = R.split_n(k, 2)
    assert np.array_equal(s2[0], a) and np.array_equal(s2[1], b)
    assert R.split_n(k, 4).shape == (4, 2) and len({tuple(r) for r in R.split_n(k, 4)}) == 4
    u = [R.uniform1(kk, -1.5, 1.5) for kk in R.split_n(k, 64)]
    assert all(-1.5 <= v < 1.5 for v in u) and np.std(u) > 0.5 and u[0].dtype == np.float32
    # the oracle's restatement of jax.random.uniform agrees bit for bit
    from oracle import planner_oracle as J
    for kk in R.split_n(k, 8):
        assert R.uniform1(kk, -0.5, 0.5) == J.jax_uniform_legacy(kk, (1,), -0.5, 0.5)[0]
    assert np.array_equal(R.split_n(k, 4), J.jax_split_legacy(k, 4))


def test_command_override_follows_the_env_key_chain():
    """BaseEnv.command_override: the random command of the next step % 500 == 0 comes from
    split(R_hit)[1] where R advances by one split per env step (unitree_go2_env.py:127,141-155), so
    looking ahead from any earlier state gives the same command as arriving there."""
    import dial_mpc_b200.envs as E
    from dial_mpc_b200 import random as R
    cfg = E.get_config("unitree_go2_walk")(randomize_tasks=True)
    env = E.get_environment("unitree_go2_walk", config=cfg)
    rng = R.PRNGKey(3)
    info = {"randomize_target": True, "step": 493, "rng": rng}
    assert env.command_override(info, 5) is None              # steps 493..497
    ov = env.command_override(info, 8)                         # reaches step 500
    assert ov is not None and ov[0] == 500
    r = rng
    for _ in range(7):
        r = R.split(r)[0]
    at = env.command_override({"randomize_target": True, "step": 500, "rng": r}, 1)
    assert at[0] == 500 and np.array_equal(at[1], ov[1]) and np.array_equal(at[2], ov[2])
    vel, ang = env.sample_command(R.split(r)[1])
    assert np.array_equal(vel, ov[1]) and np.array_equal(ang, ov[2])
    from oracle.planner_oracle import sample_command_oracle
    vo, ao = sample_command_oracle(R.split(r)[1])
    assert np.array_equal(vo.astype(np.float32), vel) and np.array_equal(ao.astype(np.float32), ang)
    assert -1.5 <= vel[0] < 1.5 and -0.5 <= vel[1] < 0.5 and vel[2] == 0 and ang[0] == ang[1] == 0
    # info bookkeeping: the step that used the command stores it ramped (ramp >= 1 here)
    nxt = env._next_info({"randomize_target": True, "step": 500, "rng": r, "vel_tar": np.zeros(3), "ang_vel_tar": np.zeros(3)})
    ramp = np.float32(500) * np.float32(env.dt) / np.float32(cfg.ramp_up_time)
    assert np.allclose(nxt["vel_tar"], np.minimum(vel * ramp, vel))
    assert env.command_override({"randomize_target": False, "step": 500, "rng": r}, 1) is None


def test_seq_jump_random_sequence_matches_oracle():
    """randomize_tasks of UnitreeGo2SeqJumpEnv (unitree_go2_env.py:383-394, 594-631): reset draws an
    11-stage jump sequence.  Host sampler against the oracle's restatement of jax.random, the stage
    tables against the oracle env built from the same sequence, and a rollout across a stage change
    through the kernel logic (warp emulator) with those tables."""
    import dial_mpc_b200.envs as E
    from dial_mpc_b200 import _capi, random as R
    from oracle import planner_oracle as J
    from oracle.envs_oracle import make_env
    from tests.emul import emul
    k = R.PRNGKey(5)
    for kk in R.split_n(k, 6):
        assert np.array_equal(R.uniform(kk, 2, -0.65, 0.65), J.jax_uniform_legacy(kk, (2,), -0.65, 0.65))
        assert R.uniform(kk, 1, -0.5, 0.5)[0] == R.uniform1(kk, -0.5, 0.5)
        assert np.array_equal(R.uniform(kk, 5, 0.0, 1.0), J.jax_uniform_legacy(kk, (5,), 0.0, 1.0))
    name = "unitree_go2_seq_jump"
    cfg = dict(ENV_CASES[name])
    env = E.get_environment(name, config=E.get_config(name)(randomize_tasks=True, **{
        kk: (np.array(v) if isinstance(v, list) else v) for kk, v in cfg.items()}))
    rng = R.split(k)[0]
    info = env._init_info(rng)
    assert info["randomize_target"] and info["contact_targets"].shape == (11, 4, 3)
    com_pos, com_yaw = J.sample_jump_sequence_oracle(rng)
    assert com_pos.shape == (11, 3) and np.abs(np.diff(com_pos[:, :2], axis=0)).max() < 0.65 and (com_pos[:, 2] == np.float32(0.27)).all()
    assert np.abs(np.diff(com_yaw)).max() < 0.5 and np.abs(np.diff(com_yaw)).min() > 0
    o = make_env(name, dict(cfg, pose_target_sequence=com_pos, yaw_target_sequence=com_yaw))
    assert np.abs(info["pose_target_sequence"] - com_pos).max() == 0 and np.abs(info["yaw_target_sequence"] - com_yaw).max() == 0
    assert np.abs(info["contact_targets"] - o.contact_targets).max() < 1e-6
    assert np.abs(info["contact_target_radius"] - o.contact_radius).max() == 0
    assert env.command_override(info, 1000) is None
    # a different key gives a different sequence; the configured one is untouched
    assert np.abs(env._init_info(R.split(rng)[0])["pose_target_sequence"] - com_pos).max() > 1e-2
    assert len(env._contact_targets) == 5
    # stage bookkeeping runs over the drawn sequence (11 stages), not the configured 5
    nxt = env._next_info(dict(info, step=399, vel_tar=np.zeros(3), ang_vel_tar=np.zeros(3)))
    assert nxt["contact_stage"] == 8
    # kernel logic with the drawn tables: rollout across the stage change at step 50
    d = env.plan_desc()
    pose, yaw, tgt, rad = env.stage_tables(info)
    d.n_stage = 11
    _capi._set(d.pose_seq, pose); _capi._set(d.yaw_seq, yaw); _capi._set(d.contact_targets, tgt); _capi._set(d.contact_radius, rad)
    s = o.reset()
    s.step[:] = 46
    us = np.clip(np.random.default_rng(2).normal(size=(2, 8, 12)) * 0.5, -1, 1)
    rew, q, qd, x = o.rollout(s, us)
    out = emul.rollout(env, d, s.qpos[0], s.qvel[0], s.qacc_warmstart[0], us=us, step0=46, stage0=0)
    assert np.abs(out["q"] - q).max() < 1e-4
    assert np.abs(out["rewss"] - rew).max() < 1e-3 * (1 + np.abs(rew).max())
    # ... and they matter: the configured sequence gives other rewards after the stage change
    base = emul.rollout(env, env.plan_desc(), s.qpos[0], s.qvel[0], s.qacc_warmstart[0], us=us, step0=46, stage0=0)
    assert np.abs(base["rewss"][:, 5:] - out["rewss"][:, 5:]).max() > 1e-3


def test_bench_issue_roofline_block():
    """bench.py's `roofline.issue`: achieved IPC per SM from ncu's instruction count and the kernel time, ceiling =
    resident warps / measured cycles per instruction of a warp streaming non-resident code."""
    import json
    import os
    import bench
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    probe = json.load(open(os.path.join(root, "profiles", "r02_icache_probe.json")))
    r = bench.issue_roofline(8494, 2049, 26, 0.691e-3, 1965.0, 148, probe)
    assert abs(r["ipc_per_sm"] - 2.25) < 0.01 and abs(r["resident_warps_per_sm"] - 2049 / 148) < 1e-9
    assert 0.9 < r["frac"] < 1.0 and abs(r["ipc_ceiling_streaming_code"] - 13.84 / 5.841) < 0.01
    r4 = bench.issue_roofline(8644, 8193, 26, 2.730e-3, 1965.0, 148, probe)           # 4 waves of 14 warps
    assert r4["resident_warps_per_sm"] == 14 and 0.9 < r4["frac"] < 1.0
    r0 = bench.issue_roofline(8775, 129, 17, 0.345e-3, 1965.0, 148, probe)            # one warp per SM: no ceiling claimed
    assert r0["ipc_ceiling_streaming_code"] is None and r0["frac"] is None and r0["ipc_per_sm"] > 0
    assert bench.issue_roofline(None, 1, 1, 1.0, 1965.0, 148, probe) is None
    assert bench.issue_roofline(8494, 2049, 26, 0.691e-3, 1965.0, 148, {})["frac"] is None
