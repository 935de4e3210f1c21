"""ORACLE (test infrastructure) — ctypes driver of the C port of the per-sample step
(oracle/c/dial_port.c): the fp32 CPU baseline of bench.py and, built in double, a second
restatement that must agree with the NumPy oracle to rounding (tests/test_c_port.py).

The C structs are the public ones of include/dial_b200.h (the shared contract); they are filled
here from the ORACLE's own model container and env objects, not from the product package's."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build_oracle
from .envs_oracle import Go2SeqJumpOracle, Go2WalkOracle, H1LocoOracle, H1WalkOracle, OracleEnv, OState

_LIBS = {}


def available() -> bool:
    return os.path.exists(build_oracle.lib_path("float")) and os.path.exists(build_oracle.lib_path("double"))


def _structs():
    from dial_mpc_b200 import _capi   # struct classes generated from include/dial_b200.h (no library is loaded)
    return _capi


def _lib(real: str):
    if real not in _LIBS:
        path = build_oracle.lib_path(real)
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: python oracle/build_oracle.py")
        lib = C.CDLL(path)
        lib.port_sizeof_real.restype = C.c_int
        P = C.c_void_p
        lib.port_rollout.argtypes = [P, P, C.c_int, C.c_int, P, P, P, C.c_int, C.c_int, P, P, P, P, P, P]
        lib.port_rollout.restype = C.c_int
        lib.port_rollout_mt.argtypes = [P, P, C.c_int, C.c_int, P, P, P, C.c_int, C.c_int, P, P, P, P, P, P, C.c_int]
        lib.port_rollout_mt.restype = C.c_int
        assert lib.port_sizeof_real() == (4 if real == "float" else 8)
        _LIBS[real] = lib
    return _LIBS[real]


def fill_model(m) -> "C.Structure":
    cap = _structs()
    d = cap.dial_model_desc()
    for k in ("nq", "nv", "nu", "nbody", "njnt", "ngeom", "nsite", "npair", "ncon", "iterations", "ls_iterations", "cone"):
        setattr(d, k, int(getattr(m, k)))
    d.eulerdamp = int(bool(m.eulerdamp))
    for k in ("timestep", "tolerance", "ls_tolerance", "impratio", "meaninertia"):
        setattr(d, k, float(getattr(m, k)))
    cap._set(d.gravity, m.gravity)
    for k in ("body_parentid", "body_rootid", "body_depth", "body_jntadr", "body_dofadr", "body_dofnum",
              "body_pos", "body_quat", "body_ipos", "body_iquat", "body_mass", "body_inertia",
              "jnt_type", "jnt_qposadr", "jnt_dofadr", "jnt_limited", "jnt_pos", "jnt_axis", "jnt_range",
              "jnt_margin", "jnt_solref", "jnt_solimp",
              "dof_bodyid", "dof_jntid", "dof_parentid", "dof_armature", "dof_damping", "dof_invweight0",
              "qpos0", "geom_type", "geom_bodyid", "geom_pos", "geom_quat", "geom_size",
              "pair_kind", "pair_geom1", "pair_geom2", "pair_ncon", "pair_condim", "pair_friction", "pair_margin",
              "pair_gap", "pair_solref", "pair_solimp", "site_bodyid", "site_pos",
              "actuator_dofadr", "actuator_qposadr", "actuator_ctrllimited", "actuator_forcelimited",
              "actuator_gear", "actuator_gain", "actuator_bias", "actuator_ctrlrange", "actuator_forcerange"):
        cap._set(getattr(d, k), getattr(m, k))
    cap._set(d.body_invweight0, m.body_invweight0[:, 0])
    cap._set(d.body_invweight0_rot, m.body_invweight0[:, 1])
    return d


def fill_plan(env: OracleEnv) -> "C.Structure":
    cap = _structs()
    d = cap.dial_plan_desc()
    ids = cap.ENV_IDS
    d.env_id = (ids["unitree_go2_seq_jump"] if isinstance(env, Go2SeqJumpOracle) else ids["unitree_go2_walk"] if isinstance(env, Go2WalkOracle)
                else ids["unitree_h1_loco"] if isinstance(env, H1LocoOracle) else ids["unitree_h1_walk"] if isinstance(env, H1WalkOracle) else -1)
    if d.env_id < 0:
        raise NotImplementedError("the C port covers the tree / pyramidal envs (Go2, H1) only")
    d.Nsample = d.Ntotal = 1
    d.Hsample, d.Hnode = 1, 1
    d.n_frames = env.n_frames
    d.leg_control_torque = int(env.leg_control == "torque")
    d.temp_sample, d.dt, d.action_scale = 1.0, float(env.dt), float(env.action_scale)
    nu = env.nu
    cap._set(d.kp, np.broadcast_to(env.kp, (nu,)))
    cap._set(d.kd, np.broadcast_to(env.kd, (nu,)))
    cap._set(d.joint_range, env.joint_range)
    cap._set(d.physical_joint_range, env.physical_joint_range)
    cap._set(d.joint_torque_range, np.clip(env.joint_torque_range, -3e38, 3e38))
    d.torso_body = int(env.torso) + 1
    d.ramp_up_time = float(getattr(env, "ramp_up_time", 1.0))
    cap._set(d.vel_cmd, getattr(env, "vel_cmd", np.zeros(3)))
    cap._set(d.ang_cmd, getattr(env, "ang_cmd", np.zeros(3)))
    cap._set(d.pos_tar, env.pos_tar)
    ov = getattr(env, "cmd_override", None)
    d.cmd_step = -1 if ov is None else int(ov[0])
    if ov is not None:
        cap._set(d.cmd_vel, ov[1]); cap._set(d.cmd_ang, ov[2])
    duty, cad, amp = env.GAIT_PARAMS[env.gait]
    d.gait_duty, d.gait_cadence, d.gait_amplitude = float(duty), float(cad), float(amp)
    ph = env.GAIT_PHASE[env.gait]
    d.nfeet = len(ph)
    cap._set(d.gait_phase, np.asarray(ph, dtype=np.float64))
    if hasattr(env, "feet_site"):
        cap._set(d.feet_site, np.asarray(env.feet_site, dtype=np.int32))
    if isinstance(env, Go2SeqJumpOracle):
        n = env.pose_seq.shape[0]
        d.n_stage, d.jump_dt = n, float(env.jump_dt)
        cap._set(d.pose_seq, env.pose_seq)
        cap._set(d.yaw_seq, env.yaw_seq)
        cap._set(d.contact_targets, env.contact_targets)
        cap._set(d.contact_radius, env.contact_radius)
    return d


class CPort:
    """rollout_us of one oracle env through the C port.  real = "float" (baseline) or "double"."""

    def __init__(self, env: OracleEnv, _cfg=None, real: str = "float"):
        self.env, self.real = env, real
        self.lib = _lib(real)
        self.md, self.pd = fill_model(env.m), fill_plan(env)

    def rollout(self, s: OState, us, want_traj=True, threads: int = 1):
        us = np.ascontiguousarray(us, dtype=np.float64)
        B, H, nu = us.shape
        m = self.env.m
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        q0, v0, w0 = f(s.qpos[0]), f(s.qvel[0]), f(s.qacc_warmstart[0])
        rew = np.zeros((B, H))
        q = np.zeros((B, H, m.nq)) if want_traj else None
        qd = np.zeros((B, H, m.nv)) if want_traj else None
        x = np.zeros((B, H, m.nbody - 1, 3)) if want_traj else None
        warm = np.zeros((B, m.nv))
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        rc = self.lib.port_rollout_mt(C.byref(self.md), C.byref(self.pd), B, H, p(q0), p(v0), p(w0), int(s.step[0]), int(s.stage[0]),
                                      p(us), p(rew), p(q), p(qd), p(x), p(warm), int(threads))
        if rc != 0:
            raise RuntimeError(f"port_rollout failed ({rc}): model / env outside the C port's scope")
        return rew, q, qd, x, warm

    def rollout_rews(self, s: OState, us, threads: int = 1):
        """Per-sample mean rewards; ``threads`` OpenMP threads over the rows (the CPU baseline of bench.py)."""
        return self.rollout(s, us, want_traj=False, threads=threads)[0].mean(-1)
