"""Unitree Go2 environments on the CUDA sampling core.

Same classes / config fields / registry names as the reference
(dial_mpc/envs/unitree_go2_env.py): ``UnitreeGo2Env`` (:36-315) and
``UnitreeGo2SeqJumpEnv`` (:327-646).  The reward code of ``step`` is fused into the
rollout kernel (csrc/dial_device.cuh: reward_lane0); this module only prepares the
constants it needs.  ``UnitreeGo2CrateEnv`` is out of scope.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, Union

import numpy as np

from dial_mpc_b200 import _capi
from dial_mpc_b200.envs.base_env import BaseEnv, BaseEnvConfig, System
from dial_mpc_b200.modelc import CompiledModel
from dial_mpc_b200.utils.io_utils import get_model_path


@dataclass
class UnitreeGo2EnvConfig(BaseEnvConfig):
    kp: Union[float, Any] = 30.0
    kd: Union[float, Any] = 0.0
    default_vx: float = 1.0
    default_vy: float = 0.0
    default_vyaw: float = 0.0
    ramp_up_time: float = 2.0
    gait: str = "trot"


class UnitreeGo2Env(BaseEnv):
    supports_randomize_tasks = True
    env_id = _capi.ENV_IDS["unitree_go2_walk"]

    def __init__(self, config: UnitreeGo2EnvConfig):
        super().__init__(config)
        self._foot_radius = 0.0175
        self._gait = config.gait
        self._gait_phase = {
            "stand": np.zeros(4),
            "walk": np.array([0.0, 0.5, 0.75, 0.25]),
            "trot": np.array([0.0, 0.5, 0.5, 0.0]),
            "canter": np.array([0.0, 0.33, 0.33, 0.66]),
            "gallop": np.array([0.0, 0.05, 0.4, 0.35]),
        }
        self._gait_params = {
            #                  ratio, cadence, amplitude
            "stand": np.array([1.0, 1.0, 0.0]),
            "walk": np.array([0.75, 1.0, 0.08]),
            "trot": np.array([0.45, 2, 0.08]),
            "canter": np.array([0.4, 4, 0.06]),
            "gallop": np.array([0.3, 3.5, 0.10]),
        }
        self._torso_idx = self.sys.body_id("base")
        self._init_q = self.sys.keyframe("home")
        self._default_pose = self.sys.keyframe("home")[7:]
        self.joint_range = np.array(
            [[-0.5, 0.5], [0.4, 1.4], [-2.3, -0.85]] * 2 + [[-0.5, 0.5], [0.4, 1.4], [-2.3, -1.3]] * 2)
        feet_site = ["FL_foot", "FR_foot", "RL_foot", "RR_foot"]
        self._feet_site_id = np.array([self.sys.site_id(f) for f in feet_site])
        self._pos_tar = np.array([0.282, 0.0, 0.3])

    def make_system(self, config: UnitreeGo2EnvConfig) -> System:
        model_path = get_model_path("unitree_go2", "mjx_scene_force.xml")
        sys = System(CompiledModel.load(model_path))
        return sys.tree_replace({"opt.timestep": config.timestep})

    def _init_info(self, rng) -> Dict[str, Any]:
        return {"rng": rng, "pos_tar": self._pos_tar.copy(), "vel_tar": np.zeros(3),
                "ang_vel_tar": np.zeros(3), "yaw_tar": 0.0, "step": 0,
                "randomize_target": self._config.randomize_tasks}

    def _fill_reward_desc(self, d) -> None:
        c = self._config
        d.torso_body = int(self._torso_idx)
        d.nfeet = 4
        _capi._set(d.feet_site, self._feet_site_id.astype(np.int32))
        duty, cadence, amp = self._gait_params[self._gait]
        d.gait_duty, d.gait_cadence, d.gait_amplitude = float(duty), float(cadence), float(amp)
        _capi._set(d.gait_phase, self._gait_phase[self._gait])
        _capi._set(d.vel_cmd, [c.default_vx, c.default_vy, 0.0])
        _capi._set(d.ang_cmd, [0.0, 0.0, c.default_vyaw])
        d.ramp_up_time = float(c.ramp_up_time)
        _capi._set(d.pos_tar, self._pos_tar)
        d.n_stage = 1
        d.jump_dt = 1.0


@dataclass
class UnitreeGo2SeqJumpEnvConfig(UnitreeGo2EnvConfig):
    jump_dt: float = 1.0
    contact_targets: Any = None
    contact_target_radius: Any = None
    pose_target_sequence: Any = None
    yaw_target_sequence: Any = None


def _euler_to_quat_deg(v):
    c1, c2, c3 = np.cos(np.asarray(v) * np.pi / 360)
    s1, s2, s3 = np.sin(np.asarray(v) * np.pi / 360)
    return np.array([c1 * c2 * c3 - s1 * s2 * s3, s1 * c2 * c3 + c1 * s2 * s3,
                     c1 * s2 * c3 - s1 * c2 * s3, c1 * c2 * s3 + s1 * s2 * c3])


def _quat_to_3x3(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


class UnitreeGo2SeqJumpEnv(UnitreeGo2Env):
    supports_randomize_tasks = True      # `reset` draws a whole jump sequence (unitree_go2_env.py:383-394, 594-631)
    env_id = _capi.ENV_IDS["unitree_go2_seq_jump"]

    def __init__(self, config: UnitreeGo2SeqJumpEnvConfig = None):
        config = config if config is not None else UnitreeGo2SeqJumpEnvConfig()
        super().__init__(config)
        if config.contact_targets is None or config.contact_target_radius is None:
            (self._contact_targets, self._contact_target_radius, self._pose_target_sequence,
             self._yaw_target_sequence) = UnitreeGo2SeqJumpEnv.generate_jumping_sequence(
                config.pose_target_sequence, config.yaw_target_sequence, 0.1)
        else:
            self._contact_targets = np.asarray(config.contact_targets, dtype=np.float64)
            self._contact_target_radius = np.asarray(config.contact_target_radius, dtype=np.float64)
            self._pose_target_sequence = np.asarray(config.pose_target_sequence, dtype=np.float64)
            self._yaw_target_sequence = np.asarray(config.yaw_target_sequence, dtype=np.float64)
        self.joint_range = np.array(
            [[-0.5, 0.5], [0.4, 2.0], [-2.3, -1.3]] * 2 + [[-0.5, 0.5], [0.4, 1.4], [-2.3, -1.3]] * 2)
        self._pos_tar = np.array([0.0, 0.0, 0.27])

    @staticmethod
    def generate_jumping_sequence(com_pos, com_heading, foot_place_radius: float):
        """unitree_go2_env.py:559-592 (heading in rad; euler_to_quat takes degrees)."""
        com_pos = np.asarray(com_pos, dtype=np.float64)
        com_heading = np.asarray(com_heading, dtype=np.float64)
        n_steps = com_pos.shape[0]
        assert n_steps == len(com_heading)
        offsets0 = np.array([[0.2, -0.135, 0.0],   # FR
                             [0.2, 0.135, 0.0],    # FL
                             [-0.2, -0.135, 0.0],  # RR
                             [-0.2, 0.135, 0.0]])  # RL
        targets = []
        for i in range(n_steps):
            R = _quat_to_3x3(_euler_to_quat_deg([0.0, 0.0, com_heading[i] * 180 / np.pi]))
            targets.append(np.repeat(com_pos[i][None], 4, axis=0) + offsets0 @ R.T)
        return (np.array(targets), np.full((n_steps, 4), foot_place_radius), com_pos.copy(),
                com_heading.copy())

    N_RANDOM_STAGES = 10    # `n_steps` of the reference's sample_command (+ the start pose = 11 stages)

    def sample_command(self, rng):
        """unitree_go2_env.py:594-631: a random walk of the COM target (xy steps uniform in
        +-0.65 m, yaw steps uniform in +-0.5 rad; fp32 running sums like the JAX scan) from
        ``jax.random.split(rng, 20)``, turned into stage tables by generate_jumping_sequence."""
        from dial_mpc_b200 import random as drandom
        n = self.N_RANDOM_STAGES
        keys = drandom.split_n(rng, 2 * n)
        pos = np.zeros((n + 1, 3), dtype=np.float32)
        yaw = np.zeros(n + 1, dtype=np.float32)
        pos[0] = [0.0, 0.0, 0.27]
        for i in range(n):
            pos[i + 1] = pos[i]
            pos[i + 1, :2] += drandom.uniform(keys[i], 2, -0.65, 0.65)
            yaw[i + 1] = yaw[i] + drandom.uniform(keys[n + i], 1, -0.5, 0.5)[0]
        return UnitreeGo2SeqJumpEnv.generate_jumping_sequence(pos, yaw, 0.1)

    def command_override(self, info, horizon):
        return None     # this env's step draws no per-step command (unitree_go2_env.py:403-521)

    def stage_tables(self, info):
        """The jump sequence a state carries (reset may have drawn its own): what the plan's stage
        constants must hold when a kernel is launched from that state."""
        return (info.get("pose_target_sequence", self._pose_target_sequence),
                info.get("yaw_target_sequence", self._yaw_target_sequence),
                info.get("contact_targets", self._contact_targets),
                info.get("contact_target_radius", self._contact_target_radius))

    def _init_info(self, rng) -> Dict[str, Any]:
        info = super()._init_info(rng)
        tables = (self._contact_targets, self._contact_target_radius, self._pose_target_sequence,
                  self._yaw_target_sequence)
        if self._config.randomize_tasks:
            tables = self.sample_command(rng)
        info.update(last_ctrl=np.zeros(12, dtype=np.float32), contact_stage=0,
                    contact_targets=tables[0], contact_target_radius=tables[1],
                    pose_target_sequence=tables[2], yaw_target_sequence=tables[3])
        return info

    _done_height = 0.1

    def _get_obs(self, pipeline_state, info):
        """unitree_go2_env.py:523-557: [vel_tar, ang_vel_tar, last_ctrl, torso position error to
        the stage's pose target, roll, pitch, wrapped yaw error, joint angles, vb, ab, joint
        velocities]."""
        import torch
        ps = pipeline_state
        w, x, y, z = ps.qpos[3], ps.qpos[4], ps.qpos[5], ps.qpos[6]
        # brax.math.quat_to_euler
        ez = torch.atan2(-2 * x * y + 2 * w * z, x * x + w * w - z * z - y * y)
        ey = torch.asin(torch.clamp(2 * x * z + 2 * w * y, -1.0, 1.0))
        ex = torch.atan2(-2 * y * z + 2 * w * x, z * z - y * y - x * x + w * w)
        stage = int(info.get("contact_stage", 0))
        pos = ps.kin[0:3] if ps.kin is not None else ps.qpos[0:3]
        dpos = pos - self._dev(ps, info.get("pose_target_sequence", self._pose_target_sequence)[stage])
        dyaw = ez - float(info.get("yaw_target_sequence", self._yaw_target_sequence)[stage])
        dyaw = torch.atan2(torch.sin(dyaw), torch.cos(dyaw)).reshape(1)
        last = info.get("last_ctrl", np.zeros(12))
        last = last if torch.is_tensor(last) else self._dev(ps, last)
        return torch.cat([self._dev(ps, info.get("vel_tar", np.zeros(3))), self._dev(ps, info.get("ang_vel_tar", np.zeros(3))), last, dpos,
                          torch.stack([ex, ey]), dyaw, ps.qpos[7:], self._vb_ab(ps), ps.qvel[6:]])

    def step(self, state, action):
        new = super().step(state, action)
        new.info["last_ctrl"] = new.pipeline_state.ctrl          # unitree_go2_env.py:516
        return new

    def _next_info(self, info):
        new = super()._next_info(info)
        for k in ("vel_tar", "ang_vel_tar"):       # no ramp in this env
            if k in info:
                new[k] = info[k]
        n = len(info.get("contact_targets", self._contact_targets))
        new["contact_stage"] = int(min(np.floor(np.float32(new["step"]) * np.float32(self.dt)
                                                / np.float32(self._config.jump_dt)), n - 1))
        return new

    def _fill_reward_desc(self, d) -> None:
        super()._fill_reward_desc(d)
        n = len(self._contact_targets)
        if n > _capi.DEFINES["DIAL_MAXSTAGE"]:
            raise ValueError("too many jump stages")
        d.n_stage = n
        d.jump_dt = float(self._config.jump_dt)
        _capi._set(d.pose_seq, self._pose_target_sequence)
        _capi._set(d.yaw_seq, self._yaw_target_sequence)
        _capi._set(d.contact_targets, self._contact_targets)
        _capi._set(d.contact_radius, self._contact_target_radius)
