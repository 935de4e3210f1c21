"""Unitree H1 environments on the CUDA sampling core.

Same classes / config fields / registry names as the reference ``UnitreeH1WalkEnv``
(dial_mpc/envs/unitree_h1_env.py:25-375) and ``UnitreeH1LocoEnv`` (:570-902).  The push-crate
variant (:378-567: box geom, slide joint with frictionloss) is not built."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Union

import numpy as np

from dial_mpc_b200 import _capi
from dial_mpc_b200.envs.base_env import BaseEnv, BaseEnvConfig, System
from dial_mpc_b200.modelc import CompiledModel
from dial_mpc_b200.utils.io_utils import get_model_path


def _h1_kp():
    return np.array([200.0, 200.0, 200.0, 200.0, 60.0] * 2 + [200.0] + [60.0] * 8)


def _h1_kd():
    return np.array([5.0, 5.0, 5.0, 5.0, 1.5] * 2 + [5.0] + [1.5] * 8)


@dataclass
class UnitreeH1WalkEnvConfig(BaseEnvConfig):
    kp: Union[float, Any] = field(default_factory=_h1_kp)
    kd: Union[float, Any] = field(default_factory=_h1_kd)
    default_vx: float = 1.0
    default_vy: float = 0.0
    default_vyaw: float = 0.0
    ramp_up_time: float = 2.0
    gait: str = "jog"


class UnitreeH1WalkEnv(BaseEnv):
    supports_randomize_tasks = True
    env_id = _capi.ENV_IDS["unitree_h1_walk"]

    def __init__(self, config: UnitreeH1WalkEnvConfig):
        super().__init__(config)
        self._pelvis_idx = self.sys.body_id("pelvis")
        self._torso_idx = self.sys.body_id("torso_link")
        self._left_foot_idx = self.sys.site_id("left_foot")
        self._right_foot_idx = self.sys.site_id("right_foot")
        self._feet_site_id = np.array([self._left_foot_idx, self._right_foot_idx], dtype=np.int32)
        self._gait = config.gait
        self._gait_phase = {"stand": np.zeros(2), "slow_walk": np.array([0.0, 0.5]),
                            "walk": np.array([0.0, 0.5]), "jog": np.array([0.0, 0.5])}
        self._gait_params = {  # ratio, cadence, amplitude
            "stand": np.array([1.0, 1.0, 0.0]), "slow_walk": np.array([0.6, 0.8, 0.15]),
            "walk": np.array([0.5, 1.0, 0.15]), "jog": np.array([0.3, 2, 0.2])}
        self._init_q = self.sys.keyframe("home")
        self._default_pose = self.sys.keyframe("home")[7:]
        self.joint_range = np.array(
            [[-0.3, 0.3], [-0.3, 0.3], [-1.0, 1.0], [0.0, 1.74], [-0.6, 0.4]] * 2 + [[-0.5, 0.5]]
            + [[-0.78, 0.78], [-0.3, 0.3], [-0.3, 0.3], [-0.3, 0.3]] * 2)
        self._pos_tar = np.array([0.0, 0.0, 1.3])

    def make_system(self, config: UnitreeH1WalkEnvConfig) -> System:
        model_path = get_model_path("unitree_h1", "mjx_scene_h1_walk.xml")
        sys = System(CompiledModel.load(model_path))
        return sys.tree_replace({"opt.timestep": config.timestep})

    def _init_info(self, rng) -> Dict[str, Any]:
        return {"rng": rng, "pos_tar": self._pos_tar.copy(), "vel_tar": np.zeros(3),
                "ang_vel_tar": np.zeros(3), "yaw_tar": 0.0, "step": 0,
                "randomize_target": self._config.randomize_tasks}

    def _fill_reward_desc(self, d) -> None:
        c = self._config
        d.torso_body = int(self._torso_idx)
        d.nfeet = 2
        _capi._set(d.feet_site, self._feet_site_id)
        duty, cadence, amp = self._gait_params[self._gait]
        d.gait_duty, d.gait_cadence, d.gait_amplitude = float(duty), float(cadence), float(amp)
        _capi._set(d.gait_phase, self._gait_phase[self._gait])
        _capi._set(d.vel_cmd, [c.default_vx, c.default_vy, 0.0])
        _capi._set(d.ang_cmd, [0.0, 0.0, c.default_vyaw])
        d.ramp_up_time = float(c.ramp_up_time)
        _capi._set(d.pos_tar, self._pos_tar)
        d.n_stage = 1
        d.jump_dt = 1.0


def _h1_loco_kp():
    return np.array([200.0, 200.0, 200.0, 200.0, 60.0] * 2 + [200.0])


def _h1_loco_kd():
    return np.array([5.0, 5.0, 5.0, 5.0, 1.5] * 2 + [5.0])


@dataclass
class UnitreeH1LocoEnvConfig(BaseEnvConfig):
    kp: Union[float, Any] = field(default_factory=_h1_loco_kp)
    kd: Union[float, Any] = field(default_factory=_h1_loco_kd)
    default_vx: float = 1.0
    default_vy: float = 0.0
    default_vyaw: float = 0.0
    ramp_up_time: float = 2.0
    gait: str = "jog"


class UnitreeH1LocoEnv(UnitreeH1WalkEnv):
    """``UnitreeH1LocoEnv`` (dial_mpc/envs/unitree_h1_env.py:609-902): legs + torso
    actuated, arms welded, two capsules per foot, one Newton / line-search iteration."""
    env_id = _capi.ENV_IDS["unitree_h1_loco"]

    def __init__(self, config: UnitreeH1LocoEnvConfig):
        super().__init__(config)
        self._gait_params = {  # ratio, cadence, amplitude
            "stand": np.array([1.0, 1.0, 0.0]), "slow_walk": np.array([0.6, 0.8, 0.15]),
            "walk": np.array([0.5, 1.5, 0.10]), "jog": np.array([0.3, 2.0, 0.2])}
        self.joint_range = np.array(
            [[-0.2, 0.2], [-0.2, 0.2], [-0.6, 0.6], [0.0, 1.5], [-0.6, 0.4]] * 2 + [[-0.5, 0.5]])

    def make_system(self, config) -> System:
        model_path = get_model_path("unitree_h1", "mjx_scene_h1_loco.xml")
        sys = System(CompiledModel.load(model_path))
        return sys.tree_replace({"opt.timestep": config.timestep})
