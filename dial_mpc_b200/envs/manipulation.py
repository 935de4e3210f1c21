"""Allegro in-hand ball reorientation on the CUDA sampling core.

Same class / config / registry name as the reference ``AllegroReorientEnv``
(dial_mpc/envs/manipulation.py:23-116): position-controlled hand (``leg_control: position``,
4 physics substeps of 5 ms per env step), reward = ball angular-velocity tracking + ball
position + joint deviation.  The model uses elliptic friction cones with condim 6 on the ball
and contacts between moving bodies, so the rollout kernel runs its dense solver path."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, Union

import numpy as np

from dial_mpc_b200 import _capi
from dial_mpc_b200.envs.base_env import BaseEnv, BaseEnvConfig, System
from dial_mpc_b200.modelc import CompiledModel
from dial_mpc_b200.utils.io_utils import get_model_path


@dataclass
class AllegroReorientEnvConfig(BaseEnvConfig):
    kp: Union[float, Any] = 1.0
    kd: Union[float, Any] = 0.1


class AllegroReorientEnv(BaseEnv):
    env_id = _capi.ENV_IDS["allegro_reorient"]

    def __init__(self, config: AllegroReorientEnvConfig):
        super().__init__(config)
        self._object_body_idx = self.sys.body_id("object")
        self._init_q = self.sys.keyframe("in_hand_reorient")
        self._ang_vel_tar = np.array([0.0, 0.0, 0.5])
        self._pos_tar = np.array([0.0, 0.0, 0.13])

    def make_system(self, config: AllegroReorientEnvConfig) -> System:
        model_path = get_model_path("wonik_allegro", "scene_left.xml")
        sys = System(CompiledModel.load(model_path))
        return sys.tree_replace({"opt.timestep": config.timestep})

    def _init_info(self, rng) -> Dict[str, Any]:
        return {"rng": rng, "ang_vel_tar": self._ang_vel_tar.copy(), "pos_tar": self._pos_tar.copy(), "step": 0}

    def _get_obs(self, pipeline_state, info):
        import torch
        return torch.zeros(1, device=pipeline_state.qpos.device)     # manipulation.py:57,97

    def _get_done(self, pipeline_state, info):
        import torch
        return torch.full((1,), float(info["step"] >= 100), device=pipeline_state.qpos.device)   # manipulation.py:86-87

    def act2joint(self, act):
        """manipulation.py:102-115: the sampling range is shifted by the initial joint pose."""
        act = np.asarray(act, dtype=np.float64)
        an = (act * self._config.action_scale + 1.0) / 2.0
        jt = self.joint_range[:, 0] + self._init_q[7:] + an * (self.joint_range[:, 1] - self.joint_range[:, 0])
        return np.clip(jt, self.physical_joint_range[:, 0], self.physical_joint_range[:, 1])

    def plan_desc(self, **kw):
        if self._config.leg_control != "position":
            raise NotImplementedError("AllegroReorientEnv supports leg_control='position' only")
        return super().plan_desc(**kw)

    def _fill_reward_desc(self, d) -> None:
        d.torso_body = int(self._object_body_idx)
        d.nfeet = 0
        _capi._set(d.ang_cmd, self._ang_vel_tar)
        _capi._set(d.pos_tar, self._pos_tar)
        _capi._set(d.joint_offset, self._init_q[7:])
        d.n_stage = 1
        d.jump_dt = 1.0
        d.ramp_up_time = 1.0
