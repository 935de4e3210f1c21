"""Base class for user-written environments (README.md:223-312 of the reference:
"Writing Custom Environment").

The reference user subclasses ``BaseEnv`` and writes ``make_system`` / ``reset`` / ``step`` in
JAX.  On the CUDA core the subclass provides

* ``make_system(config)`` — e.g. ``System(compile_mjcf("my_model.xml"))`` (free root joint
  first, actuators on the following hinge joints; capacities in include/dial_b200.h),
* ``reward_source`` — path of a ``.cuh`` file defining ``dial_custom_reward``
  (include/dial_custom_reward.h),
* optionally ``user_params()`` (≤ DIAL_MAXUSER floats handed to the reward) and ``init_q``.

``step``'s PD control / physics substeps (base_env.py:37-66 + ``pipeline_step``) are the same
fused kernel as for the built-in envs; the reward is compiled into a dedicated build of the
library on first use (``dial_mpc_b200.custom.build_library``) and cached in-tree.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from dial_mpc_b200 import _capi
from dial_mpc_b200.envs.base_env import BaseEnv, BaseEnvConfig


class CustomRewardEnv(BaseEnv):
    env_id = _capi.ENV_IDS["custom"]
    reward_source: str = ""          # path of the .cuh reward file
    build_defines: tuple = ()        # extra compile-time options of the kernels, e.g. ("DIAL_ROBUST_LS",)
    init_keyframe: Optional[str] = "home"

    def __init__(self, config: BaseEnvConfig):
        super().__init__(config)
        if not self.reward_source:
            raise ValueError(f"{type(self).__name__}.reward_source must name the CUDA reward file")
        m = self.sys.model
        if self.init_keyframe is not None and self.init_keyframe in (m.keyframes or {}):
            self._init_q = self.sys.keyframe(self.init_keyframe)
        else:
            self._init_q = np.asarray(m.qpos0, dtype=np.float64).copy()
        self._library_path: Optional[str] = None

    @property
    def library_path(self) -> str:
        """Build of libdial_b200 with this env's reward fused in (compiled on first use)."""
        if self._library_path is None:
            from dial_mpc_b200 import custom
            self._library_path = custom.build_library(self.reward_source, model=self.sys.model,
                                                       defines=self.build_defines)
        return self._library_path

    def _get_obs(self, pipeline_state, info):
        """Default observation of a custom env: [qpos, qvel] (the reference user writes their own;
        override for anything else — it is not on the sampling path)."""
        import torch
        return torch.cat([pipeline_state.qpos, pipeline_state.qvel])

    def _get_done(self, pipeline_state, info):
        import torch
        return torch.zeros((), device=pipeline_state.qpos.device)

    def user_params(self) -> np.ndarray:
        """Constants for the reward (``ctx->user``)."""
        return np.zeros(0, dtype=np.float32)

    def _fill_reward_desc(self, d) -> None:
        u = np.asarray(self.user_params(), dtype=np.float32).ravel()
        if u.size > _capi.DEFINES["DIAL_MAXUSER"]:
            raise ValueError(f"user_params: at most {_capi.DEFINES['DIAL_MAXUSER']} floats")
        d.n_user = int(u.size)
        _capi._set(d.user, u)
