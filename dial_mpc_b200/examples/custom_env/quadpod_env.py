"""Example custom environment (the reference's README.md:223-312 recipe on the CUDA core).

    python -m dial_mpc_b200.core.dial_core --config quadpod.yaml --custom-env quadpod_env

run from this directory.  The model is compiled from MJCF, the reward is the CUDA function in
``quadpod_reward.cuh``; the first run compiles a dedicated library build (about a minute).
"""
import os
from dataclasses import dataclass

import numpy as np

import dial_mpc_b200.envs as dial_envs
from dial_mpc_b200.config.base_env_config import BaseEnvConfig
from dial_mpc_b200.envs.base_env import System
from dial_mpc_b200.envs.custom_env import CustomRewardEnv
from dial_mpc_b200.modelc import compile_mjcf

_HERE = os.path.dirname(os.path.abspath(__file__))


@dataclass
class QuadpodEnvConfig(BaseEnvConfig):
    kp: float = 20.0
    kd: float = 0.5
    target_vx: float = 0.5
    ramp_up_time: float = 1.0
    target_height: float = 0.33
    energy_weight: float = 0.01
    feet_weight: float = 2.0


class QuadpodEnv(CustomRewardEnv):
    reward_source = os.path.join(_HERE, "quadpod_reward.cuh")

    def __init__(self, config: QuadpodEnvConfig):
        super().__init__(config)
        # sampling range of the joint targets (hip, knee) x 4
        self.joint_range = np.array([[-0.2, 1.2], [-2.0, -0.8]] * 4)

    def make_system(self, config: QuadpodEnvConfig) -> System:
        sys = System(compile_mjcf(os.path.join(_HERE, "quadpod.xml")))
        return sys.tree_replace({"opt.timestep": config.timestep})

    def user_params(self):
        c = self._config
        return np.array([c.target_vx, c.ramp_up_time, c.target_height, c.energy_weight, c.feet_weight, 12.0],
                        dtype=np.float32)


dial_envs.register_environment("quadpod_walk", QuadpodEnv)
dial_envs.register_config("quadpod_walk", QuadpodEnvConfig)
