"""Example custom environment on the DENSE solver path (elliptic friction cones): a ball on the floor between two
two-link fingers that close on it like tweezers (``pincher.xml``, nv = 10).  The stock library instantiates the dense solver for the
Allegro scene (nv = 22); ``dial_mpc_b200.custom`` compiles this env's library with the solver
instantiated for its own dof count (-DDIAL_DENSE_NV=10) and the reward of ``pincher_reward.cuh``.

    python -m dial_mpc_b200.core.dial_core --config pincher.yaml --custom-env pincher_env
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
class PincherEnvConfig(BaseEnvConfig):
    dt: float = 0.02
    timestep: float = 0.005           # 4 physics substeps per env step, like allegro_reorient.yaml
    leg_control: str = "position"
    target_height: float = 0.02
    target_spin: float = 2.0
    joint_weight: float = 0.05


class PincherEnv(CustomRewardEnv):
    reward_source = os.path.join(_HERE, "pincher_reward.cuh")

    def __init__(self, config: PincherEnvConfig):
        super().__init__(config)
        # sampling range of the joint targets (proximal, distal) x 2 fingers
        self.joint_range = np.array([[-0.3, 0.6], [-0.3, 0.8]] * 2)

    def make_system(self, config: PincherEnvConfig) -> System:
        sys = System(compile_mjcf(os.path.join(_HERE, "pincher.xml")))
        return sys.tree_replace({"opt.timestep": config.timestep})

    def user_params(self):
        c = self._config
        return np.array([c.target_height, c.target_spin, c.joint_weight, 0.1, 0.1, 0.1, 0.1], dtype=np.float32)


dial_envs.register_environment("pincher_spin", PincherEnv)
dial_envs.register_config("pincher_spin", PincherEnvConfig)
