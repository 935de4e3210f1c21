"""Environment + config registries.

Keeps the two registries of the reference: the env-config registry of
``dial_mpc.envs`` (dial_mpc/envs/__init__.py:14-30: ``register_config`` / ``get_config``)
and the env-class registry the reference borrows from Brax
(``brax_envs.register_environment`` / ``get_environment``; call sites
dial_mpc/envs/unitree_go2_env.py:806-808, dial_mpc/core/dial_core.py:221)."""
from typing import Any, Dict, Type

from dial_mpc_b200.envs.base_env import BaseEnv, BaseEnvConfig, PipelineState, State, System  # noqa: F401
from dial_mpc_b200.envs.unitree_go2_env import (UnitreeGo2Env, UnitreeGo2EnvConfig, UnitreeGo2SeqJumpEnv,
                                                UnitreeGo2SeqJumpEnvConfig)
from dial_mpc_b200.envs.unitree_h1_env import (UnitreeH1LocoEnv, UnitreeH1LocoEnvConfig, UnitreeH1WalkEnv,
                                               UnitreeH1WalkEnvConfig)
from dial_mpc_b200.envs.manipulation import AllegroReorientEnv, AllegroReorientEnvConfig

_configs: Dict[str, Any] = {
    "unitree_h1_walk": UnitreeH1WalkEnvConfig,
    "unitree_go2_walk": UnitreeGo2EnvConfig,
    "unitree_go2_seq_jump": UnitreeGo2SeqJumpEnvConfig,
    "allegro_reorient": AllegroReorientEnvConfig,
    "unitree_h1_loco": UnitreeH1LocoEnvConfig,
}
_envs: Dict[str, Type[BaseEnv]] = {}


def register_config(name: str, config: Any):
    _configs[name] = config


def get_config(name: str) -> Any:
    return _configs[name]


def register_environment(env_name: str, env_class: Type[BaseEnv]):
    """``brax.envs.register_environment`` stand-in."""
    _envs[env_name] = env_class


def get_environment(env_name: str, **kwargs) -> BaseEnv:
    """``brax.envs.get_environment(env_name, config=...)`` stand-in."""
    return _envs[env_name](**kwargs)


register_environment("unitree_go2_walk", UnitreeGo2Env)
register_environment("unitree_go2_seq_jump", UnitreeGo2SeqJumpEnv)
register_environment("unitree_h1_walk", UnitreeH1WalkEnv)
register_environment("allegro_reorient", AllegroReorientEnv)
register_environment("unitree_h1_loco", UnitreeH1LocoEnv)
