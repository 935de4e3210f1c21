"""Base env config — same fields and defaults as the reference ``BaseEnvConfig``
(dial_mpc/config/base_env_config.py:4-20)."""
from dataclasses import dataclass


@dataclass
class BaseEnvConfig:
    task_name: str = "default"
    randomize_tasks: bool = False  # walk envs: a one-step random command every 500 steps (dial_plan_set_command)
    kp: float = 30.0  # P gain, or a list of P gains for each joint
    kd: float = 1.0  # D gain, or a list of D gains for each joint
    debug: bool = False
    dt: float = 0.02  # dt of the environment step
    timestep: float = 0.02  # timestep of the underlying simulator step
    backend: str = "mjx"  # kept for API parity; the physics is always the CUDA restatement
    leg_control: str = "torque"  # "torque" or "position"
    action_scale: float = 1.0
