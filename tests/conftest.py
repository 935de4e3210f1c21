import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        skip = pytest.mark.skip(reason="no CUDA device")
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)


ENV_CASES = {
    "unitree_go2_walk": dict(default_vx=0.8, ramp_up_time=1.0),
    "unitree_go2_seq_jump": dict(
        pose_target_sequence=[[0, 0, 0.27], [0.4, 0, 0.27], [0.8, 0, 0.27], [1.2, 0, 0.27], [1.6, 0, 0.27]],
        yaw_target_sequence=[0.0] * 5),
    "unitree_h1_walk": dict(default_vx=2.0, ramp_up_time=3.0),
    "allegro_reorient": dict(dt=0.02, timestep=0.005, leg_control="position"),
    "unitree_h1_loco": dict(default_vx=0.6, ramp_up_time=3.0, gait="walk"),
}


def make_pair(name):
    """(product env, oracle env) with identical configuration."""
    import numpy as np
    import dial_mpc_b200.envs as E
    from oracle.envs_oracle import make_env
    cfg = ENV_CASES[name]
    cfg_t = E.get_config(name)
    ecfg = cfg_t(**{k: (np.array(v) if isinstance(v, list) else v) for k, v in cfg.items()})
    return E.get_environment(name, config=ecfg), make_env(name, cfg)


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as g
    g.build()
    return True
