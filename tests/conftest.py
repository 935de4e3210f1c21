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


from baseline_configs import ENV_CFG as ENV_CASES  # noqa: E402  (one table for bench.py and the tests)


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
