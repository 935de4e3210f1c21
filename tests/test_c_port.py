"""CPU tests of the C port of the per-sample step (oracle/c/dial_port.c, the fp32 CPU baseline
of bench.py): built in double it is a second, independent-in-code restatement that must agree
with the NumPy oracle up to the fp32 rounding of the model constants it receives through the
public C structs; built in float it must stay within fp32 tolerances (the same ones the CUDA
path is held to)."""
import numpy as np
import pytest

from baseline_configs import ENV_CFG
from oracle import build_oracle
from oracle.c_port import CPort
from oracle.envs_oracle import make_env


@pytest.fixture(scope="module")
def built_port():
    build_oracle.build()
    return True


@pytest.mark.parametrize("name,H", [("unitree_go2_walk", 17), ("unitree_go2_seq_jump", 26), ("unitree_h1_walk", 31),
                                    ("unitree_h1_loco", 21)])
def test_c_port_matches_numpy_oracle(built_port, name, H):
    o = make_env(name, ENV_CFG[name])
    s = o.reset()
    for _ in range(5):
        s, _, _ = o.step(s, np.zeros((1, o.nu)))
    s.step[:] = 40          # the seq-jump stage boundary (step 50) falls inside the rollout
    rng = np.random.default_rng(0)
    us = np.clip(rng.normal(size=(24, H, o.nu)) * 0.6, -1, 1)
    r, q, qd, x = o.rollout(s, us)
    rc, qc, qdc, xc, _ = CPort(o, real="double").rollout(s, us)
    assert np.abs(qc - q).max() < 2e-5 and np.abs(xc - x).max() < 2e-5 and np.abs(qdc - qd).max() < 1e-3
    assert np.abs(rc - r).max() < 1e-4 * (1 + np.abs(r).max())
    rf, qf, qdf, xf, _ = CPort(o, real="float").rollout(s, us)
    assert np.abs(qf - q).max() < 5e-4 and np.abs(xf - x).max() < 5e-4
    assert np.abs(rf - r).max() < 2e-3 * (1 + np.abs(r).max())


def test_c_port_rejects_models_outside_its_scope(built_port):
    o = make_env("allegro_reorient", ENV_CFG["allegro_reorient"])
    with pytest.raises(NotImplementedError):
        CPort(o)


def test_bench_reference_arm_prints_the_contract_line(built_port):
    """`bench.py --impl reference` (the driver's reference arm: the CPU restatement on the host cores, no
    GPU involved): one JSON line with the contract's keys, on the smallest BASELINE config."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "0",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, cwd=root, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "sample-steps/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1 and line["vs_baseline"] is None
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port-c", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "sample-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "configs[0]" in line["config"]["workload"]
