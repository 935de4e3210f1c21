"""Oracle pinning (CPU): known-answer tests with in-container references (scipy, Random123
vectors) and the committed golden fixtures.  Parity with the real MJX is unpinned (no
reference tests / installable reference) — see oracle/mjx_oracle.py."""
import os

import numpy as np
import pytest
from scipy.interpolate import InterpolatedUnivariateSpline
from scipy.spatial.transform import Rotation

from oracle import envs_oracle as eo
from oracle import mjx_oracle as mo
from oracle import planner_oracle as po
from tests.conftest import ENV_CASES

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_threefry_kats():
    # Random123 known-answer vectors
    assert [int(v) for v in np.array(po.threefry2x32((0, 0), np.uint32([0]), np.uint32([0]))).ravel()] == [0x6B200159, 0x99BA4EFE]
    f = 0xFFFFFFFF
    assert [int(v) for v in np.array(po.threefry2x32((f, f), np.uint32([f]), np.uint32([f]))).ravel()] == [0x1CB996FC, 0xBB002BE7]
    got = po.threefry2x32((0x13198A2E, 0x03707344), np.uint32([0x243F6A88]), np.uint32([0x85A308D3]))
    assert [int(v) for v in np.array(got).ravel()] == [0xC4923A9C, 0x483DF7A0]


def test_jax_normal_statistics():
    x = po.jax_normal_legacy((0, 42), (4096, 5, 12))
    assert abs(x.mean()) < 0.01 and abs(x.std() - 1.0) < 0.01
    assert np.isfinite(x).all()
    # split produces distinct keys
    a, b = po.jax_split_legacy((0, 0))
    assert tuple(a) != tuple(b)


@pytest.mark.parametrize("Hs,Hn", [(16, 4), (25, 5), (20, 5), (30, 5), (25, 4), (20, 4), (24, 6)])
def test_spline_matches_scipy(Hs, Hn):
    from dial_mpc_b200.utils.spline import interp_matrix
    us = np.linspace(0, 0.02 * Hs, Hs + 1)
    nd = np.linspace(0, 0.02 * Hs, Hn + 1)
    y = np.random.default_rng(0).standard_normal(Hn + 1)
    ref = InterpolatedUnivariateSpline(nd, y, k=2)(us)
    assert np.abs(interp_matrix(nd, us) @ y - ref).max() < 1e-12
    assert np.abs(po.spline_matrix(nd, us) @ y - ref).max() < 1e-12
    # interpolating: the node values are reproduced, rows sum to one
    M = interp_matrix(nd, us)
    assert np.abs(M.sum(1) - 1).max() < 1e-12
    u = np.random.default_rng(1).standard_normal(Hs + 1)
    ref2 = InterpolatedUnivariateSpline(us, u, k=2)(nd)
    assert np.abs(interp_matrix(us, nd) @ u - ref2).max() < 1e-12


def test_quaternion_helpers_vs_scipy():
    rng = np.random.default_rng(0)
    q = rng.standard_normal((50, 4))
    q /= np.linalg.norm(q, axis=-1, keepdims=True)
    v = rng.standard_normal((50, 3))
    R = Rotation.from_quat(q[:, [1, 2, 3, 0]])
    assert np.abs(eo.rotate(v, q) - R.apply(v)).max() < 1e-12
    assert np.abs(eo.inv_rotate(v, q) - R.inv().apply(v)).max() < 1e-12
    assert np.abs(mo.qmat(q) - R.as_matrix()).max() < 1e-12
    e = eo.quat_to_euler(q)
    ref = R.as_euler("XYZ")
    d = np.abs(np.arctan2(np.sin(e - ref), np.cos(e - ref)))
    assert d.max() < 1e-9
    for deg in ([10.0, -20.0, 30.0], [0.0, 0.0, 90.0]):
        qq = eo.euler_to_quat_deg(np.array(deg))
        ref_q = Rotation.from_euler("XYZ", deg, degrees=True).as_quat()[[3, 0, 1, 2]]
        assert min(np.abs(qq - ref_q).max(), np.abs(qq + ref_q).max()) < 1e-12


def test_foot_step_profile():
    t = np.linspace(0, 2, 401)
    h = eo.get_foot_step(0.45, 2, 0.08, [0.0, 0.5, 0.5, 0.0], t)
    assert h.shape == (401, 4) and h.min() >= 0 and abs(h.max() - 0.08) < 1e-3
    # duty ratio: fraction of time on the ground (height == 0) is ~ duty
    assert abs((h[:, 0] == 0).mean() - 0.45) < 0.03
    # phase offset of half a period
    assert np.abs(h[:-50, 0] - h[50:, 1]).max() < 1e-9
    from dial_mpc_b200.utils.function_utils import get_foot_step
    assert np.abs(get_foot_step(0.45, 2, 0.08, [0.0, 0.5, 0.5, 0.0], 0.37) - eo.get_foot_step(0.45, 2, 0.08, [0.0, 0.5, 0.5, 0.0], np.array([0.37]))[0]).max() < 1e-12


@pytest.mark.parametrize("name", list(ENV_CASES))
def test_oracle_reproduces_golden(name):
    g = np.load(os.path.join(GOLD, f"{name}.npz"))
    env = eo.make_env(name, ENV_CASES[name])
    s = eo.OState(g["qpos"][None], g["qvel"][None], g["qacc_warmstart"][None],
                  np.array([int(g["step"])]), np.array([int(g["stage"])]))
    pl = po.PlannerOracle(env, int(g["N"]), int(g["Hs"]), int(g["Hn"]), float(g["temp"]),
                          0.9 if "go2" in name else 1.0, 0.5)
    Ybar, info = pl.reverse_once(s, g["eps"].astype(np.float64), g["Ybar0"], g["noise_scale"])
    assert np.abs(info["rews"] - g["rews"]).max() < 1e-9
    assert np.abs(Ybar - g["Ybar"]).max() < 1e-8
    assert np.abs(info["qbar"] - g["qbar"]).max() < 1e-8
