"""Physics self-consistency of the oracle (no external reference exists): SURVEY.md §4-2."""
import os

import numpy as np
import pytest

from oracle import mjx_oracle as mo
from dial_mpc_b200.modelc import CompiledModel, mass_matrix_fp64

MODELS = os.path.join(os.path.dirname(os.path.dirname(__file__)), "dial_mpc_b200", "models")
FILES = ["unitree_go2_mjx_scene_force.json", "unitree_h1_mjx_scene_h1_walk.json"]


def _rand_state(m, rng, B=3, height=2.0):
    q = np.tile(np.array(m.keyframes["home"]["qpos"]), (B, 1))
    q[:, 2] = height
    q[:, 7:] += rng.normal(size=(B, m.nq - 7)) * 0.2
    qq = rng.normal(size=(B, 4))
    q[:, 3:7] = qq / np.linalg.norm(qq, axis=-1, keepdims=True)
    return q, rng.normal(size=(B, m.nv))


@pytest.mark.parametrize("fname", FILES)
def test_expected_dimensions(fname):
    m = mo.OModel(os.path.join(MODELS, fname))
    if "go2" in fname:
        assert (m.nq, m.nv, m.nu, m.nbody, m.ncon, m.nefc) == (19, 18, 12, 14, 4, 28)
        assert abs(m.body_mass.sum() - 16.206408) < 1e-6
        assert np.allclose(m.pair_friction[0], [1.0, 1.0, 0.02, 0.01, 0.01])
        assert np.allclose(m.pair_solimp[0], [0.4575, 0.975, 0.016, 0.5, 2.0])
        assert m.iterations == 2 and m.ls_iterations == 5 and not m.eulerdamp
    else:
        assert (m.nq, m.nv, m.nu, m.nbody, m.ncon, m.nefc) == (26, 25, 19, 21, 4, 35)
        assert np.all(m.actuator_ctrllimited == 1)


@pytest.mark.parametrize("fname", FILES)
def test_crb_mass_matrix_equals_jacobian_form(fname):
    path = os.path.join(MODELS, fname)
    m, cm = mo.OModel(path), CompiledModel.load(path)
    rng = np.random.default_rng(0)
    q, v = _rand_state(m, rng)
    d = mo.forward(m, q, v, np.zeros((3, m.nu)), np.zeros((3, m.nv)))
    for i in range(3):
        M2 = mass_matrix_fp64(cm, q[i])
        assert np.abs(M2 - d.M[i]).max() < 1e-10
        assert np.linalg.eigvalsh(d.M[i]).min() > 0


@pytest.mark.parametrize("fname", FILES)
def test_free_fall_and_momentum(fname):
    """No contact, zero control: the COM accelerates with gravity regardless of posture."""
    m = mo.OModel(os.path.join(MODELS, fname), timestep=0.001)
    m.dof_damping = np.zeros_like(m.dof_damping)
    rng = np.random.default_rng(1)
    q, v = _rand_state(m, rng, B=2)
    mass = m.body_mass

    def com_vel(q, v, dt=1e-7):
        d0 = mo.forward(m, q, v, np.zeros((2, m.nu)), np.zeros((2, m.nv)))
        q1 = mo.integrate_pos(m, d0.qpos, v, dt)
        d1 = mo.forward(m, q1, v, np.zeros((2, m.nu)), np.zeros((2, m.nv)))
        c0 = np.einsum("b,nbi->ni", mass, d0.xipos) / mass.sum()
        c1 = np.einsum("b,nbi->ni", mass, d1.xipos) / mass.sum()
        return (c1 - c0) / dt
    v0 = com_vel(q, v)
    qn, vn, _, d = mo.step(m, q, v, np.zeros((2, m.nu)), np.zeros((2, m.nv)))
    assert np.all(d.con_dist > 0)
    v1 = com_vel(qn, vn)
    acc = (v1 - v0) / m.timestep
    assert np.abs(acc - m.gravity).max() < 5e-3 * 9.81


@pytest.mark.parametrize("fname", FILES)
def test_energy_conservation_without_dissipation(fname):
    m = mo.OModel(os.path.join(MODELS, fname), timestep=2e-4)
    m.dof_damping = np.zeros_like(m.dof_damping)
    rng = np.random.default_rng(2)
    q, v = _rand_state(m, rng, B=2, height=5.0)
    v *= 0.5

    def energy(q, v):
        d = mo.forward(m, q, v, np.zeros((2, m.nu)), np.zeros((2, m.nv)))
        ke = 0.5 * np.einsum("ni,nij,nj->n", v, d.M, v)
        pe = -np.einsum("b,nbi,i->n", m.body_mass, d.xipos, m.gravity)
        return ke + pe
    e0 = energy(q, v)
    w = np.zeros((2, m.nv))
    for _ in range(200):
        q, v, w, _ = mo.step(m, q, v, np.zeros((2, m.nu)), w)
    e1 = energy(q, v)
    assert np.abs(e1 - e0).max() < 2e-3 * np.abs(e0).max()


@pytest.mark.parametrize("fname", FILES)
def test_gravity_bias_is_potential_gradient(fname):
    """qfrc_bias at zero velocity = dV/dq (finite differences on hinge coordinates)."""
    m = mo.OModel(os.path.join(MODELS, fname))
    rng = np.random.default_rng(3)
    q, _ = _rand_state(m, rng, B=1)
    z = np.zeros((1, m.nv))
    d = mo.forward(m, q, z, np.zeros((1, m.nu)), z)

    def V(qq):
        dd = mo.forward(m, qq, z, np.zeros((1, m.nu)), z)
        return -np.einsum("b,nbi,i->n", m.body_mass, dd.xipos, m.gravity)[0]
    for dof in range(6, m.nv):
        qa = 7 + (dof - 6)
        qp, qm = q.copy(), q.copy()
        qp[0, qa] += 1e-6
        qm[0, qa] -= 1e-6
        assert abs((V(qp) - V(qm)) / 2e-6 - d.qfrc_bias[0, dof]) < 1e-5 * (1 + abs(d.qfrc_bias[0, dof]))
    # free translation dofs: total weight along z
    assert abs(d.qfrc_bias[0, 2] - m.body_mass.sum() * 9.81) < 1e-8


@pytest.mark.parametrize("fname", FILES)
def test_contact_jacobian_matches_finite_difference(fname):
    """J qvel on the contact rows = contact-frame velocity of the foot point (body fixed)."""
    m = mo.OModel(os.path.join(MODELS, fname))
    rng = np.random.default_rng(4)
    q = np.array(m.keyframes["home"]["qpos"])[None].copy()
    q[:, 7:] += rng.normal(size=(1, m.nq - 7)) * 0.05
    v = rng.normal(size=(1, m.nv)) * 0.5
    z = np.zeros((1, m.nv))
    d = mo.forward(m, q, v, np.zeros((1, m.nu)), z)
    assert (d.con_dist < 0).any()
    dt = 1e-7
    q1 = mo.integrate_pos(m, d.qpos, v, dt)
    d1 = mo.forward(m, q1, v, np.zeros((1, m.nu)), z)
    c = 0
    for k in range(m.npair):
        b2 = int(m.geom_bodyid[m.pair_geom2[k]])
        for _ in range(int(m.pair_ncon[k])):
            if d.con_dist[0, c] < m.pair_margin[k]:
                loc = d.xmat[0, b2].T @ (d.con_pos[0, c] - d.xpos[0, b2])   # body-fixed point
                p1 = d1.xpos[0, b2] + d1.xmat[0, b2] @ loc
                vel = d.con_frame[0, c] @ ((p1 - d.con_pos[0, c]) / dt)
                mu = m.pair_friction[k]
                rows = d.efc_J[0, m.nlim + 4 * c: m.nlim + 4 * c + 4] @ v[0]
                ref = np.array([vel[0] + mu[0] * vel[1], vel[0] - mu[0] * vel[1],
                                vel[0] + mu[1] * vel[2], vel[0] - mu[1] * vel[2]])
                assert np.abs(rows - ref).max() < 1e-5
            c += 1


@pytest.mark.parametrize("fname", FILES)
def test_newton_solver_converges_to_kkt_point(fname):
    m = mo.OModel(os.path.join(MODELS, fname), timestep=0.02)
    rng = np.random.default_rng(5)
    q = np.array(m.keyframes["home"]["qpos"])[None].repeat(4, 0)
    q[:, 7:] += rng.normal(size=(4, m.nq - 7)) * 0.1
    v = rng.normal(size=(4, m.nv)) * 0.3
    ctrl = rng.normal(size=(4, m.nu)) * 5
    z = np.zeros((4, m.nv))
    costs = []
    for iters in (1, 2, 30):
        m.iterations = iters
        d = mo.forward(m, q, v, ctrl, z)
        jar = np.einsum("nrv,nv->nr", d.efc_J, d.qacc) - d.efc_aref
        f = d.efc_D * np.maximum(-jar, 0)
        grad = np.einsum("nvw,nw->nv", d.M, d.qacc) - d.qfrc_smooth - np.einsum("nrv,nr->nv", d.efc_J, f)
        gauss = 0.5 * np.einsum("nv,nvw,nw->n", d.qacc - d.qacc_smooth, d.M, d.qacc - d.qacc_smooth)
        costs.append(gauss + 0.5 * np.sum(d.efc_D * np.minimum(jar, 0) ** 2, -1))
    assert np.all(costs[1] <= costs[0] + 1e-9) and np.all(costs[2] <= costs[1] + 1e-9)
    assert np.abs(grad).max() < 1e-6 * (1 + np.abs(d.qfrc_smooth).max())
    assert np.all(f >= 0)   # unilateral forces on limits and pyramid edges
