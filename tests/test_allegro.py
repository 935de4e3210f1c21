"""Allegro reorient (BASELINE configs[3]): model compiler (mesh inertias), elliptic-cone oracle
pieces validated by finite differences, and the dense solver path of the kernel under the CPU
warp emulator.  GPU parity is in test_gpu_parity.py."""
import numpy as np
import pytest

from oracle import mjx_oracle as mo
from tests.conftest import make_pair


def test_mesh_mass_properties_of_a_box():
    from dial_mpc_b200.modelc.mjcf import _mesh_mass_properties, _primitive_mass_properties, GEOM_BOX
    o = np.array([0.3, -0.2, 0.5])
    v = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float) * [1, 2, 3] + o
    f = [(0, 2, 1), (0, 3, 2), (4, 5, 6), (4, 6, 7), (0, 1, 5), (0, 5, 4), (1, 2, 6), (1, 6, 5), (2, 3, 7), (2, 7, 6), (3, 0, 4), (3, 4, 7)]
    tri = np.array([[v[i] for i in t] for t in f])
    vol, com, I = _mesh_mass_properties(tri)
    vb, Ib = _primitive_mass_properties(GEOM_BOX, np.array([0.5, 1.0, 1.5]))
    assert abs(vol - 6) < 1e-12 and np.abs(com - (o + [0.5, 1, 1.5])).max() < 1e-12
    assert abs(vb - 6) < 1e-12 and np.abs(I - Ib).max() < 1e-12
    # flipped orientation gives the same answer
    vol2, com2, I2 = _mesh_mass_properties(tri[:, ::-1])
    assert abs(vol2 - 6) < 1e-12 and np.abs(I2 - I).max() < 1e-12


def test_allegro_model_dimensions():
    env, o = make_pair("allegro_reorient")
    m = o.m
    assert (m.nq, m.nv, m.nu, m.nbody, m.ncon, m.nlim) == (23, 22, 16, 23, 19, 16)
    assert m.cone == 1 and m.impratio == 10.0 and m.eulerdamp and (m.iterations, m.ls_iterations) == (100, 50)
    assert list(m.con_rows) == [3] * 14 + [6] * 5 and m.nefc == 88
    assert np.allclose(m.pair_friction[-1], [0.7, 0.7, 0.01, 0.01, 0.01])      # ball priority wins
    assert abs(m.body_mass[1] - 0.01) < 1e-12 and 0.2 < m.body_mass.sum() < 1.0
    assert env._n_frames == 4 and abs(env.dt - 0.02) < 1e-12
    a = np.random.default_rng(0).uniform(-1, 1, 16)
    assert np.abs(env.act2joint(a) - o.act2joint(a)).max() < 1e-12


@pytest.mark.parametrize("dim", [3, 6])
def test_cone_cost_force_hessian_finite_differences(dim):
    class FakeM:
        pass
    rng = np.random.default_rng(dim)
    m = FakeM()
    mu = 0.7 / np.sqrt(10)
    fri = np.array([0.7, 0.7, 0.01, 0.01, 0.01])[:dim - 1]
    m._cones = [(0, dim, mu, fri)]
    m.nefc = dim
    B = 3000
    Dn = rng.uniform(1, 5, B)
    D = np.zeros((B, dim))
    D[:, 0] = Dn
    for i in range(1, dim):
        D[:, i] = Dn * 10 * fri[i - 1] ** 2 / fri[0] ** 2
    x = rng.normal(size=(B, dim)) * np.array([1] + [0.5 / f for f in fri]) * 0.3
    sel = np.arange(B) % 3
    x[sel == 1, 0] = -np.abs(x[sel == 1, 0]) - 0.05   # deep penetration, tiny slip -> bottom zone
    x[sel == 1, 1:] *= 0.01
    x[sel == 2, 0] = np.abs(x[sel == 2, 0]) + 0.3      # separating -> top zone

    def cf(xx):
        ctx = mo._Ctx()
        ctx.Jaref, ctx.Ma, ctx.qacc, ctx.cost = xx.copy(), np.zeros((B, dim)), np.zeros((B, dim)), np.zeros(B)
        mo._update_constraint(m, ctx, np.eye(dim)[None].repeat(B, 0), D, np.zeros((B, dim)), np.zeros((B, dim)))
        return ctx.cost.copy(), ctx.efc_force.copy()
    c0, f0 = cf(x)
    _, N, U, T, bottom, middle, Dm = mo._cone_state(m._cones[0], x, D)
    top = ~bottom & ~middle
    assert top.mean() > 0.02 and middle.mean() > 0.2 and bottom.mean() > 0.02
    assert np.all(c0[top] == 0) and np.all(c0 >= 0)
    h = 1e-6
    g = np.zeros((B, dim))
    for i in range(dim):
        xp, xm = x.copy(), x.copy()
        xp[:, i] += h
        xm[:, i] -= h
        g[:, i] = (cf(xp)[0] - cf(xm)[0]) / (2 * h)
    err = np.abs(g + f0) / (1 + np.abs(f0))
    assert np.quantile(err, 0.99) < 1e-5          # force = -dcost/dx (zone-boundary samples excluded by the quantile)
    Hc = mo.cone_hessian(m._cones[0], x, D)
    Hfd = np.zeros((B, dim, dim))
    for i in range(dim):
        xp, xm = x.copy(), x.copy()
        xp[:, i] += h
        xm[:, i] -= h
        Hfd[:, :, i] = -(cf(xp)[1] - cf(xm)[1]) / (2 * h)
    e = np.abs(Hc[middle] - Hfd[middle]).reshape(middle.sum(), -1).max(1) / (1 + np.abs(Hc[middle]).reshape(middle.sum(), -1).max(1))
    assert np.quantile(e, 0.95) < 1e-5


def test_sphere_capsule_and_capsule_capsule_collisions():
    # capsule along z at origin (r=.1, half=.5); sphere r=.2 at x=.5  -> gap .2 ; frame normal +-x
    d, p, fr = mo.sphere_sphere(np.array([[0.5, 0, 0.2]]), 0.2, mo.closest_segment_point(np.array([[0, 0, -.5]]), np.array([[0, 0, .5]]), np.array([[0.5, 0, 0.2]])), 0.1)
    assert abs(d[0] - 0.2) < 1e-6 and np.allclose(fr[0, 0], [-1, 0, 0], atol=1e-5) and np.allclose(p[0], [0.2, 0, 0.2], atol=1e-5)
    # crossed capsules (axes x and y) separated by 0.3 in z
    a, b = mo.closest_segment_to_segment_points(np.array([[-1., 0, 0]]), np.array([[1., 0, 0]]), np.array([[0, -1., .3]]), np.array([[0, 1., .3]]))
    assert np.allclose(a[0], [0, 0, 0], atol=1e-5) and np.allclose(b[0], [0, 0, .3], atol=1e-5)
    # parallel offset segments: closest points clipped to the ends
    a, b = mo.closest_segment_to_segment_points(np.array([[0., 0, 0]]), np.array([[1., 0, 0]]), np.array([[2., 1, 0]]), np.array([[3., 1, 0]]))
    assert np.allclose(a[0], [1, 0, 0], atol=1e-4) and np.allclose(b[0], [2, 1, 0], atol=1e-4)


def test_allegro_oracle_physics_sanity():
    env, o = make_pair("allegro_reorient")
    s = o.reset()
    # hold the initial pose: targets = init pose -> action that maps to init_q; the ball must stay in hand
    jr = o.joint_range
    a = ((o.init_q[7:] - jr[:, 0] - o.init_q[7:]) / (jr[:, 1] - jr[:, 0])) * 2 - 1
    rews = []
    for _ in range(10):
        s, r, aux = o.step(s, np.clip(a, -1, 1)[None])
        rews.append(r[0])
    assert np.isfinite(rews).all() and s.qpos[0, 2] > 0.05          # not fallen through the fingers
    d = aux["data"]
    f_active = (d.con_dist[0] < 0).sum()
    assert f_active >= 1 and d.solver_niter[0] <= o.m.iterations


def test_emulated_dense_path_matches_oracle():
    from tests.emul import emul
    env, o = make_pair("allegro_reorient")
    s = o.reset()
    rng = np.random.default_rng(3)
    us = np.clip(rng.normal(size=(2, 5, 16)) * 0.4, -1, 1)
    rew, q, qd, x = o.rollout(s, us)
    out = emul.rollout(env, env.plan_desc(), s.qpos[0], s.qvel[0], s.qacc_warmstart[0], us=us)
    assert np.abs(out["q"] - q).max() < 5e-4
    assert np.abs(out["rewss"] - rew).max() < 2e-3 * (1 + np.abs(rew).max())


def test_line_search_cycle_detection_is_bit_identical():
    """dense_linesearch stops MJX's 50-iteration bracket loop early once its (lo, hi, swap) state
    has become periodic and replays only the remainder modulo the period (Brent).  The claim is
    that the result is the state after the full budget, bit for bit: build the device code with
    the detection compiled out and compare every output of a contact-rich rollout."""
    from tests.emul import emul
    env, o = make_pair("allegro_reorient")
    s = o.reset()
    rng = np.random.default_rng(3)
    us = np.clip(rng.normal(size=(2, 8, 16)) * 0.6, -1, 1)
    a = emul.rollout(env, env.plan_desc(), s.qpos[0], s.qvel[0], s.qacc_warmstart[0], us=us)
    b = emul.rollout(env, env.plan_desc(), s.qpos[0], s.qvel[0], s.qacc_warmstart[0], us=us,
                     defines=("DIAL_NO_LS_CYCLE",))
    for k in ("rewss", "q", "qd", "xpos", "warm_out", "qvel_out"):
        assert np.array_equal(a[k], b[k]), k
    assert np.abs(a["qd"]).max() > 0.5     # the ball is being pushed around: contacts are active


def test_emulated_dense_path_random_states():
    """Dense / elliptic path away from the reset pose: random finger configurations (some joints
    beyond their limits), displaced and spinning ball, two env steps (8 physics substeps)."""
    from oracle.envs_oracle import OState
    from tests.emul import emul
    env, o = make_pair("allegro_reorient")
    s0 = o.reset()
    nv, nu = o.m.nv, o.m.nu
    rng = np.random.default_rng(5)
    active = 0
    for _ in range(6):
        q = s0.qpos[0].copy()
        q[:3] += rng.normal(size=3) * 0.004
        lo, hi = o.physical_joint_range[:, 0], o.physical_joint_range[:, 1]
        q[7:7 + nu] = np.clip(q[7:7 + nu] + rng.normal(size=nu) * 0.15, lo - 0.02, hi + 0.02)
        v = np.r_[rng.normal(size=3) * 0.05, rng.normal(size=3), rng.normal(size=nv - 6)]
        st = OState(q[None], v[None], np.zeros((1, nv)), np.array([3]), np.array([0]))
        us = np.clip(rng.normal(size=(1, 2, nu)) * 0.5, -1, 1)
        rew, qq, qd, x = o.rollout(st, us)
        out = emul.rollout(env, env.plan_desc(), q, v, np.zeros(nv), us=us, step0=3)
        assert np.abs(out["q"] - qq).max() < 2e-4
        assert np.abs(out["qd"] - qd).max() < 5e-3
        assert np.abs(out["rewss"] - rew).max() < 1e-3 * (1 + np.abs(rew).max())
        active += int(np.abs(out["qd"] - qd).max() > 1e-4)   # contact-rich trials show fp32 noise
    assert active >= 1


def test_line_search_stall_on_first_fingertip_contact():
    """Characterisation of the restated solver (scripts/ls_stall_trace.py): when the falling ball
    first meets a fingertip, the second Newton direction overshoots a cone-zone boundary (cost at alpha = 1
    an order of magnitude above the cost at 0, minimum near alpha = 0.1), MJX's bracketing line search
    returns "no improvement", the solver stops far from convergence, and the force-based eulerdamp
    integration kicks the 10 g ball.  The kernels follow the oracle through this (bit-identical line-search
    state); what is pinned here is that the behaviour comes from the algorithm, not from rounding."""
    from oracle import mjx_oracle as mo
    env, o = make_pair("allegro_reorient")
    s = o.reset()
    m = o.m
    jr = np.asarray(o.joint_range)
    hold = 2 * (-jr[:, 0] / (jr[:, 1] - jr[:, 0])) - 1      # joint targets = reset pose
    scale = m.meaninertia * max(1, m.nv)
    qpos, qvel, warm = s.qpos.copy(), s.qvel.copy(), s.qacc_warmstart.copy()
    ctrl = o.act2joint(hold[None])
    seen = None
    for sub in range(12):
        d = mo.forward(m, qpos, qvel, ctrl, warm)
        g = np.linalg.norm((np.einsum("nvw,nw->nv", d.M, d.qacc) - d.qfrc_smooth - d.qfrc_constraint)[0]) / scale
        if (d.con_dist[0] < 0).any():
            seen = (sub, g, int(d.solver_niter[0]), qpos.copy(), qvel.copy(), warm.copy())
            break
        assert g < 1e-8                                   # free flight: nothing to solve
        qpos, qvel, warm, _ = mo.step(m, qpos, qvel, ctrl, warm)
    assert seen is not None, "the ball never reached a fingertip"
    sub, g, niter, qpos, qvel, warm = seen
    assert niter < m.iterations and g > 1.0               # stopped early, |grad| / scale 1e8 above the tolerance
    # the stall itself: second Newton direction, true cost along it, and what the line search makes of it
    d = mo.forward(m, qpos, qvel, ctrl, warm)
    M, J, D, aref, qs, qa = d.M, d.efc_J, d.efc_D, d.efc_aref, d.qfrc_smooth, d.qacc_smooth
    wc = mo._ctx_create(m, M, J, D, aref, qs, qa, warm, grad=False)
    sc = mo._ctx_create(m, M, J, D, aref, qs, qa, qa, grad=False)
    ctx = mo._ctx_create(m, M, J, D, aref, qs, qa, np.where((wc.cost < sc.cost)[:, None], warm, qa))
    mo._linesearch(m, ctx, M, J, D, qs)
    assert ctx.ls_alpha[0] > 0.5                           # first iteration: a normal Newton step
    mo._update_constraint(m, ctx, J, D, qs, qa)
    mo._update_gradient(m, ctx, M, J, D, qs)
    ctx.search = -ctx.Mgrad
    assert ctx.search[0] @ ctx.grad[0] < 0                 # a descent direction ...

    def cost_at(alpha):
        c = mo._Ctx()
        c.qacc = ctx.qacc + alpha * ctx.search
        c.Jaref = np.einsum("nrv,nv->nr", J, c.qacc) - aref
        c.Ma = np.einsum("nvw,nw->nv", M, c.qacc)
        c.cost, c.prev_cost = np.zeros(1), np.zeros(1)
        mo._update_constraint(m, c, J, D, qs, qa)
        return c.cost[0]
    c0, c01, c1 = cost_at(0.0), cost_at(0.1), cost_at(1.0)
    assert c01 < c0 < c1 and c1 > 5 * c0                   # ... that improves for small steps and overshoots at alpha = 1
    mo._linesearch(m, ctx, M, J, D, qs)
    assert ctx.ls_alpha[0] == 0.0                          # MJX's bracket rule gives up: no step, solver stops
    # consequence: the next substep integrates the forces of the unconverged point
    v0 = qvel[0, :3].copy()
    _, qvel1, _, _ = mo.step(m, qpos, qvel, ctrl, warm)
    assert np.linalg.norm(qvel1[0, :3] - v0) > 0.5         # > 0.5 m/s in one 5 ms substep on a 10 g ball


def test_robust_line_search_build_keeps_the_ball_and_matches_its_oracle():
    """-DDIAL_ROBUST_LS (opt-in builds; a documented deviation from MJX's bracket rule): the device
    code with the narrowing bracket against the oracle with the same rule (LS_NARROWING), through the first
    fingertip contact of the hold action — where the reference rule stalls and kicks the ball (test above), this
    one converges and the ball settles on the fingertips."""
    from oracle import mjx_oracle as mo
    from tests.emul import emul
    env, o = make_pair("allegro_reorient")
    s = o.reset()
    jr = np.asarray(o.joint_range)
    hold = 2 * (-jr[:, 0] / (jr[:, 1] - jr[:, 0])) - 1
    us = np.repeat(hold[None, None], 5, axis=1)                       # 5 env steps = 20 substeps: contact at step 2
    mo.LS_NARROWING = True
    try:
        rew, q, qd, x = o.rollout(s, us)
    finally:
        mo.LS_NARROWING = False
    assert np.linalg.norm(qd[0, :, :3], axis=-1).max() < 0.6          # no kick (the reference rule: 1.6 m/s)
    assert q[0, -1, 2] > 0.12                                         # the ball rests on the fingertips
    out = emul.rollout(env, env.plan_desc(), s.qpos[0], s.qvel[0], s.qacc_warmstart[0], us=us, defines=("DIAL_ROBUST_LS",))
    assert np.abs(out["q"] - q).max() < 5e-4
    assert np.abs(out["rewss"] - rew).max() < 2e-3 * (1 + np.abs(rew).max())
    # and the stock rule does kick it, in the oracle and in the device code alike
    rew0, q0, qd0, _ = o.rollout(s, us)
    out0 = emul.rollout(env, env.plan_desc(), s.qpos[0], s.qvel[0], s.qacc_warmstart[0], us=us)
    assert np.linalg.norm(qd0[0, :, :3], axis=-1).max() > 1.0 and np.linalg.norm(out0["qd"][0, :, :3], axis=-1).max() > 1.0
