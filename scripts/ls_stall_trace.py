"""Why sampled rollouts kick the ball out of the hand: MJX's bracketing line search stalls on
the substep of a fresh pinch, the Newton solver stops unconverged, and the force-based eulerdamp integration
applies the forces of that point.  CPU only (fp64 oracle):
    python scripts/ls_stall_trace.py            # Allegro reset state, hold action
Prints per physics substep: solver iterations, |grad| / scale at the returned point (tolerance 1e-8), the ball's
velocity and the constraint force on it; then, for the first stalled substep, the true cost along the Newton
direction next to the line search's 1-D model."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle.mjx_oracle as mo  # noqa: E402
from baseline_configs import ENV_CFG  # noqa: E402
from oracle.envs_oracle import make_env  # noqa: E402


def main():
    o = make_env("allegro_reorient", ENV_CFG["allegro_reorient"])
    s = o.reset()
    m = o.m
    jr = np.asarray(o.joint_range)
    hold = 2 * (-jr[:, 0] / (jr[:, 1] - jr[:, 0])) - 1          # action whose joint target is the reset pose
    scale = m.meaninertia * max(1, m.nv)
    qpos, qvel, warm = s.qpos.copy(), s.qvel.copy(), s.qacc_warmstart.copy()
    np.set_printoptions(precision=3, suppress=True, linewidth=200)
    stalled = None
    for t in range(4):
        ctrl = o.act2joint(hold[None])
        for f in range(o.n_frames):
            d = mo.forward(m, qpos, qvel, ctrl, warm)
            grad = np.einsum("nvw,nw->nv", d.M, d.qacc) - d.qfrc_smooth - d.qfrc_constraint
            g = np.linalg.norm(grad[0]) / scale
            if g > 1.0 and stalled is None:
                stalled = (qpos.copy(), qvel.copy(), warm.copy(), ctrl.copy())
            qpos, qvel, warm, _ = mo.step(m, qpos, qvel, ctrl, warm)
            print(f"env step {t} substep {f}: Newton iterations {int(d.solver_niter[0])}, |grad|/scale {g:9.3e}, "
                  f"contacts {(d.con_dist[0] < 0).sum()}, ball v {qvel[0, :3]}, constraint force on the ball {d.qfrc_constraint[0, :3]}")
    if stalled is None:
        return
    qpos, qvel, warm, ctrl = stalled
    d = mo.forward(m, qpos, qvel, ctrl, warm)
    M, J, D, aref, qs, qa = d.M, d.efc_J, d.efc_D, d.efc_aref, d.qfrc_smooth, d.qacc_smooth
    wc = mo._ctx_create(m, M, J, D, aref, qs, qa, warm, grad=False)
    sc = mo._ctx_create(m, M, J, D, aref, qs, qa, qa, grad=False)
    ctx = mo._ctx_create(m, M, J, D, aref, qs, qa, np.where((wc.cost < sc.cost)[:, None], warm, qa))
    print("\nfirst stalled substep, iteration by iteration:")

    def true_cost(c0, alpha):
        c = mo._Ctx()
        c.qacc = c0.qacc + alpha * c0.search
        c.Jaref = np.einsum("nrv,nv->nr", J, c.qacc) - aref
        c.Ma = np.einsum("nvw,nw->nv", M, c.qacc)
        c.cost, c.prev_cost = np.zeros(1), np.zeros(1)
        mo._update_constraint(m, c, J, D, qs, qa)
        return c.cost[0]
    for it in range(m.iterations):
        improvement = (ctx.prev_cost[0] - ctx.cost[0]) / scale
        gradient = np.linalg.norm(ctx.grad[0]) / scale
        print(f"  iteration {it}: cost {ctx.cost[0]:.4f}, |grad|/scale {gradient:.3e}, improvement/scale {improvement:.3e}, "
              f"search . grad {ctx.search[0] @ ctx.grad[0]:.3f}")
        if improvement < m.tolerance or gradient < m.tolerance:
            print("  -> solver stops (improvement or gradient below tolerance 1e-8)")
            break
        scan = ", ".join(f"{a:g}: {true_cost(ctx, a):.3f}" for a in (0.0, 0.02, 0.1, 0.3, 0.6, 1.0))
        mo._linesearch(m, ctx, M, J, D, qs)
        print(f"     cost along the Newton direction (alpha: cost) {scan};  line search returns alpha = {float(ctx.ls_alpha[0]):.4f}")
        mo._update_constraint(m, ctx, J, D, qs, qa)
        mo._update_gradient(m, ctx, M, J, D, qs)
        ctx.search = -ctx.Mgrad


if __name__ == "__main__":
    main()
