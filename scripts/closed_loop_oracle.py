"""Closed loop on the fp64 ORACLE (CPU, small): DIAL-MPC on the Allegro scene with the planner centred on the
action that holds the reset pose, under the restated MJX line search (A) and under the narrowing bracket (B,
what -DDIAL_ROBUST_LS compiles into the kernels) — DESIGN.md 2.  N = 48 samples, Hsample = 10, Hnode = 3,
Ndiffuse = 2, `n_steps` control steps; the same noise for both variants.
    python scripts/closed_loop_oracle.py [n_steps=8]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle.mjx_oracle as mo  # noqa: E402
from baseline_configs import ENV_CFG  # noqa: E402
from oracle.envs_oracle import make_env  # noqa: E402
from oracle.planner_oracle import PlannerOracle  # noqa: E402


def run(narrowing: bool, n_steps: int):
    mo.LS_NARROWING = narrowing
    try:
        o = make_env("allegro_reorient", ENV_CFG["allegro_reorient"])
        pl = PlannerOracle(o, 48, 10, 3, 0.05, 1.0, 0.5, sigma_scale=0.15)
        jr = np.asarray(o.joint_range)
        hold = 2 * (-jr[:, 0] / (jr[:, 1] - jr[:, 0])) - 1
        Y = np.repeat(hold[None], pl.Hn + 1, 0)
        s = o.reset()
        rng = np.random.default_rng(0)
        out = []
        for t in range(n_steps):
            for i in range(2):
                eps = rng.standard_normal((pl.N, pl.Hn + 1, o.nu))
                Y, info = pl.reverse_once(s, eps, Y, pl.schedule(2)[i])
            s, r, _ = o.step(s, Y[0][None])
            # keep the plan centred on the hold action beyond the horizon (the reference shifts in zeros = "curl")
            u = np.roll(pl.node2u(Y), -1, axis=0)
            u[-1] = hold
            Y = pl.u2node(u)
            out.append((t, float(r[0]), s.qpos[0, :3].copy(), float(np.linalg.norm(s.qvel[0, :3]))))
            print(f"{'B narrowing' if narrowing else 'A restated '} step {t}: reward {r[0]:9.3f}  ball {np.round(s.qpos[0, :3], 3)}  |v| {out[-1][3]:.2f}", flush=True)
        return out
    finally:
        mo.LS_NARROWING = False


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    t0 = time.time()
    for nar in (False, True):
        run(nar, n)
    print(f"{time.time() - t0:.0f} s")
