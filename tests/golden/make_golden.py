"""Generate the golden fixtures under tests/golden/ from the fp64 oracle.

The reference's own implementation (JAX + Brax + MuJoCo-MJX) cannot be imported in this
environment and the reference tree holds no golden vectors for this path, so these
fixtures freeze the ORACLE's outputs ("parity unpinned", see oracle/mjx_oracle.py).  They
pin (i) the oracle against accidental edits and (ii) the CUDA path on the GPU box, where
the oracle is also available but the fixtures make the comparison independent of it.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.envs_oracle import make_env  # noqa: E402
from oracle.planner_oracle import PlannerOracle  # noqa: E402
from tests.conftest import ENV_CASES  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = {"unitree_go2_walk": (16, 4, 32), "unitree_go2_seq_jump": (25, 5, 32), "unitree_h1_walk": (30, 5, 16),
         "allegro_reorient": (8, 4, 16), "unitree_h1_loco": (20, 5, 16)}


def main():
    for name, (Hs, Hn, N) in CASES.items():
        env = make_env(name, ENV_CASES[name])
        s = env.reset()
        # settle 10 env steps with zero action (BASELINE.md §3 fixture state)
        if name == "allegro_reorient":   # hold the initial grasp for 3 env steps (zero action would open the hand)
            jr = env.joint_range
            hold = np.clip(((-jr[:, 0]) / (jr[:, 1] - jr[:, 0])) * 2 - 1, -1, 1)[None]
            for _ in range(3):
                s, _, _ = env.step(s, hold)
        else:
            for _ in range(10):
                s, _, _ = env.step(s, np.zeros((1, env.nu)))
        rng = np.random.default_rng(20260922)
        eps = rng.standard_normal((N, Hn + 1, env.nu)).astype(np.float32).astype(np.float64)
        pl = PlannerOracle(env, N, Hs, Hn, 0.05, 0.9 if "go2" in name else 1.0, 0.5)
        Ybar0 = np.clip(rng.standard_normal((Hn + 1, env.nu)) * 0.2 + (hold if name == "allegro_reorient" else 0.0), -1, 1).astype(np.float32).astype(np.float64)
        Ybar, info = pl.reverse_once(s, eps, Ybar0, pl.sigma_control)
        # the oracle's own sensitivity to fp32-sized input noise (max over three re-runs with 1e-5 noise on the actions): the yardstick
        # for chaotic rows in the GPU parity tests (tests/test_gpu_at_size.py)
        prng = np.random.default_rng(7)
        rews_sens = np.zeros_like(info["rews"])
        for _ in range(3):
            rp = env.rollout(s, info["us"] + 1e-5 * prng.standard_normal(info["us"].shape))[0]
            rews_sens = np.maximum(rews_sens, np.abs(rp.mean(-1) - info["rews"]))
        np.savez_compressed(
            os.path.join(OUT, f"{name}.npz"),
            qpos=s.qpos[0], qvel=s.qvel[0], qacc_warmstart=s.qacc_warmstart[0], step=int(s.step[0]),
            stage=int(s.stage[0]), eps=eps.astype(np.float32), Ybar0=Ybar0, noise_scale=pl.sigma_control,
            us=info["us"].astype(np.float32), rewss=info["rewss"], rews=info["rews"], weights=info["weights"],
            Ybar=Ybar, qbar=info["qbar"], qdbar=info["qdbar"], xbar=info["xbar"],
            rews_sens=rews_sens, Hs=Hs, Hn=Hn, N=N, temp=0.05)
        print(name, "rews[:3]", info["rews"][:3], "max w", info["weights"].max())


if __name__ == "__main__":
    main()
