"""TEST-SIDE tool (CPU): Newton / line-search iteration counts of the dense (Allegro) path, per row and
physics substep, from the kernel logic run through the warp emulator (tests/emul, -DDIAL_EMUL_TRACE).
    python scripts/allegro_iteration_trace.py [rows=14] [noise level i: sigma * tdf**i]
Writes scratch/its_<rows>_<i>.npy / lss_<rows>_<i>.pkl (read by scripts/allegro_schedule_model.py);
profiles/r02_allegro_lockstep_model.md is built on them."""
import ctypes as C
import os
import pickle
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baseline_configs import BASELINE, product_env  # noqa: E402
from dial_mpc_b200.utils.spline import interp_matrix  # noqa: E402
from tests.conftest import make_pair  # noqa: E402
from tests.emul import emul  # noqa: E402


def main():
    b = BASELINE[3]
    env = product_env(b["env"])
    nrows = int(sys.argv[1]) if len(sys.argv) > 1 else 14
    level = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    _, o = make_pair("allegro_reorient")
    s = o.reset()
    Hs, Hn = b["Hs"], b["Hn"]
    noise = (b["hdf"] ** np.arange(Hn + 1)[::-1]) * b["tdf"] ** level
    step_us, step_nodes = np.linspace(0, 0.02 * Hs, Hs + 1), np.linspace(0, 0.02 * Hs, Hn + 1)
    desc = env.plan_desc(Nsample=nrows - 1, Hsample=Hs, Hnode=Hn, temp_sample=0.05, M_n2u=interp_matrix(step_nodes, step_us))
    lib = emul.build(defines=("DIAL_EMUL_TRACE",))
    t0 = time.time()
    out = emul.rollout(env, desc, s.qpos[0], s.qvel[0], s.qacc_warmstart[0], Ybar=np.zeros((Hn + 1, 16)), noise=noise,
                       key=(0, 1), mode=1, nrows=nrows, H=Hs + 1, defines=("DIAL_EMUL_TRACE",))
    print(f"emulated {nrows} rows in {time.time() - t0:.0f} s; mean rewards {out['rews'][:4]}")
    buf = (C.c_int * 4000000)()
    n = lib.emul_trace_take(buf, 4000000)
    tr = np.array(buf[:n]).reshape(-1, 2)
    # kind 1 entries (line-search iterations of one Newton iteration) precede the kind-0 entry of their substep
    sub, cur = [], []
    for kind, v in tr:
        if kind == 1:
            cur.append(int(v))
        else:
            sub.append((int(v), cur))
            cur = []
    nsub = (Hs + 1) * 4
    assert len(sub) == nrows * nsub, (len(sub), nrows, nsub)
    its = np.array([x[0] for x in sub]).reshape(nrows, nsub)
    lss = [[sub[r * nsub + j][1] for j in range(nsub)] for r in range(nrows)]
    os.makedirs("scratch", exist_ok=True)
    np.save(f"scratch/its_{nrows}_{level}.npy", its)
    pickle.dump(lss, open(f"scratch/lss_{nrows}_{level}.pkl", "wb"))
    allls = np.array([v for r in lss for s_ in r for v in s_])
    print(f"Newton iterations per substep: mean {its.mean():.2f}, max {its.max()}, slowest row per substep {its.max(0).mean():.2f}")
    print(f"line-search iterations: mean {allls.mean():.1f}, max {allls.max()}, histogram {np.bincount(allls)[:55]}")


if __name__ == "__main__":
    main()
