"""TEST-SIDE tool (CPU): Newton / line-search iteration counts of the dense (Allegro) path, per row and
physics substep, from the kernel logic run through the warp emulator (tests/emul, -DDIAL_EMUL_TRACE).
    python scripts/allegro_iteration_trace.py [rows=14] [noise level i: sigma * tdf**i]
Writes scratch/its_<rows>_<i>.npy / lss_<rows>_<i>.pkl; profiles/r02_allegro_lockstep_model.md is built on them."""
import sys, time, ctypes as C
import numpy as np
sys.path.insert(0, '.')
from baseline_configs import BASELINE, product_env, dial_config
from tests.emul import emul
from dial_mpc_b200.core.dial_core import MBDPI
from dial_mpc_b200.utils.spline import interp_matrix

b = BASELINE[3]
env = product_env(b["env"])
cfg = dial_config(3)
nrows = int(sys.argv[1]) if len(sys.argv) > 1 else 14
noise_pow = int(sys.argv[2]) if len(sys.argv) > 2 else 0
import torch
st = env.reset(torch.zeros(2, dtype=torch.uint32)) if False else None
# initial state from the oracle-free host env: reset state
from tests.conftest import make_pair
envp, o = make_pair("allegro_reorient")
s = o.reset()
Hs, Hn = b["Hs"], b["Hn"]
sigma = (b["hdf"] ** np.arange(Hn + 1)[::-1]) * 1.0
noise = sigma * b["tdf"] ** noise_pow
step_us = np.linspace(0, 0.02 * Hs, Hs + 1); step_nodes = np.linspace(0, 0.02 * Hs, Hn + 1)
desc = env.plan_desc(Nsample=nrows - 1, Hsample=Hs, Hnode=Hn, temp_sample=0.05, M_n2u=interp_matrix(step_nodes, step_us))
lib = emul.build(defines=("DIAL_EMUL_TRACE",))
t0 = time.time()
out = emul.rollout(env, desc, s.qpos[0], s.qvel[0], s.qacc_warmstart[0], Ybar=np.zeros((Hn + 1, 16)), noise=noise,
                   key=(0, 1), mode=1, nrows=nrows, H=Hs + 1, defines=("DIAL_EMUL_TRACE",))
print("emul time", time.time() - t0)
buf = (C.c_int * 4000000)()
n = lib.emul_trace_take(buf, 4000000)
tr = np.array(buf[:n]).reshape(-1, 2)
# per row sequence: kind 1 (ls iterations) entries precede the kind-0 entry of their substep
rows = []
cur_ls = []
sub = []
for k, v in tr:
    if k == 1: cur_ls.append(v)
    else:
        sub.append((v, list(cur_ls))); cur_ls = []
nsub = (Hs + 1) * 4
assert len(sub) == nrows * nsub, (len(sub), nrows, nsub)
its = np.array([x[0] for x in sub]).reshape(nrows, nsub)
lss = [[sub[r * nsub + j][1] for j in range(nsub)] for r in range(nrows)]
import os; os.makedirs("scratch", exist_ok=True)
np.save(f"scratch/its_{nrows}_{noise_pow}.npy", its)
import pickle; pickle.dump(lss, open(f"scratch/lss_{nrows}_{noise_pow}.pkl", "wb"))
print("newton its mean", its.mean(), "max", its.max(), "per-substep max over rows mean", its.max(0).mean())
allls = np.array([v for r in lss for s_ in r for v in s_])
print("ls its mean", allls.mean(), "max", allls.max(), "hist", np.bincount(allls)[:55])
print("rews", out["rews"][:5])
