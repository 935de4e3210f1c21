import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
np.set_printoptions(precision=6, suppress=True, linewidth=220)
from tests.conftest import make_pair
from dial_mpc_b200 import random as drandom
env, o = make_pair("allegro_reorient")
s = o.reset()
st = env.reset(drandom.PRNGKey(0))
print("reset qpos err", np.abs(st.pipeline_state.qpos.cpu().numpy()-s.qpos[0]).max(), "warm rel err", np.abs(st.pipeline_state.qacc_warmstart.cpu().numpy()-s.qacc_warmstart[0]).max()/np.abs(s.qacc_warmstart[0]).max())
rng = np.random.default_rng(1)
B, H = 6, 6
us = np.clip(rng.normal(size=(B, H, 16)) * 0.3, -1, 1)
rew, q, qd, x = o.rollout(s, us)
rg, qg, qdg, xg = env._get_plan().rollout(st, us)
torch.cuda.synchronize()
qg = qg.cpu().numpy(); rg = rg.cpu().numpy(); qdg = qdg.cpu().numpy()
print("q err per (row, step)\n", np.abs(qg - q).max(-1))
print("qd err per (row, step)\n", np.abs(qdg - qd).max(-1))
print("rew gpu\n", rg, "\nrew oracle\n", rew)
print("nan count", np.isnan(qg).sum())
