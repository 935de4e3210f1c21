import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.conftest import make_pair
from dial_mpc_b200 import random as drandom
from dial_mpc_b200.core.dial_config import DialConfig
from dial_mpc_b200.core.dial_core import MBDPI
env, o = make_pair("allegro_reorient")
N = 1024
mb = MBDPI(DialConfig(env_name="allegro_reorient", Nsample=N, Hsample=20, Hnode=4, temp_sample=0.05, horizon_diffuse_factor=1.0), env)
st = env.reset(drandom.PRNGKey(0))
rng = np.random.default_rng(5)
eps = rng.standard_normal((N, 5, 16)).astype(np.float32)
Y = torch.zeros(5, 16, device=mb.device)
mb.plan.reverse_rollout(st, mb.plan.f32(eps), None, Y, mb.sigma_control, mb._rews_local)
torch.cuda.synchronize()
r = mb._rews_local.cpu().numpy()
order = np.argsort(r[:N])
print("worst rows", order[:8], r[order[:8]])
print("quantiles", np.quantile(r, [0, 0.001, 0.01, 0.1, 0.5, 0.9, 1.0]))
os.makedirs("gpurun_out", exist_ok=True)
np.savez("gpurun_out/allegro_worst.npz", eps=eps[order[:6]], rews=r[order[:6]], rows=order[:6], median_rows_eps=eps[order[N//2:N//2+2]], median_rews=r[order[N//2:N//2+2]])
# per-step rewards of the worst rows through the explicit-action API
us = mb.node2u_vvmap(torch.clamp(torch.cat([torch.zeros(6,1,16,device=mb.device), mb.plan.f32(eps[order[:6]])[:,1:]*mb.sigma_control[None,1:,None]],1), -1, 1))
rewss, (q, qd, x) = mb.rollout_us_vmap(st, us)
print("per-step rewards of worst row\n", rewss[0].cpu().numpy().round(2))
print("ball pos of worst row at end", q[0, -1, :3].cpu().numpy(), "max |qvel| over horizon", float(qd[0].abs().max()))
