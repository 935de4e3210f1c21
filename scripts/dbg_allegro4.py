import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
np.set_printoptions(precision=4, suppress=True, linewidth=200)
from tests.conftest import make_pair
from dial_mpc_b200 import random as drandom
from dial_mpc_b200.core.dial_config import DialConfig
from dial_mpc_b200.core.dial_core import MBDPI
env, o = make_pair("allegro_reorient")
jr = o.joint_range
hold = np.clip(((-jr[:, 0]) / (jr[:, 1] - jr[:, 0])) * 2 - 1, -1, 1)
st = env.reset(drandom.PRNGKey(0))
for t in range(40):
    st = env.step(st, hold)
    if t % 8 == 0 or t == 39:
        print(t, "ball", st.pipeline_state.qpos[:3].cpu().numpy(), "rew %.3f" % float(st.reward))
# planner from the settled state: does reverse_once improve on the hold plan?
cfg = DialConfig(env_name="allegro_reorient", Nsample=2048, Hsample=20, Hnode=4, temp_sample=0.05, horizon_diffuse_factor=1.0)
mb = MBDPI(cfg, env)
Y = torch.tensor(np.tile(hold, (5, 1)), dtype=torch.float32, device=mb.device)
rng = drandom.PRNGKey(0)
for i in range(6):
    rng, Y, info = mb.reverse_once(st, rng, Y, mb.sigma_control * 0.5 ** i)
    r = info["rews"].cpu().numpy()
    print(f"iter {i}: mean-row rew {r[-1]:.4f}  best {np.nanmax(r):.4f}  median {np.nanmedian(r):.4f}  max w {float(info['weights'].max()):.3f}  finite {np.isfinite(r).mean():.3f}")
print("Y[1]-hold", (Y[1].cpu().numpy() - hold).round(2))
