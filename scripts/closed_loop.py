"""Closed-loop sanity run: synchronous MPC loop (dial_core.main semantics) for a few steps."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, yaml
import dial_mpc_b200.envs as E
from dial_mpc_b200 import random as drandom
from dial_mpc_b200.core.dial_config import DialConfig
from dial_mpc_b200.core.dial_core import MBDPI
from dial_mpc_b200.utils.io_utils import get_example_path, load_dataclass_from_dict
ex = sys.argv[1]; nsteps = int(sys.argv[2])
cfgd = yaml.safe_load(open(get_example_path(ex + ".yaml")))
dc = load_dataclass_from_dict(DialConfig, cfgd)
ec = load_dataclass_from_dict(E.get_config(dc.env_name), cfgd, convert_list_to_array=True)
env = E.get_environment(dc.env_name, config=ec)
mb = MBDPI(dc, env)
rng = drandom.PRNGKey(dc.seed)
rng, r0 = drandom.split(rng)
state = env.reset(r0)
Y = torch.zeros(dc.Hnode + 1, mb.nu, device=mb.device)
rews = []; t0 = time.time()
for t in range(nsteps):
    state = env.step(state, Y[0])
    rews.append(float(state.reward))
    Y = mb.shift(Y)
    nd = dc.Ndiffuse_init if t == 0 else dc.Ndiffuse
    rng, Y, info = mb.reverse_scan(state, rng, Y, mb.schedule(nd))
    if t % 20 == 0 or t == nsteps - 1:
        q = state.pipeline_state.qpos.cpu().numpy(); v = state.pipeline_state.qvel.cpu().numpy()
        print(f"t={t:3d} rew={rews[-1]:8.3f} plan_rew={float(info['rews'][-1]):8.3f} x={q[0]:6.3f} z={q[2]:5.3f} vx={v[0]:6.3f}")
torch.cuda.synchronize()
print(f"{ex}: mean reward {np.mean(rews):.3f} over {nsteps} steps, {nsteps/(time.time()-t0):.1f} MPC steps/s wall")
