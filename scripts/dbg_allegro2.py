import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dial_mpc_b200.envs as E
from dial_mpc_b200 import random as drandom
from dial_mpc_b200.core.dial_config import DialConfig
from dial_mpc_b200.core.dial_core import MBDPI
env = E.get_environment("allegro_reorient", config=E.AllegroReorientEnvConfig(dt=0.02, timestep=0.005, leg_control="position"))
for N in (256, 4096):
    mb = MBDPI(DialConfig(env_name="allegro_reorient", Nsample=N, Hsample=20, Hnode=4, temp_sample=0.05, horizon_diffuse_factor=1.0), env)
    st = env.reset(drandom.PRNGKey(0))
    Y = torch.zeros(5, 16, device=mb.device); key = drandom.PRNGKey(1)
    outs = []
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        mb.plan.reverse_rollout(st, None, key, Y, mb.sigma_control, mb._rews_local)
        torch.cuda.synchronize(); dt = time.time() - t0
        r = mb._rews_local.cpu().numpy().copy(); outs.append(r)
        print(f"N={N} rep={rep}: {dt*1e3:.1f} ms, nan rows {np.isnan(r).sum()}, inf {np.isinf(r).sum()}, min {np.nanmin(r):.3f} median {np.nanmedian(r):.3f}")
    same = np.array_equal(outs[0], outs[1], equal_nan=True)
    print("  deterministic:", same, "" if same else f"differing rows {np.sum(~np.isclose(outs[0], outs[1], equal_nan=True))}")
    nanrows = np.isnan(outs[0]).nonzero()[0][:8]
    print("  first nan rows", nanrows)
