import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dial_mpc_b200.envs as E
from dial_mpc_b200 import random as drandom, _capi
from dial_mpc_b200.core.dial_config import DialConfig
from dial_mpc_b200.core.dial_core import MBDPI
env = E.get_environment("allegro_reorient", config=E.AllegroReorientEnvConfig(dt=0.02, timestep=0.005, leg_control="position"))
N = int(sys.argv[1])
mb = MBDPI(DialConfig(env_name="allegro_reorient", Nsample=N, Hsample=20, Hnode=4, temp_sample=0.05, horizon_diffuse_factor=1.0), env)
st = env.reset(drandom.PRNGKey(0))
Y = torch.zeros(5, 16, device=mb.device); key = drandom.PRNGKey(1)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    mb.plan.reverse_rollout(st, None, key, Y, mb.sigma_control, mb._rews_local)
    torch.cuda.synchronize(); dt = time.time() - t0
out = (C.c_float * 8)()
if _capi.lib().dial_debug_counters(mb.plan.handle, out) == 0:
    print(f"physics steps {out[0]:.0f} newton iterations {out[1]:.0f} -> {out[1]/max(out[0],1):.2f} per step")
print(f"N={N} WPC={os.environ.get('DIAL_WPC','auto')} lockstep={'off' if os.environ.get('DIAL_NO_LOCKSTEP') else 'on'}: {dt*1e3:.1f} ms -> {N*20/dt:.3e} sample-steps/s")
z = torch.zeros_like(mb.sigma_control)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    mb.plan.reverse_rollout(st, None, key, Y, z, mb._rews_local)
    torch.cuda.synchronize(); dt = time.time() - t0
print(f"  zero-noise (all rows = nominal): {dt*1e3:.2f} ms")
small = mb.sigma_control * 0.2
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    mb.plan.reverse_rollout(st, None, key, Y, small, mb._rews_local)
    torch.cuda.synchronize(); dt = time.time() - t0
print(f"  0.2 x noise: {dt*1e3:.2f} ms")
