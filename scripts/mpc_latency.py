"""Per-MPC-step latency of the synchronous loop (what a 50 Hz controller sees): env.step +
shift + Ndiffuse x reverse_once + device->host copy of the first action, wall clock per step
vs. the GPU time of the same step (CUDA events)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, yaml
import dial_mpc_b200.envs as E
from dial_mpc_b200 import random as drandom
from dial_mpc_b200.core.dial_config import DialConfig
from dial_mpc_b200.core.dial_core import MBDPI
from dial_mpc_b200.utils.io_utils import get_example_path, load_dataclass_from_dict
ex = sys.argv[1]; nsteps = int(sys.argv[2]); bars = (sys.argv[3] != "nobars") if len(sys.argv) > 3 else True
cfgd = yaml.safe_load(open(get_example_path(ex + ".yaml")))
dc = load_dataclass_from_dict(DialConfig, cfgd)
ec = load_dataclass_from_dict(E.get_config(dc.env_name), cfgd, convert_list_to_array=True)
env = E.get_environment(dc.env_name, config=ec)
mb = MBDPI(dc, env, compute_bars=bars)
rng = drandom.PRNGKey(dc.seed)
rng, r0 = drandom.split(rng)
state = env.reset(r0)
Y = torch.zeros(dc.Hnode + 1, mb.nu, device=mb.device)
sched = mb.schedule(dc.Ndiffuse)
wall, gpu = [], []
for t in range(nsteps + 10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    state = env.step(state, Y[0])
    Y = mb.shift(Y)
    rng, Y, info = mb.reverse_scan(state, rng, Y, sched)
    a0 = Y[0].cpu()
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize()
    if t >= 10:
        wall.append((t1 - t0) * 1e3); gpu.append(e0.elapsed_time(e1))
print(f"{ex} N={dc.Nsample} Hs={dc.Hsample} Ndiffuse={dc.Ndiffuse} bars={bars}: wall median {np.median(wall):.3f} ms "
      f"(p90 {np.percentile(wall, 90):.3f}), gpu-event median {np.median(gpu):.3f} ms -> {1e3/np.median(wall):.0f} Hz")
