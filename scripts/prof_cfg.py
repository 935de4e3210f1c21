"""Run K rollout launches of BASELINE configs[i] from the bench state (for ncu / quick timing).
    python scripts/prof_cfg.py <config index> [K] [--time]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from baseline_configs import BASELINE, dial_config, product_env  # noqa: E402
from dial_mpc_b200 import random as drandom  # noqa: E402
from dial_mpc_b200.core.dial_core import MBDPI  # noqa: E402

ci = int(sys.argv[1])
K = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 3
b = BASELINE[ci]
env = product_env(b["env"])
cfg = dial_config(ci)
mb = MBDPI(cfg, env)
st = env.reset(drandom.PRNGKey(0))
for _ in range(10):
    st = env.step(st, torch.zeros(mb.nu, device=mb.device))
Y = torch.zeros(cfg.Hnode + 1, mb.nu, device=mb.device)
key = drandom.PRNGKey(1)
noise = mb.sigma_control * float(os.environ.get("DIAL_PROF_NOISE", "1"))   # 0: identical rows (no divergence between warps)
for i in range(K):
    mb.plan.reverse_rollout(st, None, key, Y, noise, mb._rews_local)
torch.cuda.synchronize()
if "--time" in sys.argv:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20 if ci != 3 else 5
    e0.record()
    for i in range(reps):
        mb.plan.reverse_rollout(st, None, key, Y, noise, mb._rews_local)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"cfg{ci} {b['name']} N={cfg.Nsample} Hs={cfg.Hsample} WPC={os.environ.get('DIAL_WPC', 'auto')} noise x{os.environ.get('DIAL_PROF_NOISE', '1')}: rollout kernel {ms:.4f} ms "
          f"-> {cfg.Nsample * cfg.Hsample / ms * 1e3:.4e} sample-steps/s  rews {float(mb._rews_local.mean()):.5f}")
else:
    print("done", float(mb._rews_local.mean()))
