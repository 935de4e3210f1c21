set -x
mkdir -p gpurun_out
python scripts/cpu_scaling.py > gpurun_out/cpu_scaling.json 2>&1; cat gpurun_out/cpu_scaling.json
(python scripts/prof_cfg.py 3 2 --time
 DIAL_WPC=6 python scripts/prof_cfg.py 3 2 --time
 DIAL_WPC=4 python scripts/prof_cfg.py 3 2 --time
 DIAL_DENSE_LOCKSTEP=2 python scripts/prof_cfg.py 3 2 --time
 DIAL_DENSE_LOCKSTEP=2 DIAL_WPC=6 python scripts/prof_cfg.py 3 2 --time
 DIAL_NO_LOCKSTEP=1 python scripts/prof_cfg.py 3 2 --time
 DIAL_NO_LOCKSTEP=1 DIAL_WPC=6 python scripts/prof_cfg.py 3 2 --time
 DIAL_NO_LOCKSTEP=1 DIAL_NO_DYNAMIC_ROWS=1 python scripts/prof_cfg.py 3 2 --time) > gpurun_out/time10.log 2>&1
grep cfg gpurun_out/time10.log
DIAL_DEBUG_COUNTERS=1 python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch, ctypes as C
from baseline_configs import BASELINE, dial_config, product_env
from dial_mpc_b200 import random as drandom
from dial_mpc_b200.core.dial_core import MBDPI
env = product_env("allegro_reorient"); cfg = dial_config(3); mb = MBDPI(cfg, env)
st = env.reset(drandom.PRNGKey(0))
for _ in range(10): st = env.step(st, torch.zeros(mb.nu, device=mb.device))
Y = torch.zeros(cfg.Hnode + 1, mb.nu, device=mb.device)
out = (C.c_float * 8)()
mb.plan.lib.dial_debug_counters(mb.plan.handle, out)
mb.plan.reverse_rollout(st, None, drandom.PRNGKey(1), Y, mb.sigma_control, mb._rews_local); torch.cuda.synchronize()
mb.plan.lib.dial_debug_counters(mb.plan.handle, out)
print("ALLEGRO substeps", out[0], "newton iterations", out[1], "avg", out[1] / max(out[0], 1))
PY
python -m pytest tests/test_gpu_at_size.py -q --tb=short -p no:cacheprovider -k "baseline_size and 3" > gpurun_out/tests10.log 2>&1; tail -4 gpurun_out/tests10.log
