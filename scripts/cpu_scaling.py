"""Thread-scaling sweep of the C port (CPU baseline) on this host: one reverse_once of configs[1]."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from baseline_configs import BASELINE, ENV_CFG  # noqa: E402
from oracle import build_oracle  # noqa: E402
from oracle.c_port import CPort  # noqa: E402
from oracle.envs_oracle import make_env  # noqa: E402
from oracle.planner_oracle import PlannerOracle  # noqa: E402

build_oracle.build()
b = BASELINE[1]
env = make_env(b["env"], ENV_CFG[b["env"]])
s = env.reset()
for _ in range(10):
    s, _, _ = env.step(s, np.zeros((1, env.nu)))
pl = PlannerOracle(env, b["N"], b["Hs"], b["Hn"], b["temp"], b["hdf"], b["tdf"])
eps = np.random.default_rng(0).standard_normal((b["N"], b["Hn"] + 1, env.nu))
us = pl.node2u(pl.make_Y0s(eps, np.zeros((b["Hn"] + 1, env.nu)), pl.sigma_control))
cp = CPort(env)
out = {"nproc": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)), "loadavg": os.getloadavg()}
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    try:
        out[f] = open(f).read().strip()
    except Exception:
        pass
res = {}
for th in (1, 2, 4, 8, 16, 32, 64, 128):
    if th > 2 * (os.cpu_count() or 1):
        break
    best = 1e9
    for _ in range(3):
        t = time.perf_counter()
        cp.rollout_rews(s, us, threads=th)
        best = min(best, time.perf_counter() - t)
    res[th] = b["N"] * b["Hs"] / best
out["sample_steps_per_s_by_threads"] = res
print(json.dumps(out))
