"""Scratch GPU check: parity of rollouts vs the oracle on all envs + first timings."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dial_mpc_b200.envs as E
from dial_mpc_b200 import random as drandom
from dial_mpc_b200.core.dial_config import DialConfig
from dial_mpc_b200.core.dial_core import MBDPI
from oracle.envs_oracle import make_env

def mk(name):
    if name == 'unitree_go2_walk':
        cfg = dict(default_vx=0.8, ramp_up_time=1.0); ecfg = E.UnitreeGo2EnvConfig(**cfg)
    elif name == 'unitree_go2_seq_jump':
        cfg = dict(pose_target_sequence=[[0,0,0.27],[0.4,0,0.27],[0.8,0,0.27],[1.2,0,0.27],[1.6,0,0.27]], yaw_target_sequence=[0.0]*5)
        ecfg = E.UnitreeGo2SeqJumpEnvConfig(**{k: np.array(v) for k, v in cfg.items()})
    else:
        cfg = dict(default_vx=2.0, ramp_up_time=3.0); ecfg = E.UnitreeH1WalkEnvConfig(**cfg)
    return E.get_environment(name, config=ecfg), make_env(name, cfg)

for name, H in [('unitree_go2_walk', 17), ('unitree_go2_seq_jump', 26), ('unitree_h1_walk', 31)]:
    env, o = mk(name)
    s = o.reset()
    st = env.reset(drandom.PRNGKey(0))
    print(name, "reset qpos err", np.abs(st.pipeline_state.qpos.cpu().numpy() - s.qpos[0]).max(),
          "warm err", np.abs(st.pipeline_state.qacc_warmstart.cpu().numpy() - s.qacc_warmstart[0]).max())
    rng = np.random.default_rng(1)
    B = 32
    us = np.clip(rng.normal(size=(B, H, env.action_size)) * 0.6, -1, 1)
    rew, q, qd, x = o.rollout(s, us)
    plan = env._get_plan()
    rg, qg, qdg, xg = plan.rollout(st, us)
    torch.cuda.synchronize()
    print("  q err", np.abs(qg.cpu().numpy() - q).max(), "qd err", np.abs(qdg.cpu().numpy() - qd).max(),
          "x err", np.abs(xg.cpu().numpy() - x).max(), "rew err", np.abs(rg.cpu().numpy() - rew).max())

# timing cfg2
for (name, N, Hs, Hn) in [('unitree_go2_seq_jump', 2048, 25, 5), ('unitree_go2_walk', 128, 16, 4), ('unitree_h1_walk', 2048, 30, 5), ('unitree_go2_walk', 8192, 25, 4)]:
    env, o = mk(name)
    cfg = DialConfig(env_name=name, Nsample=N, Hsample=Hs, Hnode=Hn, temp_sample=0.05)
    mb = MBDPI(cfg, env)
    st = env.reset(drandom.PRNGKey(0))
    Y = torch.zeros(Hn + 1, mb.nu, device=mb.device)
    rng = drandom.PRNGKey(0)
    for i in range(3):
        rng, Y2, info = mb.reverse_once(st, rng, Y, mb.sigma_control)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10
    e0.record()
    for i in range(K):
        rng, Y2, info = mb.reverse_once(st, rng, Y, mb.sigma_control)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"{name} N={N} Hs={Hs}: {ms:.3f} ms per reverse_once -> {N*Hs/ms*1e3:.3e} sample-steps/s; rews mean {float(info['rews'].mean()):.4f} finite={bool(torch.isfinite(info['rews']).all())}")
