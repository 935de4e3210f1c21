"""TEST-ONLY: oracle rollouts spread over host processes (spawned, so they never inherit a CUDA
context).  The oracle is batched NumPy; the Allegro model (100 Newton x 50 line-search
iterations, 4 substeps) needs minutes per hundred rows on one core."""
from __future__ import annotations

import os
from concurrent.futures import ProcessPoolExecutor
from multiprocessing import get_context

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rollout_chunk(args):
    import sys
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    name, cfg, qpos, qvel, warm, step, stage, us = args
    from oracle.envs_oracle import OState, make_env
    o = make_env(name, cfg)
    s = OState(qpos[None], qvel[None], warm[None], np.array([step], dtype=np.int64), np.array([stage], dtype=np.int64)).tile(us.shape[0])
    rews, qs, qds, xs, ws, sts = [], [], [], [], [], []
    for t in range(us.shape[1]):          # OracleEnv.rollout, also keeping qacc_warmstart and the stage
        s, r, aux = o.step(s, us[:, t])
        rews.append(r); qs.append(aux["q"]); qds.append(aux["qd"]); xs.append(aux["xpos"])
        ws.append(s.qacc_warmstart.copy()); sts.append(np.asarray(s.stage).copy())
    return (np.stack(rews, 1), np.stack(qs, 1), np.stack(qds, 1), np.stack(xs, 1), np.stack(ws, 1), np.stack(sts, 1))


def oracle_rollout(name, cfg, qpos, qvel, warm, step, stage, us, procs=None):
    """o.rollout(state, us) with the rows of `us` split over `procs` processes.  Returns
    (rewss [B,H], q, qd, xpos) like OracleEnv.rollout, then qacc_warmstart [B,H,nv] and the stage
    [B,H] after every step (what is needed to restart a row from the oracle's state at any step)."""
    us = np.asarray(us, dtype=np.float64)
    B = us.shape[0]
    procs = max(1, min(procs or (os.cpu_count() or 1), B, 32))
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    base = (name, dict(cfg), f64(qpos), f64(qvel), f64(warm), int(step), int(stage))
    if procs == 1:
        return _rollout_chunk(base + (us,))
    chunks = np.array_split(np.arange(B), procs)
    old = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
    for k in old:
        os.environ[k] = "1"       # inherited by the spawned workers before they import numpy
    try:
        with ProcessPoolExecutor(procs, mp_context=get_context("spawn")) as ex:
            parts = list(ex.map(_rollout_chunk, [base + (us[c],) for c in chunks]))
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return tuple(np.concatenate([p[i] for p in parts], 0) for i in range(6))


def oracle_with_yardstick(name, cfg, qpos, qvel, warm, step, stage, us, rng, K=3, scale=1e-5, procs=None):
    """Nominal oracle rollout of `us` plus the oracle's own sensitivity to fp32-sized noise, per
    output element: the max over K re-runs with the actions perturbed by `scale`*N(0,1) and — for
    the tree models, which the C port covers — the fp32 build of the C port against the fp64
    oracle (the oracle's own fp32-vs-fp64 divergence).  Returns (nominal 6-tuple incl. warm-start
    and stage trajectories, sens 4-tuple)."""
    us = np.asarray(us, dtype=np.float64)
    n = us.shape[0]
    stack = np.concatenate([us] + [us + scale * rng.standard_normal(us.shape) for _ in range(K)], 0)
    out = oracle_rollout(name, cfg, qpos, qvel, warm, step, stage, stack, procs=procs)
    nom = tuple(a[:n] for a in out)
    sens = [np.zeros_like(a) for a in nom[:4]]
    for k in range(K):
        for i in range(4):
            sens[i] = np.maximum(sens[i], np.abs(out[i][(k + 1) * n:(k + 2) * n] - nom[i]))
    try:
        from oracle import build_oracle
        from oracle.c_port import CPort
        from oracle.envs_oracle import OState, make_env
        build_oracle.build()
        o = make_env(name, cfg)
        s = OState(np.asarray(qpos, dtype=np.float64)[None], np.asarray(qvel, dtype=np.float64)[None],
                   np.asarray(warm, dtype=np.float64)[None], np.array([int(step)]), np.array([int(stage)]))
        f32 = CPort(o, real="float").rollout(s, us)[:4]
        for i in range(4):
            sens[i] = np.maximum(sens[i], np.abs(f32[i] - nom[i]))
    except NotImplementedError:
        pass        # dense / elliptic model: perturbation yardstick only
    return nom, tuple(sens)
