"""GPU parity tests (run on the B200 box): the CUDA path through the C ABI against the fp64
oracle, the committed golden fixtures, and size-independent properties at BASELINE sizes.

Tolerances (fp32 kernel vs fp64 oracle): q / x.pos 2e-4 abs, qvel 1e-2 abs,
per-step rewards 2e-3*(1+|r|), per-sample mean rewards 1e-3*(1+|r|), softmax weights total
variation 2e-2, Ybar 1e-2 (weights are exp((r-rbar)/std/0.05): reward noise is amplified)."""
import os

import numpy as np
import pytest
import torch

from tests.conftest import ENV_CASES, make_pair

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _state_from(env, qpos, qvel, warm, step=0, stage=0):
    from dial_mpc_b200.envs.base_env import PipelineState, State
    plan = env._get_plan()
    return State(PipelineState(plan.f32(qpos), plan.f32(qvel), plan.f32(warm)), None, 0.0, 0.0, {},
                 {"step": int(step), "contact_stage": int(stage)})


@pytest.mark.parametrize("name,H", [("unitree_go2_walk", 17), ("unitree_go2_seq_jump", 26), ("unitree_h1_walk", 31),
                                    ("allegro_reorient", 6), ("unitree_h1_loco", 21)])
def test_rollout_matches_oracle(built, name, H):
    from dial_mpc_b200 import random as drandom
    env, o = make_pair(name)
    s = o.reset()
    st = env.reset(drandom.PRNGKey(0))
    assert np.abs(st.pipeline_state.qpos.cpu().numpy() - s.qpos[0]).max() < 1e-6
    werr = np.abs(st.pipeline_state.qacc_warmstart.cpu().numpy() - s.qacc_warmstart[0]).max()
    assert werr < 2e-3 * (1 + np.abs(s.qacc_warmstart[0]).max())
    rng = np.random.default_rng(1)
    B = 24 if name != "allegro_reorient" else 6
    us = np.clip(rng.normal(size=(B, H, env.action_size)) * (0.6 if name != "allegro_reorient" else 0.3), -1, 1)
    rew, q, qd, x = o.rollout(s, us)
    rg, qg, qdg, xg = env._get_plan().rollout(st, us)
    torch.cuda.synchronize()
    if name == "allegro_reorient":
        # a 10 g ball between finger tips: single contact events change velocities by O(1) and
        # amplify fp32 rounding.  Deterministic statement: every row within tolerance + YARD x the
        # ORACLE's own sensitivity to fp32-sized noise on the actions (per-substep parity from
        # identical states is in test_gpu_at_size.py::test_single_physics_step_qacc_parity)
        from tests.test_gpu_at_size import YARD
        sq_, sr_ = np.zeros_like(q), np.zeros_like(rew)
        for _ in range(3):      # yardstick: max over three re-runs with 1e-5 noise on the actions
            rewp, qp, _, _ = o.rollout(s, us + 1e-5 * rng.standard_normal(us.shape))
            sq_, sr_ = np.maximum(sq_, np.abs(qp - q)), np.maximum(sr_, np.abs(rewp - rew))
        eq = np.abs(qg.cpu().numpy() - q)
        er = np.abs(rg.cpu().numpy() - rew)
        assert (eq <= 2e-4 + YARD * sq_).all(), (eq.max(), sq_.max())
        assert (er <= 2e-3 * (1 + np.abs(rew)) + YARD * sr_).all(), (er.max(), sr_.max())
        assert eq[:, :3].max() < 2e-4 and (er[:, :3] / (1 + np.abs(rew[:, :3]))).max() < 2e-3   # before the first contact event
        assert np.isfinite(rg.cpu().numpy()).all()
        return
    assert np.abs(qg.cpu().numpy() - q).max() < 2e-4
    assert np.abs(qdg.cpu().numpy() - qd).max() < 1e-2
    assert np.abs(xg.cpu().numpy() - x).max() < 2e-4
    assert (np.abs(rg.cpu().numpy() - rew) < 2e-3 * (1 + np.abs(rew))).all()


def test_env_step_sequence_matches_oracle(built):
    from dial_mpc_b200 import random as drandom
    env, o = make_pair("unitree_go2_seq_jump")
    s = o.reset()
    st = env.reset(drandom.PRNGKey(0))
    rng = np.random.default_rng(2)
    for t in range(55):       # crosses the stage boundary at step 50
        a = np.clip(rng.normal(size=12) * 0.4, -1, 1)
        s, r, aux = o.step(s, a[None])
        st = env.step(st, a)
        assert st.info["step"] == t + 1 and st.info["contact_stage"] == int(s.stage[0])
        if t < 12:            # contact dynamics are chaotic: compare while trajectories are close
            assert abs(float(st.reward) - r[0]) < 2e-3 * (1 + abs(r[0]))
            assert np.abs(st.pipeline_state.qpos.cpu().numpy() - s.qpos[0]).max() < 2e-4
    assert np.isfinite(float(st.reward))


@pytest.mark.parametrize("name", list(ENV_CASES))
def test_reverse_once_matches_golden(built, name):
    from dial_mpc_b200 import random as drandom
    from dial_mpc_b200.core.dial_config import DialConfig
    from dial_mpc_b200.core.dial_core import MBDPI
    g = np.load(os.path.join(GOLD, f"{name}.npz"))
    env, _ = make_pair(name)
    cfg = DialConfig(env_name=name, Nsample=int(g["N"]), Hsample=int(g["Hs"]), Hnode=int(g["Hn"]),
                     temp_sample=float(g["temp"]), horizon_diffuse_factor=0.9 if "go2" in name else 1.0)
    mb = MBDPI(cfg, env)
    assert np.abs(mb.sigma_control.cpu().numpy() - g["noise_scale"]).max() < 1e-6
    st = _state_from(env, g["qpos"], g["qvel"], g["qacc_warmstart"], g["step"], g["stage"])
    _, Ybar, info = mb.reverse_once(st, drandom.PRNGKey(0), g["Ybar0"], g["noise_scale"], eps=g["eps"])
    torch.cuda.synchronize()
    rews = info["rews"].cpu().numpy()
    w = info["weights"].cpu().numpy()
    assert abs(w.sum() - 1) < 1e-4
    if name == "allegro_reorient":   # chaotic contact events: budget = tolerance + YARD x oracle sensitivity (fixture)
        from tests.test_gpu_at_size import YARD
        err = np.abs(rews - g["rews"])
        assert (err <= 2e-3 * (1 + np.abs(g["rews"])) + YARD * g["rews_sens"]).all() and np.isfinite(rews).all(), (err, g["rews_sens"])
        assert int(np.argmax(w)) == int(np.argmax(g["weights"]))
        return
    assert (np.abs(rews - g["rews"]) < 1e-3 * (1 + np.abs(g["rews"]))).all()
    assert 0.5 * np.abs(w - g["weights"]).sum() < 2e-2
    assert np.abs(Ybar.cpu().numpy() - g["Ybar"]).max() < 1e-2
    assert np.abs(info["qbar"].cpu().numpy() - g["qbar"]).max() < 5e-3
    assert np.abs(info["xbar"].cpu().numpy() - g["xbar"]).max() < 5e-3
    # explicit-actions rollout of the same rows reproduces the per-step rewards
    rewss, (q, qd, x) = mb.rollout_us_vmap(st, g["us"])
    assert (np.abs(rewss.cpu().numpy() - g["rewss"]) < 2e-3 * (1 + np.abs(g["rewss"]))).all()


def test_update_stage_exact_on_given_rewards(built):
    """weights/Ybar kernels against a float64 evaluation on the SAME rewards (no chaos)."""
    from dial_mpc_b200.core.dial_config import DialConfig
    from dial_mpc_b200.core.dial_core import MBDPI
    env, _ = make_pair("unitree_go2_walk")
    N, Hn = 1000, 4
    mb = MBDPI(DialConfig(env_name="unitree_go2_walk", Nsample=N, Hsample=16, Hnode=Hn, temp_sample=0.05), env)
    rng = np.random.default_rng(0)
    rews = rng.normal(size=N + 1).astype(np.float32) * 0.3 - 1.0
    eps = rng.standard_normal((N, Hn + 1, 12)).astype(np.float32)
    Ybar = (rng.standard_normal((Hn + 1, 12)) * 0.4).astype(np.float32)
    noise = mb.sigma_control
    out, w = torch.empty(Hn + 1, 12, device="cuda"), torch.empty(N + 1, device="cuda")
    mb.plan.reverse_update(mb.plan.f32(eps), None, mb.plan.f32(Ybar), noise, mb.plan.f32(rews), out, w)
    r = rews.astype(np.float64)
    logp = (r - r[-1]) / r.std() / 0.05
    wr = np.exp(logp - logp.max()); wr /= wr.sum()
    Y0s = eps.astype(np.float64) * noise.cpu().numpy().astype(np.float64)[None, :, None] + Ybar
    Y0s[:, 0] = Ybar[0]
    Y0s = np.clip(np.concatenate([Y0s, Ybar[None].astype(np.float64)], 0), -1, 1)
    assert np.abs(w.cpu().numpy() - wr).max() < 2e-5
    assert np.abs(out.cpu().numpy() - np.einsum("n,nij->ij", wr, Y0s)).max() < 2e-5
    # diverged samples (NaN / inf reward) get weight 0 instead of poisoning the update
    rews2 = rews.copy(); rews2[5] = np.nan; rews2[7] = np.inf
    mb.plan.reverse_update(mb.plan.f32(eps), None, mb.plan.f32(Ybar), noise, mb.plan.f32(rews2), out, w)
    wn = w.cpu().numpy()
    assert wn[5] == 0 and wn[7] == 0 and np.isfinite(wn).all() and abs(wn.sum() - 1) < 1e-4
    # a diverged MEAN sample (rbar non-finite) only moves the reference point of the softmax shift
    rews3 = rews.copy(); rews3[-1] = np.nan
    mb.plan.reverse_update(mb.plan.f32(eps), None, mb.plan.f32(Ybar), noise, mb.plan.f32(rews3), out, w)
    w3 = w.cpu().numpy()
    r3 = r[:-1]
    l3 = (r3 - r3.max()) / r3.std() / 0.05
    e3 = np.exp(l3); e3 /= e3.sum()
    assert w3[-1] == 0 and np.abs(w3[:-1] - e3).max() < 5e-5 and torch.isfinite(out).all()
    # no finite reward at all: the whole weight goes to the mean sample, Ybar is kept (clipped)
    mb.plan.reverse_update(mb.plan.f32(eps), None, mb.plan.f32(Ybar), noise, mb.plan.f32(np.full(N + 1, np.nan)), out, w)
    assert float(w[-1]) == 1.0 and float(w[:-1].abs().max()) == 0.0
    assert np.abs(out.cpu().numpy() - np.clip(Ybar, -1, 1)).max() < 1e-6


def test_native_rng_matches_oracle_restatement(built):
    """In-kernel Threefry + erfinv sampler == the oracle's restatement of jax.random.normal
    (legacy counter layout): same rewards whether eps is injected or generated."""
    from dial_mpc_b200 import random as drandom
    from dial_mpc_b200.core.dial_config import DialConfig
    from dial_mpc_b200.core.dial_core import MBDPI
    from oracle.planner_oracle import jax_normal_legacy, jax_split_legacy
    env, _ = make_pair("unitree_go2_walk")
    N, Hs, Hn = 256, 16, 4
    mb = MBDPI(DialConfig(env_name="unitree_go2_walk", Nsample=N, Hsample=Hs, Hnode=Hn, temp_sample=0.05), env)
    st = env.reset(drandom.PRNGKey(0))
    rng = drandom.PRNGKey(3)
    Y0 = torch.zeros(Hn + 1, 12, device="cuda")
    rng1, Y1, i1 = mb.reverse_once(st, rng, Y0, mb.sigma_control)
    sub = jax_split_legacy((int(rng[0]), int(rng[1])))
    assert tuple(rng1) == tuple(sub[0])
    eps = jax_normal_legacy(tuple(int(v) for v in sub[1]), (N, Hn + 1, 12)).astype(np.float32)
    _, Y2, i2 = mb.reverse_once(st, rng, Y0, mb.sigma_control, eps=eps)
    # the two eps tensors differ by <= 1 ulp (device erfinv vs fp64 erfinv rounded to fp32); contact
    # switches amplify that for a few rows: 99 % of rows within 2e-5, outliers bounded by 2e-3
    diff = np.abs(i1["rews"].cpu().numpy() - i2["rews"].cpu().numpy())
    assert np.quantile(diff, 0.99) < 2e-5 and diff.max() < 2e-3, (np.quantile(diff, 0.99), diff.max())
    assert np.abs(Y1.cpu().numpy() - Y2.cpu().numpy()).max() < 1e-2


def test_full_size_properties(built):
    """BASELINE configs[1] size (Go2 seq-jump, N=2048, Hs=25): properties that do not need the
    oracle — determinism, shard invariance, simplex weights, bounded update, finite outputs."""
    from dial_mpc_b200 import random as drandom
    from dial_mpc_b200.core.dial_config import DialConfig
    from dial_mpc_b200.core.dial_core import MBDPI
    env, _ = make_pair("unitree_go2_seq_jump")
    N, Hs, Hn = 2048, 25, 5
    cfg = DialConfig(env_name="unitree_go2_seq_jump", Nsample=N, Hsample=Hs, Hnode=Hn, temp_sample=0.05)
    mb = MBDPI(cfg, env)
    st = env.reset(drandom.PRNGKey(0))
    rng = drandom.PRNGKey(0)
    Y0 = torch.zeros(Hn + 1, 12, device="cuda")
    _, Ya, ia = mb.reverse_once(st, rng, Y0, mb.sigma_control)
    _, Yb, ib = mb.reverse_once(st, rng, Y0, mb.sigma_control)
    assert torch.equal(ia["rews"], ib["rews"]) and torch.equal(Ya, Yb)          # bitwise deterministic
    w = ia["weights"]
    assert torch.isfinite(ia["rews"]).all() and abs(float(w.sum()) - 1) < 1e-4 and float(w.min()) >= 0
    assert float(Ya.abs().max()) <= 1.0 + 1e-6                                  # convex combination of clipped knots
    assert torch.allclose(Ya[0], Y0[0].clamp(-1, 1), atol=1e-6)                # node 0 is pinned
    assert ia["qbar"].shape == (Hs + 1, 19) and ia["xbar"].shape == (Hs + 1, 13, 3)
    # a rank owning only samples [1024, 2048) computes the same per-sample rewards: bitwise when both
    # launches use the same kernel instantiation (warps per CTA), to fp32 rounding otherwise (the
    # compiler fuses multiply-adds differently per instantiation)
    half = MBDPI(cfg, env, rank=1, world_size=2)
    half.plan.reverse_rollout(st, None, drandom.split(rng)[1], Y0, mb.sigma_control, half._rews_local)
    torch.cuda.synchronize()
    assert torch.allclose(half._rews_local[:1024], ia["rews"][1024:2048], rtol=0, atol=2e-4)
    os.environ["DIAL_WPC"] = "8"
    try:
        full8 = torch.empty_like(mb._rews_local)
        mb.plan.reverse_rollout(st, None, drandom.split(rng)[1], Y0, mb.sigma_control, full8)
        half.plan.reverse_rollout(st, None, drandom.split(rng)[1], Y0, mb.sigma_control, half._rews_local)
        torch.cuda.synchronize()
    finally:
        del os.environ["DIAL_WPC"]
    assert torch.equal(half._rews_local[:1024], full8[1024:2048])
    assert torch.equal(half._rews_local[1024], full8[2048])
    # mean of zero-noise rows equals the mean row
    _, Yz, iz = mb.reverse_once(st, rng, Y0, torch.zeros(Hn + 1, device="cuda"))
    assert float((iz["rews"] - iz["rews"][-1]).abs().max()) == 0.0
    # ... and std(rews) == 0 must not poison the update (the reference would return NaN): uniform
    # weights, Ybar = the clipped input knots
    assert torch.isfinite(Yz).all() and torch.allclose(Yz, Y0.clamp(-1, 1), atol=1e-6)
    assert torch.allclose(iz["weights"], torch.full_like(iz["weights"], 1.0 / (N + 1)), rtol=1e-4)
    assert torch.isfinite(iz["xbar"]).all()


def test_error_paths(built):
    from dial_mpc_b200.core.dial_config import DialConfig
    from dial_mpc_b200.core.dial_core import MBDPI
    env, _ = make_pair("unitree_go2_walk")
    with pytest.raises(ValueError):
        MBDPI(DialConfig(env_name="unitree_go2_walk", Nsample=8, Hsample=100, Hnode=4), env)   # Hsample too long
    with pytest.raises(KeyError):
        MBDPI(DialConfig(env_name="unitree_go2_walk", update_method="cma"), env)
    d = env.plan_desc()
    d.env_id = 99
    from dial_mpc_b200.plan import Plan
    with pytest.raises(RuntimeError, match="unknown env_id"):
        Plan(env, d)


def test_generic_tree_solver_fallback(built):
    """The level-scheduled compact Cholesky (used for trees that are not 'root chain + hanging
    chains') must agree with the star solve: run the golden case with the fallback forced."""
    import subprocess
    import sys
    code = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from tests.conftest import make_pair
from tests.test_gpu_parity import _state_from, GOLD
from dial_mpc_b200 import random as drandom
from dial_mpc_b200.core.dial_config import DialConfig
from dial_mpc_b200.core.dial_core import MBDPI
for name, hdf in (("unitree_go2_walk", 0.9), ("unitree_h1_walk", 1.0)):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    env, _ = make_pair(name)
    mb = MBDPI(DialConfig(env_name=name, Nsample=int(g["N"]), Hsample=int(g["Hs"]), Hnode=int(g["Hn"]),
                          temp_sample=float(g["temp"]), horizon_diffuse_factor=hdf), env)
    st = _state_from(env, g["qpos"], g["qvel"], g["qacc_warmstart"], g["step"], g["stage"])
    _, Y, info = mb.reverse_once(st, drandom.PRNGKey(0), g["Ybar0"], g["noise_scale"], eps=g["eps"])
    r = info["rews"].cpu().numpy()
    assert (np.abs(r - g["rews"]) < 1e-3 * (1 + np.abs(g["rews"]))).all(), np.abs(r - g["rews"]).max()
    assert np.abs(Y.cpu().numpy() - g["Ybar"]).max() < 1e-2
print("GENERIC_OK")
'''
    env = dict(os.environ, DIAL_FORCE_GENERIC_TREE="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True,
                         cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))), timeout=600)
    assert "GENERIC_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_deploy_planner_cycle_over_shared_memory(built):
    """the async planner (deploy/dial_plan.py) attached to the reference's six shm
    segments, two planning cycles 20.5 ms apart, against the ORACLE planner fed with the same
    Threefry noise (jax.random restatement): joint targets, torques and reference positions
    published to shm.  Reference: dial_mpc/deploy/dial_plan.py:136-139,172-229."""
    from scipy.interpolate import InterpolatedUnivariateSpline
    from dial_mpc_b200.core.dial_config import DialConfig
    from dial_mpc_b200.deploy.dial_plan import MBDPublisher
    from oracle.envs_oracle import OState, make_env
    from oracle.planner_oracle import PlannerOracle, jax_normal_legacy, jax_split_legacy
    import dial_mpc_b200.envs as E
    cfg = DialConfig(env_name="unitree_go2_walk", Nsample=256, Hsample=16, Hnode=4, Ndiffuse=1, Ndiffuse_init=2,
                     temp_sample=0.05, seed=0)
    ecfg = E.UnitreeGo2EnvConfig(default_vx=0.8, ramp_up_time=1.0)
    env = E.get_environment(cfg.env_name, config=ecfg)
    o = make_env(cfg.env_name, dict(default_vx=0.8, ramp_up_time=1.0))
    N, Hs, Hn, nu = cfg.Nsample, cfg.Hsample, cfg.Hnode, 12
    po = PlannerOracle(o, N, Hs, Hn, cfg.temp_sample, cfg.horizon_diffuse_factor, cfg.traj_diffuse_factor)

    def oracle_cycle(state, rng, Y, n_list):
        info = None
        for n in n_list:
            for i in range(n):
                sub = jax_split_legacy((int(rng[0]), int(rng[1])))
                rng, key = sub[0], sub[1]
                eps = jax_normal_legacy((int(key[0]), int(key[1])), (N, Hn + 1, nu))
                Y, info = po.reverse_once(state, eps, Y, cfg.traj_diffuse_factor ** i * np.ones(Hn + 1))   # no sigma_control
        return rng, Y, info

    def published(state, Y, info):
        us = po.node2u(Y)
        acts = np.stack([o.act2joint(u) for u in us])
        taus = np.stack([o.act2tau(u[None], state.qpos, state.qvel)[0] for u in us])
        return acts, taus, info["xbar"][:, 1:, :3]

    pub = MBDPublisher(env, ecfg, cfg, create_shm=True)      # the test plays the simulator's role
    try:
        assert pub.acts_shared.shape == (17, 12) and pub.refs_shared.shape == (17, 12, 3)
        assert pub._shm["state_shm"].size >= 37 * 32
        # ---- cycle 1 at t = 0 from the home keyframe: Ndiffuse_init then Ndiffuse iterations ----------
        pub.plan_once()
        q0 = np.asarray(pub.default_q, dtype=np.float64)
        s0 = OState(q0[None], np.zeros((1, 18)), np.zeros((1, 18)), np.array([0]), np.array([0]))   # warm-start 0: no mjx.forward
        rng, Yo, io = oracle_cycle(s0, (0, cfg.seed), np.zeros((Hn + 1, nu)), [cfg.Ndiffuse_init, cfg.Ndiffuse])
        acts, taus, refs = published(s0, Yo, io)
        assert pub.plan_time_shared[0] == 0.0
        assert np.abs(pub.Y.cpu().numpy() - Yo).max() < 1e-2
        assert np.abs(pub.acts_shared - acts).max() < 1e-2
        assert np.abs(pub.tau_shared - taus).max() < 0.5            # kp = 30: 1e-2 rad of target
        assert np.abs(pub.refs_shared - refs).max() < 5e-3
        assert np.array_equal(np.asarray(pub.rng, dtype=np.uint32), np.asarray(rng, dtype=np.uint32))
        lo, hi = env.physical_joint_range[:, 0], env.physical_joint_range[:, 1]
        assert (pub.acts_shared >= lo - 1e-5).all() and (pub.acts_shared <= hi + 1e-5).all()
        # ---- cycle 2: the simulator advanced 20.5 ms (int(t / dt) truncates like the reference) ----------
        s1, _, _ = o.step(s0, Yo[0][None])
        pub.state_shared[:19] = s1.qpos[0]
        pub.state_shared[19:37] = s1.qvel[0]
        pub.time_shared[0] = 0.0205
        Y_before = pub.Y.clone()
        pub.plan_once()
        assert pub.plan_time_shared[0] == np.float32(0.0205) and pub._state.info["step"] == 1
        shift_time = float(np.float32(0.0205)) - 0.0
        Ysh = np.stack([InterpolatedUnivariateSpline(po.step_nodes, Yo[:, a], k=2)(po.step_nodes + shift_time) for a in range(nu)], 1)
        s1p = OState(s1.qpos.astype(np.float32).astype(np.float64), s1.qvel.astype(np.float32).astype(np.float64),
                     np.zeros((1, 18)), np.array([1]), np.array([0]))
        rng2, Yo2, io2 = oracle_cycle(s1p, rng, Ysh, [cfg.Ndiffuse])
        acts2, taus2, refs2 = published(s1p, Yo2, io2)
        assert np.abs(pub.Y.cpu().numpy() - Yo2).max() < 2e-2
        assert np.abs(pub.acts_shared - acts2).max() < 2e-2
        assert np.abs(pub.tau_shared - taus2).max() < 1.0
        assert np.abs(pub.refs_shared - refs2).max() < 1e-2
        assert np.array_equal(np.asarray(pub.rng, dtype=np.uint32), np.asarray(rng2, dtype=np.uint32))
        # shifting by one node period moves node k+1 onto node k (interpolating spline)
        sh = pub.shift(Y_before, pub.mbdpi.node_dt)
        assert torch.allclose(sh[:-1], Y_before[1:], atol=1e-5)
    finally:
        pub.close(unlink=True)


@pytest.mark.parametrize("name", ["unitree_go2_seq_jump", "unitree_h1_walk"])
def test_device_loop_graph_equals_eager_loop(built, name):
    """The CUDA-graph control step (dial_mpc_step: env.step + shift + Ndiffuse x reverse_once with
    device-side key splitting / counters) reproduces the eager call sequence of the reference's
    main loop (core/dial_core.py:242-268): same keys (exact), same rewards and control knots."""
    from dial_mpc_b200 import random as drandom
    from dial_mpc_b200.core.dial_config import DialConfig
    from dial_mpc_b200.core.dial_core import DeviceLoop, MBDPI
    env, _ = make_pair(name)
    args = DialConfig(env_name=name, Nsample=256, Hsample=12, Hnode=4, Ndiffuse=3, Ndiffuse_init=5,
                      temp_sample=0.05, horizon_diffuse_factor=0.9, traj_diffuse_factor=0.5)
    mb = MBDPI(args, env)
    rng = drandom.PRNGKey(3)
    rng, r0 = drandom.split(rng)
    state = env.reset(r0)
    state.info["step"] = 48          # seq-jump: crosses into the second stage during the run
    if "contact_stage" in state.info:
        state.info["contact_stage"] = 0
    Y0 = torch.zeros(args.Hnode + 1, mb.nu, device="cuda")
    loop = DeviceLoop(mb, state, rng, Y0)
    nsteps = 4                        # step 0 eager (Ndiffuse_init), 1 eager (Ndiffuse), 2.. graph replays
    st, Y, r = state, Y0, rng
    for t in range(nsteps):
        nd = args.Ndiffuse_init if t == 0 else args.Ndiffuse
        st = env.step(st, Y[0])
        Y = mb.shift(Y)
        r, Y, info = mb.reverse_scan(st, r, Y, mb.schedule(nd))
        loop.step(nd)
        torch.cuda.synchronize()
        s2 = loop.state()
        # rng chain and counters are integer-exact
        assert np.array_equal(np.asarray(r, dtype=np.uint32), s2.info["rng"])
        assert s2.info["step"] == st.info["step"]
        if "contact_stage" in st.info:
            assert s2.info["contact_stage"] == st.info["contact_stage"]
        tol = 1e-5 if t == 0 else 2e-3   # later steps inherit fp32 reordering of the 6x6 shift matmul
        assert (s2.pipeline_state.qpos - st.pipeline_state.qpos).abs().max() < tol
        assert abs(float(s2.reward) - float(st.reward)) < tol * (1 + abs(float(st.reward)))
        assert (loop.Y - Y).abs().max() < 5 * tol
        assert (loop.info()["rews"] - info["rews"]).abs().max() < 5 * tol * (1 + float(info["rews"].abs().max()))
        assert (loop.info()["xbar"] - info["xbar"]).abs().max() < 5 * tol
    assert mb.plan.launches > 0
    # env_step=False: plan again from the same state (deploy mode) -> state and counters untouched
    q_before, c_before = loop.buf["qpos"].clone(), loop.buf["counters"].clone()
    loop.step(2, env_step=False)
    loop.step(2, env_step=False)
    loop.step(2, env_step=False)
    torch.cuda.synchronize()
    assert torch.equal(q_before, loop.buf["qpos"]) and torch.equal(c_before, loop.buf["counters"])
    assert torch.isfinite(loop.Y).all()


@pytest.mark.parametrize("name", ["unitree_go2_walk", "unitree_go2_seq_jump", "unitree_h1_walk", "unitree_h1_loco",
                                  "allegro_reorient"])
def test_env_step_obs_and_done_match_oracle(built, name):
    """State.obs / State.done of env.reset / env.step against the oracle's restatement of the
    envs' _get_obs (unitree_go2_env.py:263-286, :523-557, unitree_h1_env.py:323-346, :850-873;
    manipulation.py:57,86-97): same layout, the info targets of the state ENTERING the step."""
    from dial_mpc_b200 import random as drandom
    env, o = make_pair(name)
    s = o.reset()
    st = env.reset(drandom.PRNGKey(0))
    if name == "allegro_reorient":
        assert tuple(st.obs.shape) == (1,) and float(st.obs[0]) == 0.0
        st = env.step(st, np.zeros(env.action_size))
        assert tuple(st.obs.shape) == (1,) and float(st.done[0]) == 0.0 and st.info["step"] == 1
        return
    d0 = __import__("oracle.mjx_oracle", fromlist=["forward"]).forward(
        o.m, s.qpos, s.qvel, np.zeros((1, o.nu)), s.qacc_warmstart)
    ob0 = o.observe(s, s.qpos, s.qvel, d0, np.zeros((1, o.nu)))[0]
    assert st.obs.shape[0] == ob0.shape[0]
    assert np.abs(st.obs.cpu().numpy() - ob0).max() < 1e-5
    rng = np.random.default_rng(3)
    last = None
    for t in range(8):
        a = np.clip(rng.normal(size=env.action_size) * 0.4, -1, 1)
        s_prev = s
        s, r, aux = o.step(s, a[None])
        ob = o.observe(s_prev, aux["q"], aux["qd"], aux["data"], aux["ctrl"], last_ctrl=last)[0]
        dn = o.done(s_prev, aux["q"], aux["qd"], aux["data"])[0]
        last = aux["ctrl"]
        st = env.step(st, a)
        og = st.obs.cpu().numpy()
        assert og.shape == ob.shape
        tol = 2e-4 + 1e-2 * (np.abs(ob) > 5)          # velocities / torques: the qvel tolerance
        err = np.abs(og - ob)
        assert (err <= np.maximum(tol, 2e-3 * (1 + np.abs(ob)))).all(), (t, int(err.argmax()), err.max())
        assert float(st.done) == dn
        if "vel_tar" in st.info and name != "unitree_go2_seq_jump":
            vt, at = o.info_targets(s.step)
            assert np.abs(np.asarray(st.info["vel_tar"]) - vt[0]).max() < 1e-6
            assert np.abs(np.asarray(st.info["ang_vel_tar"]) - at[0]).max() < 1e-6


@pytest.mark.parametrize("N", [16384, 65536])
def test_fused_update_equals_unfused_at_multi_gpu_totals(built, N):
    """The control-step graph's fused update kernel (statistics + weights + weighted knot sum in
    one launch) at the reward counts an 8-GPU run sees (8 x 2048 and configs[4]'s 65536) against
    the separate weights / Ybar kernels: same rollouts (bitwise), same knots to rounding."""
    from dial_mpc_b200 import random as drandom
    from dial_mpc_b200.core.dial_config import DialConfig
    from dial_mpc_b200.core.dial_core import DeviceLoop, MBDPI
    env, _ = make_pair("unitree_go2_walk")
    args = DialConfig(env_name="unitree_go2_walk", Nsample=N, Hsample=4, Hnode=2, Ndiffuse=2, Ndiffuse_init=2,
                      temp_sample=0.05, horizon_diffuse_factor=0.9, traj_diffuse_factor=0.5)
    out = []
    for unfused in (False, True):
        if unfused:
            os.environ["DIAL_NO_FUSED_UPDATE"] = "1"
        try:
            mb = MBDPI(args, env)
            rng = drandom.PRNGKey(5)
            rng, r0 = drandom.split(rng)
            state = env.reset(r0)
            loop = DeviceLoop(mb, state, rng, torch.zeros(args.Hnode + 1, mb.nu, device="cuda"))
            for _ in range(3):           # eager, eager, graph replay
                loop.step(args.Ndiffuse)
            torch.cuda.synchronize()
            out.append((loop.Y.clone(), loop.info()["rews"].clone(), loop.state().info["rng"].copy()))
        finally:
            os.environ.pop("DIAL_NO_FUSED_UPDATE", None)
    (Ya, ra, ka), (Yb, rb, kb) = out
    assert np.array_equal(ka, kb)
    assert torch.isfinite(Ya).all() and ra.shape[0] == N + 1
    assert (Ya - Yb).abs().max() < (2e-4 if N <= 16384 else 1e-3)     # fp32 sums over N terms in two different orders
    # the last rollouts start from knots that differ by the above: rewards agree to that, amplified
    assert (ra - rb).abs().max() < 0.1 * (1 + float(rb.abs().max()))


def test_randomize_tasks_one_step_command(built):
    """randomize_tasks=True (unitree_go2_env.py:141-163): the env step at step % 500 == 0 uses a
    random command drawn from the env's key chain; rollouts whose horizon reaches that step see it
    too.  CUDA path (dial_plan_set_command) against the oracle with the same override, through
    env.step, Plan.rollout and the control-step graph."""
    import dial_mpc_b200.envs as E
    from dial_mpc_b200 import random as drandom
    from dial_mpc_b200.core.dial_config import DialConfig
    from dial_mpc_b200.core.dial_core import DeviceLoop, MBDPI
    from oracle.envs_oracle import make_env
    from oracle.planner_oracle import sample_command_oracle
    name = "unitree_go2_walk"
    cfg = dict(ENV_CASES[name])
    env = E.get_environment(name, config=E.get_config(name)(randomize_tasks=True, **cfg))
    o = make_env(name, cfg)
    st = env.reset(drandom.PRNGKey(9))
    assert st.info["randomize_target"]
    st.info["step"] = 497
    s = o.reset()
    s.step[:] = 497
    # the command step 500 will use: three splits ahead on the env's key chain
    r = st.info["rng"]
    for _ in range(3):
        r = drandom.split(r)[0]
    vel, ang = sample_command_oracle(drandom.split(r)[1])
    o.cmd_override = (500, vel, ang)
    assert abs(vel[0] - cfg.get("default_vx", 1.0)) > 1e-3
    # rollouts from step 497 cross step 500
    rng = np.random.default_rng(6)
    us = np.clip(rng.normal(size=(8, 7, env.action_size)) * 0.5, -1, 1)
    rew, q, _, _ = o.rollout(s, us)
    rg, qg, _, _ = env._get_plan().rollout(st, us)
    assert np.abs(qg.cpu().numpy() - q).max() < 2e-4
    assert (np.abs(rg.cpu().numpy() - rew) < 2e-3 * (1 + np.abs(rew))).all()
    o.cmd_override = None
    base = o.rollout(s, us)[0]
    assert np.abs(rew[:, 3] - base[:, 3]).min() > 1e-2            # the override matters at step 500 only
    o.cmd_override = (500, vel, ang)
    # env.step across the boundary: rewards and the stored info targets
    for t in range(5):
        a = us[0, t]
        s_prev = s
        s, r_o, aux = o.step(s, a[None])
        st = env.step(st, a)
        assert abs(float(st.reward) - r_o[0]) < 2e-3 * (1 + abs(r_o[0])), t
        vt, at = o.info_targets(s.step)
        assert np.abs(np.asarray(st.info["vel_tar"]) - vt[0]).max() < 1e-6, t
        assert np.abs(np.asarray(st.info["ang_vel_tar"]) - at[0]).max() < 1e-6, t
    # the control-step graph: same knots as the eager loop when the command changes between replays
    args = DialConfig(env_name=name, Nsample=128, Hsample=8, Hnode=3, Ndiffuse=2, Ndiffuse_init=2,
                      temp_sample=0.05, horizon_diffuse_factor=0.9, traj_diffuse_factor=0.5)
    mb = MBDPI(args, env)
    st0 = env.reset(drandom.PRNGKey(9))
    st0.info["step"] = 488
    prng = drandom.PRNGKey(2)
    Y0 = torch.zeros(args.Hnode + 1, mb.nu, device="cuda")
    loop = DeviceLoop(mb, st0, prng, Y0)
    stE, Y, pr = st0, Y0, prng
    for t in range(14):              # 488 .. 501: command absent, within the horizon, at the env step, gone
        stE = env.step(stE, Y[0])
        Y = mb.shift(Y)
        pr, Y, info = mb.reverse_scan(stE, pr, Y, mb.schedule(args.Ndiffuse))
        loop.step(args.Ndiffuse)
        torch.cuda.synchronize()
        assert (loop.Y - Y).abs().max() < 5e-3, t
        assert abs(float(loop.reward) - float(stE.reward)) < 2e-3 * (1 + abs(float(stE.reward))), t
        if t == 12:
            assert float(stE.reward) < -5.0          # the env step at 500 ran with the random command
        # re-synchronise (closed-loop differences are amplified by the softmax; each step is compared on its own)
        ps = stE.pipeline_state
        loop.set_state(ps.qpos, ps.qvel, ps.qacc_warmstart)
        loop.buf["Y"].copy_(Y)
    assert mb.plan._cmd is None and loop.state().info["step"] == 502


def test_seq_jump_randomize_tasks_draws_the_sequence_at_reset(built):
    """randomize_tasks=True of UnitreeGo2SeqJumpEnv (unitree_go2_env.py:383-394, 594-631): reset draws
    an 11-stage jump sequence; every launch from such a state runs with its tables
    (dial_plan_set_stages).  CUDA rollouts / env steps against the oracle env built from the oracle's
    own restatement of the sampler, across a stage change; a state with the configured sequence on
    the same plan switches the tables back."""
    import dial_mpc_b200.envs as E
    from dial_mpc_b200 import random as drandom
    from dial_mpc_b200.core.dial_config import DialConfig
    from dial_mpc_b200.core.dial_core import DeviceLoop, MBDPI
    from oracle.envs_oracle import make_env
    from oracle.planner_oracle import sample_jump_sequence_oracle
    name = "unitree_go2_seq_jump"
    cfg = {k: (np.array(v) if isinstance(v, list) else v) for k, v in ENV_CASES[name].items()}
    env = E.get_environment(name, config=E.get_config(name)(randomize_tasks=True, **cfg))
    key = drandom.PRNGKey(21)
    st = env.reset(key)
    assert st.info["randomize_target"] and st.info["contact_targets"].shape == (11, 4, 3)
    com_pos, com_yaw = sample_jump_sequence_oracle(drandom.split(key)[0])
    o = make_env(name, dict(ENV_CASES[name], pose_target_sequence=com_pos, yaw_target_sequence=com_yaw))
    assert np.abs(st.info["contact_targets"] - o.contact_targets).max() < 1e-6
    s = o.reset()
    s.step[:] = 46
    st.info["step"] = 46
    rng = np.random.default_rng(8)
    us = np.clip(rng.normal(size=(8, 9, env.action_size)) * 0.5, -1, 1)
    rew, q, _, _ = o.rollout(s, us)
    plan = env._get_plan()
    rg, qg, _, _ = plan.rollout(st, us)
    assert np.abs(qg.cpu().numpy() - q).max() < 2e-4
    assert (np.abs(rg.cpu().numpy() - rew) < 2e-3 * (1 + np.abs(rew))).all()
    # the configured 5-stage sequence gives other rewards once the stage has changed (step 50)
    env0 = E.get_environment(name, config=E.get_config(name)(**cfg))
    o0 = make_env(name, dict(ENV_CASES[name]))
    st0 = env0.reset(key)
    st0.info["step"] = 46
    r0, _, _, _ = o0.rollout(s, us)
    assert np.abs(r0[:, 5:] - rew[:, 5:]).max() > 1e-3
    # ... also on the SAME plan: a launch from a state that carries the configured tables
    # switches them back (and forth)
    st_cfg = type(st)(st.pipeline_state, st.obs, st.reward, st.done, st.metrics,
                      dict(st.info, **{k2: st0.info[k2] for k2 in ("contact_targets", "contact_target_radius",
                                                                "pose_target_sequence", "yaw_target_sequence")}))
    rb, _, _, _ = plan.rollout(st_cfg, us)
    assert (np.abs(rb.cpu().numpy() - r0) < 2e-3 * (1 + np.abs(r0))).all()
    rg2, _, _, _ = plan.rollout(st, us)
    assert torch.equal(rg2, rg)
    # env.step across the stage change: rewards, stage counter, observation
    for t in range(6):
        a = us[0, t]
        s, r_o, aux = o.step(s, a[None])
        st = env.step(st, a)
        assert abs(float(st.reward) - r_o[0]) < 2e-3 * (1 + abs(r_o[0])), t
        assert st.info["contact_stage"] == int(s.stage[0]), t
    assert st.info["contact_stage"] == 1 and st.info["contact_targets"].shape == (11, 4, 3)
    # the control-step graph binds the drawn sequence: same knots as the eager loop
    args = DialConfig(env_name=name, Nsample=128, Hsample=8, Hnode=3, Ndiffuse=2, Ndiffuse_init=2,
                      temp_sample=0.05, horizon_diffuse_factor=0.9, traj_diffuse_factor=0.5)
    mb = MBDPI(args, env)
    mb.plan.set_stages(env.stage_tables(st0.info))      # start from the "wrong" tables on purpose
    prng = drandom.PRNGKey(2)
    Y0 = torch.zeros(args.Hnode + 1, mb.nu, device="cuda")
    loop = DeviceLoop(mb, st, prng, Y0)
    stE = env.step(st, Y0[0])
    Y = mb.shift(Y0)
    pr, Y, info = mb.reverse_scan(stE, prng, Y, mb.schedule(args.Ndiffuse))
    loop.step(args.Ndiffuse)
    torch.cuda.synchronize()
    assert (loop.Y - Y).abs().max() < 5e-3
    assert abs(float(loop.reward) - float(stE.reward)) < 2e-3 * (1 + abs(float(stE.reward)))
