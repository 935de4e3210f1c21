"""Does the ball stay in the Allegro hand?  CPU experiment on the fp64 oracle: the reset state with
the action that holds the reset pose (and with small random wiggles around it), 100 env steps, under

  A  the restated MJX algorithm (what the kernels implement)
  B  A with a line-search bracket that only ever narrows (lo / hi move to `next` only if that point lies on
     their own side of the minimum)
  C  A with the implicit-damping Euler step fed with forces consistent with the solver's iterate
     (M qacc instead of qfrc_smooth + qfrc_constraint: equal at convergence, different when the solver stalls)
  D  B + C

    python scripts/ball_retention_variants.py
Prints, per variant and action sequence: substeps on which the solver returned far from convergence
(|grad| / scale > 1; tolerance 1e-8), the largest ball speed, and where the ball is after 2 s."""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle.mjx_oracle as mo  # noqa: E402
from baseline_configs import ENV_CFG  # noqa: E402
from oracle.envs_oracle import make_env  # noqa: E402


def variant(narrowing: bool, iterate_forces: bool):
    """Copy of the oracle module with (narrowing) a line search that replaces lo / hi by their Newton
    successors only when the successor is still on the same side of the minimum (derivative of the same sign,
    closer to zero), and / or (iterate_forces) an Euler step whose implicit-damping solve sees M qacc."""
    src = open(mo.__file__).read()
    a = "LS_NARROWING = False"
    c = "(d.qfrc_smooth + d.qfrc_constraint)[..., None]"
    assert a in src and c in src
    if narrowing:
        src = src.replace(a, "LS_NARROWING = True")
    if iterate_forces:
        src = src.replace(c, 'np.einsum("nvw,nw->nv", d.M, d.qacc)[..., None]')
    mod = types.ModuleType("mjx_oracle_variant")
    mod.__dict__["__file__"] = mo.__file__
    exec(compile(src, mo.__file__, "exec"), mod.__dict__)
    return mod


def run(mod, o, seqs, eulerdamp):
    """All action sequences as one batch: [(unconverged substeps, max ball speed, final ball position)]."""
    m = o.m
    m.eulerdamp = eulerdamp
    B = len(seqs)
    s = o.reset().tile(B)
    scale = m.meaninertia * max(1, m.nv)
    qpos, qvel, warm = s.qpos.copy(), s.qvel.copy(), s.qacc_warmstart.copy()
    stalls, vmax = np.zeros(B, dtype=int), np.zeros(B)
    for t in range(seqs[0].shape[0]):
        ctrl = o.act2joint(np.stack([q[t] for q in seqs]))
        for _ in range(o.n_frames):
            qpos, qvel, warm, d = mod.step(m, qpos, qvel, ctrl, warm)
            g = np.linalg.norm(np.einsum("nvw,nw->nv", d.M, d.qacc) - d.qfrc_smooth - d.qfrc_constraint, axis=-1) / scale
            stalls += g > 1.0
            vmax = np.maximum(vmax, np.linalg.norm(qvel[:, :3], axis=-1))
    return [(int(stalls[i]), float(vmax[i]), qpos[i, :3].copy()) for i in range(B)]


def main():
    o = make_env("allegro_reorient", ENV_CFG["allegro_reorient"])
    jr = np.asarray(o.joint_range)
    hold = 2 * (-jr[:, 0] / (jr[:, 1] - jr[:, 0])) - 1
    rng = np.random.default_rng(0)
    seqs = {"hold": np.repeat(hold[None], 100, 0),
            "wiggle 0.05": np.clip(hold[None] + 0.05 * rng.normal(size=(100, 16)), -1, 1),
            "wiggle 0.15": np.clip(hold[None] + 0.15 * rng.normal(size=(100, 16)), -1, 1)}
    ed0 = bool(o.m.eulerdamp)
    print("| variant | actions | unconverged substeps (of 400) | max ball speed m/s | ball after 2 s (x, y, z) | in hand |\n|---|---|---|---|---|---|")
    only = sys.argv[1:] or ["A", "B", "C", "D"]
    for name, mod, ed in (("A restated MJX", mo, ed0), ("B narrowing bracket", variant(True, False), ed0),
                          ("C forces of the iterate", variant(False, True), ed0), ("D narrowing + forces of the iterate", variant(True, True), ed0)):
        if name[0] not in only:
            continue
        for (sname, _), (st, vmax, p) in zip(seqs.items(), run(mod, o, list(seqs.values()), ed)):
            kept = abs(p[0]) < 0.08 and abs(p[1]) < 0.08 and p[2] > 0.05
            print(f"| {name} | {sname} | {st} | {vmax:.2f} | {p[0]:.3f}, {p[1]:.3f}, {p[2]:.3f} | {'yes' if kept else 'no'} |", flush=True)
    o.m.eulerdamp = ed0


if __name__ == "__main__":
    main()
