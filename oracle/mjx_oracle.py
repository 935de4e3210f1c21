"""ORACLE (test infrastructure, not product code) — fp64 NumPy restatement of the
third-party physics the reference's hot path calls.

PARITY UNPINNED: the reference (LeCAR-Lab/dial-mpc @ 871c84f) ships no tests,
golden vectors or fixtures for this path, and its physics lives in un-vendored,
un-pinned third-party packages (`mujoco` + `mujoco.mjx`, `brax`; reference
setup.py:9-21) that are not installable in this environment.  This file restates
the published MuJoCo-MJX algorithm (``mjx.step`` = ``forward`` + semi-implicit
Euler; MJX 3.1/3.2 generation) for the model features the BASELINE configs use.
It is anchored on the reference's own call sites:

* ``self.pipeline_step(state.pipeline_state, ctrl)`` -> ``brax.mjx.pipeline.step``
  -> ``mjx.step``            dial_mpc/envs/unitree_go2_env.py:135,415
                             dial_mpc/envs/unitree_h1_env.py:192
* the Brax wrapper quantities ``x, xd, contact`` as mirrored in-repo at
                             dial_mpc/deploy/dial_plan.py:45-61
* ``pipeline_init`` = make_data + set qpos/qvel + ``mjx.forward``
                             dial_mpc/envs/unitree_go2_env.py:104

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import this module.  Everything is batched over a
leading sample axis ``B`` and computed in float64.
"""

from __future__ import annotations

import json
from typing import Dict, NamedTuple, Optional

import numpy as np

JNT_FREE, JNT_SLIDE, JNT_HINGE = 0, 2, 3
PAIR_PLANE_SPHERE, PAIR_PLANE_CAPSULE = 0, 1
PAIR_SPHERE_SPHERE, PAIR_SPHERE_CAPSULE, PAIR_CAPSULE_CAPSULE = 2, 3, 4
MINVAL = 1e-15
MINIMP, MAXIMP = 1e-4, 0.9999


# ---------------------------------------------------------------------------
# model container (reads the compiled JSON blob directly; no product imports)
# ---------------------------------------------------------------------------
class OModel:
    def __init__(self, path: str, timestep: Optional[float] = None):
        with open(path) as f:
            obj = json.load(f)
        for k, v in obj["scalars"].items():
            setattr(self, k, v)
        self.gravity = np.array(self.gravity, dtype=np.float64)
        for k, v in obj["arrays"].items():
            setattr(self, k, np.array(v["data"], dtype=np.dtype(v["dtype"])).reshape(v["shape"]))
        self.names = obj["names"]
        self.keyframes = obj["keyframes"]
        if timestep is not None:
            self.timestep = float(timestep)
        # derived index helpers
        nv = self.nv
        anc = np.zeros((nv, nv), dtype=bool)  # anc[i, j]: dof j is ancestor-or-self of dof i
        for i in range(nv):
            j = i
            while j >= 0:
                anc[i, j] = True
                j = int(self.dof_parentid[j])
        self.dof_anc = anc
        banc = np.zeros((self.nbody, nv), dtype=bool)  # dof d moves body b
        for b in range(self.nbody):
            bb = b
            while bb > 0:
                if self.body_jntadr[bb] >= 0:
                    d0 = int(self.body_dofadr[bb])
                    banc[b, d0:d0 + int(self.body_dofnum[bb])] = True
                bb = int(self.body_parentid[bb])
        self.body_dofmask = banc
        self.lim_jnt = np.nonzero(self.jnt_limited)[0]
        self.nlim = len(self.lim_jnt)
        # rows per contact: pyramidal 2*(condim-1), elliptic condim; cones = (row0, dim, pair) per contact
        self.con_pair = np.repeat(np.arange(self.npair), self.pair_ncon)
        dims = self.pair_condim[self.con_pair]
        assert np.all((dims == 3) | (dims == 6) | (dims == 1)), "oracle: condim 1/3/6 only"
        rows = dims if self.cone == 1 else np.where(dims == 1, 1, 2 * (dims - 1))
        self.con_row0 = self.nlim + np.concatenate([[0], np.cumsum(rows)[:-1]]).astype(int) if self.ncon else np.zeros(0, int)
        self.con_rows = rows.astype(int)
        self.nefc = self.nlim + int(rows.sum())
        self.elliptic = self.cone == 1


# ---------------------------------------------------------------------------
# batched quaternion / vector helpers
# ---------------------------------------------------------------------------
def qmul(a, b):
    aw, ax, ay, az = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bw, bx, by, bz = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw], axis=-1)


def qrot(q, v):
    """MJX math.rotate(vec, quat)."""
    s, u = q[..., :1], q[..., 1:]
    r = 2 * (np.sum(u * v, -1, keepdims=True) * u) + (s * s - np.sum(u * u, -1, keepdims=True)) * v
    return r + 2 * s * np.cross(u, v)


def qconj(q):
    return q * np.array([1.0, -1.0, -1.0, -1.0])


def qmat(q):
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    m = np.stack([
        w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z], axis=-1)
    return m.reshape(q.shape[:-1] + (3, 3))


def axis_angle_quat(axis, angle):
    s, c = np.sin(angle * 0.5), np.cos(angle * 0.5)
    return np.concatenate([c[..., None], axis * s[..., None]], axis=-1)


def normalize(x):
    n = np.linalg.norm(x, axis=-1, keepdims=True)
    return x / (n + 1e-6 * (n == 0.0))


def motion_cross(u, v):
    ang = np.cross(u[..., :3], v[..., :3])
    vel = np.cross(u[..., 3:], v[..., :3]) + np.cross(u[..., :3], v[..., 3:])
    return np.concatenate([ang, vel], -1)


def motion_cross_force(v, f):
    ang = np.cross(v[..., :3], f[..., :3]) + np.cross(v[..., 3:], f[..., 3:])
    vel = np.cross(v[..., :3], f[..., 3:])
    return np.concatenate([ang, vel], -1)


def inert_mul(ci, v):
    """cinert (10) times motion vector (6): [Ixx Iyy Izz Ixy Ixz Iyz | m*off | m]."""
    I = np.stack([
        ci[..., 0], ci[..., 3], ci[..., 4],
        ci[..., 3], ci[..., 1], ci[..., 5],
        ci[..., 4], ci[..., 5], ci[..., 2]], -1).reshape(ci.shape[:-1] + (3, 3))
    pos, mass = ci[..., 6:9], ci[..., 9:10]
    ang = np.einsum("...ij,...j->...i", I, v[..., :3]) + np.cross(pos, v[..., 3:])
    vel = mass * v[..., 3:] - np.cross(pos, v[..., :3])
    return np.concatenate([ang, vel], -1)


# ---------------------------------------------------------------------------
# forward pipeline pieces
# ---------------------------------------------------------------------------
class Data(NamedTuple):
    qpos: np.ndarray            # [B,nq] (quaternions normalised, as mjx.kinematics does)
    xpos: np.ndarray            # [B,nb,3]
    xquat: np.ndarray           # [B,nb,4]
    xmat: np.ndarray            # [B,nb,3,3]
    xipos: np.ndarray
    ximat: np.ndarray
    xanchor: np.ndarray         # [B,nb,3] (per body's joint)
    xaxis: np.ndarray
    root_com: np.ndarray        # [B,nb,3] subtree_com[body_rootid[b]]
    cinert: np.ndarray          # [B,nb,10]
    cdof: np.ndarray            # [B,nv,6]
    M: np.ndarray               # [B,nv,nv]
    cvel: np.ndarray            # [B,nb,6]
    cdof_dot: np.ndarray
    qfrc_bias: np.ndarray
    qfrc_passive: np.ndarray
    qfrc_actuator: np.ndarray
    qfrc_smooth: np.ndarray
    qacc_smooth: np.ndarray
    site_xpos: np.ndarray       # [B,nsite,3]
    con_dist: np.ndarray        # [B,ncon]
    con_pos: np.ndarray         # [B,ncon,3]
    con_frame: np.ndarray       # [B,ncon,3,3]
    efc_J: np.ndarray           # [B,nefc,nv]
    efc_D: np.ndarray
    efc_aref: np.ndarray
    efc_pos: np.ndarray
    qacc: np.ndarray
    solver_niter: np.ndarray
    qfrc_constraint: np.ndarray


def kinematics(m: OModel, qpos: np.ndarray):
    B = qpos.shape[0]
    nb = m.nbody
    qpos = qpos.copy()
    xpos = np.zeros((B, nb, 3))
    xquat = np.zeros((B, nb, 4))
    xquat[:, 0, 0] = 1.0
    xanchor = np.zeros((B, nb, 3))
    xaxis = np.zeros((B, nb, 3))
    for b in range(1, nb):
        p = int(m.body_parentid[b])
        pos = xpos[:, p] + qrot(xquat[:, p], m.body_pos[b])
        quat = qmul(xquat[:, p], np.broadcast_to(m.body_quat[b], (B, 4)))
        j = int(m.body_jntadr[b])
        if j >= 0:
            qa = int(m.jnt_qposadr[j])
            jt = int(m.jnt_type[j])
            if jt == JNT_FREE:
                xanchor[:, b] = qpos[:, qa:qa + 3]
                xaxis[:, b] = [0.0, 0.0, 1.0]
                pos = qpos[:, qa:qa + 3].copy()
                quat = normalize(qpos[:, qa + 3:qa + 7])
                qpos[:, qa + 3:qa + 7] = quat
            else:
                anchor = qrot(quat, m.jnt_pos[j]) + pos
                axis = qrot(quat, m.jnt_axis[j])
                xanchor[:, b], xaxis[:, b] = anchor, axis
                if jt == JNT_HINGE:
                    angle = qpos[:, qa] - m.qpos0[qa]
                    qloc = axis_angle_quat(np.broadcast_to(m.jnt_axis[j], (B, 3)), angle)
                    quat = qmul(quat, qloc)
                    pos = anchor - qrot(quat, m.jnt_pos[j])
                else:  # slide
                    pos = pos + axis * (qpos[:, qa] - m.qpos0[qa])[:, None]
        xpos[:, b] = pos
        xquat[:, b] = normalize(quat)
    xmat = qmat(xquat)
    xipos = xpos + qrot(xquat, m.body_ipos[None])
    ximat = qmat(qmul(xquat, np.broadcast_to(m.body_iquat[None], xquat.shape)))
    return qpos, xpos, xquat, xmat, xipos, ximat, xanchor, xaxis


def com_pos(m: OModel, xpos, xmat, xipos, ximat, xanchor, xaxis):
    B = xpos.shape[0]
    nb, nv = m.nbody, m.nv
    # subtree COM of each kinematic tree root
    root_com = np.zeros((B, nb, 3))
    for r in np.unique(m.body_rootid):
        sel = m.body_rootid == r
        mass = m.body_mass[sel].sum()
        if mass < MINVAL:
            root_com[:, sel] = xipos[:, r][:, None]
        else:
            root_com[:, sel] = (np.einsum("b,nbi->ni", m.body_mass[sel], xipos[:, sel]) / mass)[:, None]
    off = xipos - root_com
    mass = m.body_mass[None, :, None]
    inert = np.einsum("nbij,bj,nbkj->nbik", ximat, m.body_inertia, ximat)
    o2 = np.sum(off * off, -1)[..., None, None] * np.eye(3)
    inert = inert + mass[..., None] * (o2 - off[..., :, None] * off[..., None, :])
    cinert = np.concatenate([
        inert[..., 0, 0, None], inert[..., 1, 1, None], inert[..., 2, 2, None],
        inert[..., 0, 1, None], inert[..., 0, 2, None], inert[..., 1, 2, None],
        off * mass, np.broadcast_to(mass, (B, nb, 1))], -1)
    cdof = np.zeros((B, nv, 6))
    for b in range(1, nb):
        j = int(m.body_jntadr[b])
        if j < 0:
            continue
        d = int(m.jnt_dofadr[j])
        jt = int(m.jnt_type[j])
        offset = root_com[:, b] - xanchor[:, b]
        if jt == JNT_FREE:
            for i in range(3):
                cdof[:, d + i, 3 + i] = 1.0
                ax = xmat[:, b, :, i]
                cdof[:, d + 3 + i, :3] = ax
                cdof[:, d + 3 + i, 3:] = np.cross(ax, offset)
        elif jt == JNT_HINGE:
            cdof[:, d, :3] = xaxis[:, b]
            cdof[:, d, 3:] = np.cross(xaxis[:, b], offset)
        else:
            cdof[:, d, 3:] = xaxis[:, b]
    return root_com, cinert, cdof


def crb_mass_matrix(m: OModel, cinert, cdof):
    nb = m.nbody
    crb = cinert.copy()
    for b in range(nb - 1, 0, -1):
        p = int(m.body_parentid[b])
        if p > 0:
            crb[:, p] += crb[:, b]
    crb[:, 0] = 0.0
    crb_cdof = inert_mul(crb[:, m.dof_bodyid], cdof)
    M = np.einsum("nik,njk->nij", crb_cdof, cdof)
    mask = m.dof_anc
    M = M * mask[None]
    M = M + np.transpose(M * (~np.eye(m.nv, dtype=bool))[None], (0, 2, 1))
    M = M + np.diag(m.dof_armature)[None]
    return M


def com_vel(m: OModel, cdof, qvel):
    B = qvel.shape[0]
    nb, nv = m.nbody, m.nv
    cvel = np.zeros((B, nb, 6))
    cdof_dot = np.zeros((B, nv, 6))
    cq = cdof * qvel[..., None]
    for b in range(1, nb):
        v = cvel[:, int(m.body_parentid[b])].copy()
        j = int(m.body_jntadr[b])
        if j >= 0:
            d = int(m.jnt_dofadr[j])
            if int(m.jnt_type[j]) == JNT_FREE:
                v = v + cq[:, d:d + 3].sum(1)
                for i in range(3):
                    cdof_dot[:, d + 3 + i] = motion_cross(v, cdof[:, d + 3 + i])
                v = v + cq[:, d + 3:d + 6].sum(1)
            else:
                cdof_dot[:, d] = motion_cross(v, cdof[:, d])
                v = v + cq[:, d]
        cvel[:, b] = v
    return cvel, cdof_dot


def rne(m: OModel, cinert, cdof, cdof_dot, cvel, qvel):
    B = qvel.shape[0]
    nb = m.nbody
    cacc = np.zeros((B, nb, 6))
    cacc[:, 0, 3:] = -m.gravity
    cdq = cdof_dot * qvel[..., None]
    for b in range(1, nb):
        a = cacc[:, int(m.body_parentid[b])].copy()
        j = int(m.body_jntadr[b])
        if j >= 0:
            d = int(m.jnt_dofadr[j])
            a = a + cdq[:, d:d + int(m.body_dofnum[b])].sum(1)
        cacc[:, b] = a
    cfrc = inert_mul(cinert, cacc) + motion_cross_force(cvel, inert_mul(cinert, cvel))
    for b in range(nb - 1, 0, -1):
        cfrc[:, int(m.body_parentid[b])] += cfrc[:, b]
    return np.sum(cdof * cfrc[:, m.dof_bodyid], -1)


def make_frame(a):
    a = normalize(a)
    y = np.array([0.0, 1.0, 0.0])
    z = np.array([0.0, 0.0, 1.0])
    b = np.where(((a[..., 1] > -0.5) & (a[..., 1] < 0.5))[..., None], y, z)
    b = b - a * np.sum(a * b, -1, keepdims=True)
    b = normalize(b)
    return np.stack([a, b, np.cross(a, b)], axis=-2)


def collision(m: OModel, xpos, xmat):
    """Fixed-size contact arrays (MJX): plane-sphere and plane-capsule."""
    B = xpos.shape[0]
    dist = np.zeros((B, m.ncon))
    pos = np.zeros((B, m.ncon, 3))
    frame = np.zeros((B, m.ncon, 3, 3))
    c = 0
    for k in range(m.npair):
        g1, g2 = int(m.pair_geom1[k]), int(m.pair_geom2[k])
        b1, b2 = int(m.geom_bodyid[g1]), int(m.geom_bodyid[g2])

        def gpose(g, b):
            p = xpos[:, b] + np.einsum("nij,j->ni", xmat[:, b], m.geom_pos[g])
            R = np.einsum("nij,jk->nik", xmat[:, b], qmat(m.geom_quat[g]))
            return p, R
        p1, R1 = gpose(g1, b1)
        p2, R2 = gpose(g2, b2)
        n = R1[:, :, 2]
        kind = int(m.pair_kind[k])
        if kind == PAIR_PLANE_SPHERE:
            r = m.geom_size[g2, 0]
            d = np.sum((p2 - p1) * n, -1) - r
            dist[:, c] = d
            pos[:, c] = p2 - n * (r + 0.5 * d)[:, None]
            frame[:, c] = make_frame(n)
            c += 1
        elif kind == PAIR_PLANE_CAPSULE:
            r, hl = m.geom_size[g2, 0], m.geom_size[g2, 1]
            axis = R2[:, :, 2]
            bvec = axis - n * np.sum(n * axis, -1, keepdims=True)
            bn = np.linalg.norm(bvec, axis=-1, keepdims=True)
            bdir = bvec / (bn + 1e-6 * (bn == 0.0))
            y = np.array([0.0, 1.0, 0.0])
            z = np.array([0.0, 0.0, 1.0])
            alt = np.where(((n[:, 1] > -0.5) & (n[:, 1] < 0.5))[:, None], y, z)
            bdir = np.where(bn < 0.5, alt, bdir)
            fr = np.stack([n, bdir, np.cross(n, bdir)], axis=-2)
            seg = axis * hl
            for sgn in (1.0, -1.0):
                cc = p2 + sgn * seg
                d = np.sum((cc - p1) * n, -1) - r
                dist[:, c] = d
                pos[:, c] = cc - n * (r + 0.5 * d)[:, None]
                frame[:, c] = fr
                c += 1
        elif kind in (PAIR_SPHERE_SPHERE, PAIR_SPHERE_CAPSULE, PAIR_CAPSULE_CAPSULE):
            if kind == PAIR_SPHERE_SPHERE:
                q1, q2 = p1, p2
            elif kind == PAIR_SPHERE_CAPSULE:
                seg = R2[:, :, 2] * m.geom_size[g2, 1]
                q1, q2 = p1, closest_segment_point(p2 - seg, p2 + seg, p1)
            else:
                s1, s2 = R1[:, :, 2] * m.geom_size[g1, 1], R2[:, :, 2] * m.geom_size[g2, 1]
                q1, q2 = closest_segment_to_segment_points(p1 - s1, p1 + s1, p2 - s2, p2 + s2)
            d, pp, fr = sphere_sphere(q1, m.geom_size[g1, 0], q2, m.geom_size[g2, 0])
            dist[:, c], pos[:, c], frame[:, c] = d, pp, fr
            c += 1
        else:
            raise NotImplementedError(kind)
    return dist, pos, frame


def sphere_sphere(pos1, r1, pos2, r2):
    """mjx collision_primitive._sphere_sphere."""
    dvec = pos2 - pos1
    dist = np.linalg.norm(dvec, axis=-1, keepdims=True)
    n = dvec / (dist + 1e-6 * (dist == 0.0))
    n = np.where(dist == 0.0, np.array([1.0, 0.0, 0.0]), n)
    d = dist[:, 0] - (r1 + r2)
    return d, pos1 + n * (r1 + d * 0.5)[:, None], make_frame(n)


def closest_segment_point(a, b, pt):
    ab = b - a
    t = np.sum((pt - a) * ab, -1, keepdims=True) / (np.sum(ab * ab, -1, keepdims=True) + 1e-6)
    return a + np.clip(t, 0.0, 1.0) * ab


def closest_segment_to_segment_points(a0, a1, b0, b1):
    """mjx math.closest_segment_to_segment_points."""
    def nwn(x):
        n = np.linalg.norm(x, axis=-1, keepdims=True)
        return x / (n + 1e-6 * (n == 0.0)), n
    dir_a, len_a = nwn(a1 - a0)
    dir_b, len_b = nwn(b1 - b0)
    ha, hb = len_a * 0.5, len_b * 0.5
    a_mid, b_mid = a0 + dir_a * ha, b0 + dir_b * hb
    trans = a_mid - b_mid
    dd = np.sum(dir_a * dir_b, -1, keepdims=True)
    dat = np.sum(dir_a * trans, -1, keepdims=True)
    dbt = np.sum(dir_b * trans, -1, keepdims=True)
    denom = 1 - dd * dd
    ta = (-dat + dd * dbt) / (denom + 1e-6)
    tb = dbt + ta * dd
    ta, tb = np.clip(ta, -ha, ha), np.clip(tb, -hb, hb)
    best_a, best_b = a_mid + dir_a * ta, b_mid + dir_b * tb
    new_a = closest_segment_point(a0, a1, best_b)
    new_b = closest_segment_point(b0, b1, best_a)
    d1 = np.sum((new_a - best_b) ** 2, -1, keepdims=True)
    d2 = np.sum((new_b - best_a) ** 2, -1, keepdims=True)
    return np.where(d1 < d2, new_a, best_a), np.where(d1 < d2, best_b, new_b)


def _kbi(m: OModel, solref, solimp, pos):
    timeconst, dampratio = solref[..., 0], solref[..., 1]
    timeconst = np.maximum(timeconst, 2 * m.timestep)  # refsafe
    dmin, dmax, width, mid, power = (solimp[..., i] for i in range(5))
    dmin = np.clip(dmin, MINIMP, MAXIMP)
    dmax = np.clip(dmax, MINIMP, MAXIMP)
    width = np.maximum(MINVAL, width)
    mid = np.clip(mid, MINIMP, MAXIMP)
    power = np.maximum(1, power)
    k = 1 / (dmax * dmax * timeconst * timeconst * dampratio * dampratio)
    b = 2 / (dmax * timeconst)
    k = np.where(solref[..., 0] <= 0, -solref[..., 0] / (dmax * dmax), k)
    b = np.where(solref[..., 1] <= 0, -solref[..., 1] / dmax, b)
    imp_x = np.abs(pos) / width
    imp_a = (1.0 / np.power(mid, power - 1)) * np.power(imp_x, power)
    imp_b = 1 - (1.0 / np.power(1 - mid, power - 1)) * np.power(np.maximum(1 - imp_x, 0.0), power)
    imp_y = np.where(imp_x < mid, imp_a, imp_b)
    imp = dmin + imp_y * (dmax - dmin)
    imp = np.clip(imp, dmin, dmax)
    imp = np.where(imp_x > 1.0, dmax, imp)
    return k, b, imp


def make_constraint(m: OModel, qpos, qvel, cdof, root_com, con_dist, con_pos, con_frame):
    """Rows: joint limits (joint order), then contacts: pyramidal edges (2(condim-1) per
    contact) or elliptic rows (condim per contact: normal, 2 tangents[, torsion, 2 rolling])."""
    B = qpos.shape[0]
    nv = m.nv
    J = np.zeros((B, m.nefc, nv))
    pos = np.zeros((B, m.nefc))
    D = np.zeros((B, m.nefc))
    aref = np.zeros((B, m.nefc))
    active = np.zeros((B, m.nefc), dtype=bool)
    for r, j in enumerate(m.lim_jnt):
        qa, d = int(m.jnt_qposadr[j]), int(m.jnt_dofadr[j])
        dmin = qpos[:, qa] - m.jnt_range[j, 0]
        dmax = m.jnt_range[j, 1] - qpos[:, qa]
        p = np.minimum(dmin, dmax) - m.jnt_margin[j]
        act = p < 0
        sign = (dmin < dmax) * 2.0 - 1.0
        J[:, r, d] = sign * act
        pos[:, r] = p * act
        active[:, r] = act
        k_, b_, imp = _kbi(m, np.broadcast_to(m.jnt_solref[j], (B, 2)), np.broadcast_to(m.jnt_solimp[j], (B, 5)), p * act)
        R = np.maximum(m.dof_invweight0[d] * (1 - imp) / imp, MINVAL)
        D[:, r] = np.where(act, 1.0 / R, 0.0)
        aref[:, r] = np.where(act, -b_ * (sign * qvel[:, d]) - k_ * imp * p, 0.0)
    cones = []
    for c in range(m.ncon):
        k = int(m.con_pair[c])
        g1, g2 = int(m.pair_geom1[k]), int(m.pair_geom2[k])
        b1, b2 = int(m.geom_bodyid[g1]), int(m.geom_bodyid[g2])
        fri = m.pair_friction[k]
        dim = int(m.pair_condim[k])
        incl = m.pair_margin[k] - m.pair_gap[k]
        t = m.body_invweight0[b1, 0] + m.body_invweight0[b2, 0]
        p = con_pos[:, c]
        d = con_dist[:, c] - incl
        act = d < 0

        def jac(body):
            offset = p - root_com[:, body]
            jp_ = cdof[..., 3:] + np.cross(cdof[..., :3], offset[:, None, :])
            mask = m.body_dofmask[body][None, :, None]
            return jp_ * mask, cdof[..., :3] * mask
        jp2, jr2 = jac(b2)
        jp1, jr1 = jac(b1)
        dp = np.einsum("nij,nvj->niv", con_frame[:, c], jp2 - jp1)   # [B,3,nv]
        dr = np.einsum("nij,nvj->niv", con_frame[:, c], jr2 - jr1)
        k_, b_, imp = _kbi(m, np.broadcast_to(m.pair_solref[k], (B, 2)), np.broadcast_to(m.pair_solimp[k], (B, 5)), d * act)
        r0 = int(m.con_row0[c])
        if not m.elliptic:
            iw = (t + fri[0] * fri[0] * t) * 2 * fri[0] * fri[0] / m.impratio
            R = np.maximum(iw * (1 - imp) / imp, MINVAL)
            full = np.concatenate([dp, dr], 1)
            edges = []
            for i in range(1, dim):
                edges += [full[:, 0] + fri[i - 1] * full[:, i], full[:, 0] - fri[i - 1] * full[:, i]]
            if dim == 1:
                edges = [full[:, 0]]
            for e_i, e in enumerate(edges):
                row = r0 + e_i
                J[:, row] = e * act[:, None]
                pos[:, row] = d * act
                active[:, row] = act
                D[:, row] = np.where(act, 1.0 / R, 0.0)
                aref[:, row] = np.where(act, -b_ * np.einsum("nv,nv->n", e, qvel) - k_ * imp * d, 0.0)
        else:
            rows = np.concatenate([dp, dr], 1)[:, :dim]
            Rn = np.maximum(t * (1 - imp) / imp, MINVAL)
            for i in range(dim):
                row = r0 + i
                J[:, row] = rows[:, i] * act[:, None]
                active[:, row] = act
                if i == 0:
                    Ri = Rn
                    pos[:, row] = d * act
                    aref[:, row] = np.where(act, -b_ * np.einsum("nv,nv->n", rows[:, 0], qvel) - k_ * imp * d, 0.0)
                else:
                    Ri = np.maximum(Rn / m.impratio * fri[0] * fri[0] / (fri[i - 1] * fri[i - 1]), MINVAL)
                    aref[:, row] = np.where(act, -b_ * np.einsum("nv,nv->n", rows[:, i], qvel), 0.0)
                D[:, row] = np.where(act, 1.0 / Ri, 0.0)
            cones.append((r0, dim, fri[0] / np.sqrt(m.impratio), fri[:dim - 1].copy()))
    m._cones = cones   # static per model: (row0, dim, mu, friction[dim-1]) of each elliptic contact
    return J, D, aref, pos, active


# ---------------------------------------------------------------------------
# Newton solver (MJX solver.py), batched with per-sample done masks.
# Elliptic cones follow MuJoCo's primal cone cost: with N = mu*x0, T = |fri o x_1..| the
# zones are top (N >= mu T: zero), bottom (mu N + T <= 0: plain quadratic) and middle
# (0.5 Dm (N - mu T)^2, Dm = D0 / (mu^2 (1 + mu^2))).
# ---------------------------------------------------------------------------
class _Ctx:
    pass


def _cone_state(cone, Jaref, D):
    r0, dim, mu, fri = cone
    x = Jaref[:, r0:r0 + dim]
    N = mu * x[:, 0]
    U = x[:, 1:] * fri
    T = np.sqrt(np.sum(U * U, -1))
    bottom = ((T <= 0) & (N < 0)) | ((T > 0) & (mu * N + T <= 0))
    middle = (T > 0) & (N < mu * T) & (mu * N + T > 0)
    Dm = D[:, r0] / max(mu * mu * (1 + mu * mu), MINVAL)
    return x, N, U, T, bottom, middle, Dm


def _row_active(m, ctx, D):
    """Rows treated as plain quadratics: limits with Jaref<0, pyramidal edges with Jaref<0,
    all rows of an elliptic contact that sits in the bottom zone."""
    act = ctx.Jaref < 0
    for cone in getattr(m, "_cones", []):
        r0, dim = cone[0], cone[1]
        _, _, _, _, bottom, _, _ = _cone_state(cone, ctx.Jaref, D)
        act[:, r0:r0 + dim] = bottom[:, None]
    return act


def _update_constraint(m, ctx, J, D, qfrc_smooth, qacc_smooth):
    act = _row_active(m, ctx, D)
    force = D * -ctx.Jaref * act
    cost = 0.5 * np.sum(D * ctx.Jaref * ctx.Jaref * act, -1)
    for cone in getattr(m, "_cones", []):
        r0, dim, mu, fri = cone
        x, N, U, T, bottom, middle, Dm = _cone_state(cone, ctx.Jaref, D)
        Ts = np.where(T > 0, T, 1.0)
        NmT = N - mu * T
        f0 = -Dm * NmT * mu
        force[:, r0] += np.where(middle, f0, 0.0)
        force[:, r0 + 1:r0 + dim] += np.where(middle[:, None], -(f0 / Ts)[:, None] * U * fri, 0.0)
        cost += np.where(middle, 0.5 * Dm * NmT * NmT, 0.0)
    ctx.efc_force = force
    ctx.qfrc_constraint = np.einsum("nrv,nr->nv", J, force)
    ctx.gauss = 0.5 * np.sum((ctx.Ma - qfrc_smooth) * (ctx.qacc - qacc_smooth), -1)
    ctx.prev_cost = ctx.cost
    ctx.cost = cost + ctx.gauss


def cone_hessian(cone, Jaref, D):
    """Analytic Hessian (w.r.t. the contact's rows of Jaref) of the middle-zone cone cost."""
    r0, dim, mu, fri = cone
    x, N, U, T, bottom, middle, Dm = _cone_state(cone, Jaref, D)
    B = Jaref.shape[0]
    Ts = np.where(T > 0, T, 1.0)
    s = x[:, 0] - T
    H = np.zeros((B, dim, dim))
    sc = Dm * mu * mu
    H[:, 0, 0] = sc
    H[:, 0, 1:] = -sc[:, None] * fri * U / Ts[:, None]
    H[:, 1:, 0] = H[:, 0, 1:]
    yy = U[:, :, None] * U[:, None, :] / (Ts * Ts)[:, None, None] * (1 + s / Ts)[:, None, None]
    dd = np.eye(dim - 1)[None] * (s / Ts)[:, None, None]
    H[:, 1:, 1:] = sc[:, None, None] * (fri[:, None] * fri[None, :])[None] * (yy - dd)
    return H * middle[:, None, None]


def _update_gradient(m, ctx, M, J, D, qfrc_smooth):
    ctx.grad = ctx.Ma - qfrc_smooth - ctx.qfrc_constraint
    act = _row_active(m, ctx, D)
    H = M + np.einsum("nrv,nr,nrw->nvw", J, D * act, J)
    for cone in getattr(m, "_cones", []):
        r0, dim = cone[0], cone[1]
        Hc = cone_hessian(cone, ctx.Jaref, D)
        Jc = J[:, r0:r0 + dim]
        H = H + np.einsum("nrv,nrs,nsw->nvw", Jc, Hc, Jc)
    ctx.Mgrad = np.linalg.solve(H, ctx.grad[..., None])[..., 0]
    ctx.H = H


def _ctx_create(m, M, J, D, aref, qfrc_smooth, qacc_smooth, qacc, grad=True):
    ctx = _Ctx()
    ctx.qacc = qacc.copy()
    ctx.Jaref = np.einsum("nrv,nv->nr", J, qacc) - aref
    ctx.Ma = np.einsum("nvw,nw->nv", M, qacc)
    ctx.cost = np.full(qacc.shape[0], np.inf)
    ctx.prev_cost = np.zeros(qacc.shape[0])
    _update_constraint(m, ctx, J, D, qfrc_smooth, qacc_smooth)
    if grad:
        _update_gradient(m, ctx, M, J, D, qfrc_smooth)
        ctx.search = -ctx.Mgrad
    return ctx


# False: MJX's bracket rule (the reference).  True: the checker of the -DDIAL_ROBUST_LS builds of the kernels
# (a bracket that only narrows); set by the tests that compare those builds.
LS_NARROWING = False


def _linesearch(m: OModel, ctx, M, J, D, qfrc_smooth):
    B = ctx.qacc.shape[0]
    scale = m.meaninertia * max(1, m.nv)
    smag = np.linalg.norm(ctx.search, axis=-1) * scale
    gtol = m.tolerance * m.ls_tolerance * smag
    mv = np.einsum("nvw,nw->nv", M, ctx.search)
    jv = np.einsum("nrv,nv->nr", J, ctx.search)
    quad_gauss = np.stack([
        ctx.gauss,
        np.sum(ctx.search * ctx.Ma, -1) - np.sum(ctx.search * qfrc_smooth, -1),
        0.5 * np.sum(ctx.search * mv, -1)], -1)                       # [B,3]
    quad = np.stack([0.5 * ctx.Jaref * ctx.Jaref, jv * ctx.Jaref, 0.5 * jv * jv], -1) * D[..., None]
    cones = getattr(m, "_cones", [])
    simple = np.ones(m.nefc, dtype=bool)
    cq = []
    for cone in cones:
        r0, dim, mu, fri = cone
        simple[r0:r0 + dim] = False
        x, v = ctx.Jaref[:, r0:r0 + dim], jv[:, r0:r0 + dim]
        cq.append(dict(U0=mu * x[:, 0], V0=mu * v[:, 0], UU=np.sum((x[:, 1:] * fri) ** 2, -1),
                       UV=np.sum(x[:, 1:] * v[:, 1:] * fri * fri, -1), VV=np.sum((v[:, 1:] * fri) ** 2, -1),
                       Dm=D[:, r0] / max(mu * mu * (1 + mu * mu), MINVAL), mu=mu,
                       quad=np.sum(quad[:, r0:r0 + dim], 1)))

    def point(alpha):
        x = ctx.Jaref + alpha[:, None] * jv
        act = (x < 0) & simple[None]
        qt = quad_gauss + np.sum(quad * act[..., None], 1)
        cost = alpha * alpha * qt[:, 2] + alpha * qt[:, 1] + qt[:, 0]
        d0 = 2 * alpha * qt[:, 2] + qt[:, 1]
        d1 = 2 * qt[:, 2]
        for q in cq:
            mu = q["mu"]
            N = q["U0"] + alpha * q["V0"]
            Tsq = q["UU"] + alpha * (2 * q["UV"] + alpha * q["VV"])
            T = np.sqrt(np.maximum(Tsq, 0.0))
            bottom = ((Tsq <= 0) & (N < 0)) | ((Tsq > 0) & (mu * N + T <= 0))
            middle = (Tsq > 0) & (N < mu * T) & (mu * N + T > 0)
            qq = q["quad"]
            cost = cost + np.where(bottom, alpha * alpha * qq[:, 2] + alpha * qq[:, 1] + qq[:, 0], 0.0)
            d0 = d0 + np.where(bottom, 2 * alpha * qq[:, 2] + qq[:, 1], 0.0)
            d1 = d1 + np.where(bottom, 2 * qq[:, 2], 0.0)
            Ts = np.where(T > 0, T, 1.0)
            Tsqs = np.where(Tsq > 0, Tsq, 1.0)
            T1 = (q["UV"] + alpha * q["VV"]) / Ts
            T2 = q["VV"] / Ts - (q["UV"] + alpha * q["VV"]) * T1 / Tsqs
            NmT = N - mu * T
            N1 = q["V0"]
            cost = cost + np.where(middle, 0.5 * q["Dm"] * NmT * NmT, 0.0)
            d0 = d0 + np.where(middle, q["Dm"] * NmT * (N1 - mu * T1), 0.0)
            d1 = d1 + np.where(middle, q["Dm"] * ((N1 - mu * T1) ** 2 + NmT * (-mu * T2)), 0.0)
        d1 = d1 + (d1 == 0) * MINVAL
        return np.stack([alpha, cost, d0, d1], -1)                    # [B,4]

    def sel(c, a, b):
        return np.where(c[:, None], a, b)

    p0 = point(np.zeros(B))
    lo = point(p0[:, 0] - p0[:, 2] / p0[:, 3])
    lesser = lo[:, 2] < p0[:, 2]
    hi = sel(lesser, p0, lo)
    lo = sel(lesser, lo, p0)
    swap = np.ones(B, dtype=bool)
    ls_iter = np.zeros(B, dtype=np.int64)
    for _ in range(m.ls_iterations):
        done = ls_iter >= m.ls_iterations
        done |= ~swap
        done |= (lo[:, 2] < 0) & (lo[:, 2] > -gtol)
        done |= (hi[:, 2] > 0) & (hi[:, 2] < gtol)
        go = ~done
        if not go.any():
            break
        lo_next = point(lo[:, 0] - lo[:, 2] / lo[:, 3])
        hi_next = point(hi[:, 0] - hi[:, 2] / hi[:, 3])
        mid = point(0.5 * (lo[:, 0] + hi[:, 0]))
        if LS_NARROWING:
            swap_lo_next = (lo[:, 2] < lo_next[:, 2]) & (lo_next[:, 2] < 0)
        else:
            swap_lo_next = (lo[:, 2] > 0) | (lo[:, 2] < lo_next[:, 2])
        nlo = sel(swap_lo_next, lo_next, lo)
        swap_lo_mid = (mid[:, 2] < 0) & (nlo[:, 2] < mid[:, 2])
        nlo = sel(swap_lo_mid, mid, nlo)
        if LS_NARROWING:
            swap_hi_next = (hi[:, 2] > hi_next[:, 2]) & (hi_next[:, 2] > 0)
        else:
            swap_hi_next = (hi[:, 2] < 0) | (hi[:, 2] > hi_next[:, 2])
        nhi = sel(swap_hi_next, hi_next, hi)
        swap_hi_mid = (mid[:, 2] > 0) & (nhi[:, 2] > mid[:, 2])
        nhi = sel(swap_hi_mid, mid, nhi)
        nswap = swap_lo_next | swap_lo_mid | swap_hi_next | swap_hi_mid
        lo = sel(go, nlo, lo)
        hi = sel(go, nhi, hi)
        swap = np.where(go, nswap, swap)
        ls_iter = ls_iter + go
    improved = (lo[:, 1] < p0[:, 1]) | (hi[:, 1] < p0[:, 1])
    alpha = np.where(lo[:, 1] < hi[:, 1], lo[:, 0], hi[:, 0])
    step = (improved * alpha)[:, None]
    ctx.qacc = ctx.qacc + step * ctx.search
    ctx.Ma = ctx.Ma + step * mv
    ctx.Jaref = ctx.Jaref + step * jv
    ctx.ls_alpha = improved * alpha


def solve(m: OModel, M, J, D, aref, qfrc_smooth, qacc_smooth, qacc_warmstart):
    B = qacc_smooth.shape[0]
    scale = m.meaninertia * max(1, m.nv)
    warm = _ctx_create(m, M, J, D, aref, qfrc_smooth, qacc_smooth, qacc_warmstart, grad=False)
    smth = _ctx_create(m, M, J, D, aref, qfrc_smooth, qacc_smooth, qacc_smooth, grad=False)
    qacc0 = np.where((warm.cost < smth.cost)[:, None], qacc_warmstart, qacc_smooth)
    ctx = _ctx_create(m, M, J, D, aref, qfrc_smooth, qacc_smooth, qacc0)
    niter = np.zeros(B, dtype=np.int64)
    fields = ("qacc", "Jaref", "Ma", "cost", "prev_cost", "gauss", "efc_force",
              "qfrc_constraint", "grad", "Mgrad", "search")
    for it in range(m.iterations):
        improvement = (ctx.prev_cost - ctx.cost) / scale
        gradient = np.linalg.norm(ctx.grad, axis=-1) / scale
        done = niter >= m.iterations
        if m.iterations != 1:
            done = done | (improvement < m.tolerance) | (gradient < m.tolerance)
        go = ~done
        if not go.any():
            break
        old = {f: getattr(ctx, f).copy() for f in fields}
        _linesearch(m, ctx, M, J, D, qfrc_smooth)
        _update_constraint(m, ctx, J, D, qfrc_smooth, qacc_smooth)
        _update_gradient(m, ctx, M, J, D, qfrc_smooth)
        ctx.search = -ctx.Mgrad
        for f in fields:
            new = getattr(ctx, f)
            g = go.reshape((B,) + (1,) * (new.ndim - 1))
            setattr(ctx, f, np.where(g, new, old[f]))
        niter = niter + go
    return ctx.qacc, niter, ctx.qfrc_constraint


# ---------------------------------------------------------------------------
# forward / step
# ---------------------------------------------------------------------------
def forward(m: OModel, qpos, qvel, ctrl, qacc_warmstart) -> Data:
    B = qpos.shape[0]
    qpos_n, xpos, xquat, xmat, xipos, ximat, xanchor, xaxis = kinematics(m, qpos)
    root_com, cinert, cdof = com_pos(m, xpos, xmat, xipos, ximat, xanchor, xaxis)
    M = crb_mass_matrix(m, cinert, cdof)
    site_xpos = xpos[:, m.site_bodyid] + np.einsum("nsij,sj->nsi", xmat[:, m.site_bodyid], m.site_pos)
    con_dist, con_pos, con_frame = collision(m, xpos, xmat)
    J, D, aref, epos, _ = make_constraint(m, qpos_n, qvel, cdof, root_com, con_dist, con_pos, con_frame)
    cvel, cdof_dot = com_vel(m, cdof, qvel)
    qfrc_passive = -m.dof_damping[None] * qvel
    qfrc_bias = rne(m, cinert, cdof, cdof_dot, cvel, qvel)
    # actuation (joint transmissions): force = gain*ctrl + b0 + b1*q + b2*qd
    c = np.where(m.actuator_ctrllimited[None].astype(bool),
                 np.clip(ctrl, m.actuator_ctrlrange[:, 0], m.actuator_ctrlrange[:, 1]), ctrl)
    qa, da = m.actuator_qposadr, m.actuator_dofadr
    force = (m.actuator_gain[None] * c + m.actuator_bias[None, :, 0]
             + m.actuator_bias[None, :, 1] * qpos_n[:, qa] + m.actuator_bias[None, :, 2] * qvel[:, da])
    force = np.where(m.actuator_forcelimited[None].astype(bool),
                     np.clip(force, m.actuator_forcerange[:, 0], m.actuator_forcerange[:, 1]), force)
    qfrc_actuator = np.zeros((B, m.nv))
    np.add.at(qfrc_actuator, (slice(None), da), force * m.actuator_gear[None])
    qfrc_smooth = qfrc_passive - qfrc_bias + qfrc_actuator
    qacc_smooth = np.linalg.solve(M, qfrc_smooth[..., None])[..., 0]
    if m.nefc == 0:
        qacc, niter, qfc = qacc_smooth, np.zeros(B, dtype=np.int64), np.zeros((B, m.nv))
    else:
        qacc, niter, qfc = solve(m, M, J, D, aref, qfrc_smooth, qacc_smooth, qacc_warmstart)
    return Data(qpos_n, xpos, xquat, xmat, xipos, ximat, xanchor, xaxis, root_com, cinert, cdof, M,
                cvel, cdof_dot, qfrc_bias, qfrc_passive, qfrc_actuator, qfrc_smooth, qacc_smooth,
                site_xpos, con_dist, con_pos, con_frame, J, D, aref, epos, qacc, niter, qfc)


def integrate_pos(m: OModel, qpos, qvel, dt):
    qpos = qpos.copy()
    for j in range(m.njnt):
        qa, d = int(m.jnt_qposadr[j]), int(m.jnt_dofadr[j])
        if int(m.jnt_type[j]) == JNT_FREE:
            qpos[:, qa:qa + 3] += dt * qvel[:, d:d + 3]
            w = qvel[:, d + 3:d + 6]
            n = np.linalg.norm(w, axis=-1, keepdims=True)
            axis = w / (n + 1e-6 * (n == 0.0))
            qr = axis_angle_quat(axis, dt * n[:, 0])
            qpos[:, qa + 3:qa + 7] = normalize(qmul(qpos[:, qa + 3:qa + 7], qr))
        else:
            qpos[:, qa] += dt * qvel[:, d]
    return qpos


def step(m: OModel, qpos, qvel, ctrl, qacc_warmstart):
    """``mjx.step``: forward at the current state, then semi-implicit Euler.
    Returns (qpos', qvel', qacc_warmstart', Data-of-the-forward-pass)."""
    d = forward(m, qpos, qvel, ctrl, qacc_warmstart)
    qacc = d.qacc
    if m.eulerdamp and np.any(m.dof_damping != 0):
        Mh = d.M + np.diag(m.timestep * m.dof_damping)[None]
        qacc = np.linalg.solve(Mh, (d.qfrc_smooth + d.qfrc_constraint)[..., None])[..., 0]
    qvel_new = qvel + m.timestep * qacc
    qpos_new = integrate_pos(m, d.qpos, qvel_new, m.timestep)
    return qpos_new, qvel_new, d.qacc.copy(), d


def brax_views(m: OModel, d: Data) -> Dict[str, np.ndarray]:
    """Brax ``x`` / ``xd`` of brax.mjx.pipeline (mirrored at deploy/dial_plan.py:52-59)."""
    ang = d.cvel[:, 1:, :3]
    off = d.xpos[:, 1:] - d.root_com[:, 1:]
    vel = d.cvel[:, 1:, 3:] - np.cross(off, ang)
    return dict(x_pos=d.xpos[:, 1:], x_rot=d.xquat[:, 1:], xd_ang=ang, xd_vel=vel)
