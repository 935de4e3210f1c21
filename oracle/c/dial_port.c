/*
 * dial_port.c — ORACLE / CPU BASELINE (test infrastructure, not product code).
 *
 * Plain-C, one-sample-at-a-time port of the per-sample step of the DIAL-MPC hot path for the
 * tree / pyramidal-cone models (Go2, H1): the "CPU baseline v1" of SURVEY.md §7-9 / BASELINE.md §3.
 * It restates, scalar and without any batching tricks, exactly what oracle/mjx_oracle.py and
 * oracle/envs_oracle.py restate in NumPy (PARITY UNPINNED like them: the reference ships no
 * golden vectors and its physics lives in un-vendored MJX/Brax):
 *
 *   env.step      dial_mpc/envs/unitree_go2_env.py:126-261, :403-521; unitree_h1_env.py:181-321, :686-830
 *   act2tau       dial_mpc/envs/base_env.py:37-66
 *   mjx.step      third party, call sites unitree_go2_env.py:135,415; unitree_h1_env.py:192
 *   rollout_us    dial_mpc/core/dial_core.py:36-42
 *
 * Built twice by oracle/build_oracle.py: REAL=double (validated against the NumPy oracle to
 * the rounding of the fp32 model constants, ~1e-6: tests/test_c_port.py) and REAL=float (the timed fp32 CPU baseline and the oracle's own
 * fp32-vs-fp64 yardstick).  Only tests/, smoke() and bench.py's CPU arms load it.
 * The dense / elliptic model (Allegro) is not ported: its CPU baseline stays the NumPy oracle.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#if defined(__SSE__)
#include <xmmintrin.h>
#endif
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../../include/dial_b200.h"

#ifndef REAL
#define REAL float
#endif
typedef REAL real;

#define MAXE (4 * DIAL_MAXC + DIAL_MAXV) /* constraint rows: limits + pyramid edges */
#define MINVAL ((real)1e-15)
#define MINIMP ((real)1e-4)
#define MAXIMP ((real)0.9999)
enum { JNT_FREE = 0, JNT_BALL = 1, JNT_SLIDE = 2, JNT_HINGE = 3 };
enum { PAIR_PLANE_SPHERE = 0, PAIR_PLANE_CAPSULE = 1 };

#if defined(__GNUC__)
#define SQRT(x) ((real)sqrt((double)(x)))
#endif
static inline real rsq(real x) { return sizeof(real) == 4 ? (real)sqrtf((float)x) : (real)sqrt((double)x); }
static inline real rmin(real a, real b) { return a < b ? a : b; }
static inline real rmax(real a, real b) { return a > b ? a : b; }
static inline real rclip(real x, real lo, real hi) { return rmin(rmax(x, lo), hi); }

/* ---- small vector / quaternion helpers -------------------------------------------------- */
static inline void cross3(const real* a, const real* b, real* r) {
  real x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline real dot3(const real* a, const real* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void qmul(const real* a, const real* b, real* r) {
  real w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  real x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  real y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  real z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static inline void qrot(const real* q, const real* v, real* r) { /* MJX math.rotate */
  const real* u = q + 1;
  real s = q[0], uv = dot3(u, v), uu = dot3(u, u), c[3];
  cross3(u, v, c);
  for (int i = 0; i < 3; ++i) r[i] = 2 * uv * u[i] + (s * s - uu) * v[i] + 2 * s * c[i];
}
static inline void qnormalize(real* q) {
  real n = rsq(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  real inv = 1 / (n + (real)1e-6 * (n == 0 ? 1 : 0));
  for (int i = 0; i < 4; ++i) q[i] *= inv;
}
static inline void qmat(const real* q, real* m) {
  real w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
static inline void axisangle(const real* axis, real angle, real* q) {
  real s = (real)sin((double)(angle * (real)0.5)), c = (real)cos((double)(angle * (real)0.5));
  q[0] = c; q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
static inline void mcross(const real* u, const real* v, real* r) { /* motion cross */
  real a[3], b[3], c[3];
  cross3(u, v, a); cross3(u + 3, v, b); cross3(u, v + 3, c);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
static inline void mcross_force(const real* v, const real* f, real* r) {
  real a[3], b[3], c[3];
  cross3(v, f, a); cross3(v + 3, f + 3, b); cross3(v, f + 3, c);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}
/* cinert (Ixx Iyy Izz Ixy Ixz Iyz | m*off | m) times a motion vector */
static inline void inert_mul(const real* ci, const real* v, real* r) {
  real c1[3], c2[3];
  cross3(ci + 6, v + 3, c1);
  cross3(ci + 6, v, c2);
  r[0] = ci[0] * v[0] + ci[3] * v[1] + ci[4] * v[2] + c1[0];
  r[1] = ci[3] * v[0] + ci[1] * v[1] + ci[5] * v[2] + c1[1];
  r[2] = ci[4] * v[0] + ci[5] * v[1] + ci[2] * v[2] + c1[2];
  r[3] = ci[9] * v[3] - c2[0]; r[4] = ci[9] * v[4] - c2[1]; r[5] = ci[9] * v[5] - c2[2];
}

/* ---- per-sample workspace ------------------------------------------------------------------ */
typedef struct {
  real qpos[DIAL_MAXQ], qvel[DIAL_MAXV], warm[DIAL_MAXV], ctrl[DIAL_MAXU];
  real xpos[DIAL_MAXB][3], xquat[DIAL_MAXB][4], xmat[DIAL_MAXB][9], xipos[DIAL_MAXB][3], ximat[DIAL_MAXB][9];
  real xanchor[DIAL_MAXB][3], xaxis[DIAL_MAXB][3], rcom[DIAL_MAXB][3];
  real cinert[DIAL_MAXB][10], crb[DIAL_MAXB][10], cdof[DIAL_MAXV][6], cdofdot[DIAL_MAXV][6];
  real cvel[DIAL_MAXB][6], cacc[DIAL_MAXB][6], cfrc[DIAL_MAXB][6];
  real M[DIAL_MAXV][DIAL_MAXV], H[DIAL_MAXV][DIAL_MAXV], L[DIAL_MAXV][DIAL_MAXV];
  real cdist[DIAL_MAXC], cpos[DIAL_MAXC][3], cframe[DIAL_MAXC][9];
  real J[MAXE][DIAL_MAXV], D[MAXE], aref[MAXE];
  int nefc;
  uint32_t anc[DIAL_MAXV];      /* bit j: dof j is ancestor-or-self of dof i */
  uint32_t bodymask[DIAL_MAXB]; /* bit d: dof d moves body b */
  real root_invmass[DIAL_MAXB];
  int con_pair[DIAL_MAXC], con_sub[DIAL_MAXC];
} Work;

static void derive(const dial_model_desc* m, Work* w) {
  for (int i = 0; i < m->nv; ++i) {
    uint32_t mask = 0;
    int j = i;
    while (j >= 0) { mask |= 1u << j; j = m->dof_parentid[j]; }
    w->anc[i] = mask;
  }
  for (int b = 0; b < m->nbody; ++b) {
    uint32_t mask = 0;
    int bb = b;
    while (bb > 0) {
      if (m->body_jntadr[bb] >= 0)
        for (int k = 0; k < m->body_dofnum[bb]; ++k) mask |= 1u << (m->body_dofadr[bb] + k);
      bb = m->body_parentid[bb];
    }
    w->bodymask[b] = mask;
  }
  for (int r = 0; r < m->nbody; ++r) {
    double mass = 0;
    for (int b = 1; b < m->nbody; ++b) if (m->body_rootid[b] == r) mass += m->body_mass[b];
    w->root_invmass[r] = mass > 0 ? (real)(1.0 / mass) : 0;
  }
  int c = 0;
  for (int k = 0; k < m->npair; ++k)
    for (int s = 0; s < m->pair_ncon[k]; ++s) { w->con_pair[c] = k; w->con_sub[c] = s; ++c; }
}

/* ---- mjx.forward pieces (oracle/mjx_oracle.py: kinematics, com_pos, crb, com_vel, rne) ------ */
static void kinematics(const dial_model_desc* m, Work* w) {
  const int nb = m->nbody;
  w->xpos[0][0] = w->xpos[0][1] = w->xpos[0][2] = 0;
  w->xquat[0][0] = 1; w->xquat[0][1] = w->xquat[0][2] = w->xquat[0][3] = 0;
  for (int b = 1; b < nb; ++b) {
    const int p = m->body_parentid[b];
    real pos[3], quat[4], bp[3], bq[4], t[3];
    for (int i = 0; i < 3; ++i) bp[i] = m->body_pos[b][i];
    for (int i = 0; i < 4; ++i) bq[i] = m->body_quat[b][i];
    qrot(w->xquat[p], bp, t);
    for (int i = 0; i < 3; ++i) pos[i] = w->xpos[p][i] + t[i];
    qmul(w->xquat[p], bq, quat);
    const int j = m->body_jntadr[b];
    if (j >= 0) {
      const int qa = m->jnt_qposadr[j], jt = m->jnt_type[j];
      if (jt == JNT_FREE) {
        for (int i = 0; i < 3; ++i) { pos[i] = w->qpos[qa + i]; w->xanchor[b][i] = pos[i]; }
        w->xaxis[b][0] = 0; w->xaxis[b][1] = 0; w->xaxis[b][2] = 1;
        for (int i = 0; i < 4; ++i) quat[i] = w->qpos[qa + 3 + i];
        qnormalize(quat);
        for (int i = 0; i < 4; ++i) w->qpos[qa + 3 + i] = quat[i];
      } else {
        real jp[3], ja[3], anchor[3], axis[3];
        for (int i = 0; i < 3; ++i) { jp[i] = m->jnt_pos[j][i]; ja[i] = m->jnt_axis[j][i]; }
        qrot(quat, jp, t);
        for (int i = 0; i < 3; ++i) anchor[i] = t[i] + pos[i];
        qrot(quat, ja, axis);
        for (int i = 0; i < 3; ++i) { w->xanchor[b][i] = anchor[i]; w->xaxis[b][i] = axis[i]; }
        const real dq = w->qpos[qa] - (real)m->qpos0[qa];
        if (jt == JNT_HINGE) {
          real ql[4], qn[4];
          axisangle(ja, dq, ql);
          qmul(quat, ql, qn);
          for (int i = 0; i < 4; ++i) quat[i] = qn[i];
          qrot(quat, jp, t);
          for (int i = 0; i < 3; ++i) pos[i] = anchor[i] - t[i];
        } else {
          for (int i = 0; i < 3; ++i) pos[i] += axis[i] * dq;
        }
      }
    }
    qnormalize(quat);
    for (int i = 0; i < 3; ++i) w->xpos[b][i] = pos[i];
    for (int i = 0; i < 4; ++i) w->xquat[b][i] = quat[i];
  }
  for (int b = 0; b < nb; ++b) {
    real ip[3], iq[4], t[3], q2[4];
    qmat(w->xquat[b], w->xmat[b]);
    for (int i = 0; i < 3; ++i) ip[i] = m->body_ipos[b][i];
    for (int i = 0; i < 4; ++i) iq[i] = m->body_iquat[b][i];
    qrot(w->xquat[b], ip, t);
    for (int i = 0; i < 3; ++i) w->xipos[b][i] = w->xpos[b][i] + t[i];
    qmul(w->xquat[b], iq, q2);
    qmat(q2, w->ximat[b]);
  }
}

static void com_pos(const dial_model_desc* m, Work* w) {
  const int nb = m->nbody, nv = m->nv;
  for (int b = 0; b < nb; ++b) {
    const int r = m->body_rootid[b];
    real s[3] = {0, 0, 0};
    if (w->root_invmass[r] > 0) {
      for (int c = 0; c < nb; ++c)
        if (m->body_rootid[c] == r)
          for (int i = 0; i < 3; ++i) s[i] += (real)m->body_mass[c] * w->xipos[c][i];
      for (int i = 0; i < 3; ++i) w->rcom[b][i] = s[i] * w->root_invmass[r];
    } else {
      for (int i = 0; i < 3; ++i) w->rcom[b][i] = w->xipos[r][i];
    }
  }
  for (int b = 0; b < nb; ++b) {
    real off[3], I[6];
    const real mass = (real)m->body_mass[b];
    const real* X = w->ximat[b];
    const real d0 = (real)m->body_inertia[b][0], d1 = (real)m->body_inertia[b][1], d2 = (real)m->body_inertia[b][2];
    for (int i = 0; i < 3; ++i) off[i] = w->xipos[b][i] - w->rcom[b][i];
    I[0] = X[0] * X[0] * d0 + X[1] * X[1] * d1 + X[2] * X[2] * d2;
    I[1] = X[3] * X[3] * d0 + X[4] * X[4] * d1 + X[5] * X[5] * d2;
    I[2] = X[6] * X[6] * d0 + X[7] * X[7] * d1 + X[8] * X[8] * d2;
    I[3] = X[0] * X[3] * d0 + X[1] * X[4] * d1 + X[2] * X[5] * d2;
    I[4] = X[0] * X[6] * d0 + X[1] * X[7] * d1 + X[2] * X[8] * d2;
    I[5] = X[3] * X[6] * d0 + X[4] * X[7] * d1 + X[5] * X[8] * d2;
    const real o2 = dot3(off, off);
    real* ci = w->cinert[b];
    ci[0] = I[0] + mass * (o2 - off[0] * off[0]);
    ci[1] = I[1] + mass * (o2 - off[1] * off[1]);
    ci[2] = I[2] + mass * (o2 - off[2] * off[2]);
    ci[3] = I[3] - mass * off[0] * off[1];
    ci[4] = I[4] - mass * off[0] * off[2];
    ci[5] = I[5] - mass * off[1] * off[2];
    ci[6] = mass * off[0]; ci[7] = mass * off[1]; ci[8] = mass * off[2]; ci[9] = mass;
  }
  for (int d = 0; d < nv; ++d) for (int i = 0; i < 6; ++i) w->cdof[d][i] = 0;
  for (int b = 1; b < nb; ++b) {
    const int j = m->body_jntadr[b];
    if (j < 0) continue;
    const int d = m->jnt_dofadr[j], jt = m->jnt_type[j];
    real offset[3];
    for (int i = 0; i < 3; ++i) offset[i] = w->rcom[b][i] - w->xanchor[b][i];
    if (jt == JNT_FREE) {
      for (int i = 0; i < 3; ++i) {
        w->cdof[d + i][3 + i] = 1;
        real ax[3] = {w->xmat[b][i], w->xmat[b][3 + i], w->xmat[b][6 + i]}, c[3];
        cross3(ax, offset, c);
        for (int k = 0; k < 3; ++k) { w->cdof[d + 3 + i][k] = ax[k]; w->cdof[d + 3 + i][3 + k] = c[k]; }
      }
    } else if (jt == JNT_HINGE) {
      real c[3];
      cross3(w->xaxis[b], offset, c);
      for (int k = 0; k < 3; ++k) { w->cdof[d][k] = w->xaxis[b][k]; w->cdof[d][3 + k] = c[k]; }
    } else {
      for (int k = 0; k < 3; ++k) w->cdof[d][3 + k] = w->xaxis[b][k];
    }
  }
}

static void crb_mass_matrix(const dial_model_desc* m, Work* w) {
  const int nb = m->nbody, nv = m->nv;
  memcpy(w->crb, w->cinert, sizeof(w->crb));
  for (int b = nb - 1; b > 0; --b) {
    const int p = m->body_parentid[b];
    if (p > 0) for (int i = 0; i < 10; ++i) w->crb[p][i] += w->crb[b][i];
  }
  for (int i = 0; i < nv; ++i) {
    real f[6];
    inert_mul(w->crb[m->dof_bodyid[i]], w->cdof[i], f);
    for (int j = 0; j < nv; ++j) w->M[i][j] = 0;
    for (int j = 0; j <= i; ++j) {
      if (!((w->anc[i] >> j) & 1u)) continue;
      real s = 0;
      for (int k = 0; k < 6; ++k) s += f[k] * w->cdof[j][k];
      w->M[i][j] = s;
    }
  }
  for (int i = 0; i < nv; ++i) {
    for (int j = 0; j < i; ++j) w->M[j][i] = w->M[i][j];
    w->M[i][i] += (real)m->dof_armature[i];
  }
}

static void com_vel(const dial_model_desc* m, Work* w) {
  const int nb = m->nbody;
  for (int i = 0; i < 6; ++i) w->cvel[0][i] = 0;
  for (int d = 0; d < m->nv; ++d) for (int i = 0; i < 6; ++i) w->cdofdot[d][i] = 0;
  for (int b = 1; b < nb; ++b) {
    real v[6];
    memcpy(v, w->cvel[m->body_parentid[b]], sizeof(v));
    const int j = m->body_jntadr[b];
    if (j >= 0) {
      const int d = m->jnt_dofadr[j];
      if (m->jnt_type[j] == JNT_FREE) {
        for (int k = 0; k < 3; ++k) for (int i = 0; i < 6; ++i) v[i] += w->cdof[d + k][i] * w->qvel[d + k];
        for (int k = 3; k < 6; ++k) mcross(v, w->cdof[d + k], w->cdofdot[d + k]);
        for (int k = 3; k < 6; ++k) for (int i = 0; i < 6; ++i) v[i] += w->cdof[d + k][i] * w->qvel[d + k];
      } else {
        mcross(v, w->cdof[d], w->cdofdot[d]);
        for (int i = 0; i < 6; ++i) v[i] += w->cdof[d][i] * w->qvel[d];
      }
    }
    memcpy(w->cvel[b], v, sizeof(v));
  }
}

static void rne(const dial_model_desc* m, Work* w, real* bias) {
  const int nb = m->nbody;
  for (int i = 0; i < 3; ++i) { w->cacc[0][i] = 0; w->cacc[0][3 + i] = -(real)m->gravity[i]; }
  for (int b = 1; b < nb; ++b) {
    real a[6];
    memcpy(a, w->cacc[m->body_parentid[b]], sizeof(a));
    const int j = m->body_jntadr[b];
    if (j >= 0) {
      const int d = m->jnt_dofadr[j];
      for (int k = 0; k < m->body_dofnum[b]; ++k)
        for (int i = 0; i < 6; ++i) a[i] += w->cdofdot[d + k][i] * w->qvel[d + k];
    }
    memcpy(w->cacc[b], a, sizeof(a));
  }
  for (int b = 0; b < nb; ++b) {
    real f1[6], f2[6], f3[6];
    inert_mul(w->cinert[b], w->cacc[b], f1);
    inert_mul(w->cinert[b], w->cvel[b], f2);
    mcross_force(w->cvel[b], f2, f3);
    for (int i = 0; i < 6; ++i) w->cfrc[b][i] = f1[i] + f3[i];
  }
  for (int b = nb - 1; b > 0; --b) {
    const int p = m->body_parentid[b];
    for (int i = 0; i < 6; ++i) w->cfrc[p][i] += w->cfrc[b][i];
  }
  for (int d = 0; d < m->nv; ++d) {
    real s = 0;
    const real* f = w->cfrc[m->dof_bodyid[d]];
    for (int i = 0; i < 6; ++i) s += w->cdof[d][i] * f[i];
    bias[d] = s;
  }
}

/* ---- collision: plane-sphere, plane-capsule (oracle collision()) ------------------------------ */
static void vnormalize(real* a, real* n_out) {
  real n = rsq(dot3(a, a));
  real inv = 1 / (n + (real)1e-6 * (n == 0 ? 1 : 0));
  a[0] *= inv; a[1] *= inv; a[2] *= inv;
  if (n_out) *n_out = n;
}
static void collide(const dial_model_desc* m, Work* w) {
  for (int c = 0; c < m->ncon; ++c) {
    const int k = w->con_pair[c], g1 = m->pair_geom1[k], g2 = m->pair_geom2[k];
    const int b1 = m->geom_bodyid[g1], b2 = m->geom_bodyid[g2];
    real p1[3], p2[3], R1[9], R2[9], G[9], gq[4];
    for (int i = 0; i < 3; ++i) {
      p1[i] = w->xpos[b1][i]; p2[i] = w->xpos[b2][i];
      for (int j = 0; j < 3; ++j) {
        p1[i] += w->xmat[b1][3 * i + j] * (real)m->geom_pos[g1][j];
        p2[i] += w->xmat[b2][3 * i + j] * (real)m->geom_pos[g2][j];
      }
    }
    for (int i = 0; i < 4; ++i) gq[i] = (real)m->geom_quat[g1][i];
    qmat(gq, G);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
      real s = 0;
      for (int l = 0; l < 3; ++l) s += w->xmat[b1][3 * i + l] * G[3 * l + j];
      R1[3 * i + j] = s;
    }
    for (int i = 0; i < 4; ++i) gq[i] = (real)m->geom_quat[g2][i];
    qmat(gq, G);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
      real s = 0;
      for (int l = 0; l < 3; ++l) s += w->xmat[b2][3 * i + l] * G[3 * l + j];
      R2[3 * i + j] = s;
    }
    real n[3] = {R1[2], R1[5], R1[8]}, t1[3], center[3];
    const real radius = (real)m->geom_size[g2][0];
    const real alt[3] = {0, (n[1] > -0.5 && n[1] < 0.5) ? (real)1 : (real)0, (n[1] > -0.5 && n[1] < 0.5) ? (real)0 : (real)1};
    for (int i = 0; i < 3; ++i) center[i] = p2[i];
    if (m->pair_kind[k] == PAIR_PLANE_CAPSULE) {
      real ax[3] = {R2[2], R2[5], R2[8]}, bd[3], bn;
      const real na = dot3(n, ax);
      for (int i = 0; i < 3; ++i) bd[i] = ax[i] - n[i] * na;
      vnormalize(bd, &bn);
      for (int i = 0; i < 3; ++i) t1[i] = bn < 0.5 ? alt[i] : bd[i];
      const real sgn = w->con_sub[c] == 0 ? 1 : -1;
      for (int i = 0; i < 3; ++i) center[i] = p2[i] + ax[i] * sgn * (real)m->geom_size[g2][1];
    } else {
      /* make_frame(n): normalise n, b = alt - n (n.alt), normalised */
      real nn[3] = {n[0], n[1], n[2]};
      vnormalize(nn, 0);
      const real na = dot3(nn, alt);
      for (int i = 0; i < 3; ++i) t1[i] = alt[i] - nn[i] * na;
      vnormalize(t1, 0);
      for (int i = 0; i < 3; ++i) n[i] = nn[i];
    }
    real dv[3] = {center[0] - p1[0], center[1] - p1[1], center[2] - p1[2]};
    const real n_raw[3] = {R1[2], R1[5], R1[8]};
    const real dist = dot3(dv, n_raw) - radius;
    w->cdist[c] = dist;
    for (int i = 0; i < 3; ++i) w->cpos[c][i] = center[i] - n_raw[i] * (radius + (real)0.5 * dist);
    real t2[3];
    cross3(n, t1, t2);
    for (int i = 0; i < 3; ++i) { w->cframe[c][i] = n[i]; w->cframe[c][3 + i] = t1[i]; w->cframe[c][6 + i] = t2[i]; }
  }
}

/* ---- constraint rows (oracle _kbi, make_constraint; pyramidal) -------------------------------- */
static void kbi(const dial_model_desc* m, const float* solref, const float* solimp, real pos, real* k, real* b, real* imp) {
  real timeconst = rmax((real)solref[0], 2 * (real)m->timestep), dampratio = (real)solref[1];
  real dmin = rclip((real)solimp[0], MINIMP, MAXIMP), dmax = rclip((real)solimp[1], MINIMP, MAXIMP);
  real width = rmax(MINVAL, (real)solimp[2]), mid = rclip((real)solimp[3], MINIMP, MAXIMP), power = rmax(1, (real)solimp[4]);
  *k = 1 / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  *b = 2 / (dmax * timeconst);
  if (solref[0] <= 0) *k = -(real)solref[0] / (dmax * dmax);
  if (solref[1] <= 0) *b = -(real)solref[1] / dmax;
  real x = (real)fabs((double)pos) / width;
  real ya = (1 / (real)pow((double)mid, (double)(power - 1))) * (real)pow((double)x, (double)power);
  real yb = 1 - (1 / (real)pow((double)(1 - mid), (double)(power - 1))) * (real)pow((double)rmax(1 - x, 0), (double)power);
  real y = x < mid ? ya : yb;
  real im = dmin + y * (dmax - dmin);
  im = rclip(im, dmin, dmax);
  if (x > 1) im = dmax;
  *imp = im;
}

static void make_constraint(const dial_model_desc* m, Work* w) {
  const int nv = m->nv;
  int r = 0;
  for (int j = 0; j < m->njnt; ++j) {
    if (!m->jnt_limited[j] || m->jnt_type[j] == JNT_FREE) continue;
    const int qa = m->jnt_qposadr[j], d = m->jnt_dofadr[j];
    real dmin = w->qpos[qa] - (real)m->jnt_range[j][0], dmax = (real)m->jnt_range[j][1] - w->qpos[qa];
    real p = rmin(dmin, dmax) - (real)m->jnt_margin[j];
    const int act = p < 0;
    const real sign = dmin < dmax ? 1 : -1;
    for (int v = 0; v < nv; ++v) w->J[r][v] = 0;
    w->D[r] = 0; w->aref[r] = 0;
    if (act) {
      real k_, b_, imp;
      kbi(m, m->jnt_solref[j], m->jnt_solimp[j], p, &k_, &b_, &imp);
      real R = rmax((real)m->dof_invweight0[d] * (1 - imp) / imp, MINVAL);
      w->J[r][d] = sign;
      w->D[r] = 1 / R;
      w->aref[r] = -b_ * (sign * w->qvel[d]) - k_ * imp * p;
    }
    ++r;
  }
  for (int c = 0; c < m->ncon; ++c) {
    const int k = w->con_pair[c], g1 = m->pair_geom1[k], g2 = m->pair_geom2[k];
    const int b1 = m->geom_bodyid[g1], b2 = m->geom_bodyid[g2];
    const real mu0 = (real)m->pair_friction[k][0], mu1 = (real)m->pair_friction[k][1];
    const real dist = w->cdist[c] - ((real)m->pair_margin[k] - (real)m->pair_gap[k]);
    const int act = dist < 0;
    for (int e = 0; e < 4; ++e) { for (int v = 0; v < nv; ++v) w->J[r + e][v] = 0; w->D[r + e] = 0; w->aref[r + e] = 0; }
    if (act) {
      real k_, b_, imp;
      kbi(m, m->pair_solref[k], m->pair_solimp[k], dist, &k_, &b_, &imp);
      const real t = (real)m->body_invweight0[b1] + (real)m->body_invweight0[b2];
      const real iw = (t + mu0 * mu0 * t) * 2 * mu0 * mu0 / (real)m->impratio;
      const real R = rmax(iw * (1 - imp) / imp, MINVAL);
      for (int v = 0; v < nv; ++v) {
        const real s2 = (real)((w->bodymask[b2] >> v) & 1u), s1 = (real)((w->bodymask[b1] >> v) & 1u);
        if (s2 == 0 && s1 == 0) continue;
        real o2[3], o1[3], c2[3], c1[3], jp[3];
        for (int i = 0; i < 3; ++i) { o2[i] = w->cpos[c][i] - w->rcom[b2][i]; o1[i] = w->cpos[c][i] - w->rcom[b1][i]; }
        cross3(w->cdof[v], o2, c2);
        cross3(w->cdof[v], o1, c1);
        for (int i = 0; i < 3; ++i) jp[i] = s2 * (w->cdof[v][3 + i] + c2[i]) - s1 * (w->cdof[v][3 + i] + c1[i]);
        const real jn = dot3(w->cframe[c], jp), j1 = dot3(w->cframe[c] + 3, jp), j2 = dot3(w->cframe[c] + 6, jp);
        w->J[r][v] = jn + mu0 * j1; w->J[r + 1][v] = jn - mu0 * j1;
        w->J[r + 2][v] = jn + mu1 * j2; w->J[r + 3][v] = jn - mu1 * j2;
      }
      for (int e = 0; e < 4; ++e) {
        real jv = 0;
        for (int v = 0; v < nv; ++v) jv += w->J[r + e][v] * w->qvel[v];
        w->D[r + e] = 1 / R;
        w->aref[r + e] = -b_ * jv - k_ * imp * dist;
      }
    }
    r += 4;
  }
  w->nefc = r;
}

/* ---- dense Cholesky solve of an nv x nv SPD system (np.linalg.solve in the oracle) ------------ */
static void chol_solve(int n, real A[DIAL_MAXV][DIAL_MAXV], real L[DIAL_MAXV][DIAL_MAXV], const real* g, real* x) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      real s = A[i][j];
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
      L[i][j] = (i == j) ? rsq(rmax(s, MINVAL)) : s / L[j][j];
    }
  real y[DIAL_MAXV];
  for (int i = 0; i < n; ++i) {
    real s = g[i];
    for (int k = 0; k < i; ++k) s -= L[i][k] * y[k];
    y[i] = s / L[i][i];
  }
  for (int i = n - 1; i >= 0; --i) {
    real s = y[i];
    for (int k = i + 1; k < n; ++k) s -= L[k][i] * x[k];
    x[i] = s / L[i][i];
  }
}

/* ---- Newton solver (oracle solve / _linesearch; pyramidal rows only) --------------------------- */
typedef struct {
  real qacc[DIAL_MAXV], Ma[DIAL_MAXV], grad[DIAL_MAXV], search[DIAL_MAXV], Jaref[MAXE];
  real gauss, cost, prev_cost;
} Ctx;

static void update_constraint(const dial_model_desc* m, const Work* w, Ctx* c, const real* qfs, const real* qas, real* qfc) {
  const int nv = m->nv;
  real cost = 0;
  for (int v = 0; v < nv; ++v) qfc[v] = 0;
  for (int r = 0; r < w->nefc; ++r) {
    if (c->Jaref[r] < 0) {
      const real f = -w->D[r] * c->Jaref[r];
      cost += (real)0.5 * w->D[r] * c->Jaref[r] * c->Jaref[r];
      for (int v = 0; v < nv; ++v) qfc[v] += w->J[r][v] * f;
    }
  }
  real g = 0;
  for (int v = 0; v < nv; ++v) g += (c->Ma[v] - qfs[v]) * (c->qacc[v] - qas[v]);
  c->gauss = (real)0.5 * g;
  c->prev_cost = c->cost;
  c->cost = cost + c->gauss;
}

static void ctx_create(const dial_model_desc* m, const Work* w, Ctx* c, const real* qfs, const real* qas, const real* qacc, real* qfc) {
  const int nv = m->nv;
  for (int v = 0; v < nv; ++v) c->qacc[v] = qacc[v];
  for (int r = 0; r < w->nefc; ++r) {
    real s = 0;
    for (int v = 0; v < nv; ++v) s += w->J[r][v] * qacc[v];
    c->Jaref[r] = s - w->aref[r];
  }
  for (int i = 0; i < nv; ++i) {
    real s = 0;
    for (int j = 0; j < nv; ++j) s += w->M[i][j] * qacc[j];
    c->Ma[i] = s;
  }
  c->cost = INFINITY;
  c->prev_cost = 0;
  update_constraint(m, w, c, qfs, qas, qfc);
}

static void update_gradient(const dial_model_desc* m, Work* w, Ctx* c, const real* qfs, const real* qfc) {
  const int nv = m->nv;
  for (int v = 0; v < nv; ++v) c->grad[v] = c->Ma[v] - qfs[v] - qfc[v];
  for (int i = 0; i < nv; ++i) for (int j = 0; j <= i; ++j) w->H[i][j] = w->M[i][j];
  for (int r = 0; r < w->nefc; ++r) {
    if (!(c->Jaref[r] < 0)) continue;
    const real d = w->D[r];
    const real* Jr = w->J[r];
    for (int i = 0; i < nv; ++i) {
      if (Jr[i] == 0) continue;
      const real a = d * Jr[i];
      for (int j = 0; j <= i; ++j) w->H[i][j] += a * Jr[j];
    }
  }
  real mg[DIAL_MAXV];
  chol_solve(nv, w->H, w->L, c->grad, mg);
  for (int v = 0; v < nv; ++v) c->search[v] = -mg[v];
}

typedef struct { real alpha, cost, d0, d1; } LSP;

static LSP ls_point(const Work* w, const Ctx* c, const real* jv, const real* qg, real alpha) {
  real q0 = qg[0], q1 = qg[1], q2 = qg[2];
  for (int r = 0; r < w->nefc; ++r) {
    if (c->Jaref[r] + alpha * jv[r] < 0) {
      q0 += (real)0.5 * c->Jaref[r] * c->Jaref[r] * w->D[r];
      q1 += jv[r] * c->Jaref[r] * w->D[r];
      q2 += (real)0.5 * jv[r] * jv[r] * w->D[r];
    }
  }
  LSP p;
  p.alpha = alpha;
  p.cost = alpha * alpha * q2 + alpha * q1 + q0;
  p.d0 = 2 * alpha * q2 + q1;
  p.d1 = 2 * q2;
  if (p.d1 == 0) p.d1 = MINVAL;
  return p;
}

static void linesearch(const dial_model_desc* m, const Work* w, Ctx* c, const real* qfs) {
  const int nv = m->nv;
  const real scale = (real)m->meaninertia * (real)(nv > 1 ? nv : 1);
  real mv[DIAL_MAXV], jv[MAXE], ss = 0, sMa = 0, sMv = 0;
  for (int i = 0; i < nv; ++i) {
    real s = 0;
    for (int j = 0; j < nv; ++j) s += w->M[i][j] * c->search[j];
    mv[i] = s;
  }
  for (int r = 0; r < w->nefc; ++r) {
    real s = 0;
    for (int v = 0; v < nv; ++v) s += w->J[r][v] * c->search[v];
    jv[r] = s;
  }
  for (int v = 0; v < nv; ++v) { ss += c->search[v] * c->search[v]; sMa += c->search[v] * (c->Ma[v] - qfs[v]); sMv += c->search[v] * mv[v]; }
  const real gtol = (real)m->tolerance * (real)m->ls_tolerance * rsq(ss) * scale;
  const real qg[3] = {c->gauss, sMa, (real)0.5 * sMv};
  LSP p0 = ls_point(w, c, jv, qg, 0), lo = ls_point(w, c, jv, qg, p0.alpha - p0.d0 / p0.d1), hi;
  if (lo.d0 < p0.d0) { hi = p0; } else { hi = lo; lo = p0; }
  int swap = 1;
  for (int it = 0; it < m->ls_iterations; ++it) {
    int done = !swap;
    done |= (lo.d0 < 0) && (lo.d0 > -gtol);
    done |= (hi.d0 > 0) && (hi.d0 < gtol);
    if (done) break;
    LSP lo_next = ls_point(w, c, jv, qg, lo.alpha - lo.d0 / lo.d1);
    LSP hi_next = ls_point(w, c, jv, qg, hi.alpha - hi.d0 / hi.d1);
    LSP mid = ls_point(w, c, jv, qg, (real)0.5 * (lo.alpha + hi.alpha));
    int s1 = (lo.d0 > 0) || (lo.d0 < lo_next.d0);
    if (s1) lo = lo_next;
    int s2 = (mid.d0 < 0) && (lo.d0 < mid.d0);
    if (s2) lo = mid;
    int s3 = (hi.d0 < 0) || (hi.d0 > hi_next.d0);
    if (s3) hi = hi_next;
    int s4 = (mid.d0 > 0) && (hi.d0 > mid.d0);
    if (s4) hi = mid;
    swap = s1 || s2 || s3 || s4;
  }
  const int improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
  real alpha = (lo.cost < hi.cost) ? lo.alpha : hi.alpha;
  if (!improved) alpha = 0;
  for (int v = 0; v < nv; ++v) { c->qacc[v] += alpha * c->search[v]; c->Ma[v] += alpha * mv[v]; }
  for (int r = 0; r < w->nefc; ++r) c->Jaref[r] += alpha * jv[r];
}

/* mjx.step: forward + semi-implicit Euler; leaves the kinematic arrays of the forward pass in w */
static void physics_step(const dial_model_desc* m, Work* w, int integrate) {
  const int nv = m->nv;
  kinematics(m, w);
  com_pos(m, w);
  crb_mass_matrix(m, w);
  collide(m, w);
  make_constraint(m, w);
  com_vel(m, w);
  real bias[DIAL_MAXV], qfs[DIAL_MAXV], qas[DIAL_MAXV], qfc[DIAL_MAXV];
  rne(m, w, bias);
  for (int d = 0; d < nv; ++d) qfs[d] = -(real)m->dof_damping[d] * w->qvel[d] - bias[d];
  for (int a = 0; a < m->nu; ++a) {
    real c = w->ctrl[a];
    if (m->actuator_ctrllimited[a]) c = rclip(c, (real)m->actuator_ctrlrange[a][0], (real)m->actuator_ctrlrange[a][1]);
    real force = (real)m->actuator_gain[a] * c + (real)m->actuator_bias[a][0] + (real)m->actuator_bias[a][1] * w->qpos[m->actuator_qposadr[a]]
               + (real)m->actuator_bias[a][2] * w->qvel[m->actuator_dofadr[a]];
    if (m->actuator_forcelimited[a]) force = rclip(force, (real)m->actuator_forcerange[a][0], (real)m->actuator_forcerange[a][1]);
    qfs[m->actuator_dofadr[a]] += force * (real)m->actuator_gear[a];
  }
  for (int i = 0; i < nv; ++i) for (int j = 0; j <= i; ++j) w->H[i][j] = w->M[i][j];
  chol_solve(nv, w->H, w->L, qfs, qas);
  real qacc[DIAL_MAXV];
  if (w->nefc == 0) {
    for (int v = 0; v < nv; ++v) qacc[v] = qas[v];
  } else {
    const real scale = (real)m->meaninertia * (real)(nv > 1 ? nv : 1);
    Ctx cw, cs, c;
    ctx_create(m, w, &cw, qfs, qas, w->warm, qfc);
    ctx_create(m, w, &cs, qfs, qas, qas, qfc);
    ctx_create(m, w, &c, qfs, qas, cw.cost < cs.cost ? w->warm : qas, qfc);
    update_gradient(m, w, &c, qfs, qfc);
    for (int it = 0; it < m->iterations; ++it) {
      real improvement = (c.prev_cost - c.cost) / scale, g2 = 0;
      for (int v = 0; v < nv; ++v) g2 += c.grad[v] * c.grad[v];
      real gradient = rsq(g2) / scale;
      int done = 0;
      if (m->iterations != 1) done = (improvement < (real)m->tolerance) || (gradient < (real)m->tolerance);
      if (done) break;
      linesearch(m, w, &c, qfs);
      update_constraint(m, w, &c, qfs, qas, qfc);
      update_gradient(m, w, &c, qfs, qfc);
    }
    for (int v = 0; v < nv; ++v) qacc[v] = c.qacc[v];
  }
  for (int v = 0; v < nv; ++v) w->warm[v] = qacc[v];
  if (!integrate) return;
  const real dt = (real)m->timestep;
  for (int v = 0; v < nv; ++v) w->qvel[v] += dt * qacc[v];
  for (int j = 0; j < m->njnt; ++j) {
    const int qa = m->jnt_qposadr[j], d = m->jnt_dofadr[j];
    if (m->jnt_type[j] == JNT_FREE) {
      for (int i = 0; i < 3; ++i) w->qpos[qa + i] += dt * w->qvel[d + i];
      real wv[3] = {w->qvel[d + 3], w->qvel[d + 4], w->qvel[d + 5]}, nrm, qr[4], qn[4];
      nrm = rsq(dot3(wv, wv));
      const real inv = 1 / (nrm + (real)1e-6 * (nrm == 0 ? 1 : 0));
      for (int i = 0; i < 3; ++i) wv[i] *= inv;
      axisangle(wv, dt * nrm, qr);
      qmul(w->qpos + qa + 3, qr, qn);
      qnormalize(qn);
      for (int i = 0; i < 4; ++i) w->qpos[qa + 3 + i] = qn[i];
    } else {
      w->qpos[qa] += dt * w->qvel[d];
    }
  }
}

/* ---- rewards (oracle/envs_oracle.py) ---------------------------------------------------------- */
static real foot_step(real duty, real cadence, real amplitude, real phase, real time) {
  const double PI = 3.14159265358979323846;
  double t = (double)time * 2 * PI * (double)cadence + PI;
  double a = t + PI - 2 * PI * (double)phase;
  double angle = a - floor(a / (2 * PI)) * 2 * PI - PI;
  if (duty < 1) angle *= 0.5 / (1 - (double)duty);
  double cl = angle < -PI / 2 ? -PI / 2 : (angle > PI / 2 ? PI / 2 : angle);
  double value = duty < 1 ? cos(cl) : 0.0;
  double fin = fabs(value) >= 1e-6 ? fabs(value) : 0.0;
  return amplitude * (real)fin;
}
static real quat_yaw(const real* q) {
  return (real)atan2((double)(-2 * q[1] * q[2] + 2 * q[0] * q[3]), (double)(q[1] * q[1] + q[0] * q[0] - q[3] * q[3] - q[2] * q[2]));
}

static real reward(const dial_model_desc* m, const dial_plan_desc* c, Work* w, int step, int* stage) {
  const real stepf = (real)step, dt = (real)c->dt;
  const real up_q[3] = {0, 0, 1};
  real up[3];
  qrot(w->xquat[1], up_q, up);
  const real r_upright = -(up[0] * up[0] + up[1] * up[1] + (up[2] - 1) * (up[2] - 1));
  const int tb = c->torso_body;
  const real* rot = w->xquat[tb];
  real off[3], vel[3], cr[3], qc[4] = {rot[0], -rot[1], -rot[2], -rot[3]}, vb[3], ab[3], angd[3];
  for (int i = 0; i < 3; ++i) off[i] = w->xpos[tb][i] - w->rcom[tb][i];
  cross3(off, w->cvel[tb], cr);
  for (int i = 0; i < 3; ++i) { vel[i] = w->cvel[tb][3 + i] - cr[i]; angd[i] = w->cvel[tb][i] * (real)(3.14159265358979323846 / 180.0); }
  qrot(qc, vel, vb);
  qrot(qc, angd, ab);
  if (c->env_id == DIAL_ENV_GO2_SEQJUMP) {
    const int st = *stage;
    real dp[3], r_contact = 0, pen = 0;
    for (int i = 0; i < 3; ++i) dp[i] = w->xpos[tb][i] - (real)c->pose_seq[st][i];
    const real r_pos = -dot3(dp, dp);
    const real dy = quat_yaw(rot) - (real)c->yaw_seq[st];
    for (int i = 0; i < 4; ++i) {
      const real dist = w->cdist[i];
      int penal = dist <= (real)0.001;
      for (int j = 0; j < c->n_stage; ++j) {
        const real dx = w->cpos[i][0] - (real)c->contact_targets[j][i][0], dyy = w->cpos[i][1] - (real)c->contact_targets[j][i][1];
        const int cond = dx * dx + dyy * dyy <= (real)c->contact_radius[j][i] * (real)c->contact_radius[j][i];
        if (cond && j == st) r_contact += rclip(1 - dist, 0, 1);
        penal = penal && !cond;
      }
      pen += penal ? 1 : 0;
    }
    /* fp32 like the reference's JAX arithmetic (50 * 0.02f rounds to 1.0f; in double the float dt would give 0.99999998) */
    int ns = (int)floorf((float)(step + 1) * c->dt / c->jump_dt);
    *stage = ns < c->n_stage - 1 ? ns : c->n_stage - 1;
    return r_pos + r_upright + (real)0.3 * (-dy * dy) + (real)0.1 * r_contact - (real)0.1 * pen + 10;
  }
  const real ramp = stepf * dt / (real)c->ramp_up_time;
  /* randomize_tasks: one-step command override (unitree_go2_env.py:141-163) */
  const float* vel_cmd = step == c->cmd_step ? c->cmd_vel : c->vel_cmd;
  const float* ang_cmd = step == c->cmd_step ? c->cmd_ang : c->ang_cmd;
  const real vtx = rmin((real)vel_cmd[0] * ramp, (real)vel_cmd[0]), vty = rmin((real)vel_cmd[1] * ramp, (real)vel_cmd[1]);
  const real atx = rmin((real)ang_cmd[0] * ramp, (real)ang_cmd[0]), aty = rmin((real)ang_cmd[1] * ramp, (real)ang_cmd[1]);
  const real atz = rmin((real)ang_cmd[2] * ramp, (real)ang_cmd[2]);
  real r_gaits = 0;
  for (int f = 0; f < c->nfeet; ++f) {
    const real zt = foot_step((real)c->gait_duty, (real)c->gait_cadence, (real)c->gait_amplitude, (real)c->gait_phase[f], stepf * dt);
    if (c->env_id == DIAL_ENV_GO2_WALK) {
      const int sid = c->feet_site[f], sb = m->site_bodyid[sid];
      const real* X = w->xmat[sb];
      const real z = w->xpos[sb][2] + X[6] * (real)m->site_pos[sid][0] + X[7] * (real)m->site_pos[sid][1] + X[8] * (real)m->site_pos[sid][2];
      const real e = (zt - z) / (real)0.05;
      r_gaits -= e * e;
    } else if (c->env_id == DIAL_ENV_H1_WALK) {
      const real z = rmin(w->cdist[2 * f], w->cdist[2 * f + 1]);
      r_gaits -= (zt - z) * (zt - z);
    } else {
      const real z = rmin(rmin(w->cdist[4 * f], w->cdist[4 * f + 1]), rmin(w->cdist[4 * f + 2], w->cdist[4 * f + 3]));
      r_gaits -= (zt - z) * (zt - z);
    }
  }
  const real dyaw = quat_yaw(rot) - atz * dt * stepf;
  const real wy = (real)atan2(sin((double)dyaw), cos((double)dyaw));
  const real r_yaw = -wy * wy;
  const real r_vel = -((vb[0] - vtx) * (vb[0] - vtx) + (vb[1] - vty) * (vb[1] - vty));
  const real r_ang = -(ab[2] - atz) * (ab[2] - atz);
  const real hz = w->xpos[tb][2] - (real)c->pos_tar[2];
  const real r_h = -hz * hz;
  if (c->env_id == DIAL_ENV_GO2_WALK)
    return (real)0.1 * r_gaits + (real)0.5 * r_upright + (real)0.3 * r_yaw + r_vel + r_ang + r_h;
  if (c->env_id == DIAL_ENV_H1_LOCO) {
    const real r_ang3 = -((ab[0] - atx) * (ab[0] - atx) + (ab[1] - aty) * (ab[1] - aty) + (ab[2] - atz) * (ab[2] - atz));
    real r_level = 0, r_energy = 0;
    for (int f = 0; f < c->nfeet; ++f) {
      const real* X = w->xmat[m->site_bodyid[c->feet_site[f]]];
      r_level -= X[2] * X[2] + X[5] * X[5] + (X[8] - 1) * (X[8] - 1);
    }
    for (int a = 0; a < m->nu; ++a) { const real e = w->ctrl[a] / (real)c->joint_torque_range[a][1] * w->qvel[6 + a] / 160; r_energy -= e * e; }
    return 10 * r_gaits + (real)0.5 * r_upright + (real)0.5 * r_yaw + r_vel + r_ang3 + (real)0.5 * r_h + (real)0.02 * r_level + (real)0.01 * r_energy;
  }
  real r_energy = 0;
  for (int a = 0; a < m->nu; ++a) { const real e = w->ctrl[a] / (real)c->joint_torque_range[a][1]; r_energy -= e * e; }
  return 5 * r_gaits + (real)0.5 * r_upright + (real)0.1 * r_yaw + r_vel + r_ang + (real)0.5 * r_h + (real)0.01 * r_energy;
}

/* ---- rollout_us for nrows action sequences from one state ---------------------------------------
 * us [nrows,H,nu] (double); outputs (double, each nullable except rewss):
 * rewss [nrows,H], q [nrows,H,nq], qd [nrows,H,nv], xpos [nrows,H,nbody-1,3], warm_out [nrows,nv] */
int port_sizeof_real(void) { return (int)sizeof(real); }

static void rollout_row(const dial_model_desc* m, const dial_plan_desc* c, Work* wp, int row, int H, const double* qpos0,
                        const double* qvel0, const double* warm0, int step0, int stage0, const double* us, double* rewss,
                        double* q, double* qd, double* xpos, double* warm_out) {
  Work* w_ = wp;
#define w (*w_)
  const int nq = m->nq, nv = m->nv, nu = m->nu, nb = m->nbody;
  for (int i = 0; i < nq; ++i) w.qpos[i] = (real)qpos0[i];
  for (int i = 0; i < nv; ++i) { w.qvel[i] = (real)qvel0[i]; w.warm[i] = (real)warm0[i]; }
  int step = step0, stage = stage0;
  for (int t = 0; t < H; ++t) {
    const double* u = us + ((size_t)row * H + t) * nu;
    for (int a = 0; a < nu; ++a) {
      const real an = ((real)u[a] * (real)c->action_scale + 1) * (real)0.5;
      real jt = (real)c->joint_range[a][0] + (real)c->joint_offset[a] + an * ((real)c->joint_range[a][1] - (real)c->joint_range[a][0]);
      jt = rclip(jt, (real)c->physical_joint_range[a][0], (real)c->physical_joint_range[a][1]);
      real ctrl = jt;
      if (c->leg_control_torque) {
        const real tau = (real)c->kp[a] * (jt - w.qpos[7 + a]) - (real)c->kd[a] * w.qvel[6 + a];
        ctrl = rclip(tau, (real)c->joint_torque_range[a][0], (real)c->joint_torque_range[a][1]);
      }
      w.ctrl[a] = ctrl;
    }
    for (int f = 0; f < c->n_frames; ++f) physics_step(m, &w, 1);
    const real r = reward(m, c, &w, step, &stage);
    step += 1;
    const size_t rt = (size_t)row * H + t;
    rewss[rt] = (double)r;
    if (q) for (int i = 0; i < nq; ++i) q[rt * nq + i] = (double)w.qpos[i];
    if (qd) for (int i = 0; i < nv; ++i) qd[rt * nv + i] = (double)w.qvel[i];
    if (xpos) for (int b = 1; b < nb; ++b) for (int i = 0; i < 3; ++i) xpos[(rt * (nb - 1) + (b - 1)) * 3 + i] = (double)w.xpos[b][i];
  }
  if (warm_out) for (int i = 0; i < nv; ++i) warm_out[(size_t)row * nv + i] = (double)w.warm[i];
#undef w
}

/* nthreads: OpenMP threads over the rows (each with its own workspace); <= 1: this thread only */
int port_rollout_mt(const dial_model_desc* m, const dial_plan_desc* c, int nrows, int H, const double* qpos0,
                    const double* qvel0, const double* warm0, int step0, int stage0, const double* us, double* rewss,
                    double* q, double* qd, double* xpos, double* warm_out, int nthreads) {
  if (m->cone != 0 || m->nv > DIAL_MAXV || 4 * m->ncon + m->nv > MAXE) return -1;
  if (c->env_id == DIAL_ENV_ALLEGRO || c->env_id == DIAL_ENV_CUSTOM) return -2;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads > 1 ? nthreads : 1)
#endif
  {
    static _Thread_local Work w;
    derive(m, &w);
#if defined(__SSE__)
    /* flush denormals (the GPU build does too: -use_fast_math implies ftz); without it the fp32
     * build spends most of its time in microcoded denormal arithmetic on some states */
    const unsigned int csr0 = _mm_getcsr();
    _mm_setcsr(csr0 | 0x8040u);
#endif
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 2)
#endif
    for (int row = 0; row < nrows; ++row)
      rollout_row(m, c, &w, row, H, qpos0, qvel0, warm0, step0, stage0, us, rewss, q, qd, xpos, warm_out);
#if defined(__SSE__)
    _mm_setcsr(csr0);
#endif
  }
  return 0;
}

int port_rollout(const dial_model_desc* m, const dial_plan_desc* c, int nrows, int H, const double* qpos0,
                 const double* qvel0, const double* warm0, int step0, int stage0, const double* us, double* rewss,
                 double* q, double* qd, double* xpos, double* warm_out) {
  return port_rollout_mt(m, c, nrows, H, qpos0, qvel0, warm0, step0, stage0, us, rewss, q, qd, xpos, warm_out, 1);
}
