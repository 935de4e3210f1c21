// dial_device.cuh — per-warp device code of the DIAL-MPC sampling core (sm_100a).
//
// One warp owns one sample for the whole horizon ("persistent" over Hsample+1 env
// steps): generalized state, the inertia tree and every intermediate of the
// rigid-body step live in a per-warp slab of shared memory / registers; HBM is
// touched only for the per-step outputs (reward, q, qd, x.pos).  Lanes map to
// bodies (kinematics, tree passes), dofs (mass-matrix rows, vectors) or constraint
// rows (contact pyramid edges), whichever the phase needs.
//
// The physics restates mjx.step (third party; reached by the reference through
// brax PipelineEnv.pipeline_step, dial_mpc/envs/unitree_go2_env.py:135) in fp32:
// kinematics -> COM frame (cinert, cdof) -> CRB mass matrix -> collision ->
// constraint rows -> velocity/RNE bias -> actuation -> Newton solve -> Euler.
//
// The file is also compiled by g++ with -DDIAL_HOST_EMUL (tests/emul) where a warp
// is emulated by 32 lock-step fibers; that build exists only to debug kernel logic
// on machines without a GPU and is never used by the product path.
#pragma once

#include <stdint.h>
#include <math.h>
#include "../../include/dial_b200.h"

#ifdef DIAL_HOST_EMUL
#include "warp_emul.h"
#define DEV inline
#else
#define DEV __device__ __forceinline__
#define DEVNI __device__ __noinline__
DEV void syncwarp() { __syncwarp(); }
DEV float shfl(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
DEV float shfl_xor(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
DEV int shfl_i(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
#endif
#ifndef DEVNI
#define DEVNI inline
#endif

#define DIAL_MAXCHAIN 14   // longest dof ancestor chain (H1: 6 + 5 = 11)
#define DIAL_MAXLEVEL 28
#define DIAL_MAXE 32       // contact pyramid edge rows (4 per contact)
#define DIAL_MINVAL 1e-15f
#define DIAL_MINIMP 1e-4f
#define DIAL_MAXIMP 0.9999f

enum { JNT_FREE = 0, JNT_BALL = 1, JNT_SLIDE = 2, JNT_HINGE = 3 };
enum { PAIR_PLANE_SPHERE = 0, PAIR_PLANE_CAPSULE = 1 };

// ---------------------------------------------------------------------------------
// device-side model: the C-ABI descriptor + host-derived schedules + smem offsets
// ---------------------------------------------------------------------------------
struct DevModel {
  dial_model_desc m;
  // derived on the host (dial_capi.cu: derive_model)
  int32_t maxdepth;
  int32_t nroot, root_body[4];
  float root_invmass[4];
  int32_t body_rootidx[DIAL_MAXB];
  int32_t child_adr[DIAL_MAXB], child_num[DIAL_MAXB], child_ids[DIAL_MAXB];
  int32_t dof_level[DIAL_MAXV], nlevel;
  int32_t level_adr[DIAL_MAXLEVEL + 1], level_dofs[DIAL_MAXV];
  uint32_t dof_ancmask[DIAL_MAXV];    // bit j: dof j is ancestor-or-self of dof i
  uint32_t body_dofmask[DIAL_MAXB];   // bit d: dof d moves body b
  int32_t dof_actuator[DIAL_MAXV];    // actuator driving the dof or -1
  int32_t dof_limited[DIAL_MAXV];     // joint id if the dof's joint is limited else -1
  int32_t con_pair[DIAL_MAXC], con_sub[DIAL_MAXC];
  int32_t nedge;                      // 4 * ncon
  // per-warp shared-memory layout (float offsets)
  int32_t o_xpos, o_xquat, o_xmat, o_xipos, o_cinert, o_cdof, o_cdofdot, o_cvel, o_cacc,
      o_cfrc, o_M, o_L, o_J, o_qpos, o_qvel, o_warm, o_ctrl, o_vec, o_frow, o_cpos,
      o_cframe, o_cdist, o_rcom, o_ldinv, o_site, o_misc, warp_floats;
  int32_t pad_[3];
};

struct DevPlan {
  dial_plan_desc c;
  int32_t pad_[2];
};

// arguments of one rollout launch
struct RolloutArgs {
  int32_t nrows;        // sample rows rolled by this launch
  int32_t H;            // env steps per row
  int32_t mode;         // 0: explicit us; 1: planner (Y0s from eps / key); 2: forward only (pipeline_init)
  int32_t step0, stage0;
  const float* qpos0;
  const float* qvel0;
  const float* warm0;
  const float* us;      // [nrows,H,nu]               (mode 0)
  const float* eps;     // [Ntotal,Hn+1,nu] or null   (mode 1)
  const float* Ybar;    // [Hn+1,nu]
  const float* noise;   // [Hn+1]
  uint32_t key0, key1;
  float* rewss;         // [nrows,H] or null
  float* rews;          // [nrows] mean over H or null
  float* q;             // [nrows,H,nq] or null
  float* qd;            // [nrows,H,nv] or null
  float* xpos;          // [nrows,H,nbody-1,3] or null
  float* qpos_out;      // final state of row 0 (env_step / pipeline_init) or null
  float* qvel_out;
  float* warm_out;
  float* ctrl_out;
  float* dbg;           // optional debug dump (tests)
};

// ---------------------------------------------------------------------------------
// small vector / quaternion helpers
// ---------------------------------------------------------------------------------
struct V3 { float x, y, z; };
struct Q4 { float w, x, y, z; };

DEV V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
DEV V3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }
DEV void st3(float* p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
DEV V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
DEV V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
DEV V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
DEV float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DEV V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
DEV Q4 ldq(const float* p) { Q4 q; q.w = p[0]; q.x = p[1]; q.y = p[2]; q.z = p[3]; return q; }
DEV void stq(float* p, Q4 q) { p[0] = q.w; p[1] = q.x; p[2] = q.y; p[3] = q.z; }
DEV Q4 qmul(Q4 a, Q4 b) {
  Q4 r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
  r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
  return r;
}
// MJX math.rotate(vec, quat)
DEV V3 qrot(Q4 q, V3 v) {
  V3 u = v3(q.x, q.y, q.z);
  float s = q.w;
  V3 r = u * (2.f * dot(u, v)) + v * (s * s - dot(u, u));
  return r + cross(u, v) * (2.f * s);
}
DEV Q4 qnormalize(Q4 q) {
  float n = sqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  float inv = 1.f / (n + 1e-6f * (n == 0.f ? 1.f : 0.f));
  Q4 r; r.w = q.w * inv; r.x = q.x * inv; r.y = q.y * inv; r.z = q.z * inv; return r;
}
DEV void qmat(Q4 q, float* m) {  // row-major 3x3
  float w = q.w, x = q.x, y = q.y, z = q.z;
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2.f * (x * y - w * z); m[2] = 2.f * (x * z + w * y);
  m[3] = 2.f * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2.f * (y * z - w * x);
  m[6] = 2.f * (x * z - w * y); m[7] = 2.f * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
DEV Q4 axisangle(V3 axis, float angle) {
  float s = sinf(0.5f * angle), c = cosf(0.5f * angle);
  Q4 q; q.w = c; q.x = axis.x * s; q.y = axis.y * s; q.z = axis.z * s; return q;
}

// spatial helpers on [ang(3), lin(3)] vectors
DEV void mcross(const float* u, const float* v, float* r) {  // motion cross
  V3 ua = ld3(u), ul = ld3(u + 3), va = ld3(v), vl = ld3(v + 3);
  st3(r, cross(ua, va));
  st3(r + 3, cross(ul, va) + cross(ua, vl));
}
DEV void mcross_force(const float* v, const float* f, float* r) {
  V3 va = ld3(v), vl = ld3(v + 3), fa = ld3(f), fl = ld3(f + 3);
  st3(r, cross(va, fa) + cross(vl, fl));
  st3(r + 3, cross(va, fl));
}
// cinert (10: Ixx Iyy Izz Ixy Ixz Iyz | m*off | m) times motion vector
DEV void inert_mul(const float* ci, const float* v, float* r) {
  V3 va = ld3(v), vl = ld3(v + 3), pos = ld3(ci + 6);
  float m = ci[9];
  V3 ang = v3(ci[0] * va.x + ci[3] * va.y + ci[4] * va.z,
              ci[3] * va.x + ci[1] * va.y + ci[5] * va.z,
              ci[4] * va.x + ci[5] * va.y + ci[2] * va.z) + cross(pos, vl);
  V3 vel = vl * m - cross(pos, va);
  st3(r, ang); st3(r + 3, vel);
}
DEV float dot6(const float* a, const float* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}

DEV float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += shfl_xor(v, o);
  return v;
}
DEV void warp_sum3(float& a, float& b, float& c) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += shfl_xor(a, o); b += shfl_xor(b, o); c += shfl_xor(c, o); }
}

// ---------------------------------------------------------------------------------
// Threefry-2x32 (20 rounds) and the JAX legacy normal sampler
// (jax.random.normal at core/dial_core.py:107-109; third party, restated)
// ---------------------------------------------------------------------------------
DEV uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
DEV void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  uint32_t ks2 = k0 ^ k1 ^ 0x1BD11BDAu;
  x0 += k0; x1 += k1;
#define TF_R(r) { x0 += x1; x1 = rotl32(x1, r); x1 ^= x0; }
  TF_R(13) TF_R(15) TF_R(26) TF_R(6)   x0 += k1;  x1 += ks2 + 1u;
  TF_R(17) TF_R(29) TF_R(16) TF_R(24)  x0 += ks2; x1 += k0 + 2u;
  TF_R(13) TF_R(15) TF_R(26) TF_R(6)   x0 += k0;  x1 += k1 + 3u;
  TF_R(17) TF_R(29) TF_R(16) TF_R(24)  x0 += k1;  x1 += ks2 + 4u;
  TF_R(13) TF_R(15) TF_R(26) TF_R(6)   x0 += ks2; x1 += k0 + 5u;
#undef TF_R
}
// element `i` of jax.random.bits(key, (n,)) with threefry_partitionable=False
DEV uint32_t jax_bits_legacy(uint32_t k0, uint32_t k1, uint32_t i, uint32_t n) {
  uint32_t half = (n + 1u) >> 1;
  uint32_t x0, x1;
  if (i < half) { x0 = i; x1 = i + half; } else { x0 = i - half; x1 = i; }
  bool second = i >= half;
  if (x1 >= n) x1 = 0u;  // odd n: the counter array is padded with one zero
  threefry2x32(k0, k1, x0, x1);
  return second ? x1 : x0;
}
// XLA's single-precision erfinv (Giles' polynomials)
DEV float erfinv_f32(float x) {
  float w = -log1pf(-x * x);
  float p;
  if (w < 5.f) {
    w -= 2.5f;
    p = 2.81022636e-08f;
    p = 3.43273939e-07f + p * w; p = -3.5233877e-06f + p * w; p = -4.39150654e-06f + p * w;
    p = 0.00021858087f + p * w; p = -0.00125372503f + p * w; p = -0.00417768164f + p * w;
    p = 0.246640727f + p * w; p = 1.50140941f + p * w;
  } else {
    w = sqrtf(w) - 3.f;
    p = -0.000200214257f;
    p = 0.000100950558f + p * w; p = 0.00134934322f + p * w; p = -0.00367342844f + p * w;
    p = 0.00573950773f + p * w; p = -0.0076224613f + p * w; p = 0.00943887047f + p * w;
    p = 1.00167406f + p * w; p = 2.83297682f + p * w;
  }
  return p * x;
}
DEV float jax_normal_legacy(uint32_t k0, uint32_t k1, uint32_t i, uint32_t n) {
  uint32_t bits = jax_bits_legacy(k0, k1, i, n);
  float f = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
  const float lo = -0.99999994f;  // nextafter(-1, 0)
  float u = fmaxf(lo, f * 2.0f + lo);
  return 1.41421356237f * erfinv_f32(u);
}

// ---------------------------------------------------------------------------------
// per-warp context
// ---------------------------------------------------------------------------------
struct WarpCtx {
  const DevModel* M;   // in shared memory
  const DevPlan* P;    // in shared memory
  float* s;            // this warp's slab
  int lane;
  // lane-as-dof ancestor chain (self first, then parents)
  int nchain;
  int chain[DIAL_MAXCHAIN];
};

#define SM(name) (w.s + w.M->o_##name)

// ---- sparse reverse-order Cholesky  H = L^T L  on the dof tree (no fill-in) ---------
// L (lower triangle, dense rows) is overwritten in place.  Level-scheduled: all dofs of
// one elimination level (never ancestors of each other) are pivots at the same time.
DEVNI void factor_LTL(WarpCtx& w) {
  const DevModel& M = *w.M;
  const int nv = M.m.nv, lane = w.lane;
  float* L = SM(L);
  float* ldinv = SM(ldinv);
  const int mylevel = lane < nv ? M.dof_level[lane] : -1;
  for (int lv = 0; lv < M.nlevel; ++lv) {
    if (mylevel == lv) {
      float d = sqrtf(fmaxf(L[lane * nv + lane], DIAL_MINVAL));
      float inv = 1.f / d;
      L[lane * nv + lane] = d;
      ldinv[lane] = inv;
#pragma unroll
      for (int c = 1; c < DIAL_MAXCHAIN; ++c)
        if (c < w.nchain) L[lane * nv + w.chain[c]] *= inv;
    }
    syncwarp();
    if (lane < nv && mylevel > lv) {
      for (int pi = M.level_adr[lv]; pi < M.level_adr[lv + 1]; ++pi) {
        int k = M.level_dofs[pi];
        if ((M.dof_ancmask[k] >> lane) & 1u) {
          float lki = L[k * nv + lane];
#pragma unroll
          for (int c = 0; c < DIAL_MAXCHAIN; ++c)
            if (c < w.nchain) L[lane * nv + w.chain[c]] -= lki * L[k * nv + w.chain[c]];
        }
      }
    }
    syncwarp();
  }
}

// solve (L^T L) x = g ; lane d holds g_d on entry and x_d on return.  Uses SM(vec).
DEVNI float solve_LTL(WarpCtx& w, float g) {
  const DevModel& M = *w.M;
  const int nv = M.m.nv, lane = w.lane;
  const float* L = SM(L);
  const float* ldinv = SM(ldinv);
  float* vec = SM(vec);
  const int mylevel = lane < nv ? M.dof_level[lane] : -1;
  float y = g;
  // L^T y = g : leaves -> root
  for (int lv = 0; lv < M.nlevel; ++lv) {
    if (mylevel == lv) { y *= ldinv[lane]; vec[lane] = y; }
    syncwarp();
    if (lane < nv && mylevel > lv) {
      for (int pi = M.level_adr[lv]; pi < M.level_adr[lv + 1]; ++pi) {
        int k = M.level_dofs[pi];
        if ((M.dof_ancmask[k] >> lane) & 1u) y -= L[k * nv + lane] * vec[k];
      }
    }
  }
  // L x = y : root -> leaves
  float x = 0.f;
  for (int lv = M.nlevel - 1; lv >= 0; --lv) {
    if (mylevel == lv) {
      float acc = y;
#pragma unroll
      for (int c = 1; c < DIAL_MAXCHAIN; ++c)
        if (c < w.nchain) acc -= L[lane * nv + w.chain[c]] * vec[w.chain[c]];
      x = acc * ldinv[lane];
      vec[lane] = x;
    }
    syncwarp();
  }
  return x;
}

// y = M x (symmetric, lower triangle stored); lane d holds x_d / returns y_d.  Uses SM(vec).
DEV float mul_M(WarpCtx& w, float x) {
  const int nv = w.M->m.nv, lane = w.lane;
  const float* Mm = SM(M);
  float* vec = SM(vec);
  syncwarp();
  if (lane < nv) vec[lane] = x;
  syncwarp();
  float y = 0.f;
  if (lane < nv) {
    for (int j = 0; j <= lane; ++j) y += Mm[lane * nv + j] * vec[j];
    for (int j = lane + 1; j < nv; ++j) y += Mm[j * nv + lane] * vec[j];
  }
  return y;
}

// contact-edge rows times a dof vector: lane e returns sum_d J[e][d] x_d.  Uses SM(vec).
DEV float mul_J(WarpCtx& w, float x) {
  const int nv = w.M->m.nv, ne = w.M->nedge, lane = w.lane;
  const float* J = SM(J);
  float* vec = SM(vec);
  syncwarp();
  if (lane < nv) vec[lane] = x;
  syncwarp();
  float y = 0.f;
  if (lane < ne)
    for (int d = 0; d < nv; ++d) y += J[lane * nv + d] * vec[d];
  return y;
}

// J^T f for the contact-edge rows: lane e holds f_e, lane d returns sum_e J[e][d] f_e.
DEV float mul_JT(WarpCtx& w, float f) {
  const int nv = w.M->m.nv, ne = w.M->nedge, lane = w.lane;
  const float* J = SM(J);
  float* frow = SM(frow);
  syncwarp();
  if (lane < ne) frow[lane] = f;
  syncwarp();
  float y = 0.f;
  if (lane < nv)
    for (int e = 0; e < ne; ++e) y += J[e * nv + lane] * frow[e];
  return y;
}

// stiffness / damping / impedance of a constraint row (mjx constraint._kbi)
DEV void kbi(float timestep, const float* solref, const float* solimp, float pos,
             float& k, float& b, float& imp) {
  float timeconst = fmaxf(solref[0], 2.f * timestep), dampratio = solref[1];
  float dmin = fminf(fmaxf(solimp[0], DIAL_MINIMP), DIAL_MAXIMP);
  float dmax = fminf(fmaxf(solimp[1], DIAL_MINIMP), DIAL_MAXIMP);
  float width = fmaxf(DIAL_MINVAL, solimp[2]);
  float mid = fminf(fmaxf(solimp[3], DIAL_MINIMP), DIAL_MAXIMP);
  float power = fmaxf(1.f, solimp[4]);
  k = 1.f / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  b = 2.f / (dmax * timeconst);
  if (solref[0] <= 0.f) k = -solref[0] / (dmax * dmax);
  if (solref[1] <= 0.f) b = -solref[1] / dmax;
  float x = fabsf(pos) / width;
  float a_ = (1.f / powf(mid, power - 1.f)) * powf(x, power);
  float b_ = 1.f - (1.f / powf(1.f - mid, power - 1.f)) * powf(fmaxf(1.f - x, 0.f), power);
  float y = x < mid ? a_ : b_;
  imp = dmin + y * (dmax - dmin);
  imp = fminf(fmaxf(imp, dmin), dmax);
  if (x > 1.f) imp = dmax;
}

// ---------------------------------------------------------------------------------
// Newton solver state kept in registers
//   lane d (< nv)   : dof-vector elements + the joint-limit row of dof d (if any)
//   lane e (< nedge): contact pyramid edge row e
// ---------------------------------------------------------------------------------
struct Solver {
  // dof vectors
  float qacc, Ma, grad, Mgrad, search, qfs, qas;  // qfs = qfrc_smooth, qas = qacc_smooth
  // limit row (lane = dof)
  float l_sign, l_D, l_aref, l_Jaref;
  // contact edge row (lane = edge)
  float e_D, e_aref, e_Jaref;
  // scalars (warp-uniform)
  float gauss, cost, prev_cost;
};

DEV void update_constraint(WarpCtx& w, Solver& S) {
  // efc_force = D * -Jaref * active ; qfrc_constraint = J^T force ; costs
  float fl = (S.l_Jaref < 0.f) ? -S.l_D * S.l_Jaref : 0.f;
  float fe = (S.e_Jaref < 0.f) ? -S.e_D * S.e_Jaref : 0.f;
  float qfc = mul_JT(w, fe) + S.l_sign * fl;
  float g = (S.Ma - S.qfs) * (S.qacc - S.qas);
  float c = ((S.l_Jaref < 0.f) ? S.l_D * S.l_Jaref * S.l_Jaref : 0.f)
          + ((S.e_Jaref < 0.f) ? S.e_D * S.e_Jaref * S.e_Jaref : 0.f);
  float dummy = 0.f;
  warp_sum3(g, c, dummy);
  S.gauss = 0.5f * g;
  S.prev_cost = S.cost;
  S.cost = 0.5f * c + S.gauss;
  S.grad = S.Ma - S.qfs - qfc;
}

// H = M + J^T D_active J  (tree-sparse), factor, Mgrad = H^-1 grad
DEV void update_gradient(WarpCtx& w, Solver& S) {
  const DevModel& M = *w.M;
  const int nv = M.m.nv, ne = M.nedge, lane = w.lane;
  const float* Mm = SM(M);
  const float* J = SM(J);
  float* L = SM(L);
  float* frow = SM(frow);
  syncwarp();
  if (lane < ne) frow[lane] = (S.e_Jaref < 0.f) ? S.e_D : 0.f;
  syncwarp();
  if (lane < nv) {
    float acc[DIAL_MAXCHAIN];
#pragma unroll
    for (int c = 0; c < DIAL_MAXCHAIN; ++c) acc[c] = (c < w.nchain) ? Mm[lane * nv + w.chain[c]] : 0.f;
    for (int e = 0; e < ne; ++e) {
      float a = frow[e] * J[e * nv + lane];
      if (a != 0.f) {
#pragma unroll
        for (int c = 0; c < DIAL_MAXCHAIN; ++c)
          if (c < w.nchain) acc[c] += a * J[e * nv + w.chain[c]];
      }
    }
    acc[0] += (S.l_Jaref < 0.f) ? S.l_D : 0.f;  // limit rows are +-e_d: diagonal only
#pragma unroll
    for (int c = 0; c < DIAL_MAXCHAIN; ++c)
      if (c < w.nchain) L[lane * nv + w.chain[c]] = acc[c];
  }
  syncwarp();
  factor_LTL(w);
  S.Mgrad = solve_LTL(w, S.grad);
}

struct LSPoint { float alpha, cost, d0, d1; };

// evaluate the 1-D piecewise-quadratic cost at up to three alphas at once
DEV void ls_points3(const Solver& S, float l_jv, float e_jv, const float* qg,
                    float a0, float a1, float a2, LSPoint& p0, LSPoint& p1, LSPoint& p2) {
  float s[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) s[i] = 0.f;
  const float al[3] = {a0, a1, a2};
  {
    float q0 = 0.5f * S.l_Jaref * S.l_Jaref * S.l_D, q1 = l_jv * S.l_Jaref * S.l_D, q2 = 0.5f * l_jv * l_jv * S.l_D;
#pragma unroll
    for (int i = 0; i < 3; ++i)
      if (S.l_Jaref + al[i] * l_jv < 0.f) { s[3 * i] += q0; s[3 * i + 1] += q1; s[3 * i + 2] += q2; }
    q0 = 0.5f * S.e_Jaref * S.e_Jaref * S.e_D; q1 = e_jv * S.e_Jaref * S.e_D; q2 = 0.5f * e_jv * e_jv * S.e_D;
#pragma unroll
    for (int i = 0; i < 3; ++i)
      if (S.e_Jaref + al[i] * e_jv < 0.f) { s[3 * i] += q0; s[3 * i + 1] += q1; s[3 * i + 2] += q2; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int i = 0; i < 9; ++i) s[i] += shfl_xor(s[i], o);
  }
  LSPoint* out[3] = {&p0, &p1, &p2};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float t0 = qg[0] + s[3 * i], t1 = qg[1] + s[3 * i + 1], t2 = qg[2] + s[3 * i + 2];
    float a = al[i];
    out[i]->alpha = a;
    out[i]->cost = a * a * t2 + a * t1 + t0;
    out[i]->d0 = 2.f * a * t2 + t1;
    out[i]->d1 = 2.f * t2 + (t2 == 0.f ? DIAL_MINVAL : 0.f);
  }
}

DEV void linesearch(WarpCtx& w, Solver& S) {
  const DevModel& M = *w.M;
  const int nv = M.m.nv;
  const float scale = M.m.meaninertia * (float)(nv > 1 ? nv : 1);
  float mv = mul_M(w, S.search);
  float e_jv = mul_J(w, S.search);
  float l_jv = S.l_sign * S.search;
  float ss = S.search * S.search, sMa = S.search * (S.Ma - S.qfs), sMv = S.search * mv;
  warp_sum3(ss, sMa, sMv);
  float gtol = M.m.tolerance * M.m.ls_tolerance * sqrtf(ss) * scale;
  float qg[3] = {S.gauss, sMa, 0.5f * sMv};
  LSPoint p0, lo, hi, t1, t2;
  ls_points3(S, l_jv, e_jv, qg, 0.f, 0.f, 0.f, p0, t1, t2);
  ls_points3(S, l_jv, e_jv, qg, p0.alpha - p0.d0 / p0.d1, 0.f, 0.f, lo, t1, t2);
  if (lo.d0 < p0.d0) { hi = p0; } else { hi = lo; lo = p0; }
  bool swap = true;
  for (int it = 0; it < M.m.ls_iterations; ++it) {
    bool done = !swap;
    done |= (lo.d0 < 0.f) && (lo.d0 > -gtol);
    done |= (hi.d0 > 0.f) && (hi.d0 < gtol);
    if (done) break;
    LSPoint lo_next, hi_next, mid;
    ls_points3(S, l_jv, e_jv, qg, lo.alpha - lo.d0 / lo.d1, hi.alpha - hi.d0 / hi.d1,
               0.5f * (lo.alpha + hi.alpha), lo_next, hi_next, mid);
    bool swap_lo_next = (lo.d0 > 0.f) || (lo.d0 < lo_next.d0);
    if (swap_lo_next) lo = lo_next;
    bool swap_lo_mid = (mid.d0 < 0.f) && (lo.d0 < mid.d0);
    if (swap_lo_mid) lo = mid;
    bool swap_hi_next = (hi.d0 < 0.f) || (hi.d0 > hi_next.d0);
    if (swap_hi_next) hi = hi_next;
    bool swap_hi_mid = (mid.d0 > 0.f) && (hi.d0 > mid.d0);
    if (swap_hi_mid) hi = mid;
    swap = swap_lo_next || swap_lo_mid || swap_hi_next || swap_hi_mid;
  }
  bool improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
  float alpha = (lo.cost < hi.cost) ? lo.alpha : hi.alpha;
  if (!improved) alpha = 0.f;
  S.qacc += alpha * S.search;
  S.Ma += alpha * mv;
  S.l_Jaref += alpha * l_jv;
  S.e_Jaref += alpha * e_jv;
}

// cost of a candidate start point (warm-start selection, solver.py `_Context.create(grad=False)`)
DEV float candidate_cost(WarpCtx& w, Solver& S, float qacc) {
  float Ma = mul_M(w, qacc);
  float eJ = mul_J(w, qacc) - S.e_aref;
  float lJ = S.l_sign * qacc - S.l_aref;
  float g = (Ma - S.qfs) * (qacc - S.qas);
  float c = ((lJ < 0.f) ? S.l_D * lJ * lJ : 0.f) + ((eJ < 0.f) ? S.e_D * eJ * eJ : 0.f);
  float dummy = 0.f;
  warp_sum3(g, c, dummy);
  return 0.5f * c + 0.5f * g;
}

DEV float newton_solve(WarpCtx& w, Solver& S, float warmstart) {
  const DevModel& M = *w.M;
  const int nv = M.m.nv;
  const float scale = M.m.meaninertia * (float)(nv > 1 ? nv : 1);
  float cw = candidate_cost(w, S, warmstart);
  float cs = candidate_cost(w, S, S.qas);
  S.qacc = (cw < cs) ? warmstart : S.qas;
  S.Ma = mul_M(w, S.qacc);
  S.e_Jaref = mul_J(w, S.qacc) - S.e_aref;
  S.l_Jaref = S.l_sign * S.qacc - S.l_aref;
  S.cost = INFINITY;
  S.prev_cost = 0.f;
  update_constraint(w, S);
  update_gradient(w, S);
  S.search = -S.Mgrad;
  for (int it = 0; it < M.m.iterations; ++it) {
    if (M.m.iterations != 1) {
      float g2 = warp_sum(S.grad * S.grad);
      float improvement = (S.prev_cost - S.cost) / scale;
      float gradient = sqrtf(g2) / scale;
      if (improvement < M.m.tolerance || gradient < M.m.tolerance) break;
    }
    linesearch(w, S);
    update_constraint(w, S);
    update_gradient(w, S);
    S.search = -S.Mgrad;
  }
  return S.qacc;
}

// ---------------------------------------------------------------------------------
// one physics step (mjx.step) for the warp's sample.  State (qpos,qvel,warm,ctrl) in
// the slab; kinematic arrays of the forward pass are left in the slab for the reward.
// ---------------------------------------------------------------------------------
DEVNI void physics_step(WarpCtx& w, bool integrate) {
  const DevModel& M = *w.M;
  const dial_model_desc& m = M.m;
  const int lane = w.lane, nb = m.nbody, nv = m.nv;
  float* xpos = SM(xpos); float* xquat = SM(xquat); float* xmat = SM(xmat); float* xipos = SM(xipos);
  float* cinert = SM(cinert); float* cdof = SM(cdof); float* cdofdot = SM(cdofdot);
  float* cvel = SM(cvel); float* cacc = SM(cacc); float* cfrc = SM(cfrc);
  float* Mm = SM(M); float* J = SM(J);
  float* qpos = SM(qpos); float* qvel = SM(qvel); float* warm = SM(warm); float* ctrl = SM(ctrl);
  float* cpos = SM(cpos); float* cframe = SM(cframe); float* cdist = SM(cdist); float* rcom = SM(rcom);

  // ---- 1. kinematics (lane = body, level by level) ----------------------------------
  const int b = lane;
  const bool isbody = b > 0 && b < nb;
  const int depth = isbody ? m.body_depth[b] : -1;
  const int jid = isbody ? m.body_jntadr[b] : -1;
  const int jtype = jid >= 0 ? m.jnt_type[jid] : -1;
  V3 anchor = v3(0, 0, 0), axis = v3(0, 0, 1), xip = v3(0, 0, 0);
  float ximat[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) ximat[i] = 0.f;
  for (int lv = 1; lv <= M.maxdepth; ++lv) {
    if (depth == lv) {
      int p = m.body_parentid[b];
      Q4 pq = ldq(xquat + 4 * p);
      V3 pos = ld3(xpos + 3 * p) + qrot(pq, ld3(m.body_pos[b]));
      Q4 quat = qmul(pq, ldq(m.body_quat[b]));
      if (jtype == JNT_FREE) {
        int qa = m.jnt_qposadr[jid];
        pos = ld3(qpos + qa);
        quat = qnormalize(ldq(qpos + qa + 3));
        stq(qpos + qa + 3, quat);
        anchor = pos;
        axis = v3(0, 0, 1);
      } else if (jtype == JNT_HINGE || jtype == JNT_SLIDE) {
        int qa = m.jnt_qposadr[jid];
        V3 jp = ld3(m.jnt_pos[jid]), ja = ld3(m.jnt_axis[jid]);
        anchor = qrot(quat, jp) + pos;
        axis = qrot(quat, ja);
        float dq = qpos[qa] - m.qpos0[qa];
        if (jtype == JNT_HINGE) {
          quat = qmul(quat, axisangle(ja, dq));
          pos = anchor - qrot(quat, jp);
        } else {
          pos = pos + axis * dq;
        }
      }
      quat = qnormalize(quat);
      st3(xpos + 3 * b, pos);
      stq(xquat + 4 * b, quat);
      float R[9];
      qmat(quat, R);
#pragma unroll
      for (int i = 0; i < 9; ++i) xmat[9 * b + i] = R[i];
      xip = pos + qrot(quat, ld3(m.body_ipos[b]));
      st3(xipos + 3 * b, xip);
      qmat(qmul(quat, ldq(m.body_iquat[b])), ximat);
    }
    syncwarp();
  }

  // ---- 2. subtree COM of each tree root --------------------------------------------
  const int myroot = isbody ? M.body_rootidx[b] : -1;
  V3 com = v3(0, 0, 0);
  for (int r = 0; r < M.nroot; ++r) {
    float mass = (myroot == r) ? m.body_mass[b] : 0.f;
    float sx = mass * xip.x, sy = mass * xip.y, sz = mass * xip.z;
    warp_sum3(sx, sy, sz);
    V3 c = v3(sx, sy, sz) * M.root_invmass[r];
    if (lane == 0) st3(rcom + 3 * r, c);
    if (myroot == r) com = c;
  }

  // ---- 3. cinert (lane = body) and cdof (lane = jointed body) -------------------------
  if (isbody) {
    V3 off = xip - com;
    float mass = m.body_mass[b];
    const float* di = m.body_inertia[b];
    float I[6];  // xx yy zz xy xz yz of  R diag R^T
    I[0] = ximat[0] * ximat[0] * di[0] + ximat[1] * ximat[1] * di[1] + ximat[2] * ximat[2] * di[2];
    I[1] = ximat[3] * ximat[3] * di[0] + ximat[4] * ximat[4] * di[1] + ximat[5] * ximat[5] * di[2];
    I[2] = ximat[6] * ximat[6] * di[0] + ximat[7] * ximat[7] * di[1] + ximat[8] * ximat[8] * di[2];
    I[3] = ximat[0] * ximat[3] * di[0] + ximat[1] * ximat[4] * di[1] + ximat[2] * ximat[5] * di[2];
    I[4] = ximat[0] * ximat[6] * di[0] + ximat[1] * ximat[7] * di[1] + ximat[2] * ximat[8] * di[2];
    I[5] = ximat[3] * ximat[6] * di[0] + ximat[4] * ximat[7] * di[1] + ximat[5] * ximat[8] * di[2];
    float o2 = dot(off, off);
    float* ci = cinert + 10 * b;
    ci[0] = I[0] + mass * (o2 - off.x * off.x);
    ci[1] = I[1] + mass * (o2 - off.y * off.y);
    ci[2] = I[2] + mass * (o2 - off.z * off.z);
    ci[3] = I[3] - mass * off.x * off.y;
    ci[4] = I[4] - mass * off.x * off.z;
    ci[5] = I[5] - mass * off.y * off.z;
    ci[6] = mass * off.x; ci[7] = mass * off.y; ci[8] = mass * off.z; ci[9] = mass;
    if (jid >= 0) {
      int d = m.jnt_dofadr[jid];
      V3 offset = com - anchor;
      if (jtype == JNT_FREE) {
        for (int i = 0; i < 3; ++i) {
          float* c0 = cdof + 6 * (d + i);
          c0[0] = c0[1] = c0[2] = 0.f;
          c0[3] = i == 0 ? 1.f : 0.f; c0[4] = i == 1 ? 1.f : 0.f; c0[5] = i == 2 ? 1.f : 0.f;
          V3 ax = v3(xmat[9 * b + i], xmat[9 * b + 3 + i], xmat[9 * b + 6 + i]);
          float* c1 = cdof + 6 * (d + 3 + i);
          st3(c1, ax);
          st3(c1 + 3, cross(ax, offset));
        }
      } else if (jtype == JNT_HINGE) {
        st3(cdof + 6 * d, axis);
        st3(cdof + 6 * d + 3, cross(axis, offset));
      } else {
        float* c0 = cdof + 6 * d;
        c0[0] = c0[1] = c0[2] = 0.f;
        st3(c0 + 3, axis);
      }
    }
  }
  syncwarp();

  // ---- 4. velocities / accelerations down the tree, local RNE force ---------------------
  for (int lv = 1; lv <= M.maxdepth; ++lv) {
    if (depth == lv) {
      int p = m.body_parentid[b];
      float cv[6], ca[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) { cv[i] = cvel[6 * p + i]; ca[i] = cacc[6 * p + i]; }
      if (jtype == JNT_FREE) {
        int d = m.jnt_dofadr[jid];
        for (int k = 0; k < 3; ++k) {
          float qd = qvel[d + k];
#pragma unroll
          for (int i = 0; i < 6; ++i) { cv[i] += cdof[6 * (d + k) + i] * qd; cdofdot[6 * (d + k) + i] = 0.f; }
        }
        float cvt[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) cvt[i] = cv[i];
        for (int k = 3; k < 6; ++k) {
          float dd[6];
          mcross(cvt, cdof + 6 * (d + k), dd);
          float qd = qvel[d + k];
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            cdofdot[6 * (d + k) + i] = dd[i];
            cv[i] += cdof[6 * (d + k) + i] * qd;
            ca[i] += dd[i] * qd;
          }
        }
      } else if (jid >= 0) {
        int d = m.jnt_dofadr[jid];
        float dd[6];
        mcross(cv, cdof + 6 * d, dd);
        float qd = qvel[d];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          cdofdot[6 * d + i] = dd[i];
          cv[i] += cdof[6 * d + i] * qd;
          ca[i] += dd[i] * qd;
        }
      }
      float f1[6], f2[6], f3[6];
      inert_mul(cinert + 10 * b, ca, f1);
      inert_mul(cinert + 10 * b, cv, f2);
      mcross_force(cv, f2, f3);
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        cvel[6 * b + i] = cv[i];
        cacc[6 * b + i] = ca[i];
        cfrc[6 * b + i] = f1[i] + f3[i];
      }
    }
    syncwarp();
  }

  // ---- 5. composite inertia and RNE force up the tree (parents pull children) ----------
  for (int lv = M.maxdepth - 1; lv >= 1; --lv) {
    if (depth == lv) {
      for (int ci = M.child_adr[b]; ci < M.child_adr[b] + M.child_num[b]; ++ci) {
        int c = M.child_ids[ci];
#pragma unroll
        for (int i = 0; i < 10; ++i) cinert[10 * b + i] += cinert[10 * c + i];
#pragma unroll
        for (int i = 0; i < 6; ++i) cfrc[6 * b + i] += cfrc[6 * c + i];
      }
    }
    syncwarp();
  }

  // ---- 6. mass-matrix rows, bias, smooth force (lane = dof) ---------------------------
  Solver S;
  S.qfs = 0.f; S.qas = 0.f;
  S.l_sign = 0.f; S.l_D = 0.f; S.l_aref = 0.f; S.l_Jaref = 0.f;
  S.e_D = 0.f; S.e_aref = 0.f; S.e_Jaref = 0.f;
  S.qacc = S.Ma = S.grad = S.Mgrad = S.search = 0.f;
  S.gauss = S.cost = S.prev_cost = 0.f;
  const int d = lane;
  const bool isdof = d < nv;
  float myqvel = isdof ? qvel[d] : 0.f;
  if (isdof) {
    int bi = m.dof_bodyid[d];
    float f[6];
    inert_mul(cinert + 10 * bi, cdof + 6 * d, f);
    for (int j = 0; j <= d; ++j) Mm[d * nv + j] = 0.f;
#pragma unroll
    for (int c = 0; c < DIAL_MAXCHAIN; ++c)
      if (c < w.nchain) Mm[d * nv + w.chain[c]] = dot6(f, cdof + 6 * w.chain[c]);
    Mm[d * nv + d] += m.dof_armature[d];
    float bias = dot6(cdof + 6 * d, cfrc + 6 * bi);
    float act = 0.f;
    int a = M.dof_actuator[d];
    if (a >= 0) {
      float c = ctrl[a];
      if (m.actuator_ctrllimited[a]) c = fminf(fmaxf(c, m.actuator_ctrlrange[a][0]), m.actuator_ctrlrange[a][1]);
      float force = m.actuator_gain[a] * c + m.actuator_bias[a][0]
                  + m.actuator_bias[a][1] * qpos[m.actuator_qposadr[a]] + m.actuator_bias[a][2] * myqvel;
      if (m.actuator_forcelimited[a]) force = fminf(fmaxf(force, m.actuator_forcerange[a][0]), m.actuator_forcerange[a][1]);
      act = force * m.actuator_gear[a];
    }
    S.qfs = -m.dof_damping[d] * myqvel - bias + act;
  }
  syncwarp();
  // qacc_smooth = M^-1 qfrc_smooth
  {
    float* L = SM(L);
    if (isdof) {
#pragma unroll
      for (int c = 0; c < DIAL_MAXCHAIN; ++c)
        if (c < w.nchain) L[d * nv + w.chain[c]] = Mm[d * nv + w.chain[c]];
    }
    syncwarp();
    factor_LTL(w);
    S.qas = solve_LTL(w, S.qfs);
  }

  // ---- 7. collision (lane = contact) ---------------------------------------------------
  if (lane < m.ncon) {
    int k = M.con_pair[lane];
    int g1 = m.pair_geom1[k], g2 = m.pair_geom2[k];
    int b1 = m.geom_bodyid[g1], b2 = m.geom_bodyid[g2];
    // plane (geom1): normal = z axis of its world frame
    float R1[9], Rg[9];
    qmat(ldq(m.geom_quat[g1]), Rg);
    const float* X1 = xmat + 9 * b1;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) R1[3 * r + c] = X1[3 * r] * Rg[c] + X1[3 * r + 1] * Rg[3 + c] + X1[3 * r + 2] * Rg[6 + c];
    V3 n = v3(R1[2], R1[5], R1[8]);
    V3 gp1 = v3(X1[0] * m.geom_pos[g1][0] + X1[1] * m.geom_pos[g1][1] + X1[2] * m.geom_pos[g1][2],
                X1[3] * m.geom_pos[g1][0] + X1[4] * m.geom_pos[g1][1] + X1[5] * m.geom_pos[g1][2],
                X1[6] * m.geom_pos[g1][0] + X1[7] * m.geom_pos[g1][1] + X1[8] * m.geom_pos[g1][2]) + ld3(xpos + 3 * b1);
    const float* X2 = xmat + 9 * b2;
    V3 gp2 = v3(X2[0] * m.geom_pos[g2][0] + X2[1] * m.geom_pos[g2][1] + X2[2] * m.geom_pos[g2][2],
                X2[3] * m.geom_pos[g2][0] + X2[4] * m.geom_pos[g2][1] + X2[5] * m.geom_pos[g2][2],
                X2[6] * m.geom_pos[g2][0] + X2[7] * m.geom_pos[g2][1] + X2[8] * m.geom_pos[g2][2]) + ld3(xpos + 3 * b2);
    float radius = m.geom_size[g2][0];
    V3 center = gp2;
    V3 t1;
    V3 ydir = v3(0, 1, 0), zdir = v3(0, 0, 1);
    V3 alt = (n.y > -0.5f && n.y < 0.5f) ? ydir : zdir;
    if (m.pair_kind[k] == PAIR_PLANE_CAPSULE) {
      float Rg2[9];
      qmat(ldq(m.geom_quat[g2]), Rg2);
      V3 ax = v3(X2[0] * Rg2[2] + X2[1] * Rg2[5] + X2[2] * Rg2[8],
                 X2[3] * Rg2[2] + X2[4] * Rg2[5] + X2[5] * Rg2[8],
                 X2[6] * Rg2[2] + X2[7] * Rg2[5] + X2[8] * Rg2[8]);
      V3 bv = ax - n * dot(n, ax);
      float bn = sqrtf(dot(bv, bv));
      V3 bd = bv * (1.f / (bn + 1e-6f * (bn == 0.f ? 1.f : 0.f)));
      t1 = bn < 0.5f ? alt : bd;
      float sgn = M.con_sub[lane] == 0 ? 1.f : -1.f;
      center = gp2 + ax * (sgn * m.geom_size[g2][1]);
    } else {
      // make_frame(n)
      V3 bv = alt - n * dot(n, alt);
      float bn = sqrtf(dot(bv, bv));
      t1 = bv * (1.f / (bn + 1e-6f * (bn == 0.f ? 1.f : 0.f)));
    }
    float dist = dot(center - gp1, n) - radius;
    V3 p = center - n * (radius + 0.5f * dist);
    cdist[lane] = dist;
    st3(cpos + 3 * lane, p);
    st3(cframe + 9 * lane, n);
    st3(cframe + 9 * lane + 3, t1);
    st3(cframe + 9 * lane + 6, cross(n, t1));
  }
  syncwarp();

  // ---- 8. constraint rows ----------------------------------------------------------------
  // contact Jacobian columns (lane = dof): J[4c+e][d]
  if (isdof) {
    V3 ca_ = ld3(cdof + 6 * d), cl_ = ld3(cdof + 6 * d + 3);
    int rootb = M.body_rootidx[m.dof_bodyid[d]];
    V3 rc = ld3(rcom + 3 * rootb);
    for (int c = 0; c < m.ncon; ++c) {
      int k = M.con_pair[c];
      int b1 = m.geom_bodyid[m.pair_geom1[k]], b2 = m.geom_bodyid[m.pair_geom2[k]];
      float sgn = (float)((M.body_dofmask[b2] >> d) & 1u) - (float)((M.body_dofmask[b1] >> d) & 1u);
      float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
      if (sgn != 0.f && cdist[c] - (m.pair_margin[k] - m.pair_gap[k]) < 0.f) {
        V3 p = ld3(cpos + 3 * c);
        V3 jp = (cl_ + cross(ca_, p - rc)) * sgn;
        float jn = dot(ld3(cframe + 9 * c), jp), j1 = dot(ld3(cframe + 9 * c + 3), jp), j2 = dot(ld3(cframe + 9 * c + 6), jp);
        float mu0 = m.pair_friction[k][0], mu1 = m.pair_friction[k][1];
        e0 = jn + mu0 * j1; e1 = jn - mu0 * j1; e2 = jn + mu1 * j2; e3 = jn - mu1 * j2;
      }
      J[(4 * c + 0) * nv + d] = e0; J[(4 * c + 1) * nv + d] = e1;
      J[(4 * c + 2) * nv + d] = e2; J[(4 * c + 3) * nv + d] = e3;
    }
    // joint-limit row of this dof
    int lj = M.dof_limited[d];
    if (lj >= 0) {
      float qv = qpos[m.jnt_qposadr[lj]];
      float dmin = qv - m.jnt_range[lj][0], dmax = m.jnt_range[lj][1] - qv;
      float pos = fminf(dmin, dmax) - m.jnt_margin[lj];
      if (pos < 0.f) {
        float sign = dmin < dmax ? 1.f : -1.f;
        float k_, b_, imp;
        kbi(m.timestep, m.jnt_solref[lj], m.jnt_solimp[lj], pos, k_, b_, imp);
        float R = fmaxf(m.dof_invweight0[d] * (1.f - imp) / imp, DIAL_MINVAL);
        S.l_sign = sign;
        S.l_D = 1.f / R;
        S.l_aref = -b_ * (sign * myqvel) - k_ * imp * pos;
      }
    }
  }
  // contact edge rows (lane = edge)
  float ejv = mul_J(w, myqvel);
  if (lane < M.nedge) {
    int c = lane >> 2;
    int k = M.con_pair[c];
    float pos = cdist[c] - (m.pair_margin[k] - m.pair_gap[k]);
    if (pos < 0.f) {
      int b1 = m.geom_bodyid[m.pair_geom1[k]], b2 = m.geom_bodyid[m.pair_geom2[k]];
      float mu = m.pair_friction[k][0];
      float t = m.body_invweight0[b1] + m.body_invweight0[b2];
      float iw = (t + mu * mu * t) * 2.f * mu * mu / m.impratio;
      float k_, b_, imp;
      kbi(m.timestep, m.pair_solref[k], m.pair_solimp[k], pos, k_, b_, imp);
      float R = fmaxf(iw * (1.f - imp) / imp, DIAL_MINVAL);
      S.e_D = 1.f / R;
      S.e_aref = -b_ * ejv - k_ * imp * pos;
    }
  }

  // ---- 9. Newton solve --------------------------------------------------------------------
  float mywarm = isdof ? warm[d] : 0.f;
  float qacc = newton_solve(w, S, mywarm);

  // ---- 10. semi-implicit Euler -------------------------------------------------------------
  syncwarp();
  if (isdof) {
    warm[d] = qacc;
    if (integrate) qvel[d] = myqvel + m.timestep * qacc;
  }
  syncwarp();
  if (integrate && isbody && jid >= 0) {
    int qa = m.jnt_qposadr[jid], dd = m.jnt_dofadr[jid];
    if (jtype == JNT_FREE) {
      for (int i = 0; i < 3; ++i) qpos[qa + i] += m.timestep * qvel[dd + i];
      V3 wv = ld3(qvel + dd + 3);
      float nrm = sqrtf(dot(wv, wv));
      V3 ax = wv * (1.f / (nrm + 1e-6f * (nrm == 0.f ? 1.f : 0.f)));
      Q4 qn = qnormalize(qmul(ldq(qpos + qa + 3), axisangle(ax, m.timestep * nrm)));
      stq(qpos + qa + 3, qn);
    } else {
      qpos[qa] += m.timestep * qvel[dd];
    }
  }
  syncwarp();
}

// ---------------------------------------------------------------------------------
// rewards (computed by lane 0 from the slab; kinematics are those of the last forward
// pass, q/qvel are post-integration — the reference's staleness)
// ---------------------------------------------------------------------------------
DEV float foot_step(float duty, float cadence, float amplitude, float phase, float time) {
  const float PI = 3.14159265358979f, TWO_PI = 6.28318530717959f;
  float t = time * TWO_PI * cadence + PI;
  float a = t + PI - TWO_PI * phase;
  float angle = a - floorf(a / TWO_PI) * TWO_PI - PI;
  if (duty < 1.f) angle *= 0.5f / (1.f - duty);
  float cl = fminf(fmaxf(angle, -0.5f * PI), 0.5f * PI);
  float value = duty < 1.f ? cosf(cl) : 0.f;
  float fin = fabsf(value) >= 1e-6f ? fabsf(value) : 0.f;
  return amplitude * fin;
}

DEV float quat_yaw(Q4 q) {
  return atan2f(-2.f * q.x * q.y + 2.f * q.w * q.z, q.x * q.x + q.w * q.w - q.z * q.z - q.y * q.y);
}

struct BaseKin { V3 pos, vb, ab; Q4 rot; };

// Brax x / xd of body `bid` in the body frame (brax.mjx.pipeline, deploy/dial_plan.py:52-59)
DEV BaseKin base_kin(WarpCtx& w, int bid) {
  const DevModel& M = *w.M;
  BaseKin r;
  r.pos = ld3(SM(xpos) + 3 * bid);
  r.rot = ldq(SM(xquat) + 4 * bid);
  V3 ang = ld3(SM(cvel) + 6 * bid), lin = ld3(SM(cvel) + 6 * bid + 3);
  V3 off = r.pos - ld3(SM(rcom) + 3 * M.body_rootidx[bid]);
  V3 vel = lin - cross(off, ang);
  Q4 qc; qc.w = r.rot.w; qc.x = -r.rot.x; qc.y = -r.rot.y; qc.z = -r.rot.z;
  r.vb = qrot(qc, vel);
  r.ab = qrot(qc, ang * (3.14159265358979f / 180.f));
  return r;
}

DEV float reward_lane0(WarpCtx& w, int step, int& stage) {
  const DevModel& M = *w.M;
  const dial_plan_desc& c = w.P->c;
  const dial_model_desc& m = M.m;
  const float stepf = (float)step;
  float rew = 0.f;
  Q4 rot0 = ldq(SM(xquat) + 4);  // x.rot[0]
  V3 up = qrot(rot0, v3(0, 0, 1));
  float r_upright = -(up.x * up.x + up.y * up.y + (up.z - 1.f) * (up.z - 1.f));
  BaseKin bk = base_kin(w, c.torso_body);
  if (c.env_id == DIAL_ENV_GO2_WALK || c.env_id == DIAL_ENV_H1_WALK) {
    float ramp = stepf * c.dt / c.ramp_up_time;
    float vtx = fminf(c.vel_cmd[0] * ramp, c.vel_cmd[0]), vty = fminf(c.vel_cmd[1] * ramp, c.vel_cmd[1]);
    float atz = fminf(c.ang_cmd[2] * ramp, c.ang_cmd[2]);
    float r_gaits = 0.f;
    for (int f = 0; f < c.nfeet; ++f) {
      float zt = foot_step(c.gait_duty, c.gait_cadence, c.gait_amplitude, c.gait_phase[f], stepf * c.dt);
      float z;
      if (c.env_id == DIAL_ENV_GO2_WALK) {
        int sid = c.feet_site[f], sb = m.site_bodyid[sid];
        const float* X = SM(xmat) + 9 * sb;
        z = SM(xpos)[3 * sb + 2] + X[6] * m.site_pos[sid][0] + X[7] * m.site_pos[sid][1] + X[8] * m.site_pos[sid][2];
        float e = (zt - z) / 0.05f;
        r_gaits -= e * e;
      } else {
        z = fminf(SM(cdist)[2 * f], SM(cdist)[2 * f + 1]);
        r_gaits -= (zt - z) * (zt - z);
      }
    }
    float yaw_tar = 0.f + atz * c.dt * stepf;
    float dyaw = quat_yaw(bk.rot) - yaw_tar;
    float wy = atan2f(sinf(dyaw), cosf(dyaw));
    float r_yaw = -wy * wy;
    float r_vel = -((bk.vb.x - vtx) * (bk.vb.x - vtx) + (bk.vb.y - vty) * (bk.vb.y - vty));
    float r_ang = -(bk.ab.z - atz) * (bk.ab.z - atz);
    float r_h = -(bk.pos.z - c.pos_tar[2]) * (bk.pos.z - c.pos_tar[2]);
    if (c.env_id == DIAL_ENV_GO2_WALK) {
      rew = 0.1f * r_gaits + 0.5f * r_upright + 0.3f * r_yaw + r_vel + r_ang + r_h;
    } else {
      float r_energy = 0.f;
      for (int a = 0; a < m.nu; ++a) { float e = SM(ctrl)[a] / c.joint_torque_range[a][1]; r_energy -= e * e; }
      rew = 5.f * r_gaits + 0.5f * r_upright + 0.1f * r_yaw + r_vel + r_ang + 0.5f * r_h + 0.01f * r_energy;
    }
  } else {  // DIAL_ENV_GO2_SEQJUMP
    V3 dp = bk.pos - ld3(c.pose_seq[stage]);
    float r_pos = -dot(dp, dp);
    float dy = quat_yaw(bk.rot) - c.yaw_seq[stage];
    float r_yaw = -dy * dy;
    float r_contact = 0.f, pen = 0.f;
    for (int i = 0; i < 4; ++i) {
      float dist = SM(cdist)[i];
      bool penal = dist <= 0.001f;
      float px = SM(cpos)[3 * i], py = SM(cpos)[3 * i + 1];
      for (int j = 0; j < c.n_stage; ++j) {
        float dx = px - c.contact_targets[j][i][0], dyy = py - c.contact_targets[j][i][1];
        bool cond = dx * dx + dyy * dyy <= c.contact_radius[j][i] * c.contact_radius[j][i];
        if (cond && j == stage) r_contact += fminf(fmaxf(1.f - dist, 0.f), 1.f);
        penal = penal && !cond;
      }
      pen += penal ? 1.f : 0.f;
    }
    rew = r_pos + r_upright + 0.3f * r_yaw + 0.1f * r_contact - 0.1f * pen + 10.f;
    int ns = (int)floorf((float)(step + 1) * c.dt / c.jump_dt);
    stage = ns < c.n_stage - 1 ? ns : c.n_stage - 1;
  }
  return rew;
}

// ---------------------------------------------------------------------------------
// the per-warp rollout: one sample row, H env steps
// ---------------------------------------------------------------------------------
DEV void rollout_warp(const DevModel* Mp, const DevPlan* Pp, float* slab, const RolloutArgs& A,
                      int row, int lane) {
  WarpCtx w;
  w.M = Mp; w.P = Pp; w.s = slab; w.lane = lane;
  const DevModel& M = *Mp;
  const dial_model_desc& m = M.m;
  const dial_plan_desc& c = Pp->c;
  const int nq = m.nq, nv = m.nv, nu = m.nu, nb = m.nbody;
  // ancestor chain of this lane's dof
  w.nchain = 0;
#pragma unroll
  for (int i = 0; i < DIAL_MAXCHAIN; ++i) w.chain[i] = 0;
  if (lane < nv) {
    int j = lane, n = 0;
#pragma unroll
    for (int i = 0; i < DIAL_MAXCHAIN; ++i) {
      if (j >= 0) { w.chain[i] = j; n = i + 1; j = M.m.dof_parentid[j]; }
    }
    w.nchain = n;
  }
  // initial state + world body constants
  for (int i = lane; i < nq; i += 32) SM(qpos)[i] = A.qpos0[i];
  for (int i = lane; i < nv; i += 32) { SM(qvel)[i] = A.qvel0[i]; SM(warm)[i] = A.warm0[i]; }
  if (lane == 0) {
    SM(xpos)[0] = SM(xpos)[1] = SM(xpos)[2] = 0.f;
    SM(xquat)[0] = 1.f; SM(xquat)[1] = SM(xquat)[2] = SM(xquat)[3] = 0.f;
    for (int i = 0; i < 9; ++i) SM(xmat)[i] = (i % 4 == 0) ? 1.f : 0.f;
    for (int i = 0; i < 3; ++i) SM(xipos)[i] = 0.f;
    for (int i = 0; i < 10; ++i) SM(cinert)[i] = 0.f;
    for (int i = 0; i < 6; ++i) { SM(cvel)[i] = 0.f; SM(cfrc)[i] = 0.f; }
    SM(cacc)[0] = SM(cacc)[1] = SM(cacc)[2] = 0.f;
    SM(cacc)[3] = -m.gravity[0]; SM(cacc)[4] = -m.gravity[1]; SM(cacc)[5] = -m.gravity[2];
  }
  syncwarp();

  if (A.mode == 2) {  // pipeline_init: forward only
    if (lane < nu) SM(ctrl)[lane] = 0.f;
    syncwarp();
    physics_step(w, false);
    if (A.qpos_out) for (int i = lane; i < nq; i += 32) A.qpos_out[i] = SM(qpos)[i];
    if (A.warm_out) for (int i = lane; i < nv; i += 32) A.warm_out[i] = SM(warm)[i];
    return;
  }

  // control knots of this sample (lane = actuator)
  const int Hn1 = c.Hnode + 1;
  float Y[DIAL_MAXNODE];
#pragma unroll
  for (int k = 0; k < DIAL_MAXNODE; ++k) Y[k] = 0.f;
  if (A.mode == 1 && lane < nu) {
    const bool is_mean = row == c.Nsample;
    const uint32_t gidx = (uint32_t)(c.shard_offset + row);
    const uint32_t ntot = (uint32_t)c.Ntotal * (uint32_t)Hn1 * (uint32_t)nu;
#pragma unroll
    for (int k = 0; k < DIAL_MAXNODE; ++k) {
      if (k < Hn1) {
        float yb = A.Ybar[k * nu + lane];
        float y = yb;
        if (!is_mean && k > 0) {
          uint32_t idx = (gidx * (uint32_t)Hn1 + (uint32_t)k) * (uint32_t)nu + (uint32_t)lane;
          float e = A.eps ? A.eps[idx] : jax_normal_legacy(A.key0, A.key1, idx, ntot);
          y = e * A.noise[k] + yb;
        }
        Y[k] = fminf(fmaxf(y, -1.f), 1.f);
      }
    }
  }

  int step = A.step0, stage = A.stage0;
  float rsum = 0.f;
  for (int t = 0; t < A.H; ++t) {
    // action -> joint target -> torque (base_env.py:37-66)
    if (lane < nu) {
      float u;
      if (A.mode == 0) {
        u = A.us[((size_t)row * A.H + t) * nu + lane];
      } else {
        u = 0.f;
#pragma unroll
        for (int k = 0; k < DIAL_MAXNODE; ++k)
          if (k < Hn1) u += c.M_n2u[t][k] * Y[k];
      }
      float an = (u * c.action_scale + 1.f) * 0.5f;
      float jt = c.joint_range[lane][0] + an * (c.joint_range[lane][1] - c.joint_range[lane][0]);
      jt = fminf(fmaxf(jt, c.physical_joint_range[lane][0]), c.physical_joint_range[lane][1]);
      float ctrl = jt;
      if (c.leg_control_torque) {
        float tau = c.kp[lane] * (jt - SM(qpos)[7 + lane]) - c.kd[lane] * SM(qvel)[6 + lane];
        ctrl = fminf(fmaxf(tau, c.joint_torque_range[lane][0]), c.joint_torque_range[lane][1]);
      }
      SM(ctrl)[lane] = ctrl;
    }
    syncwarp();
    for (int f = 0; f < c.n_frames; ++f) physics_step(w, true);
    float rew = 0.f;
    if (lane == 0) rew = reward_lane0(w, step, stage);
    stage = shfl_i(stage, 0);
    step += 1;
    rsum += rew;
    // per-step outputs (coalesced: consecutive lanes -> consecutive addresses)
    size_t rt = (size_t)row * A.H + t;
    if (A.rewss && lane == 0) A.rewss[rt] = rew;
    if (A.q) for (int i = lane; i < nq; i += 32) A.q[rt * nq + i] = SM(qpos)[i];
    if (A.qd) for (int i = lane; i < nv; i += 32) A.qd[rt * nv + i] = SM(qvel)[i];
    if (A.xpos) for (int i = lane; i < 3 * (nb - 1); i += 32) A.xpos[rt * 3 * (nb - 1) + i] = SM(xpos)[3 + i];
  }
  if (A.rews && lane == 0) A.rews[row] = rsum / (float)A.H;
  if (row == 0) {
    if (A.qpos_out) for (int i = lane; i < nq; i += 32) A.qpos_out[i] = SM(qpos)[i];
    if (A.qvel_out) for (int i = lane; i < nv; i += 32) A.qvel_out[i] = SM(qvel)[i];
    if (A.warm_out) for (int i = lane; i < nv; i += 32) A.warm_out[i] = SM(warm)[i];
    if (A.ctrl_out) for (int i = lane; i < nu; i += 32) A.ctrl_out[i] = SM(ctrl)[i];
  }
}
