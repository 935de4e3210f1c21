// dial_kernels.cu — sm_100a kernels + the C ABI of include/dial_b200.h.
//
// Kernels
//   rollout_kernel<WPC>   one warp per sample row, persistent over the horizon; the
//                         compiled model + plan constants are staged into shared memory
//                         with one TMA bulk copy (cp.async.bulk + mbarrier) per CTA.
//   weights_kernel        population std + max-subtracted softmax over all rewards
//                         (warp-shuffle + one smem stage reductions), single CTA.
//   ybar_kernel           Ybar = sum_n w_n Y0s_n with Y0s regenerated from eps / Threefry;
//                         per-CTA partials, last CTA reduces in fixed order (deterministic).
//   trajbar_kernel        qbar / qdbar / xbar weighted sums over the stored trajectories.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include <new>
#include "dial_host.h"
#define DIAL_STR2(x) #x
#define DIAL_STR(x) DIAL_STR2(x)

static thread_local std::string g_err;
static int fail(const std::string& s) { g_err = s; return -1; }
#define CUDA_OK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

// The rollout kernel lives in dial_rollout_variant.cu, compiled once per solver variant
// (-DDIAL_VARIANT=v) so that the five instantiations build in parallel; each object exports
// one launcher.  Custom-reward builds (-DDIAL_ONLY_VARIANT=v) compile a single translation unit.
#ifdef DIAL_ONLY_VARIANT
#define DIAL_HAS_VARIANT(v) ((v) == DIAL_ONLY_VARIANT)
#define DIAL_VARIANT DIAL_ONLY_VARIANT
#include "dial_rollout_variant.cu"
#else
#define DIAL_HAS_VARIANT(v) 1
#endif
#define DIAL_DECL_LAUNCH(v) cudaError_t dial_launch_rollout_v##v(const DevModel*, const DevPlan*, const RolloutArgs&, int, int, size_t, cudaStream_t);
#if DIAL_HAS_VARIANT(0)
DIAL_DECL_LAUNCH(0)
#endif
#if DIAL_HAS_VARIANT(1)
DIAL_DECL_LAUNCH(1)
#endif
#if DIAL_HAS_VARIANT(2)
DIAL_DECL_LAUNCH(2)
#endif
#if DIAL_HAS_VARIANT(3)
DIAL_DECL_LAUNCH(3)
#endif
#if DIAL_HAS_VARIANT(4)
DIAL_DECL_LAUNCH(4)
#endif

// ---------------------------------------------------------------------------------
// softmax weights over all rewards (core/dial_core.py:125-128), single CTA
// ---------------------------------------------------------------------------------
// block-wide sum of K values per thread (+ optionally the max of one): shuffle tree, one smem
// stage, result identical in every thread
template <int K>
__device__ __forceinline__ void block_reduce(float (&v)[K], float* mx, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if (mx) *mx = fmaxf(*mx, __shfl_xor_sync(0xffffffffu, *mx, o));
  }
  __syncthreads();   // previous use of `red` is over
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) red[k * 32 + wid] = v[k];
    if (mx) red[K * 32 + wid] = *mx;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = lane < nw ? red[k * 32 + lane] : 0.f;
  if (mx) *mx = lane < nw ? red[K * 32 + lane] : -INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if (mx) *mx = fmaxf(*mx, __shfl_xor_sync(0xffffffffu, *mx, o));
  }
}

// rews [n] (mean sample last) -> weights [n] = softmax((rews - rews[n-1]) / std(rews) / temp)
// (core/dial_core.py:125-128).  Three passes: statistics of d = r - rbar (count, sum, sum of
// squares, max: the shift keeps the one-pass variance accurate), exp + normaliser, scale.
// Deviations from the reference, which has no guards:
//   * non-finite rewards (diverged samples) get weight 0 and are left out of the statistics;
//   * std == 0 (all finite rewards equal, e.g. zero noise): uniform weights over the finite
//     samples instead of 0/0 = NaN;
//   * no finite reward at all: the whole weight goes to the mean sample (Ybar is kept);
//   * a non-finite rbar only changes the reference point of the shift (softmax is shift-invariant).
// Multi-GPU consumer side of the reward exchange (dial_exchange_*): wait until every rank's flag
// in the local mailbox carries the current sequence number.  Bounded spin (~4 s of %globaltimer):
// a peer that never arrives sets *err instead of hanging the GPU.
struct XchWait {
  const float* mbox;            // local mailbox [2][n]; null: no exchange, `rews` is used as given
  const uint32_t* flags;        // local flags [2][DIAL_MAXRANK]
  uint32_t* seq;                // local sequence number, bumped when the weights are done
  uint32_t* err;                // local error word (1: timeout)
  float* rews_copy;             // optional compact copy of the gathered rewards [n]
  int world;
};
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// `acc` (nullable): running total of the time this rank spent waiting for its slowest peer, in ns
// (dial_exchange_status words 4 / 5) — the part of a sharded step that is skew + NVLink latency.
__device__ __forceinline__ void xch_wait_flags(const uint32_t* flags, uint32_t want, int world, uint32_t* err,
                                               uint32_t* acc = nullptr) {
  if ((int)threadIdx.x < world) {
    const volatile uint32_t* f = flags + threadIdx.x;
    const unsigned long long t0 = globaltimer_ns();
    while (*f < want) {
      if (globaltimer_ns() - t0 > 4000000000ull) { *err = 1u; break; }
      __nanosleep(100);
    }
    if (acc) {
      const unsigned long long dt = globaltimer_ns() - t0;
      const unsigned m = __reduce_max_sync(__activemask(), (unsigned)(dt > 0xffffffffull ? 0xffffffffull : dt));
      if (threadIdx.x == 0) *acc += m;
    }
    __threadfence_system();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(1024) weights_kernel(const float* __restrict__ rews, int n, float temp,
                                                        float* __restrict__ weights, const XchWait X) {
  __shared__ float red[4 * 32];
  const int tid = threadIdx.x;
  if (X.mbox) {
    const uint32_t seq = *X.seq, buf = seq & 1u;
    xch_wait_flags(X.flags + buf * DIAL_MAXRANK, seq + 1u, X.world, X.err, X.err + 2);
    rews = X.mbox + (size_t)buf * n;
    if (X.rews_copy)
      for (int i = tid; i < n; i += blockDim.x) X.rews_copy[i] = __ldcg(rews + i);
  }
  float rbar = __ldcg(rews + n - 1);
  if (!isfinite(rbar)) rbar = 0.f;
  float st[3] = {0.f, 0.f, 0.f}, dmax = -INFINITY;
  for (int i = tid; i < n; i += blockDim.x) {
    const float r = __ldcg(rews + i);
    if (isfinite(r)) { const float d = r - rbar; st[0] += 1.f; st[1] += d; st[2] += d * d; dmax = fmaxf(dmax, d); }
  }
  block_reduce<3>(st, &dmax, red);
  const float cnt = st[0];
  if (cnt == 0.f) {
    for (int i = tid; i < n; i += blockDim.x) weights[i] = (i == n - 1) ? 1.f : 0.f;
    if (X.mbox && tid == 0) *X.seq = *X.seq + 1u;
    return;
  }
  const float mean = st[1] / cnt;
  const float sd = sqrtf(fmaxf(st[2] / cnt - mean * mean, 0.f));
  const bool flat = !(sd > 0.f);
  const float inv = flat ? 0.f : 1.f / sd / temp;   // flat: every logit 0 -> uniform weights
  const float mx = dmax * inv;
  float z[1] = {0.f};
  for (int i = tid; i < n; i += blockDim.x) {
    const float r = __ldcg(rews + i);
    const float e = isfinite(r) ? expf((r - rbar) * inv - mx) : 0.f;
    weights[i] = e;
    z[0] += e;
  }
  block_reduce<1>(z, nullptr, red);
  const float iz = 1.f / z[0];
  for (int i = tid; i < n; i += blockDim.x) weights[i] *= iz;
  if (X.mbox && tid == 0) *X.seq = *X.seq + 1u;   // the next reverse_once uses the other mailbox half
}

// Sum of the per-rank partial bars (qbar|qdbar|xbar, core/dial_core.py:133-135) over NVLink peer
// memory: push my partial into slot [rank] of every rank's bars mailbox, raise my flag there,
// wait for all flags here, add the slots in rank order (bitwise identical on every rank).
struct BarsXch {
  float* mbox[DIAL_MAXRANK];        // bars mailbox of rank p: [2][DIAL_MAXRANK][nbar]
  uint32_t* flags[DIAL_MAXRANK];    // bars flags of rank p:   [2][DIAL_MAXRANK]
  uint32_t* seq;                    // local bars sequence number
  uint32_t* err;
  int world, rank, nbar;
  const float* partial;             // local partial [nbar]
  float* out[3];
  int n0, n1;                       // nbar = n0 (q) + n1 (qd) + rest (x)
};
__global__ void __launch_bounds__(1024) bars_allreduce_kernel(const BarsXch B) {
  const uint32_t seq = *B.seq, buf = seq & 1u;
  const size_t half = (size_t)buf * DIAL_MAXRANK * B.nbar;
  for (int p = 0; p < B.world; ++p) {
    float* dst = B.mbox[p] + half + (size_t)B.rank * B.nbar;
    for (int i = threadIdx.x; i < B.nbar; i += blockDim.x) dst[i] = B.partial[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    for (int p = 0; p < B.world; ++p)
      *reinterpret_cast<volatile uint32_t*>(B.flags[p] + buf * DIAL_MAXRANK + B.rank) = seq + 1u;
  }
  __syncthreads();
  xch_wait_flags(B.flags[B.rank] + buf * DIAL_MAXRANK, seq + 1u, B.world, B.err, B.err + 3);
  const float* mine = B.mbox[B.rank] + half;
  for (int i = threadIdx.x; i < B.nbar; i += blockDim.x) {
    float s_ = 0.f;
    for (int p = 0; p < B.world; ++p) s_ += __ldcg(mine + (size_t)p * B.nbar + i);
    float* o = i < B.n0 ? B.out[0] + i : (i < B.n0 + B.n1 ? B.out[1] + (i - B.n0) : B.out[2] + (i - B.n0 - B.n1));
    *o = s_;
  }
  __syncthreads();
  if (threadIdx.x == 0) *B.seq = seq + 1u;
}

// sum_b col[b * stride], b = 0..count-1, in that order, with 16 independent L2 loads in flight (the
// last-CTA reductions of ybar_kernel / update_kernel: a plain loop pays one L2 round trip per term)
__device__ __forceinline__ float ordered_column_sum(const float* __restrict__ col, int stride, unsigned count, bool on) {
  float tot = 0.f;
  if (!on) return tot;
  unsigned b = 0;
  for (; b + 16 <= count; b += 16) {
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = __ldcg(col + (size_t)(b + k) * stride);
#pragma unroll
    for (int k = 0; k < 16; ++k) tot += v[k];
  }
  for (; b < count; ++b) tot += __ldcg(col + (size_t)b * stride);
  return tot;
}

// ---------------------------------------------------------------------------------
// Ybar = sum_n w_n * Y0s_n  (core/dial_core.py:129-132), Y0s regenerated
// ---------------------------------------------------------------------------------
#define YBAR_THREADS 256
__global__ void __launch_bounds__(YBAR_THREADS) ybar_kernel(const float* __restrict__ weights, const float* __restrict__ eps,
                                                             uint32_t key0, uint32_t key1, const float* __restrict__ Ybar,
                                                             const float* __restrict__ noise, int Ntotal, int Hn1, int nu,
                                                             float* __restrict__ partial, unsigned int* __restrict__ counter,
                                                             float* __restrict__ Ybar_out,
                                                             const uint32_t* __restrict__ key_dev) {
  if (key_dev) { key0 = key_dev[0]; key1 = key_dev[1]; }
  // thread -> (sample slot, output element); elements = Hn1*nu <= 160
  const int ne = Hn1 * nu;
  __shared__ float acc[YBAR_THREADS];
  __shared__ bool is_last;
  const int slots = YBAR_THREADS / ne;  // samples processed concurrently per CTA
  const int slot = threadIdx.x / ne, el = threadIdx.x - slot * ne;
  float a = 0.f;
  if (slot < slots) {
    const int k = el / nu;
    const float yb = Ybar[el], ns = noise[k];
    const uint32_t ntot = (uint32_t)Ntotal * (uint32_t)ne;
    for (int n = blockIdx.x * slots + slot; n <= Ntotal; n += gridDim.x * slots) {
      float y = yb;
      if (n < Ntotal && k > 0) {
        uint32_t idx = (uint32_t)n * (uint32_t)ne + (uint32_t)el;
        float e = eps ? eps[idx] : jax_normal_legacy(key0, key1, idx, ntot);
        y = e * ns + yb;
      }
      y = fminf(fmaxf(y, -1.f), 1.f);
      a += weights[n] * y;
    }
  }
  acc[threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.x < ne) {
    float s = 0.f;
    for (int sl = 0; sl < slots; ++sl) s += acc[sl * ne + threadIdx.x];
    partial[blockIdx.x * ne + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last) {
    if (threadIdx.x < ne) Ybar_out[threadIdx.x] = ordered_column_sum(partial + threadIdx.x, ne, gridDim.x, true);
    if (threadIdx.x == 0) *counter = 0u;
  }
}

// ---------------------------------------------------------------------------------
// Fused update of the control step graph: weights + Ybar + rng advance in ONE multi-CTA kernel
// (dial_core.py:106,125-132).  Every CTA recomputes the reward statistics (n <= 131072: up to 512
// L2-resident loads per thread at the 65536-sample config, a few at 2048), accumulates sum_n e_n Y0s_n and sum_n e_n over its share of the samples with
// e_n = exp((r_n - rbar) / std / temp - max); the last CTA adds the partials in fixed order,
// divides, normalises the stored weights, and advances the planner rng.  Two graph nodes fewer per
// reverse_once; measured against the three-kernel sequence it is a wash (3.073 vs 3.068 ms per
// control step at N = 2048: a launch boundary inside a graph costs about a microsecond).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(YBAR_THREADS) update_kernel(const float* __restrict__ rews, int n, float temp,
                                                               float* __restrict__ weights, const XchWait X,
                                                               uint32_t* __restrict__ rng, const float* __restrict__ Ybar,
                                                               const float* __restrict__ noise, int Ntotal, int Hn1, int nu,
                                                               float* __restrict__ partial, unsigned int* __restrict__ counter,
                                                               float* __restrict__ Ybar_out) {
  __shared__ float red[4 * 32];
  __shared__ float acc[YBAR_THREADS];
  __shared__ bool is_last;
  const int tid = threadIdx.x;
  if (X.mbox) {
    const uint32_t seq = *X.seq, buf = seq & 1u;
    xch_wait_flags(X.flags + buf * DIAL_MAXRANK, seq + 1u, X.world, X.err, blockIdx.x == 0 ? X.err + 2 : nullptr);
    rews = X.mbox + (size_t)buf * n;
  }
  uint32_t key0, key1;
  const uint32_t r0 = rng[0], r1 = rng[1];
  split_key(r0, r1, key0, key1);
  // ---- statistics of d = r - rbar over the finite rewards (see weights_kernel) --------------------
  float rbar = __ldcg(rews + n - 1);
  if (!isfinite(rbar)) rbar = 0.f;
  float st[3] = {0.f, 0.f, 0.f}, dmax = -INFINITY;
  for (int i = tid; i < n; i += blockDim.x) {
    const float r = __ldcg(rews + i);
    if (isfinite(r)) { const float d = r - rbar; st[0] += 1.f; st[1] += d; st[2] += d * d; dmax = fmaxf(dmax, d); }
  }
  block_reduce<3>(st, &dmax, red);
  const float cnt = st[0];
  const float mean = st[1] / fmaxf(cnt, 1.f);
  const float sd = sqrtf(fmaxf(st[2] / fmaxf(cnt, 1.f) - mean * mean, 0.f));
  const float inv = (sd > 0.f) ? 1.f / sd / temp : 0.f;
  const float mx = dmax * inv;
  // ---- this CTA's share of sum e_n Y0s_n (thread -> (sample slot, knot element)) and sum e_n ------------
  const int ne = Hn1 * nu;
  const int slots = YBAR_THREADS / ne;
  const int slot = tid / ne, el = tid - slot * ne;
  float a = 0.f, z = 0.f;
  if (slot < slots) {
    const int k = el / nu;
    const float yb = Ybar[el], ns = noise[k];
    const uint32_t ntot = (uint32_t)Ntotal * (uint32_t)ne;
    for (int s_ = blockIdx.x * slots + slot; s_ <= Ntotal; s_ += gridDim.x * slots) {
      const float r = __ldcg(rews + s_);
      float e = isfinite(r) ? expf((r - rbar) * inv - mx) : 0.f;
      if (cnt == 0.f) e = (s_ == Ntotal) ? 1.f : 0.f;       // no finite reward: keep the mean sample
      float y = yb;
      if (s_ < Ntotal && k > 0) {
        const uint32_t idx = (uint32_t)s_ * (uint32_t)ne + (uint32_t)el;
        y = jax_normal_legacy(key0, key1, idx, ntot) * ns + yb;
      }
      y = fminf(fmaxf(y, -1.f), 1.f);
      a += e * y;
      if (el == 0) { z += e; weights[s_] = e; }
    }
  }
  acc[tid] = a;
  float zz[1] = {z};
  block_reduce<1>(zz, nullptr, red);
  __syncthreads();
  if (tid < ne) {
    float s_ = 0.f;
    for (int sl = 0; sl < slots; ++sl) s_ += acc[sl * ne + tid];
    partial[blockIdx.x * (ne + 1) + tid] = s_;
  }
  if (tid == 0) partial[blockIdx.x * (ne + 1) + ne] = zz[0];
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // one thread per column of the partials (columns 0..ne-1: sum e Y0s, column ne: sum e), one pass,
  // 16 L2 loads in flight, added in CTA order (bitwise deterministic)
  __shared__ float zsh;
  float tot = ordered_column_sum(partial + tid, ne + 1, gridDim.x, tid <= ne);
  if (tid == ne) zsh = tot;
  if (ne == (int)blockDim.x && tid == 0) zsh = ordered_column_sum(partial + ne, ne + 1, gridDim.x, true);
  __syncthreads();
  const float iz = 1.f / zsh;
  if (tid < ne) Ybar_out[tid] = tot * iz;
  for (int i = tid; i < n; i += blockDim.x) {
    weights[i] = __ldcg(weights + i) * iz;
    if (X.mbox && X.rews_copy) X.rews_copy[i] = __ldcg(rews + i);
  }
  if (tid == 0) {
    *counter = 0u;
    uint32_t n0, n1;
    split_rng(r0, r1, n0, n1);
    rng[0] = n0; rng[1] = n1;
    if (X.mbox) *X.seq = *X.seq + 1u;
  }
}

// ---------------------------------------------------------------------------------
// glue of the device-resident MPC loop (dial_mpc_step): everything the reference's Python loop
// does between kernels (core/dial_core.py:242-268) as tiny kernels, so that one MPC step is one
// CUDA graph with no host work inside
// ---------------------------------------------------------------------------------
// rng, key = jax.random.split(rng)   (dial_core.py:106), legacy layout as dial_key_split
__global__ void mpc_split_kernel(uint32_t* __restrict__ rng, uint32_t* __restrict__ key) {
  if (threadIdx.x == 0) {
    const uint32_t k0 = rng[0], k1 = rng[1];
    uint32_t a0 = 0, b0 = 2, a1 = 1, b1 = 3;
    threefry2x32(k0, k1, a0, b0);
    threefry2x32(k0, k1, a1, b1);
    rng[0] = a0; rng[1] = a1; key[0] = b0; key[1] = b1;
  }
}

// Y <- shift(Y) = u2node(roll(node2u(Y), -1), last row 0)  (dial_core.py:160-165) as one constant
// (Hn+1)x(Hn+1) matrix; one thread per output element
__global__ void mpc_shift_kernel(const float* __restrict__ Msh, const float* __restrict__ Yin,
                                 float* __restrict__ Yout, int n1, int nu) {
  const int i = threadIdx.x;
  if (i < n1 * nu) {
    const int k = i / nu, a = i - k * nu;
    float s = 0.f;
    for (int j = 0; j < n1; ++j) s += Msh[k * n1 + j] * Yin[j * nu + a];
    Yout[i] = s;
  }
}

// ---------------------------------------------------------------------------------
// qbar / qdbar / xbar  (core/dial_core.py:133-135): weighted sums over stored trajectories.
// A row of a trajectory array is H * ncol contiguous floats, so out[j] = sum_r w_r traj[r][j] with
// one thread per j reads whole 128-byte lines (consecutive lanes -> consecutive addresses).
//   stage 1: grid (ceil(max_j / 256), TB_CHUNKS row chunks, 3 arrays): per-chunk partials, rows in
//            order, 8 independent loads in flight per thread
//   stage 2: grid H: partials summed in fixed chunk order (bitwise deterministic)
// The only bandwidth-shaped kernel of the path: rows * H * (nq + nv + 3 (nbody-1)) * 4 bytes read
// once from L2 / HBM (cfg1: 16 MB, cfg4 shard: 65 MB).
// ---------------------------------------------------------------------------------
#define TB_CHUNKS 32
struct TrajArgs {
  const float* traj[3];
  float* out[3];
  int ncol[3], coloff[3];
  int coltot, nrows, H;
  const float* weights;
  int w_offset, mean_row, mean_weight_index, include_mean;
  float* partial;  // [TB_CHUNKS][H][coltot]
};

// (32 registers: one CTA fits beside the 448-thread rollout CTA of the next iteration, which the bars overlap)
__global__ void __launch_bounds__(256, 8) trajbar_partial_kernel(const TrajArgs T) {
  const int chunk = blockIdx.y, arr = blockIdx.z;
  const int ncol = arr == 0 ? T.ncol[0] : (arr == 1 ? T.ncol[1] : T.ncol[2]), len = T.H * ncol;
  const int coloff = arr == 0 ? T.coloff[0] : (arr == 1 ? T.coloff[1] : T.coloff[2]);
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (blockIdx.x * 256 >= len) return;
  const float* __restrict__ traj = arr == 0 ? T.traj[0] : (arr == 1 ? T.traj[1] : T.traj[2]);
  const int per = (T.nrows + TB_CHUNKS - 1) / TB_CHUNKS;
  const int r0 = chunk * per, r1 = min(T.nrows, r0 + per);
  const int jj = j < len ? j : len - 1;
  float a = 0.f;
  for (int r = r0; r < r1; r += 8) {
    float wv[8], xv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int rr = r + k;
      float wgt = 0.f;
      if (rr < r1) {
        if (rr == T.mean_row) wgt = T.include_mean ? T.weights[T.mean_weight_index] : 0.f;
        else wgt = T.weights[T.w_offset + rr];
      }
      wv[k] = wgt;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)   // weight 0: a diverged sample (NaN trajectory) or a row beyond the chunk: not read
      xv[k] = (wv[k] != 0.f) ? traj[(size_t)(r + k) * len + jj] : 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) a += wv[k] * xv[k];
  }
  if (j < len) {
    const int t = j / ncol, c = j - t * ncol;
    T.partial[((size_t)chunk * T.H + t) * T.coltot + coloff + c] = a;
  }
}

__global__ void __launch_bounds__(128, 8) trajbar_final_kernel(const TrajArgs T) {
  const int t = blockIdx.x;
  for (int col = threadIdx.x; col < T.coltot; col += blockDim.x) {
    float s = 0.f;
#pragma unroll 8
    for (int ch = 0; ch < TB_CHUNKS; ++ch) s += T.partial[((size_t)ch * T.H + t) * T.coltot + col];
    float* out; int ncol, off;   // (no dynamic indexing of the kernel parameter: it would be copied to the stack)
    if (col >= T.coloff[2]) { out = T.out[2]; ncol = T.ncol[2]; off = T.coloff[2]; }
    else if (col >= T.coloff[1]) { out = T.out[1]; ncol = T.ncol[1]; off = T.coloff[1]; }
    else { out = T.out[0]; ncol = T.ncol[0]; off = T.coloff[0]; }
    if (out) out[(size_t)t * ncol + (col - off)] = s;
  }
}

// ---------------------------------------------------------------------------------
// plan object
// ---------------------------------------------------------------------------------
struct dial_plan {
  DevModel hM;
  DevPlan hP;
  DevModel* dM = nullptr;
  DevPlan* dP = nullptr;
  int variant = 0;
  int num_sms = 148;
  size_t smem_bytes = 0;
  // workspaces
  // trajectory workspaces [Nsample+1, Hs+1, *], double-buffered so that the bars of iteration i
  // (side stream) can overlap the rollout of iteration i+1
  float *traj_q[2] = {nullptr, nullptr}, *traj_qd[2] = {nullptr, nullptr}, *traj_x[2] = {nullptr, nullptr};
  int cur = 0;
  float* weights = nullptr;                                        // [Ntotal+1]
  float* weights2 = nullptr;                                       // second buffer: the bars of iteration i overlap iteration i+1
  cudaStream_t side = nullptr;                                     // bars branch of the control-step graph
  cudaEvent_t ev_main[2] = {nullptr, nullptr}, ev_side[2] = {nullptr, nullptr};
  float* partial = nullptr;
  float* tb_partial = nullptr;
  unsigned int* counter = nullptr;
  unsigned int* row_counter = nullptr;
  float* zeros = nullptr;                                          // [nv]
  int ybar_grid = 0;
  int upd_grid = 0;             // grid of the fused update kernel: about 8 samples per thread slot
  int64_t launches = 0;
  float* dbg = nullptr;  // optional device counters (DIAL_DEBUG_COUNTERS=1)
  // device-resident MPC loop
  dial_mpc_buffers mpc{};       // caller-owned state block (dial_mpc_bind)
  bool mpc_bound = false;
  float* mpc_Msh = nullptr;     // [Hn+1][Hn+1] shift matrix
  float* mpc_Y1 = nullptr;      // ping-pong partner of mpc.Y
  uint32_t* mpc_key = nullptr;  // sampling key of the current reverse_once
  struct MpcGraph { int n_diffuse, env_step, seen; cudaGraphExec_t exec; int64_t launches; };
  std::vector<MpcGraph> mpc_graphs;
  // multi-GPU exchange over NVLink peer memory (dial_exchange_*): one cudaMalloc per rank, mapped
  // into every peer with CUDA IPC.  Word offsets inside the block are the same on every rank.
  struct Exchange {
    bool on = false;
    int rank = 0, world = 1, nbar = 0;
    uint32_t* base[DIAL_MAXRANK] = {nullptr};   // base[rank] = own allocation, others = IPC mappings
    size_t words = 0, o_mbox = 0, o_bars = 0, o_flags = 0, o_bflags = 0, o_local = 0;
    float* bars_partial = nullptr;              // local staging of this rank's partial bars [nbar]
    float* mbox(int p) const { return reinterpret_cast<float*>(base[p] + o_mbox); }
    float* bars(int p) const { return reinterpret_cast<float*>(base[p] + o_bars); }
    uint32_t* flags(int p) const { return base[p] + o_flags; }
    uint32_t* bflags(int p) const { return base[p] + o_bflags; }
    uint32_t* seq() const { return base[rank] + o_local; }
    unsigned int* done() const { return base[rank] + o_local + 1; }
    uint32_t* err() const { return base[rank] + o_local + 2; }
    uint32_t* bseq() const { return base[rank] + o_local + 3; }
  } xch;
};

static void fill_xch(const dial_plan* p, RolloutArgs& A) {
  if (!p->xch.on) return;
  A.xch_world = p->xch.world; A.xch_rank = p->xch.rank;
  for (int r = 0; r < p->xch.world; ++r) { A.xch_mbox[r] = p->xch.mbox(r); A.xch_flags[r] = p->xch.flags(r); }
  A.xch_seq = p->xch.seq(); A.xch_done = p->xch.done();
}
static XchWait xch_wait_args(const dial_plan* p, float* rews_copy) {
  XchWait X; memset(&X, 0, sizeof(X));
  if (p->xch.on) {
    X.mbox = p->xch.mbox(p->xch.rank); X.flags = p->xch.flags(p->xch.rank); X.seq = p->xch.seq(); X.err = p->xch.err();
    X.rews_copy = rews_copy; X.world = p->xch.world;
  }
  return X;
}

extern "C" int dial_abi_version(void) { return DIAL_ABI_VERSION; }
extern "C" const char* dial_last_error(void) { return g_err.c_str(); }
extern "C" size_t dial_sizeof(int which) {
  return which == 0 ? sizeof(dial_model_desc) : which == 1 ? sizeof(dial_plan_desc) : which == 2 ? sizeof(dial_state)
       : which == 3 ? sizeof(dial_mpc_buffers) : 0;
}

// solver instantiation by tree shape: star<3,6> (quadruped), star<5,7> (humanoid), star<5,6>,
// dense<22> (elliptic cones), generic tree.  `wpc` warps per CTA (1..16; one kernel serves all).
static cudaError_t launch_rollout(dial_plan* p, const RolloutArgs& A, int wpc, cudaStream_t st) {
  const size_t smem = sizeof(DevModel) + sizeof(DevPlan) + (size_t)wpc * p->hM.warp_floats * sizeof(float);
  int grid = (A.nrows + wpc - 1) / wpc;
  if (A.row_counter) {
    const int resident = p->num_sms * (wpc > 8 ? 1 : 16 / wpc);
    grid = grid < resident ? grid : resident;
    cudaError_t e = cudaMemsetAsync(A.row_counter, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return e;
  }
  p->launches++;
  switch (p->variant) {
#if DIAL_HAS_VARIANT(1)
    case 1: return dial_launch_rollout_v1(p->dM, p->dP, A, grid, wpc, smem, st);
#endif
#if DIAL_HAS_VARIANT(2)
    case 2: return dial_launch_rollout_v2(p->dM, p->dP, A, grid, wpc, smem, st);
#endif
#if DIAL_HAS_VARIANT(3)
    case 3: return dial_launch_rollout_v3(p->dM, p->dP, A, grid, wpc, smem, st);
#endif
#if DIAL_HAS_VARIANT(4)
    case 4: return dial_launch_rollout_v4(p->dM, p->dP, A, grid, wpc, smem, st);
#endif
#if DIAL_HAS_VARIANT(0)
    case 0: return dial_launch_rollout_v0(p->dM, p->dP, A, grid, wpc, smem, st);
#endif
    default: return cudaErrorInvalidDeviceFunction;
  }
}

// Launch shape.  One CTA per SM with up to 16 warps that re-converge at every env step
// ("lock-step"): the warps of an SM then walk the ~230 KB of straight-line kernel code together
// and share instruction fetches (measured 1.4x over independent 4-warp CTAs at N=2048, r1d).
static cudaError_t launch_rollout_any(dial_plan* p, const RolloutArgs& A0, cudaStream_t st) {
  RolloutArgs A = A0;
  const char* f = getenv("DIAL_WPC");
  int wpc = f ? atoi(f) : 0;
  if (wpc == 0) {
    // one CTA per SM holding that SM's share of the rows (any warp count 1..16: the kernel is
    // not specialised on it)
    const int per_sm = (A.nrows + p->num_sms - 1) / p->num_sms;
    constexpr int MAXW = DIAL_MAXTHREADS / 32;   // the kernel's __launch_bounds__ (dial_host.h)
    wpc = per_sm < MAXW ? per_sm : MAXW;
    // more than one wave of rows: balance the waves of one CTA per SM.  Measured at N=8192
    // (56 rows per SM): H1 4 waves of 14 warps 5.57 ms, 16 warps 5.76 ms, two resident 8-warp
    // CTAs 5.91 ms; Go2 3.85 / 3.93 / 3.81 ms.
    if (per_sm > MAXW) {
      const int waves = (per_sm + MAXW - 1) / MAXW;
      wpc = (per_sm + waves - 1) / waves;
    }
    // respect the 227 KB shared-memory limit of one CTA
    const size_t fixed = sizeof(DevModel) + sizeof(DevPlan), slab = (size_t)p->hM.warp_floats * sizeof(float);
    while (wpc > 1 && fixed + wpc * slab > 227 * 1024) --wpc;
  }
  if (wpc < 1 || wpc > DIAL_MAXTHREADS / 32) return cudaErrorInvalidValue;
  // lock-step pays off on both solver paths.  The dense (elliptic) path used to run free with
  // dynamic row assignment because MJX's 50-iteration line searches made its rows heavy-tailed;
  // since the line search stops at the detected cycle, sharing the instruction fetch wins there
  // too, and the finer the better: Allegro N=4096, 14 warps per CTA: 91.9 ms free-running with
  // dynamic rows, 76.9 ms with a barrier per env step, 52.8 ms per physics substep, 48.4 ms per
  // Newton iteration (level 3, the dense default).  DIAL_NO_LOCKSTEP=1 restores the free-running
  // warps, DIAL_NO_MIDSYNC=1 / DIAL_DENSE_LOCKSTEP=2 select the coarser levels.
  A.lockstep = (wpc >= 2 && !getenv("DIAL_NO_LOCKSTEP")) ? 1 : 0;
  if (p->hM.dense && !A.lockstep && A.nrows > wpc * p->num_sms && !getenv("DIAL_NO_DYNAMIC_ROWS")) A.row_counter = p->row_counter;
  // second barrier before the constraint solve: the dense path needs it (and a third per Newton
  // iteration); on the star paths it stopped paying once the solver shrank (r02: Go2 0.709 -> 0.705 ms,
  // H1 1.060 -> 1.043 ms without it) — DIAL_MIDSYNC=1 / DIAL_NO_MIDSYNC=1 override
  const bool mid = getenv("DIAL_MIDSYNC") ? true : (getenv("DIAL_NO_MIDSYNC") ? false : (p->hM.dense || !p->hM.s_on));
  if (A.lockstep && mid) A.lockstep = 2;
  { const char* se = getenv("DIAL_SYNC_EVERY"); A.sync_every = se ? atoi(se) : 1; if (A.sync_every < 1) A.sync_every = 1; }
  const char* dl = getenv("DIAL_DENSE_LOCKSTEP");
  if (A.lockstep == 2 && p->hM.dense && !(dl && atoi(dl) == 2)) A.lockstep = 3;
  return launch_rollout(p, A, wpc, st);
}

extern "C" dial_plan* dial_plan_create(const dial_model_desc* model, const dial_plan_desc* cfg) {
  if (!model || !cfg) { g_err = "null descriptor"; return nullptr; }
  dial_plan* p = new (std::nothrow) dial_plan();
  if (!p) { g_err = "out of memory"; return nullptr; }
  std::string err;
  if (!derive_model(*model, p->hM, err)) { g_err = err; delete p; return nullptr; }
  {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
      p->num_sms = sms;
  }
  p->variant = (getenv("DIAL_FORCE_GENERIC_TREE") && !p->hM.dense) ? 0 : star_variant(p->hM);
  if (p->variant < 0) { g_err = "this build instantiates the dense (elliptic) solver path for nv = " DIAL_STR(DIAL_DENSE_NV) " only (custom builds: dial_mpc_b200.custom compiles it for the model's nv)"; delete p; return nullptr; }
  memset(&p->hP, 0, sizeof(DevPlan));
  p->hP.c = *cfg;
  p->hP.c.cmd_step = -1;      // command overrides come through dial_plan_set_command only
  const dial_plan_desc& c = *cfg;
  if (c.Hsample + 1 > DIAL_MAXH || c.Hnode + 1 > DIAL_MAXNODE || c.Hnode < 1 || c.Nsample < 1 || c.Ntotal < c.Nsample ||
      c.n_frames < 1 || (c.Hnode + 1) * model->nu > YBAR_THREADS || c.n_stage > DIAL_MAXSTAGE) {
    g_err = "invalid plan configuration (Hsample/Hnode/Nsample/n_frames out of range)";
    delete p;
    return nullptr;
  }
  if (c.env_id < DIAL_ENV_GO2_WALK || c.env_id > DIAL_ENV_CUSTOM) { g_err = "unknown env_id"; delete p; return nullptr; }
#ifndef DIAL_CUSTOM_REWARD_FILE
  if (c.env_id == DIAL_ENV_CUSTOM) {
    g_err = "env_id DIAL_ENV_CUSTOM needs a library built with a reward source (dial_mpc_b200.custom.build_library); this is the stock build";
    delete p; return nullptr;
  }
#endif
#ifdef DIAL_ONLY_VARIANT
  if (p->variant != DIAL_ONLY_VARIANT) {
    g_err = "this custom build holds solver variant " + std::to_string(DIAL_ONLY_VARIANT) + " only, the model needs " + std::to_string(p->variant);
    delete p; return nullptr;
  }
#endif
  if (c.n_user < 0 || c.n_user > DIAL_MAXUSER) { g_err = "n_user out of range"; delete p; return nullptr; }
  auto bad = [&](cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    dial_plan_destroy(p);
    return (dial_plan*)nullptr;
  };
  cudaError_t e;
  if ((e = cudaMalloc(&p->dM, sizeof(DevModel))) != cudaSuccess) return bad(e, "cudaMalloc(model)");
  if ((e = cudaMalloc(&p->dP, sizeof(DevPlan))) != cudaSuccess) return bad(e, "cudaMalloc(plan)");
  if ((e = cudaMemcpy(p->dM, &p->hM, sizeof(DevModel), cudaMemcpyHostToDevice)) != cudaSuccess) return bad(e, "cudaMemcpy(model)");
  if ((e = cudaMemcpy(p->dP, &p->hP, sizeof(DevPlan), cudaMemcpyHostToDevice)) != cudaSuccess) return bad(e, "cudaMemcpy(plan)");
  const size_t rows = (size_t)c.Nsample + 1, H = (size_t)c.Hsample + 1;
  const dial_model_desc& m = *model;
  for (int b = 0; b < 2; ++b) {
    if ((e = cudaMalloc(&p->traj_q[b], rows * H * m.nq * sizeof(float))) != cudaSuccess) return bad(e, "cudaMalloc(traj_q)");
    if ((e = cudaMalloc(&p->traj_qd[b], rows * H * m.nv * sizeof(float))) != cudaSuccess) return bad(e, "cudaMalloc(traj_qd)");
    if ((e = cudaMalloc(&p->traj_x[b], rows * H * 3 * (m.nbody - 1) * sizeof(float))) != cudaSuccess) return bad(e, "cudaMalloc(traj_x)");
  }
  if ((e = cudaMalloc(&p->weights, ((size_t)c.Ntotal + 1) * sizeof(float))) != cudaSuccess) return bad(e, "cudaMalloc(weights)");
  if ((e = cudaMalloc(&p->weights2, ((size_t)c.Ntotal + 1) * sizeof(float))) != cudaSuccess) return bad(e, "cudaMalloc(weights2)");
  if ((e = cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking)) != cudaSuccess) return bad(e, "cudaStreamCreate(side)");
  for (int i = 0; i < 2; ++i) {
    if ((e = cudaEventCreateWithFlags(&p->ev_main[i], cudaEventDisableTiming)) != cudaSuccess) return bad(e, "cudaEventCreate");
    if ((e = cudaEventCreateWithFlags(&p->ev_side[i], cudaEventDisableTiming)) != cudaSuccess) return bad(e, "cudaEventCreate");
  }
  const int ne = (c.Hnode + 1) * m.nu, slots = YBAR_THREADS / ne;
  int g = (c.Ntotal + 1 + slots - 1) / slots;
  p->ybar_grid = g < 1 ? 1 : (g > 296 ? 296 : g);
  {
    int gu = (c.Ntotal + 1 + slots * 8 - 1) / (slots * 8);
    p->upd_grid = gu < 1 ? 1 : (gu > p->ybar_grid ? p->ybar_grid : gu);
  }
  if ((e = cudaMalloc(&p->partial, (size_t)p->ybar_grid * (ne + 1) * sizeof(float))) != cudaSuccess) return bad(e, "cudaMalloc(partial)");
  if ((e = cudaMalloc(&p->tb_partial, (size_t)TB_CHUNKS * H * (m.nq + m.nv + 3 * (m.nbody - 1)) * sizeof(float))) != cudaSuccess) return bad(e, "cudaMalloc(tb_partial)");
  if ((e = cudaMalloc(&p->counter, sizeof(unsigned int))) != cudaSuccess) return bad(e, "cudaMalloc(counter)");
  if ((e = cudaMemset(p->counter, 0, sizeof(unsigned int))) != cudaSuccess) return bad(e, "cudaMemset(counter)");
  if ((e = cudaMalloc(&p->row_counter, sizeof(unsigned int))) != cudaSuccess) return bad(e, "cudaMalloc(row_counter)");
  if (getenv("DIAL_DEBUG_COUNTERS")) {
    if ((e = cudaMalloc(&p->dbg, 8 * sizeof(float))) != cudaSuccess) return bad(e, "cudaMalloc(dbg)");
    cudaMemset(p->dbg, 0, 8 * sizeof(float));
  }
  if ((e = cudaMalloc(&p->zeros, DIAL_MAXV * sizeof(float))) != cudaSuccess) return bad(e, "cudaMalloc(zeros)");
  if ((e = cudaMemset(p->zeros, 0, DIAL_MAXV * sizeof(float))) != cudaSuccess) return bad(e, "cudaMemset(zeros)");
  return p;
}

extern "C" void dial_plan_destroy(dial_plan* p) {
  if (!p) return;
  cudaFree(p->dM); cudaFree(p->dP);
  for (int b = 0; b < 2; ++b) { cudaFree(p->traj_q[b]); cudaFree(p->traj_qd[b]); cudaFree(p->traj_x[b]); }
  for (auto& g : p->mpc_graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  for (int r = 0; r < DIAL_MAXRANK; ++r) {
    if (!p->xch.base[r]) continue;
    if (r == p->xch.rank) cudaFree(p->xch.base[r]); else cudaIpcCloseMemHandle(p->xch.base[r]);
  }
  cudaFree(p->xch.bars_partial);
  cudaFree(p->mpc_Msh); cudaFree(p->mpc_Y1); cudaFree(p->mpc_key);
  for (int i = 0; i < 2; ++i) { if (p->ev_main[i]) cudaEventDestroy(p->ev_main[i]); if (p->ev_side[i]) cudaEventDestroy(p->ev_side[i]); }
  if (p->side) cudaStreamDestroy(p->side);
  cudaFree(p->weights2);
  cudaFree(p->weights); cudaFree(p->partial); cudaFree(p->tb_partial); cudaFree(p->counter); cudaFree(p->row_counter); cudaFree(p->zeros); cudaFree(p->dbg);
  delete p;
}

static void fill_state(RolloutArgs& A, const dial_state* s) {
  A.qpos0 = s->qpos; A.qvel0 = s->qvel; A.warm0 = s->qacc_warmstart; A.step0 = s->step; A.stage0 = s->stage;
}

extern "C" int dial_rollout(dial_plan* p, const dial_state* s, const float* us, int B, int H, float* rewss,
                            float* q, float* qd, float* xpos, void* stream) {
  if (!p || !s || !us || !rewss) return fail("dial_rollout: null argument");
  if (B < 1 || H < 1) return fail("dial_rollout: B and H must be positive");
  RolloutArgs A; memset(&A, 0, sizeof(A));
  fill_state(A, s);
  A.nrows = B; A.H = H; A.mode = 0; A.us = us; A.rewss = rewss; A.q = q; A.qd = qd; A.xpos = xpos;
  CUDA_OK(launch_rollout_any(p, A, (cudaStream_t)stream));
  return 0;
}

extern "C" int dial_plan_set_command(dial_plan* p, int cmd_step, const float* vel, const float* ang, void* stream) {
  if (!p) return fail("dial_plan_set_command: null plan");
  if (cmd_step >= 0 && (!vel || !ang)) return fail("dial_plan_set_command: null command");
  dial_plan_desc& c = p->hP.c;
  c.cmd_step = cmd_step;
  for (int i = 0; i < 3; ++i) { c.cmd_vel[i] = cmd_step >= 0 ? vel[i] : 0.f; c.cmd_ang[i] = cmd_step >= 0 ? ang[i] : 0.f; }
  // 28 bytes from pageable host memory: staged by the driver before the call returns, stream-ordered on the device
  const size_t off = offsetof(dial_plan_desc, cmd_step);
  CUDA_OK(cudaMemcpyAsync((char*)p->dP + off, (const char*)&p->hP.c + off, sizeof(int32_t) + 6 * sizeof(float),
                          cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return 0;
}

extern "C" int dial_plan_set_stages(dial_plan* p, int n_stage, const float* pose_seq, const float* yaw_seq,
                                    const float* contact_targets, const float* contact_radius, void* stream) {
  if (!p) return fail("dial_plan_set_stages: null plan");
  if (n_stage < 1 || n_stage > DIAL_MAXSTAGE) return fail("dial_plan_set_stages: n_stage out of range (1..DIAL_MAXSTAGE)");
  if (!pose_seq || !yaw_seq || !contact_targets || !contact_radius) return fail("dial_plan_set_stages: null table");
  dial_plan_desc& c = p->hP.c;
  c.n_stage = n_stage;
  memset(c.pose_seq, 0, sizeof(c.pose_seq)); memset(c.yaw_seq, 0, sizeof(c.yaw_seq));
  memset(c.contact_targets, 0, sizeof(c.contact_targets)); memset(c.contact_radius, 0, sizeof(c.contact_radius));
  memcpy(c.pose_seq, pose_seq, sizeof(float) * 3 * n_stage);
  memcpy(c.yaw_seq, yaw_seq, sizeof(float) * n_stage);
  memcpy(c.contact_targets, contact_targets, sizeof(float) * 12 * n_stage);
  memcpy(c.contact_radius, contact_radius, sizeof(float) * 4 * n_stage);
  // one stream-ordered copy of [n_stage .. n_user) from the plan's own host mirror (pageable: staged
  // by the driver before the call returns)
  const size_t off = offsetof(dial_plan_desc, n_stage), end = offsetof(dial_plan_desc, n_user);
  CUDA_OK(cudaMemcpyAsync((char*)p->dP + off, (const char*)&p->hP.c + off, end - off, cudaMemcpyHostToDevice,
                          (cudaStream_t)stream));
  return 0;
}

extern "C" int dial_env_step(dial_plan* p, const dial_state* s, const float* action, float* qpos_out,
                             float* qvel_out, float* warm_out, float* reward, float* ctrl_out, void* stream) {
  return dial_env_step_kin(p, s, action, qpos_out, qvel_out, warm_out, reward, ctrl_out, nullptr, stream);
}

extern "C" int dial_env_step_kin(dial_plan* p, const dial_state* s, const float* action, float* qpos_out,
                                 float* qvel_out, float* warm_out, float* reward, float* ctrl_out, float* kin_out,
                                 void* stream) {
  if (!p || !s || !action || !qpos_out || !qvel_out || !warm_out || !reward) return fail("dial_env_step: null argument");
  RolloutArgs A; memset(&A, 0, sizeof(A));
  fill_state(A, s);
  A.nrows = 1; A.H = 1; A.mode = 0; A.us = action; A.rewss = reward;
  A.qpos_out = qpos_out; A.qvel_out = qvel_out; A.warm_out = warm_out; A.ctrl_out = ctrl_out; A.kin_out = kin_out;
  CUDA_OK(launch_rollout(p, A, 1, (cudaStream_t)stream));
  return 0;
}

extern "C" int dial_pipeline_init(dial_plan* p, const float* qpos, const float* qvel, float* qpos_out,
                                  float* warm_out, void* stream) {
  if (!p || !qpos || !qvel || !qpos_out || !warm_out) return fail("dial_pipeline_init: null argument");
  RolloutArgs A; memset(&A, 0, sizeof(A));
  A.qpos0 = qpos; A.qvel0 = qvel; A.warm0 = p->zeros;  // mjx.make_data: qacc_warmstart = 0
  A.nrows = 1; A.H = 1; A.mode = 2; A.qpos_out = qpos_out; A.warm_out = warm_out;
  CUDA_OK(launch_rollout(p, A, 1, (cudaStream_t)stream));
  return 0;
}

extern "C" int dial_reverse_rollout(dial_plan* p, const dial_state* s, const float* eps, const uint32_t key[2],
                                    const float* Ybar, const float* noise_scale, float* rews_local, void* stream) {
  if (!p || !s || !Ybar || !noise_scale || !rews_local) return fail("dial_reverse_rollout: null argument");
  if (!eps && !key) return fail("dial_reverse_rollout: need eps or key");
  RolloutArgs A; memset(&A, 0, sizeof(A));
  fill_state(A, s);
  const dial_plan_desc& c = p->hP.c;
  A.nrows = c.Nsample + 1; A.H = c.Hsample + 1; A.mode = 1;
  A.eps = eps; A.Ybar = Ybar; A.noise = noise_scale;
  if (key) { A.key0 = key[0]; A.key1 = key[1]; }
  p->cur ^= 1;
  A.rews = rews_local; A.q = p->traj_q[p->cur]; A.qd = p->traj_qd[p->cur]; A.xpos = p->traj_x[p->cur];
  A.dbg = p->dbg;
  fill_xch(p, A);   // sharded plans with a connected exchange: rewards go straight to every rank's mailbox
  CUDA_OK(launch_rollout_any(p, A, (cudaStream_t)stream));
  return 0;
}

extern "C" int dial_reverse_update(dial_plan* p, const float* eps, const uint32_t key[2], const float* Ybar,
                                   const float* noise_scale, const float* rews_all, float* Ybar_out,
                                   float* weights, void* stream) {
  return dial_reverse_update_x(p, eps, key, Ybar, noise_scale, rews_all, Ybar_out, weights, nullptr, stream);
}

extern "C" int dial_reverse_update_x(dial_plan* p, const float* eps, const uint32_t key[2], const float* Ybar,
                                     const float* noise_scale, const float* rews_all, float* Ybar_out,
                                     float* weights, float* rews_gathered, void* stream) {
  if (!p || !Ybar || !noise_scale || !Ybar_out) return fail("dial_reverse_update: null argument");
  if (!rews_all && !p->xch.on) return fail("dial_reverse_update: rews_all may be NULL only with a connected exchange");
  if (!eps && !key) return fail("dial_reverse_update: need eps or key");
  const dial_plan_desc& c = p->hP.c;
  cudaStream_t st = (cudaStream_t)stream;
  float* w = weights ? weights : p->weights;
  // rews_all == NULL: the rewards are in this rank's mailbox (written by every rank's rollout
  // epilogue over NVLink); the kernel waits for the flags, and rews_gathered gets a compact copy
  XchWait X = xch_wait_args(p, rews_gathered);
  if (rews_all) X.mbox = nullptr;
  weights_kernel<<<1, 1024, 0, st>>>(rews_all, c.Ntotal + 1, c.temp_sample, w, X);
  p->launches++;
  CUDA_OK(cudaGetLastError());
  ybar_kernel<<<p->ybar_grid, YBAR_THREADS, 0, st>>>(w, eps, key ? key[0] : 0u, key ? key[1] : 0u, Ybar, noise_scale,
                                                     c.Ntotal, c.Hnode + 1, p->hM.m.nu, p->partial, p->counter, Ybar_out, nullptr);
  p->launches++;
  CUDA_OK(cudaGetLastError());
  if (weights && weights != p->weights)
    CUDA_OK(cudaMemcpyAsync(p->weights, weights, ((size_t)c.Ntotal + 1) * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

extern "C" int dial_reverse_trajbar(dial_plan* p, const float* weights, int rank, float* qbar, float* qdbar,
                                    float* xbar, void* stream) {
  if (!p) return fail("dial_reverse_trajbar: null plan");
  const dial_plan_desc& c = p->hP.c;
  const dial_model_desc& m = p->hM.m;
  cudaStream_t st = (cudaStream_t)stream;
  const float* w = weights ? weights : p->weights;
  const int H = c.Hsample + 1, rows = c.Nsample + 1;
  TrajArgs T;
  T.traj[0] = p->traj_q[p->cur]; T.traj[1] = p->traj_qd[p->cur]; T.traj[2] = p->traj_x[p->cur];
  T.out[0] = qbar; T.out[1] = qdbar; T.out[2] = xbar;
  const bool xsum = p->xch.on && qbar && qdbar && xbar;   // sum over ranks on the device (peer memory)
  if (xsum) {
    T.out[0] = p->xch.bars_partial; T.out[1] = T.out[0] + (size_t)H * m.nq; T.out[2] = T.out[1] + (size_t)H * m.nv;
  }
  T.ncol[0] = m.nq; T.ncol[1] = m.nv; T.ncol[2] = 3 * (m.nbody - 1);
  T.coloff[0] = 0; T.coloff[1] = m.nq; T.coloff[2] = m.nq + m.nv;
  T.coltot = m.nq + m.nv + 3 * (m.nbody - 1);
  T.nrows = rows; T.H = H; T.weights = w; T.w_offset = c.shard_offset; T.mean_row = c.Nsample;
  T.mean_weight_index = c.Ntotal; T.include_mean = rank == 0 ? 1 : 0; T.partial = p->tb_partial;
  const int maxlen = H * (T.ncol[2] > T.ncol[0] ? T.ncol[2] : T.ncol[0]);   // nq = nv + 1 > nv always
  trajbar_partial_kernel<<<dim3((maxlen + 255) / 256, TB_CHUNKS, 3), 256, 0, st>>>(T);
  p->launches++;
  CUDA_OK(cudaGetLastError());
  trajbar_final_kernel<<<H, 128, 0, st>>>(T);
  p->launches++;
  CUDA_OK(cudaGetLastError());
  if (xsum) {
    BarsXch Bx; memset(&Bx, 0, sizeof(Bx));
    for (int r = 0; r < p->xch.world; ++r) { Bx.mbox[r] = p->xch.bars(r); Bx.flags[r] = p->xch.bflags(r); }
    Bx.seq = p->xch.bseq(); Bx.err = p->xch.err(); Bx.world = p->xch.world; Bx.rank = p->xch.rank; Bx.nbar = p->xch.nbar;
    Bx.partial = p->xch.bars_partial; Bx.out[0] = qbar; Bx.out[1] = qdbar; Bx.out[2] = xbar;
    Bx.n0 = H * m.nq; Bx.n1 = H * m.nv;
    bars_allreduce_kernel<<<1, 1024, 0, st>>>(Bx);
    p->launches++;
    CUDA_OK(cudaGetLastError());
  }
  return 0;
}

// ---- multi-GPU exchange over NVLink peer memory -------------------------------------------------
extern "C" int dial_exchange_create(dial_plan* p, int rank, int world, unsigned char handle_out[DIAL_IPC_HANDLE_BYTES]) {
  if (!p || !handle_out) return fail("dial_exchange_create: null argument");
  if (world < 2 || world > DIAL_MAXRANK || rank < 0 || rank >= world) return fail("dial_exchange_create: need 2 <= world <= DIAL_MAXRANK");
  static_assert(sizeof(cudaIpcMemHandle_t) <= DIAL_IPC_HANDLE_BYTES, "IPC handle does not fit");
  const dial_plan_desc& c = p->hP.c;
  const dial_model_desc& m = p->hM.m;
  if (c.Ntotal != c.Nsample * world || c.shard_offset != rank * c.Nsample) return fail("dial_exchange_create: plan shard does not match rank/world");
  if (p->xch.base[p->xch.rank]) return fail("dial_exchange_create: already created");
  dial_plan::Exchange& x = p->xch;
  x.rank = rank; x.world = world;
  x.nbar = (c.Hsample + 1) * (m.nq + m.nv + 3 * (m.nbody - 1));
  auto up = [](size_t n) { return (n + 31) & ~(size_t)31; };
  size_t o = 0;
  x.o_mbox = o; o += up(2 * ((size_t)c.Ntotal + 1));
  x.o_bars = o; o += up(2 * (size_t)DIAL_MAXRANK * x.nbar);
  x.o_flags = o; o += up(2 * DIAL_MAXRANK);
  x.o_bflags = o; o += up(2 * DIAL_MAXRANK);
  x.o_local = o; o += 32;
  x.words = o;
  void* d = nullptr;
  CUDA_OK(cudaMalloc(&d, x.words * sizeof(uint32_t)));
  CUDA_OK(cudaMemset(d, 0, x.words * sizeof(uint32_t)));
  x.base[rank] = reinterpret_cast<uint32_t*>(d);
  CUDA_OK(cudaMalloc(&x.bars_partial, (size_t)x.nbar * sizeof(float)));
  cudaIpcMemHandle_t h;
  CUDA_OK(cudaIpcGetMemHandle(&h, d));
  memset(handle_out, 0, DIAL_IPC_HANDLE_BYTES);
  memcpy(handle_out, &h, sizeof(h));
  return 0;
}

extern "C" int dial_exchange_connect(dial_plan* p, const unsigned char* handles) {
  if (!p || !handles) return fail("dial_exchange_connect: null argument");
  dial_plan::Exchange& x = p->xch;
  if (!x.base[x.rank]) return fail("dial_exchange_connect: call dial_exchange_create first");
  for (int r = 0; r < x.world; ++r) {
    if (r == x.rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * DIAL_IPC_HANDLE_BYTES, sizeof(h));
    void* d = nullptr;
    CUDA_OK(cudaIpcOpenMemHandle(&d, h, cudaIpcMemLazyEnablePeerAccess));
    x.base[r] = reinterpret_cast<uint32_t*>(d);
  }
  x.on = true;
  return 0;
}

extern "C" int dial_exchange_status(dial_plan* p, uint32_t out[6]) {
  if (!p || !out) return fail("dial_exchange_status: null argument");
  if (!p->xch.on) { for (int i = 0; i < 6; ++i) out[i] = 0; return 0; }
  CUDA_OK(cudaMemcpy(out, p->xch.base[p->xch.rank] + p->xch.o_local, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int dial_reverse_trajectories(dial_plan* p, float* q, float* qd, float* xpos, void* stream) {
  if (!p) return fail("dial_reverse_trajectories: null plan");
  const dial_plan_desc& c = p->hP.c;
  const dial_model_desc& m = p->hM.m;
  const size_t n = ((size_t)c.Nsample + 1) * (c.Hsample + 1) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (q) CUDA_OK(cudaMemcpyAsync(q, p->traj_q[p->cur], n * m.nq, cudaMemcpyDeviceToDevice, st));
  if (qd) CUDA_OK(cudaMemcpyAsync(qd, p->traj_qd[p->cur], n * m.nv, cudaMemcpyDeviceToDevice, st));
  if (xpos) CUDA_OK(cudaMemcpyAsync(xpos, p->traj_x[p->cur], n * 3 * (m.nbody - 1), cudaMemcpyDeviceToDevice, st));
  return 0;
}

// ---- device-resident synchronous MPC loop ---------------------------------------------------
extern "C" int dial_mpc_bind(dial_plan* p, const dial_mpc_buffers* b, const float* M_shift) {
  if (!p || !b || !M_shift) return fail("dial_mpc_bind: null argument");
  if (!b->qpos || !b->qvel || !b->qacc_warmstart || !b->counters || !b->rng || !b->Y || !b->ctrl || !b->reward ||
      !b->rews || !b->noise)
    return fail("dial_mpc_bind: only qbar/qdbar/xbar may be null");
  const dial_plan_desc& c = p->hP.c;
  if (c.Ntotal != c.Nsample && !p->xch.on)
    return fail("dial_mpc_bind: a sharded plan needs a connected exchange (dial_exchange_create / dial_exchange_connect) for the device-resident loop");
  if (c.Ntotal != c.Nsample && !b->rews_all) return fail("dial_mpc_bind: sharded plans need rews_all [Ntotal+1]");
  const int n1 = c.Hnode + 1, nu = p->hM.m.nu;
  for (auto& g : p->mpc_graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  p->mpc_graphs.clear();
  if (!p->mpc_Msh) CUDA_OK(cudaMalloc(&p->mpc_Msh, DIAL_MAXNODE * DIAL_MAXNODE * sizeof(float)));
  if (!p->mpc_Y1) CUDA_OK(cudaMalloc(&p->mpc_Y1, DIAL_MAXNODE * DIAL_MAXU * sizeof(float)));
  if (!p->mpc_key) CUDA_OK(cudaMalloc(&p->mpc_key, 2 * sizeof(uint32_t)));
  CUDA_OK(cudaMemcpy(p->mpc_Msh, M_shift, (size_t)n1 * n1 * sizeof(float), cudaMemcpyHostToDevice));
  (void)nu;
  p->mpc = *b;
  p->mpc_bound = true;
  return 0;
}

// enqueue one MPC step on `st` (eagerly or into a capture)
static int mpc_enqueue(dial_plan* p, int n_diffuse, int env_step, cudaStream_t st) {
  const dial_plan_desc& c = p->hP.c;
  const dial_mpc_buffers& B = p->mpc;
  const int n1 = c.Hnode + 1, nu = p->hM.m.nu;
  float* Y[2] = {B.Y, p->mpc_Y1};
  int cur = 0;
  if (env_step == 1) {
    // state = step_env(state, Y0[0])  (dial_core.py:245): in place, counters advanced by the kernel
    RolloutArgs A; memset(&A, 0, sizeof(A));
    A.qpos0 = B.qpos; A.qvel0 = B.qvel; A.warm0 = B.qacc_warmstart;
    A.counters_in = B.counters; A.counters_out = B.counters;
    A.nrows = 1; A.H = 1; A.mode = 0; A.us = Y[cur]; A.rewss = B.reward;
    A.qpos_out = B.qpos; A.qvel_out = B.qvel; A.warm_out = B.qacc_warmstart; A.ctrl_out = B.ctrl;
    CUDA_OK(launch_rollout(p, A, 1, st));
  }
  if (env_step == 1 || env_step == 2) {
    // Y0 = shift(Y0)  (dial_core.py:252)
    mpc_shift_kernel<<<1, DIAL_MAXNODE * DIAL_MAXU, 0, st>>>(p->mpc_Msh, Y[cur], Y[cur ^ 1], n1, nu);
    p->launches++;
    CUDA_OK(cudaGetLastError());
    cur ^= 1;
  }
  // The info-only bars (qbar, qdbar, xbar; dial_core.py:133-135) are computed for EVERY iteration,
  // like the reference's scan does (the caller sees those of the last one), on a side branch of
  // the graph: the bars of iteration i read trajectory buffer i&1 and weights buffer i&1 while
  // iteration i+1 rolls into the other pair; iteration i+2 waits for them.
  const bool bars = B.qbar && B.qdbar && B.xbar;
  float* wts[2] = {p->weights, p->weights2};
  for (int i = 0; i < n_diffuse; ++i) {
    const float* noise = B.noise + (size_t)i * n1;
    // rng split folded into the rollout (key = split(rng)[1]) and the fused update kernel (which also
    // advances rng); beyond 131072 samples the three-kernel sequence is kept
    const bool fused = c.Ntotal + 1 <= (1 << 17) && !getenv("DIAL_NO_FUSED_UPDATE");
    if (!fused) {
      mpc_split_kernel<<<1, 32, 0, st>>>(B.rng, p->mpc_key);
      p->launches++;
      CUDA_OK(cudaGetLastError());
    }
    if (bars && i >= 2) CUDA_OK(cudaStreamWaitEvent(st, p->ev_side[i & 1], 0));
    RolloutArgs A; memset(&A, 0, sizeof(A));
    A.qpos0 = B.qpos; A.qvel0 = B.qvel; A.warm0 = B.qacc_warmstart; A.counters_in = B.counters;
    A.nrows = c.Nsample + 1; A.H = c.Hsample + 1; A.mode = 1;
    A.Ybar = Y[cur]; A.noise = noise;
    if (fused) A.rng_dev = B.rng; else A.key_dev = p->mpc_key;
    p->cur ^= 1;
    A.rews = B.rews; A.q = p->traj_q[p->cur]; A.qd = p->traj_qd[p->cur]; A.xpos = p->traj_x[p->cur];
    A.dbg = p->dbg;
    fill_xch(p, A);
    CUDA_OK(launch_rollout_any(p, A, st));
    float* w = wts[i & 1];
    XchWait X = xch_wait_args(p, B.rews_all);
    if (fused) {
      update_kernel<<<p->upd_grid, YBAR_THREADS, 0, st>>>(B.rews, c.Ntotal + 1, c.temp_sample, w, X, B.rng, Y[cur], noise, c.Ntotal,
                                                          n1, nu, p->partial, p->counter, Y[cur ^ 1]);
      p->launches++;
      CUDA_OK(cudaGetLastError());
    } else {
      weights_kernel<<<1, 1024, 0, st>>>(B.rews, c.Ntotal + 1, c.temp_sample, w, X);
      p->launches++;
      CUDA_OK(cudaGetLastError());
      ybar_kernel<<<p->ybar_grid, YBAR_THREADS, 0, st>>>(w, nullptr, 0u, 0u, Y[cur], noise, c.Ntotal, n1, nu,
                                                         p->partial, p->counter, Y[cur ^ 1], p->mpc_key);
      p->launches++;
      CUDA_OK(cudaGetLastError());
    }
    cur ^= 1;
    if (bars) {
      CUDA_OK(cudaEventRecord(p->ev_main[i & 1], st));
      CUDA_OK(cudaStreamWaitEvent(p->side, p->ev_main[i & 1], 0));
      int rc = dial_reverse_trajbar(p, w, p->xch.on ? p->xch.rank : 0, B.qbar, B.qdbar, B.xbar, (void*)p->side);
      if (rc) return rc;
      CUDA_OK(cudaEventRecord(p->ev_side[i & 1], p->side));
    }
  }
  if (cur != 0) CUDA_OK(cudaMemcpyAsync(Y[0], Y[1], (size_t)n1 * nu * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (bars && n_diffuse > 0) {   // join the bars branch (both outstanding iterations)
    if (n_diffuse >= 2) CUDA_OK(cudaStreamWaitEvent(st, p->ev_side[(n_diffuse - 2) & 1], 0));
    CUDA_OK(cudaStreamWaitEvent(st, p->ev_side[(n_diffuse - 1) & 1], 0));
  }
  if (n_diffuse > 0 && wts[(n_diffuse - 1) & 1] != p->weights)   // p->weights always holds the last iteration's weights
    CUDA_OK(cudaMemcpyAsync(p->weights, p->weights2, ((size_t)c.Ntotal + 1) * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

extern "C" int dial_mpc_step(dial_plan* p, int n_diffuse, int env_step, void* stream) {
  if (!p) return fail("dial_mpc_step: null plan");
  if (!p->mpc_bound) return fail("dial_mpc_step: call dial_mpc_bind first");
  if (n_diffuse < 0 || n_diffuse > 64) return fail("dial_mpc_step: n_diffuse out of range");
  if (env_step < 0 || env_step > 2) return fail("dial_mpc_step: env_step must be 0 (plan only), 1 (env step + shift) or 2 (shift only)");
  cudaStream_t st = (cudaStream_t)stream;
  dial_plan::MpcGraph* g = nullptr;
  for (auto& e : p->mpc_graphs) if (e.n_diffuse == n_diffuse && e.env_step == env_step) g = &e;
  if (!g) {
    // first use of this shape: run it eagerly (also configures the kernels' shared-memory limits)
    p->mpc_graphs.push_back({n_diffuse, env_step, 1, nullptr, 0});
    return mpc_enqueue(p, n_diffuse, env_step, st);
  }
  if (!g->exec) {
    if (getenv("DIAL_NO_GRAPH")) return mpc_enqueue(p, n_diffuse, env_step, st);
    // second use: capture the same sequence into a graph, then replay it from now on
    cudaStream_t cs;
    CUDA_OK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    const int64_t l0 = p->launches;
    cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { cudaStreamDestroy(cs); CUDA_OK(e); }
    int rc = mpc_enqueue(p, n_diffuse, env_step, cs);
    cudaGraph_t graph = nullptr;
    e = cudaStreamEndCapture(cs, &graph);
    cudaStreamDestroy(cs);
    g->launches = p->launches - l0;
    p->launches = l0;
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    CUDA_OK(e);
    e = cudaGraphInstantiate(&g->exec, graph, 0);
    cudaGraphDestroy(graph);
    CUDA_OK(e);
  }
  CUDA_OK(cudaGraphLaunch(g->exec, st));
  p->launches += g->launches;
  return 0;
}

extern "C" void dial_key_split(const uint32_t key[2], uint32_t out0[2], uint32_t out1[2]) {
  // jax.random.split(key, 2), legacy layout: counters [0,1,2,3] -> halves (0,1) | (2,3)
  uint32_t a0 = 0, b0 = 2, a1 = 1, b1 = 3;
  threefry2x32(key[0], key[1], a0, b0);
  threefry2x32(key[0], key[1], a1, b1);
  out0[0] = a0; out0[1] = a1; out1[0] = b0; out1[1] = b1;
}

extern "C" int dial_solver_variant(const dial_model_desc* model) {
  if (!model) return fail("null descriptor");
  DevModel* D = new (std::nothrow) DevModel();
  if (!D) return fail("out of memory");
  std::string err;
  int v = -1;
  if (!derive_model(*model, *D, err)) g_err = err;
  else if ((v = star_variant(*D)) < 0) g_err = "this build instantiates the dense (elliptic) solver path for nv = " DIAL_STR(DIAL_DENSE_NV) " only (custom builds: dial_mpc_b200.custom compiles it for the model's nv)";
  delete D;
  return v;
}

extern "C" const char* dial_custom_reward_id(void) {
#if defined(DIAL_CUSTOM_REWARD_FILE) && defined(DIAL_CUSTOM_REWARD_ID)
  return DIAL_STR(DIAL_CUSTOM_REWARD_ID);
#elif defined(DIAL_CUSTOM_REWARD_FILE)
  return "custom";
#else
  return "";
#endif
}

// ---- in-run fp32 peak (roofline denominator): independent FFMA chains at full occupancy ------
__global__ void __launch_bounds__(1024) fp32_peak_kernel(float* out, int iters, float a, float b) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = (float)(threadIdx.x + i) * 1e-3f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 123.456f) out[0] = s;   // never true: keeps the chains alive
}

extern "C" int dial_fp32_peak(int iters, float* tflops_out) {
  if (!tflops_out || iters < 1) return fail("dial_fp32_peak: bad argument");
  int dev = 0, sms = 0;
  CUDA_OK(cudaGetDevice(&dev));
  CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  float* d = nullptr;
  CUDA_OK(cudaMalloc(&d, sizeof(float)));
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0));
  CUDA_OK(cudaEventCreate(&e1));
  const int grid = sms * 2;
  fp32_peak_kernel<<<grid, 1024>>>(d, 64, 0.999f, 1e-3f);   // warm-up
  float best = 0.f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    fp32_peak_kernel<<<grid, 1024>>>(d, iters, 0.999f, 1e-3f);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) { cudaFree(d); return fail(std::string("fp32 peak kernel: ") + cudaGetErrorString(e)); }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * 8 * 16 * (double)iters * 1024.0 * grid;
    const float tf = (float)(flop / (ms * 1e-3) / 1e12);
    best = tf > best ? tf : best;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  *tflops_out = best;
  return 0;
}

extern "C" int64_t dial_launch_count(const dial_plan* p) { return p ? p->launches : 0; }

extern "C" int dial_debug_counters(dial_plan* p, float out[8]) {
  if (!p || !p->dbg) return fail("debug counters are off (set DIAL_DEBUG_COUNTERS=1 before dial_plan_create)");
  CUDA_OK(cudaMemcpy(out, p->dbg, 8 * sizeof(float), cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemset(p->dbg, 0, 8 * sizeof(float)));
  return 0;
}
