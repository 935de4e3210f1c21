// dial_rollout_variant.cu — the rollout kernel of ONE solver variant (-DDIAL_VARIANT=v):
//   1 star<3,6> (quadruped)   2 star<5,7> (humanoid)   3 dense<DIAL_DENSE_NV> (elliptic cones; 22 in the stock build)
//   4 star<5,6>               0 generic tree (level-scheduled compact Cholesky)
// Compiled once per variant so that the instantiations build in parallel (each is ~200 KB of
// straight-line SASS); exports one launcher, dial_launch_rollout_v<v>.
//
//   rollout_kernel<NL,NR>   one warp per sample row, persistent over the horizon; 1..16 warps
//                           per CTA (blockDim, chosen by the host's launch policy); the compiled
//                           model + plan constants are staged into shared memory with one TMA
//                           bulk copy (cp.async.bulk + mbarrier) per CTA.
#include <cuda_runtime.h>
#include "dial_host.h"

#ifndef DIAL_VARIANT
#error "compile with -DDIAL_VARIANT=<0..4>"
#endif
#if DIAL_VARIANT == 1
#define DIAL_V_NL 3
#define DIAL_V_NR 6
#elif DIAL_VARIANT == 2
#define DIAL_V_NL 5
#define DIAL_V_NR 7
#elif DIAL_VARIANT == 3
#define DIAL_V_NL -1
#define DIAL_V_NR DIAL_DENSE_NV
#elif DIAL_VARIANT == 4
#define DIAL_V_NL 5
#define DIAL_V_NR 6
#else
#define DIAL_V_NL 0
#define DIAL_V_NR 0
#endif

// ---------------------------------------------------------------------------------
// TMA bulk copy global -> shared (1-D), completion on an mbarrier
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

static_assert(sizeof(DevModel) % 16 == 0, "DevModel must be a multiple of 16 bytes for cp.async.bulk");
static_assert(sizeof(DevPlan) % 16 == 0, "DevPlan must be a multiple of 16 bytes for cp.async.bulk");

// 512 threads, one CTA per SM -> 128 registers per thread; the same binary serves every CTA
// shape (1..16 warps), so per-sample results do not depend on the launch shape.
template <int NL, int NR>
__global__ void __launch_bounds__(DIAL_MAXTHREADS, 1) rollout_kernel(const DevModel* __restrict__ gM,
                                                         const DevPlan* __restrict__ gP,
                                                         const RolloutArgs A) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  DevModel* sM = reinterpret_cast<DevModel*>(smem);
  DevPlan* sP = reinterpret_cast<DevPlan*>(smem + sizeof(DevModel));
  float* slabs = reinterpret_cast<float*>(smem + sizeof(DevModel) + sizeof(DevPlan));
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_expect_tx(&bar, (uint32_t)(sizeof(DevModel) + sizeof(DevPlan)));
    tma_bulk_g2s(sM, gM, (uint32_t)sizeof(DevModel), &bar);
    tma_bulk_g2s(sP, gP, (uint32_t)sizeof(DevPlan), &bar);
  }
  __syncthreads();
  mbar_wait(&bar, 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  float* slab = slabs + (size_t)warp * sM->warp_floats;
  if (A.row_counter) {
    // persistent warps + dynamic row assignment (free-running dense path: heavy-tailed cost per row)
    for (;;) {
      int row = 0;
      if (lane == 0) row = (int)atomicAdd(A.row_counter, 1u);
      row = __shfl_sync(0xffffffffu, row, 0);
      if (row >= A.nrows) return;
      rollout_warp<NL, NR>(sM, sP, slab, A, row, lane);
      __syncwarp();
    }
  }
  int row = blockIdx.x * wpc + warp;
  bool active = true;
  if (row >= A.nrows) {
    active = A.lockstep != 0;
    row = A.nrows - 1;  // lock-step CTAs need every warp at the barriers: duplicate the last row (benign)
  }
  if (active) rollout_warp<NL, NR>(sM, sP, slab, A, row, lane);
#ifndef DIAL_NO_XCH
  if (A.xch_world > 1) {
    // reward exchange epilogue: the rows of this CTA are in every rank's mailbox (peer stores over
    // NVLink); the last CTA of the grid publishes them by raising this rank's flag everywhere
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      const unsigned int old = atomicAdd(A.xch_done, 1u);
      if (old == gridDim.x - 1) {
        __threadfence_system();
        const uint32_t seq = *A.xch_seq, buf = seq & 1u;
        for (int p = 0; p < A.xch_world; ++p)
          *reinterpret_cast<volatile uint32_t*>(A.xch_flags[p] + buf * DIAL_MAXRANK + A.xch_rank) = seq + 1u;
        *A.xch_done = 0u;
        __threadfence_system();
      }
    }
  }
#endif
}

#define DIAL_CAT2(a, b) a##b
#define DIAL_CAT(a, b) DIAL_CAT2(a, b)

cudaError_t DIAL_CAT(dial_launch_rollout_v, DIAL_VARIANT)(const DevModel* dM, const DevPlan* dP, const RolloutArgs& A,
                                                          int grid, int wpc, size_t smem, cudaStream_t st) {
  // the dynamic shared-memory opt-in is per device (and per kernel): remember what each device has
  static size_t configured[64] = {0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64 || smem > configured[dev]) {
    e = cudaFuncSetAttribute(rollout_kernel<DIAL_V_NL, DIAL_V_NR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = smem;
  }
  rollout_kernel<DIAL_V_NL, DIAL_V_NR><<<grid, wpc * 32, smem, st>>>(dM, dP, A);
  return cudaGetLastError();
}
