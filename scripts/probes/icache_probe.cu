// icache_probe.cu — how fast does a B200 SM execute straight-line code that does not fit its instruction
// caches, (a) when all warps of the SM walk it together and (b) when every warp is somewhere else in it?
// The design of the rollout kernel (one CTA per SM, warps in lock-step) rests on the answer.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache_probe icache_probe.cu && ./icache_probe
// Body: SEGS segments of 256 independent-chain FFMAs (4 accumulators: not latency-bound), entered through a
// jump table so that a warp can start at any segment; KB = SEGS * 256 * 16 / 1024.  "scattered": warp w enters
// the first pass at segment 7 w SEGS/16 and keeps that distance to the others afterwards; `sync` > 0: a CTA
// barrier in front of every sync-th segment (lock-step at that period).  One JSON line per run; results and
// reading: profiles/r02_icache_probe.{md,json,jsonl}.  (The 16 KB body — four segments, one jump-table
// dispatch per 1024 instructions — is slower than the 32 KB one for reasons of its own and is left out there.)
#include <cstdio>
#include <cuda_runtime.h>

#define F4 a0 = fmaf(a0, x, y); a1 = fmaf(a1, x, y); a2 = fmaf(a2, x, y); a3 = fmaf(a3, x, y);
#define F16 F4 F4 F4 F4
#define F64 F16 F16 F16 F16
#define F256 F64 F64 F64 F64
#define SEG(i) case i: if (sync > 0 && ((i) % sync) == 0) __syncthreads(); F256

template <int SEGS>
__global__ void __launch_bounds__(512, 1) probe(float* out, long long* cyc, int iters, int scatter, int sync, float x, float y) {
  float a0 = threadIdx.x, a1 = 1.f, a2 = 2.f, a3 = 3.f;
  const int warp = threadIdx.x >> 5;
  int start = scatter ? (warp * (SEGS / 16 > 0 ? SEGS / 16 : 1) * 7) % SEGS : 0;   // warps spread over the body
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    switch (start) {
      SEG(0) SEG(1) SEG(2) SEG(3) SEG(4) SEG(5) SEG(6) SEG(7)
      if (SEGS <= 8) break;
      SEG(8) SEG(9) SEG(10) SEG(11) SEG(12) SEG(13) SEG(14) SEG(15)
      if (SEGS <= 16) break;
      SEG(16) SEG(17) SEG(18) SEG(19) SEG(20) SEG(21) SEG(22) SEG(23)
      SEG(24) SEG(25) SEG(26) SEG(27) SEG(28) SEG(29) SEG(30) SEG(31)
      if (SEGS <= 32) break;
      SEG(32) SEG(33) SEG(34) SEG(35) SEG(36) SEG(37) SEG(38) SEG(39)
      SEG(40) SEG(41) SEG(42) SEG(43) SEG(44) SEG(45) SEG(46) SEG(47)
      SEG(48) SEG(49) SEG(50) SEG(51) SEG(52) SEG(53) SEG(54) SEG(55)
      SEG(56) SEG(57) SEG(58) SEG(59) SEG(60) SEG(61) SEG(62) SEG(63)
      default: break;
    }
    start = 0;   // later passes run the whole body: the warps keep their distance in the code (a fixed drift)
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 16 + warp] = t1 - t0;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[gridDim.x * 16] = 0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}

template <int SEGS>
void run(int warps, int scatter, float* out, long long* cyc, int sms, int sync = 0) {
  const int iters = 64;
  probe<SEGS><<<sms, warps * 32, 0>>>(out, cyc, 4, scatter, sync, 1.0001f, 0.5f);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe<SEGS><<<sms, warps * 32, 0>>>(out, cyc, iters, scatter, sync, 1.0001f, 0.5f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[16];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double worst = 0.0, mean = 0.0, total_instr = 0.0;
  long long mx = 0;
  for (int w = 0; w < warps; ++w) {
    const int start = scatter ? (w * (SEGS / 16 > 0 ? SEGS / 16 : 1) * 7) % SEGS : 0;
    const double instr = ((double)iters * SEGS - start) * 256;   // FFMAs only (loop / switch overhead ignored)
    const double cpi = h[w] / instr;
    worst = cpi > worst ? cpi : worst; mean += cpi / warps; total_instr += instr;
    mx = h[w] > mx ? h[w] : mx;
  }
  printf("{\"body_kb\": %d, \"warps_per_sm\": %d, \"mode\": \"%s\", \"barrier_every_instr\": %d, \"ms\": %.4f, \"cycles_per_instr_per_warp_mean\": %.3f, "
         "\"cycles_per_instr_per_warp_worst\": %.3f, \"ipc_sm\": %.3f}\n",
         SEGS * 4, warps, scatter ? "scattered" : "together", sync * 256, ms, mean, worst, total_instr / mx);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * 512);
  cudaMalloc(&cyc, sizeof(long long) * (sms * 16 + 1));
  for (int scatter = 0; scatter < 2; ++scatter)
    for (int warps : {1, 4, 14}) {
      run<4>(warps, scatter, out, cyc, sms);     // 16 KB: fits L1.5
      run<8>(warps, scatter, out, cyc, sms);     // 32 KB
      run<16>(warps, scatter, out, cyc, sms);    // 64 KB
      run<32>(warps, scatter, out, cyc, sms);    // 128 KB (the rollout kernel executes ~136 KB per env step)
      run<64>(warps, scatter, out, cyc, sms);    // 256 KB
    }
  // warps per SM and CTA barriers (lock-step) on the 128 KB / 256 KB bodies
  for (int warps : {8, 12, 14, 16}) {
    for (int sync : {0, 32, 8, 2}) {             // barrier every 8192 / 2048 / 512 instructions
      run<32>(warps, 0, out, cyc, sms, sync);
      run<64>(warps, 0, out, cyc, sms, sync);
    }
  }
  printf("done %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
