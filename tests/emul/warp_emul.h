// warp_emul.h — TEST-ONLY lock-step warp emulator (32 ucontext fibers).
// Lets g++ compile dial_device.cuh (-DDIAL_HOST_EMUL) so kernel logic can be debugged on
// machines without a GPU.  Not part of the product: nothing in dial_mpc_b200 loads it.
#pragma once
#include <ucontext.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <functional>

namespace emul {
struct Warp {
  ucontext_t main_ctx;
  ucontext_t ctx[32];
  char* stacks[32];
  bool done[32];
  int cur;
  float xf[32];
  int xi[32];
  std::function<void(int)> body;
};
inline Warp*& cur_warp() { static thread_local Warp* w = nullptr; return w; }
inline void barrier() { Warp* w = cur_warp(); swapcontext(&w->ctx[w->cur], &w->main_ctx); }
inline void trampoline() {
  Warp* w = cur_warp();
  int lane = w->cur;
  w->body(lane);
  w->done[lane] = true;
  swapcontext(&w->ctx[lane], &w->main_ctx);
}
inline void run_warp(std::function<void(int)> body) {
  Warp w;
  w.body = body;
  cur_warp() = &w;
  const size_t STK = 1 << 20;
  for (int l = 0; l < 32; ++l) {
    w.stacks[l] = (char*)malloc(STK);
    w.done[l] = false;
    getcontext(&w.ctx[l]);
    w.ctx[l].uc_stack.ss_sp = w.stacks[l];
    w.ctx[l].uc_stack.ss_size = STK;
    w.ctx[l].uc_link = &w.main_ctx;
    makecontext(&w.ctx[l], (void (*)())trampoline, 0);
  }
  bool any = true;
  while (any) {
    any = false;
    for (int l = 0; l < 32; ++l) {
      if (w.done[l]) continue;
      any = true;
      w.cur = l;
      swapcontext(&w.main_ctx, &w.ctx[l]);
    }
  }
  for (int l = 0; l < 32; ++l) free(w.stacks[l]);
  cur_warp() = nullptr;
}
}  // namespace emul

inline void syncwarp() { emul::barrier(); }
inline float shfl(float v, int src) {
  emul::Warp* w = emul::cur_warp();
  w->xf[w->cur] = v; emul::barrier();
  float r = w->xf[src & 31]; emul::barrier();
  return r;
}
inline float shfl_xor(float v, int m) { return shfl(v, emul::cur_warp()->cur ^ m); }
inline int shfl_i(int v, int src) {
  emul::Warp* w = emul::cur_warp();
  w->xi[w->cur] = v; emul::barrier();
  int r = w->xi[src & 31]; emul::barrier();
  return r;
}
inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline void cta_sync() {}
inline bool cta_sync_or(bool p) { return p; }  // one emulated warp per CTA
inline void fast_sincos(float x, float& s, float& c) { s = sinf(x); c = cosf(x); }
inline float fast_cos(float x) { return cosf(x); }
