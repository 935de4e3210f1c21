// TEST-ONLY: runs the per-warp device code of csrc/dial_device.cuh on the CPU through the
// lock-step fiber emulator (warp_emul.h).  Used by tests/test_emul_*.py to debug kernel
// logic without a GPU.  Never loaded by the dial_mpc_b200 package.
#define DIAL_HOST_EMUL 1
#include <vector>
#include <string>
#include <stdio.h>
#ifdef DIAL_EMUL_TRACE
static std::vector<int> g_trace;
extern "C" void emul_trace(int kind, int val) { g_trace.push_back(kind); g_trace.push_back(val); }
extern "C" int emul_trace_take(int* out, int cap) {
  int n = (int)g_trace.size() < cap ? (int)g_trace.size() : cap;
  for (int i = 0; i < n; ++i) out[i] = g_trace[i];
  g_trace.clear();
  return n;
}
#endif
#include "../../dial_mpc_b200/csrc/dial_host.h"

int forced_variant = -1;
extern "C" void emul_force_variant(int v) { forced_variant = v; }

// derived-model flags (tests): out = {star variant, s_on, sb_on, warp_floats}
extern "C" int emul_model_flags(const dial_model_desc* m, int out[4]) {
  static DevModel D;
  std::string err;
  if (!derive_model(*m, D, err)) { fprintf(stderr, "emul: %s\n", err.c_str()); return -1; }
  out[0] = star_variant(D); out[1] = D.s_on; out[2] = D.sb_on; out[3] = D.warp_floats;
  return 0;
}

extern "C" int emul_rollout(const dial_model_desc* m, const dial_plan_desc* c, int mode, int nrows,
                            int H, int step0, int stage0, const float* qpos0, const float* qvel0,
                            const float* warm0, const float* us, const float* eps, const float* Ybar,
                            const float* noise, uint32_t key0, uint32_t key1, float* rewss, float* rews,
                            float* q, float* qd, float* xpos, float* qpos_out, float* qvel_out,
                            float* warm_out, float* ctrl_out, float* slab_out) {
  static DevModel D;
  static DevPlan P;
  std::string err;
  if (!derive_model(*m, D, err)) { fprintf(stderr, "emul: %s\n", err.c_str()); return -1; }
  P.c = *c;
  RolloutArgs A;
  memset(&A, 0, sizeof(A));
  A.nrows = nrows; A.H = H; A.mode = mode; A.step0 = step0; A.stage0 = stage0;
  A.qpos0 = qpos0; A.qvel0 = qvel0; A.warm0 = warm0; A.us = us; A.eps = eps; A.Ybar = Ybar;
  A.noise = noise; A.key0 = key0; A.key1 = key1; A.rewss = rewss; A.rews = rews; A.q = q; A.qd = qd;
  A.xpos = xpos; A.qpos_out = qpos_out; A.qvel_out = qvel_out; A.warm_out = warm_out; A.ctrl_out = ctrl_out;
  std::vector<float> slab(D.warp_floats, 0.f);
  for (int row = 0; row < nrows; ++row) {
    const int variant = forced_variant >= 0 ? forced_variant : star_variant(D);
    emul::run_warp([&](int lane) {
      if (variant == 1) rollout_warp<3, 6>(&D, &P, slab.data(), A, row, lane);
      else if (variant == 2) rollout_warp<5, 7>(&D, &P, slab.data(), A, row, lane);
      else if (variant == 3) rollout_warp<-1, DIAL_DENSE_NV>(&D, &P, slab.data(), A, row, lane);
      else if (variant == 4) rollout_warp<5, 6>(&D, &P, slab.data(), A, row, lane);
      else rollout_warp<0, 0>(&D, &P, slab.data(), A, row, lane);
    });
    if (slab_out && row == 0) memcpy(slab_out, slab.data(), sizeof(float) * D.warp_floats);
  }
  return 0;
}

