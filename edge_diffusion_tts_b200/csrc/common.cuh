// Shared helpers for libedtts.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/edtts.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libedtts is written for sm_100a (B200) only"
#endif

namespace edtts {

constexpr int H = EDTTS_HIDDEN;        // 160
constexpr int M = EDTTS_N_MELS;        // 80
constexpr int NH = EDTTS_HEADS;        // 4
constexpr int HD = EDTTS_HEAD_DIM;     // 40
constexpr int RANK = EDTTS_KV_RANK;    // 80
constexpr int FFN = EDTTS_FFN_HIDDEN;  // 320
constexpr int NL = EDTTS_LAYERS;       // 4
constexpr int WIN = EDTTS_WINDOW;      // 64

// ---- error reporting across the C ABI -------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);          // cudaPeekAtLastError -> EDTTS_ECUDA

#define EDTTS_REQUIRE(cond, code, ...)  \
  do {                                  \
    if (!(cond)) {                      \
      ::edtts::set_error(__VA_ARGS__);  \
      return (code);                    \
    }                                   \
  } while (0)

// Kernel classes for the launch counter and the optional per-class event timing
// (edtts_prof_*): bench.py uses them to time the dominant kernel live.
enum KernelClass : int {
  KC_GEMM_SIMT = 0, KC_ATTN_WINDOW_SIMT, KC_ATTN_CROSS_SIMT, KC_COND, KC_EMBED, KC_VQ, KC_SCHEDULE, KC_DSCONV,
  KC_TC_GEMM, KC_TC_ATTN_WINDOW, KC_TC_ATTN_CROSS, KC_TC_MISC, KC_TC_LAYER, KC_MEL, KC_T3_GEMM, KC_T3_ATTN_WINDOW,
  KC_T3_ATTN_CROSS, KC_COUNT
};

// RAII around one kernel launch: counts it and, when profiling is on and the stream is
// not capturing, brackets it with CUDA events on the launching stream.
struct LaunchScope {
  LaunchScope(int cls, cudaStream_t stream);
  ~LaunchScope();
  int cls_;
  cudaStream_t stream_;
  void* slot_;
};

// ---- device facts (never the literal 148): cached per device ordinal -------------------------------------
int current_device();                        // cudaGetDevice, -1 on error
int sm_count();                              // cudaDevAttrMultiProcessorCount of the current device
// cudaFuncSetAttribute is per DEVICE: a function-local `static PerDeviceOnce once;` guards the attribute set-up of a kernel
struct PerDeviceOnce {
  bool done[64] = {};
  bool need() const { const int d = current_device(); return d < 0 || d >= 64 || !done[d]; }
  void set() { const int d = current_device(); if (d >= 0 && d < 64) done[d] = true; }
};
// grid cap of a grid-stride streaming kernel: `per_sm` resident CTAs on every SM of the current device
static inline int64_t stream_grid_cap(int per_sm) { return (int64_t)sm_count() * per_sm; }

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// ---- device math ------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) {          // nn.GELU() default (exact erf)
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// DDIM eta=0 / eta>0 update with torch's operation order and no FMA contraction
// (schedule.py:179-202), so that it is bit-identical to the reference given the
// same eps.  ab_t, ab_p are alpha_bar[t], alpha_bar[max(t_prev,0)] (or 1 if t_prev<0).
__device__ __forceinline__ void ddim_update(float x, float e, float nz, float ab_t, float ab_p, float eta,
                                            float& x_prev, float& x0) {
  const float s1 = __fsqrt_rn(__fsub_rn(1.0f, ab_t));
  float v = __fdiv_rn(__fsub_rn(x, __fmul_rn(s1, e)), __fsqrt_rn(ab_t));
  v = fminf(fmaxf(v, -3.0f), 3.0f);
  x0 = v;
  const float om_p = __fsub_rn(1.0f, ab_p);
  const float ratio = __fmul_rn(__fdiv_rn(om_p, __fsub_rn(1.0f, ab_t)), __fsub_rn(1.0f, __fdiv_rn(ab_t, ab_p)));
  const float sigma = __fmul_rn(eta, __fsqrt_rn(ratio));
  const float dir = __fmul_rn(__fsqrt_rn(__fsub_rn(om_p, __fmul_rn(sigma, sigma))), e);
  x_prev = __fadd_rn(__fadd_rn(__fmul_rn(__fsqrt_rn(ab_p), v), dir), __fmul_rn(sigma, nz));
}

// DDPM ancestral update (schedule.py:221-238), torch operation order.
__device__ __forceinline__ float ddpm_update(float x, float e, float nz, float alpha, float alpha_bar, float beta,
                                             float var, float nonzero) {
  const float coef1 = __fdiv_rn(1.0f, __fsqrt_rn(alpha));
  const float coef2 = __fdiv_rn(beta, __fsqrt_rn(__fsub_rn(1.0f, alpha_bar)));
  const float mean = __fmul_rn(coef1, __fsub_rn(x, __fmul_rn(coef2, e)));
  return __fadd_rn(mean, __fmul_rn(__fmul_rn(nonzero, __fsqrt_rn(var)), nz));
}

// DPM-Solver++ update (schedule.py:326-438, 479-481), torch operation order, one rounding per operation.
// c = {sa, sb, c0, c1, c2, 1/r, c3, -}; mode 0: mo is v, 1: mo is x0 (both clamped to +-3), 2: mo is x0 used as given.
__device__ __forceinline__ void dpm_update(float x, float mo, float h1, float h2, const float* c, int order, int mode,
                                           float& x_prev, float& x0) {
  float p0 = mode ? mo : __fsub_rn(__fmul_rn(c[0], x), __fmul_rn(c[1], mo));
  if (mode != 2) p0 = fminf(fmaxf(p0, -3.0f), 3.0f);
  x0 = p0;
  float r = __fadd_rn(__fmul_rn(c[2], x), __fmul_rn(c[3], p0));
  if (order == 2) {
    const float d1 = __fmul_rn(c[5], __fsub_rn(p0, h1));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(c[4], d1), 0.5f));
  } else if (order >= 3) {
    const float d1 = __fsub_rn(p0, h1);
    const float d2 = __fadd_rn(__fsub_rn(p0, __fmul_rn(2.0f, h1)), h2);
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(c[4], d1), 0.5f));
    r = __fadd_rn(r, __fdiv_rn(__fmul_rn(c[6], d2), 6.0f));
  }
  x_prev = r;
}

}  // namespace edtts
