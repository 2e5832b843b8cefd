// Stand-alone DiffusionSchedule update rules (schedule.py:157-238) -- used when a
// caller drives DiffusionSchedule.get_ddim_step / ddpm_step directly; inside
// generate_mel the same arithmetic runs fused in the last kernel of the step.
// Pure streaming kernels: read x, eps (+noise), write x_prev (+x0); 128-bit
// accesses, grid sized in multiples of the SM count.
#include "common.cuh"

namespace edtts {

template <int VEC>
__global__ void __launch_bounds__(256) ddim_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                                                   const float* __restrict__ noise, const float* __restrict__ alpha_bar,
                                                   const int64_t* __restrict__ t, const int64_t* __restrict__ t_prev,
                                                   float eta, float* __restrict__ x_prev, float* __restrict__ x0,
                                                   int64_t total, int64_t n) {
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC; i < total;
       i += (int64_t)gridDim.x * blockDim.x * VEC) {
    const int b = (int)(i / n);
    const float ab_t = alpha_bar[t[b]];
    const int64_t tp = t_prev[b];
    const float ab_p = tp >= 0 ? alpha_bar[tp] : 1.0f;
    float xv[VEC], ev[VEC], nv[VEC], xp[VEC], xz[VEC];
    if (VEC == 4) {
      *reinterpret_cast<float4*>(xv) = *reinterpret_cast<const float4*>(x + i);
      *reinterpret_cast<float4*>(ev) = *reinterpret_cast<const float4*>(eps + i);
      if (noise) *reinterpret_cast<float4*>(nv) = *reinterpret_cast<const float4*>(noise + i);
    } else {
      xv[0] = x[i];
      ev[0] = eps[i];
      if (noise) nv[0] = noise[i];
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) ddim_update(xv[j], ev[j], noise ? nv[j] : 0.f, ab_t, ab_p, eta, xp[j], xz[j]);
    if (VEC == 4) {
      if (x_prev) *reinterpret_cast<float4*>(x_prev + i) = *reinterpret_cast<float4*>(xp);
      if (x0) *reinterpret_cast<float4*>(x0 + i) = *reinterpret_cast<float4*>(xz);
    } else {
      if (x_prev) x_prev[i] = xp[0];
      if (x0) x0[i] = xz[0];
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) ddpm_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                                                   const float* __restrict__ noise, const float* __restrict__ alphas,
                                                   const float* __restrict__ alpha_bar, const float* __restrict__ betas,
                                                   const float* __restrict__ post_var, const int64_t* __restrict__ t,
                                                   float* __restrict__ x_prev, int64_t total, int64_t n) {
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC; i < total;
       i += (int64_t)gridDim.x * blockDim.x * VEC) {
    const int b = (int)(i / n);
    const int64_t tt = t[b];
    const float al = alphas[tt], ab = alpha_bar[tt], be = betas[tt], pv = post_var[tt];
    const float nzm = tt > 0 ? 1.0f : 0.0f;
    float xv[VEC], ev[VEC], nv[VEC], xp[VEC];
    if (VEC == 4) {
      *reinterpret_cast<float4*>(xv) = *reinterpret_cast<const float4*>(x + i);
      *reinterpret_cast<float4*>(ev) = *reinterpret_cast<const float4*>(eps + i);
      *reinterpret_cast<float4*>(nv) = *reinterpret_cast<const float4*>(noise + i);
    } else {
      xv[0] = x[i];
      ev[0] = eps[i];
      nv[0] = noise[i];
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) xp[j] = ddpm_update(xv[j], ev[j], nv[j], al, ab, be, pv, nzm);
    if (VEC == 4) *reinterpret_cast<float4*>(x_prev + i) = *reinterpret_cast<float4*>(xp);
    else x_prev[i] = xp[0];
  }
}

// DPM-Solver++ update (schedule.py:339-438) with torch's operation order, one rounding per operation (no FMA contraction).
template <int VEC>
__global__ void __launch_bounds__(256) dpm_kernel(const float* __restrict__ x, const float* __restrict__ mo,
                                                  const float* __restrict__ h1, const float* __restrict__ h2,
                                                  const float* __restrict__ coef, int order, int predict_x0,
                                                  float* __restrict__ x_prev, float* __restrict__ x0_out, int64_t total,
                                                  int64_t n) {
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC; i < total;
       i += (int64_t)gridDim.x * blockDim.x * VEC) {
    const float* c = coef + (i / n) * 8;
    float xv[VEC], mv[VEC], av[VEC], bv[VEC], xp[VEC], xz[VEC];
    if (VEC == 4) {
      *reinterpret_cast<float4*>(xv) = *reinterpret_cast<const float4*>(x + i);
      *reinterpret_cast<float4*>(mv) = *reinterpret_cast<const float4*>(mo + i);
      if (order >= 2) *reinterpret_cast<float4*>(av) = *reinterpret_cast<const float4*>(h1 + i);
      if (order >= 3) *reinterpret_cast<float4*>(bv) = *reinterpret_cast<const float4*>(h2 + i);
    } else {
      xv[0] = x[i];
      mv[0] = mo[i];
      if (order >= 2) av[0] = h1[i];
      if (order >= 3) bv[0] = h2[i];
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j)
      dpm_update(xv[j], mv[j], order >= 2 ? av[j] : 0.f, order >= 3 ? bv[j] : 0.f, c, order, predict_x0, xp[j], xz[j]);
    if (VEC == 4) {
      *reinterpret_cast<float4*>(x_prev + i) = *reinterpret_cast<float4*>(xp);
      if (x0_out) *reinterpret_cast<float4*>(x0_out + i) = *reinterpret_cast<float4*>(xz);
    } else {
      x_prev[i] = xp[0];
      if (x0_out) x0_out[i] = xz[0];
    }
  }
}

// v-prediction DDIM step with optional classifier-free guidance (inference_pipeline.py:176-192), torch operation order
__global__ void __launch_bounds__(256) vddim_kernel(const float* __restrict__ x, const float* __restrict__ vc,
                                                    const float* __restrict__ vu, float s, const float* __restrict__ coef,
                                                    float* __restrict__ xn, float* __restrict__ x0_out, int64_t total, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float* c = coef + (i / n) * 4;
    const float xv = x[i];
    float v = vc[i];
    if (vu) v = __fadd_rn(vu[i], __fmul_rn(s, __fsub_rn(v, vu[i])));
    float x0 = __fsub_rn(__fmul_rn(c[0], xv), __fmul_rn(c[1], v));
    x0 = fminf(fmaxf(x0, -3.0f), 3.0f);
    const float eps = __fadd_rn(__fmul_rn(c[1], xv), __fmul_rn(c[0], v));
    if (x0_out) x0_out[i] = x0;
    xn[i] = __fadd_rn(__fmul_rn(c[2], x0), __fmul_rn(c[3], eps));
  }
}
// x[b, :L, :] = sa known + sb noise
__global__ void __launch_bounds__(256) inpaint_inject_kernel(float* __restrict__ x, const float* __restrict__ known,
                                                             const float* __restrict__ noise, const float* __restrict__ coef,
                                                             int64_t total, int64_t per_b, int64_t x_per_b) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / per_b, r = i % per_b;
    const float* c = coef + b * 4;
    x[b * x_per_b + r] = __fadd_rn(__fmul_rn(c[0], known[i]), __fmul_rn(c[1], noise[i]));
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline unsigned stream_grid(int64_t work_items) {
  const int64_t blocks = (work_items + 255) / 256;
  const int64_t cap = stream_grid_cap(8);   // 8 resident 256-thread CTAs per SM
  return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace edtts

using namespace edtts;

extern "C" int edtts_ddim_step(const float* x_t, const float* eps, const float* noise, const float* alpha_bar,
                               const int64_t* t, const int64_t* t_prev, float eta, float* x_prev_out, float* x0_out,
                               int32_t B, int64_t n, void* stream) {
  EDTTS_REQUIRE(x_t && eps && alpha_bar && t && t_prev && (x_prev_out || x0_out) && B > 0 && n > 0, EDTTS_EINVAL,
                "ddim_step: null argument");
  EDTTS_REQUIRE(eta == 0.f || noise, EDTTS_EINVAL, "ddim_step: eta > 0 needs noise");
  const int64_t total = (int64_t)B * n;
  const bool vec = n % 4 == 0 && aligned16(x_t) && aligned16(eps) && (!noise || aligned16(noise)) &&
                   (!x_prev_out || aligned16(x_prev_out)) && (!x0_out || aligned16(x0_out));
  LaunchScope ls(KC_SCHEDULE, as_stream(stream));
  if (vec)
    ddim_kernel<4><<<stream_grid(total / 4), 256, 0, as_stream(stream)>>>(x_t, eps, eta == 0.f ? nullptr : noise,
                                                                          alpha_bar, t, t_prev, eta, x_prev_out, x0_out,
                                                                          total, n);
  else
    ddim_kernel<1><<<stream_grid(total), 256, 0, as_stream(stream)>>>(x_t, eps, eta == 0.f ? nullptr : noise, alpha_bar,
                                                                      t, t_prev, eta, x_prev_out, x0_out, total, n);
  return check_launch("ddim_step");
}

extern "C" int edtts_ddpm_step(const float* x_t, const float* eps, const float* noise, const float* alphas,
                               const float* alpha_bar, const float* betas, const float* posterior_var, const int64_t* t,
                               float* x_prev_out, int32_t B, int64_t n, void* stream) {
  EDTTS_REQUIRE(x_t && eps && noise && alphas && alpha_bar && betas && posterior_var && t && x_prev_out && B > 0 &&
                    n > 0,
                EDTTS_EINVAL, "ddpm_step: null argument");
  const int64_t total = (int64_t)B * n;
  const bool vec = n % 4 == 0 && aligned16(x_t) && aligned16(eps) && aligned16(noise) && aligned16(x_prev_out);
  LaunchScope ls(KC_SCHEDULE, as_stream(stream));
  if (vec)
    ddpm_kernel<4><<<stream_grid(total / 4), 256, 0, as_stream(stream)>>>(x_t, eps, noise, alphas, alpha_bar, betas,
                                                                          posterior_var, t, x_prev_out, total, n);
  else
    ddpm_kernel<1><<<stream_grid(total), 256, 0, as_stream(stream)>>>(x_t, eps, noise, alphas, alpha_bar, betas,
                                                                      posterior_var, t, x_prev_out, total, n);
  return check_launch("ddpm_step");
}

extern "C" int edtts_dpm_step(const float* x_t, const float* model_out, const float* hist1, const float* hist2,
                              const float* coef, int32_t order_used, int32_t predict_x0, float* x_prev_out, float* x0_out,
                              int32_t B, int64_t n, void* stream) {
  EDTTS_REQUIRE(x_t && model_out && coef && x_prev_out && B > 0 && n > 0, EDTTS_EINVAL, "dpm_step: null argument");
  EDTTS_REQUIRE(order_used >= 1 && order_used <= 3, EDTTS_EINVAL, "dpm_step: order_used=%d", order_used);
  EDTTS_REQUIRE((order_used < 2 || hist1) && (order_used < 3 || hist2), EDTTS_EINVAL,
                "dpm_step: order %d needs %d history tensor(s)", order_used, order_used - 1);
  const int64_t total = (int64_t)B * n;
  const bool vec = n % 4 == 0 && aligned16(x_t) && aligned16(model_out) && (!hist1 || aligned16(hist1)) &&
                   (!hist2 || aligned16(hist2)) && aligned16(x_prev_out) && (!x0_out || aligned16(x0_out));
  LaunchScope ls(KC_SCHEDULE, as_stream(stream));
  if (vec)
    dpm_kernel<4><<<stream_grid(total / 4), 256, 0, as_stream(stream)>>>(x_t, model_out, hist1, hist2, coef, order_used,
                                                                         predict_x0, x_prev_out, x0_out, total, n);
  else
    dpm_kernel<1><<<stream_grid(total), 256, 0, as_stream(stream)>>>(x_t, model_out, hist1, hist2, coef, order_used,
                                                                     predict_x0, x_prev_out, x0_out, total, n);
  return check_launch("dpm_step");
}

extern "C" int edtts_vddim_step(const float* x_t, const float* v_cond, const float* v_uncond, float cfg_scale,
                                const float* coef, float* x_next_out, float* x0_out, int32_t B, int64_t n, void* stream) {
  EDTTS_REQUIRE(x_t && v_cond && coef && x_next_out && B > 0 && n > 0, EDTTS_EINVAL, "vddim_step: null argument");
  const int64_t total = (int64_t)B * n;
  LaunchScope ls(KC_SCHEDULE, as_stream(stream));
  vddim_kernel<<<stream_grid(total), 256, 0, as_stream(stream)>>>(x_t, v_cond, v_uncond, cfg_scale, coef, x_next_out, x0_out,
                                                                  total, n);
  return check_launch("vddim_step");
}

extern "C" int edtts_inpaint_inject(float* x, const float* known, const float* noise, const float* coef, int32_t B, int32_t T,
                                    int32_t L, int32_t D, void* stream) {
  EDTTS_REQUIRE(x && known && noise && coef && B > 0 && T > 0 && D > 0 && L >= 0 && L <= T, EDTTS_EINVAL,
                "inpaint_inject: B=%d T=%d L=%d D=%d", B, T, L, D);
  if (L == 0) return EDTTS_OK;
  const int64_t per_b = (int64_t)L * D, total = per_b * B;
  LaunchScope ls(KC_SCHEDULE, as_stream(stream));
  inpaint_inject_kernel<<<stream_grid(total), 256, 0, as_stream(stream)>>>(x, known, noise, coef, total, per_b, (int64_t)T * D);
  return check_launch("inpaint_inject");
}
