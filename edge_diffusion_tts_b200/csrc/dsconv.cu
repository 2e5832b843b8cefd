// DepthwiseSeparableConv (layers/conv.py:10-64), operator level:
//   depthwise k-tap Conv1d (groups=C, no bias, padding k/2, stride s)
//   -> pointwise 1x1 Conv1d (+bias) -> GroupNorm(min(8,C_out), eps 1e-5) -> GELU(erf)
// on [B, C, T] tensors.
//
// Kernel 1 (per utterance, per 64-frame tile): the input window of every channel is
//   staged in shared memory with coalesced loads, the depthwise taps are applied from
//   shared memory, and the pointwise channel mix runs as a register-blocked product
//   against the staged tile (4 output channels per thread, weights read warp-uniform).
// Kernel 2 (per utterance, per group): two-pass mean / variance in a fixed order
//   (deterministic: no float atomics).
// Kernel 3: normalise + affine + GELU in place, 128-bit accesses.
#include "common.cuh"

namespace edtts {

constexpr int DC_TT = 64;        // output frames per tile
constexpr int DC_THREADS = 256;

__global__ void __launch_bounds__(DC_THREADS) dsconv_mix_kernel(const float* __restrict__ x, const float* __restrict__ dw,
                                                                const float* __restrict__ pw,
                                                                const float* __restrict__ pb, float* __restrict__ y,
                                                                int c_in, int c_out, int T, int t_out, int k,
                                                                int stride) {
  extern __shared__ float sm[];
  const int span = (DC_TT - 1) * stride + k;
  float* xs = sm;                      // [c_in][span]
  float* ds = sm + (size_t)c_in * span;   // [c_in][DC_TT]
  const int b = blockIdx.y, t0 = blockIdx.x * DC_TT;
  const int pad = k / 2;
  const int in0 = t0 * stride - pad;
  const float* xb = x + (int64_t)b * c_in * T;
  for (int i = threadIdx.x; i < c_in * span; i += DC_THREADS) {
    const int c = i / span, p = i % span;
    const int ti = in0 + p;
    xs[i] = (ti >= 0 && ti < T) ? xb[(int64_t)c * T + ti] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c_in * DC_TT; i += DC_THREADS) {
    const int c = i / DC_TT, tt = i % DC_TT;
    float s = 0.f;
    for (int j = 0; j < k; ++j) s = fmaf(dw[c * k + j], xs[c * span + tt * stride + j], s);
    ds[i] = s;
  }
  __syncthreads();
  const int tt = threadIdx.x % DC_TT, g = threadIdx.x / DC_TT;   // 4 channel groups, warp-uniform
  const int t = t0 + tt;
  float* yb = y + (int64_t)b * c_out * t_out;
  for (int co = g * 4; co < c_out; co += 16) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const int c1 = min(co + 1, c_out - 1), c2 = min(co + 2, c_out - 1), c3 = min(co + 3, c_out - 1);
    for (int ci = 0; ci < c_in; ++ci) {
      const float d = ds[ci * DC_TT + tt];
      a0 = fmaf(pw[(int64_t)co * c_in + ci], d, a0);
      a1 = fmaf(pw[(int64_t)c1 * c_in + ci], d, a1);
      a2 = fmaf(pw[(int64_t)c2 * c_in + ci], d, a2);
      a3 = fmaf(pw[(int64_t)c3 * c_in + ci], d, a3);
    }
    if (t < t_out) {
      yb[(int64_t)co * t_out + t] = a0 + pb[co];
      if (co + 1 < c_out) yb[(int64_t)(co + 1) * t_out + t] = a1 + pb[co + 1];
      if (co + 2 < c_out) yb[(int64_t)(co + 2) * t_out + t] = a2 + pb[co + 2];
      if (co + 3 < c_out) yb[(int64_t)(co + 3) * t_out + t] = a3 + pb[co + 3];
    }
  }
}

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w];
  return s;
}

// stats[b][g] = (mean, rstd) over the contiguous (c_out/G) x t_out slab of group g
__global__ void __launch_bounds__(256) dsconv_stats_kernel(const float* __restrict__ y, float* __restrict__ stats,
                                                           int64_t slab, float eps) {
  __shared__ float red[8];
  const float* p = y + (int64_t)blockIdx.x * slab;
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < slab; i += 256) s += p[i];
  const float mean = block_sum_256(s, red) / (float)slab;
  float q = 0.f;
  for (int64_t i = threadIdx.x; i < slab; i += 256) {
    const float d = p[i] - mean;
    q = fmaf(d, d, q);
  }
  const float var = block_sum_256(q, red) / (float)slab;
  if (threadIdx.x == 0) {
    stats[2 * blockIdx.x] = mean;
    stats[2 * blockIdx.x + 1] = 1.0f / sqrtf(var + eps);
  }
}

__global__ void __launch_bounds__(256) dsconv_norm_gelu_kernel(float* __restrict__ y, const float* __restrict__ stats,
                                                               const float* __restrict__ gw, const float* __restrict__ gb,
                                                               int64_t total, int c_out, int t_out, int ch_per_group) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bc = i / t_out;            // b * c_out + c
    const int c = (int)(bc % c_out);
    const int64_t bg = (bc / c_out) * (c_out / ch_per_group) + c / ch_per_group;
    const float mean = stats[2 * bg], rstd = stats[2 * bg + 1];
    y[i] = gelu_erf((y[i] - mean) * rstd * gw[c] + gb[c]);
  }
}

}  // namespace edtts

using namespace edtts;

extern "C" int64_t edtts_dsconv_workspace_bytes(int32_t B, int32_t c_out, int32_t t_out) {
  (void)t_out;
  const int groups = c_out < 8 ? c_out : 8;
  return align_up((int64_t)B * groups * 2 * 4, 256);
}

extern "C" int edtts_dsconv_forward(const float* x, const float* dw_w, const float* pw_w, const float* pw_b,
                                    const float* gn_w, const float* gn_b, float* y_out, void* workspace,
                                    int64_t workspace_bytes, int32_t B, int32_t c_in, int32_t c_out, int32_t T,
                                    int32_t kernel_size, int32_t stride, void* stream) {
  EDTTS_REQUIRE(x && dw_w && pw_w && pw_b && gn_w && gn_b && y_out && workspace, EDTTS_EINVAL, "dsconv: null argument");
  EDTTS_REQUIRE(B > 0 && c_in > 0 && c_out > 0 && T > 0 && kernel_size > 0 && stride > 0, EDTTS_EINVAL,
                "dsconv: bad sizes");
  const int groups = c_out < 8 ? c_out : 8;     // conv.py:48
  EDTTS_REQUIRE(c_out % groups == 0, EDTTS_EINVAL, "dsconv: C_out=%d not divisible by %d groups", c_out, groups);
  const int pad = kernel_size / 2;
  const int t_out = (T + 2 * pad - kernel_size) / stride + 1;
  EDTTS_REQUIRE(t_out > 0, EDTTS_EINVAL, "dsconv: empty output");
  EDTTS_REQUIRE(workspace_bytes >= edtts_dsconv_workspace_bytes(B, c_out, t_out), EDTTS_ENOSPC, "dsconv: workspace");
  const int span = (DC_TT - 1) * stride + kernel_size;
  const size_t smem = ((size_t)c_in * span + (size_t)c_in * DC_TT) * 4;
  EDTTS_REQUIRE(smem <= 227 * 1024, EDTTS_ENOTSUP, "dsconv: C_in=%d stride=%d k=%d needs %zu B of shared memory", c_in,
                stride, kernel_size, smem);
  cudaStream_t st = as_stream(stream);
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(dsconv_mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return check_launch("dsconv smem attribute");
  dim3 grid((t_out + DC_TT - 1) / DC_TT, B);
  int rc;
  {
    LaunchScope ls(KC_DSCONV, st);
    dsconv_mix_kernel<<<grid, DC_THREADS, smem, st>>>(x, dw_w, pw_w, pw_b, y_out, c_in, c_out, T, t_out, kernel_size,
                                                      stride);
    rc = check_launch("dsconv_mix");
  }
  if (rc) return rc;
  const int cpg = c_out / groups;
  float* stats = reinterpret_cast<float*>(workspace);
  {
    LaunchScope ls(KC_DSCONV, st);
    dsconv_stats_kernel<<<B * groups, 256, 0, st>>>(y_out, stats, (int64_t)cpg * t_out, 1e-5f);
    rc = check_launch("dsconv_stats");
  }
  if (rc) return rc;
  const int64_t total = (int64_t)B * c_out * t_out;
  const int64_t blocks = (total + 255) / 256;
  LaunchScope ls(KC_DSCONV, st);
  dsconv_norm_gelu_kernel<<<(unsigned)(blocks > stream_grid_cap(8) ? stream_grid_cap(8) : blocks), 256, 0, st>>>(y_out, stats, gn_w, gn_b,
                                                                                          total, c_out, t_out, cpg);
  return check_launch("dsconv_norm_gelu");
}
