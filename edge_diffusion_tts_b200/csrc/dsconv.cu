// DepthwiseSeparableConv (layers/conv.py:10-64), operator level:
//   depthwise k-tap Conv1d (groups=C, no bias, padding k/2, stride s)
//   -> pointwise 1x1 Conv1d (+bias) -> GroupNorm(min(8,C_out), eps 1e-5) -> GELU(erf)
// on [B, C, T] tensors.
//
// Tensor-core route (C_in % 8 == 0, C_out % 16 == 0, C_out <= 256, 8 groups): THREE passes over y instead of five.
//   dsconv_tc_kernel, one CTA per (utterance, 128 output frames), two CTAs per SM:
//     * the input window of 32 channels at a time is staged in shared memory with 128-bit loads (x rows are contiguous in T),
//     * the depthwise taps are applied from shared memory in fp32; the result (128 frames x 32 channels) is written as the A
//       operand of the pointwise product -- split into a tf32 head and a tf32 tail (a = hi + lo, 22 mantissa bits),
//     * the pointwise 1x1 convolution is a tcgen05 GEMM (M = 128 frames, N = C_out, K = C_in) with the accumulator in tensor
//       memory: per 8-channel k-step three kind::tf32 MMAs (hi hi, lo hi, hi lo) -- fp32-grade accuracy (the parity bar of
//       this operator is 2e-5 max-abs, which single-pass bf16 / tf32 cannot meet); the weight image (hi | lo, core-matrix
//       layout) is packed once per call and fetched per 32-channel chunk with one bulk copy (TMA engine),
//     * epilogue: accumulator -> registers, + bias, y written once (coalesced along T), GroupNorm statistics of the tile
//       accumulated on the way: per column a warp-shuffle sum over the 32 rows, per (tile, group) one fp64 (sum, sum of
//       squares) partial in the workspace -- deterministic, no float atomics.
//   dsconv_finish_kernel: every (utterance, channel) row combines the <= ceil(T'/128) partials of its group in fp64
//     (mean, rstd) and applies normalise + affine + GELU in place with 128-bit accesses.
// Other shapes take the CUDA-core route below (kernel 1-3, the round-1 implementation).
//
// CUDA-core route:
// Kernel 1 (per utterance, per 64-frame tile): the input window of every channel is
//   staged in shared memory with coalesced loads, the depthwise taps are applied from
//   shared memory, and the pointwise channel mix runs as a register-blocked product
//   against the staged tile (4 output channels per thread, weights read warp-uniform).
// Kernel 2 (per utterance, per group): two-pass mean / variance in a fixed order
//   (deterministic: no float atomics).
// Kernel 3: normalise + affine + GELU in place.
#include "common.cuh"
#include "umma.cuh"
#include "tf32x3.cuh"
#include <stdlib.h>


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


// =====================================================================================================================
// tensor-core route
// =====================================================================================================================
namespace dctc {
using namespace tc;
using namespace t3;

constexpr int THREADS = 256;
constexpr int A_BYTES = A_HALF;

template <int K>
__device__ __forceinline__ float taps(const float* __restrict__ wr, const float* __restrict__ xr, int k) {
  float s = 0.f;
  if (K > 0) {
#pragma unroll
    for (int j = 0; j < K; ++j) s = fmaf(wr[j], xr[j], s);
  } else {
    for (int j = 0; j < k; ++j) s = fmaf(wr[j], xr[j], s);
  }
  return s;
}

struct Args {
  const float* x;        // [B][c_in][T]
  const float* dw;       // [c_in][k]
  const float* wimg;     // packed pointwise weights
  const float* pb;       // [c_out]
  float* y;              // [B][c_out][t_out]
  double* part;          // [B][ntiles][8][2]  (sum, sum of squares) per (tile, group)
  int c_in, c_out, T, t_out, k, stride, ntiles, nchunk, span, xs_ld;
};

__global__ void __launch_bounds__(THREADS, 2) dsconv_tc_kernel(const Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  // map: A hi | A lo | W hi | W lo (this chunk) | xs [KC][xs_ld] | dw [c_in * k] | colsum [4 warps][c_out][2] | barriers
  uint8_t* sAh = smem;
  uint8_t* sAl = smem + A_BYTES;
  uint8_t* sW = smem + 2 * A_BYTES;
  const int w_half = 8 * a.c_out * 16;                   // bytes of one half (hi or lo) of a weight chunk
  float* xs0 = reinterpret_cast<float*>(sW + 2 * w_half);  // two window buffers: chunk c + 1 streams in (cp.async) under chunk c
  float* sdw = xs0 + 2 * KC * a.xs_ld;
  float* sbias = sdw + ((a.c_in * a.k + 3) & ~3);         // [c_out]
  float* sgrp = sbias + a.c_out;                          // [4 row quarters][8 groups][2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sgrp + 64);
  uint64_t* bar_w = bars;
  uint64_t* bar_mma = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, tile = blockIdx.x;
  const int t0 = tile * TM;
  const int pad = a.k / 2;
  const int in0 = t0 * a.stride - pad;                   // first input frame of the window (may be negative)
  const int a0 = (in0 >= 0 ? in0 : in0 - 3) / 4 * 4;     // aligned down to a multiple of 4 (floor for negatives)
  const int off = in0 - a0;                              // 0..3
  const int ngrp = a.xs_ld / 4;
  const bool vec_ok = (a.T & 3) == 0;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  for (int i = tid; i < a.c_in * a.k; i += THREADS) sdw[i] = a.dw[i];
  for (int i = tid; i < a.c_out; i += THREADS) sbias[i] = a.pb[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = idesc_tf32(TM, (uint32_t)a.c_out);
  const float* xb = a.x + (int64_t)b * a.c_in * a.T;
  uint32_t ph_w = 0, ph_m = 0;

  // stage the input window of chunk c's channels into buffer c & 1: 16-byte asynchronous copies along T (x rows are
  // contiguous in T); pieces that touch the padding or a row end are assembled with plain stores
  auto stage = [&](int c) {
    const int ch0 = c * KC, nch = min(KC, a.c_in - ch0);
    float* xs = xs0 + (c & 1) * KC * a.xs_ld;
    for (int i = tid; i < nch * ngrp; i += THREADS) {
      const int ch = i / ngrp, g = i - ch * ngrp;
      const int ti = a0 + 4 * g;
      const float* row = xb + (int64_t)(ch0 + ch) * a.T;
      float* dst = xs + ch * a.xs_ld + 4 * g;
      if (vec_ok && ti >= 0 && ti + 3 < a.T) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(row + ti) : "memory");
      } else {
        float4 v;
        v.x = (ti >= 0 && ti < a.T) ? row[ti] : 0.f;
        v.y = (ti + 1 >= 0 && ti + 1 < a.T) ? row[ti + 1] : 0.f;
        v.z = (ti + 2 >= 0 && ti + 2 < a.T) ? row[ti + 2] : 0.f;
        v.w = (ti + 3 >= 0 && ti + 3 < a.T) ? row[ti + 3] : 0.f;
        *reinterpret_cast<float4*>(dst) = v;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage(0);

  for (int c = 0; c < a.nchunk; ++c) {
    const int ch0 = c * KC;
    const int nch = min(KC, a.c_in - ch0);               // multiple of 8
    const float* xs = xs0 + (c & 1) * KC * a.xs_ld;
    asm volatile("cp.async.wait_group 0;" ::: "memory"); // this thread's pieces of chunk c have landed ...
    __syncthreads();                                     // ... and everybody's; buffer (c + 1) & 1 is no longer read
    if (c + 1 < a.nchunk) stage(c + 1);
    if (c > 0) {                                         // the previous chunk's MMAs have read A and W
      mbar_wait(bar_mma, ph_m);
      ph_m ^= 1;
      tc_fence_after();
    }
    if (tid == 0) {
      mbar_expect_tx(bar_w, 2 * w_half);
      bulk_g2s(sW, a.wimg + (int64_t)c * (2 * w_half / 4), 2 * w_half, bar_w);
    }
    // ---- depthwise taps (fp32) -> A operand, split hi / lo; thread = (frame r, 16 of the 32 channels) ------------
    {
      const int r = tid & (TM - 1), h = tid >> 7;
      for (int q = 0; q < 4; ++q) {                      // 4 channels -> one 16-byte piece of slab (4 h + q)
        const int cl = 16 * h + 4 * q;
        float v[4];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const int ch = cl + j4;
          float s = 0.f;
          if (ch < nch) {
            const float* xr = xs + ch * a.xs_ld + off + r * a.stride;
            const float* wr = sdw + (ch0 + ch) * a.k;
            s = a.k == 3 ? taps<3>(wr, xr, 3) : a.k == 5 ? taps<5>(wr, xr, 5) : taps<0>(wr, xr, a.k);
          }
          v[j4] = s;
        }
        float4 hi, lo;
        hi.x = tf32_rna(v[0]); hi.y = tf32_rna(v[1]); hi.z = tf32_rna(v[2]); hi.w = tf32_rna(v[3]);
        lo.x = v[0] - hi.x; lo.y = v[1] - hi.y; lo.z = v[2] - hi.z; lo.w = v[3] - hi.w;
        const int slab = 4 * h + q;
        *reinterpret_cast<float4*>(sAh + slab * (TM * 16) + r * 16) = hi;
        *reinterpret_cast<float4*>(sAl + slab * (TM * 16) + r * 16) = lo;
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    // ---- pointwise product on the tensor cores: three tf32 MMAs per 8-channel k-step --------------------------------
    if (tid == 0) {
      mbar_wait(bar_w, ph_w);
      tc_fence_after();
      const uint32_t ah = smem_u32(sAh), al = smem_u32(sAl), wh = smem_u32(sW), wl = wh + w_half;
      const int nks = nch / 8;
      for (int ks = 0; ks < nks; ++ks) {
        const uint32_t ao = ks * 2 * (TM * 16), wo = ks * 2 * (a.c_out * 16);
        const uint64_t dah = make_desc(ah + ao, TM * 16, 128), dal = make_desc(al + ao, TM * 16, 128);
        const uint64_t dwh = make_desc(wh + wo, a.c_out * 16, 128), dwl = make_desc(wl + wo, a.c_out * 16, 128);
        umma_tf32(tmem, dah, dwh, idesc, c > 0 || ks > 0);
        umma_tf32(tmem, dal, dwh, idesc, true);
        umma_tf32(tmem, dah, dwl, idesc, true);
      }
      umma_commit(bar_mma);
    }
    ph_w ^= 1;
  }
  mbar_wait(bar_mma, ph_m);
  tc_fence_after();

  // ---- epilogue: + bias, y out (one pass), GroupNorm partials of the tile -------------------------------------------
  // thread = (frame r, half of the channels = 4 of the 8 groups); per group the thread adds its cpg values, the warp its 32
  // rows (shuffles), one thread per group the 4 row quarters in fp64
  {
    const int lq = warp & 3, half = warp >> 2;
    const int r = lq * 32 + lane;
    const int t = t0 + r;
    const bool valid = t < a.t_out;
    const int cpg = a.c_out / 8;                         // a multiple of 4
    const uint32_t trow = tmem + ((uint32_t)(lq * 32) << 16);
    float* yb = a.y + (int64_t)b * a.c_out * a.t_out + t;
    for (int gi = 0; gi < 4; ++gi) {
      const int g = 4 * half + gi;
      float s1 = 0.f, s2 = 0.f;
      for (int c4 = g * cpg; c4 < (g + 1) * cpg; c4 += 4) {
        float v[4];
        tmem_ld4_nw(trow + c4, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float o = v[j] + sbias[c4 + j];
          if (valid) {
            yb[(int64_t)(c4 + j) * a.t_out] = o;
            s1 += o;
            s2 = fmaf(o, o, s2);
          }
        }
      }
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (lane == 0) {
        sgrp[(lq * 8 + g) * 2] = s1;
        sgrp[(lq * 8 + g) * 2 + 1] = s2;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 8) {                                         // one thread per group: fixed summation order, fp64
    double s1 = 0.0, s2 = 0.0;
    for (int w = 0; w < 4; ++w) {
      s1 += (double)sgrp[(w * 8 + tid) * 2];
      s2 += (double)sgrp[(w * 8 + tid) * 2 + 1];
    }
    double* p = a.part + (((int64_t)b * a.ntiles + tile) * 8 + tid) * 2;
    p[0] = s1;
    p[1] = s2;
  }
  if (warp == 0) tmem_dealloc<256>(tmem);
}

// One block per (utterance, group): the group's statistics from the tile partials (one thread, fp64, fixed order), then
// y = gelu((y - mean) rstd w + b) in place over the group's cpg contiguous channel rows, 128-bit accesses.
__global__ void __launch_bounds__(256) dsconv_finish_kernel(float* __restrict__ y, const double* __restrict__ part,
                                                            const float* __restrict__ gw, const float* __restrict__ gb, int c_out,
                                                            int t_out, int ntiles, float eps) {
  __shared__ float s_mean, s_rstd;
  const int cpg = c_out / 8;
  const int64_t b = blockIdx.x >> 3;
  const int g = blockIdx.x & 7;
  if (threadIdx.x == 0) {
    double s1 = 0.0, s2 = 0.0;
    for (int i = 0; i < ntiles; ++i) {
      const double* p = part + ((b * ntiles + i) * 8 + g) * 2;
      s1 += p[0];
      s2 += p[1];
    }
    const double n = (double)cpg * t_out;
    const double mean = s1 / n;
    double var = s2 / n - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean = (float)mean;
    s_rstd = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const float mean = s_mean, rstd = s_rstd;
  float* base = y + (b * c_out + (int64_t)g * cpg) * t_out;
  if ((t_out & 3) == 0) {
    const int n4 = t_out >> 2;
    for (int i = threadIdx.x; i < cpg * n4; i += 256) {
      const int c = g * cpg + i / n4;
      const float sc = rstd * gw[c], sh = gb[c] - mean * sc;
      float4 v = reinterpret_cast<float4*>(base)[i];
      v.x = gelu_erf(fmaf(v.x, sc, sh)); v.y = gelu_erf(fmaf(v.y, sc, sh));
      v.z = gelu_erf(fmaf(v.z, sc, sh)); v.w = gelu_erf(fmaf(v.w, sc, sh));
      reinterpret_cast<float4*>(base)[i] = v;
    }
  } else {
    for (int i = threadIdx.x; i < cpg * t_out; i += 256) {
      const int c = g * cpg + i / t_out;
      const float sc = rstd * gw[c], sh = gb[c] - mean * sc;
      base[i] = gelu_erf(fmaf(base[i], sc, sh));
    }
  }
}

}  // namespace dctc

}  // namespace edtts

using namespace edtts;

static bool dsconv_tc_ok(int c_in, int c_out, int k, int stride) {
  if (c_in % 8 || c_out % 32 || c_out > 256 || k > 15) return false;    // cpg = c_out / 8 a multiple of 4
  const int span = (dctc::TM - 1) * stride + k, xs_ld = ((span + 3 + 3) / 4) * 4;
  const int64_t smem = 2 * dctc::A_BYTES + 2 * 8 * c_out * 16 + 2 * (int64_t)dctc::KC * xs_ld * 4 + ((c_in * k + 3) & ~3) * 4 +
                       (c_out + 64) * 4 + 64;
  return smem <= 112 * 1024;                             // two CTAs per SM
}

extern "C" int64_t edtts_dsconv_workspace_bytes(int32_t B, int32_t c_in, int32_t c_out, int32_t t_out) {
  const int groups = c_out < 8 ? c_out : 8;
  const int64_t stats = align_up((int64_t)B * groups * 2 * 4, 256);
  const int64_t nchunk = (c_in + dctc::KC - 1) / dctc::KC, ntiles = (t_out + dctc::TM - 1) / dctc::TM;
  const int64_t wimg = align_up(nchunk * 2 * 8 * (int64_t)c_out * 16, 256);
  const int64_t part = align_up((int64_t)B * ntiles * 8 * 2 * 8, 256);
  return stats + wimg + part;
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
  EDTTS_REQUIRE(workspace_bytes >= edtts_dsconv_workspace_bytes(B, c_in, c_out, t_out), EDTTS_ENOSPC, "dsconv: workspace");
  cudaStream_t st = as_stream(stream);
  static const bool no_tc = getenv("EDTTS_DSCONV_SIMT") != nullptr;          // development: force the CUDA-core route
  if (!no_tc && dsconv_tc_ok(c_in, c_out, kernel_size, stride)) {
    using namespace dctc;
    Args a;
    a.nchunk = (c_in + KC - 1) / KC;
    a.ntiles = (t_out + TM - 1) / TM;
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace) + align_up((int64_t)B * groups * 2 * 4, 256);
    float* wimg = reinterpret_cast<float*>(ws);
    a.part = reinterpret_cast<double*>(ws + align_up((int64_t)a.nchunk * 2 * 8 * c_out * 16, 256));
    a.x = x; a.dw = dw_w; a.wimg = wimg; a.pb = pw_b; a.y = y_out;
    a.c_in = c_in; a.c_out = c_out; a.T = T; a.t_out = t_out; a.k = kernel_size; a.stride = stride;
    a.span = (TM - 1) * stride + kernel_size;
    a.xs_ld = ((a.span + 3 + 3) / 4) * 4;
    const int smem = 2 * A_BYTES + 2 * 8 * c_out * 16 + 2 * KC * a.xs_ld * 4 + ((c_in * kernel_size + 3) & ~3) * 4 + (c_out + 64) * 4 + 64;
    static PerDeviceOnce configured;
    if (configured.need()) {
      if (cudaFuncSetAttribute(dsconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024) != cudaSuccess)
        return check_launch("dsconv_tc smem attribute");
      configured.set();
    }
    {
      LaunchScope ls(KC_DSCONV, st);
      if (int rc = pack_w_tf32(pw_w, wimg, c_in, c_out, c_in, st)) return rc;
    }
    {
      LaunchScope ls(KC_DSCONV, st);
      dsconv_tc_kernel<<<dim3(a.ntiles, B), THREADS, smem, st>>>(a);
      if (int rc = check_launch("dsconv_tc")) return rc;
    }
    LaunchScope ls(KC_DSCONV, st);
    dsconv_finish_kernel<<<(unsigned)(B * 8), 256, 0, st>>>(y_out, a.part, gn_w, gn_b, c_out, t_out, a.ntiles, 1e-5f);
    return check_launch("dsconv_finish");
  }
  const int span = (DC_TT - 1) * stride + kernel_size;
  const size_t smem = ((size_t)c_in * span + (size_t)c_in * DC_TT) * 4;
  EDTTS_REQUIRE(smem <= 227 * 1024, EDTTS_ENOTSUP, "dsconv: C_in=%d stride=%d k=%d needs %zu B of shared memory", c_in,
                stride, kernel_size, smem);
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
