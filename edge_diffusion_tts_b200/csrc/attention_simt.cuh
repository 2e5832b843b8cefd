// fp32 CUDA-core attention for the parity path (EDTTS_PREC_FP32).
//
// One kernel serves both attentions of a DiffusionTransformerBlock:
//   window >= 0 : EfficientAttention's banded SDPA, |i-j| <= window
//                 (layers/attention.py:94-111) -- computed over the band only,
//                 the dense [T,T] mask of the reference is never built;
//   window <  0 : MultiHeadLatentAttention's full SDPA over the S context tokens
//                 (layers/mla.py:176-180).
// Thread <-> query: each thread keeps its query row, the running (max, sum) and the
// 40-wide output accumulator in registers and walks the keys of a shared-memory
// chunk, reading K/V rows as warp-broadcast 128-bit loads.  No shuffles, no score
// matrix in memory; keys are consumed 8 at a time so the online-softmax rescale is
// amortised.  Results for a row do not depend on the other rows in the launch.
#pragma once
#include "common.cuh"

namespace edtts {

struct AttnArgs {
  const float* q;  int q_stride;      // q row (b*Tq + i) at q + row*q_stride, head h at + h*40
  const float* k;  const float* v;  int kv_stride;
  float* o;        int o_stride;
  int Tq, Tk, window;
  float scale;
};

constexpr int AT_Q = 128;   // queries per block (= threads)
constexpr int AT_KC = 64;   // keys per shared-memory chunk

#ifndef EDTTS_DECL_ONLY
__global__ void __launch_bounds__(AT_Q) attn_simt_kernel(const AttnArgs a) {
  __shared__ __align__(16) float Ks[AT_KC * HD];
  __shared__ __align__(16) float Vs[AT_KC * HD];
  const int b = blockIdx.z, h = blockIdx.y;
  const int q0 = blockIdx.x * AT_Q;
  const int qi = q0 + threadIdx.x;
  const bool active = qi < a.Tq;
  const int W = a.window;

  float q[HD], acc[HD];
  {
    const float* qp = a.q + ((int64_t)b * a.Tq + (active ? qi : 0)) * a.q_stride + h * HD;
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
      const float4 t = *reinterpret_cast<const float4*>(qp + d);
      q[d] = t.x; q[d + 1] = t.y; q[d + 2] = t.z; q[d + 3] = t.w;
    }
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = 0.f;
  }
  float m = -INFINITY, l = 0.f;

  int klo = 0, khi = a.Tk;
  if (W >= 0) {
    klo = max(0, q0 - W);
    khi = min(a.Tk, q0 + AT_Q - 1 + W + 1);
  }
  // keys needed by this warp's 32 queries
  const int wq0 = q0 + (threadIdx.x & ~31);
  const int wlo = (W >= 0) ? wq0 - W : 0;
  const int whi = (W >= 0) ? wq0 + 31 + W : a.Tk - 1;

  const float* kbase = a.k + (int64_t)b * a.Tk * a.kv_stride + h * HD;
  const float* vbase = a.v + (int64_t)b * a.Tk * a.kv_stride + h * HD;

  for (int kc = klo; kc < khi; kc += AT_KC) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < AT_KC * (HD / 4); idx += AT_Q) {
      const int r = idx / (HD / 4), c = (idx % (HD / 4)) * 4;
      float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
      if (kc + r < khi) {
        kk = *reinterpret_cast<const float4*>(kbase + (int64_t)(kc + r) * a.kv_stride + c);
        vv = *reinterpret_cast<const float4*>(vbase + (int64_t)(kc + r) * a.kv_stride + c);
      }
      *reinterpret_cast<float4*>(&Ks[r * HD + c]) = kk;
      *reinterpret_cast<float4*>(&Vs[r * HD + c]) = vv;
    }
    __syncthreads();
    if (kc > whi || kc + AT_KC - 1 < wlo) continue;   // warp-uniform: chunk outside this warp's band
    for (int j0 = 0; j0 < AT_KC; j0 += 8) {
      const int key0 = kc + j0;
      if (key0 >= khi) break;
      if (key0 > whi || key0 + 7 < wlo) continue;      // warp-uniform
      float s[8];
      float gmax = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float* kr = &Ks[(j0 + jj) * HD];
        float dot = 0.f;
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
          const float4 t = *reinterpret_cast<const float4*>(kr + d);
          dot = fmaf(q[d], t.x, dot);
          dot = fmaf(q[d + 1], t.y, dot);
          dot = fmaf(q[d + 2], t.z, dot);
          dot = fmaf(q[d + 3], t.w, dot);
        }
        const int key = key0 + jj;
        const bool ok = active && key < khi && (W < 0 || (key >= qi - W && key <= qi + W));
        s[jj] = ok ? dot * a.scale : -INFINITY;
        gmax = fmaxf(gmax, s[jj]);
      }
      const float m_new = fmaxf(m, gmax);
      const float corr = (m_new == -INFINITY) ? 1.0f : expf(m - m_new);
      l *= corr;
#pragma unroll
      for (int d = 0; d < HD; ++d) acc[d] *= corr;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float p = (s[jj] == -INFINITY) ? 0.f : expf(s[jj] - m_new);
        l += p;
        const float* vr = &Vs[(j0 + jj) * HD];
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
          const float4 t = *reinterpret_cast<const float4*>(vr + d);
          acc[d] = fmaf(p, t.x, acc[d]);
          acc[d + 1] = fmaf(p, t.y, acc[d + 1]);
          acc[d + 2] = fmaf(p, t.z, acc[d + 2]);
          acc[d + 3] = fmaf(p, t.w, acc[d + 3]);
        }
      }
      m = m_new;
    }
  }
  if (active) {
    const float inv = 1.0f / l;
    float* op = a.o + ((int64_t)b * a.Tq + qi) * a.o_stride + h * HD;
#pragma unroll
    for (int d = 0; d < HD; d += 4)
      *reinterpret_cast<float4*>(op + d) = make_float4(acc[d] * inv, acc[d + 1] * inv, acc[d + 2] * inv, acc[d + 3] * inv);
  }
}

#endif  // EDTTS_DECL_ONLY

int launch_attn_simt(const AttnArgs& a, int B, cudaStream_t stream);

}  // namespace edtts
