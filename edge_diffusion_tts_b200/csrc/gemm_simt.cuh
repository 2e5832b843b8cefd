// fp32 CUDA-core GEMM with fused normalisation prologues and epilogues.
// This is the EDTTS_PREC_FP32 (parity, 1e-4) arithmetic of every Linear on the
// path: y[rows,N] = pro(A)[rows,K] @ W[N,K]^T, W exactly as nn.Linear stores it.
//
// Block tile 128 rows x (16*TN) columns, K sliced by 16, 256 threads each owning
// an 8 x TN register tile.  A and W slices are staged transposed ([k][row]) in
// shared memory so the inner product reads them with 128-bit / conflict-free
// loads.  A row's result never depends on which other rows share the launch
// (batch invariance: the multi-GPU batch shards reproduce the 1-GPU bits).
#pragma once
#include "common.cuh"

namespace edtts {

enum GemmPro : int { PRO_NONE = 0, PRO_RMS = 1, PRO_ADARMS = 2, PRO_LN = 3 };
enum GemmEpi : int { EPI_STORE = 0, EPI_GELU = 1, EPI_RESID = 2, EPI_PE = 3, EPI_SWIGLU = 4, EPI_STEP = 5 };

struct GemmArgs {
  const float* A = nullptr;  int64_t rows = 0;  int K = 0;  int lda = 0;
  const float* W = nullptr;  int N = 0;          // EPI_SWIGLU: W has 2N rows (x rows, then gate rows)
  const float* bias = nullptr;                   // [N] ([2N] for EPI_SWIGLU) or null
  float* out = nullptr;      int ldo = 0;
  int pro = PRO_NONE;        int epi = EPI_STORE;
  // prologue: RMSNorm (mla.py:53-58), AdaRMSNorm (transformer.py:64-68), LayerNorm
  const float* norm_w = nullptr;  const float* norm_b = nullptr;  float norm_eps = 1e-6f;
  const float* mod = nullptr;     int mod_stride = 0;             // scale = mod[b*stride + k], shift = +K
  int rows_per_batch = 1;                                         // b = row / rows_per_batch
  // epilogue
  const float* resid = nullptr;                                   // EPI_RESID (may alias out)
  const float* pe = nullptr;      int pe_period = 1;              // EPI_PE: + pe[(row % period)*N + col]
  const float* x_t = nullptr;     edtts_step_args step{};         // EPI_STEP
};

constexpr int G_BM = 128, G_BK = 16, G_THREADS = 256;
constexpr int G_ASLD = G_BM + 4;

#ifndef EDTTS_DECL_ONLY
template <int TN, bool DUAL>
__global__ void __launch_bounds__(G_THREADS) gemm_simt_kernel(const GemmArgs g) {
  constexpr int BN = 16 * TN;
  constexpr int WSLD = BN + 4;
  __shared__ __align__(16) float As[G_BK * G_ASLD];
  __shared__ __align__(16) float Ws[G_BK * WSLD];
  __shared__ __align__(16) float Wg[DUAL ? G_BK * WSLD : 4];
  __shared__ float s_rstd[G_BM];
  __shared__ float s_mean[G_BM];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)blockIdx.x * G_BM;
  const int n0 = blockIdx.y * BN;
  const int K = g.K;

  // ---- per-row statistics for the normalisation prologues (K <= 160) ----------
  if (g.pro != PRO_NONE) {
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < G_BM; r += G_THREADS / 32) {
      const int64_t row = row0 + r;
      float v[6];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int k = lane + 32 * i;
        v[i] = (row < g.rows && k < K) ? g.A[row * g.lda + k] : 0.f;
        s += (g.pro == PRO_LN) ? v[i] : v[i] * v[i];
      }
      s = warp_sum(s);
      float mean = 0.f, var;
      if (g.pro == PRO_LN) {
        mean = s / (float)K;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const int k = lane + 32 * i;
          const float d = (k < K) ? v[i] - mean : 0.f;
          q += d * d;
        }
        var = warp_sum(q) / (float)K;
      } else {
        var = s / (float)K;
      }
      if (lane == 0) {
        s_rstd[r] = 1.0f / sqrtf(var + g.norm_eps);
        s_mean[r] = mean;
      }
    }
    __syncthreads();
  }

  float acc[8][TN];
  float acc2[DUAL ? 8 : 1][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      acc[i][j] = 0.f;
      if (DUAL) acc2[i][j] = 0.f;
    }

  for (int k0 = 0; k0 < K; k0 += G_BK) {
    // ---- A slice: 128 rows x 16 k, prologue applied on the way in ------------
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int idx = tid + it * G_THREADS;
      const int r = idx >> 2, kq = (idx & 3) * 4;
      const int64_t row = row0 + r;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < g.rows) a = *reinterpret_cast<const float4*>(g.A + row * g.lda + k0 + kq);
      float av[4] = {a.x, a.y, a.z, a.w};
      if (g.pro != PRO_NONE) {
        const float rstd = s_rstd[r], mean = s_mean[r];
        const int b = (int)(row / g.rows_per_batch);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = k0 + kq + j;
          float x = av[j];
          if (g.pro == PRO_LN) {
            x = (x - mean) * rstd * g.norm_w[k] + g.norm_b[k];
          } else {
            x = (x * rstd) * g.norm_w[k];
            if (g.pro == PRO_ADARMS && row < g.rows) {
              const float* m = g.mod + (int64_t)b * g.mod_stride;
              x = x * (1.0f + m[k]) + m[K + k];
            }
          }
          av[j] = (row < g.rows) ? x : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) As[(kq + j) * G_ASLD + r] = av[j];
    }
    // ---- W slice: BN rows x 16 k ------------------------------------------
    for (int idx = tid; idx < BN * 4; idx += G_THREADS) {
      const int n = idx >> 2, kq = (idx & 3) * 4;
      const float4 w = *reinterpret_cast<const float4*>(g.W + (int64_t)(n0 + n) * K + k0 + kq);
      Ws[(kq + 0) * WSLD + n] = w.x;
      Ws[(kq + 1) * WSLD + n] = w.y;
      Ws[(kq + 2) * WSLD + n] = w.z;
      Ws[(kq + 3) * WSLD + n] = w.w;
      if (DUAL) {
        const float4 u = *reinterpret_cast<const float4*>(g.W + (int64_t)(g.N + n0 + n) * K + k0 + kq);
        Wg[(kq + 0) * WSLD + n] = u.x;
        Wg[(kq + 1) * WSLD + n] = u.y;
        Wg[(kq + 2) * WSLD + n] = u.z;
        Wg[(kq + 3) * WSLD + n] = u.w;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < G_BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk * G_ASLD + ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk * G_ASLD + ty * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN], b2[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        b[j] = Ws[kk * WSLD + tx * TN + j];
        if (DUAL) b2[j] = Wg[kk * WSLD + tx * TN + j];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
          if (DUAL) acc2[i][j] = fmaf(a[i], b2[j], acc2[i][j]);
        }
    }
    __syncthreads();
  }

  // ---- epilogue -----------------------------------------------------------
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t row = row0 + ty * 8 + i;
    if (row >= g.rows) continue;
    float ab_t = 0.f, ab_p = 1.f, al = 0.f, be = 0.f, pv = 0.f, nzm = 0.f;
    if (g.epi == EPI_STEP && (g.step.mode == EDTTS_STEP_DDIM || g.step.mode == EDTTS_STEP_DDPM)) {
      const int b = (int)(row / g.rows_per_batch);
      const int64_t t = g.step.t[b];
      ab_t = g.step.alpha_bar[t];
      if (g.step.mode == EDTTS_STEP_DDIM) {
        const int64_t tp = g.step.t_prev[b];
        ab_p = (tp >= 0) ? g.step.alpha_bar[tp] : 1.0f;
      } else {
        al = g.step.alphas[t];
        be = g.step.betas[t];
        pv = g.step.posterior_var[t];
        nzm = (t > 0) ? 1.0f : 0.0f;
      }
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int col = n0 + tx * TN + j;
      float v = acc[i][j];
      if (g.bias) v += g.bias[col];
      const int64_t o = row * g.ldo + col;
      switch (g.epi) {
        case EPI_GELU: v = gelu_erf(v); break;
        case EPI_RESID: v = g.resid[o] + v; break;
        case EPI_PE: v += g.pe[(int64_t)(row % g.pe_period) * g.N + col]; break;
        case EPI_SWIGLU: {
          float gate = DUAL ? acc2[i][j] : 0.f;
          if (g.bias) gate += g.bias[g.N + col];
          v = v * silu(gate);
        } break;
        default: break;
      }
      if (g.epi == EPI_STEP) {
        if (g.step.eps_out) g.step.eps_out[o] = v;
        if (g.step.mode == EDTTS_STEP_DDIM) {
          float xp, x0;
          ddim_update(g.x_t[o], v, 0.f, ab_t, ab_p, 0.f, xp, x0);
          if (g.step.x0_out) g.step.x0_out[o] = x0;
          if (g.step.write_x_prev && g.step.x_prev_out) g.step.x_prev_out[o] = xp;
        } else if (g.step.mode == EDTTS_STEP_DDPM) {
          g.step.x_prev_out[o] = ddpm_update(g.x_t[o], v, g.step.noise[o], al, ab_t, be, pv, nzm);
        } else if (g.step.mode == EDTTS_STEP_DPM) {
          const int ord = g.step.dpm_order;
          float xp, x0;
          dpm_update(g.x_t[o], v, ord >= 2 ? g.step.dpm_hist1[o] : 0.f, ord >= 3 ? g.step.dpm_hist2[o] : 0.f,
                     g.step.dpm_coef + (row / g.rows_per_batch) * 8, ord, g.step.dpm_predict_x0 ? 1 : 0, xp, x0);
          if (g.step.x0_out) g.step.x0_out[o] = x0;
          g.step.x_prev_out[o] = xp;
        }
      } else {
        g.out[o] = v;
      }
    }
  }
}

#endif  // EDTTS_DECL_ONLY

int launch_gemm_simt(const GemmArgs& g, cudaStream_t stream);

}  // namespace edtts
