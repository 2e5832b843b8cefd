// Persistent tcgen05 GEMM for the Linear layers of the decoder (EDTTS_PREC_BF16):
//   D[128 x N_CTA] (TMEM, fp32) = A[128 x K] (smem, bf16) * W[N_CTA x K]^T (smem, bf16)
//
//   * the CTA's weight slab (N_CTA x K bf16, pre-packed in the UMMA operand image by
//     edtts_pack_weights_bf16) is fetched ONCE with a single bulk async copy and stays
//     resident in shared memory while the CTA walks 128-row tiles of the activations;
//   * the A tile comes either from the fp32 residual stream with the normalisation fused
//     into the load (RMSNorm / AdaRMSNorm / LayerNorm prologue: warp per row, coalesced
//     reads, bf16 written straight into the operand image), or from a bf16 chunk-major
//     activation written by a previous kernel (one bulk copy per K-slab, no register pass;
//     the next tile's copies are issued as soon as the MMAs of the current one retire);
//   * one elected thread issues K/16 tcgen05.mma (M=128, N=N_CTA) into TMEM and commits to
//     an mbarrier; the 4 warps then drain TMEM with tcgen05.ld (thread <-> row) and run the
//     fused epilogue: bias, SwiGLU, residual add, positional table, or the DDIM/DDPM update.
//
// Column blocks (blockIdx.y) let one Linear be split so that weights + A tile fit two CTAs
// per SM (51 KB + 40 KB for K=160, N_CTA=160): the second CTA's loads overlap the first
// one's MMA/epilogue without an intra-CTA pipeline.
#pragma once
#include "umma.cuh"
#include "gemm_simt.cuh"   // GemmPro enum (prologue kinds)

namespace edtts {
namespace tc {

enum AMode : int { A_F32 = 0, A_CHUNK = 1 };
enum TcEpi : int { TE_CHUNK = 0, TE_SWIGLU = 1, TE_RESID = 2, TE_PE = 3, TE_STEP = 4, TE_F32 = 5 };

struct TcGemmArgs {
  int amode = A_F32;
  const float* A_f32 = nullptr;  int lda = 0;        // [R][lda] fp32
  const __nv_bfloat16* A_chunk = nullptr;            // [K/8][R][8] bf16
  int64_t R = 0;  int T = 1;                         // rows, rows per utterance
  const __nv_bfloat16* W_img = nullptr;              // [ny][K/8][N_CTA][8] bf16
  const float* bias = nullptr;                       // packed column order, [ny*N_CTA] or null
  int pro = PRO_NONE;  const float* norm_w = nullptr;  const float* norm_b = nullptr;  float norm_eps = 1e-6f;
  const float* mod = nullptr;  int mod_stride = 0;
  int epi = TE_F32;
  float* out_f32 = nullptr;  int ldo = 0;            // TE_RESID / TE_PE / TE_F32 (row-major fp32)
  __nv_bfloat16* out_chunk = nullptr;                // TE_CHUNK / TE_SWIGLU: [cols/8][R][8] bf16
  const float* pe = nullptr;                         // TE_PE: [T][ldo]
  const float* x_t = nullptr;  edtts_step_args step{};   // TE_STEP
};

constexpr int TILE_M = 128;
constexpr int TC_THREADS = 128;
constexpr int STAGE_LD = 33;

template <int K, int N_CTA>
struct TcGemmSmem {
  static constexpr int W_BYTES = K * N_CTA * 2;
  static constexpr int A_BYTES = K * TILE_M * 2;
  static constexpr int STAGE_BYTES = 4 * 32 * STAGE_LD * 4;
  static constexpr int OFF_A = W_BYTES;
  static constexpr int OFF_STAGE = OFF_A + A_BYTES;
  static constexpr int OFF_BAR = OFF_STAGE + STAGE_BYTES;
  static constexpr int TOTAL = OFF_BAR + 64;
  static constexpr uint32_t TMEM_COLS = N_CTA <= 32 ? 32 : N_CTA <= 64 ? 64 : N_CTA <= 128 ? 128 : 256;
};

template <int K, int N_CTA>
__global__ void __launch_bounds__(TC_THREADS) tc_gemm_kernel(const TcGemmArgs g) {
  using L = TcGemmSmem<K, N_CTA>;
  static_assert(K % 16 == 0 && N_CTA % 16 == 0 && N_CTA <= 256, "UMMA shape");
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sA = smem + L::OFF_A;
  float* stage = reinterpret_cast<float*>(smem + L::OFF_STAGE);
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* bar_a = bar_w + 1;
  uint64_t* bar_mma = bar_w + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 3);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.y * N_CTA;
  const int64_t ntiles = (g.R + TILE_M - 1) / TILE_M;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_a, 1);
    mbar_init(bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<L::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto issue_a_copy = [&](int64_t tile) {   // thread 0 only, A_CHUNK mode
    const int64_t row0 = tile * TILE_M;
    const uint32_t valid = (uint32_t)min((int64_t)TILE_M, g.R - row0);
    mbar_expect_tx(bar_a, (K / 8) * valid * 16);
#pragma unroll 1
    for (int c = 0; c < K / 8; ++c)
      bulk_g2s(sA + c * (TILE_M * 16), g.A_chunk + ((int64_t)c * g.R + row0) * 8, valid * 16, bar_a);
  };

  if (tid == 0) {
    mbar_expect_tx(bar_w, L::W_BYTES);
    bulk_g2s(sW, g.W_img + (int64_t)blockIdx.y * K * N_CTA, L::W_BYTES, bar_w);
    if (g.amode == A_CHUNK && blockIdx.x < ntiles) issue_a_copy(blockIdx.x);
  }

  // per-lane constants of the fp32 prologue: this lane owns k = 64*i + 2*lane (+1)
  constexpr int NI = (K + 63) / 64;
  float2 nw[NI], nb[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int k = 64 * i + 2 * lane;
    nw[i] = make_float2(1.f, 1.f);
    nb[i] = make_float2(0.f, 0.f);
    if (g.amode == A_F32 && g.pro != PRO_NONE && k < K) {
      nw[i] = *reinterpret_cast<const float2*>(g.norm_w + k);
      if (g.pro == PRO_LN) nb[i] = *reinterpret_cast<const float2*>(g.norm_b + k);
    }
  }

  constexpr uint32_t IDESC = make_idesc(TILE_M, N_CTA);
  uint32_t phase_a = 0, phase_mma = 0;
  bool w_ready = false;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * TILE_M;

    if (g.amode == A_F32) {
      // ---- fused normalisation prologue: warp per row, fp32 -> bf16 operand image --------
#pragma unroll 1
      for (int rr = 0; rr < 32; ++rr) {
        const int rl = warp * 32 + rr;
        const int64_t row = row0 + rl;
        float2 v[NI];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const int k = 64 * i + 2 * lane;
          v[i] = make_float2(0.f, 0.f);
          if (row < g.R && k < K) v[i] = *reinterpret_cast<const float2*>(g.A_f32 + row * g.lda + k);
          s += (g.pro == PRO_LN) ? (v[i].x + v[i].y) : (v[i].x * v[i].x + v[i].y * v[i].y);
        }
        float mean = 0.f, rstd = 1.f;
        if (g.pro != PRO_NONE) {
          s = warp_sum(s);
          if (g.pro == PRO_LN) {
            mean = s / (float)K;
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < NI; ++i) {
              const int k = 64 * i + 2 * lane;
              if (k < K) {
                const float dx = v[i].x - mean, dy = v[i].y - mean;
                q += dx * dx + dy * dy;
              }
            }
            rstd = 1.0f / sqrtf(warp_sum(q) / (float)K + g.norm_eps);
          } else {
            rstd = 1.0f / sqrtf(s / (float)K + g.norm_eps);
          }
        }
        const float* m = (g.pro == PRO_ADARMS && row < g.R) ? g.mod + (row / g.T) * g.mod_stride : nullptr;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const int k = 64 * i + 2 * lane;
          if (k < K) {
            float x = v[i].x, y = v[i].y;
            if (g.pro == PRO_LN) {
              x = (x - mean) * rstd * nw[i].x + nb[i].x;
              y = (y - mean) * rstd * nw[i].y + nb[i].y;
            } else if (g.pro != PRO_NONE) {
              x = (x * rstd) * nw[i].x;
              y = (y * rstd) * nw[i].y;
              if (m) {
                const float2 sc = *reinterpret_cast<const float2*>(m + k);
                const float2 sh = *reinterpret_cast<const float2*>(m + K + k);
                x = x * (1.0f + sc.x) + sh.x;
                y = y * (1.0f + sc.y) + sh.y;
              }
            }
            *reinterpret_cast<uint32_t*>(sA + (k >> 3) * (TILE_M * 16) + rl * 16 + (k & 7) * 2) = pack_bf16x2(x, y);
          }
        }
      }
      fence_proxy_async();
    }
    __syncthreads();

    if (tid == 0) {
      if (!w_ready) {
        mbar_wait(bar_w, 0);
        w_ready = true;
      }
      if (g.amode == A_CHUNK) mbar_wait(bar_a, phase_a);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(sA), w_addr = smem_u32(sW);
#pragma unroll
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint64_t ad = make_desc(a_addr + ks * 2 * (TILE_M * 16), TILE_M * 16, 128);
        const uint64_t bd = make_desc(w_addr + ks * 2 * (N_CTA * 16), N_CTA * 16, 128);
        umma_bf16(tmem_base, ad, bd, IDESC, ks > 0);
      }
      umma_commit(bar_mma);
    }
    phase_a ^= 1;
    mbar_wait(bar_mma, phase_mma);
    phase_mma ^= 1;
    tc_fence_after();
    // A tile is free again: prefetch the next one while the epilogue runs
    if (tid == 0 && g.amode == A_CHUNK && tile + gridDim.x < ntiles) issue_a_copy(tile + gridDim.x);

    // ---- epilogue: TMEM -> registers -> global ------------------------------------
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int64_t row = row0 + tid;
    if (g.epi == TE_CHUNK) {
#pragma unroll 1
      for (int c0 = 0; c0 < N_CTA; c0 += 16) {
        float v[16];
        tmem_ld16(trow + c0, v);
        if (g.bias) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += g.bias[n0 + c0 + j];
        }
        if (row < g.R) {
          __nv_bfloat16* o = g.out_chunk + ((int64_t)((n0 + c0) >> 3) * g.R + row) * 8;
          *reinterpret_cast<uint4*>(o) = pack_bf16x8(v);
          *reinterpret_cast<uint4*>(o + g.R * 8) = pack_bf16x8(v + 8);
        }
      }
    } else if (g.epi == TE_SWIGLU) {
      constexpr int NU = N_CTA / 2;
#pragma unroll 1
      for (int c0 = 0; c0 < NU; c0 += 16) {
        float a[16], gt[16];
        tmem_ld16(trow + c0, a);
        tmem_ld16(trow + NU + c0, gt);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float xv = a[j] + (g.bias ? g.bias[n0 + c0 + j] : 0.f);
          const float gv = gt[j] + (g.bias ? g.bias[n0 + NU + c0 + j] : 0.f);
          a[j] = xv * silu(gv);
        }
        if (row < g.R) {
          __nv_bfloat16* o = g.out_chunk + ((int64_t)((blockIdx.y * NU + c0) >> 3) * g.R + row) * 8;
          *reinterpret_cast<uint4*>(o) = pack_bf16x8(a);
          *reinterpret_cast<uint4*>(o + g.R * 8) = pack_bf16x8(a + 8);
        }
      }
    } else {
      // fp32 row-major outputs: transpose 32x32 blocks through a per-warp staging tile so that
      // global accesses are 128-byte rows instead of one 16-byte piece per thread row
      float* st = stage + warp * 32 * STAGE_LD;
#pragma unroll 1
      for (int c0 = 0; c0 < N_CTA; c0 += 32) {
        float v[32];
        tmem_ld16(trow + c0, v);
        tmem_ld16(trow + c0 + 16, v + 16);
#pragma unroll
        for (int j = 0; j < 32; ++j) st[lane * STAGE_LD + j] = v[j];
        __syncwarp();
        const int col = c0 + lane;
        const bool col_ok = col < N_CTA;
        const float bv = (g.bias && col_ok) ? g.bias[n0 + col] : 0.f;
#pragma unroll 1
        for (int rr = 0; rr < 32; ++rr) {
          const int64_t r = row0 + warp * 32 + rr;
          if (r >= g.R || !col_ok) continue;
          float val = st[rr * STAGE_LD + lane] + bv;
          const int64_t o = r * g.ldo + n0 + col;
          if (g.epi == TE_RESID) {
            g.out_f32[o] += val;
          } else if (g.epi == TE_PE) {
            g.out_f32[o] = val + g.pe[(r % g.T) * g.ldo + n0 + col];
          } else if (g.epi == TE_F32) {
            g.out_f32[o] = val;
          } else {   // TE_STEP
            const edtts_step_args& s = g.step;
            if (s.eps_out) s.eps_out[o] = val;
            if (s.mode == EDTTS_STEP_DDIM) {
              const int b = (int)(r / g.T);
              const float ab_t = s.alpha_bar[s.t[b]];
              const int64_t tp = s.t_prev[b];
              const float ab_p = tp >= 0 ? s.alpha_bar[tp] : 1.0f;
              float xp, x0;
              ddim_update(g.x_t[o], val, 0.f, ab_t, ab_p, 0.f, xp, x0);
              if (s.x0_out) s.x0_out[o] = x0;
              if (s.write_x_prev && s.x_prev_out) s.x_prev_out[o] = xp;
            } else if (s.mode == EDTTS_STEP_DDPM) {
              const int b = (int)(r / g.T);
              const int64_t tt = s.t[b];
              s.x_prev_out[o] = ddpm_update(g.x_t[o], val, s.noise[o], s.alphas[tt], s.alpha_bar[tt], s.betas[tt],
                                            s.posterior_var[tt], tt > 0 ? 1.0f : 0.0f);
            }
          }
        }
        __syncwarp();
      }
    }
    tc_fence_before();
    __syncthreads();   // TMEM drained and staging free before the next tile's MMA / prologue
  }

  __syncthreads();
  if (warp == 0) tmem_dealloc<L::TMEM_COLS>(tmem_base);
}

}  // namespace tc
}  // namespace edtts
