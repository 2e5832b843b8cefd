// Persistent tcgen05 GEMM for the Linear layers of the decoder (EDTTS_PREC_BF16):
//   D[128 x N_CTA] (TMEM, fp32) = A[128 x K] (smem, bf16) * W[N_CTA x K]^T (smem, bf16)
//
// One CTA per SM.  Every byte that crosses HBM moves through the TMA engine where the layout
// allows it, so no thread ever waits on a global load inside the tile loop:
//   * W slab (N_CTA x K bf16, pre-packed UMMA operand image): ONE bulk copy per CTA, resident
//     in shared memory for the whole launch;
//   * A tile, bf16 chunk-major activation of a previous kernel: one bulk copy per K-slab straight
//     into the operand image (prefetched for tile i+1 as soon as the MMAs of tile i retire);
//   * A tile, fp32 residual stream: one bulk copy of the contiguous 128 x K fp32 block into a
//     staging buffer; the fused RMSNorm / AdaRMSNorm / LayerNorm prologue then runs shared->shared
//     (warp per row) and writes bf16 into the operand image; the next tile's block is already in
//     flight while the tensor core and the epilogue work on the current one;
//   * residual epilogue (h += D + bias): TMEM -> registers -> padded fp32 staging tile, then
//     per-row cp.reduce.async.bulk .add.f32: the add happens at L2, nothing is read back;
//   * bf16 chunk-major outputs (q/k/v, SwiGLU u): 16-byte stores, 512 contiguous bytes per warp.
// One elected thread issues the tcgen05.mma chain (M=128, N<=256, K/16 steps) and commits to an
// mbarrier; 8 warps drain TMEM with tcgen05.ld (warp w: lanes 32*(w%4).., column chunks of parity w/4).
#pragma once
#include "umma.cuh"
#include "gemm_simt.cuh"   // GemmPro enum (prologue kinds)

namespace edtts {
namespace tc {

enum AMode : int { A_F32 = 0, A_CHUNK = 1 };
enum TcEpi : int { TE_CHUNK = 0, TE_SWIGLU = 1, TE_RESID = 2, TE_PE = 3, TE_STEP = 4, TE_F32 = 5 };

struct TcGemmArgs {
  int amode = A_F32;
  const float* A_f32 = nullptr;                      // [R][K] fp32, contiguous rows (lda == K)
  const __nv_bfloat16* A_chunk = nullptr;            // [K/8][R][8] bf16
  int64_t R = 0;  int T = 1;                         // rows, rows per utterance
  const __nv_bfloat16* W_img = nullptr;              // [ny][K/8][N_CTA][8] bf16
  const float* bias = nullptr;                       // packed column order, [ny*N_CTA] or null
  int pro = PRO_NONE;  const float* norm_w = nullptr;  const float* norm_b = nullptr;  float norm_eps = 1e-6f;
  const float* mod = nullptr;  int mod_stride = 0;
  int epi = TE_F32;
  float* out_f32 = nullptr;  int ldo = 0;            // TE_RESID / TE_PE / TE_F32 (row-major fp32)
  __nv_bfloat16* out_chunk = nullptr;                // TE_CHUNK / TE_SWIGLU: [cols/8][R][8] bf16
  int f16_from_chunk = 1 << 30;                      // TE_CHUNK: chunks >= this are written as f16 (attention V operand)
  const float* pe = nullptr;                         // TE_PE: [T][ldo]
  const float* x_t = nullptr;  edtts_step_args step{};   // TE_STEP
};

constexpr int TILE_M = 128;
constexpr int TC_THREADS = 256;

template <int K, int N_CTA>
struct TcGemmSmem {
  static constexpr int W_BYTES = K * N_CTA * 2;
  static constexpr int A_BYTES = K * TILE_M * 2;
  static constexpr int IN_BYTES = TILE_M * K * 4;                // fp32 input block (A_F32)
  static constexpr int OUT_LD = N_CTA + 4;                       // padded: conflict-free 16-byte row stores
  static constexpr int OUT_BYTES = TILE_M * OUT_LD * 4;          // fp32 output staging (TE_RESID)
  static constexpr int OFF_A = W_BYTES;
  static constexpr int OFF_IO = OFF_A + A_BYTES;
  static constexpr uint32_t TMEM_COLS = N_CTA <= 32 ? 32 : N_CTA <= 64 ? 64 : N_CTA <= 128 ? 128 : N_CTA <= 256 ? 256 : 512;
  static constexpr int N_INSTR = N_CTA <= 256 ? N_CTA : N_CTA / 2;   // columns per tcgen05.mma
  static int total(int amode, int epi) {
    int io = 0;
    if (amode == A_F32) io = IN_BYTES;
    if (epi == TE_RESID && OUT_BYTES > io) io = OUT_BYTES;
    return OFF_IO + io + 64;
  }
};

template <int K, int N_CTA>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_gemm_kernel(const TcGemmArgs g, const int io_bytes) {
  using L = TcGemmSmem<K, N_CTA>;
  static_assert(K % 16 == 0 && L::N_INSTR % 16 == 0 && L::N_INSTR <= 256, "UMMA shape");
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sA = smem + L::OFF_A;
  float* sIO = reinterpret_cast<float*>(smem + L::OFF_IO);
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + L::OFF_IO + io_bytes);
  uint64_t* bar_a = bar_w + 1;
  uint64_t* bar_mma = bar_w + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 3);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lgroup = warp & 3, half = warp >> 2;
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

  auto issue_input = [&](int64_t tile) {   // thread 0 only
    const int64_t row0 = tile * TILE_M;
    const uint32_t valid = (uint32_t)min((int64_t)TILE_M, g.R - row0);
    if (g.amode == A_CHUNK) {
      mbar_expect_tx(bar_a, (K / 8) * valid * 16);
#pragma unroll 1
      for (int c = 0; c < K / 8; ++c)
        bulk_g2s(sA + c * (TILE_M * 16), g.A_chunk + ((int64_t)c * g.R + row0) * 8, valid * 16, bar_a);
    } else {
      mbar_expect_tx(bar_a, valid * K * 4);
      bulk_g2s(sIO, g.A_f32 + row0 * K, valid * K * 4, bar_a);
    }
  };

  if (tid == 0) {
    mbar_expect_tx(bar_w, L::W_BYTES);
    bulk_g2s(sW, g.W_img + (int64_t)blockIdx.y * K * N_CTA, L::W_BYTES, bar_w);
    if (blockIdx.x < ntiles) issue_input(blockIdx.x);
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

  constexpr uint32_t IDESC = make_idesc(TILE_M, L::N_INSTR);
  uint32_t phase_a = 0, phase_mma = 0;
  bool w_ready = false, reduce_pending = false;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * TILE_M;
    const int64_t next = tile + gridDim.x;

    if (g.amode == A_F32) {
      // ---- fused normalisation prologue, shared -> shared: warp per row, fp32 -> bf16 operand image
      mbar_wait(bar_a, phase_a);
#pragma unroll 4
      for (int rr = 0; rr < 16; ++rr) {
        const int rl = warp * 16 + rr;
        const int64_t row = row0 + rl;
        const bool live = row < g.R;
        float2 v[NI];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const int k = 64 * i + 2 * lane;
          v[i] = make_float2(0.f, 0.f);
          if (live && k < K) v[i] = *reinterpret_cast<const float2*>(sIO + rl * K + k);
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
        const float* m = (g.pro == PRO_ADARMS && live) ? g.mod + (row / g.T) * g.mod_stride : nullptr;
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
            if (!live) x = y = 0.f;
            *reinterpret_cast<uint32_t*>(sA + (k >> 3) * (TILE_M * 16) + rl * 16 + (k & 7) * 2) = pack_bf16x2(x, y);
          }
        }
      }
      fence_proxy_async();
    }
    __syncthreads();   // operand image complete; fp32 input block consumed

    if (tid == 0) {
      if (g.amode == A_F32 && next < ntiles) issue_input(next);   // next block flies during MMA + epilogue
      if (!w_ready) {
        mbar_wait(bar_w, 0);
        w_ready = true;
      }
      if (g.amode == A_CHUNK) mbar_wait(bar_a, phase_a);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(sA), w_addr = smem_u32(sW);
#pragma unroll
      for (int nb_ = 0; nb_ < N_CTA / L::N_INSTR; ++nb_) {
#pragma unroll
        for (int ks = 0; ks < K / 16; ++ks) {
          const uint64_t ad = make_desc(a_addr + ks * 2 * (TILE_M * 16), TILE_M * 16, 128);
          const uint64_t bd = make_desc(w_addr + ks * 2 * (N_CTA * 16) + nb_ * L::N_INSTR * 16, N_CTA * 16, 128);
          umma_bf16(tmem_base + nb_ * L::N_INSTR, ad, bd, IDESC, ks > 0);
        }
      }
      umma_commit(bar_mma);
    }
    phase_a ^= 1;
    mbar_wait(bar_mma, phase_mma);
    phase_mma ^= 1;
    tc_fence_after();
    if (tid == 0 && g.amode == A_CHUNK && next < ntiles) issue_input(next);   // operand image is free again

    // ---- epilogue: TMEM -> registers -> global ------------------------------------
    const uint32_t trow = tmem_base + ((uint32_t)(lgroup * 32) << 16);
    const int rl = lgroup * 32 + lane;                 // tile row whose TMEM lane this thread reads
    const int64_t row = row0 + rl;
    if (g.epi == TE_CHUNK) {
#pragma unroll 1
      for (int c0 = half * 16; c0 < N_CTA; c0 += 32) {
        float v[16];
        tmem_ld16(trow + c0, v);
        if (g.bias) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += g.bias[n0 + c0 + j];
        }
        if (row < g.R) {
          __nv_bfloat16* o = g.out_chunk + ((int64_t)((n0 + c0) >> 3) * g.R + row) * 8;
          const bool f16 = ((n0 + c0) >> 3) >= g.f16_from_chunk;
          *reinterpret_cast<uint4*>(o) = f16 ? pack_f16x8(v) : pack_bf16x8(v);
          *reinterpret_cast<uint4*>(o + g.R * 8) = f16 ? pack_f16x8(v + 8) : pack_bf16x8(v + 8);
        }
      }
    } else if (g.epi == TE_SWIGLU) {
      constexpr int NU = N_CTA / 2;
#pragma unroll 1
      for (int c0 = half * 16; c0 < NU; c0 += 32) {
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
    } else if (g.epi == TE_RESID) {
      // h[row0.., n0..] += D + bias through the TMA engine: stage fp32 rows, reduce-add at L2
      if (reduce_pending) {
        if (warp == 0) bulk_wait_read_all();           // previous tile's staging fully read
        __syncthreads();
      }
#pragma unroll 1
      for (int c0 = half * 16; c0 < N_CTA; c0 += 32) {
        float v[16];
        tmem_ld16(trow + c0, v);
        if (g.bias) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += g.bias[n0 + c0 + j];
        }
        float4* d = reinterpret_cast<float4*>(sIO + rl * L::OUT_LD + c0);
        d[0] = make_float4(v[0], v[1], v[2], v[3]);
        d[1] = make_float4(v[4], v[5], v[6], v[7]);
        d[2] = make_float4(v[8], v[9], v[10], v[11]);
        d[3] = make_float4(v[12], v[13], v[14], v[15]);
      }
      fence_proxy_async();
      __syncthreads();
      if (warp == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = lane + 32 * i;
          if (row0 + r < g.R) bulk_reduce_add_f32(g.out_f32 + (row0 + r) * g.ldo + n0, sIO + r * L::OUT_LD, N_CTA * 4);
        }
        bulk_commit();
      }
      reduce_pending = true;
    } else {
      // TE_PE / TE_STEP / TE_F32: thread <-> row, 64 contiguous bytes per thread and chunk
      const edtts_step_args& sa = g.step;
      float ab_t = 0.f, ab_p = 1.f, al = 0.f, be = 0.f, pv = 0.f, nzm = 0.f;
      if (g.epi == TE_STEP && sa.mode != EDTTS_STEP_EPS && row < g.R) {
        const int b = (int)(row / g.T);
        const int64_t tt = sa.t[b];
        ab_t = sa.alpha_bar[tt];
        if (sa.mode == EDTTS_STEP_DDIM) {
          const int64_t tp = sa.t_prev[b];
          ab_p = tp >= 0 ? sa.alpha_bar[tp] : 1.0f;
        } else {
          al = sa.alphas[tt];
          be = sa.betas[tt];
          pv = sa.posterior_var[tt];
          nzm = tt > 0 ? 1.0f : 0.0f;
        }
      }
#pragma unroll 1
      for (int c0 = half * 16; c0 < N_CTA; c0 += 32) {
        float v[16];
        tmem_ld16(trow + c0, v);
        if (row >= g.R) continue;
        if (g.bias) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += g.bias[n0 + c0 + j];
        }
        const int64_t o = row * g.ldo + n0 + c0;
        if (g.epi == TE_PE) {
          const float4* pp = reinterpret_cast<const float4*>(g.pe + (row % g.T) * g.ldo + n0 + c0);
          float4 p[4] = {pp[0], pp[1], pp[2], pp[3]};
#pragma unroll
          for (int q = 0; q < 4; ++q)
            reinterpret_cast<float4*>(g.out_f32 + o)[q] =
                make_float4(v[4 * q] + p[q].x, v[4 * q + 1] + p[q].y, v[4 * q + 2] + p[q].z, v[4 * q + 3] + p[q].w);
        } else if (g.epi == TE_F32) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            reinterpret_cast<float4*>(g.out_f32 + o)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else {   // TE_STEP
          if (sa.eps_out) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              reinterpret_cast<float4*>(sa.eps_out + o)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
          if (sa.mode != EDTTS_STEP_EPS) {
            float x[16], nz[16], xp[16], x0[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              *reinterpret_cast<float4*>(x + 4 * q) = reinterpret_cast<const float4*>(g.x_t + o)[q];
              if (sa.mode == EDTTS_STEP_DDPM) *reinterpret_cast<float4*>(nz + 4 * q) = reinterpret_cast<const float4*>(sa.noise + o)[q];
            }
            if (sa.mode == EDTTS_STEP_DDIM) {
#pragma unroll
              for (int j = 0; j < 16; ++j) ddim_update(x[j], v[j], 0.f, ab_t, ab_p, 0.f, xp[j], x0[j]);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                if (sa.x0_out) reinterpret_cast<float4*>(sa.x0_out + o)[q] = *reinterpret_cast<float4*>(x0 + 4 * q);
                if (sa.write_x_prev && sa.x_prev_out)
                  reinterpret_cast<float4*>(sa.x_prev_out + o)[q] = *reinterpret_cast<float4*>(xp + 4 * q);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) xp[j] = ddpm_update(x[j], v[j], nz[j], al, ab_t, be, pv, nzm);
#pragma unroll
              for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(sa.x_prev_out + o)[q] = *reinterpret_cast<float4*>(xp + 4 * q);
            }
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();   // TMEM drained before the next tile's MMA
  }

  if (warp == 0 && reduce_pending) bulk_wait_all();   // staging must outlive the engine's reads
  __syncthreads();
  if (warp == 0) tmem_dealloc<L::TMEM_COLS>(tmem_base);
}

}  // namespace tc
}  // namespace edtts
