// Fused DiffusionTransformerBlock for the bf16 tensor-core path (layers/transformer.py:141-160).
//
// One persistent CTA per SM; a work item is one (utterance, 128-frame tile).  For its tile the
// CTA runs the whole block on chip -- the only HBM traffic of a layer is: q/k/v of the tile
// (+ the +-64 frame halo of k, v) in, the fp32 residual stream h in and out, and the layer's
// weights / the utterance's context K, V streamed from L2:
//
//   h (128 x 160 fp32) lives in TENSOR MEMORY for the whole layer: every output projection is a
//   tcgen05.mma that ACCUMULATES into the h columns, so the three residual adds cost nothing.
//
//   window attention (attention.py:94-111)   4 heads, two at a time (one per warpgroup):
//        S = Q_h K_h^T (128 x 128 keys) -> softmax (thread <-> frame, out of TMEM) -> P (bf16, smem)
//        -> O_h += P V_h ; two key blocks cover the band t0-64 .. t0+191, online softmax
//   h += O Wproj^T                            tcgen05.mma into h
//   n2 = RMSNorm(h + b_proj) w2               row pass, TMEM -> regs -> bf16 A operand (smem)
//   q  = n2 Wq^T                              tcgen05.mma into scratch columns, -> bf16 Q operand
//   cross attention (mla.py:176-180)          same kernel body, keys = the S context tokens
//   h += O Wout^T
//   n3 = AdaRMSNorm(h)                        row pass
//   u  = swiglu(n3 W0^T + b0)                 two halves of 160 u-columns, each 128 x 320 in TMEM
//   h += u W3^T (+ b3)                        tcgen05.mma into h, then h -> HBM
//
// 256 threads = 2 warpgroups.  In the attention phases each warpgroup owns two heads and its
// own K/V/P buffers, TMEM columns and mbarriers, so the tensor core works on one head while the
// CUDA cores run the other head's softmax.  In the row passes warpgroup g handles columns
// 80g .. 80g+79 of every row (thread <-> row, TMEM lane = row).  Weights are streamed from L2
// in 160 x 160 bf16 chunks (51,200 B, pre-packed UMMA operand images) through two slots by the
// TMA engine, prefetched one GEMM ahead.
#pragma once
#include "umma.cuh"

namespace edtts {
namespace tc {

constexpr int LY_THREADS = 256;
constexpr int LY_SLAB = 128 * 16;             // one 8-wide K slab of a 128-row operand
constexpr int LY_WSLAB = 160 * 16;            // one 8-wide K slab of a 160-row weight chunk
constexpr int LY_WCHUNK = 20 * LY_WSLAB;      // 51,200 B: W[160 out][160 in] bf16
constexpr int LY_NCHUNK = 9;                  // proj, q_proj, out_proj, ffn0 x4, ffn3 x2
enum LyChunk : int { WC_PROJ = 0, WC_Q, WC_OUT, WC_F0_X0, WC_F0_G0, WC_F0_X1, WC_F0_G1, WC_F3_K0, WC_F3_K1 };

// per-layer constant vector (floats), packed by tc_layer.cu
constexpr int LC_PROJ_B = 0, LC_N2W = 160, LC_N3W = 320, LC_F0B = 480, LC_F3B = 1120, LC_COUNT = 1280;
// shared-memory copy: the above + the tile's AdaLN vectors g3 = w3 * (1 + scale), sh3 = shift
constexpr int LS_G3 = 1280, LS_SH3 = 1440, LS_COUNT = 1600;

// shared memory map (bytes)
constexpr int LO_A = 0;                                   // 21 slabs: A operand / Q / attention output
constexpr int LO_W0 = LO_A + 21 * LY_SLAB;                // weight slot 0
constexpr int LO_X = LO_W0 + LY_WCHUNK;                   // overlay region
constexpr int LO_W1 = LO_X;                               //   GEMM chain: weight slot 1
constexpr int LO_U = LO_X + LY_WCHUNK;                    //   GEMM chain: u half (128 x 160 bf16)
constexpr int LO_P = LO_X;                                //   attention: P of warpgroup g at + g * 32 KB
constexpr int LO_KV = LO_X + 2 * 16 * LY_SLAB;            //   attention: K | V of warpgroup g at + g * 24 KB
constexpr int LO_X_END = LO_KV + 2 * 12 * LY_SLAB;
constexpr int LO_CONST = LO_X_END;
constexpr int LO_RED = LO_CONST + LS_COUNT * 4;
constexpr int LO_BAR = LO_RED + 2 * 128 * 4;
constexpr int LY_SMEM = LO_BAR + 16 * 8 + 16;
static_assert(LO_U + 20 * LY_SLAB <= LO_X_END, "u half must fit in the overlay region");
static_assert(LY_SMEM <= 232448, "shared memory budget");

// tensor memory map (columns)
constexpr uint32_t TM_H = 0;                              // residual stream, 160 columns
constexpr uint32_t TM_G = 160;                            // GEMM chain scratch, 320 columns
constexpr uint32_t TM_S0 = 160, TM_WG = 176;              // attention: S (128) | O (48) per warpgroup

struct LayerArgs {
  float* h;                          // [R][160] fp32 residual stream, in place
  const __nv_bfloat16* qkv;          // [60][R][8] chunk-major q | k | v of this layer (finite slack around it)
  const __nv_bfloat16* kvx;          // [40][RS][8] chunk-major context k | v of this layer
  const __nv_bfloat16* wimg;         // LY_NCHUNK weight chunks
  const float* consts;               // LC_COUNT floats
  const float* mod3;                 // norm3 (scale | shift) of utterance 0; + b * mod_stride
  int mod_stride;
  int64_t R, RS;
  int B, T, S;
  int tiles_per_utt;
  float scale_log2e;                 // head_dim^-0.5 * log2(e)
  int stop_phase;                    // debug: 1 = stop after attention + proj, 2 = after cross, 0 = whole block
};

__device__ __forceinline__ float fast_silu(float g) {
  return __fdividef(g, 1.0f + ex2_approx(-1.4426950408889634f * g));
}

// D[128 x 160] (+)= A[128 x 160] * W[160 x 160]^T, one elected thread
__device__ __forceinline__ void ly_issue_gemm(uint32_t d_tmem, uint32_t a_addr, uint32_t w_addr, bool accumulate) {
  constexpr uint32_t IDESC = make_idesc(128, 160);
#pragma unroll
  for (int ks = 0; ks < 10; ++ks)
    umma_bf16(d_tmem, make_desc(a_addr + ks * 2 * LY_SLAB, LY_SLAB, 128), make_desc(w_addr + ks * 2 * LY_WSLAB, LY_WSLAB, 128),
              IDESC, accumulate || ks > 0);
}

struct LyTile {
  int b, t0, nq;
  int64_t row0;
};

// One attention phase of one warpgroup (heads wg and wg + 2).  Q is in sA (slabs 5*head ..), the
// normalised output replaces it there.  K/V of the first (head, block) item must already be in flight.
template <bool WINDOW>
__device__ __forceinline__ void ly_attention(const LayerArgs& a, const LyTile& tl, uint8_t* smem, uint32_t tmem_base,
                                             uint64_t* bars, uint32_t& ph_k, uint32_t& ph_v, uint32_t& ph_s,
                                             uint32_t& ph_o) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wg = warp >> 2, lq = warp & 3, row = lq * 32 + lane;
  const bool wl = (tid & 127) == 0;                       // warpgroup leader: issues copies and MMAs
  uint8_t* sA = smem + LO_A;
  uint8_t* sP = smem + LO_P + wg * (16 * LY_SLAB);
  uint8_t* sK = smem + LO_KV + wg * (12 * LY_SLAB);
  uint8_t* sV = sK + 6 * LY_SLAB;
  uint64_t* bar_k = bars + 4 + wg;
  uint64_t* bar_v = bars + 6 + wg;
  uint64_t* bar_s = bars + 8 + wg;
  uint64_t* bar_o = bars + 10 + wg;
  const uint32_t trow = tmem_base + ((uint32_t)(lq * 32) << 16);
  const uint32_t tS = trow + TM_S0 + wg * TM_WG;
  const uint32_t tO = tS + 128;
  const uint32_t dS = tmem_base + TM_S0 + wg * TM_WG;     // MMA destinations (lane 0)
  const uint32_t dO = dS + 128;

  const int nblocks = WINDOW ? ((tl.t0 + WIN < a.T) ? 2 : 1) : (a.S + 127) / 128;
  const int n_items = 2 * nblocks;

  auto issue_k = [&](int it) {                            // leader only
    const int head = wg + 2 * (it / nblocks), kb = it % nblocks;
    if (WINDOW) {
      const int64_t g0 = tl.row0 - WIN + 128 * kb;
      mbar_expect_tx(bar_k, 5 * LY_SLAB);
#pragma unroll
      for (int g = 0; g < 5; ++g)
        bulk_g2s(sK + g * LY_SLAB, a.qkv + ((int64_t)(20 + head * 5 + g) * a.R + g0) * 8, LY_SLAB, bar_k);
    } else {
      const int nv = min(128, a.S - kb * 128);
      const int64_t g0 = (int64_t)tl.b * a.S + kb * 128;
      mbar_expect_tx(bar_k, 5 * nv * 16);
#pragma unroll
      for (int g = 0; g < 5; ++g)
        bulk_g2s(sK + g * LY_SLAB, a.kvx + ((int64_t)(head * 5 + g) * a.RS + g0) * 8, nv * 16, bar_k);
    }
  };
  auto issue_v = [&](int it) {
    const int head = wg + 2 * (it / nblocks), kb = it % nblocks;
    if (WINDOW) {
      const int64_t g0 = tl.row0 - WIN + 128 * kb;
      mbar_expect_tx(bar_v, 5 * LY_SLAB);
#pragma unroll
      for (int g = 0; g < 5; ++g)
        bulk_g2s(sV + g * LY_SLAB, a.qkv + ((int64_t)(40 + head * 5 + g) * a.R + g0) * 8, LY_SLAB, bar_v);
    } else {
      const int nv = min(128, a.S - kb * 128);
      const int64_t g0 = (int64_t)tl.b * a.S + kb * 128;
      mbar_expect_tx(bar_v, 5 * nv * 16);
#pragma unroll
      for (int g = 0; g < 5; ++g)
        bulk_g2s(sV + g * LY_SLAB, a.kvx + ((int64_t)(20 + head * 5 + g) * a.RS + g0) * 8, nv * 16, bar_v);
    }
  };

  float o_acc[HD];
  float m_run = -INFINITY, l_run = 0.f;

  for (int it = 0; it < n_items; ++it) {
    const int head = wg + 2 * (it / nblocks), kb = it % nblocks;
    if (kb == 0) {
      m_run = -INFINITY;
      l_run = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) o_acc[d] = 0.f;
    }
    // valid key columns [lo, hi] of this thread's row; 32-column chunks [cfirst, clast] of this warp
    int lo, hi, cfirst, clast, nkeys;
    if (WINDOW) {
      const int j0 = tl.t0 - WIN + 128 * kb;              // frame index of key column 0
      nkeys = 128;
      if (kb == 0) {
        lo = max(row, -j0);
        hi = min(127, a.T - 1 - j0);
        cfirst = max(lq, (j0 < 0 ? -j0 : 0) >> 5);
        clast = 3;
      } else {
        lo = 0;
        hi = min(row, a.T - 1 - j0);
        cfirst = 0;
        clast = min(lq, (a.T - 1 - j0) >> 5);
      }
    } else {
      const int nv = min(128, a.S - kb * 128);
      nkeys = (nv + 15) & ~15;
      lo = 0;
      hi = nv - 1;
      cfirst = 0;
      clast = (nv - 1) >> 5;
    }
    const int nchunks = (nkeys + 31) >> 5;

    // ---- S = Q_h K^T ---------------------------------------------------------------------
    if (wl) {
      mbar_wait(bar_k, ph_k);
      tc_fence_after();
      const uint32_t idesc = make_idesc(128, (uint32_t)nkeys);
      const uint32_t qa = smem_u32(sA) + head * 5 * LY_SLAB, ka = smem_u32(sK);
#pragma unroll
      for (int ks = 0; ks < 3; ++ks)
        umma_bf16(dS, make_desc(qa + ks * 2 * LY_SLAB, LY_SLAB, 128), make_desc(ka + ks * 2 * LY_SLAB, LY_SLAB, 128), idesc,
                  ks > 0);
      umma_commit(bar_s);
    }
    ph_k ^= 1;
    mbar_wait(bar_s, ph_s);
    ph_s ^= 1;
    tc_fence_after();
    if (wl && it + 1 < n_items) issue_k(it + 1);          // K buffer is free again

    // ---- softmax ------------------------------------------------------------------------------
    float bmax = -INFINITY;
    for (int ch = cfirst; ch <= clast; ++ch) {
      float s[32];
      tmem_ld32(tS + ch * 32, s);
      const bool full = __all_sync(0xffffffffu, lo <= ch * 32 && hi >= ch * 32 + 31);
      if (full) {
#pragma unroll
        for (int j = 0; j < 32; ++j) bmax = fmaxf(bmax, s[j]);
      } else {
        const int base = ch * 32 - lo;
        const unsigned span = (unsigned)(hi - lo);
        const bool any = hi >= lo;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (any && (unsigned)(base + j) <= span) bmax = fmaxf(bmax, s[j]);
      }
    }
    const float m_new = fmaxf(m_run, bmax * a.scale_log2e);
    const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
    const float alpha = ex2_approx(m_run - m_use);        // m_run = -inf -> 0
    float lsum = 0.f;
    for (int ch = 0; ch < nchunks; ++ch) {
      uint4 pk[4];
      if (ch < cfirst || ch > clast) {                    // warp-uniform: columns no row of this warp needs
#pragma unroll
        for (int g = 0; g < 4; ++g) pk[g] = make_uint4(0, 0, 0, 0);
      } else {
        float s[32];
        tmem_ld32(tS + ch * 32, s);
        const bool full = __all_sync(0xffffffffu, lo <= ch * 32 && hi >= ch * 32 + 31);
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            s[j] = ex2_approx(fmaf(s[j], a.scale_log2e, -m_use));
            lsum += s[j];
          }
        } else {
          const int base = ch * 32 - lo;
          const unsigned span = (unsigned)(hi - lo);
          const bool any = hi >= lo;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float p = ex2_approx(fmaf(s[j], a.scale_log2e, -m_use));
            s[j] = (any && (unsigned)(base + j) <= span) ? p : 0.f;
            lsum += s[j];
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) pk[g] = pack_bf16x8(s + 8 * g);
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(sP + (ch * 4 + g) * LY_SLAB + row * 16) = pk[g];
    }
    l_run = l_run * alpha + lsum;
    m_run = m_new;
    fence_proxy_async();
    tc_fence_before();
    named_bar_sync(1 + wg, 128);                          // P complete, S columns drained

    // ---- O_blk = P V -------------------------------------------------------------------------
    if (wl) {
      mbar_wait(bar_v, ph_v);
      tc_fence_after();
      constexpr uint32_t IDESC_O = make_idesc(128, 48, /*b_mn_major=*/true);
      const uint32_t pa = smem_u32(sP), va = smem_u32(sV);
      const int nks = nkeys >> 4;
      for (int ks = 0; ks < nks; ++ks)
        umma_bf16(dO, make_desc(pa + ks * 2 * LY_SLAB, LY_SLAB, 128),
                  make_desc(va + ks * 2 * 128, /*LBO: next 8 keys*/ 128, /*SBO: next 8 dims*/ LY_SLAB), IDESC_O, ks > 0);
      umma_commit(bar_o);
    }
    ph_v ^= 1;
    mbar_wait(bar_o, ph_o);
    ph_o ^= 1;
    tc_fence_after();
    if (wl && it + 1 < n_items) issue_v(it + 1);          // V buffer is free again
    {
      float ob[HD];
      tmem_ld32(tO, ob);
      tmem_ld8(tO + 32, ob + 32);
#pragma unroll
      for (int d = 0; d < HD; ++d) o_acc[d] = fmaf(o_acc[d], alpha, ob[d]);
    }
    if (kb == nblocks - 1) {                              // head done: normalised output replaces Q_h in sA
      const float inv = __fdividef(1.0f, l_run);
#pragma unroll
      for (int g = 0; g < 5; ++g) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = o_acc[g * 8 + j] * inv;
        *reinterpret_cast<uint4*>(sA + (head * 5 + g) * LY_SLAB + row * 16) = pack_bf16x8(v);
      }
    }
    tc_fence_before();
  }
  fence_proxy_async();
}

__global__ void __launch_bounds__(LY_THREADS, 1) tc_layer_kernel(const LayerArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem + LO_A;
  uint8_t* sW0 = smem + LO_W0;
  uint8_t* sW1 = smem + LO_W1;
  uint8_t* sU = smem + LO_U;
  float* sC = reinterpret_cast<float*>(smem + LO_CONST);
  float* sRed = reinterpret_cast<float*>(smem + LO_RED);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LO_BAR);
  uint64_t* bar_w0 = bars + 0;
  uint64_t* bar_w1 = bars + 1;
  uint64_t* bar_q = bars + 2;
  uint64_t* bar_g = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wg = warp >> 2, lq = warp & 3, row = lq * 32 + lane;
  const int cb = 80 * wg;                                 // this thread's column half in the row passes
  const bool wl = (tid & 127) == 0;

  // finite shared memory everywhere (stale rows enter MMAs as 0 * x), zero pad slabs
  for (int i = tid * 16; i < LO_BAR; i += LY_THREADS * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int i = tid; i < LC_COUNT; i += LY_THREADS) sC[i] = a.consts[i];
  if (tid == 0) {
    for (int i = 0; i < 12; ++i) mbar_init(bars + i, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t trow = tmem_base + ((uint32_t)(lq * 32) << 16);

  const int ntiles = a.B * a.tiles_per_utt;
  uint32_t ph_w0 = 0, ph_w1 = 0, ph_q = 0, ph_g = 0, ph_k = 0, ph_v = 0, ph_s = 0, ph_o = 0;
  auto load_w = [&](int chunk, uint8_t* slot, uint64_t* bar) {   // tid 0 only
    mbar_expect_tx(bar, LY_WCHUNK);
    bulk_g2s(slot, a.wimg + (int64_t)chunk * (LY_WCHUNK / 2), LY_WCHUNK, bar);
  };
  auto gemm_wait = [&]() {
    mbar_wait(bar_g, ph_g);
    ph_g ^= 1;
    tc_fence_after();
  };
  // store h (TMEM, + optional bias) of the valid rows to HBM
  auto store_h = [&](const LyTile& tl, const float* bias) {
#pragma unroll 1
    for (int i = 0; i < 5; ++i) {
      float v[16];
      tmem_ld16(trow + TM_H + cb + 16 * i, v);
      if (row < tl.nq) {
        float4* dst = reinterpret_cast<float4*>(a.h + (tl.row0 + row) * H + cb + 16 * i);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 o = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          if (bias) {
            const float4 bb = *reinterpret_cast<const float4*>(bias + cb + 16 * i + 4 * q);
            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
          }
          dst[q] = o;
        }
      }
    }
  };

  if (tid == 0 && blockIdx.x < ntiles) load_w(WC_PROJ, sW0, bar_w0);

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    LyTile tl;
    tl.b = tile / a.tiles_per_utt;
    tl.t0 = (tile % a.tiles_per_utt) * 128;
    tl.nq = min(128, a.T - tl.t0);
    tl.row0 = (int64_t)tl.b * a.T + tl.t0;

    // ---- tile prologue: Q -> sA, first K/V blocks, h -> TMEM, AdaLN vectors ----------------
    if (tid == 0) {
      mbar_expect_tx(bar_q, 20 * tl.nq * 16);
#pragma unroll 1
      for (int c = 0; c < 20; ++c) bulk_g2s(sA + c * LY_SLAB, a.qkv + ((int64_t)c * a.R + tl.row0) * 8, tl.nq * 16, bar_q);
    }
    if (wl) {
      // first window item of this warpgroup: head wg, key block 0
      const int64_t g0 = tl.row0 - WIN;
      uint8_t* sK = smem + LO_KV + wg * (12 * LY_SLAB);
      uint8_t* sV = sK + 6 * LY_SLAB;
      mbar_expect_tx(bars + 4 + wg, 5 * LY_SLAB);
      mbar_expect_tx(bars + 6 + wg, 5 * LY_SLAB);
#pragma unroll
      for (int g = 0; g < 5; ++g) {
        bulk_g2s(sK + g * LY_SLAB, a.qkv + ((int64_t)(20 + wg * 5 + g) * a.R + g0) * 8, LY_SLAB, bars + 4 + wg);
        bulk_g2s(sV + g * LY_SLAB, a.qkv + ((int64_t)(40 + wg * 5 + g) * a.R + g0) * 8, LY_SLAB, bars + 6 + wg);
      }
    }
    {
      const float* src = a.h + (tl.row0 + row) * H + cb;
#pragma unroll 1
      for (int i = 0; i < 5; ++i) {
        float v[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row < tl.nq) x = *reinterpret_cast<const float4*>(src + 16 * i + 4 * q);
          v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
        }
        tmem_st16(trow + TM_H + cb + 16 * i, v);
      }
      tmem_st_wait();
    }
    if (tid < H) {
      const float* m = a.mod3 + (int64_t)tl.b * a.mod_stride;
      sC[LS_G3 + tid] = sC[LC_N3W + tid] * (1.0f + m[tid]);
      sC[LS_SH3 + tid] = m[H + tid];
    }
    mbar_wait(bar_q, ph_q);
    ph_q ^= 1;
    tc_fence_before();
    __syncthreads();

    // ---- banded self-attention ---------------------------------------------------------------
    ly_attention<true>(a, tl, smem, tmem_base, bars, ph_k, ph_v, ph_s, ph_o);
    __syncthreads();
    if (wl) {                                             // context K/V of the first cross item (overlay tail is free)
      const int nv = min(128, a.S);
      const int64_t g0 = (int64_t)tl.b * a.S;
      uint8_t* sK = smem + LO_KV + wg * (12 * LY_SLAB);
      uint8_t* sV = sK + 6 * LY_SLAB;
      mbar_expect_tx(bars + 4 + wg, 5 * nv * 16);
      mbar_expect_tx(bars + 6 + wg, 5 * nv * 16);
#pragma unroll
      for (int g = 0; g < 5; ++g) {
        bulk_g2s(sK + g * LY_SLAB, a.kvx + ((int64_t)(wg * 5 + g) * a.RS + g0) * 8, nv * 16, bars + 4 + wg);
        bulk_g2s(sV + g * LY_SLAB, a.kvx + ((int64_t)(20 + wg * 5 + g) * a.RS + g0) * 8, nv * 16, bars + 6 + wg);
      }
    }

    // ---- h += O Wproj^T ------------------------------------------------------------------------
    if (tid == 0) {
      load_w(WC_Q, sW1, bar_w1);
      mbar_wait(bar_w0, ph_w0);
      tc_fence_after();
      ly_issue_gemm(tmem_base + TM_H, smem_u32(sA), smem_u32(sW0), true);
      umma_commit(bar_g);
    }
    ph_w0 ^= 1;
    gemm_wait();
    if (tid == 0) load_w(WC_OUT, sW0, bar_w0);

    // ---- n2 = RMSNorm(h + b_proj) * w2 -> sA ; h + b_proj back to TMEM ---------------------------
    {
      float v[80];
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        tmem_ld16(trow + TM_H + cb + 16 * i, v + 16 * i);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          v[16 * i + j] += sC[LC_PROJ_B + cb + 16 * i + j];
          ss = fmaf(v[16 * i + j], v[16 * i + j], ss);
        }
        tmem_st16(trow + TM_H + cb + 16 * i, v + 16 * i);
      }
      sRed[wg * 128 + row] = ss;
      tmem_st_wait();
      __syncthreads();
      const float rstd = rsqrtf((sRed[row] + sRed[128 + row]) * (1.0f / H) + 1e-6f);
#pragma unroll
      for (int g = 0; g < 10; ++g) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = v[8 * g + j] * rstd * sC[LC_N2W + cb + 8 * g + j];
        *reinterpret_cast<uint4*>(sA + (cb / 8 + g) * LY_SLAB + row * 16) = pack_bf16x8(o);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (a.stop_phase == 1) {
      tc_fence_after();
      store_h(tl, nullptr);
      // drain the prefetches so that the barrier phases stay consistent
      if (tid == 0) { mbar_wait(bar_w1, ph_w1); mbar_wait(bar_w0, ph_w0); }
      if (wl) { mbar_wait(bars + 4 + wg, ph_k); mbar_wait(bars + 6 + wg, ph_v); }
      ph_w1 ^= 1; ph_w0 ^= 1; ph_k ^= 1; ph_v ^= 1;
      tc_fence_before();
      __syncthreads();
      if (tid == 0 && tile + gridDim.x < ntiles) load_w(WC_PROJ, sW0, bar_w0);
      continue;
    }

    // ---- q = n2 Wq^T -> bf16 Q operand in sA -------------------------------------------------------
    if (tid == 0) {
      mbar_wait(bar_w1, ph_w1);
      tc_fence_after();
      ly_issue_gemm(tmem_base + TM_G, smem_u32(sA), smem_u32(sW1), false);
      umma_commit(bar_g);
    }
    ph_w1 ^= 1;
    gemm_wait();
#pragma unroll 1
    for (int i = 0; i < 5; ++i) {
      float v[16];
      tmem_ld16(trow + TM_G + cb + 16 * i, v);
      *reinterpret_cast<uint4*>(sA + (cb / 8 + 2 * i) * LY_SLAB + row * 16) = pack_bf16x8(v);
      *reinterpret_cast<uint4*>(sA + (cb / 8 + 2 * i + 1) * LY_SLAB + row * 16) = pack_bf16x8(v + 8);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();

    // ---- cross attention over the context tokens ---------------------------------------------------
    ly_attention<false>(a, tl, smem, tmem_base, bars, ph_k, ph_v, ph_s, ph_o);
    __syncthreads();

    // ---- h += O Wout^T -----------------------------------------------------------------------------
    if (tid == 0) {
      load_w(WC_F0_X0, sW1, bar_w1);
      mbar_wait(bar_w0, ph_w0);
      tc_fence_after();
      ly_issue_gemm(tmem_base + TM_H, smem_u32(sA), smem_u32(sW0), true);
      umma_commit(bar_g);
    }
    ph_w0 ^= 1;
    gemm_wait();
    if (tid == 0) load_w(WC_F0_G0, sW0, bar_w0);

    // ---- n3 = AdaRMSNorm(h) -> sA --------------------------------------------------------------------
    {
      float v[80];
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        tmem_ld16(trow + TM_H + cb + 16 * i, v + 16 * i);
#pragma unroll
        for (int j = 0; j < 16; ++j) ss = fmaf(v[16 * i + j], v[16 * i + j], ss);
      }
      sRed[wg * 128 + row] = ss;
      __syncthreads();
      const float rstd = rsqrtf((sRed[row] + sRed[128 + row]) * (1.0f / H) + 1e-6f);
#pragma unroll
      for (int g = 0; g < 10; ++g) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o[j] = fmaf(v[8 * g + j] * rstd, sC[LS_G3 + cb + 8 * g + j], sC[LS_SH3 + cb + 8 * g + j]);
        *reinterpret_cast<uint4*>(sA + (cb / 8 + g) * LY_SLAB + row * 16) = pack_bf16x8(o);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (a.stop_phase == 2) {
      tc_fence_after();
      store_h(tl, nullptr);
      if (tid == 0) { mbar_wait(bar_w1, ph_w1); mbar_wait(bar_w0, ph_w0); }
      ph_w1 ^= 1; ph_w0 ^= 1;
      tc_fence_before();
      __syncthreads();
      if (tid == 0 && tile + gridDim.x < ntiles) load_w(WC_PROJ, sW0, bar_w0);
      continue;
    }

    // ---- feed-forward: two halves of 160 u columns ---------------------------------------------------
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      if (tid == 0) {
        mbar_wait(bar_w1, ph_w1);
        mbar_wait(bar_w0, ph_w0);
        tc_fence_after();
        ly_issue_gemm(tmem_base + TM_G, smem_u32(sA), smem_u32(sW1), false);          // x part
        ly_issue_gemm(tmem_base + TM_G + 160, smem_u32(sA), smem_u32(sW0), false);    // gate part
        umma_commit(bar_g);
      }
      ph_w1 ^= 1;
      ph_w0 ^= 1;
      gemm_wait();
      if (tid == 0) {
        load_w(half == 0 ? WC_F0_X1 : WC_F3_K0, sW1, bar_w1);
        load_w(half == 0 ? WC_F0_G1 : WC_F3_K1, sW0, bar_w0);
      }
      uint8_t* dst = half == 0 ? sU : sA;                 // sA (n3) is dead once the second half's MMAs retired
      const float* bx = sC + LC_F0B + half * 320 + cb;
      const float* bg = bx + 160;
#pragma unroll 1
      for (int i = 0; i < 5; ++i) {
        float x[16], g[16];
        tmem_ld16(trow + TM_G + cb + 16 * i, x);
        tmem_ld16(trow + TM_G + 160 + cb + 16 * i, g);
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = (x[j] + bx[16 * i + j]) * fast_silu(g[j] + bg[16 * i + j]);
        *reinterpret_cast<uint4*>(dst + (cb / 8 + 2 * i) * LY_SLAB + row * 16) = pack_bf16x8(x);
        *reinterpret_cast<uint4*>(dst + (cb / 8 + 2 * i + 1) * LY_SLAB + row * 16) = pack_bf16x8(x + 8);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
    }

    // ---- h += u W3^T ; h + b3 -> HBM --------------------------------------------------------------------
    if (tid == 0) {
      mbar_wait(bar_w1, ph_w1);
      mbar_wait(bar_w0, ph_w0);
      tc_fence_after();
      ly_issue_gemm(tmem_base + TM_H, smem_u32(sU), smem_u32(sW1), true);
      ly_issue_gemm(tmem_base + TM_H, smem_u32(sA), smem_u32(sW0), true);
      umma_commit(bar_g);
    }
    ph_w1 ^= 1;
    ph_w0 ^= 1;
    gemm_wait();
    if (tid == 0 && tile + gridDim.x < ntiles) load_w(WC_PROJ, sW0, bar_w0);
    store_h(tl, sC + LC_F3B);
    tc_fence_before();
    __syncthreads();
  }

  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

}  // namespace tc
}  // namespace edtts
