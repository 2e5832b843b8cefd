// tcgen05 flash-style attention specialised for hidden=160, 4 heads, head_dim=40
// (EDTTS_PREC_BF16).  One CTA = one (utterance, 128-query tile, head).
//
//   WINDOW : EfficientAttention, |i-j| <= 64 (layers/attention.py:94-111).  The 128 queries
//            need keys t0-64 .. t0+191: ONE 256-key block, S = Q K^T is a single
//            128 x 256 x 48 UMMA into TMEM; the dense [T,T] mask never exists.
//   CROSS  : MultiHeadLatentAttention over the S context tokens (layers/mla.py:176-180):
//            256-key blocks with an online softmax; the per-block P V lands in TMEM and is
//            merged into a register accumulator (thread <-> query row).
//
// Operands arrive as bf16 chunk-major slabs ([d/8][row][8], see umma.cuh) written by the QKV /
// q_proj / context kernels, so Q, K and V tiles are fetched with bulk async copies straight into
// the UMMA operand image: Q and K as K-major operands (head_dim padded 40 -> 48 with a zeroed
// slab), V as the MN-major B operand of P V.  Softmax runs thread <-> row out of TMEM
// (tcgen05.ld), P is written to shared memory as the bf16 A operand (16-byte conflict-free
// stores), and the S columns of TMEM are reused for the P V accumulator.
#pragma once
#include "umma.cuh"

namespace edtts {
namespace tc {

constexpr int ATT_PAD_BYTES = 4096;     // finite slack around the qkv buffer (band halo of the first/last rows)
constexpr int AT_M = 128;               // queries per CTA
constexpr int AT_KB = 256;              // keys per block
constexpr int AT_DG = 6;                // head_dim 40 padded to 48 = 6 slabs of 8

struct TcAttnArgs {
  const __nv_bfloat16* q;   int64_t q_rows;     // chunk-major [..][q_rows][8]; head h, slab g at chunk q_chunk0 + 5h + g
  const __nv_bfloat16* kv;  int64_t kv_rows;    // chunk-major [..][kv_rows][8]
  int q_chunk0, k_chunk0, v_chunk0;
  __nv_bfloat16* o;                              // chunk-major [20][q_rows][8]
  int Tq, Tk;                                    // per-utterance lengths
  float scale_log2e;                             // head_dim^-0.5 * log2(e)
};

struct AttnSmem {
  static constexpr int Q_BYTES = AT_DG * AT_M * 16;        // 12 KB
  static constexpr int K_BYTES = AT_DG * AT_KB * 16;       // 24 KB
  static constexpr int V_BYTES = AT_DG * AT_KB * 16;       // 24 KB (slab 5 never read into a stored column)
  static constexpr int P_BYTES = (AT_KB / 8) * AT_M * 16;  // 64 KB
  // P overlays Q|K (dead once S = Q K^T has retired) and extends past them
  static constexpr int OFF_Q = 0, OFF_K = Q_BYTES, OFF_P = 0;
  static constexpr int OFF_V = P_BYTES;
  static constexpr int OFF_BAR = OFF_V + V_BYTES;
  static constexpr int TOTAL = OFF_BAR + 64;
  static constexpr uint32_t TMEM_COLS = 256;
};

template <bool WINDOW>
__global__ void __launch_bounds__(128) tc_attn_kernel(const TcAttnArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sQ = smem + AttnSmem::OFF_Q;
  uint8_t* sK = smem + AttnSmem::OFF_K;
  uint8_t* sP = smem + AttnSmem::OFF_P;
  uint8_t* sV = smem + AttnSmem::OFF_V;
  uint64_t* bar_qk = reinterpret_cast<uint64_t*>(smem + AttnSmem::OFF_BAR);
  uint64_t* bar_v = bar_qk + 1;
  uint64_t* bar_mma = bar_qk + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_qk + 3);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int head = blockIdx.y, b = blockIdx.z;
  const int t0 = blockIdx.x * AT_M;
  const int qi = t0 + tid;                                  // this thread's query (local index)
  const int64_t qrow0 = (int64_t)b * a.Tq + t0;             // global row of the tile's first query
  const int nq = min(AT_M, a.Tq - t0);

  if (tid == 0) {
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<AttnSmem::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);

  float o_acc[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) o_acc[d] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;

  const int nblocks = WINDOW ? 1 : (a.Tk + AT_KB - 1) / AT_KB;
  uint32_t ph_qk = 0, ph_v = 0, ph_mma = 0;

  for (int kb = 0; kb < nblocks; ++kb) {
    // key block: local key index of column c is j0 + c
    const int j0 = WINDOW ? t0 - WIN : kb * AT_KB;
    const int64_t krow0 = (int64_t)b * a.Tk + j0;           // may point into the zeroed slack / neighbours
    int ncopy = AT_KB;                                      // rows fetched (CROSS: only this utterance's keys)
    if (!WINDOW) ncopy = min(AT_KB, a.Tk - j0);

    // ---- loads: Q (first block only) + K on bar_qk, V on bar_v ------------------------
    if (kb > 0) {
      // P of the previous block overlays Q|K: Q must be re-fetched together with K
      __syncthreads();
    }
    if (tid == 0) {
      mbar_expect_tx(bar_qk, (uint32_t)(5 * nq * 16 + 5 * ncopy * 16));
      for (int g = 0; g < 5; ++g) {
        bulk_g2s(sQ + g * (AT_M * 16), a.q + ((int64_t)(a.q_chunk0 + head * 5 + g) * a.q_rows + qrow0) * 8, nq * 16,
                 bar_qk);
        bulk_g2s(sK + g * (AT_KB * 16), a.kv + ((int64_t)(a.k_chunk0 + head * 5 + g) * a.kv_rows + krow0) * 8,
                 ncopy * 16, bar_qk);
      }
      mbar_expect_tx(bar_v, (uint32_t)(5 * ncopy * 16));
      for (int g = 0; g < 5; ++g)
        bulk_g2s(sV + g * (AT_KB * 16), a.kv + ((int64_t)(a.v_chunk0 + head * 5 + g) * a.kv_rows + krow0) * 8,
                 ncopy * 16, bar_v);
    }
    // zero the padding slab (d = 40..47) of Q and K, and V rows that were not fetched
    {
      const uint4 z = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(sQ + 5 * (AT_M * 16) + tid * 16) = z;
      *reinterpret_cast<uint4*>(sK + 5 * (AT_KB * 16) + tid * 16) = z;
      *reinterpret_cast<uint4*>(sK + 5 * (AT_KB * 16) + (tid + 128) * 16) = z;
      if (!WINDOW && ncopy < AT_KB)
        for (int i = tid; i < 5 * (AT_KB - ncopy); i += 128) {
          const int g = i / (AT_KB - ncopy), r = ncopy + i % (AT_KB - ncopy);
          *reinterpret_cast<uint4*>(sV + g * (AT_KB * 16) + r * 16) = z;
        }
      fence_proxy_async();
    }
    __syncthreads();

    // ---- S = Q K^T : 3 x (128 x 256 x 16) ----------------------------------------------
    if (tid == 0) {
      mbar_wait(bar_qk, ph_qk);
      tc_fence_after();
      constexpr uint32_t IDESC_S = make_idesc(AT_M, AT_KB);
      const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK);
#pragma unroll
      for (int ks = 0; ks < 3; ++ks)
        umma_bf16(tmem_base, make_desc(qa + ks * 2 * (AT_M * 16), AT_M * 16, 128),
                  make_desc(ka + ks * 2 * (AT_KB * 16), AT_KB * 16, 128), IDESC_S, ks > 0);
      umma_commit(bar_mma);
    }
    ph_qk ^= 1;
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tc_fence_after();

    // ---- softmax, thread <-> row ------------------------------------------------------------
    // columns this warp can need: WINDOW [32w, 32w+160), CROSS all 256
    const int cbeg = WINDOW ? warp * 32 : 0;
    const int cend = WINDOW ? warp * 32 + 160 : AT_KB;
    auto valid = [&](int c) {
      const int j = j0 + c;
      if (WINDOW) return j >= 0 && j < a.Tk && j >= qi - WIN && j <= qi + WIN;
      return j < a.Tk;
    };
    float bmax = -INFINITY;
    for (int c0 = cbeg; c0 < cend; c0 += 16) {
      float s[16];
      tmem_ld16(trow + c0, s);
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (valid(c0 + j)) bmax = fmaxf(bmax, s[j]);
    }
    const float m_new = fmaxf(m_run, bmax * a.scale_log2e);
    const float alpha = (m_run == -INFINITY) ? 0.f : exp2f(m_run - m_new);
    float lsum = 0.f;
    // all warps must be done reading Q|K through the tensor core (they are: bar_mma) before P overwrites them
    for (int c0 = 0; c0 < AT_KB; c0 += 16) {
      uint4 p0 = make_uint4(0, 0, 0, 0), p1 = p0;
      if (c0 >= cbeg && c0 < cend) {      // warp-uniform
        float s[16];
        tmem_ld16(trow + c0, s);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float p = valid(c0 + j) ? exp2f(s[j] * a.scale_log2e - m_new) : 0.f;
          lsum += p;
          s[j] = p;
        }
        p0 = pack_f16x8(s);                  // P and V are f16 operands
        p1 = pack_f16x8(s + 8);
      }
      *reinterpret_cast<uint4*>(sP + (c0 >> 3) * (AT_M * 16) + tid * 16) = p0;
      *reinterpret_cast<uint4*>(sP + ((c0 >> 3) + 1) * (AT_M * 16) + tid * 16) = p1;
    }
    l_run = l_run * alpha + lsum;
    m_run = m_new;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();

    // ---- O_blk = P V : 16 x (128 x 48 x 16), V is the MN-major B operand ----------------
    if (tid == 0) {
      mbar_wait(bar_v, ph_v);
      tc_fence_after();
      constexpr uint32_t IDESC_O = make_idesc_f16(AT_M, 48, /*b_mn_major=*/true);
      const uint32_t pa = smem_u32(sP), va = smem_u32(sV);
#pragma unroll
      for (int ks = 0; ks < AT_KB / 16; ++ks)
        umma_bf16(tmem_base, make_desc(pa + ks * 2 * (AT_M * 16), AT_M * 16, 128),
                  make_desc(va + ks * 2 * 128, /*LBO: next 8 keys*/ 128, /*SBO: next 8 dims*/ AT_KB * 16), IDESC_O,
                  ks > 0);
      umma_commit(bar_mma);
    }
    ph_v ^= 1;
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tc_fence_after();
    {
      float ob[HD];
      tmem_ld16(trow, ob);
      tmem_ld16(trow + 16, ob + 16);
      tmem_ld8(trow + 32, ob + 32);
#pragma unroll
      for (int d = 0; d < HD; ++d) o_acc[d] = o_acc[d] * alpha + ob[d];
    }
    tc_fence_before();
  }

  // ---- epilogue: normalise, bf16, chunk-major store (A operand of the following proj GEMM) ----
  if (qi < a.Tq) {
    const float inv = 1.0f / l_run;
#pragma unroll
    for (int g = 0; g < 5; ++g) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = o_acc[g * 8 + j] * inv;
      *reinterpret_cast<uint4*>(a.o + ((int64_t)(head * 5 + g) * a.q_rows + qrow0 + tid) * 8) = pack_bf16x8(v);
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc<AttnSmem::TMEM_COLS>(tmem_base);
}

int launch_tc_attn_window(const __nv_bfloat16* qkv, __nv_bfloat16* o, int B, int T, cudaStream_t st);
int launch_tc_attn_cross(const __nv_bfloat16* q, const void* kv_chunk, __nv_bfloat16* o, int B, int T, int S,
                         cudaStream_t st);

}  // namespace tc
}  // namespace edtts
