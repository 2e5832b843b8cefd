// Fused DiffusionTransformerBlock for the bf16 tensor-core path (layers/transformer.py:141-160).
//
// One persistent CTA per SM; a work item is one (layer, utterance, 128-frame tile), and ONE launch runs a whole decoder
// evaluation: the head (in_proj + positional embedding + QKV of block 0) and the four blocks, items numbered layer-major
// and dealt round-robin, per-item completion flags between the layers (MegaArgs below).  For its tile the
// CTA runs the whole block on chip -- the only HBM traffic of a layer is: q/k/v of the tile
// (+ the +-64 frame halo of k, v) in, the fp32 residual stream h in and out, and the layer's
// weights / the utterance's context K, V streamed from L2:
//
//   h (128 x 160 fp32) lives in TENSOR MEMORY for the whole layer: every output projection is a
//   tcgen05.mma that ACCUMULATES into the h columns, so the three residual adds cost nothing.
//
//   window attention (attention.py:94-111)   4 heads, two in flight (one per compute warpgroup)
//   h += O Wproj^T                            tcgen05.mma into h
//   n2 = RMSNorm(h + b_proj) w2               row pass, TMEM -> regs -> bf16 A operand (smem)
//   q  = n2 Wq^T                              tcgen05.mma into scratch columns, -> bf16 Q operand
//   cross attention (mla.py:176-180)          same machinery, keys = the S context tokens
//   h += O Wout^T
//   n3 = AdaRMSNorm(h)                        row pass
//   u  = swiglu(n3 W0^T + b0)                 four quarters of 80 u-columns (x 80 | gate 80 = one 160-row weight chunk), two
//                                             TMEM buffers: the GEMM of quarter k+1 runs under the SwiGLU pass of quarter k
//   h += u W3^T (+ b3)                        tcgen05.mma into h (first half under the last SwiGLU pass), then h -> HBM
//
// 384 threads, warp-specialised:
//   warps 0-3 / 4-7   compute warpgroups (thread <-> frame, TMEM lane = frame).  In the attention
//                     phases warpgroup g owns heads g and g+2; in the row passes it handles columns
//                     80g .. 80g+79 of every row.
//   warp 8 / 10       TMA producer of warpgroup 0 / 1 (lane 0): K (4 stages) and V (2 stages) blocks of 64 keys; between
//                     two items it prefetches the next item's h / q | k | v rows into L2.  Lane 1 of warp 8 is the
//                     "agent" of the merged launch: gpu-scope acquire / release of the per-item flags
//   warp 9 / 11       MMA issuer of warpgroup 0 / 1 (converged warp, elected lane):  S = Q K^T into a
//                     double-buffered TMEM block, O += P V accumulating in TMEM.  Warp 11 is also the GEMM-CHAIN issuer: between
//                     the attention phases it streams the weight chunks and issues every projection / FFN / QKV GEMM of the
//                     item, woken by per-warp arrivals of the compute warps (bar_go / bar_free), so no compute warp is ever
//                     held by an MMA queue and the row pass of one GEMM overlaps the next GEMM
// Attention is a single streaming pass per head in which no thread waits for a tensor-core round
// trip: S blocks (64 keys) arrive in a double-buffered TMEM block, p = exp2(s*c - m) is written back as
// f16x2 over the first 32 columns of the SAME S block (tcgen05.st) and feeds O += P V as a tensor-memory
// A operand, so P never touches shared memory.  m is a lazily updated running row maximum: only when a
// block raises it by more than 2^6 does the warp rescale its O rows in TMEM (tcgen05.ld / st) -- after
// the first block or two of a head that never happens -- so O is read once per head and the result is
// exact (O and the row sum, delivered by the tensor core through V's constant-one column, carry the same
// factor).
// All hand-offs are mbarriers (full/free pairs per buffer); the only CTA-wide barriers are the
// named barriers between GEMM-chain stages.  Weights are streamed from L2 in 160 x 160 bf16 chunks
// (51,200 B, pre-packed UMMA operand images) through two slots, prefetched one GEMM ahead.
#pragma once
#include "umma.cuh"
#include "tc_path.cuh"

namespace edtts {
namespace tc {

constexpr int LY_THREADS = 384;
constexpr int LY_CTHREADS = 256;              // compute threads
constexpr int LY_SLAB = 128 * 16;             // one 8-wide K slab of a 128-row operand
constexpr int LY_WSLAB = 160 * 16;            // one 8-wide K slab of a 160-row weight chunk
constexpr int LY_WCHUNK = 20 * LY_WSLAB;      // 51,200 B: W[160 out][160 in] bf16
constexpr int LY_NCHUNK = 9;                  // proj, q_proj, out_proj, ffn0 x4, ffn3 x2
// ffn0 quarter q: rows 0..79 = ffn.net.0 rows 80q.. (x part of u columns 80q..80q+79), rows 80..159 = rows 320+80q.. (their gates)
enum LyChunk : int { WC_PROJ = 0, WC_Q, WC_OUT, WC_F0_Q0, WC_F0_Q1, WC_F0_Q2, WC_F0_Q3, WC_F3_K0, WC_F3_K1 };
constexpr int LY_KB = 64;                     // keys per attention block
constexpr int LY_KSLAB = LY_KB * 16;          // one 8-wide slab of a 64-key K / V block
constexpr int LY_KBUF = 6 * LY_KSLAB;         // K (or V) block, head_dim padded 40 -> 48

// per-layer constant vector (floats), packed by tc_layer.cu
constexpr int LC_PROJ_B = 0, LC_N2W = 160, LC_N3W = 320, LC_F0B = 480, LC_F3B = 1120, LC_COUNT = 1280;
// shared-memory copy: the above + the tile's AdaLN vectors g3 = w3 * (1 + scale), sh3 = shift
constexpr int LS_G3 = 1280, LS_SH3 = 1440;
// tail: bias added to the finished h (b3 or in_proj.bias), next norm's (gain, shift) or final_norm (w, b), out_proj.bias
constexpr int LS_TB = 1600, LS_TG = 1760, LS_TS = 1920, LS_OB = 2080, LS_COUNT = 2160;

// shared memory map (bytes)
constexpr int LO_A = 0;                                   // 21 slabs: A operand / Q / attention output
constexpr int LO_W0 = LO_A + 21 * LY_SLAB;                // weight slot 0
constexpr int LO_X = LO_W0 + LY_WCHUNK;                   // overlay region
constexpr int LO_W1 = LO_X;                               //   GEMM chain: weight slot 1
constexpr int LO_U = LO_X + LY_WCHUNK;                    //   GEMM chain: u quarters 0..2 (128 x 240 bf16; quarter 3 goes to sA)
constexpr int LY_KST = 4, LY_VST = 2;                     //   K / V stages per warpgroup (S blocks: 2, so K stage = S op % 4)
constexpr int LO_KV = LO_U;                               //   attention: wg: K0 | K1 | K2 | V0 | V1 (K/V stream in while
                                                          //   weight slot 1 is in use, so they only overlay the u half)
constexpr int LO_X_END = LO_KV + 2 * (LY_KST + LY_VST) * LY_KBUF;
static_assert(LO_U + 30 * LY_SLAB <= LO_X_END, "u quarters 0..2 must fit in the overlay region");
constexpr int LO_CONST = LO_X_END;
constexpr int LO_RED = LO_CONST + LS_COUNT * 4;
constexpr int LO_BAR = LO_RED + 4 * 128 * 4;
constexpr int LY_NBAR = 13 + 2 * 16;
constexpr int LY_SMEM = LO_BAR + LY_NBAR * 8 + 16;
static_assert(LY_SMEM <= 232448, "shared memory budget");

// mbarrier indices
// LB_G / LB_G2: GEMM-chain completions (tcgen05.commit); LB_GO: "A operand written" and LB_FREE[2]: "TMEM buffer b read out / result
// seen" -- one arrival per compute warp (count 8), waited for by the GEMM-chain issuer
enum LyBar : int { LB_W0 = 0, LB_W1, LB_Q, LB_G, LB_KVGO, LB_ATTGO, LB_DEP = 6 /* two */, LB_FIN = 8, LB_G2 = 9, LB_GO = 10,
                   LB_FREE = 11 /* two */, LB_WG0 = 13 };
// One commit per MMA group: WB_SK[n % 4] = "S op n retired" tells the softmax warps that S block n % 2 is full AND the TMA
// producer that K stage n % 4 is free; WB_PV[n % 2] = "P V op n retired" frees V stage / P block n % 2 (lazy rescale).
enum LyWgBar : int { WB_KFULL = 0, WB_SK = 4, WB_VFULL = 8, WB_PV = 10, WB_PFULL = 12, WB_OFULL = 14, WB_OFREE = 15, WB_COUNT = 16 };
static_assert(LY_KST == 4 && LY_VST == 2, "barrier indexing assumes 4 K stages, 2 V stages, 2 S blocks");

// tensor memory map (columns)
constexpr uint32_t TM_H = 0;                              // residual stream, 160 columns
constexpr uint32_t TM_G = 160;                            // GEMM chain scratch, 320 columns (FFN: two buffers of x 80 | gate 80)
constexpr uint32_t TM_S0 = 160, TM_WG = 176;              // attention, per warpgroup: S0 (64) | S1 (64) | O (48); the f16
                                                          // probabilities P(i) overwrite columns 0..31 of their S block


struct LayerArgs {
  int mode;                          // LM_BLOCK: one transformer block; LM_HEAD: h = in_proj(x_t) + pe instead (decoder.py:96-97)
  int tail;                          // what follows on the finished h rows: LT_QKV = the NEXT block's norm1 + QKV projection,
                                     // LT_FINAL = final_norm + out_proj + update rule (decoder.py:108-109, schedule.py)
  float* hc;                         // residual stream between launches, chunk-major fp32 [40][R][4], in place
  const __nv_bfloat16* qkv;          // [60][R][8] chunk-major q | k | v (f16) of this block (finite slack around it)
  const __nv_bfloat16* kvx;          // [40][RS][8] chunk-major context k | v (f16) of this block
  const __nv_bfloat16* wimg;         // LY_NCHUNK weight chunks
  const float* consts;               // LC_COUNT floats
  const float* mod3;                 // norm3 (scale | shift) of utterance 0; + b * mod_stride
  int mod_stride;
  int64_t R, RS;
  int B, T, S;
  int tiles_per_utt;
  float scale_log2e;                 // head_dim^-0.5 * log2(e)
  // LM_HEAD
  const float* x_t;                  // [R][80] fp32 (also the x_t of the fused update rule, LT_FINAL)
  const __nv_bfloat16* w_in;         // in_proj.weight as [10][160][8]
  const float* in_b;                 // [160]
  const float* pe;                   // pos_emb.pe [>= T][160]
  const float* pe_cm;                // the same table chunk-major [40][pe_rows][4] (coalesced thread-per-row reads), or null
  int pe_rows;
  // LT_QKV
  const __nv_bfloat16* w_qkv;        // 3 chunks (q, k, v) of the next block's attn.qkv.weight
  const float* n1w;                  // next block's norm1.norm.weight [160]
  const float* mod1;                 // next block's norm1 (scale | shift) of utterance 0; + b * mod_stride
  __nv_bfloat16* qkv_out;            // [60][R][8] of the next block
  // LT_FINAL
  const __nv_bfloat16* w_out;        // out_proj.weight as [20][80][8]
  const float* fn_w;                 // final_norm.weight / bias [160]
  const float* fn_b;
  const float* out_b;                // [80]
  int stop_phase;                    // debug: 1 = stop after attention + proj, 2 = after cross, 0 = whole block
  long long* phase_clocks;           // debug: [gridDim.x][24] cycles per phase / attention section (thread 0), or null
};

// One launch runs up to LY_MAXL consecutive "layer launches" of a decoder step (head, blocks 0..3) as ONE persistent
// grid: work item g = l * ntiles + tile, handed out round-robin in that order, so the CTAs that would idle in the last
// wave of layer l start on layer l + 1 instead (1792 tiles on 148 SMs: 13 rounds per layer separately, 12.1 merged).
// Item (l, n) reads what items (l-1, n-1), (l-1, n), (l-1, n+1) wrote (its h rows, q | k | v rows and their +-64 frame
// halo); `done` holds one flag per item, set (release, gpu scope) when the item's stores are out and polled (acquire)
// before the item starts.  Both are the job of one otherwise idle lane (the "agent": warp 8, lane 1), which hands the
// result to the compute threads / TMA producers through mbarriers, so no compute warp ever waits on a gpu-scope fence.
// The flags also order the write-after-read hazards of the double-buffered q | k | v (the readers of the rows an item
// overwrites are exactly its three predecessors) and the in-place h (row-local).
constexpr int LY_MAXL = 5;
struct MegaArgs {
  LayerArgs a[LY_MAXL];
  edtts_step_args step;              // update rule of the LT_FINAL tail (last layer only)
  int n_launch;
  int* done;                         // [n_launch][ntiles], zeroed before the launch; null: no dependencies (n_launch == 1)
};

// Development trace (-DLY_TRACE): CTA 0 records (event id, clock) pairs of four threads -- compute warp 0 / warp 4 lane 0,
// MMA issuer and TMA producer of warpgroup 0 -- into phase_clocks[slot * 1024 ..]; the launcher prints them.
#ifdef LY_TRACE
#define LY_TR(slot, id)                                                              \
  { if (blockIdx.x == 0 && a.phase_clocks && (threadIdx.x & 31) == 0) {                \
    long long* tb_ = a.phase_clocks + (slot) * 1024;                                 \
    const long long n_ = tb_[0];                                                     \
    if (n_ < 510) { tb_[2 + 2 * n_] = (id); tb_[3 + 2 * n_] = clock64(); tb_[0] = n_ + 1; } \
  } }
#else
#define LY_TR(slot, id) {}
#endif

// silu(g) = g * sigmoid(g) = 0.5 g (1 + tanh(g / 2)): one MUFU operation instead of ex2 + rcp
__device__ __forceinline__ float fast_silu(float g) {
  float t;
  const float hg = 0.5f * g;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(hg));
  return fmaf(hg, t, hg);
}

// D[128 x 160] (+)= A[128 x 160] * W[160 x 160]^T, one elected thread
__device__ __forceinline__ void ly_issue_gemm(uint32_t d_tmem, uint32_t a_addr, uint32_t w_addr, bool accumulate) {
  constexpr uint32_t IDESC = make_idesc(128, 160);
#pragma unroll
  for (int ks = 0; ks < 10; ++ks)
    umma_bf16(d_tmem, make_desc(a_addr + ks * 2 * LY_SLAB, LY_SLAB, 128), make_desc(w_addr + ks * 2 * LY_WSLAB, LY_WSLAB, 128),
              IDESC, accumulate || ks > 0);
}

// k-steps ks0 .. ks0+nks-1 of a 160-row weight chunk against A slabs starting at a_addr (accumulating)
__device__ __forceinline__ void ly_issue_gemm_part(uint32_t d_tmem, uint32_t a_addr, uint32_t w_addr, int ks0, int nks) {
  constexpr uint32_t IDESC = make_idesc(128, 160);
#pragma unroll
  for (int ks = 0; ks < nks; ++ks)
    umma_bf16(d_tmem, make_desc(a_addr + ks * 2 * LY_SLAB, LY_SLAB, 128), make_desc(w_addr + (ks0 + ks) * 2 * LY_WSLAB, LY_WSLAB, 128),
              IDESC, true);
}

// D[128 x n] (+)= A[128 x 16 ksteps] * W[n x 16 ksteps]^T; W slabs are n rows of 16 bytes
__device__ __forceinline__ void ly_issue_gemm_ex(uint32_t d_tmem, uint32_t a_addr, uint32_t w_addr, int ksteps, int n,
                                                 bool accumulate) {
  const uint32_t idesc = make_idesc(128, (uint32_t)n);
  for (int ks = 0; ks < ksteps; ++ks)
    umma_bf16(d_tmem, make_desc(a_addr + ks * 2 * LY_SLAB, LY_SLAB, 128), make_desc(w_addr + ks * 2 * n * 16, n * 16, 128), idesc,
              accumulate || ks > 0);
}

__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu(int* p, int v) { asm volatile("st.relaxed.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// all three predecessors of item g (previous layer, tiles n-1 .. n+1) have published their results
__device__ __forceinline__ bool ly_deps_ready(const int* done, int g, int ntiles) {
  const int l = g / ntiles;
  if (l == 0) return true;
  const int n = g - l * ntiles;
  const int* f = done + (l - 1) * ntiles;
  int ok = ld_relaxed_gpu(f + n);
  if (n > 0) ok &= ld_relaxed_gpu(f + n - 1);
  if (n + 1 < ntiles) ok &= ld_relaxed_gpu(f + n + 1);
  return ok != 0;
}

struct LyTile {
  int b, t0, nq;
  int64_t row0;
};
__device__ __forceinline__ LyTile ly_tile(const LayerArgs& a, int tile) {
  LyTile tl;
  tl.b = tile / a.tiles_per_utt;
  tl.t0 = (tile % a.tiles_per_utt) * 128;
  tl.nq = min(128, a.T - tl.t0);
  tl.row0 = (int64_t)tl.b * a.T + tl.t0;
  return tl;
}

// Key blocks of one attention phase.  Window: block kb (0..3) holds frames t0-64+64kb .. +63, blocks with no
// frame inside the utterance are skipped; cross: block i holds context tokens 64i .. 64i+63.
struct LyPlan {
  int kb0, nb;
};
template <bool WINDOW>
__device__ __forceinline__ LyPlan ly_plan(const LayerArgs& a, const LyTile& tl) {
  LyPlan p;
  if (WINDOW) {
    p.kb0 = tl.t0 == 0 ? 1 : 0;
    const int kb1 = min(3, (a.T - 1 - (tl.t0 - WIN)) >> 6);
    p.nb = kb1 - p.kb0 + 1;
  } else {
    p.kb0 = 0;
    p.nb = (a.S + LY_KB - 1) / LY_KB;
  }
  return p;
}
// keys the MMAs of block i cover (a multiple of 16)
template <bool WINDOW>
__device__ __forceinline__ int ly_nkeys(const LayerArgs& a, int i) {
  if (WINDOW) return LY_KB;
  return (min(LY_KB, a.S - i * LY_KB) + 15) & ~15;
}

// ---- TMA producer of one warpgroup, one attention phase (one thread) ------------------------------------------
template <bool WINDOW>
__device__ __forceinline__ void ly_tma_phase(const LayerArgs& a, const __nv_bfloat16* qkv, const __nv_bfloat16* kvx,
                                             const LyTile& tl, uint8_t* smem, uint64_t* wb, int wg, uint32_t& nk, uint32_t& nv) {
  const LyPlan pl = ly_plan<WINDOW>(a, tl);
  uint8_t* kv = smem + LO_KV + wg * ((LY_KST + LY_VST) * LY_KBUF);
  auto load = [&](int head, int i, bool is_v, uint32_t& cnt) {
    const uint32_t nst = is_v ? LY_VST : LY_KST;
    const uint32_t buf = cnt % nst, use = cnt / nst;
    uint64_t* full = wb + (is_v ? WB_VFULL : WB_KFULL) + buf;
    mbar_wait(wb + (is_v ? WB_PV : WB_SK) + buf, (use & 1) ^ 1);
    if (wg == 0) LY_TR(3, is_v ? 31 : 30)
    uint8_t* dst = kv + ((is_v ? LY_KST : 0) + buf) * LY_KBUF;
    if (WINDOW) {
      const int64_t g0 = tl.row0 - WIN + LY_KB * (pl.kb0 + i);
      const int c0 = (is_v ? 40 : 20) + head * 5;
#ifdef LY_KO_KV   // timing experiment only (wrong results): no K/V copies
      mbar_arrive(full);
#else
      mbar_expect_tx(full, 5 * LY_KSLAB);
#pragma unroll
      for (int g = 0; g < 5; ++g) bulk_g2s(dst + g * LY_KSLAB, qkv + ((int64_t)(c0 + g) * a.R + g0) * 8, LY_KSLAB, full);
#endif
    } else {
      const int nvalid = min(LY_KB, a.S - i * LY_KB);
      const int64_t g0 = (int64_t)tl.b * a.S + i * LY_KB;
      const int c0 = (is_v ? 20 : 0) + head * 5;
      if (is_v && (nvalid & 15)) {                        // keys nvalid .. round16(nvalid)-1 enter P V with p = 0: finite v needed
        for (int r = nvalid; r < ((nvalid + 15) & ~15); ++r)
#pragma unroll
          for (int g = 0; g < 5; ++g) *reinterpret_cast<uint4*>(dst + g * LY_KSLAB + r * 16) = make_uint4(0, 0, 0, 0);
        fence_proxy_async();
      }
#ifdef LY_KO_KV
      mbar_arrive(full);
#else
      mbar_expect_tx(full, 5 * nvalid * 16);
#pragma unroll
      for (int g = 0; g < 5; ++g) bulk_g2s(dst + g * LY_KSLAB, kvx + ((int64_t)(c0 + g) * a.RS + g0) * 8, nvalid * 16, full);
#endif
    }
    ++cnt;
  };
  for (int hh = 0; hh < 2; ++hh) {
    const int head = wg + 2 * hh;
    for (int i = 0; i < LY_KST && i < pl.nb; ++i) load(head, i, false, nk);
    for (int i = 0; i < LY_VST && i < pl.nb; ++i) load(head, i, true, nv);
    for (int i = 0; i < pl.nb; ++i) {                     // K(i+3) needs S(i) retired, V(i+2) needs P V(i) retired
      if (i + LY_KST < pl.nb) load(head, i + LY_KST, false, nk);
      if (i + LY_VST < pl.nb) load(head, i + LY_VST, true, nv);
    }
  }
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- MMA issuer of one warpgroup, one attention phase (a converged warp; one elected lane issues) ------------------
template <bool WINDOW>
__device__ __forceinline__ void ly_mma_phase(const LayerArgs& a, const LyTile& tl, uint8_t* smem, uint32_t tmem_base,
                                             uint64_t* wb, int wg, uint32_t& ns, uint32_t& nv, uint32_t& nh) {
  const LyPlan pl = ly_plan<WINDOW>(a, tl);
  const uint32_t sKV = smem_u32(smem + LO_KV + wg * ((LY_KST + LY_VST) * LY_KBUF));
  const uint32_t dS = tmem_base + TM_S0 + wg * TM_WG;
  const uint32_t dO = dS + 2 * LY_KB;
  // operand descriptors are built once per phase; a k-step only adds its byte offset (>> 4) to the start-address field
  const uint64_t dq0 = make_desc(smem_u32(smem + LO_A), LY_SLAB, 128);
  const uint64_t dk0 = make_desc(sKV, LY_KSLAB, 128);
  const uint64_t dv0 = make_desc(sKV + LY_KST * LY_KBUF, /*LBO: next 8 keys*/ 128, /*SBO: next 8 dims*/ LY_KSLAB);
  constexpr uint32_t IDESC_S = make_idesc(128, LY_KB);
  constexpr uint32_t IDESC_O = make_idesc_f16(128, 48, /*b_mn_major=*/true);     // P, V are f16
  auto s_op = [&](int head, int i) {
    const uint32_t sbuf = ns & 1, kbuf = ns & 3;
    mbar_wait(wb + WB_KFULL + kbuf, (ns >> 2) & 1);
    if (wg == 0) LY_TR(2, 20)
    // S block sbuf is free: its last reader is P V (ns - 2) (P lives in the block), issued before this MMA by this
    // very thread, and the tensor core executes its MMAs in issue order.
    tc_fence_after();
    const int nk = ly_nkeys<WINDOW>(a, i);
    const uint32_t idesc = (WINDOW || nk == LY_KB) ? IDESC_S : make_idesc(128, (uint32_t)nk);
    const uint64_t qa = dq0 + (uint64_t)(head * 5 * (LY_SLAB >> 4)), ka = dk0 + (uint64_t)(kbuf * (LY_KBUF >> 4));
    if (elect_one()) {
#ifndef LY_KO_AMMA   // (timing experiment only: no attention MMAs)
#pragma unroll
      for (int ks = 0; ks < 3; ++ks)
        umma_bf16(dS + sbuf * LY_KB, qa + (uint64_t)(ks * 2 * (LY_SLAB >> 4)), ka + (uint64_t)(ks * 2 * (LY_KSLAB >> 4)), idesc, ks > 0);
#endif
      umma_commit(wb + WB_SK + kbuf);
    }
    if (wg == 0) LY_TR(2, 21)
    ++ns;
  };
  auto pv_op = [&](int i, bool last) {
    const uint32_t buf = nv & 1, par = (nv >> 1) & 1;
    mbar_wait(wb + WB_PFULL + buf, par);
    if (wg == 0) LY_TR(2, 22)
    mbar_wait(wb + WB_VFULL + buf, par);
    if (wg == 0) LY_TR(2, 23)
    if (i == 0) mbar_wait(wb + WB_OFREE, (nh & 1) ^ 1);
    tc_fence_after();
    const uint32_t pa = dS + buf * LY_KB;                 // P: A operand in tensor memory
    const uint64_t va = dv0 + (uint64_t)(buf * (LY_KBUF >> 4));
    const int nks = ly_nkeys<WINDOW>(a, i) >> 4;
    if (elect_one()) {
#ifdef LY_KO_AMMA
      for (int ks = 0; ks < (i == 0 ? 1 : 0); ++ks)
#else
      for (int ks = 0; ks < nks; ++ks)
#endif
        umma_f16_ts(dO, pa + ks * 8, va + (uint64_t)(ks * (2 * 128 >> 4)), IDESC_O, i > 0 || ks > 0);
      umma_commit(wb + WB_PV + buf);
      if (last) umma_commit(wb + WB_OFULL);
    }
    if (wg == 0) LY_TR(2, 24)
    ++nv;
  };
  for (int hh = 0; hh < 2; ++hh) {
    const int head = wg + 2 * hh;
    s_op(head, 0);
    if (pl.nb > 1) s_op(head, 1);
    for (int i = 0; i < pl.nb; ++i) {
      pv_op(i, i == pl.nb - 1);
      if (i + 2 < pl.nb) s_op(head, i + 2);
    }
    ++nh;
  }
}

// one arrival per warp on a count-4 mbarrier, after every lane's preceding work
__device__ __forceinline__ void ly_warp_arrive(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// ---- compute warpgroup, one attention phase: heads wg and wg + 2 ------------------------------------------------------
// Q is in sA (slabs 5*head ..); the normalised output replaces it there.
//
// Per 64-key block and thread (= frame): 64 scores come out of TMEM into registers (the TMEM block is released
// at once, so the next S = Q K^T runs under this block's arithmetic), the running maximum is checked, and
//     p = cvt.f16x2(ex2(s * c - m))
// goes straight to the P operand (f16, 8 keys = one 16-byte store).  Two scores per MUFU operation and no
// running sum: V carries a constant 1 in its padding dimension 40 (ly_init_pads), so the tensor core delivers
// the row sum of exactly the rounded p as column 40 of O.
// two probabilities -> one packed f16x2.  MUFU.EX2 on f32 and one F2FP pack: measured 4.5-5 cycles per score and
// sub-partition (tools/ubench/expmix.cu); ex2.approx.f16x2 splits into two half-rate MUFU.EX2.F16 (8.1 per score).
__device__ __forceinline__ uint32_t ly_exp2_f16x2(float x_lo, float x_hi) {
#ifdef LY_KO_EXP   // timing experiment only (wrong results): no MUFU
  const __half2 h = __floats2half2_rn(x_lo * 0.001f, x_hi * 0.001f);
#else
  const __half2 h = __floats2half2_rn(ex2_approx(x_lo), ex2_approx(x_hi));
#endif
  return *reinterpret_cast<const uint32_t*>(&h);
}
// 2^x on the FMA / ALU pipes (no MUFU): x = n + r, r in [-0.5, 0.5]; 2^r by a cubic (max relative error 7.5e-5, below
// the f16 rounding of P), 2^n by adding n to the exponent field.  MUFU.EX2 issues once per 8 cycles per sub-partition
// (tools/ubench/expord.cu), so moving every LY_POLY_EVERY-th pair of scores here shortens the MUFU-bound pass.
#ifndef LY_POLY_EVERY
#define LY_POLY_EVERY 0
#endif
// -DEDTTS_DEBUG_CLOCKS builds: per-phase clocks of thread 0 only (cheap); -DLY_SOFTMAX_CLOCKS=1 adds the per-section counters
// inside the softmax steps (~50 registers, they change the schedule of the loop they measure)
#ifndef LY_SOFTMAX_CLOCKS
#define LY_SOFTMAX_CLOCKS 0
#endif
__device__ __forceinline__ float ly_exp2_poly(float x) {
  x = fmaxf(x, -30.0f);                                   // 2^-30 is far below the smallest f16: no exponent underflow
  const float t = x + 12582912.0f;                        // 1.5 * 2^23: the low mantissa bits of t hold round(x)
  const float r = x - (t - 12582912.0f);
  float p = fmaf(0.05517164245247841f, r, 0.2426111251115799f);
  p = fmaf(p, r, 0.6932609677314758f);
  p = fmaf(p, r, 0.9999280571937561f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ uint32_t ly_pack_f16x2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// (x0, x1) = (s0, s1) * c - m in ONE instruction: packed fp32 FMA (fma.rn.f32x2, sm_100)
__device__ __forceinline__ void ly_scale2(uint32_t s0, uint32_t s1, float c, float m, float& x0, float& x1) {
  uint64_t sp, cp2, mp, d;
  const float nm = -m;
  asm("mov.b64 %0, {%1, %2};" : "=l"(sp) : "r"(s0), "r"(s1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(cp2) : "f"(c));
  asm("mov.b64 %0, {%1, %1};" : "=l"(mp) : "f"(nm));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(sp), "l"(cp2), "l"(mp));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(d));
}
// ---- packed fp32 pairs (sm_100: add / mul / fma .f32x2 issue once for two values) -------------------------------------
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// row-pass arithmetic on PAIRS of columns (one issue slot for two values)
__device__ __forceinline__ void f2_add_to(float* v, const float* b) {           // v[0..1] += b[0..1]
  f2_unpack(f2_add(f2_pack(v[0], v[1]), f2_pack(b[0], b[1])), v[0], v[1]);
}
__device__ __forceinline__ uint64_t f2_sq_acc(const float* v, uint64_t acc) {    // acc += v * v (two partial sums)
  const uint64_t p = f2_pack(v[0], v[1]);
  return f2_fma(p, p, acc);
}
__device__ __forceinline__ float f2_hsum(uint64_t acc) {
  float a, b;
  f2_unpack(acc, a, b);
  return a + b;
}
// (x + bx) * silu(g + bg) for two columns: 5 packed operations + 2 tanh instead of 12 scalar ones
__device__ __forceinline__ void swiglu2(float* x, const float* g, const float* bx, const float* bg) {
  const uint64_t xb = f2_add(f2_pack(x[0], x[1]), f2_pack(bx[0], bx[1]));
  const uint64_t hg = f2_mul(f2_add(f2_pack(g[0], g[1]), f2_pack(bg[0], bg[1])), f2_pack(0.5f, 0.5f));
  float h0, h1, t0, t1;
  f2_unpack(hg, h0, h1);
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
  f2_unpack(f2_mul(xb, f2_fma(hg, f2_pack(t0, t1), hg)), x[0], x[1]);             // silu(g) = hg * tanh(hg) + hg
}
// 2^x for a pair on the FMA / ALU pipes (no MUFU), see ly_exp2_poly: 2 FMNMX + 3 packed adds + 3 packed FMAs + 2 exponent
// inserts for two results (11 issue slots with the pack, against 16 MUFU cycles for the same two on the MUFU pipe)
__device__ __forceinline__ uint32_t ly_exp2_poly2_f16x2(float x0, float x1) {
  const uint64_t x = f2_pack(fmaxf(x0, -30.0f), fmaxf(x1, -30.0f));
  const uint64_t magic = f2_pack(12582912.0f, 12582912.0f), nmagic = f2_pack(-12582912.0f, -12582912.0f);
  const uint64_t t = f2_add(x, magic);                    // low mantissa bits of t: round(x)
  const uint64_t tn = f2_add(t, nmagic);                  // round(x) as a float
  float tn0, tn1;
  f2_unpack(tn, tn0, tn1);
  const uint64_t r = f2_add(x, f2_pack(-tn0, -tn1));      // x - round(x) in [-0.5, 0.5]
  uint64_t p = f2_fma(f2_pack(0.05517164245247841f, 0.05517164245247841f), r, f2_pack(0.2426111251115799f, 0.2426111251115799f));
  p = f2_fma(p, r, f2_pack(0.6932609677314758f, 0.6932609677314758f));
  p = f2_fma(p, r, f2_pack(0.9999280571937561f, 0.9999280571937561f));
  float p0, p1, t0, t1;
  f2_unpack(p, p0, p1);
  f2_unpack(t, t0, t1);
  const float e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  const float e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
  const __half2 h = __floats2half2_rn(e0, e1);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// TMEM -> registers, 64 columns, NOT waited for
__device__ __forceinline__ void ly_s_issue(uint32_t taddr, uint32_t (&r)[64]) {
#pragma unroll
  for (int c = 0; c < 2; ++c)
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[32 * c + 0]), "=r"(r[32 * c + 1]), "=r"(r[32 * c + 2]), "=r"(r[32 * c + 3]), "=r"(r[32 * c + 4]),
          "=r"(r[32 * c + 5]), "=r"(r[32 * c + 6]), "=r"(r[32 * c + 7]), "=r"(r[32 * c + 8]), "=r"(r[32 * c + 9]),
          "=r"(r[32 * c + 10]), "=r"(r[32 * c + 11]), "=r"(r[32 * c + 12]), "=r"(r[32 * c + 13]), "=r"(r[32 * c + 14]),
          "=r"(r[32 * c + 15]), "=r"(r[32 * c + 16]), "=r"(r[32 * c + 17]), "=r"(r[32 * c + 18]), "=r"(r[32 * c + 19]),
          "=r"(r[32 * c + 20]), "=r"(r[32 * c + 21]), "=r"(r[32 * c + 22]), "=r"(r[32 * c + 23]), "=r"(r[32 * c + 24]),
          "=r"(r[32 * c + 25]), "=r"(r[32 * c + 26]), "=r"(r[32 * c + 27]), "=r"(r[32 * c + 28]), "=r"(r[32 * c + 29]),
          "=r"(r[32 * c + 30]), "=r"(r[32 * c + 31])
        : "r"(taddr + 32 * c)
        : "memory");
}
// wait for the loads above; the registers are tied to the wait so that no use can be scheduled before it
__device__ __forceinline__ void ly_s_wait(uint32_t (&r)[64]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
  asm volatile(""
               : "+r"(r[32]), "+r"(r[33]), "+r"(r[34]), "+r"(r[35]), "+r"(r[36]), "+r"(r[37]), "+r"(r[38]), "+r"(r[39]),
                 "+r"(r[40]), "+r"(r[41]), "+r"(r[42]), "+r"(r[43]), "+r"(r[44]), "+r"(r[45]), "+r"(r[46]), "+r"(r[47]),
                 "+r"(r[48]), "+r"(r[49]), "+r"(r[50]), "+r"(r[51]), "+r"(r[52]), "+r"(r[53]), "+r"(r[54]), "+r"(r[55]),
                 "+r"(r[56]), "+r"(r[57]), "+r"(r[58]), "+r"(r[59]), "+r"(r[60]), "+r"(r[61]), "+r"(r[62]), "+r"(r[63])
               :
               : "memory");
}

// K padding slabs (dims 40..47) = 0, V padding slabs = (1, 0, .., 0): the GEMM-chain buffers overlay them, so the
// compute threads rewrite them before every attention phase (640 16-byte stores per CTA)
__device__ __forceinline__ void ly_init_pads(uint8_t* smem, int tid) {
  for (int i = tid; i < 2 * (LY_KST + LY_VST) * LY_KB; i += LY_CTHREADS) {
    const int r = i % LY_KB, b = (i / LY_KB) % (LY_KST + LY_VST), w = i / (LY_KB * (LY_KST + LY_VST));
    uint8_t* slab = smem + LO_KV + (w * (LY_KST + LY_VST) + b) * LY_KBUF + 5 * LY_KSLAB;
    *reinterpret_cast<uint4*>(slab + r * 16) = make_uint4(b >= LY_KST ? 0x00003C00u : 0u, 0u, 0u, 0u);
  }
}

template <bool WINDOW, bool PROF>
__device__ __forceinline__ void ly_softmax_phase(const LayerArgs& a, const LyTile& tl, uint8_t* smem, uint32_t tmem_base,
                                                 uint64_t* wb, uint32_t& cs, uint32_t& cp, uint32_t& co, long long* fc) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  long long fc_last = PROF ? clock64() : 0;
  long long fcr[PROF ? 8 : 1] = {0};                      // register-resident (static indices); flushed at the end
#define LY_FC(i)                              \
  if (PROF) {                                 \
    const long long now_ = clock64();         \
    fcr[i] += now_ - fc_last;                 \
    fc_last = now_;                           \
  }
  const int wg = warp >> 2, lq = warp & 3, row = lq * 32 + lane;
  const LyPlan pl = ly_plan<WINDOW>(a, tl);
  uint8_t* sA = smem + LO_A;
  const uint32_t tS = tmem_base + ((uint32_t)(lq * 32) << 16) + TM_S0 + wg * TM_WG;
  const uint32_t tO = tS + 2 * LY_KB;
  const float c = a.scale_log2e;

  float m_run = -INFINITY;                                // running row maximum in exp2 units (s * c), lazily updated
  const int row0u = __shfl_sync(0xffffffffu, lq, 0) * 32; // first row of the warp, provably warp-uniform

  // one block of 64 keys
  auto step = [&](uint32_t (&cur)[64], int i) {
    const uint32_t pbuf = cp & 1;
    LY_FC(0)
    if (lq == 0) LY_TR(wg, 1)
    mbar_wait(wb + WB_SK + (cs & 3), (cs >> 2) & 1);
    if (lq == 0) LY_TR(wg, 2)
    tc_fence_after();
    ly_s_issue(tS + (cs & 1) * LY_KB, cur);               // both 32-column loads in flight, one wait
    ly_s_wait(cur);
    const uint32_t tP = tS + (cs & 1) * LY_KB;            // P(i) replaces S(i) in place (this thread's lane only)
    ++cs;
    if (lq == 0) LY_TR(wg, 3)
    LY_FC(1)
    // valid key columns [lo, hi] of this thread's row inside block i (empty if hi < lo)
    int lo, hi;
    if (WINDOW) {
      const int off0 = LY_KB * (pl.kb0 + i);              // band offset (frame - (t0 - 64)) of column 0
      lo = max(max(row, WIN - tl.t0) - off0, 0);
      hi = min(min(row + 2 * WIN, a.T - 1 - tl.t0 + WIN) - off0, LY_KB - 1);
    } else {
      lo = 0;
      hi = min(LY_KB, a.S - i * LY_KB) - 1;
    }
    int lo_a = 0, lo_b = 0, hi_a = 0, hi_b = 0;           // the same bounds for the warp's first (a) and last (b) row
    if (WINDOW) {
      const int off0 = LY_KB * (pl.kb0 + i);
      lo_a = max(max(row0u, WIN - tl.t0) - off0, 0);
      lo_b = max(max(row0u + 31, WIN - tl.t0) - off0, 0);
      hi_a = min(min(row0u + 2 * WIN, a.T - 1 - tl.t0 + WIN) - off0, LY_KB - 1);
      hi_b = min(min(row0u + 31 + 2 * WIN, a.T - 1 - tl.t0 + WIN) - off0, LY_KB - 1);
    }
    const bool any = hi >= lo;
    if (!any) lo = hi = LY_KB;                            // empty: no column index below 64 passes the unsigned range test
    const int base = -lo;
    const unsigned span = (unsigned)(hi - lo);
    const int nch = (ly_nkeys<WINDOW>(a, i) + 31) >> 5;
    bool need[2], full[2];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      if (WINDOW) {
        // warp-uniform without votes: lo and hi are non-decreasing in the row, so the warp's first / last row bound them.
        // need may be true for a chunk no single row touches (then every p is masked to 0); full is exact.
        need[ch] = ch < nch && lo_a <= ch * 32 + 31 && hi_b >= ch * 32 && hi_b >= lo_a;
        full[ch] = lo_b <= ch * 32 && hi_a >= ch * 32 + 31;
      } else {                                            // context keys: the same range for every row, no votes needed
        need[ch] = ch < nch;
        full[ch] = hi >= ch * 32 + 31;
      }
    }
    if (lq == 0) LY_TR(wg, 9)
    // ---- block maximum, lazy update of the running maximum ------------------------------------------------
    float b0 = -INFINITY, b1 = -INFINITY, b2 = -INFINITY, b3 = -INFINITY;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      if (!need[ch]) continue;
      if (full[ch]) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {                 // FMNMX3: two scores per instruction, four chains
          b0 = fmaxf(fmaxf(b0, __uint_as_float(cur[32 * ch + j])), __uint_as_float(cur[32 * ch + j + 1]));
          b1 = fmaxf(fmaxf(b1, __uint_as_float(cur[32 * ch + j + 2])), __uint_as_float(cur[32 * ch + j + 3]));
          b2 = fmaxf(fmaxf(b2, __uint_as_float(cur[32 * ch + j + 4])), __uint_as_float(cur[32 * ch + j + 5]));
          b3 = fmaxf(fmaxf(b3, __uint_as_float(cur[32 * ch + j + 6])), __uint_as_float(cur[32 * ch + j + 7]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          b0 = fmaxf(b0, (unsigned)(base + 32 * ch + j) <= span ? __uint_as_float(cur[32 * ch + j]) : -INFINITY);
      }
    }
#ifdef LY_KO_MAX   // timing experiment only (wrong results): no block maximum
    const float bmax = (i == 0 ? 8.0f : m_run);
#else
    const float bmax = fmaxf(fmaxf(b0, b1), fmaxf(b2, b3)) * c;
#endif
    if (lq == 0) LY_TR(wg, 10)
    // The new running maximum is a per-thread decision, so the exponentials below start at once; whether the WARP has
    // to rescale its O rows (tcgen05.ld / st are warp-collective) is voted on after the P store, off the critical chain.
    const bool grow = bmax > m_run + 6.0f;                // also true for the first finite block maximum
    const float m_old = m_run;
    m_run = grow ? bmax : m_run;
    const float m_use = (m_run == -INFINITY) ? 0.f : m_run;
    if (lq == 0) LY_TR(wg, 11)
    LY_FC(2)
    // ---- p -> P (f16x2 per 32-bit column, tensor memory): the A operand of P V, no shared-memory round trip ----
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      if (ch >= nch) continue;
      uint32_t pk[16];
      if (!need[ch]) {
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = 0u;
      } else if (full[ch]) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float x0, x1;
          ly_scale2(cur[32 * ch + 2 * j], cur[32 * ch + 2 * j + 1], c, m_use, x0, x1);
          if (LY_POLY_EVERY > 0 && j % (LY_POLY_EVERY > 0 ? LY_POLY_EVERY : 1) == 0) pk[j] = ly_exp2_poly2_f16x2(x0, x1);
          else pk[j] = ly_exp2_f16x2(x0, x1);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float x0, x1;
          ly_scale2(cur[32 * ch + 2 * j], cur[32 * ch + 2 * j + 1], c, m_use, x0, x1);
          pk[j] = ly_exp2_f16x2((unsigned)(base + 32 * ch + 2 * j) <= span ? x0 : -INFINITY,
                                (unsigned)(base + 32 * ch + 2 * j + 1) <= span ? x1 : -INFINITY);
        }
      }
      tmem_st16u(tP + 16 * ch, pk);
    }
    LY_FC(3)
    if (lq == 0) LY_TR(wg, 4)
    if (__any_sync(0xffffffffu, grow) && i > 0) {         // rescale this warp's O rows (and the row sum in column 40)
      const float f = (m_run == m_old) ? 1.0f : ex2_approx(m_old - m_run);       // m_old = -inf -> 0
      mbar_wait(wb + WB_PV + (pbuf ^ 1), ((cp - 1) >> 1) & 1);                   // P V (i-1) must have retired
      tc_fence_after();
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        float ob[16];
        tmem_ld16(tO + 16 * q, ob);
#pragma unroll
        for (int d = 0; d < 16; ++d) ob[d] *= f;
        tmem_st16(tO + 16 * q, ob);
      }
    }
    tmem_st_wait();
    tc_fence_before();
    if (lq == 0) LY_TR(wg, 5)
    ly_warp_arrive(wb + WB_PFULL + pbuf, lane);
    if (lq == 0) LY_TR(wg, 6)
    ++cp;
    LY_FC(4)
  };

  for (int hh = 0; hh < 2; ++hh) {
    const int head = wg + 2 * hh;
    uint32_t sc[64];
    m_run = -INFINITY;
    for (int i = 0; i < pl.nb; ++i) step(sc, i);

    // ---- head done: O / l replaces Q_h in sA ---------------------------------------------------------------------
    LY_FC(6)
    if (lq == 0) LY_TR(wg, 7)
    mbar_wait(wb + WB_OFULL, co & 1);
    if (lq == 0) LY_TR(wg, 8)
    tc_fence_after();
    LY_FC(7)
    float ob[48];
    tmem_ld32(tO, ob);
    tmem_ld16(tO + 32, ob + 32);
    tc_fence_before();
    ly_warp_arrive(wb + WB_OFREE, lane);
    ++co;
    const float inv = __fdividef(1.0f, ob[HD]);           // column 40 = sum of p (V's constant-one dimension)
#pragma unroll
    for (int g = 0; g < 5; ++g) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = ob[g * 8 + j] * inv;
      *reinterpret_cast<uint4*>(sA + (head * 5 + g) * LY_SLAB + row * 16) = pack_bf16x8(v);
    }
  }
  fence_proxy_async();
  if (PROF && fc)
    for (int i = 0; i < 8; ++i) fc[i] += fcr[i];
}

// PROF = true: per-phase cycle counters (debug builds of the launch only; they cost ~50 registers)
template <bool PROF>
__global__ void __launch_bounds__(LY_THREADS, 1) tc_layer_kernel(const __grid_constant__ MegaArgs p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem + LO_A;
  uint8_t* sW0 = smem + LO_W0;
  uint8_t* sW1 = smem + LO_W1;
  uint8_t* sU = smem + LO_U;
  float* sC = reinterpret_cast<float*>(smem + LO_CONST);
  float* sRed = reinterpret_cast<float*>(smem + LO_RED);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LO_BAR);
  uint64_t* bar_w0 = bars + LB_W0;
  uint64_t* bar_w1 = bars + LB_W1;
  uint64_t* bar_q = bars + LB_Q;
  uint64_t* bar_g = bars + LB_G;
  uint64_t* bar_kvgo = bars + LB_KVGO;
  uint64_t* bar_attgo = bars + LB_ATTGO;
  uint64_t* bar_dep = bars + LB_DEP;                      // [2]: item it's predecessors are visible (agent -> compute, TMA)
  uint64_t* bar_fin = bars + LB_FIN;                      // item finished, its stores are issued (compute -> agent)
  uint64_t* bar_g2 = bars + LB_G2;
  uint64_t* bar_go = bars + LB_GO;
  uint64_t* bar_free = bars + LB_FREE;                    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + LY_NBAR);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // finite shared memory everywhere (stale rows enter MMAs as 0 * x), zero pad slabs
  for (int i = tid * 16; i < LO_BAR; i += LY_THREADS * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < LB_WG0; ++i) mbar_init(bars + i, (i == LB_GO || i == LB_FREE || i == LB_FREE + 1) ? 8 : 1);
    for (int w = 0; w < 2; ++w) {
      uint64_t* wbi = bars + LB_WG0 + w * WB_COUNT;
      for (int i = 0; i < WB_COUNT; ++i) {
        const bool per_warp = (i >= WB_PFULL && i < WB_PFULL + 2) || i == WB_OFREE;
        mbar_init(wbi + i, per_warp ? 4 : 1);
      }
    }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // d: the fields every layer of the launch shares (sizes, scale, h / x_t pointers) -- read through a compile-time
  // index they stay direct constant-bank operands; only the per-layer pointers and modes go through p.a[l]
  const LayerArgs& d = p.a[0];
  const int ntiles = d.B * d.tiles_per_utt;
  const int nitems = ntiles * p.n_launch;

  // =========================== control warps: TMA producers and MMA issuers ===========================
  if (warp >= 8) {
    // register budget: the launch allocates 168 x 384 = 64,512; 128 x 88 + 256 x 208 = 64,512 (an .inc beyond the
    // pool would block forever)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    // The MMA issuer warps run CONVERGED (all 32 lanes walk the loop, one elected lane issues): the compiler then keeps
    // descriptors and barrier addresses in uniform registers and a tcgen05.mma costs its hardware floor to issue (24-48
    // cycles for these shapes) instead of ~56 cycles through per-instruction R2UR moves inside a divergent single-thread
    // branch (tools/ubench/mmarate.cu).  The warp index is made provably warp-uniform with a shuffle.
    const int cw = __shfl_sync(0xffffffffu, warp, 0) - 8;
    const int cwg = cw >> 1;
    const bool is_mma = cw & 1;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    if (cw == 3) {
      // ================= warp 11: attention MMAs of warpgroup 1 AND the whole GEMM chain (converged warp, elected lane) =========
      // The script mirrors the compute threads' item script step by step.  wait_go(): all eight compute warps have written (and
      // fenced) the A operand of the next GEMM and are done with the TMEM columns it overwrites; wait_free(b): they have read out
      // TMEM buffer b (FFN) / seen the previous completion of bar_g (QKV tail).  Every weight chunk is requested as soon as its
      // slot is free; completions go to bar_g / bar_g2 alternately so that two GEMMs can be in flight.
      uint64_t* cwb = bars + LB_WG0 + cwg * WB_COUNT;
      uint32_t n_phase = 0, c0 = 0, c1 = 0, c2 = 0;
      uint32_t pw0 = 0, pw1 = 0, ng = 0, ng2 = 0, ngo = 0, nf0 = 0, nf1 = 0;
      const uint32_t aA = smem_u32(sA), aW0 = smem_u32(sW0), aW1 = smem_u32(sW1), aU = smem_u32(sU), aU2 = aU + 20 * LY_SLAB;
      auto ld = [&](const __nv_bfloat16* src, int bytes, uint8_t* slot, uint64_t* bar) {
        if (elect_one()) {
          mbar_expect_tx(bar, bytes);
          bulk_g2s(slot, src, bytes, bar);
        }
        __syncwarp();
      };
      auto ld_first = [&](const LayerArgs& x) {            // first weight chunk of an item (slot 0)
        if (x.mode == LM_HEAD) ld(x.w_in, LY_WCHUNK / 2, sW0, bar_w0);
        else ld(x.wimg + (int64_t)WC_PROJ * (LY_WCHUNK / 2), LY_WCHUNK, sW0, bar_w0);
      };
      auto wait_w0 = [&]() { mbar_wait(bar_w0, pw0); pw0 ^= 1; tc_fence_after(); };
      auto wait_w1 = [&]() { mbar_wait(bar_w1, pw1); pw1 ^= 1; tc_fence_after(); };
      auto wait_go = [&]() { mbar_wait(bar_go, ngo & 1); ++ngo; tc_fence_after(); };
      auto wait_free = [&](int b) {
        if (b == 0) { mbar_wait(bar_free, nf0 & 1); ++nf0; } else { mbar_wait(bar_free + 1, nf1 & 1); ++nf1; }
        tc_fence_after();
      };
      auto wait_g = [&]() { mbar_wait(bar_g, (ng - 1) & 1); tc_fence_after(); };      // the latest commit on bar_g has retired
      auto wait_g2 = [&]() { mbar_wait(bar_g2, (ng2 - 1) & 1); tc_fence_after(); };
      // one 160 x 160 chunk GEMM, completion on bar_g (which = 0), bar_g2 (1) or none (-1)
      auto gemm = [&](uint32_t dcol, uint32_t a_addr, uint32_t w_addr, bool acc, int which) {
        if (elect_one()) {
          ly_issue_gemm(tmem_u + dcol, a_addr, w_addr, acc);
          if (which == 0) umma_commit(bar_g);
          else if (which == 1) umma_commit(bar_g2);
        }
        __syncwarp();
        if (which == 0) ++ng; else if (which == 1) ++ng2;
      };
      if ((int)blockIdx.x < nitems) ld_first(p.a[blockIdx.x / ntiles]);
      for (int g = blockIdx.x; g < nitems; g += gridDim.x) {
        const int l = g / ntiles;
        const LayerArgs& a = p.a[l];
        const LyTile tl = ly_tile(d, g - l * ntiles);
        const bool more = g + (int)gridDim.x < nitems;
        const LayerArgs& an = p.a[more ? (g + (int)gridDim.x) / ntiles : l];
        auto ldc = [&](int chunk, uint8_t* slot, uint64_t* bar) { ld(a.wimg + (int64_t)chunk * (LY_WCHUNK / 2), LY_WCHUNK, slot, bar); };
        if (a.mode == LM_HEAD) {
          wait_go();                                        // x tile (bf16) in sA
          wait_w0();
          if (elect_one()) {
            ly_issue_gemm_ex(tmem_u + TM_H, aA, aW0, M / 16, H, false);
            umma_commit(bar_g);
          }
          __syncwarp();
          ++ng;
          wait_g();                                         // slot 0 is free again
        } else {
          // ---- banded self-attention (warpgroup 1's heads) ----
          mbar_wait(bar_attgo, n_phase & 1);
          tc_fence_after();
          ly_mma_phase<true>(d, tl, smem, tmem_u, cwb, cwg, c0, c1, c2);
          ++n_phase;
          // ---- h += O Wproj^T ----
          wait_go();                                        // all heads' outputs are in sA; every attention MMA of the phase has retired
          if (a.stop_phase != 1) {
            if (elect_one()) mbar_arrive(bar_kvgo);         // context K/V may stream in during the GEMM chain
            __syncwarp();
          }
          ldc(WC_Q, sW1, bar_w1);
          wait_w0();
          gemm(TM_H, aA, aW0, true, 0);
          wait_g();
          if (a.stop_phase == 1) {                          // debug stop: drain the prefetch, hand slot 0 to the next item
            wait_w1();
            if (more) ld_first(an);
            continue;
          }
          ldc(WC_OUT, sW0, bar_w0);
          // ---- q = n2 Wq^T ----
          wait_go();
          wait_w1();
          gemm(TM_G, aA, aW1, false, 0);
          // ---- cross attention ----
          mbar_wait(bar_attgo, n_phase & 1);
          tc_fence_after();
          ly_mma_phase<false>(d, tl, smem, tmem_u, cwb, cwg, c0, c1, c2);
          ++n_phase;
          // ---- h += O Wout^T ----
          wait_go();
          ldc(WC_F0_Q0, sW1, bar_w1);
          wait_w0();
          gemm(TM_H, aA, aW0, true, 0);
          wait_g();
          if (a.stop_phase == 2) {
            wait_w1();
            if (more) ld_first(an);
            continue;
          }
          ldc(WC_F0_Q1, sW0, bar_w0);
          // ---- feed-forward: quarters 0..3 into TMEM buffers 0 / 1, then h += u W3^T ----
          wait_go();                                        // n3 in sA
          wait_w1();
          gemm(TM_G, aA, aW1, false, 0);                    // quarter 0
          wait_w0();
          gemm(TM_G + 160, aA, aW0, false, 1);              // quarter 1
          wait_g();
          ldc(WC_F0_Q2, sW1, bar_w1);
          wait_free(0);                                     // buffer 0 read out (u quarter 0 written)
          wait_w1();
          gemm(TM_G, aA, aW1, false, 0);                    // quarter 2
          wait_g2();
          ldc(WC_F0_Q3, sW0, bar_w0);
          wait_free(1);
          wait_w0();
          gemm(TM_G + 160, aA, aW0, false, 1);              // quarter 3
          wait_g();
          ldc(WC_F3_K0, sW1, bar_w1);
          wait_free(0);                                     // u quarters 0..2 written
          wait_w1();
          gemm(TM_H, aU, aW1, true, -1);                    // h += u[0..159] W3[:, 0..159]^T
          wait_g2();
          ldc(WC_F3_K1, sW0, bar_w0);
          wait_free(1);                                     // u quarter 3 written (sA)
          wait_w0();
          if (elect_one()) {                                // h += u[160..319] W3[:, 160..319]^T: quarter 2 from sU2, quarter 3 from sA
            ly_issue_gemm_part(tmem_u + TM_H, aU2, aW0, 0, 5);
            ly_issue_gemm_part(tmem_u + TM_H, aA, aW0, 5, 5);
            umma_commit(bar_g);
          }
          __syncwarp();
          ++ng;
          wait_g();                                         // both weight slots are free
        }
        // ---- tail ----
        if (a.tail == LT_QKV) {
          ld(a.w_qkv, LY_WCHUNK, sW1, bar_w1);
          ld(a.w_qkv + LY_WCHUNK / 2, LY_WCHUNK, sW0, bar_w0);
          wait_go();                                        // norm1 of the next block in sA, h read out
          wait_w1();
          gemm(0, aA, aW1, false, 0);                       // q
          wait_w0();
          gemm(160, aA, aW0, false, 1);                     // k
          wait_g();
          ld(a.w_qkv + 2 * (LY_WCHUNK / 2), LY_WCHUNK, sW1, bar_w1);
          wait_free(0);                                     // every compute warp has seen q's completion on bar_g
          wait_w1();
          gemm(320, aA, aW1, false, 0);                     // v
          wait_g2();
          if (more) ld_first(an);
        } else if (a.tail == LT_FINAL) {
          ld(a.w_out, LY_WCHUNK / 2, sW1, bar_w1);
          if (more) ld_first(an);
          wait_go();
          wait_w1();
          if (elect_one()) {
            ly_issue_gemm_ex(tmem_u + 0, aA, aW1, H / 16, M, false);
            umma_commit(bar_g);
          }
          __syncwarp();
          ++ng;
        } else {
          if (more) ld_first(an);
        }
      }
    } else if (is_mma || lane == 0) {
      uint64_t* cwb = bars + LB_WG0 + cwg * WB_COUNT;
      uint32_t n_phase = 0, c0 = 0, c1 = 0, c2 = 0;       // TMA: c0 = K loads, c1 = V loads; MMA: S ops, PV ops, heads
      uint32_t it = 0;                                    // the CTA's item counter (head items included)
      for (int g = blockIdx.x; g < nitems; g += gridDim.x, ++it) {
        const int l = g / ntiles;
        const LayerArgs& a = p.a[l];
        if (a.mode != LM_BLOCK) continue;
        const LyTile tl = ly_tile(d, g - l * ntiles);
        for (int ph = 0; ph < 2; ++ph) {
          if (ph == 1 && a.stop_phase == 1) break;
          if (!is_mma) {
            mbar_wait(bar_kvgo, n_phase & 1);
            if (ph == 0 && p.done) {                      // the window K/V (and halo) rows come from the previous layer's items;
                                                          // the agent's acquire + proxy fence precede the arrive on bar_dep
              mbar_wait(bar_dep + (it & 1), (it >> 1) & 1);
            }
            if (ph == 0) {
              ly_tma_phase<true>(d, a.qkv, a.kvx, tl, smem, cwb, cwg, c0, c1);
              // (prefetching the utterance's context K / V into L2 here was measured: 10.47 against 10.02 ms -- the seven CTAs of
              // an utterance ask for the same 256 KB and the producers fall behind; the 4 K / 2 V stages already cover it)
            } else {
              ly_tma_phase<false>(d, a.qkv, a.kvx, tl, smem, cwb, cwg, c0, c1);
            }
          } else {
            mbar_wait(bar_attgo, n_phase & 1);
            tc_fence_after();
            if (ph == 0) ly_mma_phase<true>(d, tl, smem, tmem_u, cwb, cwg, c0, c1, c2);
            else ly_mma_phase<false>(d, tl, smem, tmem_u, cwb, cwg, c0, c1, c2);
          }
          ++n_phase;
        }
#ifndef LY_NO_L2_PREFETCH
        // The producers idle from here to the next item's window phase: pull that item's h tile and q | k | v rows (+ halo)
        // into L2 -- they were written a dozen rounds ago by other SMs and have left it -- so that its prologue and first
        // K/V stages see L2 latency instead of HBM latency.  wg 0: h and q; wg 1: k and v.
        if (!is_mma) {
          const int gn = g + (int)gridDim.x;
          if (gn < nitems) {
            const int ln = gn / ntiles;
            const LayerArgs& an = p.a[ln];
            if (an.mode == LM_BLOCK) {
              const LyTile tn = ly_tile(d, gn - ln * ntiles);
              if (cwg == 0) {
                for (int c = 0; c < 40; ++c) bulk_prefetch_l2(d.hc + ((int64_t)c * d.R + tn.row0) * 4, tn.nq * 16);
                for (int c = 0; c < 20; ++c) bulk_prefetch_l2(an.qkv + ((int64_t)c * d.R + tn.row0) * 8, tn.nq * 16);
              } else {
                for (int c = 20; c < 60; ++c) bulk_prefetch_l2(an.qkv + ((int64_t)c * d.R + tn.row0 - WIN) * 8, 4 * LY_KB * 16);
              }
            }
          }
        }
#endif
      }
    } else if (cw == 0 && lane == 1 && p.done) {
      // ---- the agent: gpu-scope acquire of every item's predecessors, gpu-scope release of every finished item ----
      auto acquire = [&](int g, uint32_t slot) {          // blocking
        for (uint32_t spin = 0; !ly_deps_ready(p.done, g, ntiles); ++spin) {
          if (spin > (1u << 22)) __trap();
          __nanosleep(64);
        }
        fence_acq_rel_gpu();
        fence_proxy_async_all();                          // the rows were written through the generic proxy; q / k / v are read by bulk copies (async proxy)
        mbar_arrive(bar_dep + slot);
      };
      uint32_t it = 0;
      if ((int)blockIdx.x < nitems) acquire(blockIdx.x, 0);
      for (int g = blockIdx.x; g < nitems; g += gridDim.x, ++it) {
        const int gn = g + (int)gridDim.x;
        bool acq_done = gn >= nitems;
        // while the item runs: pick up the next item's predecessors as soon as they are there (normally at once)
        for (uint32_t spin = 0; !mbar_try_wait(bar_fin, it & 1); ++spin) {
          if (!acq_done && ly_deps_ready(p.done, gn, ntiles)) {
            fence_acq_rel_gpu();
            fence_proxy_async_all();
            mbar_arrive(bar_dep + ((it + 1) & 1));
            acq_done = true;
          }
          if (spin > (1u << 24)) __trap();
        }
        fence_acq_rel_gpu();                              // the CTA's stores of item g (ordered before the arrive on bar_fin) ...
        st_relaxed_gpu(p.done + g, 1);                    // ... are visible to whoever sees this flag
        if (!acq_done) acquire(gn, (it + 1) & 1);         // release first: the next item may depend on items that wait for this one
      }
    }
    return;
  }

  // =========================== compute warpgroups ===========================================================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
  const int wg = warp >> 2, lq = warp & 3, row = lq * 32 + lane;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);   // the warp index, provably warp-uniform
  const int cb = 80 * wg;                                 // this thread's column half in the row passes
  uint64_t* wb = bars + LB_WG0 + wg * WB_COUNT;
  const uint32_t trow = tmem_base + ((uint32_t)(lq * 32) << 16);
  uint32_t ph_q = 0, ph_g = 0, ph_g2 = 0, cs = 0, cp = 0, co = 0;
  auto csync = [&]() { named_bar_sync(3, LY_CTHREADS); };
  // completion of the next GEMM the issuer commits on bar_g / bar_g2 (every commit is waited for exactly once, in order)
  auto gemm_wait = [&]() {
    mbar_wait(bar_g, ph_g);
    ph_g ^= 1;
    tc_fence_after();
  };
  auto gemm_wait2 = [&]() {
    mbar_wait(bar_g2, ph_g2);
    ph_g2 ^= 1;
    tc_fence_after();
  };
  // hand-off to the GEMM-chain issuer: this warp's part of the A operand is written and fenced (fence.proxy.async), its
  // reads of the TMEM columns the GEMM overwrites are complete (tcgen05.fence::before_thread_sync); one arrival per warp
  auto go = [&]() { ly_warp_arrive(bar_go, lane); };
  // h (TMEM) of the valid rows -> HBM, chunk-major (debug stops only; the tail stores from registers)
  auto store_h = [&](const LayerArgs& a, const LyTile& tl) {
#pragma unroll 1
    for (int i = 0; i < 5; ++i) {
      float v[16];
      tmem_ld16(trow + TM_H + cb + 16 * i, v);
      if (row < tl.nq) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(d.hc + ((int64_t)(cb / 4 + 4 * i + q) * d.R + tl.row0 + row) * 4) =
              make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    }
  };

  long long pc[PROF ? 16 : 1] = {0}, fcw[PROF ? 8 : 1] = {0}, fcx[PROF ? 8 : 1] = {0};
  long long pc_last = PROF ? clock64() : 0;
#define LY_PHASE(i)                                   \
  if (PROF && a.phase_clocks && tid == 0) {           \
    const long long now_ = clock64();                 \
    pc[i] += now_ - pc_last;                          \
    pc_last = now_;                                   \
  }
  if (tid == 0 && (int)blockIdx.x < nitems) {
    const LayerArgs& a0 = p.a[blockIdx.x / ntiles];
    if (a0.mode == LM_BLOCK) mbar_arrive(bar_kvgo);       // K/V buffers are free: first window phase may load
  }

  int cur_l = -1;
  uint32_t it = 0;
  for (int g = blockIdx.x; g < nitems; g += gridDim.x, ++it) {
    const int l = g / ntiles;
    const LayerArgs& a = p.a[l];
    const LyTile tl = ly_tile(d, g - l * ntiles);
    const bool more = g + (int)gridDim.x < nitems;
    const LayerArgs& an = p.a[more ? (g + (int)gridDim.x) / ntiles : l];      // the next item's layer
    // end of an item: every compute thread has issued its stores (csync before); the agent publishes them
    auto item_fin = [&]() {
      if (p.done && tid == 0) mbar_arrive(bar_fin);
    };

    if (l != cur_l) {                                     // first item of a layer on this CTA: its constants
      cur_l = l;
      if (a.mode == LM_BLOCK)
        for (int i = tid; i < LC_COUNT; i += LY_CTHREADS) sC[i] = a.consts[i];
      csync();
      if (tid < H) {                                      // tail constants that do not depend on the tile
        sC[LS_TB + tid] = a.mode == LM_HEAD ? a.in_b[tid] : sC[LC_F3B + tid];
        if (a.tail == LT_FINAL) {
          sC[LS_TG + tid] = a.fn_w[tid];
          sC[LS_TS + tid] = a.fn_b[tid];
          if (tid < M) sC[LS_OB + tid] = a.out_b[tid];
        }
      }
    }
    if (p.done) mbar_wait(bar_dep + (it & 1), (it >> 1) & 1);   // the predecessors' h / q | k | v rows are visible
#ifndef LY_NO_L2_PREFETCH
    if (tid == 0) {                                       // contiguous row-major tiles this CTA reads later: one L2 prefetch each
      if (a.tail == LT_FINAL) {                           // the update rule's operands, read at the very end of this item
        const edtts_step_args& sp = p.step;
        const int64_t o = tl.row0 * M;
        const uint32_t nb = (uint32_t)tl.nq * M * 4;
        if (sp.mode != EDTTS_STEP_EPS) bulk_prefetch_l2(d.x_t + o, nb);
        if (sp.mode == EDTTS_STEP_DDPM && sp.noise) bulk_prefetch_l2(sp.noise + o, nb);
        if (sp.mode == EDTTS_STEP_DPM && sp.dpm_order >= 2 && sp.dpm_hist1) bulk_prefetch_l2(sp.dpm_hist1 + o, nb);
        if (sp.mode == EDTTS_STEP_DPM && sp.dpm_order >= 3 && sp.dpm_hist2) bulk_prefetch_l2(sp.dpm_hist2 + o, nb);
      }
      if (more && an.mode == LM_HEAD) {                   // the next head item's x_t tile
        const LyTile tn = ly_tile(d, g + (int)gridDim.x - ((g + (int)gridDim.x) / ntiles) * ntiles);
        bulk_prefetch_l2(d.x_t + tn.row0 * M, (uint32_t)tn.nq * M * 4);
      }
    }
#endif

    // per-utterance AdaLN vectors: gain = w * (1 + scale), shift (read after the next CTA barrier at the earliest: n3 / the tail)
    auto load_adaln = [&]() {
      if (tid < H) {
        if (a.mode == LM_BLOCK) {
          const float* m = a.mod3 + (int64_t)tl.b * d.mod_stride;
          sC[LS_G3 + tid] = sC[LC_N3W + tid] * (1.0f + m[tid]);
          sC[LS_SH3 + tid] = m[H + tid];
        }
        if (a.tail == LT_QKV) {
          const float* m = a.mod1 + (int64_t)tl.b * d.mod_stride;
          sC[LS_TG + tid] = a.n1w[tid] * (1.0f + m[tid]);
          sC[LS_TS + tid] = m[H + tid];
        }
      }
    };

    if (a.mode == LM_HEAD) {
      load_adaln();
      // ---- h = in_proj(x_t): x tile -> bf16 A operand (K = 80), one MMA chain into the h columns -----------------------
      {
        // 128 rows x 10 groups of 8 columns, consecutive threads on consecutive 32-byte pieces of the (contiguous) tile
        float4 x[10];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const int e = i * LY_CTHREADS + tid, r = e / 10, g = e - r * 10;
          const float4* src = reinterpret_cast<const float4*>(d.x_t + (tl.row0 + r) * M + 8 * g);
          x[2 * i] = r < tl.nq ? src[0] : make_float4(0.f, 0.f, 0.f, 0.f);      // plain loads: x_prev of the same launch may alias x_t
          x[2 * i + 1] = r < tl.nq ? src[1] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const int e = i * LY_CTHREADS + tid, r = e / 10, g = e - r * 10;
          const float o[8] = {x[2 * i].x, x[2 * i].y, x[2 * i].z, x[2 * i].w, x[2 * i + 1].x, x[2 * i + 1].y, x[2 * i + 1].z, x[2 * i + 1].w};
          *reinterpret_cast<uint4*>(sA + g * LY_SLAB + r * 16) = pack_bf16x8(o);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      go();                                               // issuer: h = x W_in^T
      gemm_wait();
      LY_PHASE(0)
    } else {
    // ---- tile prologue: Q -> sA, h -> TMEM ---------------------------------------------------------------------------
    if (lq == 0) LY_TR(wg, 40)
    if (warp_u == 1) {                                    // converged warp + elected lane: uniform-register addressing
      if (elect_one()) {
        mbar_expect_tx(bar_q, 20 * tl.nq * 16);
#pragma unroll
        for (int c = 0; c < 20; ++c) bulk_g2s(sA + c * LY_SLAB, a.qkv + ((int64_t)c * d.R + tl.row0) * 8, tl.nq * 16, bar_q);
      }
      __syncwarp();
    }
    {
      const float* src = d.hc + ((int64_t)(cb / 4) * d.R + tl.row0 + row) * 4;
      float4 x[20];
#pragma unroll
      for (int q = 0; q < 20; ++q)
        x[q] = row < tl.nq ? __ldcs(reinterpret_cast<const float4*>(src + (int64_t)q * d.R * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        float v[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          v[4 * q] = x[4 * i + q].x; v[4 * q + 1] = x[4 * i + q].y; v[4 * q + 2] = x[4 * i + q].z; v[4 * q + 3] = x[4 * i + q].w;
        }
        tmem_st16(trow + TM_H + cb + 16 * i, v);
      }
      if (lq == 0) LY_TR(wg, 41)
      tmem_st_wait();
    }
    if (lq == 0) LY_TR(wg, 42)
    ly_init_pads(smem, tid);
    fence_proxy_async();
    if (lq == 0) LY_TR(wg, 43)
    mbar_wait(bar_q, ph_q);
    ph_q ^= 1;
    tc_fence_before();
    if (lq == 0) LY_TR(wg, 44)
    csync();
    if (lq == 0) LY_TR(wg, 45)
    if (tid == 0) mbar_arrive(bar_attgo);                 // Q in place, attention TMEM columns and P buffers free
    load_adaln();                                         // (after the go-ahead: read at n3 / in the tail only)
    LY_PHASE(0)

    // ---- banded self-attention -------------------------------------------------------------------------------
    ly_softmax_phase<true, PROF && LY_SOFTMAX_CLOCKS>(d, tl, smem, tmem_base, wb, cs, cp, co, (PROF && a.phase_clocks && tid == 0) ? fcw : nullptr);
    go();                                                 // issuer: context K/V go-ahead, h += O Wproj^T
    LY_PHASE(1)

    // ---- h += O Wproj^T ------------------------------------------------------------------------------------------
    gemm_wait();

    // ---- n2 = RMSNorm(h + b_proj) * w2 -> sA ; h + b_proj back to TMEM ---------------------------------------------
    {
      float v[80];
      uint64_t ss2 = 0;
      tmem_ld80(trow + TM_H + cb, v);
#pragma unroll
      for (int i = 0; i < 5; ++i) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          f2_add_to(v + 16 * i + j, sC + LC_PROJ_B + cb + 16 * i + j);
          ss2 = f2_sq_acc(v + 16 * i + j, ss2);
        }
        tmem_st16(trow + TM_H + cb + 16 * i, v + 16 * i);
      }
      const float ss = f2_hsum(ss2);
      sRed[wg * 128 + row] = ss;
      tmem_st_wait();
      csync();
      const float rstd = rsqrtf((sRed[row] + sRed[128 + row]) * (1.0f / H) + 1e-6f);
#pragma unroll
      for (int g = 0; g < 10; ++g) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; j += 2)
          f2_unpack(f2_mul(f2_mul(f2_pack(v[8 * g + j], v[8 * g + j + 1]), f2_pack(rstd, rstd)),
                           f2_pack(sC[LC_N2W + cb + 8 * g + j], sC[LC_N2W + cb + 8 * g + j + 1])), o[j], o[j + 1]);
        *reinterpret_cast<uint4*>(sA + (cb / 8 + g) * LY_SLAB + row * 16) = pack_bf16x8(o);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    if (a.stop_phase == 1) {                              // debug stop (the issuer drains its prefetch and moves on)
      store_h(a, tl);
      tc_fence_before();
      csync();
      if (tid == 0 && more) mbar_arrive(bar_kvgo);
      item_fin();
      continue;
    }
    go();                                                 // issuer: q = n2 Wq^T
    LY_PHASE(2)

    // ---- q = n2 Wq^T -> bf16 Q operand in sA -------------------------------------------------------------------------
    gemm_wait();
    {
      float v[80];
      tmem_ld80(trow + TM_G + cb, v);
#pragma unroll
      for (int g = 0; g < 10; ++g) *reinterpret_cast<uint4*>(sA + (cb / 8 + g) * LY_SLAB + row * 16) = pack_bf16x8(v + 8 * g);
    }
    fence_proxy_async();
    tc_fence_before();
    csync();
    if (tid == 0) mbar_arrive(bar_attgo);
    LY_PHASE(3)

    // ---- cross attention over the context tokens ---------------------------------------------------------------------
    ly_softmax_phase<false, PROF && LY_SOFTMAX_CLOCKS>(d, tl, smem, tmem_base, wb, cs, cp, co, (PROF && a.phase_clocks && tid == 0) ? fcx : nullptr);
    go();                                                 // issuer: h += O Wout^T
    LY_PHASE(4)

    // ---- h += O Wout^T -----------------------------------------------------------------------------------------------
    gemm_wait();

    // ---- n3 = AdaRMSNorm(h) -> sA --------------------------------------------------------------------------------------
    {
      float v[80];
      uint64_t ss2 = 0;
      tmem_ld80(trow + TM_H + cb, v);
#pragma unroll
      for (int j = 0; j < 80; j += 2) ss2 = f2_sq_acc(v + j, ss2);
      sRed[wg * 128 + row] = f2_hsum(ss2);
      csync();
      const float rstd = rsqrtf((sRed[row] + sRed[128 + row]) * (1.0f / H) + 1e-6f);
#pragma unroll
      for (int g = 0; g < 10; ++g) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; j += 2)
          f2_unpack(f2_fma(f2_mul(f2_pack(v[8 * g + j], v[8 * g + j + 1]), f2_pack(rstd, rstd)),
                           f2_pack(sC[LS_G3 + cb + 8 * g + j], sC[LS_G3 + cb + 8 * g + j + 1]),
                           f2_pack(sC[LS_SH3 + cb + 8 * g + j], sC[LS_SH3 + cb + 8 * g + j + 1])), o[j], o[j + 1]);
        *reinterpret_cast<uint4*>(sA + (cb / 8 + g) * LY_SLAB + row * 16) = pack_bf16x8(o);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    if (a.stop_phase == 2) {
      store_h(a, tl);
      tc_fence_before();
      csync();
      if (tid == 0 && more) mbar_arrive(bar_kvgo);
      item_fin();
      continue;
    }
    go();                                                 // issuer: FFN quarters 0 and 1
    LY_PHASE(5)

    // ---- feed-forward in quarters of 80 u columns ------------------------------------------------------------------------
    // Quarter q = one weight chunk (x rows | gate rows) = one N = 160 GEMM into TMEM buffer q & 1 (TM_G + 160 (q & 1): x 80 | gate 80),
    // completion on bar_g (even) / bar_g2 (odd).  While the compute threads run the SwiGLU pass of quarter q the tensor core
    // runs quarter q + 1; a warp's arrival on bar_free[q & 1] tells the issuer that it has read the buffer out and written its
    // part of u.  u quarters 0..2 go to sU (30 slabs), quarter 3 to sA (n3 is dead once quarter 3's MMAs retired);
    // h += u W3^T: the k-half of quarters 0 | 1 runs under the last SwiGLU pass.
    {
      uint8_t* sU2 = sU + 20 * LY_SLAB;
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        const int b = q & 1;
        if (b == 0) gemm_wait(); else gemm_wait2();       // quarter q is in TMEM buffer b
        uint8_t* dst = q < 2 ? sU + q * 10 * LY_SLAB : q == 2 ? sU2 : sA;
        const float* bx = sC + LC_F0B + q * 160 + 40 * wg;
        const float* bg = bx + 80;
        {
          float x[40], g[40];
          tmem_ld40(trow + TM_G + 160 * b + 40 * wg, x);
          tmem_ld40(trow + TM_G + 160 * b + 80 + 40 * wg, g);
#pragma unroll
          for (int j = 0; j < 40; j += 2) swiglu2(x + j, g + j, bx + j, bg + j);
#pragma unroll
          for (int i = 0; i < 5; ++i) *reinterpret_cast<uint4*>(dst + (5 * wg + i) * LY_SLAB + row * 16) = pack_bf16x8(x + 8 * i);
        }
        fence_proxy_async();
        tc_fence_before();
        ly_warp_arrive(bar_free + b, lane);
      }
    }
    LY_PHASE(6)

    // ---- h += u W3^T (issued by the issuer behind quarter 3) ------------------------------------------------------------------
    gemm_wait();
    if (tid == 0 && more) mbar_arrive(bar_kvgo);          // overlay K/V region is free: next tile's window K/V may load
    }   // LM_BLOCK
    LY_PHASE(7)

    // =============== tail: the finished h rows leave the SM; what consumes them next runs right here ===============
    if (lq == 0) LY_TR(wg, 50)
    {
      float v[80];
      float s1 = 0.f;
      tmem_ld80(trow + TM_H + cb, v);
      if (lq == 0) LY_TR(wg, 51)
#pragma unroll
      for (int j = 0; j < 80; j += 2) f2_add_to(v + j, sC + LS_TB + cb + j);
      if (a.mode == LM_HEAD && row < tl.nq) {             // + pos_emb.pe[t]
        if (d.pe_cm) {                                    // chunk-major table: consecutive rows are 16 bytes apart
          const float* pp = d.pe_cm + ((int64_t)(cb / 4) * d.pe_rows + tl.t0 + row) * 4;
#pragma unroll
          for (int q = 0; q < 20; ++q) {
            const float4 pv = __ldg(reinterpret_cast<const float4*>(pp + (int64_t)q * d.pe_rows * 4));
            f2_add_to(v + 4 * q, &pv.x);
            f2_add_to(v + 4 * q + 2, &pv.z);
          }
        } else {
          const float4* pp = reinterpret_cast<const float4*>(d.pe + (int64_t)(tl.t0 + row) * H + cb);
#pragma unroll
          for (int q = 0; q < 20; ++q) {
            const float4 pv = pp[q];
            f2_add_to(v + 4 * q, &pv.x);
            f2_add_to(v + 4 * q + 2, &pv.z);
          }
        }
      }
      // h -> HBM: 20 x 16 bytes per thread (the store path makes this ~3 k cycles per tile).  With a QKV tail the stores are
      // issued AFTER the hand-off to the issuer below, i.e. under the q | k GEMMs
      auto store_rows = [&]() {
        if (row < tl.nq) {
#pragma unroll
          for (int q = 0; q < 20; ++q)
            __stcs(reinterpret_cast<float4*>(d.hc + ((int64_t)(cb / 4 + q) * d.R + tl.row0 + row) * 4),
                   make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));    // streaming: read next by another SM's launch
        }
      };
      if (a.tail == LT_NONE) store_rows();
      if (lq == 0) LY_TR(wg, 52)
      if (a.tail != LT_NONE) {
        // next norm: AdaRMSNorm (LT_QKV) or LayerNorm (LT_FINAL, exact two-step variance) -> bf16 A operand
        if (a.tail == LT_FINAL) {
#pragma unroll
          for (int j = 0; j < 80; ++j) s1 += v[j];
        } else {
          uint64_t ss2 = 0;
#pragma unroll
          for (int j = 0; j < 80; j += 2) ss2 = f2_sq_acc(v + j, ss2);
          s1 = f2_hsum(ss2);
        }
        sRed[wg * 128 + row] = s1;
        if (lq == 0) LY_TR(wg, 53)
        csync();
        if (lq == 0) LY_TR(wg, 54)
        const float tot = sRed[row] + sRed[128 + row];
        float mean = 0.f, rstd;
        if (a.tail == LT_FINAL) {
          mean = tot * (1.0f / H);
          float s2 = 0.f;
#pragma unroll
          for (int j = 0; j < 80; ++j) s2 = fmaf(v[j] - mean, v[j] - mean, s2);
          sRed[256 + wg * 128 + row] = s2;
          csync();
          rstd = rsqrtf((sRed[256 + row] + sRed[384 + row]) * (1.0f / H) + 1e-5f);
        } else {
          rstd = rsqrtf(tot * (1.0f / H) + 1e-6f);
        }
#pragma unroll
        for (int g = 0; g < 10; ++g) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; j += 2)
            f2_unpack(f2_fma(f2_mul(f2_add(f2_pack(v[8 * g + j], v[8 * g + j + 1]), f2_pack(-mean, -mean)), f2_pack(rstd, rstd)),
                             f2_pack(sC[LS_TG + cb + 8 * g + j], sC[LS_TG + cb + 8 * g + j + 1]),
                             f2_pack(sC[LS_TS + cb + 8 * g + j], sC[LS_TS + cb + 8 * g + j + 1])), o[j], o[j + 1]);
          *reinterpret_cast<uint4*>(sA + (cb / 8 + g) * LY_SLAB + row * 16) = pack_bf16x8(o);
        }
        if (lq == 0) LY_TR(wg, 55)
        fence_proxy_async();
        tc_fence_before();
        go();                                             // issuer: the tail's GEMMs (h has been read out, the A operand is in sA)
        if (a.tail == LT_QKV) store_rows();
      }
    }
    tc_fence_before();
    if (lq == 0) LY_TR(wg, 57)
    LY_PHASE(8)

    if (a.tail == LT_QKV) {
      // ---- q | k | v of the next block: three 160-column chunks into TMEM columns 0..479 (h is dead), completions on
      // bar_g, bar_g2, bar_g: each part is converted and stored while the next one is still in the tensor pipe ----
#pragma unroll 1
      for (int part = 0; part < 3; ++part) {              // q, k as bf16; v as f16 (P V operand)
        if (part == 1) gemm_wait2(); else gemm_wait();
        if (part == 0) {
          ly_warp_arrive(bar_free, lane);                 // q's completion seen: the issuer may commit v on bar_g
          LY_PHASE(9)
        }
        if (part == 2) {
          LY_PHASE(12)
        }
        {
          float v[80];
          tmem_ld80(trow + 160 * part + cb, v);
          if (row < tl.nq) {
            __nv_bfloat16* o = a.qkv_out + ((int64_t)(20 * part + cb / 8) * d.R + tl.row0 + row) * 8;
#pragma unroll
            for (int g = 0; g < 10; ++g)
              *reinterpret_cast<uint4*>(o + (int64_t)g * d.R * 8) = part == 2 ? pack_f16x8(v + 8 * g) : pack_bf16x8(v + 8 * g);
          }
        }
      }
    } else if (a.tail == LT_FINAL) {
      // ---- eps = out_proj(final_norm(h)) with the update rule applied in registers -----------------------------------
      gemm_wait();
      const edtts_step_args& sa = p.step;
      // Per-utterance coefficients: the same for every row of the tile (all threads load them, the loads broadcast).
      float ab_t = 0.f, ab_p = 1.f, al = 0.f, be = 0.f, pv = 0.f, nzm = 0.f;
      if (sa.mode == EDTTS_STEP_DDIM || sa.mode == EDTTS_STEP_DDPM) {
        const int64_t tt = sa.t[tl.b];
        ab_t = sa.alpha_bar[tt];
        if (sa.mode == EDTTS_STEP_DDIM) {
          const int64_t tp = sa.t_prev[tl.b];
          ab_p = tp >= 0 ? sa.alpha_bar[tp] : 1.0f;
        } else {
          al = sa.alphas[tt];
          be = sa.betas[tt];
          pv = sa.posterior_var[tt];
          nzm = tt > 0 ? 1.0f : 0.0f;
        }
      }
      // eps (+ bias) of this thread's row goes through shared memory (sA: the out_proj A operand is dead once the MMA
      // retired), so that the update below walks x_t / x_prev / x0 -- row-major [R][80] fp32 -- with consecutive threads on
      // consecutive 16-byte pieces: 4 LSU wavefronts per warp request instead of the 32 of a thread-per-row walk (each
      // lane in its own 320-byte row), which made this phase the slowest per byte of the whole item.
      // The tile takes slabs 0..19 of sA exactly (128 x 80 x 4 B): slab 20 is head 3's zero padding of the S = Q K^T
      // operand and must stay zero.  The 16-byte pieces of a row are rotated by row / 2 so that the thread-per-row
      // stores of a warp spread over all banks (row stride 80 words = 16 mod 32).
      constexpr int FE_LD = 80;
      static_assert(128 * FE_LD * 4 <= 20 * LY_SLAB, "staged eps tile must not reach the padding slab of sA");
      float* sE = reinterpret_cast<float*>(sA);
      {
        const int c0 = 40 * wg;                           // this thread's 40 of the 80 mel columns
        float e[40];
        tmem_ld32(trow + c0, e);
        tmem_ld8(trow + c0 + 32, e + 32);
#pragma unroll
        for (int q = 0; q < 10; ++q)
          *reinterpret_cast<float4*>(sE + row * FE_LD + 4 * ((10 * wg + q + (row >> 1)) % 20)) =
              make_float4(e[4 * q] + sC[LS_OB + c0 + 4 * q], e[4 * q + 1] + sC[LS_OB + c0 + 4 * q + 1],
                          e[4 * q + 2] + sC[LS_OB + c0 + 4 * q + 2], e[4 * q + 3] + sC[LS_OB + c0 + 4 * q + 3]);
      }
      csync();
      const float* dc = sa.mode == EDTTS_STEP_DPM ? sa.dpm_coef + tl.b * 8 : nullptr;
      const int ord = sa.dpm_order, pm = sa.dpm_predict_x0 ? 1 : 0;
#pragma unroll 2
      for (int i = 0; i < 10; ++i) {                      // 128 rows x 20 float4 = 10 per thread
        const int idx = i * LY_CTHREADS + tid, r = idx / 20, q = idx - r * 20;
        if (r >= tl.nq) break;                            // idx grows with i: no later piece of this thread is valid either
        const float4 e4 = *reinterpret_cast<const float4*>(sE + r * FE_LD + 4 * ((q + (r >> 1)) % 20));
        const int64_t o = (tl.row0 + r) * M + 4 * q;
        if (sa.eps_out) *reinterpret_cast<float4*>(sa.eps_out + o) = e4;
        if (sa.mode == EDTTS_STEP_DDIM) {
          const float4 x = *reinterpret_cast<const float4*>(d.x_t + o);
          float4 xp, x0;
          ddim_update(x.x, e4.x, 0.f, ab_t, ab_p, 0.f, xp.x, x0.x);
          ddim_update(x.y, e4.y, 0.f, ab_t, ab_p, 0.f, xp.y, x0.y);
          ddim_update(x.z, e4.z, 0.f, ab_t, ab_p, 0.f, xp.z, x0.z);
          ddim_update(x.w, e4.w, 0.f, ab_t, ab_p, 0.f, xp.w, x0.w);
          if (sa.x0_out) __stcs(reinterpret_cast<float4*>(sa.x0_out + o), x0);      // streaming: read back by the host / next history step only
          if (sa.write_x_prev && sa.x_prev_out) *reinterpret_cast<float4*>(sa.x_prev_out + o) = xp;
        } else if (sa.mode == EDTTS_STEP_DDPM) {
          const float4 x = *reinterpret_cast<const float4*>(d.x_t + o);
          const float4 nz = *reinterpret_cast<const float4*>(sa.noise + o);
          float4 xp;
          xp.x = ddpm_update(x.x, e4.x, nz.x, al, ab_t, be, pv, nzm);
          xp.y = ddpm_update(x.y, e4.y, nz.y, al, ab_t, be, pv, nzm);
          xp.z = ddpm_update(x.z, e4.z, nz.z, al, ab_t, be, pv, nzm);
          xp.w = ddpm_update(x.w, e4.w, nz.w, al, ab_t, be, pv, nzm);
          *reinterpret_cast<float4*>(sa.x_prev_out + o) = xp;
        } else if (sa.mode == EDTTS_STEP_DPM) {
          const float4 x = *reinterpret_cast<const float4*>(d.x_t + o);
          const float4 ha = ord >= 2 ? *reinterpret_cast<const float4*>(sa.dpm_hist1 + o) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 hb = ord >= 3 ? *reinterpret_cast<const float4*>(sa.dpm_hist2 + o) : make_float4(0.f, 0.f, 0.f, 0.f);
          float4 xp, x0;
          dpm_update(x.x, e4.x, ha.x, hb.x, dc, ord, pm, xp.x, x0.x);
          dpm_update(x.y, e4.y, ha.y, hb.y, dc, ord, pm, xp.y, x0.y);
          dpm_update(x.z, e4.z, ha.z, hb.z, dc, ord, pm, xp.z, x0.z);
          dpm_update(x.w, e4.w, ha.w, hb.w, dc, ord, pm, xp.w, x0.w);
          if (sa.x0_out) *reinterpret_cast<float4*>(sa.x0_out + o) = x0;
          *reinterpret_cast<float4*>(sa.x_prev_out + o) = xp;
        }
      }
      fence_proxy_async();                                // sA is written next by the following item's bulk copies
    }
    // a head item is followed by a block item: its K/V buffers are free
    if (tid == 0 && a.mode == LM_HEAD && more && an.mode == LM_BLOCK) mbar_arrive(bar_kvgo);
    tc_fence_before();
    csync();
    item_fin();
    LY_PHASE(13)
  }
  if (PROF && p.a[0].phase_clocks && tid == 0)
    for (int i = 0; i < 8; ++i) {
      p.a[0].phase_clocks[blockIdx.x * 32 + i] = pc[i];
      p.a[0].phase_clocks[blockIdx.x * 32 + 8 + i] = fcw[i];
      p.a[0].phase_clocks[blockIdx.x * 32 + 16 + i] = fcx[i];
      p.a[0].phase_clocks[blockIdx.x * 32 + 24 + i] = pc[8 + i];
    }

  csync();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

}  // namespace tc
}  // namespace edtts
