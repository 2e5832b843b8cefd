// EDTTS_PREC_TF32X3: the fp32-grade (max-abs 1e-4) decoder step on the tensor cores.
//
// The CUDA-core parity path (decoder.cu: gemm_simt_kernel + attn_simt_kernel) spends 115 ms in its GEMMs and 55 ms in its
// attentions per cfg3 generate.  Here the same sequence of launches (models/decoder.py:96-109, layers/transformer.py:141-160)
// runs on tcgen05 with every operand split into a tf32 head and tail (tf32x3.cuh: three kind::tf32 MMAs per k-step, fp32
// accumulator in tensor memory, ~2^-22 relative per operand: measured 4 - 5 x the rounding error of the FFMA chain, 5.8e-6 on eps):
//
//   t3_gemm_kernel   y = epi(pro(A) W^T + b) for 128 rows x one NB-column block of W, 288 threads, two CTAs per SM.  Warps 0-7
//                    fetch the A values of a 16-element chunk (three chunks ahead), apply the RMSNorm / AdaRMSNorm / LayerNorm
//                    prologue of gemm_simt.cuh (row statistics: t3_rowstats_kernel, a small launch in front), split hi / lo into
//                    one of two operand-image buffers and hand it over by per-warp mbarrier arrivals; warp 8 streams the weight
//                    chunks (bulk copies into three buffers, from images packed per step by ONE launch of t3_pack_jobs_kernel)
//                    and issues the MMAs.  Epilogues: bias, GELU, SwiGLU (x 80 | gate 80 per block) -> a shared-memory tile ->
//                    the TMA engine (bulk copy per row; the residual GEMMs as bulk ADD in L2), or coalesced walks for the
//                    positional table and the fused DDIM / DDPM / DPM update rule of the last GEMM.
//   t3_attn_kernel   one CTA = (utterance, head, 128 queries), 160 threads (thread <-> query <-> tensor-memory lane; warp 4 issues
//                    the MMAs), keys in blocks of 32: S = Q K^T (15 MMAs, one block ahead in a second tensor-memory block) ->
//                    online softmax in registers (exp2, fp32) -> P split hi / lo into a shared-memory operand image -> O += P V
//                    (12 MMAs, V staged transposed) accumulating in tensor memory with a lazily raised running maximum.  Band
//                    (|i - j| <= 64, attention.py:94-111) or full context (mla.py:176-180) by a per-element mask with
//                    warp-uniform fast paths; only the key blocks a tile can see are visited.  Two CTAs per SM.  Cross-attention:
//                    K | V blocks arrive as ready-made operand images (t3_kvimg_kernel, built once per generate by
//                    edtts_context_prepare) by bulk copies of the issuer lane -- the softmax warps issue no global load.
//
// A row's result does not depend on the other rows of a launch (batch invariance).
#define EDTTS_DECL_ONLY
#include "t3_decoder.cuh"
#include "tf32x3.cuh"
#include <algorithm>

namespace edtts {
namespace t3 {

constexpr int GT = 256;                               // compute threads of the GEMM kernel
constexpr int GT_ALL = GT + 32;                       // + the MMA issuer warp

// mbarrier wait of an MMA-issuer lane: a short sleep between probes, so that the issuer's spin does not take issue slots from the
// compute warp sharing its scheduler (12.8 % of the attention kernel's issued instructions were this loop)
__device__ __forceinline__ void mbar_wait_issuer(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
    __nanosleep(32);
    if (spin > (1u << 22)) __trap();
  }
}

// ---- weight images ------------------------------------------------------------------------------------------------------
struct PackJob {
  const float* W;        // first source row
  float* img;            // image of the block this job writes into
  int k, ldw, nrows, row_off, ntot, nchunk;            // rows [row_off, row_off + nrows) of an image with ntot rows
};
constexpr int MAX_JOBS = 64;
struct PackJobs {
  PackJob j[MAX_JOBS];
};

// image of a block, per 32-element chunk c: [hi | lo][slab s < 8][row < ntot][4] = W[row][32 c + 4 s + j] (tf32x3.cuh)
__global__ void t3_pack_jobs_kernel(const __grid_constant__ PackJobs jobs) {
  const PackJob& J = jobs.j[blockIdx.y];
  const int total = J.nchunk * 8 * J.nrows * 4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i & 3, row = (i >> 2) % J.nrows, s = ((i >> 2) / J.nrows) & 7, c = (i >> 2) / (J.nrows * 8);
    const int kk = KC * c + 4 * s + j;
    const float w = kk < J.k ? J.W[(int64_t)row * J.ldw + kk] : 0.f;
    const float hi = tf32_rna(w);
    const int64_t o = (int64_t)c * (64 * J.ntot) + ((int64_t)s * J.ntot + J.row_off + row) * 4 + j;
    J.img[o] = hi;
    J.img[o + 8 * J.ntot * 4] = tf32_rna(w - hi);
  }
}

int64_t t3_gemm_block_stride(int K, int NB) { return (int64_t)((K + KC - 1) / KC) * 64 * NB; }
int64_t t3_gemm_image_floats(int K, int N, int NB, bool swiglu) {
  const int nblocks = swiglu ? N / (NB / 2) : N / NB;
  return nblocks * t3_gemm_block_stride(K, NB);
}

// appends the jobs of one matrix; returns the number of jobs added (or -1)
static int add_jobs(PackJobs& pj, int n, const float* W, float* img, int K, int N, int NB, bool swiglu) {
  const int nchunk = (K + KC - 1) / KC;
  const int64_t stride = t3_gemm_block_stride(K, NB);
  if (swiglu) {
    const int half = NB / 2, nblocks = N / half;
    if (n + 2 * nblocks > MAX_JOBS) return -1;
    for (int b = 0; b < nblocks; ++b) {
      pj.j[n++] = PackJob{W + (int64_t)b * half * K, img + b * stride, K, K, half, 0, NB, nchunk};
      pj.j[n++] = PackJob{W + (int64_t)(N + b * half) * K, img + b * stride, K, K, half, half, NB, nchunk};
    }
    return 2 * nblocks;
  }
  const int nblocks = N / NB;
  if (n + nblocks > MAX_JOBS) return -1;
  for (int b = 0; b < nblocks; ++b) pj.j[n++] = PackJob{W + (int64_t)b * NB * K, img + b * stride, K, K, NB, 0, NB, nchunk};
  return nblocks;
}
static int launch_pack(const PackJobs& pj, int n, cudaStream_t st) {
  LaunchScope ls(KC_TC_MISC, st);
  t3_pack_jobs_kernel<<<dim3(40, n), 256, 0, st>>>(pj);
  return check_launch("t3_pack_jobs");
}
int pack_w_blocks(const float* W, float* img, int K, int N, int NB, bool swiglu, cudaStream_t st) {
  PackJobs pj;
  const int n = add_jobs(pj, 0, W, img, K, N, NB, swiglu);
  EDTTS_REQUIRE(n > 0, EDTTS_EINVAL, "t3 pack: K=%d N=%d NB=%d", K, N, NB);
  return launch_pack(pj, n, st);
}

// ---- GEMM -----------------------------------------------------------------------------------------------------------------
// One CTA = 128 rows x one NB-column block, 256 threads, two CTAs per SM.  The contraction runs in chunks of 16 elements (two
// k-steps, six MMAs): the A values of a chunk are read straight from global memory into registers two chunks ahead (a warp
// instruction covers 8 rows x 64 contiguous bytes), get the normalisation prologue, are split hi / lo and stored into one of TWO
// operand-image buffers; the weight chunk arrives by two bulk copies (hi, lo) into one of THREE buffers, requested a whole
// iteration before its MMAs; so chunk c + 1 is prepared while the MMAs of chunk c run.  Epilogue: accumulator -> registers
// (+ bias, SwiGLU) -> a shared-memory tile over the operand buffers -> coalesced 16-byte global accesses for the residual / PE
// / update-rule operands and the output (row-strided per-thread accesses cost a 32-line transaction per warp instruction).
struct T3GemmArgs {
  GemmArgs g;
  const float* wimg;
  int64_t img_stride;
  const float* stats;                                 // normalisation prologue: (rstd, mean) per row, written by t3_rowstats_kernel
  int NB, nchunk, main_bytes;
  int ahead;                                          // CTAs resident on the device at a time: the tile `ahead` positions on is prefetched into L2
};
// Row statistics of the normalisation prologues, once per row (every NB-column block of a GEMM needs them, and inside the GEMM
// they cost a pass over the tile before the first MMA can be fed: 14 k of a CTA's ~45 k cycles).  A warp per row, four rows in
// flight per warp; stats[row] = (rstd, mean).
__global__ void __launch_bounds__(256) t3_rowstats_kernel(const float* __restrict__ A, int64_t rows, int K, int lda, int pro, float eps,
                                                          float2* __restrict__ stats) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int RB = 4;
  for (int64_t r0 = ((int64_t)blockIdx.x * 8 + warp) * RB; r0 < rows; r0 += (int64_t)gridDim.x * 8 * RB) {
    float v[RB][6];
#pragma unroll
    for (int u = 0; u < RB; ++u)
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int k = lane + 32 * i;
        v[u][i] = (r0 + u < rows && k < K) ? __ldg(A + (r0 + u) * lda + k) : 0.f;
      }
#pragma unroll
    for (int u = 0; u < RB; ++u) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 6; ++i) s += (pro == PRO_LN) ? v[u][i] : v[u][i] * v[u][i];
      s = warp_sum(s);
      float mean = 0.f, var;
      if (pro == PRO_LN) {
        mean = s / (float)K;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const int k = lane + 32 * i;
          const float d = (k < K) ? v[u][i] - mean : 0.f;
          q += d * d;
        }
        var = warp_sum(q) / (float)K;
      } else {
        var = s / (float)K;
      }
      if (lane == 0 && r0 + u < rows) stats[r0 + u] = make_float2(1.0f / sqrtf(var + eps), mean);
    }
  }
}

constexpr int GK = 16;                                // contraction elements per pipeline chunk
constexpr int G_A_BUF = 2 * (GK / 4) * TM * 16;       // hi | lo operand image of one chunk: 16,384 B
constexpr int G_NW = 3;                               // weight chunk buffers

__global__ void __launch_bounds__(GT_ALL, 2) t3_gemm_kernel(const __grid_constant__ T3GemmArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const GemmArgs& g = a.g;
  const int NB = a.NB;
  const int w_half = (GK / 4) * NB * 16;              // hi (or lo) part of a weight chunk
  uint8_t* sA = smem;
  uint8_t* sW = smem + 2 * G_A_BUF;
  float* s_rstd = reinterpret_cast<float*>(smem + a.main_bytes);
  float* s_mean = s_rstd + TM;
  float* s_bias = s_mean + TM;                                    // [NB <= 160] bias of the block's image rows
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(s_bias + 160);    // [G_NW] weight chunk landed (bulk-copy bytes)
  uint64_t* bar_mma = bar_w + G_NW;                               // [2] MMAs of chunk c retired (tcgen05.commit)
  uint64_t* bar_a = bar_mma + 2;                                  // [2] operand image of chunk c written (one arrival per compute warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_a + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t row0 = (int64_t)blockIdx.x * TM;
  const float* wimg = a.wimg + (int64_t)blockIdx.y * a.img_stride;
  const int K = g.K, nchunk = a.nchunk;
#ifdef T3_CLOCKS   // development: cycle counts of one CTA's phases (thread 0), printed
  long long ck0 = clock64(), ck_m = 0, ck_split = 0, ck1 = 0, ck2 = 0, ck3 = 0, ck4 = 0;
#define T3_CK(var, stmt) { const long long t_ = clock64(); stmt; var += clock64() - t_; }
#else
#define T3_CK(var, stmt) { stmt; }
#endif

  if (tid == 0) {
    for (int i = 0; i < G_NW; ++i) mbar_init(bar_w + i, 1);
    mbar_init(bar_mma, 1);
    mbar_init(bar_mma + 1, 1);
    mbar_init(bar_a, GT / 32);
    mbar_init(bar_a + 1, GT / 32);
    mbar_fence_init();
  }
  if (g.pro != PRO_NONE && tid < TM) {                 // row statistics of the tile (t3_rowstats_kernel)
    const float2 st = row0 + tid < g.rows ? __ldg(reinterpret_cast<const float2*>(a.stats) + row0 + tid) : make_float2(0.f, 0.f);
    s_rstd[tid] = st.x;
    s_mean[tid] = st.y;
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  if (tid < NB) {                                      // image row j of the block <-> bias index (SwiGLU: x rows, then their gates)
    const int nout = g.epi == EPI_SWIGLU ? NB / 2 : NB;
    const int j = tid < nout ? blockIdx.y * nout + tid : g.N + blockIdx.y * nout + (tid - nout);
    s_bias[tid] = g.bias ? __ldg(g.bias + j) : 0.f;
  }
  tc_fence_before();
  __syncthreads();                                     // barriers initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // ================= MMA issuer / weight streamer: warp 8, one elected lane =================
  // No compute thread ever sits in the tensor core's queue: the issuer waits for "operand image c written" (bar_a) and "weight
  // chunk c landed" (bar_w), issues the six MMAs of the chunk and commits them to bar_mma; a weight buffer is refilled as soon as
  // the chunk that used it has retired (three buffers: chunk c + 2 is requested while chunks c and c + 1 are queued).
  if (warp == GT / 32) {
    if (lane == 0) {
      auto request_w = [&](int c) {                    // chunk c (16 elements = half of a 32-element image chunk): hi slabs, lo slabs
        uint64_t* bar = bar_w + c % G_NW;
        uint8_t* dst = sW + (c % G_NW) * 2 * w_half;
        const float* src = wimg + (int64_t)(c >> 1) * (64 * NB) + (c & 1) * (16 * NB);
        mbar_expect_tx(bar, 2 * w_half);
        bulk_g2s(dst, src, w_half, bar);
        bulk_g2s(dst + w_half, src + 32 * NB, w_half, bar);
      };
      for (int w = 0; w < 2 && g.lda == K; ++w) {      // L2 prefetch (fire and forget) of this tile's rows and of the tile a later
        const int64_t r0 = row0 + (int64_t)w * a.ahead * TM;   // CTA of this SM will work on
        if (r0 < g.rows) {
          const int64_t nbytes = (min((int64_t)TM, g.rows - r0)) * K * 4;
          const char* src = reinterpret_cast<const char*>(g.A + r0 * g.lda);
          for (int64_t o = 0; o < nbytes; o += 16384) bulk_prefetch_l2(src + o, (uint32_t)min((int64_t)16384, nbytes - o));
        }
      }
      for (int c = 0; c < G_NW && c < nchunk; ++c) request_w(c);
      for (int c = 0; c < nchunk; ++c) {
        mbar_wait_issuer(bar_a + (c & 1), (c >> 1) & 1);
        mbar_wait_issuer(bar_w + c % G_NW, (c / G_NW) & 1);
        tc_fence_after();
        const uint32_t ah = smem_u32(sA + (c & 1) * G_A_BUF), wh = smem_u32(sW + (c % G_NW) * 2 * w_half);
        issue_chunk(tmem, ah, ah + G_A_BUF / 2, wh, wh + w_half, GK / 8, NB, c > 0);
        umma_commit(bar_mma + (c & 1));
        if (c >= 1 && c + 2 < nchunk) {                // buffer (c + 2) % 3 was chunk c - 1's
          mbar_wait_issuer(bar_mma + ((c - 1) & 1), ((c - 1) >> 1) & 1);
          request_w(c + 2);
        }
      }
    }
    return;
  }

  // ================= compute warps 0-7 =================
  // this thread's two (row, 16-byte piece) slots of every chunk
  const int ap = lane >> 3;
  int ar[2];
  const float* aptr[2];
  const float* am[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    ar[s] = (warp * 2 + s) * 8 + (lane & 7);
    const int64_t row = row0 + ar[s];
    aptr[s] = row < g.rows ? g.A + row * g.lda + 4 * ap : nullptr;
    am[s] = (g.pro == PRO_ADARMS && row < g.rows) ? g.mod + (row / g.rows_per_batch) * (int64_t)g.mod_stride : nullptr;
  }
  auto fetch = [&](int c, float4 (&dst)[2]) {
#pragma unroll
    for (int s = 0; s < 2; ++s)
      dst[s] = (aptr[s] && c < nchunk) ? __ldg(reinterpret_cast<const float4*>(aptr[s] + GK * c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  float4 cur[2], nx1[2], nx2[2], nx3[2];

  fetch(0, cur);
  fetch(1, nx1);
  fetch(2, nx2);
#ifdef T3_CLOCKS
  ck1 = clock64();
#endif

  for (int c = 0; c < nchunk; ++c) {
    fetch(c + 3, nx3);
    // the chunk's prologue vectors (16-byte loads, issued before the wait below): norm weight / bias of columns 16 c + 4 ap ..,
    // AdaLN scale and shift of this thread's two rows
    float4 w4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f), sc4[2], sh4[2];
    if (g.pro != PRO_NONE) {
      const int k0 = GK * c + 4 * ap;
      w4 = __ldg(reinterpret_cast<const float4*>(g.norm_w + k0));
      if (g.pro == PRO_LN) b4 = __ldg(reinterpret_cast<const float4*>(g.norm_b + k0));
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        sc4[s] = make_float4(0.f, 0.f, 0.f, 0.f);
        sh4[s] = sc4[s];
        if (am[s]) {
          sc4[s] = __ldg(reinterpret_cast<const float4*>(am[s] + k0));
          sh4[s] = __ldg(reinterpret_cast<const float4*>(am[s] + K + k0));
        }
      }
    }
    if (c >= 2) {                                      // operand buffer c & 1 is free once chunk c - 2 retired
      T3_CK(ck_m, mbar_wait(bar_mma + (c & 1), ((c >> 1) - 1) & 1))
    }
    uint8_t* sAh = sA + (c & 1) * G_A_BUF;
    uint8_t* sAl = sAh + G_A_BUF / 2;
#ifdef T3_CLOCKS
    const long long t_split = clock64();
#endif
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      float v[4] = {cur[s].x, cur[s].y, cur[s].z, cur[s].w};
      if (g.pro != PRO_NONE && aptr[s]) {
        const float rstd = s_rstd[ar[s]], mean = s_mean[ar[s]];
        const float wv[4] = {w4.x, w4.y, w4.z, w4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
        const float sc[4] = {sc4[s].x, sc4[s].y, sc4[s].z, sc4[s].w}, sh[4] = {sh4[s].x, sh4[s].y, sh4[s].z, sh4[s].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float x = v[j];
          if (g.pro == PRO_LN) {
            x = (x - mean) * rstd * wv[j] + bv[j];
          } else {
            x = (x * rstd) * wv[j];
            if (am[s]) x = x * (1.0f + sc[j]) + sh[j];
          }
          v[j] = x;
        }
      }
      split_store(sAh, sAl, ap, ar[s], v);
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_a + (c & 1));
#ifdef T3_CLOCKS
    ck_split += clock64() - t_split;
#endif
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      cur[s] = nx1[s];
      nx1[s] = nx2[s];
      nx2[s] = nx3[s];
    }
  }
#ifdef T3_CLOCKS
  ck2 = clock64();
#endif
  mbar_wait(bar_mma + ((nchunk - 1) & 1), ((nchunk - 1) >> 1) & 1);   // the last commit covers every earlier MMA
  tc_fence_after();
#ifdef T3_CLOCKS
  ck3 = clock64();
#endif

  // ---- epilogue, phase 1: thread = (row r, half of the block's output columns) -> stage[r][col] (+ bias, SwiGLU) -----------
  const bool swi = g.epi == EPI_SWIGLU;
  const int NOUT = swi ? NB / 2 : NB;                  // output columns of this block
  const int n_out0 = blockIdx.y * NOUT;
  const int LD = NOUT + 4;                             // stage row stride (words): 4 mod 32, conflict-free 16-byte accesses
  float* stage = reinterpret_cast<float*>(smem);       // over the operand / weight buffers: every MMA and bulk copy has retired
  {
    const int lq = warp & 3, half = warp >> 2;
    const int r = lq * 32 + lane;
    const int ncol = NOUT / 2;                         // a multiple of 8
    const uint32_t trow = tmem + ((uint32_t)(lq * 32) << 16);
    for (int c8 = 0; c8 < ncol; c8 += 8) {
      const int cc = half * ncol + c8;                 // column inside the block
      float v[8], gt[8];
      tmem_ld8(trow + cc, v);
      if (swi) tmem_ld8(trow + NOUT + cc, gt);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float x = v[j] + s_bias[cc + j];
        if (swi) {                                      // x * silu(gate) with ex2.approx and the approximate reciprocal (2 ulp each)
          const float gate = gt[j] + s_bias[NOUT + cc + j];
          x = x * __fdividef(gate, 1.0f + __expf(-gate));
        }
        if (g.epi == EPI_GELU) x = gelu_erf(x);
        v[j] = x;
      }
      *reinterpret_cast<float4*>(stage + r * LD + cc) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(stage + r * LD + cc + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
  fence_proxy_async();                                 // the staged tile is read by bulk copies below
  tc_fence_before();
  named_bar_sync(1, GT);
  if (warp == 0) tmem_dealloc<256>(tmem);
#ifdef T3_CLOCKS
  ck4 = clock64();
#endif

  // ---- phase 2a: plain / residual outputs leave by the TMA engine, one bulk copy (or bulk add: global += shared, performed in
  // L2 -- the residual never travels to the SM) per row --------------------------------------------------------------------------
  if (g.epi == EPI_STORE || g.epi == EPI_GELU || g.epi == EPI_SWIGLU || g.epi == EPI_RESID) {
    const int r = warp * (TM / (GT / 32)) + lane;     // 16 rows per warp: the bulk instructions of a warp's lanes are serialised
    if (lane < TM / (GT / 32) && row0 + r < g.rows) {
      float* dst = g.out + (row0 + r) * g.ldo + n_out0;
      if (g.epi == EPI_RESID) bulk_reduce_add_f32(dst, stage + r * LD, NOUT * 4);
      else bulk_s2g(dst, stage + r * LD, NOUT * 4);
      bulk_commit();
      bulk_wait_all();                                 // shared memory must outlive the engine's reads; writes performed
    }
#ifdef T3_CLOCKS
    if (tid == 0 && blockIdx.x == 640 && blockIdx.y == 0)
      printf("t3clk K=%d NB=%d pro=%d epi=%d: setup %lld loop %lld (mma-wait %lld split+arrive %lld) final %lld phase1 %lld bulk %lld\n", K, NB,
             g.pro, g.epi, ck1 - ck0, ck2 - ck1, ck_m, ck_split, ck3 - ck2, ck4 - ck3, clock64() - ck4);
#endif
    return;
  }

  // ---- phase 2b (positional table, update rule): consecutive threads walk consecutive 16-byte pieces of a row --------------------
  const int n4 = NOUT / 4;
  for (int i = tid; i < TM * n4; i += GT) {
    const int r = i / n4, c4 = i - r * n4;
    const int64_t row = row0 + r;
    if (row >= g.rows) break;                          // rows ascend with i
    const float4 s4 = *reinterpret_cast<const float4*>(stage + r * LD + 4 * c4);
    float v[4] = {s4.x, s4.y, s4.z, s4.w};
    const int col0 = n_out0 + 4 * c4;
    const int64_t o0 = row * g.ldo + col0;
    if (g.epi == EPI_STEP) {
      const int64_t bidx = row / g.rows_per_batch;
      float ab_t = 0.f, ab_p = 1.f, al = 0.f, be = 0.f, pv = 0.f, nzm = 0.f;
      if (g.step.mode == EDTTS_STEP_DDIM || g.step.mode == EDTTS_STEP_DDPM) {
        const int64_t t = g.step.t[bidx];
        ab_t = g.step.alpha_bar[t];
        if (g.step.mode == EDTTS_STEP_DDIM) {
          const int64_t tp = g.step.t_prev[bidx];
          ab_p = (tp >= 0) ? g.step.alpha_bar[tp] : 1.0f;
        } else {
          al = g.step.alphas[t];
          be = g.step.betas[t];
          pv = g.step.posterior_var[t];
          nzm = (t > 0) ? 1.0f : 0.0f;
        }
      }
      const float4 x4 = *reinterpret_cast<const float4*>(g.x_t + o0);
      const float xt[4] = {x4.x, x4.y, x4.z, x4.w};
      float xp[4], x0[4];
      if (g.step.eps_out) *reinterpret_cast<float4*>(g.step.eps_out + o0) = s4;
      if (g.step.mode == EDTTS_STEP_DDIM) {
#pragma unroll
        for (int j = 0; j < 4; ++j) ddim_update(xt[j], v[j], 0.f, ab_t, ab_p, 0.f, xp[j], x0[j]);
        if (g.step.x0_out) *reinterpret_cast<float4*>(g.step.x0_out + o0) = make_float4(x0[0], x0[1], x0[2], x0[3]);
        if (g.step.write_x_prev && g.step.x_prev_out)
          *reinterpret_cast<float4*>(g.step.x_prev_out + o0) = make_float4(xp[0], xp[1], xp[2], xp[3]);
      } else if (g.step.mode == EDTTS_STEP_DDPM) {
        const float4 n4v = *reinterpret_cast<const float4*>(g.step.noise + o0);
        const float nz[4] = {n4v.x, n4v.y, n4v.z, n4v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) xp[j] = ddpm_update(xt[j], v[j], nz[j], al, ab_t, be, pv, nzm);
        *reinterpret_cast<float4*>(g.step.x_prev_out + o0) = make_float4(xp[0], xp[1], xp[2], xp[3]);
      } else if (g.step.mode == EDTTS_STEP_DPM) {
        const int ord = g.step.dpm_order;
        float h1[4] = {0.f, 0.f, 0.f, 0.f}, h2[4] = {0.f, 0.f, 0.f, 0.f};
        if (ord >= 2) {
          const float4 t4 = *reinterpret_cast<const float4*>(g.step.dpm_hist1 + o0);
          h1[0] = t4.x; h1[1] = t4.y; h1[2] = t4.z; h1[3] = t4.w;
        }
        if (ord >= 3) {
          const float4 t4 = *reinterpret_cast<const float4*>(g.step.dpm_hist2 + o0);
          h2[0] = t4.x; h2[1] = t4.y; h2[2] = t4.z; h2[3] = t4.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dpm_update(xt[j], v[j], h1[j], h2[j], g.step.dpm_coef + bidx * 8, ord, g.step.dpm_predict_x0 ? 1 : 0, xp[j], x0[j]);
        if (g.step.x0_out) *reinterpret_cast<float4*>(g.step.x0_out + o0) = make_float4(x0[0], x0[1], x0[2], x0[3]);
        *reinterpret_cast<float4*>(g.step.x_prev_out + o0) = make_float4(xp[0], xp[1], xp[2], xp[3]);
      }
      continue;
    }
    if (g.epi == EPI_PE) {
      const float4 p4 = __ldg(reinterpret_cast<const float4*>(g.pe + (int64_t)(row % g.pe_period) * g.N + col0));
      v[0] += p4.x; v[1] += p4.y; v[2] += p4.z; v[3] += p4.w;
    }
    *reinterpret_cast<float4*>(g.out + o0) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// pipeline buffers, overlaid by the epilogue tile
static int gemm_main_bytes(int NB, bool swi) {
  const int nout = swi ? NB / 2 : NB;
  const int pipe = 2 * G_A_BUF + G_NW * 2 * (GK / 4) * NB * 16, stage = TM * (nout + 4) * 4;
  return (int)align_up(std::max(pipe, stage), 128);
}
static int gemm_smem(int NB, bool swi) { return gemm_main_bytes(NB, swi) + 2 * TM * 4 + 160 * 4 + (G_NW + 6) * 8 + 16; }

int launch_t3_gemm(const GemmArgs& g, const float* wimg, int64_t img_stride, int NB, float* stats, cudaStream_t st) {
  const bool swi = g.epi == EPI_SWIGLU;
  const int nout = swi ? NB / 2 : NB;
  EDTTS_REQUIRE(g.rows > 0 && g.K % GK == 0 && g.lda % 4 == 0 && g.ldo % 4 == 0 && NB % 16 == 0 && NB <= 160 && nout % 16 == 0 &&
                    g.N % nout == 0,
                EDTTS_EINVAL, "t3_gemm: rows=%lld K=%d N=%d NB=%d unsupported", (long long)g.rows, g.K, g.N, NB);
  EDTTS_REQUIRE(g.pro == PRO_NONE || g.K <= 192, EDTTS_EINVAL, "t3_gemm: norm prologue needs K <= 192 (K=%d)", g.K);
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  EDTTS_REQUIRE(al16(g.A) && al16(g.norm_w) && al16(g.norm_b) && al16(g.mod) && g.mod_stride % 4 == 0 && al16(g.resid) && al16(g.pe) &&
                    al16(g.out) && al16(g.x_t),
                EDTTS_EINVAL, "t3_gemm: operands must be 16-byte aligned");
  EDTTS_REQUIRE(g.epi != EPI_RESID || g.resid == g.out, EDTTS_EINVAL, "t3_gemm: the residual epilogue accumulates in place (resid == out)");
  T3GemmArgs a;
  a.g = g; a.wimg = wimg; a.img_stride = img_stride; a.NB = NB; 
  a.nchunk = g.K / GK; a.main_bytes = gemm_main_bytes(NB, swi); a.stats = stats;
  if (g.pro != PRO_NONE) {
    EDTTS_REQUIRE(stats && (reinterpret_cast<uintptr_t>(stats) & 7) == 0, EDTTS_EINVAL, "t3_gemm: the normalisation prologue needs a statistics buffer");
    LaunchScope ls(KC_TC_MISC, st);
    const int64_t blocks = std::min<int64_t>((g.rows + 31) / 32, stream_grid_cap(8));
    t3_rowstats_kernel<<<(unsigned)blocks, 256, 0, st>>>(g.A, g.rows, g.K, g.lda, g.pro, g.norm_eps, reinterpret_cast<float2*>(stats));
    int rc = check_launch("t3_rowstats");
    if (rc) return rc;
  }
  a.ahead = 2 * sm_count();
  static PerDeviceOnce configured;
  if (configured.need()) {
    if (cudaFuncSetAttribute(t3_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem(160, false)) != cudaSuccess)
      return check_launch("t3_gemm smem attribute");
    configured.set();
  }
  LaunchScope ls(KC_T3_GEMM, st);
  t3_gemm_kernel<<<dim3((unsigned)((g.rows + TM - 1) / TM), g.N / nout), GT_ALL, gemm_smem(NB, swi), st>>>(a);
  return check_launch("t3_gemm");
}

// ---- attention --------------------------------------------------------------------------------------------------------------
// Software pipeline of one CTA (all 128 threads walk the same steps; the MMAs run behind them):
//   iteration i:  A  K(i+1) registers -> operand image (second K buffer)           | tensor core: P V(i-1)
//                 B  sync, S(i+1) = Q K(i+1)^T issued into the second S block       |
//                 C  S(i) (issued an iteration ago) -> registers, softmax           | S(i+1)
//                 D  wait P V(i-1); lazy rescale of O; V(i), P(i) -> operand images |
//                 E  sync, O += P(i) V(i) issued (accumulates in tensor memory)     |
// The MMAs are issued by a fifth warp (one elected lane), woken by per-warp mbarrier arrivals (bar_k: K image written and the S
// block free; bar_p: V and P images written, O rescaled), so no softmax thread sits in the tensor core's queue or waits for an
// MMA round trip it has just asked for, and the four softmax warps never wait for each other.  O stays in tensor memory for the whole head: the running
// maximum is only raised when a block exceeds it by more than 2^8 (p <= 256 otherwise, harmless in fp32 / tf32 x 2), and only
// then does the warp rescale its O rows (tcgen05.ld / st) -- after the first block or two that does not happen.
constexpr int AQ = 128;                               // queries per CTA (= threads = MMA M)
constexpr int AKB = 32;                               // keys per block
constexpr int A_Q_HALF = (HD / 4) * AQ * 16;          // 10 slabs x 128 rows x 16 B
constexpr int A_K_HALF = (HD / 4) * AKB * 16;         // 10 slabs x 32 keys x 16 B
constexpr int A_VN = 48;                              // head_dim padded to an MMA N (rows 40..47 of the V^T image stay zero)
constexpr int A_V_HALF = (AKB / 4) * A_VN * 16;       // 8 slabs x 48 rows x 16 B
constexpr int A_P_HALF = (AKB / 4) * AQ * 16;         // 8 slabs x 128 rows x 16 B
constexpr int A_SMEM = 2 * (A_Q_HALF + 2 * A_K_HALF + A_V_HALF + A_P_HALF) + 128;
constexpr int A_NLD = (AKB * (HD / 4) + AQ - 1) / AQ; // 16-byte pieces of a K (or V) block per thread: 320 / 128 -> 3
constexpr float A_GROW = 8.0f;                        // lazy running maximum: raise it only beyond 2^8
constexpr int A_THREADS = AQ + 32;                    // softmax warps + the MMA issuer warp

// Operand images of context K | V (cross-attention): per (utterance, head, 32-key block) [K hi | K lo | V^T hi | V^T lo], exactly the
// shared-memory layout of the attention kernel, so that its issuer lane streams a block in with one bulk copy each for K and V
// and the softmax warps issue no global load at all (their proxy fences / arrivals wait for every outstanding load of a thread).
constexpr int A_IMG_K = 2 * A_K_HALF, A_IMG_V = 2 * A_V_HALF, A_IMG_BLK = A_IMG_K + A_IMG_V;   // 10,240 + 12,288 bytes

__global__ void __launch_bounds__(AQ) t3_kvimg_kernel(const float* __restrict__ k, const float* __restrict__ v, int kv_stride, int Tk,
                                                      uint8_t* __restrict__ img) {
  __shared__ __align__(16) uint8_t blk[A_IMG_BLK];
  const int tid = threadIdx.x, i = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nb = gridDim.x, kc = i * AKB;
  uint8_t* sKh = blk;
  uint8_t* sKl = blk + A_K_HALF;
  uint8_t* sVh = blk + A_IMG_K;
  uint8_t* sVl = sVh + A_V_HALF;
  if (tid < 64) {                                      // padding rows 40..47 of the V^T images
    const int slab = tid >> 3, row = HD + (tid & 7);
    *reinterpret_cast<float4*>(sVh + slab * (A_VN * 16) + row * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(sVl + slab * (A_VN * 16) + row * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float* kbase = k + (int64_t)b * Tk * kv_stride + h * HD;
  const float* vbase = v + (int64_t)b * Tk * kv_stride + h * HD;
#pragma unroll
  for (int it = 0; it < A_NLD; ++it) {
    const int idx = tid + it * AQ;
    if (idx >= AKB * (HD / 4)) continue;
    const int key = idx / (HD / 4), c4 = idx % (HD / 4);
    float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
    if (kc + key < Tk) {
      kk = *reinterpret_cast<const float4*>(kbase + (int64_t)(kc + key) * kv_stride + 4 * c4);
      vv = *reinterpret_cast<const float4*>(vbase + (int64_t)(kc + key) * kv_stride + 4 * c4);
    }
    {
      const float x[4] = {kk.x, kk.y, kk.z, kk.w};
      float4 hi, lo;
      hi.x = tf32_rna(x[0]); hi.y = tf32_rna(x[1]); hi.z = tf32_rna(x[2]); hi.w = tf32_rna(x[3]);
      lo.x = x[0] - hi.x; lo.y = x[1] - hi.y; lo.z = x[2] - hi.z; lo.w = x[3] - hi.w;
      *reinterpret_cast<float4*>(sKh + c4 * (AKB * 16) + key * 16) = hi;
      *reinterpret_cast<float4*>(sKl + c4 * (AKB * 16) + key * 16) = lo;
    }
    {
      const float x[4] = {vv.x, vv.y, vv.z, vv.w};
      const int off = (key >> 2) * (A_VN * 16) + (key & 3) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float hi = tf32_rna(x[j]);
        *reinterpret_cast<float*>(sVh + off + (4 * c4 + j) * 16) = hi;
        *reinterpret_cast<float*>(sVl + off + (4 * c4 + j) * 16) = x[j] - hi;
      }
    }
  }
  __syncthreads();
  float4* dst = reinterpret_cast<float4*>(img + (((int64_t)b * NH + h) * nb + i) * A_IMG_BLK);
  for (int j = tid; j < A_IMG_BLK / 16; j += AQ) dst[j] = reinterpret_cast<const float4*>(blk)[j];
}

// IMG: the K / V blocks come as ready-made operand images (t3_kvimg_kernel) by bulk copies of the issuer lane; K double-buffered,
// V single-buffered (requested when P V of the previous block has retired); the softmax warps neither load nor convert K / V.
template <bool IMG>
__global__ void __launch_bounds__(A_THREADS, 2) t3_attn_kernel(const AttnArgs a, const uint8_t* __restrict__ kvimg) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sQh = smem;
  uint8_t* sQl = sQh + A_Q_HALF;
  uint8_t* sK = sQl + A_Q_HALF;                        // two buffers of [hi | lo]
  uint8_t* sVh = sK + 4 * A_K_HALF;
  uint8_t* sVl = sVh + A_V_HALF;
  uint8_t* sPh = sVl + A_V_HALF;
  uint8_t* sPl = sPh + A_P_HALF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sPl + A_P_HALF);
  uint64_t* bar_s = bars;                              // [2] S block (i & 1) computed
  uint64_t* bar_o = bars + 2;                          // P V(i) retired
  uint64_t* bar_k = bars + 3;                          // [2] K image of block i written, S block (i & 1) read out (4 warp arrivals)
  uint64_t* bar_p = bars + 5;                          // V and P images of block i written, O rescaled (4 warp arrivals)
  uint64_t* bar_kl = bars + 6;                         // [2] IMG: K image of block i landed in buffer i & 1
  uint64_t* bar_vl = bars + 8;                         // IMG: V image of block i landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool issuer = warp == AQ / 32;
  const int b = blockIdx.z, h = blockIdx.y;
  const int q0 = blockIdx.x * AQ;
  const int qi = q0 + tid;
  const bool active = qi < a.Tq;
  const int W = a.window;

  if (tid == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_s + 1, 1);
    mbar_init(bar_o, 1);
    mbar_init(bar_k, AQ / 32);
    mbar_init(bar_k + 1, AQ / 32);
    mbar_init(bar_p, AQ / 32);
    mbar_init(bar_kl, 1);
    mbar_init(bar_kl + 1, 1);
    mbar_init(bar_vl, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<128>(tmem_slot);
  if (tid < 64) {                                      // padding rows 40..47 of the V^T images
    const int slab = tid >> 3, row = HD + (tid & 7);
    *reinterpret_cast<float4*>(sVh + slab * (A_VN * 16) + row * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(sVl + slab * (A_VN * 16) + row * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (!issuer) {                                       // this thread's query row, pre-scaled by scale * log2(e)
    const float qs = a.scale * 1.4426950408889634f;
    const float* qp = a.q + ((int64_t)b * a.Tq + (active ? qi : 0)) * a.q_stride + h * HD;
    float4 t[HD / 4];
#pragma unroll
    for (int d = 0; d < HD / 4; ++d) t[d] = active ? *reinterpret_cast<const float4*>(qp + 4 * d) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int d = 0; d < HD / 4; ++d) {
      const float v[4] = {t[d].x * qs, t[d].y * qs, t[d].z * qs, t[d].w * qs};
      split_store(sQh, sQl, d, tid, v);
    }
  }

  int klo = 0, khi = a.Tk;
  if (W >= 0) {
    klo = max(0, q0 - W);
    khi = min(a.Tk, q0 + AQ + W);
  }
  const int nblk = (khi - klo + AKB - 1) / AKB;
  const float* kbase = a.k + (int64_t)b * a.Tk * a.kv_stride + h * HD;
  const float* vbase = a.v + (int64_t)b * a.Tk * a.kv_stride + h * HD;

  // this thread's (key, 4 dims) pieces of a 32-key block
  int pkey[A_NLD], pc4[A_NLD];
#pragma unroll
  for (int it = 0; it < A_NLD; ++it) {
    const int idx = tid + it * AQ;
    pkey[it] = idx < AKB * (HD / 4) ? idx / (HD / 4) : -1;
    pc4[it] = idx % (HD / 4);
  }
  float4 kreg[A_NLD], vreg[A_NLD];
  auto fetch = [&](const float* base, int kc, float4 (&dst)[A_NLD]) {       // rows kc .. kc + 31 -> registers (zeros beyond khi)
#pragma unroll
    for (int it = 0; it < A_NLD; ++it) {
      dst[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pkey[it] >= 0 && kc + pkey[it] < khi)
        dst[it] = *reinterpret_cast<const float4*>(base + (int64_t)(kc + pkey[it]) * a.kv_stride + 4 * pc4[it]);
    }
  };
  auto stash_k = [&](int buf) {                        // K image [slab = dim / 4][key][4]
    uint8_t* kh = sK + buf * 2 * A_K_HALF;
#pragma unroll
    for (int it = 0; it < A_NLD; ++it) {
      if (pkey[it] < 0) continue;
      const float v[4] = {kreg[it].x, kreg[it].y, kreg[it].z, kreg[it].w};
      float4 hi, lo;
      hi.x = tf32_rna(v[0]); hi.y = tf32_rna(v[1]); hi.z = tf32_rna(v[2]); hi.w = tf32_rna(v[3]);
      lo.x = v[0] - hi.x; lo.y = v[1] - hi.y; lo.z = v[2] - hi.z; lo.w = v[3] - hi.w;
      *reinterpret_cast<float4*>(kh + pc4[it] * (AKB * 16) + pkey[it] * 16) = hi;
      *reinterpret_cast<float4*>(kh + A_K_HALF + pc4[it] * (AKB * 16) + pkey[it] * 16) = lo;
    }
  };
  auto stash_v = [&]() {                               // V^T image [slab = key / 4][dim][4]
#pragma unroll
    for (int it = 0; it < A_NLD; ++it) {
      if (pkey[it] < 0) continue;
      const float v[4] = {vreg[it].x, vreg[it].y, vreg[it].z, vreg[it].w};
      const int off = (pkey[it] >> 2) * (A_VN * 16) + (pkey[it] & 3) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float hi = tf32_rna(v[j]);
        *reinterpret_cast<float*>(sVh + off + (4 * pc4[it] + j) * 16) = hi;
        *reinterpret_cast<float*>(sVl + off + (4 * pc4[it] + j) * 16) = v[j] - hi;
      }
    }
  };

  // ---- prologue: K(0) staged ----
  if (!issuer) {
    if (!IMG) {
      fetch(kbase, klo, kreg);
      stash_k(0);
      if (nblk > 1) fetch(kbase, klo + AKB, kreg);
      fetch(vbase, klo, vreg);
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tO = 2 * AKB;                         // columns: S0 (32) | S1 (32) | O (48)

  // ================= MMA issuer: warp 4, one elected lane =================
  if (issuer) {
    if (lane == 0) {
      const uint8_t* img = IMG ? kvimg + ((int64_t)b * NH + h) * nblk * A_IMG_BLK : nullptr;   // klo = 0: block i <-> image block i
      auto req_k = [&](int i) {
        mbar_expect_tx(bar_kl + (i & 1), A_IMG_K);
        bulk_g2s(sK + (i & 1) * A_IMG_K, img + (int64_t)i * A_IMG_BLK, A_IMG_K, bar_kl + (i & 1));
      };
      auto req_v = [&](int i) {
        mbar_expect_tx(bar_vl, A_IMG_V);
        bulk_g2s(sVh, img + (int64_t)i * A_IMG_BLK + A_IMG_K, A_IMG_V, bar_vl);
      };
      if (IMG) {
        req_k(0);
        if (nblk > 1) req_k(1);
        req_v(0);
      }
      auto issue_s = [&](int i) {
        mbar_wait_issuer(bar_k + (i & 1), (i >> 1) & 1);
        if (IMG) mbar_wait_issuer(bar_kl + (i & 1), (i >> 1) & 1);
        tc_fence_after();
        const uint32_t kh = smem_u32(sK + (i & 1) * 2 * A_K_HALF);
        issue_chunk(tmem + (i & 1) * AKB, smem_u32(sQh), smem_u32(sQl), kh, kh + A_K_HALF, HD / 8, AKB, false);
        umma_commit(bar_s + (i & 1));
      };
      issue_s(0);
      for (int blk = 0; blk < nblk; ++blk) {
        if (blk + 1 < nblk) issue_s(blk + 1);
        if (IMG && blk + 2 < nblk) {                   // K buffer blk & 1 is free once S(blk) has retired
          mbar_wait_issuer(bar_s + (blk & 1), (blk >> 1) & 1);
          req_k(blk + 2);
        }
        mbar_wait_issuer(bar_p, blk & 1);
        if (IMG) mbar_wait_issuer(bar_vl, blk & 1);
        tc_fence_after();
        issue_chunk(tmem + tO, smem_u32(sPh), smem_u32(sPl), smem_u32(sVh), smem_u32(sVl), AKB / 8, A_VN, blk > 0);
        umma_commit(bar_o);
        if (IMG && blk + 1 < nblk) {                   // the single V buffer is free once P V(blk) has retired
          mbar_wait_issuer(bar_o, blk & 1);
          req_v(blk + 1);
        }
      }
    }
    return;
  }

  // ================= softmax warps 0-3 =================
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  __syncwarp();
  if (lane == 0) mbar_arrive(bar_k);                   // K(0) (and Q, the V padding) written

  float m = -INFINITY, l = 0.f;
  const int r0w = q0 + warp * 32;                      // first query row of this warp

  for (int blk = 0; blk < nblk; ++blk) {
    const int kc = klo + blk * AKB;
    // ---- A: K(blk+1) -> the other K buffer (S(blk-1), its last reader, retired: this thread waited for it); this warp has read
    // S(blk-1) out of the tensor-memory block that S(blk+1) will overwrite ----
    if (blk + 1 < nblk) {
      if (!IMG) {
        stash_k((blk + 1) & 1);
        if (blk + 2 < nblk) fetch(kbase, kc + 2 * AKB, kreg);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_k + ((blk + 1) & 1));
    }
    // ---- C: S(blk) -> registers, softmax ----
    mbar_wait(bar_s + (blk & 1), (blk >> 1) & 1);
    tc_fence_after();
    // What this WARP's 32 rows see of the block's keys (warp-uniform, so the collective tensor-memory accesses stay converged):
    // empty -> p = 0, nothing else to do; full -> every (row, key) pair is valid, no masks; else per-element masks.
    bool empty, full;
    if (W >= 0) {
      empty = r0w >= a.Tq || kc > r0w + 31 + W || kc + AKB - 1 < r0w - W;
      full = r0w + 31 < a.Tq && kc >= r0w + 31 - W && kc + AKB - 1 <= r0w + W && kc + AKB <= khi;
    } else {
      empty = r0w >= a.Tq;
      full = r0w + 31 < a.Tq && kc + AKB <= khi;
    }
    float f = 1.0f;                                    // factor this thread's O row needs before P V(blk) is added
    bool grow = false;
    float s[AKB];
    if (empty) {
#pragma unroll
      for (int j = 0; j < AKB; ++j) s[j] = 0.f;
    } else {
      tmem_ld32(trow + (blk & 1) * AKB, s);
      if (!full) {
#pragma unroll
        for (int j = 0; j < AKB; ++j) {
          const int key = kc + j;
          const bool ok = active && key < khi && (W < 0 || (key >= qi - W && key <= qi + W));
          s[j] = ok ? s[j] : -INFINITY;
        }
      }
      float b0 = s[0], b1 = s[1], b2 = s[2], b3 = s[3];
#pragma unroll
      for (int j = 4; j < AKB; j += 4) {
        b0 = fmaxf(b0, s[j]); b1 = fmaxf(b1, s[j + 1]); b2 = fmaxf(b2, s[j + 2]); b3 = fmaxf(b3, s[j + 3]);
      }
      const float bm = fmaxf(fmaxf(b0, b1), fmaxf(b2, b3));
      grow = bm > m + A_GROW;                          // also the first finite block maximum (m = -inf)
      const float m_old = m;
      m = grow ? bm : m;
      const float m_use = (m == -INFINITY) ? 0.f : m;  // a row that has not met a valid key yet: every p = 0
      f = grow ? ex2_approx(m_old - m_use) : 1.0f;     // m_old = -inf -> 0
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
      for (int j = 0; j < AKB; j += 4) {
        s[j] = ex2_approx(s[j] - m_use);               // masked: ex2(-inf) = 0
        s[j + 1] = ex2_approx(s[j + 1] - m_use);
        s[j + 2] = ex2_approx(s[j + 2] - m_use);
        s[j + 3] = ex2_approx(s[j + 3] - m_use);
        l0 += s[j]; l1 += s[j + 1]; l2 += s[j + 2]; l3 += s[j + 3];
      }
      l = fmaf(l, f, (l0 + l1) + (l2 + l3));
    }
    // ---- D: P V(blk-1) retired -> O may be rescaled, the V and P images rewritten ----
    if (blk > 0) {
      mbar_wait(bar_o, (blk - 1) & 1);
      tc_fence_after();
      if (!empty && __any_sync(0xffffffffu, grow)) {
        float o[HD];
        tmem_ld40(trow + tO, o);
#pragma unroll
        for (int d = 0; d < HD; ++d) o[d] *= f;
        tmem_st40(trow + tO, o);
        tmem_st_wait();
      }
    }
    if (!IMG) {
      stash_v();
      if (blk + 1 < nblk) fetch(vbase, kc + AKB, vreg);
    }
#pragma unroll
    for (int q4 = 0; q4 < AKB / 4; ++q4) split_store(sPh, sPl, q4, tid, s + 4 * q4);
    // ---- E: hand P(blk), V(blk) to the issuer ----
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p);
  }
  mbar_wait(bar_o, (nblk - 1) & 1);
  tc_fence_after();
  {
    float o[HD];
    tmem_ld40(trow + tO, o);
    if (active) {
      const float inv = 1.0f / l;
      float* op = a.o + ((int64_t)b * a.Tq + qi) * a.o_stride + h * HD;
#pragma unroll
      for (int d = 0; d < HD; d += 4)
        *reinterpret_cast<float4*>(op + d) = make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
    }
  }
  tc_fence_before();
  named_bar_sync(1, AQ);
  if (warp == 0) tmem_dealloc<128>(tmem);
}

int64_t t3_kvimg_bytes(int B, int Tk) { return (int64_t)B * NH * ((Tk + AKB - 1) / AKB) * A_IMG_BLK; }

// kvimg: scratch of t3_kvimg_bytes(B, Tk) bytes for the operand images of a full-context (window < 0) attention, or null: the softmax
// warps then stage K / V themselves (always for the band attention, whose k | v change with every layer and step)
int64_t t3_kv_rows_bytes(int B, int S) { return align_up((int64_t)NL * B * S * 2 * H * 4, 256); }
int64_t t3_kv_total_bytes(int B, int S) { return t3_kv_rows_bytes(B, S) + NL * align_up(t3_kvimg_bytes(B, S), 256); }

int launch_t3_kvimg(const float* k, const float* v, int kv_stride, int Tk, int B, void* kvimg, cudaStream_t st) {
  LaunchScope ls(KC_TC_MISC, st);
  t3_kvimg_kernel<<<dim3((Tk + AKB - 1) / AKB, NH, B), AQ, 0, st>>>(k, v, kv_stride, Tk, reinterpret_cast<uint8_t*>(kvimg));
  return check_launch("t3_kvimg");
}

int launch_t3_attn(const AttnArgs& a, int B, void* kvimg, bool build, cudaStream_t st) {
  EDTTS_REQUIRE(a.q_stride % 4 == 0 && a.kv_stride % 4 == 0 && a.o_stride % 4 == 0, EDTTS_EINVAL, "t3_attn: strides must be multiples of 4");
  static PerDeviceOnce configured;
  if (configured.need()) {
    if (cudaFuncSetAttribute(t3_attn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, A_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(t3_attn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, A_SMEM) != cudaSuccess)
      return check_launch("t3_attn smem attribute");
    configured.set();
  }
  dim3 grid((a.Tq + AQ - 1) / AQ, NH, B);
  if (kvimg && a.window < 0) {
    if (build) {
      int rc = launch_t3_kvimg(a.k, a.v, a.kv_stride, a.Tk, B, kvimg, st);
      if (rc) return rc;
    }
    LaunchScope ls(KC_T3_ATTN_CROSS, st);
    t3_attn_kernel<true><<<grid, A_THREADS, A_SMEM, st>>>(a, reinterpret_cast<const uint8_t*>(kvimg));
    return check_launch("t3_attn");
  }
  LaunchScope ls(a.window >= 0 ? KC_T3_ATTN_WINDOW : KC_T3_ATTN_CROSS, st);
  t3_attn_kernel<false><<<grid, A_THREADS, A_SMEM, st>>>(a, nullptr);
  return check_launch("t3_attn");
}

// ---- context K | V (once per utterance) ----------------------------------------------------------------------------------------
static int64_t ctx_img_floats() { return t3_gemm_image_floats(H, RANK, 80, false) + t3_gemm_image_floats(RANK, 2 * H, 160, false); }
int64_t t3_context_scratch_bytes(int64_t rows) { return align_up(NL * ctx_img_floats() * 4, 256) + align_up(rows * 8, 256); }

int t3_context_kv(const edtts_decoder_weights* w, const float* ctx, float* craw, float* kv_out, void* scratch, int64_t rows, int B, int S,
                  cudaStream_t st) {
  float* img = reinterpret_cast<float*>(scratch);
  float* stats = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + align_up(NL * ctx_img_floats() * 4, 256));
  const int64_t o_up = t3_gemm_image_floats(H, RANK, 80, false);
  int rc;
  {
    PackJobs pj;
    int n = 0, k;
    for (int l = 0; l < NL; ++l) {
      const edtts_layer_weights& L = w->layers[l];
      if ((k = add_jobs(pj, n, L.kv_down_w, img + l * ctx_img_floats(), H, RANK, 80, false)) < 0) return EDTTS_EINVAL; n += k;
      if ((k = add_jobs(pj, n, L.kv_up_w, img + l * ctx_img_floats() + o_up, RANK, 2 * H, 160, false)) < 0) return EDTTS_EINVAL; n += k;
    }
    if ((rc = launch_pack(pj, n, st))) return rc;
  }
  for (int l = 0; l < NL; ++l) {
    const edtts_layer_weights& L = w->layers[l];
    const float* li = img + l * ctx_img_floats();
    GemmArgs d;   // kv_down_proj (mla.py:146)
    d.A = ctx; d.rows = rows; d.K = H; d.lda = H; d.W = L.kv_down_w; d.N = RANK; d.out = craw; d.ldo = RANK;
    if ((rc = launch_t3_gemm(d, li, t3_gemm_block_stride(H, 80), 80, stats, st))) return rc;
    GemmArgs u;   // kv_norm + kv_up_proj (mla.py:147-153)
    u.A = craw; u.rows = rows; u.K = RANK; u.lda = RANK; u.W = L.kv_up_w; u.N = 2 * H;
    u.out = kv_out + (int64_t)l * rows * 2 * H; u.ldo = 2 * H;
    u.pro = PRO_RMS; u.norm_w = L.kv_norm_w; u.norm_eps = 1e-6f;
    if ((rc = launch_t3_gemm(u, li + o_up, t3_gemm_block_stride(RANK, 160), 160, stats, st))) return rc;
    // the layer's operand images for the cross-attention of every decoder step of this generate, behind the fp32 rows
    char* img_l = reinterpret_cast<char*>(kv_out) + t3_kv_rows_bytes(B, S) + l * align_up(t3_kvimg_bytes(B, S), 256);
    if ((rc = launch_t3_kvimg(u.out, u.out + H, 2 * H, S, B, img_l, st))) return rc;
  }
  return EDTTS_OK;
}

// ---- one decoder evaluation ---------------------------------------------------------------------------------------------------
// weight images, in the order the step uses them (floats)
struct ImgLayout {
  int64_t in_proj, out_proj, layer[NL], total;
  int64_t qkv, proj, q, out, f0, f3;                   // offsets inside a layer
};
static ImgLayout img_layout() {
  ImgLayout L;
  int64_t o = 0;
  L.in_proj = o; o += t3_gemm_image_floats(M, H, 160, false);
  L.out_proj = o; o += t3_gemm_image_floats(H, M, 80, false);
  int64_t p = 0;
  L.qkv = p; p += t3_gemm_image_floats(H, 3 * H, 160, false);
  L.proj = p; p += t3_gemm_image_floats(H, H, 160, false);
  L.q = p; p += t3_gemm_image_floats(H, H, 160, false);
  L.out = p; p += t3_gemm_image_floats(H, H, 160, false);
  L.f0 = p; p += t3_gemm_image_floats(H, FFN, 160, true);
  L.f3 = p; p += t3_gemm_image_floats(FFN, H, 160, false);
  for (int l = 0; l < NL; ++l) { L.layer[l] = o; o += p; }
  L.total = o;
  return L;
}

int64_t t3_decoder_workspace_bytes(int B, int T, int S) {
  (void)S;
  const int64_t R = (int64_t)B * T;
  return align_up(R * H * 4, 256) * 2 + align_up(R * 3 * H * 4, 256) + align_up(img_layout().total * 4, 256) + align_up(R * 8, 256);
}

int t3_decoder_step(const edtts_decoder_weights* w, const float* x_t, const float* mod, const float* kv, const edtts_step_args* args,
                    void* workspace, int B, int T, int S, cudaStream_t st) {
  const int64_t R = (int64_t)B * T;
  char* ws = reinterpret_cast<char*>(workspace);
  float* h = reinterpret_cast<float*>(ws);
  float* a = reinterpret_cast<float*>(ws + align_up(R * H * 4, 256));
  float* big = reinterpret_cast<float*>(ws + 2 * align_up(R * H * 4, 256));
  float* img = reinterpret_cast<float*>(ws + 2 * align_up(R * H * 4, 256) + align_up(R * 3 * H * 4, 256));
  float* stats = reinterpret_cast<float*>(ws + 2 * align_up(R * H * 4, 256) + align_up(R * 3 * H * 4, 256) + align_up(img_layout().total * 4, 256));
  const ImgLayout IL = img_layout();
  const float scale = 1.0f / sqrtf((float)HD);
  const int64_t s160 = t3_gemm_block_stride(H, 160), s80 = t3_gemm_block_stride(H, 80), s_in = t3_gemm_block_stride(M, 160),
                s320 = t3_gemm_block_stride(FFN, 160);
  int rc;
  {  // weight images of this step (the parameters are read live, as on the CUDA-core path): one launch
    PackJobs pj;
    int n = 0, k;
    if ((k = add_jobs(pj, n, w->in_proj_w, img + IL.in_proj, M, H, 160, false)) < 0) return EDTTS_EINVAL; n += k;
    if ((k = add_jobs(pj, n, w->out_proj_w, img + IL.out_proj, H, M, 80, false)) < 0) return EDTTS_EINVAL; n += k;
    for (int l = 0; l < NL; ++l) {
      const edtts_layer_weights& L = w->layers[l];
      float* li = img + IL.layer[l];
      if ((k = add_jobs(pj, n, L.attn_qkv_w, li + IL.qkv, H, 3 * H, 160, false)) < 0) return EDTTS_EINVAL; n += k;
      if ((k = add_jobs(pj, n, L.attn_proj_w, li + IL.proj, H, H, 160, false)) < 0) return EDTTS_EINVAL; n += k;
      if ((k = add_jobs(pj, n, L.q_proj_w, li + IL.q, H, H, 160, false)) < 0) return EDTTS_EINVAL; n += k;
      if ((k = add_jobs(pj, n, L.cross_out_w, li + IL.out, H, H, 160, false)) < 0) return EDTTS_EINVAL; n += k;
      if ((k = add_jobs(pj, n, L.ffn0_w, li + IL.f0, H, FFN, 160, true)) < 0) return EDTTS_EINVAL; n += k;
      if ((k = add_jobs(pj, n, L.ffn3_w, li + IL.f3, FFN, H, 160, false)) < 0) return EDTTS_EINVAL; n += k;
    }
    if ((rc = launch_pack(pj, n, st))) return rc;
  }
  {  // h = in_proj(x_t) + pe[:T]   (decoder.py:96-97)
    GemmArgs g;
    g.A = x_t; g.rows = R; g.K = M; g.lda = M; g.W = w->in_proj_w; g.N = H; g.bias = w->in_proj_b;
    g.out = h; g.ldo = H; g.epi = EPI_PE; g.pe = w->pos_pe; g.pe_period = T;
    if ((rc = launch_t3_gemm(g, img + IL.in_proj, s_in, 160, stats, st))) return rc;
  }
  for (int l = 0; l < NL; ++l) {
    const edtts_layer_weights& L = w->layers[l];
    const float* li = img + IL.layer[l];
    {  // qkv = attn.qkv(norm1(h, cond))   (transformer.py:143, attention.py:90)
      GemmArgs g;
      g.A = h; g.rows = R; g.K = H; g.lda = H; g.W = L.attn_qkv_w; g.N = 3 * H; g.out = big; g.ldo = 3 * H;
      g.pro = PRO_ADARMS; g.norm_w = L.norm1_norm_w; g.mod = mod + (int64_t)(2 * l) * 2 * H;
      g.mod_stride = 2 * NL * 2 * H; g.rows_per_batch = T;
      if ((rc = launch_t3_gemm(g, li + IL.qkv, s160, 160, stats, st))) return rc;
    }
    {  // banded self-attention (attention.py:94-111)
      AttnArgs at{big, 3 * H, big + H, big + 2 * H, 3 * H, a, H, T, T, WIN, scale};
      if ((rc = launch_t3_attn(at, B, nullptr, false, st))) return rc;
    }
    {  // h += attn.proj(o)   (attention.py:123, transformer.py:146)
      GemmArgs g;
      g.A = a; g.rows = R; g.K = H; g.lda = H; g.W = L.attn_proj_w; g.N = H; g.bias = L.attn_proj_b;
      g.out = h; g.ldo = H; g.epi = EPI_RESID; g.resid = h;
      if ((rc = launch_t3_gemm(g, li + IL.proj, s160, 160, stats, st))) return rc;
    }
    {  // q = q_proj(norm2(h))   (transformer.py:151, mla.py:139)
      GemmArgs g;
      g.A = h; g.rows = R; g.K = H; g.lda = H; g.W = L.q_proj_w; g.N = H; g.out = big; g.ldo = H;
      g.pro = PRO_RMS; g.norm_w = L.norm2_w;
      if ((rc = launch_t3_gemm(g, li + IL.q, s160, 160, stats, st))) return rc;
    }
    {  // full cross-attention over the S context tokens (mla.py:176-180)
      const float* kvl = kv + (int64_t)l * B * S * 2 * H;
      AttnArgs at{big, H, kvl, kvl + H, 2 * H, a, H, T, S, -1, scale};
      // operand images of this layer's context k | v, written once per generate by edtts_context_prepare behind the fp32 rows of kv
      void* kvimg = const_cast<char*>(reinterpret_cast<const char*>(kv)) + t3_kv_rows_bytes(B, S) + l * align_up(t3_kvimg_bytes(B, S), 256);
      if ((rc = launch_t3_attn(at, B, kvimg, false, st))) return rc;
    }
    {  // h += out_proj(o)   (mla.py:194)
      GemmArgs g;
      g.A = a; g.rows = R; g.K = H; g.lda = H; g.W = L.cross_out_w; g.N = H; g.out = h; g.ldo = H;
      g.epi = EPI_RESID; g.resid = h;
      if ((rc = launch_t3_gemm(g, li + IL.out, s160, 160, stats, st))) return rc;
    }
    {  // u = swiglu(ffn.net.0(norm3(h, cond)))   (transformer.py:155, :13-23)
      GemmArgs g;
      g.A = h; g.rows = R; g.K = H; g.lda = H; g.W = L.ffn0_w; g.N = FFN; g.bias = L.ffn0_b;
      g.out = big; g.ldo = FFN; g.epi = EPI_SWIGLU;
      g.pro = PRO_ADARMS; g.norm_w = L.norm3_norm_w; g.mod = mod + (int64_t)(2 * l + 1) * 2 * H;
      g.mod_stride = 2 * NL * 2 * H; g.rows_per_batch = T;
      if ((rc = launch_t3_gemm(g, li + IL.f0, s160, 160, stats, st))) return rc;
    }
    {  // h += ffn.net.3(u)
      GemmArgs g;
      g.A = big; g.rows = R; g.K = FFN; g.lda = FFN; g.W = L.ffn3_w; g.N = H; g.bias = L.ffn3_b;
      g.out = h; g.ldo = H; g.epi = EPI_RESID; g.resid = h;
      if ((rc = launch_t3_gemm(g, li + IL.f3, s320, 160, stats, st))) return rc;
    }
  }
  {  // eps = out_proj(final_norm(h)) + fused update   (decoder.py:108-109, schedule.py)
    GemmArgs g;
    g.A = h; g.rows = R; g.K = H; g.lda = H; g.W = w->out_proj_w; g.N = M; g.bias = w->out_proj_b;
    g.out = nullptr; g.ldo = M; g.epi = EPI_STEP; g.pro = PRO_LN; g.norm_w = w->final_norm_w;
    g.norm_b = w->final_norm_b; g.norm_eps = 1e-5f; g.rows_per_batch = T; g.x_t = x_t; g.step = *args;
    if ((rc = launch_t3_gemm(g, img + IL.out_proj, s80, 80, stats, st))) return rc;
  }
  return EDTTS_OK;
}

}  // namespace t3
}  // namespace edtts
