// EdgeDiffusionDecoder on B200: conditioning, context/KV preparation and one
// decoder evaluation with the DDIM/DDPM update fused into the last kernel.
// Reference: models/decoder.py:66-109, layers/transformer.py:129-160.
#include "common.cuh"
#include "gemm_simt.cuh"
#include "attention_simt.cuh"
#include "tc_path.cuh"
#include "t3_decoder.cuh"

namespace edtts {

int launch_gemm_simt(const GemmArgs& g, cudaStream_t stream) {
  EDTTS_REQUIRE(g.rows > 0 && g.K % G_BK == 0 && g.lda % 4 == 0, EDTTS_EINVAL,
                "gemm: rows=%lld K=%d lda=%d unsupported", (long long)g.rows, g.K, g.lda);
  EDTTS_REQUIRE(g.pro == PRO_NONE || g.K <= 192, EDTTS_EINVAL, "gemm: norm prologue needs K <= 192 (K=%d)", g.K);
  const unsigned gx = (unsigned)((g.rows + G_BM - 1) / G_BM);
  const bool dual = g.epi == EPI_SWIGLU;
  LaunchScope ls(KC_GEMM_SIMT, stream);
  if (g.N % 80 == 0) {
    dim3 grid(gx, g.N / 80);
    if (dual) gemm_simt_kernel<5, true><<<grid, G_THREADS, 0, stream>>>(g);
    else gemm_simt_kernel<5, false><<<grid, G_THREADS, 0, stream>>>(g);
  } else if (g.N % 64 == 0 && !dual) {
    dim3 grid(gx, g.N / 64);
    gemm_simt_kernel<4, false><<<grid, G_THREADS, 0, stream>>>(g);
  } else {
    set_error("gemm: N=%d must be a multiple of 80 or 64", g.N);
    return EDTTS_EINVAL;
  }
  return check_launch("gemm_simt");
}

int launch_attn_simt(const AttnArgs& a, int B, cudaStream_t stream) {
  dim3 grid((a.Tq + AT_Q - 1) / AT_Q, NH, B);
  LaunchScope ls(a.window >= 0 ? KC_ATTN_WINDOW_SIMT : KC_ATTN_CROSS_SIMT, stream);
  attn_simt_kernel<<<grid, AT_Q, 0, stream>>>(a);
  return check_launch("attn_simt");
}

// ---------------------------------------------------------------------------
// Conditioning: t -> sinusoid -> Linear -> GELU -> Linear (+ step_emb) -> cond,
// then the 8 AdaLayerNorm projections (decoder.py:77-80, transformer.py:64-66).
// One block per CU utterances and per quarter of one AdaLayerNorm projection (the small time MLP is recomputed per block:
// 2 x 102 KB of weights from L2).  A WARP owns an output feature: the lanes split the 160-long weight row (one coalesced
// 640-byte read, 5 values per lane), multiply it with the CU input rows in shared memory and reduce with shuffles -- the
// weight read latency of several features overlaps (unrolled), instead of one thread walking a row serially.  A row's
// result does not depend on its position in the batch (fixed lane partition, fixed shuffle tree).
// ---------------------------------------------------------------------------
constexpr int CU = 8;
constexpr int CSPLIT = 4;                 // blocks per AdaLayerNorm projection (80 of its 320 outputs each)

constexpr int CTHREADS = 1024, CWARPS = CTHREADS / 32;

// One stage of NOUT_ outputs starting at j0: warp w owns outputs j0 + w, j0 + w + 32, ...  ALL weight values of the warp's
// outputs are requested before the first dot product (the rows come from L2 or HBM: one exposed latency per stage instead of
// one per output); lane l takes k = l, l + 32, ..., l + 128 of a row.  fn(j, u, value) receives the dot product (without
// bias) of output j for row u in lane u.
template <int NOUT_, typename F>
__device__ __forceinline__ void warp_stage(const float* __restrict__ W, int j0, const float (*x)[H], int warp, int lane, F fn) {
  constexpr int PER = (NOUT_ + CWARPS - 1) / CWARPS;
  float wv[PER][5];
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int j = warp + CWARPS * i;
#pragma unroll
    for (int q = 0; q < 5; ++q) wv[i][q] = j < NOUT_ ? __ldg(W + (int64_t)(j0 + j) * H + lane + 32 * q) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int j = warp + CWARPS * i;
    if (j >= NOUT_) break;                                // warp-uniform
    float mine = 0.f;
#pragma unroll
    for (int u = 0; u < CU; ++u) {
      float p = 0.f;
#pragma unroll
      for (int q = 0; q < 5; ++q) p = fmaf(wv[i][q], x[u][lane + 32 * q], p);
      p = warp_sum(p);
      mine = lane == u ? p : mine;
    }
    if (lane < CU) fn(j0 + j, lane, mine);
  }
}

__global__ void __launch_bounds__(CTHREADS) cond_kernel(const edtts_decoder_weights w, const int64_t* __restrict__ t,
                                                        const int64_t* __restrict__ step_idx, float* __restrict__ cond_out,
                                                        float* __restrict__ mod_out, int B) {
  __shared__ __align__(16) float e[CU][H], h1[CU][H], c[CU][H];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b0 = blockIdx.x * CU;
  const int nu = min(CU, B - b0);
  for (int i = tid; i < CU * (H / 2); i += CTHREADS) {
    const int u = i / (H / 2), k = i % (H / 2);
    const float arg = u < nu ? (float)t[b0 + u] * w.time_freqs[k] : 0.f;   // embeddings.py:42-43
    e[u][k] = sinf(arg);
    e[u][k + H / 2] = cosf(arg);
  }
  __syncthreads();
  warp_stage<H>(w.time1_w, 0, e, warp, lane, [&](int o, int u, float v) { h1[u][o] = gelu_erf(v + w.time1_b[o]); });
  __syncthreads();
  warp_stage<H>(w.time3_w, 0, h1, warp, lane, [&](int o, int u, float v) {
    float s = v + w.time3_b[o];
    if (u < nu) {
      if (step_idx) s += w.step_emb[step_idx[b0 + u] * H + o];
      if (cond_out && blockIdx.y == 0) cond_out[(int64_t)(b0 + u) * H + o] = s;
    }
    c[u][o] = s;
  });
  __syncthreads();
  if (!mod_out) return;
  // blockIdx.y: which of the 8 AdaLayerNorms, and which quarter of its 2 H outputs
  constexpr int NOUT = 2 * H;
  const int which = blockIdx.y / CSPLIT, part = blockIdx.y % CSPLIT;
  const edtts_layer_weights& L = w.layers[which >> 1];
  const float* pw = (which & 1) ? L.norm3_proj_w : L.norm1_proj_w;
  const float* pb = (which & 1) ? L.norm3_proj_b : L.norm1_proj_b;
  warp_stage<NOUT / CSPLIT>(pw, part * (NOUT / CSPLIT), c, warp, lane, [&](int j, int u, float v) {
    if (u < nu) mod_out[((int64_t)(b0 + u) * 2 * NL + which) * 2 * H + j] = v + pb[j];
  });
}

// ctx[b,s,:] = token_emb[sem_idx[b,s]] + pe[s]   (decoder.py:88,93)
__global__ void embed_ctx_kernel(const float* __restrict__ emb, const float* __restrict__ pe,
                                 const int64_t* __restrict__ idx, float* __restrict__ ctx, int64_t rows, int S,
                                 int codebook) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // one float4 each
  if (i >= rows * (H / 4)) return;
  const int64_t r = i / (H / 4);
  const int c = (int)(i % (H / 4)) * 4;
  int64_t tok = idx[r];
  tok = tok < 0 ? 0 : (tok >= codebook ? codebook - 1 : tok);
  const float4 a = *reinterpret_cast<const float4*>(emb + tok * H + c);
  const float4 p = *reinterpret_cast<const float4*>(pe + (r % S) * H + c);
  *reinterpret_cast<float4*>(ctx + r * H + c) = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
}

}  // namespace edtts

using namespace edtts;

extern "C" int edtts_cond_prepare(const edtts_decoder_weights* w, const int64_t* t, const int64_t* step_idx,
                                  float* cond_out, float* mod_out, int32_t B, void* stream) {
  EDTTS_REQUIRE(w && t && B > 0 && (cond_out || mod_out), EDTTS_EINVAL, "cond_prepare: null argument");
  LaunchScope ls(KC_COND, as_stream(stream));
  cond_kernel<<<dim3((B + CU - 1) / CU, mod_out ? 2 * NL * CSPLIT : 1), CTHREADS, 0, as_stream(stream)>>>(*w, t, step_idx, cond_out, mod_out, B);
  return check_launch("cond_kernel");
}

extern "C" int64_t edtts_context_kv_bytes(int32_t B, int32_t S, int32_t precision) {
  if (precision == EDTTS_PREC_TF32X3) return t3::t3_kv_total_bytes(B, S);   // fp32 rows + the cross-attention operand images
  return (int64_t)NL * B * S * 2 * H * 4;
}

extern "C" int64_t edtts_context_workspace_bytes(int32_t B, int32_t S) {
  const int64_t rows = (int64_t)B * S;
  return align_up(rows * H * 4, 256) + align_up(rows * RANK * 4, 256) + align_up(rows * 2 * H * 4, 256) + t3::t3_context_scratch_bytes(rows);
}

extern "C" int edtts_context_prepare(const edtts_decoder_weights* w, const int64_t* sem_idx, const float* sem_features,
                                     float* kv_out, void* workspace, int64_t workspace_bytes, int32_t B, int32_t S,
                                     int32_t precision, void* stream) {
  EDTTS_REQUIRE(w && kv_out && workspace && B > 0 && S > 0, EDTTS_EINVAL, "context_prepare: null argument");
  EDTTS_REQUIRE((sem_idx != nullptr) != (sem_features != nullptr), EDTTS_EINVAL,
                "Either sem_idx or sem_features must be provided");   // decoder.py:90
  EDTTS_REQUIRE(S <= w->ctx_rows, EDTTS_EINVAL, "context_prepare: S=%d exceeds the %d-row context PE table", S,
                w->ctx_rows);
  EDTTS_REQUIRE(workspace_bytes >= edtts_context_workspace_bytes(B, S), EDTTS_ENOSPC, "context_prepare: workspace");
  // step-invariant and <4% of the FLOPs: always computed by the fp32 kernels; for the bf16 path the
  // result is then stored as the bf16 chunk-major operand image the tcgen05 cross-attention fetches
  cudaStream_t st = as_stream(stream);
  const int64_t rows = (int64_t)B * S;
  float* ctx = reinterpret_cast<float*>(workspace);
  float* craw = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + align_up(rows * H * 4, 256));
  if (sem_idx && precision == EDTTS_PREC_BF16)   // embedding + positional rows -> bf16 operand image -> k | v of all layers
    return tc_context_kv(w, nullptr, sem_idx, S, craw, kv_out, rows, st);
  if (sem_idx) {
    const int64_t n4 = rows * (H / 4);
    LaunchScope ls(KC_EMBED, st);
    embed_ctx_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(w->token_emb, w->ctx_pe, sem_idx, ctx, rows, S,
                                                                   w->codebook_size);
    int rc = check_launch("embed_ctx");
    if (rc) return rc;
  } else {  // sem_proj (decoder.py:85) + context PE
    GemmArgs g;
    g.A = sem_features; g.rows = rows; g.K = EDTTS_SEMANTIC_DIM; g.lda = EDTTS_SEMANTIC_DIM;
    g.W = w->sem_proj_w; g.N = H; g.bias = w->sem_proj_b; g.out = ctx; g.ldo = H;
    g.epi = EPI_PE; g.pe = w->ctx_pe; g.pe_period = S;
    int rc = launch_gemm_simt(g, st);
    if (rc) return rc;
  }
  if (precision == EDTTS_PREC_BF16)   // kv_down -> kv_norm -> kv_up on the tensor cores, stored as the attention operand image
    return tc_context_kv(w, ctx, nullptr, S, craw, kv_out, rows, st);
  if (precision == EDTTS_PREC_TF32X3)   // the same projections as tf32 x 3 GEMMs, fp32 rows out
    return t3::t3_context_kv(w, ctx, craw, kv_out,
                             reinterpret_cast<char*>(workspace) + align_up(rows * H * 4, 256) + align_up(rows * RANK * 4, 256) +
                                 align_up(rows * 2 * H * 4, 256),
                             rows, B, S, st);
  for (int l = 0; l < NL; ++l) {
    const edtts_layer_weights& L = w->layers[l];
    GemmArgs d;   // kv_down_proj (mla.py:146)
    d.A = ctx; d.rows = rows; d.K = H; d.lda = H; d.W = L.kv_down_w; d.N = RANK; d.out = craw; d.ldo = RANK;
    int rc = launch_gemm_simt(d, st);
    if (rc) return rc;
    GemmArgs u;   // kv_norm + kv_up_proj (mla.py:147-153)
    u.A = craw; u.rows = rows; u.K = RANK; u.lda = RANK; u.W = L.kv_up_w; u.N = 2 * H;
    u.out = kv_out + (int64_t)l * rows * 2 * H; u.ldo = 2 * H;
    u.pro = PRO_RMS; u.norm_w = L.kv_norm_w; u.norm_eps = 1e-6f;
    rc = launch_gemm_simt(u, st);
    if (rc) return rc;
  }
  return EDTTS_OK;
}

extern "C" int64_t edtts_decoder_workspace_bytes(int32_t B, int32_t T, int32_t S, int32_t precision) {
  const int64_t R = (int64_t)B * T;
  (void)S;
  if (precision == EDTTS_PREC_BF16) return tc_decoder_workspace_bytes(B, T, S);
  if (precision == EDTTS_PREC_TF32X3) return t3::t3_decoder_workspace_bytes(B, T, S);
  return align_up(R * H * 4, 256) * 2 + align_up(R * 3 * H * 4, 256);
}

static int decoder_step_fp32(const edtts_decoder_weights* w, const float* x_t, const float* mod, const float* kv,
                             const edtts_step_args* args, void* workspace, int32_t B, int32_t T, int32_t S,
                             cudaStream_t st) {
  const int64_t R = (int64_t)B * T;
  char* ws = reinterpret_cast<char*>(workspace);
  float* h = reinterpret_cast<float*>(ws);
  float* a = reinterpret_cast<float*>(ws + align_up(R * H * 4, 256));
  float* big = reinterpret_cast<float*>(ws + 2 * align_up(R * H * 4, 256));
  const float scale = 1.0f / sqrtf((float)HD);
  int rc;
  {  // h = in_proj(x_t) + pe[:T]   (decoder.py:96-97)
    GemmArgs g;
    g.A = x_t; g.rows = R; g.K = M; g.lda = M; g.W = w->in_proj_w; g.N = H; g.bias = w->in_proj_b;
    g.out = h; g.ldo = H; g.epi = EPI_PE; g.pe = w->pos_pe; g.pe_period = T;
    if ((rc = launch_gemm_simt(g, st))) return rc;
  }
  for (int l = 0; l < NL; ++l) {
    const edtts_layer_weights& L = w->layers[l];
    {  // qkv = attn.qkv(norm1(h, cond))   (transformer.py:143, attention.py:90)
      GemmArgs g;
      g.A = h; g.rows = R; g.K = H; g.lda = H; g.W = L.attn_qkv_w; g.N = 3 * H; g.out = big; g.ldo = 3 * H;
      g.pro = PRO_ADARMS; g.norm_w = L.norm1_norm_w; g.mod = mod + (int64_t)(2 * l) * 2 * H;
      g.mod_stride = 2 * NL * 2 * H; g.rows_per_batch = T;
      if ((rc = launch_gemm_simt(g, st))) return rc;
    }
    {  // banded self-attention (attention.py:94-111)
      AttnArgs at{big, 3 * H, big + H, big + 2 * H, 3 * H, a, H, T, T, WIN, scale};
      if ((rc = launch_attn_simt(at, B, st))) return rc;
    }
    {  // h += attn.proj(o)   (attention.py:123, transformer.py:146)
      GemmArgs g;
      g.A = a; g.rows = R; g.K = H; g.lda = H; g.W = L.attn_proj_w; g.N = H; g.bias = L.attn_proj_b;
      g.out = h; g.ldo = H; g.epi = EPI_RESID; g.resid = h;
      if ((rc = launch_gemm_simt(g, st))) return rc;
    }
    {  // q = q_proj(norm2(h))   (transformer.py:151, mla.py:139)
      GemmArgs g;
      g.A = h; g.rows = R; g.K = H; g.lda = H; g.W = L.q_proj_w; g.N = H; g.out = big; g.ldo = H;
      g.pro = PRO_RMS; g.norm_w = L.norm2_w;
      if ((rc = launch_gemm_simt(g, st))) return rc;
    }
    {  // full cross-attention over the S context tokens (mla.py:176-180)
      const float* kvl = kv + (int64_t)l * B * S * 2 * H;
      AttnArgs at{big, H, kvl, kvl + H, 2 * H, a, H, T, S, -1, scale};
      if ((rc = launch_attn_simt(at, B, st))) return rc;
    }
    {  // h += out_proj(o)   (mla.py:194)
      GemmArgs g;
      g.A = a; g.rows = R; g.K = H; g.lda = H; g.W = L.cross_out_w; g.N = H; g.out = h; g.ldo = H;
      g.epi = EPI_RESID; g.resid = h;
      if ((rc = launch_gemm_simt(g, st))) return rc;
    }
    {  // u = swiglu(ffn.net.0(norm3(h, cond)))   (transformer.py:155, :13-23)
      GemmArgs g;
      g.A = h; g.rows = R; g.K = H; g.lda = H; g.W = L.ffn0_w; g.N = FFN; g.bias = L.ffn0_b;
      g.out = big; g.ldo = FFN; g.epi = EPI_SWIGLU;
      g.pro = PRO_ADARMS; g.norm_w = L.norm3_norm_w; g.mod = mod + (int64_t)(2 * l + 1) * 2 * H;
      g.mod_stride = 2 * NL * 2 * H; g.rows_per_batch = T;
      if ((rc = launch_gemm_simt(g, st))) return rc;
    }
    {  // h += ffn.net.3(u)
      GemmArgs g;
      g.A = big; g.rows = R; g.K = FFN; g.lda = FFN; g.W = L.ffn3_w; g.N = H; g.bias = L.ffn3_b;
      g.out = h; g.ldo = H; g.epi = EPI_RESID; g.resid = h;
      if ((rc = launch_gemm_simt(g, st))) return rc;
    }
  }
  {  // eps = out_proj(final_norm(h)) + fused update   (decoder.py:108-109, schedule.py)
    GemmArgs g;
    g.A = h; g.rows = R; g.K = H; g.lda = H; g.W = w->out_proj_w; g.N = M; g.bias = w->out_proj_b;
    g.out = nullptr; g.ldo = M; g.epi = EPI_STEP; g.pro = PRO_LN; g.norm_w = w->final_norm_w;
    g.norm_b = w->final_norm_b; g.norm_eps = 1e-5f; g.rows_per_batch = T; g.x_t = x_t; g.step = *args;
    if ((rc = launch_gemm_simt(g, st))) return rc;
  }
  return EDTTS_OK;
}

extern "C" int edtts_decoder_step(const edtts_decoder_weights* w, const float* x_t, const float* mod, const float* kv,
                                  const edtts_step_args* args, void* workspace, int64_t workspace_bytes, int32_t B,
                                  int32_t T, int32_t S, int32_t precision, void* stream) {
  EDTTS_REQUIRE(w && x_t && mod && kv && args && workspace, EDTTS_EINVAL, "decoder_step: null argument");
  EDTTS_REQUIRE(B > 0 && T > 0 && S > 0, EDTTS_EINVAL, "decoder_step: B=%d T=%d S=%d", B, T, S);
  EDTTS_REQUIRE(T <= w->pos_rows, EDTTS_EINVAL, "decoder_step: T=%d exceeds the %d-row PE table", T, w->pos_rows);
  EDTTS_REQUIRE(workspace_bytes >= edtts_decoder_workspace_bytes(B, T, S, precision), EDTTS_ENOSPC,
                "decoder_step: workspace too small");
  switch (args->mode) {
    case EDTTS_STEP_EPS:
      EDTTS_REQUIRE(args->eps_out, EDTTS_EINVAL, "decoder_step: eps_out is null");
      break;
    case EDTTS_STEP_DDIM:
      EDTTS_REQUIRE(args->t && args->t_prev && args->alpha_bar && (args->x0_out || args->x_prev_out), EDTTS_EINVAL,
                    "decoder_step: DDIM needs t, t_prev, alpha_bar and an output");
      break;
    case EDTTS_STEP_DDPM:
      EDTTS_REQUIRE(args->t && args->alpha_bar && args->alphas && args->betas && args->posterior_var &&
                        args->noise && args->x_prev_out,
                    EDTTS_EINVAL, "decoder_step: DDPM needs t, tables, noise and x_prev_out");
      break;
    case EDTTS_STEP_DPM:
      EDTTS_REQUIRE(args->dpm_coef && args->x_prev_out && args->dpm_order >= 1 && args->dpm_order <= 3 &&
                        (args->dpm_order < 2 || args->dpm_hist1) && (args->dpm_order < 3 || args->dpm_hist2),
                    EDTTS_EINVAL, "decoder_step: DPM needs coefficients, x_prev_out and order-1 history tensors");
      break;
    default:
      set_error("decoder_step: unknown mode %d", args->mode);
      return EDTTS_EINVAL;
  }
  if (precision == EDTTS_PREC_FP32) return decoder_step_fp32(w, x_t, mod, kv, args, workspace, B, T, S, as_stream(stream));
  if (precision == EDTTS_PREC_BF16)
    return tc_decoder_step(w, x_t, mod, kv, args, workspace, B, T, S, as_stream(stream));
  if (precision == EDTTS_PREC_TF32X3) return t3::t3_decoder_step(w, x_t, mod, kv, args, workspace, B, T, S, as_stream(stream));
  set_error("decoder_step: unknown precision %d", precision);
  return EDTTS_EINVAL;
}
