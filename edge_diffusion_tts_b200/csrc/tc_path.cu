// bf16 tensor-core path (EDTTS_PREC_BF16): weight packing, kernel launchers and the
// decoder step built from the tcgen05 kernels in tc_gemm.cuh / tc_attention.cuh.
#include "tc_path.cuh"
#include "tc_gemm.cuh"
#include "tc_attention.cuh"
#include <stdlib.h>

namespace edtts {
namespace tc {

// ---- packed weight image --------------------------------------------------------------
// Every matrix is stored per column block y as the UMMA operand image [K/8][N_CTA][8] bf16,
// so a CTA fetches its slab with one bulk copy.  ffn.net.0 is row-permuted so that block y
// holds the 80 "x" rows and the 80 matching "gate" rows of SwiGLU output columns 80y..80y+79.
struct LayerOff {
  int64_t qkv, attn_proj, q_proj, cross_out, ffn0, ffn0_bias, ffn3, kv_down, kv_up;
};
struct PackedOff {
  int64_t in_proj, out_proj;
  LayerOff layer[NL];
  int64_t total;
};

static PackedOff packed_offsets() {
  PackedOff p;
  int64_t o = 0;
  auto take = [&](int64_t bytes) {
    const int64_t at = o;
    o += align_up(bytes, 128);
    return at;
  };
  p.in_proj = take((int64_t)H * M * 2);
  p.out_proj = take((int64_t)M * H * 2);
  for (int l = 0; l < NL; ++l) {
    p.layer[l].qkv = take((int64_t)3 * H * H * 2);
    p.layer[l].attn_proj = take((int64_t)H * H * 2);
    p.layer[l].q_proj = take((int64_t)H * H * 2);
    p.layer[l].cross_out = take((int64_t)H * H * 2);
    p.layer[l].ffn0 = take((int64_t)2 * FFN * H * 2);
    p.layer[l].ffn0_bias = take((int64_t)2 * FFN * 4);
    p.layer[l].ffn3 = take((int64_t)H * FFN * 2);
    p.layer[l].kv_down = take((int64_t)RANK * H * 2);
    p.layer[l].kv_up = take((int64_t)2 * H * RANK * 2);
  }
  p.total = o;
  return p;
}

__device__ __forceinline__ int swiglu_row(int p, int n_cta, int ffn) {
  const int y = p / n_cta, n = p % n_cta, nu = n_cta / 2;
  return n < nu ? y * nu + n : ffn + y * nu + (n - nu);
}

// src fp32 [N_total][K] (nn.Linear weight) -> dst bf16 [N_total/N_CTA][K/8][N_CTA][8]
__global__ void pack_weight_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int n_total, int K,
                                   int n_cta, int swiglu_ffn) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_total * K) return;
  const int j = i & 7;
  const int n = (i >> 3) % n_cta;
  const int c = (i >> 3) / n_cta % (K / 8);
  const int y = (i >> 3) / n_cta / (K / 8);
  const int p = y * n_cta + n;
  const int row = swiglu_ffn ? swiglu_row(p, n_cta, swiglu_ffn) : p;
  dst[i] = __float2bfloat16_rn(src[(int64_t)row * K + c * 8 + j]);
}
__global__ void pack_bias_swiglu_kernel(const float* __restrict__ src, float* __restrict__ dst, int n_total, int n_cta,
                                        int ffn) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_total) dst[p] = src[swiglu_row(p, n_cta, ffn)];
}
// test helpers: fp32 row-major [R][K] <-> bf16 chunk-major [K/8][R][8]
// columns >= f16_from_col are stored as f16 (attention V operand), the rest as bf16
__global__ void pack_act_kernel(const float* __restrict__ src, int lda, __nv_bfloat16* __restrict__ dst, int64_t R,
                                int K, int f16_from_col) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * K) return;
  const int j = (int)(i & 7);
  const int64_t r = (i >> 3) % R;
  const int c = (int)((i >> 3) / R);
  const float v = src[r * lda + c * 8 + j];
  if (c * 8 >= f16_from_col) reinterpret_cast<__half*>(dst)[i] = __float2half_rn(v);
  else dst[i] = __float2bfloat16_rn(v);
}
__global__ void unpack_act_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int64_t R, int N) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * N) return;
  const int j = (int)(i & 7);
  const int64_t r = (i >> 3) % R;
  const int c = (int)((i >> 3) / R);
  dst[r * N + c * 8 + j] = __bfloat162float(src[i]);
}

static int pack_weight(const float* src, void* dst, int n_total, int K, int n_cta, int swiglu_ffn, cudaStream_t st) {
  LaunchScope ls(KC_TC_MISC, st);
  const int n = n_total * K;
  pack_weight_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n_total, K, n_cta,
                                                      swiglu_ffn);
  return check_launch("pack_weight");
}

// ---- GEMM launcher -----------------------------------------------------------------------
template <int K, int N_CTA>
static int launch_tc_gemm_t(const TcGemmArgs& g, int ny, cudaStream_t st) {
  using L = TcGemmSmem<K, N_CTA>;
  EDTTS_REQUIRE(!(g.amode == A_F32 && g.epi == TE_RESID), EDTTS_EINVAL, "tc_gemm: fp32 input with residual epilogue");
  static PerDeviceOnce configured;
  if (configured.need()) {
    const int max_smem = L::OFF_IO + (L::IN_BYTES > L::OUT_BYTES ? L::IN_BYTES : L::OUT_BYTES) + 64;
    if (cudaFuncSetAttribute(tc_gemm_kernel<K, N_CTA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             max_smem > 232448 ? 232448 : max_smem) != cudaSuccess)
      return check_launch("tc_gemm smem attribute");
    configured.set();
  }
  const int smem = L::total(g.amode, g.epi);
  EDTTS_REQUIRE(smem <= 232448, EDTTS_ENOTSUP, "tc_gemm<%d,%d>: %d B of shared memory", K, N_CTA, smem);
  const int64_t ntiles = (g.R + TILE_M - 1) / TILE_M;
  // persistent CTAs: as many per SM as shared memory and tensor memory allow (the small context projections are latency-
  // bound per tile; two or three resident CTAs overlap one tile's load / prologue with another's MMA / epilogue)
  int per_sm = 232448 / (smem + 1024);
  const int tmem_fit = 512 / (int)L::TMEM_COLS;
  if (per_sm > tmem_fit) per_sm = tmem_fit;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 3) per_sm = 3;
  int64_t nx = (int64_t)sm_count() * per_sm / ny;
  if (nx > ntiles) nx = ntiles;
  if (nx < 1) nx = 1;
  LaunchScope ls(KC_TC_GEMM, st);
  tc_gemm_kernel<K, N_CTA><<<dim3((unsigned)nx, ny), TC_THREADS, smem, st>>>(g, smem - L::OFF_IO - 64);
  return check_launch("tc_gemm");
}

int launch_tc_gemm(const TcGemmArgs& g, int K, int n_cta, int ny, cudaStream_t st) {
  if (K == 80 && n_cta == 160) return launch_tc_gemm_t<80, 160>(g, ny, st);
  if (K == 160 && n_cta == 240) return launch_tc_gemm_t<160, 240>(g, ny, st);
  if (K == 160 && n_cta == 160) return launch_tc_gemm_t<160, 160>(g, ny, st);
  if (K == 160 && n_cta == 320) return launch_tc_gemm_t<160, 320>(g, ny, st);
  if (K == 320 && n_cta == 80) return launch_tc_gemm_t<320, 80>(g, ny, st);
  if (K == 160 && n_cta == 80) return launch_tc_gemm_t<160, 80>(g, ny, st);
  set_error("tc_gemm: no kernel for K=%d N_CTA=%d", K, n_cta);
  return EDTTS_ENOTSUP;
}

int pack_activation(const float* src, int lda, void* dst_chunk, int64_t R, int K, int f16_from_col, cudaStream_t st) {
  LaunchScope ls(KC_TC_MISC, st);
  const int64_t n = R * K;
  pack_act_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, lda, reinterpret_cast<__nv_bfloat16*>(dst_chunk), R, K,
                                                               f16_from_col);
  return check_launch("pack_activation");
}

template <bool WINDOW>
static int launch_tc_attn(const TcAttnArgs& a, int B, cudaStream_t st) {
  static PerDeviceOnce configured;
  if (configured.need()) {
    if (cudaFuncSetAttribute(tc_attn_kernel<WINDOW>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::TOTAL) !=
        cudaSuccess)
      return check_launch("tc_attn smem attribute");
    cudaFuncSetAttribute(tc_attn_kernel<WINDOW>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    configured.set();
  }
  LaunchScope ls(WINDOW ? KC_TC_ATTN_WINDOW : KC_TC_ATTN_CROSS, st);
  tc_attn_kernel<WINDOW><<<dim3((a.Tq + AT_M - 1) / AT_M, NH, B), 128, AttnSmem::TOTAL, st>>>(a);
  return check_launch(WINDOW ? "tc_attn_window" : "tc_attn_cross");
}

static const float kScaleLog2e = 1.4426950408889634f / sqrtf((float)HD);

int launch_tc_attn_window(const __nv_bfloat16* qkv, __nv_bfloat16* o, int B, int T, cudaStream_t st) {
  TcAttnArgs a;
  a.q = qkv; a.q_rows = (int64_t)B * T; a.kv = qkv; a.kv_rows = (int64_t)B * T;
  a.q_chunk0 = 0; a.k_chunk0 = 20; a.v_chunk0 = 40; a.o = o; a.Tq = T; a.Tk = T; a.scale_log2e = kScaleLog2e;
  return launch_tc_attn<true>(a, B, st);
}

int launch_tc_attn_cross(const __nv_bfloat16* q, const void* kv_chunk, __nv_bfloat16* o, int B, int T, int S,
                         cudaStream_t st) {
  TcAttnArgs a;
  a.q = q; a.q_rows = (int64_t)B * T; a.kv = reinterpret_cast<const __nv_bfloat16*>(kv_chunk); a.kv_rows = (int64_t)B * S;
  a.q_chunk0 = 0; a.k_chunk0 = 0; a.v_chunk0 = 20; a.o = o; a.Tq = T; a.Tk = S; a.scale_log2e = kScaleLog2e;
  return launch_tc_attn<false>(a, B, st);
}

}  // namespace tc

using namespace tc;

// ---- workspace ------------------------------------------------------------------------------
struct TcWs {
  int64_t h, qkv, qkv2, q, o, u, flags, total;
};
static TcWs tc_ws_layout(int64_t R, int64_t flag_bytes = 0) {
  TcWs w;
  int64_t o = 0;
  auto take = [&](int64_t bytes) {
    const int64_t at = o;
    o += align_up(bytes, 256);
    return at;
  };
  w.h = take(R * H * 4);
  take(ATT_PAD_BYTES);                 // finite (zeroed) slack below qkv for the band halo of row 0
  w.qkv = take(R * 3 * H * 2);
  take(ATT_PAD_BYTES);                 // and above the last row
  w.qkv2 = take(R * 3 * H * 2);        // fused path: q | k | v of the next block (written while this one is read)
  take(ATT_PAD_BYTES);
  w.q = take(R * H * 2);
  w.o = take(R * H * 2);
  w.u = take(R * FFN * 2);
  w.flags = take(flag_bytes);          // fused path: per-item completion flags of the merged launch (last: offsets above are fixed)
  w.total = o;
  return w;
}

int64_t tc_decoder_workspace_bytes(int32_t B, int32_t T, int32_t S) {
  (void)S;
  return tc_ws_layout((int64_t)B * T, tc::tc_layer_flag_bytes(B, T)).total;
}

static bool env_flag(const char* name, bool dflt) {
  const char* v = getenv(name);
  if (!v || !*v) return dflt;
  return v[0] != '0';
}

// in_proj, then `n_layers` transformer blocks; fused = the whole block after the QKV GEMM is one launch
// (tc_layer.cuh), otherwise 7 launches.  stop_phase != 0 stops the LAST block early (test hook, fused only).
static int tc_run_layers(const edtts_decoder_weights* w, const float* x_t, const float* mod, const float* kv,
                         void* workspace, int32_t B, int32_t T, int32_t S, int n_layers, int stop_phase, bool fused,
                         const edtts_step_args* step, cudaStream_t st) {
  EDTTS_REQUIRE(w->packed_bf16, EDTTS_EINVAL, "decoder_step(bf16): weights.packed_bf16 is null; call "
                                              "edtts_pack_weights_bf16 first");
  const PackedOff po = packed_offsets();
  const uint8_t* pk = reinterpret_cast<const uint8_t*>(w->packed_bf16);
  auto img = [&](int64_t off) { return reinterpret_cast<const __nv_bfloat16*>(pk + off); };
  const int64_t R = (int64_t)B * T;
  const TcWs wl = tc_ws_layout(R, tc::tc_layer_flag_bytes(B, T));
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  float* h = reinterpret_cast<float*>(ws + wl.h);
  __nv_bfloat16* qkv = reinterpret_cast<__nv_bfloat16*>(ws + wl.qkv);
  __nv_bfloat16* qx = reinterpret_cast<__nv_bfloat16*>(ws + wl.q);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(ws + wl.o);
  __nv_bfloat16* u = reinterpret_cast<__nv_bfloat16*>(ws + wl.u);
  int rc;
  // the halo slack must hold finite numbers (masked keys still enter P.V as 0 * v)
  if (cudaMemsetAsync(ws + wl.qkv - ATT_PAD_BYTES, 0, ATT_PAD_BYTES, st) != cudaSuccess ||
      cudaMemsetAsync(ws + wl.qkv + R * 3 * H * 2, 0, ATT_PAD_BYTES, st) != cudaSuccess ||
      cudaMemsetAsync(ws + wl.qkv2 + R * 3 * H * 2, 0, ATT_PAD_BYTES, st) != cudaSuccess)
    return check_launch("tc workspace memset");
  if (fused) {
    // head (h = in_proj(x_t) + pe, q|k|v of block 0) and the blocks, each also producing the next block's q|k|v (or,
    // after the last block, final_norm + out_proj + the update rule when `step` is given): ONE persistent launch with
    // per-tile dependencies between the layers (EDTTS_MERGED_LAYERS=0: one launch per layer)
    const bool merged = env_flag("EDTTS_MERGED_LAYERS", true);     // read per call: tests compare both routes in one process
    void* qb[2] = {qkv, ws + wl.qkv2};
    return launch_tc_layers(w, pk + po.total, n_layers, h, qb, kv, mod, x_t, step, B, T, S, stop_phase, merged, ws + wl.flags, st);
  }
  {  // h = in_proj(x_t) + pe[:T]
    TcGemmArgs g;
    g.amode = A_F32; g.A_f32 = x_t; g.R = R; g.T = T; g.W_img = img(po.in_proj); g.bias = w->in_proj_b;
    g.epi = TE_PE; g.out_f32 = h; g.ldo = H; g.pe = w->pos_pe;
    if ((rc = launch_tc_gemm(g, M, H, 1, st))) return rc;
  }
  for (int l = 0; l < n_layers; ++l) {
    const edtts_layer_weights& L = w->layers[l];
    const LayerOff& lo = po.layer[l];
    const __nv_bfloat16* kvl = reinterpret_cast<const __nv_bfloat16*>(kv) + (int64_t)l * B * S * 2 * H;
    {  // q,k,v = attn.qkv(norm1(h, cond)) -> bf16 chunk-major [60][R][8]
      TcGemmArgs g;
      g.amode = A_F32; g.A_f32 = h; g.R = R; g.T = T; g.W_img = img(lo.qkv);
      g.pro = PRO_ADARMS; g.norm_w = L.norm1_norm_w; g.mod = mod + (int64_t)(2 * l) * 2 * H; g.mod_stride = 2 * NL * 2 * H;
      g.epi = TE_CHUNK; g.out_chunk = qkv; g.f16_from_chunk = 40;   // v is the f16 operand of P V
      if ((rc = launch_tc_gemm(g, H, 240, 2, st))) return rc;
    }
    if ((rc = launch_tc_attn_window(qkv, o, B, T, st))) return rc;
    {  // h += attn.proj(o) + bias
      TcGemmArgs g;
      g.amode = A_CHUNK; g.A_chunk = o; g.R = R; g.T = T; g.W_img = img(lo.attn_proj); g.bias = L.attn_proj_b;
      g.epi = TE_RESID; g.out_f32 = h; g.ldo = H;
      if ((rc = launch_tc_gemm(g, H, H, 1, st))) return rc;
    }
    if (l == n_layers - 1 && stop_phase == 1) break;
    {  // q = q_proj(norm2(h)) -> bf16 chunk-major [20][R][8]
      TcGemmArgs g;
      g.amode = A_F32; g.A_f32 = h; g.R = R; g.T = T; g.W_img = img(lo.q_proj);
      g.pro = PRO_RMS; g.norm_w = L.norm2_w; g.epi = TE_CHUNK; g.out_chunk = qx;
      if ((rc = launch_tc_gemm(g, H, H, 1, st))) return rc;
    }
    // bf16 path: kv is the chunk-major bf16 image written by edtts_context_prepare(precision=BF16)
    if ((rc = launch_tc_attn_cross(qx, kvl, o, B, T, S, st))) return rc;
    {  // h += out_proj(o)
      TcGemmArgs g;
      g.amode = A_CHUNK; g.A_chunk = o; g.R = R; g.T = T; g.W_img = img(lo.cross_out);
      g.epi = TE_RESID; g.out_f32 = h; g.ldo = H;
      if ((rc = launch_tc_gemm(g, H, H, 1, st))) return rc;
    }
    if (l == n_layers - 1 && stop_phase == 2) break;
    {  // u = swiglu(ffn.net.0(norm3(h, cond))) -> bf16 chunk-major [40][R][8]
      TcGemmArgs g;
      g.amode = A_F32; g.A_f32 = h; g.R = R; g.T = T; g.W_img = img(lo.ffn0);
      g.bias = reinterpret_cast<const float*>(pk + lo.ffn0_bias);
      g.pro = PRO_ADARMS; g.norm_w = L.norm3_norm_w; g.mod = mod + (int64_t)(2 * l + 1) * 2 * H;
      g.mod_stride = 2 * NL * 2 * H; g.epi = TE_SWIGLU; g.out_chunk = u;
      if ((rc = launch_tc_gemm(g, H, 2 * H, 2, st))) return rc;
    }
    {  // h += ffn.net.3(u) + bias
      TcGemmArgs g;
      g.amode = A_CHUNK; g.A_chunk = u; g.R = R; g.T = T; g.W_img = img(lo.ffn3); g.bias = L.ffn3_b;
      g.epi = TE_RESID; g.out_f32 = h; g.ldo = H;
      if ((rc = launch_tc_gemm(g, FFN, M, 2, st))) return rc;
    }
  }
  return EDTTS_OK;
}

int tc_decoder_step(const edtts_decoder_weights* w, const float* x_t, const float* mod, const float* kv,
                    const edtts_step_args* args, void* workspace, int32_t B, int32_t T, int32_t S, cudaStream_t st) {
  static const bool fused = env_flag("EDTTS_FUSED_LAYER", true);
  int rc = tc_run_layers(w, x_t, mod, kv, workspace, B, T, S, NL, 0, fused, args, st);
  if (rc || fused) return rc;
  const PackedOff po = packed_offsets();
  const uint8_t* pk = reinterpret_cast<const uint8_t*>(w->packed_bf16);
  const int64_t R = (int64_t)B * T;
  float* h = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + tc_ws_layout(R).h);
  {  // eps = out_proj(final_norm(h)) with the DDIM/DDPM update fused
    TcGemmArgs g;
    g.amode = A_F32; g.A_f32 = h; g.R = R; g.T = T; g.W_img = reinterpret_cast<const __nv_bfloat16*>(pk + po.out_proj);
    g.bias = w->out_proj_b;
    g.pro = PRO_LN; g.norm_w = w->final_norm_w; g.norm_b = w->final_norm_b; g.norm_eps = 1e-5f;
    g.epi = TE_STEP; g.ldo = M; g.x_t = x_t; g.step = *args;
    if ((rc = launch_tc_gemm(g, H, M, 1, st))) return rc;
  }
  return EDTTS_OK;
}

// ---- context K | V of all four layers in ONE launch (mla.py:144-153, step-invariant) -----------------------
// k | v = kv_up_proj(kv_norm(kv_down_proj(ctx))) per layer.  Persistent CTAs, blockIdx.y = layer: the layer's two weight
// images (kv_down 25.6 KB, kv_up 51.2 KB) are fetched once per CTA; per 128-token tile the bf16 chunk-major context rows
// arrive by bulk copies (double-buffered: the next tile streams in under this one), c = ctx Wd^T is a tcgen05 GEMM into
// tensor memory, the RMSNorm runs on the accumulator rows in registers (thread = token, the two column halves meet in
// shared memory) and writes the bf16 A operand of the second GEMM, k | v = n Wu^T (two 160-column accumulators) leaves as the
// chunk-major attention operand image (k bf16, v f16).  The fp32 intermediate never touches HBM; 1 launch instead of 8.
constexpr int CK_THREADS = 256;
constexpr int CK_SLAB = TILE_M * 16;
constexpr int CK_OFF_X = 0;                                  // two context tiles: 2 x 20 slabs
constexpr int CK_OFF_WD = CK_OFF_X + 2 * 20 * CK_SLAB;       // kv_down image [20][80][8]
constexpr int CK_OFF_WU = CK_OFF_WD + RANK * H * 2;          // kv_up image [2][10][160][8]
constexpr int CK_OFF_C = CK_OFF_WU + 2 * H * RANK * 2;       // normalised c: 10 slabs
constexpr int CK_OFF_RED = CK_OFF_C + 10 * CK_SLAB;          // [2][128] partial sums of squares
constexpr int CK_OFF_BAR = CK_OFF_RED + 2 * TILE_M * 4;
constexpr int CK_SMEM = CK_OFF_BAR + 64;

struct CtxKvArgs {
  const __nv_bfloat16* ctx_chunk;    // [20][R][8] bf16
  const uint8_t* packed;             // packed bf16 weight image
  int64_t off_down[NL], off_up[NL];
  const float* norm_w[NL];           // kv_norm.weight [80]
  __nv_bfloat16* kv_out;             // [NL][40][R][8]: k chunks 0..19 bf16, v chunks 20..39 f16
  int64_t R;
};

__global__ void __launch_bounds__(CK_THREADS, 1) ctx_kv_kernel(const CtxKvArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sX = smem + CK_OFF_X;
  uint8_t* sWd = smem + CK_OFF_WD;
  uint8_t* sWu = smem + CK_OFF_WU;
  uint8_t* sC = smem + CK_OFF_C;
  float* sRed = reinterpret_cast<float*>(smem + CK_OFF_RED);
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + CK_OFF_BAR);
  uint64_t* bar_x = bar_w + 1;                                // [2]
  uint64_t* bar_mma = bar_w + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int l = blockIdx.y;
  const int64_t ntiles = (a.R + TILE_M - 1) / TILE_M;

  for (int i = tid * 16; i < CK_OFF_WD; i += CK_THREADS * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);   // finite rows
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_x, 1);
    mbar_init(bar_x + 1, 1);
    mbar_init(bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t TM_C = 0, TM_KV = 128;                   // accumulators: c (80 columns), k | v (320 columns)

  auto issue_x = [&](int64_t tile, int buf) {                 // thread 0
    const int64_t row0 = tile * TILE_M;
    const uint32_t valid = (uint32_t)min((int64_t)TILE_M, a.R - row0);
    mbar_expect_tx(bar_x + buf, 20 * valid * 16);
#pragma unroll 1
    for (int c = 0; c < 20; ++c)
      bulk_g2s(sX + (buf * 20 + c) * CK_SLAB, a.ctx_chunk + ((int64_t)c * a.R + row0) * 8, valid * 16, bar_x + buf);
  };
  if (tid == 0) {
    mbar_expect_tx(bar_w, RANK * H * 2 + 2 * H * RANK * 2);
    bulk_g2s(sWd, a.packed + a.off_down[l], RANK * H * 2, bar_w);
    bulk_g2s(sWu, a.packed + a.off_up[l], 2 * H * RANK * 2, bar_w);
    if (blockIdx.x < ntiles) issue_x(blockIdx.x, 0);
  }
  const int lq = warp & 3, half = warp >> 2;
  const int r = lq * 32 + lane;
  const uint32_t trow = tmem + ((uint32_t)(lq * 32) << 16);
  float nw[40];
#pragma unroll
  for (int j = 0; j < 40; ++j) nw[j] = a.norm_w[l][40 * half + j];
  __nv_bfloat16* kvl = a.kv_out + (int64_t)l * a.R * 2 * H;
  uint32_t ph_m = 0;
  uint32_t it = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    const int64_t row0 = tile * TILE_M;
    if (tid == 0) {
      if (tile + gridDim.x < ntiles) issue_x(tile + gridDim.x, buf ^ 1);      // its last reader (tile it - 1's GEMM) has retired
      if (it == 0) mbar_wait(bar_w, 0);
      mbar_wait(bar_x + buf, (it >> 1) & 1);
      tc_fence_after();
      constexpr uint32_t ID_C = make_idesc(TILE_M, RANK);
#pragma unroll
      for (int ks = 0; ks < H / 16; ++ks)
        umma_bf16(tmem + TM_C, make_desc(smem_u32(sX) + (buf * 20 + 2 * ks) * CK_SLAB, CK_SLAB, 128),
                  make_desc(smem_u32(sWd) + ks * 2 * RANK * 16, RANK * 16, 128), ID_C, ks > 0);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, ph_m);
    ph_m ^= 1;
    tc_fence_after();
    {  // kv_norm on the accumulator row: this thread's 40 of the 80 columns
      float v[40];
      tmem_ld32(trow + TM_C + 40 * half, v);
      tmem_ld8(trow + TM_C + 40 * half + 32, v + 32);
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < 40; ++j) ss = fmaf(v[j], v[j], ss);
      sRed[half * TILE_M + r] = ss;
      __syncthreads();
      const float rstd = 1.0f / sqrtf((sRed[r] + sRed[TILE_M + r]) / (float)RANK + 1e-6f);
#pragma unroll
      for (int g = 0; g < 5; ++g) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[8 * g + j] * rstd) * nw[8 * g + j];
        *reinterpret_cast<uint4*>(sC + (5 * half + g) * CK_SLAB + r * 16) = pack_bf16x8(o);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t ID_KV = make_idesc(TILE_M, H);
#pragma unroll
      for (int y = 0; y < 2; ++y)
#pragma unroll
        for (int ks = 0; ks < RANK / 16; ++ks)
          umma_bf16(tmem + TM_KV + y * H, make_desc(smem_u32(sC) + ks * 2 * CK_SLAB, CK_SLAB, 128),
                    make_desc(smem_u32(sWu) + y * (H * RANK * 2) + ks * 2 * H * 16, H * 16, 128), ID_KV, ks > 0);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, ph_m);
    ph_m ^= 1;
    tc_fence_after();
    {  // k (half 0, bf16) | v (half 1, f16) -> chunk-major operand image
      const int64_t row = row0 + r;
#pragma unroll 1
      for (int c0 = 0; c0 < H; c0 += 16) {
        float v[16];
        tmem_ld16(trow + TM_KV + half * H + c0, v);
        if (row < a.R) {
          __nv_bfloat16* o = kvl + ((int64_t)((half * H + c0) >> 3) * a.R + row) * 8;
          *reinterpret_cast<uint4*>(o) = half ? pack_f16x8(v) : pack_bf16x8(v);
          *reinterpret_cast<uint4*>(o + a.R * 8) = half ? pack_f16x8(v + 8) : pack_bf16x8(v + 8);
        }
      }
    }
    tc_fence_before();
    __syncthreads();                                           // accumulators and sC are free for the next tile
  }
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// ctx[b, s, :] = token_emb[sem_idx[b, s]] + pe[s] (decoder.py:88,93) written straight as the bf16 chunk-major operand image
// [20][rows][8]: a thread = (row, 8-column group); the fp32 context rows never exist
__global__ void embed_ctx_chunk_kernel(const float* __restrict__ emb, const float* __restrict__ pe, const int64_t* __restrict__ idx,
                                       __nv_bfloat16* __restrict__ out, int64_t rows, int S, int codebook) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * (H / 8)) return;
  const int64_t r = i % rows;
  const int c = (int)(i / rows);
  int64_t tok = idx[r];
  tok = tok < 0 ? 0 : (tok >= codebook ? codebook - 1 : tok);
  const float4* e = reinterpret_cast<const float4*>(emb + tok * H + 8 * c);
  const float4* p = reinterpret_cast<const float4*>(pe + (r % S) * H + 8 * c);
  const float4 e0 = e[0], e1 = e[1], p0 = p[0], p1 = p[1];
  const float v[8] = {e0.x + p0.x, e0.y + p0.y, e0.z + p0.z, e0.w + p0.w, e1.x + p1.x, e1.y + p1.y, e1.z + p1.z, e1.w + p1.w};
  *reinterpret_cast<uint4*>(out + i * 8) = pack_bf16x8(v);
}

// ctx: fp32 context rows [rows][160] (packed to the operand image here), or -- sem_idx given -- built directly from the
// token embedding (ctx may then be null)
int tc_context_kv(const edtts_decoder_weights* w, const float* ctx, const int64_t* sem_idx, int S, float* craw, void* kv_out,
                  int64_t rows, cudaStream_t st) {
  EDTTS_REQUIRE(w->packed_bf16, EDTTS_EINVAL, "context_prepare(bf16): weights.packed_bf16 is null");
  const PackedOff po = packed_offsets();
  const uint8_t* pk = reinterpret_cast<const uint8_t*>(w->packed_bf16);
  // the context rows once as a bf16 chunk-major operand image (read through the TMA engine, no fp32 staging / conversion
  // prologue per layer); it lives in the third region of the context workspace
  __nv_bfloat16* ctx_chunk = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(craw) + align_up(rows * RANK * 4, 256));
  if (sem_idx) {
    const int64_t n = rows * (H / 8);
    LaunchScope ls(KC_EMBED, st);
    embed_ctx_chunk_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w->token_emb, w->ctx_pe, sem_idx, ctx_chunk, rows, S, w->codebook_size);
    if (int rc0 = check_launch("embed_ctx_chunk")) return rc0;
  } else if (int rc0 = pack_activation(ctx, H, ctx_chunk, rows, H, 1 << 30, st)) {
    return rc0;
  }
  if (env_flag("EDTTS_CTX_FUSED", true)) {
    CtxKvArgs a;
    a.ctx_chunk = ctx_chunk; a.packed = pk; a.kv_out = reinterpret_cast<__nv_bfloat16*>(kv_out); a.R = rows;
    for (int l = 0; l < NL; ++l) {
      a.off_down[l] = po.layer[l].kv_down;
      a.off_up[l] = po.layer[l].kv_up;
      a.norm_w[l] = w->layers[l].kv_norm_w;
    }
    static PerDeviceOnce configured;
    if (configured.need()) {
      if (cudaFuncSetAttribute(ctx_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CK_SMEM) != cudaSuccess)
        return check_launch("ctx_kv smem attribute");
      configured.set();
    }
    const int64_t ntiles = (rows + TILE_M - 1) / TILE_M;
    int64_t nx = sm_count() / NL;
    if (nx < 1) nx = 1;
    if (nx > ntiles) nx = ntiles;
    LaunchScope ls(KC_TC_GEMM, st);
    ctx_kv_kernel<<<dim3((unsigned)nx, NL), CK_THREADS, CK_SMEM, st>>>(a);
    return check_launch("ctx_kv");
  }
  for (int l = 0; l < NL; ++l) {
    const LayerOff& lo = po.layer[l];
    int rc;
    {  // c = kv_down_proj(ctx)   (mla.py:146)
      TcGemmArgs g;
      g.amode = A_CHUNK; g.A_chunk = ctx_chunk; g.R = rows; g.T = (int)rows; g.W_img = reinterpret_cast<const __nv_bfloat16*>(pk + lo.kv_down);
      g.epi = TE_F32; g.out_f32 = craw; g.ldo = RANK;
      if ((rc = launch_tc_gemm(g, H, RANK, 1, st))) return rc;
    }
    {  // k | v = kv_up_proj(kv_norm(c)) -> chunk-major, k (chunks 0..19) bf16, v (20..39) f16   (mla.py:147-153)
      TcGemmArgs g;
      g.amode = A_F32; g.A_f32 = craw; g.R = rows; g.T = (int)rows; g.W_img = reinterpret_cast<const __nv_bfloat16*>(pk + lo.kv_up);
      g.pro = PRO_RMS; g.norm_w = w->layers[l].kv_norm_w; g.norm_eps = 1e-6f;
      g.epi = TE_CHUNK; g.out_chunk = reinterpret_cast<__nv_bfloat16*>(kv_out) + (int64_t)l * rows * 2 * H; g.f16_from_chunk = 20;
      if ((rc = launch_tc_gemm(g, RANK, H, 2, st))) return rc;
    }
  }
  return EDTTS_OK;
}

int tc_test_hidden(const edtts_decoder_weights* w, const float* x_t, const float* mod, const float* kv, float* h_out,
                   void* workspace, int32_t B, int32_t T, int32_t S, int n_layers, int stop_phase, int fused,
                   cudaStream_t st) {
  EDTTS_REQUIRE(n_layers >= 0 && n_layers <= NL && stop_phase >= 0 && stop_phase <= 2, EDTTS_EINVAL,
                "test_hidden: n_layers=%d stop_phase=%d", n_layers, stop_phase);
  int rc = tc_run_layers(w, x_t, mod, kv, workspace, B, T, S, n_layers, stop_phase, fused != 0, nullptr, st);
  if (rc) return rc;
  const int64_t R = (int64_t)B * T;
  const float* h = reinterpret_cast<const float*>(reinterpret_cast<uint8_t*>(workspace) + tc_ws_layout(R).h);
  if (fused) return unpack_hc(h, h_out, R, st);           // the fused path keeps h chunk-major
  if (cudaMemcpyAsync(h_out, h, R * H * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess) return check_launch("test_hidden copy");
  return EDTTS_OK;
}

// ---- single-kernel test hook ---------------------------------------------------------------
// mode 1: fp32 A (prologue path), fp32 out; mode 2: bf16 chunk-major A (bulk-copy path), fp32 out;
// mode 3: fp32 A, bf16 chunk-major out (TE_CHUNK) unpacked back to fp32.  Allocates temporaries
// (test hook only -- never used on the sampling path).
int tc_test_linear(const float* x, const float* w, const float* bias, float* y, int64_t rows, int32_t K, int32_t N,
                   int mode, cudaStream_t st) {
  int n_cta = 0;
  if (K == 80 && N == 160) n_cta = 160;
  if (K == 160) n_cta = (N == 80) ? 80 : (N == 160) ? 160 : (N == 480) ? 240 : (N == 640) ? 320 : 0;
  if (K == 320 && N == 160) n_cta = 80;
  EDTTS_REQUIRE(n_cta > 0, EDTTS_ENOTSUP, "tc_test_linear: K=%d N=%d has no tcgen05 kernel", K, N);
  const int ny = N / n_cta;
  __nv_bfloat16 *wimg = nullptr, *achunk = nullptr, *ochunk = nullptr;
  if (cudaMalloc(&wimg, (size_t)N * K * 2) != cudaSuccess) return check_launch("cudaMalloc");
  int rc = pack_weight(w, wimg, N, K, n_cta, 0, st);
  TcGemmArgs g;
  g.R = rows; g.T = (int)rows; g.W_img = wimg; g.bias = bias; g.out_f32 = y; g.ldo = N; g.epi = TE_F32;
  g.amode = A_F32; g.A_f32 = x;
  if (!rc && mode == 2) {
    if (cudaMalloc(&achunk, (size_t)rows * K * 2) != cudaSuccess) rc = check_launch("cudaMalloc");
    if (!rc) {
      const int64_t n = rows * K;
      pack_act_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, K, achunk, rows, K, 1 << 30);
      rc = check_launch("pack_act");
      g.amode = A_CHUNK; g.A_chunk = achunk;
    }
  }
  if (!rc && mode == 3) {
    if (cudaMalloc(&ochunk, (size_t)rows * N * 2) != cudaSuccess) rc = check_launch("cudaMalloc");
    g.epi = TE_CHUNK; g.out_chunk = ochunk;
  }
  if (!rc) rc = launch_tc_gemm(g, K, n_cta, ny, st);
  if (!rc && mode == 3) {
    const int64_t n = rows * N;
    unpack_act_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ochunk, y, rows, N);
    rc = check_launch("unpack_act");
  }
  const cudaError_t e = cudaStreamSynchronize(st);
  if (!rc && e != cudaSuccess) {
    set_error("tc_test_linear: %s", cudaGetErrorString(e));
    rc = EDTTS_ECUDA;
  }
  cudaFree(wimg);
  cudaFree(achunk);
  cudaFree(ochunk);
  return rc;
}

// fp32 row-major q [B*Tq][q_stride], k / v [B*Tk][kv_stride] -> o [B*Tq][160], through the tcgen05 kernels.
int tc_test_attention(const float* q, int q_stride, const float* k, const float* v, int kv_stride, float* o, int B,
                      int Tq, int Tk, int window, cudaStream_t st) {
  const int64_t Rq = (int64_t)B * Tq, Rk = (int64_t)B * Tk;
  EDTTS_REQUIRE(window < 0 || (window == WIN && Tq == Tk), EDTTS_ENOTSUP, "tc_test_attention: window must be %d", WIN);
  uint8_t* buf = nullptr;
  const size_t q_bytes = (size_t)Rq * H * 2, kv_bytes = (size_t)Rk * 2 * H * 2, o_bytes = (size_t)Rq * H * 2;
  const size_t total = ATT_PAD_BYTES + q_bytes + kv_bytes + ATT_PAD_BYTES + o_bytes;
  if (cudaMalloc(&buf, total) != cudaSuccess) return check_launch("cudaMalloc");
  cudaMemsetAsync(buf, 0, total, st);
  __nv_bfloat16* qc = reinterpret_cast<__nv_bfloat16*>(buf + ATT_PAD_BYTES);
  __nv_bfloat16* kc = reinterpret_cast<__nv_bfloat16*>(buf + ATT_PAD_BYTES + q_bytes);
  __nv_bfloat16* vc = kc + Rk * H;
  __nv_bfloat16* oc = reinterpret_cast<__nv_bfloat16*>(buf + ATT_PAD_BYTES + q_bytes + kv_bytes + ATT_PAD_BYTES);
  int rc = pack_activation(q, q_stride, qc, Rq, H, 1 << 30, st);
  if (!rc) rc = pack_activation(k, kv_stride, kc, Rk, H, 1 << 30, st);
  if (!rc) rc = pack_activation(v, kv_stride, vc, Rk, H, 0, st);
  if (!rc) rc = window >= 0 ? launch_tc_attn_window(qc, oc, B, Tq, st) : launch_tc_attn_cross(qc, kc, oc, B, Tq, Tk, st);
  if (!rc) {
    const int64_t n = Rq * H;
    unpack_act_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(oc, o, Rq, H);
    rc = check_launch("unpack_act");
  }
  const cudaError_t e = cudaStreamSynchronize(st);
  if (!rc && e != cudaSuccess) {
    set_error("tc_test_attention: %s", cudaGetErrorString(e));
    rc = EDTTS_ECUDA;
  }
  cudaFree(buf);
  return rc;
}

}  // namespace edtts

using namespace edtts;

extern "C" int64_t edtts_packed_bf16_bytes(void) { return tc::packed_offsets().total + tc::tc_layer_packed_bytes(); }

extern "C" int edtts_pack_weights_bf16(const edtts_decoder_weights* w, void* packed_out, void* stream) {
  EDTTS_REQUIRE(w && packed_out, EDTTS_EINVAL, "pack_weights_bf16: null argument");
  const tc::PackedOff po = tc::packed_offsets();
  uint8_t* pk = reinterpret_cast<uint8_t*>(packed_out);
  cudaStream_t st = as_stream(stream);
  int rc;
  if ((rc = tc::pack_weight(w->in_proj_w, pk + po.in_proj, H, M, H, 0, st))) return rc;
  if ((rc = tc::pack_weight(w->out_proj_w, pk + po.out_proj, M, H, M, 0, st))) return rc;
  for (int l = 0; l < NL; ++l) {
    const edtts_layer_weights& L = w->layers[l];
    const tc::LayerOff& lo = po.layer[l];
    if ((rc = tc::pack_weight(L.attn_qkv_w, pk + lo.qkv, 3 * H, H, 240, 0, st))) return rc;
    if ((rc = tc::pack_weight(L.attn_proj_w, pk + lo.attn_proj, H, H, H, 0, st))) return rc;
    if ((rc = tc::pack_weight(L.q_proj_w, pk + lo.q_proj, H, H, H, 0, st))) return rc;
    if ((rc = tc::pack_weight(L.cross_out_w, pk + lo.cross_out, H, H, H, 0, st))) return rc;
    if ((rc = tc::pack_weight(L.ffn0_w, pk + lo.ffn0, 2 * FFN, H, 2 * H, FFN, st))) return rc;
    {
      LaunchScope ls(KC_TC_MISC, st);
      tc::pack_bias_swiglu_kernel<<<(2 * FFN + 255) / 256, 256, 0, st>>>(
          L.ffn0_b, reinterpret_cast<float*>(pk + lo.ffn0_bias), 2 * FFN, 2 * H, FFN);
      if ((rc = check_launch("pack_bias"))) return rc;
    }
    if ((rc = tc::pack_weight(L.ffn3_w, pk + lo.ffn3, H, FFN, M, 0, st))) return rc;
    if ((rc = tc::pack_weight(L.kv_down_w, pk + lo.kv_down, RANK, H, RANK, 0, st))) return rc;
    if ((rc = tc::pack_weight(L.kv_up_w, pk + lo.kv_up, 2 * H, RANK, H, 0, st))) return rc;
  }
  // images of the fused transformer-block kernel follow the per-GEMM images
  return tc::tc_layer_pack(w, pk + po.total, st);
}
