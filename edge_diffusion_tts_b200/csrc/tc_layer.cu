// Fused transformer-block kernel of the bf16 tensor-core path: weight image packing and launcher.
#include "tc_path.cuh"
#include "tc_layer.cuh"
#include <stdlib.h>
#include <string.h>

namespace edtts {
namespace tc {

// ---- per-layer image: LY_NCHUNK weight chunks [20][160][8] bf16, then LC_COUNT fp32 constants -----------
constexpr int64_t LY_IMG_W_BYTES = (int64_t)LY_NCHUNK * LY_WCHUNK;
constexpr int64_t LY_IMG_BYTES = LY_IMG_W_BYTES + LC_COUNT * 4;          // 465,920 (128-byte multiple)
static_assert(LY_IMG_BYTES % 128 == 0, "layer image alignment");

// extras behind the NL block images: in_proj [10][160][8], out_proj [20][80][8], then per block attn.qkv as 3 chunks
constexpr int64_t LY_EXTRA_OFF = NL * LY_IMG_BYTES;
constexpr int64_t LY_EX_IN = 0, LY_EX_OUT = LY_WCHUNK / 2, LY_EX_QKV = LY_WCHUNK;
constexpr int64_t LY_EXTRA_BYTES = LY_EX_QKV + (int64_t)NL * 3 * LY_WCHUNK;
int64_t tc_layer_packed_bytes() { return LY_EXTRA_OFF + LY_EXTRA_BYTES; }

// dst[c][n][j] = src[(r(n)) * ld + col0 + 8 c + j], n < nrows, c < kcols / 8   (block of an nn.Linear weight);
// r(n) = row0 + n for n < split, row1 + (n - split) beyond (two row ranges of the weight in one chunk)
__global__ void pack_chunk_kernel(const float* __restrict__ src, int ld, int row0, int row1, int split, int col0, int nrows, int kcols,
                                  __nv_bfloat16* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows * kcols) return;
  const int j = i & 7, n = (i >> 3) % nrows, c = (i >> 3) / nrows;
  const int r = n < split ? row0 + n : row1 + (n - split);
  dst[i] = __float2bfloat16_rn(src[(int64_t)r * ld + col0 + 8 * c + j]);
}
static int pack_chunk2(const float* src, int ld, int row0, int row1, int split, int col0, int nrows, int kcols, void* dst,
                       cudaStream_t st) {
  LaunchScope ls(KC_TC_MISC, st);
  pack_chunk_kernel<<<(nrows * kcols + 255) / 256, 256, 0, st>>>(src, ld, row0, row1, split, col0, nrows, kcols,
                                                                reinterpret_cast<__nv_bfloat16*>(dst));
  return check_launch("pack_chunk");
}
static int pack_chunk(const float* src, int ld, int row0, int col0, int nrows, int kcols, void* dst, cudaStream_t st) {
  return pack_chunk2(src, ld, row0, 0, nrows, col0, nrows, kcols, dst, st);
}

__global__ void pack_layer_consts_kernel(const edtts_layer_weights L, float* __restrict__ dst) {
  for (int i = threadIdx.x; i < LC_COUNT; i += blockDim.x) {
    float v;
    if (i < LC_N2W) v = L.attn_proj_b[i - LC_PROJ_B];
    else if (i < LC_N3W) v = L.norm2_w[i - LC_N2W];
    else if (i < LC_F0B) v = L.norm3_norm_w[i - LC_N3W];
    else if (i < LC_F3B) {
      // [quarter][x 80 | gate 80]: u column 80 q + c <- ffn.net.0 rows (80 q + c) and (320 + 80 q + c)
      const int k = i - LC_F0B, q = k / 160, r = k % 160;
      v = L.ffn0_b[r < 80 ? 80 * q + r : FFN + 80 * q + (r - 80)];
    } else v = L.ffn3_b[i - LC_F3B];
    dst[i] = v;
  }
}

int tc_layer_pack(const edtts_decoder_weights* w, void* dst, cudaStream_t st) {
  uint8_t* base = reinterpret_cast<uint8_t*>(dst);
  for (int l = 0; l < NL; ++l) {
    const edtts_layer_weights& L = w->layers[l];
    uint8_t* img = base + l * LY_IMG_BYTES;
    // ffn0 quarter q: rows 80 q .. (x) then FFN + 80 q .. (gate)
    struct Src { const float* p; int ld, row0, row1, split, col0; };
    const Src src[LY_NCHUNK] = {
        {L.attn_proj_w, H, 0, 0, 160, 0},  {L.q_proj_w, H, 0, 0, 160, 0},       {L.cross_out_w, H, 0, 0, 160, 0},
        {L.ffn0_w, H, 0, FFN, 80, 0},      {L.ffn0_w, H, 80, FFN + 80, 80, 0},  {L.ffn0_w, H, 160, FFN + 160, 80, 0},
        {L.ffn0_w, H, 240, FFN + 240, 80, 0}, {L.ffn3_w, FFN, 0, 0, 160, 0},    {L.ffn3_w, FFN, 0, 0, 160, 160}};
    for (int c = 0; c < LY_NCHUNK; ++c) {
      int rc = pack_chunk2(src[c].p, src[c].ld, src[c].row0, src[c].row1, src[c].split, src[c].col0, 160, 160,
                           img + (int64_t)c * LY_WCHUNK, st);
      if (rc) return rc;
    }
    LaunchScope ls(KC_TC_MISC, st);
    pack_layer_consts_kernel<<<1, 256, 0, st>>>(L, reinterpret_cast<float*>(img + LY_IMG_W_BYTES));
    int rc = check_launch("pack_layer_consts");
    if (rc) return rc;
  }
  uint8_t* ex = base + LY_EXTRA_OFF;
  int rc = pack_chunk(w->in_proj_w, M, 0, 0, H, M, ex + LY_EX_IN, st);
  if (!rc) rc = pack_chunk(w->out_proj_w, H, 0, 0, M, H, ex + LY_EX_OUT, st);
  for (int l = 0; l < NL && !rc; ++l)
    for (int c = 0; c < 3 && !rc; ++c)
      rc = pack_chunk(w->layers[l].attn_qkv_w, H, 160 * c, 0, 160, 160, ex + LY_EX_QKV + (int64_t)(3 * l + c) * LY_WCHUNK, st);
  return rc;
}

// Arguments of one "layer launch" of the fused kernel.  layer = -1: head (in_proj + pe); otherwise transformer block `layer`.
// tail: LT_NONE / LT_QKV (norm1 + QKV of block layer+1, written to qkv_out) / LT_FINAL (final_norm + out_proj + step).
static int fill_layer_args(LayerArgs& a, const edtts_decoder_weights* w, const void* layer_img_base, int layer, int tail,
                           float* hc, const void* qkv_in, void* qkv_out, const void* kvx, const float* mod, const float* x_t,
                           const edtts_step_args* step, int B, int T, int S, int stop_phase) {
  const uint8_t* base = reinterpret_cast<const uint8_t*>(layer_img_base);
  const uint8_t* ex = base + LY_EXTRA_OFF;
  const int mod_stride = 2 * NL * 2 * H;                  // mod[b][2 l + {0: norm1, 1: norm3}][scale 160 | shift 160]
  memset(&a, 0, sizeof(a));
  a.mode = layer < 0 ? LM_HEAD : LM_BLOCK;
  a.tail = tail;
  a.hc = hc;
  a.mod_stride = mod_stride;
  a.R = (int64_t)B * T;
  a.RS = (int64_t)B * S;
  a.B = B; a.T = T; a.S = S;
  a.tiles_per_utt = (T + 127) / 128;
  a.scale_log2e = 1.4426950408889634f / sqrtf((float)HD);
  a.x_t = x_t;
  if (layer >= 0) {
    const uint8_t* img = base + (int64_t)layer * LY_IMG_BYTES;
    a.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv_in);
    a.kvx = reinterpret_cast<const __nv_bfloat16*>(kvx);
    a.wimg = reinterpret_cast<const __nv_bfloat16*>(img);
    a.consts = reinterpret_cast<const float*>(img + LY_IMG_W_BYTES);
    a.mod3 = mod + (int64_t)(2 * layer + 1) * 2 * H;
  } else {
    a.w_in = reinterpret_cast<const __nv_bfloat16*>(ex + LY_EX_IN);
    a.in_b = w->in_proj_b;
    a.pe = w->pos_pe;
    a.pe_cm = w->pos_pe_cm;
    a.pe_rows = w->pos_rows;
  }
  if (tail == LT_QKV) {
    const int nl = layer + 1;
    EDTTS_REQUIRE(nl < NL && qkv_out, EDTTS_EINVAL, "tc_layer: QKV tail after block %d", layer);
    a.w_qkv = reinterpret_cast<const __nv_bfloat16*>(ex + LY_EX_QKV + (int64_t)3 * nl * LY_WCHUNK);
    a.n1w = w->layers[nl].norm1_norm_w;
    a.mod1 = mod + (int64_t)(2 * nl) * 2 * H;
    a.qkv_out = reinterpret_cast<__nv_bfloat16*>(qkv_out);
  } else if (tail == LT_FINAL) {
    EDTTS_REQUIRE(step && x_t, EDTTS_EINVAL, "tc_layer: final tail needs the step arguments");
    a.w_out = reinterpret_cast<const __nv_bfloat16*>(ex + LY_EX_OUT);
    a.fn_w = w->final_norm_w;
    a.fn_b = w->final_norm_b;
    a.out_b = w->out_proj_b;
  }
  a.stop_phase = stop_phase;
  a.phase_clocks = nullptr;
  return EDTTS_OK;
}

int64_t tc_layer_flag_bytes(int B, int T) { return align_up((int64_t)LY_MAXL * B * ((T + 127) / 128) * 4, 256); }

// Resident CTAs of the fused kernel on the current device: SM count x occupancy (1: the kernel takes the whole shared memory
// and all 512 tensor-memory columns of an SM).  EDTTS_MAX_CTAS (development / tests) lowers it.
static int layer_resident_ctas() {
  static int cached[64] = {};
  const int dev = current_device();
  if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tc_layer_kernel<false>, LY_THREADS, LY_SMEM) != cudaSuccess || per_sm < 1) {
    (void)cudaGetLastError();
    per_sm = 1;
  }
  if (per_sm > 1) per_sm = 1;                                 // TMEM: one CTA allocates all 512 columns
  int n = sm_count() * per_sm;
  if (const char* e = getenv("EDTTS_MAX_CTAS")) {
    const int cap = atoi(e);
    if (cap > 0 && cap < n) n = cap;
  }
  if (dev >= 0 && dev < 64) cached[dev] = n;
  return n;
}

// The head and `n_layers` transformer blocks of one decoder evaluation.  merged: one persistent launch over all of them
// (items ordered layer-major, per-item dependency flags in `flags`, tc_layer_flag_bytes, zeroed here on the stream);
// otherwise one launch per layer.  qb[0] receives the head's q | k | v; block l reads qb[l & 1] and writes qb[(l + 1) & 1].
//
// Co-residency: in the merged launch CTAs wait on flags other CTAs of the same grid publish, so every CTA of the grid must be
// resident at once.  The launch is therefore COOPERATIVE (cudaLaunchAttributeCooperative): the driver starts it only when
// the whole grid fits on the SMs the context may use (MPS limits, green contexts, kernels of other streams holding SMs) and
// refuses it (cudaErrorCooperativeLaunchTooLarge) when it never can; a refusal falls back to one launch per layer, which
// has no inter-CTA waits.  The grid is min(items, SM count x occupancy) of the CURRENT device, never a literal.
int launch_tc_layers(const edtts_decoder_weights* w, const void* layer_img_base, int n_layers, float* hc, void* const qb[2],
                     const void* kv, const float* mod, const float* x_t, const edtts_step_args* step, int B, int T, int S,
                     int stop_phase, bool merged, void* flags, cudaStream_t st) {
  static PerDeviceOnce configured;                            // the attribute is per device, not per process
  if (configured.need()) {
    if (cudaFuncSetAttribute(tc_layer_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LY_SMEM) != cudaSuccess)
      return check_launch("tc_layer smem attribute");
#ifdef EDTTS_DEBUG_CLOCKS
    if (cudaFuncSetAttribute(tc_layer_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LY_SMEM) != cudaSuccess)
      return check_launch("tc_layer smem attribute");
#endif
    configured.set();
  }
  EDTTS_REQUIRE(n_layers >= 0 && n_layers + 1 <= LY_MAXL, EDTTS_EINVAL, "tc_layer: %d layers", n_layers);
  MegaArgs all;
  memset(&all, 0, sizeof(all));
  int rc = fill_layer_args(all.a[0], w, layer_img_base, -1, n_layers > 0 ? LT_QKV : LT_NONE, hc, nullptr, qb[0], nullptr, mod, x_t,
                           nullptr, B, T, S, 0);
  if (rc) return rc;
  for (int l = 0; l < n_layers; ++l) {
    const bool last = l == n_layers - 1;
    const int tail = !last ? LT_QKV : (step && stop_phase == 0 ? LT_FINAL : LT_NONE);
    const __nv_bfloat16* kvl = reinterpret_cast<const __nv_bfloat16*>(kv) + (int64_t)l * B * S * 2 * H;
    if ((rc = fill_layer_args(all.a[l + 1], w, layer_img_base, l, tail, hc, qb[l & 1], qb[(l + 1) & 1], kvl, mod, x_t, step, B, T, S,
                              last ? stop_phase : 0)))
      return rc;
  }
  const int n_launch = n_layers + 1;
  if (step) all.step = *step;
  const int ntiles = B * ((T + 127) / 128);
  const int resident = layer_resident_ctas();
#ifdef EDTTS_DEBUG_CLOCKS
  // development build only (-DEDTTS_DEBUG_CLOCKS): per-phase cycle counters of CTA 0, read back synchronously.  Allocates and
  // synchronises, which the ABI forbids -- never compiled into the shipped library.
  static long long* clk_buf = nullptr;
  static const bool want_clocks = getenv("EDTTS_LAYER_CLOCKS") != nullptr;
  if (want_clocks) {
    if (!clk_buf) cudaMalloc(&clk_buf, 1024 * 32 * sizeof(long long));
    for (int l = 0; l < n_launch; ++l) all.a[l].phase_clocks = clk_buf;
  }
#endif
#ifdef LY_TRACE
  static long long* tr_buf = nullptr;
  static int tr_done = 0;
  const bool tr_now = tr_done < 1 && B >= 64;
  if (tr_now) {
    if (!tr_buf) cudaMalloc(&tr_buf, 4 * 1024 * sizeof(long long));
    cudaMemsetAsync(tr_buf, 0, 4 * 1024 * sizeof(long long), st);
    for (int l = 0; l < n_launch; ++l) all.a[l].phase_clocks = tr_buf;
  }
#endif
  auto launch = [&](const MegaArgs& m, int items, bool cooperative) -> cudaError_t {
    LaunchScope ls(KC_TC_LAYER, st);
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof(lc));
    lc.gridDim = dim3((unsigned)(items < resident ? items : resident));
    lc.blockDim = dim3(LY_THREADS);
    lc.dynamicSmemBytes = LY_SMEM;
    lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    lc.attrs = at;
    lc.numAttrs = cooperative ? 1 : 0;
#ifdef EDTTS_DEBUG_CLOCKS
    if (want_clocks) return cudaLaunchKernelEx(&lc, tc_layer_kernel<true>, m);
#endif
    return cudaLaunchKernelEx(&lc, tc_layer_kernel<false>, m);
  };
  bool done_merged = false;
  if (merged && n_launch > 1) {
    EDTTS_REQUIRE(flags, EDTTS_EINVAL, "tc_layer: merged launch needs the flag buffer");
    if (cudaMemsetAsync(flags, 0, (size_t)n_launch * ntiles * 4, st) != cudaSuccess) return check_launch("tc_layer flags memset");
    all.n_launch = n_launch;
    all.done = reinterpret_cast<int*>(flags);
    const cudaError_t e = launch(all, n_launch * ntiles, true);
    if (e == cudaSuccess) done_merged = true;
    else if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported) (void)cudaGetLastError();   // -> per layer
    else return check_launch("tc_layer (merged)");
  }
  if (!done_merged) {
    for (int l = 0; l < n_launch; ++l) {
      MegaArgs one;
      memset(&one, 0, sizeof(one));
      one.a[0] = all.a[l];
      one.step = all.step;
      one.n_launch = 1;
      if (launch(one, ntiles, false) != cudaSuccess) return check_launch("tc_layer");
    }
  }
#ifdef LY_TRACE
  if (tr_now) {
    static long long hb[4 * 1024];
    cudaStreamSynchronize(st);
    cudaMemcpy(hb, tr_buf, sizeof(hb), cudaMemcpyDeviceToHost);
    ++tr_done;
    long long t0 = -1;
    for (int s_ = 0; s_ < 4; ++s_)
      for (long long e = 0; e < hb[s_ * 1024]; ++e)
        if (t0 < 0 || hb[s_ * 1024 + 3 + 2 * e] < t0) t0 = hb[s_ * 1024 + 3 + 2 * e];
    for (int s_ = 0; s_ < 4; ++s_)
      for (long long e = 0; e < hb[s_ * 1024]; ++e)
        fprintf(stderr, "TRACE %d %lld %lld\n", s_, hb[s_ * 1024 + 2 + 2 * e], hb[s_ * 1024 + 3 + 2 * e] - t0);
  }
#endif
#ifdef EDTTS_DEBUG_CLOCKS
  if (want_clocks) {   // synchronous read-back of the per-phase cycle counters of CTA 0 (summed over its items)
    long long hc_[32];
    cudaMemcpy(hc_, clk_buf, sizeof(hc_), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[tc_layer clocks] prologue %lld window %lld proj+n2 %lld q %lld cross %lld out+n3 %lld ffn %lld f3 %lld\n",
            hc_[0], hc_[1], hc_[2], hc_[3], hc_[4], hc_[5], hc_[6], hc_[7]);
    fprintf(stderr, "[tc_layer tail] norm+store_h %lld mma01 %lld issue2 %lld cvt_qk %lld wait2 %lld cvt_v+sync %lld\n", hc_[24], hc_[25],
            hc_[26], hc_[27], hc_[28], hc_[29]);
    for (int k = 1; k < 3; ++k)
      fprintf(stderr, "[tc_layer %s] other %lld waitS+load %lld max %lld exp+store %lld arriveP %lld - %lld head %lld waitO %lld\n",
              k == 1 ? "window" : "cross ", hc_[8 * k], hc_[8 * k + 1], hc_[8 * k + 2], hc_[8 * k + 3], hc_[8 * k + 4],
              hc_[8 * k + 5], hc_[8 * k + 6], hc_[8 * k + 7]);
  }
#endif
  return check_launch("tc_layer");
}

// chunk-major fp32 [40][R][4] -> row-major [R][160] (test hook)
__global__ void unpack_hc_kernel(const float* __restrict__ hc, float* __restrict__ h, int64_t R) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // one float4
  if (i >= R * 40) return;
  const int64_t r = i % R;
  const int c = (int)(i / R);
  *reinterpret_cast<float4*>(h + r * H + 4 * c) = *reinterpret_cast<const float4*>(hc + i * 4);
}
int unpack_hc(const float* hc, float* h, int64_t R, cudaStream_t st) {
  LaunchScope ls(KC_TC_MISC, st);
  unpack_hc_kernel<<<(unsigned)((R * 40 + 255) / 256), 256, 0, st>>>(hc, h, R);
  return check_launch("unpack_hc");
}

}  // namespace tc
}  // namespace edtts
