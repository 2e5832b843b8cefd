// Fused transformer-block kernel of the bf16 tensor-core path: weight image packing and launcher.
#include "tc_path.cuh"
#include "tc_layer.cuh"
#include <stdlib.h>

namespace edtts {
namespace tc {

// ---- per-layer image: LY_NCHUNK weight chunks [20][160][8] bf16, then LC_COUNT fp32 constants -----------
constexpr int64_t LY_IMG_W_BYTES = (int64_t)LY_NCHUNK * LY_WCHUNK;
constexpr int64_t LY_IMG_BYTES = LY_IMG_W_BYTES + LC_COUNT * 4;          // 465,920 (128-byte multiple)
static_assert(LY_IMG_BYTES % 128 == 0, "layer image alignment");

int64_t tc_layer_packed_bytes() { return NL * LY_IMG_BYTES; }

// dst[c][n][j] = src[(row0 + n) * ld + col0 + 8 c + j]   (160 x 160 block of an nn.Linear weight)
__global__ void pack_chunk_kernel(const float* __restrict__ src, int ld, int row0, int col0,
                                  __nv_bfloat16* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 160 * 160) return;
  const int j = i & 7, n = (i >> 3) % 160, c = (i >> 3) / 160;
  dst[i] = __float2bfloat16_rn(src[(int64_t)(row0 + n) * ld + col0 + 8 * c + j]);
}

__global__ void pack_layer_consts_kernel(const edtts_layer_weights L, float* __restrict__ dst) {
  for (int i = threadIdx.x; i < LC_COUNT; i += blockDim.x) {
    float v;
    if (i < LC_N2W) v = L.attn_proj_b[i - LC_PROJ_B];
    else if (i < LC_N3W) v = L.norm2_w[i - LC_N2W];
    else if (i < LC_F0B) v = L.norm3_norm_w[i - LC_N3W];
    else if (i < LC_F3B) {
      // [half][x 160 | gate 160]: u column 160 half + c <- ffn.net.0 rows (160 half + c) and (320 + 160 half + c)
      const int k = i - LC_F0B, half = k / 320, r = k % 320;
      v = L.ffn0_b[r < 160 ? 160 * half + r : FFN + 160 * half + (r - 160)];
    } else v = L.ffn3_b[i - LC_F3B];
    dst[i] = v;
  }
}

int tc_layer_pack(const edtts_decoder_weights* w, void* dst, cudaStream_t st) {
  uint8_t* base = reinterpret_cast<uint8_t*>(dst);
  for (int l = 0; l < NL; ++l) {
    const edtts_layer_weights& L = w->layers[l];
    uint8_t* img = base + l * LY_IMG_BYTES;
    struct Src { const float* p; int ld, row0, col0; };
    const Src src[LY_NCHUNK] = {
        {L.attn_proj_w, H, 0, 0},   {L.q_proj_w, H, 0, 0},       {L.cross_out_w, H, 0, 0},
        {L.ffn0_w, H, 0, 0},        {L.ffn0_w, H, FFN, 0},       {L.ffn0_w, H, 160, 0},
        {L.ffn0_w, H, FFN + 160, 0}, {L.ffn3_w, FFN, 0, 0},      {L.ffn3_w, FFN, 0, 160}};
    for (int c = 0; c < LY_NCHUNK; ++c) {
      LaunchScope ls(KC_TC_MISC, st);
      pack_chunk_kernel<<<(160 * 160 + 255) / 256, 256, 0, st>>>(src[c].p, src[c].ld, src[c].row0, src[c].col0,
                                                                 reinterpret_cast<__nv_bfloat16*>(img + (int64_t)c * LY_WCHUNK));
      int rc = check_launch("pack_chunk");
      if (rc) return rc;
    }
    LaunchScope ls(KC_TC_MISC, st);
    pack_layer_consts_kernel<<<1, 256, 0, st>>>(L, reinterpret_cast<float*>(img + LY_IMG_W_BYTES));
    int rc = check_launch("pack_layer_consts");
    if (rc) return rc;
  }
  return EDTTS_OK;
}

int launch_tc_layer(const void* layer_img_base, int layer, float* h, const void* qkv, const void* kvx,
                    const float* mod3, int mod_stride, int B, int T, int S, int stop_phase, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(tc_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LY_SMEM) != cudaSuccess)
      return check_launch("tc_layer smem attribute");
    configured = true;
  }
  const uint8_t* img = reinterpret_cast<const uint8_t*>(layer_img_base) + (int64_t)layer * LY_IMG_BYTES;
  LayerArgs a;
  a.h = h;
  a.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv);
  a.kvx = reinterpret_cast<const __nv_bfloat16*>(kvx);
  a.wimg = reinterpret_cast<const __nv_bfloat16*>(img);
  a.consts = reinterpret_cast<const float*>(img + LY_IMG_W_BYTES);
  a.mod3 = mod3;
  a.mod_stride = mod_stride;
  a.R = (int64_t)B * T;
  a.RS = (int64_t)B * S;
  a.B = B; a.T = T; a.S = S;
  a.tiles_per_utt = (T + 127) / 128;
  a.scale_log2e = 1.4426950408889634f / sqrtf((float)HD);
  a.stop_phase = stop_phase;
  a.phase_clocks = nullptr;
  static long long* clk_buf = nullptr;
  static const bool want_clocks = getenv("EDTTS_LAYER_CLOCKS") != nullptr;
  if (want_clocks) {
    if (!clk_buf) cudaMalloc(&clk_buf, 148 * 24 * sizeof(long long));
    a.phase_clocks = clk_buf;
  }
  const int ntiles = B * a.tiles_per_utt;
  LaunchScope ls(KC_TC_LAYER, st);
  tc_layer_kernel<<<ntiles < 148 ? ntiles : 148, LY_THREADS, LY_SMEM, st>>>(a);
  if (want_clocks) {   // debug only: synchronous read-back of the per-phase cycle counters of CTA 0
    long long hc[24];
    cudaMemcpy(hc, clk_buf, sizeof(hc), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[tc_layer clocks] prologue %lld window %lld proj+n2 %lld q %lld cross %lld out+n3 %lld ffn %lld f3+store %lld\n",
            hc[0], hc[1], hc[2], hc[3], hc[4], hc[5], hc[6], hc[7]);
    for (int k = 1; k < 3; ++k)
      fprintf(stderr, "[tc_layer %s] other %lld waitS+load %lld max %lld exp+store %lld arriveP %lld - %lld head %lld waitO %lld\n",
              k == 1 ? "window" : "cross ", hc[8 * k], hc[8 * k + 1], hc[8 * k + 2], hc[8 * k + 3], hc[8 * k + 4], hc[8 * k + 5],
              hc[8 * k + 6], hc[8 * k + 7]);
  }
  return check_launch("tc_layer");
}

}  // namespace tc
}  // namespace edtts
