// bf16 tcgen05 (tensor-core) path entry points, implemented in tc_path.cu.
#pragma once
#include "common.cuh"

namespace edtts {
int64_t tc_decoder_workspace_bytes(int32_t B, int32_t T, int32_t S);
int tc_decoder_step(const edtts_decoder_weights* w, const float* x_t, const float* mod, const float* kv,
                    const edtts_step_args* args, void* workspace, int32_t B, int32_t T, int32_t S,
                    cudaStream_t stream);
int tc_test_linear(const float* x, const float* w, const float* bias, float* y, int64_t rows, int32_t K, int32_t N,
                   int mode, cudaStream_t stream);
int tc_test_attention(const float* q, int q_stride, const float* k, const float* v, int kv_stride, float* o, int B,
                      int Tq, int Tk, int window, cudaStream_t stream);
// hidden state after `n_layers` blocks (the last one optionally stopped after phase `stop_phase`), test hook
int tc_test_hidden(const edtts_decoder_weights* w, const float* x_t, const float* mod, const float* kv, float* h_out,
                   void* workspace, int32_t B, int32_t T, int32_t S, int n_layers, int stop_phase, int fused,
                   cudaStream_t stream);
// cross-attention K | V of all blocks from the context embeddings (bf16 path): kv_out[l] = chunk-major [40][rows][8],
// k as bf16, v as f16; craw = scratch rows x 80 fp32
int tc_context_kv(const edtts_decoder_weights* w, const float* ctx, const int64_t* sem_idx, int S, float* craw, void* kv_out,
                  int64_t rows, cudaStream_t st);
namespace tc {
enum LyMode : int { LM_BLOCK = 0, LM_HEAD = 1 };
enum LyTail : int { LT_NONE = 0, LT_QKV = 1, LT_FINAL = 2 };
int64_t tc_layer_packed_bytes();
int tc_layer_pack(const edtts_decoder_weights* w, void* dst, cudaStream_t st);
// the head + `n_layers` blocks of a decoder evaluation on the fused kernel (tc_layer.cuh): one merged persistent launch
// (flags: tc_layer_flag_bytes, zeroed inside) or one launch per layer
int64_t tc_layer_flag_bytes(int B, int T);
int launch_tc_layers(const edtts_decoder_weights* w, const void* layer_img_base, int n_layers, float* hc, void* const qb[2],
                     const void* kv, const float* mod, const float* x_t, const edtts_step_args* step, int B, int T, int S,
                     int stop_phase, bool merged, void* flags, cudaStream_t st);
int unpack_hc(const float* hc, float* h, int64_t R, cudaStream_t st);
// fp32 row-major [R][lda] (first K columns) -> 16-bit chunk-major [K/8][R][8]: bf16, columns >= f16_from_col as f16
int pack_activation(const float* src, int lda, void* dst_chunk, int64_t R, int K, int f16_from_col, cudaStream_t st);
}
}  // namespace edtts
