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
namespace tc {
// fp32 row-major [R][lda] (first K columns) -> bf16 chunk-major [K/8][R][8]
int pack_activation(const float* src, int lda, void* dst_chunk, int64_t R, int K, cudaStream_t st);
}
}  // namespace edtts
