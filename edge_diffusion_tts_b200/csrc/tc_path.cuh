// bf16 tcgen05 (tensor-core) path entry points, implemented in tc_path.cu.
#pragma once
#include "common.cuh"

namespace edtts {
int64_t tc_decoder_workspace_bytes(int32_t B, int32_t T, int32_t S);
int tc_decoder_step(const edtts_decoder_weights* w, const float* x_t, const float* mod, const float* kv,
                    const edtts_step_args* args, void* workspace, int32_t B, int32_t T, int32_t S,
                    cudaStream_t stream);
int tc_test_linear(const float* x, const float* w, const float* bias, float* y, int64_t rows, int32_t K, int32_t N,
                   cudaStream_t stream);
}  // namespace edtts
