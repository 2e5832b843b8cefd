// placeholder until the tcgen05 path lands
#include "tc_path.cuh"
namespace edtts {
int64_t tc_decoder_workspace_bytes(int32_t, int32_t, int32_t) { return 0; }
int tc_decoder_step(const edtts_decoder_weights*, const float*, const float*, const float*, const edtts_step_args*,
                    void*, int32_t, int32_t, int32_t, cudaStream_t) {
  set_error("bf16 tensor-core path not built");
  return EDTTS_ENOTSUP;
}
int tc_test_linear(const float*, const float*, const float*, float*, int64_t, int32_t, int32_t, cudaStream_t) {
  set_error("bf16 tensor-core path not built");
  return EDTTS_ENOTSUP;
}
}  // namespace edtts
extern "C" int64_t edtts_packed_bf16_bytes(void) { return 0; }
extern "C" int edtts_pack_weights_bf16(const edtts_decoder_weights*, void*, void*) {
  edtts::set_error("bf16 tensor-core path not built");
  return EDTTS_ENOTSUP;
}
