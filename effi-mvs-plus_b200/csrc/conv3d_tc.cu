// placeholder until the tcgen05 path lands
#include "common.cuh"
namespace effimvs {
size_t costreg_bf16_workspace_bytes(int, int, int, int) { return 0; }
int costreg_bf16(const float*, const float* const*, const float* const*, int, int, int, int, void*, size_t, float*, cudaStream_t) {
    set_error("bf16 regularization not built"); return EFFIMVS_EUNSUPPORTED; }
size_t cost_up_bf16_workspace_bytes(int, int, int, int) { return 0; }
int cost_up_bf16(const float*, const float*, const float* const*, const float* const*, int, int, int, int, void*, size_t, float*, cudaStream_t) {
    set_error("bf16 regularization not built"); return EFFIMVS_EUNSUPPORTED; }
}
