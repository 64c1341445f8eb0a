// i8gemm_sm100.cu -- int8 tensor-core projection (placeholder until the tcgen05 kernel lands)
#include "common.cuh"
namespace gb {
bool i8_available(gb200_ctx *) { return false; }
int i8_prepare(gb200_ctx *ctx) { return set_err(ctx, GB200_ERR_UNSUPPORTED, "int8 path not built"); }
int i8_project_bed(gb200_ctx *ctx, const unsigned char *, const int *, size_t, size_t, size_t, double *) {
  return set_err(ctx, GB200_ERR_UNSUPPORTED, "int8 path not built");
}
}  // namespace gb
