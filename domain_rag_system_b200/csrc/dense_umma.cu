// placeholder until the tcgen05 kernel lands
#include "common.cuh"
namespace drs {
bool dense_umma_supported(int, int64_t, int, int) { return false; }
int dense_umma_parts(int) { return 0; }
int launch_dense_umma(const float*, int, const void*, const float*, const uint8_t*, int64_t, int, int, uint64_t*, int,
                      cudaStream_t) {
  return fail(DRS_ERR_UNSUPPORTED, "tensor-core kernel not built");
}
}  // namespace drs
