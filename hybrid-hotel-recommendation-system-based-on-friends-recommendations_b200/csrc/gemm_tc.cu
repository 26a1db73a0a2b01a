// tcgen05 tensor-core GEMM (placeholder until the kernel lands: every shape reports unsupported,
// so the dispatcher in linear.cu uses the CUDA-core path).
#include "kernels.cuh"

namespace dcnr {

bool gemm_tc_supported(int, bool, bool, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int) { return false; }

int launch_gemm_tc(int, const float *, int64_t, bool, const float *, int64_t, bool, float *, int64_t, int64_t, int64_t,
                   int64_t, int, const GemmEpilogue &, cudaStream_t) {
    set_error("tcgen05 GEMM not built");
    return DCNR_ERR_INVALID;
}

}  // namespace dcnr
