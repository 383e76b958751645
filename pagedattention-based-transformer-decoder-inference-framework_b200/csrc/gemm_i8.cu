// gemm_i8.cu -- placeholder until the tcgen05 kind::i8 kernel lands (same round).
#include "pa_common.cuh"

PA_API int pa_gemm_i8(const int8_t*, const int8_t*, int8_t*, int32_t*, int, int, int, int, float,
                      float, float, const float*, int, pa_stream_t) {
    return PA_ERR_UNSUPPORTED;
}
