// k_tile_tma.cu — instantiates gemm_tma_f64 for the four operand layouts.
#include "launch_api.cuh"
namespace eb200 {
int k_launch_tma(cudaStream_t s, const CUtensorMap &tmA, const CUtensorMap &tmB, const TmaGemmArgs &a, dim3 grid, size_t smem, bool ak, bool bk) {
#define EB_TMA_CASE(AK, BK_)                                                    \
    {                                                                           \
        auto kern = gemm_tma_f64<AK, BK_>;                                      \
        int  rc = set_smem_attr(kern, smem);                                    \
        if (rc) return rc;                                                      \
        kern<<<grid, 256, smem, s>>>(tmA, tmB, a);                              \
    }
    if (ak && bk) EB_TMA_CASE(true, true)
    else if (ak && !bk) EB_TMA_CASE(true, false)
    else if (!ak && bk) EB_TMA_CASE(false, true)
    else EB_TMA_CASE(false, false)
#undef EB_TMA_CASE
    EB_POST_LAUNCH("gemm_tma_f64");
    return EB200_OK;
}
} // namespace eb200
