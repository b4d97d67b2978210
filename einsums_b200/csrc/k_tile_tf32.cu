// k_tile_tf32.cu — instantiates the tcgen05 3xTF32 FP32 GEMM kernel for the four operand layouts.
#define EB200_TF32_KERNEL_IMPL
#include "gemm_tf32x3.cuh"
#include "launch_api.cuh"
namespace eb200 {
int k_launch_tf32x3(cudaStream_t s, const CUtensorMap &tmA, const CUtensorMap &tmB, const Tf32GemmArgs &a, dim3 grid, size_t smem, bool a_mn, bool b_mn) {
#define EB_TF32_CASE(AM, BM_)                                                   \
    {                                                                           \
        auto kern = gemm_tf32x3_tcgen05<AM, BM_>;                               \
        int  rc = set_smem_attr(kern, smem);                                    \
        if (rc) return rc;                                                      \
        kern<<<grid, 192, smem, s>>>(tmA, tmB, a);                              \
    }
    if (a_mn && b_mn) EB_TF32_CASE(true, true)
    else if (a_mn && !b_mn) EB_TF32_CASE(true, false)
    else if (!a_mn && b_mn) EB_TF32_CASE(false, true)
    else EB_TF32_CASE(false, false)
#undef EB_TF32_CASE
    EB_POST_LAUNCH("gemm_tf32x3_tcgen05");
    return EB200_OK;
}
} // namespace eb200
