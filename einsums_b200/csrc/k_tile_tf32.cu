// k_tile_tf32.cu — instantiates the tcgen05 3xTF32 FP32 GEMM kernel for the four operand layouts.
#define EB200_TF32_KERNEL_IMPL
#include "gemm_tf32x3.cuh"
#include "launch_api.cuh"
namespace eb200 {
int k_launch_tf32x3(cudaStream_t s, const CUtensorMap &tmA, const CUtensorMap &tmB, const Tf32GemmArgs &a, dim3 grid, size_t smem, bool a_mn, bool b_mn) {
    const bool multi = a.K > 512; // more than one accumulation chunk per tile
#define EB_TF32_CASE1(AM, BM_, MU)                                              \
    {                                                                           \
        auto kern = gemm_tf32x3_tcgen05<AM, BM_, MU>;                           \
        int  rc = set_smem_attr(kern, smem);                                    \
        if (rc) return rc;                                                      \
        kern<<<grid, 320, smem, s>>>(tmA, tmB, a);                              \
    }
#define EB_TF32_CASE(AM, BM_)                                                   \
    {                                                                           \
        if (multi) EB_TF32_CASE1(AM, BM_, true) else EB_TF32_CASE1(AM, BM_, false) \
    }
    if (a_mn && b_mn) EB_TF32_CASE(true, true)
    else if (a_mn && !b_mn) EB_TF32_CASE(true, false)
    else if (!a_mn && b_mn) EB_TF32_CASE(false, true)
    else EB_TF32_CASE(false, false)
#undef EB_TF32_CASE
#undef EB_TF32_CASE1
    EB_POST_LAUNCH("gemm_tf32x3_tcgen05");
    return EB200_OK;
}
int k_launch_tf32x3_generic(cudaStream_t s, const CUtensorMap &tmA, const CUtensorMap &tmB, const Tf32GemmArgs &a, dim3 grid, size_t smem, bool a_mn, bool b_mn) {
#define EB_TF32_GEN(AM, BM_)                                                    \
    {                                                                           \
        auto kern = gemm_tf32x3_tcgen05<AM, BM_, true, true>;                   \
        int  rc = set_smem_attr(kern, smem);                                    \
        if (rc) return rc;                                                      \
        kern<<<grid, 320, smem, s>>>(tmA, tmB, a);                              \
    }
    if (a_mn && b_mn) EB_TF32_GEN(true, true)
    else if (a_mn && !b_mn) EB_TF32_GEN(true, false)
    else if (!a_mn && b_mn) EB_TF32_GEN(false, true)
    else EB_TF32_GEN(false, false)
#undef EB_TF32_GEN
    EB_POST_LAUNCH("gemm_tf32x3_tcgen05<generic>");
    return EB200_OK;
}
} // namespace eb200
