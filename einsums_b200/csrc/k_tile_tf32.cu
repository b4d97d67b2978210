// k_tile_tf32.cu — instantiates the tcgen05 3xTF32 FP32 GEMM kernel.
#define EB200_TF32_KERNEL_IMPL
#include "gemm_tf32x3.cuh"
#include "launch_api.cuh"
namespace eb200 {
int k_launch_tf32x3(cudaStream_t s, const CUtensorMap &tmA, const CUtensorMap &tmB, const Tf32GemmArgs &a, dim3 grid, size_t smem) {
    int rc = set_smem_attr(gemm_tf32x3_tcgen05, smem);
    if (rc) return rc;
    gemm_tf32x3_tcgen05<<<grid, 192, smem, s>>>(tmA, tmB, a);
    EB_POST_LAUNCH("gemm_tf32x3_tcgen05");
    return EB200_OK;
}
} // namespace eb200
