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
// CTA-pair variant: clusters of two CTAs (cta_group::2), grid = an even number of CTAs
int k_launch_tf32x3_pair(cudaStream_t s, const CUtensorMap &tmA, const CUtensorMap &tmB, const Tf32GemmArgs &a, dim3 grid, size_t smem, bool a_mn, bool b_mn) {
    const bool multi = a.K > 512;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = dim3(320, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
#define EB_TF32_PAIR1(AM, BM_, MU)                                              \
    {                                                                           \
        auto kern = gemm_tf32x3_tcgen05<AM, BM_, MU, false, true>;              \
        int  rc = set_smem_attr(kern, smem);                                    \
        if (rc) return rc;                                                      \
        EB_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, a));                   \
    }
#define EB_TF32_PAIR(AM, BM_)                                                   \
    {                                                                           \
        if (multi) EB_TF32_PAIR1(AM, BM_, true) else EB_TF32_PAIR1(AM, BM_, false) \
    }
    if (a_mn && b_mn) EB_TF32_PAIR(true, true)
    else if (a_mn && !b_mn) EB_TF32_PAIR(true, false)
    else if (!a_mn && b_mn) EB_TF32_PAIR(false, true)
    else EB_TF32_PAIR(false, false)
#undef EB_TF32_PAIR
#undef EB_TF32_PAIR1
    EB_POST_LAUNCH("gemm_tf32x3_tcgen05<pair>");
    return EB200_OK;
}
int k_launch_tf32x3_generic(cudaStream_t s, const CUtensorMap &tmA, const CUtensorMap &tmB, const Tf32GemmArgs &a, dim3 grid, size_t smem, bool a_mn, bool b_mn,
                            bool pair) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = dim3(320, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
#define EB_TF32_GEN(AM, BM_)                                                    \
    {                                                                           \
        if (pair) {                                                             \
            auto kern = gemm_tf32x3_tcgen05<AM, BM_, true, true, true>;         \
            int  rc = set_smem_attr(kern, smem);                                \
            if (rc) return rc;                                                  \
            EB_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, a));               \
        } else {                                                                \
            auto kern = gemm_tf32x3_tcgen05<AM, BM_, true, true>;               \
            int  rc = set_smem_attr(kern, smem);                                \
            if (rc) return rc;                                                  \
            kern<<<grid, 320, smem, s>>>(tmA, tmB, a);                          \
        }                                                                       \
    }
    if (a_mn && b_mn) EB_TF32_GEN(true, true)
    else if (a_mn && !b_mn) EB_TF32_GEN(true, false)
    else if (!a_mn && b_mn) EB_TF32_GEN(false, true)
    else EB_TF32_GEN(false, false)
#undef EB_TF32_GEN
    EB_POST_LAUNCH(pair ? "gemm_tf32x3_tcgen05<generic,pair>" : "gemm_tf32x3_tcgen05<generic>");
    return EB200_OK;
}

// B' of the c64 path (contract.cu, try_run_c64_tcgen05): Bc = B gathered to [rows = batch x n][K] complex; B' holds, per batch, a plane
// of real-output rows (Bre, -Bim interleaved along k) and a plane of imaginary-output rows (Bim, Bre); rows are 2 * Kp floats long
__global__ void __launch_bounds__(256) c64_expand_b(const float2 *__restrict__ Bc, float *__restrict__ Bp, int64_t total, int64_t N, int64_t K, int64_t Kp) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = idx / K, k = idx - r * K, b = r / N, n = r - b * N;
        const float2  v = Bc[idx];
        float2       *re_row = reinterpret_cast<float2 *>(Bp + ((b * 2 + 0) * N + n) * 2 * Kp);
        float2       *im_row = reinterpret_cast<float2 *>(Bp + ((b * 2 + 1) * N + n) * 2 * Kp);
        re_row[k] = make_float2(v.x, -v.y);
        im_row[k] = make_float2(v.y, v.x);
    }
}
int k_launch_c64_expand(cudaStream_t s, const void *Bc, void *Bp, int64_t rows, int64_t N, int64_t K, int64_t Kp) {
    const int64_t total = rows * K;
    if (total <= 0) return EB200_OK;
    const unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 32);
    c64_expand_b<<<grid, 256, 0, s>>>((const float2 *)Bc, (float *)Bp, total, N, K, Kp);
    EB_POST_LAUNCH("c64_expand_b");
    return EB200_OK;
}
} // namespace eb200
