// k_tile_tma.cu — instantiates gemm_tma_f64 (plain products, 128x64) and contract_tma_f64 (multi-index / batched / split-K,
// 128x64 and 80x80) for the four operand layouts.
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

int k_launch_tma_z(cudaStream_t s, const CUtensorMap &tmA, const CUtensorMap &tmB, const TmaZGemmArgs &a, dim3 grid, size_t smem, bool ak, bool bk) {
#define EB_TMA_CASE(AK, BK_)                                                    \
    {                                                                           \
        auto kern = gemm_tma_z<AK, BK_>;                                        \
        int  rc = set_smem_attr(kern, smem);                                    \
        if (rc) return rc;                                                      \
        kern<<<grid, 256, smem, s>>>(tmA, tmB, a);                              \
    }
    if (ak && bk) EB_TMA_CASE(true, true)
    else if (ak && !bk) EB_TMA_CASE(true, false)
    else if (!ak && bk) EB_TMA_CASE(false, true)
    else EB_TMA_CASE(false, false)
#undef EB_TMA_CASE
    EB_POST_LAUNCH("gemm_tma_z");
    return EB200_OK;
}

namespace {
template <class CFG> int launch_generic(cudaStream_t s, const CUtensorMap &tmA, const CUtensorMap &tmB, const TmaGenArgs &a, dim3 grid, bool ak, bool bk) {
#define EB_TMA_CASE(AK, BK_)                                                    \
    {                                                                           \
        auto kern = contract_tma_f64<CFG, AK, BK_>;                             \
        int  rc = set_smem_attr(kern, CFG::SMEM);                               \
        if (rc) return rc;                                                      \
        kern<<<grid, CFG::NT, CFG::SMEM, s>>>(tmA, tmB, a);                     \
    }
    if (ak && bk) EB_TMA_CASE(true, true)
    else if (ak && !bk) EB_TMA_CASE(true, false)
    else if (!ak && bk) EB_TMA_CASE(false, true)
    else EB_TMA_CASE(false, false)
#undef EB_TMA_CASE
    EB_POST_LAUNCH("contract_tma_f64");
    return EB200_OK;
}
} // namespace

int k_launch_tma_generic(cudaStream_t s, const CUtensorMap &tmA, const CUtensorMap &tmB, const TmaGenArgs &a, dim3 grid, int cfg, bool ak, bool bk) {
    return cfg ? launch_generic<TmaTile80x80>(s, tmA, tmB, a, grid, ak, bk) : launch_generic<TmaTile128x64>(s, tmA, tmB, a, grid, ak, bk);
}
} // namespace eb200
