// k_tile_mma.cu — instantiates tile_mma_f64 (cp.async gather kernel, 128x64x16, 2 CTAs/SM) for the 4 layouts x {8,16}-byte loads.
#include "launch_api.cuh"
namespace eb200 {
namespace {
template <int BM, int BN, int BK, int WM, int WN, int STAGES, int MINB, bool VEC>
int launch_mma_vec(cudaStream_t s, TileArgs &args, dim3 grid) {
    const bool ak = args.p.a_kmajor, bk = args.p.b_kmajor;
#define EB_MMA_CASE(AK, BKM)                                                                          \
    {                                                                                                 \
        using Cfg = TileCfg<BM, BN, BK, WM, WN, STAGES, AK, BKM>;                                     \
        auto kern = tile_mma_f64<BM, BN, BK, WM, WN, STAGES, AK, BKM, MINB, VEC>;                     \
        int  rc = set_smem_attr(kern, Cfg::SMEM);                                                          \
        if (rc) return rc;                                                                            \
        kern<<<grid, Cfg::NT, Cfg::SMEM, s>>>(args);                                                  \
    }
    if (ak && bk) EB_MMA_CASE(true, true)
    else if (ak && !bk) EB_MMA_CASE(true, false)
    else if (!ak && bk) EB_MMA_CASE(false, true)
    else EB_MMA_CASE(false, false)
#undef EB_MMA_CASE
    EB_POST_LAUNCH("tile_mma_f64");
    return EB200_OK;
}
template <int BM, int BN, int BK, int WM, int WN, int STAGES, int MINB>
int launch_mma_cfg(cudaStream_t s, TileArgs &args, dim3 grid) {
    if (args.vec_ab) return launch_mma_vec<BM, BN, BK, WM, WN, STAGES, MINB, true>(s, args, grid);
    return launch_mma_vec<BM, BN, BK, WM, WN, STAGES, MINB, false>(s, args, grid);
}

} // namespace
int k_launch_mma(cudaStream_t s, TileArgs &args, dim3 grid) { return launch_mma_cfg<128, 64, 16, 32, 32, 3, 2>(s, args, grid); }
} // namespace eb200
