// k_tile_simt.cu — instantiates tile_simt for f32 / f64 / c64 / c128.
#include "launch_api.cuh"
namespace eb200 {
namespace {
template <typename T, bool CA, bool CB> int launch_simt(cudaStream_t s, TileArgs &args, dim3 grid) {
    tile_simt<T, CA, CB><<<grid, 256, 0, s>>>(args);
    EB_POST_LAUNCH("tile_simt");
    return EB200_OK;
}
template <typename T> int launch_simt_conj(cudaStream_t s, TileArgs &args, dim3 grid) {
    if constexpr (Scalar<T>::is_complex) {
        if (args.conj_a && args.conj_b) return launch_simt<T, true, true>(s, args, grid);
        if (args.conj_a) return launch_simt<T, true, false>(s, args, grid);
        if (args.conj_b) return launch_simt<T, false, true>(s, args, grid);
    }
    return launch_simt<T, false, false>(s, args, grid);
}

} // namespace
int k_launch_simt(int dtype, cudaStream_t s, TileArgs &args, dim3 grid) {
    switch (dtype) {
    case EB200_F32: return launch_simt_conj<float>(s, args, grid);
    case EB200_F64: return launch_simt_conj<double>(s, args, grid);
    case EB200_C64: return launch_simt_conj<float2>(s, args, grid);
    default: return launch_simt_conj<double2>(s, args, grid);
    }
}
} // namespace eb200
