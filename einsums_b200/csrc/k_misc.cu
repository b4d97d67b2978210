// k_misc.cu — elementwise / scale / split finalize / type conversion kernels.
#include "launch_api.cuh"
namespace eb200 {
namespace {
template <typename T, bool CA, bool CB> int launch_elem(cudaStream_t s, ElemArgs &a, int grid) {
    // pair path: innermost target index unit-stride in C, even extent, every other C stride even, C base aligned to two elements
    bool pairs = sizeof(T) <= 8 && a.t.n >= 1 && a.t.s[2][0] == 1 && (a.t.ext[0] % 2) == 0 && ((uintptr_t)a.C % (2 * sizeof(T))) == 0;
    for (int i = 1; i < a.t.n && pairs; i++)
        if (a.t.s[2][i] % 2) pairs = false;
    if (pairs) {
        elementwise_pairs<T, CA, CB, 4><<<grid, 256, 0, s>>>(a);
        EB_POST_LAUNCH("elementwise_pairs");
        return EB200_OK;
    }
    elementwise<T, CA, CB><<<grid, 256, 0, s>>>(a);
    EB_POST_LAUNCH("elementwise");
    return EB200_OK;
}
template <typename S, typename D> __device__ __forceinline__ D convert_scalar(S v) {
    if constexpr (Scalar<S>::is_complex && Scalar<D>::is_complex) return make_scalar<D>((double)v.x, (double)v.y);
    else if constexpr (Scalar<S>::is_complex) return make_scalar<D>((double)v.x, 0.0); // complex -> real keeps the real part
    else return make_scalar<D>((double)v, 0.0);
}
template <typename S, typename D> __global__ void __launch_bounds__(256) convert_tensor(const __grid_constant__ ConvArgs a) {
    const S *src = reinterpret_cast<const S *>(a.src);
    D       *dst = reinterpret_cast<D *>(a.dst);
    for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < a.g.tot; x += (int64_t)gridDim.x * blockDim.x) {
        int64_t os, od;
        grp_offset2(a.g, x, ROLE_A, ROLE_C, os, od);
        dst[od] = convert_scalar<S, D>(src[os]);
    }
}
template <typename S, typename D> int launch_convert(cudaStream_t s, const ConvArgs &a, int grid) {
    convert_tensor<S, D><<<grid, 256, 0, s>>>(a);
    EB_POST_LAUNCH("convert_tensor");
    return EB200_OK;
}
template <typename S> int launch_convert_to(cudaStream_t s, int dst_dtype, const ConvArgs &a, int grid) {
    switch (dst_dtype) {
    case EB200_F32: return launch_convert<S, float>(s, a, grid);
    case EB200_F64: return launch_convert<S, double>(s, a, grid);
    case EB200_C64: return launch_convert<S, float2>(s, a, grid);
    default: return launch_convert<S, double2>(s, a, grid);
    }
}
int launch_convert_any(cudaStream_t s, int src_dtype, int dst_dtype, const ConvArgs &a, int grid) {
    switch (src_dtype) {
    case EB200_F32: return launch_convert_to<float>(s, dst_dtype, a, grid);
    case EB200_F64: return launch_convert_to<double>(s, dst_dtype, a, grid);
    case EB200_C64: return launch_convert_to<float2>(s, dst_dtype, a, grid);
    default: return launch_convert_to<double2>(s, dst_dtype, a, grid);
    }
}

template <typename T> int elem_typed(cudaStream_t s, ElemArgs &a, int ca, int cb, int grid) {
    if constexpr (Scalar<T>::is_complex) {
        if (ca && cb) return launch_elem<T, true, true>(s, a, grid);
        if (ca) return launch_elem<T, true, false>(s, a, grid);
        if (cb) return launch_elem<T, false, true>(s, a, grid);
    }
    return launch_elem<T, false, false>(s, a, grid);
}
} // namespace
int k_launch_elem(int dtype, cudaStream_t s, ElemArgs &a, int ca, int cb, int grid) {
    switch (dtype) {
    case EB200_F32: return elem_typed<float>(s, a, ca, cb, grid);
    case EB200_F64: return elem_typed<double>(s, a, ca, cb, grid);
    case EB200_C64: return elem_typed<float2>(s, a, ca, cb, grid);
    default: return elem_typed<double2>(s, a, ca, cb, grid);
    }
}
#define EB_BY_DTYPE(KERNEL, NAME)                                                     \
    switch (dtype) {                                                                  \
    case EB200_F32: KERNEL<float><<<grid, 256, 0, s>>>(a); break;                     \
    case EB200_F64: KERNEL<double><<<grid, 256, 0, s>>>(a); break;                    \
    case EB200_C64: KERNEL<float2><<<grid, 256, 0, s>>>(a); break;                    \
    default: KERNEL<double2><<<grid, 256, 0, s>>>(a); break;                          \
    }                                                                                 \
    EB_POST_LAUNCH(NAME);                                                             \
    return EB200_OK;
int k_launch_scale(int dtype, cudaStream_t s, ScaleArgs &a, int grid) { EB_BY_DTYPE(scale_tensor, "scale_tensor") }
int k_launch_finalize(int dtype, cudaStream_t s, FinalizeArgs &a, int grid) { EB_BY_DTYPE(splitk_finalize, "splitk_finalize") }
int k_launch_dot(int dtype, cudaStream_t s, ReduceArgs &a, int grid) { EB_BY_DTYPE(dot_contig, "dot_contig") }
#undef EB_BY_DTYPE
int k_launch_convert(cudaStream_t s, int src_dtype, int dst_dtype, const ConvArgs &a, int grid) { return launch_convert_any(s, src_dtype, dst_dtype, a, grid); }
} // namespace eb200
