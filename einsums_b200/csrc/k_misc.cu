// k_misc.cu — elementwise / scale / split finalize / type conversion kernels.
#include "launch_api.cuh"

#include <algorithm>
namespace eb200 {
namespace {
template <typename T, bool CA, bool CB> int launch_elem(cudaStream_t s, ElemArgs &a, int grid) {
    // pair path: innermost target index unit-stride in C, even extent, every other C stride even, C base aligned to two elements
    bool pairs = sizeof(T) <= 8 && a.t.n >= 1 && a.t.s[2][0] == 1 && (a.t.ext[0] % 2) == 0 && ((uintptr_t)a.C % (2 * sizeof(T))) == 0;
    for (int i = 1; i < a.t.n && pairs; i++)
        if (a.t.s[2][i] % 2) pairs = false;
    if (pairs && a.t.n == 1 && a.t.ext[0] >= 256) { // one merged index (contiguous direct products): a single "row" for the 2-D kernel
        a.t.ext[1] = 1;
        a.t.s[0][1] = a.t.s[1][1] = a.t.s[2][1] = 0;
        a.t.n = 2;
    }
    if (pairs && a.t.n == 2 && a.t.ext[0] >= 256 && (a.t.ext[1] + 1) / 2 <= 65535 && (a.t.ext[0] / 2 + 127) / 128 <= 0x7fffffffLL) {
        dim3 g2((unsigned)((a.t.ext[0] / 2 + 127) / 128), (unsigned)((a.t.ext[1] + 1) / 2), 1);
        elementwise_pairs_2d<T, CA, CB><<<g2, 128, 0, s>>>(a);
        EB_POST_LAUNCH("elementwise_pairs_2d");
        return EB200_OK;
    }
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

// ---------------------------------------------------------------- K-split sharding: fixed-order sum of the ranks' partial results
namespace eb200 {
namespace {
// One thread owns VEC consecutive elements; the n_slots loads of a thread are independent (all issued before the adds), the sum runs
// slot 0, 1, ... so that the result does not depend on which rank finished first. Streams n_slots * count elements once: HBM bound.
template <typename T, int VEC> __global__ void __launch_bounds__(256) reduce_slots_kernel(int n_slots, const T *__restrict__ slots, int64_t slot_stride,
                                                                                       int64_t count, T beta, T *__restrict__ C) {
    struct alignas(sizeof(T) * VEC) Pack { T v[VEC]; };
    const int64_t npack = count / VEC, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npack; i += stride) {
        Pack acc;
        if (beta != T(0)) {
            acc = *reinterpret_cast<const Pack *>(C + i * VEC);
#pragma unroll
            for (int e = 0; e < VEC; e++) acc.v[e] *= beta;
        } else {
#pragma unroll
            for (int e = 0; e < VEC; e++) acc.v[e] = T(0);
        }
        constexpr int BATCH = 8;
        for (int s0 = 0; s0 < n_slots; s0 += BATCH) {
            Pack p[BATCH];
#pragma unroll
            for (int b = 0; b < BATCH; b++)
                if (s0 + b < n_slots) p[b] = *reinterpret_cast<const Pack *>(slots + (int64_t)(s0 + b) * slot_stride + i * VEC);
#pragma unroll
            for (int b = 0; b < BATCH; b++)
                if (s0 + b < n_slots) {
#pragma unroll
                    for (int e = 0; e < VEC; e++) acc.v[e] += p[b].v[e];
                }
        }
        *reinterpret_cast<Pack *>(C + i * VEC) = acc;
    }
    // tail (count % VEC elements), by the first threads of the grid
    const int64_t t = npack * VEC + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < count) {
        T acc = (beta != T(0)) ? beta * C[t] : T(0);
        for (int s = 0; s < n_slots; s++) acc += slots[(int64_t)s * slot_stride + t];
        C[t] = acc;
    }
}
template <typename T> int reduce_slots_typed(cudaStream_t st, int n_slots, const void *slots, int64_t slot_stride, int64_t count, double beta, void *C) {
    constexpr int VEC = 16 / sizeof(T);
    const bool    vec = ((uintptr_t)slots % 16) == 0 && ((uintptr_t)C % 16) == 0 && (slot_stride % VEC) == 0;
    const int64_t work = vec ? (count + VEC - 1) / VEC : count;
    const int     grid = (int)std::max<int64_t>(1, std::min<int64_t>((work + 255) / 256, (int64_t)sm_count() * 8));
    if (vec) reduce_slots_kernel<T, VEC><<<grid, 256, 0, st>>>(n_slots, (const T *)slots, slot_stride, count, (T)beta, (T *)C);
    else reduce_slots_kernel<T, 1><<<grid, 256, 0, st>>>(n_slots, (const T *)slots, slot_stride, count, (T)beta, (T *)C);
    EB_POST_LAUNCH("reduce_slots");
    return EB200_OK;
}
} // namespace
} // namespace eb200

extern "C" int eb200_reduce_slots(void *stream, int dtype, int n_slots, const void *slots, int64_t slot_stride, int64_t count, double beta, void *C) {
    using namespace eb200;
    if (n_slots < 1 || count < 0 || slot_stride < count) return fail(EB200_ERR_DOMAIN, "reduce_slots: n_slots %d, count %lld, slot_stride %lld", n_slots, (long long)count, (long long)slot_stride);
    if (count == 0) return EB200_OK;
    if (!slots || !C) return fail(EB200_ERR_INVALID_ARGUMENT, "reduce_slots: NULL pointer");
    switch (dtype) {
    case EB200_F32: return reduce_slots_typed<float>(as_stream(stream), n_slots, slots, slot_stride, count, beta, C);
    case EB200_F64: return reduce_slots_typed<double>(as_stream(stream), n_slots, slots, slot_stride, count, beta, C);
    default: return fail(EB200_ERR_INVALID_ARGUMENT, "reduce_slots: dtype %d (pass complex data as reals)", dtype);
    }
}

