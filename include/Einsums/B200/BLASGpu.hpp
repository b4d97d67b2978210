// BLASGpu.hpp — einsums::blas::gpu::*<T>: the typed facade over the flat device-BLAS seam, under the reference's names.
//
// Reference: libs/Einsums/hipBLAS/include/Einsums/hipBLAS.hpp (gemm :63-94, gemv :132-160, scal :423, rscl :482, dot :545,
// dotc :586, axpy :621, axpby :667, ger :719, gerc :747, dirprod) forwarding to blas::hip::{s,d,c,z}* (hipBLASVendor/Vendor.hpp).
// Here the four value types forward to eb200_{s,d,c,z}* of the C ABI: column-major, DEVICE pointers, the calling thread's stream
// (gpu::get_stream()). Error behaviour is the vendor layer's: a bad trans flag throws std::invalid_argument, a bad size or
// leading dimension std::domain_error (gemm.hip:39-80) — the C ABI reports them as status codes and EB200_CHECK maps them back.
// The decomposition entries of the reference facade (syev, geev, gesv, getrf, gesdd, ...; hipSolver) are out of scope.
#pragma once
#include <Einsums/B200/Errors.hpp>
#include <Einsums/B200/Runtime.hpp>

#include <complex>
#include <cstdint>
#include <type_traits>

namespace einsums::blas {
using int_t = std::int64_t; // reference: BLAS/Types.hpp (ILP64 build)
}

namespace einsums::blas::gpu {

namespace detail {
template <typename T> struct Seam; // component type and the entry points of one value type
#define EB200_SEAM(P, T, R)                                                                                                                      \
    template <> struct Seam<T> {                                                                                                                 \
        using Real = R;                                                                                                                          \
        static constexpr auto gemm = eb200_##P##gemm;                                                                                            \
        static constexpr auto gemm_strided_batched = eb200_##P##gemm_strided_batched;                                                            \
        static constexpr auto gemv = eb200_##P##gemv;                                                                                            \
        static constexpr auto dot = eb200_##P##dot;                                                                                              \
        static constexpr auto ger = eb200_##P##ger;                                                                                              \
        static constexpr auto axpy = eb200_##P##axpy;                                                                                            \
        static constexpr auto axpby = eb200_##P##axpby;                                                                                          \
        static constexpr auto scal = eb200_##P##scal;                                                                                            \
        static constexpr auto copy = eb200_##P##copy;                                                                                            \
        static constexpr auto dirprod = eb200_##P##dirprod;                                                                                      \
    };
EB200_SEAM(s, float, float)
EB200_SEAM(d, double, double)
EB200_SEAM(c, std::complex<float>, float)
EB200_SEAM(z, std::complex<double>, double)
#undef EB200_SEAM
template <typename T> auto *rp(T *p) { return reinterpret_cast<typename Seam<std::remove_const_t<T>>::Real *>(p); }
template <typename T> auto const *rp(T const *p) { return reinterpret_cast<typename Seam<T>::Real const *>(p); }
template <typename T> inline constexpr bool is_complex_v = !std::is_arithmetic_v<T>;
} // namespace detail

/// C := alpha op(A) op(B) + beta C (column-major)
template <typename T>
void gemm(char transa, char transb, int_t m, int_t n, int_t k, T alpha, T const *a, int_t lda, T const *b, int_t ldb, T beta, T *c, int_t ldc) {
    using namespace detail;
    EB200_CHECK(Seam<T>::gemm(::einsums::gpu::get_stream(), transa, transb, m, n, k, rp(&alpha), rp(a), lda, rp(b), ldb, rp(&beta), rp(c), ldc));
}
/// `batch` products whose operands are `stride*` elements apart (the block/tiled algebra's equal-size blocks)
template <typename T>
void gemm_strided_batched(char transa, char transb, int_t m, int_t n, int_t k, T alpha, T const *a, int_t lda, int_t stridea, T const *b, int_t ldb,
                          int_t strideb, T beta, T *c, int_t ldc, int_t stridec, int_t batch) {
    using namespace detail;
    EB200_CHECK(Seam<T>::gemm_strided_batched(::einsums::gpu::get_stream(), transa, transb, m, n, k, rp(&alpha), rp(a), lda, stridea, rp(b), ldb, strideb,
                                              rp(&beta), rp(c), ldc, stridec, batch));
}
/// y := alpha op(A) x + beta y
template <typename T> void gemv(char transa, int_t m, int_t n, T alpha, T const *a, int_t lda, T const *x, int_t incx, T beta, T *y, int_t incy) {
    using namespace detail;
    EB200_CHECK(Seam<T>::gemv(::einsums::gpu::get_stream(), transa, m, n, rp(&alpha), rp(a), lda, rp(x), incx, rp(&beta), rp(y), incy));
}
/// x *= alpha
template <typename T> void scal(int_t n, T const alpha, T *vec, int_t inc) {
    using namespace detail;
    EB200_CHECK(Seam<T>::scal(::einsums::gpu::get_stream(), n, rp(&alpha), rp(vec), inc));
}
/// complex vector times a real scalar (reference: csscal / zdscal overloads)
template <typename T>
    requires detail::is_complex_v<T>
void scal(int_t n, typename T::value_type const alpha, T *vec, int_t inc) {
    scal<T>(n, T(alpha), vec, inc);
}
/// x /= alpha
template <typename T> void rscl(int_t n, T const alpha, T *vec, int_t inc) { scal<T>(n, T(1) / alpha, vec, inc); }
/// sum_i x_i y_i (no conjugation); the result comes back to the host, the stream is synchronised
template <typename T> auto dot(int_t n, T const *x, int_t incx, T const *y, int_t incy) -> T {
    using namespace detail;
    T out{};
    EB200_CHECK(Seam<T>::dot(::einsums::gpu::get_stream(), n, rp(x), incx, rp(y), incy, rp(&out)));
    return out;
}
/// sum_i conj(x_i) y_i
template <typename T> auto dotc(int_t n, T const *x, int_t incx, T const *y, int_t incy) -> T {
    using namespace detail;
    if constexpr (!is_complex_v<T>) return dot<T>(n, x, incx, y, incy);
    else {
        T out{};
        if constexpr (std::is_same_v<T, std::complex<float>>)
            EB200_CHECK(eb200_cdotc(::einsums::gpu::get_stream(), n, rp(x), incx, rp(y), incy, rp(&out)));
        else EB200_CHECK(eb200_zdotc(::einsums::gpu::get_stream(), n, rp(x), incx, rp(y), incy, rp(&out)));
        return out;
    }
}
/// y += alpha x
template <typename T> void axpy(int_t n, T alpha_x, T const *x, int_t inc_x, T *y, int_t inc_y) {
    using namespace detail;
    EB200_CHECK(Seam<T>::axpy(::einsums::gpu::get_stream(), n, rp(&alpha_x), rp(x), inc_x, rp(y), inc_y));
}
/// y := alpha x + b y
template <typename T> void axpby(int_t n, T alpha_x, T const *x, int_t inc_x, T b, T *y, int_t inc_y) {
    using namespace detail;
    EB200_CHECK(Seam<T>::axpby(::einsums::gpu::get_stream(), n, rp(&alpha_x), rp(x), inc_x, rp(&b), rp(y), inc_y));
}
/// A += alpha x y^T
template <typename T> void ger(int_t m, int_t n, T alpha, T const *x, int_t inc_x, T const *y, int_t inc_y, T *a, int_t lda) {
    using namespace detail;
    EB200_CHECK(Seam<T>::ger(::einsums::gpu::get_stream(), m, n, rp(&alpha), rp(x), inc_x, rp(y), inc_y, rp(a), lda));
}
/// A += alpha x y^H
template <typename T> void gerc(int_t m, int_t n, T alpha, T const *x, int_t inc_x, T const *y, int_t inc_y, T *a, int_t lda) {
    using namespace detail;
    if constexpr (!is_complex_v<T>) ger<T>(m, n, alpha, x, inc_x, y, inc_y, a, lda);
    else if constexpr (std::is_same_v<T, std::complex<float>>)
        EB200_CHECK(eb200_cgerc(::einsums::gpu::get_stream(), m, n, rp(&alpha), rp(x), inc_x, rp(y), inc_y, rp(a), lda));
    else EB200_CHECK(eb200_zgerc(::einsums::gpu::get_stream(), m, n, rp(&alpha), rp(x), inc_x, rp(y), inc_y, rp(a), lda));
}
/// y := x
template <typename T> void copy(int_t n, T const *x, int_t inc_x, T *y, int_t inc_y) {
    using namespace detail;
    EB200_CHECK(Seam<T>::copy(::einsums::gpu::get_stream(), n, rp(x), inc_x, rp(y), inc_y));
}
/// x := conj(x) (no-op for real types)
template <typename T> void lacgv(int_t n, T *x, int_t incx) {
    using namespace detail;
    if constexpr (std::is_same_v<T, std::complex<float>>) EB200_CHECK(eb200_clacgv(::einsums::gpu::get_stream(), n, rp(x), incx));
    else if constexpr (std::is_same_v<T, std::complex<double>>) EB200_CHECK(eb200_zlacgv(::einsums::gpu::get_stream(), n, rp(x), incx));
}
/// z += alpha x o y elementwise
template <typename T> void dirprod(int_t n, T alpha, T const *x, int_t incx, T const *y, int_t incy, T *z, int_t incz) {
    using namespace detail;
    EB200_CHECK(Seam<T>::dirprod(::einsums::gpu::get_stream(), n, rp(&alpha), rp(x), incx, rp(y), incy, rp(z), incz));
}

} // namespace einsums::blas::gpu
