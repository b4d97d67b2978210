// blas.cu — the BLAS-like flat seam (column-major, device pointers, 64-bit sizes).
// Replaces libs/Einsums/hipBLASVendor/src/{gemm,gemv,dot,ger,axpy,scal,copy,direct_prod,lacgv}.hip
// (reference paths). Level-2/3 entry points build a contraction descriptor and run the engine in
// contract.cu; level-1 entry points are single-pass streaming kernels.
#include "contract.cuh"

#include <algorithm>
#include <cctype>

namespace eb200 {
namespace {

int grid_for_n(int64_t n, int threads, int per_thread) {
    int64_t need = (n + (int64_t)threads * per_thread - 1) / ((int64_t)threads * per_thread);
    int64_t cap = (int64_t)sm_count() * 16;
    return (int)std::max<int64_t>(1, std::min(need, cap));
}

// ---------------------------------------------------------------- level-1 kernels
enum L1Op { L1_AXPY, L1_AXPBY, L1_SCAL, L1_COPY, L1_FILL, L1_DIRPROD, L1_CONJ, L1_DIRDIV, L1_DIRMUL };
__device__ __forceinline__ float   div_(float a, float b) { return a / b; }
__device__ __forceinline__ double  div_(double a, double b) { return a / b; }
__device__ __forceinline__ float2  div_(float2 a, float2 b) { float d = b.x * b.x + b.y * b.y; return make_float2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d); }
__device__ __forceinline__ double2 div_(double2 a, double2 b) { double d = b.x * b.x + b.y * b.y; return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d); }

template <typename T, int OP> __device__ __forceinline__ void l1_apply(T alpha, T beta, const T *x, int64_t ix, const T *y, int64_t iy, T *z, int64_t iz) {
    if constexpr (OP == L1_AXPY) z[iz] = fma_(alpha, x[ix], z[iz]);
    else if constexpr (OP == L1_AXPBY) z[iz] = is_zero(beta) ? mul_(alpha, x[ix]) : fma_(alpha, x[ix], mul_(beta, z[iz]));
    else if constexpr (OP == L1_SCAL) z[iz] = mul_(alpha, z[iz]);
    else if constexpr (OP == L1_COPY) z[iz] = x[ix];
    else if constexpr (OP == L1_FILL) z[iz] = alpha;
    else if constexpr (OP == L1_DIRPROD) z[iz] = fma_(mul_(alpha, x[ix]), y[iy], z[iz]);
    else if constexpr (OP == L1_CONJ) z[iz] = conj_(z[iz]);
    else if constexpr (OP == L1_DIRDIV) z[iz] = div_(z[iz], x[ix]);
    else if constexpr (OP == L1_DIRMUL) z[iz] = mul_(z[iz], x[ix]);
}

template <typename T, int OP>
__global__ void __launch_bounds__(256) l1_kernel(int64_t n, T alpha, T beta, const T *__restrict__ x, int64_t incx, const T *__restrict__ y,
                                                int64_t incy, T *__restrict__ z, int64_t incz) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t       i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // 4 independent elements in flight per thread
    for (; i + 3 * stride < n; i += 4 * stride) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            int64_t j = i + u * stride;
            l1_apply<T, OP>(alpha, beta, x, j * incx, y, j * incy, z, j * incz);
        }
    }
    for (; i < n; i += stride) l1_apply<T, OP>(alpha, beta, x, i * incx, y, i * incy, z, i * incz);
}

// Unit increments, 16-byte aligned: one small CTA per 2 KB chunk (128 threads x one 16-byte pack of every stream), no loop, ~20 registers,
// so 2048 threads per SM are resident. In-place streams (scal: 5.5 -> ... TB/s) need that many independent CTAs in different phases of
// "load, then store the same lines"; the grid-stride kernel above keeps 8 x fewer threads resident.
template <typename T, int OP>
__global__ void __launch_bounds__(128) l1_kernel_contig(int64_t npack, T alpha, T beta, const T *__restrict__ x, const T *__restrict__ y, T *__restrict__ z) {
    constexpr int VEC = 16 / sizeof(T) > 0 ? 16 / sizeof(T) : 1;
    struct alignas(sizeof(T) * VEC) Pack { T v[VEC]; };
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npack) return;
    Pack px, py, pz;
    constexpr bool reads_x = OP == L1_AXPY || OP == L1_AXPBY || OP == L1_COPY || OP == L1_DIRPROD || OP == L1_DIRDIV || OP == L1_DIRMUL;
    constexpr bool reads_y = OP == L1_DIRPROD;
    constexpr bool reads_z = OP != L1_COPY && OP != L1_FILL;
    if constexpr (reads_x) px = reinterpret_cast<const Pack *>(x)[i];
    if constexpr (reads_y) py = reinterpret_cast<const Pack *>(y)[i];
    if constexpr (reads_z) {
        if (OP != L1_AXPBY || !is_zero(beta)) pz = reinterpret_cast<const Pack *>(z)[i];
    }
#pragma unroll
    for (int e = 0; e < VEC; e++) {
        if constexpr (OP == L1_AXPY) pz.v[e] = fma_(alpha, px.v[e], pz.v[e]);
        else if constexpr (OP == L1_AXPBY) pz.v[e] = is_zero(beta) ? mul_(alpha, px.v[e]) : fma_(alpha, px.v[e], mul_(beta, pz.v[e]));
        else if constexpr (OP == L1_SCAL) pz.v[e] = mul_(alpha, pz.v[e]);
        else if constexpr (OP == L1_COPY) pz.v[e] = px.v[e];
        else if constexpr (OP == L1_FILL) pz.v[e] = alpha;
        else if constexpr (OP == L1_DIRPROD) pz.v[e] = fma_(mul_(alpha, px.v[e]), py.v[e], pz.v[e]);
        else if constexpr (OP == L1_CONJ) pz.v[e] = conj_(pz.v[e]);
        else if constexpr (OP == L1_DIRDIV) pz.v[e] = div_(pz.v[e], px.v[e]);
        else if constexpr (OP == L1_DIRMUL) pz.v[e] = mul_(pz.v[e], px.v[e]);
    }
    reinterpret_cast<Pack *>(z)[i] = pz;
}

template <typename T, int OP>
int l1_launch(void *stream, int64_t n, T alpha, T beta, const T *x, int64_t incx, const T *y, int64_t incy, T *z, int64_t incz, const char *name) {
    if (n <= 0) return EB200_OK;
    constexpr int VEC = 16 / sizeof(T) > 0 ? 16 / sizeof(T) : 1;
    const bool    contig = incz == 1 && (!x || incx == 1) && (!y || incy == 1) && (((uintptr_t)z | (uintptr_t)x | (uintptr_t)y) % 16) == 0 && n >= 4096;
    if (contig) {
        const int64_t npack = n / VEC, blocks = (npack + 127) / 128;
        if (blocks <= 0x7fffffffLL) {
            l1_kernel_contig<T, OP><<<(unsigned)blocks, 128, 0, as_stream(stream)>>>(npack, alpha, beta, x, y, z);
            EB_POST_LAUNCH(name);
            const int64_t done = npack * VEC;
            if (done == n) return EB200_OK;
            // tail of n % VEC elements through the strided kernel
            l1_kernel<T, OP><<<1, 32, 0, as_stream(stream)>>>(n - done, alpha, beta, x ? x + done : x, 1, y ? y + done : y, 1, z + done, 1);
            EB_POST_LAUNCH(name);
            return EB200_OK;
        }
    }
    l1_kernel<T, OP><<<grid_for_n(n, 256, 4), 256, 0, as_stream(stream)>>>(n, alpha, beta, x, incx, y, incy, z, incz);
    EB_POST_LAUNCH(name);
    return EB200_OK;
}

// BLAS negative-increment convention: element i lives at (n-1-i)*|inc| ... we fold it into a base shift
template <typename T> const T *blas_base(const T *p, int64_t n, int64_t inc) { return (inc < 0 && n > 0) ? p - (n - 1) * inc : p; }
template <typename T> T       *blas_base(T *p, int64_t n, int64_t inc) { return (inc < 0 && n > 0) ? p - (n - 1) * inc : p; }

// ---------------------------------------------------------------- descriptor builders
template <typename T> void desc_init(eb200_contract_desc &d, int n_unique, T alpha, T beta) {
    memset(&d, 0, sizeof(d));
    d.n_unique = n_unique;
    d.C.dtype = d.A.dtype = d.B.dtype = Scalar<T>::dtype;
    if constexpr (Scalar<T>::is_complex) {
        d.alpha[0] = alpha.x; d.alpha[1] = alpha.y;
        d.beta[0] = beta.x;   d.beta[1] = beta.y;
    } else {
        d.alpha[0] = alpha; d.beta[0] = beta;
    }
}
void desc_axis(eb200_tensor_desc &t, int idx, int64_t stride) {
    t.index[t.rank] = idx;
    t.stride[t.rank] = stride;
    t.rank++;
}

int check_trans(char t, const char *which, const char *fn) {
    char c = (char)std::tolower((unsigned char)t);
    if (c != 'n' && c != 't' && c != 'c')
        return fail(EB200_ERR_INVALID_ARGUMENT, "The %s argument (%s) to %s call was invalid! Expected n, t, or c, case-insensitive, got %c.",
                    which, which[0] == 'f' ? "transA" : "transB", fn, t);
    return EB200_OK;
}

template <typename T>
int gemm_impl(void *stream, char transa, char transb, int64_t m, int64_t n, int64_t k, const T *alpha, const T *A, int64_t lda, int64_t strideA,
              const T *B, int64_t ldb, int64_t strideB, const T *beta, T *C, int64_t ldc, int64_t strideC, int64_t batch) {
    // reference order: early return first, then validation (HBV/src/gemm.hip:101-104)
    if (m == 0 || n == 0 || k == 0 || batch == 0) return EB200_OK;
    int rc;
    if ((rc = check_trans(transa, "first", "gemm"))) return rc;
    if ((rc = check_trans(transb, "second", "gemm"))) return rc;
    const bool notA = std::tolower((unsigned char)transa) == 'n', notB = std::tolower((unsigned char)transb) == 'n';
    const int64_t nrowa = notA ? m : k, nrowb = notB ? k : n;
    if (m < 0) return fail(EB200_ERR_DOMAIN, "The third argument (m) to gemm call was invalid! It must be greater than or equal to zero. Got %lld.", (long long)m);
    if (n < 0) return fail(EB200_ERR_DOMAIN, "The fourth argument (n) to gemm call was invalid! It must be greater than or equal to zero. Got %lld.", (long long)n);
    if (k < 0) return fail(EB200_ERR_DOMAIN, "The fifth argument (k) to gemm call was invalid! It must be greater than or equal to zero. Got %lld.", (long long)k);
    if (lda < std::max<int64_t>(1, nrowa))
        return fail(EB200_ERR_DOMAIN, "The eighth argument (lda) to gemm call was invalid! It must be at least 1 and at least the number of rows (%lld). Got %lld.",
                    (long long)nrowa, (long long)lda);
    if (ldb < std::max<int64_t>(1, nrowb))
        return fail(EB200_ERR_DOMAIN, "The tenth argument (ldb) to gemm call was invalid! It must be at least 1 and at least the number of rows (%lld). Got %lld.",
                    (long long)nrowb, (long long)ldb);
    if (ldc < std::max<int64_t>(1, m))
        return fail(EB200_ERR_DOMAIN, "The thirteenth argument (ldc) to gemm call was invalid! It must be at least 1 and at least the number of rows (%lld). Got %lld.",
                    (long long)m, (long long)ldc);
    if (batch < 0) return fail(EB200_ERR_DOMAIN, "gemm_strided_batched: batch must be >= 0, got %lld", (long long)batch);
    if (!alpha || !beta) return fail(EB200_ERR_INVALID_ARGUMENT, "gemm: alpha/beta pointer is NULL");

    // unique indices: 0=i (rows of C), 1=j (cols of C), 2=l (contracted), 3=batch
    eb200_contract_desc d;
    desc_init<T>(d, batch > 1 ? 4 : 3, *alpha, *beta);
    d.extent[0] = m; d.extent[1] = n; d.extent[2] = k; d.extent[3] = batch;
    d.C.data = C; d.A.data = const_cast<T *>(A); d.B.data = const_cast<T *>(B);
    desc_axis(d.C, 0, 1); desc_axis(d.C, 1, ldc);
    if (notA) { desc_axis(d.A, 0, 1); desc_axis(d.A, 2, lda); } else { desc_axis(d.A, 2, 1); desc_axis(d.A, 0, lda); }
    if (notB) { desc_axis(d.B, 2, 1); desc_axis(d.B, 1, ldb); } else { desc_axis(d.B, 1, 1); desc_axis(d.B, 2, ldb); }
    if (batch > 1) { desc_axis(d.C, 3, strideC); desc_axis(d.A, 3, strideA); desc_axis(d.B, 3, strideB); }
    d.conj_a = std::tolower((unsigned char)transa) == 'c';
    d.conj_b = std::tolower((unsigned char)transb) == 'c';
    return contract_dispatch(as_stream(stream), d);
}

template <typename T>
int gemv_impl(void *stream, char trans, int64_t m, int64_t n, const T *alpha, const T *A, int64_t lda, const T *x, int64_t incx, const T *beta, T *y,
              int64_t incy) {
    if (m == 0 || n == 0) return EB200_OK; // HBV/src/gemv.hip early return
    char t = (char)std::tolower((unsigned char)trans);
    if (t != 'n' && t != 't' && t != 'c')
        return fail(EB200_ERR_INVALID_ARGUMENT, "The first argument (transA) to gemv call was invalid! Expected n, t, or c, case-insensitive, got %c.", trans);
    if (m < 0) return fail(EB200_ERR_DOMAIN, "The second argument (m) to gemv call was invalid! It must be greater than or equal to zero. Got %lld.", (long long)m);
    if (n < 0) return fail(EB200_ERR_DOMAIN, "The third argument (n) to gemv call was invalid! It must be greater than or equal to zero. Got %lld.", (long long)n);
    if (lda < std::max<int64_t>(1, m))
        return fail(EB200_ERR_DOMAIN, "The sixth argument (lda) to gemv call was invalid! It must be at least 1 and at least the number of rows (%lld). Got %lld.",
                    (long long)m, (long long)lda);
    if (incx == 0) return fail(EB200_ERR_DOMAIN, "The eighth argument (incx) to gemv call was invalid! It must be non-zero.");
    if (incy == 0) return fail(EB200_ERR_DOMAIN, "The eleventh argument (incy) to gemv call was invalid! It must be non-zero.");
    if (!alpha || !beta) return fail(EB200_ERR_INVALID_ARGUMENT, "gemv: alpha/beta pointer is NULL");
    const int64_t leny = (t == 'n') ? m : n, lenx = (t == 'n') ? n : m;
    // indices: 0 = output index, 1 = contracted index
    eb200_contract_desc d;
    desc_init<T>(d, 2, *alpha, *beta);
    d.extent[0] = leny; d.extent[1] = lenx;
    d.C.data = blas_base(y, leny, incy);
    d.A.data = const_cast<T *>(A);
    d.B.data = const_cast<T *>(blas_base(x, lenx, incx));
    desc_axis(d.C, 0, incy);
    if (t == 'n') { desc_axis(d.A, 0, 1); desc_axis(d.A, 1, lda); } else { desc_axis(d.A, 1, 1); desc_axis(d.A, 0, lda); }
    desc_axis(d.B, 1, incx);
    d.conj_a = (t == 'c');
    return contract_dispatch(as_stream(stream), d);
}

template <typename T>
int ger_impl(void *stream, bool conj_y, int64_t m, int64_t n, const T *alpha, const T *x, int64_t incx, const T *y, int64_t incy, T *A, int64_t lda) {
    if (m == 0 || n == 0) return EB200_OK;
    if (m < 0) return fail(EB200_ERR_DOMAIN, "The first argument (m) to ger call was invalid! It must be greater than or equal to zero. Got %lld.", (long long)m);
    if (n < 0) return fail(EB200_ERR_DOMAIN, "The second argument (n) to ger call was invalid! It must be greater than or equal to zero. Got %lld.", (long long)n);
    if (incx == 0) return fail(EB200_ERR_DOMAIN, "The fifth argument (incx) to ger call was invalid! It must be non-zero.");
    if (incy == 0) return fail(EB200_ERR_DOMAIN, "The seventh argument (incy) to ger call was invalid! It must be non-zero.");
    if (lda < std::max<int64_t>(1, m))
        return fail(EB200_ERR_DOMAIN, "The ninth argument (lda) to ger call was invalid! It must be at least 1 and at least the number of rows (%lld). Got %lld.",
                    (long long)m, (long long)lda);
    if (!alpha) return fail(EB200_ERR_INVALID_ARGUMENT, "ger: alpha pointer is NULL");
    eb200_contract_desc d;
    T one = make_scalar<T>(1.0, 0.0);
    desc_init<T>(d, 2, *alpha, one);
    d.extent[0] = m; d.extent[1] = n;
    d.C.data = A;
    d.A.data = const_cast<T *>(blas_base(x, m, incx));
    d.B.data = const_cast<T *>(blas_base(y, n, incy));
    desc_axis(d.C, 0, 1); desc_axis(d.C, 1, lda);
    desc_axis(d.A, 0, incx);
    desc_axis(d.B, 1, incy);
    d.conj_b = conj_y;
    return contract_dispatch(as_stream(stream), d);
}

template <typename T> int dot_dev_impl(void *stream, bool conj_x, int64_t n, const T *x, int64_t incx, const T *y, int64_t incy, T *result_dev) {
    if (!result_dev) return fail(EB200_ERR_INVALID_ARGUMENT, "dot: result pointer is NULL");
    eb200_contract_desc d;
    desc_init<T>(d, 1, make_scalar<T>(1.0, 0.0), make_scalar<T>(0.0, 0.0));
    if (n <= 0) { // empty sum: result = 0
        EB_CUDA(cudaMemsetAsync(result_dev, 0, sizeof(T), as_stream(stream)));
        return EB200_OK;
    }
    d.extent[0] = n;
    d.C.data = result_dev;
    d.A.data = const_cast<T *>(blas_base(x, n, incx));
    d.B.data = const_cast<T *>(blas_base(y, n, incy));
    desc_axis(d.A, 0, incx);
    desc_axis(d.B, 0, incy);
    d.conj_a = conj_x;
    return contract_dispatch(as_stream(stream), d);
}

template <typename T> int dot_host_impl(void *stream, bool conj_x, int64_t n, const T *x, int64_t incx, const T *y, int64_t incy, T *result_host) {
    if (!result_host) return fail(EB200_ERR_INVALID_ARGUMENT, "dot: result pointer is NULL");
    T  *tmp = nullptr;
    int rc = scratch_alloc((void **)&tmp, sizeof(T), as_stream(stream));
    if (rc) return rc;
    rc = dot_dev_impl<T>(stream, conj_x, n, x, incx, y, incy, tmp);
    if (rc == EB200_OK) {
        cudaError_t e = cudaMemcpyAsync(result_host, tmp, sizeof(T), cudaMemcpyDeviceToHost, as_stream(stream));
        if (e == cudaSuccess) e = cudaStreamSynchronize(as_stream(stream));
        if (e != cudaSuccess) rc = cuda_fail(e, "dot result copy");
    }
    cudaFreeAsync(tmp, as_stream(stream));
    return rc;
}

} // namespace
} // namespace eb200

using namespace eb200;

#define EB_CPLX(T2, p) (*reinterpret_cast<const T2 *>(p))

#define EB200_DEF_BLAS(P, CT, KT)                                                                                                        \
    extern "C" int eb200_##P##gemm(void *stream, char transa, char transb, int64_t m, int64_t n, int64_t k, const CT *alpha, const CT *A, \
                                   int64_t lda, const CT *B, int64_t ldb, const CT *beta, CT *C, int64_t ldc) {                           \
        return gemm_impl<KT>(stream, transa, transb, m, n, k, (const KT *)alpha, (const KT *)A, lda, 0, (const KT *)B, ldb, 0,            \
                             (const KT *)beta, (KT *)C, ldc, 0, 1);                                                                       \
    }                                                                                                                                     \
    extern "C" int eb200_##P##gemm_strided_batched(void *stream, char transa, char transb, int64_t m, int64_t n, int64_t k,             \
                                                   const CT *alpha, const CT *A, int64_t lda, int64_t strideA, const CT *B, int64_t ldb,  \
                                                   int64_t strideB, const CT *beta, CT *C, int64_t ldc, int64_t strideC, int64_t batch) { \
        return gemm_impl<KT>(stream, transa, transb, m, n, k, (const KT *)alpha, (const KT *)A, lda, strideA, (const KT *)B, ldb,         \
                             strideB, (const KT *)beta, (KT *)C, ldc, strideC, batch);                                                    \
    }                                                                                                                                     \
    extern "C" int eb200_##P##gemv(void *stream, char trans, int64_t m, int64_t n, const CT *alpha, const CT *A, int64_t lda,            \
                                   const CT *x, int64_t incx, const CT *beta, CT *y, int64_t incy) {                                      \
        return gemv_impl<KT>(stream, trans, m, n, (const KT *)alpha, (const KT *)A, lda, (const KT *)x, incx, (const KT *)beta,           \
                             (KT *)y, incy);                                                                                              \
    }                                                                                                                                     \
    extern "C" int eb200_##P##dot(void *stream, int64_t n, const CT *x, int64_t incx, const CT *y, int64_t incy, CT *result_host) {      \
        return dot_host_impl<KT>(stream, false, n, (const KT *)x, incx, (const KT *)y, incy, (KT *)result_host);                          \
    }                                                                                                                                     \
    extern "C" int eb200_##P##dot_dev(void *stream, int64_t n, const CT *x, int64_t incx, const CT *y, int64_t incy, CT *result_dev) {   \
        return dot_dev_impl<KT>(stream, false, n, (const KT *)x, incx, (const KT *)y, incy, (KT *)result_dev);                            \
    }                                                                                                                                     \
    extern "C" int eb200_##P##ger(void *stream, int64_t m, int64_t n, const CT *alpha, const CT *x, int64_t incx, const CT *y,           \
                                  int64_t incy, CT *A, int64_t lda) {                                                                     \
        return ger_impl<KT>(stream, false, m, n, (const KT *)alpha, (const KT *)x, incx, (const KT *)y, incy, (KT *)A, lda);              \
    }                                                                                                                                     \
    extern "C" int eb200_##P##axpy(void *stream, int64_t n, const CT *alpha, const CT *x, int64_t incx, CT *y, int64_t incy) {           \
        if (!alpha) return fail(EB200_ERR_INVALID_ARGUMENT, "axpy: alpha is NULL");                                                       \
        return l1_launch<KT, L1_AXPY>(stream, n, *(const KT *)alpha, zero_of<KT>(), blas_base((const KT *)x, n, incx), incx, nullptr, 0,  \
                                      blas_base((KT *)y, n, incy), incy, "axpy");                                                         \
    }                                                                                                                                     \
    extern "C" int eb200_##P##axpby(void *stream, int64_t n, const CT *alpha, const CT *x, int64_t incx, const CT *beta, CT *y,          \
                                    int64_t incy) {                                                                                       \
        if (!alpha || !beta) return fail(EB200_ERR_INVALID_ARGUMENT, "axpby: alpha/beta is NULL");                                        \
        return l1_launch<KT, L1_AXPBY>(stream, n, *(const KT *)alpha, *(const KT *)beta, blas_base((const KT *)x, n, incx), incx,         \
                                       nullptr, 0, blas_base((KT *)y, n, incy), incy, "axpby");                                           \
    }                                                                                                                                     \
    extern "C" int eb200_##P##scal(void *stream, int64_t n, const CT *alpha, CT *x, int64_t incx) {                                      \
        if (!alpha) return fail(EB200_ERR_INVALID_ARGUMENT, "scal: alpha is NULL");                                                       \
        return l1_launch<KT, L1_SCAL>(stream, n, *(const KT *)alpha, zero_of<KT>(), nullptr, 0, nullptr, 0, blas_base((KT *)x, n, incx),  \
                                      incx, "scal");                                                                                      \
    }                                                                                                                                     \
    extern "C" int eb200_##P##copy(void *stream, int64_t n, const CT *x, int64_t incx, CT *y, int64_t incy) {                            \
        return l1_launch<KT, L1_COPY>(stream, n, zero_of<KT>(), zero_of<KT>(), blas_base((const KT *)x, n, incx), incx, nullptr, 0,       \
                                      blas_base((KT *)y, n, incy), incy, "copy");                                                         \
    }                                                                                                                                     \
    extern "C" int eb200_##P##fill(void *stream, int64_t n, const CT *value, CT *x, int64_t incx) {                                      \
        if (!value) return fail(EB200_ERR_INVALID_ARGUMENT, "fill: value is NULL");                                                       \
        return l1_launch<KT, L1_FILL>(stream, n, *(const KT *)value, zero_of<KT>(), nullptr, 0, nullptr, 0, blas_base((KT *)x, n, incx),  \
                                      incx, "fill");                                                                                      \
    }                                                                                                                                     \
    extern "C" int eb200_##P##dirprod(void *stream, int64_t n, const CT *alpha, const CT *x, int64_t incx, const CT *y, int64_t incy,    \
                                      CT *z, int64_t incz) {                                                                              \
        if (!alpha) return fail(EB200_ERR_INVALID_ARGUMENT, "dirprod: alpha is NULL");                                                    \
        return l1_launch<KT, L1_DIRPROD>(stream, n, *(const KT *)alpha, zero_of<KT>(), blas_base((const KT *)x, n, incx), incx,           \
                                         blas_base((const KT *)y, n, incy), incy, blas_base((KT *)z, n, incz), incz, "dirprod");          \
    }                                                                                                                                     \
    extern "C" int eb200_##P##dirdiv(void *stream, int64_t n, const CT *x, int64_t incx, CT *y, int64_t incy) {                          \
        return l1_launch<KT, L1_DIRDIV>(stream, n, zero_of<KT>(), zero_of<KT>(), blas_base((const KT *)x, n, incx), incx, nullptr, 0,     \
                                        blas_base((KT *)y, n, incy), incy, "dirdiv");                                                     \
    }                                                                                                                                     \
    extern "C" int eb200_##P##dirmul(void *stream, int64_t n, const CT *x, int64_t incx, CT *y, int64_t incy) {                          \
        return l1_launch<KT, L1_DIRMUL>(stream, n, zero_of<KT>(), zero_of<KT>(), blas_base((const KT *)x, n, incx), incx, nullptr, 0,     \
                                        blas_base((KT *)y, n, incy), incy, "dirmul");                                                     \
    }

EB200_DEF_BLAS(s, float, float)
EB200_DEF_BLAS(d, double, double)
EB200_DEF_BLAS(c, float, float2)
EB200_DEF_BLAS(z, double, double2)

extern "C" {
int eb200_cdotc(void *stream, int64_t n, const float *x, int64_t incx, const float *y, int64_t incy, float *result_host) {
    return dot_host_impl<float2>(stream, true, n, (const float2 *)x, incx, (const float2 *)y, incy, (float2 *)result_host);
}
int eb200_zdotc(void *stream, int64_t n, const double *x, int64_t incx, const double *y, int64_t incy, double *result_host) {
    return dot_host_impl<double2>(stream, true, n, (const double2 *)x, incx, (const double2 *)y, incy, (double2 *)result_host);
}
int eb200_cgerc(void *stream, int64_t m, int64_t n, const float *alpha, const float *x, int64_t incx, const float *y, int64_t incy, float *A,
                int64_t lda) {
    return ger_impl<float2>(stream, true, m, n, (const float2 *)alpha, (const float2 *)x, incx, (const float2 *)y, incy, (float2 *)A, lda);
}
int eb200_zgerc(void *stream, int64_t m, int64_t n, const double *alpha, const double *x, int64_t incx, const double *y, int64_t incy,
                double *A, int64_t lda) {
    return ger_impl<double2>(stream, true, m, n, (const double2 *)alpha, (const double2 *)x, incx, (const double2 *)y, incy, (double2 *)A, lda);
}
int eb200_clacgv(void *stream, int64_t n, float *x, int64_t incx) {
    return l1_launch<float2, L1_CONJ>(stream, n, zero_of<float2>(), zero_of<float2>(), nullptr, 0, nullptr, 0, blas_base((float2 *)x, n, incx), incx,
                                      "lacgv");
}
int eb200_zlacgv(void *stream, int64_t n, double *x, int64_t incx) {
    return l1_launch<double2, L1_CONJ>(stream, n, zero_of<double2>(), zero_of<double2>(), nullptr, 0, nullptr, 0, blas_base((double2 *)x, n, incx),
                                       incx, "lacgv");
}
}
