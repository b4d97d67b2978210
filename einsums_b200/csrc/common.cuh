// common.cuh — shared plumbing for libeinsums_b200.so (sm_100a only).
// Error state, launch accounting and the scalar-type algebra used by every kernel.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/einsums_b200.h"

namespace eb200 {

// ---------------------------------------------------------------- error state (thread local)
struct ErrorState {
    int  code = 0;
    int  cuda = 0;
    char msg[768] = {0};
};
ErrorState &err_state();
int          fail(int code, const char *fmt, ...);
int          cuda_fail(cudaError_t e, const char *what);
int64_t     &launch_counter();

#define EB_CUDA(expr)                                                  \
    do {                                                               \
        cudaError_t eb_e_ = (expr);                                    \
        if (eb_e_ != cudaSuccess) return ::eb200::cuda_fail(eb_e_, #expr); \
    } while (0)

// call after every <<<>>>: counts the launch and surfaces launch-configuration errors
#define EB_POST_LAUNCH(name)                                           \
    do {                                                               \
        ::eb200::launch_counter()++;                                   \
        cudaError_t eb_e_ = cudaGetLastError();                        \
        if (eb_e_ != cudaSuccess) return ::eb200::cuda_fail(eb_e_, name); \
    } while (0)

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }
int                 sm_count();
/// name of the tile kernel this thread launched last (eb200_last_kernel_name: tests and the benchmark report it)
void        set_last_kernel(const char *name);
/// tuning / testing knobs behind eb200_set_option
enum { OPT_TMA = 0, OPT_TMA_GENERIC, OPT_TMA_TILE, OPT_TMA_DIST, OPT_PERMUTE_TMA, OPT_TCGEN05_GENERIC, OPT_SCRATCH_KEEP_MB, OPT_TCGEN05_PAIR, OPT_COUNT };
int option(int which);
/// stream-ordered scratch from the library's private memory pool (runtime.cu); release with cudaFreeAsync on the same stream
int scratch_alloc(void **p, size_t bytes, cudaStream_t s);
/// a call about to take `bytes` of scratch in total: lets the pool keep that much cached (bounded by 1/8 of the device memory)
void scratch_expect(size_t bytes);

// ---------------------------------------------------------------- scalar algebra
// real types are float/double; complex are float2/double2 (interleaved re,im = the C ABI layout)
template <typename T> struct Scalar;
template <> struct Scalar<float> {
    using real = float;
    static constexpr bool is_complex = false;
    static constexpr int  dtype = EB200_F32;
};
template <> struct Scalar<double> {
    using real = double;
    static constexpr bool is_complex = false;
    static constexpr int  dtype = EB200_F64;
};
template <> struct Scalar<float2> {
    using real = float;
    static constexpr bool is_complex = true;
    static constexpr int  dtype = EB200_C64;
};
template <> struct Scalar<double2> {
    using real = double;
    static constexpr bool is_complex = true;
    static constexpr int  dtype = EB200_C128;
};

template <typename T> __host__ __device__ inline T make_scalar(double re, double im);
template <> __host__ __device__ inline float   make_scalar<float>(double re, double) { return (float)re; }
template <> __host__ __device__ inline double  make_scalar<double>(double re, double) { return re; }
template <> __host__ __device__ inline float2  make_scalar<float2>(double re, double im) { return make_float2((float)re, (float)im); }
template <> __host__ __device__ inline double2 make_scalar<double2>(double re, double im) { return make_double2(re, im); }

template <typename T> __host__ __device__ inline T zero_of() { return make_scalar<T>(0.0, 0.0); }

__host__ __device__ inline bool is_zero(float a) { return a == 0.0f; }
__host__ __device__ inline bool is_zero(double a) { return a == 0.0; }
__host__ __device__ inline bool is_zero(float2 a) { return a.x == 0.0f && a.y == 0.0f; }
__host__ __device__ inline bool is_zero(double2 a) { return a.x == 0.0 && a.y == 0.0; }
__host__ __device__ inline bool is_one(float a) { return a == 1.0f; }
__host__ __device__ inline bool is_one(double a) { return a == 1.0; }
__host__ __device__ inline bool is_one(float2 a) { return a.x == 1.0f && a.y == 0.0f; }
__host__ __device__ inline bool is_one(double2 a) { return a.x == 1.0 && a.y == 0.0; }

__device__ __forceinline__ float   add_(float a, float b) { return a + b; }
__device__ __forceinline__ double  add_(double a, double b) { return a + b; }
__device__ __forceinline__ float2  add_(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 add_(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }

__device__ __forceinline__ float   mul_(float a, float b) { return a * b; }
__device__ __forceinline__ double  mul_(double a, double b) { return a * b; }
__device__ __forceinline__ float2  mul_(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 mul_(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// a*b + c
__device__ __forceinline__ float  fma_(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float2 fma_(float2 a, float2 b, float2 c) {
    float re = fmaf(a.x, b.x, c.x); re = fmaf(-a.y, b.y, re);
    float im = fmaf(a.x, b.y, c.y); im = fmaf(a.y, b.x, im);
    return make_float2(re, im);
}
__device__ __forceinline__ double2 fma_(double2 a, double2 b, double2 c) {
    double re = fma(a.x, b.x, c.x); re = fma(-a.y, b.y, re);
    double im = fma(a.x, b.y, c.y); im = fma(a.y, b.x, im);
    return make_double2(re, im);
}

__device__ __forceinline__ float   conj_(float a) { return a; }
__device__ __forceinline__ double  conj_(double a) { return a; }
__device__ __forceinline__ float2  conj_(float2 a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ double2 conj_(double2 a) { return make_double2(a.x, -a.y); }

template <bool CONJ, typename T> __device__ __forceinline__ T maybe_conj(T a) {
    if constexpr (CONJ) return conj_(a);
    else return a;
}

// run-time conjugation switch (uniform flag); identity for real types
template <typename T> __device__ __forceinline__ T conj_if(T a, bool flag) {
    if constexpr (Scalar<T>::is_complex) return flag ? conj_(a) : a;
    else return a;
}

// out = alpha*v (+ beta*old when beta != 0). `has_beta` is uniform.
template <typename T> __device__ __forceinline__ T axpby_(T alpha, T v, T beta, T old, bool has_beta) {
    T r = mul_(alpha, v);
    if (has_beta) r = fma_(beta, old, r);
    return r;
}

// warp/block reductions with a fixed combination order (deterministic run to run)
template <typename T> __device__ __forceinline__ T shfl_down_(T v, int d);
template <> __device__ __forceinline__ float  shfl_down_(float v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
template <> __device__ __forceinline__ double shfl_down_(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
template <> __device__ __forceinline__ float2 shfl_down_(float2 v, int d) {
    return make_float2(__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d));
}
template <> __device__ __forceinline__ double2 shfl_down_(double2 v, int d) {
    return make_double2(__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d));
}
template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = add_(v, shfl_down_(v, d));
    return v;  // valid in lane 0
}

} // namespace eb200
