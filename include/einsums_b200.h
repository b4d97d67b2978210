/*
 * einsums_b200.h — C ABI of libeinsums_b200.so: the sm_100a device backend for Einsums'
 * contraction hot path (einsum / permute / linear_algebra::{gemm,gemv,dot,ger,axpy,scal,
 * direct_product} on DeviceTensor<T,Rank>).
 *
 * Every entry point replaces one function of the reference's "flat seam" (the closest thing
 * Einsums has to an FFI), cited per declaration as file:line relative to the reference tree:
 *   HBV  = libs/Einsums/hipBLASVendor
 *   TA   = libs/Einsums/TensorAlgebra/include/Einsums/TensorAlgebra
 *   T    = libs/Einsums/Tensor/include/Einsums/Tensor
 *   GS   = libs/Einsums/GPUStreams
 *
 * Conventions (same as the reference's vendor layer unless stated):
 *   - plain pointers and 64-bit sizes; no C++/torch types;
 *   - BLAS-style entry points have Fortran (column-major) semantics on DEVICE pointers,
 *     scalars are passed by pointer to HOST memory (read before return);
 *   - complex numbers are interleaved (re,im) pairs: float[2] / double[2];
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream). Calls only
 *     enqueue work; they do not synchronise unless documented ("host result");
 *   - every function returns an eb200_status; on error eb200_last_error_string() (thread
 *     local) describes it with the same wording class as the reference's exceptions;
 *   - there is NO CPU fallback: without a usable CUDA device every compute call returns
 *     EB200_ERR_NO_DEVICE / EB200_ERR_CUDA.
 */
#ifndef EINSUMS_B200_H
#define EINSUMS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EB200_VERSION_MAJOR 0
#define EB200_VERSION_MINOR 1
#define EB200_MAX_RANK 12   /* max rank of one tensor in a contraction / permute          */
#define EB200_MAX_UNIQUE 24 /* max number of distinct indices in one contraction          */

typedef enum eb200_status {
    EB200_OK = 0,
    /* maps to std::invalid_argument in the C++ shim (HBV/src/gemm.hip:39-49, gemv.hip:27-33) */
    EB200_ERR_INVALID_ARGUMENT = 1,
    /* maps to std::domain_error (HBV/src/gemm.hip:51-80: negative sizes, too-small ld)       */
    EB200_ERR_DOMAIN = 2,
    /* maps to einsums::dimension_error (TA/Backends/Dispatch.hpp:125-127, DevicePermute.hpp:121-126) */
    EB200_ERR_DIMENSION = 3,
    /* maps to einsums::rank_error / tensor_compat_error                                        */
    EB200_ERR_RANK = 4,
    EB200_ERR_UNSUPPORTED = 5,
    /* cudaErrorMemoryAllocation; maps to hip_exception<hipErrorOutOfMemory> (Errors/Error.hpp:1476-1500) */
    EB200_ERR_OUT_OF_MEMORY = 6,
    /* any other CUDA runtime error; eb200_last_cuda_error() returns the cudaError_t           */
    EB200_ERR_CUDA = 7,
    EB200_ERR_NO_DEVICE = 8,
    EB200_ERR_NOT_INITIALIZED = 9
} eb200_status;

/* element types (value of eb200_contract_desc::dtype etc.) */
typedef enum eb200_dtype { EB200_F32 = 0, EB200_F64 = 1, EB200_C64 = 2, EB200_C128 = 3 } eb200_dtype;

/* AlgorithmChoice as reported by the reference's analyser (TA/Detail/Utilities.hpp:26-34).
 * eb200_contract reports which device kernel family actually ran through
 * eb200_contract_desc::kernel_used using the values below 100; */
typedef enum eb200_algorithm {
    EB200_ALG_GENERIC = 0,
    EB200_ALG_DOT = 1,
    EB200_ALG_DIRECT = 2,
    EB200_ALG_GER = 3,
    EB200_ALG_GEMV = 4,
    EB200_ALG_GEMM = 5,
    EB200_ALG_INDETERMINATE = 6
} eb200_algorithm;

/* device kernel families of the contraction engine (eb200_contract_desc::kernel_used) */
typedef enum eb200_kernel_family {
    EB200_KF_NONE = 0,        /* nothing launched (zero-size)                                  */
    EB200_KF_TILE_MMA = 1,    /* tiled tensor-pipe kernel (FP64 DMMA) with fused index gathers */
    EB200_KF_TILE_SIMT = 2,   /* tiled FMA kernel (f32 / complex) with fused index gathers     */
    EB200_KF_ROWREDUCE = 3,   /* streaming reduction: few outputs per row of A (gemv-like)     */
    EB200_KF_COLREDUCE = 4,   /* streaming reduction, target index contiguous in A (gemv-T)    */
    EB200_KF_FULLREDUCE = 5,  /* dot-like: everything is contracted                             */
    EB200_KF_ELEMENTWISE = 6, /* no contracted index: outer / direct / broadcast products      */
    EB200_KF_SCALE_ONLY = 7,  /* empty contraction space: C = beta*C only                       */
    EB200_KF_TILE_TCGEN05 = 8 /* FP32 plain GEMM on tcgen05/TMEM with 3xTF32 split accumulation  */
} eb200_kernel_family;

/* ------------------------------------------------------------------------------------------
 * Runtime: replaces GS/src/GPUStreams.hip:58-136 (initialize/finalize), :145-166 (stream_wait),
 * GS/include/Einsums/GPUStreams/GPUStreams.hpp:194-280 (get_stream & friends) and
 * libs/Einsums/GPUMemory (GPUAllocator: hipMalloc under a budget).
 * ---------------------------------------------------------------------------------------- */
int eb200_version(void);
/* Selects `device` for the calling thread, verifies it is sm_100, warms the context. */
int eb200_init(int device);
int eb200_finalize(void);
int eb200_device_count(int *count);
int eb200_set_device(int device);
const char *eb200_last_error_string(void);
int eb200_last_cuda_error(void);

int eb200_stream_create(void **stream);
int eb200_stream_destroy(void *stream);
int eb200_stream_synchronize(void *stream); /* gpu::stream_wait()                              */
int eb200_device_synchronize(void);         /* gpu::device_synchronize()                       */

/* events, so that callers without the CUDA runtime can time on the launching stream */
int eb200_event_create(void **event);
int eb200_event_destroy(void *event);
int eb200_event_record(void *event, void *stream);
int eb200_event_synchronize(void *event);
int eb200_event_elapsed_ms(void *start, void *stop, float *ms);

/* T/Backends/DeviceTensor.hpp:62-116 (ctor: hipMalloc) and :33-60 (dtor) */
int eb200_malloc(void **ptr, size_t bytes);
int eb200_free(void *ptr);
/* Byte budget of eb200_malloc: replaces gpu::GPUAllocator / Einsums_GPUMemory_vars and the --einsums:gpu-buffer-size
 * option (libs/Einsums/GPUMemory/include/Einsums/GPUMemory/GPUAllocator.hpp:157-191, src/ModuleVars.cpp:43-63,
 * src/InitModule.cpp:39-55). An allocation that would push the live total past the budget fails with
 * EB200_ERR_OUT_OF_MEMORY before any CUDA call ("Could not allocate enough memory on the GPU device. Requested N bytes.").
 * 0 = unlimited (default). Process-wide, thread-safe. */
int eb200_set_memory_budget(size_t bytes);
int eb200_memory_stats(size_t *budget_bytes, size_t *in_use_bytes, size_t *peak_bytes, size_t *live_allocations);
int eb200_malloc_host(void **ptr, size_t bytes); /* pinned host memory (PINNED mode)           */
int eb200_free_host(void *ptr);
/* MAPPED mode (T/DeviceTensor.hpp:1309-1312, 1354-1357: hipHostRegister + hipHostGetDevicePointer): page-locks an existing
 * host range and returns the device-side alias kernels may dereference (zero-copy over PCIe). */
int eb200_host_register(void *ptr, size_t bytes, void **dev_ptr);
int eb200_host_unregister(void *ptr);
/* device-side alias of memory obtained from eb200_malloc_host (PINNED mode, :1313-1316) */
int eb200_host_get_device_pointer(void *host_ptr, void **dev_ptr);
int eb200_mem_get_info(size_t *free_bytes, size_t *total_bytes);
/* Peer memory across the processes of one node (one process per GPU): the owner exports an allocation made with eb200_malloc, the
 * peers open it and may then pass the returned pointer to ANY entry point of this library as a device pointer -- loads and stores go
 * over NVLink / NVSwitch. Used by the K-split sharding (einsums_b200/sharding.py): the partial result of a contraction is written by
 * the GEMM epilogue straight into the owning rank's receive slot. `handle` is EB200_IPC_HANDLE_BYTES opaque bytes, to be sent to the
 * peers by any means (torch.distributed.all_gather_object here). */
#define EB200_IPC_HANDLE_BYTES 64
int eb200_ipc_get_handle(void *dev_ptr, void *handle);
int eb200_ipc_open_handle(const void *handle, void **peer_ptr);
int eb200_ipc_close_handle(void *peer_ptr);
/* C[i] = beta * C[i] + sum_{s = 0 .. n_slots-1} slots[s * slot_stride + i],  i < count: the fixed-order (slot 0 first) sum of the partial
 * results the ranks deposited in this rank's receive slots; beta == 0 never reads C. dtype: EB200_F32 / EB200_F64 (complex data can
 * be passed as twice as many reals when beta is real). */
int eb200_reduce_slots(void *stream, int dtype, int n_slots, const void *slots, int64_t slot_stride, int64_t count, double beta, void *C);

/* contiguous copies; *_async variants only enqueue */
int eb200_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream);
int eb200_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream);
int eb200_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream);
int eb200_memcpy_h2d_async(void *dst, const void *src, size_t bytes, void *stream);
int eb200_memcpy_d2h_async(void *dst, const void *src, size_t bytes, void *stream);
int eb200_memset_zero(void *dst, size_t bytes, void *stream);

/* Strided device->device copy with layout change: dst[sum i_k*dst_strides[k]] =
 * src[sum i_k*src_strides[k]] over `dims` (element strides). Replaces the copy_to_tensor
 * kernels of T/Backends/DeviceTensor.hpp:295-402 and DeviceTensorView.hpp:59 (view <- tensor,
 * host column-major -> device row-major conversion after a raw upload).                     */
int eb200_copy_strided(void *stream, int dtype, int rank, const int64_t *dims, const void *src,
                       const int64_t *src_strides, void *dst, const int64_t *dst_strides);

/* ------------------------------------------------------------------------------------------
 * BLAS-like flat seam. Column-major, device pointers, 64-bit sizes.
 * eb200_Xgemm replaces blas::hip::Xgemm            HBV/src/gemm.hip:82-154 (checks :23-80)
 * eb200_Xgemv replaces blas::hip::Xgemv            HBV/src/gemv.hip:57-122 (checks :27-55)
 * eb200_Xdot/Xdotc replaces blas::hip::Xdot(c)     HBV/src/dot.hip:24-118
 * eb200_Xger/Xgerc replaces blas::hip::Xger(c)     HBV/src/ger.hip:46-122
 * eb200_Xaxpy/axpby/scal/copy/lacgv                HBV/src/axpy.hip:22-100, scal.hip:25-70,
 *                                                  copy.hip:24-49, lacgv.hip
 * eb200_Xdirprod replaces blas::hip::Xdirprod      HBV/src/direct_prod.hip:202-270
 * Early-return rules are the reference's: gemm returns at once if m, n or k is 0
 * (gemm.hip:101-102), gemv if m or n is 0.
 * ---------------------------------------------------------------------------------------- */
#define EB200_DECL_BLAS(P, T)                                                                                   \
    int eb200_##P##gemm(void *stream, char transa, char transb, int64_t m, int64_t n, int64_t k,              \
                        const T *alpha, const T *A, int64_t lda, const T *B, int64_t ldb, const T *beta, T *C, \
                        int64_t ldc);                                                                           \
    int eb200_##P##gemm_strided_batched(void *stream, char transa, char transb, int64_t m, int64_t n,         \
                                        int64_t k, const T *alpha, const T *A, int64_t lda, int64_t strideA,   \
                                        const T *B, int64_t ldb, int64_t strideB, const T *beta, T *C,         \
                                        int64_t ldc, int64_t strideC, int64_t batch);                          \
    int eb200_##P##gemv(void *stream, char trans, int64_t m, int64_t n, const T *alpha, const T *A,           \
                        int64_t lda, const T *x, int64_t incx, const T *beta, T *y, int64_t incy);             \
    /* result written to HOST memory; synchronises the stream (reference: hipblasXdot + return) */             \
    int eb200_##P##dot(void *stream, int64_t n, const T *x, int64_t incx, const T *y, int64_t incy,           \
                       T *result_host);                                                                         \
    /* result written to DEVICE memory; no synchronisation */                                                  \
    int eb200_##P##dot_dev(void *stream, int64_t n, const T *x, int64_t incx, const T *y, int64_t incy,       \
                           T *result_dev);                                                                      \
    int eb200_##P##ger(void *stream, int64_t m, int64_t n, const T *alpha, const T *x, int64_t incx,          \
                       const T *y, int64_t incy, T *A, int64_t lda);                                            \
    int eb200_##P##axpy(void *stream, int64_t n, const T *alpha, const T *x, int64_t incx, T *y,              \
                        int64_t incy);                                                                          \
    int eb200_##P##axpby(void *stream, int64_t n, const T *alpha, const T *x, int64_t incx, const T *beta,    \
                         T *y, int64_t incy);                                                                   \
    int eb200_##P##scal(void *stream, int64_t n, const T *alpha, T *x, int64_t incx);                         \
    int eb200_##P##copy(void *stream, int64_t n, const T *x, int64_t incx, T *y, int64_t incy);               \
    int eb200_##P##fill(void *stream, int64_t n, const T *value, T *x, int64_t incx);                         \
    /* z += alpha * x * y elementwise */                                                                       \
    int eb200_##P##dirprod(void *stream, int64_t n, const T *alpha, const T *x, int64_t incx, const T *y,     \
                           int64_t incy, T *z, int64_t incz);                                                  \
    /* y /= x and y *= x elementwise (BlockTensor operator/= and operator*=, Tensor/BlockTensor.hpp) */        \
    int eb200_##P##dirdiv(void *stream, int64_t n, const T *x, int64_t incx, T *y, int64_t incy);             \
    int eb200_##P##dirmul(void *stream, int64_t n, const T *x, int64_t incx, T *y, int64_t incy);

EB200_DECL_BLAS(s, float)
EB200_DECL_BLAS(d, double)
/* complex: T is the scalar component type; all counts are in complex elements */
EB200_DECL_BLAS(c, float)
EB200_DECL_BLAS(z, double)
#undef EB200_DECL_BLAS

/* conjugated variants (complex only): x^H y, A += alpha x y^H, x = conj(x) */
int eb200_cdotc(void *stream, int64_t n, const float *x, int64_t incx, const float *y, int64_t incy, float *result_host);
int eb200_zdotc(void *stream, int64_t n, const double *x, int64_t incx, const double *y, int64_t incy, double *result_host);
int eb200_cgerc(void *stream, int64_t m, int64_t n, const float *alpha, const float *x, int64_t incx, const float *y,
                int64_t incy, float *A, int64_t lda);
int eb200_zgerc(void *stream, int64_t m, int64_t n, const double *alpha, const double *x, int64_t incx,
                const double *y, int64_t incy, double *A, int64_t lda);
int eb200_clacgv(void *stream, int64_t n, float *x, int64_t incx);
int eb200_zlacgv(void *stream, int64_t n, double *x, int64_t incx);

/* ------------------------------------------------------------------------------------------
 * permute: C[perm] = beta*C + alpha*A.  Replaces tensor_algebra::detail::gpu_permute
 * (TA/Backends/DevicePermute.hpp:25-32 -> LibreTT) and permute_kernel (:34-73).
 *   rank        number of axes (<= EB200_MAX_RANK)
 *   dimsA[i]    extent of axis i of A
 *   stridesA[i] element stride of axis i of A
 *   perm[i]     axis of A that axis i of C enumerates   (C.dim(i) >= A.dim(perm[i]))
 *   stridesC[i] element stride of axis i of C
 *   conj        non-zero: conjugate A (complex types)
 * beta == 0 overwrites C without reading it (TA/Permute.hpp:377-385).
 * ---------------------------------------------------------------------------------------- */
int eb200_spermute(void *stream, int rank, const int64_t *dimsA, const int64_t *stridesA, const int *perm,
                   const int64_t *stridesC, const float *alpha, const float *A, const float *beta, float *C, int conj);
int eb200_dpermute(void *stream, int rank, const int64_t *dimsA, const int64_t *stridesA, const int *perm,
                   const int64_t *stridesC, const double *alpha, const double *A, const double *beta, double *C,
                   int conj);
int eb200_cpermute(void *stream, int rank, const int64_t *dimsA, const int64_t *stridesA, const int *perm,
                   const int64_t *stridesC, const float *alpha, const float *A, const float *beta, float *C, int conj);
int eb200_zpermute(void *stream, int rank, const int64_t *dimsA, const int64_t *stridesA, const int *perm,
                   const int64_t *stridesC, const double *alpha, const double *A, const double *beta, double *C,
                   int conj);

/* ------------------------------------------------------------------------------------------
 * Generic contraction  C[c...] = beta*C[c...] + alpha * sum_{links} op(A[a...]) * op(B[b...])
 * Replaces the device einsum_generic_algorithm + its three kernels
 * (TA/Backends/GPUTensorAlgebra.hpp:28-342). The descriptor is exactly what that launcher
 * builds on the host: the list of unique indices (order: C's, then A's, then B's first
 * occurrences — UniqueT<C...,A...,B...>, :225), one extent per unique index resolved
 * C -> A -> B (get_dim_ranges_for_many, TensorBase/IndexUtilities.hpp:24-120), and per tensor
 * the compile_index_table output (position in the unique list of each axis) plus strides.
 * Semantics = the reference device kernel: sum over the FULL unique-index space; an index
 * repeated inside one tensor (Hadamard) walks the diagonal; an index that appears only in C
 * broadcasts; beta == 0 zero-fills first (never reads C).
 * The index permutation is fused into the operand loads: no transposed copy is ever made.
 * ---------------------------------------------------------------------------------------- */
typedef struct eb200_tensor_desc {
    void *data;                            /* device pointer (element 0)                        */
    int32_t dtype;                         /* eb200_dtype                                       */
    int32_t rank;                          /* 0 for a scalar                                    */
    int32_t index[EB200_MAX_RANK];         /* position of axis i in the unique-index list       */
    int64_t stride[EB200_MAX_RANK];        /* element stride of axis i                          */
} eb200_tensor_desc;

typedef struct eb200_contract_desc {
    int32_t n_unique;
    int32_t conj_a, conj_b;
    int64_t extent[EB200_MAX_UNIQUE];      /* extent of each unique index                       */
    eb200_tensor_desc C, A, B;
    /* prefactors, as (re, im); im ignored for real C. Interpreted in C's type.                */
    double alpha[2];
    double beta[2];
    /* tuning / testing knobs: 0 = automatic */
    int32_t force_kernel;                  /* eb200_kernel_family to force (testing), 0 = auto  */
    int32_t split_k;                       /* 0 = auto                                          */
    /* outputs */
    int32_t kernel_used;                   /* eb200_kernel_family that ran                      */
    int32_t launches;                      /* number of kernels enqueued                        */
} eb200_contract_desc;

int eb200_contract(void *stream, eb200_contract_desc *desc);

/* Grouped contraction: n descriptors enqueued on `stream` with no host synchronisation between them. Replaces the
 * per-block `einsum` loops of the reference's block / tiled tensor algebra
 * (libs/Einsums/TensorAlgebra/include/Einsums/TensorAlgebra/Backends/BlockAlgebra.hpp:36-47, TileAlgebra.hpp).
 * Runs of consecutive descriptors that are identical except for base pointers advancing by a constant step (equal-size
 * blocks of one allocation) are folded into a single launch by turning the step into an extra batch index.
 * *n_launched_groups (optional) receives the number of contractions actually dispatched (<= n). */
int eb200_contract_grouped(void *stream, int64_t n, eb200_contract_desc *descs, int32_t *n_launched_groups);

/* Contraction with tensors in HOST memory (pinned memory recommended; pageable works but copies do not overlap).
 * Replaces the reference's host-tensor offload (LinearAlgebra/Bases/gemm.hpp:248-486 impl_gemm_gpu,
 * TensorImpl/TensorImpl.hpp:1676-1776 gpu_cache_tensor). Every data pointer of the descriptor is inspected: device
 * pointers are used in place, host operands are uploaded once, and a host C is processed in slabs of about `slab_bytes`
 * (0 = 256 MiB) along its slowest axis with H2D | contract | D2H overlapped on three streams; C is updated in place.
 * Host-resident tensors must be dense row-major; the slowest index of a host C must be an uncontracted index carried by
 * exactly one operand. Returns after the result is in host memory. Device scratch is cached per host thread
 * (eb200_contract_host_release frees it). */
typedef struct eb200_host_stats {
    int64_t h2d_bytes, d2h_bytes, slabs;
} eb200_host_stats;
int eb200_contract_host(void *stream, const eb200_contract_desc *desc, size_t slab_bytes, eb200_host_stats *stats);
int eb200_contract_host_release(void);

/* Host-side only (no device needed): classification the engine would pick for this
 * descriptor, written to desc->kernel_used. Lets CPU tests pin the planner. */
int eb200_contract_plan(eb200_contract_desc *desc);

/* Name of the tile kernel the calling thread's last contraction launched ("" before the first one). The descriptor's kernel_used
 * names the family; this names the kernel inside the family:
 *   FP64   "gemm_tma_f64", "contract_tma_f64<128x64>", "contract_tma_f64<80x80>", "contract_tma_f64<packed>", "tile_mma_f64"
 *   FP32   "gemm_tf32x3_tcgen05", "gemm_tf32x3_tcgen05<pair>", "gemm_tf32x3_tcgen05<generic>", "gemm_tf32x3_tcgen05<generic,pair>",
 *          "gemm_tf32x3_tcgen05<packed>", "tile_simt"
 *   c64    "gemm_tf32x3_tcgen05<c64>", "tile_simt"
 *   c128   "gemm_tma_z", "gemm_tma_z<packed>", "tile_simt" */
const char *eb200_last_kernel_name(void);
/* Tuning / testing knobs, process-wide (defaults from EB200_USE_TMA, EB200_USE_TMA_GENERIC, EB200_TMA_TILE, EB200_TMA_DIST, EB200_PERMUTE_TMA):
 *   "tma"          1 | 0: use / avoid both TMA kernels (0 = cp.async gather kernel everywhere)
 *   "tma_generic"  1 | 0 | 2: multi-index TMA kernel by the padding rule / never / whenever it is legal
 *   "tma_tile"     -1 | 0 | 1: automatic / 128x64 / 80x80 tiles in the multi-index TMA kernel
 *   "tma_dist"     0 | n: automatic / issue the loads of step t + n while step t is consumed (1 <= n < stages)
 *   "permute_tma"  1 | 0: TMA-staged FP64 plane transpose whenever the layout has a tensor map / never (register-staged kernels)
 *   "tcgen05_generic" 1 | 0 | 2 | 3: multi-index FP32 / c64 contractions on the tcgen05 kernel by the padding / fill rules / never /
 *                  whenever legal (tests) / FP32 through packed operands even when tensor maps exist (A/B experiments);
 *                  default from EB200_TCGEN05_GENERIC
 *   "tcgen05_pair" 1 | 0 | 2: plain / batched FP32 products on CTA pairs (cta_group::2, 256 x 128 tile pairs) when M pads by <= 5 % and
 *                  every pair has work / never / always (default from EB200_TCGEN05_PAIR)
 *   "scratch_keep_mb" n: freed scratch (split-K partials, packed operands, widened copies) the library's PRIVATE memory pool keeps
 *                  cached across synchronisations, in MB (default 4096, EB200_SCRATCH_KEEP_MB); the rest goes back to the driver. A
 *                  contraction that needs more scratch than that in one call raises the threshold to its own demand, bounded by an
 *                  eighth of the device memory, so that repeated calls do not re-acquire it; 0 switches caching and growth off */
int eb200_set_option(const char *name, int value);
int eb200_get_option(const char *name, int *value);

/* Number of kernel launches enqueued by this thread through this library since the last
 * eb200_launch_count_reset(); bench.py reports it as "gpu_launches". */
int64_t eb200_launch_count(void);
void eb200_launch_count_reset(void);

#ifdef __cplusplus
}
#endif
#endif /* EINSUMS_B200_H */
