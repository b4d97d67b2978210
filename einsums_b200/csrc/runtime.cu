// runtime.cu — device runtime of libeinsums_b200.so: init/finalize, streams, events, memory.
// Replaces libs/Einsums/GPUStreams/src/GPUStreams.hip:58-166 and the hipMalloc/hipMemcpy calls in
// libs/Einsums/Tensor/include/Einsums/Tensor/Backends/DeviceTensor.hpp:33-116 (reference paths).
#include "common.cuh"

#include <atomic>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

namespace eb200 {

ErrorState &err_state() {
    static thread_local ErrorState st;
    return st;
}
int64_t &launch_counter() {
    static thread_local int64_t n = 0;
    return n;
}

int fail(int code, const char *fmt, ...) {
    ErrorState &st = err_state();
    st.code = code;
    st.cuda = 0;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(st.msg, sizeof(st.msg), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char *what) {
    ErrorState &st = err_state();
    st.cuda = (int)e;
    if (e == cudaErrorMemoryAllocation) st.code = EB200_ERR_OUT_OF_MEMORY;
    else if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) st.code = EB200_ERR_NO_DEVICE;
    else st.code = EB200_ERR_CUDA;
    snprintf(st.msg, sizeof(st.msg), "CUDA error %d (%s) in %s: %s", (int)e, cudaGetErrorName(e), what, cudaGetErrorString(e));
    // clear the sticky "last error" for non-fatal failures so later calls start clean
    (void)cudaGetLastError();
    return st.code;
}

static int g_sm_count[64] = {0};
int        sm_count() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (g_sm_count[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        g_sm_count[dev] = n;
    }
    return g_sm_count[dev];
}

// ---------------------------------------------------------------- knobs and introspection
namespace {
struct Options {
    std::atomic<int> v[OPT_COUNT];
    Options() {
        auto env = [](const char *n, int dflt) { const char *e = getenv(n); return e ? atoi(e) : dflt; };
        v[OPT_TMA] = env("EB200_USE_TMA", 1);
        v[OPT_TMA_GENERIC] = env("EB200_USE_TMA_GENERIC", 1);
        v[OPT_TMA_TILE] = env("EB200_TMA_TILE", -1);
        v[OPT_TMA_DIST] = env("EB200_TMA_DIST", 0);
        v[OPT_PERMUTE_TMA] = env("EB200_PERMUTE_TMA", 1);
        v[OPT_TCGEN05_GENERIC] = env("EB200_TCGEN05_GENERIC", 1);
        v[OPT_SCRATCH_KEEP_MB] = env("EB200_SCRATCH_KEEP_MB", 4096);
        v[OPT_TCGEN05_PAIR] = env("EB200_TCGEN05_PAIR", 1);
    }
};
Options &options() {
    static Options o;
    return o;
}
const char *const kOptionNames[OPT_COUNT] = {"tma", "tma_generic", "tma_tile", "tma_dist", "permute_tma", "tcgen05_generic", "scratch_keep_mb", "tcgen05_pair"};
thread_local const char *t_last_kernel = "";
} // namespace
int  option(int which) { return options().v[which].load(std::memory_order_relaxed); }
void set_last_kernel(const char *name) { t_last_kernel = name; }

// ---------------------------------------------------------------- stream-ordered scratch
// Split-K partials, mixed-type widened copies and device-side scalars come from a PRIVATE memory pool per device. The
// device's default pool (shared with the embedding application, e.g. torch) is never touched. Freed scratch up to
// `keep` bytes stays cached across synchronisations (a release threshold of 0 would hand it back at every sync, ~1 ms
// per split-K call); anything above is returned to the driver at the next synchronisation point. `keep` is the option
// "scratch_keep_mb" (default 4096 MB = 2 % of the B200's HBM: the packed complex paths hold three operand-sized buffers per call, and
// re-acquiring 2 GB from the driver costs ~16 ms per call); eb200_set_option updates live pools, 0 returns everything at each sync.
namespace {
struct ScratchPools {
    std::mutex    mu;
    cudaMemPool_t pool[64] = {};
    uint64_t      keep[64] = {}; // current release threshold (bytes)
};
ScratchPools &scratch_pools() {
    static ScratchPools *p = new ScratchPools; // leaked on purpose: the context may be gone at static destruction
    return *p;
}
} // namespace

int scratch_alloc(void **p, size_t bytes, cudaStream_t s) {
    int dev = 0;
    EB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(EB200_ERR_UNSUPPORTED, "scratch_alloc: device index %d not supported", dev);
    ScratchPools &sp = scratch_pools();
    cudaMemPool_t pool;
    {
        std::lock_guard<std::mutex> lock(sp.mu);
        if (!sp.pool[dev]) {
            cudaMemPoolProps props;
            memset(&props, 0, sizeof(props));
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            EB_CUDA(cudaMemPoolCreate(&sp.pool[dev], &props));
            uint64_t keep = (uint64_t)std::max(0, option(OPT_SCRATCH_KEEP_MB)) << 20;
            EB_CUDA(cudaMemPoolSetAttribute(sp.pool[dev], cudaMemPoolAttrReleaseThreshold, &keep));
            sp.keep[dev] = keep;
        }
        pool = sp.pool[dev];
    }
    EB_CUDA(cudaMallocFromPoolAsync(p, bytes ? bytes : 1, pool, s));
    return EB200_OK;
}

// A call that is about to take `bytes` of scratch in total (packed operands of a large contraction) tells the pool: the release threshold
// grows to cover it, up to an eighth of the device's memory, so that the NEXT such call does not re-acquire the memory from the driver
// (Wmbej c64 o=60 v=240 needs 8 GB: 406 ms per call with 4 GB kept, 146 ms with the scratch cached). "scratch_keep_mb" = 0 switches this off.
void scratch_expect(size_t bytes) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
    const int base_mb = option(OPT_SCRATCH_KEEP_MB);
    if (base_mb <= 0) return;
    ScratchPools               &sp = scratch_pools();
    std::lock_guard<std::mutex> lock(sp.mu);
    if (!sp.pool[dev]) return; // the first allocation creates the pool with the base threshold; the next call grows it
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { (void)cudaGetLastError(); return; }
    const uint64_t cap = std::max<uint64_t>((uint64_t)base_mb << 20, (uint64_t)total_b / 8);
    const uint64_t want = std::min<uint64_t>((uint64_t)bytes + (64ull << 20), cap);
    if (want > sp.keep[dev] && cudaMemPoolSetAttribute(sp.pool[dev], cudaMemPoolAttrReleaseThreshold, (void *)&want) == cudaSuccess) sp.keep[dev] = want;
    else (void)cudaGetLastError();
}

// ---------------------------------------------------------------- budgeted device allocator
// Replaces gpu::GPUAllocator + Einsums_GPUMemory_vars (reference: libs/Einsums/GPUMemory/include/Einsums/GPUMemory/
// GPUAllocator.hpp:157-191 allocate/deallocate, src/ModuleVars.cpp:43-63 try_allocate/deallocate, src/InitModule.cpp:39-55
// --einsums:gpu-buffer-size): every eb200_malloc is charged against a byte budget and refused (no CUDA call) when it
// would exceed it. Budget 0 = unlimited (B200-first default: 180 GB of HBM belong to the tensors). The size of each live
// allocation is remembered so eb200_free needs no size argument.
namespace {
struct Budget {
    std::mutex                         mu;
    size_t                             budget = 0, in_use = 0, peak = 0;
    std::unordered_map<void *, size_t> live;
};
Budget &budget_state() {
    static Budget *b = new Budget;
    return *b;
}
} // namespace

} // namespace eb200

using namespace eb200;

extern "C" {

int eb200_version(void) { return EB200_VERSION_MAJOR * 1000 + EB200_VERSION_MINOR; }

const char *eb200_last_error_string(void) { return err_state().msg; }
int         eb200_last_cuda_error(void) { return err_state().cuda; }

int eb200_device_count(int *count) {
    if (!count) return fail(EB200_ERR_INVALID_ARGUMENT, "eb200_device_count: count is NULL");
    int         n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return cuda_fail(e, "cudaGetDeviceCount");
    }
    *count = n;
    return EB200_OK;
}

int eb200_set_device(int device) {
    EB_CUDA(cudaSetDevice(device));
    return EB200_OK;
}

int eb200_init(int device) {
    int n = 0;
    int rc = eb200_device_count(&n);
    if (rc != EB200_OK) return rc;
    if (n == 0) return fail(EB200_ERR_NO_DEVICE, "eb200_init: no CUDA device is visible; this backend has no CPU fallback");
    if (device < 0 || device >= n) return fail(EB200_ERR_INVALID_ARGUMENT, "eb200_init: device %d out of range [0,%d)", device, n);
    EB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    EB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        return fail(EB200_ERR_UNSUPPORTED, "eb200_init: device %d is sm_%d%d; libeinsums_b200 carries sm_100a code only", device, prop.major,
                    prop.minor);
    }
    EB_CUDA(cudaFree(nullptr)); // same context warm-up the reference does (GPUStreams.hip:60)
    (void)sm_count();
    return EB200_OK;
}

int eb200_finalize(void) {
    // Unlike the reference (hipDeviceReset, GPUStreams.hip:134) we only drain: other libraries
    // (torch, NCCL) may share the context.
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess && e != cudaErrorNoDevice) return cuda_fail(e, "cudaDeviceSynchronize");
    return EB200_OK;
}

int eb200_stream_create(void **stream) {
    if (!stream) return fail(EB200_ERR_INVALID_ARGUMENT, "eb200_stream_create: NULL");
    cudaStream_t s;
    EB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (void *)s;
    return EB200_OK;
}
int eb200_stream_destroy(void *stream) {
    EB_CUDA(cudaStreamDestroy(as_stream(stream)));
    return EB200_OK;
}
int eb200_stream_synchronize(void *stream) {
    EB_CUDA(cudaStreamSynchronize(as_stream(stream)));
    return EB200_OK;
}
int eb200_device_synchronize(void) {
    EB_CUDA(cudaDeviceSynchronize());
    return EB200_OK;
}

int eb200_event_create(void **event) {
    if (!event) return fail(EB200_ERR_INVALID_ARGUMENT, "eb200_event_create: NULL");
    cudaEvent_t e;
    EB_CUDA(cudaEventCreate(&e));
    *event = (void *)e;
    return EB200_OK;
}
int eb200_event_destroy(void *event) {
    EB_CUDA(cudaEventDestroy((cudaEvent_t)event));
    return EB200_OK;
}
int eb200_event_record(void *event, void *stream) {
    EB_CUDA(cudaEventRecord((cudaEvent_t)event, as_stream(stream)));
    return EB200_OK;
}
int eb200_event_synchronize(void *event) {
    EB_CUDA(cudaEventSynchronize((cudaEvent_t)event));
    return EB200_OK;
}
int eb200_event_elapsed_ms(void *start, void *stop, float *ms) {
    if (!ms) return fail(EB200_ERR_INVALID_ARGUMENT, "eb200_event_elapsed_ms: NULL");
    EB_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return EB200_OK;
}

int eb200_malloc(void **ptr, size_t bytes) {
    if (!ptr) return fail(EB200_ERR_INVALID_ARGUMENT, "eb200_malloc: NULL");
    *ptr = nullptr;
    if (bytes == 0) return EB200_OK;
    Budget &b = budget_state();
    {
        std::lock_guard<std::mutex> lock(b.mu);
        if (b.budget && (bytes > b.budget || b.in_use > b.budget - bytes))
            return fail(EB200_ERR_OUT_OF_MEMORY, "Could not allocate enough memory on the GPU device. Requested %zu bytes (budget %zu, in use %zu).",
                        bytes, b.budget, b.in_use);
        b.in_use += bytes; // reserved before the CUDA call so that concurrent callers cannot overshoot the budget
    }
    cudaError_t e = cudaMalloc(ptr, bytes);
    std::lock_guard<std::mutex> lock(b.mu);
    if (e != cudaSuccess) {
        b.in_use -= bytes;
        *ptr = nullptr;
        return cuda_fail(e, "cudaMalloc");
    }
    b.live[*ptr] = bytes;
    if (b.in_use > b.peak) b.peak = b.in_use;
    return EB200_OK;
}
int eb200_free(void *ptr) {
    if (!ptr) return EB200_OK;
    Budget &b = budget_state();
    {
        std::lock_guard<std::mutex> lock(b.mu);
        auto it = b.live.find(ptr);
        if (it != b.live.end()) {
            b.in_use = b.in_use < it->second ? 0 : b.in_use - it->second;
            b.live.erase(it);
        }
    }
    EB_CUDA(cudaFree(ptr));
    return EB200_OK;
}
int eb200_set_memory_budget(size_t bytes) {
    Budget                     &b = budget_state();
    std::lock_guard<std::mutex> lock(b.mu);
    b.budget = bytes;
    return EB200_OK;
}
int eb200_memory_stats(size_t *budget_bytes, size_t *in_use_bytes, size_t *peak_bytes, size_t *live_allocations) {
    Budget                     &b = budget_state();
    std::lock_guard<std::mutex> lock(b.mu);
    if (budget_bytes) *budget_bytes = b.budget;
    if (in_use_bytes) *in_use_bytes = b.in_use;
    if (peak_bytes) *peak_bytes = b.peak;
    if (live_allocations) *live_allocations = b.live.size();
    return EB200_OK;
}
int eb200_malloc_host(void **ptr, size_t bytes) {
    if (!ptr) return fail(EB200_ERR_INVALID_ARGUMENT, "eb200_malloc_host: NULL");
    *ptr = nullptr;
    if (bytes == 0) return EB200_OK;
    EB_CUDA(cudaMallocHost(ptr, bytes));
    return EB200_OK;
}
int eb200_free_host(void *ptr) {
    if (!ptr) return EB200_OK;
    EB_CUDA(cudaFreeHost(ptr));
    return EB200_OK;
}
int eb200_host_register(void *ptr, size_t bytes, void **dev_ptr) {
    if (!ptr || !dev_ptr) return fail(EB200_ERR_INVALID_ARGUMENT, "eb200_host_register: NULL");
    *dev_ptr = nullptr;
    if (bytes == 0) return EB200_OK;
    EB_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterMapped));
    cudaError_t e = cudaHostGetDevicePointer(dev_ptr, ptr, 0);
    if (e != cudaSuccess) {
        cudaHostUnregister(ptr);
        return cuda_fail(e, "cudaHostGetDevicePointer");
    }
    return EB200_OK;
}
int eb200_host_unregister(void *ptr) {
    if (!ptr) return EB200_OK;
    EB_CUDA(cudaHostUnregister(ptr));
    return EB200_OK;
}
int eb200_host_get_device_pointer(void *host_ptr, void **dev_ptr) {
    if (!host_ptr || !dev_ptr) return fail(EB200_ERR_INVALID_ARGUMENT, "eb200_host_get_device_pointer: NULL");
    EB_CUDA(cudaHostGetDevicePointer(dev_ptr, host_ptr, 0));
    return EB200_OK;
}
int eb200_ipc_get_handle(void *dev_ptr, void *handle) {
    if (!dev_ptr || !handle) return fail(EB200_ERR_INVALID_ARGUMENT, "eb200_ipc_get_handle: NULL");
    static_assert(sizeof(cudaIpcMemHandle_t) <= EB200_IPC_HANDLE_BYTES, "IPC handle does not fit");
    cudaIpcMemHandle_t h;
    EB_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
    memset(handle, 0, EB200_IPC_HANDLE_BYTES);
    memcpy(handle, &h, sizeof(h));
    return EB200_OK;
}
int eb200_ipc_open_handle(const void *handle, void **peer_ptr) {
    if (!handle || !peer_ptr) return fail(EB200_ERR_INVALID_ARGUMENT, "eb200_ipc_open_handle: NULL");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    EB_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return EB200_OK;
}
int eb200_ipc_close_handle(void *peer_ptr) {
    if (!peer_ptr) return EB200_OK;
    EB_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return EB200_OK;
}
int eb200_mem_get_info(size_t *free_bytes, size_t *total_bytes) {
    size_t f = 0, t = 0;
    EB_CUDA(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return EB200_OK;
}

int eb200_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream) {
    if (bytes == 0) return EB200_OK;
    EB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, as_stream(stream)));
    EB_CUDA(cudaStreamSynchronize(as_stream(stream)));
    return EB200_OK;
}
int eb200_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream) {
    if (bytes == 0) return EB200_OK;
    EB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, as_stream(stream)));
    EB_CUDA(cudaStreamSynchronize(as_stream(stream)));
    return EB200_OK;
}
int eb200_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream) {
    if (bytes == 0) return EB200_OK;
    EB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, as_stream(stream)));
    return EB200_OK;
}
int eb200_memcpy_h2d_async(void *dst, const void *src, size_t bytes, void *stream) {
    if (bytes == 0) return EB200_OK;
    EB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, as_stream(stream)));
    return EB200_OK;
}
int eb200_memcpy_d2h_async(void *dst, const void *src, size_t bytes, void *stream) {
    if (bytes == 0) return EB200_OK;
    EB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, as_stream(stream)));
    return EB200_OK;
}
int eb200_memset_zero(void *dst, size_t bytes, void *stream) {
    if (bytes == 0) return EB200_OK;
    EB_CUDA(cudaMemsetAsync(dst, 0, bytes, as_stream(stream)));
    return EB200_OK;
}

int eb200_set_option(const char *name, int value) {
    if (!name) return fail(EB200_ERR_INVALID_ARGUMENT, "eb200_set_option: name is NULL");
    for (int i = 0; i < OPT_COUNT; i++)
        if (!strcmp(name, kOptionNames[i])) {
            options().v[i].store(value, std::memory_order_relaxed);
            if (i == OPT_SCRATCH_KEEP_MB) { // live pools follow the new threshold
                ScratchPools &sp = scratch_pools();
                std::lock_guard<std::mutex> lock(sp.mu);
                uint64_t keep = (uint64_t)std::max(0, value) << 20;
                for (int dv = 0; dv < 64; dv++)
                    if (sp.pool[dv]) {
                        EB_CUDA(cudaMemPoolSetAttribute(sp.pool[dv], cudaMemPoolAttrReleaseThreshold, &keep));
                        sp.keep[dv] = keep;
                    }
            }
            return EB200_OK;
        }
    return fail(EB200_ERR_INVALID_ARGUMENT, "eb200_set_option: unknown option '%s'", name);
}
int eb200_get_option(const char *name, int *value) {
    if (!name || !value) return fail(EB200_ERR_INVALID_ARGUMENT, "eb200_get_option: NULL argument");
    for (int i = 0; i < OPT_COUNT; i++)
        if (!strcmp(name, kOptionNames[i])) {
            *value = option(i);
            return EB200_OK;
        }
    return fail(EB200_ERR_INVALID_ARGUMENT, "eb200_get_option: unknown option '%s'", name);
}
const char *eb200_last_kernel_name(void) { return t_last_kernel; }

int64_t eb200_launch_count(void) { return launch_counter(); }
void    eb200_launch_count_reset(void) { launch_counter() = 0; }

} // extern "C"
