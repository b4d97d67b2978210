// Runtime.hpp — einsums::gpu runtime of the host shim: one non-blocking CUDA stream per host thread, created on
// first use, and the synchronous-by-default contract of the reference (every op ends in stream_wait()).
// Replaces libs/Einsums/GPUStreams (reference: include/Einsums/GPUStreams/GPUStreams.hpp:194-280, src/GPUStreams.hip:58-263).
// There is no cuBLAS/cuSOLVER handle: the backend is self-contained.
#pragma once
#include <Einsums/B200/Errors.hpp>

#include <atomic>

namespace einsums {
namespace gpu {

namespace detail {
struct ThreadStream {
    void *stream = nullptr;
    bool  skip_next = false;
    ~ThreadStream() {
        if (stream) eb200_stream_destroy(stream);
    }
};
inline ThreadStream &thread_stream() {
    static thread_local ThreadStream ts;
    return ts;
}
inline std::atomic<bool> &async_mode() {
    static std::atomic<bool> flag{false};
    return flag;
}
inline std::atomic<int> &device_index() {
    static std::atomic<int> dev{-1};
    return dev;
}
} // namespace detail

/// einsums::initialize() equivalent for this path (reference: initialize_Einsums_GPUStreams, GPUStreams.hip:58-102).
inline void initialize(int device = 0) {
    EB200_CHECK(eb200_init(device));
    detail::device_index() = device;
}
inline void finalize() { EB200_CHECK(eb200_finalize()); }

/// The calling thread's stream (reference: get_stream() by omp_get_thread_num()).
inline void *get_stream() {
    auto &ts = detail::thread_stream();
    if (!ts.stream) {
        if (detail::device_index() < 0) initialize(0);
        else EB200_CHECK(eb200_set_device(detail::device_index()));
        EB200_CHECK(eb200_stream_create(&ts.stream));
    }
    return ts.stream;
}

/// Ops return only after their kernels finished (reference behaviour) unless async mode is on (benchmarks).
inline void set_async(bool on) { detail::async_mode() = on; }
inline bool is_async() { return detail::async_mode(); }

inline void skip_next_wait() { detail::thread_stream().skip_next = true; }

inline void stream_wait(bool may_skip = false) {
    auto &ts = detail::thread_stream();
    if (may_skip && (ts.skip_next || is_async())) {
        ts.skip_next = false;
        return;
    }
    if (ts.stream) EB200_CHECK(eb200_stream_synchronize(ts.stream));
}
/// wait at the end of an op: elided in async mode
inline void op_wait() { stream_wait(true); }
inline void device_synchronize() { EB200_CHECK(eb200_device_synchronize()); }
inline void all_stream_wait() { device_synchronize(); }

} // namespace gpu

inline void initialize(int device = 0) { gpu::initialize(device); }
inline void finalize() { gpu::finalize(); }

} // namespace einsums
