// host_path.cu — eb200_contract_host: a contraction whose tensors may live in HOST memory.
// Replaces the reference's host-tensor offload (impl_gemm_gpu tiling in 500-blocks with a GPU cache,
// LinearAlgebra/Bases/gemm.hpp:248-486, TensorImpl.hpp:1676-1776): C is cut into slabs along its slowest axis (an
// uncontracted index) and three streams overlap   H2D(slab s+1) | contract(slab s) | D2H(slab s-1)   over a ring of device
// slab buffers, so a PCIe-bound call costs ~max(H2D, D2H). Operands found in host memory are uploaded once per call;
// operands already on the device are used in place. Nothing is ever permuted.
#include "common.cuh"
#include "contract.cuh"

#include <algorithm>
#include <cstring>

namespace eb200 {
namespace {

size_t elem_size(int dt) { return dt == EB200_F32 ? 4 : (dt == EB200_C128 ? 16 : 8); }

bool on_device(const void *p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

/// grow-only device scratch + streams + events of the calling host thread
struct HostPathState {
    static constexpr int RING = 3;
    int          device = -1;
    void        *opA = nullptr, *opB = nullptr, *ring[RING] = {};
    size_t       capA = 0, capB = 0, capRing = 0;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t  ev_start = nullptr, ev_ops = nullptr, ev_in[RING] = {}, ev_cmp[RING] = {}, ev_out[RING] = {};

    int ensure() {
        int dev = -1;
        EB_CUDA(cudaGetDevice(&dev));
        if (dev == device) return EB200_OK;
        release();
        EB_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
        EB_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
        EB_CUDA(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
        EB_CUDA(cudaEventCreateWithFlags(&ev_ops, cudaEventDisableTiming));
        for (int i = 0; i < RING; i++) {
            EB_CUDA(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
            EB_CUDA(cudaEventCreateWithFlags(&ev_cmp[i], cudaEventDisableTiming));
            EB_CUDA(cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming));
        }
        device = dev;
        return EB200_OK;
    }
    static int grow(void *&p, size_t &cap, size_t need) {
        if (need <= cap) return EB200_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, need);
        if (e != cudaSuccess) return cuda_fail(e, "eb200_contract_host: device scratch");
        cap = need;
        return EB200_OK;
    }
    void release() {
        if (device < 0) return;
        cudaFree(opA); cudaFree(opB);
        for (int i = 0; i < RING; i++) {
            cudaFree(ring[i]);
            if (ev_in[i]) cudaEventDestroy(ev_in[i]);
            if (ev_cmp[i]) cudaEventDestroy(ev_cmp[i]);
            if (ev_out[i]) cudaEventDestroy(ev_out[i]);
            ring[i] = nullptr; ev_in[i] = ev_cmp[i] = ev_out[i] = nullptr;
        }
        if (ev_start) cudaEventDestroy(ev_start);
        if (ev_ops) cudaEventDestroy(ev_ops);
        if (s_in) cudaStreamDestroy(s_in);
        if (s_out) cudaStreamDestroy(s_out);
        opA = opB = nullptr; capA = capB = capRing = 0; ev_start = ev_ops = nullptr; s_in = s_out = nullptr;
        device = -1;
    }
};
HostPathState &state() {
    static thread_local HostPathState st;
    return st;
}

/// dense row-major check; returns the element count
bool dense_row_major(const eb200_tensor_desc &t, const int64_t *extent, int64_t &elems) {
    int64_t acc = 1;
    for (int i = t.rank - 1; i >= 0; i--) {
        if (t.stride[i] != acc) return false;
        acc *= extent[t.index[i]];
    }
    elems = acc;
    return true;
}

} // namespace
} // namespace eb200

extern "C" {

int eb200_contract_host(void *stream, const eb200_contract_desc *desc, size_t slab_bytes, eb200_host_stats *stats) {
    using namespace eb200;
    if (!desc) return fail(EB200_ERR_INVALID_ARGUMENT, "eb200_contract_host: desc is NULL");
    eb200_contract_desc d = *desc;
    if (stats) memset(stats, 0, sizeof(*stats));
    const bool c_dev = on_device(d.C.data), a_dev = on_device(d.A.data), b_dev = on_device(d.B.data);
    cudaStream_t s_cmp = as_stream(stream);
    if (c_dev && a_dev && b_dev) { // nothing to stage
        int rc = contract_dispatch(s_cmp, d);
        if (rc == EB200_OK) EB_CUDA(cudaStreamSynchronize(s_cmp));
        return rc;
    }
    for (int u = 0; u < d.n_unique; u++)
        if (d.extent[u] == 0) return EB200_OK;
    int64_t nA = 0, nB = 0, nC = 0;
    if (!a_dev && !dense_row_major(d.A, d.extent, nA)) return fail(EB200_ERR_UNSUPPORTED, "eb200_contract_host: a host-resident A must be dense row-major");
    if (!b_dev && !dense_row_major(d.B, d.extent, nB)) return fail(EB200_ERR_UNSUPPORTED, "eb200_contract_host: a host-resident B must be dense row-major");
    HostPathState &st = state();
    int            rc = st.ensure();
    if (rc) return rc;
    const size_t esA = elem_size(d.A.dtype), esB = elem_size(d.B.dtype), esC = elem_size(d.C.dtype);
    const bool   beta_zero = d.beta[0] == 0.0 && d.beta[1] == 0.0;

    EB_CUDA(cudaEventRecord(st.ev_start, s_cmp));
    EB_CUDA(cudaStreamWaitEvent(st.s_in, st.ev_start, 0));
    EB_CUDA(cudaStreamWaitEvent(st.s_out, st.ev_start, 0));
    if (!a_dev) {
        if ((rc = HostPathState::grow(st.opA, st.capA, (size_t)nA * esA))) return rc;
        EB_CUDA(cudaMemcpyAsync(st.opA, d.A.data, (size_t)nA * esA, cudaMemcpyHostToDevice, st.s_in));
        d.A.data = st.opA;
        if (stats) stats->h2d_bytes += (int64_t)((size_t)nA * esA);
    }
    if (!b_dev) {
        if ((rc = HostPathState::grow(st.opB, st.capB, (size_t)nB * esB))) return rc;
        EB_CUDA(cudaMemcpyAsync(st.opB, d.B.data, (size_t)nB * esB, cudaMemcpyHostToDevice, st.s_in));
        d.B.data = st.opB;
        if (stats) stats->h2d_bytes += (int64_t)((size_t)nB * esB);
    }
    EB_CUDA(cudaEventRecord(st.ev_ops, st.s_in));
    EB_CUDA(cudaStreamWaitEvent(s_cmp, st.ev_ops, 0));

    if (c_dev) { // only operands were on the host
        rc = contract_dispatch(s_cmp, d);
        if (rc) return rc;
        EB_CUDA(cudaStreamSynchronize(s_cmp));
        if (stats) stats->slabs = 1;
        return EB200_OK;
    }

    // ---- C on the host: slabs along its slowest axis -------------------------------------------------------------
    if (d.C.rank < 1) { // a scalar result: one element through a stream-ordered temporary
        void *host_c = d.C.data, *tmp = nullptr;
        if ((rc = scratch_alloc(&tmp, 16, s_cmp))) return rc;
        if (!beta_zero) EB_CUDA(cudaMemcpyAsync(tmp, host_c, esC, cudaMemcpyHostToDevice, s_cmp));
        d.C.data = tmp;
        rc = contract_dispatch(s_cmp, d);
        if (rc == EB200_OK) {
            cudaError_t e = cudaMemcpyAsync(host_c, tmp, esC, cudaMemcpyDeviceToHost, s_cmp);
            if (e != cudaSuccess) rc = cuda_fail(e, "eb200_contract_host: D2H of the scalar");
        }
        cudaFreeAsync(tmp, s_cmp);
        cudaError_t e = cudaStreamSynchronize(s_cmp);
        if (rc) return rc;
        if (e != cudaSuccess) return cuda_fail(e, "eb200_contract_host: compute stream");
        if (stats) { stats->slabs = 1; stats->d2h_bytes += (int64_t)esC; if (!beta_zero) stats->h2d_bytes += (int64_t)esC; }
        return EB200_OK;
    }
    if (!dense_row_major(d.C, d.extent, nC)) return fail(EB200_ERR_UNSUPPORTED, "eb200_contract_host: a host-resident C must be dense row-major");
    const int u0 = d.C.index[0];
    int       in_c = 0, in_a = 0, in_b = 0, ax_a = -1, ax_b = -1;
    for (int i = 0; i < d.C.rank; i++) in_c += d.C.index[i] == u0;
    for (int i = 0; i < d.A.rank; i++) if (d.A.index[i] == u0) { in_a++; ax_a = i; }
    for (int i = 0; i < d.B.rank; i++) if (d.B.index[i] == u0) { in_b++; ax_b = i; }
    if (in_c != 1 || in_a + in_b != 1)
        return fail(EB200_ERR_UNSUPPORTED, "eb200_contract_host: the slowest index of C must be an uncontracted index carried by exactly one operand");
    const int64_t n = d.extent[u0], row_elems = d.C.stride[0];
    const int64_t rows = std::max<int64_t>(1, std::min<int64_t>(n, (int64_t)(slab_bytes ? slab_bytes : ((size_t)256 << 20)) / std::max<int64_t>(1, row_elems * (int64_t)esC)));
    const size_t  slab_cap = (size_t)rows * (size_t)row_elems * esC;
    if (slab_cap > st.capRing) {
        for (int i = 0; i < HostPathState::RING; i++) {
            size_t cap = st.capRing;
            if ((rc = HostPathState::grow(st.ring[i], cap, slab_cap))) { st.capRing = 0; return rc; }
        }
        st.capRing = slab_cap;
    }
    char       *host_c = static_cast<char *>(d.C.data);
    char       *baseA = static_cast<char *>(d.A.data), *baseB = static_cast<char *>(d.B.data);
    bool        used[HostPathState::RING] = {};
    int64_t     nslab = 0;
    for (int64_t lo = 0; lo < n && rc == EB200_OK; lo += rows, nslab++) {
        const int64_t hi = std::min(n, lo + rows);
        const int     slot = (int)(nslab % HostPathState::RING);
        const size_t  bytes = (size_t)(hi - lo) * (size_t)row_elems * esC;
        if (used[slot]) EB_CUDA(cudaStreamWaitEvent(st.s_in, st.ev_out[slot], 0)); // the slot's previous D2H has finished
        if (!beta_zero) {
            EB_CUDA(cudaMemcpyAsync(st.ring[slot], host_c + (size_t)lo * (size_t)row_elems * esC, bytes, cudaMemcpyHostToDevice, st.s_in));
            if (stats) stats->h2d_bytes += (int64_t)bytes;
        }
        EB_CUDA(cudaEventRecord(st.ev_in[slot], st.s_in));
        EB_CUDA(cudaStreamWaitEvent(s_cmp, st.ev_in[slot], 0));
        eb200_contract_desc ds = d;
        ds.extent[u0] = hi - lo;
        ds.C.data = st.ring[slot];
        if (ax_a >= 0) ds.A.data = baseA + (size_t)lo * (size_t)d.A.stride[ax_a] * esA;
        if (ax_b >= 0) ds.B.data = baseB + (size_t)lo * (size_t)d.B.stride[ax_b] * esB;
        rc = contract_dispatch(s_cmp, ds);
        if (rc) break;
        EB_CUDA(cudaEventRecord(st.ev_cmp[slot], s_cmp));
        EB_CUDA(cudaStreamWaitEvent(st.s_out, st.ev_cmp[slot], 0));
        EB_CUDA(cudaMemcpyAsync(host_c + (size_t)lo * (size_t)row_elems * esC, st.ring[slot], bytes, cudaMemcpyDeviceToHost, st.s_out));
        EB_CUDA(cudaEventRecord(st.ev_out[slot], st.s_out));
        used[slot] = true;
        if (stats) stats->d2h_bytes += (int64_t)bytes;
    }
    // host-synchronous: the result is in host memory when the call returns (also drains after an error)
    cudaError_t e1 = cudaStreamSynchronize(st.s_out), e2 = cudaStreamSynchronize(s_cmp), e3 = cudaStreamSynchronize(st.s_in);
    if (rc) return rc;
    if (e1 != cudaSuccess) return cuda_fail(e1, "eb200_contract_host: D2H stream");
    if (e2 != cudaSuccess) return cuda_fail(e2, "eb200_contract_host: compute stream");
    if (e3 != cudaSuccess) return cuda_fail(e3, "eb200_contract_host: H2D stream");
    if (stats) stats->slabs = nslab;
    return EB200_OK;
}

int eb200_contract_host_release(void) {
    eb200::state().release();
    return EB200_OK;
}

} // extern "C"
