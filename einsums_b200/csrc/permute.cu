// permute.cu — tensor transpose C[perm] = beta*C + alpha*op(A) for any rank and any axis order.
// Replaces tensor_algebra::detail::gpu_permute (LibreTT plan+execute,
// reference libs/Einsums/TensorAlgebra/src/DevicePermute.hip:20-59) and permute_kernel
// (TensorAlgebra/Backends/DevicePermute.hpp:34-73), plus eb200_copy_strided for the layout-changing
// copies of Tensor/Backends/DeviceTensor.hpp:295-402.
//
// Method: axes are sorted by A stride and merged where A and C agree on contiguity; then
//   * if the fastest axis of A is also the fastest axis of C -> row kernel (both sides coalesced,
//     16-byte accesses when alignment allows);
//   * otherwise -> 2-D tile transpose through shared memory over the plane spanned by
//     (fastest axis of A, fastest axis of C); both the read and the write are coalesced, all other
//     axes are decoded once per tile. FP64 overwrites (beta == 0) stage the tile with the Tensor Memory
//     Accelerator (permute_tma_f64: one <= 5-D tensor map over A, 128-byte-swizzled boxes, tile extents
//     fitted to the axis extents so that rows stay long: 80 x 80 for 160, 96 x 96 for 96).
#include "plan.cuh"
#include "tma_util.cuh"

#include <algorithm>
#include <vector>

namespace eb200 {
namespace {

struct PermArgs {
    Grp         rest;   // remaining axes: strides ROLE_A / ROLE_C
    int64_t     ea, sa_a, sc_a; // axis "a": fastest in A   (extent, A stride, C stride)
    int64_t     ec, sa_c, sc_c; // axis "c": fastest in C
    const void *A;
    void       *C;
    double      alpha[2], beta[2];
    int64_t     tiles_a, tiles_c;
    uint32_t    cpr_mul, cpr_shift; // fast division by chunks-per-row (row kernel)
};

template <typename T, bool CONJ> __device__ __forceinline__ T perm_out(T alpha, T a, T beta, T *dst, bool has_beta) {
    T v = mul_(alpha, maybe_conj<CONJ>(a));
    if (has_beta) v = fma_(beta, *dst, v);
    return v;
}

// plane transpose: TILE x TILE elements per block, block = (TILE, ROWS)
template <typename T, bool CONJ, int TILE, int ROWS> __global__ void __launch_bounds__(TILE *ROWS) permute_tile(const __grid_constant__ PermArgs a) {
    __shared__ T tile[TILE][TILE + 1];
    const T     *A = reinterpret_cast<const T *>(a.A);
    T           *C = reinterpret_cast<T *>(a.C);
    const T      alpha = make_scalar<T>(a.alpha[0], a.alpha[1]), beta = make_scalar<T>(a.beta[0], a.beta[1]);
    const bool   has_beta = !is_zero(beta);
    const int    tx = threadIdx.x, ty = threadIdx.y;
    const int64_t per_rest = a.tiles_a * a.tiles_c;
    const int64_t ntiles = per_rest * a.rest.tot;

    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int64_t rest = t / per_rest, in = t - rest * per_rest;
        int64_t tc = in / a.tiles_a, ta = in - tc * a.tiles_a;
        int64_t offA, offC;
        grp_offset2(a.rest, rest, ROLE_A, ROLE_C, offA, offC);
        const int64_t a0 = ta * TILE, c0 = tc * TILE;
        // read: lanes along axis a (contiguous in A), rows along axis c
        {
            const int64_t ia = a0 + tx;
#pragma unroll
            for (int r = ty; r < TILE; r += ROWS) {
                int64_t ic = c0 + r;
                if (ia < a.ea && ic < a.ec) tile[r][tx] = A[offA + ia * a.sa_a + ic * a.sa_c];
            }
        }
        __syncthreads();
        // write: lanes along axis c (contiguous in C), rows along axis a
        {
            const int64_t ic = c0 + tx;
#pragma unroll
            for (int r = ty; r < TILE; r += ROWS) {
                int64_t ia = a0 + r;
                if (ia < a.ea && ic < a.ec) {
                    T *dst = C + offC + ia * a.sc_a + ic * a.sc_c;
                    *dst = perm_out<T, CONJ>(alpha, tile[tx][r], beta, dst, has_beta);
                }
            }
        }
        __syncthreads();
    }
}

// 16-byte plane transpose for 4/8-byte real types: 64 x 64 elements per tile, 256 threads.
// Phase 1: every thread issues 8 independent 2-element loads along axis a (all in flight before the
// first use); the pairs are scattered into an XOR-swizzled tile[a][c]. Phase 2: 2-element reads along
// axis c, 8 independent stores. Needs even extents/strides and 2-element-aligned bases (checked on the host).
template <typename T> struct alignas(2 * sizeof(T)) Pair { T x, y; };

// TILE = 64: 8 pairs per thread; TILE = 32 (extents that fill 64-tiles badly, e.g. 96 or 160): 2 pairs per thread, more CTAs per SM.
template <typename T, bool HAS_BETA, int TILE> __global__ void __launch_bounds__(256, TILE == 64 ? (HAS_BETA ? 2 : 3) : 6) permute_tile_vec(const __grid_constant__ PermArgs a) {
    constexpr int LX = TILE / 2, ROWS = 256 / LX, NI = TILE / ROWS; // lanes along the pair direction, rows per pass, passes
    __shared__ __align__(16) T tile[TILE * TILE];
    const T      *A = reinterpret_cast<const T *>(a.A);
    T            *C = reinterpret_cast<T *>(a.C);
    const T       alpha = make_scalar<T>(a.alpha[0], a.alpha[1]), beta = make_scalar<T>(a.beta[0], a.beta[1]);
    constexpr bool has_beta = HAS_BETA;
    const bool    plain = is_one(alpha) && !has_beta;
    const int     tid = threadIdx.x, lx = tid % LX, ly = tid / LX;
    const int64_t per_rest = a.tiles_a * a.tiles_c;
    const int64_t ntiles = per_rest * a.rest.tot;
    // swizzle: element (ia, ic) lives at tile[ia*TILE + (ic ^ sw(ia))], sw(ia) = 2*((ia>>1)&7) keeps (even,odd) pairs together
    auto sw = [](int ia) { return ((ia >> 1) & 7) << 1; };

    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int64_t rest, ta, tc;
        if (ntiles <= 0x7fffffffLL) {
            uint32_t tt = (uint32_t)t, pr = (uint32_t)per_rest, r_ = tt / pr, in = tt - r_ * pr;
            uint32_t c_ = in / (uint32_t)a.tiles_a;
            rest = r_; tc = c_; ta = in - c_ * (uint32_t)a.tiles_a;
        } else {
            rest = t / per_rest;
            int64_t in = t - rest * per_rest;
            tc = in / a.tiles_a; ta = in - tc * a.tiles_a;
        }
        int64_t offA, offC;
        grp_offset2(a.rest, rest, ROLE_A, ROLE_C, offA, offC);
        const int64_t a0 = ta * TILE, c0 = tc * TILE;
        Pair<T> v[NI], old[HAS_BETA ? NI : 1];
        if constexpr (HAS_BETA) { // the old C values of this thread's phase-2 pairs are fetched together with the A tile
            const int64_t ic = c0 + 2 * lx;
#pragma unroll
            for (int i = 0; i < NI; i++) {
                const int64_t ia = a0 + ly + ROWS * i;
                if (ia < a.ea && ic < a.ec) old[i] = *reinterpret_cast<const Pair<T> *>(C + offC + ia * a.sc_a + ic);
            }
        }
        {
            const int64_t ia = a0 + 2 * lx;
#pragma unroll
            for (int i = 0; i < NI; i++) {
                const int64_t ic = c0 + ly + ROWS * i;
                if (ia < a.ea && ic < a.ec) v[i] = *reinterpret_cast<const Pair<T> *>(A + offA + ia + ic * a.sa_c);
            }
#pragma unroll
            for (int i = 0; i < NI; i++) {
                const int icl = ly + ROWS * i, ial = 2 * lx;
                tile[ial * TILE + (icl ^ sw(ial))] = v[i].x;
                tile[(ial + 1) * TILE + (icl ^ sw(ial + 1))] = v[i].y;
            }
        }
        __syncthreads();
        {
            const int64_t ic = c0 + 2 * lx;
#pragma unroll
            for (int i = 0; i < NI; i++) {
                const int     ial = ly + ROWS * i;
                const int64_t ia = a0 + ial;
                if (ia < a.ea && ic < a.ec) {
                    Pair<T>  w = *reinterpret_cast<const Pair<T> *>(&tile[ial * TILE + ((2 * lx) ^ sw(ial))]);
                    Pair<T> *dst = reinterpret_cast<Pair<T> *>(C + offC + ia * a.sc_a + ic);
                    if (!plain) {
                        w.x = mul_(alpha, w.x); w.y = mul_(alpha, w.y);
                        if constexpr (HAS_BETA) { w.x = fma_(beta, old[i].x, w.x); w.y = fma_(beta, old[i].y, w.y); }
                    }
                    *dst = w;
                }
            }
        }
        __syncthreads();
    }
}

// beta == 0 and extents that fill 64-tiles badly (96, 160 ...): U consecutive 32 x 32 tiles per CTA. All U*2 loads of a thread are
// issued before the first shared-memory store, so a CTA keeps as many bytes in flight as the 64-tile kernel while every tile is full.
template <typename T, int U> __global__ void __launch_bounds__(256, 3) permute_tile32_multi(const __grid_constant__ PermArgs a) {
    constexpr int TILE = 32, LX = 16, ROWS = 16, NI = 2;
    __shared__ __align__(16) T tile[U][TILE * TILE];
    const T      *A = reinterpret_cast<const T *>(a.A);
    T            *C = reinterpret_cast<T *>(a.C);
    const T       alpha = make_scalar<T>(a.alpha[0], a.alpha[1]);
    const bool    plain = is_one(alpha);
    const int     tid = threadIdx.x, lx = tid % LX, ly = tid / LX;
    const int64_t per_rest = a.tiles_a * a.tiles_c;
    const int64_t ntiles = per_rest * a.rest.tot;
    auto sw = [](int ia) { return ((ia >> 1) & 7) << 1; };

    const int64_t t0 = (int64_t)blockIdx.x * U;
    int64_t       offA[U], offC[U], a0[U], c0[U];
    Pair<T>       v[U][NI];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int64_t t = t0 + u;
        a0[u] = c0[u] = a.ea; // out of range => nothing is loaded or stored
        offA[u] = offC[u] = 0;
        if (t >= ntiles) continue;
        int64_t rest, ta, tc;
        if (ntiles <= 0x7fffffffLL) {
            uint32_t tt = (uint32_t)t, pr = (uint32_t)per_rest, r_ = tt / pr, in = tt - r_ * pr;
            uint32_t c_ = in / (uint32_t)a.tiles_a;
            rest = r_; tc = c_; ta = in - c_ * (uint32_t)a.tiles_a;
        } else {
            rest = t / per_rest;
            int64_t in = t - rest * per_rest;
            tc = in / a.tiles_a; ta = in - tc * a.tiles_a;
        }
        grp_offset2(a.rest, rest, ROLE_A, ROLE_C, offA[u], offC[u]);
        a0[u] = ta * TILE; c0[u] = tc * TILE;
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int64_t ia = a0[u] + 2 * lx;
#pragma unroll
        for (int i = 0; i < NI; i++) {
            const int64_t ic = c0[u] + ly + ROWS * i;
            if (ia < a.ea && ic < a.ec) v[u][i] = *reinterpret_cast<const Pair<T> *>(A + offA[u] + ia + ic * a.sa_c);
        }
    }
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
        for (int i = 0; i < NI; i++) {
            const int icl = ly + ROWS * i, ial = 2 * lx;
            tile[u][ial * TILE + (icl ^ sw(ial))] = v[u][i].x;
            tile[u][(ial + 1) * TILE + (icl ^ sw(ial + 1))] = v[u][i].y;
        }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int64_t ic = c0[u] + 2 * lx;
#pragma unroll
        for (int i = 0; i < NI; i++) {
            const int     ial = ly + ROWS * i;
            const int64_t ia = a0[u] + ial;
            if (ia < a.ea && ic < a.ec) {
                Pair<T> w = *reinterpret_cast<const Pair<T> *>(&tile[u][ial * TILE + ((2 * lx) ^ sw(ial))]);
                if (!plain) { w.x = mul_(alpha, w.x); w.y = mul_(alpha, w.y); }
                *reinterpret_cast<Pair<T> *>(C + offC[u] + ia * a.sc_a + ic) = w;
            }
        }
    }
}

// Row-type permutations (the same axis is fastest in A and in C) with at least one more axis: a CTA copies a 64 x 64 patch of the
// plane (fastest axis, one other axis). Every thread issues 8 independent 2-element loads, then 8 stores; the only index decode
// is one per CTA for the remaining axes. No shared memory: nothing is transposed inside a row.
template <typename T, bool HAS_BETA> __global__ void __launch_bounds__(256, 4) permute_rows2d(const __grid_constant__ PermArgs a) {
    const T      *A = reinterpret_cast<const T *>(a.A);
    T            *C = reinterpret_cast<T *>(a.C);
    const T       alpha = make_scalar<T>(a.alpha[0], a.alpha[1]), beta = make_scalar<T>(a.beta[0], a.beta[1]);
    const bool    plain = is_one(alpha) && !HAS_BETA;
    const int     lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int64_t per_rest = a.tiles_a * a.tiles_c;
    const int64_t t = blockIdx.x;
    int64_t       rest, ta, tc;
    {
        uint32_t tt = (uint32_t)t, pr = (uint32_t)per_rest, r_ = tt / pr, in = tt - r_ * pr;
        uint32_t c_ = in / (uint32_t)a.tiles_a;
        rest = r_; tc = c_; ta = in - c_ * (uint32_t)a.tiles_a;
    }
    int64_t offA, offC;
    grp_offset2(a.rest, rest, ROLE_A, ROLE_C, offA, offC);
    const int64_t ia = ta * 64 + 2 * lx, c0 = tc * 64 + ly;
    if (ia >= a.ea) return;
    Pair<T> v[8], o[HAS_BETA ? 8 : 1];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int64_t ic = c0 + 8 * i;
        if (ic < a.ec) {
            v[i] = *reinterpret_cast<const Pair<T> *>(A + offA + ia + ic * a.sa_c);
            if constexpr (HAS_BETA) o[i] = *reinterpret_cast<const Pair<T> *>(C + offC + ia + ic * a.sc_c);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int64_t ic = c0 + 8 * i;
        if (ic < a.ec) {
            Pair<T> w = v[i];
            if (!plain) {
                w.x = mul_(alpha, w.x); w.y = mul_(alpha, w.y);
                if constexpr (HAS_BETA) { w.x = fma_(beta, o[i].x, w.x); w.y = fma_(beta, o[i].y, w.y); }
            }
            *reinterpret_cast<Pair<T> *>(C + offC + ia + ic * a.sc_c) = w;
        }
    }
}

// row kernel: the same axis is fastest on both sides. VEC elements per thread per step.
template <typename T, bool CONJ, int VEC, bool HAS_BETA> __global__ void __launch_bounds__(256, 3) permute_rows(const __grid_constant__ PermArgs a) {
    const T      *A = reinterpret_cast<const T *>(a.A);
    T            *C = reinterpret_cast<T *>(a.C);
    const T       alpha = make_scalar<T>(a.alpha[0], a.alpha[1]), beta = make_scalar<T>(a.beta[0], a.beta[1]);
    constexpr bool has_beta = HAS_BETA;
    const int64_t chunks_per_row = (a.ea + VEC - 1) / VEC;
    const int64_t total = chunks_per_row * a.rest.tot;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    constexpr int UNR = 4;
    struct alignas(sizeof(T) * VEC) Pack { T v[VEC]; };
    auto locate = [&](int64_t x, int64_t &oa, int64_t &oc) {
        int64_t row, ch;
        if (a.rest.tot == 1) { row = 0; ch = x; }
        else if (total <= 0x7fffffffLL) { uint32_t q = fastdiv((uint32_t)x, a.cpr_mul, a.cpr_shift); row = q; ch = (uint32_t)x - q * (uint32_t)chunks_per_row; }
        else { row = x / chunks_per_row; ch = x - row * chunks_per_row; }
        grp_offset2(a.rest, row, ROLE_A, ROLE_C, oa, oc);
        oa += ch * VEC * (VEC > 1 ? 1 : a.sa_a);
        oc += ch * VEC * (VEC > 1 ? 1 : a.sc_a);
    };
    for (int64_t x0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x0 < total; x0 += stride * UNR) {
        int64_t oa[UNR], oc[UNR];
        Pack    in[UNR], old[HAS_BETA ? UNR : 1];
#pragma unroll
        for (int u = 0; u < UNR; u++) {
            int64_t x = x0 + u * stride;
            if (x < total) {
                locate(x, oa[u], oc[u]);
                in[u] = *reinterpret_cast<const Pack *>(A + oa[u]);
                if constexpr (HAS_BETA) old[u] = *reinterpret_cast<const Pack *>(C + oc[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; u++) {
            int64_t x = x0 + u * stride;
            if (x < total) {
                Pack out;
#pragma unroll
                for (int e = 0; e < VEC; e++) {
                    T v = mul_(alpha, maybe_conj<CONJ>(in[u].v[e]));
                    if constexpr (HAS_BETA) v = fma_(beta, old[u].v[e], v);
                    out.v[e] = v;
                }
                *reinterpret_cast<Pack *>(C + oc[u]) = out;
            }
        }
    }
}


// ---------------------------------------------------------------- TMA-staged plane transpose (FP64, beta == 0)
// One CTA moves one TA x TC tile of the (a, c) plane: lane 0 issues TA/16 bulk tensor loads (boxes of 16 a x TC c, 128-byte rows,
// hardware swizzle) against one mbarrier; the 8 warps then read the tile TRANSPOSED (ld.shared.v2.f64 of two consecutive a for 32
// consecutive c: conflict-free under the swizzle) and store rows of C with lanes along c. No thread computes an A address, edge
// tiles are clipped by the TMA unit (out-of-range elements arrive as zeros and are never stored), and tile extents are run-time
// values fitted to the extents (ncu, n = 160: 64-tiles leave 256-byte row fragments and reach 69 % of DRAM cycles; see profiles/r2).
struct PermTmaArgs {
    int64_t  ea, ec, sc_a;     // extents of the plane axes; C stride of axis a (C stride of axis c is 1)
    int64_t  rext[3], rsc[3];  // remaining axes: extents and C strides (tensor-map dimensions 2..4 in the same order)
    double  *C;
    double   alpha;
    int32_t  ta, tc, tiles_a, tiles_c;
};

__global__ void __launch_bounds__(256, 3) permute_tma_f64(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ PermTmaArgs a) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = ((uint32_t)__cvta_generic_to_shared(smem_dyn) + 1023u) & ~1023u;
    const int      tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t box_bytes = (uint32_t)a.tc * 128u, bar = base + (uint32_t)(a.ta >> 4) * box_bytes;

    uint32_t t = blockIdx.x;
    const uint32_t ti_a = t % (uint32_t)a.tiles_a; t /= (uint32_t)a.tiles_a;
    const uint32_t ti_c = t % (uint32_t)a.tiles_c; t /= (uint32_t)a.tiles_c;
    const uint32_t r0 = t % (uint32_t)a.rext[0]; t /= (uint32_t)a.rext[0];
    const uint32_t r1 = t % (uint32_t)a.rext[1]; t /= (uint32_t)a.rext[1];
    const uint32_t r2 = t;
    const int64_t  a0 = (int64_t)ti_a * a.ta, c0 = (int64_t)ti_c * a.tc;
    const int      na = (int)min((int64_t)a.ta, a.ea - a0), nc = (int)min((int64_t)a.tc, a.ec - c0); // valid part of the tile (na is even)

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const int nbox = (na + 15) >> 4;
        mbar_expect_tx(bar, (uint32_t)nbox * box_bytes);
        for (int b = 0; b < nbox; b++) tma_load_5d(base + b * box_bytes, &tmA, bar, (int)a0 + 16 * b, (int)c0, (int)r0, (int)r1, (int)r2);
    }
    __syncthreads();
    double *C = a.C + (int64_t)r0 * a.rsc[0] + (int64_t)r1 * a.rsc[1] + (int64_t)r2 * a.rsc[2] + a0 * a.sc_a + c0;
    const double alpha = a.alpha;
    const bool   plain = alpha == 1.0;
    mbar_wait(bar, 0);
    for (int p = warp; 2 * p < na; p += 8) { // rows 2p, 2p + 1 of the tile (axis a)
        const uint32_t row = base + (uint32_t)(p >> 3) * box_bytes;
        const uint32_t chunk = (uint32_t)(p & 7);
        double        *dst = C + (int64_t)(2 * p) * a.sc_a;
#pragma unroll 4
        for (int c = lane; c < nc; c += 32) {
            double2 v = lds_f64x2(row + (uint32_t)c * 128u + ((chunk ^ ((uint32_t)c & 7u)) << 4));
            if (!plain) { v.x *= alpha; v.y *= alpha; }
            dst[c] = v.x;
            dst[c + a.sc_a] = v.y;
        }
    }
}

// tile extent for an axis of extent e: as few tiles as possible of at most 128, equal sizes, rounded up to a multiple of q
inline int64_t perm_tma_tile(int64_t e, int64_t q) {
    const int64_t nt = (e + 127) / 128, t = (e + nt - 1) / nt;
    return (t + q - 1) / q * q;
}

// returns 1 when the TMA kernel ran, 0 when the layout does not qualify, < 0 on error
int try_permute_tma(cudaStream_t s, const std::vector<int64_t> &ext, const std::vector<int64_t> &sa, const std::vector<int64_t> &sc, size_t ia, size_t ic,
                    const double *A, double *C, double alpha, int mode) {
    EncodeTiledFn enc = encode_tiled_fn();
    const size_t  n = ext.size();
    if (!enc || n > 5 || n < 2) return 0;
    PermTmaArgs pa;
    memset(&pa, 0, sizeof(pa));
    cuuint64_t gdim[5] = {1, 1, 1, 1, 1}, gstr[4];
    cuuint32_t box[5] = {16, 1, 1, 1, 1}, es[5] = {1, 1, 1, 1, 1};
    int64_t    span = 2;
    for (size_t i = 0; i < n; i++) {
        if (ext[i] >= (1LL << 31) || sa[i] <= 0 || sa[i] >= (1LL << 36)) return 0;
        span = std::max(span, sa[i] * ext[i]);
    }
    pa.ea = ext[ia]; pa.ec = ext[ic]; pa.sc_a = sc[ia];
    gdim[0] = (cuuint64_t)ext[ia];
    gdim[1] = (cuuint64_t)ext[ic];
    gstr[0] = (cuuint64_t)sa[ic] * 8;
    int     d = 2;
    int64_t rest_tot = 1;
    for (int k = 0; k < 3; k++) pa.rext[k] = 1;
    for (size_t i = 0; i < n; i++) {
        if (i == ia || i == ic) continue;
        gdim[d] = (cuuint64_t)ext[i];
        gstr[d - 1] = (cuuint64_t)sa[i] * 8;
        pa.rext[d - 2] = ext[i];
        pa.rsc[d - 2] = sc[i];
        rest_tot *= ext[i];
        d++;
    }
    for (; d < 5; d++) gstr[d - 1] = (cuuint64_t)((span + 1) & ~1LL) * 8; // unused dimensions: extent 1, coordinate 0
    for (int k = 0; k < 4; k++)
        if (gstr[k] >= (1ULL << 40) || (gstr[k] & 15)) return 0;
    int64_t ta = perm_tma_tile(pa.ea, 16), tc = perm_tma_tile(pa.ec, 8);
    while (ta * tc > 9216) { // 72 KB: three CTAs per SM
        if (ta >= tc) ta = ((ta / 2) + 15) / 16 * 16;
        else tc = ((tc / 2) + 7) / 8 * 8;
    }
    // measured over the 24 rank-4 orders (profiles/r2): n = 96: 5.09 -> 6.24 TB/s worst order, n = 160: 5.35 -> 5.93, n = 128: 6.14 -> 6.27,
    // n = 64: equal within 2 % -- so whenever the layout has a tensor map this kernel is the one (mode 1 and 2 are the same today)
    (void)mode;
    box[1] = (cuuint32_t)tc;
    pa.ta = (int32_t)ta; pa.tc = (int32_t)tc;
    const int64_t tiles_a = (pa.ea + ta - 1) / ta, tiles_c = (pa.ec + tc - 1) / tc, ntiles = tiles_a * tiles_c * rest_tot;
    if (ntiles > 0x7fffffffLL) return 0;
    pa.tiles_a = (int32_t)tiles_a; pa.tiles_c = (int32_t)tiles_c;
    pa.C = C; pa.alpha = alpha;
    CUtensorMap tm;
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, (void *)A, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 0;
    const size_t smem = (size_t)ta * tc * 8 + 1024 + 16;
    {   // per call: the attribute belongs to the (function, device) pair and several host threads / devices may come through here
        cudaError_t e = cudaFuncSetAttribute(permute_tma_f64, cudaFuncAttributeMaxDynamicSharedMemorySize, 9216 * 8 + 1024 + 16);
        if (e != cudaSuccess) return -cuda_fail(e, "cudaFuncSetAttribute(permute_tma_f64)");
    }
    permute_tma_f64<<<(unsigned)ntiles, 256, smem, s>>>(tm, pa);
    launch_counter()++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return -cuda_fail(e, "permute_tma_f64");
    return 1;
}

struct Axis { int64_t ext, sa, sc; };

template <typename T>
int permute_impl(void *stream, int rank, const int64_t *dimsA, const int64_t *stridesA, const int *perm, const int64_t *stridesC, T alpha,
                 const T *A, T beta, T *C, int conj) {
    if (rank < 0 || rank > EB200_MAX_RANK) return fail(EB200_ERR_RANK, "permute: rank %d outside [0,%d]", rank, EB200_MAX_RANK);
    if (rank > 0 && (!dimsA || !stridesA || !perm || !stridesC)) return fail(EB200_ERR_INVALID_ARGUMENT, "permute: NULL dims/strides/perm");
    bool seen[EB200_MAX_RANK] = {false};
    for (int i = 0; i < rank; i++) {
        if (perm[i] < 0 || perm[i] >= rank || seen[perm[i]])
            return fail(EB200_ERR_INVALID_ARGUMENT, "permute: perm is not a permutation of 0..%d (perm[%d]=%d)", rank - 1, i, perm[i]);
        seen[perm[i]] = true;
    }
    std::vector<Axis> ax;
    for (int i = 0; i < rank; i++) {
        int     ai = perm[i];
        int64_t e = dimsA[ai];
        if (e < 0) return fail(EB200_ERR_DOMAIN, "permute: negative extent %lld on axis %d", (long long)e, ai);
        if (e == 0) return EB200_OK; // empty tensor: nothing to do
        if (e == 1) continue;
        ax.push_back({e, stridesA[ai], stridesC[i]});
    }
    if (!A || !C) return fail(EB200_ERR_INVALID_ARGUMENT, "permute: NULL tensor pointer");
    std::stable_sort(ax.begin(), ax.end(), [](const Axis &x, const Axis &y) { return std::llabs(x.sa) < std::llabs(y.sa); });
    for (size_t i = 0; i + 1 < ax.size();) {
        if (ax[i + 1].sa == ax[i].sa * ax[i].ext && ax[i + 1].sc == ax[i].sc * ax[i].ext) {
            ax[i].ext *= ax[i + 1].ext;
            ax.erase(ax.begin() + i + 1);
        } else i++;
    }
    if (ax.empty()) ax.push_back({1, 1, 1}); // scalar / all-ones tensor
    // a plain contiguous copy merges into ONE axis; the 2-D patch kernel (8 independent 16-byte loads per thread, 6.8 TB/s) wants two:
    // split off an inner run of 128..4096 elements when the extent allows it (the 1-D row kernel measured 6.1-6.3 TB/s here)
    if (ax.size() == 1 && ax[0].sa == 1 && ax[0].sc == 1 && ax[0].ext >= (1 << 16)) {
        for (int64_t inner = 4096; inner >= 128; inner >>= 1)
            if (ax[0].ext % inner == 0) {
                const int64_t outer = ax[0].ext / inner;
                ax[0].ext = inner;
                ax.push_back({outer, inner, inner});
                break;
            }
    }

    PermArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.rest.tot = 1;
    pa.A = A; pa.C = C;
    if constexpr (Scalar<T>::is_complex) { pa.alpha[0] = alpha.x; pa.alpha[1] = alpha.y; pa.beta[0] = beta.x; pa.beta[1] = beta.y; }
    else { pa.alpha[0] = alpha; pa.beta[0] = beta; }

    size_t ia = 0, ic = 0;
    for (size_t i = 1; i < ax.size(); i++)
        if (std::llabs(ax[i].sc) < std::llabs(ax[ic].sc)) ic = i;
    auto push_rest = [&](const Axis &x) -> bool {
        if (pa.rest.n >= MAXG) return false;
        pa.rest.ext[pa.rest.n] = x.ext;
        pa.rest.s[ROLE_A][pa.rest.n] = x.sa;
        pa.rest.s[ROLE_C][pa.rest.n] = x.sc;
        pa.rest.n++;
        pa.rest.tot *= x.ext;
        return true;
    };
    cudaStream_t s = as_stream(stream);
    const bool   cj = conj && Scalar<T>::is_complex;
    if (ia == ic) {
        pa.ea = ax[ia].ext; pa.sa_a = ax[ia].sa; pa.sc_a = ax[ia].sc;
        for (size_t i = 0; i < ax.size(); i++)
            if (i != ia && !push_rest(ax[i])) return fail(EB200_ERR_UNSUPPORTED, "permute: more than %d non-mergeable axes", MAXG + 2);
        if constexpr (!Scalar<T>::is_complex) {
            // 2-D patch kernel: needs a second axis, pairs along the fastest axis, everything 2-element aligned
            bool ok2d = !cj && ax.size() >= 2 && pa.sa_a == 1 && pa.sc_a == 1 && pa.ea % 2 == 0 && ((uintptr_t)A % (2 * sizeof(T)) == 0) &&
                        ((uintptr_t)C % (2 * sizeof(T)) == 0);
            for (size_t i = 0; i < ax.size() && ok2d; i++)
                if (i != ia && (ax[i].sa % 2 || ax[i].sc % 2)) ok2d = false;
            if (ok2d) {
                // second axis: the remaining axis with the smallest C stride (keeps the written rows close together)
                size_t i2 = (ia == 0) ? 1 : 0;
                for (size_t i = 0; i < ax.size(); i++)
                    if (i != ia && std::llabs(ax[i].sc) < std::llabs(ax[i2].sc)) i2 = i;
                PermArgs pb = pa;
                memset(&pb.rest, 0, sizeof(pb.rest));
                pb.rest.tot = 1;
                pb.ec = ax[i2].ext; pb.sa_c = ax[i2].sa; pb.sc_c = ax[i2].sc;
                bool fits = true;
                for (size_t i = 0; i < ax.size(); i++)
                    if (i != ia && i != i2) {
                        if (pb.rest.n >= MAXG) { fits = false; break; }
                        pb.rest.ext[pb.rest.n] = ax[i].ext;
                        pb.rest.s[ROLE_A][pb.rest.n] = ax[i].sa;
                        pb.rest.s[ROLE_C][pb.rest.n] = ax[i].sc;
                        pb.rest.n++;
                        pb.rest.tot *= ax[i].ext;
                    }
                pb.tiles_a = (pb.ea + 63) / 64;
                pb.tiles_c = (pb.ec + 63) / 64;
                const int64_t ntiles = pb.tiles_a * pb.tiles_c * pb.rest.tot;
                if (fits && ntiles < 0x7fffffffLL) {
                    grp_finalize(pb.rest);
                    if (!is_zero(beta)) permute_rows2d<T, true><<<(unsigned)ntiles, 256, 0, s>>>(pb);
                    else permute_rows2d<T, false><<<(unsigned)ntiles, 256, 0, s>>>(pb);
                    EB_POST_LAUNCH("permute_rows2d");
                    return EB200_OK;
                }
            }
        }
        grp_finalize(pa.rest);
        constexpr int VEC = 16 / sizeof(T) > 1 ? 16 / sizeof(T) : 1;
        bool          vec_ok = VEC > 1 && pa.sa_a == 1 && pa.sc_a == 1 && pa.ea % VEC == 0 && ((uintptr_t)A % 16 == 0) && ((uintptr_t)C % 16 == 0);
        for (int i = 0; i < pa.rest.n && vec_ok; i++)
            if (pa.rest.s[ROLE_A][i] % VEC || pa.rest.s[ROLE_C][i] % VEC) vec_ok = false;
        int64_t chunks = ((pa.ea + (vec_ok ? VEC : 1) - 1) / (vec_ok ? VEC : 1)) * pa.rest.tot;
        fastdiv_init((pa.ea + (vec_ok ? VEC : 1) - 1) / (vec_ok ? VEC : 1), pa.cpr_mul, pa.cpr_shift);
        int     grid = (int)std::max<int64_t>(1, std::min<int64_t>((chunks + 255) / 256, (int64_t)sm_count() * 16));
        const bool hb = !is_zero(beta);
#define EB_ROWS(CJ, V)                                                            \
    {                                                                             \
        if (hb) permute_rows<T, CJ, V, true><<<grid, 256, 0, s>>>(pa);            \
        else permute_rows<T, CJ, V, false><<<grid, 256, 0, s>>>(pa);              \
    }
        if (vec_ok) {
            if (cj) EB_ROWS(true, VEC) else EB_ROWS(false, VEC)
        } else {
            if (cj) EB_ROWS(true, 1) else EB_ROWS(false, 1)
        }
#undef EB_ROWS
        EB_POST_LAUNCH("permute_rows");
    } else {
        pa.ea = ax[ia].ext; pa.sa_a = ax[ia].sa; pa.sc_a = ax[ia].sc;
        pa.ec = ax[ic].ext; pa.sa_c = ax[ic].sa; pa.sc_c = ax[ic].sc;
        for (size_t i = 0; i < ax.size(); i++)
            if (i != ia && i != ic && !push_rest(ax[i])) return fail(EB200_ERR_UNSUPPORTED, "permute: more than %d non-mergeable axes", MAXG + 2);
        grp_finalize(pa.rest);
        if constexpr (!Scalar<T>::is_complex) {
            bool vec_ok = !cj && pa.sa_a == 1 && pa.sc_c == 1 && pa.ea % 2 == 0 && pa.ec % 2 == 0 && pa.sa_c % 2 == 0 && pa.sc_a % 2 == 0 &&
                          ((uintptr_t)A % (2 * sizeof(T)) == 0) && ((uintptr_t)C % (2 * sizeof(T)) == 0);
            for (int i = 0; i < pa.rest.n && vec_ok; i++)
                if (pa.rest.s[ROLE_A][i] % 2 || pa.rest.s[ROLE_C][i] % 2) vec_ok = false;
            if constexpr (std::is_same_v<T, double>) {
                const int mode = option(OPT_PERMUTE_TMA);
                if (vec_ok && mode != 0 && is_zero(beta) && ax.size() <= 5 && ((uintptr_t)A % 16) == 0) {
                    std::vector<int64_t> e_, sa_, sc_;
                    for (auto &x : ax) { e_.push_back(x.ext); sa_.push_back(x.sa); sc_.push_back(x.sc); }
                    const int r = try_permute_tma(s, e_, sa_, sc_, ia, ic, A, C, alpha, mode);
                    if (r == 1) return EB200_OK;
                    if (r < 0) return -r;
                }
            }
            if (vec_ok) {
                pa.tiles_a = (pa.ea + 63) / 64;
                pa.tiles_c = (pa.ec + 63) / 64;
                // how full are 64-tiles / 32-tiles? (96 -> 56 % vs 100 %, 160 -> 69 % vs 100 %)
                const int64_t ta32 = (pa.ea + 31) / 32, tc32 = (pa.ec + 31) / 32;
                const double  fill64 = double(pa.ea) * double(pa.ec) / (double(pa.tiles_a * 64) * double(pa.tiles_c * 64));
                const double  fill32 = double(pa.ea) * double(pa.ec) / (double(ta32 * 32) * double(tc32 * 32));
                // measured (n = 96, 160): with beta the old-C prefetch doubles the loads in flight and 32-tiles reach 99 % of the copy
                // rate; without it they only pay off when 64-tiles are less than 60 % full
                const bool    small = is_zero(beta) ? (fill32 > fill64 + 0.1 && fill64 < 0.6) : fill32 >= fill64;
                if (small) { pa.tiles_a = ta32; pa.tiles_c = tc32; }
                int64_t ntiles = pa.tiles_a * pa.tiles_c * pa.rest.tot;
                int     grid = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, 0x3fffffffLL)); // one tile per CTA: the hardware scheduler balances the tail
                if (small && is_zero(beta)) {
                    constexpr int U = 4;
                    permute_tile32_multi<T, U><<<(unsigned)std::min<int64_t>((ntiles + U - 1) / U, 0x3fffffffLL), 256, 0, s>>>(pa);
                } else if (small) {
                    permute_tile_vec<T, true, 32><<<grid, 256, 0, s>>>(pa);
                } else {
                    if (!is_zero(beta)) permute_tile_vec<T, true, 64><<<grid, 256, 0, s>>>(pa);
                    else permute_tile_vec<T, false, 64><<<grid, 256, 0, s>>>(pa);
                }
                EB_POST_LAUNCH("permute_tile_vec");
                return EB200_OK;
            }
        }
        constexpr int TILE = 32, ROWS = 8;
        pa.tiles_a = (pa.ea + TILE - 1) / TILE;
        pa.tiles_c = (pa.ec + TILE - 1) / TILE;
        int64_t ntiles = pa.tiles_a * pa.tiles_c * pa.rest.tot;
        int     grid = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, 0x3fffffffLL));
        dim3    block(TILE, ROWS);
        if (cj) permute_tile<T, true, TILE, ROWS><<<grid, block, 0, s>>>(pa);
        else permute_tile<T, false, TILE, ROWS><<<grid, block, 0, s>>>(pa);
        EB_POST_LAUNCH("permute_tile");
    }
    return EB200_OK;
}

} // namespace
} // namespace eb200

using namespace eb200;

extern "C" {

int eb200_spermute(void *stream, int rank, const int64_t *dimsA, const int64_t *stridesA, const int *perm, const int64_t *stridesC,
                   const float *alpha, const float *A, const float *beta, float *C, int conj) {
    if (!alpha || !beta) return fail(EB200_ERR_INVALID_ARGUMENT, "permute: alpha/beta is NULL");
    return permute_impl<float>(stream, rank, dimsA, stridesA, perm, stridesC, *alpha, A, *beta, C, conj);
}
int eb200_dpermute(void *stream, int rank, const int64_t *dimsA, const int64_t *stridesA, const int *perm, const int64_t *stridesC,
                   const double *alpha, const double *A, const double *beta, double *C, int conj) {
    if (!alpha || !beta) return fail(EB200_ERR_INVALID_ARGUMENT, "permute: alpha/beta is NULL");
    return permute_impl<double>(stream, rank, dimsA, stridesA, perm, stridesC, *alpha, A, *beta, C, conj);
}
int eb200_cpermute(void *stream, int rank, const int64_t *dimsA, const int64_t *stridesA, const int *perm, const int64_t *stridesC,
                   const float *alpha, const float *A, const float *beta, float *C, int conj) {
    if (!alpha || !beta) return fail(EB200_ERR_INVALID_ARGUMENT, "permute: alpha/beta is NULL");
    return permute_impl<float2>(stream, rank, dimsA, stridesA, perm, stridesC, *(const float2 *)alpha, (const float2 *)A, *(const float2 *)beta,
                                (float2 *)C, conj);
}
int eb200_zpermute(void *stream, int rank, const int64_t *dimsA, const int64_t *stridesA, const int *perm, const int64_t *stridesC,
                   const double *alpha, const double *A, const double *beta, double *C, int conj) {
    if (!alpha || !beta) return fail(EB200_ERR_INVALID_ARGUMENT, "permute: alpha/beta is NULL");
    return permute_impl<double2>(stream, rank, dimsA, stridesA, perm, stridesC, *(const double2 *)alpha, (const double2 *)A,
                                 *(const double2 *)beta, (double2 *)C, conj);
}

int eb200_copy_strided(void *stream, int dtype, int rank, const int64_t *dims, const void *src, const int64_t *src_strides, void *dst,
                       const int64_t *dst_strides) {
    int perm[EB200_MAX_RANK];
    if (rank < 0 || rank > EB200_MAX_RANK) return fail(EB200_ERR_RANK, "copy_strided: rank %d outside [0,%d]", rank, EB200_MAX_RANK);
    for (int i = 0; i < rank; i++) perm[i] = i;
    switch (dtype) {
    case EB200_F32: return permute_impl<float>(stream, rank, dims, src_strides, perm, dst_strides, 1.0f, (const float *)src, 0.0f, (float *)dst, 0);
    case EB200_F64: return permute_impl<double>(stream, rank, dims, src_strides, perm, dst_strides, 1.0, (const double *)src, 0.0, (double *)dst, 0);
    case EB200_C64:
        return permute_impl<float2>(stream, rank, dims, src_strides, perm, dst_strides, make_float2(1.f, 0.f), (const float2 *)src,
                                    make_float2(0.f, 0.f), (float2 *)dst, 0);
    case EB200_C128:
        return permute_impl<double2>(stream, rank, dims, src_strides, perm, dst_strides, make_double2(1., 0.), (const double2 *)src,
                                     make_double2(0., 0.), (double2 *)dst, 0);
    default: return fail(EB200_ERR_INVALID_ARGUMENT, "copy_strided: bad dtype %d", dtype);
    }
}
}
