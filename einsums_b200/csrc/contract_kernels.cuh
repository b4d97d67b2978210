// contract_kernels.cuh — device kernels of the contraction engine (sm_100a).
//
//   tile_mma_f64     C tile = A tile x B tile on the FP64 tensor pipe (DMMA.8x8x4 via
//                    mma.sync.m8n8k4.f64; tcgen05 has no f64 kind), operands gathered straight from
//                    their native index order into padded shared memory by a multi-stage cp.async
//                    pipeline — the index permutation is fused into the loads (transpose-free TTGT).
//   tile_simt        same tiling on the FMA pipes for f32 / complex.
//   row_reduce       streaming kernel for contractions with <= 8 output columns per row and the
//                    contracted direction contiguous in A (gemv 'T'-like, Fock J/K, dot).
//   col_reduce       same, row direction contiguous in A (gemv 'N'-like).
//   elementwise      no contracted index (outer / direct / broadcast products).
//   splitk_finalize  fixed-order sum of split partials + alpha/beta epilogue (deterministic).
//   scale_tensor     C = beta*C over C's full extent (beta == 0 writes zeros without reading).
#pragma once
#include "plan.cuh"

namespace eb200 {

// ------------------------------------------------------------------------------------------
// small PTX helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 8-byte async copy global->shared; src_bytes == 0 zero-fills the destination
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct TileArgs {
    Plan          p;
    const void   *A;
    const void   *B;
    void         *C;
    void         *ws;        // split-K partials [split][batch][M][N] (C's type), or nullptr
    double        alpha[2];
    double        beta[2];
    int32_t       splits;    // >= 1
    int32_t       tiles_m, tiles_n;
    int32_t       conj_a, conj_b;
    int32_t       vec_ab;    // A and B tiles can be fetched as aligned 16-byte pairs along their fast direction
    int32_t       vec_c;     // C pairs (n, n+1) are adjacent and 16-byte aligned
    int64_t       batch0;    // first batch of this launch (grid.z is limited to 65535)
};

// 16-byte async copy global->shared (L2 only); src_bytes == 0 zero-fills
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
}
// element-sized async copy (4, 8 or 16 bytes); valid == false zero-fills
template <int BYTES> __device__ __forceinline__ void cp_async_elem(uint32_t dst, const void *src, bool valid) {
    static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "cp.async size");
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;\n" ::"r"(dst), "l"(src), "n"(BYTES), "r"(valid ? BYTES : 0));
}

// ------------------------------------------------------------------------------------------
// FP64 tensor-pipe tile kernel
// ------------------------------------------------------------------------------------------
template <int BM, int BN, int BK, int WM, int WN, int STAGES, bool A_KMAJOR, bool B_KMAJOR> struct TileCfg {
    static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
    static constexpr int NT      = WARPS_M * WARPS_N * 32;
    static constexpr int PAD     = 4; // pitch == 4 (mod 16) doubles: conflict-free DMMA fragment loads
    static constexpr int A_PITCH = A_KMAJOR ? (BK + PAD) : (BM + PAD);
    static constexpr int B_PITCH = B_KMAJOR ? (BK + PAD) : (BN + PAD);
    static constexpr int A_ELEMS = A_KMAJOR ? BM * A_PITCH : BK * A_PITCH;
    static constexpr int B_ELEMS = B_KMAJOR ? BN * B_PITCH : BK * B_PITCH;
    static constexpr int RING    = STAGES + 1;
    static constexpr size_t SMEM = (size_t)STAGES * (A_ELEMS + B_ELEMS) * 8 + (size_t)(2 * BM + 2 * BN) * 8 + (size_t)RING * BK * 2 * 8;
};

template <int BM, int BN, int BK, int WM, int WN, int STAGES, bool A_KMAJOR, bool B_KMAJOR, int MINB, bool VEC>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32, MINB) tile_mma_f64(const __grid_constant__ TileArgs args) {
    using Cfg = TileCfg<BM, BN, BK, WM, WN, STAGES, A_KMAJOR, B_KMAJOR>;
    constexpr int NT = Cfg::NT;
    constexpr int FM = WM / 8, FN = WN / 8;
    static_assert(BK % 4 == 0 && (BM * BK) % NT == 0 && (BN * BK) % NT == 0, "tile shape");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double  *As    = reinterpret_cast<double *>(smem_raw);
    double  *Bs    = As + STAGES * Cfg::A_ELEMS;
    int64_t *moffA = reinterpret_cast<int64_t *>(Bs + STAGES * Cfg::B_ELEMS);
    int64_t *moffC = moffA + BM;
    int64_t *noffB = moffC + BM;
    int64_t *noffC = noffB + BN;
    int64_t *koffA = noffC + BN;            // [RING][BK]
    int64_t *koffB = koffA + Cfg::RING * BK; // [RING][BK]

    const Plan &p   = args.p;
    const int   tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int   wm0 = (warp % Cfg::WARPS_M) * WM, wn0 = (warp / Cfg::WARPS_M) * WN;

    // grouped rasterisation: consecutive CTAs share A/B slabs in L2
    int tile_m, tile_n;
    {
        constexpr int GROUP = 8;
        int id = blockIdx.x, per_group = GROUP * args.tiles_n;
        int g = id / per_group, first = g * GROUP;
        int gsize = min(args.tiles_m - first, GROUP);
        int in_g = id - g * per_group;
        tile_m = first + in_g % gsize;
        tile_n = in_g / gsize;
    }
    const int     split = blockIdx.y;
    const int64_t batch = blockIdx.z + args.batch0;

    int64_t baseA, baseB, baseC;
    grp_offset3(p.b, batch, baseA, baseB, baseC);
    const double *A = reinterpret_cast<const double *>(args.A) + baseA;
    const double *B = reinterpret_cast<const double *>(args.B) + baseB;

    const int64_t Mtot = p.m.tot, Ntot = p.n.tot, Ktot = p.k.tot;
    const int     KT = (int)((Ktot + BK - 1) / BK);
    const int     ks0 = (int)((int64_t)KT * split / args.splits), ks1 = (int)((int64_t)KT * (split + 1) / args.splits);
    const int     nsteps = ks1 - ks0;

    for (int i = tid; i < BM; i += NT) {
        int64_t m = (int64_t)tile_m * BM + i, oa = OFF_INVALID, oc = OFF_INVALID;
        if (m < Mtot) grp_offset2(p.m, m, ROLE_A, ROLE_C, oa, oc);
        moffA[i] = oa;
        moffC[i] = oc;
    }
    for (int i = tid; i < BN; i += NT) {
        int64_t n = (int64_t)tile_n * BN + i, ob = OFF_INVALID, oc = OFF_INVALID;
        if (n < Ntot) grp_offset2(p.n, n, ROLE_B, ROLE_C, ob, oc);
        noffB[i] = ob;
        noffC[i] = oc;
    }
    auto compute_koff = [&](int step /*local*/) {
        // one entry per thread, BK threads
        int     slot = step % Cfg::RING;
        int64_t k = (int64_t)(ks0 + step) * BK + tid, oa = OFF_INVALID, ob = OFF_INVALID;
        if (k < Ktot) grp_offset2(p.k, k, ROLE_A, ROLE_B, oa, ob);
        koffA[slot * BK + tid] = oa;
        koffB[slot * BK + tid] = ob;
    };
    if (tid < BK)
        for (int s = 0; s < STAGES && s < nsteps; s++) compute_koff(s);
    __syncthreads();

    // The stage loader is cut into BK/4 parts so that it can be interleaved with the k4 sub-steps of the
    // MMA loop (keeps the tensor pipe fed while addresses are generated). part < 0 issues everything.
    constexpr int NPART = BK / 4;
    constexpr int A_OPS = VEC ? BM * BK / 2 / NT : BM * BK / NT;
    constexpr int B_OPS = VEC ? BN * BK / 2 / NT : BN * BK / NT;
    static_assert(A_OPS >= 1 && B_OPS >= 1, "tile too small for the thread count");
    auto load_part = [&](int step /*local*/, int part) {
        const int      buf  = step % STAGES;
        const int      slot = step % Cfg::RING;
        const uint32_t as = smem_u32(As + buf * Cfg::A_ELEMS), bs = smem_u32(Bs + buf * Cfg::B_ELEMS);
#pragma unroll
        for (int j = 0; j < A_OPS; j++) {
            if (part >= 0 && (j % NPART) != part) continue;
            int e = tid + j * NT, ml, kl;
            if constexpr (VEC) {
                if constexpr (A_KMAJOR) { kl = 2 * (e % (BK / 2)); ml = e / (BK / 2); }
                else { ml = 2 * (e % (BM / 2)); kl = e / (BM / 2); }
            } else {
                if constexpr (A_KMAJOR) { kl = e % BK; ml = e / BK; }
                else { ml = e % BM; kl = e / BM; }
            }
            int64_t  om = moffA[ml], ok = koffA[slot * BK + kl];
            bool     v = (om != OFF_INVALID) && (ok != OFF_INVALID);
            uint32_t dst = as + 8u * (A_KMAJOR ? ml * Cfg::A_PITCH + kl : kl * Cfg::A_PITCH + ml);
            if constexpr (VEC) cp_async16(dst, v ? (const void *)(A + om + ok) : (const void *)A, v ? 16 : 0);
            else cp_async8(dst, v ? (const void *)(A + om + ok) : (const void *)A, v ? 8 : 0);
        }
#pragma unroll
        for (int j = 0; j < B_OPS; j++) {
            if (part >= 0 && (j % NPART) != part) continue;
            int e = tid + j * NT, nl, kl;
            if constexpr (VEC) {
                if constexpr (B_KMAJOR) { kl = 2 * (e % (BK / 2)); nl = e / (BK / 2); }
                else { nl = 2 * (e % (BN / 2)); kl = e / (BN / 2); }
            } else {
                if constexpr (B_KMAJOR) { kl = e % BK; nl = e / BK; }
                else { nl = e % BN; kl = e / BN; }
            }
            int64_t  on = noffB[nl], ok = koffB[slot * BK + kl];
            bool     v = (on != OFF_INVALID) && (ok != OFF_INVALID);
            uint32_t dst = bs + 8u * (B_KMAJOR ? nl * Cfg::B_PITCH + kl : kl * Cfg::B_PITCH + nl);
            if constexpr (VEC) cp_async16(dst, v ? (const void *)(B + on + ok) : (const void *)B, v ? 16 : 0);
            else cp_async8(dst, v ? (const void *)(B + on + ok) : (const void *)B, v ? 8 : 0);
        }
    };
    auto load_stage = [&](int step) { load_part(step, -1); };

    double acc[FM][FN][2];
#pragma unroll
    for (int i = 0; i < FM; i++)
#pragma unroll
        for (int j = 0; j < FN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nsteps) load_stage(s);
        cp_async_commit();
    }

    const int fr = lane >> 2, fk = lane & 3; // fragment row (m or n) and k within the k4 step
    for (int it = 0; it < nsteps; it++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const bool prefetch = it + STAGES - 1 < nsteps;
        if (tid < BK && it + STAGES < nsteps) compute_koff(it + STAGES);

        const double *a_s = As + (it % STAGES) * Cfg::A_ELEMS;
        const double *b_s = Bs + (it % STAGES) * Cfg::B_ELEMS;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double a[FM], b[FN];
#pragma unroll
            for (int i = 0; i < FM; i++) {
                int m = wm0 + 8 * i + fr, k = kk + fk;
                a[i] = a_s[A_KMAJOR ? m * Cfg::A_PITCH + k : k * Cfg::A_PITCH + m];
            }
#pragma unroll
            for (int j = 0; j < FN; j++) {
                int n = wn0 + 8 * j + fr, k = kk + fk;
                b[j] = b_s[B_KMAJOR ? n * Cfg::B_PITCH + k : k * Cfg::B_PITCH + n];
            }
            // one slice of next stage's copies is issued while the fragment loads are in flight
            if (prefetch) load_part(it + STAGES - 1, kk / 4);
#pragma unroll
            for (int i = 0; i < FM; i++)
#pragma unroll
                for (int j = 0; j < FN; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        cp_async_commit();
    }
    cp_async_wait<0>();

    // ---- epilogue: C fragment (row = lane/4, cols = 2*(lane%4)+{0,1}) straight from registers
    if (args.splits == 1) {
        double      *C = reinterpret_cast<double *>(args.C) + baseC;
        const double alpha = args.alpha[0], beta = args.beta[0];
        const bool   has_beta = beta != 0.0;
        // two passes over the fragment rows; inside a pass every C load is issued before the first use
        constexpr int HM = (FM + 1) / 2;
#pragma unroll
        for (int h = 0; h < FM; h += HM) {
            double  cold[HM][FN][2];
            int64_t offs[HM][FN];
#pragma unroll
            for (int ii = 0; ii < HM; ii++) {
                const int     i = h + ii;
                const int64_t om = (i < FM) ? moffC[wm0 + 8 * i + fr] : OFF_INVALID;
#pragma unroll
                for (int j = 0; j < FN; j++) {
                    const int64_t on0 = noffC[wn0 + 8 * j + 2 * fk], on1 = noffC[wn0 + 8 * j + 2 * fk + 1];
                    cold[ii][j][0] = cold[ii][j][1] = 0.0;
                    if (args.vec_c) {
                        offs[ii][j] = (om != OFF_INVALID && on0 != OFF_INVALID) ? om + on0 : OFF_INVALID;
                        if (has_beta && offs[ii][j] != OFF_INVALID) {
                            double2 v = *reinterpret_cast<const double2 *>(C + offs[ii][j]);
                            cold[ii][j][0] = v.x; cold[ii][j][1] = v.y;
                        }
                    } else {
                        offs[ii][j] = om; // row offset; columns looked up again below
                        if (has_beta && om != OFF_INVALID) {
                            if (on0 != OFF_INVALID) cold[ii][j][0] = C[om + on0];
                            if (on1 != OFF_INVALID) cold[ii][j][1] = C[om + on1];
                        }
                    }
                }
            }
#pragma unroll
            for (int ii = 0; ii < HM; ii++) {
                const int i = h + ii;
                if (i >= FM) continue;
#pragma unroll
                for (int j = 0; j < FN; j++) {
                    if (offs[ii][j] == OFF_INVALID) continue;
                    double v0 = alpha * acc[i][j][0], v1 = alpha * acc[i][j][1];
                    if (has_beta) { v0 = fma(beta, cold[ii][j][0], v0); v1 = fma(beta, cold[ii][j][1], v1); }
                    if (args.vec_c) {
                        *reinterpret_cast<double2 *>(C + offs[ii][j]) = make_double2(v0, v1);
                    } else {
                        const int64_t on0 = noffC[wn0 + 8 * j + 2 * fk], on1 = noffC[wn0 + 8 * j + 2 * fk + 1];
                        if (on0 != OFF_INVALID) C[offs[ii][j] + on0] = v0;
                        if (on1 != OFF_INVALID) C[offs[ii][j] + on1] = v1;
                    }
                }
            }
        }
    } else {
        double *ws = reinterpret_cast<double *>(args.ws) + (((int64_t)split * p.b.tot + batch) * Mtot) * Ntot;
#pragma unroll
        for (int i = 0; i < FM; i++) {
            int64_t m = (int64_t)tile_m * BM + wm0 + 8 * i + fr;
            if (m >= Mtot) continue;
#pragma unroll
            for (int j = 0; j < FN; j++) {
                int64_t n = (int64_t)tile_n * BN + wn0 + 8 * j + 2 * fk;
                if (n < Ntot) ws[m * Ntot + n] = acc[i][j][0];
                if (n + 1 < Ntot) ws[m * Ntot + n + 1] = acc[i][j][1];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// FMA-pipe tile kernel (f32 generic contractions, complex, and f64 for cross-checks).
// 256 threads as 16 x 16, TM x TN outputs per thread (8x8 for 4-byte types, 4x4 otherwise), BK = 16 (8 for 16-byte types),
// double-buffered shared memory filled by cp.async: the gathers of step it+1 are in flight while step it is multiplied.
// Same fused-permutation addressing as tile_mma_f64: base + moff[m] + koff[k].
// ------------------------------------------------------------------------------------------
template <typename T> struct SimtCfg {
    static constexpr int TM = sizeof(T) == 4 ? 8 : 4, TN = TM;
    static constexpr int BM = 16 * TM, BN = 16 * TN;
    static constexpr int BK = sizeof(T) == 16 ? 8 : 16;
};

template <typename T, bool CONJ_A, bool CONJ_B> __global__ void __launch_bounds__(256, 2) tile_simt(const __grid_constant__ TileArgs args) {
    using Cfg = SimtCfg<T>;
    constexpr int TM = Cfg::TM, TN = Cfg::TN, BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, NT = 256;
    constexpr int EA = BM * BK / NT, EB = BN * BK / NT;
    constexpr bool V4 = sizeof(T) == 4; // 4-byte types read their 8 + 8 operands as four 16-byte shared loads
    constexpr int  PAD = V4 ? 4 : 1;
    __shared__ __align__(16) T As[2][BK][BM + PAD];
    __shared__ __align__(16) T Bs[2][BK][BN + PAD];
    __shared__ int64_t moffA[BM], moffC[BM], noffB[BN], noffC[BN], koffA[3][BK], koffB[3][BK];

    const Plan &p = args.p;
    const int   tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int   tile_m = blockIdx.x % args.tiles_m, tile_n = blockIdx.x / args.tiles_m;
    const int   split = blockIdx.y;
    const int64_t batch = blockIdx.z + args.batch0;

    int64_t baseA, baseB, baseC;
    grp_offset3(p.b, batch, baseA, baseB, baseC);
    const T *A = reinterpret_cast<const T *>(args.A) + baseA;
    const T *B = reinterpret_cast<const T *>(args.B) + baseB;

    const int64_t Mtot = p.m.tot, Ntot = p.n.tot, Ktot = p.k.tot;
    const int     KT = (int)((Ktot + BK - 1) / BK);
    const int     ks0 = (int)((int64_t)KT * split / args.splits), ks1 = (int)((int64_t)KT * (split + 1) / args.splits);
    const int     nsteps = ks1 - ks0;

    for (int i = tid; i < BM; i += NT) {
        int64_t m = (int64_t)tile_m * BM + i, oa = OFF_INVALID, oc = OFF_INVALID;
        if (m < Mtot) grp_offset2(p.m, m, ROLE_A, ROLE_C, oa, oc);
        moffA[i] = oa; moffC[i] = oc;
    }
    for (int i = tid; i < BN; i += NT) {
        int64_t n = (int64_t)tile_n * BN + i, ob = OFF_INVALID, oc = OFF_INVALID;
        if (n < Ntot) grp_offset2(p.n, n, ROLE_B, ROLE_C, ob, oc);
        noffB[i] = ob; noffC[i] = oc;
    }
    auto compute_koff = [&](int step) { // threads < BK
        int64_t k = (int64_t)(ks0 + step) * BK + tid, oa = OFF_INVALID, ob = OFF_INVALID;
        if (k < Ktot) grp_offset2(p.k, k, ROLE_A, ROLE_B, oa, ob);
        koffA[step % 3][tid] = oa;
        koffB[step % 3][tid] = ob;
    };
    if (tid < BK) {
        compute_koff(0);
        if (nsteps > 1) compute_koff(1);
    }
    __syncthreads();

    // Each thread's EA (EB) elements share one coordinate (k when the operand is k-major, else m/n) and step through the
    // other, so one offset is hoisted and the other comes from a single table walk.
    constexpr int PA = BM + PAD, PB = BN + PAD;
    const bool    akm = p.a_kmajor, bkm = p.b_kmajor;
    const int     a_fix = akm ? tid % BK : tid % BM, a_v0 = akm ? tid / BK : tid / BM, a_vs = akm ? NT / BK : NT / BM;
    const int     b_fix = bkm ? tid % BK : tid % BN, b_v0 = bkm ? tid / BK : tid / BN, b_vs = bkm ? NT / BK : NT / BN;
    const int     a_s0 = akm ? a_fix * PA + a_v0 : a_v0 * PA + a_fix, a_ss = akm ? a_vs : a_vs * PA;
    const int     b_s0 = bkm ? b_fix * PB + b_v0 : b_v0 * PB + b_fix, b_ss = bkm ? b_vs : b_vs * PB;
    const int64_t a_mfix = akm ? 0 : moffA[a_fix], b_nfix = bkm ? 0 : noffB[b_fix];

    auto issue = [&](int step) { // global -> shared, asynchronously; out-of-range elements are zero-filled
        const int64_t *ka = koffA[step % 3], *kb = koffB[step % 3];
        {
            const int64_t  ofix = akm ? ka[a_fix] : a_mfix;
            const int64_t *tab  = akm ? moffA : ka;
            const uint32_t dst  = smem_u32(&As[step & 1][0][0] + a_s0);
#pragma unroll
            for (int j = 0; j < EA; j++) {
                int64_t ov = tab[a_v0 + j * a_vs];
                bool    v  = ofix != OFF_INVALID && ov != OFF_INVALID;
                cp_async_elem<sizeof(T)>(dst + (uint32_t)(j * a_ss * sizeof(T)), v ? (const void *)(A + ofix + ov) : (const void *)A, v);
            }
        }
        {
            const int64_t  ofix = bkm ? kb[b_fix] : b_nfix;
            const int64_t *tab  = bkm ? noffB : kb;
            const uint32_t dst  = smem_u32(&Bs[step & 1][0][0] + b_s0);
#pragma unroll
            for (int j = 0; j < EB; j++) {
                int64_t ov = tab[b_v0 + j * b_vs];
                bool    v  = ofix != OFF_INVALID && ov != OFF_INVALID;
                cp_async_elem<sizeof(T)>(dst + (uint32_t)(j * b_ss * sizeof(T)), v ? (const void *)(B + ofix + ov) : (const void *)B, v);
            }
        }
    };

    // tile-local row of accumulator i / column of accumulator j
    auto row_of = [&](int i) { return V4 ? 64 * (i >> 2) + 4 * ty + (i & 3) : ty + 16 * i; };
    auto col_of = [&](int j) { return V4 ? 64 * (j >> 2) + 4 * tx + (j & 3) : tx + 16 * j; };

    alignas(16) T acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = zero_of<T>();

    if (nsteps > 0) issue(0);
    cp_async_commit();
    for (int it = 0; it < nsteps; it++) {
        cp_async_wait<0>();
        __syncthreads(); // step `it` has landed; every thread is done with step it-1, so its buffer and k-offset slot are free
        if (tid < BK && it + 2 < nsteps) compute_koff(it + 2);
        if (it + 1 < nsteps) issue(it + 1); // in flight during the multiply below
        cp_async_commit();
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            T a[TM], b[TN];
            if constexpr (V4) {
#pragma unroll
                for (int h = 0; h < TM / 4; h++) *reinterpret_cast<float4 *>(&a[4 * h]) = *reinterpret_cast<const float4 *>(&As[it & 1][kk][64 * h + 4 * ty]);
#pragma unroll
                for (int h = 0; h < TN / 4; h++) *reinterpret_cast<float4 *>(&b[4 * h]) = *reinterpret_cast<const float4 *>(&Bs[it & 1][kk][64 * h + 4 * tx]);
            } else {
#pragma unroll
                for (int i = 0; i < TM; i++) a[i] = maybe_conj<CONJ_A>(As[it & 1][kk][ty + 16 * i]);
#pragma unroll
                for (int j = 0; j < TN; j++) b[j] = maybe_conj<CONJ_B>(Bs[it & 1][kk][tx + 16 * j]);
            }
            // packed FP32 FMAs (FFMA2: two lanes per instruction) halve the issue slots of the f32 inner product (+15-20 % measured);
            // for c64 the packing moves cost more than they save (-19 %), so complex stays on scalar FMAs
            if constexpr (std::is_same_v<T, float>) {
#pragma unroll
                for (int i = 0; i < TM; i++) {
                    const float2 a2 = make_float2(a[i], a[i]);
#pragma unroll
                    for (int j = 0; j < TN; j += 2) {
                        float2 c = make_float2(acc[i][j], acc[i][j + 1]);
                        c = __ffma2_rn(a2, make_float2(b[j], b[j + 1]), c);
                        acc[i][j] = c.x; acc[i][j + 1] = c.y;
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < TM; i++)
#pragma unroll
                    for (int j = 0; j < TN; j++) acc[i][j] = fma_(a[i], b[j], acc[i][j]);
            }
        }
    }

    if (args.splits == 1) {
        T         *C = reinterpret_cast<T *>(args.C) + baseC;
        const T    alpha = make_scalar<T>(args.alpha[0], args.alpha[1]), beta = make_scalar<T>(args.beta[0], args.beta[1]);
        const bool has_beta = !is_zero(beta);
#pragma unroll
        for (int i = 0; i < TM; i++) {
            int64_t om = moffC[row_of(i)];
            if (om == OFF_INVALID) continue;
#pragma unroll
            for (int j = 0; j < TN; j++) {
                int64_t on = noffC[col_of(j)];
                if (on == OFF_INVALID) continue;
                T *dst = C + om + on;
                *dst = axpby_(alpha, acc[i][j], beta, has_beta ? *dst : zero_of<T>(), has_beta);
            }
        }
    } else {
        T *ws = reinterpret_cast<T *>(args.ws) + (((int64_t)split * p.b.tot + batch) * Mtot) * Ntot;
#pragma unroll
        for (int i = 0; i < TM; i++) {
            int64_t m = (int64_t)tile_m * BM + row_of(i);
            if (m >= Mtot) continue;
#pragma unroll
            for (int j = 0; j < TN; j++) {
                int64_t n = (int64_t)tile_n * BN + col_of(j);
                if (n < Ntot) ws[m * Ntot + n] = acc[i][j];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// split-K / split-reduction finalize: C = beta*C + alpha * sum_s ws[s][row][col]   (fixed order)
// rows are enumerated by (gb, gm) [gb outer], columns by gn; only ROLE_C strides are used.
// ------------------------------------------------------------------------------------------
struct FinalizeArgs {
    Grp         gb, gm, gn;
    const void *ws;
    void       *C;
    double      alpha[2], beta[2];
    int32_t     splits;
};

template <typename T> __global__ void __launch_bounds__(256) splitk_finalize(const __grid_constant__ FinalizeArgs a) {
    const int64_t rows = a.gb.tot * a.gm.tot, cols = a.gn.tot, total = rows * cols;
    const T       alpha = make_scalar<T>(a.alpha[0], a.alpha[1]), beta = make_scalar<T>(a.beta[0], a.beta[1]);
    const bool    has_beta = !is_zero(beta);
    const T      *ws = reinterpret_cast<const T *>(a.ws);
    T            *C = reinterpret_cast<T *>(a.C);
    for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (int64_t)gridDim.x * blockDim.x) {
        int64_t row = x / cols, col = x - row * cols;
        int64_t bi = row / a.gm.tot, mi = row - bi * a.gm.tot;
        T       s = zero_of<T>();
        for (int sp = 0; sp < a.splits; sp++) s = add_(s, ws[(int64_t)sp * total + x]);
        int64_t off = grp_offset(a.gb, bi, ROLE_C) + grp_offset(a.gm, mi, ROLE_C) + grp_offset(a.gn, col, ROLE_C);
        C[off] = axpby_(alpha, s, beta, has_beta ? C[off] : zero_of<T>(), has_beta);
    }
}

// ------------------------------------------------------------------------------------------
// streaming reductions
// ------------------------------------------------------------------------------------------
struct ReduceArgs {
    Grp         rows, n, k;   // rows: (A,B,C) strides; n: (B,C); k: (A,B)
    const void *A;
    const void *B;
    void       *C;
    void       *ws;           // partials [split][rows.tot][n.tot] when splits > 1
    double      alpha[2], beta[2];
    int32_t     splits;
    int32_t     conj_a, conj_b;
    uint32_t    ppr_mul, ppr_shift; // fast division by (k.ext[0] / VEC): packs per inner run (row_reduce_seg, flat lane mapping)
    int32_t     flat_lanes;         // 1: lanes enumerate (ko, pack) pairs linearly (no idle lanes on short runs)
};

// row_reduce: block = RB consecutive rows x one K segment. Lanes run along k (contiguous in A).
// Each B value is loaded once and reused by the RB rows, so B traffic is 1/RB of A traffic.
template <typename T, int NN, int RB>
__global__ void __launch_bounds__(256) row_reduce(const __grid_constant__ ReduceArgs a) {
    constexpr int NT = 256, NW = NT / 32, UNR = 4;
    __shared__ T  red[NW][RB * NN];
    const int     tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t R = a.rows.tot, K = a.k.tot;
    const int     ncols = (int)a.n.tot;
    const int64_t r0 = (int64_t)blockIdx.x * RB;
    const int     split = blockIdx.y;
    const int64_t k_begin = K * split / a.splits, k_end = K * (split + 1) / a.splits;

    const T *A = reinterpret_cast<const T *>(a.A);
    const T *B = reinterpret_cast<const T *>(a.B);

    int64_t rowA[RB], rowB[RB];
    bool    same_b = true; // B does not depend on the row (no batch index): share the B loads
#pragma unroll
    for (int j = 0; j < RB; j++) {
        int64_t r = r0 + j, oa = 0, ob = 0, oc = 0;
        if (r < R) grp_offset3(a.rows, r, oa, ob, oc);
        rowA[j] = (r < R) ? oa : OFF_INVALID;
        rowB[j] = ob;
        if (ob != 0) same_b = false;
    }
    int64_t colB[NN];
#pragma unroll
    for (int c = 0; c < NN; c++) colB[c] = (c < ncols) ? grp_offset(a.n, c, ROLE_B) : 0;

    T acc[RB][NN];
#pragma unroll
    for (int j = 0; j < RB; j++)
#pragma unroll
        for (int c = 0; c < NN; c++) acc[j][c] = zero_of<T>();

    const int64_t e0 = a.k.n > 0 ? a.k.ext[0] : 1;
    for (int64_t kb = k_begin + (int64_t)tid * 1; kb < k_end; kb += (int64_t)NT * UNR) {
        int64_t oa[UNR], ob[UNR];
        bool    ok[UNR];
#pragma unroll
        for (int u = 0; u < UNR; u++) {
            int64_t k = kb + (int64_t)u * NT;
            ok[u] = k < k_end;
            oa[u] = ob[u] = 0;
            if (ok[u]) {
                if (a.k.n == 1) {
                    oa[u] = k * a.k.s[ROLE_A][0];
                    ob[u] = k * a.k.s[ROLE_B][0];
                } else {
                    grp_offset2(a.k, k, ROLE_A, ROLE_B, oa[u], ob[u]);
                }
            }
        }
        (void)e0;
        T av[UNR][RB];
#pragma unroll
        for (int u = 0; u < UNR; u++)
#pragma unroll
            for (int j = 0; j < RB; j++) av[u][j] = (ok[u] && rowA[j] != OFF_INVALID) ? A[rowA[j] + oa[u]] : zero_of<T>();
#pragma unroll
        for (int u = 0; u < UNR; u++) {
            if (!ok[u]) continue;
            if (same_b) {
#pragma unroll
                for (int c = 0; c < NN; c++) {
                    if (c >= ncols) break;
                    T bv = conj_if(B[ob[u] + colB[c]], a.conj_b != 0);
#pragma unroll
                    for (int j = 0; j < RB; j++) acc[j][c] = fma_(conj_if(av[u][j], a.conj_a != 0), bv, acc[j][c]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < RB; j++) {
                    if (rowA[j] == OFF_INVALID) continue;
#pragma unroll
                    for (int c = 0; c < NN; c++) {
                        if (c >= ncols) break;
                        T bv = conj_if(B[rowB[j] + ob[u] + colB[c]], a.conj_b != 0);
                        acc[j][c] = fma_(conj_if(av[u][j], a.conj_a != 0), bv, acc[j][c]);
                    }
                }
            }
        }
    }
    (void)lane;
    // block reduction, fixed order
#pragma unroll
    for (int j = 0; j < RB; j++)
#pragma unroll
        for (int c = 0; c < NN; c++) {
            T v = warp_sum(acc[j][c]);
            if (lane == 0) red[warp][j * NN + c] = v;
        }
    __syncthreads();
    if (tid < RB * NN) {
        int     j = tid / NN, c = tid % NN;
        int64_t r = r0 + j;
        if (r < R && c < ncols) {
            T s = zero_of<T>();
#pragma unroll
            for (int w = 0; w < NW; w++) s = add_(s, red[w][tid]);
            if (a.splits == 1) {
                const T    alpha = make_scalar<T>(a.alpha[0], a.alpha[1]), beta = make_scalar<T>(a.beta[0], a.beta[1]);
                const bool has_beta = !is_zero(beta);
                T         *dst = reinterpret_cast<T *>(a.C) + grp_offset(a.rows, r, ROLE_C) + grp_offset(a.n, c, ROLE_C);
                *dst = axpby_(alpha, s, beta, has_beta ? *dst : zero_of<T>(), has_beta);
            } else {
                reinterpret_cast<T *>(a.ws)[((int64_t)split * R + r) * ncols + c] = s;
            }
        }
    }
}

// row_reduce_seg: the bandwidth path for contractions whose contracted direction is contiguous in A
// (gemv, Fock J/K, dot). The K space is seen as (outer composite ko) x (inner index ki of extent E0,
// unit stride in A). A work unit = one run of 32*VEC consecutive ki for one ko; a warp decodes ko once
// per unit (no per-element division), lanes stream A with VEC*sizeof(T)-byte loads for RB rows at a
// time and reuse each B value RB times. Units are split over blockIdx.y; partial sums are combined
// in a fixed order (deterministic).
template <typename T, int NN, int RB, int VEC>
__global__ void __launch_bounds__(256, (RB >= 8 ? 2 : (NN * RB * sizeof(T) <= 64 ? 3 : 1))) row_reduce_seg(const __grid_constant__ ReduceArgs a) {
    constexpr int NT = 256, NW = NT / 32;
    __shared__ T  red[NW][RB * NN];
    const int     tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t R = a.rows.tot;
    const int     ncols = (int)a.n.tot;
    const int64_t r0 = (int64_t)blockIdx.x * RB;
    const int     split = blockIdx.y;
    const T      *A = reinterpret_cast<const T *>(a.A);
    const T      *B = reinterpret_cast<const T *>(a.B);

    const int64_t E0 = a.k.ext[0], sA0 = a.k.s[ROLE_A][0], sB0 = a.k.s[ROLE_B][0];
    const int64_t KO = a.k.tot / E0;
    constexpr int CH = 32 * VEC;
    const int64_t chunks = (E0 + CH - 1) / CH;
    const int64_t units = KO * chunks;
    const int64_t u_begin = units * split / a.splits, u_end = units * (split + 1) / a.splits;

    int64_t rowA[RB], rowB[RB];
    bool    same_b = true;
#pragma unroll
    for (int j = 0; j < RB; j++) {
        int64_t r = r0 + j, oa = 0, ob = 0, oc = 0;
        if (r < R) grp_offset3(a.rows, r, oa, ob, oc);
        rowA[j] = (r < R) ? oa : rowA[0]; // out-of-range rows re-read row 0 (results discarded)
        rowB[j] = (r < R) ? ob : 0;
        if (rowB[j] != 0) same_b = false;
    }
    int64_t colB[NN];
#pragma unroll
    for (int c = 0; c < NN; c++) colB[c] = (c < ncols) ? grp_offset(a.n, c, ROLE_B) : 0;

    T acc[RB][NN];
#pragma unroll
    for (int j = 0; j < RB; j++)
#pragma unroll
        for (int c = 0; c < NN; c++) acc[j][c] = zero_of<T>();

    struct alignas(sizeof(T) * VEC) Pack { T v[VEC]; };

    // flat mapping: one "pack" (VEC elements) per lane, packs of all runs enumerated linearly; segmented mapping: a warp owns
    // one 32-pack chunk of one run. Both share the body below.
    const int64_t ppr = E0 / VEC;                                   // packs per run (flat mapping needs E0 % VEC == 0)
    const int64_t total_packs = KO * ppr;
    const int64_t it_begin = a.flat_lanes ? (total_packs + 31) / 32 * split / a.splits : u_begin;
    const int64_t it_end = a.flat_lanes ? (total_packs + 31) / 32 * (split + 1) / a.splits : u_end;
    for (int64_t u = it_begin + warp; u < it_end; u += NW) {
        int64_t ko, ki;
        if (a.flat_lanes) {
            const int64_t idx = u * 32 + lane;
            if (idx >= total_packs) continue;
            const uint32_t q = fastdiv((uint32_t)idx, a.ppr_mul, a.ppr_shift);
            ko = q;
            ki = ((uint32_t)idx - q * (uint32_t)ppr) * VEC;
        } else {
            int64_t ch;
            if (chunks == 1) { ko = u; ch = 0; }
            else if (units <= 0x7fffffffLL) { uint32_t q = (uint32_t)u / (uint32_t)chunks; ko = q; ch = (uint32_t)u - q * (uint32_t)chunks; }
            else { ko = u / chunks; ch = u - ko * chunks; }
            ki = ch * CH + (int64_t)lane * VEC;
            if (ki >= E0) continue;
        }
        int64_t offA = 0, offB = 0;
        if (a.k.n > 1) { // decode the outer K indices (index 0 is the inner run)
            if (a.k.tot <= 0x7fffffffLL) {
                uint32_t x = (uint32_t)ko;
                for (int i = 1; i < a.k.n; i++) {
                    uint32_t q = (i + 1 < a.k.n) ? fastdiv(x, a.k.fmul[i], a.k.fshift[i]) : 0u, r = x - q * (uint32_t)a.k.ext[i];
                    offA += (int64_t)r * a.k.s[ROLE_A][i];
                    offB += (int64_t)r * a.k.s[ROLE_B][i];
                    x = q;
                }
            } else {
                int64_t x = ko;
                for (int i = 1; i < a.k.n; i++) {
                    int64_t q = x / a.k.ext[i], r = x - q * a.k.ext[i];
                    offA += r * a.k.s[ROLE_A][i];
                    offB += r * a.k.s[ROLE_B][i];
                    x = q;
                }
            }
        }
        Pack av[RB];
        if constexpr (VEC > 1) {
#pragma unroll
            for (int j = 0; j < RB; j++) av[j] = *reinterpret_cast<const Pack *>(A + rowA[j] + offA + ki);
        } else {
#pragma unroll
            for (int j = 0; j < RB; j++) av[j].v[0] = A[rowA[j] + offA + ki * sA0];
        }
        if (same_b) {
#pragma unroll
            for (int c = 0; c < NN; c++) {
                if (c >= ncols) break;
                Pack bv;
                if constexpr (VEC > 1) bv = *reinterpret_cast<const Pack *>(B + offB + ki + colB[c]);
                else bv.v[0] = B[offB + ki * sB0 + colB[c]];
#pragma unroll
                for (int e = 0; e < VEC; e++) {
                    T b = conj_if(bv.v[e], a.conj_b != 0);
#pragma unroll
                    for (int j = 0; j < RB; j++) acc[j][c] = fma_(conj_if(av[j].v[e], a.conj_a != 0), b, acc[j][c]);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < RB; j++)
#pragma unroll
                for (int c = 0; c < NN; c++) {
                    if (c >= ncols) break;
#pragma unroll
                    for (int e = 0; e < VEC; e++) {
                        T b = conj_if(B[rowB[j] + offB + (ki + e) * sB0 + colB[c]], a.conj_b != 0);
                        acc[j][c] = fma_(conj_if(av[j].v[e], a.conj_a != 0), b, acc[j][c]);
                    }
                }
        }
    }
#pragma unroll
    for (int j = 0; j < RB; j++)
#pragma unroll
        for (int c = 0; c < NN; c++) {
            T v = warp_sum(acc[j][c]);
            if (lane == 0) red[warp][j * NN + c] = v;
        }
    __syncthreads();
    if (tid < RB * NN) {
        int     j = tid / NN, c = tid % NN;
        int64_t r = r0 + j;
        if (r < R && c < ncols) {
            T s = zero_of<T>();
#pragma unroll
            for (int w = 0; w < NW; w++) s = add_(s, red[w][tid]);
            if (a.splits == 1) {
                const T    alpha = make_scalar<T>(a.alpha[0], a.alpha[1]), beta = make_scalar<T>(a.beta[0], a.beta[1]);
                const bool has_beta = !is_zero(beta);
                T         *dst = reinterpret_cast<T *>(a.C) + grp_offset(a.rows, r, ROLE_C) + grp_offset(a.n, c, ROLE_C);
                *dst = axpby_(alpha, s, beta, has_beta ? *dst : zero_of<T>(), has_beta);
            } else {
                reinterpret_cast<T *>(a.ws)[((int64_t)split * R + r) * ncols + c] = s;
            }
        }
    }
}

// dot_contig: the whole-tensor inner product of two unit-stride operands (dot / true_dot / energy-style contractions to a scalar).
// One CTA per K segment, 8 independent element pairs per thread in flight, fixed-order reduction; partials go to a.ws[split].
template <typename T> __global__ void __launch_bounds__(256) dot_contig(const __grid_constant__ ReduceArgs a) {
    constexpr int NT = 256, NW = NT / 32, UNR = 8;
    __shared__ T  red[NW];
    const T      *A = reinterpret_cast<const T *>(a.A);
    const T      *B = reinterpret_cast<const T *>(a.B);
    const int64_t K = a.k.tot;
    const int64_t k0 = K * blockIdx.x / a.splits, k1 = K * ((int64_t)blockIdx.x + 1) / a.splits;
    const bool    ca = a.conj_a != 0, cb = a.conj_b != 0;
    T             acc[UNR];
#pragma unroll
    for (int u = 0; u < UNR; u++) acc[u] = zero_of<T>();
    int64_t k = k0 + threadIdx.x;
    if constexpr (sizeof(T) == 8 && !Scalar<T>::is_complex) {
        // 16-byte loads when both segments start on a 16-byte boundary: half the load instructions for the same bytes in flight
        if ((((uintptr_t)(A + k0) | (uintptr_t)(B + k0)) & 15) == 0) {
            struct alignas(16) P2 { T x, y; };
            const P2     *A2 = reinterpret_cast<const P2 *>(A + k0), *B2 = reinterpret_cast<const P2 *>(B + k0);
            const int64_t np = (k1 - k0) / 2;
            int64_t       q = threadIdx.x;
            for (; q + (int64_t)(UNR - 1) * NT < np; q += (int64_t)UNR * NT) {
                P2 x[UNR], y[UNR];
#pragma unroll
                for (int u = 0; u < UNR; u++) { x[u] = A2[q + (int64_t)u * NT]; y[u] = B2[q + (int64_t)u * NT]; }
#pragma unroll
                for (int u = 0; u < UNR; u++) { acc[u] = fma_(x[u].x, y[u].x, acc[u]); acc[u] = fma_(x[u].y, y[u].y, acc[u]); }
            }
            for (; q < np; q += NT) { acc[0] = fma_(A2[q].x, B2[q].x, acc[0]); acc[0] = fma_(A2[q].y, B2[q].y, acc[0]); }
            k = k0 + 2 * np + threadIdx.x; // an odd tail element (at most one) goes through the scalar loop below
        }
    }
    for (; k + (int64_t)(UNR - 1) * NT < k1; k += (int64_t)UNR * NT) {
        T x[UNR], y[UNR];
#pragma unroll
        for (int u = 0; u < UNR; u++) { x[u] = A[k + (int64_t)u * NT]; y[u] = B[k + (int64_t)u * NT]; }
#pragma unroll
        for (int u = 0; u < UNR; u++) acc[u] = fma_(conj_if(x[u], ca), conj_if(y[u], cb), acc[u]);
    }
    for (; k < k1; k += NT) acc[0] = fma_(conj_if(A[k], ca), conj_if(B[k], cb), acc[0]);
    T sum = acc[0];
#pragma unroll
    for (int u = 1; u < UNR; u++) sum = add_(sum, acc[u]);
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        T t = red[0];
#pragma unroll
        for (int w = 1; w < NW; w++) t = add_(t, red[w]);
        reinterpret_cast<T *>(a.ws)[blockIdx.x] = t;
    }
}

// col_reduce: threads run along rows (contiguous in A); each thread owns RT rows and walks a K segment.
template <typename T, int NN, int RT>
__global__ void __launch_bounds__(256) col_reduce(const __grid_constant__ ReduceArgs a) {
    constexpr int NT = 256;
    const int     tid = threadIdx.x;
    const int64_t R = a.rows.tot, K = a.k.tot;
    const int     ncols = (int)a.n.tot;
    const int     split = blockIdx.y;
    const int64_t k_begin = K * split / a.splits, k_end = K * (split + 1) / a.splits;
    const T      *A = reinterpret_cast<const T *>(a.A);
    const T      *B = reinterpret_cast<const T *>(a.B);

    int64_t rowA[RT], rowB[RT];
#pragma unroll
    for (int j = 0; j < RT; j++) {
        int64_t r = (int64_t)blockIdx.x * NT * RT + (int64_t)j * NT + tid, oa = 0, ob = 0, oc = 0;
        if (r < R) grp_offset3(a.rows, r, oa, ob, oc);
        rowA[j] = (r < R) ? oa : OFF_INVALID;
        rowB[j] = ob;
    }
    int64_t colB[NN];
#pragma unroll
    for (int c = 0; c < NN; c++) colB[c] = (c < ncols) ? grp_offset(a.n, c, ROLE_B) : 0;

    T acc[RT][NN];
#pragma unroll
    for (int j = 0; j < RT; j++)
#pragma unroll
        for (int c = 0; c < NN; c++) acc[j][c] = zero_of<T>();

    // uniform incremental walk over the K multi-index
    int64_t idx[MAXG], offA = 0, offB = 0;
    {
        int64_t x = k_begin;
        for (int i = 0; i < MAXG; i++) {
            idx[i] = 0;
            if (i < a.k.n) {
                int64_t q = x / a.k.ext[i];
                idx[i] = x - q * a.k.ext[i];
                offA += idx[i] * a.k.s[ROLE_A][i];
                offB += idx[i] * a.k.s[ROLE_B][i];
                x = q;
            }
        }
    }
    int64_t k = k_begin;
    if (a.k.n == 1) { // a single contracted index (plain gemv): four k positions per iteration, all loads issued before the FMAs
        constexpr int UK = 4;
        const int64_t sA = a.k.s[ROLE_A][0], sB = a.k.s[ROLE_B][0];
        for (; k + UK <= k_end; k += UK) {
            T av[UK][RT], bv[UK][RT][NN];
#pragma unroll
            for (int u = 0; u < UK; u++)
#pragma unroll
                for (int j = 0; j < RT; j++) {
                    av[u][j] = (rowA[j] != OFF_INVALID) ? A[rowA[j] + offA + u * sA] : zero_of<T>();
#pragma unroll
                    for (int c = 0; c < NN; c++) bv[u][j][c] = (c < ncols) ? B[rowB[j] + offB + u * sB + colB[c]] : zero_of<T>();
                }
#pragma unroll
            for (int u = 0; u < UK; u++)
#pragma unroll
                for (int j = 0; j < RT; j++)
#pragma unroll
                    for (int c = 0; c < NN; c++) acc[j][c] = fma_(conj_if(av[u][j], a.conj_a != 0), conj_if(bv[u][j][c], a.conj_b != 0), acc[j][c]);
            offA += UK * sA;
            offB += UK * sB;
        }
        idx[0] = k; // k.n == 1: the running index is the position itself (the generic loop below finishes the tail)
    }
    for (; k < k_end; k++) {
        T av[RT];
#pragma unroll
        for (int j = 0; j < RT; j++) av[j] = (rowA[j] != OFF_INVALID) ? conj_if(A[rowA[j] + offA], a.conj_a != 0) : zero_of<T>();
#pragma unroll
        for (int c = 0; c < NN; c++) {
            if (c >= ncols) break;
#pragma unroll
            for (int j = 0; j < RT; j++) {
                T bv = conj_if(B[rowB[j] + offB + colB[c]], a.conj_b != 0);
                acc[j][c] = fma_(av[j], bv, acc[j][c]);
            }
        }
        // advance
        offA += a.k.s[ROLE_A][0];
        offB += a.k.s[ROLE_B][0];
        if (++idx[0] == a.k.ext[0]) {
#pragma unroll
            for (int i = 0; i < MAXG - 1; i++) {
                if (i < a.k.n && idx[i] == a.k.ext[i]) {
                    offA -= idx[i] * a.k.s[ROLE_A][i];
                    offB -= idx[i] * a.k.s[ROLE_B][i];
                    idx[i] = 0;
                    if (i + 1 < a.k.n) {
                        idx[i + 1]++;
                        offA += a.k.s[ROLE_A][i + 1];
                        offB += a.k.s[ROLE_B][i + 1];
                    }
                }
            }
        }
    }
    const T    alpha = make_scalar<T>(a.alpha[0], a.alpha[1]), beta = make_scalar<T>(a.beta[0], a.beta[1]);
    const bool has_beta = !is_zero(beta);
#pragma unroll
    for (int j = 0; j < RT; j++) {
        int64_t r = (int64_t)blockIdx.x * NT * RT + (int64_t)j * NT + tid;
        if (r >= R) continue;
#pragma unroll
        for (int c = 0; c < NN; c++) {
            if (c >= ncols) break;
            if (a.splits == 1) {
                T *dst = reinterpret_cast<T *>(a.C) + grp_offset(a.rows, r, ROLE_C) + grp_offset(a.n, c, ROLE_C);
                *dst = axpby_(alpha, acc[j][c], beta, has_beta ? *dst : zero_of<T>(), has_beta);
            } else {
                reinterpret_cast<T *>(a.ws)[((int64_t)split * R + r) * ncols + c] = acc[j][c];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// elementwise (no contracted index) and whole-tensor scaling
// ------------------------------------------------------------------------------------------
struct ElemArgs {
    Grp         t; // all target indices, inner-first by C stride; strides (A,B,C)
    const void *A;
    const void *B;
    void       *C;
    double      alpha[2], beta[2];
};

template <typename T, bool CONJ_A, bool CONJ_B> __global__ void __launch_bounds__(256) elementwise(const __grid_constant__ ElemArgs a) {
    const T      *A = reinterpret_cast<const T *>(a.A);
    const T      *B = reinterpret_cast<const T *>(a.B);
    T            *C = reinterpret_cast<T *>(a.C);
    const T       alpha = make_scalar<T>(a.alpha[0], a.alpha[1]), beta = make_scalar<T>(a.beta[0], a.beta[1]);
    const bool    has_beta = !is_zero(beta);
    const int64_t total = a.t.tot;
    for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (int64_t)gridDim.x * blockDim.x) {
        int64_t oa, ob, oc;
        grp_offset3(a.t, x, oa, ob, oc);
        T v = mul_(maybe_conj<CONJ_A>(A[oa]), maybe_conj<CONJ_B>(B[ob]));
        C[oc] = axpby_(alpha, v, beta, has_beta ? C[oc] : zero_of<T>(), has_beta);
    }
}

// elementwise_pairs: the same operation when the innermost target index is unit-stride in C with an even extent and every other C
// stride is even (outer products, direct products on contiguous tensors): a thread owns two neighbouring elements (one 16-byte C
// access for 8-byte types), decodes the index once per pair and keeps UNR = 4 pairs in flight (8 was tried for ger, where C is the only
// large stream: no gain, 4.7 TB/s either way). A and B inner strides may be 0 or anything.
template <typename T, bool CONJ_A, bool CONJ_B, int UNR> __global__ void __launch_bounds__(256) elementwise_pairs(const __grid_constant__ ElemArgs a) {
    struct alignas(2 * sizeof(T)) P2 { T x, y; };
    const T      *A = reinterpret_cast<const T *>(a.A);
    const T      *B = reinterpret_cast<const T *>(a.B);
    T            *C = reinterpret_cast<T *>(a.C);
    const T       alpha = make_scalar<T>(a.alpha[0], a.alpha[1]), beta = make_scalar<T>(a.beta[0], a.beta[1]);
    const bool    has_beta = !is_zero(beta);
    const int64_t pairs = a.t.tot / 2, stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t sa0 = a.t.s[0][0], sb0 = a.t.s[1][0];
    for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < pairs; x += stride * UNR) {
        int64_t oc[UNR];
        T       a0[UNR], a1[UNR], b0[UNR], b1[UNR];
        P2      c[UNR];
        // decode and issue the loads of one pair at a time: the A / B offsets die at once (all three offset sets of four pairs kept live
        // cost 118 registers = 2 CTAs per SM, and a C-only stream such as ger then had too few bytes in flight: 4.7 TB/s)
#pragma unroll
        for (int u = 0; u < UNR; u++) {
            const int64_t xx = x + u * stride;
            oc[u] = -1;
            if (xx < pairs) {
                int64_t oa, ob;
                grp_offset3(a.t, 2 * xx, oa, ob, oc[u]);
                a0[u] = A[oa]; a1[u] = A[oa + sa0];
                b0[u] = B[ob]; b1[u] = B[ob + sb0];
                if (has_beta) c[u] = *reinterpret_cast<const P2 *>(C + oc[u]);
                else c[u].x = c[u].y = zero_of<T>();
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; u++)
            if (oc[u] >= 0) {
                P2 r;
                r.x = axpby_(alpha, mul_(maybe_conj<CONJ_A>(a0[u]), maybe_conj<CONJ_B>(b0[u])), beta, c[u].x, has_beta);
                r.y = axpby_(alpha, mul_(maybe_conj<CONJ_A>(a1[u]), maybe_conj<CONJ_B>(b1[u])), beta, c[u].y, has_beta);
                *reinterpret_cast<P2 *>(C + oc[u]) = r;
            }
    }
}

// elementwise_pairs_2d: the pair kernel for TWO target indices (outer products / ger, direct products and broadcasts on matrices).
// One small CTA (128 threads) owns 256 columns x 2 rows: a thread loads its two 16-byte C pairs and the operand values, no index decode and
// no loop (ncu on dger with the generic pair kernel: two thirds of the warp samples in the multiply-shift decode, a third waiting on the C
// loads, 16 warps per SM). ~32 registers, so 2048 threads per SM are resident, like the contiguous level-1 kernels.
template <typename T, bool CONJ_A, bool CONJ_B> __global__ void __launch_bounds__(128) elementwise_pairs_2d(const __grid_constant__ ElemArgs a) {
    struct alignas(2 * sizeof(T)) P2 { T x, y; };
    constexpr int ROWS = 2;
    const T      *A = reinterpret_cast<const T *>(a.A);
    const T      *B = reinterpret_cast<const T *>(a.B);
    T            *C = reinterpret_cast<T *>(a.C);
    const T       alpha = make_scalar<T>(a.alpha[0], a.alpha[1]), beta = make_scalar<T>(a.beta[0], a.beta[1]);
    const bool    has_beta = !is_zero(beta);
    const int64_t e0 = a.t.ext[0], e1 = a.t.ext[1];
    const int64_t sa0 = a.t.s[0][0], sa1 = a.t.s[0][1], sb0 = a.t.s[1][0], sb1 = a.t.s[1][1], sc1 = a.t.s[2][1];
    const int64_t c0 = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (c0 >= e0) return;
    const int64_t r0 = (int64_t)blockIdx.y * ROWS;
    P2 c[ROWS];
    T  a0[ROWS], a1[ROWS], b0[ROWS], b1[ROWS];
#pragma unroll
    for (int u = 0; u < ROWS; u++) {
        const int64_t r = r0 + u;
        if (r < e1) {
            if (has_beta) c[u] = *reinterpret_cast<const P2 *>(C + c0 + r * sc1);
            a0[u] = A[c0 * sa0 + r * sa1]; a1[u] = A[(c0 + 1) * sa0 + r * sa1];
            b0[u] = B[c0 * sb0 + r * sb1]; b1[u] = B[(c0 + 1) * sb0 + r * sb1];
        }
    }
#pragma unroll
    for (int u = 0; u < ROWS; u++) {
        const int64_t r = r0 + u;
        if (r < e1) {
            P2 o;
            o.x = axpby_(alpha, mul_(maybe_conj<CONJ_A>(a0[u]), maybe_conj<CONJ_B>(b0[u])), beta, has_beta ? c[u].x : zero_of<T>(), has_beta);
            o.y = axpby_(alpha, mul_(maybe_conj<CONJ_A>(a1[u]), maybe_conj<CONJ_B>(b1[u])), beta, has_beta ? c[u].y : zero_of<T>(), has_beta);
            *reinterpret_cast<P2 *>(C + c0 + r * sc1) = o;
        }
    }
}

struct ScaleArgs {
    Grp    t; // ROLE_C strides only
    void  *C;
    double beta[2];
};
template <typename T> __global__ void __launch_bounds__(256) scale_tensor(const __grid_constant__ ScaleArgs a) {
    T            *C = reinterpret_cast<T *>(a.C);
    const T       beta = make_scalar<T>(a.beta[0], a.beta[1]);
    const bool    zero = is_zero(beta);
    const int64_t total = a.t.tot;
    for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (int64_t)gridDim.x * blockDim.x) {
        int64_t oc = grp_offset(a.t, x, ROLE_C);
        C[oc] = zero ? zero_of<T>() : mul_(beta, C[oc]);
    }
}

} // namespace eb200
