// gemm_tma_z.cuh — complex FP64 (c128) GEMM on the FP64 tensor pipe: TMA-staged tiles, mbarrier pipeline, four real DMMA products per
// complex product. Serves eb200_zgemm and every c128 contraction the planner reduces to a plain (single M, N, K index) product; the
// FMA-pipe tile kernel (tile_simt, 24 TFLOP/s of real flops) remains for multi-index / batched / split-K complex shapes.
//
//   * Operands stay interleaved (re, im) as the C ABI stores them. A tensor map sees a complex matrix as FP64 with a doubled inner
//     extent; a 128-byte swizzle row holds 8 complex elements. K-major operands arrive as one box {8 k, BM rows}, M/N-major operands
//     as boxes {8 m, 8 k}.
//   * One ld.shared.v2.f64 fetches (re, im) of a fragment element; the k of a k4 step are taken as k = 2*(lane%4) + kk (kk = 0, 1), which
//     keeps the 16-byte fragment reads of a quarter warp on 8 different 16-byte columns under the swizzle for both tile orientations.
//   * C_re += a_re*b_re + (-a_im)*b_im ; C_im += a_re*b_im + a_im*b_re : four DMMA.8x8x4 per 8x8 fragment pair, sign flips (and the
//     conjugation flags) are integer XORs on the sign bit, so the FP64 pipe executes nothing but DMMA in the main loop.
//   * CTA tile 128 x 32 complex, BK = 8 complex, 4 stages (80 KB), 8 warps of 32 x 16 (64 accumulator registers), 2 CTAs per SM.
#pragma once
#include <cuda.h>

#include "gemm_tma.cuh"

namespace eb200 {

struct TmaZGemmArgs {
    int64_t  M, N, K;
    int64_t  sCm, sCn; // strides of C in complex elements
    double2 *C;
    double   alpha[2], beta[2];
    int32_t  tiles_m, tiles_n, conj_a, conj_b;
};

__device__ __forceinline__ double flip_sign_if(double x, uint32_t mask_hi) { // mask_hi = 0x80000000 or 0
    return __hiloint2double(__double2hiint(x) ^ (int)mask_hi, __double2loint(x));
}

template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(256, 2) gemm_tma_z(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                     const __grid_constant__ TmaZGemmArgs args) {
    constexpr int BM = 128, BN = 32, BK = 8, STAGES = 4, NCONS = 8;
    constexpr int A_BYTES = BM * BK * 16, B_BYTES = BN * BK * 16, STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int FM = 4, FN = 2; // warp tile 32 x 16 complex

    extern __shared__ unsigned char smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const uint32_t bars = smem_base + STAGES * STAGE_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
    const int nsteps = (int)((args.K + BK - 1) / BK);

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), NCONS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // tensor-map coordinates are in doubles along the inner (interleaved) dimension
    const int m0_tile = tile_m * BM, n0_tile = tile_n * BN;
    auto issue_loads = [&](int it) {
        const int      s = it % STAGES;
        const uint32_t a_dst = smem_base + s * STAGE_BYTES, b_dst = a_dst + A_BYTES;
        const int      k0 = it * BK;
        mbar_expect_tx(full_bar(s), STAGE_BYTES);
        if constexpr (A_KMAJOR) tma_load_2d(a_dst, &tmA, full_bar(s), 2 * k0, m0_tile); // box {8 k, 128 m}
        else {
#pragma unroll
            for (int b = 0; b < BM / 8; b++) tma_load_2d(a_dst + b * 1024, &tmA, full_bar(s), 2 * (m0_tile + 8 * b), k0); // box {8 m, 8 k}
        }
        if constexpr (B_KMAJOR) tma_load_2d(b_dst, &tmB, full_bar(s), 2 * k0, n0_tile);
        else {
#pragma unroll
            for (int b = 0; b < BN / 8; b++) tma_load_2d(b_dst + b * 1024, &tmB, full_bar(s), 2 * (n0_tile + 8 * b), k0);
        }
    };
    if (tid == 0)
        for (int it = 0; it < STAGES - 1 && it < nsteps; it++) issue_loads(it);

    const int wm0 = (warp % 4) * 32, wn0 = (warp / 4) * 16;
    const int r = lane >> 2, j = lane & 3;
    // byte offsets of this lane's fragment element for the two k4 sub-steps of a stage (k = 2j + kk)
    uint32_t a_off[2], b_off[2];
#pragma unroll
    for (int kk = 0; kk < 2; kk++) {
        const int k = 2 * j + kk;
        a_off[kk] = A_KMAJOR ? (uint32_t)((wm0 + r) * 128 + ((k ^ r) << 4)) : (uint32_t)((wm0 >> 3) * 1024 + k * 128 + ((r ^ k) << 4));
        b_off[kk] = B_KMAJOR ? (uint32_t)((wn0 + r) * 128 + ((k ^ r) << 4)) : (uint32_t)((wn0 >> 3) * 1024 + k * 128 + ((r ^ k) << 4));
    }
    const uint32_t ca = args.conj_a ? 0x80000000u : 0u, cb = args.conj_b ? 0x80000000u : 0u;

    double acc_re[FM][FN][2], acc_im[FM][FN][2];
#pragma unroll
    for (int i = 0; i < FM; i++)
#pragma unroll
        for (int jj = 0; jj < FN; jj++) acc_re[i][jj][0] = acc_re[i][jj][1] = acc_im[i][jj][0] = acc_im[i][jj][1] = 0.0;

    for (int it = 0; it < nsteps; it++) {
        const int s = it % STAGES;
        if (tid == 0) {
            const int nxt = it + STAGES - 1;
            if (nxt < nsteps) {
                if (nxt >= STAGES) mbar_wait(empty_bar(nxt % STAGES), ((nxt / STAGES) - 1) & 1);
                issue_loads(nxt);
            }
        }
        __syncwarp();
        mbar_wait(full_bar(s), (it / STAGES) & 1);
        const uint32_t a_base = smem_base + s * STAGE_BYTES, b_base = a_base + A_BYTES;
#pragma unroll
        for (int kk = 0; kk < 2; kk++) {
            double a_re[FM], a_im[FM], a_nim[FM], b_re[FN], b_im[FN];
#pragma unroll
            for (int i = 0; i < FM; i++) { // fragment i: rows wm0 + 8 i .. + 7 (K-major: 8 rows = 1024 bytes; M-major: the next 8-m box = 1024 bytes)
                const double2 v = lds_f64x2(a_base + a_off[kk] + i * 1024);
                a_re[i] = v.x;
                a_im[i] = flip_sign_if(v.y, ca);
                a_nim[i] = flip_sign_if(v.y, ca ^ 0x80000000u);
            }
#pragma unroll
            for (int jj = 0; jj < FN; jj++) {
                const double2 v = lds_f64x2(b_base + b_off[kk] + jj * 1024);
                b_re[jj] = v.x;
                b_im[jj] = flip_sign_if(v.y, cb);
            }
#pragma unroll
            for (int i = 0; i < FM; i++)
#pragma unroll
                for (int jj = 0; jj < FN; jj++) {
                    dmma884(acc_re[i][jj][0], acc_re[i][jj][1], a_re[i], b_re[jj]);
                    dmma884(acc_im[i][jj][0], acc_im[i][jj][1], a_re[i], b_im[jj]);
                    dmma884(acc_re[i][jj][0], acc_re[i][jj][1], a_nim[i], b_im[jj]);
                    dmma884(acc_im[i][jj][0], acc_im[i][jj][1], a_im[i], b_re[jj]);
                }
        }
        __threadfence_block(); // every fragment load has completed before the stage is handed back (see gemm_tma_f64)
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(s));
    }

    // ---- epilogue (C fragment: row = lane/4, cols = 2*(lane%4) + {0,1}): out = alpha * acc + beta * C in complex arithmetic
    const int64_t m_base = (int64_t)tile_m * BM + wm0 + r, n_base = (int64_t)tile_n * BN + wn0 + 2 * j;
    const double  ar = args.alpha[0], ai = args.alpha[1], br = args.beta[0], bi = args.beta[1];
    const bool    has_beta = br != 0.0 || bi != 0.0;
#pragma unroll
    for (int i = 0; i < FM; i++) {
        const int64_t m = m_base + 8 * i;
        if (m >= args.M) continue;
        double2 old[FN][2];
#pragma unroll
        for (int jj = 0; jj < FN; jj++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int64_t n = n_base + 8 * jj + e;
                old[jj][e] = make_double2(0.0, 0.0);
                if (has_beta && n < args.N) old[jj][e] = args.C[m * args.sCm + n * args.sCn];
            }
#pragma unroll
        for (int jj = 0; jj < FN; jj++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int64_t n = n_base + 8 * jj + e;
                if (n >= args.N) continue;
                const double xr = acc_re[i][jj][e], xi = acc_im[i][jj][e];
                double       vr = ar * xr - ai * xi, vi = ar * xi + ai * xr;
                if (has_beta) {
                    vr += br * old[jj][e].x - bi * old[jj][e].y;
                    vi += br * old[jj][e].y + bi * old[jj][e].x;
                }
                args.C[m * args.sCm + n * args.sCn] = make_double2(vr, vi);
            }
    }
}

} // namespace eb200
