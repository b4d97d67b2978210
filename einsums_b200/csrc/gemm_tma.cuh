// gemm_tma.cuh — TMA-staged, warp-specialised FP64 GEMM for contractions that the planner reduces to a plain
// (single M, N, K index) matrix product: Wabef, Wmnij, the sharded vvvv slabs, eb200_dgemm.
//
//   * operand tiles are brought in by the Tensor Memory Accelerator (cp.async.bulk.tensor.2d, SASS UTMALDG) from
//     2-D tensor maps with the hardware 128-byte swizzle; one elected thread issues them STAGES-1 steps ahead,
//     so the 8 warps execute nothing but LDS + DMMA in the main loop;
//   * producer/consumer hand-off uses mbarriers (full[s]: TMA complete_tx; empty[s]: one arrive per consumer warp):
//     no CTA-wide barrier in the main loop, warps drift apart and keep the FP64 tensor pipe busy;
//   * shared-memory tiles are dense 128-byte rows (no padding). The DMMA.8x8x4 fragment reads stay bank-conflict free
//     under the swizzle because the four k of a k4 step are taken in the permuted order
//         k(kk, j) = 2*((j + kk) & 3) + 8*(j >> 1) + (j & 1)          (j = lane % 4, kk = 0..3)
//     (a sum over k does not care about the order, and A and B use the same permutation); this works for both
//     K-major tiles (row = m, 16 k per row) and M-major tiles (row = k, 16 m per row), see DESIGN.md §2.1;
//   * FP64 has no tcgen05 kind: the math is mma.sync.m8n8k4.f64 (DMMA), accumulators in registers;
//   * tile 128x64x16, 4 stages (96 KB), 256 threads, 2 CTAs per SM; split-K, batches, multi-index groups and the 80x80 tile
//     are served by contract_tma_f64 (contract_tma.cuh).
#pragma once
#include <cuda.h>

#include "contract_kernels.cuh"
#include "tma_util.cuh"

namespace eb200 {

struct TmaGemmArgs {
    int64_t M, N, K;
    int64_t sCm, sCn;      // element strides of C
    double *C;
    double  alpha, beta;
    int32_t tiles_m, tiles_n, vec_c;
    int32_t dist;          // prefetch distance in k steps (1 .. STAGES-1), see contract_tma.cuh
    int32_t group;         // tile rows per rasterisation group (B panels are re-read from DRAM once per group when they exceed L2)
    int32_t prefetch_desc; // 1: prefetch.tensormap for both maps before the barrier set-up
};

template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(256, 2) gemm_tma_f64(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                       const __grid_constant__ TmaGemmArgs args) {
    constexpr int BM = 128, BN = 64, BK = 16, STAGES = 4, NCONS = 8;
    constexpr int A_BYTES = BM * BK * 8, B_BYTES = BN * BK * 8, STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int FM = 4, FN = 4; // 32x32 warp tile

    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte alignment for the swizzle atoms
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const uint32_t bars = smem_base + STAGES * STAGE_BYTES; // full[STAGES], empty[STAGES]
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    int tile_m, tile_n;
    {
        const int GROUP = args.group;
        int id = blockIdx.x, per_group = GROUP * args.tiles_n;
        int g = id / per_group, first = g * GROUP;
        int gsize = min(args.tiles_m - first, GROUP);
        int in_g = id - g * per_group;
        tile_m = first + in_g % gsize;
        tile_n = in_g / gsize;
    }
    // split-K and other tile shapes are served by contract_tma_f64 (contract_tma.cuh); this kernel always owns the whole K range
    const int ks0 = 0, nsteps = (int)((args.K + BK - 1) / BK);

    if (tid == 0) {
        // the descriptor fetch of this CTA's first loads overlaps the barrier set-up (EB200_TMA_PREFETCH_DESC=1; measured: no effect, profiles/r2/experiment_prefetch_tensormap_no_effect.txt)
        if (args.prefetch_desc) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        }
        for (int s = 0; s < STAGES; s++) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), NCONS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ===================== producer duty: lane 0 of warp 0 drives the TMA, STAGES-1 steps ahead =====================
    // (a dedicated producer warp would push the CTA to 288 threads and the consumers below 100 registers)
    const int m0_tile = tile_m * BM, n0_tile = tile_n * BN;
    auto issue_loads = [&](int it) {
        const int      s = it % STAGES;
        const uint32_t a_dst = smem_base + s * STAGE_BYTES, b_dst = a_dst + A_BYTES;
        const int      k0 = (ks0 + it) * BK;
        mbar_expect_tx(full_bar(s), STAGE_BYTES);
        if constexpr (A_KMAJOR) tma_load_2d(a_dst, &tmA, full_bar(s), k0, m0_tile); // box {16 k, 128 m}
        else {
#pragma unroll
            for (int b = 0; b < BM / 16; b++) tma_load_2d(a_dst + b * 2048, &tmA, full_bar(s), m0_tile + 16 * b, k0); // box {16 m, 16 k}
        }
        if constexpr (B_KMAJOR) tma_load_2d(b_dst, &tmB, full_bar(s), k0, n0_tile);
        else {
#pragma unroll
            for (int b = 0; b < BN / 16; b++) tma_load_2d(b_dst + b * 2048, &tmB, full_bar(s), n0_tile + 16 * b, k0);
        }
    };
    const int dist = args.dist;
    if (tid == 0)
        for (int it = 0; it < dist && it < nsteps; it++) issue_loads(it);

    // ===================== consumer warps =====================
    const int wm0 = (warp % 4) * 32, wn0 = (warp / 4) * 32;
    const int r = lane >> 2, j = lane & 3;
    // per-thread byte offsets of the fragment element for the four k4 sub-steps (swizzle + k permutation folded in)
    uint32_t a_off[4], b_off[4];
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
        const int u = (j + kk) & 3, v = j >> 1, p = j & 1;
        const int k = 2 * u + 8 * v + p;
        if constexpr (A_KMAJOR) a_off[kk] = (uint32_t)(wm0 * 128 + r * 128 + ((((k >> 1) ^ r) & 7) << 4) + (k & 1) * 8);
        else a_off[kk] = (uint32_t)((wm0 >> 4) * 2048 + k * 128 + ((((r >> 1) ^ (k & 7)) & 7) << 4) + (r & 1) * 8);
        if constexpr (B_KMAJOR) b_off[kk] = (uint32_t)(wn0 * 128 + r * 128 + ((((k >> 1) ^ r) & 7) << 4) + (k & 1) * 8);
        else b_off[kk] = (uint32_t)((wn0 >> 4) * 2048 + k * 128 + ((((r >> 1) ^ (k & 7)) & 7) << 4) + (r & 1) * 8);
    }

    double acc[FM][FN][2];
#pragma unroll
    for (int i = 0; i < FM; i++)
#pragma unroll
        for (int jj = 0; jj < FN; jj++) acc[i][jj][0] = acc[i][jj][1] = 0.0;

    for (int it = 0; it < nsteps; it++) {
        const int s = it % STAGES;
        if (tid == 0) {
            // refill the stage everybody finished STAGES - dist iterations ago
            const int nxt = it + dist;
            if (nxt < nsteps) {
                if (nxt >= STAGES) mbar_wait(empty_bar(nxt % STAGES), ((nxt / STAGES) - 1) & 1);
                issue_loads(nxt);
            }
        }
        __syncwarp();
        mbar_wait(full_bar(s), (it / STAGES) & 1);
        const uint32_t a_base = smem_base + s * STAGE_BYTES, b_base = a_base + A_BYTES;
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
            double a[FM], b[FN];
#pragma unroll
            for (int i = 0; i < FM; i++) {
                // fragment i covers rows wm0 + 8i .. +7
                if constexpr (A_KMAJOR) a[i] = lds_f64(a_base + a_off[kk] + i * 1024);
                else a[i] = lds_f64(a_base + ((a_off[kk] + (i >> 1) * 2048) ^ ((i & 1) * 64)));
            }
#pragma unroll
            for (int jj = 0; jj < FN; jj++) {
                if constexpr (B_KMAJOR) b[jj] = lds_f64(b_base + b_off[kk] + jj * 1024);
                else b[jj] = lds_f64(b_base + ((b_off[kk] + (jj >> 1) * 2048) ^ ((jj & 1) * 64)));
            }
#pragma unroll
            for (int i = 0; i < FM; i++)
#pragma unroll
                for (int jj = 0; jj < FN; jj++) dmma884(acc[i][jj][0], acc[i][jj][1], a[i], b[jj]);
        }
        // The stage may be handed back only when every fragment load has COMPLETED, not merely issued: ptxas schedules the
        // last LDS of the step right before the arrive (its consumer DMMA comes later), and a refill landing in that
        // window corrupted one fragment per ~10^6 warp-steps (found with tools/dbg_tma.py; tests/test_gpu_parity.py::
        // test_tma_gemm_stress guards it). A CTA-scope fence drains the outstanding loads first.
        __threadfence_block();
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(s));
    }

    // ---- epilogue from registers (C fragment: row = lane/4, cols = 2*(lane%4) + {0,1})
    const int64_t m_base = (int64_t)tile_m * BM + wm0 + r, n_base = (int64_t)tile_n * BN + wn0 + 2 * j;
    const double alpha = args.alpha, beta = args.beta;
    const bool   has_beta = beta != 0.0;
#pragma unroll
    for (int h = 0; h < FM; h += 2) {
        double cold[2][FN][2];
#pragma unroll
        for (int ii = 0; ii < 2; ii++) {
            const int64_t m = m_base + 8 * (h + ii);
#pragma unroll
            for (int jj = 0; jj < FN; jj++) {
                const int64_t n = n_base + 8 * jj;
                cold[ii][jj][0] = cold[ii][jj][1] = 0.0;
                if (has_beta && m < args.M) {
                    const double *src = args.C + m * args.sCm + n * args.sCn;
                    if (args.vec_c) {
                        if (n < args.N) {
                            double2 v = *reinterpret_cast<const double2 *>(src);
                            cold[ii][jj][0] = v.x; cold[ii][jj][1] = v.y;
                        }
                    } else {
                        if (n < args.N) cold[ii][jj][0] = src[0];
                        if (n + 1 < args.N) cold[ii][jj][1] = src[args.sCn];
                    }
                }
            }
        }
#pragma unroll
        for (int ii = 0; ii < 2; ii++) {
            const int64_t m = m_base + 8 * (h + ii);
            if (m >= args.M) continue;
#pragma unroll
            for (int jj = 0; jj < FN; jj++) {
                const int64_t n = n_base + 8 * jj;
                if (n >= args.N) continue;
                double v0 = alpha * acc[h + ii][jj][0], v1 = alpha * acc[h + ii][jj][1];
                if (has_beta) { v0 = fma(beta, cold[ii][jj][0], v0); v1 = fma(beta, cold[ii][jj][1], v1); }
                double *dst = args.C + m * args.sCm + n * args.sCn;
                if (args.vec_c) *reinterpret_cast<double2 *>(dst) = make_double2(v0, v1);
                else {
                    dst[0] = v0;
                    if (n + 1 < args.N) dst[args.sCn] = v1;
                }
            }
        }
    }
}

} // namespace eb200
