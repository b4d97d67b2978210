// contract_tma.cuh — the generic FP64 contraction kernel on the TMA + mbarrier pipeline (kernel (b) of the north star).
//
// Replaces the device einsum_generic_algorithm of the reference (TensorAlgebra/Backends/GPUTensorAlgebra.hpp:28-82: one
// thread per output element, integer div/mod per operand access, atomicAdd) for every contraction whose operands have a
// unit-stride index: the index permutation is folded into <= 5-dimensional tensor maps, so the operands are never
// transposed and never gathered element by element.
//
//   * Every operand tile is a set of 16 x 16 boxes (2 KB, one 128-byte swizzle row per line). A box is addressed by the
//     coordinates of up to five tensor indices: {the operand's unit-stride index, the inner index of the other group,
//     outer target indices, outer contracted indices, batch indices}. Composite M / N / K positions are enumerated with the
//     INNER index padded to a multiple of 16, so a box never straddles two values of an outer index; ragged edges and
//     padding are zero-filled by the TMA out-of-bounds rule (no predicates anywhere in the main loop).
//   * Producer duty rotates over the warps: step t is issued by warp t % NWARPS, STAGES-2 steps ahead of its use (a dedicated
//     producer warp would cap the CTA at 96 registers per thread and spill the accumulators; a fixed producer would make its
//     warp the slowest one). Inside the duty warp lane b owns box b, so the twelve boxes of a step cost ONE coordinate
//     computation and ONE cp.async.bulk.tensor.5d instruction (SASS UTMALDG.5D) in SIMT. The warp walks the contracted indices
//     with counters (no division in the loop) and adds the per-step K coordinates to the lane's target coordinates.
//     full[s] mbarriers carry the transaction bytes, empty[s] one arrival per warp: no CTA-wide barrier in the main loop.
//   * Consumer warps run LDS.64 + DMMA.8x8x4 only, with the same swizzle-aware fragment addressing as gemm_tma_f64
//     (conflict-free for K-major and M-major tiles alike).
//   * Split-K is fused: every split CTA stores its partial tile (thread-major, fully coalesced), the LAST CTA to arrive on
//     the tile's counter adds the partials in split order 0..S-1 (bit-reproducible, whichever CTA happens to be last)
//     and runs the alpha/beta epilogue. No finalize kernel.
//   * Tile shape is a template parameter (warps x fragments); besides 128x64 (8 warps, 2 CTAs/SM) an 80x80 variant
//     (4 warps of 40x40, 3 CTAs/SM) serves outputs that 128x64 tiles cover badly (Wmnij: 400 x 400 = 5 x 5 tiles of 80).
#pragma once
#include <cuda.h>

#include "gemm_tma.cuh"

namespace eb200 {

constexpr int TG_MAXI = 3; // indices per group the kernel walks (M, N, K, batch)

struct TmaGenArgs {
    double       *C;
    double       *ws;        // split partials: [batch in launch][tile][split][BM*BN], thread-major
    unsigned int *counters;  // [batch in launch][tile], zeroed by the host before the launch
    double        alpha, beta;
    int32_t       splits, tiles_m, tiles_n, vec_c;
    int32_t       dist;      // prefetch distance in k steps (1 .. STAGES-1)
    int32_t       em[TG_MAXI], en[TG_MAXI], ek[TG_MAXI], eb[TG_MAXI]; // extents, inner first (unused slots = 1)
    int32_t       mblk0, nblk0, kstep0;                                // 16-blocks along the inner M / N / K index
    int32_t       KT;                                                  // composite k steps = kstep0 * ek[1] * ek[2]
    int64_t       MB, NB;                                              // composite 16-row block counts
    int32_t       dAm[TG_MAXI], dAk[TG_MAXI], dAb[TG_MAXI];            // tensor-map dimension of each index in A's map (-1: absent)
    int32_t       dBn[TG_MAXI], dBk[TG_MAXI], dBb[TG_MAXI];
    int64_t       sCm[TG_MAXI], sCn[TG_MAXI], sCb[TG_MAXI];            // element strides in C
    int64_t       batch0;                                              // first batch of this launch (grid.z <= 65535)
};

__device__ __forceinline__ double2 ld_cg_f64x2(const double2 *p) {
    double2 v;
    asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

template <int WARPS_M_, int WARPS_N_, int FM_, int FN_, int STAGES_, int MINB_> struct TmaTile {
    static constexpr int WARPS_M = WARPS_M_, WARPS_N = WARPS_N_, FM = FM_, FN = FN_, STAGES = STAGES_, MINB = MINB_;
    static constexpr int BM = WARPS_M * FM * 8, BN = WARPS_N * FN * 8, BK = 16;
    static constexpr int BMB = BM / 16, BNB = BN / 16;
    static constexpr int NCONS = WARPS_M * WARPS_N, NT = NCONS * 32;
    static constexpr int A_BYTES = BM * BK * 8, B_BYTES = BN * BK * 8, STAGE_BYTES = A_BYTES + B_BYTES;
    // tiles | mbarriers | per-block coordinates (8 ints each) | per-block decode (4 ints each) | C offset tables | flag
    static constexpr int    TAIL = 2 * STAGES * 8 + (BMB + BNB) * 32 + (BMB + BNB) * 16 + (BM + BN) * 8 + 16;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 + TAIL;
    static_assert(BM % 16 == 0 && BN % 16 == 0, "tiles are made of 16-row boxes");
};

template <class CFG, bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(CFG::NT, CFG::MINB) contract_tma_f64(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                                       const __grid_constant__ TmaGenArgs args) {
    constexpr int BM = CFG::BM, BN = CFG::BN, STAGES = CFG::STAGES, NCONS = CFG::NCONS, FM = CFG::FM, FN = CFG::FN;
    constexpr int BMB = CFG::BMB, BNB = CFG::BNB, A_BYTES = CFG::A_BYTES, STAGE_BYTES = CFG::STAGE_BYTES;
    constexpr int WM = FM * 8, WN = FN * 8, NTC = CFG::NT;

    extern __shared__ unsigned char smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char *gen_base = smem_dyn + (smem_base - smem_u32(smem_dyn));
    const uint32_t bars = smem_base + STAGES * STAGE_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    int32_t *coordA = reinterpret_cast<int32_t *>(gen_base + STAGES * STAGE_BYTES + 2 * STAGES * 8); // [BMB][8]
    int32_t *coordB = coordA + BMB * 8;                                                              // [BNB][8]
    int32_t *blkM = coordB + BNB * 8;                                                                // [BMB][4]: blk0, i1, i2, valid
    int32_t *blkN = blkM + BMB * 4;                                                                  // [BNB][4]
    int64_t *moffC = reinterpret_cast<int64_t *>(blkN + BNB * 4);                                    // [BM]
    int64_t *noffC = moffC + BM;                                                                     // [BN]
    int32_t *flag = reinterpret_cast<int32_t *>(noffC + BN);

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
    const int split = blockIdx.y;
    const int ks0 = (int)((int64_t)args.KT * split / args.splits), ks1 = (int)((int64_t)args.KT * (split + 1) / args.splits);
    const int nsteps = ks1 - ks0;

    // ---- once per CTA: decode the tile's 16-row blocks into tensor-map coordinates and C offsets
    int     bi[TG_MAXI];
    int64_t baseC = 0;
    {
        int64_t x = (int64_t)blockIdx.z + args.batch0;
#pragma unroll
        for (int i = 0; i < TG_MAXI; i++) {
            const int64_t q = x / args.eb[i];
            bi[i] = (int)(x - q * args.eb[i]);
            baseC += bi[i] * args.sCb[i];
            x = q;
        }
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), NCONS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int b = tid; b < BMB + BNB; b += CFG::NT) {
        const bool     is_m = b < BMB;
        const int      bb = is_m ? b : b - BMB;
        const int64_t  cb = (int64_t)(is_m ? tile_m : tile_n) * (is_m ? BMB : BNB) + bb;
        const int32_t *ext = is_m ? args.em : args.en;
        const int      blk0n = is_m ? args.mblk0 : args.nblk0;
        const bool     valid = cb < (is_m ? args.MB : args.NB);
        int            blk0 = blk0n, i1 = 0, i2 = 0; // out of range: an inner coordinate past the extent makes the whole box zero-fill
        if (valid) {
            const int64_t q0 = cb / blk0n;
            blk0 = (int)(cb - q0 * blk0n);
            const int64_t q1 = q0 / ext[1];
            i1 = (int)(q0 - q1 * ext[1]);
            i2 = (int)q1;
        }
        int32_t       *co = (is_m ? coordA : coordB) + bb * 8;
        const int32_t *dt = is_m ? args.dAm : args.dBn, *db = is_m ? args.dAb : args.dBb;
#pragma unroll
        for (int d = 0; d < 8; d++) co[d] = 0;
        co[dt[0]] = 16 * blk0;
        if (dt[1] >= 0) co[dt[1]] = i1;
        if (dt[2] >= 0) co[dt[2]] = i2;
#pragma unroll
        for (int i = 0; i < TG_MAXI; i++)
            if (db[i] >= 0) co[db[i]] = bi[i];
        int32_t *bk = (is_m ? blkM : blkN) + bb * 4;
        bk[0] = blk0; bk[1] = i1; bk[2] = i2; bk[3] = valid ? 1 : 0;
    }
    __syncthreads();
    for (int i = tid; i < BM + BN; i += CFG::NT) {
        const bool     is_m = i < BM;
        const int      r = is_m ? i : i - BM;
        const int32_t *bk = (is_m ? blkM : blkN) + (r >> 4) * 4;
        const int64_t *sc = is_m ? args.sCm : args.sCn;
        const int      x0 = bk[0] * 16 + (r & 15);
        int64_t        off = OFF_INVALID;
        if (bk[3] && x0 < (is_m ? args.em[0] : args.en[0])) off = x0 * sc[0] + bk[1] * sc[1] + bk[2] * sc[2];
        (is_m ? moffC : noffC)[r] = off;
    }
    __syncthreads();

    // ===================== producer duty (rotating) =====================
    // step t (local, 0-based) belongs to warp t % NCONS; a warp keeps the K counters of ITS next step and moves them NCONS steps at a time.
    // The boxes of a step are issued by DIFFERENT LANES of the duty warp in one SIMT pass (lane b owns box b: BMB boxes of A, then BNB
    // boxes of B): one coordinate computation and one UTMALDG instruction per step instead of twelve serial ones by a single lane.
    static_assert(BMB + BNB <= 32, "one lane per box");
    int my_step = warp, s0 = 0, k1 = 0, k2 = 0;
    {
        const int t = ks0 + warp, q0 = t / args.kstep0;
        s0 = t - q0 * args.kstep0;
        k2 = q0 / args.ek[1];
        k1 = q0 - k2 * args.ek[1];
    }
    auto issue_my_step = [&]() { // the whole duty warp
        const int s = my_step % STAGES;
        if (lane == 0) { // one lane polls the barrier (32 lanes polling load the mbarrier unit 32 x), the others park at the __syncwarp
            if (my_step >= STAGES) mbar_wait(empty_bar(s), ((my_step / STAGES) - 1) & 1);
            mbar_expect_tx(full_bar(s), STAGE_BYTES);
        }
        __syncwarp();
        if (lane < BMB + BNB) {
            // this lane's box: operand, slot inside the stage, base coordinates (target + batch part, decoded once per CTA into shared
            // memory: re-read here rather than kept in registers, the accumulators need them), K dimension slots
            const bool     is_a = lane < BMB;
            const int      b = is_a ? lane : lane - BMB;
            const int32_t *co = (is_a ? coordA : coordB) + b * 8;
            const int4     c03 = *reinterpret_cast<const int4 *>(co);
            const int      kd0 = is_a ? args.dAk[0] : args.dBk[0], kd1 = is_a ? args.dAk[1] : args.dBk[1], kd2 = is_a ? args.dAk[2] : args.dBk[2];
            int            c[5] = {c03.x, c03.y, c03.z, c03.w, co[4]};
#pragma unroll
            for (int d = 0; d < 5; d++) c[d] += (kd0 == d ? 16 * s0 : 0) + (kd1 == d ? k1 : 0) + (kd2 == d ? k2 : 0);
            tma_load_5d(smem_base + s * STAGE_BYTES + (is_a ? 0u : (uint32_t)A_BYTES) + (uint32_t)b * 2048u, is_a ? &tmA : &tmB, full_bar(s), c[0], c[1],
                        c[2], c[3], c[4]);
        }
        my_step += NCONS;
        s0 += NCONS;
        while (s0 >= args.kstep0) {
            s0 -= args.kstep0;
            if (++k1 == args.ek[1]) { k1 = 0; ++k2; }
        }
    };
    // Prefetch distance: the loads of step it + dist are issued while step it is consumed. dist = STAGES - 1 keeps every stage busy but
    // makes the duty warp wait until the SLOWEST warp has left step it - 1, every step (ncu: 22 % of the samples in that spin, tensor
    // pipe 85 %); dist = STAGES - 2 refills the stage everybody left one step earlier, so the wait is over before it starts.
    const int dist = args.dist;
    while (my_step < dist && my_step < nsteps) issue_my_step(); // prologue: the first `dist` steps

    // ===================== consumers =====================
    const int wm0 = (warp % CFG::WARPS_M) * WM, wn0 = (warp / CFG::WARPS_M) * WN;
    const int r = lane >> 2, j = lane & 3;
    // byte offset of this lane's fragment element inside a 16-row block for the four k4 sub-steps (swizzle and the k permutation
    // of gemm_tma_f64 folded in); fragment i then adds its block (2 KB apiece) and, for M-major tiles, flips the 64-byte half
    uint32_t a_off[4], b_off[4];
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
        const int u = (j + kk) & 3, v = j >> 1, p = j & 1;
        const int k = 2 * u + 8 * v + p;
        const uint32_t kmaj = (uint32_t)(r * 128 + ((((k >> 1) ^ r) & 7) << 4) + (k & 1) * 8);      // row = r (+8 for odd fragments), 16 k per row
        const uint32_t mmaj = (uint32_t)(k * 128 + ((((r >> 1) ^ (k & 7)) & 7) << 4) + (r & 1) * 8); // row = k, 16 m per row
        a_off[kk] = A_KMAJOR ? kmaj : mmaj;
        b_off[kk] = B_KMAJOR ? kmaj : mmaj;
    }
    // fragment i of the warp covers tile rows t = w0 + 8 i .. + 7: block t >> 4, upper half of the block when (t >> 3) & 1
    auto frag_addr = [](uint32_t base, uint32_t off, int t, bool kmajor) -> uint32_t {
        const uint32_t blk = (uint32_t)(t >> 4) * 2048u, hi = (uint32_t)((t >> 3) & 1);
        // K-major: rows 8..15 of the block are 1024 bytes further (their swizzle phase (row & 7) is the same);
        // M-major: elements 8..15 of a 128-byte row are the upper 64 bytes: flips bit 6 of the (already swizzled) offset
        return kmajor ? base + blk + off + hi * 1024u : base + blk + (off ^ (hi * 64u));
    };

    // warp tiles that start on a 16-row block boundary (the 128x64 configuration): the block of fragment i is a compile-time offset from
    // the warp's first block, as in gemm_tma_f64 -- LDS with immediate offsets instead of an address computation per fragment load
    constexpr bool ALIGNED_WARP_TILES = (WM % 16 == 0) && (WN % 16 == 0);
    if constexpr (ALIGNED_WARP_TILES) {
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
            a_off[kk] += (uint32_t)(wm0 >> 4) * 2048u;
            b_off[kk] += (uint32_t)(wn0 >> 4) * 2048u;
        }
    }
    auto frag_a = [&](uint32_t base, uint32_t off, int i) -> uint32_t {
        if constexpr (ALIGNED_WARP_TILES) return A_KMAJOR ? base + off + (uint32_t)i * 1024u : base + ((off + (uint32_t)(i >> 1) * 2048u) ^ ((uint32_t)(i & 1) * 64u));
        else return frag_addr(base, off, wm0 + 8 * i, A_KMAJOR);
    };
    auto frag_b = [&](uint32_t base, uint32_t off, int jj) -> uint32_t {
        if constexpr (ALIGNED_WARP_TILES) return B_KMAJOR ? base + off + (uint32_t)jj * 1024u : base + ((off + (uint32_t)(jj >> 1) * 2048u) ^ ((uint32_t)(jj & 1) * 64u));
        else return frag_addr(base, off, wn0 + 8 * jj, B_KMAJOR);
    };

    double acc[FM][FN][2];
#pragma unroll
    for (int i = 0; i < FM; i++)
#pragma unroll
        for (int jj = 0; jj < FN; jj++) acc[i][jj][0] = acc[i][jj][1] = 0.0;

    for (int it = 0; it < nsteps; it++) {
        const int s = it % STAGES;
        if (my_step == it + dist && my_step < nsteps) issue_my_step(); // (warp-uniform) refills the stage everybody left at it + dist - STAGES
        __syncwarp();
        mbar_wait(full_bar(s), (it / STAGES) & 1);
        const uint32_t a_base = smem_base + s * STAGE_BYTES, b_base = a_base + A_BYTES;
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
            double a[FM], b[FN];
#pragma unroll
            for (int i = 0; i < FM; i++) a[i] = lds_f64(frag_a(a_base, a_off[kk], i));
#pragma unroll
            for (int jj = 0; jj < FN; jj++) b[jj] = lds_f64(frag_b(b_base, b_off[kk], jj));
#pragma unroll
            for (int i = 0; i < FM; i++)
#pragma unroll
                for (int jj = 0; jj < FN; jj++) dmma884(acc[i][jj][0], acc[i][jj][1], a[i], b[jj]);
        }
        // hand the stage back only when every fragment load has completed (see gemm_tma_f64)
        __threadfence_block();
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(s));
    }

    // ---- fused split-K: partial tiles to scratch; the last CTA of the tile sums them in split order
    if (args.splits > 1) {
        const int64_t  tile_id = (int64_t)blockIdx.z * gridDim.x + blockIdx.x;
        double2       *part = reinterpret_cast<double2 *>(args.ws) + (tile_id * args.splits) * (int64_t)(BM * BN / 2);
        double2       *mine = part + (int64_t)split * (BM * BN / 2);
#pragma unroll
        for (int i = 0; i < FM; i++)
#pragma unroll
            for (int jj = 0; jj < FN; jj++) mine[(i * FN + jj) * NTC + tid] = make_double2(acc[i][jj][0], acc[i][jj][1]);
        __threadfence();
        __syncthreads();
        if (tid == 0) *flag = (atomicAdd(args.counters + tile_id, 1u) == (unsigned)(args.splits - 1));
        __syncthreads();
        if (!*flag) return;
        __threadfence();
#pragma unroll
        for (int i = 0; i < FM; i++)
#pragma unroll
            for (int jj = 0; jj < FN; jj++) acc[i][jj][0] = acc[i][jj][1] = 0.0;
        for (int sp = 0; sp < args.splits; sp++) {
            const double2 *src = part + (int64_t)sp * (BM * BN / 2);
#pragma unroll
            for (int i = 0; i < FM; i++)
#pragma unroll
                for (int jj = 0; jj < FN; jj++) {
                    const double2 v = ld_cg_f64x2(src + (i * FN + jj) * NTC + tid);
                    acc[i][jj][0] += v.x;
                    acc[i][jj][1] += v.y;
                }
        }
    }

    // ---- epilogue from registers (C fragment: row = lane/4, cols = 2*(lane%4) + {0,1}); addresses from the offset tables
    double      *C = args.C + baseC;
    const double alpha = args.alpha, beta = args.beta;
    const bool   has_beta = beta != 0.0;
    constexpr int HM = (FM * FN > 16) ? 1 : (FM + 1) / 2; // fragment rows per pass (all C loads of a pass are issued before their first use)
#pragma unroll
    for (int h = 0; h < FM; h += HM) {
        double  cold[HM][FN][2];
        int64_t offs[HM][FN];
#pragma unroll
        for (int ii = 0; ii < HM; ii++) {
            const int     i = h + ii;
            const int64_t om = (i < FM) ? moffC[wm0 + 8 * i + r] : OFF_INVALID;
#pragma unroll
            for (int jj = 0; jj < FN; jj++) {
                const int64_t on0 = noffC[wn0 + 8 * jj + 2 * j], on1 = noffC[wn0 + 8 * jj + 2 * j + 1];
                cold[ii][jj][0] = cold[ii][jj][1] = 0.0;
                if (args.vec_c) {
                    offs[ii][jj] = (om != OFF_INVALID && on0 != OFF_INVALID) ? om + on0 : OFF_INVALID;
                    if (has_beta && offs[ii][jj] != OFF_INVALID) {
                        const double2 v = *reinterpret_cast<const double2 *>(C + offs[ii][jj]);
                        cold[ii][jj][0] = v.x; cold[ii][jj][1] = v.y;
                    }
                } else {
                    offs[ii][jj] = om;
                    if (has_beta && om != OFF_INVALID) {
                        if (on0 != OFF_INVALID) cold[ii][jj][0] = C[om + on0];
                        if (on1 != OFF_INVALID) cold[ii][jj][1] = C[om + on1];
                    }
                }
            }
        }
#pragma unroll
        for (int ii = 0; ii < HM; ii++) {
            const int i = h + ii;
            if (i >= FM) continue;
#pragma unroll
            for (int jj = 0; jj < FN; jj++) {
                if (offs[ii][jj] == OFF_INVALID) continue;
                double v0 = alpha * acc[i][jj][0], v1 = alpha * acc[i][jj][1];
                if (has_beta) { v0 = fma(beta, cold[ii][jj][0], v0); v1 = fma(beta, cold[ii][jj][1], v1); }
                if (args.vec_c) {
                    *reinterpret_cast<double2 *>(C + offs[ii][jj]) = make_double2(v0, v1);
                } else {
                    const int64_t on0 = noffC[wn0 + 8 * jj + 2 * j], on1 = noffC[wn0 + 8 * jj + 2 * j + 1];
                    if (on0 != OFF_INVALID) C[offs[ii][jj] + on0] = v0;
                    if (on1 != OFF_INVALID) C[offs[ii][jj] + on1] = v1;
                }
            }
        }
    }
}

using TmaTile128x64 = TmaTile<4, 2, 4, 4, 4, 2>; // 8 consumer warps of 32x32, 4 stages (96 KB), 2 CTAs/SM
using TmaTile80x80 = TmaTile<2, 2, 5, 5, 3, 3>;  // 4 consumer warps of 40x40, 3 stages (60 KB), 3 CTAs/SM

} // namespace eb200
