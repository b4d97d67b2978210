// gemm_tf32x3.cuh — FP32 GEMM on the 5th-generation tensor cores (tcgen05 / TMEM) with 3xTF32 split accumulation.
//
// FP32 results must match the CPU path to 1e-5; a single TF32 product (10-bit mantissa) cannot, so every operand is split on
// the fly into hi = trunc_tf32(x) and lo = x - hi (exact in FP32) and the tile product is formed as
//        A_hi*B_hi + A_hi*B_lo + A_lo*B_hi          (the lo*lo term is below 2^-20 relative and dropped)
// with all three tcgen05.mma accumulating into the same FP32 accumulator in tensor memory.
//
// Persistent kernel: one CTA per SM walks a static list of 128x128 output tiles (and batch members); BK = 32 floats = one 128-byte
// swizzle row; 320 threads, warp-specialised:
//   warp 0, one lane : TMA producer — cp.async.bulk.tensor.3d {k, m|n, batch} of the raw A and B tiles (SWIZZLE_128B), 3 stages that
//                      run straight across tile boundaries (the next tile's loads start while the last MMAs of this one are in flight)
//   warps 2..5       : splitters — read the raw tile and write lo = x - trunc_tf32(x) to a second buffer in the same swizzled layout
//                      (the raw tile itself serves as hi: kind::tf32 ignores the 13 low mantissa bits), fence.proxy.async, arrive
//   warp 1, one lane : tcgen05.mma issuer — 4 k-slices x 3 products per stage, kind::tf32, M=128 N=128 K=8, cta_group::1;
//                      tcgen05.commit releases the stage to the producer and hands finished accumulators to the epilogue warps
//   warps 6..9       : epilogue — tcgen05.ld 32x32b (each warp its own TMEM lane quadrant) -> alpha*acc + beta*C -> global, one tile
//                      behind the issuer, so the store of tile i overlaps the main loop of tile i + 1
// Two TMEM accumulator pairs (hi*hi | cross terms; all 512 columns) alternate over chunks of 512 k and over consecutive tiles; finished
// chunks are added into registers with round-to-nearest (the tensor core's own FP32 accumulation truncates). K <= 512
// uses the MULTI = false instantiation, whose epilogue streams 32-column blocks TMEM -> registers -> global without a row buffer.
#pragma once
#include <cuda.h>

#include "gemm_tma.cuh"

namespace eb200 {

struct Tf32GemmArgs {
    int64_t M, N, K;
    int64_t sCm, sCn;
    float  *C;
    float   alpha, beta;
    int32_t tiles_m, tiles_n;
    int64_t batch, sCb;    // number of independent products (1 = plain GEMM) and the element stride of C between them
    // ---- GENERIC instantiation only (multi-index M / N / K / batch groups, <= 5-D tensor maps with 32 x 32 boxes; see contract_tma.cuh
    //      for the scheme: composite positions are enumerated with the INNER index padded to a multiple of 32)
    int32_t em[3], en[3], ek[3], eb[3];                 // extents, inner first (unused slots = 1)
    int32_t mblk0, nblk0, kstep0, KT;                   // 32-blocks along the inner M / N / K index; composite k steps
    int64_t MB, NB;                                     // composite 32-row block counts
    int32_t dAm[3], dAk[3], dAb[3], dBn[3], dBk[3], dBb[3]; // tensor-map dimension of each index (-1: absent)
    int64_t gCm[3], gCn[3], gCb[3];                     // element strides in C
    int32_t c_n_unit;                                   // the inner N index is unit-stride in C: transposed (row-segment) stores
    // ---- c64 through the real kernel (contract.cu, try_run_c64_tcgen05): A' = packed [b][m][re,im interleaved along k], B' = two planes of
    //      rows (real-output rows | imaginary-output rows). A tile covers 64 complex columns: boxes 0,1 come from plane 0, boxes 2,3 from
    //      plane 1 of the same columns, so accumulator column j is the real and column j + 64 the imaginary part of complex column j.
    int32_t cplx;                                       // 1: complex-interleaved C (strides given in floats), 64 complex columns per tile
    int32_t dBc;                                        // tensor-map dimension of B'-plane index
    int32_t b_flat;                                     // the operands carry ONE merged batch dimension (dAb[0] / dBb[0]); C keeps eb / gCb
    int32_t c_vec4;                                     // plain kernel, sCn == 1: C rows, batch stride and N are multiples of 4 floats, C 16-byte aligned
};

#ifdef EB200_TF32_KERNEL_IMPL // the kernel itself is compiled by k_tile_tf32.cu only
// Shared-memory matrix descriptor, SWIZZLE_128B (layout type 2), descriptor version 1 (Blackwell):
//   start address >> 4 at bit 0 | leading byte offset >> 4 at bit 16 | stride byte offset >> 4 at bit 32 | version at bit 46 | layout at bit 61
// K-major tile  (128-byte rows run along K, row = m; TMA SWIZZLE_128B):
//     layout type 2 (SWIZZLE_128B), LBO = 16 B (unused), SBO = 1024 B (next 8 rows of m)
// MN-major tile (128-byte rows run along M/N, row = k; 32-bit elements need the 32-byte-atom swizzle, TMA SWIZZLE_128B_ATOM_32B):
//     layout type 1 (SWIZZLE_128B_BASE32B), LBO = 4096 B (next 32 m = next TMA box), SBO = 512 B (next 4 k: the swizzle atom is 4 rows)
// (the MN-major field values were confirmed by a parameter sweep on the B200, tools/dbg_tf32_mn.py)
template <bool MN_MAJOR> __device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_addr) {
    constexpr uint64_t LBO = MN_MAJOR ? (4096 >> 4) : 1, SBO = MN_MAJOR ? (512 >> 4) : (1024 >> 4), LT = MN_MAJOR ? 1 : 2;
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (LBO << 16) | (SBO << 32) | ((uint64_t)1 << 46) | (LT << 61);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- CTA-pair (cta_group::2) variants: one thread of the LEADER CTA issues M = 256 MMAs over both CTAs' shared and tensor memory
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// completion of all prior MMAs arrives on the barrier at the same shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
}
// arrive on the barrier at this offset in the LEADER CTA (rank 0 of the cluster). Default (CTA-scope release) semantics, as CUTLASS's
// ClusterBarrier::arrive(cta_id): a .release.cluster here costs a MEMBAR.ALL.GPU per arriving thread (17 % of the samples, ncu)
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, 0;\n"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
        "}\n" ::"r"(bar)
        : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAITC_LOOP:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAITC_DONE;\n"
        "bra WAITC_LOOP;\n"
        "WAITC_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// PAIR: two CTAs of a cluster share one 256 x 128 tile: each loads and splits its own 128 rows of A and its own 64-row HALF of B, the
// leader's tcgen05.mma.cta_group::2 (M = 256) reads both halves of B, and each CTA ends up with its own 128 x 128 accumulators in its own
// tensor memory (the epilogue is the single-CTA one). Per k-step a CTA then moves 24 KB (TMA) + 48 KB (splitters) + 72 KB (operand
// reads) instead of 32 + 64 + 96 KB through shared memory, which is what bounds the single-CTA kernel; 4 stages of 48 KB.
template <bool A_MN, bool B_MN, bool MULTI, bool GENERIC = false, bool PAIR = false>
__global__ void __launch_bounds__(320, 1) gemm_tf32x3_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                             const __grid_constant__ Tf32GemmArgs args) {
    constexpr int      BM = 128, BN = 128, BK = 32, STAGES = PAIR ? 4 : 3, NACC = 2;
    constexpr uint32_t TILE_BYTES = BM * BK * 4;                 // 16 KB per A tile
    constexpr uint32_t TILE_B_BYTES = PAIR ? TILE_BYTES / 2 : TILE_BYTES; // PAIR: this CTA's 64-row half of the B tile
    constexpr uint32_t RAW_BYTES = TILE_BYTES + TILE_B_BYTES;    // A_raw | B_raw
    constexpr uint32_t STAGE_BYTES = 2 * RAW_BYTES;              // + A_lo | B_lo
    constexpr uint32_t TMEM_COLS = 512;                          // two accumulator PAIRS (hi*hi | cross terms, 128 columns each): chunks and tiles alternate
    constexpr uint32_t PATCH_BYTES = 32 * 33 * 4;                // per epilogue warp: 32 x 32 transposition patch
    constexpr int      KCB = 16;                                 // k-blocks per accumulation chunk (512 k)
    // instruction descriptor: D=F32 (1<<4), A=TF32 (2<<7), B=TF32 (2<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) |
                               ((uint32_t)((PAIR ? 2 * BM : BM) >> 4) << 24);

    extern __shared__ unsigned char smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const uint32_t patches = smem_base + STAGES * STAGE_BYTES;
    const uint32_t bars = patches + 4 * PATCH_BYTES;
    auto full_raw = [&](int s) { return bars + 8u * s; };
    auto full_lo = [&](int s) { return bars + 8u * (STAGES + s); };
    auto empty = [&](int s) { return bars + 8u * (2 * STAGES + s); };
    auto chunk_full = [&](int b) { return bars + 8u * (3 * STAGES + b); };
    auto chunk_drained = [&](int b) { return bars + 8u * (3 * STAGES + NACC + b); };
    const uint32_t tmem_slot = bars + 8u * (3 * STAGES + 2 * NACC);    // 4 bytes: TMEM base address written by tcgen05.alloc
    unsigned char *smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nsteps = GENERIC ? args.KT : (int)((args.K + BK - 1) / BK);
    const int nchunks = (nsteps + KCB - 1) / KCB;
    // static tile schedule: this CTA owns tiles blockIdx.x, blockIdx.x + gridDim.x, ... ; tile_m runs fastest, so the CTAs that are
    // resident together share B panels (and, a little later, A panels) in L2
    // PAIR: the schedule runs over CTA pairs and 256-row tile pairs (args.tiles_m counts those); rank r of the pair owns rows 128 r .. 128 r + 127
    uint32_t rank = 0;
    if constexpr (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int64_t sched_id = PAIR ? (blockIdx.x >> 1) : blockIdx.x, sched_n = PAIR ? (gridDim.x >> 1) : gridDim.x;
    const int64_t tiles_per_batch = (int64_t)args.tiles_m * args.tiles_n, total_tiles = tiles_per_batch * args.batch;
    const int     my_tiles = (int)((total_tiles - sched_id + sched_n - 1) / sched_n);
    auto decode = [&](int i, int &tile_m, int &tile_n, int &b) {
        const int64_t t = sched_id + (int64_t)i * sched_n;
        b = (int)(t / tiles_per_batch);
        const int in = (int)(t - (int64_t)b * tiles_per_batch);
        tile_n = in / args.tiles_m;
        tile_m = in - tile_n * args.tiles_m;
        if constexpr (PAIR) tile_m = 2 * tile_m + (int)rank;
    };

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(full_raw(s), 1);
            mbar_init(full_lo(s), PAIR ? 8 : 128); // PAIR: the leader's barrier collects one arrival per splitter warp of both CTAs
            mbar_init(empty(s), 1);
        }
        for (int b = 0; b < NACC; b++) {
            mbar_init(chunk_full(b), 1);
            mbar_init(chunk_drained(b), PAIR ? 256 : 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if constexpr (PAIR) { // one warp of EACH CTA of the pair takes part
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if constexpr (PAIR) cluster_sync_all(); // the peer's barriers must be initialised before anything arrives on them
    else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ===================== TMA producer =====================
        if constexpr (GENERIC) {
            // lanes 0..3 own the four 32-row M blocks of the tile, lanes 4..7 the four N blocks: per tile each lane decodes its block into
            // tensor-map coordinates, per k-step it adds the K coordinates and issues its box (one SIMT pass per step); lane 0 alone polls
            // the empty barrier and posts the transaction bytes
            // PAIR: lanes 4..5 own this CTA's TWO N blocks (its half of the tile's columns; complex mode: rank 0 = the real-output plane,
            // rank 1 = the imaginary-output plane of the same 64 complex columns)
            constexpr int NLANES = PAIR ? 6 : 8;
            const bool    is_a = lane < 4;
            const int     blk = lane & 3;
            int        gs = 0;
            for (int i = 0; i < my_tiles; i++) {
                int tile_m, tile_n, b;
                decode(i, tile_m, tile_n, b);
                int c0[5] = {0, 0, 0, 0, 0};
                if (lane < NLANES) {
                    const int64_t  cb = is_a ? (int64_t)tile_m * 4 + blk
                                            : (args.cplx ? (int64_t)tile_n * 2 + (blk & 1) : (int64_t)tile_n * 4 + (PAIR ? 2 * (int)rank : 0) + blk);
                    const int32_t *ext = is_a ? args.em : args.en, *dt = is_a ? args.dAm : args.dBn, *db = is_a ? args.dAb : args.dBb;
                    const int      blk0n = is_a ? args.mblk0 : args.nblk0;
                    int            blk0 = blk0n, i1 = 0, i2 = 0; // out of range: an inner coordinate past the extent zero-fills the whole box
                    if (cb < (is_a ? args.MB : args.NB)) {
                        const int64_t q0 = cb / blk0n;
                        blk0 = (int)(cb - q0 * blk0n);
                        const int64_t q1 = q0 / ext[1];
                        i1 = (int)(q0 - q1 * ext[1]);
                        i2 = (int)q1;
                    }
#pragma unroll
                    for (int d = 0; d < 5; d++) c0[d] = (dt[0] == d ? 32 * blk0 : 0) + (dt[1] == d ? i1 : 0) + (dt[2] == d ? i2 : 0);
                    if (args.b_flat) {
#pragma unroll
                        for (int d = 0; d < 5; d++) c0[d] += (db[0] == d ? b : 0);
                    } else {
                        int x = b;
#pragma unroll
                        for (int q = 0; q < 3; q++) {
                            const int bi = x % args.eb[q];
                            x /= args.eb[q];
#pragma unroll
                            for (int d = 0; d < 5; d++) c0[d] += (db[q] == d ? bi : 0);
                        }
                    }
                    if (args.cplx && !is_a) {
#pragma unroll
                        for (int d = 0; d < 5; d++) c0[d] += (args.dBc == d ? (PAIR ? (int)rank : (blk >> 1)) : 0);
                    }
                }
                const int kd0 = is_a ? args.dAk[0] : args.dBk[0], kd1 = is_a ? args.dAk[1] : args.dBk[1], kd2 = is_a ? args.dAk[2] : args.dBk[2];
                int       s0 = 0, k1 = 0, k2 = 0;
                for (int it = 0; it < nsteps; it++, gs++) {
                    const int s = gs % STAGES;
                    if (lane == 0) {
                        if (gs >= STAGES) mbar_wait(empty(s), ((gs / STAGES) - 1) & 1);
                        mbar_expect_tx(full_raw(s), RAW_BYTES);
                    }
                    __syncwarp();
                    if (lane < NLANES) {
                        int c[5];
#pragma unroll
                        for (int d = 0; d < 5; d++) c[d] = c0[d] + (kd0 == d ? 32 * s0 : 0) + (kd1 == d ? k1 : 0) + (kd2 == d ? k2 : 0);
                        tma_load_5d(smem_base + s * STAGE_BYTES + (is_a ? 0u : TILE_BYTES) + (uint32_t)blk * 4096u, is_a ? &tmA : &tmB, full_raw(s), c[0], c[1],
                                    c[2], c[3], c[4]);
                    }
                    if (++s0 == args.kstep0) {
                        s0 = 0;
                        if (++k1 == args.ek[1]) { k1 = 0; ++k2; }
                    }
                }
            }
        } else if (lane == 0) {
            int gs = 0; // k-steps issued by this CTA since the start (all tiles): stage = gs % STAGES
            for (int i = 0; i < my_tiles; i++) {
                int tile_m, tile_n, b;
                decode(i, tile_m, tile_n, b);
                for (int it = 0; it < nsteps; it++, gs++) {
                    const int s = gs % STAGES;
                    if (gs >= STAGES) mbar_wait(empty(s), ((gs / STAGES) - 1) & 1);
                    const uint32_t a_dst = smem_base + s * STAGE_BYTES, b_dst = a_dst + TILE_BYTES;
                    mbar_expect_tx(full_raw(s), RAW_BYTES);
                    if constexpr (A_MN) { // box {32 m, 32 k}: four boxes cover the 128 rows of the tile
#pragma unroll
                        for (int q = 0; q < 4; q++) tma_load_3d(a_dst + q * 4096, &tmA, full_raw(s), tile_m * BM + 32 * q, it * BK, b);
                    } else tma_load_3d(a_dst, &tmA, full_raw(s), it * BK, tile_m * BM, b); // box {32 k, 128 m}
                    const int n0 = tile_n * BN + (PAIR ? (int)rank * (BN / 2) : 0); // PAIR: this CTA's half of the tile's columns
                    if constexpr (B_MN) {
#pragma unroll
                        for (int q = 0; q < (PAIR ? 2 : 4); q++) tma_load_3d(b_dst + q * 4096, &tmB, full_raw(s), n0 + 32 * q, it * BK, b);
                    } else tma_load_3d(b_dst, &tmB, full_raw(s), it * BK, n0, b); // box {32 k, 128 n} (PAIR: {32 k, 64 n})
                }
            }
        }
    } else if (warp == 1) {
        // ===================== tcgen05.mma issuer (PAIR: the leader CTA's only) =====================
        if (lane == 0 && (!PAIR || rank == 0)) {
            int gs = 0, gc = 0; // global k-step and accumulation-chunk counters of this CTA
            for (int i = 0; i < my_tiles; i++) {
                for (int it = 0; it < nsteps; it++, gs++) {
                    const int s = gs % STAGES, ab = gc & (NACC - 1);
                    if (it % KCB == 0 && gc >= NACC) { // the accumulator must have been drained (chunk gc - NACC)
                        if constexpr (PAIR) mbar_wait_cluster(chunk_drained(ab), ((gc / NACC) - 1) & 1);
                        else mbar_wait(chunk_drained(ab), ((gc / NACC) - 1) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    if constexpr (PAIR) mbar_wait_cluster(full_lo(s), (gs / STAGES) & 1);
                    else mbar_wait(full_lo(s), (gs / STAGES) & 1); // splitters are done => the raw tile landed too
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = smem_base + s * STAGE_BYTES, b_hi = a_hi + TILE_BYTES, a_lo = a_hi + RAW_BYTES, b_lo = a_lo + TILE_BYTES;
                    // The hi*hi product and the two cross terms go to DIFFERENT accumulators: every accumulation step truncates the FP32
                    // accumulator (a one-sided error of up to 1 ulp of the running sum), so the large sum should absorb as few steps as
                    // possible; the cross terms are 2^-10 of it and their own truncation does not matter.
                    const uint32_t tacc_hi = tmem_base + (uint32_t)ab * 256u, tacc_x = tacc_hi + 128u;
#pragma unroll
                    for (int kk = 0; kk < BK / 8; kk++) {
                        // one MMA consumes 8 k: 32 bytes along the swizzle row of a K-major tile, one 8-row (1024-byte) atom of an MN-major tile
                        const uint32_t koa = kk * (A_MN ? 1024 : 32), kob = kk * (B_MN ? 1024 : 32);
                        const uint64_t dah = umma_smem_desc_sw128<A_MN>(a_hi + koa), dbh = umma_smem_desc_sw128<B_MN>(b_hi + kob);
                        const uint64_t dal = umma_smem_desc_sw128<A_MN>(a_lo + koa), dbl = umma_smem_desc_sw128<B_MN>(b_lo + kob);
                        const uint32_t first = ((it % KCB) | kk) == 0;
                        if constexpr (PAIR) {
                            umma_tf32_pair(tacc_x, dal, dbh, IDESC, !first);
                            umma_tf32_pair(tacc_x, dah, dbl, IDESC, 1);
                            umma_tf32_pair(tacc_hi, dah, dbh, IDESC, !first);
                        } else {
                            umma_tf32(tacc_x, dal, dbh, IDESC, !first);
                            umma_tf32(tacc_x, dah, dbl, IDESC, 1);
                            umma_tf32(tacc_hi, dah, dbh, IDESC, !first);
                        }
                    }
                    if constexpr (PAIR) umma_commit_pair(empty(s)); // frees the stage in both CTAs
                    else umma_commit(empty(s));
                    if (it % KCB == KCB - 1 || it == nsteps - 1) {
                        if constexpr (PAIR) umma_commit_pair(chunk_full(ab));
                        else umma_commit(chunk_full(ab));
                        gc++;
                    }
                }
            }
        }
    } else if (warp < 6) {
        // ===================== splitters (warps 2..5, 128 threads) =====================
        // The tensor core reads only the upper 19 bits of an FP32 container, so the raw tile IS the hi part (x truncated to TF32);
        // lo = x - trunc(x) is exact in FP32 and goes to the second buffer in the same swizzled layout.
        const int t = tid - 64;
        const int total_steps = my_tiles * nsteps;
        for (int gs = 0; gs < total_steps; gs++) {
            const int s = gs % STAGES;
            mbar_wait(full_raw(s), (gs / STAGES) & 1);
            const float4 *raw = reinterpret_cast<const float4 *>(smem_gen + s * STAGE_BYTES);
            float4       *lo = reinterpret_cast<float4 *>(smem_gen + s * STAGE_BYTES + RAW_BYTES);
#pragma unroll 8
            for (int i = 0; i < (int)(RAW_BYTES / 16 / 128); i++) {
                const int    idx = i * 128 + t;
                const float4 x = raw[idx];
                float4       l;
                auto tr = [](float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); };
                l.x = x.x - tr(x.x);
                l.y = x.y - tr(x.y);
                l.z = x.z - tr(x.z);
                l.w = x.w - tr(x.w);
                lo[idx] = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy stores -> visible to the tensor core's async proxy
            if constexpr (PAIR) {
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(full_lo(s));
            } else mbar_arrive(full_lo(s));
        }
    } else {
        // ===================== epilogue warps (6..9): drain finished chunks, store finished tiles =====================
        // They run one tile behind the issuer: while tile i is converted and stored from here, the tensor core already accumulates
        // tile i + 1 into the other accumulators.
        const int quad = warp & 3;                       // TMEM lane quadrant of this warp: warps 6,7,8,9 -> 2,3,0,1
        float    *patch = reinterpret_cast<float *>(smem_gen + STAGES * STAGE_BYTES) + (warp - 6) * (32 * 33);
        const float alpha = args.alpha, beta = args.beta;
        auto arrive_drained = [&](int ab) {
            if constexpr (PAIR) mbar_arrive_leader(chunk_drained(ab)); // the leader's issuer waits for both CTAs' epilogue warps
            else mbar_arrive(chunk_drained(ab));
        };
        auto tmem_ld32 = [&](uint32_t taddr, uint32_t(&r)[32]) {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                  "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                  "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
                  "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        };
        auto tmem_ld16 = [&](uint32_t taddr, uint32_t(&r)[16]) { // no wait: the caller waits once for several loads
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                           "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(taddr));
        };
        auto tmem_ld8 = [&](uint32_t taddr, uint32_t(&r)[8]) { // no wait: the caller waits once for several loads
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                         : "r"(taddr));
        };
        // store 32 columns [n0, n0 + 32) of this thread's row: out = alpha * v + beta * C
        auto store_block = [&](float *Cb, int tile_m, int64_t n0, const float(&v)[32]) {
            const int64_t m = (int64_t)tile_m * BM + quad * 32 + lane;
            if (args.sCn == 1) {
                // C rows are contiguous along n: transpose the 32x32 block through a warp-private shared-memory patch so that every
                // store/load instruction covers one 128-byte row segment
#pragma unroll
                for (int j = 0; j < 32; j++) patch[lane * 33 + j] = v[j];
                __syncwarp();
                const int64_t n = n0 + lane, mrow0 = (int64_t)tile_m * BM + quad * 32;
                if (n < args.N) {
#pragma unroll
                    for (int r0 = 0; r0 < 32; r0 += 16) { // the 16 old C values of a half block are requested before their first use
                        float old[16];
                        if (beta != 0.f) {
#pragma unroll
                            for (int rr = 0; rr < 16; rr++) old[rr] = (mrow0 + r0 + rr < args.M) ? Cb[(mrow0 + r0 + rr) * args.sCm + n] : 0.f;
                        }
#pragma unroll
                        for (int rr = 0; rr < 16; rr++) {
                            const int64_t mm = mrow0 + r0 + rr;
                            if (mm < args.M) {
                                float o = alpha * patch[(r0 + rr) * 33 + lane];
                                if (beta != 0.f) o = fmaf(beta, old[rr], o);
                                Cb[mm * args.sCm + n] = o;
                            }
                        }
                    }
                }
                __syncwarp();
            } else if (m < args.M) {
                // lanes are consecutive rows m: coalesced when C is contiguous along m (the "transposed C" case, e.g. Wabef)
                float *crow = Cb + m * args.sCm;
#pragma unroll
                for (int j0 = 0; j0 < 32; j0 += 16) {
                    float old[16];
                    if (beta != 0.f) {
#pragma unroll
                        for (int j = 0; j < 16; j++) old[j] = (n0 + j0 + j < args.N) ? crow[(n0 + j0 + j) * args.sCn] : 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const int64_t n = n0 + j0 + j;
                        if (n < args.N) {
                            float o = alpha * v[j0 + j];
                            if (beta != 0.f) o = fmaf(beta, old[j], o);
                            crow[n * args.sCn] = o;
                        }
                    }
                }
            }
        };
        int gc = 0;
        for (int i = 0; i < my_tiles; i++) {
            int tile_m, tile_n, b;
            decode(i, tile_m, tile_n, b);
            float *Cb = args.C + (int64_t)b * args.sCb;
            if constexpr (GENERIC) {
                // ---- multi-index C: offset tables for the tile's 128 rows and 128 columns (thread e fills entry e of both), accumulate the
                // chunks in registers as in the MULTI path, then store through the tables
                constexpr int64_t INVALID = INT64_MIN;
                int64_t          *moffC = reinterpret_cast<int64_t *>(smem_gen + STAGES * STAGE_BYTES + 4 * PATCH_BYTES + 256);
                int64_t          *noffC = moffC + BM;
                const int         e = tid - 192; // 0 .. 127
                {
                    int64_t baseC = 0;
                    int     x = b;
#pragma unroll
                    for (int q = 0; q < 3; q++) {
                        baseC += (int64_t)(x % args.eb[q]) * args.gCb[q];
                        x /= args.eb[q];
                    }
                    Cb = args.C + baseC;
                    auto offset_of = [&](int64_t p, const int32_t *ext, int blk0n, int64_t nblocks, const int64_t *st) -> int64_t {
                        const int64_t cb = p >> 5;
                        if (cb >= nblocks) return INVALID;
                        const int64_t q0 = cb / blk0n, q1 = q0 / ext[1];
                        const int64_t x0 = (cb - q0 * blk0n) * 32 + (p & 31);
                        if (x0 >= ext[0]) return INVALID;
                        return x0 * st[0] + (q0 - q1 * ext[1]) * st[1] + q1 * st[2];
                    };
                    asm volatile("bar.sync 1, 128;" ::: "memory"); // the previous tile's stores have read the tables
                    moffC[e] = offset_of((int64_t)tile_m * BM + e, args.em, args.mblk0, args.MB, args.gCm);
                    if (args.cplx) { // column e: complex column e & 63 of the tile, real (e < 64) or imaginary part
                        const int64_t o = offset_of((int64_t)tile_n * (BN / 2) + (e & 63), args.en, args.nblk0, args.NB, args.gCn);
                        noffC[e] = (o == INVALID) ? INVALID : o + (e >> 6);
                    } else noffC[e] = offset_of((int64_t)tile_n * BN + e, args.en, args.nblk0, args.NB, args.gCn);
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
                float acc[BN];
                for (int ch = 0; ch < nchunks; ch++, gc++) {
                    const int ab = gc & (NACC - 1);
                    mbar_wait(chunk_full(ab), (gc / NACC) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                    for (int c0 = 0; c0 < BN; c0 += 32) {
                        uint32_t       r[32];
                        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ab * 256 + c0);
                        tmem_ld32(taddr, r);
                        if (ch == 0) {
#pragma unroll
                            for (int j = 0; j < 32; j++) acc[c0 + j] = __uint_as_float(r[j]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; j++) acc[c0 + j] += __uint_as_float(r[j]);
                        }
                        tmem_ld32(taddr + 128u, r);
#pragma unroll
                        for (int j = 0; j < 32; j++) acc[c0 + j] += __uint_as_float(r[j]);
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    arrive_drained(ab);
                }
#pragma unroll
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    if (args.c_n_unit) {
                        // lanes along n (contiguous in C): transpose the 32 x 32 block through the warp's patch
#pragma unroll
                        for (int j = 0; j < 32; j++) patch[lane * 33 + j] = acc[c0 + j];
                        __syncwarp();
                        const int64_t on = noffC[c0 + lane];
                        if (on != INVALID) {
#pragma unroll
                            for (int r0 = 0; r0 < 32; r0 += 16) { // the 16 old C values of a half block are requested before their first use
                                int64_t om[16];
                                float   old[16];
#pragma unroll
                                for (int rr = 0; rr < 16; rr++) om[rr] = moffC[quad * 32 + r0 + rr];
                                if (beta != 0.f) {
#pragma unroll
                                    for (int rr = 0; rr < 16; rr++) old[rr] = (om[rr] != INVALID) ? Cb[om[rr] + on] : 0.f;
                                }
#pragma unroll
                                for (int rr = 0; rr < 16; rr++) {
                                    if (om[rr] != INVALID) {
                                        float o = alpha * patch[(r0 + rr) * 33 + lane];
                                        if (beta != 0.f) o = fmaf(beta, old[rr], o);
                                        Cb[om[rr] + on] = o;
                                    }
                                }
                            }
                        }
                        __syncwarp();
                    } else {
                        // lanes along m
                        const int64_t om = moffC[quad * 32 + lane];
                        if (args.cplx) {
                            // (re, im) = accumulator columns (j, j + 64): one 8-byte access per complex element
                            if (c0 < BN / 2 && om != INVALID) {
#pragma unroll
                                for (int j0 = 0; j0 < 32; j0 += 16) {
                                    int64_t on[16];
                                    float2  old[16];
#pragma unroll
                                    for (int j = 0; j < 16; j++) on[j] = noffC[c0 + j0 + j];
                                    if (beta != 0.f) {
#pragma unroll
                                        for (int j = 0; j < 16; j++)
                                            old[j] = (on[j] != INVALID) ? *reinterpret_cast<const float2 *>(Cb + om + on[j]) : make_float2(0.f, 0.f);
                                    }
#pragma unroll
                                    for (int j = 0; j < 16; j++) {
                                        if (on[j] != INVALID) {
                                            float2 o = make_float2(alpha * acc[c0 + j0 + j], alpha * acc[(c0 + j0 + j + BN / 2) & (BN - 1)]);
                                            if (beta != 0.f) {
                                                o.x = fmaf(beta, old[j].x, o.x);
                                                o.y = fmaf(beta, old[j].y, o.y);
                                            }
                                            *reinterpret_cast<float2 *>(Cb + om + on[j]) = o;
                                        }
                                    }
                                }
                            }
                        } else if (om != INVALID) {
#pragma unroll
                            for (int j0 = 0; j0 < 32; j0 += 16) { // fully unrolled: acc[] must stay in registers
                                int64_t on[16];
                                float   old[16];
#pragma unroll
                                for (int j = 0; j < 16; j++) on[j] = noffC[c0 + j0 + j];
                                if (beta != 0.f) {
#pragma unroll
                                    for (int j = 0; j < 16; j++) old[j] = (on[j] != INVALID) ? Cb[om + on[j]] : 0.f;
                                }
#pragma unroll
                                for (int j = 0; j < 16; j++) {
                                    if (on[j] != INVALID) {
                                        float o = alpha * acc[c0 + j0 + j];
                                        if (beta != 0.f) o = fmaf(beta, old[j], o);
                                        Cb[om + on[j]] = o;
                                    }
                                }
                            }
                        }
                    }
                }
            } else if constexpr (!MULTI) {
                // K <= 512: one chunk per tile. Each 32-column block goes TMEM -> registers -> global; the accumulator pair is handed
                // back after the last block has been read.
                const int ab = gc & (NACC - 1);
                if (args.sCn == 1) {
                    // C rows contiguous along n (the usual case). With beta != 0 the old C tile of the NEXT tile is pulled into L2 now
                    // (prefetch.global.L2: no registers, no shared memory), and the 32 old values a lane needs for a block are requested
                    // before the block's accumulator columns are read, so their (L2-hit) latency overlaps the TMEM load and the transposition.
                    const int64_t mrow0 = (int64_t)tile_m * BM + quad * 32, nbase = (int64_t)tile_n * BN + lane;
                    if (beta != 0.f) {
                        auto prefetch_tile = [&](int ti) {
                            int pm, pn, pb;
                            decode(ti, pm, pn, pb);
                            const int64_t row = (int64_t)pm * BM + quad * 32 + lane, col = (int64_t)pn * BN;
                            if (row < args.M) {
                                const float *pr = args.C + (int64_t)pb * args.sCb + row * args.sCm;
#pragma unroll
                                for (int c = 0; c < BN; c += 32)
                                    if (col + c < args.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(pr + col + c));
                            }
                        };
                        if (i == 0) prefetch_tile(0);
                        if (i + 1 < my_tiles) prefetch_tile(i + 1);
                    }
                    if (args.c_vec4) {
                        // 16-byte accesses: lane = (row r0 = lane / 8, column group cg = lane % 8) handles rows r0 + 4 i (i = 0..7) x columns 4 cg .. 4 cg + 3
                        // of a 32 x 32 block — a quarter of the load / store instructions and addresses of the scalar form
                        const int     r0 = lane >> 3, cg = lane & 7;
                        const int64_t ncol = (int64_t)tile_n * BN + 4 * cg;
                        float4        oldA[8]; // old C values of the block (a second buffer for the next block spills at the 168-register cap: 130 vs 148 TFLOP/s)
#define EB_LOAD_OLD(Q, O)                                                                                                                         \
    {                                                                                                                                             \
        const int64_t nq_ = ncol + 32 * (Q);                                                                                                      \
        _Pragma("unroll") for (int i2 = 0; i2 < 8; i2++) {                                                                                        \
            const int64_t mm_ = mrow0 + r0 + 4 * i2;                                                                                              \
            O[i2] = (mm_ < args.M && nq_ < args.N) ? *reinterpret_cast<const float4 *>(Cb + mm_ * args.sCm + nq_) : make_float4(0.f, 0.f, 0.f, 0.f); \
        }                                                                                                                                         \
    }
#define EB_BLOCK(Q, CUR, NXT, HAS_NEXT)                                                                                                           \
    {                                                                                                                                             \
        if (beta != 0.f) EB_LOAD_OLD((Q), CUR)                                                                                                    \
        _Pragma("unroll") for (int h = 0; h < 4; h++) {                                                                                           \
            uint32_t       r[8], x[8];                                                                                                            \
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ab * 256 + 32 * (Q) + 8 * h);                           \
            tmem_ld8(taddr, r);                                                                                                                   \
            tmem_ld8(taddr + 128u, x);                                                                                                            \
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");                                                                          \
            _Pragma("unroll") for (int j = 0; j < 8; j++) patch[lane * 33 + 8 * h + j] = __uint_as_float(r[j]) + __uint_as_float(x[j]);           \
        }                                                                                                                                         \
        if ((Q) == BN / 32 - 1) {                                                                                                                 \
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");                                                                      \
            arrive_drained(ab);                                                                                                                   \
        }                                                                                                                                         \
        __syncwarp();                                                                                                                             \
        const int64_t n = ncol + 32 * (Q);                                                                                                        \
        if (n < args.N) {                                                                                                                         \
            _Pragma("unroll") for (int i2 = 0; i2 < 8; i2++) {                                                                                    \
                const int     rr = r0 + 4 * i2;                                                                                                   \
                const int64_t mm = mrow0 + rr;                                                                                                    \
                if (mm < args.M) {                                                                                                                \
                    const float *pp = patch + rr * 33 + 4 * cg;                                                                                   \
                    float4       o = make_float4(alpha * pp[0], alpha * pp[1], alpha * pp[2], alpha * pp[3]);                                     \
                    if (beta != 0.f) {                                                                                                            \
                        const float4 c = CUR[i2];                                                                                                 \
                        o.x = fmaf(beta, c.x, o.x); o.y = fmaf(beta, c.y, o.y); o.z = fmaf(beta, c.z, o.z); o.w = fmaf(beta, c.w, o.w);           \
                    }                                                                                                                             \
                    *reinterpret_cast<float4 *>(Cb + mm * args.sCm + n) = o;                                                                      \
                }                                                                                                                                 \
            }                                                                                                                                     \
        }                                                                                                                                         \
        __syncwarp();                                                                                                                             \
    }
                        mbar_wait(chunk_full(ab), (gc / NACC) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        EB_BLOCK(0, oldA, oldA, true)
                        EB_BLOCK(1, oldA, oldA, true)
                        EB_BLOCK(2, oldA, oldA, true)
                        EB_BLOCK(3, oldA, oldA, false)
#undef EB_BLOCK
#undef EB_LOAD_OLD
                    } else {
                    mbar_wait(chunk_full(ab), (gc / NACC) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
                    for (int q = 0; q < BN / 32; q++) {
                        const int64_t n = nbase + 32 * q;
                        float         old[32];
                        if (beta != 0.f && n < args.N) {
#pragma unroll
                            for (int rr = 0; rr < 32; rr++) old[rr] = (mrow0 + rr < args.M) ? Cb[(mrow0 + rr) * args.sCm + n] : 0.f;
                        }
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            uint32_t       r[16], x[16];
                            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ab * 256 + 32 * q + 16 * h);
                            tmem_ld16(taddr, r);
                            tmem_ld16(taddr + 128u, x);
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                            for (int j = 0; j < 16; j++) patch[lane * 33 + 16 * h + j] = __uint_as_float(r[j]) + __uint_as_float(x[j]);
                        }
                        if (q == BN / 32 - 1) {
                            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                            arrive_drained(ab);
                        }
                        __syncwarp();
                        if (n < args.N) {
#pragma unroll
                            for (int rr = 0; rr < 32; rr++) {
                                const int64_t mm = mrow0 + rr;
                                if (mm < args.M) {
                                    float o = alpha * patch[rr * 33 + lane];
                                    if (beta != 0.f) o = fmaf(beta, old[rr], o);
                                    Cb[mm * args.sCm + n] = o;
                                }
                            }
                        }
                        __syncwarp();
                    }
                    }
                } else {
                    mbar_wait(chunk_full(ab), (gc / NACC) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
                    for (int c0 = 0; c0 < BN; c0 += 32) {
                        uint32_t       r[32], x[32];
                        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ab * 256 + c0);
                        tmem_ld32(taddr, r);
                        tmem_ld32(taddr + 128u, x);
                        if (c0 == BN - 32) {
                            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                            arrive_drained(ab);
                        }
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]) + __uint_as_float(x[j]);
                        store_block(Cb, tile_m, (int64_t)tile_n * BN + c0, v);
                    }
                }
                gc++;
            } else {
                // K > 512: finished chunks are added into registers with round-to-nearest (the error of an accumulator grows with the
                // number of k-steps it has absorbed; chunk biases of mixed-sign data are independent, so the total stays at the one-chunk level)
                float acc[BN];
                for (int ch = 0; ch < nchunks; ch++, gc++) {
                    const int ab = gc & (NACC - 1);
                    mbar_wait(chunk_full(ab), (gc / NACC) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                    for (int c0 = 0; c0 < BN; c0 += 32) {
                        uint32_t       r[32];
                        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ab * 256 + c0);
                        tmem_ld32(taddr, r);
                        if (ch == 0) {
#pragma unroll
                            for (int j = 0; j < 32; j++) acc[c0 + j] = __uint_as_float(r[j]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; j++) acc[c0 + j] += __uint_as_float(r[j]);
                        }
                        tmem_ld32(taddr + 128u, r);
#pragma unroll
                        for (int j = 0; j < 32; j++) acc[c0 + j] += __uint_as_float(r[j]);
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    arrive_drained(ab);
                }
#pragma unroll
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] = acc[c0 + j];
                    store_block(Cb, tile_m, (int64_t)tile_n * BN + c0, v);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    if constexpr (PAIR) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        cluster_sync_all(); // neither CTA may leave (or free tensor memory) while the other still arrives on its barriers
    } else __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

#endif // EB200_TF32_KERNEL_IMPL

} // namespace eb200
