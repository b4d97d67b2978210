// gemm_tf32x3.cuh — FP32 GEMM on the 5th-generation tensor cores (tcgen05 / TMEM) with 3xTF32 split accumulation.
//
// FP32 results must match the CPU path to 1e-5; a single TF32 product (10-bit mantissa) cannot, so every operand is split on
// the fly into hi = trunc_tf32(x) and lo = x - hi (exact in FP32) and the tile product is formed as
//        A_hi*B_hi + A_hi*B_lo + A_lo*B_hi          (the lo*lo term is below 2^-20 relative and dropped)
// with all three tcgen05.mma accumulating into the same FP32 accumulator in tensor memory.
//
// Roles (192 threads, warp-specialised, one 128x128 output tile per CTA, BK = 32 floats = one 128-byte swizzle row):
//   warp 0, one lane : TMA producer — cp.async.bulk.tensor.2d of the raw A and B tiles (SWIZZLE_128B, K-major)
//   warps 2..5       : splitters — read the raw tile, write hi back in place and lo to a second buffer (same swizzled layout,
//                      so they just stream 16-byte chunks), fence.proxy.async, arrive; later they run the epilogue
//   warp 1, one lane : tcgen05.mma issuer — 4 k-slices x 3 products per stage, kind::tf32, M=128 N=128 K=8, cta_group::1;
//                      tcgen05.commit releases the stage to the producer and finally signals the epilogue
//   epilogue         : tcgen05.ld 32x32b (each warp its own TMEM lane quadrant) -> alpha*acc + beta*C -> global
// Two 128-column TMEM accumulators alternate between chunks of 512 k; finished chunks are added into registers with
// round-to-nearest (the tensor core's own FP32 accumulation truncates).
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

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(192, 1) gemm_tf32x3_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                             const __grid_constant__ Tf32GemmArgs args) {
    constexpr int      BM = 128, BN = 128, BK = 32, STAGES = 3;
    constexpr uint32_t TILE_BYTES = BM * BK * 4;                 // 16 KB per operand tile
    constexpr uint32_t RAW_BYTES = 2 * TILE_BYTES;               // A_raw | B_raw
    constexpr uint32_t STAGE_BYTES = 2 * RAW_BYTES;              // + A_lo | B_lo
    constexpr uint32_t TMEM_COLS = 256;                          // two 128-column accumulators (ping-pong between K chunks)
    constexpr int      KCB = 16;                                 // k-blocks per accumulation chunk (512 k)
    // instruction descriptor: D=F32 (1<<4), A=TF32 (2<<7), B=TF32 (2<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) |
                               ((uint32_t)(BM >> 4) << 24);

    extern __shared__ unsigned char smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const uint32_t bars = smem_base + STAGES * STAGE_BYTES;
    auto full_raw = [&](int s) { return bars + 8u * s; };
    auto full_lo = [&](int s) { return bars + 8u * (STAGES + s); };
    auto empty = [&](int s) { return bars + 8u * (2 * STAGES + s); };
    auto chunk_full = [&](int b) { return bars + 8u * (3 * STAGES + b); };
    auto chunk_drained = [&](int b) { return bars + 8u * (3 * STAGES + 2 + b); };
    const uint32_t tmem_slot = bars + 8u * (3 * STAGES + 4);    // 4 bytes: TMEM base address written by tcgen05.alloc
    unsigned char *smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_m = blockIdx.x % args.tiles_m, tile_n = blockIdx.x / args.tiles_m;
    const int nsteps = (int)((args.K + BK - 1) / BK);

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(full_raw(s), 1);
            mbar_init(full_lo(s), 128);
            mbar_init(empty(s), 1);
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(chunk_full(b), 1);
            mbar_init(chunk_drained(b), 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < nsteps; it++) {
                const int s = it % STAGES;
                if (it >= STAGES) mbar_wait(empty(s), ((it / STAGES) - 1) & 1);
                const uint32_t a_dst = smem_base + s * STAGE_BYTES, b_dst = a_dst + TILE_BYTES;
                mbar_expect_tx(full_raw(s), RAW_BYTES);
                if constexpr (A_MN) { // box {32 m, 32 k}: four boxes cover the 128 rows of the tile
#pragma unroll
                    for (int b = 0; b < 4; b++) tma_load_2d(a_dst + b * 4096, &tmA, full_raw(s), tile_m * BM + 32 * b, it * BK);
                } else tma_load_2d(a_dst, &tmA, full_raw(s), it * BK, tile_m * BM); // box {32 k, 128 m}
                if constexpr (B_MN) {
#pragma unroll
                    for (int b = 0; b < 4; b++) tma_load_2d(b_dst + b * 4096, &tmB, full_raw(s), tile_n * BN + 32 * b, it * BK);
                } else tma_load_2d(b_dst, &tmB, full_raw(s), it * BK, tile_n * BN);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int it = 0; it < nsteps; it++) {
                const int s = it % STAGES;
                const int c = it / KCB, cb = c & 1;               // accumulation chunk and its TMEM accumulator
                if (it % KCB == 0 && c >= 2) {                     // the accumulator must have been drained (chunk c-2)
                    mbar_wait(chunk_drained(cb), ((c >> 1) - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                mbar_wait(full_lo(s), (it / STAGES) & 1);   // splitters are done => the raw tile landed too
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = smem_base + s * STAGE_BYTES, b_hi = a_hi + TILE_BYTES, a_lo = a_hi + RAW_BYTES, b_lo = a_lo + TILE_BYTES;
                const uint32_t tacc = tmem_base + (uint32_t)cb * 128u;
#pragma unroll
                for (int kk = 0; kk < BK / 8; kk++) {
                    // one MMA consumes 8 k: 32 bytes along the swizzle row of a K-major tile, one 8-row (1024-byte) atom of an MN-major tile
                    const uint32_t koa = kk * (A_MN ? 1024 : 32), kob = kk * (B_MN ? 1024 : 32);
                    const uint64_t dah = umma_smem_desc_sw128<A_MN>(a_hi + koa), dbh = umma_smem_desc_sw128<B_MN>(b_hi + kob);
                    const uint64_t dal = umma_smem_desc_sw128<A_MN>(a_lo + koa), dbl = umma_smem_desc_sw128<B_MN>(b_lo + kob);
                    umma_tf32(tacc, dal, dbh, IDESC, ((it % KCB) | kk) != 0);
                    umma_tf32(tacc, dah, dbl, IDESC, 1);
                    umma_tf32(tacc, dah, dbh, IDESC, 1);
                }
                umma_commit(empty(s));
                if (it % KCB == KCB - 1 || it == nsteps - 1) umma_commit(chunk_full(cb));
            }
        }
    } else {
        // ===================== splitters (warps 2..5, 128 threads); they also drain finished K chunks =====================
        const int     t = tid - 64;
        const int     quad = warp & 3;                       // TMEM lane quadrant of this warp: warps 2,3,4,5 -> 2,3,0,1
        const int     nchunks = (nsteps + KCB - 1) / KCB;
        float         acc[BN];                               // this thread's output row (TMEM lane), summed over chunks in FP32 RN
#pragma unroll
        for (int j = 0; j < BN; j++) acc[j] = 0.f;
        int drained = 0;
        // The tensor core accumulates into TMEM with truncation, so the error of one accumulator grows linearly with the number of
        // k-steps it has absorbed. Every KCB k-blocks the issuer switches to the other accumulator and the finished one is added
        // into registers with round-to-nearest here; signs of the chunk biases are independent, so the total stays at the one-chunk level.
        auto drain = [&](int d) {
            const int db = d & 1;
            mbar_wait(chunk_full(db), (d >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t       r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(db * 128 + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                      "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                      "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
                      "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; j++) acc[c0 + j] += __uint_as_float(r[j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(chunk_drained(db));
        };
        for (int it = 0; it < nsteps; it++) {
            const int s = it % STAGES;
            mbar_wait(full_raw(s), (it / STAGES) & 1);
            float4 *raw = reinterpret_cast<float4 *>(smem_gen + s * STAGE_BYTES);
            float4 *lo = reinterpret_cast<float4 *>(smem_gen + s * STAGE_BYTES + RAW_BYTES);
#pragma unroll 4
            for (int i = 0; i < (int)(RAW_BYTES / 16 / 128); i++) {
                const int idx = i * 128 + t;
                float4    x = raw[idx], h, l;
                // hi = x rounded to nearest TF32 (10-bit mantissa), lo = x - hi (exact) rounded to nearest TF32 as well
                auto rn_tf32 = [](float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); };
                h.x = rn_tf32(x.x); l.x = rn_tf32(x.x - h.x);
                h.y = rn_tf32(x.y); l.y = rn_tf32(x.y - h.y);
                h.z = rn_tf32(x.z); l.z = rn_tf32(x.z - h.z);
                h.w = rn_tf32(x.w); l.w = rn_tf32(x.w - h.w);
                raw[idx] = h;
                lo[idx] = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy stores -> visible to the tensor core's async proxy
            mbar_arrive(full_lo(s));
            // a chunk is drained two k-blocks after its last k-block was split: its MMAs have normally completed by then, so neither
            // side waits; the issuer only re-uses an accumulator after this drain (chunk_drained)
            if (drained < nchunks - 1 && it >= (drained + 1) * KCB + 1) drain(drained++);
        }
        while (drained < nchunks) drain(drained++);

        // ===================== epilogue: alpha*acc + beta*C, one output row per thread =====================
        const int64_t m = (int64_t)tile_m * BM + quad * 32 + lane;
        const float   alpha = args.alpha, beta = args.beta;
        float        *patch = reinterpret_cast<float *>(smem_gen) + (warp - 2) * (32 * 33);
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
            const int64_t n0 = (int64_t)tile_n * BN + c0;
            if (args.sCn == 1) {
                // C rows are contiguous along n: transpose the 32x32 chunk through a warp-private shared-memory patch (the pipeline
                // stages are idle by now) so that every store/load instruction covers one 128-byte row segment
#pragma unroll
                for (int j = 0; j < 32; j++) patch[lane * 33 + j] = acc[c0 + j];
                __syncwarp();
                const int64_t n = n0 + lane, mrow0 = (int64_t)tile_m * BM + quad * 32;
                if (n < args.N) {
                    float old[32];
                    if (beta != 0.f) { // every old C value of the chunk is requested before the first use
#pragma unroll
                        for (int rr = 0; rr < 32; rr++) old[rr] = (mrow0 + rr < args.M) ? args.C[(mrow0 + rr) * args.sCm + n] : 0.f;
                    }
#pragma unroll
                    for (int rr = 0; rr < 32; rr++) {
                        const int64_t mm = mrow0 + rr;
                        if (mm < args.M) {
                            float v = alpha * patch[rr * 33 + lane];
                            if (beta != 0.f) v = fmaf(beta, old[rr], v);
                            args.C[mm * args.sCm + n] = v;
                        }
                    }
                }
                __syncwarp();
            } else if (m < args.M) {
                // lanes are consecutive rows m: coalesced when C is contiguous along m (the "transposed C" case, e.g. Wabef)
                float *crow = args.C + m * args.sCm;
                float  old[32];
                if (beta != 0.f) {
#pragma unroll
                    for (int j = 0; j < 32; j++) old[j] = (n0 + j < args.N) ? crow[(n0 + j) * args.sCn] : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const int64_t n = n0 + j;
                    if (n < args.N) {
                        float v = alpha * acc[c0 + j];
                        if (beta != 0.f) v = fmaf(beta, old[j], v);
                        crow[n * args.sCn] = v;
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

#endif // EB200_TF32_KERNEL_IMPL

} // namespace eb200
