// plan.cuh — host-side planner of the contraction engine.
//
// Input: eb200_contract_desc (= what the reference's device launcher builds on the host,
// TensorAlgebra/Backends/GPUTensorAlgebra.hpp:225-239: unique index list, extents, index tables).
// Output: a Plan that classifies every unique index by which tensors carry it
//     batch (C,A,B) | M (C,A) | N (C,B) | K (A,B) | A-only / B-only sums -> K | C-only broadcast -> M
// orders each group inner-first by memory stride, merges indices that are contiguous in every
// tensor that carries them, and picks the kernel family. No tensor is ever physically permuted.
#pragma once
#include "common.cuh"

namespace eb200 {

constexpr int     MAXG        = 8;            // max indices per group after merging
constexpr int64_t OFF_INVALID = INT64_MIN;    // offset-table sentinel for out-of-range rows/cols

struct Grp {
    int32_t n;                // number of indices in the group (0 => tot == 1)
    int32_t pad_;
    int64_t tot;              // product of extents
    int64_t ext[MAXG];        // inner-first
    int64_t s[3][MAXG];       // element strides per tensor: [0]=A, [1]=B, [2]=C (0 when absent)
    // division by ext[i] as multiply+shift, valid for dividends < 2^31 (used when tot < 2^31): q = (n * fmul) >> fshift
    uint32_t fmul[MAXG];
    uint32_t fshift[MAXG];
};

// Granlund-Montgomery round-up constants for n < 2^31: s = ceil(log2 d), M = ceil(2^(31+s) / d) < 2^32 + ...
inline void fastdiv_init(int64_t d, uint32_t &mul, uint32_t &shift) {
    if (d <= 0 || d > 0x7fffffffLL) { mul = 0; shift = 0; return; }
    int s = 0;
    while ((1LL << s) < d) s++;
    shift = 31 + s;
    unsigned __int128 num = (unsigned __int128)1 << shift;
    mul = (uint32_t)((num + (unsigned __int128)d - 1) / (unsigned __int128)d);
}
inline void grp_finalize(Grp &g) {
    for (int i = 0; i < MAXG; i++) {
        g.fmul[i] = 0;
        g.fshift[i] = 0;
        if (i < g.n) fastdiv_init(g.ext[i], g.fmul[i], g.fshift[i]);
    }
}
__host__ __device__ inline uint32_t fastdiv(uint32_t n, uint32_t mul, uint32_t shift) { return (uint32_t)(((uint64_t)n * mul) >> shift); }

enum { ROLE_A = 0, ROLE_B = 1, ROLE_C = 2 };

struct Plan {
    Grp     b, m, n, k;       // batch, M, N, K groups
    Grp     t;                // all target indices (batch+M+N) sorted by C stride: elementwise / scaling
    Grp     rows;             // batch+M merged as "rows" for the streaming reductions
    int32_t family;           // eb200_kernel_family
    int32_t swapped;          // A and B (and M,N) were exchanged so that the streamed operand is "A"
    int32_t a_kmajor;         // tile kernels: A's fastest-varying direction is K (else M)
    int32_t b_kmajor;         // tile kernels: B's fastest-varying direction is K (else N)
    int32_t c_has_repeat;     // C carries a repeated index: the whole C must be pre-scaled
    int32_t split_k;
    int32_t empty_c;          // C (as enumerated) has zero elements: nothing to do
    int32_t empty_k;          // some contracted extent is 0: C = beta*C only
};

// Builds the plan. Returns EB200_OK or an error (message set).
int make_plan(const eb200_contract_desc &d, Plan &p, int sm_count_hint);

// device+host mixed-radix decode: linear position x in [0,g.tot) -> sum of idx*stride for one role
__host__ __device__ inline int64_t grp_offset(const Grp &g, int64_t x, int role) {
    int64_t off = 0;
    if (g.n == 1) return x * g.s[role][0];
    for (int i = 0; i < g.n; i++) {
        int64_t q = x / g.ext[i];
        int64_t r = x - q * g.ext[i];
        off += r * g.s[role][i];
        x = q;
    }
    return off;
}

// two roles at once (shares the divisions)
__host__ __device__ inline void grp_offset2(const Grp &g, int64_t x, int r0, int r1, int64_t &o0, int64_t &o1) {
    o0 = 0;
    o1 = 0;
    if (g.tot <= 0x7fffffffLL) {
        uint32_t y = (uint32_t)x;
        for (int i = 0; i < g.n; i++) {
            uint32_t e = (uint32_t)g.ext[i];
            uint32_t q = fastdiv(y, g.fmul[i], g.fshift[i]);
            uint32_t r = y - q * e;
            o0 += (int64_t)r * g.s[r0][i];
            o1 += (int64_t)r * g.s[r1][i];
            y = q;
        }
    } else {
        for (int i = 0; i < g.n; i++) {
            int64_t q = x / g.ext[i];
            int64_t r = x - q * g.ext[i];
            o0 += r * g.s[r0][i];
            o1 += r * g.s[r1][i];
            x = q;
        }
    }
}

__host__ __device__ inline void grp_offset3(const Grp &g, int64_t x, int64_t &oa, int64_t &ob, int64_t &oc) {
    oa = ob = oc = 0;
    if (g.tot <= 0x7fffffffLL) {
        uint32_t y = (uint32_t)x;
        for (int i = 0; i < g.n; i++) {
            uint32_t e = (uint32_t)g.ext[i];
            uint32_t q = fastdiv(y, g.fmul[i], g.fshift[i]);
            uint32_t r = y - q * e;
            oa += (int64_t)r * g.s[0][i];
            ob += (int64_t)r * g.s[1][i];
            oc += (int64_t)r * g.s[2][i];
            y = q;
        }
    } else {
        for (int i = 0; i < g.n; i++) {
            int64_t q = x / g.ext[i];
            int64_t r = x - q * g.ext[i];
            oa += r * g.s[0][i];
            ob += r * g.s[1][i];
            oc += r * g.s[2][i];
            x = q;
        }
    }
}

} // namespace eb200
