// k_reduce.cuh — launch dispatch of the streaming reductions, shared by the real and complex translation units.
#pragma once
#include "launch_api.cuh"
#include <type_traits>
namespace eb200 {
namespace {
// 0: flat kernel; 1: segmented, scalar loads; 2: segmented, 16-byte pairs
template <typename T> int reduce_seg_mode(const ReduceArgs &a) {
    if (a.k.n == 0 || a.k.ext[0] < 32) return 0;
    if (sizeof(T) > 8 || a.k.s[ROLE_A][0] != 1 || a.k.s[ROLE_B][0] != 1 || (a.k.ext[0] & 1)) return 1;
    auto all_even = [](const Grp &g, int role, int skip) {
        for (int i = 0; i < g.n; i++)
            if (i != skip && (g.s[role][i] & 1)) return false;
        return true;
    };
    if (!all_even(a.k, ROLE_A, 0) || !all_even(a.k, ROLE_B, 0) || !all_even(a.rows, ROLE_A, -1) || !all_even(a.rows, ROLE_B, -1) ||
        !all_even(a.n, ROLE_B, -1))
        return 1;
    constexpr uintptr_t AL = 2 * sizeof(T);
    if ((uintptr_t)a.A % AL || (uintptr_t)a.B % AL) return 1;
    return 2;
}

template <typename T, int NN> int launch_reduce_nn(cudaStream_t s, bool row, ReduceArgs &a, dim3 grid) {
    if (row) {
        int mode = reduce_seg_mode<T>(a);
        a.flat_lanes = 0;
        if (mode >= 1) {
            const int64_t vec = (mode == 2) ? 2 : 1, ppr = a.k.ext[0] / vec;
            // flat lane mapping only where whole-warp chunks would leave lanes idle (run length not a multiple of 32 packs)
            if (a.k.tot / vec < 0x7fffffffLL && a.k.ext[0] % vec == 0 && (ppr % 32) != 0 && ppr < 32 * 16) {
                a.flat_lanes = 1;
                fastdiv_init(ppr, a.ppr_mul, a.ppr_shift);
            }
        }
        auto go = [&](auto rb_tag) {
            constexpr int RB = decltype(rb_tag)::value;
            if (mode == 2) {
                if constexpr (sizeof(T) <= 8) row_reduce_seg<T, NN, RB, 2><<<grid, 256, 0, s>>>(a);
            } else if (mode == 1) {
                row_reduce_seg<T, NN, RB, 1><<<grid, 256, 0, s>>>(a);
            } else {
                row_reduce<T, NN, RB><<<grid, 256, 0, s>>>(a);
            }
        };
        // must agree with run_reduce's grid (contract.cu): both call reduce_rows_per_block
        if constexpr (NN == 1) {
            if (reduce_rows_per_block(a.n.tot, a.k.tot) == 8) go(std::integral_constant<int, 8>{});
            else go(std::integral_constant<int, 4>{});
        } else {
            go(std::integral_constant<int, 4>{});
        }
        EB_POST_LAUNCH("row_reduce");
    } else {
        col_reduce<T, NN, 2><<<grid, 256, 0, s>>>(a);
        EB_POST_LAUNCH("col_reduce");
    }
    return EB200_OK;
}
template <typename T> int launch_reduce_c(cudaStream_t s, bool row, ReduceArgs &a, dim3 grid) {
    int nn = (int)a.n.tot;
    if (nn <= 1) return launch_reduce_nn<T, 1>(s, row, a, grid);
    if (nn <= 2) return launch_reduce_nn<T, 2>(s, row, a, grid);
    if (nn <= 4) return launch_reduce_nn<T, 4>(s, row, a, grid);
    return launch_reduce_nn<T, 8>(s, row, a, grid);
}
template <typename T> int launch_reduce(cudaStream_t s, bool row, ReduceArgs &a, dim3 grid) { return launch_reduce_c<T>(s, row, a, grid); }
} // namespace
} // namespace eb200
