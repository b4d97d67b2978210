// launch_api.cuh — kernel launch entry points, one translation unit per kernel family so that the library builds in
// parallel (contract.cu holds only host logic).
#pragma once
#include "contract_kernels.cuh"
#include "gemm_tma.cuh"
#include "contract_tma.cuh"
#include "gemm_tf32x3.cuh"
#include "gemm_tma_z.cuh"

namespace eb200 {
/// rows a CTA of the row-direction reductions owns. One output column and long rows: 8 rows = 8 independent 16-byte loads per
/// lane in flight (Fock J/K at n=300: 6.5 -> 7.1 TB/s); short rows (K < 16384) keep 4 (more CTAs per SM hide the per-CTA prologue better).
inline int reduce_rows_per_block(int64_t nn, int64_t k_tot) { return (nn <= 1 && k_tot >= 16384) ? 8 : 4; }
struct ConvArgs {
    Grp         g; // ROLE_A = source strides, ROLE_C = destination strides
    const void *src;
    void       *dst;
};
int k_launch_mma(cudaStream_t s, TileArgs &args, dim3 grid);                                    // k_tile_mma.cu
int k_launch_simt(int dtype, cudaStream_t s, TileArgs &args, dim3 grid);                         // k_tile_simt.cu
int k_launch_tma(cudaStream_t s, const CUtensorMap &tmA, const CUtensorMap &tmB, const TmaGemmArgs &a, dim3 grid, size_t smem, bool ak,
                 bool bk);                                                                       // k_tile_tma.cu
int k_launch_tma_generic(cudaStream_t s, const CUtensorMap &tmA, const CUtensorMap &tmB, const TmaGenArgs &a, dim3 grid, int cfg, bool ak,
                         bool bk);                                                               // k_tile_tma.cu
int k_launch_tma_z(cudaStream_t s, const CUtensorMap &tmA, const CUtensorMap &tmB, const TmaZGemmArgs &a, dim3 grid, size_t smem, bool ak,
                   bool bk);                                                                     // k_tile_tma.cu
int k_launch_tf32x3(cudaStream_t s, const CUtensorMap &tmA, const CUtensorMap &tmB, const Tf32GemmArgs &a, dim3 grid, size_t smem, bool a_mn,
                    bool b_mn); // k_tile_tf32.cu
int k_launch_tf32x3_pair(cudaStream_t s, const CUtensorMap &tmA, const CUtensorMap &tmB, const Tf32GemmArgs &a, dim3 grid, size_t smem, bool a_mn,
                         bool b_mn); // k_tile_tf32.cu
int k_launch_tf32x3_generic(cudaStream_t s, const CUtensorMap &tmA, const CUtensorMap &tmB, const Tf32GemmArgs &a, dim3 grid, size_t smem,
                            bool a_mn, bool b_mn, bool pair); // k_tile_tf32.cu
int k_launch_c64_expand(cudaStream_t s, const void *Bc, void *Bp, int64_t rows, int64_t N, int64_t K, int64_t Kp); // k_tile_tf32.cu
int k_launch_reduce(int dtype, cudaStream_t s, bool row, ReduceArgs &a, dim3 grid);              // k_reduce_{real,cplx}.cu
int k_launch_reduce_real(int dtype, cudaStream_t s, bool row, ReduceArgs &a, dim3 grid);
int k_launch_reduce_cplx(int dtype, cudaStream_t s, bool row, ReduceArgs &a, dim3 grid);
int k_launch_dot(int dtype, cudaStream_t s, ReduceArgs &a, int grid);                          // k_misc.cu
int k_launch_elem(int dtype, cudaStream_t s, ElemArgs &a, int conj_a, int conj_b, int grid);     // k_misc.cu
int k_launch_scale(int dtype, cudaStream_t s, ScaleArgs &a, int grid);
int k_launch_finalize(int dtype, cudaStream_t s, FinalizeArgs &a, int grid);
int k_launch_convert(cudaStream_t s, int src_dtype, int dst_dtype, const ConvArgs &a, int grid);
template <typename K> inline int set_smem_attr(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) EB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return EB200_OK;
}
} // namespace eb200
