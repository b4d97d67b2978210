#include "k_reduce.cuh"
namespace eb200 {
int k_launch_reduce_real(int dtype, cudaStream_t s, bool row, ReduceArgs &a, dim3 grid) {
    return dtype == EB200_F32 ? launch_reduce<float>(s, row, a, grid) : launch_reduce<double>(s, row, a, grid);
}
int k_launch_reduce(int dtype, cudaStream_t s, bool row, ReduceArgs &a, dim3 grid) {
    return (dtype == EB200_F32 || dtype == EB200_F64) ? k_launch_reduce_real(dtype, s, row, a, grid) : k_launch_reduce_cplx(dtype, s, row, a, grid);
}
} // namespace eb200
