#include "k_reduce.cuh"
namespace eb200 {
int k_launch_reduce_cplx(int dtype, cudaStream_t s, bool row, ReduceArgs &a, dim3 grid) {
    return dtype == EB200_C64 ? launch_reduce<float2>(s, row, a, grid) : launch_reduce<double2>(s, row, a, grid);
}
} // namespace eb200
