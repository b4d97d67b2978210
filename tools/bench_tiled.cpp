// Tiled GEMM throughput through the C++ shim (TiledDeviceTensor + linear_algebra::gemm): dense grid and half-empty grid,
// next to the same product on one dense DeviceTensor. FP64. Build: make -C tests/cpp bench_tiled
#include <Einsums/LinearAlgebra.hpp>
#include <Einsums/Tensor/DeviceTensor.hpp>
#include <Einsums/Tensor/TiledTensor.hpp>
#include <Einsums/TensorAlgebra.hpp>

#include <chrono>
#include <cstdio>

using namespace einsums;

template <typename F> double time_ms(F &&fn, int reps) {
    fn();
    gpu::device_synchronize();
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; r++) fn();
    gpu::device_synchronize();
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / reps;
}

int main() {
    einsums::initialize(0);
    gpu::set_async(true);
    for (auto [grid, tile] : {std::pair{6, 1024}, std::pair{12, 512}, std::pair{24, 256}, std::pair{48, 128}}) {
        std::vector<int64_t> sizes(size_t(grid), tile);
        for (int keep_mod : {1, 2}) { // 1: every tile present; 2: checkerboard (half the tiles absent)
            TiledDeviceTensor<double, 2> A("A", sizes), B("B", sizes), C("C", sizes);
            std::vector<double> h(size_t(tile) * tile, 0.25);
            double products = 0;
            for (int i = 0; i < grid; i++)
                for (int j = 0; j < grid; j++)
                    if ((i + j) % keep_mod == 0) {
                        A.tile(i, j).read(h);
                        B.tile(i, j).read(h);
                    }
            for (int i = 0; i < grid; i++)
                for (int j = 0; j < grid; j++)
                    for (int k = 0; k < grid; k++)
                        if ((i + k) % keep_mod == 0 && (k + j) % keep_mod == 0) products += 1;
            double flops = products * 2.0 * tile * double(tile) * tile;
            eb200_launch_count_reset();
            double ms = time_ms([&] { linear_algebra::gemm<false, false>(1.0, A, B, 0.0, &C); }, 5);
            long launches = long(eb200_launch_count()) / 6;
            std::printf("{\"grid\": %d, \"tile\": %d, \"present\": \"%s\", \"tile_products\": %.0f, \"ms\": %.3f, \"tflops\": %.2f, \"launches_per_call\": %ld}\n", grid, tile,
                        keep_mod == 1 ? "all" : "checkerboard", products, ms, flops / ms / 1e9, launches);
        }
    }
    {
        int64_t n = 6144;
        DeviceTensor<double, 2> A("A", n, n), B("B", n, n), C("C", n, n);
        A.set_all(0.25); B.set_all(0.25);
        double ms = time_ms([&] { linear_algebra::gemm<false, false>(1.0, A, B, 0.0, &C); }, 5);
        std::printf("{\"dense\": %ld, \"ms\": %.3f, \"tflops\": %.2f}\n", long(n), ms, 2.0 * n * n * n / ms / 1e9);
    }
    einsums::finalize();
    return 0;
}
