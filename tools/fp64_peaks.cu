// fp64_peaks.cu — B200 FP64 roofline denominators measured live (test infrastructure, not product).
// Measures DMMA.8x8x4 (mma.sync m8n8k4 f64) and DFMA issue-bound peaks, their mix, and a plain
// HBM copy, with CUDA events. Prints one JSON object. Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

template <int NACC>
__global__ void __launch_bounds__(256) dmma_loop(double* out, int iters, double seed) {
    double a = seed + threadIdx.x * 1e-9, b = seed - threadIdx.x * 1e-9;
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dfma_loop(double* out, int iters, double seed) {
    double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) c[i] = fma(a, c[i], b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i];
    if (s == 123.456) out[0] = s;
}

// mixed: NM DMMAs and NF DFMAs per iteration
template <int NM, int NF>
__global__ void __launch_bounds__(256) mix_loop(double* out, int iters, double seed) {
    double a = seed + threadIdx.x * 1e-9, b = seed - threadIdx.x * 1e-9;
    double c[NM][2], f[NF];
#pragma unroll
    for (int i = 0; i < NM; i++) { c[i][0] = 0; c[i][1] = 0; }
#pragma unroll
    for (int i = 0; i < NF; i++) f[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NM; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
#pragma unroll
        for (int i = 0; i < NF; i++) f[i] = fma(a, f[i], b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NM; i++) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < NF; i++) s += f[i];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) copy_kernel(const double2* __restrict__ in, double2* __restrict__ out, size_t n2) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n2; i += 4 * stride) {
        double2 v0 = in[i], v1 = in[i + stride], v2 = in[i + 2 * stride], v3 = in[i + 3 * stride];
        out[i] = v0; out[i + stride] = v1; out[i + 2 * stride] = v2; out[i + 3 * stride] = v3;
    }
    for (; i < n2; i += stride) out[i] = in[i];
}

template <typename F>
static float time_ms(F f, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 64));
    const int iters = 20000;
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    // DMMA: 256 FMA per warp-instruction = 512 flop
    {
        double best = 0; int bc = 0, bw = 0;
        for (int ctas = 1; ctas <= 4; ctas *= 2) {
            float ms = time_ms([&] { dmma_loop<8><<<sms * ctas, 256>>>(out, iters, 1.0); }, 5);
            double tf = (double)sms * ctas * 8 /*warps*/ * 8 /*acc*/ * iters * 512.0 / (ms * 1e-3) / 1e12;
            if (tf > best) { best = tf; bc = ctas; bw = 8; }
        }
        printf(", \"dmma_tflops\": %.3f, \"dmma_ctas_per_sm\": %d", best, bc); (void)bw;
        float ms = time_ms([&] { dmma_loop<4><<<sms, 128>>>(out, iters, 1.0); }, 5);
        printf(", \"dmma_1warp_per_smsp_4acc_tflops\": %.3f", (double)sms * 4 * 4 * iters * 512.0 / (ms * 1e-3) / 1e12);
        ms = time_ms([&] { dmma_loop<1><<<sms, 128>>>(out, iters, 1.0); }, 5);
        printf(", \"dmma_latency_ns_per_dep_instr\": %.3f", ms * 1e6 / iters);
    }
    {
        double best = 0;
        for (int ctas = 1; ctas <= 4; ctas *= 2) {
            float ms = time_ms([&] { dfma_loop<16><<<sms * ctas, 256>>>(out, iters, 1.0); }, 5);
            double tf = (double)sms * ctas * 256 * 16.0 * iters * 2.0 / (ms * 1e-3) / 1e12;
            if (tf > best) best = tf;
        }
        printf(", \"dfma_tflops\": %.3f", best);
    }
    {
        // 8 DMMA (2048 FMA/warp) + 16 DFMA (512 FMA/warp)
        float ms = time_ms([&] { mix_loop<8, 16><<<sms * 2, 256>>>(out, iters, 1.0); }, 5);
        double fl = (double)sms * 2 * 8 * iters * (8 * 512.0 + 16 * 64.0);
        printf(", \"mix_8dmma_16dfma_tflops\": %.3f", fl / (ms * 1e-3) / 1e12);
        ms = time_ms([&] { mix_loop<8, 64><<<sms * 2, 256>>>(out, iters / 4, 1.0); }, 5);
        fl = (double)sms * 2 * 8 * (iters / 4) * (8 * 512.0 + 64 * 64.0);
        printf(", \"mix_8dmma_64dfma_tflops\": %.3f", fl / (ms * 1e-3) / 1e12);
    }
    {
        size_t n = (size_t)1 << 28;  // 2 GiB of doubles in, 2 GiB out
        double2 *a, *b; CK(cudaMalloc(&a, n * 8)); CK(cudaMalloc(&b, n * 8));
        CK(cudaMemset(a, 1, n * 8)); CK(cudaMemset(b, 0, n * 8));
        float ms = time_ms([&] { copy_kernel<<<sms * 8, 256>>>(a, b, n / 2); }, 10);
        printf(", \"copy_kernel_gbs\": %.1f", 2.0 * n * 8 / (ms * 1e-3) / 1e9);
        ms = time_ms([&] { CK(cudaMemcpyAsync(b, a, n * 8, cudaMemcpyDeviceToDevice)); }, 10);
        printf(", \"memcpy_d2d_gbs\": %.1f", 2.0 * n * 8 / (ms * 1e-3) / 1e9);
        cudaFree(a); cudaFree(b);
    }
    int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf(", \"sm_clock_khz_max\": %d}\n", clk);
    return 0;
}
