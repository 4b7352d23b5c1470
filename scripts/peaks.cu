// Micro-benchmark of the CUDA-core pipes the K-fused control-cycle kernel is bound by on B200:
// FP32 FFMA, FP64 DFMA and MUFU (ex2) throughput, plus the SM clock seen while running them.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peaks scripts/peaks.cu && ./peaks
#include <cstdio>
#include <cuda_runtime.h>

template <typename T>
__global__ void fma_kernel(T* out, int iters, T a, T b) {
    T x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll 4
        for (int k = 0; k < 4; ++k) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void mufu_kernel(float* out, int iters) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 0.1f, x2 = x0 + 0.2f, x3 = x0 + 0.3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll 8
        for (int k = 0; k < 8; ++k) {
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x2));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x3));
            x0 -= 1.0f; x1 -= 1.0f; x2 -= 1.0f; x3 -= 1.0f;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}

template <typename F>
static double time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 4096;
    void* buf; cudaMalloc(&buf, (size_t)blocks * threads * 8);
    const double fma_per_launch = (double)blocks * threads * iters * 32.0;
    double ms32 = time_ms([&] { fma_kernel<float><<<blocks, threads>>>((float*)buf, iters, 1.0000001f, 1e-9f); });
    double ms64 = time_ms([&] { fma_kernel<double><<<blocks, threads>>>((double*)buf, iters, 1.0000001, 1e-9); });
    const double mufu_per_launch = (double)blocks * threads * iters * 32.0;
    double msmu = time_ms([&] { mufu_kernel<<<blocks, threads>>>((float*)buf, iters); });
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp32_ffma_tflops\": %.2f, \"fp64_dfma_tflops\": %.2f, \"mufu_ex2_gops\": %.1f, "
           "\"fp32_warp_instr_per_clk_per_sm_at_1965MHz\": %.2f, \"how\": \"8 independent FMA chains per thread, %d blocks x %d threads, best of 5; "
           "2 flops per FMA; MUFU: 4 independent ex2 chains\"}\n",
           p.name, p.multiProcessorCount, 2e-9 * fma_per_launch / ms32, 2e-9 * fma_per_launch / ms64, 1e-6 * mufu_per_launch / msmu,
           fma_per_launch / 32.0 / (ms32 * 1e-3) / 1.965e9 / p.multiProcessorCount, blocks, threads);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
