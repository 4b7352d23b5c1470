// PCIe probe: how fast can SMs read / write page-locked host memory directly, as a function of request size?
// Decides the request granularity of the host-buffer session's direct I/O path (vfk_session_cycle).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pcie_probe scripts/pcie_probe.cu && ./pcie_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../vfclik_b200/csrc/vfk_tma.cuh"
using namespace vfk;

__device__ __forceinline__ void bulk_s2g(void* dst, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}

// every warp: lane 0 bulk-reads `sz` bytes at grid-strided offsets, double buffered
__global__ void rd_bulk(const char* src, size_t total, int sz, float* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem) + warp * 2;
    unsigned char* buf = smem + 1024 + (size_t)warp * 2 * sz;
    if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    __syncwarp();
    const size_t nreq = total / sz, stride = (size_t)gridDim.x * wpb;
    size_t r = (size_t)blockIdx.x * wpb + warp;
    float acc = 0;
    if (r < nreq && lane == 0) { mbar_arrive_expect_tx(&bars[0], sz); bulk_g2s(buf, src + r * sz, sz, &bars[0]); }
    for (int it = 0; r < nreq; r += stride, ++it) {
        if (r + stride < nreq && lane == 0) {
            mbar_arrive_expect_tx(&bars[(it + 1) & 1], sz);
            bulk_g2s(buf + (size_t)((it + 1) & 1) * sz, src + (r + stride) * sz, sz, &bars[(it + 1) & 1]);
        }
        mbar_wait(&bars[it & 1], (it >> 1) & 1);
        acc += reinterpret_cast<float*>(buf + (size_t)(it & 1) * sz)[lane];
        __syncwarp();
    }
    if (acc == 123.456f) sink[0] = acc;
}

// every warp: plain coalesced loads of 128 B (one float per lane), `rows` independent loads in flight
__global__ void rd_ldg(const float* src, size_t total_floats, float* sink) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_floats; i += stride) acc += src[i];
    if (acc == 123.456f) sink[0] = acc;
}

// every warp: plain coalesced 128 B stores
__global__ void wr_st(float* dst, size_t total_floats) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_floats; i += stride) dst[i] = (float)i;
}

// every warp: stage `sz` bytes in smem, one bulk store
__global__ void wr_bulk(char* dst, size_t total, int sz) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    unsigned char* buf = smem + (size_t)warp * sz;
    for (int k = lane * 4; k < sz; k += 128) *reinterpret_cast<float*>(buf + k) = (float)k;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const size_t nreq = total / sz, stride = (size_t)gridDim.x * wpb;
    for (size_t r = (size_t)blockIdx.x * wpb + warp; r < nreq; r += stride) {
        if (lane == 0) {
            bulk_s2g(dst + r * sz, buf, sz);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
static float time_ms(F f, cudaStream_t s, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaStreamSynchronize(s);
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) { cudaEventRecord(a, s); f(); cudaEventRecord(b, s); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}

int main() {
    const size_t total = 29360128;   // 7 x 1M x 4 B, the session's q block
    char *hin, *hout; float* sink;
    cudaHostAlloc(&hin, total, cudaHostAllocDefault); cudaHostAlloc(&hout, total, cudaHostAllocDefault);
    cudaMalloc(&sink, 4);
    for (size_t i = 0; i < total; ++i) hin[i] = (char)i;
    cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
    const int blocks = 148 * 3, threads = 128;
    printf("request  read-bulk GB/s  write-bulk GB/s  both GB/s(each way)\n");
    for (int sz : {128, 256, 512, 1024, 2048, 4096}) {
        const size_t smem_r = 1024 + (size_t)4 * 2 * sz, smem_w = (size_t)4 * sz;
        cudaFuncSetAttribute(rd_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r);
        cudaFuncSetAttribute(wr_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w);
        float r = time_ms([&] { rd_bulk<<<blocks, threads, smem_r, s1>>>(hin, total, sz, sink); }, s1);
        float w = time_ms([&] { wr_bulk<<<blocks, threads, smem_w, s1>>>(hout, total, sz); }, s1);
        // both directions at once: time on the host around two streams
        cudaDeviceSynchronize();
        cudaEvent_t a, b1, b2; cudaEventCreate(&a); cudaEventCreate(&b1); cudaEventCreate(&b2);
        float best = 1e30f;
        for (int i = 0; i < 5; ++i) {
            cudaEventRecord(a, s1); cudaStreamWaitEvent(s2, a, 0);
            rd_bulk<<<blocks / 2, threads, smem_r, s1>>>(hin, total, sz, sink);
            wr_bulk<<<blocks / 2, threads, smem_w, s2>>>(hout, total, sz);
            cudaEventRecord(b1, s1); cudaEventRecord(b2, s2); cudaEventSynchronize(b1); cudaEventSynchronize(b2);
            float m1, m2; cudaEventElapsedTime(&m1, a, b1); cudaEventElapsedTime(&m2, a, b2);
            float m = m1 > m2 ? m1 : m2; if (m < best) best = m;
        }
        printf("%6d   %8.1f        %8.1f         %8.1f\n", sz, total / r / 1e6, total / w / 1e6, total / best / 1e6);
    }
    float r = time_ms([&] { rd_ldg<<<blocks, 256, 0, s1>>>((const float*)hin, total / 4, sink); }, s1);
    float w = time_ms([&] { wr_st<<<blocks, 256, 0, s1>>>((float*)hout, total / 4); }, s1);
    printf("plain 128 B per warp: read (LDG) %.1f GB/s, write (STG) %.1f GB/s\n", total / r / 1e6, total / w / 1e6);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
