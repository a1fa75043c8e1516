// Development aid: how fast can one B200 WRITE zeros?  K3 on an alignment-restricted batch (c5) is 95 % zero rows,
// so its floor is the write-only bandwidth, not the copy bandwidth.  Variants: cudaMemsetAsync; grid-stride
// STG.128 (default and .cs); bulk shared->global copies of a zeroed shared-memory buffer issued by one thread per
// CTA (and by one lane of each of several warps).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../include/mrnnt_b200/common.cuh"
using namespace mrnnt;

template <bool CS>
__global__ void stg_zero(uint4 *out, int64_t n4) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        if (CS) st_stream_u4(out + i, z); else out[i] = z;
    }
}

// rows of `row_bytes`; a warp zeroes one row at a time, like K3's consumer warps do
__global__ void warp_rows_zero(uint4 *out, int64_t rows, int row_u4) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int64_t r = (int64_t)blockIdx.x * nw + warp; r < rows; r += (int64_t)gridDim.x * nw) {
        uint4 *p = out + r * row_u4;
        for (int j = lane; j < row_u4; j += 32) st_stream_u4(p + j, z);
    }
}

__device__ __forceinline__ void bulk_s2g(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// each issuing warp (lane 0) pushes `chunk` bytes of zeros per bulk copy
__global__ void bulk_zero(unsigned char *out, int64_t bytes, int chunk) {
    extern __shared__ __align__(128) unsigned char zbuf[];
    for (int i = threadIdx.x * 16; i < chunk; i += blockDim.x * 16) *reinterpret_cast<uint4 *>(zbuf + i) = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (lane != 0) return;
    const int64_t nchunks = bytes / chunk;
    for (int64_t c = (int64_t)blockIdx.x * nw + warp; c < nchunks; c += (int64_t)gridDim.x * nw) {
        bulk_s2g(out + c * chunk, zbuf, (uint32_t)chunk);
        bulk_commit();
        bulk_wait_read<8>();
    }
    bulk_wait<0>();
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <class F> static float time_it(F f, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) {
        cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best;
    }
    return best;
}

int main(int argc, char **argv) {
    const int64_t rows = argc > 1 ? atoll(argv[1]) : 585600;
    const int V = argc > 2 ? atoi(argv[2]) : 2000;
    const int64_t bytes = rows * V * 4;
    unsigned char *buf; CK(cudaMalloc(&buf, bytes));
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    auto rep = [&](const char *name, float ms) { printf("%-44s %8.1f us  %7.1f GB/s\n", name, ms * 1e3, bytes / ms * 1e-6); };
    rep("cudaMemsetAsync", time_it([&] { cudaMemsetAsync(buf, 0, bytes); }));
    for (int bpsm : {2, 4, 8}) {
        char n[64];
        snprintf(n, 64, "STG.128 grid-stride %dx%d blocks x 512", sms, bpsm);
        rep(n, time_it([&] { stg_zero<false><<<sms * bpsm, 512>>>((uint4 *)buf, bytes / 16); }));
        snprintf(n, 64, "STG.128.cs grid-stride %dx%d blocks x 512", sms, bpsm);
        rep(n, time_it([&] { stg_zero<true><<<sms * bpsm, 512>>>((uint4 *)buf, bytes / 16); }));
    }
    rep("warp-per-row .cs, 148 x 768 threads", time_it([&] { warp_rows_zero<<<sms, 768>>>((uint4 *)buf, rows, V / 4); }));
    rep("warp-per-row .cs, 296 x 768 threads", time_it([&] { warp_rows_zero<<<sms * 2, 768>>>((uint4 *)buf, rows, V / 4); }));
    for (int chunk : {8000, 16000, 32000, 64000}) {
        for (int warps : {1, 4, 8}) {
            char n[64];
            snprintf(n, 64, "bulk S2G %d B x %d issuing warps/CTA", chunk, warps);
            CK(cudaFuncSetAttribute(bulk_zero, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
            rep(n, time_it([&] { bulk_zero<<<sms, warps * 32, chunk>>>(buf, bytes, chunk); }));
        }
    }
    // per-SM caps: the same two ways of storing, from a subset of the SMs, on 1 GB
    {
        const int64_t sub_bytes = (int64_t)1 << 30;
        auto rep2 = [&](const char *name, int ctas, float ms) {
            printf("%-40s %3d CTAs %8.1f us  %7.1f GB/s  %6.1f GB/s per SM\n", name, ctas, ms * 1e3, sub_bytes / ms * 1e-6, sub_bytes / ms * 1e-6 / ctas);
        };
        for (int ctas : {16, 32, 64, 148}) {
            for (int warps : {1, 2, 4}) {
                char n[64];
                snprintf(n, 64, "bulk S2G 8192 B, %d warps", warps);
                rep2(n, ctas, time_it([&] { bulk_zero<<<ctas, warps * 32, 8192>>>(buf, sub_bytes, 8192); }));
            }
            for (int warps : {2, 4, 8, 16}) {
                char n[64];
                snprintf(n, 64, "STG.128.cs, %d warps", warps);
                rep2(n, ctas, time_it([&] { stg_zero<true><<<ctas, warps * 32>>>((uint4 *)buf, sub_bytes / 16); }));
            }
        }
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
