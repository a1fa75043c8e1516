// How fast can a kernel pull pinned host memory over PCIe?  (mrnnt_upload_acts reads the live rows straight
// from the host.)  Compares the copy engine with (a) 16-byte loads from the device's view of the block, at several
// occupancies / loads in flight, and (b) bulk copies (cp.async.bulk, TMA engine) host -> shared -> device in chunks.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -Iinclude tools/h2d_probe.cu -o tools/h2d_probe && tools/h2d_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "mrnnt_b200/common.cuh"
#include "mrnnt_b200/zero_fill.cuh"

using namespace mrnnt;

template <int U>
__global__ void ldg_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, int64_t n) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * U;
    for (int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x) * U + threadIdx.x; i < n; i += stride) {
        uint4 v[U];
#pragma unroll
        for (int j = 0; j < U; ++j)
            if (i + j * blockDim.x < n) v[j] = __ldcs(src + i + j * blockDim.x);
#pragma unroll
        for (int j = 0; j < U; ++j)
            if (i + j * blockDim.x < n) dst[i + j * blockDim.x] = v[j];
    }
}

// one thread per CTA drives NBUF buffers of `chunk` bytes: bulk load host -> shared, bulk store shared -> device
template <int NBUF>
__global__ void bulk_kernel(const unsigned char *__restrict__ src, unsigned char *__restrict__ dst, int64_t bytes, int chunk) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + static_cast<size_t>(NBUF) * chunk);
    if (threadIdx.x != 0) return;
    for (int i = 0; i < NBUF; ++i) mbar_init(bar + i, 1);
    mbar_init_fence();
    const int64_t nchunks = (bytes + chunk - 1) / chunk;
    const int64_t nloc = blockIdx.x < nchunks ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto load = [&](int64_t k) {
        const int b = static_cast<int>(k % NBUF);
        const int64_t off = (blockIdx.x + k * gridDim.x) * chunk;
        const uint32_t nb = static_cast<uint32_t>(bytes - off < chunk ? bytes - off : chunk);
        mbar_arrive_expect_tx(bar + b, nb);
        bulk_g2s(smem + static_cast<size_t>(b) * chunk, src + off, nb, bar + b);
    };
    for (int64_t k = 0; k < NBUF && k < nloc; ++k) load(k);
    for (int64_t k = 0; k < nloc; ++k) {
        const int b = static_cast<int>(k % NBUF);
        mbar_wait(bar + b, static_cast<uint32_t>((k / NBUF) & 1));
        const int64_t off = (blockIdx.x + k * gridDim.x) * chunk;
        const uint32_t nb = static_cast<uint32_t>(bytes - off < chunk ? bytes - off : chunk);
        bulk_s2g(dst + off, smem + static_cast<size_t>(b) * chunk, nb);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (k + NBUF < nloc) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            load(k + NBUF);
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
static void timeit(const char *name, int64_t bytes, F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        f();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const cudaError_t err = cudaGetLastError();
    printf("%-44s %8.3f ms  %6.1f GB/s  %s\n", name, best, bytes / best / 1e6, err == cudaSuccess ? "" : cudaGetErrorString(err));
    fflush(stdout);
}

int main() {
    const int64_t bytes = 768ll << 20;
    unsigned char *h, *d;
    if (cudaHostAlloc(&h, bytes, cudaHostAllocDefault) != cudaSuccess) return 1;
    if (cudaMalloc(&d, bytes) != cudaSuccess) return 1;
    for (int64_t i = 0; i < bytes; i += 4096) h[i] = static_cast<unsigned char>(i >> 12);
    void *hd = nullptr;
    cudaHostGetDevicePointer(&hd, h, 0);
    const uint4 *s4 = static_cast<const uint4 *>(hd);
    uint4 *d4 = reinterpret_cast<uint4 *>(d);
    const int64_t n4 = bytes / 16;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);

    timeit("copy engine (cudaMemcpyAsync)", bytes, [&] { cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, 0); });
    timeit("ldg U=8  grid=8/SM x256", bytes, [&] { ldg_kernel<8><<<sms * 8, 256>>>(s4, d4, n4); });
    timeit("ldg U=8  grid=2/SM x256", bytes, [&] { ldg_kernel<8><<<sms * 2, 256>>>(s4, d4, n4); });
    timeit("ldg U=8  grid=1/SM x128", bytes, [&] { ldg_kernel<8><<<sms, 128>>>(s4, d4, n4); });
    timeit("ldg U=4  grid=1/SM x64", bytes, [&] { ldg_kernel<4><<<sms, 64>>>(s4, d4, n4); });
    timeit("ldg U=2  grid=4/SM x256", bytes, [&] { ldg_kernel<2><<<sms * 4, 256>>>(s4, d4, n4); });
    timeit("ldg U=8  grid=32 CTAs x256", bytes, [&] { ldg_kernel<8><<<32, 256>>>(s4, d4, n4); });
    const int chunks[] = {2048, 4096, 16384, 32768};
    for (int c : chunks) {
        const size_t smem = 4 * static_cast<size_t>(c) + 64;
        cudaFuncSetAttribute(bulk_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        char name[96];
        for (int per_sm : {1, 2}) {
            snprintf(name, sizeof name, "bulk chunk=%d NBUF=4 grid=%d/SM", c, per_sm);
            timeit(name, bytes, [&] { bulk_kernel<4><<<sms * per_sm, 32, smem>>>(h == hd ? h : static_cast<unsigned char *>(hd), d, bytes, c); });
        }
    }
    // check the last copy
    unsigned char probe[4];
    for (int i = 0; i < 4; ++i) cudaMemcpy(probe + i, d + (int64_t(i) * 1234567 / 4096) * 4096, 1, cudaMemcpyDeviceToHost);
    for (int i = 0; i < 4; ++i)
        if (probe[i] != static_cast<unsigned char>((int64_t(i) * 1234567 / 4096))) printf("MISMATCH at probe %d\n", i);
    return 0;
}
