// Development aid: how fast can the bulk-copy ring (and plain vector loads) pull a 787 MB array through one B200?
// Variants: (1) ring, consumers only arrive; (2) ring, consumers LDS the row and sum it; (3) plain LDG.128
// grid-stride sum; (4) LDG.128 + STG.128 copy (read+write, the shape of the "measured peak").
#include <cstdio>
#include <cuda_runtime.h>
#include "../include/mrnnt_b200/k1_lse.cuh"
using namespace mrnnt;

template <int NW, int MODE>
__global__ void __launch_bounds__((NW + 1) * 32, 1) ring(const float *acts, float *out, int64_t rows, int V, int G, int stages) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const size_t tile_floats = (size_t)G * V;
    float *tiles = (float *)smem_raw;
    uint64_t *full = (uint64_t *)(smem_raw + (size_t)stages * tile_floats * 4);
    uint64_t *empty = full + stages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int i = 0; i < stages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, G); } mbar_init_fence(); }
    __syncthreads();
    const int64_t ntiles = (rows + G - 1) / G;
    const int64_t nloc = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (warp == NW) {
        const uint64_t policy = l2_policy_evict_first();
        int stage = 0; uint32_t phase = 0;
        for (int64_t k = 0; k < nloc; ++k) {
            const int64_t row0 = (blockIdx.x + k * gridDim.x) * G;
            mbar_wait(empty + stage, phase ^ 1u);
            if (lane == 0) {
                const int n = (int)min((int64_t)G, rows - row0);
                mbar_arrive_expect_tx(full + stage, (uint32_t)n * V * 4u);
                bulk_g2s_hint(tiles + stage * tile_floats, acts + row0 * V, (uint32_t)n * V * 4u, full + stage, policy);
            }
            if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
    } else {
        float acc = 0.f;
        const int V4 = V >> 2;
        for (int64_t q = warp; q < nloc * G; q += NW) {
            const int64_t k = q / G; const int r = (int)(q - k * G);
            const int stage = (int)(k % stages); const uint32_t phase = (uint32_t)((k / stages) & 1);
            mbar_wait(full + stage, phase);
            if (MODE == 2) {
                const float4 *x4 = (const float4 *)(tiles + stage * tile_floats + (size_t)r * V);
                for (int j = lane; j < V4; j += 32) { float4 v = x4[j]; acc += (v.x + v.y) + (v.z + v.w); }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + stage);
        }
        if (MODE == 2 && acc == 123.f) out[0] = acc;
    }
}

__global__ void ldg_sum(const float4 *in, float *out, int64_t n4) {
    float acc = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        float4 a = __ldcs(in + i), b = __ldcs(in + i + stride), c = __ldcs(in + i + 2 * stride), d = __ldcs(in + i + 3 * stride);
        acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w + c.x + c.y + c.z + c.w + d.x + d.y + d.z + d.w;
    }
    for (; i < n4; i += stride) { float4 a = __ldcs(in + i); acc += a.x + a.y + a.z + a.w; }
    if (acc == 123.f) out[0] = acc;
}
__global__ void ldg_copy(const float4 *in, float4 *out, int64_t n4) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        float4 a = __ldcs(in + i), b = __ldcs(in + i + stride), c = __ldcs(in + i + 2 * stride), d = __ldcs(in + i + 3 * stride);
        __stcs(out + i, a); __stcs(out + i + stride, b); __stcs(out + i + 2 * stride, c); __stcs(out + i + 3 * stride, d);
    }
    for (; i < n4; i += stride) __stcs(out + i, __ldcs(in + i));
}

template <typename F> float timeit(F f, int iters = 10) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f(); cudaEventRecord(a);
    for (int i = 0; i < iters; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / iters;
}

int main() {
    const int V = 1000; const int64_t rows = 196800; const int64_t n = rows * V;
    float *acts, *out; cudaMalloc(&acts, n * 4); cudaMalloc(&out, n * 4); cudaMemset(acts, 0, n * 4);
    const double gb = n * 4 / 1e9;
    for (int nw : {8, 16}) {
        StreamTiling tl; stream_tiling(V, 4, 0, nw, kK3TileTarget, true, &tl);
        auto run = [&](auto kern, const char *name) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tl.smem_bytes);
            float ms = timeit([&] { kern<<<148, (nw + 1) * 32, tl.smem_bytes>>>(acts, out, rows, V, tl.G, tl.stages); });
            printf("%-34s NW=%2d G=%d stages=%d: %.1f us  %.0f GB/s (%s)\n", name, nw, tl.G, tl.stages, ms * 1e3, gb / ms * 1e3, cudaGetErrorString(cudaGetLastError()));
        };
        if (nw == 8) { run(ring<8, 1>, "ring, consumers arrive only"); run(ring<8, 2>, "ring, consumers LDS+sum"); }
        else { run(ring<16, 1>, "ring, consumers arrive only"); run(ring<16, 2>, "ring, consumers LDS+sum"); }
    }
    for (int bpsm : {4, 8, 16}) {
        float ms = timeit([&] { ldg_sum<<<148 * bpsm, 256>>>((const float4 *)acts, out, n / 4); });
        printf("LDG.128 grid-stride sum, %2d CTA/SM x256: %.1f us  %.0f GB/s\n", bpsm, ms * 1e3, gb / ms * 1e3);
    }
    for (int bpsm : {4, 8, 16}) {
        float ms = timeit([&] { ldg_copy<<<148 * bpsm, 256>>>((const float4 *)acts, (float4 *)out, n / 4); });
        printf("LDG.128+STG.128 copy, %2d CTA/SM x256: %.1f us  %.0f GB/s (read+write)\n", bpsm, ms * 1e3, 2 * gb / ms * 1e3);
    }
    { float ms = timeit([&] { cudaMemcpyAsync(out, acts, n * 4, cudaMemcpyDeviceToDevice); });
      printf("cudaMemcpy D2D: %.1f us  %.0f GB/s (read+write)\n", ms * 1e3, 2 * gb / ms * 1e3); }
    return 0;
}
