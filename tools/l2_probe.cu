// Development aid for SURVEY 8f-f3 (per-utterance pipelining / L2 reuse): how much of what one pass over a chunk of
// logits leaves in the 126 MB L2 is still there when a later pass reads the chunk again, while other chunks are being
// read and gradients are being written in between?
//
// The pipeline that is modelled, inside ONE persistent kernel (separate launches per step are launch-bound at these
// sizes, and ncu empties the L2 between launches): every CTA walks the same sequence of steps on its own slice of the
// chunks, without any barrier -- step i: read chunk i ("K1", with the L2 policy under test); read chunk i - depth again
// ("K3", evict-first: last use) and write that chunk's gradients with streaming stores.  With reuse = 0 the second read
// goes to memory nobody has touched (what three separate kernels do today).  Printed: the time of the whole pipeline
// for both; under  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum  the DRAM bytes of each variant.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/l2_probe tools/l2_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

enum { POL_NORMAL = 0, POL_EVICT_LAST = 1, POL_EVICT_FIRST = 2 };

template <int POL>
__device__ __forceinline__ uint64_t make_policy() {
    uint64_t p = 0;
    if (POL == POL_EVICT_LAST) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    if (POL == POL_EVICT_FIRST) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

template <int POL>
__device__ __forceinline__ uint4 ld_hint(const uint4 *p, uint64_t pol) {
    uint4 v;
    if (POL == POL_NORMAL)
        asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else
        asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}


template <int POL>
__global__ void __launch_bounds__(512, 2) pipeline(const uint4 *pool, uint4 *grads, size_t n16, size_t g16, int nchunks, int depth,
                                                   int reuse, unsigned *sink) {
    const uint64_t pol = make_policy<POL>();
    const uint64_t pol_last_use = make_policy<POL_EVICT_FIRST>();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned acc = 0;
    for (int i = 0; i < nchunks + depth; ++i) {
        if (i < nchunks) {
            const uint4 *src = pool + (size_t)i * n16;
            size_t k = t0;
            for (; k + 3 * stride < n16; k += 4 * stride) {
                uint4 a = ld_hint<POL>(src + k, pol), b = ld_hint<POL>(src + k + stride, pol), c = ld_hint<POL>(src + k + 2 * stride, pol),
                      d = ld_hint<POL>(src + k + 3 * stride, pol);
                acc += a.x ^ b.y ^ c.z ^ d.w;
            }
            for (; k < n16; k += stride) acc += ld_hint<POL>(src + k, pol).x;
        }
        const int j = i - depth;
        if (j >= 0) {
            const uint4 *src = pool + (size_t)(reuse ? j : nchunks + j) * n16;
            uint4 *dst = grads + (size_t)j * g16;
            size_t k = t0;
            for (; k + 3 * stride < n16; k += 4 * stride) {
                uint4 a = ld_hint<POL_EVICT_FIRST>(src + k, pol_last_use), b = ld_hint<POL_EVICT_FIRST>(src + k + stride, pol_last_use),
                      c = ld_hint<POL_EVICT_FIRST>(src + k + 2 * stride, pol_last_use), d = ld_hint<POL_EVICT_FIRST>(src + k + 3 * stride, pol_last_use);
                acc += a.x ^ b.y ^ c.z ^ d.w;
            }
            for (; k < n16; k += stride) acc += ld_hint<POL_EVICT_FIRST>(src + k, pol_last_use).x;
            for (k = t0; k < g16; k += stride)
                asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(dst + k), "r"(acc), "r"(1u), "r"(2u), "r"(3u) : "memory");
        }
    }
    if (acc == 0x12345u) *sink = acc;
}

__global__ void __launch_bounds__(512) flush_write(uint4 *dst, size_t n16) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = make_uint4(7u, 7u, 7u, 7u);
}

int main(int argc, char **argv) {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t MB = 1 << 20;
    const int nchunks = 40;
    if (getenv("L2P_PERSIST_MB") != nullptr) {  // set aside part of the L2 for evict_last ("persisting") lines
        int max_persist = 0;
        CK(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, 0));
        size_t want = (size_t)atoi(getenv("L2P_PERSIST_MB")) * MB;
        if (want > (size_t)max_persist) want = (size_t)max_persist;
        CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
        size_t got = 0;
        CK(cudaDeviceGetLimit(&got, cudaLimitPersistingL2CacheSize));
        printf("persisting L2 set-aside: max %.1f MB, set %.1f MB\n", max_persist / 1048576.0, got / 1048576.0);
    }
    const float write_ratio = argc > 1 ? (float)atof(argv[1]) : 1.37f;  // gradient bytes per live logit byte (c2: 24.6 / 18)
    const bool quick = getenv("L2P_QUICK") != nullptr;                  // the subset that is run under ncu
    const int chunk_mb_list[] = {18, 9, 36};
    const int depth_list[] = {1, 2, 3, 4, 5, 6, 8};
    const size_t pool_bytes = (size_t)2 * nchunks * 36 * MB;
    uint4 *pool, *grads, *flush;
    unsigned *sink;
    CK(cudaMalloc(&pool, pool_bytes));
    CK(cudaMalloc(&grads, (size_t)(nchunks * 36 * MB * write_ratio) + MB));
    CK(cudaMalloc(&flush, 512 * MB));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(pool, 1, pool_bytes));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const char *pol_name[2] = {"normal", "evict_last"};
    for (int chunk_mb : chunk_mb_list) {
        if (quick && chunk_mb != 18) continue;
        for (int pol = 0; pol < 2; ++pol) {
            for (int depth : depth_list) {
                if (quick && (depth == 5 || depth == 8)) continue;
                const size_t cb = (size_t)chunk_mb * MB, n16 = cb / 16;
                const size_t g16 = (size_t)(cb * write_ratio) / 16;
                float ms[2];
                for (int reuse = 0; reuse < 2; ++reuse) {
                    flush_write<<<sms * 4, 512>>>(flush, 512 * MB / 16);
                    CK(cudaEventRecord(e0));
                    if (pol == 0) pipeline<POL_NORMAL><<<2 * sms, 512>>>(pool, grads, n16, g16, nchunks, depth, reuse, sink);
                    else pipeline<POL_EVICT_LAST><<<2 * sms, 512>>>(pool, grads, n16, g16, nchunks, depth, reuse, sink);
                    CK(cudaEventRecord(e1));
                    CK(cudaDeviceSynchronize());
                    CK(cudaEventElapsedTime(&ms[reuse], e0, e1));
                }
                const double bytes = (double)nchunks * (2.0 * cb + g16 * 16.0);
                printf("chunk %2d MB policy %-10s depth %d (%3d MB of logits between the two reads): separate %.1f us (%.0f GB/s)  "
                       "reuse %.1f us (%.0f GB/s algorithmic)  ratio %.3f\n",
                       chunk_mb, pol_name[pol], depth, chunk_mb * depth, ms[0] * 1e3, bytes / ms[0] / 1e6, ms[1] * 1e3,
                       bytes / ms[1] / 1e6, ms[1] / ms[0]);
                fflush(stdout);
            }
        }
    }
    return 0;
}
