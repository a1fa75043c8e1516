// Development aid: where the warps of the REAL K1 (k1_lse.cuh) spend their cycles, on synthetic rows.
//   k1_probe [rows] [V] [bf16: 0|1] [NW: 8|16|24] [dead_every: 0 = all rows live, n = every n-th row dead]
#define MRNNT_K1_TRACE
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../include/mrnnt_b200/k1_lse.cuh"
using namespace mrnnt;

template <typename E, int NW, int C>
static void run(const void *acts, const int *labels, const int *meta, RawRow *lp, int64_t rows, int V, const StreamTiling &tl) {
    auto kern = k1_lse_tma_kernel<E, NW, C, false>;
    cudaError_t ea = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tl.smem_bytes);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    printf("regs=%d smem_dyn_max=%d static=%zu attr=%s\n", fa.numRegs, fa.maxDynamicSharedSizeBytes, fa.sharedSizeBytes, cudaGetErrorString(ea));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        long long zero[32][4] = {};
        cudaMemcpyToSymbol(g_k1_trace, zero, sizeof(zero));
        cudaEventRecord(e0);
        kern<<<148, (NW + 1) * 32, tl.smem_bytes>>>((const E *)acts, labels, meta, lp, rows, V, 0, tl.G, tl.stages, ZeroFill{}, tl.smem_bytes);
        cudaError_t el = cudaGetLastError();
        cudaEventRecord(e1); cudaError_t es = cudaEventSynchronize(e1);
        if (rep == 0) printf("launch: %s, sync: %s\n", cudaGetErrorString(el), cudaGetErrorString(es));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    long long h[32][4]; cudaMemcpyFromSymbol(h, g_k1_trace, sizeof(h));
    double cw = 0, cb = 0, n = 0;
    for (int w = 0; w < NW; ++w) { cw += h[w][0]; cb += h[w][1]; n += h[w][2]; }
    printf("rows=%lld V=%d %s NW=%d G=%d stages=%d: %.1f us, %.0f GB/s (%s)\n", (long long)rows, V, sizeof(E) == 2 ? "bf16" : "f32",
           NW, tl.G, tl.stages, best * 1e3, rows * (double)V * sizeof(E) / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    printf("  CTA 0 consumers: %.0f rows per warp; per row: %.0f cycles waiting for the tile, %.0f cycles working\n", n / NW,
           cw / n, cb / n);
    printf("  CTA 0 producer : %lld tiles; per tile %.0f cycles waiting for a free stage\n", h[NW][2],
           (double)h[NW][0] / (double)h[NW][2]);
}

int main(int argc, char **argv) {
    const int64_t rows = argc > 1 ? atoll(argv[1]) : 196800;
    const int V = argc > 2 ? atoi(argv[2]) : 1000;
    const int bf16 = argc > 3 ? atoi(argv[3]) : 0;
    const int NW = argc > 4 ? atoi(argv[4]) : 24;
    const int dead_every = argc > 5 ? atoi(argv[5]) : 0;
    const size_t es = bf16 ? 2 : 4;
    void *acts; int *labels, *meta; RawRow *lp;
    cudaMalloc(&acts, rows * V * es); cudaMemset(acts, 0x3c, rows * V * es);
    cudaMalloc(&labels, 4096); cudaMemset(labels, 0, 4096);
    std::vector<int> hm(rows);
    for (int64_t i = 0; i < rows; ++i) hm[i] = (dead_every > 0 && i % dead_every == 0) ? kRowDead : (int)(i % 1000);
    cudaMalloc(&meta, rows * 4); cudaMemcpy(meta, hm.data(), rows * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&lp, rows * sizeof(RawRow));
    StreamTiling tl;
    if (!stream_tiling(V, es, sizeof(int), NW, kK1TileTarget, false, &tl)) { printf("no tiling\n"); return 1; }
    const int NV = V / (16 / (int)es);
#define DISPATCH(E, NWc) do { if (NV <= (32 / Elem<E>::kPerVec) * 32) run<E, NWc, 32 / Elem<E>::kPerVec>(acts, labels, meta, lp, rows, V, tl); \
                              else run<E, NWc, 0>(acts, labels, meta, lp, rows, V, tl); } while (0)
    if (bf16) { if (NW == 8) DISPATCH(__nv_bfloat16, 8); else if (NW == 16) DISPATCH(__nv_bfloat16, 16); else DISPATCH(__nv_bfloat16, 24); }
    else { if (NW == 8) DISPATCH(float, 8); else if (NW == 16) DISPATCH(float, 16); else DISPATCH(float, 24); }
    return 0;
}
