// Development aid: where the time of a loss call goes IN THE STREAM (back-to-back calls, dependent launches, early return)
// -- %globaltimer stamps written by the kernels themselves (common.cuh, MRNNT_TIMELINE), per call.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DMRNNT_TIMELINE -Iinclude -o tools/timeline_probe tools/timeline_probe.cu
//   timeline_probe [T] [S] [B] [V] [calls]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "gpu_rnnt.h"
#include "gpu_workspace_manager.h"
using namespace mrnnt;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void fill_uniform(float *x, size_t n, unsigned seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned h = (unsigned)i * 2654435761u ^ seed;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        x[i] = (h >> 8) * (1.0f / 16777216.0f);
    }
}

int main(int argc, char **argv) {
    const int T = argc > 1 ? atoi(argv[1]) : 150, S = argc > 2 ? atoi(argv[2]) : 40, B = argc > 3 ? atoi(argv[3]) : 32;
    const int V = argc > 4 ? atoi(argv[4]) : 1000, calls = std::min(argc > 5 ? atoi(argv[5]) : 8, kTimelineSlots);
    const size_t rows = (size_t)B * T * (S + 1), n = rows * V;
    float *acts, *grads;
    int *labels, *Td, *Sd;
    CK(cudaMalloc(&acts, n * 4)); CK(cudaMalloc(&grads, n * 4));
    fill_uniform<<<1024, 256>>>(acts, n, 12345u);
    std::vector<int> lab((size_t)B * S), Th(B, T), Sh(B, S);
    for (size_t i = 0; i < lab.size(); ++i) lab[i] = 1 + (int)((i * 7919u) % (V - 1));
    CK(cudaMalloc(&labels, lab.size() * 4)); CK(cudaMalloc(&Td, B * 4)); CK(cudaMalloc(&Sd, B * 4));
    CK(cudaMemcpy(labels, lab.data(), lab.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(Td, Th.data(), B * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(Sd, Sh.data(), B * 4, cudaMemcpyHostToDevice));
    cudaStream_t stream;
    CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    GpuRNNTWorkspaceManager<float> mgr(acts, labels, B, Td, Sd, V);
    if (mgr.create_workspace() != RNNT_STATUS_SUCCESS) { printf("create_workspace failed\n"); return 1; }
    GpuRNNTComputer<float> comp(mgr, 0, stream);
    std::vector<float> costs(B);
    for (int i = 0; i < 5; ++i) comp.cost_and_grad(costs.data(), grads);   // warm-up
    CK(cudaDeviceSynchronize());
    std::vector<unsigned long long> init((size_t)kTimelineSlots * kTimelineEvents);
    for (int s = 0; s < kTimelineSlots; ++s)
        for (int e = 0; e < kTimelineEvents; ++e)
            init[(size_t)s * kTimelineEvents + e] = (e == 0 || e == 1 || e == 3 || e == 4 || e == 8 || e == 9) ? ~0ull : 0ull;
    CK(cudaMemcpyToSymbol(g_timeline, init.data(), init.size() * 8));
    const int first = mgr.engine().timeline_slot() + 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, stream);
    for (int i = 0; i < calls; ++i) comp.cost_and_grad(costs.data(), grads);
    cudaEventRecord(e1, stream);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<unsigned long long> tl(init.size());
    CK(cudaMemcpyFromSymbol(tl.data(), g_timeline, tl.size() * 8));
    printf("T=%d S=%d B=%d V=%d: %d calls back to back, %.1f us per call (CUDA events); cost[0] = %f\n", T, S, B, V, calls, ms * 1e3 / calls, costs[0]);
    printf("per call, us after the FIRST call's K1 start:  K1 in / past wait / out | K2 in / past wait / recursions done / phases done / fill done | K3 in / past wait / out\n");
    const unsigned long long t0 = tl[(size_t)((first) & (kTimelineSlots - 1)) * kTimelineEvents + 0];
    double prev_end = 0;
    for (int c = 0; c < calls; ++c) {
        const unsigned long long *r = &tl[(size_t)((first + c) & (kTimelineSlots - 1)) * kTimelineEvents];
        auto us = [&](int e) { return ((double)r[e] - (double)t0) * 1e-3; };
        printf("call %d: K1 %7.1f %7.1f %7.1f | K2 %7.1f %7.1f %7.1f %7.1f %7.1f | K3 %7.1f %7.1f %7.1f   (K1 %5.1f  K2 past-wait..phases %5.1f  ..fill %5.1f  K3 past-wait..out %5.1f; gap K3 out -> next K1 past wait: see next line; since previous K3 out: %5.1f)\n",
               c, us(0), us(1), us(2), us(3), us(4), us(5), us(6), us(7), us(8), us(9), us(10),
               us(2) - us(1), us(6) - us(4), us(7) - us(4), us(10) - us(9), us(1) - prev_end);
        prev_end = us(10);
    }
    mgr.free_workspace();
    return 0;
}
