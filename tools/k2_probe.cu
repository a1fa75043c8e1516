// Development aid: cycles per frame of the REAL K2 alpha/beta passes (k2_lattice.cuh) on synthetic lp data,
// one CTA, with clock64 around each pass.
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../include/mrnnt_b200/k2_lattice.cuh"
using namespace mrnnt;

template <int K>
__global__ void __launch_bounds__(kK2Threads) probe(K2Args a, long long *cyc) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int warp = threadIdx.x >> 5;
    long long t0 = clock64();
    if (warp == 0) { k2_alpha_pass<K>(a, 0, sm); if ((threadIdx.x & 31) == 0) cyc[0] = clock64() - t0; }
    else if (warp == 1 && a.need_beta) { k2_beta_pass<K>(a, 0, sm + (kK2ChunkBufs * (size_t)a.chunk_frames * (a.S_max + 1) * sizeof(double2) + 64)); if ((threadIdx.x & 31) == 0) cyc[1] = clock64() - t0; }
    __syncthreads();
    long long t1 = clock64();
    if (a.need_beta) k2_coef_rows(a, 0);
    __syncthreads();
    if (threadIdx.x == 0) { cyc[2] = clock64() - t1; cyc[3] = clock64() - t0; }
}

int main(int argc, char **argv) {
    const int T = argc > 1 ? atoi(argv[1]) : 150, S = argc > 2 ? atoi(argv[2]) : 40, W = S + 1;
    const size_t rows = (size_t)T * W, slack = 8 * W;
    std::vector<double2> lp(rows + 2 * slack, make_double2(-1.2, -7.1));
    std::vector<int2> band(T + 64, make_int2(0, S));
    std::vector<double> denom(rows, -7.0);
    std::vector<int> labels(S, 3);
    int hT = T, hS = S; int64_t rs[2] = {0, (int64_t)rows};
    K2Args a{};
    int *dT, *dS, *dl; int64_t *drs; int2 *dband; double2 *dlp; double *dden, *dal, *dbe, *dll; float4 *dco; float *dc; long long *dcyc;
    cudaMalloc(&dT, 4); cudaMalloc(&dS, 4); cudaMalloc(&dl, 4 * S); cudaMalloc(&drs, 16);
    cudaMalloc(&dband, band.size() * 8); cudaMalloc(&dlp, lp.size() * 16); cudaMalloc(&dden, rows * 8);
    cudaMalloc(&dal, rows * 8); cudaMalloc(&dbe, rows * 8); cudaMalloc(&dll, 32); cudaMalloc(&dco, rows * 16);
    cudaMalloc(&dc, 16); cudaMalloc(&dcyc, 64);
    cudaMemcpy(dT, &hT, 4, cudaMemcpyHostToDevice); cudaMemcpy(dS, &hS, 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dl, labels.data(), 4 * S, cudaMemcpyHostToDevice); cudaMemcpy(drs, rs, 16, cudaMemcpyHostToDevice);
    cudaMemcpy(dband, band.data(), band.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dlp, lp.data(), lp.size() * 16, cudaMemcpyHostToDevice);
    cudaMemcpy(dden, denom.data(), rows * 8, cudaMemcpyHostToDevice);
    a.T = dT; a.S = dS; a.labels = dl; a.row_start = drs; a.band = dband + 32; a.lp = dlp + slack; a.denom = dden;
    a.alpha = dal; a.beta = dbe; a.coef = dco; a.ll_fwd = dll; a.ll_bwd = dll + 1; a.costs = dc;
    a.T_max = T; a.S_max = S; a.V = 1000; a.blank = 0;
    for (int nb = 0; nb < 2; ++nb) {
        a.need_beta = nb;
        long long h[4] = {0, 0, 0, 0};
        for (int rep = 0; rep < 2; ++rep) {
            const size_t sm = k2_smem_bytes(S);
            a.chunk_frames = k2_chunk_frames(S);
            cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            cudaFuncSetAttribute(probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            if (W <= 32) probe<1><<<1, kK2Threads, sm>>>(a, dcyc);
            else if (W <= 64) probe<2><<<1, kK2Threads, sm>>>(a, dcyc);
            else probe<4><<<1, kK2Threads, sm>>>(a, dcyc);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(h, dcyc, 32, cudaMemcpyDeviceToHost);
        printf("T=%d S=%d need_beta=%d: alpha %.1f cyc/frame, beta %.1f cyc/frame, coef epilogue %lld cyc, total %lld cyc (%s)\n",
               T, S, nb, (double)h[0] / T, (double)h[1] / T, h[2], h[3], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
