// Development aid: where the time of the REAL lattice kernel (k2_lattice.cuh) goes, on synthetic weights.
//   k2_probe [T] [S] [B] [parts] [K] [zero_warps] [V]  -> event time of the kernel, clock64 stamps of the alpha pass of
//   utterance 0 (and of its CTA's first zero-fill warp)
#define MRNNT_K2_TRACE
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../include/mrnnt_b200/k2_lattice.cuh"
using namespace mrnnt;

static const RawRow *g_lp_host = nullptr;  // K1-style records; the kernel rewrites them in place, so every launch gets a fresh copy
static RawRow *g_lp_dev = nullptr;
static size_t g_lp_bytes = 0;

template <int K>
static float run(const K2Args &a, int B, size_t sm, int reps) {
    cudaFuncSetAttribute(k2_lattice_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    K2Args b = a;
    float best = 1e9f;
    for (int r = 0; r < reps; ++r) {
        b.epoch = a.epoch + r;
        cudaMemcpy(g_lp_dev, g_lp_host, g_lp_bytes, cudaMemcpyHostToDevice);
        cudaEventRecord(e0);
        b.phase_ctas = B * a.parts;
        k2_lattice_kernel<K><<<(a.zero_warps > 0 && a.need_beta && B * a.parts < 148) ? 148 : B * a.parts, kK2Threads, sm>>>(b);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char **argv) {
    const int T = argc > 1 ? atoi(argv[1]) : 150, S = argc > 2 ? atoi(argv[2]) : 40, W = S + 1;
    const int B = argc > 3 ? atoi(argv[3]) : 32, parts = argc > 4 ? atoi(argv[4]) : 4;
    const size_t rows1 = (size_t)T * W, rows = rows1 * B, slack = 8 * W;
    RawRow w0; w0.xb = 5.5f; w0.xl = 0.3f; w0.dh = 9.0f; w0.dl = 2.1f;
    std::vector<RawRow> lp(rows + 2 * slack, w0);
    std::vector<int2> band((size_t)B * T + 64, make_int2(0, S));
        std::vector<int> labels((size_t)B * S, 3), hT(B, T), hS(B, S);
    std::vector<int64_t> rs(B + 1);
    for (int b = 0; b <= B; ++b) rs[b] = (int64_t)b * rows1;
    K2Args a{};
    int *dT, *dS, *dl; int64_t *drs; int2 *dband; RawRow *dlp; Weight *dw; double *dll; Cell *dal, *dbe; float4 *dco; float *dc; unsigned *dfl;
    cudaMalloc(&dT, 4 * B); cudaMalloc(&dS, 4 * B); cudaMalloc(&dl, 4 * B * S); cudaMalloc(&drs, 8 * (B + 1));
    cudaMalloc(&dband, band.size() * 8); cudaMalloc(&dlp, lp.size() * 16);
    cudaMalloc(&dal, rows * 8); cudaMalloc(&dbe, rows * 8); cudaMalloc(&dll, 16 * B); cudaMalloc(&dco, rows * 16);
    cudaMalloc(&dc, 4 * B); cudaMalloc(&dfl, 4 * k2_flag_words(B)); cudaMemset(dfl, 0, 4 * k2_flag_words(B)); cudaMalloc(&dw, rows * 16);
    cudaMemcpy(dT, hT.data(), 4 * B, cudaMemcpyHostToDevice); cudaMemcpy(dS, hS.data(), 4 * B, cudaMemcpyHostToDevice);
    cudaMemcpy(dl, labels.data(), 4 * B * S, cudaMemcpyHostToDevice); cudaMemcpy(drs, rs.data(), 8 * (B + 1), cudaMemcpyHostToDevice);
    cudaMemcpy(dband, band.data(), band.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dlp, lp.data(), lp.size() * 16, cudaMemcpyHostToDevice);
    g_lp_host = lp.data(); g_lp_dev = dlp; g_lp_bytes = lp.size() * 16;
    a.T = dT; a.S = dS; a.labels = dl; a.row_start = drs; a.band = dband + 32; a.lp = dlp + slack; a.wts = dw;
    int *drl; cudaMalloc(&drl, rows * 4); a.rowlab = drl;
    a.alpha = dal; a.beta = dbe; a.coef = dco; a.ll_fwd = dll; a.ll_bwd = dll + B; a.costs = dc; a.costs_mapped = nullptr; a.flags = dfl;
    a.T_max = T; a.S_max = S; a.V = 1000; a.blank = 0;
    a.chunk_frames = k2_chunk_frames(W);
    const int K = argc > 5 ? atoi(argv[5]) : k2_states_per_lane(W);  // optional: force the states per lane
    a.row_warps = k2_row_warps(W, K);
    a.chunk_bufs = k2_chunk_bufs(a.row_warps);
    const size_t sm = k2_smem_bytes(W, a.row_warps);
    // zero fill: the plan's row flags of an unrestricted lattice, and a gradient buffer to fill
    const int zero_warps = argc > 6 ? atoi(argv[6]) : 0, V = argc > 7 ? atoi(argv[7]) : 1000;
    std::vector<int> meta(rows);
    size_t ndead = 0;
    for (int b = 0; b < B; ++b)
        for (int t = 0; t < T; ++t)
            for (int s2 = 0; s2 <= S; ++s2) {
                const bool live = t == 0 ? s2 == 0 : (s2 <= t && (S - s2) <= (T - t));
                meta[(size_t)b * rows1 + (size_t)t * W + s2] = live ? kRowNoLabel : kRowDead;
                ndead += !live;
            }
    int *dmeta; unsigned char *dgr = nullptr;
    cudaMalloc(&dmeta, rows * 4); cudaMemcpy(dmeta, meta.data(), rows * 4, cudaMemcpyHostToDevice);
    if (zero_warps > 0) cudaMalloc(&dgr, rows * (size_t)V * 4);
    a.rowmeta = dmeta; a.zero_dst = dgr; a.row_bytes = (unsigned)V * 4u; a.rows = (int64_t)rows; a.B = B;
    printf("dead rows: %zu of %zu = %.1f MB to zero\n", ndead, rows, ndead * (double)V * 4e-6);
    unsigned epoch = 1;
    for (int nb = 0; nb < 2; ++nb) {
        a.need_beta = nb; a.parts = nb ? parts : 1; a.epoch = epoch; epoch += 100;
        a.zero_warps = nb ? zero_warps : 0;
        float ms;
        if (K == 1) ms = run<1>(a, B, sm, 10);
        else if (K == 2) ms = run<2>(a, B, sm, 10);
        else ms = run<4>(a, B, sm, 10);
        long long h[64];
        cudaMemcpyFromSymbol(h, g_k2_trace, sizeof(h));
        double ll;
        cudaMemcpy(&ll, dll, 8, cudaMemcpyDeviceToHost);
        printf("T=%d S=%d B=%d parts=%d need_beta=%d K=%d row_warps=%d: kernel %.2f us (best of 10), ll=%.6f (%s)\n", T, S, B, a.parts, nb, K, a.row_warps,
               ms * 1e3, ll, cudaGetErrorString(cudaGetLastError()));
        const int nch = (T + a.chunk_frames - 1) / a.chunk_frames;
        printf("  alpha pass of utterance 0: setup %lld cyc, total %lld cyc = %.1f cyc/frame; per chunk (wait, run): ",
               h[1] - h[0], h[50] - h[0], (double)(h[50] - h[3]) / T);
        for (int c = 0; c < nch && c < 20; ++c)
            printf("(%lld, %lld) ", h[3 + 2 * c] - h[2 + 2 * c], (c + 1 < nch && c < 19 ? h[4 + 2 * c] : h[50]) - h[3 + 2 * c]);
        printf("\n  phase A (own share) %lld cyc, wait for the other parts %lld cyc, both passes done at %lld cyc, coefficient phase %lld cyc\n",
               h[60] - h[0], h[61] - h[60], h[51] - h[0], h[52] - h[51]);
        printf("  phase A inside (last call of CTA 0): start %lld, loads landed %lld, done %lld cyc after the kernel's start\n", h[55] - h[0], h[56] - h[0], h[57] - h[0]);
        if (a.zero_warps > 0) printf("  zero fill of CTA 0's first warp: from %lld to %lld cyc\n", h[53] - h[0], h[54] - h[0]);
    }
    return 0;
}
