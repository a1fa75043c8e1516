// Development aid: the dead-row zero fill (zero_fill.cuh) ALONE, on the row pattern of a named shape -- how far is it from
// a contiguous write of the same bytes, and why?  Variants:
//   counter   zero_dead_rows() as the lattice kernel runs it (units of 32 rows through a counter), W warps per CTA
//   contig    the same code on a plan whose dead rows are one contiguous block of the same size
//   runlist   a precomputed list of dead runs cut into equal BYTE shares per warp, every lane issuing 8 KB pieces
//   fill_probe [T] [S] [B] [V]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../include/mrnnt_b200/zero_fill.cuh"
using namespace mrnnt;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void counter_fill(ZeroFill z) {
    extern __shared__ __align__(128) unsigned char zb[];
    const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    zero_dead_rows(z, blockIdx.x * nw + warp, gridDim.x * nw, zb);
}

// runs: (first byte, bytes) sorted by address; pre[i] = bytes before run i; the grid's warps take equal byte shares
__global__ void runlist_fill(unsigned char *dst, const long long *run_off, const long long *run_len, const long long *pre,
                             int nruns, long long total) {
    extern __shared__ __align__(128) unsigned char zb[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = threadIdx.x * 16; i < kZeroFillBytes; i += blockDim.x * 16) *reinterpret_cast<uint4 *>(zb + i) = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const long long gw = (long long)blockIdx.x * nw + warp, ngw = (long long)gridDim.x * nw;
    // shares in whole 8 KB pieces
    const long long pieces = (total + kZeroFillBytes - 1) / kZeroFillBytes;
    const long long p0 = pieces * gw / ngw, p1 = pieces * (gw + 1) / ngw;
    long long b0 = p0 * kZeroFillBytes, b1 = std::min<long long>(p1 * kZeroFillBytes, total);
    if (b0 >= b1) return;
    // first run that ends behind b0 (binary search)
    int lo = 0, hi = nruns - 1;
    while (lo < hi) {
        const int mid = (lo + hi) / 2;
        if (pre[mid] + run_len[mid] > b0) hi = mid; else lo = mid + 1;
    }
    for (int r = lo; r < nruns && pre[r] < b1; ++r) {
        const long long s = std::max(b0, pre[r]) - pre[r], e = std::min(b1, pre[r] + run_len[r]) - pre[r];
        unsigned char *p = dst + run_off[r];
        for (long long o = s + (long long)lane * kZeroFillBytes; o < e; o += 32ll * kZeroFillBytes)
            bulk_s2g(p + o, zb, (uint32_t)std::min<long long>(kZeroFillBytes, e - o));
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// the same run list, but the 8 KB pieces of the compacted byte space are dealt out round-robin to ALL threads of the
// grid (piece p -> thread p mod nthreads): at any moment the whole GPU writes one compact window of the address space
__global__ void runlist_interleaved(unsigned char *dst, const long long *run_off, const long long *run_len, const long long *pre,
                                    int nruns, long long total, int piece_bytes) {
    extern __shared__ __align__(128) unsigned char zb[];
    for (int i = threadIdx.x * 16; i < kZeroFillBytes; i += blockDim.x * 16) *reinterpret_cast<uint4 *>(zb + i) = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    // thread order: lane-major across the grid, so that one warp instruction issues 32 adjacent pieces
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    const long long g = (long long)(threadIdx.x >> 5) * gridDim.x * 32 + (long long)blockIdx.x * 32 + (threadIdx.x & 31);
    const long long pieces = (total + piece_bytes - 1) / piece_bytes;
    int r = 0;
    for (long long p = g; p < pieces; p += nthreads) {
        long long b0 = p * piece_bytes;
        const long long b1 = std::min<long long>(b0 + piece_bytes, total);
        int lo = r, hi = nruns - 1;
        while (lo < hi) {
            const int mid = (lo + hi) / 2;
            if (pre[mid] + run_len[mid] > b0) hi = mid; else lo = mid + 1;
        }
        r = lo;
        while (b0 < b1) {
            const long long e = std::min(b1, pre[r] + run_len[r]);
            bulk_s2g(dst + run_off[r] + (b0 - pre[r]), zb, (uint32_t)(e - b0));
            b0 = e;
            if (b0 < b1) ++r;
        }
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// the run list again, written with ordinary 16-byte stores: piece p (of `piece_bytes`) -> warp p mod nwarps, the 32 lanes
// of the warp store it together (512 contiguous bytes per instruction)
__global__ void runlist_stg(unsigned char *dst, const long long *run_off, const long long *run_len, const long long *pre,
                            int nruns, long long total, int piece_bytes) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long gw = (long long)(threadIdx.x >> 5) * gridDim.x + blockIdx.x;
    const long long pieces = (total + piece_bytes - 1) / piece_bytes;
    int r = 0;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (long long p = gw; p < pieces; p += nwarps) {
        long long b0 = p * piece_bytes;
        const long long b1 = std::min<long long>(b0 + piece_bytes, total);
        int lo = r, hi = nruns - 1;
        while (lo < hi) {
            const int mid = (lo + hi) / 2;
            if (pre[mid] + run_len[mid] > b0) hi = mid; else lo = mid + 1;
        }
        r = lo;
        while (b0 < b1) {
            const long long e = std::min(b1, pre[r] + run_len[r]);
            unsigned char *q = dst + run_off[r] + (b0 - pre[r]);
            for (long long o = lane * 16; o < e - b0; o += 512) __stcs(reinterpret_cast<uint4 *>(q + o), z);
            b0 = e;
            if (b0 < b1) ++r;
        }
    }
}

template <class F> static float time_it(F f, int reps = 10) {
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
    const int T = argc > 1 ? atoi(argv[1]) : 150, S = argc > 2 ? atoi(argv[2]) : 40, B = argc > 3 ? atoi(argv[3]) : 32;
    const int V = argc > 4 ? atoi(argv[4]) : 1000, W = S + 1;
    const size_t rows1 = (size_t)T * W, rows = rows1 * B, row_bytes = (size_t)V * 4;
    std::vector<int> meta(rows), contig(rows, kRowNoLabel);
    size_t ndead = 0;
    for (int b = 0; b < B; ++b)
        for (int t = 0; t < T; ++t)
            for (int s = 0; s <= S; ++s) {
                const bool live = t == 0 ? s == 0 : (s <= t && (S - s) <= (T - t));
                meta[b * rows1 + (size_t)t * W + s] = live ? kRowNoLabel : kRowDead;
                ndead += !live;
            }
    for (size_t i = 0; i < ndead; ++i) contig[i] = kRowDead;
    std::vector<long long> off, len, pre;
    long long total = 0;
    for (size_t i = 0; i < rows;) {
        if (meta[i] != kRowDead) { ++i; continue; }
        size_t j = i;
        while (j < rows && meta[j] == kRowDead) ++j;
        off.push_back((long long)(i * row_bytes)); len.push_back((long long)((j - i) * row_bytes)); pre.push_back(total);
        total += (long long)((j - i) * row_bytes);
        i = j;
    }
    printf("T=%d S=%d B=%d V=%d: %zu dead rows of %zu in %zu runs = %.1f MB to zero (contiguous floor at 6.3 TB/s: %.1f us)\n", T, S, B, V, ndead,
           rows, off.size(), total * 1e-6, total / 6.3e6);
    int *dmeta, *dcontig; unsigned char *dst; unsigned *ctr; long long *doff, *dlen, *dpre;
    CK(cudaMalloc(&dmeta, rows * 4)); CK(cudaMalloc(&dcontig, rows * 4)); CK(cudaMalloc(&dst, rows * row_bytes)); CK(cudaMalloc(&ctr, 256));
    CK(cudaMalloc(&doff, off.size() * 8)); CK(cudaMalloc(&dlen, off.size() * 8)); CK(cudaMalloc(&dpre, off.size() * 8));
    CK(cudaMemcpy(dmeta, meta.data(), rows * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dcontig, contig.data(), rows * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(doff, off.data(), off.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dlen, len.data(), off.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dpre, pre.data(), off.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemset(ctr, 0, 256));
    ZeroFill z; z.dst = dst; z.rowmeta = dmeta; z.rows = (int64_t)rows; z.row_bytes = (unsigned)row_bytes; z.ctr = ctr;
    printf("cudaMemsetAsync of the same bytes             %7.1f us\n", 1e3 * time_it([&] { cudaMemsetAsync(dst, 0, total); }));
    for (int w : {1, 2, 4, 8}) {
        z.rowmeta = dmeta;
        const float a = time_it([&] { counter_fill<<<148, w * 32, kZeroFillBytes>>>(z); });
        z.rowmeta = dcontig;
        const float c = time_it([&] { counter_fill<<<148, w * 32, kZeroFillBytes>>>(z); });
        const float r = time_it([&] { runlist_fill<<<148, w * 32, kZeroFillBytes>>>(dst, doff, dlen, dpre, (int)off.size(), total); });
        printf("%d warps/CTA: counter %7.1f us (%6.0f GB/s)   counter on a contiguous block %7.1f us   run list, equal bytes %7.1f us (%6.0f GB/s)\n", w,
               a * 1e3, total / a * 1e-6, c * 1e3, r * 1e3, total / r * 1e-6);
    }
    for (int w : {1, 2, 4}) {
        for (int pb : {2048, 4096, 8192}) {
            const float r = time_it([&] { runlist_interleaved<<<148, w * 32, kZeroFillBytes>>>(dst, doff, dlen, dpre, (int)off.size(), total, pb); });
            printf("%d warps/CTA: run list, interleaved %4d B pieces %7.1f us (%6.0f GB/s)\n", w, pb, r * 1e3, total / r * 1e-6);
        }
    }
    {   // the same two kernels on ONE run of the same size (is it the scatter, or the way of writing?)
        long long one_off = 0, one_len = total, one_pre = 0, *d1;
        CK(cudaMalloc(&d1, 24));
        CK(cudaMemcpy(d1, &one_off, 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d1 + 1, &one_len, 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d1 + 2, &one_pre, 8, cudaMemcpyHostToDevice));
        for (int w : {2, 4}) {
            const float r = time_it([&] { runlist_interleaved<<<148, w * 32, kZeroFillBytes>>>(dst, d1, d1 + 1, d1 + 2, 1, total, 8192); });
            printf("%d warps/CTA: ONE run, interleaved 8192 B pieces %7.1f us (%6.0f GB/s)\n", w, r * 1e3, total / r * 1e-6);
        }
        for (int w : {2, 8, 24}) {
            for (int pb : {8192, 32768}) {
                const float r1 = time_it([&] { runlist_stg<<<148, w * 32>>>(dst, d1, d1 + 1, d1 + 2, 1, total, pb); });
                const float r2 = time_it([&] { runlist_stg<<<148, w * 32>>>(dst, doff, dlen, dpre, (int)off.size(), total, pb); });
                printf("%2d warps/CTA: 16-byte stores, %5d B pieces: ONE run %7.1f us (%6.0f GB/s)   the run list %7.1f us (%6.0f GB/s)\n", w, pb,
                       r1 * 1e3, total / r1 * 1e-6, r2 * 1e3, total / r2 * 1e-6);
            }
        }
    }
    CK(cudaDeviceSynchronize());
    {   // check of the interleaved variant
        CK(cudaMemset(dst, 0xff, rows * row_bytes));
        runlist_interleaved<<<148, 64, kZeroFillBytes>>>(dst, doff, dlen, dpre, (int)off.size(), total, 8192);
        CK(cudaDeviceSynchronize());
        std::vector<unsigned> hr(V);
        size_t bad2 = 0;
        for (size_t i = 0; i < rows; i += 97) {
            CK(cudaMemcpy(hr.data(), dst + i * row_bytes, row_bytes, cudaMemcpyDeviceToHost));
            const unsigned want = meta[i] == kRowDead ? 0u : 0xffffffffu;
            for (int v = 0; v < V; ++v) bad2 += hr[v] != want;
        }
        printf("interleaved check: %zu wrong words\n", bad2);
    }
    // the zeros really are everywhere they belong (run list variant ran last)
    CK(cudaMemset(dst, 0xff, rows * row_bytes));
    runlist_fill<<<148, 64, kZeroFillBytes>>>(dst, doff, dlen, dpre, (int)off.size(), total);
    CK(cudaDeviceSynchronize());
    std::vector<unsigned> hostrow(V);
    size_t bad = 0;
    for (size_t i = 0; i < rows; i += 97) {
        CK(cudaMemcpy(hostrow.data(), dst + i * row_bytes, row_bytes, cudaMemcpyDeviceToHost));
        const unsigned want = meta[i] == kRowDead ? 0u : 0xffffffffu;
        for (int v = 0; v < V; ++v) bad += hostrow[v] != want;
    }
    printf("run list check: %zu wrong words\n", bad);
    return 0;
}
