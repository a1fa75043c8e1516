// K1 -- one pass over the packed logits: per-row log-softmax denominator plus the gather of the
// two log-probabilities the lattice needs (blank, and the label leaving state s).
//
// Replaces reduce_max + reduce_exp (reference include/reduce.h:79-139, gpu_rnnt.h:242-248: two full
// reads of the logits with 4-byte loads and two stream syncs) and the strided `log_p` gathers inside
// the alpha/beta kernels (gpu_rnnt_kernel.h:80-84,144-149).  CPU twin: cpu_rnnt.h:98-115.
//
//   denom[row] = -(max_v x + log sum_v exp(x - max))            (reference sign convention)
//   lp[row]    = (x[blank] + denom, x[label(b,s)] + denom)
//
// Streaming design (HBM-bound; algorithmic bytes = 4*V per live row, nothing for dead rows):
//   * persistent CTAs, one per SM; the flat row space is cut into tiles of G consecutive rows;
//   * a producer warp stages each tile's LIVE rows into a shared-memory ring with 1-D bulk async
//     copies (TMA engine, cp.async.bulk / SASS UBLKCP) that complete on an mbarrier;
//   * NW consumer warps take one row each: 128-bit shared loads, online max / sum-of-2^x in
//     registers (one MUFU.EX2 per element), a 10-shuffle warp combine, and one 24-byte result.
// A generic variant (direct global loads, any V / alignment) covers V % 4 != 0, unaligned bases and
// rows too large for the ring.
#pragma once

#include "common.cuh"

namespace mrnnt {

// Running log2-sum-exp2 of the values one lane has seen: sum_i 2^(x_i*log2e - mL), mL = max*log2e.
struct LaneLse {
    float m = kNegInfF;   // running max of x
    float mL = kNegInfF;  // m * log2e, rounded once; every term and the final result use this value
    float neg = 0.0f;          // -mL once a finite max has been seen, 0 before (keeps -inf inputs NaN-free)
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;

    __device__ __forceinline__ void raise(float cm) {
        if (cm > m) {
            const float nmL = cm * kLog2e;
            const float f = (m == kNegInfF) ? 0.0f : ex2_approx(mL - nmL);
            s0 *= f; s1 *= f; s2 *= f; s3 *= f;
            m = cm; mL = nmL; neg = -nmL;
        }
    }
    __device__ __forceinline__ void add(const float4 &a) {
        s0 += ex2_approx(fmaf(a.x, kLog2e, neg));
        s1 += ex2_approx(fmaf(a.y, kLog2e, neg));
        s2 += ex2_approx(fmaf(a.z, kLog2e, neg));
        s3 += ex2_approx(fmaf(a.w, kLog2e, neg));
    }
    __device__ __forceinline__ void add1(float x) { s0 += ex2_approx(fmaf(x, kLog2e, neg)); }

    // Warp-wide result: natural-log denominator  -(max + log sum exp(x - max)), in double.
    __device__ __forceinline__ double finish() const {
        const float ML = warp_max(mL);
        const float scale = (mL == kNegInfF) ? 0.0f : ex2_approx(mL - ML);
        const double mine = (static_cast<double>(s0) + static_cast<double>(s1) +
                             static_cast<double>(s2) + static_cast<double>(s3)) * static_cast<double>(scale);
        const double tot = warp_sum(mine);
        int e;
        const double f = frexp(tot, &e);  // tot = f * 2^e, f in [0.5, 1)
        const double lse2 = static_cast<double>(ML) + static_cast<double>(e - 1) +
                            static_cast<double>(log2f(static_cast<float>(f + f)));
        return -lse2 * kLn2D;
    }
};

__device__ __forceinline__ float max4(const float4 &a) { return fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)); }

// What lane 0 writes for one row.
__device__ __forceinline__ void k1_store_row(double2 *__restrict__ lp, double *__restrict__ denom, int64_t row,
                                             double den, float x_blank, float x_label, bool has_label) {
    denom[row] = den;
    lp[row] = make_double2(static_cast<double>(x_blank) + den,
                           has_label ? static_cast<double>(x_label) + den : kNegInf);
}

// ---------------------------------------------------------------------------------------------
// Generic variant: one warp per row, scalar global loads.  Any V, any alignment.
// ---------------------------------------------------------------------------------------------
constexpr int kGenericWarps = 8;

static __global__ void __launch_bounds__(kGenericWarps * kWarp)
    k1_lse_generic_kernel(const float *__restrict__ acts, const int *__restrict__ labels,
                          const int *__restrict__ rowmeta, double2 *__restrict__ lp, double *__restrict__ denom,
                          int64_t rows, int V, int blank) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * kGenericWarps + (threadIdx.x >> 5);
    const int64_t nwarps = static_cast<int64_t>(gridDim.x) * kGenericWarps;
    for (int64_t row = warp0; row < rows; row += nwarps) {
        const int meta = rowmeta[row];
        if (meta == kRowDead) {
            if (lane == 0) {
                denom[row] = 0.0;
                lp[row] = make_double2(0.0, 0.0);
            }
            continue;
        }
        const float *x = acts + row * V;
        LaneLse acc;
        for (int v = lane; v < V; v += kWarp) {
            const float xv = __ldg(x + v);
            acc.raise(xv);
            acc.add1(xv);
        }
        const double den = acc.finish();
        if (lane == 0) {
            const int lab = meta >= 0 ? __ldg(labels + meta) : -1;
            const bool has = lab >= 0 && lab < V;
            k1_store_row(lp, denom, row, den, __ldg(x + blank), has ? __ldg(x + lab) : 0.0f, has);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// TMA-staged variant.  Requirements (checked on the host): V % 4 == 0, acts 16-byte aligned,
// G*V*4 <= ring slot size.
// Shared memory: [stages][G*V] floats | full[stages] | empty[stages] | meta[stages][32]
// ---------------------------------------------------------------------------------------------
struct StreamTiling {
    int G = 0;       // rows per tile (1..32)
    int stages = 0;  // ring depth
    size_t smem_bytes = 0;
};

constexpr int kStreamMaxStages = 12;
constexpr size_t kStreamSmemBudget = 200 * 1024;  // of the 227 KB a CTA may use
constexpr int kStreamTileTarget = 32 * 1024;      // bytes per ring slot we aim for

// extra_per_row: additional per-row shared bytes a kernel keeps next to the tile (K3: its coefficients)
inline bool stream_tiling(int V, size_t extra_per_row, StreamTiling *out) {
    if (V <= 0 || (V % 4) != 0) return false;
    const size_t row_bytes = static_cast<size_t>(V) * 4;
    int G = static_cast<int>(kStreamTileTarget / row_bytes);
    if (G < 1) G = 1;
    if (G > 32) G = 32;
    const size_t slot = static_cast<size_t>(G) * row_bytes + 32 * (sizeof(int) + extra_per_row) + 16;
    int stages = static_cast<int>(kStreamSmemBudget / slot);
    if (stages < 3) return false;  // rows this large go through the generic kernels
    if (stages > kStreamMaxStages) stages = kStreamMaxStages;
    out->G = G;
    out->stages = stages;
    out->smem_bytes = static_cast<size_t>(stages) * slot + 128;
    return true;
}

// Issue one bulk copy per maximal run of live rows of a tile (called by one lane).
__device__ __forceinline__ void issue_live_runs(uint32_t mask, float *tile, const float *src_row0, int V,
                                                uint64_t *bar, uint64_t policy) {
    const uint32_t row_bytes = static_cast<uint32_t>(V) * 4u;
    while (mask) {
        const int r0 = __ffs(mask) - 1;
        const uint32_t inv = ~(mask >> r0);
        const int len = inv ? (__ffs(inv) - 1) : 32;
        bulk_g2s_hint(tile + static_cast<size_t>(r0) * V, src_row0 + static_cast<size_t>(r0) * V,
                      static_cast<uint32_t>(len) * row_bytes, bar, policy);
        mask = (len >= 32) ? 0u : (mask & ~(((1u << len) - 1u) << r0));
    }
}

template <int NW>
__global__ void __launch_bounds__((NW + 1) * kWarp, 1)
    k1_lse_tma_kernel(const float *__restrict__ acts, const int *__restrict__ labels,
                      const int *__restrict__ rowmeta, double2 *__restrict__ lp, double *__restrict__ denom,
                      int64_t rows, int V, int blank, int G, int stages) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const size_t tile_floats = static_cast<size_t>(G) * V;
    float *tiles = reinterpret_cast<float *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + static_cast<size_t>(stages) * tile_floats * 4);
    uint64_t *empty = full + stages;
    int *meta_sh = reinterpret_cast<int *>(empty + stages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, static_cast<uint32_t>(G));
        }
        mbar_init_fence();
    }
    __syncthreads();

    const int64_t ntiles = (rows + G - 1) / G;
    const int64_t nloc = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == NW) {
        // ---------------- producer warp ----------------
        const uint64_t policy = l2_policy_evict_first();
        for (int64_t k = 0; k < nloc; ++k) {
            const int stage = static_cast<int>(k % stages);
            const uint32_t phase = static_cast<uint32_t>((k / stages) & 1);
            const int64_t row0 = (blockIdx.x + k * gridDim.x) * G;
            int m = kRowDead;
            if (lane < G && row0 + lane < rows) m = __ldg(rowmeta + row0 + lane);
            const uint32_t mask = __ballot_sync(0xffffffffu, m != kRowDead);
            mbar_wait(empty + stage, phase ^ 1u);
            meta_sh[stage * 32 + lane] = m;
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_expect_tx(full + stage, static_cast<uint32_t>(__popc(mask)) * static_cast<uint32_t>(V) * 4u);
                issue_live_runs(mask, tiles + stage * tile_floats, acts + row0 * V, V, full + stage, policy);
            }
        }
    } else {
        // ---------------- consumer warps: one row at a time ----------------
        const int V4 = V >> 2;
        const int64_t nq = nloc * G;
        for (int64_t q = warp; q < nq; q += NW) {
            const int64_t k = q / G;
            const int r = static_cast<int>(q - k * G);
            const int stage = static_cast<int>(k % stages);
            const uint32_t phase = static_cast<uint32_t>((k / stages) & 1);
            const int64_t row = (blockIdx.x + k * gridDim.x) * G + r;
            mbar_wait(full + stage, phase);
            const int meta = meta_sh[stage * 32 + r];
            if (row < rows) {
                if (meta == kRowDead) {
                    if (lane == 0) {
                        denom[row] = 0.0;
                        lp[row] = make_double2(0.0, 0.0);
                    }
                } else {
                    const float *xrow = tiles + stage * tile_floats + static_cast<size_t>(r) * V;
                    const float4 *x4 = reinterpret_cast<const float4 *>(xrow);
                    int lab = -1;
                    if (lane == 0 && meta >= 0) lab = __ldg(labels + meta);  // latency hides under the row loop
                    LaneLse acc;
                    const float4 ninf = make_float4(kNegInfF, kNegInfF, kNegInfF, kNegInfF);
                    for (int j = lane; j < V4; j += 2 * kWarp) {
                        const float4 a = x4[j];
                        const float4 b = (j + kWarp < V4) ? x4[j + kWarp] : ninf;
                        acc.raise(fmaxf(max4(a), max4(b)));
                        acc.add(a);
                        acc.add(b);
                    }
                    const double den = acc.finish();
                    if (lane == 0) {
                        const bool has = lab >= 0 && lab < V;
                        k1_store_row(lp, denom, row, den, xrow[blank], has ? xrow[lab] : 0.0f, has);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + stage);
        }
    }
}

}  // namespace mrnnt
