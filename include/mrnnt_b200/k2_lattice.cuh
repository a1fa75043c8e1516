// K2 -- forward (alpha) and backward (beta) lattice recursions and the per-row gradient
// coefficients.
//
// Replaces compute_alphas_kernel / compute_betas_kernel (reference include/gpu_rnnt_kernel.h:121-161,
// :197-237: one thread per label state, __syncthreads and >= 6 dependent global loads per frame,
// alpha then beta serialised on one stream) and the per-element lattice lookups of
// compute_grad_kernel (:273-284).  CPU twin: cpu_rnnt.h:155-214, accessors
// cpu_workspace_manager.h:67-86,161-205.
//
//   alpha(t,s) = lse( alpha(t-1,s-1) + lp_label(t,s-1),  alpha(t-1,s) + lp_blank(t,s) )
//   beta (t,s) = lse( beta(t+1,s+1)  + lp_label(t,s),    beta(t+1,s)  + lp_blank(t,s) )
//   cells outside [alpha_s_min, alpha_s_max] / [beta_s_min, beta_s_max] are -inf.
//
// Design: one CTA per utterance.  Warp 0 runs alpha, warp 1 runs beta, concurrently.  A lane owns K
// consecutive label states in registers; the only cross-lane traffic per frame is ONE shuffle of the
// boundary value.  The recursion is a pure latency chain on one in-order warp, so everything else is
// kept out of its instruction stream: the (lp_blank, lp_label) pairs of a whole CHUNK of frames (they
// are contiguous in memory) arrive in shared memory through one bulk async copy (TMA engine) per chunk,
// completing on an mbarrier that the warp polls once per chunk; band limits are fetched a chunk ahead,
// one frame per lane, and broadcast by shuffle.  State is kept in double (|alpha| grows like T*log V; a
// float ulp there is already ~6e-5, the whole reason the reference's float path is only good to ~4e-4
// on the gradients); the bounded correction term of each log-sum-exp is evaluated in float.
// After both passes the whole CTA folds alpha, beta, ll and the denominators into three float
// coefficients per row, so that the gradient kernel is a pure stream:
//   c0 = alpha(t-1,s) + beta(t,s)     - ll + denom      (x log2 e)
//   cb = alpha(t-1,s) + beta(t+1,s)   - ll + denom
//   cl = alpha(t-1,s) + beta(t+1,s+1) - ll + denom
#pragma once

#include "common.cuh"
#include "plan.cuh"

namespace mrnnt {

constexpr int kK2Threads = 512;
constexpr int kK2ChunkBufs = 3;             // chunk buffers per direction (two in flight while one is consumed)
constexpr int kK2MaxChunkFrames = 16;       // <= 32: one band entry per lane
constexpr int kK2ChunkTargetBytes = 12288;  // bytes of lp per chunk we aim for

struct K2Args {
    const int *T;
    const int *S;
    const int *labels;
    const int64_t *row_start;
    const int2 *band;
    const double2 *lp;
    const double *denom;
    double *alpha;
    double *beta;
    float4 *coef;
    double *ll_fwd;
    double *ll_bwd;
    float *costs;
    int T_max;
    int S_max;
    int V;
    int blank;
    int need_beta;     // 0: cost only (alpha pass), 1: alpha + beta + coefficients
    int chunk_frames;  // frames per bulk-copied chunk (host: k2_chunk_frames)
};

// frames per chunk for a launch whose widest utterance has S_max + 1 states
inline int k2_chunk_frames(int S_max) {
    const int frame_bytes = (S_max + 1) * static_cast<int>(sizeof(double2));
    int ch = kK2ChunkTargetBytes / frame_bytes;
    if (ch < 1) ch = 1;
    if (ch > kK2MaxChunkFrames) ch = kK2MaxChunkFrames;
    return ch;
}
// dynamic shared memory of k2_lattice_kernel: per direction kK2ChunkBufs chunk buffers + their mbarriers
inline size_t k2_smem_bytes(int S_max) {
    const size_t chunk = static_cast<size_t>(k2_chunk_frames(S_max)) * (S_max + 1) * sizeof(double2);
    return 2 * (kK2ChunkBufs * chunk + 64);
}

// One direction's chunk ring.  `base` is 16-byte aligned shared memory of kK2ChunkBufs*chunk_bytes + 64.
struct K2Ring {
    unsigned char *buf;
    uint64_t *full;
    size_t chunk_bytes;
    __device__ __forceinline__ K2Ring(unsigned char *base, size_t chunk_bytes_)
        : buf(base), full(reinterpret_cast<uint64_t *>(base + kK2ChunkBufs * chunk_bytes_)), chunk_bytes(chunk_bytes_) {}
    __device__ __forceinline__ void init() {  // one lane
        for (int i = 0; i < kK2ChunkBufs; ++i) mbar_init(full + i, 1);
        mbar_init_fence();
    }
    __device__ __forceinline__ const double2 *slot(int i) const {
        return reinterpret_cast<const double2 *>(buf + i * chunk_bytes);
    }
    // one lane: copy `frames` frames of W states starting at src into slot i
    __device__ __forceinline__ void fill(int i, const double2 *src, int frames, int W) {
        const uint32_t bytes = static_cast<uint32_t>(frames) * W * sizeof(double2);
        mbar_arrive_expect_tx(full + i, bytes);
        bulk_g2s(buf + i * chunk_bytes, src, bytes, full + i);
    }
};

__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- alpha pass: lane owns states s = lane*K .. lane*K+K-1 -----------------------------------------
template <int K>
__device__ __forceinline__ void k2_alpha_pass(const K2Args &a, int b, unsigned char *smem) {
    const int lane = threadIdx.x & 31;
    const int Tb = a.T[b], Sb = a.S[b];
    const int W = Sb + 1;
    const int CH = a.chunk_frames;
    const int64_t R = a.row_start[b];
    const double2 *lp = a.lp + R;
    const int2 *band = a.band + static_cast<size_t>(b) * a.T_max;
    const int s0 = lane * K;
    K2Ring ring(smem, static_cast<size_t>(CH) * (a.S_max + 1) * sizeof(double2));
    const int nchunks = (Tb + CH - 1) / CH;

    if (lane == 0) {
        ring.init();
        for (int c = 0; c < kK2ChunkBufs && c < nchunks; ++c) ring.fill(c, lp + static_cast<int64_t>(c) * CH * W, min(CH, Tb - c * CH), W);
    }
    __syncwarp();

    // clamped columns this lane reads in every frame
    int col[K + 1];
#pragma unroll
    for (int j = 0; j < K; ++j) col[j] = min(s0 + j, Sb);
    col[K] = min(max(s0 - 1, 0), Sb);

    double st[K];
#pragma unroll
    for (int j = 0; j < K; ++j) st[j] = (s0 + j == 0) ? 0.0 : kNegInf;  // alpha(-1, .)

    double *out = a.alpha + R + s0;
    const int reach = Tb - 1 - Sb;  // alpha_s_min(t) = max(lo[t], t - reach)
    int2 lh_next = (lane < min(CH, Tb)) ? band[lane] : make_int2(0, 0);
    int slot = 0;
    uint32_t phase = 0;
    for (int c = 0; c < nchunks; ++c) {
        const int t0 = c * CH;
        const int nf = min(CH, Tb - t0);
        const int2 lh_cur = lh_next;
        {   // band limits of the NEXT chunk, one frame per lane; latency hides behind this chunk
            const int tn = t0 + CH + lane;
            lh_next = (lane < CH && tn < Tb) ? band[tn] : make_int2(0, 0);
        }
        mbar_wait(ring.full + slot, phase);
        const double2 *frame = ring.slot(slot);
        for (int f = 0; f < nf; ++f, frame += W) {
            const int t = t0 + f;
            double lpb[K], lpl_prev[K];  // lp_blank(t, s_j), lp_label(t, s_j - 1)
#pragma unroll
            for (int j = 0; j < K; ++j) {
#ifdef MRNNT_X_NOLDS
                const double2 v = make_double2(-1.2 - 1e-9 * t, -7.1);
#else
                const double2 v = frame[col[j]];
#endif
                lpb[j] = v.x;
                if (j + 1 < K) lpl_prev[j + 1] = v.y;
            }
#ifdef MRNNT_X_NOLDS
            lpl_prev[0] = -7.1;
#else
            lpl_prev[0] = frame[col[K]].y;
#endif
#ifdef MRNNT_X_NOBAND
            const int lo = 0, hi = Sb; (void)lh_cur;
#else
            const int lo = __shfl_sync(0xffffffffu, lh_cur.x, f);
            const int hi = __shfl_sync(0xffffffffu, lh_cur.y, f);
#endif
            const int smin = max(lo, t - reach);
            const int smax = min(hi, t + 1);
            double up = __shfl_up_sync(0xffffffffu, st[K - 1], 1);
            if (lane == 0) up = kNegInf;  // alpha(t-1, -1)
            double nxt[K];
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const int s = s0 + j;
                const double below = (j == 0) ? up : st[j - 1];
                const double emit = below + lpl_prev[j];  // s == 0: `below` is -inf and so is the sum
                const double stay = st[j] + lpb[j];
                nxt[j] = lse_pair_masked(emit, stay, s >= smin && s <= smax);
            }
#pragma unroll
            for (int j = 0; j < K; ++j) {
                st[j] = nxt[j];
#ifndef MRNNT_X_NOSTORE
                if (s0 + j <= Sb) out[j] = nxt[j];
#endif
            }
            out += W;
        }
        // refill this slot with the chunk kK2ChunkBufs ahead
        __syncwarp();
        if (lane == 0 && c + kK2ChunkBufs < nchunks) {
            fence_proxy_async_smem();
            const int cn = c + kK2ChunkBufs;
            ring.fill(slot, lp + static_cast<int64_t>(cn) * CH * W, min(CH, Tb - cn * CH), W);
        }
        if (++slot == kK2ChunkBufs) {
            slot = 0;
            phase ^= 1u;
        }
    }
    // ll = alpha(T-1, S): held by the lane that owns state S
#pragma unroll
    for (int j = 0; j < K; ++j) {
        if (s0 + j == Sb) {
            a.ll_fwd[b] = st[j];
            a.costs[b] = static_cast<float>(-st[j]);
        }
    }
}

// ---- beta pass: frames in descending order; chunk c covers frames [Tb-(c+1)*CH, Tb-c*CH) ------------
template <int K>
__device__ __forceinline__ void k2_beta_pass(const K2Args &a, int b, unsigned char *smem) {
    const int lane = threadIdx.x & 31;
    const int Tb = a.T[b], Sb = a.S[b];
    const int W = Sb + 1;
    const int CH = a.chunk_frames;
    const int64_t R = a.row_start[b];
    const double2 *lp = a.lp + R;
    const int2 *band = a.band + static_cast<size_t>(b) * a.T_max;
    const int s0 = lane * K;
    K2Ring ring(smem, static_cast<size_t>(CH) * (a.S_max + 1) * sizeof(double2));
    const int nchunks = (Tb + CH - 1) / CH;

    auto chunk_lo = [&](int c) { return max(Tb - (c + 1) * CH, 0); };
    auto chunk_hi = [&](int c) { return Tb - c * CH; };  // exclusive
    if (lane == 0) {
        ring.init();
        for (int c = 0; c < kK2ChunkBufs && c < nchunks; ++c)
            ring.fill(c, lp + static_cast<int64_t>(chunk_lo(c)) * W, chunk_hi(c) - chunk_lo(c), W);
    }
    __syncwarp();

    int col[K];
#pragma unroll
    for (int j = 0; j < K; ++j) col[j] = min(s0 + j, Sb);

    double st[K];
#pragma unroll
    for (int j = 0; j < K; ++j) st[j] = (s0 + j == Sb) ? 0.0 : kNegInf;  // beta(T, .)

    const int reach = Tb - Sb;  // beta_s_min(t) = max(lo[t-1], t - reach)
    // lane l of a chunk holds band[tlo - 1 + l] = the limits frame t = tlo + l needs (unused at t == 0)
    auto load_band = [&](int c) {
        const int tlo = chunk_lo(c);
        const int idx = tlo - 1 + lane;
        return (c < nchunks && lane < CH && idx >= 0 && idx < Tb) ? band[idx] : make_int2(0, 0);
    };
    int2 lh_next = load_band(0);
    int slot = 0;
    uint32_t phase = 0;
    for (int c = 0; c < nchunks; ++c) {
        const int tlo = chunk_lo(c);
        const int nf = chunk_hi(c) - tlo;
        const int2 lh_cur = lh_next;
        lh_next = load_band(c + 1);
        mbar_wait(ring.full + slot, phase);
        const double2 *frame = ring.slot(slot) + static_cast<size_t>(nf - 1) * W;
        double *out = a.beta + R + static_cast<int64_t>(tlo + nf - 1) * W + s0;
        for (int f = nf - 1; f >= 0; --f, frame -= W, out -= W) {
            const int t = tlo + f;
            double lpb[K], lpl[K];
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const double2 v = frame[col[j]];
                lpb[j] = v.x;
                lpl[j] = v.y;
            }
            const int lo = __shfl_sync(0xffffffffu, lh_cur.x, f);
            const int hi = __shfl_sync(0xffffffffu, lh_cur.y, f);
            const int smin = (t == 0) ? 0 : max(lo, t - reach);
            const int smax = (t == 0) ? 0 : min(hi, t);
            double dn = __shfl_down_sync(0xffffffffu, st[0], 1);
            if (lane == 31) dn = kNegInf;
            double nxt[K];
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const int s = s0 + j;
                const double above = (j == K - 1) ? dn : st[j + 1];
                const double emit = (s < Sb) ? above + lpl[j] : kNegInf;  // beta(t+1, S+1) = -inf
                const double stay = st[j] + lpb[j];
                nxt[j] = lse_pair_masked(emit, stay, s >= smin && s <= smax);
            }
#pragma unroll
            for (int j = 0; j < K; ++j) {
                st[j] = nxt[j];
                if (s0 + j <= Sb) out[j] = nxt[j];
            }
        }
        __syncwarp();
        if (lane == 0 && c + kK2ChunkBufs < nchunks) {
            fence_proxy_async_smem();
            const int cn = c + kK2ChunkBufs;
            ring.fill(slot, lp + static_cast<int64_t>(chunk_lo(cn)) * W, chunk_hi(cn) - chunk_lo(cn), W);
        }
        if (++slot == kK2ChunkBufs) {
            slot = 0;
            phase ^= 1u;
        }
    }
    if (lane == 0) a.ll_bwd[b] = st[0];  // beta(0, 0)
}

// ---- per-row gradient coefficients (whole CTA) -----------------------------------------------------
// Semantics of the lookups follow the reference accessors (gpu_rnnt_kernel.h:10-56): alpha(-1,0)=0,
// alpha(-1,s>0)=-inf, beta(T,S)=0, beta(T,s<S)=-inf, beta(.,S+1)=-inf; everything else comes from
// the stored grids, which already hold -inf outside the band.
__device__ __forceinline__ void k2_coef_rows(const K2Args &a, int b) {
    const int Tb = a.T[b], Sb = a.S[b];
    const int W = Sb + 1;
    const int64_t R = a.row_start[b];
    const double *alpha = a.alpha + R;
    const double *beta = a.beta + R;
    const double *denom = a.denom + R;
    const int *labels = a.labels + static_cast<size_t>(b) * a.S_max;
    float4 *coef = a.coef + R;
    const double ll = a.ll_fwd[b];
    const bool feasible = ll > kNegInf;
    const int n = Tb * W;
    const float qnan = __int_as_float(0x7fc00000);
    constexpr int U = 4;  // rows per thread per batch: 5*U independent loads in flight before any use
    for (int base = threadIdx.x; base < n; base += U * kK2Threads) {
        double al[U], b0[U], b1[U], b2[U], dn[U];
        int lab[U], tt[U], ss[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = min(base + u * kK2Threads, n - 1);  // clamped: loads are unconditional, stores are not
            const int t = i / W;
            const int s = i - t * W;
            tt[u] = t;
            ss[u] = s;
            al[u] = alpha[max(i - W, 0)];
            b0[u] = beta[i];
            b1[u] = beta[min(i + W, n - 1)];
            b2[u] = beta[min(i + W + 1, n - 1)];
            dn[u] = denom[i];
            lab[u] = (Sb > 0) ? labels[min(s, Sb - 1)] : -1;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = base + u * kK2Threads;
            if (i >= n) break;
            const int t = tt[u], s = ss[u];
            const bool last = (t == Tb - 1);
            const double av = (t == 0) ? (s == 0 ? 0.0 : kNegInf) : al[u];
            const double v1 = last ? (s == Sb ? 0.0 : kNegInf) : b1[u];
            const double v2 = (s == Sb) ? kNegInf : (last ? (s + 1 == Sb ? 0.0 : kNegInf) : b2[u]);
            int lb = lab[u];
            if (s >= Sb || lb == a.blank || lb < 0 || lb >= a.V) lb = -1;  // blank branch wins (cpu_rnnt.h:224-232)
            float4 c;
            c.w = __int_as_float(lb);
            if (!feasible) {
                // infeasible utterance (e.g. the alignment band excludes the terminal state): cost = +inf and,
                // as in the reference, no finite gradient exists.  NaN is written on purpose.
                c.x = c.y = c.z = qnan;
            } else if (av == kNegInf) {
                c.x = c.y = c.z = kNegInfF;
            } else {
                const double bs = av - ll + dn[u];
                c.x = static_cast<float>((bs + b0[u]) * kLog2eD);
                c.y = static_cast<float>((bs + v1) * kLog2eD);
                c.z = static_cast<float>((bs + v2) * kLog2eD);
            }
            coef[i] = c;
        }
    }
}

template <int K>
static __global__ void __launch_bounds__(kK2Threads) k2_lattice_kernel(K2Args a, int b_begin) {
    extern __shared__ __align__(128) unsigned char k2_smem[];
    const int b = b_begin + blockIdx.x;
    const int warp = threadIdx.x >> 5;
    const size_t dir_bytes = kK2ChunkBufs * static_cast<size_t>(a.chunk_frames) * (a.S_max + 1) * sizeof(double2) + 64;
    if (warp == 0) {
        k2_alpha_pass<K>(a, b, k2_smem);
    } else if (warp == 1 && a.need_beta) {
        k2_beta_pass<K>(a, b, k2_smem + dir_bytes);
    }
    if (a.need_beta) {
        __syncthreads();
        k2_coef_rows(a, b);
    }
}

// ---- fallback for very long label sequences (S_max + 1 > 32 * 16) ----------------------------------
// One CTA per utterance; states strided over the threads, previous frame in shared memory, one
// __syncthreads per frame.  Alpha first, then beta.  Slow path, same arithmetic.
static __global__ void __launch_bounds__(kK2Threads) k2_lattice_wide_kernel(K2Args a, int b_begin) {
    extern __shared__ __align__(16) unsigned char k2w_smem[];
    double *prev = reinterpret_cast<double *>(k2w_smem);  // [S_max + 2]
    const int b = b_begin + blockIdx.x;
    const int Tb = a.T[b], Sb = a.S[b];
    const int W = Sb + 1;
    const int64_t R = a.row_start[b];
    const double2 *lp = a.lp + R;
    const int2 *band = a.band + static_cast<size_t>(b) * a.T_max;
    double *alpha = a.alpha + R;
    double *beta = a.beta + R;
    const int tid = threadIdx.x;

    for (int s = tid; s <= Sb + 1; s += kK2Threads) prev[s] = (s == 0) ? 0.0 : kNegInf;
    __syncthreads();
    for (int t = 0; t < Tb; ++t) {
        const int2 lh = band[t];
        const int smin = max(lh.x, t - (Tb - 1 - Sb));
        const int smax = min(lh.y, t + 1);
        const double2 *frame = lp + static_cast<int64_t>(t) * W;
        for (int s = tid; s <= Sb; s += kK2Threads) {
            const double emit = (s > 0) ? prev[s - 1] + frame[s - 1].y : kNegInf;
            const double stay = prev[s] + frame[s].x;
            const double v = (s >= smin && s <= smax) ? lse_pair_fast(emit, stay) : kNegInf;
            alpha[static_cast<int64_t>(t) * W + s] = v;
        }
        __syncthreads();
        for (int s = tid; s <= Sb; s += kK2Threads) prev[s] = alpha[static_cast<int64_t>(t) * W + s];
        __syncthreads();
    }
    if (tid == 0) {
        a.ll_fwd[b] = prev[Sb];
        a.costs[b] = static_cast<float>(-prev[Sb]);
    }
    if (!a.need_beta) return;
    __syncthreads();
    for (int s = tid; s <= Sb + 1; s += kK2Threads) prev[s] = (s == Sb) ? 0.0 : kNegInf;
    __syncthreads();
    for (int t = Tb - 1; t >= 0; --t) {
        const int2 lh = band[max(t - 1, 0)];
        const int smin = (t == 0) ? 0 : max(lh.x, t - (Tb - Sb));
        const int smax = (t == 0) ? 0 : min(lh.y, t);
        const double2 *frame = lp + static_cast<int64_t>(t) * W;
        for (int s = tid; s <= Sb; s += kK2Threads) {
            const double2 f = frame[s];
            const double emit = (s < Sb) ? prev[s + 1] + f.y : kNegInf;
            const double stay = prev[s] + f.x;
            const double v = (s >= smin && s <= smax) ? lse_pair_fast(emit, stay) : kNegInf;
            beta[static_cast<int64_t>(t) * W + s] = v;
        }
        __syncthreads();
        for (int s = tid; s <= Sb; s += kK2Threads) prev[s] = beta[static_cast<int64_t>(t) * W + s];
        __syncthreads();
    }
    if (tid == 0) a.ll_bwd[b] = prev[0];
    __syncthreads();
    k2_coef_rows(a, b);
}

}  // namespace mrnnt
