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
// boundary value.  The (lp_blank, lp_label) pairs and the band limits of future frames stream into a
// per-lane shared-memory FIFO with cp.async, so no global load sits on the dependent chain.  State is
// kept in double (|alpha| grows like T*log V; a float ulp there is already ~6e-5, the whole reason
// the reference's float path is only good to ~4e-4 on the gradients); the bounded
// correction term of each log-sum-exp is evaluated in float.
// After both passes the whole CTA folds alpha, beta, ll and the denominators into three float
// coefficients per row, so that the gradient kernel is a pure stream:
//   c0 = alpha(t-1,s) + beta(t,s)     - ll + denom      (x log2 e)
//   cb = alpha(t-1,s) + beta(t+1,s)   - ll + denom
//   cl = alpha(t-1,s) + beta(t+1,s+1) - ll + denom
#pragma once

#include "common.cuh"
#include "plan.cuh"

namespace mrnnt {

constexpr int kK2Threads = 256;

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
    int need_beta;  // 0: cost only (alpha pass), 1: alpha + beta + coefficients
};

template <int K>
struct K2Fifo {
    static constexpr int kDepth = (K <= 4) ? 8 : 4;   // frames in flight
    static constexpr int kStages = kDepth + 1;        // +1: never overwrite the frame just read
    static constexpr int kSlots = K + 2;              // K own states, 1 neighbour, 1 band
    static constexpr size_t kBytesPerWarp = static_cast<size_t>(kStages) * kSlots * kWarp * 16;
    static constexpr size_t kSmemBytes = 2 * kBytesPerWarp;
};

// ---- alpha pass: lane owns states s = lane*K .. lane*K+K-1 -----------------------------------------
template <int K>
__device__ __forceinline__ void k2_alpha_pass(const K2Args &a, int b, unsigned char *fifo) {
    using F = K2Fifo<K>;
    const int lane = threadIdx.x & 31;
    const int Tb = a.T[b], Sb = a.S[b];
    const int W = Sb + 1;
    const int64_t R = a.row_start[b];
    const double2 *lp = a.lp + R;
    const int2 *band = a.band + static_cast<size_t>(b) * a.T_max;
    double *alpha = a.alpha + R;
    const int s0 = lane * K;

    auto slot = [&](int stage, int j) -> unsigned char * {
        return fifo + (static_cast<size_t>(stage) * F::kSlots + j) * (kWarp * 16) + lane * 16;
    };
    auto prefetch = [&](int t) {
        if (t < Tb) {
            const int stage = t % F::kStages;
            const double2 *frame = lp + static_cast<int64_t>(t) * W;
#pragma unroll
            for (int j = 0; j < K; ++j) cp_async_16(slot(stage, j), frame + min(s0 + j, Sb));
            cp_async_16(slot(stage, K), frame + min(max(s0 - 1, 0), Sb));
            cp_async_8(slot(stage, K + 1), band + t);
        }
        cp_async_commit();
    };

    double st[K];
#pragma unroll
    for (int j = 0; j < K; ++j) st[j] = (s0 + j == 0) ? 0.0 : kNegInf;  // alpha(-1, .)

#pragma unroll
    for (int d = 0; d < F::kDepth; ++d) prefetch(d);

    for (int t = 0; t < Tb; ++t) {
        cp_async_wait<F::kDepth - 1>();
        const int stage = t % F::kStages;
        double lpb[K], lpl_prev[K];  // lp_blank(t, s_j), lp_label(t, s_j - 1)
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const double2 v = *reinterpret_cast<const double2 *>(slot(stage, j));
            lpb[j] = v.x;
            if (j + 1 < K) lpl_prev[j + 1] = v.y;
        }
        lpl_prev[0] = reinterpret_cast<const double2 *>(slot(stage, K))->y;
        const int2 lh = *reinterpret_cast<const int2 *>(slot(stage, K + 1));
        prefetch(t + F::kDepth);

        const int smin = max(lh.x, t - (Tb - 1 - Sb));
        const int smax = min(lh.y, t + 1);
        double up = __shfl_up_sync(0xffffffffu, st[K - 1], 1);
        if (lane == 0) up = kNegInf;  // alpha(t-1, -1)
        double nxt[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int s = s0 + j;
            const double below = (j == 0) ? up : st[j - 1];
            const double emit = (s > 0) ? below + lpl_prev[j] : kNegInf;
            const double stay = st[j] + lpb[j];
            const double v = lse_pair(emit, stay);
            nxt[j] = (s >= smin && s <= smax) ? v : kNegInf;
        }
#pragma unroll
        for (int j = 0; j < K; ++j) {
            st[j] = nxt[j];
            if (s0 + j <= Sb) alpha[static_cast<int64_t>(t) * W + s0 + j] = nxt[j];
        }
    }
    cp_async_wait<0>();
    // ll = alpha(T-1, S): held by the lane that owns state S
#pragma unroll
    for (int j = 0; j < K; ++j) {
        if (s0 + j == Sb) {
            a.ll_fwd[b] = st[j];
            a.costs[b] = static_cast<float>(-st[j]);
        }
    }
}

// ---- beta pass -------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void k2_beta_pass(const K2Args &a, int b, unsigned char *fifo) {
    using F = K2Fifo<K>;
    const int lane = threadIdx.x & 31;
    const int Tb = a.T[b], Sb = a.S[b];
    const int W = Sb + 1;
    const int64_t R = a.row_start[b];
    const double2 *lp = a.lp + R;
    const int2 *band = a.band + static_cast<size_t>(b) * a.T_max;
    double *beta = a.beta + R;
    const int s0 = lane * K;

    auto slot = [&](int stage, int j) -> unsigned char * {
        return fifo + (static_cast<size_t>(stage) * F::kSlots + j) * (kWarp * 16) + lane * 16;
    };
    // step i handles frame t = Tb-1-i
    auto prefetch = [&](int i) {
        if (i < Tb) {
            const int t = Tb - 1 - i;
            const int stage = i % F::kStages;
            const double2 *frame = lp + static_cast<int64_t>(t) * W;
#pragma unroll
            for (int j = 0; j < K; ++j) cp_async_16(slot(stage, j), frame + min(s0 + j, Sb));
            cp_async_8(slot(stage, K + 1), band + max(t - 1, 0));
        }
        cp_async_commit();
    };

    double st[K];
#pragma unroll
    for (int j = 0; j < K; ++j) st[j] = (s0 + j == Sb) ? 0.0 : kNegInf;  // beta(T, .)

#pragma unroll
    for (int d = 0; d < F::kDepth; ++d) prefetch(d);

    for (int i = 0; i < Tb; ++i) {
        const int t = Tb - 1 - i;
        cp_async_wait<F::kDepth - 1>();
        const int stage = i % F::kStages;
        double lpb[K], lpl[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const double2 v = *reinterpret_cast<const double2 *>(slot(stage, j));
            lpb[j] = v.x;
            lpl[j] = v.y;
        }
        const int2 lh = *reinterpret_cast<const int2 *>(slot(stage, K + 1));
        prefetch(i + F::kDepth);

        const int smin = (t == 0) ? 0 : max(lh.x, t - (Tb - Sb));
        const int smax = (t == 0) ? 0 : min(lh.y, t);
        double dn = __shfl_down_sync(0xffffffffu, st[0], 1);
        if (lane == 31) dn = kNegInf;
        double nxt[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int s = s0 + j;
            const double above = (j == K - 1) ? dn : st[j + 1];
            const double emit = (s < Sb) ? above + lpl[j] : kNegInf;  // beta(t+1, S+1) = -inf
            const double stay = st[j] + lpb[j];
            const double v = lse_pair(emit, stay);
            nxt[j] = (s >= smin && s <= smax) ? v : kNegInf;
        }
#pragma unroll
        for (int j = 0; j < K; ++j) {
            st[j] = nxt[j];
            if (s0 + j <= Sb) beta[static_cast<int64_t>(t) * W + s0 + j] = nxt[j];
        }
    }
    cp_async_wait<0>();
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
    const int n = Tb * W;
    const float qnan = __int_as_float(0x7fc00000);
    for (int i = threadIdx.x; i < n; i += kK2Threads) {
        const int t = i / W;
        const int s = i - t * W;
        int lab = -1;
        if (s < Sb) {
            lab = labels[s];
            if (lab == a.blank || lab < 0 || lab >= a.V) lab = -1;  // blank branch wins (cpu_rnnt.h:224-232)
        }
        float4 c;
        c.w = __int_as_float(lab);
        if (!(ll > kNegInf)) {
            // infeasible utterance (e.g. the alignment band excludes the terminal state): cost = +inf and,
            // as in the reference, no finite gradient exists.  NaN is written on purpose.
            c.x = c.y = c.z = qnan;
        } else {
            const double al = (t == 0) ? (s == 0 ? 0.0 : kNegInf) : alpha[i - W];
            const bool last = (t == Tb - 1);
            const double b0 = beta[i];
            const double b1 = last ? (s == Sb ? 0.0 : kNegInf) : beta[i + W];
            const double b2 = (s == Sb) ? kNegInf : (last ? (s + 1 == Sb ? 0.0 : kNegInf) : beta[i + W + 1]);
            if (al == kNegInf) {
                c.x = c.y = c.z = kNegInfF;
            } else {
                const double base = al - ll + denom[i];
                c.x = static_cast<float>((base + b0) * kLog2eD);
                c.y = static_cast<float>((base + b1) * kLog2eD);
                c.z = static_cast<float>((base + b2) * kLog2eD);
            }
        }
        coef[i] = c;
    }
}

template <int K>
__global__ void __launch_bounds__(kK2Threads) k2_lattice_kernel(K2Args a, int b_begin) {
    extern __shared__ __align__(128) unsigned char k2_smem[];
    const int b = b_begin + blockIdx.x;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        k2_alpha_pass<K>(a, b, k2_smem);
    } else if (warp == 1 && a.need_beta) {
        k2_beta_pass<K>(a, b, k2_smem + K2Fifo<K>::kBytesPerWarp);
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
            const double v = (s >= smin && s <= smax) ? lse_pair(emit, stay) : kNegInf;
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
            const double v = (s >= smin && s <= smax) ? lse_pair(emit, stay) : kNegInf;
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
