// Workspace layout and the small set-up kernels (row offsets, alignment band, per-row liveness).
//
// Replaces the host-side carve-up and the ~20 blocking cudaMemcpy calls of the reference's
// GpuRNNTWorkspaceManager (include/gpu_workspace_manager.h:228-329, :191-219).  Differences that
// matter: offsets are 64-bit (the reference's int offsets overflow above 2^31-1 logits),
// nothing here synchronises the host, and the alignment band is built on the device.
//
// Layout of the packed inputs (unchanged from the reference, cpu_workspace_manager.h:46-49,117-135):
//   acts   [rows, V]   rows = sum_b T_b*(S_b+1); row of (b,t,s) = row_start[b] + t*(S_b+1) + s
//   labels [B, S_max]  S_max = max_b S_b
//   alignment [B, T_max], T_max = max_b T_b   (or [B, align_stride] with align_stride >= T_max, an extension)
#pragma once

#include <cstddef>
#include <cstdint>

#include "common.cuh"

namespace mrnnt {

// Host-side facts about one batch, derived once from the host copies of T[] and S[].
struct Shape {
    int B = 0;
    int V = 0;
    int T_max = 0;
    int S_max = 0;
    int64_t rows = 0;  // rows of acts / gradients: sum_b T_b * (S_b + 1) packed, B * T_dim * U padded
    // Padded layout: acts is the joint network's own [B, T_dim, U, V] tensor, row of (b,t,s) =
    // (b*T_dim + t)*U + s; rows with t >= T_b or s > S_b are dead (never read, gradient zero).  0 = packed.
    int T_dim = 0;
    int U = 0;
    int label_stride = 0;  // ints per utterance in labels[] (packed reference layout: max_b S_b)
    int width() const { return U > 0 ? U : S_max + 1; }  // rows per frame at most
};

// Device arrays carved out of the caller's workspace buffer (all 256-byte aligned).
struct Workspace {
    int64_t *row_start = nullptr;  // [B+1]   first row of each utterance
    int2 *band = nullptr;          // [B*T_max] (min_allowed_s, max_allowed_s) per frame
    int *rowmeta = nullptr;        // [rows]  >=0: index into labels[], kRowNoLabel, kRowDead
    int *rowutt = nullptr;         // [rows]  utterance of the row (the gradient kernel's per-utterance scale)
    RawRow *lp = nullptr;          // [rows]  (x[blank], x[label_s], max, sum) per live row from K1; K2 turns (max, sum)
                                   //         into -log2 sum_v exp x[v] in place
    Weight *wts = nullptr;         // [rows]  transition weights (m * 2^e pairs, band folded in), K2 phase A
    Cell *alpha = nullptr;         // [rows]  full T x (S+1) grid per utterance (m * 2^e), zero outside the band
    Cell *beta = nullptr;          // [rows]
    float4 *coef = nullptr;        // [rows]  per-row gradient record (H, qb, ql, L), see k2_lattice.cuh
    int *rowlab = nullptr;         // [rows]  the row's label as the gradient kernel needs it (-1 none, kRowDead dead)
    double *ll_fwd = nullptr;      // [B]     alpha(T-1, S)
    double *ll_bwd = nullptr;      // [B]     beta(0, 0)  (diagnostic, as in the reference)
    float *costs = nullptr;        // [B]     -ll_fwd
    unsigned *k2_flags = nullptr;  // [k2_flag_words(B)] two alternating sets of kK2FlagWords hand-over words per utterance;
                                   //         then, in cache lines of their own, the zero fill's and the tile hand-out's counters
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bytes needed for `shape`.  A function of (B, T[], S[]) only, like the reference
// (gpu_workspace_manager.h:242-247); the value differs (76 B/row + O(B*T_max) instead of 12 B/row).
// words of Workspace::k2_flags: two SETS of kK2FlagWords per utterance that take turns from launch to launch (a lattice
// launch uses set `epoch & 1` and clears the other one for its successor: nothing in a launch ever has to reset a word
// that a late CTA of the same launch might still touch), then (256 bytes clear of the flags that waiting CTAs poll)
// the counters of the zero fill (zero_fill.cuh): two of the OWNED protocol, then the two alternating ones of the
// SHARED protocol
constexpr int kK2FlagWords = 24;  // per utterance: phase-A block claims [8], phase-C block claims [8], phase-A blocks done
                                  // by helpers, recursion done (k2_lattice.cuh; 8 = kK2MaxParts)
__host__ __device__ inline size_t k2_zero_ctr_word(int B) {
    return (2 * kK2FlagWords * static_cast<size_t>(B) + 63) / 64 * 64 + 64;
}
// ... and, a cache line further on, the tile hand-out counter of the gradient kernel's dynamic mode (k3_grad.cuh):
// {tiles handed out, producers finished}; zero between launches
__host__ __device__ inline size_t stream_ctr_word(int B) { return k2_zero_ctr_word(B) + 64; }
__host__ __device__ inline size_t k2_flag_words(int B) { return stream_ctr_word(B) + 4; }
inline size_t workspace_bytes(const Shape &sh) {
    const size_t rows = static_cast<size_t>(sh.rows);
    const size_t B = static_cast<size_t>(sh.B);
    size_t n = 256;  // slack to align an arbitrarily aligned base pointer
    n += align_up((B + 1) * sizeof(int64_t), 256);
    n += align_up(B * static_cast<size_t>(sh.T_max) * sizeof(int2), 256);
    n += 2 * align_up(rows * sizeof(int), 256);
    n += align_up(rows * sizeof(RawRow), 256);
    n += align_up(rows * sizeof(Weight), 256);
    n += 2 * align_up(rows * sizeof(Cell), 256);
    n += align_up(rows * sizeof(float4), 256);
    n += align_up(rows * sizeof(int), 256);
    n += 2 * align_up(B * sizeof(double), 256);
    n += align_up(B * sizeof(float), 256);
    n += align_up(k2_flag_words(sh.B) * sizeof(unsigned), 256);
    return n;
}

inline Workspace carve_workspace(void *base, const Shape &sh) {
    const size_t rows = static_cast<size_t>(sh.rows);
    const size_t B = static_cast<size_t>(sh.B);
    char *p = reinterpret_cast<char *>(align_up(reinterpret_cast<size_t>(base), 256));
    auto take = [&p](size_t bytes) {
        char *r = p;
        p += align_up(bytes, 256);
        return r;
    };
    Workspace w;
    w.row_start = reinterpret_cast<int64_t *>(take((B + 1) * sizeof(int64_t)));
    w.band = reinterpret_cast<int2 *>(take(B * static_cast<size_t>(sh.T_max) * sizeof(int2)));
    w.rowmeta = reinterpret_cast<int *>(take(rows * sizeof(int)));
    w.rowutt = reinterpret_cast<int *>(take(rows * sizeof(int)));
    w.lp = reinterpret_cast<RawRow *>(take(rows * sizeof(RawRow)));
    w.wts = reinterpret_cast<Weight *>(take(rows * sizeof(Weight)));
    w.alpha = reinterpret_cast<Cell *>(take(rows * sizeof(Cell)));
    w.beta = reinterpret_cast<Cell *>(take(rows * sizeof(Cell)));
    w.coef = reinterpret_cast<float4 *>(take(rows * sizeof(float4)));
    w.rowlab = reinterpret_cast<int *>(take(rows * sizeof(int)));
    w.ll_fwd = reinterpret_cast<double *>(take(B * sizeof(double)));
    w.ll_bwd = reinterpret_cast<double *>(take(B * sizeof(double)));
    w.costs = reinterpret_cast<float *>(take(B * sizeof(float)));
    w.k2_flags = reinterpret_cast<unsigned *>(take(k2_flag_words(sh.B) * sizeof(unsigned)));
    return w;
}

// ---------------------------------------------------------------------------------------------
// row_start[b] = sum_{b' < b} T_b' * (S_b' + 1)      (one CTA; block-wide exclusive scan, 64-bit)
// Also clears the per-utterance hand-over flags of the lattice kernel.
// Replaces the host loops + H2D copies at gpu_workspace_manager.h:262-289.
// ---------------------------------------------------------------------------------------------
constexpr int kPlanThreads = 1024;

static __global__ void __launch_bounds__(kPlanThreads) plan_row_start_kernel(const int *__restrict__ T,
                                                                       const int *__restrict__ S, int B,
                                                                       int64_t *__restrict__ row_start,
                                                                       unsigned *__restrict__ k2_flags,
                                                                       int64_t padded_block_rows) {
    __shared__ int64_t warp_tot[kPlanThreads / kWarp];
    const int tid = threadIdx.x;
    const int per = (B + kPlanThreads - 1) / kPlanThreads;
    const int b0 = tid * per;
    for (int b = tid; b < static_cast<int>(k2_flag_words(B)); b += kPlanThreads) k2_flags[b] = 0u;
    if (padded_block_rows > 0) {  // padded layout: every utterance owns a block of T_dim * U rows
        for (int b = tid; b <= B; b += kPlanThreads) row_start[b] = b * padded_block_rows;
        return;
    }
    int64_t local = 0;
    for (int i = 0; i < per; ++i) {
        const int b = b0 + i;
        if (b < B) local += static_cast<int64_t>(T[b]) * (S[b] + 1);
    }
    // inclusive scan of `local` across the block
    int64_t incl = local;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int64_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int64_t w = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t up = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += up;
        }
        warp_tot[lane] = w;  // inclusive totals of warps 0..lane
    }
    __syncthreads();
    int64_t run = incl - local + (warp > 0 ? warp_tot[warp - 1] : 0);  // exclusive prefix of this thread
    for (int i = 0; i < per; ++i) {
        const int b = b0 + i;
        if (b < B) {
            row_start[b] = run;
            run += static_cast<int64_t>(T[b]) * (S[b] + 1);
            if (b == B - 1) row_start[B] = run;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Alignment band.  Unrestricted: (0, S_b) for every frame (gpu_workspace_manager.h:315-328).
// Restricted (restrict_to_alignment, gpu_workspace_manager.h:191-219 / cpu twin :207-224):
//   m[0] = 0, m[t+1] = m[t] + (alignment[b*T_max + t] != blank_idx)
//   lo[t] = m[max(0, t+1-shift)],  hi[t] = m[min(T_b, t+1+shift)]
// One CTA per utterance; m[] lives in shared memory ((T_max+1) ints).
// ---------------------------------------------------------------------------------------------
constexpr int kBandThreads = 256;

static __global__ void __launch_bounds__(kBandThreads) band_kernel(const int *__restrict__ T, const int *__restrict__ S,
                                                             int T_max, const int *__restrict__ alignment,
                                                             int align_stride, int max_shift, int blank_idx,
                                                             int2 *__restrict__ band) {
    extern __shared__ int m_sh[];  // [T_b + 1] when alignment != nullptr
    __shared__ int warp_tot[kBandThreads / kWarp];
    __shared__ int carry_sh;
    const int b = blockIdx.x;
    const int Tb = T[b], Sb = S[b];
    int2 *band_b = band + static_cast<size_t>(b) * T_max;
    if (alignment == nullptr) {
        for (int t = threadIdx.x; t < T_max; t += kBandThreads) band_b[t] = make_int2(0, Sb);
        return;
    }
    const int *al = alignment + static_cast<size_t>(b) * align_stride;  // (the reference's stride: T_max)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        carry_sh = 0;
        m_sh[0] = 0;
    }
    __syncthreads();
    for (int base = 0; base < Tb; base += kBandThreads) {
        const int t = base + tid;
        const int flag = (t < Tb && al[t] != blank_idx) ? 1 : 0;
        int incl = flag;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        int before = carry_sh;
        for (int w = 0; w < warp; ++w) before += warp_tot[w];
        if (t < Tb) m_sh[t + 1] = before + incl;
        __syncthreads();
        if (tid == kBandThreads - 1) carry_sh = before + incl;
        __syncthreads();
    }
    const int shift = max_shift < 0 ? 0 : (max_shift > T_max ? T_max : max_shift);
    for (int t = tid; t < T_max; t += kBandThreads) {
        if (t < Tb) {
            const int a = t + 1 - shift;
            const int c = t + 1 + shift;
            band_b[t] = make_int2(m_sh[a > 0 ? a : 0], m_sh[c < Tb ? c : Tb]);
        } else {
            band_b[t] = make_int2(0, Sb);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// rowmeta[row]: which packed rows can carry probability mass at all.
// Row (b,t,s) multiplies alpha(t-1, s) everywhere it is used (alpha/beta recursions and the
// gradient, cpu_rnnt.h:158-166,188-196,220-231), so it is "live" iff alpha(t-1, s) lies inside the
// lattice:  t == 0: s == 0;  t >= 1: lo[t-1] <= s <= hi[t-1], s <= t, S-s <= T-t
// (gpu_rnnt_kernel.h:10-32).  Dead rows are never read by the streaming kernels: their gradient
// is written as zeros (the reference does the same for the geometric part, gpu_rnnt_kernel.h:266-271).
// Live rows carry the index of their label in labels[] (labels are read at compute time).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_utterance(const int64_t *__restrict__ row_start, int B, int64_t row) {
    int lo = 0, hi = B;  // invariant: row_start[lo] <= row < row_start[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (row_start[mid] <= row) lo = mid; else hi = mid;
    }
    return lo;
}

static __global__ void __launch_bounds__(256) rowmeta_kernel(const int *__restrict__ T, const int *__restrict__ S, int B,
                                                       int T_max, int label_stride, int ld_fixed,
                                                       const int64_t *__restrict__ row_start,
                                                       const int2 *__restrict__ band, int *__restrict__ rowmeta,
                                                       int *__restrict__ rowutt) {
    const int64_t rows = row_start[B];
    for (int64_t row = blockIdx.x * 256ll + threadIdx.x; row < rows; row += 256ll * gridDim.x) {
        const int b = find_utterance(row_start, B, row);
        const int Tb = T[b], Sb = S[b];
        const int local = static_cast<int>(row - row_start[b]);
        const int ld = ld_fixed > 0 ? ld_fixed : Sb + 1;
        const int t = local / ld;
        const int s = local - t * ld;
        bool live;
        if (t >= Tb || s > Sb) {
            live = false;  // padding of a padded tensor
        } else if (t == 0) {
            live = (s == 0);
        } else {
            const int2 lh = band[static_cast<size_t>(b) * T_max + (t - 1)];
            live = s >= lh.x && s <= lh.y && s <= t && (Sb - s) <= (Tb - t);
        }
        rowmeta[row] = live ? (s < Sb ? b * label_stride + s : kRowNoLabel) : kRowDead;
        rowutt[row] = b;
    }
}

// ---------------------------------------------------------------------------------------------
// The three set-up kernels above as ONE launch (batches of up to kPlanFusedMaxB utterances): grid (parts, B), CTA (p, b)
// works out row_start[b] (a block-wide sum over the utterances before b), the alignment counts m[] of utterance b in
// shared memory, and from them rowmeta / rowutt of slice p of the utterance's rows; the CTAs with p == 0 also leave
// row_start[] and band[] in the workspace for the lattice kernel, CTA (0, 0) clears the hand-over words.  The results
// are those of the three kernels, bit for bit (tests: DBG_ROWSTART / DBG_BAND / DBG_ROWMETA against the host
// restatement); what is saved are two launches and the two round trips through global memory between them -- ~6 us per
// manager, which a caller like the reference's torch binding pays on every loss call (a new manager per call).
// Shared memory: (T_max + 1) ints when there is an alignment, else none.
// ---------------------------------------------------------------------------------------------
constexpr int kPlanFusedThreads = 256;
constexpr int kPlanFusedMaxB = 1024;
constexpr int kPlanFusedRowsPerCta = 2048;
constexpr int kPlanFusedMaxParts = 64;

static __global__ void __launch_bounds__(kPlanFusedThreads)
    plan_fused_kernel(const int *__restrict__ T, const int *__restrict__ S, int B, int T_max, int label_stride, int ld_fixed,
                      int64_t padded_block_rows, const int *__restrict__ alignment, int align_stride, int max_shift,
                      int blank_idx, int64_t *__restrict__ row_start, int2 *__restrict__ band, int *__restrict__ rowmeta,
                      int *__restrict__ rowutt, unsigned *__restrict__ k2_flags) {
    extern __shared__ int m_sh[];  // [T_b + 1] when alignment != nullptr
    __shared__ int64_t red_sh[kPlanFusedThreads / kWarp];
    __shared__ int warp_tot[kPlanFusedThreads / kWarp];
    __shared__ int carry_sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, p = blockIdx.x, nparts = gridDim.x;
    const int Tb = T[b], Sb = S[b];
    if (b == 0 && p == 0)
        for (int i = tid; i < static_cast<int>(k2_flag_words(B)); i += kPlanFusedThreads) k2_flags[i] = 0u;

    // ---- row_start[b] ----
    int64_t start;
    if (padded_block_rows > 0) {
        start = b * padded_block_rows;
    } else {
        int64_t local = 0;
        for (int i = tid; i < b; i += kPlanFusedThreads) local += static_cast<int64_t>(T[i]) * (S[i] + 1);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if (lane == 0) red_sh[warp] = local;
        __syncthreads();
        start = 0;
        for (int w = 0; w < kPlanFusedThreads / kWarp; ++w) start += red_sh[w];
    }
    const int64_t nrows = padded_block_rows > 0 ? padded_block_rows : static_cast<int64_t>(Tb) * (Sb + 1);
    if (p == 0 && tid == 0) {
        row_start[b] = start;
        if (b == B - 1) row_start[B] = start + nrows;
    }

    // ---- alignment counts m[0 .. T_b] (band_kernel above) ----
    const bool restricted = alignment != nullptr;
    const int shift = max_shift < 0 ? 0 : (max_shift > T_max ? T_max : max_shift);
    if (restricted) {
        const int *al = alignment + static_cast<size_t>(b) * align_stride;
        if (tid == 0) {
            carry_sh = 0;
            m_sh[0] = 0;
        }
        __syncthreads();
        for (int base = 0; base < Tb; base += kPlanFusedThreads) {
            const int t = base + tid;
            int incl = (t < Tb && al[t] != blank_idx) ? 1 : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            if (lane == 31) warp_tot[warp] = incl;
            __syncthreads();
            int before = carry_sh;
            for (int w = 0; w < warp; ++w) before += warp_tot[w];
            if (t < Tb) m_sh[t + 1] = before + incl;
            __syncthreads();
            if (tid == kPlanFusedThreads - 1) carry_sh = before + incl;
            __syncthreads();
        }
    }
    auto band_of = [&](int t) {  // (lo, hi) of frame t < T_b
        if (!restricted) return make_int2(0, Sb);
        const int a = t + 1 - shift, c = t + 1 + shift;
        return make_int2(m_sh[a > 0 ? a : 0], m_sh[c < Tb ? c : Tb]);
    };
    if (p == 0) {
        int2 *band_b = band + static_cast<size_t>(b) * T_max;
        for (int t = tid; t < T_max; t += kPlanFusedThreads) band_b[t] = t < Tb ? band_of(t) : make_int2(0, Sb);
    }

    // ---- rowmeta / rowutt of this CTA's slice of the utterance's rows (rowmeta_kernel above) ----
    const int ld = ld_fixed > 0 ? ld_fixed : Sb + 1;
    const int64_t r0 = nrows * p / nparts, r1 = nrows * (p + 1) / nparts;
    for (int64_t local = r0 + tid; local < r1; local += kPlanFusedThreads) {
        const int t = static_cast<int>(local / ld);
        const int s = static_cast<int>(local - static_cast<int64_t>(t) * ld);
        bool live;
        if (t >= Tb || s > Sb) {
            live = false;
        } else if (t == 0) {
            live = (s == 0);
        } else {
            const int2 lh = band_of(t - 1);
            live = s >= lh.x && s <= lh.y && s <= t && (Sb - s) <= (Tb - t);
        }
        rowmeta[start + local] = live ? (s < Sb ? b * label_stride + s : kRowNoLabel) : kRowDead;
        rowutt[start + local] = b;
    }
}

// Host -> device upload of the logits that will be read (Engine::upload_live_rows): a warp per row, live rows only,
// straight out of pinned host memory over PCIe into the device array the kernels stream from.  The copy engine moves
// whole tensors; which rows the lattice reads is known from the plan (rowmeta), and the rows it calls dead -- 27 % of
// c2, ~95 % under c5's alignment band -- are never read by any kernel, so they need not cross the bus.  UNIT is the
// widest access the row size and the two base addresses allow (16 bytes wherever the streaming kernels apply).
constexpr int kUploadThreads = 256;
// The frames of an utterance all of whose states are live when no alignment band is set: row (t, s) is live iff
// alpha(t-1, s) lies inside the lattice, s <= t and S - s <= T - t (cpu_workspace_manager.h:161-181), for every s in
// [0, S] iff S <= t <= T - S.  One contiguous block of (T - 2 S + 1)(S + 1) packed rows per utterance: the copy engine
// takes it (Engine::upload_live_rows), the kernel below the ragged frames before and behind it.  Host and device use
// this one function, so the two never disagree about who brings a row.
struct MiddleBlock {
    int64_t first = 0;  // row index inside the utterance
    int64_t rows = 0;   // 0: none worth a copy of its own
};
__host__ __device__ inline MiddleBlock upload_middle_block(int Tb, int Sb, int64_t min_rows) {
    MiddleBlock m;
    const int64_t frames = static_cast<int64_t>(Tb) - 2 * static_cast<int64_t>(Sb) + 1;
    if (frames <= 0) return m;
    const int64_t n = frames * (Sb + 1);
    if (n < min_rows) return m;
    m.first = static_cast<int64_t>(Sb) * (Sb + 1);
    m.rows = n;
    return m;
}
// mid_min_rows > 0: the middle blocks of at least that many rows are somebody else's (the copy engine's) to bring.
template <typename UNIT>
static __global__ void __launch_bounds__(kUploadThreads) upload_live_rows_kernel(const UNIT *__restrict__ src,
                                                                                 UNIT *__restrict__ dst,
                                                                                 const int *__restrict__ rowmeta,
                                                                                 int64_t rows, int units_per_row,
                                                                                 int64_t mid_min_rows = 0,
                                                                                 const int *__restrict__ rowutt = nullptr,
                                                                                 const int64_t *__restrict__ row_start = nullptr,
                                                                                 const int *__restrict__ T = nullptr,
                                                                                 const int *__restrict__ S = nullptr) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(kUploadThreads) + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * kUploadThreads) >> 5;
    for (int64_t row = warp0; row < rows; row += nwarps) {
        if (__ldg(rowmeta + row) == kRowDead) continue;  // (warp-uniform)
        if (mid_min_rows > 0) {
            const int b = __ldg(rowutt + row);
            const MiddleBlock m = upload_middle_block(__ldg(T + b), __ldg(S + b), mid_min_rows);
            const int64_t r = row - __ldg(row_start + b);
            if (r >= m.first && r < m.first + m.rows) continue;
        }
        const UNIT *s = src + row * units_per_row;
        UNIT *d = dst + row * units_per_row;
        // eight loads in flight per lane (4 KB per warp at 16 bytes), ALL issued before the first store waits for one:
        // a round trip over PCIe takes microseconds, so the ragged end of a row is predicated, not a loop of its own
        for (int base = lane; base < units_per_row; base += 8 * 32) {
            UNIT v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (base + j * 32 < units_per_row) v[j] = __ldcs(s + base + j * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (base + j * 32 < units_per_row) d[base + j * 32] = v[j];
        }
    }
}

}  // namespace mrnnt
