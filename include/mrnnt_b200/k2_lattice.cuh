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
// The recursion is a latency chain of T frames, so the design keeps everything that is not the chain out
// of the warps that walk it:
//
//  * Arithmetic in a SCALED LINEAR domain: a cell is m * 2^e (float mantissa, int exponent, common.cuh),
//    which turns each log-sum-exp (conversions, two MUFU ops and three double adds on the chain) into two
//    multiplies and one exponent-aligned multiply-add.  The explicit per-cell exponent keeps the dynamic
//    range of the log domain (nothing underflows, whatever the logits); every term is non-negative, so
//    there is no cancellation and the relative error of a cell grows like sqrt(frames) * 2^-24.
//  * A SYSTOLIC ROW OF WARPS per direction: warp i owns the label states [32*K*i, 32*K*(i+1)) (K states per
//    lane, K = 1 up to 192 states).  Inside a warp the neighbour cell moves by one shuffle; between warps
//    it moves through a small tagged FIFO in shared memory that the downstream warp polls, so warps never
//    meet at a barrier: warp i+1 simply runs a few frames behind warp i.
//  * `parts` CTAs per utterance (sized so that the whole grid is co-resident: the helpers sit on SMs that
//    would otherwise idle during this kernel) and three phases separated by per-utterance flags:
//    A. all parts turn the (x_blank, x_label, max, sum) records K1 left per row into the row's denominator
//       -log2 sum exp (a float pair, written back in place) and its transition weights (mantissa, exponent), in
//       float arithmetic, and fold the WHOLE band logic into them: a weight is
//       zeroed when the cell it leads into lies outside the lattice (alpha_s_min/max of the target frame,
//       cpu_workspace_manager.h:67-71; the beta limits :73-86 are the same set shifted by one frame).  The
//       chain warps therefore contain no band arithmetic at all.
//    B. part 0 runs the two systolic rows; the weights of a whole CHUNK of frames arrive in shared memory by
//       one bulk async copy (TMA engine) per chunk, issued by an otherwise idle warp per direction.
//    C. all parts fold alpha, beta, the likelihood and the denominators into one record per row, so that the
//       gradient kernel is a pure stream.  With r0 = log2( alpha(t-1,s) beta(t,s) / Z ) (the row's occupancy, <= 0) and
//       D = dh + dl the row's base-2 log-softmax denominator:
//         (H, L) = dh + (r0 + dl) as an error-free pair:   g[v] = 2^((x[v] kLog2e + H) + L) - ...
//         qb = log2( alpha(t-1,s) p_blank(t,s) beta(t+1,s)   / Z )   what is subtracted at v == blank:    2^qb
//         ql = log2( alpha(t-1,s) p_label(t,s) beta(t+1,s+1) / Z )   what is subtracted at v == label_s:  2^ql
//       The large parts (H ~ -max * log2 e: hundreds for logits of magnitude 100) meet the logit inside one fused
//       multiply-add whose RESULT is small.  qb and ql are formed HERE, from the row's blank / label logit in two-float
//       arithmetic: where a path is forced through an improbable transition (p = 2^-400, posterior ~ 1) the lattice
//       term is +400 and log2 p is -400, and rounding them separately before they cancel costs an ulp of 400 in the
//       exponent, 4e-5 in the gradient.
#pragma once

#include "common.cuh"
#include "plan.cuh"
#include "zero_fill.cuh"

namespace mrnnt {

#ifdef MRNNT_K2_TRACE  // development aid (tools/k2_probe.cu): clock64 stamps of the alpha pass
__device__ long long g_k2_trace[64];
#define MRNNT_K2_STAMP(i) do { if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) g_k2_trace[i] = clock64(); } while (0)
#else
#define MRNNT_K2_STAMP(i) do { } while (0)
#endif

constexpr int kK2Threads = 512;
constexpr int kK2Warps = kK2Threads / kWarp;
constexpr int kK2MaxChunkBufs = 8;          // chunk buffers per direction at most (k2_chunk_bufs)
constexpr int kK2MaxChunkFrames = 16;       // frames per chunk at most; also the renormalisation period
constexpr int kK2ChunkTargetBytes = 12288;  // bytes of weights per chunk we aim for
constexpr int kK2MaxRowWarps = 6;           // chain warps per direction at most
constexpr int kK2FifoDepth = 256;           // > kK2MaxChunkBufs * kK2MaxChunkFrames: a warp cannot lap its neighbour
constexpr int kK2MaxParts = 8;              // CTAs per utterance for the coefficient phase

struct K2Args {
    const int *T;
    const int *S;
    const int *labels;
    const int64_t *row_start;
    const int2 *band;
    RawRow *lp;        // [rows] K1's records; phase A replaces their (max, sum) by the denominator pair in place
    Weight *wts;       // [rows] transition weights, written by phase A
    Cell *alpha;
    Cell *beta;
    float4 *coef;      // [rows] (H, qb, ql, L) per row, see the head of this file; H == -inf: a zero gradient row
    int *rowlab;       // [rows] the row's label for the gradient kernel: >= 0, -1 none (or blank: the blank branch wins),
                       //        kRowDead: a row the plan calls dead (whoever zeroes those rows has done so or will)
    double *ll_fwd;
    double *ll_bwd;
    float *costs;
    float *costs_mapped;  // optional second copy of the costs in host-mapped pinned memory (nullptr: none)
    unsigned *flags;   // two alternating sets of kK2FlagWords hand-over words per utterance (plan.cuh), zeroed at set-up
    unsigned epoch;    // launch counter (never 0): picks the set, and is the value part 0 publishes when its recursions are done
    int T_max;
    int S_max;
    int ld;            // rows per frame: 0 = packed layout (S_b + 1 per utterance), else the fixed U of a padded tensor
    int T_dim;         // padded layout: frames per utterance block (0 = packed)
    int label_stride;  // ints per utterance in labels[]
    int V;
    int blank;
    int need_beta;     // 0: cost only (alpha pass), 1: alpha + beta + coefficients
    int chunk_frames;  // frames per bulk-copied chunk (host: k2_chunk_frames)
    int row_warps;     // chain warps per direction (host: k2_row_warps)
    int chunk_bufs;    // chunk buffers per direction (host: k2_chunk_bufs)
    int parts;         // CTAs per utterance; CTA index = b * parts + part
    // Dead-row zero fill (optional): while the recursions run, the memory system has nothing to do; warps
    // 0..zero_warps-1 of every CTA spend that time writing the zero rows of the gradient (rowmeta == kRowDead),
    // so that the gradient kernel only touches live rows.  The grid is filled up to one CTA per SM with CTAs that do
    // nothing else.
    unsigned char *zero_dst;  // the gradient buffer (nullptr: the gradient kernel writes the zeros itself)
    const int *rowmeta;       // [rows]
    unsigned row_bytes;       // V * sizeof(element), a multiple of 16
    int zero_warps;           // 0: off
    int64_t zero_unit_end;    // this kernel's fill covers the units (of 32 rows) before this one (< 0: all of them); the
                              // gradient kernel's zero-fill warp takes the rest
    int64_t rows;             // rows of the whole batch (= row_start[B])
    int B;                    // utterances; flags[k2_zero_ctr_word(B)...]: the zero fill's two counters
    unsigned *zero_clear;     // SHARED zero fill: the counter of the NEXT call, cleared here (nullptr: none)
    unsigned *zero_shared_ctr;  // SHARED zero fill: this call's counter -- the fill warps of this kernel take units from it
                              // between the LSE kernel's zero-fill warp and the gradient kernel's, for as long as the
                              // recursions run (nullptr: the OWNED protocol, this kernel does the whole fill)
    int phase_ctas;           // B * parts; CTAs behind them (zero fill only) do nothing else
    int tl_slot;              // (MRNNT_TIMELINE: the call's slot of the timeline)
};

// States per lane for a launch whose widest utterance has `states` states; 0: use the wide kernel.  Measured
// (tools/k2_probe.cu, cycles per frame of the leading warp): K = 1 alone 80, K = 2 alone 115, K = 4 alone 300;
// a link to a neighbour warp adds ~10 and one chunk of lag.  Hence as few states per lane as the row of warps
// allows.
inline int k2_states_per_lane(int states) {
    for (int K = 1; K <= 4; K *= 2)
        if (states <= kWarp * K * kK2MaxRowWarps) return K;
    return 0;
}
inline int k2_row_warps(int states, int K) { return (states + kWarp * K - 1) / (kWarp * K); }
// frames per chunk; `width` = rows per frame in memory (S_max + 1 packed, U padded)
inline int k2_chunk_frames(int width) {
    const int frame_bytes = width * static_cast<int>(sizeof(Weight));
    int ch = kK2ChunkTargetBytes / frame_bytes;
    if (ch < 1) ch = 1;
    if (ch > kK2MaxChunkFrames) ch = kK2MaxChunkFrames;
    return ch;
}
// Chunk buffers per direction: every chain warp runs one chunk behind its upstream neighbour (the hand-over
// between warps is checked once per chunk), so the ring has to span the whole row of warps plus what is in
// flight from memory.
inline int k2_chunk_bufs(int row_warps) {
    const int n = row_warps + 3;
    return n > kK2MaxChunkBufs ? kK2MaxChunkBufs : n;
}
// shared memory of one direction: chunk ring | 2 mbarriers per slot | FIFOs between neighbouring chain warps
__host__ __device__ inline size_t k2_dir_bytes(size_t chunk_bytes, int bufs, int row_warps) {
    return bufs * chunk_bytes + 2 * kK2MaxChunkBufs * sizeof(uint64_t) + 64 +
           static_cast<size_t>(row_warps > 1 ? row_warps - 1 : 0) * kK2FifoDepth * sizeof(Cell);
}
constexpr int kK2ZeroBytes = kZeroFillBytes;
__host__ __device__ inline size_t k2_zero_offset(size_t dir_bytes) { return (2 * dir_bytes + 127) / 128 * 128; }
inline size_t k2_smem_bytes(int width, int row_warps) {
    return k2_zero_offset(k2_dir_bytes(static_cast<size_t>(k2_chunk_frames(width)) * width * sizeof(Weight),
                                       k2_chunk_bufs(row_warps), row_warps)) + kK2ZeroBytes;
}

__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// One direction's shared memory.
struct K2Dir {
    unsigned char *buf;  // [bufs][chunk_bytes] of Weight
    uint64_t *w_full;    // [bufs] bulk copy landed            (1 arrival + transaction bytes)
    uint64_t *empty;     // [bufs] consumed by the chain warps (one arrival per chain warp)
    Cell *fifo;          // [row_warps - 1][kK2FifoDepth] what crosses from one chain warp to the next, per frame
    int *progress;       // [kK2MaxRowWarps] per link: the last exchange whose entry (and all before it) is in place
    size_t chunk_bytes;
    int bufs;
    __device__ __forceinline__ K2Dir(unsigned char *base, size_t chunk_bytes_, int bufs_)
        : buf(base), chunk_bytes(chunk_bytes_), bufs(bufs_) {
        w_full = reinterpret_cast<uint64_t *>(base + bufs_ * chunk_bytes_);
        empty = w_full + kK2MaxChunkBufs;
        progress = reinterpret_cast<int *>(empty + kK2MaxChunkBufs);  // 64 bytes reserved
        fifo = reinterpret_cast<Cell *>(reinterpret_cast<unsigned char *>(empty + kK2MaxChunkBufs) + 64);
    }
    __device__ __forceinline__ unsigned char *slot(int i) const { return buf + i * chunk_bytes; }
};

constexpr int kK2NoProgress = -(1 << 30);

// rows per frame of utterance b in memory
__device__ __forceinline__ int k2_ld(const K2Args &a, int Sb) { return a.ld > 0 ? a.ld : Sb + 1; }

// Frames of chunk c.  Direction 0 (alpha) walks the frames upwards, direction 1 (beta) downwards.
__device__ __forceinline__ int k2_chunk_lo(int dir, int c, int CH, int Tb) {
    return dir == 0 ? c * CH : max(Tb - (c + 1) * CH, 0);
}
__device__ __forceinline__ int k2_chunk_hi(int dir, int c, int CH, int Tb) {  // exclusive
    return dir == 0 ? min((c + 1) * CH, Tb) : Tb - c * CH;
}

// alpha(t, s) lies inside the lattice (cpu_workspace_manager.h:67-71): lh = band[t]
__device__ __forceinline__ bool k2_alpha_valid(int t, int s, int2 lh, int Tb, int Sb) {
    return s >= max(lh.x, t - (Tb - 1 - Sb)) && s <= min(min(lh.y, t + 1), Sb);
}

// Which transitions out of row (t, s) exist.  The row is live iff alpha(t-1, s) is inside the lattice (t == 0:
// the virtual alpha(-1, .) = [1, 0, ...]); a transition is kept iff its target alpha(t, s) / alpha(t, s+1) is.
// lh = band[t], lq = band[t-1].  beta_last: the variant the beta recursion needs at t = T-1, where its
// source beta(T, .) = [s == S] whatever the band says about alpha(T-1, .) (cpu_workspace_manager.h:185-205).
__device__ __forceinline__ void k2_row_flags(int t, int s, int2 lh, int2 lq, int Tb, int Sb, bool beta_last,
                                             bool &stay_ok, bool &emit_ok) {
    const int reach = Tb - 1 - Sb;  // alpha_s_min(t) = max(lo[t], t - reach)
    const int amin = max(lh.x, t - reach), amax = min(min(lh.y, t + 1), Sb);
    const bool live = (t == 0) ? (s == 0) : (s >= max(lq.x, t - 1 - reach) && s <= min(min(lq.y, t), Sb));
    stay_ok = live && (beta_last ? s == Sb : (s >= amin && s <= amax));              // blank: (t-1,s) -> (t,s)
    emit_ok = live && (beta_last ? s + 1 == Sb : (s + 1 >= amin && s + 1 <= amax));  // label: (t-1,s) -> (t,s+1)
}

__device__ __forceinline__ Weight k2_make_weight(const RawRow &raw, bool stay_ok, bool emit_ok) {
    Weight w;
    weight_from_logit(raw.xb, raw.dh, raw.dl, stay_ok, w.mb, w.eb);
    weight_from_logit(raw.xl, raw.dh, raw.dl, emit_ok, w.ml, w.el);
    return w;
}

// ---- phase A: rows [part*n/parts, (part+1)*n/parts) of utterance b, whole CTA ---------------------------
__device__ __forceinline__ void k2_weight_rows(const K2Args &a, int b, int part, int parts, int tid = threadIdx.x,
                                               int nthreads = kK2Threads) {
    const int Tb = a.T[b], Sb = a.S[b];
    const int W = Sb + 1, ld = k2_ld(a, Sb);
    const int64_t R = a.row_start[b];
    RawRow *lp = a.lp + R;
    Weight *wts = a.wts + R;
    const int2 *band = a.band + static_cast<size_t>(b) * a.T_max;
    const int n = Tb * W;
    const int i_begin = static_cast<int>(static_cast<int64_t>(n) * part / parts);
    const int i_end = static_cast<int>(static_cast<int64_t>(n) * (part + 1) / parts);
    constexpr int U = 4;  // rows per thread per batch: every load first, then 2*U independent conversions
    MRNNT_K2_STAMP(55);
    for (int base = i_begin + tid; base < i_end; base += U * nthreads) {
        RawRow raw[U];
        int2 lh[U], lq[U];
        int tt[U], ss[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = min(base + u * nthreads, i_end - 1);  // clamped: loads unconditional, stores are not
            const int t = i / W;
            tt[u] = t;
            ss[u] = i - t * W;
            lh[u] = band[t];
            lq[u] = band[max(t - 1, 0)];
            // never written for dead rows: whatever is there is masked by the flags below
            raw[u] = lp[static_cast<int64_t>(t) * ld + ss[u]];
        }
#ifdef MRNNT_K2_TRACE
        if (raw[0].xb + raw[1].xb + raw[2].xb + raw[3].xb == 123.456f) return;  // (the stamp below waits for the loads)
#endif
        MRNNT_K2_STAMP(56);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            bool stay_ok, emit_ok;
            k2_row_flags(tt[u], ss[u], lh[u], lq[u], Tb, Sb, false, stay_ok, emit_ok);
            // K1 left (max * log2 e, sum of 2^(x log2 e - max)); from here on the record holds -log2 sum exp
            const Denominator d = lse_finish(raw[u].dh, raw[u].dl);
            raw[u].dh = d.hi;
            raw[u].dl = d.lo;
            const int i = base + u * nthreads;
            if (i < i_end) {
                const int64_t idx = static_cast<int64_t>(tt[u]) * ld + ss[u];
                *reinterpret_cast<float2 *>(&lp[idx].dh) = make_float2(d.hi, d.lo);
                wts[idx] = k2_make_weight(raw[u], stay_ok, emit_ok);
            }
        }
    }
    MRNNT_K2_STAMP(57);
}

// ---- phase B, one lane per direction: keep the chunk ring full --------------------------------------------
__device__ __forceinline__ void k2_issue_chunks(const K2Args &a, int b, int dir, K2Dir &ring) {
    const int Tb = a.T[b], W = a.S[b] + 1, ld = k2_ld(a, a.S[b]);
    const int CH = a.chunk_frames;
    const Weight *wts = a.wts + a.row_start[b];
    const int nchunks = (Tb + CH - 1) / CH;
    for (int c = 0; c < nchunks; ++c) {
        const int slot = c % ring.bufs;
        const int lap = c / ring.bufs;
        if (lap > 0) mbar_wait(ring.empty + slot, static_cast<uint32_t>((lap - 1) & 1));  // chain warps released it
        const int lo = k2_chunk_lo(dir, c, CH, Tb), hi = k2_chunk_hi(dir, c, CH, Tb);
        // whole frames as they lie in memory (in a padded tensor: with the unused states behind S_b), except that
        // the last frame stops at its last state so that the copy never leaves the utterance's block
        const uint32_t bytes = static_cast<uint32_t>((hi - lo - 1) * ld + W) * sizeof(Weight);
        mbar_arrive_expect_tx(ring.w_full + slot, bytes);
        bulk_g2s(ring.slot(slot), wts + static_cast<int64_t>(lo) * ld, bytes, ring.w_full + slot);
    }
}

// ---- chain warps -------------------------------------------------------------------------------------
// Warp `wi` of direction DIR owns the states s0 .. s0+K-1 per lane, s0 = (wi*32 + lane)*K.
//   alpha: new(s) = old(s) * wb(t,s) + old(s-1) * wl(t,s-1)     -- the PRODUCT old(s-1)*wl moves up one lane
//   beta:  new(s) = old(s) * wb(t,s) + old(s+1) * wl(t,s)       -- the STATE old(s+1) moves down one lane
// (old = frame t-1 for alpha, t+1 for beta).  Masks are already in the weights.  HAS_IN / HAS_OUT: the warp has
// an upstream / downstream neighbour warp (compile-time, so that the frame loop is branch-free apart from the
// FIFO poll).
template <int K, int DIR, bool HAS_IN, bool HAS_OUT>
__device__ __forceinline__ void k2_chain_warp(const K2Args &a, int b, K2Dir &ring, int wi) {
    const int lane = threadIdx.x & 31;
    const int Tb = a.T[b], Sb = a.S[b];
    const int ld = k2_ld(a, Sb);
    const int CH = a.chunk_frames;
    const int64_t R = a.row_start[b];
    const int s0 = (wi * kWarp + lane) * K;
    const int nchunks = (Tb + CH - 1) / CH;
    if (DIR == 0 && wi == 0) MRNNT_K2_STAMP(1);

    int col[K];
#pragma unroll
    for (int j = 0; j < K; ++j) col[j] = min(s0 + j, Sb);

    float m[K];
    int e[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {  // alpha(-1, .) = [1, 0, 0, ...]   beta(T, .) = [0, ..., 0, 1]
        const bool one = (DIR == 0) ? (s0 + j == 0) : (s0 + j == Sb);
        m[j] = one ? 1.0f : 0.0f;
        e[j] = one ? 0 : kZeroExp;
    }

    // links to the neighbouring chain warps: alpha receives from wi-1 and sends to wi+1, beta the other way
    const int link_in = (DIR == 0) ? max(wi - 1, 0) : wi, link_out = (DIR == 0) ? wi : max(wi - 1, 0);
    const Cell *fifo_in = ring.fifo + static_cast<size_t>(link_in) * kK2FifoDepth;
    Cell *fifo_out = ring.fifo + static_cast<size_t>(link_out) * kK2FifoDepth;
    const volatile int *progress_in = ring.progress + link_in;
    volatile int *progress_out = ring.progress + link_out;
    const bool edge_in = lane == ((DIR == 0) ? 0 : 31);  // lane whose neighbour lives in the other warp
    const bool edge_out = lane == ((DIR == 0) ? 31 : 0);

    Cell *grid = (DIR == 0 ? a.alpha : a.beta) + R + s0;
    const bool store = a.need_beta != 0;  // the stored grids only feed the coefficient phase
    bool st[K];
#pragma unroll
    for (int j = 0; j < K; ++j) st[j] = store && (s0 + j <= Sb);
    const int wstep = (DIR == 0) ? ld : -ld;

    int slot = 0;
    uint32_t parity = 0;
    for (int c = 0; c < nchunks; ++c) {
        const int tlo = k2_chunk_lo(DIR, c, CH, Tb);
        const int nf = k2_chunk_hi(DIR, c, CH, Tb) - tlo;
        const int t_first = (DIR == 0) ? tlo : tlo + nf - 1;
        if (DIR == 0 && wi == 0) MRNNT_K2_STAMP(2 + 2 * min(c, 20));
        mbar_wait(ring.w_full + slot, parity);
        if (DIR == 0 && wi == 0) MRNNT_K2_STAMP(3 + 2 * min(c, 20));
        const Weight *wp = reinterpret_cast<const Weight *>(ring.slot(slot)) + static_cast<size_t>(t_first - tlo) * ld;
        Cell *out = grid + static_cast<int64_t>(t_first) * ld;
        // sequence number of the exchange with the neighbour warp, counted from alpha's / beta's first frame
        // (beta numbers its first frame -1: what it sends there is its initial state beta(T, .))
        int seq = (DIR == 0) ? t_first : Tb - 2 - t_first;
        if (HAS_IN) {
            // Hand-over from the upstream warp, once per chunk: wait (politely) until it has published this chunk's
            // LAST exchange; the frame loop then reads its entries with plain loads, no polling, no branches.
            // (Polling per frame makes this warp chase its neighbour and flood the shared-memory pipe they share.)
            const int last = seq + nf - 1;
            while (*progress_in < last) __nanosleep(64);
            __threadfence_block();
        }
        Weight wn[K];
#pragma unroll
        for (int j = 0; j < K; ++j) wn[j] = wp[col[j]];
        if (DIR == 1 && c == 0) {
            // frame T-1 for beta: its own masking rule (k2_row_flags), straight from K1's records
            const int2 *band = a.band + static_cast<size_t>(b) * a.T_max;
            const int2 lh = band[Tb - 1], lq = band[max(Tb - 2, 0)];
#pragma unroll
            for (int j = 0; j < K; ++j) {
                bool stay_ok, emit_ok;
                k2_row_flags(Tb - 1, col[j], lh, lq, Tb, Sb, true, stay_ok, emit_ok);
                RawRow raw;
                raw.xb = raw.xl = kNegInfF;
                raw.dh = raw.dl = 0.0f;
                if (stay_ok || emit_ok) raw = a.lp[R + static_cast<int64_t>(Tb - 1) * ld + col[j]];
                wn[j] = k2_make_weight(raw, stay_ok, emit_ok);
            }
        }
        for (int i = 0; i < nf; ++i, ++seq, out += wstep) {
            Weight w[K];
#pragma unroll
            for (int j = 0; j < K; ++j) w[j] = wn[j];
            // next frame's weights, off the chain.  Unconditional: behind the chunk's last frame this reads a frame
            // of a neighbouring ring slot (or of the barrier/FIFO area), inside this CTA's shared memory, unused.
            wp += wstep;
#pragma unroll
            for (int j = 0; j < K; ++j) wn[j] = wp[col[j]];
            Cell q;
            q.m = 0.0f;
            q.e = kZeroExp;
            if (HAS_IN) q = fifo_in[seq & (kK2FifoDepth - 1)];  // in place since the chunk's hand-over

            // what crosses the lane boundary: alpha sends the product of its top state, beta its bottom state
            float x_m;
            int x_e;
            if (DIR == 0) {
                x_m = m[K - 1] * w[K - 1].ml;
                x_e = e[K - 1] + w[K - 1].el;
            } else {
                x_m = m[0];
                x_e = e[0];
            }
            float n_m = (DIR == 0) ? __shfl_up_sync(0xffffffffu, x_m, 1) : __shfl_down_sync(0xffffffffu, x_m, 1);
            int n_e = (DIR == 0) ? __shfl_up_sync(0xffffffffu, x_e, 1) : __shfl_down_sync(0xffffffffu, x_e, 1);
            if (HAS_OUT && edge_out) {
                Cell x;
                x.m = x_m;
                x.e = x_e;
                fifo_out[seq & (kK2FifoDepth - 1)] = x;
            }
            if (HAS_IN) {
                n_m = edge_in ? q.m : n_m;
                n_e = edge_in ? q.e : n_e;
            } else {
                n_m = edge_in ? 0.0f : n_m;  // alpha(t-1, -1) = 0 / beta(t+1, beyond the last warp's states) = 0
                n_e = edge_in ? kZeroExp : n_e;
            }

            float nm[K];
            int ne[K];
#pragma unroll
            for (int j = 0; j < K; ++j) {
                float in_m;
                int in_e;
                if (DIR == 0) {  // product arriving from state s-1 (its own label weight already applied)
                    in_m = (j == 0) ? n_m : m[j - 1] * w[j - 1].ml;
                    in_e = (j == 0) ? n_e : e[j - 1] + w[j - 1].el;
                } else {         // state s+1 times this row's label weight
                    in_m = ((j == K - 1) ? n_m : m[j + 1]) * w[j].ml;
                    in_e = ((j == K - 1) ? n_e : e[j + 1]) + w[j].el;
                }
                const Cell r = cell_add(in_m, in_e, m[j] * w[j].mb, e[j] + w[j].eb);
                nm[j] = r.m;
                ne[j] = r.e;
            }
#pragma unroll
            for (int j = 0; j < K; ++j) {
                m[j] = nm[j];
                e[j] = ne[j];
                if (st[j]) {
                    Cell cval;
                    cval.m = nm[j];
                    cval.e = ne[j];
                    out[j] = cval;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < K; ++j) cell_renorm(m[j], e[j]);
        if (HAS_OUT) {
            __threadfence_block();  // this chunk's FIFO entries before the word that announces them
            if (edge_out) *progress_out = seq - 1;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(ring.empty + slot);
        if (++slot == ring.bufs) {
            slot = 0;
            parity ^= 1u;
        }
    }
    if (DIR == 0 && wi == 0) MRNNT_K2_STAMP(50);
    if (DIR == 0) {
        // ll = log alpha(T-1, S): held by the lane that owns state S
#pragma unroll
        for (int j = 0; j < K; ++j) {
            if (s0 + j == Sb) {
                const double ll = cell_log(m[j], e[j]);
                a.ll_fwd[b] = ll;
                a.costs[b] = static_cast<float>(-ll);
                if (a.costs_mapped != nullptr) a.costs_mapped[b] = static_cast<float>(-ll);
            }
        }
    } else if (wi == 0 && lane == 0) {
        a.ll_bwd[b] = cell_log(m[0], e[0]);  // beta(0, 0)
    }
}

template <int K, int DIR>
__device__ __forceinline__ void k2_chain_dispatch(const K2Args &a, int b, K2Dir &ring, int wi, int nrow_warps) {
    const bool has_in = (DIR == 0) ? (wi > 0) : (wi + 1 < nrow_warps);
    const bool has_out = (DIR == 0) ? (wi + 1 < nrow_warps) : (wi > 0);
    if (has_in) {
        if (has_out) k2_chain_warp<K, DIR, true, true>(a, b, ring, wi);
        else k2_chain_warp<K, DIR, true, false>(a, b, ring, wi);
    } else {
        if (has_out) k2_chain_warp<K, DIR, false, true>(a, b, ring, wi);
        else k2_chain_warp<K, DIR, false, false>(a, b, ring, wi);
    }
}

// ---- per-row gradient coefficients (rows [part*n/parts, (part+1)*n/parts) of utterance b, whole CTA) ---
// Semantics of the lookups follow the reference accessors (gpu_rnnt_kernel.h:10-56): alpha(-1,0)=1,
// alpha(-1,s>0)=0, beta(T,S)=1, beta(T,s<S)=0, beta(.,S+1)=0 (linear domain); everything else comes from
// the stored grids, which hold zeros outside the lattice.
//   log2( alpha * beta / Z ) = (e_a + e_b - e_Z) + log2( m_a * m_b / m_Z ): the integer part exactly, and the
// fraction in [0,1) through MUFU.LG2 on a mantissa in [1,2) (absolute error 2^-22.6: 1e-7 relative on the gradient).
// false: the cell is zero.
__device__ __forceinline__ bool k2_log2_ratio(const Cell &x, float ratio, int base_e, float &ipart, float &fpart) {
    if (x.m == 0.0f) return false;
    const int bits = __float_as_int(ratio * x.m);
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(fpart) : "f"(__int_as_float((bits & 0x007fffff) | 0x3f800000)));
    ipart = static_cast<float>(base_e + x.e + (bits >> 23) - 127);
    return true;
}
// The occupancy term of a row as the pair (H, L) = dh + (log2(alpha beta / Z) + lo), H + L exact to ~48 bits.
__device__ __forceinline__ void k2_coef_main(const Cell &x, float ratio, int base_e, float dh, float lo, float &H, float &L) {
    float ip, fp;
    if (!k2_log2_ratio(x, ratio, base_e, ip, fp)) {
        H = kNegInfF;
        L = 0.0f;
        return;
    }
    float s, e1;
    two_sum(ip, dh, s, e1);
    const float rest = (e1 + lo) + fp;
    two_sum(s, rest, H, L);
}
// log2 of a transition's posterior: log2(alpha beta' / Z) + x * log2(e) + dh + dl, every large part added error-free
// (as in weight_from_logit); the result is <= 0 up to rounding.
__device__ __forceinline__ float k2_coef_patch(const Cell &x, float ratio, int base_e, float logit, float dh, float dl) {
    float ip, fp;
    if (!k2_log2_ratio(x, ratio, base_e, ip, fp) || !(logit > kNegInfF)) return kNegInfF;
    const float ph = logit * kLog2e;
    const float pl = fmaf(logit, kLog2eLo, fmaf(logit, kLog2e, -ph));
    float s, e1, h, e2;
    two_sum(ph, dh, s, e1);   // log2 p, high part
    two_sum(ip, s, h, e2);    // + the lattice term's integer part: what is left is small
    return h + (((e2 + e1) + (pl + dl)) + fp);
}

__device__ __forceinline__ void k2_coef_rows(const K2Args &a, int b, int part, int parts, int tid = threadIdx.x,
                                             int nthreads = kK2Threads) {
    const int Tb = a.T[b], Sb = a.S[b];
    const int ld = k2_ld(a, Sb);
    const int64_t R = a.row_start[b];
    const Cell *alpha = a.alpha + R;
    const Cell *beta = a.beta + R;
    const RawRow *lp = a.lp + R;
    const int *labels = a.labels + static_cast<size_t>(b) * a.label_stride;
    float4 *coef = a.coef + R;
    int *rowlab = a.rowlab + R;
    const int *rowmeta = a.rowmeta + R;
    // every row of the utterance's block gets a coefficient record (K3 takes row liveness from it); in a padded
    // tensor that includes the rows behind T_b and the states behind S_b, which are dead
    const int n = (a.ld > 0 ? a.T_dim : Tb) * ld;
    const int n_valid = (Tb - 1) * ld + Sb + 1;  // one past the last row that belongs to the lattice
    const Cell Z = alpha[n_valid - 1];           // alpha(T-1, S); zero when the band excludes the terminal state
    const bool feasible = Z.m > 0.0f;
    const float inv_mz = feasible ? 1.0f / Z.m : 0.0f;
    const int i_begin = static_cast<int>(static_cast<int64_t>(n) * part / parts);
    const int i_end = static_cast<int>(static_cast<int64_t>(n) * (part + 1) / parts);
    const float qnan = __int_as_float(0x7fc00000);
    Cell one;
    one.m = 1.0f;
    one.e = 0;
    Cell zero;
    zero.m = 0.0f;
    zero.e = kZeroExp;
    constexpr int U = 4;  // rows per thread per batch: 5*U independent loads in flight before any use
    for (int base = i_begin + tid; base < i_end; base += U * nthreads) {
        Cell al[U], b0[U], b1[U], b2[U];
        RawRow rw[U];
        int lab[U], tt[U], ss[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = min(base + u * nthreads, i_end - 1);  // clamped: loads are unconditional, stores are not
            const int t = i / ld;
            const int s = i - t * ld;
            tt[u] = t;
            ss[u] = s;
            // clamped into the lattice's own rows: whatever a clamped load returns is replaced or unused below
            al[u] = alpha[min(max(i - ld, 0), n_valid - 1)];
            b0[u] = beta[min(i, n_valid - 1)];
            b1[u] = beta[min(i + ld, n_valid - 1)];
            b2[u] = beta[min(i + ld + 1, n_valid - 1)];
            // K1's record with the denominator pair: garbage for dead rows, but only used when alpha(t-1,s) != 0, i.e.
            // when the row is live
            rw[u] = lp[min(i, n_valid - 1)];
            lab[u] = (Sb > 0) ? labels[min(s, Sb - 1)] : -1;
            if (rowmeta[i] == kRowDead) lab[u] = kRowDead;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = base + u * nthreads;
            if (i >= i_end) break;
            const int t = tt[u], s = ss[u];
            const bool inside = t < Tb && s <= Sb;
            const bool last = (t == Tb - 1);
            const Cell av = !inside ? zero : (t == 0) ? (s == 0 ? one : zero) : al[u];
            const Cell v1 = last ? (s == Sb ? one : zero) : b1[u];
            const Cell v2 = (s == Sb) ? zero : (last ? (s + 1 == Sb ? one : zero) : b2[u]);
            int lb = lab[u];
            // The plan's dead rows (alpha(t-1, s) outside the lattice) are marked as such: whoever zeroes them -- the
            // zero fill of this kernel or the gradient kernel -- a row whose coefficient merely came out as zero
            // (logits of -inf) is the gradient kernel's to write.
            if (lb != kRowDead && (s >= Sb || lb == a.blank || lb < 0 || lb >= a.V)) lb = -1;  // blank branch wins (cpu_rnnt.h:224-232)
            float4 c;  // (H, qb, ql, L)
            c.w = 0.0f;
            if (!inside || (feasible && av.m == 0.0f)) {
                c.x = c.y = c.z = kNegInfF;
            } else if (!feasible) {
                // infeasible utterance (e.g. the alignment band excludes the terminal state): cost = +inf and,
                // as in the reference, no finite gradient exists.  NaN is written on purpose.
                c.x = c.y = c.z = qnan;
            } else {
                const float ratio = av.m * inv_mz;
                const int base_e = av.e - Z.e;
                // The gradient kernel multiplies the logit by kLog2e alone; the denominator's -max * kLog2eLo
                // (lse_finish) stands for the missing x * kLog2eLo of the elements near the row's maximum -- the only
                // ones whose gradient is not negligible -- so it is taken out again here (max * kLog2e ~ -dh: off by
                // log2 sum <= 20, i.e. by < 3e-7 in the exponent).
                const float lo = fmaf(rw[u].dh, -kLog2eLoRel, rw[u].dl);  // dl + max * kLog2eLo
                k2_coef_main(b0[u], ratio, base_e, rw[u].dh, lo, c.x, c.w);
                c.y = k2_coef_patch(v1, ratio, base_e, rw[u].xb, rw[u].dh, rw[u].dl);
                c.z = lb >= 0 ? k2_coef_patch(v2, ratio, base_e, rw[u].xl, rw[u].dh, rw[u].dl) : kNegInfF;
            }
            coef[i] = c;
            rowlab[i] = lb;
        }
    }
}

// ---- cross-CTA hand-over of one utterance ----------------------------------------------------------------
// ---- dead-row zero fill (zero_fill.cuh) ----------------------------------------------------------------------
// done_word: the "recursions done" word of an utterance (SHARED protocol: this kernel's fill stops when it carries this
// launch's epoch -- what is left belongs to the gradient kernel's zero-fill warp, which is about to start)
__device__ __forceinline__ void k2_zero_dead_rows(const K2Args &a, int fw, int nfw, unsigned char *zbuf, const unsigned *done_word) {
    ZeroFill z;
    z.dst = a.zero_dst;
    z.rowmeta = a.rowmeta;
    z.rows = a.rows;
    z.row_bytes = a.row_bytes;
    z.unit_begin = 0;
    if (a.zero_shared_ctr != nullptr) {
        z.ctr = a.zero_shared_ctr;
        z.unit_end = -1;
        // (the word is looked at one unit late: the load is in flight while a unit is being written)
        unsigned seen = 0u;
        const unsigned epoch = a.epoch;
        zero_dead_rows_impl<false>(z, 0, 0, zbuf, [&] {
            const bool over = seen == epoch;
            seen = *reinterpret_cast<const volatile unsigned *>(done_word);
            return over;
        });
        return;
    }
    z.ctr = a.flags + k2_zero_ctr_word(a.B);
    z.unit_end = a.zero_unit_end;
    zero_dead_rows(z, fw, nfw, zbuf);
}

__device__ __forceinline__ unsigned k2_ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void k2_st_release(unsigned *p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void k2_await(const unsigned *flag, unsigned value) {
    while (k2_ld_acquire(flag) != value) __nanosleep(100);
}
__device__ __forceinline__ void k2_await_ge(const unsigned *flag, unsigned value) {
    while (k2_ld_acquire(flag) < value) __nanosleep(100);
}

// grid B * parts: CTA index = b * parts + part.  No CTA ever waits for a CTA that may not have been dispatched yet, so
// the kernel cannot deadlock whatever else shares the device (a second lattice launch on another stream, another
// process under MPS, fewer SMs than the occupancy query saw):
//   * the rows of phases A and C are cut into `parts` blocks; block p belongs to part p, which CLAIMS it (an atomic
//     exchange on the block's word) before it touches it.  Part 0 does block 0 without asking -- nobody else ever takes
//     it -- and then claims whatever has not been claimed yet and does that too: a helper that has not been dispatched
//     by then is not waited for, it finds its block taken and leaves.  What part 0 does wait for (phase A only) are the
//     blocks that WERE claimed by helpers -- CTAs that are therefore running;
//   * a helper waits for its own utterance's part 0 only, a CTA with a lower index (CTAs are dispatched in index order).
// The hand-over words come in two sets that take turns (plan.cuh: kK2FlagWords); this launch clears the other set.
// Barrier among the warps that run the phases (all of them, or all but the zero-fill warps -- those take part in
// nothing else: their stores wait on the memory system).
__device__ __forceinline__ void k2_phase_sync(int zero_warps) {
    if (zero_warps == 0) __syncthreads();
    else asm volatile("bar.sync 1, %0;" ::"r"((kK2Warps - zero_warps) * kWarp) : "memory");
}

// Part 0, after its own block: claim the blocks 1 .. parts-1 that nobody has claimed and do them (fn(block)).  Returns the
// number of blocks taken over.  `claim`: the phase's words, one per block.  Whole phase-warp set.
template <typename Fn>
__device__ __forceinline__ int k2_take_over_unclaimed(unsigned *claim, int parts, int tid, int zw, unsigned *ticket_sh, Fn fn) {
    if (tid >= 1 && tid < parts) ticket_sh[tid] = atomicExch(claim + tid, 1u);  // (all requests in flight together)
    k2_phase_sync(zw);
    int taken = 0;
    for (int blk = 1; blk < parts; ++blk) {
        if (ticket_sh[blk] == 0u) {  // (block-uniform)
            fn(blk);
            ++taken;
        }
    }
    return taken;
}

template <int K>
static __global__ void __launch_bounds__(kK2Threads) k2_lattice_kernel(K2Args a) {
    extern __shared__ __align__(128) unsigned char k2_smem[];
    __shared__ unsigned ticket_sh[kK2MaxParts];
    const int parts = a.parts;
    const int warp = threadIdx.x >> 5;
    const int zw = a.zero_warps;           // warps 0..zw-1 do the zero fill and nothing else
    const int nrw = a.row_warps;
    const size_t chunk_bytes = static_cast<size_t>(a.chunk_frames) * (a.ld > 0 ? a.ld : a.S_max + 1) * sizeof(Weight);
    const int bufs = a.chunk_bufs;
    const size_t dir_bytes = k2_dir_bytes(chunk_bytes, bufs, nrw);
    if (threadIdx.x == 0) MRNNT_TL_MIN(a.tl_slot, 3);
    pdl_launch_dependents();  // the gradient kernel may be scheduled as our CTAs retire (it waits for all of us)
    if (static_cast<int>(blockIdx.x) >= a.phase_ctas) {
        // a CTA that only fills zeros: needs nothing K1 wrote, so it does not wait for it either
        if (warp < zw)
            k2_zero_dead_rows(a, blockIdx.x * zw + warp, gridDim.x * zw, k2_smem + k2_zero_offset(dir_bytes),
                              a.flags + (static_cast<size_t>(a.epoch & 1u) * a.B + blockIdx.x % a.B) * kK2FlagWords + 2 * kK2MaxParts + 1);
        if (warp < zw && (threadIdx.x & 31) == 0) MRNNT_TL_MAX(a.tl_slot, 7);
        return;
    }
    const int b = blockIdx.x / parts;
    const int part = blockIdx.x - b * parts;
    const int tid = threadIdx.x - zw * kWarp;  // thread index among the phase warps; tid 0 leads
    unsigned *fl = a.flags + (static_cast<size_t>(a.epoch & 1u) * a.B + b) * kK2FlagWords;
    unsigned *a_claim = fl, *c_claim = fl + kK2MaxParts, *a_done = fl + 2 * kK2MaxParts, *done = a_done + 1;
    if (warp < zw) {
        // ---- zero fill: this warp's share of the batch's dead rows, nothing else (and nothing K1 wrote) ----
        if (warp == 0) MRNNT_K2_STAMP(53);
        k2_zero_dead_rows(a, blockIdx.x * zw + warp, gridDim.x * zw, k2_smem + k2_zero_offset(dir_bytes), done);
        if (warp == 0) MRNNT_K2_STAMP(54);
        if ((threadIdx.x & 31) == 0) MRNNT_TL_MAX(a.tl_slot, 7);
        return;
    }
    const int nphase = kK2Threads - zw * kWarp;
    MRNNT_K2_STAMP(0);
    pdl_wait();  // K1's records
    if (tid == 0) MRNNT_TL_MIN(a.tl_slot, 4);
    // (Nothing of the hand-over words is touched before this wait: under programmatic dependent launch this CTA can be
    // running while the PREVIOUS lattice launch has not yet cleared this set -- e.g. forward halves back to back with
    // small kernels that all fit on the device together.  Behind the wait the predecessor chain has completed.)
    if (blockIdx.x == 0) {
        if (tid == 0 && a.zero_clear != nullptr) *a.zero_clear = 0u;
        // the other set of hand-over words, for the next launch (its last user, the previous launch, is long gone)
        unsigned *other = a.flags + static_cast<size_t>((a.epoch + 1u) & 1u) * a.B * kK2FlagWords;
        for (int i = tid; i < a.B * kK2FlagWords; i += nphase) other[i] = 0u;
    }

    // ---- phase A: transition weights ----
    if (part == 0) {
        k2_weight_rows(a, b, 0, parts, tid, nphase);
    } else {
        if (tid == 0) ticket_sh[0] = atomicExch(a_claim + part, 1u);
        k2_phase_sync(zw);
        if (ticket_sh[0] == 0u) {
            k2_weight_rows(a, b, part, parts, tid, nphase);
            __threadfence();
            k2_phase_sync(zw);
            if (tid == 0) atomicAdd(a_done, 1u);
        }
    }

    unsigned c_mine = 1u;  // (tid 0 of a helper: what its claim of its phase-C block returned)
    if (part == 0) {
        // ---- phase B: the two recursions ----
        MRNNT_K2_STAMP(60);
        int taken = 0;
        if (parts > 1)
            taken = k2_take_over_unclaimed(a_claim, parts, tid, zw, ticket_sh,
                                           [&](int blk) { k2_weight_rows(a, b, blk, parts, tid, nphase); });
        const int ndir = a.need_beta ? 2 : 1;
        for (int d = 0; d < ndir; ++d) {
            K2Dir ring(k2_smem + d * dir_bytes, chunk_bytes, bufs);
            if (tid == 0) {
                for (int i = 0; i < bufs; ++i) {
                    mbar_init(ring.w_full + i, 1);
                    mbar_init(ring.empty + i, static_cast<uint32_t>(nrw));
                }
            }
            if (tid < kK2MaxRowWarps) ring.progress[tid] = kK2NoProgress;
        }
        if (tid == 0) {
            mbar_init_fence();
            // the blocks that helpers claimed: those CTAs are running
            if (parts > 1) k2_await_ge(a_done, static_cast<unsigned>(parts - 1 - taken));
        }
        k2_phase_sync(zw);  // (parts == 1: this is also what orders phase A before the bulk copies)
        MRNNT_K2_STAMP(61);
        // The SM sub-partition schedulers favour the highest warp id among the eligible warps of a partition
        // (warp id % 4): the chain warps get the highest ids, one per partition as far as possible.
        const int vw = kK2Warps - 1 - warp;
        if (vw < ndir * nrw) {
            const int d = vw / nrw, wi = vw - d * nrw;
            K2Dir ring(k2_smem + d * dir_bytes, chunk_bytes, bufs);
            if (d == 0) k2_chain_dispatch<K, 0>(a, b, ring, wi, nrw);
            else k2_chain_dispatch<K, 1>(a, b, ring, wi, nrw);
        } else if (vw < ndir * nrw + ndir) {
            const int d = vw - ndir * nrw;
            K2Dir ring(k2_smem + d * dir_bytes, chunk_bytes, bufs);
            if ((threadIdx.x & 31) == 0) {
                // the weights were written with ordinary stores (phase A); the bulk copies read them through the
                // async proxy
                asm volatile("fence.proxy.async;" ::: "memory");
                k2_issue_chunks(a, b, d, ring);
            }
        }
        if (!a.need_beta) return;
        k2_phase_sync(zw);
        if ((parts > 1 || a.zero_shared_ctr != nullptr) && tid == 0) {
            __threadfence();
            k2_st_release(done, a.epoch);
        }
        MRNNT_K2_STAMP(51);
        if (tid == 0) MRNNT_TL_MAX(a.tl_slot, 5);
    } else {
        if (!a.need_beta) return;
        // a helper claims its block of phase C while it waits for the recursions of its utterance's part 0
        if (tid == 0) {
            c_mine = atomicExch(c_claim + part, 1u);
            k2_await(done, a.epoch);
            ticket_sh[1] = c_mine;  // (not word 0: a thread may still be looking at phase A's answer there)
        }
        k2_phase_sync(zw);
    }
    // ---- phase C: gradient coefficients ----
    if (part == 0) {
        k2_coef_rows(a, b, 0, parts, tid, nphase);
        if (parts > 1)
            (void)k2_take_over_unclaimed(c_claim, parts, tid, zw, ticket_sh,
                                         [&](int blk) { k2_coef_rows(a, b, blk, parts, tid, nphase); });
    } else if (ticket_sh[1] == 0u) {
        k2_coef_rows(a, b, part, parts, tid, nphase);
    }
    k2_phase_sync(zw);
    MRNNT_K2_STAMP(52);
    if (tid == 0) MRNNT_TL_MAX(a.tl_slot, 6);
}

// ---- fallback for very long label sequences (more than 32 * 4 * kK2MaxRowWarps states) --------------------
// One CTA per utterance; states strided over the threads, previous frame in shared memory, one
// __syncthreads per frame.  Alpha first, then beta.  Slow path, same arithmetic: every thread converts the
// weights it needs on the fly and masks its own result.
static __global__ void __launch_bounds__(kK2Threads) k2_lattice_wide_kernel(K2Args a) {
    extern __shared__ __align__(16) unsigned char k2w_smem[];
    Cell *prev = reinterpret_cast<Cell *>(k2w_smem);  // [S_max + 2]
    const int b = blockIdx.x;
    const int Tb = a.T[b], Sb = a.S[b];
    const int W = k2_ld(a, Sb);  // rows per frame in memory
    const int64_t R = a.row_start[b];
    RawRow *lp = a.lp + R;
    const int2 *band = a.band + static_cast<size_t>(b) * a.T_max;
    Cell *alpha = a.alpha + R;
    Cell *beta = a.beta + R;
    const int tid = threadIdx.x;
    if (b == 0 && tid == 0 && a.zero_clear != nullptr) *a.zero_clear = 0u;
    if (b == 0) {  // (keeps the alternation of the hand-over words intact should launches of the two kernels ever mix)
        unsigned *other = a.flags + static_cast<size_t>((a.epoch + 1u) & 1u) * a.B * kK2FlagWords;
        for (int i = tid; i < a.B * kK2FlagWords; i += kK2Threads) other[i] = 0u;
    }

    // K1's (max, sum) -> the denominator pair, in place (what phase A of the main kernel does); rows K1 skipped
    // hold garbage before and after, and are masked wherever they are read
    for (int i = tid; i < Tb * (Sb + 1); i += kK2Threads) {
        const int t = i / (Sb + 1);
        RawRow &r = lp[static_cast<int64_t>(t) * W + (i - t * (Sb + 1))];
        const Denominator d = lse_finish(r.dh, r.dl);
        r.dh = d.hi;
        r.dl = d.lo;
    }
    __syncthreads();

    auto put = [](float m, int e) {
        Cell c;
        cell_renorm(m, e);
        c.m = m;
        c.e = e;
        return c;
    };
    auto row_live = [&](int t, int s) {
        return (t == 0) ? (s == 0) : k2_alpha_valid(t - 1, s, band[t - 1], Tb, Sb);
    };
    auto blank_weight = [&](int t, int s, float &m, int &e) {
        const bool live = row_live(t, s);
        const RawRow r = lp[static_cast<int64_t>(t) * W + s];
        weight_from_logit(r.xb, r.dh, r.dl, live, m, e);
    };
    auto label_weight = [&](int t, int s, float &m, int &e) {
        const bool live = s < Sb && row_live(t, s);
        const RawRow r = lp[static_cast<int64_t>(t) * W + s];
        weight_from_logit(r.xl, r.dh, r.dl, live, m, e);
    };

    for (int s = tid; s <= Sb + 1; s += kK2Threads) prev[s] = put(s == 0 ? 1.0f : 0.0f, 0);
    __syncthreads();
    for (int t = 0; t < Tb; ++t) {
        const int2 lh = band[t];
        for (int s = tid; s <= Sb; s += kK2Threads) {
            float wm, lm = 0.0f;
            int we, le = kZeroExp;
            blank_weight(t, s, wm, we);
            Cell below;
            below.m = 0.0f;
            below.e = kZeroExp;
            if (s > 0) {
                label_weight(t, s - 1, lm, le);
                below = prev[s - 1];
            }
            const Cell stay = prev[s];
            const Cell r = cell_add(below.m * lm, below.e + le, stay.m * wm, stay.e + we);
            alpha[static_cast<int64_t>(t) * W + s] = put(k2_alpha_valid(t, s, lh, Tb, Sb) ? r.m : 0.0f, r.e);
        }
        __syncthreads();
        for (int s = tid; s <= Sb; s += kK2Threads) prev[s] = alpha[static_cast<int64_t>(t) * W + s];
        __syncthreads();
    }
    if (tid == 0) {
        const double ll = cell_log(prev[Sb].m, prev[Sb].e);
        a.ll_fwd[b] = ll;
        a.costs[b] = static_cast<float>(-ll);
        if (a.costs_mapped != nullptr) a.costs_mapped[b] = static_cast<float>(-ll);
    }
    if (!a.need_beta) return;
    __syncthreads();
    for (int s = tid; s <= Sb + 1; s += kK2Threads) prev[s] = put(s == Sb ? 1.0f : 0.0f, 0);
    __syncthreads();
    for (int t = Tb - 1; t >= 0; --t) {
        for (int s = tid; s <= Sb; s += kK2Threads) {
            float wm, lm;
            int we, le;
            blank_weight(t, s, wm, we);  // zero for a dead row: beta(t, s) is then outside its band
            label_weight(t, s, lm, le);
            const Cell stay = prev[s];
            const Cell above = prev[s + 1];  // prev[Sb + 1] = 0
            const Cell r = cell_add(above.m * lm, above.e + le, stay.m * wm, stay.e + we);
            beta[static_cast<int64_t>(t) * W + s] = put(r.m, r.e);
        }
        __syncthreads();
        for (int s = tid; s <= Sb; s += kK2Threads) prev[s] = beta[static_cast<int64_t>(t) * W + s];
        __syncthreads();
    }
    if (tid == 0) a.ll_bwd[b] = cell_log(prev[0].m, prev[0].e);
    __syncthreads();
    k2_coef_rows(a, b, 0, 1);
}

}  // namespace mrnnt
