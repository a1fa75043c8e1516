// K1 -- one pass over the packed logits: per-row log-softmax denominator plus the gather of the
// two log-probabilities the lattice needs (blank, and the label leaving state s).
//
// Replaces reduce_max + reduce_exp (reference include/reduce.h:79-139, gpu_rnnt.h:242-248: two full
// reads of the logits with 4-byte loads and two stream syncs) and the strided `log_p` gathers inside
// the alpha/beta kernels (gpu_rnnt_kernel.h:80-84,144-149).  CPU twin: cpu_rnnt.h:98-115.
//
//   ML = max_v x[v] * log2 e,  sum = sum_v 2^(x[v] log2 e - ML)   (the reference's denominator, reduce.h:112-139, is
//                                                                 -(ML + log2 sum) ln 2; K2's phase A forms it)
//   lp[row] = (x[blank], x[label(b,s)], ML, sum)                  one 16-byte record per live row
//
// Streaming design (HBM-bound; algorithmic bytes = 4*V per live row, nothing for dead rows):
//   * persistent CTAs, one per SM; the flat row space is cut into tiles of G consecutive rows;
//   * a producer warp stages each tile's LIVE rows into a shared-memory ring with 1-D bulk async
//     copies (TMA engine, cp.async.bulk / SASS UBLKCP) that complete on an mbarrier;
//   * NW consumer warps take one row each.  For V <= 2048 the row is pulled into registers with
//     128-bit shared loads once; the exact row max comes from one redux.sync on order-preserving
//     integer keys, the sum of 2^((x - max) log2 e) costs one FFMA + one MUFU.EX2 per element with no
//     branches, and five shuffles finish the sum.  Larger rows make two passes over shared memory.
// A generic variant (direct global loads, any V / alignment) covers V % 4 != 0, unaligned bases and
// rows too large for the ring.
#pragma once

#include "common.cuh"
#include "zero_fill.cuh"

namespace mrnnt {

// max over / exp-sum of the NE floats of one unpacked 16-byte vector
template <int NE>
__device__ __forceinline__ float vec_max(const float (&f)[NE]) {
    float m = fmaxf(f[0], f[1]);
#pragma unroll
    for (int i = 2; i < NE; i += 2) m = fmaxf(m, fmaxf(f[i], f[i + 1]));
    return m;
}

// Warp-wide float max with ONE redux.sync: floats are mapped to signed ints that sort the same way.
__device__ __forceinline__ float warp_max_redux(float v) {
    int k = __float_as_int(v);
    k ^= (k >> 31) & 0x7fffffff;
    k = __reduce_max_sync(0xffffffffu, k);
    k ^= (k >> 31) & 0x7fffffff;
    return __int_as_float(k);
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// four running sums (two packed pairs) of 2^(x*log2e + neg) over the NE floats of a vector; the multiply-adds
// and the additions are the packed two-float instructions of sm_100 (FFMA2 / FADD2), the exponentials MUFU.EX2
template <int NE>
__device__ __forceinline__ void exp_acc(const float (&f)[NE], float neg, float2 (&s)[2]) {
    const float2 l2 = make_float2(kLog2e, kLog2e), n2 = make_float2(neg, neg);
#pragma unroll
    for (int i = 0; i < NE; i += 2) {
        const float2 t = __ffma2_rn(make_float2(f[i], f[i + 1]), l2, n2);
        s[(i >> 1) & 1] = __fadd2_rn(s[(i >> 1) & 1], make_float2(ex2_approx(t.x), ex2_approx(t.y)));
    }
}
// the same with the lanes whose vector lies beyond the row masked out (their registers hold a copy of the row's
// last vector, see row_sums)
template <int NE>
__device__ __forceinline__ void exp_acc_masked(const float (&f)[NE], float neg, bool valid, float2 (&s)[2]) {
    const float2 l2 = make_float2(kLog2e, kLog2e), n2 = make_float2(neg, neg);
#pragma unroll
    for (int i = 0; i < NE; i += 2) {
        const float2 t = __ffma2_rn(make_float2(f[i], f[i + 1]), l2, n2);
        const float2 e2 = make_float2(valid ? ex2_approx(t.x) : 0.0f, valid ? ex2_approx(t.y) : 0.0f);
        s[(i >> 1) & 1] = __fadd2_rn(s[(i >> 1) & 1], e2);
    }
}

// What K1 knows about a row's denominator: the warp-wide max (times log2 e, rounded once and used for every term)
// and the sum of 2^(x log2e - ML).  Turning them into -log2 sum_v exp(x[v]) costs a logarithm and a few
// error-free additions per row; that is left to the lattice kernel's helper CTAs (k2_lattice.cuh, phase A) -- K1
// at V ~ 1000 is bound by the instruction stream of its consumer warps, not by HBM (tools/k1_probe.cu).
struct RowSum {
    float ML, sum;
};

// A row whose bytes do not start / end on a 16-byte boundary (V % kPerVec != 0: UNALIGNED) is split into its INTERIOR --
// the aligned 16-byte vectors that lie wholly inside it, which go through exactly the code of an aligned row -- and up to
// 2 * (kPerVec - 1) EDGE elements before and behind the interior, which lanes 0, 1, ... take one each (a scalar load;
// -inf for the lanes that have none: no effect on the max, exactly zero in the sum).  The streaming kernels are bound by
// their consumers' instruction stream at V ~ 1000, so the geometry is 32-bit arithmetic and the edge costs a dozen
// instructions per row.  (Masking inside the vector loop costs eight predicated instructions on EVERY vector.)
struct RowSplit {
    unsigned off;    // byte offset of the row's first element in its tile's slot (the slot mirrors global memory from the
                     // 16-byte boundary at or below the tile's first byte, see StreamWindow)
    unsigned vec0;   // index (in 16-byte vectors from the slot's start) of the first interior vector
    int ninterior;   // interior vectors
    int nhead;       // edge elements before the interior: the row's elements [0, nhead)
    int ntail;       // edge elements behind it: the row's elements [V - ntail, V)
};
template <typename E>
__device__ __forceinline__ RowSplit row_split(int64_t tile_row0, int r, int V) {
    const unsigned rb = static_cast<unsigned>(V) * static_cast<unsigned>(sizeof(E));
    // (tile_row0 * rb) mod 16 from the two factors mod 16; the row's offset in the slot stays below 32 rows' bytes
    const unsigned t16 = ((static_cast<unsigned>(tile_row0) & 15u) * (rb & 15u)) & 15u;
    RowSplit w;
    w.off = t16 + static_cast<unsigned>(r) * rb;
    w.vec0 = (w.off + 15u) >> 4;
    const unsigned end = w.off + rb;
    w.ninterior = static_cast<int>((end >> 4) - w.vec0);
    w.nhead = static_cast<int>((w.vec0 * 16u - w.off) / sizeof(E));
    w.ntail = static_cast<int>((end & 15u) / sizeof(E));
    return w;
}
// which element of the row this lane takes as its edge element (-1: none)
__device__ __forceinline__ int edge_element(const RowSplit &w, int V, int lane) {
    const int e = lane < w.nhead ? lane : V - w.ntail + (lane - w.nhead);
    return lane < w.nhead + w.ntail ? e : -1;
}

// One row resident in shared memory: NV >= 1 aligned vectors of 16 bytes = NV * Elem<E>::kPerVec logits, all of them
// the row's own (an unaligned row: its interior; `edge` is then the lane's edge element, -inf where it has none).
// C > 0: the lane's <= C vectors live in registers, unpacked (NV <= 32*C); C == 0: two passes over shared
// memory, any NV.  Vector slots beyond the row: whole slots (c*32 >= NV, warp-uniform) are skipped; in the one
// partly filled slot the surplus lanes load the row's last vector again -- harmless for the max, masked in the sum.
template <typename E, int C, bool EDGE = false>
__device__ __forceinline__ RowSum row_sums(const uint4 *__restrict__ xv, int NV, int lane, float edge = kNegInfF) {
    constexpr int NE = Elem<E>::kPerVec;
    float2 s[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    float ML;
    if constexpr (C > 0) {
        float f[C][NE];
        Elem<E>::unpack(xv[min(lane, NV - 1)], f[0]);
        float m = EDGE ? fmaxf(edge, vec_max<NE>(f[0])) : vec_max<NE>(f[0]);
#pragma unroll
        for (int c = 1; c < C; ++c) {
            if (c * kWarp < NV) {
                Elem<E>::unpack(xv[min(lane + c * kWarp, NV - 1)], f[c]);
                m = fmaxf(m, vec_max<NE>(f[c]));
            }
        }
        ML = warp_max_redux(m) * kLog2e;
        const float neg = -ML;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            if ((c + 1) * kWarp <= NV) {
                exp_acc<NE>(f[c], neg, s);
            } else if (c * kWarp < NV) {
                exp_acc_masked<NE>(f[c], neg, lane + c * kWarp < NV, s);
            }
        }
        if constexpr (EDGE) s[0].x += ex2_approx(fmaf(edge, kLog2e, neg));  // (2^-inf = 0 for a lane without an edge element)
    } else {
        float m = edge;
#pragma unroll 4
        for (int j = lane; j < NV; j += kWarp) {
            float f[NE];
            Elem<E>::unpack(xv[j], f);
            m = fmaxf(m, vec_max<NE>(f));
        }
        ML = warp_max_redux(m) * kLog2e;
        const float neg = -ML;
#pragma unroll 4
        for (int j = lane; j < NV; j += kWarp) {
            float f[NE];
            Elem<E>::unpack(xv[j], f);
            exp_acc<NE>(f, neg, s);
        }
        if constexpr (EDGE) s[0].x += ex2_approx(fmaf(edge, kLog2e, neg));
    }
    RowSum r;
    r.ML = ML;
    r.sum = warp_sum_f((s[0].x + s[0].y) + (s[1].x + s[1].y));
    return r;
}

// What lane 0 writes for one live row: the two gathered logits and the row's (max, sum) (one 16-byte store).  The
// lattice kernel's phase A turns them into the denominator and the transition weights (k2_lattice.cuh); dead rows
// are never read again, so nothing is written for them.
__device__ __forceinline__ void k1_store_row(RawRow *__restrict__ lp, int64_t row, const RowSum &rs, float x_blank,
                                             float x_label) {
    RawRow r;
    r.xb = x_blank;
    r.xl = x_label;
    r.dh = rs.ML;
    r.dl = rs.sum;
    lp[row] = r;
}

// ---------------------------------------------------------------------------------------------
// Generic variant: one warp per row, scalar global loads (second pass hits L1).  Any V, any alignment.
// ---------------------------------------------------------------------------------------------
constexpr int kGenericWarps = 8;

template <typename E>
static __global__ void __launch_bounds__(kGenericWarps * kWarp)
    k1_lse_generic_kernel(const E *__restrict__ acts, const int *__restrict__ labels,
                          const int *__restrict__ rowmeta, RawRow *__restrict__ lp, int64_t rows, int V, int blank) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * kGenericWarps + (threadIdx.x >> 5);
    const int64_t nwarps = static_cast<int64_t>(gridDim.x) * kGenericWarps;
    for (int64_t row = warp0; row < rows; row += nwarps) {
        const int meta = rowmeta[row];
        if (meta == kRowDead) continue;
        const E *x = acts + row * V;
        float m = kNegInfF;
        for (int v = lane; v < V; v += kWarp) m = fmaxf(m, Elem<E>::to_float(x[v]));
        const float ML = warp_max_redux(m) * kLog2e;
        float s = 0.f;
        for (int v = lane; v < V; v += kWarp) s += ex2_approx(fmaf(Elem<E>::to_float(x[v]), kLog2e, -ML));
        RowSum den;
        den.ML = ML;
        den.sum = warp_sum_f(s);
        if (lane == 0) {
            const int lab = meta >= 0 ? __ldg(labels + meta) : -1;
            const bool has = lab >= 0 && lab < V;
            k1_store_row(lp, row, den, Elem<E>::to_float(x[blank]), has ? Elem<E>::to_float(x[lab]) : kNegInfF);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// TMA-staged variant.  Requirements (checked on the host): rows are whole 16-byte vectors (V % 4 == 0 for
// float, V % 8 == 0 for bfloat16), acts 16-byte aligned, a tile fits a ring slot.
// Shared memory: [stages][G*V] elements | full[stages] | empty[stages] | (meta, tile)[stages][32]
// ---------------------------------------------------------------------------------------------
#ifdef MRNNT_K1_TRACE  // development aid (tools/k1_probe.cu): where the warps of CTA 0 spend their cycles
__device__ long long g_k1_trace[32][4];  // per warp: cycles waiting, cycles working, rows/tiles, -
#define MRNNT_K1_CLOCK() clock64()
#define MRNNT_K1_ADD(w, i, v) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) g_k1_trace[w][i] += (v); } while (0)
#else
#define MRNNT_K1_CLOCK() 0ll
#define MRNNT_K1_ADD(w, i, v) do { } while (0)
#endif

// Dynamic tile hand-out (the gradient kernel's producer, kK3Dynamic in k3_grad.cuh): tiles come from a counter,
// `kGrabDepth` requests in flight per producer -- an atomic on one hot word takes a microsecond or two to come back while
// the GPU streams at full bandwidth, a tile lasts about one; every request in flight is also a tile bound to its CTA
// before the CTA gets to it, i.e. a hand-out that follows the SMs less closely (tools/gpu_grab.sh, K3 on c2 in the
// stream: depth 1..4 189-190 us, 8 193, 16 197, 32 230; with one request the call as a whole was slower).  The first two tiles of a CTA are fixed (blockIdx.x, blockIdx.x + gridDim.x), the counter hands
// out the ones from 2 * gridDim.x on, in order: the CTAs keep working on one moving window of the input.
#ifndef MRNNT_GRAB_DEPTH
#define MRNNT_GRAB_DEPTH 3
#endif
constexpr int kGrabDepth = MRNNT_GRAB_DEPTH;
struct TileGrabber {
    unsigned raw[kGrabDepth];
    unsigned *ctr;  // {tiles handed out, producers finished}; zero between launches
    int lane;
    int64_t first;  // the counter's tile 0: the CTAs' fixed tiles come before it
    __device__ __forceinline__ unsigned grab() { return lane == 0 ? atomicAdd(ctr, 1u) : 0u; }
    __device__ __forceinline__ void start(unsigned *c, int l, int fixed_per_cta = 2) {
        ctr = c;
        lane = l;
        first = static_cast<int64_t>(fixed_per_cta) * gridDim.x;
#pragma unroll
        for (int i = 0; i < kGrabDepth; ++i) raw[i] = grab();
    }
    // the next tile (warp-uniform), and one more request
    __device__ __forceinline__ int64_t next() {
        const int64_t t = static_cast<int64_t>(__shfl_sync(0xffffffffu, raw[0], 0)) + first;
#pragma unroll
        for (int i = 0; i + 1 < kGrabDepth; ++i) raw[i] = raw[i + 1];
        raw[kGrabDepth - 1] = grab();
        return t;
    }
    // The last producer to finish leaves the counters as it found them (every request of this launch has been
    // answered by then: a producer looks at its own outstanding ones before it reports).
    __device__ __forceinline__ void finish() {
        unsigned seen = 0u;
#pragma unroll
        for (int i = 0; i < kGrabDepth; ++i) seen |= __shfl_sync(0xffffffffu, raw[i], 0);
        // (the report must not be issued before the answers have arrived: the asm turns `seen` -- and with it the
        // outstanding atomics' results -- into a zero the compiler cannot fold away, and the fence orders lane 0's
        // grabs before its report whatever the compiler does)
        asm volatile("and.b32 %0, %0, 0;" : "+r"(seen));
        __threadfence();
        if (lane == 0 && atomicAdd(ctr + 1, seen + 1u) == gridDim.x - 1u) {
            ctr[0] = 0u;
            ctr[1] = 0u;
            __threadfence();
        }
    }
};

struct StreamTiling {
    int G = 0;       // rows per tile: a power of two in 1..32
    int stages = 0;  // ring depth
    int warps = 0;   // consumer warps (8 or 16)
    size_t slot_bytes = 0;  // data bytes of one ring slot: G rows, plus -- when rows are not whole 16-byte vectors -- the
                            // slack around the aligned window that covers the tile (stream_window)
    bool unaligned = false; // V * sizeof(element) is not a multiple of 16
    size_t smem_bytes = 0;
};

// Rows that are not whole 16-byte vectors (V % 4 != 0 for float): a bulk copy wants 16-byte aligned addresses and sizes,
// so what is copied for a run of rows [b0, b1) (bytes of the whole array) is the aligned window that covers it, and a ring
// slot mirrors global memory from the aligned address at or below its tile's first byte: slot byte k <-> global byte
// A0 + k, A0 = tile_first_byte & ~15.  The consumers read (and the gradient kernel writes) a row through the aligned
// vectors that cover it and mask the few elements at the two ends that are not the row's (mask_row_edges).
struct StreamWindow {
    size_t slot_off;  // offset of the window in the slot
    size_t g_off;     // offset of the window in the array
    uint32_t bytes;   // multiple of 16
    uint32_t tail;    // bytes of the run beyond the window, at the very end of the array (< 16; copied by hand)
};
__device__ __forceinline__ StreamWindow stream_window(size_t tile_b0, size_t b0, size_t b1, size_t total_bytes) {
    const size_t a0 = tile_b0 & ~static_cast<size_t>(15);
    const size_t w0 = b0 & ~static_cast<size_t>(15);
    size_t w1 = (b1 + 15) & ~static_cast<size_t>(15);
    StreamWindow w;
    w.tail = 0;
    if (w1 > total_bytes) {  // only the array's last row can get here: never read behind the caller's buffer
        w1 = total_bytes & ~static_cast<size_t>(15);
        w.tail = static_cast<uint32_t>(b1 - (w1 > b0 ? w1 : b0));
        if (w1 < w0) w1 = w0;
    }
    w.slot_off = w0 - a0;
    w.g_off = w0;
    w.bytes = static_cast<uint32_t>(w1 - w0);
    return w;
}

constexpr int kStreamMaxStages = 12;
constexpr size_t kStreamSmemBudget = 200 * 1024;  // of the 227 KB a CTA may use
// Bytes per ring slot we aim for.  Measured on B200 (c2/c3, float and bfloat16): K1 gains from larger tiles (the
// cost of a tile hand-over is per tile, not per byte: 8 KB tiles are 2x slower than 32 KB ones, 64 KB ones up to
// 6 % faster on c3 and 17 % faster on bfloat16 rows), K3 is best at 32 KB.
constexpr int kK1TileTarget = 64 * 1024;
constexpr int kK1TileTargetSmall = 32 * 1024;            // ... except on small inputs (c2: 100 us against 103)
constexpr size_t kK1SmallInputBytes = size_t(2) << 30;  // "small": less than 2 GiB of logits
constexpr int kK3TileTarget = 32 * 1024;

// Tiling of the streaming kernels for vocabulary size V with `warps` consumer warps.
// extra_per_row: additional per-row shared bytes a kernel keeps next to the tile (K3: its coefficients).
//
// Ring-safety rule: consumer warp w handles the rows q = w, w+NW, ... of the CTA's row sequence.  When
// NW > G a warp only touches every (NW/G)-th tile; a parity wait on an mbarrier is only meaningful if the
// waiting warp also consumed the PREVIOUS use of that ring stage (bulk copies complete out of order, so
// "an earlier tile was issued first" proves nothing).  Hence G is a power of two and the ring depth is a
// multiple of the period NW/gcd(NW,G) of a warp's tile pattern: every warp then visits a fixed subset of the
// stages, each of them at every one of its uses.
// whole_tiles_per_warp_set: only accept warp counts that are a multiple of G when they exceed it (K3 measured
// slower with 24 warps on 16-row tiles than with 16)
inline bool stream_tiling(int V, size_t elem_bytes, size_t extra_per_row, int warps, int tile_target,
                          bool whole_tiles_per_warp_set, StreamTiling *out) {
    const size_t row_bytes = static_cast<size_t>(V) * elem_bytes;
    if (V <= 0 || (warps != 8 && warps != 16 && warps != 24)) return false;
    const bool unaligned = (row_bytes % 16) != 0;
    if (unaligned && row_bytes < 64) return false;  // (windows of neighbouring runs must not overlap; tiny rows: generic kernels)
    // ring depth for tiles of G rows: what fits the budget, rounded down to the period of a warp's tile pattern (warp w's
    // uses repeat with period lcm(NW, G) rows = NW / gcd(NW, G) tiles; 24 warps on 16-row tiles: 3); 0: no safe ring
    auto depth_for = [&](int G, size_t *data_bytes) -> int {
        // unaligned rows: the aligned window around a tile is up to 15 bytes longer at either end
        const size_t data = unaligned ? (static_cast<size_t>(G) * row_bytes + 15) / 16 * 16 + 32 : static_cast<size_t>(G) * row_bytes;
        const size_t slot = data + 32 * (sizeof(int) + extra_per_row) + 16;
        int stages = static_cast<int>(kStreamSmemBudget / slot);
        if (stages > kStreamMaxStages) stages = kStreamMaxStages;
        int g = warps, h = G;
        while (h != 0) {
            const int t = g % h;
            g = h;
            h = t;
        }
        if (whole_tiles_per_warp_set && warps > G && warps % G != 0) return 0;
        const int stride = warps > G ? warps / g : 1;
        stages = stages / stride * stride;
        *data_bytes = data;
        return stages < 3 ? 0 : stages;
    };
    int G = 1;
    while (G < 32 && static_cast<size_t>(2 * G) * row_bytes <= static_cast<size_t>(tile_target)) G *= 2;
    size_t data = 0;
    int stages = depth_for(G, &data);
    // A vocabulary just above a power of two (V = 1025: 4100-byte rows) would halve the tile (the cost of a slot hand-over
    // is per tile, not per byte: 16 KB tiles stream ~40 % slower than 32 KB ones): take the next size up when it is within
    // a quarter of the target and the ring stays as deep as the consumer warps need (two tiles' worth of rows ahead).
    if (G < 32 && static_cast<size_t>(2 * G) * row_bytes <= static_cast<size_t>(tile_target) + static_cast<size_t>(tile_target) / 4) {
        size_t data2 = 0;
        const int stages2 = depth_for(2 * G, &data2);
        if (stages2 >= 6 || (stages2 >= 3 && stages2 >= stages)) {
            G *= 2;
            stages = stages2;
            data = data2;
        }
    }
    if (stages == 0) return false;
    const size_t slot = data + 32 * (sizeof(int) + extra_per_row) + 16;
    out->G = G;
    out->stages = stages;
    out->warps = warps;
    out->slot_bytes = data;
    out->unaligned = unaligned;
    out->smem_bytes = static_cast<size_t>(stages) * slot + 128;
    return true;
}

// Issue one bulk copy per maximal run of live rows of a tile (called by one lane).
template <typename E>
__device__ __forceinline__ void issue_live_runs(uint32_t mask, E *tile, const E *src_row0, int V, uint64_t *bar,
                                                uint64_t policy) {
    const uint32_t row_bytes = static_cast<uint32_t>(V) * static_cast<uint32_t>(sizeof(E));
    while (mask) {
        const int r0 = __ffs(mask) - 1;
        const uint32_t inv = ~(mask >> r0);
        const int len = inv ? (__ffs(inv) - 1) : 32;
        bulk_g2s_hint(tile + static_cast<size_t>(r0) * V, src_row0 + static_cast<size_t>(r0) * V,
                      static_cast<uint32_t>(len) * row_bytes, bar, policy);
        mask = (len >= 32) ? 0u : (mask & ~(((1u << len) - 1u) << r0));
    }
}

// The same for rows that are not whole 16-byte vectors: per run the aligned window that covers it (stream_window).
// Called by the WHOLE producer warp (the address arithmetic of a window is a few dozen 64-bit instructions; done by one
// lane for every run of a tile it was what the ring waited for): lane r computes and issues the run that starts at row
// r of the tile, lane 0 first posts the byte count of all of them.  `slot` mirrors the array from (row0 * row_bytes) & ~15 on.
template <typename E>
__device__ __forceinline__ void issue_live_runs_unaligned(uint32_t mask, unsigned char *slot, const E *acts, int64_t row0,
                                                          int V, int64_t rows, uint64_t *bar, uint64_t policy, int lane) {
    const size_t rb = static_cast<size_t>(V) * sizeof(E);
    const size_t total = static_cast<size_t>(rows) * rb;
    const size_t tile_b0 = static_cast<size_t>(row0) * rb;
    const unsigned char *g = reinterpret_cast<const unsigned char *>(acts);
    // this lane's row starts a run of live rows?
    const bool starts = ((mask >> lane) & 1u) != 0u && (lane == 0 || ((mask >> (lane - 1)) & 1u) == 0u);
    StreamWindow w;
    w.bytes = 0;
    w.tail = 0;
    size_t b1 = 0;
    if (starts) {
        const uint32_t inv = ~(mask >> lane);  // (zeros shifted in: inv != 0 unless lane == 0 and all 32 rows are live)
        const int len = inv ? (__ffs(inv) - 1) : 32;
        b1 = tile_b0 + (lane + len) * rb;
        w = stream_window(tile_b0, tile_b0 + lane * rb, b1, total);
        if (w.tail != 0) {  // the array's last bytes, behind the last whole 16-byte vector: by hand, element by element
            const size_t a0 = tile_b0 & ~static_cast<size_t>(15);
            for (size_t b = b1 - w.tail; b < b1; b += sizeof(E))
                *reinterpret_cast<E *>(slot + (b - a0)) = *reinterpret_cast<const E *>(g + b);
        }
    }
    uint32_t tx = w.bytes;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tx += __shfl_xor_sync(0xffffffffu, tx, o);
    __syncwarp();  // (the hand-copied tail is in place before the arrival that publishes the slot)
    if (lane == 0) mbar_arrive_expect_tx(bar, tx);
    __syncwarp();
    if (w.bytes != 0) bulk_g2s_hint(slot + w.slot_off, g + w.g_off, w.bytes, bar, policy);
}

// COMPACT: tiles without a live row take no ring slot (the slots carry their tile's index and the consumers walk
// the slots, not the tiles; the producer ends the sequence with one terminator per slot).  Worth it when many
// tiles are dead -- alignment-restricted lattices, padded inputs -- and costs ~2 % on dense inputs (the end of the
// kernel waits for the terminators), hence a compile-time choice made by the engine.
// The COMPACT variant can carry one more warp that zeroes dead rows of the GRADIENT while this kernel runs (`zero`,
// zero_fill.cuh, SHARED protocol: the gradient kernel's own zero-fill warp continues where this one stops): under an
// alignment band this kernel leaves the memory system nearly idle, and the zeros are most of the call's traffic.
inline size_t k1_smem_bytes(size_t ring_bytes, bool zero_warp) {
    return zero_warp ? (ring_bytes + 127) / 128 * 128 + kZeroFillBytes + 16 : ring_bytes;
}

// UNALIGNED: rows are not whole 16-byte vectors (StreamWindow above); slot_bytes = StreamTiling::slot_bytes.
template <typename E, int NW, int C, bool COMPACT, bool UNALIGNED = false>
static __global__ void __launch_bounds__((NW + (COMPACT ? 2 : 1)) * kWarp, 1)
    k1_lse_tma_kernel(const E *__restrict__ acts, const int *__restrict__ labels,
                      const int *__restrict__ rowmeta, RawRow *__restrict__ lp, int64_t rows, int V, int blank,
                      int G, int stages, ZeroFill zero, size_t ring_bytes, size_t slot_bytes) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *tiles = smem_raw;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + static_cast<size_t>(stages) * slot_bytes);
    uint64_t *empty = full + stages;
    // per slot and row: (rowmeta, which of this CTA's tiles the slot holds; -1: no more tiles) -- one 8-byte load
    int2 *meta_sh = reinterpret_cast<int2 *>(empty + stages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *zbuf = smem_raw + (ring_bytes + 127) / 128 * 128;          // (only there with a zero-fill warp)
    int *consumers_done = reinterpret_cast<int *>(zbuf + kZeroFillBytes);
    const bool zero_warp = COMPACT && zero.dst != nullptr;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, static_cast<uint32_t>(G));
        }
        mbar_init_fence();
        if (zero_warp) *consumers_done = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) MRNNT_TL_MIN(zero.tl_slot, 0);
    pdl_launch_dependents();  // the lattice kernel may be scheduled as our CTAs retire (it waits for all of us)
    pdl_wait();               // (first kernel of a call: its predecessor is the previous call or a set-up kernel)
    if (threadIdx.x == 0) MRNNT_TL_MIN(zero.tl_slot, 1);
    if (COMPACT && warp == NW + 1) {
        // ---------------- zero-fill warp: for as long as this CTA's consumer warps have work ----------------
        zero_dead_rows_impl<false>(zero, 0, 0, zbuf,
                                   [&] { return *reinterpret_cast<volatile int *>(consumers_done) >= NW; });
        return;
    }

    const int64_t ntiles = (rows + G - 1) / G;
    const int64_t nloc = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == NW) {
        // ---------------- producer warp ----------------
        const uint64_t policy = l2_policy_evict_first();
        // rowmeta is fetched for 32/G tiles per load (lane l: tile k0 + l/G, row l%G) and two such batches ahead:
        // with mostly dead tiles the producer does nothing but wait for these loads
        const int TPB = 32 / G;  // tiles per batch (G is a power of two <= 32)
        const uint32_t gmask = G == 32 ? 0xffffffffu : ((1u << G) - 1u);
        auto load_batch = [&](int64_t k0) {
            const int64_t k = k0 + lane / G;
            const int64_t row = (blockIdx.x + k * gridDim.x) * G + (lane % G);
            return (k < nloc && row < rows) ? __ldg(rowmeta + row) : kRowDead;
        };
        int stage = 0;
        uint32_t phase = 0;
        int mb0 = load_batch(0), mb1 = load_batch(TPB);
        for (int64_t k0 = 0; k0 < nloc; k0 += TPB) {
            const int mb = mb0;
            mb0 = mb1;
            mb1 = load_batch(k0 + 2 * TPB);
            const uint32_t ball = __ballot_sync(0xffffffffu, mb != kRowDead);
            for (int j = 0; j < TPB && k0 + j < nloc; ++j) {
                const int64_t k = k0 + j;
                const uint32_t mask = (ball >> (j * G)) & gmask;
                // COMPACT: a tile without a single live row takes no ring slot at all (a slot hand-over costs about
                // as much as streaming 8 KB)
                if (COMPACT && mask == 0u) continue;
                const int m = __shfl_sync(0xffffffffu, mb, j * G + (lane % G));
                const int64_t row0 = (blockIdx.x + k * gridDim.x) * G;
                [[maybe_unused]] const long long tw0 = MRNNT_K1_CLOCK();
                mbar_wait(empty + stage, phase ^ 1u);
                MRNNT_K1_ADD(NW, 0, MRNNT_K1_CLOCK() - tw0);
                MRNNT_K1_ADD(NW, 2, 1);
                meta_sh[stage * 32 + lane] = make_int2(lane < G ? m : kRowDead, static_cast<int>(k));
                __syncwarp();
                if constexpr (UNALIGNED) {
                    issue_live_runs_unaligned<E>(mask, tiles + stage * slot_bytes, acts, row0, V, rows, full + stage, policy, lane);
                } else if (lane == 0) {
                    mbar_arrive_expect_tx(full + stage, static_cast<uint32_t>(__popc(mask)) * static_cast<uint32_t>(V) *
                                                            static_cast<uint32_t>(sizeof(E)));
                    issue_live_runs<E>(mask, reinterpret_cast<E *>(tiles + stage * slot_bytes), acts + row0 * V, V,
                                       full + stage, policy);
                }
                if (++stage == stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
        // one terminator per slot: every consumer warp meets one within its next `stages` slot uses
        for (int i = 0; COMPACT && i < stages; ++i) {
            mbar_wait(empty + stage, phase ^ 1u);
            meta_sh[stage * 32 + lane] = make_int2(kRowDead, -1);
            __syncwarp();
            if (lane == 0) mbar_arrive_expect_tx(full + stage, 0u);
            if (++stage == stages) {
                stage = 0;
                phase ^= 1u;
            }
        }
    } else {
        // ---------------- consumer warps: one row at a time ----------------
        // q walks the rows of the slot uses u = 0, 1, ... in order (G rows each); which tile a use holds is read
        // from the slot
        const int NV = V / Elem<E>::kPerVec;
        const int64_t nq = nloc * G;
        for (int64_t q = warp; COMPACT || q < nq; q += NW) {
            const int64_t u = q / G;
            const int r = static_cast<int>(q - u * G);
            const int stage = static_cast<int>(u % stages);
            const uint32_t phase = static_cast<uint32_t>((u / stages) & 1);
            [[maybe_unused]] const long long tw0 = MRNNT_K1_CLOCK();
            mbar_wait(full + stage, phase);
            [[maybe_unused]] const long long tw1 = MRNNT_K1_CLOCK();
            const int2 mk = meta_sh[stage * 32 + r];
            const int64_t k = COMPACT ? mk.y : u;
            if (COMPACT && k < 0) break;
            const int64_t row = (blockIdx.x + k * gridDim.x) * G + r;
            MRNNT_K1_ADD(warp, 0, tw1 - tw0);
            MRNNT_K1_ADD(warp, 2, 1);
            const int meta = mk.x;
            if (row < rows && meta != kRowDead) {
                const unsigned char *slot = tiles + stage * slot_bytes;
                int lab = -1;
                if (lane == 0 && meta >= 0) lab = __ldg(labels + meta);  // latency hides under the row math
                const E *xrow;
                RowSum den;
                if constexpr (UNALIGNED) {
                    const RowSplit w = row_split<E>(row - r, r, V);
                    xrow = reinterpret_cast<const E *>(slot + w.off);
                    const int e = edge_element(w, V, lane);
                    const float xe = e >= 0 ? Elem<E>::to_float(xrow[e]) : kNegInfF;
                    den = row_sums<E, C, true>(reinterpret_cast<const uint4 *>(slot) + w.vec0, w.ninterior, lane, xe);
                } else {
                    xrow = reinterpret_cast<const E *>(slot) + static_cast<size_t>(r) * V;
                    den = row_sums<E, C>(reinterpret_cast<const uint4 *>(xrow), NV, lane);
                }
                if (lane == 0) {
                    const bool has = lab >= 0 && lab < V;
                    k1_store_row(lp, row, den, Elem<E>::to_float(xrow[blank]),
                                 has ? Elem<E>::to_float(xrow[lab]) : kNegInfF);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + stage);
            MRNNT_K1_ADD(warp, 1, MRNNT_K1_CLOCK() - tw1);
        }
        if (zero_warp && lane == 0) atomicAdd(consumers_done, 1);
        if (lane == 0) MRNNT_TL_MAX(zero.tl_slot, 2);
    }
}

}  // namespace mrnnt
