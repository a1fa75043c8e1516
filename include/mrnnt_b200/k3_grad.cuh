// K3 -- gradient of the loss w.r.t. the logits (softmax folded in), written exactly once.
//
// Replaces compute_grad_kernel<128> (reference include/gpu_rnnt_kernel.h:239-288: one 128-thread
// block per row, a linear search for the utterance per block, alpha()/beta() re-evaluated from
// global memory per element, up to two exp per element, int32 element index).  CPU twin:
// cpu_rnnt.h:216-236.
//
//   g[row, v] = 2^((x[v] * kLog2e + H) + L) - [v == blank] 2^qb - [v == label_s] 2^ql
// with the per-row record (H, qb, ql, L) and the row's label prepared by K2 (k2_lattice.cuh): H + L = the row's base-2
// log-softmax denominator plus the log2 of the row's occupancy, so the exponent is formed by ONE fused multiply-add
// whose result is small (whatever the magnitude of the logits) plus a small correction, and the two subtracted terms
// are complete.  Rows whose H is -inf (alpha(t-1,s) outside the lattice) are never read: zeros are stored
// (reference: gpu_rnnt_kernel.h:266-271 does this for the geometric part only).
// SCALED variants multiply utterance b's rows by scale[b] on the way out: the chain rule of the reference's
// autograd glue (pytorch_binding/monotonic_rnnt_op.py:97-118, a separate read+write pass over the gradients
// there) at no extra memory traffic.
//
// Streaming design: same persistent-CTA / bulk-copy ring as K1 for the loads; each consumer warp
// turns a row into 2^(x*log2e + c0) with one FFMA + one MUFU.EX2 per element, patches the (at most)
// two special elements, and stores 128-bit vectors straight to global memory (512 contiguous bytes
// per warp instruction).  Algorithmic bytes: 4*V read per live row + 4*V written per row.
#pragma once

#include "common.cuh"
#include "zero_fill.cuh"
#include "k1_lse.cuh"
#include "peer_reduce.cuh"

namespace mrnnt {

// Optional hand-over of the B costs to the host (a synchronous call's staging buffer in host-mapped pinned memory,
// engine.cuh): done here, by the first warp of the first CTA of the LAST kernel of the call, so that the PCIe write
// overlaps the whole gradient pass instead of sitting between the lattice kernel and this one.  With `ready` the warp
// then publishes `seq` there (release, system scope): the host, polling that word, has the costs ~2 us after the
// lattice kernel's end and returns to its caller while this kernel is still streaming (Engine::compute, early return).
struct CostMirror {
    const float *costs = nullptr;  // device, final since the lattice kernel
    float *mapped = nullptr;       // host-mapped copy, or nullptr
    int B = 0;
    unsigned *ready = nullptr;     // host-mapped word behind the costs, or nullptr
    unsigned seq = 0u;             // what `ready` is set to once the costs are in place
};
// (called by one whole warp of the first CTA)
__device__ __forceinline__ void mirror_costs(const CostMirror &m) {
    if (m.mapped == nullptr) return;
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < m.B; i += kWarp) m.mapped[i] = m.costs[i];
    if (m.ready != nullptr) {
        __threadfence_system();
        __syncwarp();
        if (lane == 0) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(m.ready), "r"(m.seq) : "memory");
    }
}

// subtract d from component k of a vector of NE gradients (k is warp-divergent, so the component is picked with
// selects instead of a dynamically indexed register array)
template <int NE>
__device__ __forceinline__ void patch_component(float (&g)[NE], int k, float d) {
#pragma unroll
    for (int i = 0; i < NE; ++i) g[i] -= (k == i) ? d : 0.0f;
}

// ---------------------------------------------------------------------------------------------
// Generic variant: one warp per row, scalar accesses.  Any V, any alignment.
// ---------------------------------------------------------------------------------------------
template <typename E>
static __global__ void __launch_bounds__(kGenericWarps * kWarp)
    k3_grad_generic_kernel(const E *__restrict__ acts, const float4 *__restrict__ coef,
                           const int *__restrict__ rowlab, E *__restrict__ grads, int64_t rows, int V, int blank,
                           const int *__restrict__ rowutt, const float *__restrict__ scale, CostMirror mirror,
                           PeerReduce peer) {
    const bool peer_warp = blockIdx.x == 0 && threadIdx.x < kWarp;  // (peer_reduce.cuh)
    if (peer_warp) {
        mirror_costs(mirror);
        peer_publish(peer);
    }
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * kGenericWarps + (threadIdx.x >> 5);
    const int64_t nwarps = static_cast<int64_t>(gridDim.x) * kGenericWarps;
    for (int64_t row = warp0; row < rows; row += nwarps) {
        const float4 c = __ldg(coef + row);
        E *g = grads + row * V;
        if (c.x == kNegInfF) {
            for (int v = lane; v < V; v += kWarp) g[v] = Elem<E>::from_float(0.0f);
            continue;
        }
        const E *x = acts + row * V;
        const int lab = __ldg(rowlab + row);
        const float sc = scale != nullptr ? __ldg(scale + __ldg(rowutt + row)) : 1.0f;
        for (int v = lane; v < V; v += kWarp) {
            float gv = ex2_approx(fmaf(Elem<E>::to_float(x[v]), kLog2e, c.x) + c.w);
            if (v == blank) gv -= ex2_approx(c.y);
            else if (v == lab) gv -= ex2_approx(c.z);
            g[v] = Elem<E>::from_float(gv * sc);
        }
    }
    if (peer_warp) peer_collect(peer);
}

// ---------------------------------------------------------------------------------------------
// TMA-staged variant.  Requirements: rows are whole 16-byte vectors, acts and grads 16-byte aligned.
// Shared memory: [stages][G*V] elements | full[stages] | empty[stages] | scale[stages][32] floats | coef[stages][32] float4
// | label[stages][32] ints | tile[stages] ints
// ---------------------------------------------------------------------------------------------
// flags
constexpr int kK3WriteDead = 1;  // write the zero rows (off: somebody else zeroes the rows the plan calls dead)
constexpr int kK3Compact = 2;    // tiles without a live row take no ring slot (needs kK3WriteDead off)
constexpr int kK3ZeroShared = 4; // the zero-fill warp continues a fill the LSE kernel's zero-fill warp has begun
constexpr int kK3FixedShift = 8; // flags >> 8: with kK3Dynamic, how many tiles of a CTA are fixed before the counter takes over
constexpr int kK3Dynamic = 8;    // tiles are handed out through a counter (`dyn`) instead of round-robin by CTA index: a
                                 // CTA on an SM that streams faster takes more of them, and the kernel ends when the
                                 // work does, not when the unluckiest CTA has got through its fixed share.  The first
                                 // two tiles of a CTA are fixed (no waiting for the counter at the start); the slots
                                 // carry their tile's index as in the compact mode, which this mode includes
// With a ZeroFill (dst != nullptr) the kernel is launched with one more warp, which zeroes the plan's dead rows next to
// the consumer warps (zero_fill.cuh) out of kZeroFillBytes of shared memory behind the ring: where nearly all rows
// are dead, writing the zeros takes longer than everything else, and it need not wait for anything.
inline size_t k3_smem_bytes(size_t ring_bytes, bool zero_warp) {
    return zero_warp ? (ring_bytes + 127) / 128 * 128 + kZeroFillBytes : ring_bytes;
}

// UNALIGNED: rows are not whole 16-byte vectors (k1_lse.cuh: StreamWindow); slot_bytes = StreamTiling::slot_bytes.
template <typename E, int NW, bool SCALED, bool UNALIGNED = false>
static __global__ void __launch_bounds__((NW + 2) * kWarp, 1)
    k3_grad_tma_kernel(const E *__restrict__ acts, const float4 *__restrict__ coef, const int *__restrict__ rowlab,
                       E *__restrict__ grads, int64_t rows, int V, int blank, int G, int stages,
                       const int *__restrict__ rowutt,
                       const float *__restrict__ scale, CostMirror mirror, int flags, ZeroFill zero, size_t ring_bytes,
                       unsigned *__restrict__ dyn, PeerReduce peer, size_t slot_bytes, int64_t own_dead_from) {
    // own_dead_from: without kK3WriteDead, the plan's dead rows from this row on are nevertheless this kernel's to write
    // (the lattice kernel's fill took the rows before it: Engine::k2_fill_share); INT64_MAX: none
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NE = Elem<E>::kPerVec;
    unsigned char *tiles = smem_raw;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + static_cast<size_t>(stages) * slot_bytes);
    uint64_t *empty = full + stages;
    float *scale_sh = reinterpret_cast<float *>(empty + stages);
    float4 *coef_sh = reinterpret_cast<float4 *>(reinterpret_cast<unsigned char *>(empty + stages) +
                                                 static_cast<size_t>(stages) * 32 * sizeof(float));
    int *lab_sh = reinterpret_cast<int *>(coef_sh + static_cast<size_t>(stages) * 32);
    int *tile_sh = lab_sh + static_cast<size_t>(stages) * 32;  // [stages] tile held by the slot

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, static_cast<uint32_t>(G));
        }
        mbar_init_fence();
    }
    __syncthreads();
    if (warp == NW + 1) {
        // ---------------- zero-fill warp (only launched with a ZeroFill): needs nothing the lattice kernel wrote ----
        unsigned char *zbuf = smem_raw + (ring_bytes + 127) / 128 * 128;
        if (flags & kK3ZeroShared) zero_dead_rows_impl<false>(zero, 0, 0, zbuf, [] { return false; });
        else zero_dead_rows(zero, blockIdx.x, gridDim.x, zbuf);
        return;
    }
    if (threadIdx.x == 0) MRNNT_TL_MIN(zero.tl_slot, 8);
    pdl_launch_dependents();  // the next call's LSE kernel may be scheduled as our CTAs retire (it waits for all of us)
    pdl_wait();  // the coefficients come from the lattice kernel; everything above overlapped its tail
    if (threadIdx.x == 0) MRNNT_TL_MIN(zero.tl_slot, 9);
    // the sum of this GPU's costs goes out to the peers now and the world's sum is picked up when this warp has got
    // through its rows (peer_reduce.cuh): the exchange rides along with the gradient pass
    const bool peer_warp = blockIdx.x == 0 && warp == 0;
    if (peer_warp) {
        mirror_costs(mirror);
        peer_publish(peer);
    }

    const int64_t ntiles = (rows + G - 1) / G;
    const int64_t nloc = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const bool write_dead = (flags & kK3WriteDead) != 0;
    // compact (only without write_dead): tiles without a live row take no ring slot; the slots carry their tile's
    // index, the consumers walk the slots and stop at a terminator (as in the LSE kernel's COMPACT variant)
    const bool dynamic = (flags & kK3Dynamic) != 0;
    const bool compact = (flags & kK3Compact) != 0 && !dynamic;
    const bool slot_names_tile = compact || dynamic;

    if (warp == NW) {
        // ---------------- producer warp ----------------
        const uint64_t policy = l2_policy_evict_first();
        int stage = 0;
        uint32_t phase = 0;
        auto fill_slot = [&](int64_t k, const float4 &c, int lb, float sc, uint32_t mask) {
            const int64_t row0 = (dynamic ? k : blockIdx.x + k * gridDim.x) * G;  // (dynamic: k is the tile itself)
            mbar_wait(empty + stage, phase ^ 1u);
            coef_sh[stage * 32 + lane] = c;
            lab_sh[stage * 32 + lane] = lb;
            if (SCALED) scale_sh[stage * 32 + lane] = sc;
            if (lane == 0) tile_sh[stage] = static_cast<int>(k);
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
        };
        if (dynamic) {
            // the first `fixed` tiles of a CTA are its round-robin share (flags >> kK3FixedShift, at least two), the
            // counter hands out the rest: the hand-out only has to even out the END of the kernel
            int fixed = flags >> kK3FixedShift;
            if (static_cast<int64_t>(fixed) * gridDim.x > ntiles) fixed = static_cast<int>(ntiles / gridDim.x);
            if (fixed < 2) fixed = 2;
            TileGrabber grabber;  // (k1_lse.cuh)
            grabber.start(dyn, lane, fixed);
            int kfix = 2;
            auto next_tile = [&]() -> int64_t {
                if (kfix < fixed) return blockIdx.x + static_cast<int64_t>(kfix++) * gridDim.x;
                return grabber.next();
            };
            auto load_coef = [&](int64_t tile) {
                const int64_t row = tile * G + lane;
                return (tile < ntiles && lane < G && row < rows) ? __ldg(coef + row) : make_float4(kNegInfF, 0.f, 0.f, 0.f);
            };
            auto load_scale = [&](int64_t tile) {
                const int64_t row = tile * G + lane;
                return (SCALED && tile < ntiles && lane < G && row < rows) ? __ldg(scale + __ldg(rowutt + row)) : 1.0f;
            };
            auto load_lab = [&](int64_t tile) {
                const int64_t row = tile * G + lane;
                return (tile < ntiles && lane < G && row < rows) ? __ldg(rowlab + row) : kRowDead;
            };
            int64_t t0 = blockIdx.x, t1 = static_cast<int64_t>(blockIdx.x) + gridDim.x;
            float4 c0 = load_coef(t0), c1 = load_coef(t1);
            float s0 = load_scale(t0), s1 = load_scale(t1);
            int l0 = load_lab(t0), l1 = load_lab(t1);
            while (t0 < ntiles) {
                const int64_t t2 = next_tile();
                const float4 c2 = load_coef(t2);
                const float s2 = load_scale(t2);
                const int l2 = load_lab(t2);
                const uint32_t mask = __ballot_sync(0xffffffffu, !(c0.x == kNegInfF));
                if (write_dead || mask != 0u || (t0 + 1) * G > own_dead_from) fill_slot(t0, c0, l0, s0, mask);
                t0 = t1; c0 = c1; s0 = s1; l0 = l1;
                t1 = t2; c1 = c2; s1 = s2; l1 = l2;
            }
            for (int i = 0; i < stages; ++i) {  // one terminator per slot, as in the compact mode
                mbar_wait(empty + stage, phase ^ 1u);
                if (lane == 0) {
                    tile_sh[stage] = -1;
                    mbar_arrive_expect_tx(full + stage, 0u);
                }
                if (++stage == stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            grabber.finish();
        } else if (!compact) {
            auto load_coef = [&](int64_t k) {
                const int64_t row = (blockIdx.x + k * gridDim.x) * G + lane;
                return (k < nloc && lane < G && row < rows) ? __ldg(coef + row) : make_float4(kNegInfF, 0.f, 0.f, 0.f);
            };
            auto load_scale = [&](int64_t k) {
                const int64_t row = (blockIdx.x + k * gridDim.x) * G + lane;
                return (SCALED && k < nloc && lane < G && row < rows) ? __ldg(scale + __ldg(rowutt + row)) : 1.0f;
            };
            auto load_lab = [&](int64_t k) {
                const int64_t row = (blockIdx.x + k * gridDim.x) * G + lane;
                return (k < nloc && lane < G && row < rows) ? __ldg(rowlab + row) : kRowDead;
            };
            float4 c_next = load_coef(0), c_next2 = load_coef(1);
            float s_next = load_scale(0), s_next2 = load_scale(1);
            int l_next = load_lab(0), l_next2 = load_lab(1);
            for (int64_t k = 0; k < nloc; ++k) {
                const float4 c = c_next;
                const float sc = s_next;
                const int lb = l_next;
                c_next = c_next2;  // two tiles ahead: the latency hides behind two tiles' waits
                s_next = s_next2;
                l_next = l_next2;
                c_next2 = load_coef(k + 2);
                s_next2 = load_scale(k + 2);
                l_next2 = load_lab(k + 2);
                fill_slot(k, c, lb, sc, __ballot_sync(0xffffffffu, !(c.x == kNegInfF)));
            }
        } else {
            // coefficients for 32/G tiles per load (lane l: tile k0 + l/G, row l%G), two such batches ahead: with
            // mostly dead tiles the producer does nothing but wait for these loads
            const int TPB = 32 / G;  // (G is a power of two <= 32)
            const uint32_t gmask = G == 32 ? 0xffffffffu : ((1u << G) - 1u);
            auto batch_row = [&](int64_t k0) {
                const int64_t kk = k0 + lane / G;
                const int64_t row = (blockIdx.x + kk * gridDim.x) * G + (lane % G);
                return (kk < nloc && row < rows) ? row : static_cast<int64_t>(-1);
            };
            auto load_coef = [&](int64_t k0) {
                const int64_t row = batch_row(k0);
                return row >= 0 ? __ldg(coef + row) : make_float4(kNegInfF, 0.f, 0.f, 0.f);
            };
            auto load_scale = [&](int64_t k0) {
                const int64_t row = batch_row(k0);
                return (SCALED && row >= 0) ? __ldg(scale + __ldg(rowutt + row)) : 1.0f;
            };
            auto load_lab = [&](int64_t k0) {
                const int64_t row = batch_row(k0);
                return row >= 0 ? __ldg(rowlab + row) : kRowDead;
            };
            float4 cb0 = load_coef(0), cb1 = load_coef(TPB);
            float sb0 = load_scale(0), sb1 = load_scale(TPB);
            int lb0 = load_lab(0), lb1 = load_lab(TPB);
            for (int64_t k0 = 0; k0 < nloc; k0 += TPB) {
                const float4 cb = cb0;
                const float sb = sb0;
                const int lbb = lb0;
                cb0 = cb1;
                sb0 = sb1;
                lb0 = lb1;
                cb1 = load_coef(k0 + 2 * TPB);
                sb1 = load_scale(k0 + 2 * TPB);
                lb1 = load_lab(k0 + 2 * TPB);
                const uint32_t ball = __ballot_sync(0xffffffffu, !(cb.x == kNegInfF));
                for (int j = 0; j < TPB && k0 + j < nloc; ++j) {
                    const uint32_t mask = (ball >> (j * G)) & gmask;
                    if (mask == 0u && (blockIdx.x + (k0 + j) * gridDim.x + 1) * G <= own_dead_from) continue;
                    const int src = j * G + (lane % G);
                    float4 c;
                    c.x = __shfl_sync(0xffffffffu, cb.x, src);
                    c.y = __shfl_sync(0xffffffffu, cb.y, src);
                    c.z = __shfl_sync(0xffffffffu, cb.z, src);
                    c.w = __shfl_sync(0xffffffffu, cb.w, src);
                    const float sc = SCALED ? __shfl_sync(0xffffffffu, sb, src) : 1.0f;
                    const int lb = __shfl_sync(0xffffffffu, lbb, src);
                    if (lane >= G) c.x = kNegInfF;
                    fill_slot(k0 + j, c, lb, sc, mask);
                }
            }
            // one terminator per slot: every consumer warp meets one within its next `stages` slot uses
            for (int i = 0; i < stages; ++i) {
                mbar_wait(empty + stage, phase ^ 1u);
                if (lane == 0) {
                    tile_sh[stage] = -1;
                    mbar_arrive_expect_tx(full + stage, 0u);
                }
                if (++stage == stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else {
        // ---------------- consumer warps ----------------
        // q walks the rows of the slot uses u = 0, 1, ... in order (G rows each)
        const int NV = V / NE;
        const int64_t nq = nloc * G;
        for (int64_t q = warp; slot_names_tile || q < nq; q += NW) {
            const int64_t u = q / G;
            const int r = static_cast<int>(q - u * G);
            const int stage = static_cast<int>(u % stages);
            const uint32_t phase = static_cast<uint32_t>((u / stages) & 1);
            mbar_wait(full + stage, phase);
            const int64_t k = slot_names_tile ? tile_sh[stage] : u;
            if (k < 0) break;
            const int64_t row = (dynamic ? k : blockIdx.x + k * gridDim.x) * G + r;
            if (row < rows) {
                const float4 c = coef_sh[stage * 32 + r];  // (H, qb, ql, L)
                const int lab = lab_sh[stage * 32 + r];    // -1 when the row has no (non-blank) label
                // The row's vectors in the slot and in the gradient array.  A row that is not whole 16-byte vectors
                // (UNALIGNED; k1_lse.cuh: RowSplit): its interior -- the aligned vectors wholly inside it, the same ones in
                // the slot and in global memory -- plus up to 2 * (NE - 1) edge elements, one per lane 0, 1, ...
                const unsigned char *slot = tiles + stage * slot_bytes;
                const uint4 *xv;
                uint4 *gv;
                int nvec = NV, nhead = 0, e_edge = -1;
                [[maybe_unused]] const E *xrow = nullptr;
                E *grow = grads + row * V;
                if constexpr (UNALIGNED) {
                    const RowSplit w = row_split<E>(row - r, r, V);
                    xrow = reinterpret_cast<const E *>(slot + w.off);
                    xv = reinterpret_cast<const uint4 *>(slot) + w.vec0;
                    gv = reinterpret_cast<uint4 *>(grow + w.nhead);
                    nvec = w.ninterior;
                    nhead = w.nhead;
                    e_edge = edge_element(w, V, lane);
                } else {
                    xv = reinterpret_cast<const uint4 *>(slot) + static_cast<size_t>(r) * NV;
                    gv = reinterpret_cast<uint4 *>(grow);
                }
                if (c.x == kNegInfF) {
                    // a zero row; without write_dead somebody else zeroes the rows the plan calls dead (marked in the
                    // label slot; the lattice kernel's fill or a zero-fill warp, zero_fill.cuh), and only a row
                    // INSIDE the lattice that came out as zero (masked logits) is written here
                    if (write_dead || lab != kRowDead || row >= own_dead_from) {
                        const uint4 z = make_uint4(0u, 0u, 0u, 0u);  // +0.0 in either element type
                        for (int j = lane; j < nvec; j += kWarp) st_stream_u4(gv + j, z);
                        if (UNALIGNED && e_edge >= 0) grow[e_edge] = Elem<E>::from_float(0.0f);
                    }
                } else {
                    const float sc = SCALED ? scale_sh[stage * 32 + r] : 1.0f;
                    const float2 l2 = make_float2(kLog2e, kLog2e), h2 = make_float2(c.x, c.x), lo2 = make_float2(c.w, c.w);
                    // (element e of the row is component (e - nhead) % NE of interior vector (e - nhead) / NE; an edge
                    // element is nobody's component: jb / jl stay -1 and the edge code below patches it)
                    const int ib = blank - nhead, il = lab - nhead;
                    const bool b_in = ib >= 0 && ib < nvec * NE, l_in = lab >= 0 && il >= 0 && il < nvec * NE;
                    const int jb = b_in ? ib / NE : -1, kb = ib - jb * NE;
                    const int jl = l_in ? il / NE : -1, kl = il - jl * NE;
#pragma unroll 2
                    for (int j = lane; j < nvec; j += kWarp) {
                        float x[NE], g[NE];
                        Elem<E>::unpack(xv[j], x);
#pragma unroll
                        for (int i = 0; i < NE; i += 2) {  // two elements at a time (FFMA2 / FADD2)
                            const float2 ee = __fadd2_rn(__ffma2_rn(make_float2(x[i], x[i + 1]), l2, h2), lo2);
                            g[i] = ex2_approx(ee.x);
                            g[i + 1] = ex2_approx(ee.y);
                        }
                        // the (at most) two elements that lose something: behind real branches (the exponential inside
                        // keeps the compiler from turning them into predicated subtractions on every vector)
                        if (j == jb) patch_component<NE>(g, kb, ex2_approx(c.y));
                        if (j == jl) patch_component<NE>(g, kl, ex2_approx(c.z));
                        if (SCALED) {
#pragma unroll
                            for (int i = 0; i < NE; ++i) g[i] *= sc;
                        }
                        st_stream_u4(gv + j, Elem<E>::pack(g));
                    }
                    if (UNALIGNED && e_edge >= 0) {
                        float gval = ex2_approx(fmaf(Elem<E>::to_float(xrow[e_edge]), kLog2e, c.x) + c.w);
                        if (e_edge == blank) gval -= ex2_approx(c.y);
                        else if (e_edge == lab) gval -= ex2_approx(c.z);
                        if (SCALED) gval *= sc;
                        grow[e_edge] = Elem<E>::from_float(gval);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + stage);
        }
        if (peer_warp) peer_collect(peer);
        if (lane == 0) MRNNT_TL_MAX(zero.tl_slot, 10);
    }
}

}  // namespace mrnnt
