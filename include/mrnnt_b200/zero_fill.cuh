// Zero fill: writing the gradient rows that are zero by construction (the plan's dead rows, rowmeta == kRowDead) with
// bulk shared->global copies, by warps that do nothing else -- inside the lattice kernel while its recursions leave the
// memory system idle (k2_lattice.cuh), or inside the gradient kernel next to its consumer warps where nearly all rows
// are dead (k3_grad.cuh).
#pragma once

#include "common.cuh"
#include "plan.cuh"

namespace mrnnt {

constexpr int kZeroFillBytes = 8192;  // zeroed shared memory every bulk store reads from (tools/zero_probe.cu: 8 KB
                                      // stores from one warp per SM already reach the write bandwidth of the GPU)
#ifndef MRNNT_ZERO_FILL_DEPTH
#define MRNNT_ZERO_FILL_DEPTH 2
#endif
constexpr int kZeroFillDepth = MRNNT_ZERO_FILL_DEPTH;  // units a zero-fill warp holds while as many grabs are in flight
struct ZeroFill {
    unsigned char *dst;   // the gradient buffer
    const int *rowmeta;   // [rows]
    int64_t rows;         // rows of the whole batch
    unsigned row_bytes;   // V * sizeof(element), a multiple of 4
    unsigned *ctr;        // the hand-out counter (OWNED: {units handed out, warps finished}, zero between launches)
    int64_t unit_begin = 0;   // the units (of 32 rows) this fill is responsible for: [unit_begin, unit_end);
    int tl_slot = 0;          // (MRNNT_TIMELINE: the call's slot of the timeline; rides in this struct into K1 and K3)
    int64_t unit_end = -1;    // unit_end < 0: up to the last one.  (The lattice kernel's fill takes the front of the
                              // batch, the gradient kernel's zero-fill warp the rest: Engine::k2_fill_share.)
};

__device__ __forceinline__ void bulk_s2g(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"(static_cast<uint32_t>(__cvta_generic_to_shared(ssrc))), "r"(bytes)
                 : "memory");
}
// Called by whole warps.  The rows of the whole batch are handed out in units of 32, in order, through a counter: the
// dead rows sit at the two ends of every utterance (or everywhere, under an alignment band), and a fixed split leaves
// some warps with twice the bytes of others.  The grab for the unit three ahead and the row flags of the unit two
// ahead are in flight while a unit is processed.  A lane that sees the first row of a run of dead rows stores the
// whole run, 8 KB at a time.  Two protocols:
//  * OWNED (one kernel does the whole fill): warp `fw` of the `nfw` zero-fill warps of the grid takes the units fw,
//    fw + nfw, ... (kZeroFillDepth of them) without asking, the counter hands out the ones behind those, and the last warp to finish sets
//    `ctr` = {units handed out, warps finished} back to zero for the next launch.
//  * SHARED (!OWNED; the kernels of one call take turns at one counter, each for as long as it runs): every unit comes
//    from the counter, which nobody resets here (the engine alternates between two, and the lattice kernel of a call
//    clears the other one); `stop()` is looked at once per unit, and a warp that sees it true takes no more units
//    (the up to three it already holds it still finishes: whoever continues starts behind them).
template <bool OWNED, typename Stop>
__device__ __forceinline__ void zero_dead_rows_impl(const ZeroFill &a, int fw, int nfw, unsigned char *zbuf, Stop stop) {
    const int lane = threadIdx.x & 31;
    for (int i = lane * 16; i < kZeroFillBytes; i += kWarp * 16) *reinterpret_cast<uint4 *>(zbuf + i) = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const int64_t rows = a.rows;
    const int *meta = a.rowmeta;
    unsigned *ctr = a.ctr;
    const int64_t all_units = (rows + kWarp - 1) / kWarp;
    const int64_t nunits = a.unit_end < 0 || a.unit_end > all_units ? all_units : a.unit_end;  // (one past the last unit)
    const int64_t ubase = a.unit_begin;
    constexpr unsigned kNoUnit = 0xffffffffu;
    bool stopped = false;
    auto grab = [&]() {  // (the value is only looked at two units later)
        if (!OWNED && stopped) return kNoUnit;
        return lane == 0 ? atomicAdd(ctr, 1u) : 0u;
    };
    auto unit_of = [&](unsigned raw) {
        const unsigned v = __shfl_sync(0xffffffffu, raw, 0);
        return (!OWNED && v == kNoUnit) ? nunits
                                        : ubase + static_cast<int64_t>(v) + (OWNED ? static_cast<int64_t>(kZeroFillDepth) * nfw : 0);
    };
    auto load = [&](int64_t u) {
        const int64_t r = u * kWarp + lane;
        return (u < nunits && r < rows) ? __ldg(meta + r) : 0;
    };
    // kZeroFillDepth units are held (index known, row flags loaded or on their way) and as many grabs are in flight
    constexpr int D = kZeroFillDepth;
    unsigned raw[D];
    int64_t u[D];
    int m[D];
#pragma unroll
    for (int i = 0; i < D; ++i) raw[i] = grab();
    if (OWNED) {
#pragma unroll
        for (int i = 0; i < D; ++i) u[i] = ubase + fw + static_cast<int64_t>(i) * nfw;
    } else {
#pragma unroll
        for (int i = 0; i < D; ++i) u[i] = unit_of(raw[i]);
#pragma unroll
        for (int i = 0; i < D; ++i) raw[i] = grab();
    }
#pragma unroll
    for (int i = 0; i < D; ++i) m[i] = load(u[i]);
    while (u[0] < nunits) {
        if (!OWNED) stopped = stopped || stop();
        const unsigned raw_new = grab();
        const int64_t u_new = unit_of(raw[0]);
        const int m_new = load(u_new);
        const int64_t u0 = u[0];
        const bool dead = m[0] == kRowDead;
        const uint32_t mask = __ballot_sync(0xffffffffu, dead);
        if (dead && (lane == 0 || ((mask >> (lane - 1)) & 1u) == 0u)) {
            const uint32_t inv = ~(mask >> lane);  // (the shift fills with zeros: inv != 0 unless lane == 0 and all dead)
            const int len = inv ? __ffs(inv) - 1 : kWarp;
            unsigned char *p = a.dst + static_cast<size_t>(u0 * kWarp + lane) * a.row_bytes;
            size_t left = static_cast<size_t>(len) * a.row_bytes;
            // rows that are not whole 16-byte vectors (row_bytes % 16 != 0, a multiple of 4): the up to 12 bytes before
            // the run's first and behind its last 16-byte boundary go out as ordinary 4-byte stores, the rest in bulk
            const size_t head = min(static_cast<size_t>((16u - (reinterpret_cast<uintptr_t>(p) & 15u)) & 15u), left);
            for (size_t b = 0; b < head; b += 4) *reinterpret_cast<unsigned *>(p + b) = 0u;
            p += head;
            left -= head;
            const size_t tail = left & 15u;
            left -= tail;
            for (size_t b = 0; b < tail; b += 4) *reinterpret_cast<unsigned *>(p + left + b) = 0u;
            while (left > 0) {
                const uint32_t nbytes = left < static_cast<size_t>(kZeroFillBytes) ? static_cast<uint32_t>(left) : kZeroFillBytes;
                bulk_s2g(p, zbuf, nbytes);
                p += nbytes;
                left -= nbytes;
            }
        }
        if (!OWNED) {
            // a kernel that stops must not leave megabytes of queued stores behind (measured: the bulk-copy pipe of an
            // SM takes them by the thousand, 170 us to drain): at most two units' stores in flight per lane
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
#pragma unroll
        for (int i = 0; i + 1 < D; ++i) {
            u[i] = u[i + 1];
            m[i] = m[i + 1];
            raw[i] = raw[i + 1];
        }
        u[D - 1] = u_new;
        m[D - 1] = m_new;
        raw[D - 1] = raw_new;
    }
    // the stores must have left shared memory before the CTA gives it up; the kernel's end makes them visible
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if (OWNED) {
        // the last warp to finish leaves the counters as it found them (every grab of this launch has been made by
        // then: a warp looks at its outstanding grabs before it reports)
        unsigned seen = 0u;
#pragma unroll
        for (int i = 0; i < D; ++i) seen += __shfl_sync(0xffffffffu, raw[i], 0);
        asm volatile("and.b32 %0, %0, 0;" : "+r"(seen));  // (a zero the compiler cannot fold: the grabs' answers are inputs of the report)
        __threadfence();
        if (lane == 0 && atomicAdd(ctr + 1, seen + 1u) == static_cast<unsigned>(nfw) - 1u) {
            ctr[0] = 0u;
            ctr[1] = 0u;
            __threadfence();
        }
    }
}

__device__ __forceinline__ void zero_dead_rows(const ZeroFill &a, int fw, int nfw, unsigned char *zbuf) {
    zero_dead_rows_impl<true>(a, fw, nfw, zbuf, [] { return false; });
}

}  // namespace mrnnt
