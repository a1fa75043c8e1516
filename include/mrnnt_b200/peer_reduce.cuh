// The path's only collective -- the sum over all GPUs of the summed cost (4 bytes per step) -- done by the
// kernels themselves over peer memory instead of by a collective library's kernel behind them.
//
// Every rank owns a "board" in its device memory, mapped into every peer (CUDA IPC over NVLink / NVSwitch;
// c_api.cu: mrnnt_peer_board_*): 2 * world slots of 8 bytes, slot [epoch & 1][r] = {sum of rank r's costs, epoch}.
//   publish: as soon as this rank's costs are final (the lattice kernel has ended; the first CTA of the gradient
//            kernel does it before it touches a logit) one warp adds them up and stores {sum, epoch} into slot
//            [epoch & 1][rank] of EVERY rank's board, one 8-byte release store per peer;
//   collect: at the END of the gradient kernel (~the whole gradient pass later, so the peers' stores have long
//            landed) the same warp reads its OWN board, waits for the world's slots to carry this epoch, adds them
//            in rank order (every rank gets the same bits) and leaves the total where the host finds it.
// Why not a library all-reduce on a side stream: the gradient kernel is persistent with one CTA of 25 warps and 196 KB
// of shared memory per SM; a collective kernel launched next to it gets no SM until the gradient kernel's CTAs retire
// and then runs -- launch, cross-GPU handshake, copy -- in the open: +20 us per step on 2 GPUs, +33 us on 4 (measured,
// bench.py at N = 2 / 4 before this).  Here the exchange costs one store per peer and one poll.
//
// The two parities make a slot safe to overwrite: rank r can only publish epoch e + 2 after it has collected epoch
// e + 1, which every peer published after it had collected epoch e.
// A peer that never shows up (crashed process) must not hang the GPU for ever: the collect gives up after
// `timeout_ns` (PeerReduce; the host's default is kPeerDefaultTimeoutNs = 60 s, ranks of a training job drift by
// seconds around checkpoints, evaluation and first-step initialisation; 0 waits without limit) and leaves NaN.
// Giving up is FINAL for this rank's board: the slot invariant above no longer holds once a rank has moved on without
// collecting, so the board's `failed` word is set, every later publish through that board is suppressed (the peers
// then time out as well instead of reading sums that belong to another step) and every later collect leaves NaN at
// once; the host sees it in `status_out` and mrnnt_* calls of the handle return RNNT_STATUS_EXECUTION_FAILED from
// then on (Engine::compute).  Recovery = new boards.
#pragma once

#include <cstdint>

#include "common.cuh"

namespace mrnnt {

constexpr int kPeerMaxWorld = 8;                          // one NVSwitch domain
constexpr unsigned long long kPeerDefaultTimeoutNs = 60000000000ull;  // 60 s

struct PeerReduce {
    unsigned long long *boards[kPeerMaxWorld];  // rank r's board as mapped into this process (boards[rank]: our own)
    const float *costs;                         // this rank's B costs (device), final since the lattice kernel
    float *total_out;                           // where the world's sum goes (device or host-mapped); may be nullptr
    int B;
    int rank, world;                            // world == 0: no reduce
    unsigned epoch;                             // > 0, the same on every rank for the same step
    unsigned long long timeout_ns;              // collect gives up after this long (0: never)
    unsigned *status_out;                       // optional (device or host-mapped): set to 1 when a collect gave up
};

// 2 * world slots, then the board's `failed` word (8 bytes: only its owner writes it)
__host__ __device__ inline size_t peer_board_bytes(int world) { return 2 * static_cast<size_t>(world) * 8 + 8; }
__device__ __forceinline__ unsigned long long *peer_failed_word(const PeerReduce &p) {
    return p.boards[p.rank] + 2 * static_cast<size_t>(p.world);
}

__device__ __forceinline__ unsigned long long peer_pack(float v, unsigned epoch) {
    return (static_cast<unsigned long long>(epoch) << 32) | static_cast<unsigned long long>(__float_as_uint(v));
}

// Called by one whole warp once the costs are final.
__device__ __forceinline__ void peer_publish(const PeerReduce &p) {
    if (p.world <= 0) return;
    const int lane = threadIdx.x & 31;
    if (*reinterpret_cast<volatile unsigned long long *>(peer_failed_word(p)) != 0ull) return;  // (warp-uniform)
    float s = 0.0f;
    for (int i = lane; i < p.B; i += 32) s += p.costs[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane < p.world) {
        unsigned long long *slot = p.boards[lane] + (p.epoch & 1u) * static_cast<unsigned>(p.world) + p.rank;
        const unsigned long long v = peer_pack(s, p.epoch);
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(slot), "l"(v) : "memory");
    }
}

// Called by the same whole warp at the end of the kernel.
__device__ __forceinline__ void peer_collect(const PeerReduce &p) {
    if (p.world <= 0) return;
    const int lane = threadIdx.x & 31;
    float mine = 0.0f;
    bool ok = *reinterpret_cast<volatile unsigned long long *>(peer_failed_word(p)) == 0ull;
    if (ok && lane < p.world) {
        const unsigned long long *slot = p.boards[p.rank] + (p.epoch & 1u) * static_cast<unsigned>(p.world) + lane;
        unsigned long long t0 = 0ull, v;
        for (unsigned spins = 0;; ++spins) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(slot) : "memory");
            if (static_cast<unsigned>(v >> 32) == p.epoch) break;
            if ((spins & 1023u) == 1023u) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0ull) t0 = now;
                else if (p.timeout_ns != 0ull && now - t0 > p.timeout_ns) {
                    ok = false;
                    break;
                }
            }
        }
        mine = __uint_as_float(static_cast<unsigned>(v & 0xffffffffull));
    }
    ok = __all_sync(0xffffffffu, ok);
    float total = 0.0f;
    for (int r = 0; r < p.world; ++r) total += __shfl_sync(0xffffffffu, mine, r);  // rank order: the same bits everywhere
    if (lane == 0) {
        if (!ok) {
            *peer_failed_word(p) = 1ull;  // final: see the head of this file
            if (p.status_out != nullptr) *p.status_out = 1u;
        }
        if (p.total_out != nullptr) *p.total_out = ok ? total : __int_as_float(0x7fc00000);
        __threadfence_system();
    }
}

// The whole exchange as a launch of its own (<<<1, 32>>>): behind a cost-only call, which has no gradient kernel.
static __global__ void peer_reduce_kernel(PeerReduce p) {
    peer_publish(p);
    peer_collect(p);
}

}  // namespace mrnnt
