// K3 -- gradient of the loss w.r.t. the logits (softmax folded in), written exactly once.
//
// Replaces compute_grad_kernel<128> (reference include/gpu_rnnt_kernel.h:239-288: one 128-thread
// block per row, a linear search for the utterance per block, alpha()/beta() re-evaluated from
// global memory per element, up to two exp per element, int32 element index).  CPU twin:
// cpu_rnnt.h:216-236.
//
//   g[row, v] = exp(x + c0)  - [v == blank] exp(x + cb)  - [v == label_s] exp(x + cl)
// with the three per-row coefficients (already multiplied by log2 e) and the row's label prepared by
// K2 (k2_lattice.cuh).  Rows whose c0 is -inf (alpha(t-1,s) outside the lattice) are never read:
// zeros are stored (reference: gpu_rnnt_kernel.h:266-271 does this for the geometric part only).
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
#include "k1_lse.cuh"

namespace mrnnt {

__device__ __forceinline__ float grad_elem(float x, float c) { return ex2_approx(fmaf(x, kLog2e, c)); }

// subtract the blank / label term from component (idx & 3) of a vector whose first element is 4*j
__device__ __forceinline__ void patch_component(float4 &g, const float4 &x, int j, int idx, float c) {
    if ((idx >> 2) == j) {
        const int k = idx & 3;
        const float xv = k == 0 ? x.x : (k == 1 ? x.y : (k == 2 ? x.z : x.w));
        const float d = grad_elem(xv, c);
        if (k == 0) g.x -= d; else if (k == 1) g.y -= d; else if (k == 2) g.z -= d; else g.w -= d;
    }
}

// ---------------------------------------------------------------------------------------------
// Generic variant: one warp per row, scalar accesses.  Any V, any alignment.
// ---------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(kGenericWarps * kWarp)
    k3_grad_generic_kernel(const float *__restrict__ acts, const float4 *__restrict__ coef,
                           float *__restrict__ grads, int64_t rows, int V, int blank,
                           const int *__restrict__ rowutt, const float *__restrict__ scale) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * kGenericWarps + (threadIdx.x >> 5);
    const int64_t nwarps = static_cast<int64_t>(gridDim.x) * kGenericWarps;
    for (int64_t row = warp0; row < rows; row += nwarps) {
        const float4 c = __ldg(coef + row);
        float *g = grads + row * V;
        if (c.x == kNegInfF) {
            for (int v = lane; v < V; v += kWarp) g[v] = 0.0f;
            continue;
        }
        const float *x = acts + row * V;
        const int lab = __float_as_int(c.w);
        const float sc = scale != nullptr ? __ldg(scale + __ldg(rowutt + row)) : 1.0f;
        for (int v = lane; v < V; v += kWarp) {
            const float xv = __ldg(x + v);
            float gv = grad_elem(xv, c.x);
            if (v == blank) gv -= grad_elem(xv, c.y);
            else if (v == lab) gv -= grad_elem(xv, c.z);
            g[v] = gv * sc;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// TMA-staged variant.  Requirements: V % 4 == 0, acts and grads 16-byte aligned.
// Shared memory: [stages][G*V] floats | full[stages] | empty[stages] | scale[stages][32] floats | coef[stages][32] float4
// ---------------------------------------------------------------------------------------------
template <int NW, bool SCALED>
static __global__ void __launch_bounds__((NW + 1) * kWarp, 1)
    k3_grad_tma_kernel(const float *__restrict__ acts, const float4 *__restrict__ coef, float *__restrict__ grads,
                       int64_t rows, int V, int blank, int G, int stages, const int *__restrict__ rowutt,
                       const float *__restrict__ scale) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const size_t tile_floats = static_cast<size_t>(G) * V;
    float *tiles = reinterpret_cast<float *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + static_cast<size_t>(stages) * tile_floats * 4);
    uint64_t *empty = full + stages;
    float *scale_sh = reinterpret_cast<float *>(empty + stages);
    float4 *coef_sh = reinterpret_cast<float4 *>(reinterpret_cast<unsigned char *>(empty + stages) +
                                                 static_cast<size_t>(stages) * 32 * sizeof(float));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, static_cast<uint32_t>(G));
        }
        mbar_init_fence();
    }
    __syncthreads();
    pdl_wait();  // the coefficients come from the lattice kernel; everything above overlapped its tail

    const int64_t ntiles = (rows + G - 1) / G;
    const int64_t nloc = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == NW) {
        // ---------------- producer warp ----------------
        const uint64_t policy = l2_policy_evict_first();
        auto load_coef = [&](int64_t k) {
            const int64_t row = (blockIdx.x + k * gridDim.x) * G + lane;
            return (k < nloc && lane < G && row < rows) ? __ldg(coef + row) : make_float4(kNegInfF, 0.f, 0.f, 0.f);
        };
        auto load_scale = [&](int64_t k) {
            const int64_t row = (blockIdx.x + k * gridDim.x) * G + lane;
            return (SCALED && k < nloc && lane < G && row < rows) ? __ldg(scale + __ldg(rowutt + row)) : 1.0f;
        };
        float4 c_next = load_coef(0);
        float s_next = load_scale(0);
        int stage = 0;
        uint32_t phase = 0;
        for (int64_t k = 0; k < nloc; ++k) {
            const float4 c = c_next;
            const float sc = s_next;
            c_next = load_coef(k + 1);  // one tile ahead: its latency hides behind this tile's wait
            s_next = load_scale(k + 1);
            const int64_t row0 = (blockIdx.x + k * gridDim.x) * G;
            const uint32_t mask = __ballot_sync(0xffffffffu, !(c.x == kNegInfF));
            mbar_wait(empty + stage, phase ^ 1u);
            coef_sh[stage * 32 + lane] = c;
            if (SCALED) scale_sh[stage * 32 + lane] = sc;
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_expect_tx(full + stage, static_cast<uint32_t>(__popc(mask)) * static_cast<uint32_t>(V) * 4u);
                issue_live_runs(mask, tiles + stage * tile_floats, acts + row0 * V, V, full + stage, policy);
            }
            if (++stage == stages) {
                stage = 0;
                phase ^= 1u;
            }
        }
    } else {
        // ---------------- consumer warps ----------------
        const int V4 = V >> 2;
        const int64_t nq = nloc * G;
        for (int64_t q = warp; q < nq; q += NW) {
            const int64_t k = q / G;
            const int r = static_cast<int>(q - k * G);
            const int stage = static_cast<int>(k % stages);
            const uint32_t phase = static_cast<uint32_t>((k / stages) & 1);
            const int64_t row = (blockIdx.x + k * gridDim.x) * G + r;
            mbar_wait(full + stage, phase);
            if (row < rows) {
                const float4 c = coef_sh[stage * 32 + r];
                float4 *g4 = reinterpret_cast<float4 *>(grads + row * V);
                if (c.x == kNegInfF) {
                    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int j = lane; j < V4; j += kWarp) st_stream_f4(g4 + j, z);
                } else {
                    const float4 *x4 = reinterpret_cast<const float4 *>(tiles + stage * tile_floats +
                                                                        static_cast<size_t>(r) * V);
                    const int lab = __float_as_int(c.w);  // -1 when the row has no (non-blank) label
                    const int jb = blank >> 2, jl = lab >> 2;
                    const float sc = SCALED ? scale_sh[stage * 32 + r] : 1.0f;
#pragma unroll 2
                    for (int j = lane; j < V4; j += kWarp) {
                        const float4 x = x4[j];
                        float4 g;
                        g.x = grad_elem(x.x, c.x);
                        g.y = grad_elem(x.y, c.x);
                        g.z = grad_elem(x.z, c.x);
                        g.w = grad_elem(x.w, c.x);
                        if (j == jb) patch_component(g, x, j, blank, c.y);
                        if (j == jl) patch_component(g, x, j, lab, c.z);
                        if (SCALED) {
                            g.x *= sc;
                            g.y *= sc;
                            g.z *= sc;
                            g.w *= sc;
                        }
                        st_stream_f4(g4 + j, g);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + stage);
        }
    }
}

}  // namespace mrnnt
