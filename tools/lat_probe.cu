// Latency probes for the lattice recursion (development aid): cycles per frame of the dependent chain,
// single warp, no memory traffic, for several arithmetic designs.
#include <cstdio>
#include <cuda_runtime.h>
#include "../include/mrnnt_b200/common.cuh"
namespace mrnnt {
// log-domain log-sum-exp of the first K2 design (kept here for the latency comparison only)
// log(exp(x) + exp(y)) for the lattice recursion (reference rnnt_helper.h:21-30).  The large
// parts are kept in double (|alpha| grows like T*log V, where a float ulp is already ~6e-5),
// only the bounded correction log1p(exp(-|x-y|)) in (0, ln 2] is evaluated in float.
__device__ __forceinline__ double lse_pair(double x, double y) {
    const double mx = fmax(x, y);
    const double mn = fmin(x, y);
    const float d = static_cast<float>(mn - mx);  // <= 0; NaN only when both are -inf
    const float r = log1pf(expf(d));
    const double out = mx + static_cast<double>(r);
    return (mn == kNegInf) ? mx : out;
}

// Same, with the correction term from the MUFU units: 2^(d log2 e) and log2(1 + u) in float
// (absolute error of the term ~1e-7; it enters a double accumulator, so errors add up like a random
// walk over the T frames instead of being amplified by |alpha|).  ~70 cycles of dependent latency
// instead of ~250 for expf + log1pf; the lattice recursion is a pure latency chain, so this is what
// sets the duration of K2.
__device__ __forceinline__ double lse_pair_fast(double x, double y) {
    const bool gt = x > y;
    const double mx = gt ? x : y;
    const double mn = gt ? y : x;
    const float d = __double2float_rn(mn - mx);  // <= 0, -inf when only one side is -inf, NaN when both are
    const float u = ex2_approx(d * kLog2e);
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + u));
    const double out = mx + static_cast<double>(r * 0.69314718055994531f);
    return (mx == kNegInf) ? kNegInf : out;
}

// lse_pair_fast with the band mask folded into the same select: -inf when the cell is outside the band.
__device__ __forceinline__ double lse_pair_masked(double x, double y, bool in_band) {
    const bool gt = x > y;
    const double mx = gt ? x : y;
    const double mn = gt ? y : x;
    const float d = __double2float_rn(mn - mx);
    const float u = ex2_approx(d * kLog2e);
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + u));
    const double out = mx + static_cast<double>(r * 0.69314718055994531f);
    return (in_band && mx != kNegInf) ? out : kNegInf;
}

}  // namespace mrnnt
using namespace mrnnt;
#define STEPS 2048

// A: current design: double state, float softplus via MUFU, 64-bit shuffle, K states per lane
template <int K> __global__ void chainA(double *out, long long *cyc, double lpv) {
    double st[K]; for (int j = 0; j < K; ++j) st[j] = -1.0 * (threadIdx.x + j);
    long long t0 = clock64();
    for (int t = 0; t < STEPS; ++t) {
        double up = __shfl_up_sync(0xffffffffu, st[K - 1], 1);
        double nxt[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            double below = j == 0 ? up : st[j - 1];
            nxt[j] = lse_pair_masked(below + lpv, st[j] + lpv * 0.5, (t + j) & 1023);
        }
#pragma unroll
        for (int j = 0; j < K; ++j) st[j] = nxt[j];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = st[0] + st[K - 1];
}

// B: float state relative to a running offset (no double on the chain): v = lse(a + lpa, b + lpb) all float
template <int K> __global__ void chainB(float *out, long long *cyc, float lpv) {
    float st[K]; for (int j = 0; j < K; ++j) st[j] = -1.0f * (threadIdx.x + j);
    long long t0 = clock64();
    for (int t = 0; t < STEPS; ++t) {
        float up = __shfl_up_sync(0xffffffffu, st[K - 1], 1);
        float nxt[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            float below = j == 0 ? up : st[j - 1];
            float x = below + lpv, y = st[j] + lpv * 0.5f;
            float mx = fmaxf(x, y), mn = fminf(x, y);
            float u = ex2_approx((mn - mx) * kLog2e);
            float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + u));
            nxt[j] = ((t + j) & 1023) ? fmaf(r, 0.6931472f, mx) : -1e30f;
        }
#pragma unroll
        for (int j = 0; j < K; ++j) st[j] = nxt[j];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = st[0] + st[K - 1];
}

// C: probability domain: a' = a*pb + below*pl  (2 FMA + shuffle)
template <int K> __global__ void chainC(float *out, long long *cyc, float p) {
    float st[K]; for (int j = 0; j < K; ++j) st[j] = 1.0f / (1 + threadIdx.x + j);
    long long t0 = clock64();
    for (int t = 0; t < STEPS; ++t) {
        float up = __shfl_up_sync(0xffffffffu, st[K - 1], 1);
        float nxt[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            float below = j == 0 ? up : st[j - 1];
            nxt[j] = fmaf(below, p, st[j] * (1.0f - p));
        }
#pragma unroll
        for (int j = 0; j < K; ++j) st[j] = nxt[j];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = st[0] + st[K - 1];
}

// D: shuffle only (32-bit and 64-bit) dependent chain
__global__ void chainShfl(float *out, long long *cyc) {
    float f = threadIdx.x; double d = threadIdx.x;
    long long t0 = clock64();
    for (int t = 0; t < STEPS; ++t) f = __shfl_up_sync(0xffffffffu, f, 1) * 1.0001f;
    long long t1 = clock64();
    for (int t = 0; t < STEPS; ++t) d = __shfl_up_sync(0xffffffffu, d, 1) * 1.0001;
    long long t2 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
    out[threadIdx.x] = f + (float)d;
}

int main() {
    double *od; float *of; long long *cyc, h[2];
    cudaMalloc(&od, 32 * 8); cudaMalloc(&of, 32 * 4); cudaMalloc(&cyc, 16);
#define RUN(name, call) call; call; cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost); \
    printf("%-44s %.1f cycles/frame\n", name, (double)h[0] / STEPS);
    RUN("A double+MUFU softplus K=1", (chainA<1><<<1, 32>>>(od, cyc, -6.9)));
    RUN("A double+MUFU softplus K=2", (chainA<2><<<1, 32>>>(od, cyc, -6.9)));
    RUN("A double+MUFU softplus K=4", (chainA<4><<<1, 32>>>(od, cyc, -6.9)));
    RUN("B float log-domain K=1", (chainB<1><<<1, 32>>>(of, cyc, -6.9f)));
    RUN("B float log-domain K=2", (chainB<2><<<1, 32>>>(of, cyc, -6.9f)));
    RUN("B float log-domain K=4", (chainB<4><<<1, 32>>>(of, cyc, -6.9f)));
    RUN("C probability-domain K=2", (chainC<2><<<1, 32>>>(of, cyc, 0.3f)));
    RUN("C probability-domain K=4", (chainC<4><<<1, 32>>>(of, cyc, 0.3f)));
    chainShfl<<<1, 32>>>(of, cyc); chainShfl<<<1, 32>>>(of, cyc);
    cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
    printf("shuffle32+FMUL %.1f, shuffle64+DMUL %.1f cycles/iter\n", (double)h[0] / STEPS, (double)h[1] / STEPS);
    return 0;
}
