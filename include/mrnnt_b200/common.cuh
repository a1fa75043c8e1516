// Device-side building blocks shared by the sm_100a kernels: PTX wrappers for the
// mbarrier / bulk-copy (TMA engine, SASS UBLKCP) / cp.async machinery, warp reductions and the
// numerics helpers.  Header-only on purpose: the reference's framework bindings
// (pytorch_binding/monotonic_rnnt.cu, tensorflow_binding/*.cu) compile the loss from headers with
// nothing but `-I include`, so the whole hot path has to be reachable that way.
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>

#include <cstdint>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 900)
#error "monotonic-rnnt_b200 targets sm_100a (needs mbarrier transaction counts and cp.async.bulk); build with -gencode arch=compute_100a,code=sm_100a"
#endif

namespace mrnnt {

constexpr int kWarp = 32;
constexpr float kLog2e = 1.4426950408889634f;
constexpr double kLog2eD = 1.4426950408889634074;
constexpr double kLn2D = 0.69314718055994530942;
constexpr float kNegInfF = -__builtin_huge_valf();  // usable in host and device code alike
constexpr double kNegInf = -__builtin_huge_val();

// rowmeta encoding (see plan.cuh)
constexpr int kRowNoLabel = -1;  // live row with s == S_b (no label transition leaves it)
constexpr int kRowDead = -2;     // alpha(t-1, s) is outside the lattice: gradient row is exactly zero

// ---------------------------------------------------------------------------------------------
// shared-memory addresses, mbarrier, bulk async copy
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}

// make freshly initialised barriers visible to the async (TMA) proxy
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t tx_bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(tx_bytes)
                 : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// L2 eviction policies for the bulk copies (streamed-once data should not displace reusable lines)
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// 1-D bulk copy global -> shared through the TMA engine; completion is reported as `bytes`
// transaction units on `bar`.  dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar,
                                              uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// ---------------------------------------------------------------------------------------------
// per-thread cp.async (LDGSTS) used as a private prefetch FIFO by the lattice kernel
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_16(void *dst_smem, const void *src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_8(void *dst_smem, const void *src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------------------------
// streaming global stores / loads
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_stream_f4(float4 *p, const float4 &v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// ---------------------------------------------------------------------------------------------
// numerics
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {  // 2^x, MUFU.EX2; ex2(-inf) = +0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// log2 of a positive finite float to ~6e-8 absolute: exponent exactly, mantissa in [1,2) through
// log2f (whose error is relative to a result below 1).
__device__ __forceinline__ double log2_split(float s) {
    int e;
    const float f = frexpf(s, &e);  // s = f * 2^e, f in [0.5, 1)
    return static_cast<double>(e - 1) + static_cast<double>(log2f(f + f));
}

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
