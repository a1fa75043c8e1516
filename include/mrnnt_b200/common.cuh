// Device-side building blocks shared by the sm_100a kernels: PTX wrappers for the
// mbarrier / bulk-copy (TMA engine, SASS UBLKCP) / cp.async machinery, warp reductions and the
// numerics helpers.  Header-only on purpose: the reference's framework bindings
// (pytorch_binding/monotonic_rnnt.cu, tensorflow_binding/*.cu) compile the loss from headers with
// nothing but `-I include`, so the whole hot path has to be reachable that way.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include <cmath>
#include <cstdint>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 900)
#error "monotonic-rnnt_b200 targets sm_100a (needs mbarrier transaction counts and cp.async.bulk); build with -gencode arch=compute_100a,code=sm_100a"
#endif

namespace mrnnt {

constexpr int kWarp = 32;
constexpr float kLog2e = 1.4426950408889634f;  // float(log2 e): the constant of every x * log2(e) in the streaming kernels
// log2 e = kLog2e + kLog2eLo to ~48 bits.  Only the lattice's transition weights use the full value for their own
// logit; the per-row sums (K1) and the gradient (K3) use kLog2e alone, and the denominator carries the difference for
// the row's maximum (lse_finish) -- see the note there.
constexpr float kLog2eLo = 1.925963033500011e-8f;
constexpr float kLog2eLoRel = 1.3349827e-8f;   // kLog2eLo / kLog2e
constexpr double kLog2eD = 1.4426950408889634074;
constexpr double kLn2D = 0.69314718055994530942;
constexpr float kNegInfF = -__builtin_huge_valf();  // usable in host and device code alike
constexpr double kNegInf = -__builtin_huge_val();

// rowmeta encoding (see plan.cuh)
constexpr int kRowNoLabel = -1;  // live row with s == S_b (no label transition leaves it)
constexpr int kRowDead = -2;     // alpha(t-1, s) is outside the lattice: gradient row is exactly zero

// ---------------------------------------------------------------------------------------------
// Scaled linear-domain numbers used by the lattice (K2): value = m * 2^e with a float mantissa and an int
// exponent per cell, so the recursion is multiply-add instead of log-sum-exp and still cannot underflow.
// Non-zero mantissas are >= 1 (weights are normalised to [1,2], the state is brought back to [1,2) once per
// chunk of frames, see cell_renorm).  Zero (an unreachable cell, a masked transition) is m == 0 with the
// exponent kZeroExp: far below every exponent a non-zero value can have, so that a zero term always loses
// the exponent comparison of cell_add without any test of the mantissa; products and sums of zeros drift
// further down (by at most kZeroExp per frame, 16 frames between renormalisations) and never come back.
// ---------------------------------------------------------------------------------------------
struct __align__(8) Cell {  // alpha / beta lattice cell
    float m;
    int e;
};
struct __align__(16) Weight {  // transition weights of one packed row: p(blank) = mb * 2^eb, p(label) = ml * 2^el
    float mb;
    int eb;
    float ml;
    int el;
};
struct __align__(16) RawRow {  // per live packed row (dead rows are never written nor used)
    float xb;  // logit of the blank
    float xl;  // logit of the row's label (-inf without one)
    float dh;  // as K1 leaves it: (dh, dl) = (max_v x[v] * log2 e, sum_v 2^(x[v] log2 e - max)).  K2's phase A replaces
    float dl;  // the pair IN PLACE by dh + dl = -log2 sum_v exp(x[v]) as an unevaluated sum of two floats (~48 bits),
               // so that log2 p(v) = x[v] * log2(e) + dh + dl
};
constexpr int kZeroExp = -(1 << 24);  // non-zero values reach down to 2^-(2^23); 17 frames of drift stay below 2^29

// error-free sum of two floats: s + err == a + b exactly
__host__ __device__ inline void two_sum(float a, float b, float &s, float &err) {
    s = a + b;
    const float bb = s - a;
    err = (a - (s - bb)) + (b - bb);
}

// natural log of m * 2^e in double (-inf for a zero cell)
__host__ __device__ inline double cell_log(float m, int e) {
    if (!(m > 0.0f)) return kNegInf;
    return (static_cast<double>(e) + log2(static_cast<double>(m))) * kLn2D;
}

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

// Development aid (tools/timeline_probe.cu, -DMRNNT_TIMELINE): %globaltimer stamps of the three kernels of a call IN THE
// STREAM -- first CTA in, first CTA past its wait for the predecessor, last warp out -- per call slot, so that the gaps
// between the kernels of back-to-back calls can be read off.  Compiled out otherwise.
#ifdef MRNNT_TIMELINE
constexpr int kTimelineSlots = 16, kTimelineEvents = 16;
__device__ unsigned long long g_timeline[kTimelineSlots][kTimelineEvents];
__device__ __forceinline__ unsigned long long timeline_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define MRNNT_TL_MIN(slot, ev) atomicMin(&g_timeline[(slot) & (kTimelineSlots - 1)][ev], timeline_now())
#define MRNNT_TL_MAX(slot, ev) atomicMax(&g_timeline[(slot) & (kTimelineSlots - 1)][ev], timeline_now())
#else
#define MRNNT_TL_MIN(slot, ev) ((void)0)
#define MRNNT_TL_MAX(slot, ev) ((void)0)
#endif

// L2 eviction policies for the bulk copies (streamed-once data should not displace reusable lines)
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
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
// Element types of the logits / gradients.  The reference computes on float32 only
// (pytorch_binding/monotonic_rnnt.cu:84); bfloat16 is an extension that halves the bytes the two
// streaming kernels move.  All arithmetic stays in float: a 16-byte vector is unpacked to kPerVec floats on
// arrival and packed (round to nearest even) on the way out.
// ---------------------------------------------------------------------------------------------
template <typename E>
struct Elem;
template <>
struct Elem<float> {
    static constexpr int kPerVec = 4;
    static __device__ __forceinline__ void unpack(const uint4 &r, float (&f)[4]) {
        f[0] = __uint_as_float(r.x);
        f[1] = __uint_as_float(r.y);
        f[2] = __uint_as_float(r.z);
        f[3] = __uint_as_float(r.w);
    }
    static __device__ __forceinline__ uint4 pack(const float (&f)[4]) {
        return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
    }
    static __device__ __forceinline__ float to_float(float v) { return v; }
    static __device__ __forceinline__ float from_float(float v) { return v; }
};
template <>
struct Elem<__nv_bfloat16> {
    static constexpr int kPerVec = 8;
    static __device__ __forceinline__ void unpack(const uint4 &r, float (&f)[8]) {
        const unsigned w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // a bfloat16 is the upper half of the float with the same value
            f[2 * i] = __uint_as_float(w[i] << 16);
            f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ uint4 pack(const float (&f)[8]) {
        unsigned w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 p = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
            w[i] = *reinterpret_cast<const unsigned *>(&p);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
    static __device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
    static __device__ __forceinline__ __nv_bfloat16 from_float(float v) { return __float2bfloat16_rn(v); }
};

// ---------------------------------------------------------------------------------------------
// programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// start (launch latency, prologue) while its predecessor in the stream is still draining; it must not touch the
// predecessor's results before pdl_wait() (which returns once the predecessor has completed and flushed).  Both
// are no-ops in a kernel launched the ordinary way.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// streaming global stores / loads
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_stream_u4(uint4 *p, const uint4 &v) {
    asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
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

// 2^-k for 0 <= k < 127; exactly 0 for k >= 127 (the scaled term is then either an exact zero or ~2^-100 below
// the other operand's resolution)
__device__ __forceinline__ float pow2_neg(unsigned k) {
    return __int_as_float(static_cast<int>(min(k, 127u)) * -0x00800000 + 0x3f800000);
}

// (m1,e1) + (m2,e2) in the scaled linear domain.  Inputs are non-negative; zeros carry exponents around
// kZeroExp (see above), so the larger exponent is always the right common exponent.
__device__ __forceinline__ Cell cell_add(float m1, int e1, float m2, int e2) {
    const int d = e1 - e2;
    const bool first = d >= 0;
    const float big = first ? m1 : m2;
    const float small = first ? m2 : m1;
    Cell r;
    r.m = fmaf(small, pow2_neg(static_cast<unsigned>(abs(d))), big);
    r.e = max(e1, e2);
    return r;
}

// bring the mantissa back to [1,2) (exact: only exponent bits move); zeros become the canonical zero
__device__ __forceinline__ void cell_renorm(float &m, int &e) {
    const int bits = __float_as_int(m);
    const bool zero = (m == 0.0f);
    e = zero ? kZeroExp : e + (bits >> 23) - 127;
    m = zero ? 0.0f : __int_as_float((bits & 0x007fffff) | 0x3f800000);
}

// Transition weight p = 2^(x * log2(e) + dh + dl) as m * 2^e with m in [2^-1/2, 2^1/2], in float arithmetic
// only (the lattice kernel converts two weights per row on the fly; double-precision instructions are too
// slow for that).  y = x*log2(e) + dh + dl is formed as an unevaluated float pair (exact product split,
// error-free sum), its nearest integer becomes the exponent, and 2^(fraction) is a degree-7 Taylor
// polynomial (truncation 5e-9; the result carries about one float ulp of rounding).  `ok == false`, -inf, NaN
// and anything below 2^-(2^22) give the zero weight.
__device__ __forceinline__ void weight_from_logit(float x, float dh, float dl, bool ok, float &m, int &e) {
    constexpr float kL1 = kLog2e;    // float(log2 e)
    constexpr float kL2 = kLog2eLo;  // log2 e - kL1
    constexpr float kMagic = 12582912.0f;              // 1.5 * 2^23: adding it rounds to the nearest integer
    const float ph = x * kL1;
    const float pl = fmaf(x, kL2, fmaf(x, kL1, -ph));
    float s, err;
    two_sum(ph, dh, s, err);
    const float lo = (err + pl) + dl;
    const float r = s + kMagic;
    const float n = r - kMagic;
    const float f = (s - n) + lo;  // in [-1/2, 1/2] (+- a few ulp)
    float p = 1.5252733804059841e-05f;
    p = fmaf(p, f, 1.5403530393381608e-04f);
    p = fmaf(p, f, 1.3333558146428443e-03f);
    p = fmaf(p, f, 9.6181291076284770e-03f);
    p = fmaf(p, f, 5.5504108664821580e-02f);
    p = fmaf(p, f, 2.4022650695910072e-01f);
    p = fmaf(p, f, 6.9314718055994531e-01f);
    p = fmaf(p, f, 1.0f);
    ok = ok && (s > -4194304.0f) && (s < 4194304.0f);
    m = ok ? p : 0.0f;
    e = ok ? __float_as_int(r) - __float_as_int(kMagic) : kZeroExp;
}

// log2 of a positive finite float as an integer part and a float in [0,1) (absolute error ~6e-8)
__device__ __forceinline__ void log2_parts(float v, int &ip, float &fp) {
    int ex;
    const float f = frexpf(v, &ex);  // v = f * 2^ex, f in [0.5, 1)
    ip = ex - 1;
    fp = log2f(f + f);
}

// From a row's max (times kLog2e, rounded once and used for every term) and its sum of 2^(x kLog2e - ML): the base-2
// denominator D = -log2 sum_v exp(x[v]) as an unevaluated sum of two floats (error-free additions; the only error is
// log2f's ~6e-8 on a value in [0,1)), such that log2 p(v) = x[v] * (kLog2e + kLog2eLo) + D.
// K1 forms its sum with kLog2e alone: sum_v 2^(x kLog2e - ML) = 2^-ML sum_v e^x 2^(-x kLog2eLo), and the last factor
// is 2^(-max kLog2eLo) for every term that matters (|x - max| kLog2eLo < 1e-6 for the terms within e^-50 of the
// largest).  Hence D = -(ML + log2 sum) - max * kLog2eLo; without the last term log2 p is off by max * 1.9e-8 per row,
// which does not cancel between rows of different maxima (logits of magnitude 100: 2e-6 per row, T = 800 rows per path).
struct Denominator {
    float hi, lo;
};
__device__ __forceinline__ Denominator lse_finish(float ML, float sum) {
    int ip;
    float fp, h, l1, h2, l2;
    log2_parts(sum, ip, fp);
    two_sum(ML, static_cast<float>(ip), h, l1);
    two_sum(h, fp, h2, l2);
    Denominator d;
    d.hi = -h2;
    d.lo = -fmaf(ML, kLog2eLoRel, l1 + l2);  // (max * kLog2eLo = ML * kLog2eLo / kLog2e)
    return d;
}

}  // namespace mrnnt
