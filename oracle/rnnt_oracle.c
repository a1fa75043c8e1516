/* ============================================================================
 * TEST INFRASTRUCTURE ONLY -- NOT PART OF THE PRODUCT.
 *
 * CPU restatement ("oracle") of the monotonic RNN-T loss-and-gradient path of
 * SimBe195/monotonic-rnnt, in plain C.  It follows, function by function, the
 * reference's CPU implementation:
 *     include/cpu_rnnt.h                (denominators :98-115, alphas :155-183,
 *                                        betas :185-214, grads :216-236,
 *                                        drivers :42-94, :254-263)
 *     include/cpu_workspace_manager.h   (layout :33-57, band limits :67-86,
 *                                        validation :99-107, accessors :161-205,
 *                                        restrict_to_alignment :207-224)
 *     include/rnnt_helper.h             (log_sum_exp :21-30)
 * Nothing is copied; the arithmetic order is restated so that the float build is
 * bit-identical to the reference's `CpuRNNTComputer<float>` (checked in
 * tests/test_oracle_cpu.py against tests/golden/, which was produced by the
 * reference itself through oracle/ref_shim.cpp -> oracle/_ref/).
 *
 * PARITY PINNED: yes -- against the reference's own golden values
 * (tests/test_cpu.cpp fixtures: costs -log{0.363,0.39,0.072,0.2958,0.0672,0.192},
 * the README gradients) and against outputs of the compiled reference on seeded
 * random ragged batches, with and without alignment restriction.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library, and only as the checker or the
 * timed CPU baseline.  The product (include/, monotonic-rnnt_b200/) never calls
 * it and has no CPU fallback.
 *
 * Build: see oracle/Makefile  (gcc -O2 -fopenmp -shared -fPIC).
 * ==========================================================================*/
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

/* ---- float instantiation: mirrors CpuRNNTComputer<float> ------------------*/
#define REAL float
#define FN(name) CAT(name, _f32)
#define REAL_EXP(x) expf(x)
#include "rnnt_oracle_body.inc"
#undef REAL
#undef FN
#undef REAL_EXP

/* ---- double instantiation: mirrors CpuRNNTComputer<double>, the "truth" ---*/
#define REAL double
#define FN(name) CAT(name, _f64)
#define REAL_EXP(x) exp(x)
#include "rnnt_oracle_body.inc"
#undef REAL
#undef FN
#undef REAL_EXP

/* Double-precision truth for float inputs: widens the logits once and runs the
 * f64 path; used by the tests to measure distance from exact arithmetic. */
int mrnnt_oracle_f64_from_f32(const float *acts, const int *labels, int B, const int *T, const int *S, int V,
                              int blank, const int *alignment, int max_shift, int align_blank, int num_threads,
                              double *costs, double *grads, double *denom_out, double *alpha_out,
                              double *beta_out, double *ll_backward_out) {
    if (B <= 0) return 2;
    int64_t rows = 0;
    for (int b = 0; b < B; ++b) {
        if (T[b] <= 0 || S[b] < 0 || T[b] < S[b]) return 2;
        rows += (int64_t)T[b] * (S[b] + 1);
    }
    const int64_t n = rows * V;
    double *wide = (double *)malloc((size_t)n * sizeof(double));
    if (!wide) return 1;
    for (int64_t i = 0; i < n; ++i) wide[i] = (double)acts[i];
    int rc = mrnnt_oracle_f64(wide, labels, B, T, S, V, blank, alignment, max_shift, align_blank, num_threads,
                              costs, grads, denom_out, alpha_out, beta_out, ll_backward_out);
    free(wide);
    return rc;
}

int mrnnt_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
