// TEST INFRASTRUCTURE ONLY -- NOT PART OF THE PRODUCT.
//
// Thin extern "C" shim around the UNMODIFIED reference CPU implementation.  It
// is compiled from the reference sources where they lie (`-I$RNNT_REF_DIR/include`,
// plus $RNNT_REF_DIR/tests/random.cpp for the reference's fixture generators) by
// oracle/Makefile into oracle/_ref/libmrnnt_ref.so.  No reference source is copied
// into this repository; this file only *calls* the reference's public classes
//   CpuRNNTWorkspaceManager<T>  (include/cpu_workspace_manager.h:33-57, :207-224, :264-277)
//   CpuRNNTComputer<T>          (include/cpu_rnnt.h:27-94)
// exactly as tests/test_cpu.cpp:47-56 and pytorch_binding/monotonic_rnnt.cu:29-41 do.
//
// Uses: (1) generating tests/golden/ (tests/golden/make_golden.py), (2) pinning
// oracle/rnnt_oracle.c, (3) the timed CPU baseline (`cpu_baseline.kind ==
// "reference"`, `bench.py --impl reference`).
//
// The reference indexes with `int` (cpu_workspace_manager.h:48,125-134) and
// overflows above 2^31-1 logits; batches larger than that are run one utterance
// at a time (utterances are independent and packed), see SURVEY D5.
#include <climits>
#include <cstdint>
#include <vector>

#include "cpu_rnnt.h"
#include "cpu_workspace_manager.h"

// reference fixture generators, tests/random.cpp:13-37
void genActs(std::vector<float> &arr);
std::vector<int> genLabels(int V, int S);

namespace {

template <typename T>
int run_batch(const T *acts, const int *labels, int B, const int *Tl, const int *Sl, int V, int blank,
              const int *alignment, int max_shift, int align_blank, int num_threads, T *costs, T *grads) {
    CpuRNNTWorkspaceManager<T> wm(acts, labels, B, Tl, Sl, V);
    RNNTStatus st = wm.create_workspace();
    if (st != RNNT_STATUS_SUCCESS) return static_cast<int>(st);
    if (alignment != nullptr) wm.restrict_to_alignment(alignment, max_shift, align_blank);
    CpuRNNTComputer<T> computer(wm, blank, num_threads);
    st = grads != nullptr ? computer.cost_and_grad(costs, grads) : computer.cost(costs);
    wm.free_workspace();
    return static_cast<int>(st);
}

template <typename T>
int run(const T *acts, const int *labels, int B, const int *Tl, const int *Sl, int V, int blank,
        const int *alignment, int max_shift, int align_blank, int num_threads, T *costs, T *grads) {
    if (B <= 0) return RNNT_STATUS_INVALID_VALUE;
    int64_t total = 0;
    int S_max = 0, T_max = 0;
    for (int b = 0; b < B; ++b) {
        if (Tl[b] <= 0 || Sl[b] < 0 || Tl[b] < Sl[b]) return RNNT_STATUS_INVALID_VALUE;
        total += static_cast<int64_t>(Tl[b]) * (Sl[b] + 1) * V;
        S_max = Sl[b] > S_max ? Sl[b] : S_max;
        T_max = Tl[b] > T_max ? Tl[b] : T_max;
    }
    if (total <= INT_MAX) {
        return run_batch<T>(acts, labels, B, Tl, Sl, V, blank, alignment, max_shift, align_blank, num_threads,
                            costs, grads);
    }
    // > 2^31-1 logits: one reference call per utterance (B=1 slices), utterances in parallel.
    std::vector<int64_t> start(B + 1, 0);
    for (int b = 0; b < B; ++b) start[b + 1] = start[b] + static_cast<int64_t>(Tl[b]) * (Sl[b] + 1) * V;
    int rc = 0;
#ifndef RNNT_DISABLE_OMP
    if (num_threads > 0) omp_set_num_threads(num_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        int r = run_batch<T>(acts + start[b], labels + static_cast<int64_t>(b) * S_max, 1, Tl + b, Sl + b, V, blank,
                             alignment ? alignment + static_cast<int64_t>(b) * T_max : nullptr, max_shift,
                             align_blank, 1, costs + b, grads ? grads + start[b] : nullptr);
        if (r != 0) {
#pragma omp critical
            rc = r;
        }
    }
    return rc;
}

}  // namespace

extern "C" {

int mrnnt_ref_f32(const float *acts, const int *labels, int B, const int *T, const int *S, int V, int blank,
                  const int *alignment, int max_shift, int align_blank, int num_threads, float *costs,
                  float *grads) {
    return run<float>(acts, labels, B, T, S, V, blank, alignment, max_shift, align_blank, num_threads, costs, grads);
}

int mrnnt_ref_f64(const double *acts, const int *labels, int B, const int *T, const int *S, int V, int blank,
                  const int *alignment, int max_shift, int align_blank, int num_threads, double *costs,
                  double *grads) {
    return run<double>(acts, labels, B, T, S, V, blank, alignment, max_shift, align_blank, num_threads, costs,
                       grads);
}

// reference's own fixture generators (mt19937 seeds 0 / 1), tests/random.cpp:13-37
void mrnnt_ref_gen_acts(float *out, int n) {
    std::vector<float> v(n);
    genActs(v);
    for (int i = 0; i < n; ++i) out[i] = v[i];
}

void mrnnt_ref_gen_labels(int V, int S, int *out) {
    std::vector<int> l = genLabels(V, S);
    for (int i = 0; i < S; ++i) out[i] = l[i];
}

int mrnnt_ref_num_threads(void) {
#ifndef RNNT_DISABLE_OMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
