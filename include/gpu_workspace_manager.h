// GpuRNNTWorkspaceManager<float> -- source-compatible with the reference class of the same name
// (reference include/gpu_workspace_manager.h:14-345), so that the reference's framework bindings
// compile against this directory unchanged:
//   pytorch_binding/monotonic_rnnt.cu:99-111,135-149     (create_workspace / restrict_to_alignment / free_workspace)
//   tensorflow_binding/monotonic_rnnt_op.cu:103-128      (get_workspace_size / set_workspace)
//   tests/test_gpu.cu:62-73                              (one manager reused across several cost() calls)
//
// Contract kept: constructor arguments and their residency (acts, labels, T, S are DEVICE pointers,
// B and V host ints), validation rules of get_workspace_size (B <= 0, T_b <= 0, S_b < 0, T_b < S_b ->
// RNNT_STATUS_INVALID_VALUE), workspace = one opaque device buffer whose size depends on (B, T[], S[])
// only, caller-owned via set_workspace or manager-owned via create_workspace/free_workspace.
// Not kept (private to the reference's own gpu_rnnt.h): the public data members and *_host() debug
// mirrors, and the byte count itself (ours is larger: 64-bit offsets, mantissa/exponent lattice cells,
// transition weights, per-row coefficients -- see mrnnt_b200/plan.cuh).
//
// Host synchronisation: the FIRST of get_workspace_size / create_workspace / set_workspace copies
// T[] and S[] to the host once (2*B ints); nothing else in this class blocks.  The reference does
// ~14 blocking copies here (gpu_workspace_manager.h:87-95,228-329).
// restrict_to_alignment only records its arguments; the band is built on the device, on the compute
// stream, by the next cost()/cost_and_grad() -- `alignments` must stay valid until then (it does in
// every caller above).
#pragma once
#ifndef MONOTONIC_RNNT_GPU_WORKSPACE_MANAGER_H
#define MONOTONIC_RNNT_GPU_WORKSPACE_MANAGER_H

#include <cstddef>
#include <type_traits>

#include "mrnnt_b200/engine.cuh"
#include "status.h"
#include "workspace_manager.h"

template <typename dtype>
class GpuRNNTWorkspaceManager : public RNNTWorkspaceManager {
    static_assert(std::is_same<dtype, float>::value,
                  "the sm_100a path computes on float logits only (as the reference bindings do)");

   public:
    explicit GpuRNNTWorkspaceManager(const dtype *const acts, const int *const labels, const int B, const int *T,
                                     const int *S, const int V)
        : engine_(acts, labels, B, T, S, V) {}

    GpuRNNTWorkspaceManager(const GpuRNNTWorkspaceManager &) = delete;
    ~GpuRNNTWorkspaceManager() override = default;

    // Optional fast path with no reference equivalent: hand over host copies of T[] and S[] so that
    // not even the one blocking copy happens.
    RNNTStatus set_host_lengths(const int *T_host, const int *S_host) {
        return engine_.set_host_lengths(T_host, S_host);
    }

    // Extension with no reference equivalent: acts / gradients are the joint network's padded
    // [B, T_dim, U, V] tensor (U = label positions + 1) and labels is [B, label_stride]; call before anything else.
    void set_padded_layout(int T_dim, int U, int label_stride) { engine_.set_padded_layout(T_dim, U, label_stride); }

    RNNTStatus get_workspace_size(size_t *size_bytes) const { return engine_.workspace_size(size_bytes); }

    void set_workspace(void *workspace) { (void)engine_.set_workspace(workspace); }

    RNNTStatus create_workspace() { return engine_.create_workspace(); }

    void free_workspace() { engine_.free_workspace(); }

    void restrict_to_alignment(const int *const alignments, int max_shift, int blank_idx) {
        engine_.restrict_to_alignment(alignments, max_shift, blank_idx);
    }

    [[nodiscard]] int B_host() const { return engine_.B(); }

    mrnnt::Engine &engine() { return engine_; }

   private:
    mutable mrnnt::Engine engine_;  // get_workspace_size is const in the reference API but caches T/S
};

#endif  // MONOTONIC_RNNT_GPU_WORKSPACE_MANAGER_H
