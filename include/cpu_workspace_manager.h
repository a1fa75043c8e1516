// CpuRNNTWorkspaceManager<T> -- NAME-ONLY compatibility shell.
//
// The reference's PyTorch binding includes cpu_workspace_manager.h / cpu_rnnt.h unconditionally and
// defines cpu_monotonic_rnnt* around them (pytorch_binding/monotonic_rnnt.cu:12-13,16-77), so the
// class must exist for that translation unit to compile.  This library is GPU-only by design (no CPU
// fallback): every operation here reports RNNT_STATUS_EXECUTION_FAILED, which the bindings turn into
// an exception.  The reference's real CPU implementation (include/cpu_workspace_manager.h,
// include/cpu_rnnt.h) is used by this repository only as the test oracle, from outside the product.
#pragma once
#ifndef MONOTONIC_RNNT_CPU_WORKSPACE_MANAGER_H
#define MONOTONIC_RNNT_CPU_WORKSPACE_MANAGER_H

#include <cstddef>

#include "status.h"
#include "workspace_manager.h"

template <typename dtype>
class CpuRNNTWorkspaceManager : public RNNTWorkspaceManager {
   public:
    explicit CpuRNNTWorkspaceManager(const dtype *const, const int *const, const int, const int *, const int *,
                                     const int) {}
    CpuRNNTWorkspaceManager(const CpuRNNTWorkspaceManager &) = delete;
    ~CpuRNNTWorkspaceManager() override = default;

    RNNTStatus get_workspace_size(size_t *size_bytes) const {
        if (size_bytes != nullptr) *size_bytes = 0;
        return RNNT_STATUS_EXECUTION_FAILED;
    }
    void set_workspace(void *) {}
    RNNTStatus create_workspace() { return RNNT_STATUS_EXECUTION_FAILED; }
    void free_workspace() {}
    void restrict_to_alignment(const int *const, int, int) {}
};

#endif  // MONOTONIC_RNNT_CPU_WORKSPACE_MANAGER_H
