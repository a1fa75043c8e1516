// CpuRNNTComputer<T> -- NAME-ONLY compatibility shell, see cpu_workspace_manager.h in this directory.
// There is no CPU implementation of the loss in this library; both entry points fail loudly.
#pragma once
#ifndef MONOTONIC_RNNT_CPU_RNNT_H
#define MONOTONIC_RNNT_CPU_RNNT_H

#include "cpu_workspace_manager.h"
#include "status.h"

template <typename ProbT>
class CpuRNNTComputer {
   public:
    CpuRNNTComputer(CpuRNNTWorkspaceManager<ProbT> &, int /*blank*/, int /*num_threads*/) {}
    CpuRNNTComputer(const CpuRNNTComputer &) = delete;
    CpuRNNTComputer &operator=(const CpuRNNTComputer &) = delete;

    RNNTStatus cost_and_grad(ProbT *, ProbT *) { return RNNT_STATUS_EXECUTION_FAILED; }
    RNNTStatus cost(ProbT *) { return RNNT_STATUS_EXECUTION_FAILED; }
};

#endif  // MONOTONIC_RNNT_CPU_RNNT_H
