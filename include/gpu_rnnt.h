// GpuRNNTComputer<float> -- source-compatible with the reference class (include/gpu_rnnt.h:18-251):
//   GpuRNNTComputer(GpuRNNTWorkspaceManager<float>&, int blank, CUstream stream)
//   RNNTStatus cost_and_grad(float *costs_HOST, float *grads_DEVICE)   (grads == nullptr -> cost only)
//   RNNTStatus cost(float *costs_HOST)
// Used exactly so by pytorch_binding/monotonic_rnnt.cu:107-109, tensorflow_binding/monotonic_rnnt_op.cu:123-125
// and tests/test_gpu.cu:68-71.
//
// costs are valid when the call returns (one cudaStreamSynchronize on `stream`; the reference's final
// blocking cudaMemcpy on the legacy stream, gpu_rnnt.h:229-232, is replaced by an async copy on the SAME
// stream, which also removes its ordering hazard with non-blocking streams).  Every element of `grads`
// is written exactly once, so callers need not pre-zero it (the TensorFlow op relies on that).
// Launch errors are reported (the reference always returns SUCCESS on the GPU path).
#pragma once
#ifndef MONOTONIC_RNNT_GPU_RNNT_H
#define MONOTONIC_RNNT_GPU_RNNT_H

#include "gpu_workspace_manager.h"
#include "options.h"
#include "status.h"

template <typename ProbT>
class GpuRNNTComputer {
   public:
    GpuRNNTComputer(GpuRNNTWorkspaceManager<ProbT> &workspace_manager, int blank, CUstream stream)
        : workspace_manager_(workspace_manager), blank_(blank), stream_(stream) {}

    GpuRNNTComputer(const GpuRNNTComputer &) = delete;
    GpuRNNTComputer &operator=(const GpuRNNTComputer &) = delete;

    RNNTStatus cost_and_grad(ProbT *costs, ProbT *grads) {
        return workspace_manager_.engine().compute(blank_, reinterpret_cast<cudaStream_t>(stream_), costs, grads);
    }

    RNNTStatus cost(ProbT *costs) { return cost_and_grad(costs, nullptr); }

   private:
    GpuRNNTWorkspaceManager<ProbT> &workspace_manager_;
    int blank_;
    CUstream stream_;
};

#endif  // MONOTONIC_RNNT_GPU_RNNT_H
