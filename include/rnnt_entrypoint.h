// C entry point of libmonotonic_rnnt.so (sm_100a build).
//
// Same symbol, signature and semantics as the reference's include/rnnt_entrypoint.h:24-25 /
// src/rnnt_entrypoint.cpp:16-48:
//   * `workspace_manager` must be a GpuRNNTWorkspaceManager<float> whose workspace has been
//     created (create_workspace) or handed over (set_workspace);
//   * `options` is passed by value; options.loc must be RNNT_GPU (RNNT_CPU returns
//     RNNT_STATUS_EXECUTION_FAILED: this library has no CPU path);
//   * `costs` is a HOST array of B floats, valid on return (the call synchronises the stream once);
//   * `gradients` is a DEVICE array shaped like the packed logits, or nullptr for cost only;
//   * costs == nullptr or an unknown loc / wrong manager type -> RNNT_STATUS_INVALID_VALUE.
// The flat, handle-based C ABI that carries no C++ types lives in mrnnt_c_api.h.
#pragma once
#ifndef MONOTONIC_RNNT_ENTRYPOINT_H
#define MONOTONIC_RNNT_ENTRYPOINT_H

#include "options.h"
#include "status.h"
#include "workspace_manager.h"

extern "C" {

RNNTStatus compute_rnnt_loss(RNNTWorkspaceManager &workspace_manager, RNNTOptions options, float *costs,
                             float *gradients);

}  // extern "C"

#endif  // MONOTONIC_RNNT_ENTRYPOINT_H
