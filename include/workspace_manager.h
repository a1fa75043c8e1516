// Polymorphic base of the workspace managers: compute_rnnt_loss receives one by reference and
// down-casts it (reference include/workspace_manager.h:4-11, src/rnnt_entrypoint.cpp:24,35).
// Non-copyable, virtual destructor -- that is the whole contract.
#pragma once
#ifndef MONOTONIC_RNNT_WORKSPACE_MANAGER_H
#define MONOTONIC_RNNT_WORKSPACE_MANAGER_H

class RNNTWorkspaceManager {
   public:
    RNNTWorkspaceManager() = default;
    virtual ~RNNTWorkspaceManager() = default;

    RNNTWorkspaceManager(const RNNTWorkspaceManager &) = delete;
    RNNTWorkspaceManager &operator=(const RNNTWorkspaceManager &) = delete;
};

#endif  // MONOTONIC_RNNT_WORKSPACE_MANAGER_H
