// Status codes of the monotonic RNN-T loss API.
//
// ABI contract: the enumerator VALUES (0..4) and the spelling of the enumerators and of
// rnntGetStatusString are what callers of the reference compile against
// (reference include/status.h:4-31; used at pytorch_binding/monotonic_rnnt.cu:105,109 and
// tensorflow_binding/monotonic_rnnt_op.cu:114-116).  Everything else in this file is ours.
#pragma once
#ifndef MONOTONIC_RNNT_STATUS_H
#define MONOTONIC_RNNT_STATUS_H

typedef enum {
    RNNT_STATUS_SUCCESS = 0,           // costs (and gradients) are valid
    RNNT_STATUS_MEMOPS_FAILED = 1,     // a cudaMalloc / cudaMemcpy / cudaMemset reported an error
    RNNT_STATUS_INVALID_VALUE = 2,     // B <= 0, T_b <= 0, S_b < 0, T_b < S_b, null costs, wrong manager type ...
    RNNT_STATUS_EXECUTION_FAILED = 3,  // a kernel launch or the stream reported an error; also: no CPU path here
    RNNT_STATUS_UNKNOWN_ERROR = 4
} RNNTStatus;

// Human-readable text for a status (same strings as the reference so log scrapers keep working).
static inline const char *rnntGetStatusString(RNNTStatus status) {
    static const char *const text[] = {"no error", "cuda memcpy or memset failed", "invalid value",
                                       "execution failed", "unknown error"};
    const int i = (int)status;
    return text[(i >= 0 && i < 4) ? i : 4];
}

#endif  // MONOTONIC_RNNT_STATUS_H
