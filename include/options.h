// Call options of compute_rnnt_loss.
//
// ABI contract (reference include/options.h:5-24): field order and types of RNNTOptions
// {int num_threads; CUstream stream; int blank_label; rnntComputeLocation loc;} -- 24 bytes on
// LP64, passed BY VALUE to compute_rnnt_loss -- and the values RNNT_CPU = 0 / RNNT_GPU = 1.
#pragma once
#ifndef MONOTONIC_RNNT_OPTIONS_H
#define MONOTONIC_RNNT_OPTIONS_H

// CUDA's own opaque stream handle type, re-declared so that C/C++ callers do not need cuda.h
// (cudaStream_t is the same pointer type).
typedef struct CUstream_st *CUstream;

typedef enum {
    RNNT_CPU = 0,  // not implemented by this library: there is deliberately no CPU fallback
    RNNT_GPU = 1   // sm_100a kernels on `stream`
} rnntComputeLocation;

struct RNNTOptions {
    int num_threads;          // ignored on the GPU path (the reference only uses it for OpenMP)
    CUstream stream;          // all kernels and the final costs copy are ordered on this stream
    int blank_label;          // index of the blank symbol in [0, V)
    rnntComputeLocation loc;  // must be RNNT_GPU
};

#endif  // MONOTONIC_RNNT_OPTIONS_H
