/* Flat C ABI of libmonotonic_rnnt.so (sm_100a).  Plain pointers, ints and sizes only -- this is the
 * boundary a foreign-function binding (ctypes, cgo, JNI, a C program) links against.  The C++
 * surface the reference's own bindings compile against (GpuRNNTWorkspaceManager / GpuRNNTComputer /
 * compute_rnnt_loss) sits on the same engine; see INTEGRATION.md.
 *
 * Pointer residency everywhere (reference include/gpu_workspace_manager.h:63-69, gpu_rnnt.h:229):
 *   DEVICE: acts, labels, T_dev, S_dev, alignments, gradients, workspace        HOST: costs, *_host
 * Layouts (reference include/cpu_workspace_manager.h:46-49,117-135, pytorch_binding/monotonic_rnnt_op.py:133-140):
 *   acts       float32 [sum_b T_b*(S_b+1), V], utterances concatenated without padding,
 *              row of (b,t,s) = row_start(b) + t*(S_b+1) + s
 *   labels     int32 [B, max_b S_b]          alignments int32 [B, max_b T_b]
 *   T, S       int32 [B]                     costs float32 [B]      gradients: same shape as acts
 * All functions return an RNNTStatus (include/status.h).  None of them falls back to the CPU.
 */
#ifndef MONOTONIC_RNNT_B200_C_API_H
#define MONOTONIC_RNNT_B200_C_API_H

#include <stddef.h>
#include <stdint.h>

#include "status.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mrnnt_handle_st *mrnnt_handle_t;

/* Host-only size query.  Replaces GpuRNNTWorkspaceManager<float>::get_workspace_size
 * (reference include/gpu_workspace_manager.h:228-254) for callers that hold the lengths on the host.
 * Same validation: B <= 0, T_b <= 0, S_b < 0 or T_b < S_b -> RNNT_STATUS_INVALID_VALUE. */
RNNTStatus mrnnt_get_workspace_size(const int *T_host, const int *S_host, int B, int V, size_t *size_bytes);
/* The same under the name the reference's manager uses for it (gpu_workspace_manager.h:228): */
RNNTStatus get_workspace_size(const int *T_host, const int *S_host, int B, int V, size_t *size_bytes);

/* Handle = one GpuRNNTWorkspaceManager<float> (reference include/gpu_workspace_manager.h:33-52).
 * T_host / S_host may be NULL; the lengths are then fetched from the device once (blocking). */
RNNTStatus mrnnt_create(mrnnt_handle_t *out, const float *acts, const int *labels, int B, const int *T_dev,
                        const int *S_dev, int V, const int *T_host, const int *S_host);
/* Same, for the joint network's own PADDED tensor (no reference equivalent; the reference makes its callers
 * gather the valid rows into the packed layout first, pytorch_binding/monotonic_rnnt_op.py:133-140, and scatter
 * the gradients back):
 *   acts / gradients  float32 [B, T_dim, U, V], row of (b,t,s) = (b*T_dim + t)*U + s,  T_dim >= max T_b, U >= max S_b + 1
 *   labels            int32   [B, label_stride], label_stride >= max S_b;   alignments stay [B, max_b T_b]
 * Rows with t >= T_b or s > S_b are never read and their gradient rows are written as zeros. */
RNNTStatus mrnnt_get_workspace_size_padded(const int *T_host, const int *S_host, int B, int V, int T_dim, int U,
                                           int label_stride, size_t *size_bytes);
RNNTStatus mrnnt_create_padded(mrnnt_handle_t *out, const float *acts, const int *labels, int B, const int *T_dev,
                               const int *S_dev, int V, int T_dim, int U, int label_stride, const int *T_host,
                               const int *S_host);
/* Extension (no reference equivalent; the reference is float32 only, pytorch_binding/monotonic_rnnt.cu:84):
 * dtype 1 declares acts AND gradients of this handle to be bfloat16 arrays (same layouts; pass the pointers through
 * the float* parameters); all arithmetic stays float32, costs stay float32.  dtype 0 = float32 (default).
 * Call right after mrnnt_create / mrnnt_create_padded. */
enum { MRNNT_DTYPE_F32 = 0, MRNNT_DTYPE_BF16 = 1 };
RNNTStatus mrnnt_set_dtype(mrnnt_handle_t h, int dtype);
void mrnnt_destroy(mrnnt_handle_t h);

/* gpu_workspace_manager.h:228 (get_workspace_size), :256 (set_workspace), :331-342 (create/free). */
RNNTStatus mrnnt_workspace_size(mrnnt_handle_t h, size_t *size_bytes);
RNNTStatus mrnnt_set_workspace(mrnnt_handle_t h, void *workspace);
RNNTStatus mrnnt_create_workspace(mrnnt_handle_t h);
void mrnnt_free_workspace(mrnnt_handle_t h);
/* Blocks handed back by mrnnt_free_workspace / GpuRNNTWorkspaceManager::free_workspace are kept for the next
 * create_workspace of this library instead of going through cudaFree + cudaMalloc (the reference's torch binding
 * allocates and frees the workspace on every loss call: 2.9 ms against 0.34 ms on B=32 T=150 S=40 V=1000).  At most 4
 * blocks and `bytes` bytes are kept (default 1 GiB); 0 turns the cache off, free_workspace is then a cudaFree as in the
 * reference.  mrnnt_trim_workspace_cache returns everything that is cached to the driver now. */
void mrnnt_set_workspace_cache_limit(size_t bytes);
void mrnnt_trim_workspace_cache(void);

/* gpu_workspace_manager.h:191-219.  Recorded here, applied on the device by the next compute call on
 * that call's stream; `alignments` must stay valid until then.  May be called repeatedly. */
RNNTStatus mrnnt_restrict_to_alignment(mrnnt_handle_t h, const int *alignments, int max_shift, int blank_idx);
/* Same for an alignment array that is wider than the reference's [B, max_b T_b] (e.g. [B, T_dim] next to a padded acts
 * tensor): `stride` ints per utterance, >= max_b T_b -- a smaller stride makes the next compute call return
 * RNNT_STATUS_INVALID_VALUE (the reference mis-indexes such an array silently, cpu_workspace_manager.h:208). */
RNNTStatus mrnnt_restrict_to_alignment_strided(mrnnt_handle_t h, const int *alignments, int stride, int max_shift,
                                               int blank_idx);
/* max_b T_b, max_b S_b and the row count of acts as this handle understands them (from the lengths it was given or has
 * fetched): what a binding needs to check the second dimension of its labels / alignment tensors against, which the
 * reference never does. */
RNNTStatus mrnnt_get_shape(mrnnt_handle_t h, int *T_max, int *S_max, int64_t *rows);

/* No counterpart in the reference, whose callers bring the logits to the device themselves (pytorch_binding/
 * monotonic_rnnt.cu:85-88 takes CUDA tensors): fill the handle's device `acts` from PINNED host memory of the same layout
 * and type, moving only the rows the lattice reads (the plan's dead rows -- a quarter of a plain batch, ~95 % under a
 * tight alignment band -- are never read by any kernel and stay as they are).  Asynchronous on `stream`; needs the
 * workspace; call after mrnnt_restrict_to_alignment.  RNNT_STATUS_INVALID_VALUE if `host_acts` is not pinned. */
RNNTStatus mrnnt_upload_acts(mrnnt_handle_t h, const void *host_acts, void *stream);

/* GpuRNNTComputer<float>::cost_and_grad / cost (reference include/gpu_rnnt.h:27-235) and therefore
 * compute_rnnt_loss (include/rnnt_entrypoint.h:24-25).  `stream` is a cudaStream_t / CUstream.
 * gradients == NULL selects cost only.  costs are valid on return; with gradients the call returns as soon as they
 * are (MRNNT_OPT_RETURN_EARLY below): the gradients are complete in stream order. */
RNNTStatus mrnnt_cost_and_grad(mrnnt_handle_t h, int blank_label, void *stream, float *costs_host, float *gradients);

/* Same work, no host synchronisation: costs stay on the device (mrnnt_device_costs) for callers that
 * consume them there (e.g. an NCCL all-reduce of the summed cost on the same stream).
 * Not for replay from a captured CUDA graph: every call carries host-side counters in its kernels' arguments (the
 * lattice kernel's hand-over epoch, the alternating fill counters, the peer exchange's epoch), so a replayed launch
 * would meet the words its first run left behind.  Launches are already off the critical path (dependent launches). */
RNNTStatus mrnnt_enqueue(mrnnt_handle_t h, int blank_label, void *stream, float *gradients);
const float *mrnnt_device_costs(mrnnt_handle_t h);

/* The same call in two halves, for frameworks that compute gradients in their backward pass:
 *   forward : K1 + K2 (costs on the device, per-row gradient coefficients in the workspace when want_grads != 0)
 *   backward: K3 writes d(sum_b scale[b] * cost_b)/d(acts) into `gradients`; scale_dev_or_null = NULL means 1.
 * Replaces the reference's autograd glue around cost_and_grad (pytorch_binding/monotonic_rnnt_op.py:61-118: a
 * zeros_like memset before the call and a repeat_interleave * grads pass in backward, 3 more passes over
 * the logits-sized arrays).  acts and the workspace must stay untouched between the two halves. */
RNNTStatus mrnnt_enqueue_forward(mrnnt_handle_t h, int blank_label, void *stream, int want_grads);
RNNTStatus mrnnt_enqueue_backward(mrnnt_handle_t h, void *stream, float *gradients, const float *scale_dev_or_null);
/* Forward half for a caller that already owns the gradient buffer the backward half will fill (want_grads implied):
 * the lattice kernel writes that buffer's zero rows while its recursions run, and mrnnt_enqueue_backward, given the
 * same pointer, only writes the others.  The buffer must not be written in between. */
RNNTStatus mrnnt_enqueue_forward_into(mrnnt_handle_t h, int blank_label, void *stream, float *gradients);

/* ---- multi-GPU: the all-GPU sum of the summed cost, the path's only collective ------------------------
 * Done by the kernels themselves over peer memory (NVLink / NVSwitch) instead of by a collective library's kernel
 * behind them (include/mrnnt_b200/peer_reduce.cuh): every rank owns a "board" of a few bytes in device memory, mapped
 * into all peers with CUDA IPC.  One process per GPU:
 *   1. mrnnt_peer_board_create(world, &own, handle)       own board (zeroed) + its 64-byte IPC handle
 *   2. exchange the handles (any host-side transport), mrnnt_peer_board_open(handle_r, &boards[r]) for r != rank,
 *      boards[rank] = own; a host barrier, so that nobody publishes into a board that is not there yet
 *   3. mrnnt_set_peer_reduce(h, rank, world, boards, total_out, epoch) on every handle that takes part
 * From then on every mrnnt_cost_and_grad / mrnnt_enqueue / mrnnt_enqueue_forward[_into] of that handle also leaves
 * sum_over_ranks(sum_b cost_b) in *total_out (device or host-mapped pinned memory; NULL: nowhere): with gradients the
 * exchange rides inside the gradient kernel (one store per peer when the costs are final, one poll at the kernel's
 * end), else in a one-warp launch behind the lattice kernel.  All ranks must make the same sequence of such calls
 * (it is a collective).  The epoch counts the exchanges a set of boards has carried: pass 0 for fresh boards, and
 * when a NEW handle takes over boards already in use, mrnnt_peer_epoch() of the handle that used them last.
 * A rank whose peers do not show up within the time-out (mrnnt_set_peer_timeout_ms) gets NaN instead of hanging the
 * GPU, and an error status from then on.  world <= 8 (one NVSwitch
 * domain); in one process (several handles on one device or on peer-enabled devices) plain device pointers do. */
RNNTStatus mrnnt_peer_board_create(int world, void **board_dev, unsigned char ipc_handle[64]);
RNNTStatus mrnnt_peer_board_open(const unsigned char ipc_handle[64], void **peer_ptr);
RNNTStatus mrnnt_peer_board_close(void *peer_ptr);
RNNTStatus mrnnt_peer_board_destroy(void *board_dev);
RNNTStatus mrnnt_set_peer_reduce(mrnnt_handle_t h, int rank, int world, void *const *boards, float *total_out,
                                 unsigned epoch);
unsigned mrnnt_peer_epoch(mrnnt_handle_t h);
/* How long a rank waits for its slowest peer (default 60 s; 0 = without limit).  A rank that gives up gets NaN in
 * *total_out, its board is marked failed for good -- it publishes nothing any more, so that its peers run into their own
 * time-outs instead of reading sums of another step -- and mrnnt_cost_and_grad returns RNNT_STATUS_EXECUTION_FAILED from
 * that call on (mrnnt_peer_failed).  Recovery: new boards (mrnnt_peer_board_create ...) and mrnnt_set_peer_reduce. */
RNNTStatus mrnnt_set_peer_timeout_ms(mrnnt_handle_t h, unsigned milliseconds);
int mrnnt_peer_failed(mrnnt_handle_t h);

/* One-shot convenience: size check + set_workspace + optional restrict_to_alignment + cost_and_grad. */
RNNTStatus rnnt_loss_grad_gpu(const float *acts, const int *labels, const int *T_dev, const int *S_dev,
                              const int *T_host, const int *S_host, int B, int V, int blank_label,
                              const int *alignments_or_null, int max_shift, void *workspace, size_t workspace_bytes,
                              void *stream, float *costs_host, float *gradients_or_null);

/* ---- test / bench support (not part of the reference's surface) ---------------------------------*/
enum {
    MRNNT_OPT_FORCE_GENERIC = 1, /* value != 0: use the generic (non-TMA) streaming kernels */
    MRNNT_OPT_TIMING = 2,        /* value != 0: record CUDA events around K1 / K2 / K3 of every call */
    MRNNT_OPT_K1_WARPS = 3,      /* consumer warps per CTA of K1 (8, 16, 24)                            */
    MRNNT_OPT_K3_WARPS = 4,      /* consumer warps per CTA of K3 (8, 16, 24)                            */
    MRNNT_OPT_K2_PARTS = 5,      /* upper limit of CTAs per utterance in K2's coefficient phase (0: auto) */
    MRNNT_OPT_RESERVED_SMS = 6,  /* SMs the gradient kernel leaves free for a concurrent collective (0)    */
    MRNNT_OPT_PDL = 7,           /* programmatic dependent launch of K2 behind K1 and K3 behind K2 (1)     */
    MRNNT_OPT_K1_COMPACT = 8,    /* K1 variant that gives dead tiles no ring slot: 1 / 0 forced, -1 automatic */
    MRNNT_OPT_K2_ZERO_FILL = 9,  /* who zeroes the gradient's dead rows: 0 the gradient kernel's consumer warps; 1..4
                                    that many warps per lattice CTA, while the recursions run; 32 one more warp of
                                    the gradient kernel, next to its consumers; -1 automatic                    */
    MRNNT_OPT_DYNAMIC_TILES = 10,/* the gradient kernel hands its tiles out through a counter instead of round-robin by
                                    CTA index: 1 / 0 forced, -1 automatic; 2..100: through the counter once that
                                    percentage of a CTA's round-robin share has been worked off                  */
    MRNNT_OPT_LAUNCH_COUNT = 11, /* mrnnt_get_option only: kernel launches this handle has made so far (low 31 bits) */
    MRNNT_OPT_UPLOAD_COPY_ENGINE = 12,/* mrnnt_upload_acts: the all-live block in the middle of an utterance (packed layout,
                                    no alignment band) goes through the copy engine next to the upload kernel when it has
                                    at least this many bytes; 0: the kernel brings every live row; -1: default (1 MiB)  */
    MRNNT_OPT_K2_FILL_SHARE = 14,/* percent (1..100) of the zero fill (units of 32 rows from the front of the batch) the
                                    lattice kernel writes; the gradient kernel's zero-fill warp writes the rest; -1
                                    automatic (100: measured, the split loses on every named shape).  mrnnt_get_option:
                                    what the last call used                                                          */
    MRNNT_OPT_FUSED_PLAN = 15,   /* 1 (default): row starts, alignment band and row flags are set up by ONE kernel launch (up
                                    to 1024 utterances); 0: by the three kernels it replaces.  Before the first call.    */
    MRNNT_OPT_RETURN_EARLY = 13  /* 1 (default): mrnnt_cost_and_grad / compute_rnnt_loss with gradients return as soon as
                                    the costs are on the host; the gradient kernel may still be running and the gradients
                                    (and *total_out of a peer reduce) are complete in STREAM ORDER, like the result of any
                                    kernel launch -- both of the reference's bindings consume them on the same stream.
                                    A handle with a peer reduce still waits for everything (the world's sum and the
                                    exchange's verdict are promised on return) unless the value is 2: *total_out is then
                                    valid in stream order and an exchange that gave up is reported by the NEXT call.
                                    0: return only when everything the call launched has completed (the reference's
                                    blocking copy on the legacy stream, gpu_rnnt.h:229).  Also readable.             */
};
RNNTStatus mrnnt_set_option(mrnnt_handle_t h, int option, int value);
/* What the last call actually did: MRNNT_OPT_K2_ZERO_FILL -> 0, 1..4 or 32 as above; MRNNT_OPT_LAUNCH_COUNT.  Other
   options: RNNT_STATUS_INVALID_VALUE. */
RNNTStatus mrnnt_get_option(mrnnt_handle_t h, int option, int *value);
/* Durations in ms of K1, K2, K3 of the last call (MRNNT_OPT_TIMING on, stream synchronised). */
RNNTStatus mrnnt_last_timings(mrnnt_handle_t h, float ms_k1_k2_k3[3]);

enum {
    MRNNT_DBG_DENOM = 1,    /* double [rows] (meaningful for live rows)         */
    MRNNT_DBG_ALPHA = 2,    /* double [rows] (log alpha), -inf outside the band */
    MRNNT_DBG_BETA = 3,     /* double [rows] (log beta)                      */
    MRNNT_DBG_LP = 4,       /* double [rows][2] (log p blank, log p label)   */
    MRNNT_DBG_BAND = 5,     /* int32  [B][T_max][2] (min, max allowed s)     */
    MRNNT_DBG_ROWMETA = 6,  /* int32  [rows]                                 */
    MRNNT_DBG_LL = 7,       /* double [2][B] (ll_forward, ll_backward)       */
    MRNNT_DBG_ROWSTART = 8  /* int64  [B+1]                                  */
};
/* Blocking copy of an intermediate array to the host (device-synchronises first). */
RNNTStatus mrnnt_debug_copy(mrnnt_handle_t h, int what, void *dst_host, size_t dst_bytes);

/* Counter-based synthetic logits: x[i] = (splitmix64(seed ^ (index_offset + i)) >> 40) * 2^-24 in [0,1),
 * the distribution of the reference's genActs (tests/random.cpp:4-20).  Device-side fill. */
RNNTStatus mrnnt_synth_uniform(float *dst_dev, int64_t n, uint64_t seed, int64_t index_offset, void *stream);

/* Version / build information string (static storage). */
const char *mrnnt_build_info(void);

#ifdef __cplusplus
}
#endif

#endif /* MONOTONIC_RNNT_B200_C_API_H */
