// libmonotonic_rnnt.so -- the one translation unit of the library.
//
// Exports (extern "C"):
//   compute_rnnt_loss           the reference's C entry point (include/rnnt_entrypoint.h:24-25,
//                               src/rnnt_entrypoint.cpp:16-48), GPU only
//   mrnnt_* / rnnt_loss_grad_gpu  the flat C ABI declared in include/mrnnt_c_api.h
// All of them are thin shells over mrnnt::Engine (include/mrnnt_b200/engine.cuh), which is header-only
// so that the reference's framework bindings can compile the same kernels from `-I include` alone.
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "gpu_rnnt.h"
#include "gpu_workspace_manager.h"
#include "mrnnt_c_api.h"
#include "rnnt_entrypoint.h"

struct mrnnt_handle_st {
    GpuRNNTWorkspaceManager<float> manager;
    mrnnt_handle_st(const float *acts, const int *labels, int B, const int *T, const int *S, int V)
        : manager(acts, labels, B, T, S, V) {}
};

namespace {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) synth_uniform_kernel(float *__restrict__ dst, int64_t n, uint64_t seed,
                                                             int64_t offset) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t u = splitmix64(seed ^ static_cast<uint64_t>(offset + i));
        dst[i] = static_cast<float>(u >> 40) * (1.0f / 16777216.0f);
    }
}

}  // namespace

extern "C" {

RNNTStatus compute_rnnt_loss(RNNTWorkspaceManager &workspace_manager, RNNTOptions options, float *costs,
                             float *gradients) {
    if (costs == nullptr) return RNNT_STATUS_INVALID_VALUE;
    if (options.loc == RNNT_CPU) {
        std::fprintf(stderr, "monotonic-rnnt_b200: CPU execution requested, but this library is GPU-only\n");
        return RNNT_STATUS_EXECUTION_FAILED;
    }
    if (options.loc != RNNT_GPU) return RNNT_STATUS_INVALID_VALUE;
    auto *manager = dynamic_cast<GpuRNNTWorkspaceManager<float> *>(&workspace_manager);
    if (manager == nullptr) return RNNT_STATUS_INVALID_VALUE;  // the reference would throw std::bad_cast here
    GpuRNNTComputer<float> computer(*manager, options.blank_label, options.stream);
    return gradients != nullptr ? computer.cost_and_grad(costs, gradients) : computer.cost(costs);
}

RNNTStatus mrnnt_get_workspace_size(const int *T_host, const int *S_host, int B, int V, size_t *size_bytes) {
    if (size_bytes == nullptr) return RNNT_STATUS_INVALID_VALUE;
    mrnnt::Shape sh;
    const RNNTStatus st = mrnnt::validate_lengths(T_host, S_host, B, V, &sh);
    if (st != RNNT_STATUS_SUCCESS) return st;
    *size_bytes = mrnnt::workspace_bytes(sh);
    return RNNT_STATUS_SUCCESS;
}

RNNTStatus get_workspace_size(const int *T_host, const int *S_host, int B, int V, size_t *size_bytes) {
    return mrnnt_get_workspace_size(T_host, S_host, B, V, size_bytes);
}

RNNTStatus mrnnt_create(mrnnt_handle_t *out, const float *acts, const int *labels, int B, const int *T_dev,
                        const int *S_dev, int V, const int *T_host, const int *S_host) {
    if (out == nullptr) return RNNT_STATUS_INVALID_VALUE;
    *out = nullptr;
    if (B <= 0 || V <= 0) return RNNT_STATUS_INVALID_VALUE;
    auto *h = new (std::nothrow) mrnnt_handle_st(acts, labels, B, T_dev, S_dev, V);
    if (h == nullptr) return RNNT_STATUS_UNKNOWN_ERROR;
    if (T_host != nullptr && S_host != nullptr) {
        const RNNTStatus st = h->manager.set_host_lengths(T_host, S_host);
        if (st != RNNT_STATUS_SUCCESS) {
            delete h;
            return st;
        }
    }
    *out = h;
    return RNNT_STATUS_SUCCESS;
}

RNNTStatus mrnnt_get_workspace_size_padded(const int *T_host, const int *S_host, int B, int V, int T_dim, int U,
                                           int label_stride, size_t *size_bytes) {
    if (size_bytes == nullptr || T_dim <= 0 || U <= 0 || label_stride < 0) return RNNT_STATUS_INVALID_VALUE;
    mrnnt::Shape sh;
    const RNNTStatus st = mrnnt::validate_lengths(T_host, S_host, B, V, &sh, T_dim, U, label_stride);
    if (st != RNNT_STATUS_SUCCESS) return st;
    *size_bytes = mrnnt::workspace_bytes(sh);
    return RNNT_STATUS_SUCCESS;
}

RNNTStatus mrnnt_create_padded(mrnnt_handle_t *out, const float *acts, const int *labels, int B, const int *T_dev,
                               const int *S_dev, int V, int T_dim, int U, int label_stride, const int *T_host,
                               const int *S_host) {
    if (out == nullptr) return RNNT_STATUS_INVALID_VALUE;
    *out = nullptr;
    if (B <= 0 || V <= 0 || T_dim <= 0 || U <= 0 || label_stride < 0) return RNNT_STATUS_INVALID_VALUE;
    auto *h = new (std::nothrow) mrnnt_handle_st(acts, labels, B, T_dev, S_dev, V);
    if (h == nullptr) return RNNT_STATUS_UNKNOWN_ERROR;
    h->manager.set_padded_layout(T_dim, U, label_stride);
    if (T_host != nullptr && S_host != nullptr) {
        const RNNTStatus st = h->manager.set_host_lengths(T_host, S_host);
        if (st != RNNT_STATUS_SUCCESS) {
            delete h;
            return st;
        }
    }
    *out = h;
    return RNNT_STATUS_SUCCESS;
}

RNNTStatus mrnnt_set_dtype(mrnnt_handle_t h, int dtype) {
    if (h == nullptr || (dtype != MRNNT_DTYPE_F32 && dtype != MRNNT_DTYPE_BF16)) return RNNT_STATUS_INVALID_VALUE;
    h->manager.engine().set_bf16(dtype == MRNNT_DTYPE_BF16);
    return RNNT_STATUS_SUCCESS;
}

void mrnnt_destroy(mrnnt_handle_t h) { delete h; }

RNNTStatus mrnnt_workspace_size(mrnnt_handle_t h, size_t *size_bytes) {
    if (h == nullptr || size_bytes == nullptr) return RNNT_STATUS_INVALID_VALUE;
    return h->manager.get_workspace_size(size_bytes);
}

RNNTStatus mrnnt_set_workspace(mrnnt_handle_t h, void *workspace) {
    if (h == nullptr) return RNNT_STATUS_INVALID_VALUE;
    return h->manager.engine().set_workspace(workspace);
}

RNNTStatus mrnnt_create_workspace(mrnnt_handle_t h) {
    if (h == nullptr) return RNNT_STATUS_INVALID_VALUE;
    return h->manager.create_workspace();
}

void mrnnt_free_workspace(mrnnt_handle_t h) {
    if (h != nullptr) h->manager.free_workspace();
}

RNNTStatus mrnnt_restrict_to_alignment(mrnnt_handle_t h, const int *alignments, int max_shift, int blank_idx) {
    if (h == nullptr || alignments == nullptr) return RNNT_STATUS_INVALID_VALUE;
    h->manager.restrict_to_alignment(alignments, max_shift, blank_idx);
    return RNNT_STATUS_SUCCESS;
}

RNNTStatus mrnnt_restrict_to_alignment_strided(mrnnt_handle_t h, const int *alignments, int stride, int max_shift,
                                               int blank_idx) {
    if (h == nullptr || alignments == nullptr || stride <= 0) return RNNT_STATUS_INVALID_VALUE;
    h->manager.engine().restrict_to_alignment(alignments, max_shift, blank_idx, stride);
    return RNNT_STATUS_SUCCESS;
}

RNNTStatus mrnnt_get_shape(mrnnt_handle_t h, int *T_max, int *S_max, int64_t *rows) {
    if (h == nullptr) return RNNT_STATUS_INVALID_VALUE;
    mrnnt::Engine &e = h->manager.engine();
    const RNNTStatus st = e.ensure_shape();
    if (st != RNNT_STATUS_SUCCESS) return st;
    if (T_max != nullptr) *T_max = e.shape().T_max;
    if (S_max != nullptr) *S_max = e.shape().S_max;
    if (rows != nullptr) *rows = e.shape().rows;
    return RNNT_STATUS_SUCCESS;
}

void mrnnt_set_workspace_cache_limit(size_t bytes) { mrnnt::workspace_cache_set_limit(bytes); }
void mrnnt_trim_workspace_cache(void) { mrnnt::workspace_cache_trim(); }

RNNTStatus mrnnt_upload_acts(mrnnt_handle_t h, const void *host_acts, void *stream) {
    if (h == nullptr) return RNNT_STATUS_INVALID_VALUE;
    return h->manager.engine().upload_live_rows(host_acts, static_cast<cudaStream_t>(stream));
}

RNNTStatus mrnnt_cost_and_grad(mrnnt_handle_t h, int blank_label, void *stream, float *costs_host, float *gradients) {
    if (h == nullptr) return RNNT_STATUS_INVALID_VALUE;
    return h->manager.engine().compute(blank_label, static_cast<cudaStream_t>(stream), costs_host, gradients);
}

RNNTStatus mrnnt_enqueue(mrnnt_handle_t h, int blank_label, void *stream, float *gradients) {
    if (h == nullptr) return RNNT_STATUS_INVALID_VALUE;
    mrnnt::Engine &e = h->manager.engine();
    const RNNTStatus st = e.ensure_shape();
    if (st != RNNT_STATUS_SUCCESS) return st;
    if (!e.has_workspace() || blank_label < 0 || blank_label >= e.shape().V) return RNNT_STATUS_INVALID_VALUE;
    return e.enqueue(blank_label, static_cast<cudaStream_t>(stream), gradients);
}

RNNTStatus mrnnt_enqueue_forward(mrnnt_handle_t h, int blank_label, void *stream, int want_grads) {
    if (h == nullptr) return RNNT_STATUS_INVALID_VALUE;
    mrnnt::Engine &e = h->manager.engine();
    const RNNTStatus st = e.ensure_shape();
    if (st != RNNT_STATUS_SUCCESS) return st;
    if (!e.has_workspace() || blank_label < 0 || blank_label >= e.shape().V) return RNNT_STATUS_INVALID_VALUE;
    return e.enqueue_forward(blank_label, static_cast<cudaStream_t>(stream), want_grads != 0);
}

RNNTStatus mrnnt_enqueue_forward_into(mrnnt_handle_t h, int blank_label, void *stream, float *gradients) {
    if (h == nullptr || gradients == nullptr) return RNNT_STATUS_INVALID_VALUE;
    mrnnt::Engine &e = h->manager.engine();
    const RNNTStatus st = e.ensure_shape();
    if (st != RNNT_STATUS_SUCCESS) return st;
    if (!e.has_workspace() || blank_label < 0 || blank_label >= e.shape().V) return RNNT_STATUS_INVALID_VALUE;
    return e.enqueue_forward(blank_label, static_cast<cudaStream_t>(stream), true, gradients);
}

RNNTStatus mrnnt_enqueue_backward(mrnnt_handle_t h, void *stream, float *gradients, const float *scale_dev_or_null) {
    if (h == nullptr || gradients == nullptr) return RNNT_STATUS_INVALID_VALUE;
    mrnnt::Engine &e = h->manager.engine();
    if (!e.has_workspace()) return RNNT_STATUS_INVALID_VALUE;
    return e.enqueue_backward(static_cast<cudaStream_t>(stream), gradients, scale_dev_or_null);
}

const float *mrnnt_device_costs(mrnnt_handle_t h) {
    return (h != nullptr && h->manager.engine().has_workspace()) ? h->manager.engine().workspace().costs : nullptr;
}

RNNTStatus rnnt_loss_grad_gpu(const float *acts, const int *labels, const int *T_dev, const int *S_dev,
                              const int *T_host, const int *S_host, int B, int V, int blank_label,
                              const int *alignments_or_null, int max_shift, void *workspace, size_t workspace_bytes,
                              void *stream, float *costs_host, float *gradients_or_null) {
    if (costs_host == nullptr || workspace == nullptr || B <= 0 || V <= 0) return RNNT_STATUS_INVALID_VALUE;
    GpuRNNTWorkspaceManager<float> manager(acts, labels, B, T_dev, S_dev, V);
    if (T_host != nullptr && S_host != nullptr) {
        const RNNTStatus st = manager.set_host_lengths(T_host, S_host);
        if (st != RNNT_STATUS_SUCCESS) return st;
    }
    size_t need = 0;
    RNNTStatus st = manager.get_workspace_size(&need);
    if (st != RNNT_STATUS_SUCCESS) return st;
    if (workspace_bytes < need) return RNNT_STATUS_INVALID_VALUE;
    st = manager.engine().set_workspace(workspace);
    if (st != RNNT_STATUS_SUCCESS) return st;
    if (alignments_or_null != nullptr) manager.restrict_to_alignment(alignments_or_null, max_shift, blank_label);
    return manager.engine().compute(blank_label, static_cast<cudaStream_t>(stream), costs_host, gradients_or_null);
}

RNNTStatus mrnnt_set_option(mrnnt_handle_t h, int option, int value) {
    if (h == nullptr) return RNNT_STATUS_INVALID_VALUE;
    switch (option) {
        case MRNNT_OPT_FORCE_GENERIC:
            h->manager.engine().set_force_generic(value != 0);
            return RNNT_STATUS_SUCCESS;
        case MRNNT_OPT_TIMING:
            return h->manager.engine().set_timing(value != 0);
        case MRNNT_OPT_K1_WARPS:
            h->manager.engine().set_stream_warps(value, 0);
            return RNNT_STATUS_SUCCESS;
        case MRNNT_OPT_K3_WARPS:
            h->manager.engine().set_stream_warps(0, value);
            return RNNT_STATUS_SUCCESS;
        case MRNNT_OPT_K1_COMPACT:
            h->manager.engine().set_k1_compact(value);
            return RNNT_STATUS_SUCCESS;
        case MRNNT_OPT_PDL:
            h->manager.engine().set_pdl(value != 0);
            return RNNT_STATUS_SUCCESS;
        case MRNNT_OPT_RESERVED_SMS:
            h->manager.engine().set_reserved_sms(value);
            return RNNT_STATUS_SUCCESS;
        case MRNNT_OPT_K2_PARTS:
            h->manager.engine().set_k2_parts(value);
            return RNNT_STATUS_SUCCESS;
        case MRNNT_OPT_K2_ZERO_FILL:
            h->manager.engine().set_k2_zero_fill(value);
            return RNNT_STATUS_SUCCESS;
        case MRNNT_OPT_DYNAMIC_TILES:
            h->manager.engine().set_dynamic_tiles(value);
            return RNNT_STATUS_SUCCESS;
        case MRNNT_OPT_UPLOAD_COPY_ENGINE:
            h->manager.engine().set_upload_copy_engine(value);
            return RNNT_STATUS_SUCCESS;
        case MRNNT_OPT_RETURN_EARLY:
            h->manager.engine().set_return_early(value);
            return RNNT_STATUS_SUCCESS;
        case MRNNT_OPT_K2_FILL_SHARE:
            h->manager.engine().set_k2_fill_share(value);
            return RNNT_STATUS_SUCCESS;
        case MRNNT_OPT_FUSED_PLAN:
            h->manager.engine().set_fused_plan(value != 0);
            return RNNT_STATUS_SUCCESS;
        default:
            return RNNT_STATUS_INVALID_VALUE;
    }
}

RNNTStatus mrnnt_get_option(mrnnt_handle_t h, int option, int *value) {
    if (h == nullptr || value == nullptr) return RNNT_STATUS_INVALID_VALUE;
    switch (option) {
        case MRNNT_OPT_K2_ZERO_FILL:
            *value = h->manager.engine().last_k2_zero_warps();
            return RNNT_STATUS_SUCCESS;
        case MRNNT_OPT_LAUNCH_COUNT:
            *value = static_cast<int>(h->manager.engine().launch_count() & 0x7fffffffull);
            return RNNT_STATUS_SUCCESS;
        case MRNNT_OPT_RETURN_EARLY:
            *value = h->manager.engine().return_early();
            return RNNT_STATUS_SUCCESS;
        case MRNNT_OPT_K2_FILL_SHARE:
            *value = h->manager.engine().last_k2_fill_share();
            return RNNT_STATUS_SUCCESS;
        default:
            return RNNT_STATUS_INVALID_VALUE;
    }
}

RNNTStatus mrnnt_last_timings(mrnnt_handle_t h, float ms_k1_k2_k3[3]) {
    if (h == nullptr || ms_k1_k2_k3 == nullptr) return RNNT_STATUS_INVALID_VALUE;
    return h->manager.engine().last_timings(ms_k1_k2_k3);
}

RNNTStatus mrnnt_debug_copy(mrnnt_handle_t h, int what, void *dst_host, size_t dst_bytes) {
    if (h == nullptr || dst_host == nullptr) return RNNT_STATUS_INVALID_VALUE;
    mrnnt::Engine &e = h->manager.engine();
    if (!e.has_workspace()) return RNNT_STATUS_INVALID_VALUE;
    const mrnnt::Shape &sh = e.shape();
    const mrnnt::Workspace &w = e.workspace();
    const size_t rows = static_cast<size_t>(sh.rows), B = static_cast<size_t>(sh.B);
    const void *src = nullptr;
    size_t bytes = 0;
    switch (what) {
        case MRNNT_DBG_DENOM:
        case MRNNT_DBG_LP: {
            // K1 leaves (x_blank, x_label, denominator) per live row; handed out as denominators / log-probs
            const size_t per_row = what == MRNNT_DBG_LP ? 2 : 1;
            if (dst_bytes < rows * per_row * sizeof(double)) return RNNT_STATUS_INVALID_VALUE;
            std::vector<mrnnt::RawRow> tmp(rows);
            if (cudaDeviceSynchronize() != cudaSuccess) return RNNT_STATUS_EXECUTION_FAILED;
            if (cudaMemcpy(tmp.data(), w.lp, rows * sizeof(mrnnt::RawRow), cudaMemcpyDeviceToHost) != cudaSuccess)
                return RNNT_STATUS_MEMOPS_FAILED;
            double *d = static_cast<double *>(dst_host);
            for (size_t i = 0; i < rows; ++i) {
                const double den = (static_cast<double>(tmp[i].dh) + static_cast<double>(tmp[i].dl)) * mrnnt::kLn2D;
                if (what == MRNNT_DBG_DENOM) {
                    d[i] = den;
                } else {
                    d[2 * i] = static_cast<double>(tmp[i].xb) + den;
                    d[2 * i + 1] = static_cast<double>(tmp[i].xl) + den;
                }
            }
            return RNNT_STATUS_SUCCESS;
        }
        case MRNNT_DBG_ALPHA:
        case MRNNT_DBG_BETA: {
            // stored as (float mantissa, int exponent) cells; handed out as natural logs in double
            if (dst_bytes < rows * sizeof(double)) return RNNT_STATUS_INVALID_VALUE;
            std::vector<mrnnt::Cell> tmp(rows);
            if (cudaDeviceSynchronize() != cudaSuccess) return RNNT_STATUS_EXECUTION_FAILED;
            if (cudaMemcpy(tmp.data(), what == MRNNT_DBG_ALPHA ? w.alpha : w.beta, rows * sizeof(mrnnt::Cell),
                           cudaMemcpyDeviceToHost) != cudaSuccess)
                return RNNT_STATUS_MEMOPS_FAILED;
            double *d = static_cast<double *>(dst_host);
            for (size_t i = 0; i < rows; ++i) d[i] = mrnnt::cell_log(tmp[i].m, tmp[i].e);
            return RNNT_STATUS_SUCCESS;
        }
        case MRNNT_DBG_BAND: src = w.band; bytes = B * static_cast<size_t>(sh.T_max) * sizeof(int2); break;
        case MRNNT_DBG_ROWMETA: src = w.rowmeta; bytes = rows * sizeof(int); break;
        case MRNNT_DBG_ROWSTART: src = w.row_start; bytes = (B + 1) * sizeof(int64_t); break;
        case MRNNT_DBG_LL: {
            if (dst_bytes < 2 * B * sizeof(double)) return RNNT_STATUS_INVALID_VALUE;
            if (cudaDeviceSynchronize() != cudaSuccess) return RNNT_STATUS_EXECUTION_FAILED;
            char *d = static_cast<char *>(dst_host);
            if (cudaMemcpy(d, w.ll_fwd, B * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess ||
                cudaMemcpy(d + B * sizeof(double), w.ll_bwd, B * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
                return RNNT_STATUS_MEMOPS_FAILED;
            return RNNT_STATUS_SUCCESS;
        }
        default: return RNNT_STATUS_INVALID_VALUE;
    }
    if (dst_bytes < bytes) return RNNT_STATUS_INVALID_VALUE;
    if (cudaDeviceSynchronize() != cudaSuccess) return RNNT_STATUS_EXECUTION_FAILED;
    if (cudaMemcpy(dst_host, src, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return RNNT_STATUS_MEMOPS_FAILED;
    return RNNT_STATUS_SUCCESS;
}

/* ---- the all-GPU sum of the summed cost over peer memory (include/mrnnt_b200/peer_reduce.cuh) ---- */
RNNTStatus mrnnt_peer_board_create(int world, void **board_dev, unsigned char ipc_handle[64]) {
    if (board_dev == nullptr || world <= 0 || world > mrnnt::kPeerMaxWorld) return RNNT_STATUS_INVALID_VALUE;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the C ABI carries the IPC handle as 64 bytes");
    void *p = nullptr;
    // (an allocation of its own: an IPC handle names a whole cudaMalloc block)
    if (cudaMalloc(&p, mrnnt::peer_board_bytes(mrnnt::kPeerMaxWorld)) != cudaSuccess) return RNNT_STATUS_MEMOPS_FAILED;
    if (cudaMemset(p, 0, mrnnt::peer_board_bytes(mrnnt::kPeerMaxWorld)) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
        (void)cudaFree(p);
        return RNNT_STATUS_MEMOPS_FAILED;
    }
    if (ipc_handle != nullptr) {
        cudaIpcMemHandle_t hnd;
        if (cudaIpcGetMemHandle(&hnd, p) != cudaSuccess) {
            (void)cudaGetLastError();
            (void)cudaFree(p);
            return RNNT_STATUS_EXECUTION_FAILED;
        }
        std::memcpy(ipc_handle, &hnd, 64);
    }
    *board_dev = p;
    return RNNT_STATUS_SUCCESS;
}

RNNTStatus mrnnt_peer_board_open(const unsigned char ipc_handle[64], void **peer_ptr) {
    if (ipc_handle == nullptr || peer_ptr == nullptr) return RNNT_STATUS_INVALID_VALUE;
    cudaIpcMemHandle_t hnd;
    std::memcpy(&hnd, ipc_handle, 64);
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        (void)cudaGetLastError();
        return RNNT_STATUS_EXECUTION_FAILED;
    }
    *peer_ptr = p;
    return RNNT_STATUS_SUCCESS;
}

RNNTStatus mrnnt_peer_board_close(void *peer_ptr) {
    if (peer_ptr == nullptr) return RNNT_STATUS_SUCCESS;
    return cudaIpcCloseMemHandle(peer_ptr) == cudaSuccess ? RNNT_STATUS_SUCCESS : RNNT_STATUS_MEMOPS_FAILED;
}

RNNTStatus mrnnt_peer_board_destroy(void *board_dev) {
    if (board_dev == nullptr) return RNNT_STATUS_SUCCESS;
    return cudaFree(board_dev) == cudaSuccess ? RNNT_STATUS_SUCCESS : RNNT_STATUS_MEMOPS_FAILED;
}

RNNTStatus mrnnt_set_peer_reduce(mrnnt_handle_t h, int rank, int world, void *const *boards, float *total_out,
                                 unsigned epoch) {
    if (h == nullptr) return RNNT_STATUS_INVALID_VALUE;
    const RNNTStatus st = h->manager.engine().set_peer_reduce(rank, world, boards, total_out);
    if (st == RNNT_STATUS_SUCCESS) h->manager.engine().set_peer_epoch(epoch);
    return st;
}

unsigned mrnnt_peer_epoch(mrnnt_handle_t h) { return h == nullptr ? 0u : h->manager.engine().peer_epoch(); }

RNNTStatus mrnnt_set_peer_timeout_ms(mrnnt_handle_t h, unsigned milliseconds) {
    if (h == nullptr) return RNNT_STATUS_INVALID_VALUE;
    h->manager.engine().set_peer_timeout_ms(milliseconds);
    return RNNT_STATUS_SUCCESS;
}

int mrnnt_peer_failed(mrnnt_handle_t h) { return (h != nullptr && h->manager.engine().peer_failed()) ? 1 : 0; }

RNNTStatus mrnnt_synth_uniform(float *dst_dev, int64_t n, uint64_t seed, int64_t index_offset, void *stream) {
    if (dst_dev == nullptr || n < 0) return RNNT_STATUS_INVALID_VALUE;
    if (n == 0) return RNNT_STATUS_SUCCESS;
    const int64_t want = (n + 255) / 256;
    const int blocks = static_cast<int>(want < 148 * 16 ? want : 148 * 16);
    synth_uniform_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(dst_dev, n, seed, index_offset);
    return cudaGetLastError() == cudaSuccess ? RNNT_STATUS_SUCCESS : RNNT_STATUS_EXECUTION_FAILED;
}

const char *mrnnt_build_info(void) {
    return "monotonic-rnnt_b200 (sm_100a; K1 lse+gather / K2 lattice / K3 grad; no CPU fallback) built " __DATE__;
}

}  // extern "C"
