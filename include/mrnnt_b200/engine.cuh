// Host-side orchestration of the loss-and-gradient hot path: validation, workspace sizing, lazy
// device set-up, kernel selection and the K1 -> K2 -> K3 launch sequence.
//
// This is what GpuRNNTWorkspaceManager / GpuRNNTComputer (include/gpu_workspace_manager.h,
// include/gpu_rnnt.h) and the flat C ABI (include/mrnnt_c_api.h) are thin shells around.  Compared
// with the reference's GpuRNNTComputer::cost_and_grad (include/gpu_rnnt.h:27-234, ~21 blocking D2H
// copies, 8 H2D copies and 2 stream syncs per call) a call here is: 3 kernel launches,
// one async D2H of B floats and ONE stream synchronisation.  T[] and S[] are fetched to the host
// once per manager (they are needed for the workspace size, which the API returns on the host).
//
// There is no CPU path in this file or anywhere else in the product.
#pragma once

#include <cuda_runtime.h>

#include <atomic>

#include <cstdint>
#include <cstring>
#include <mutex>
#include <type_traits>
#include <vector>

#include "../status.h"
#include "k1_lse.cuh"
#include "k2_lattice.cuh"
#include "k3_grad.cuh"
#include "plan.cuh"

namespace mrnnt {


struct DeviceInfo {
    int sm_count = 0;
    int max_smem_optin = 0;
    bool ok = false;
};

inline const DeviceInfo &device_info() {
    // queried once per process per device is not needed: the hot path runs on one device per process
    static thread_local int cached_dev = -1;
    static thread_local DeviceInfo info;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        info.ok = false;
        return info;
    }
    if (dev != cached_dev) {
        info.ok = cudaDeviceGetAttribute(&info.sm_count, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
                  cudaDeviceGetAttribute(&info.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) ==
                      cudaSuccess;
        cached_dev = dev;
    }
    return info;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) costs host microseconds; do it once per (kernel, device).
// The table has INTERNAL linkage on purpose: every module that carries these headers (libmonotonic_rnnt.so, a
// framework extension compiled from the same headers) owns its own copies of the kernels, so the "already
// configured" state must not be shared between modules (a function-local static in a template would be: it
// is emitted as a process-wide unique symbol).
constexpr int kMaxDevices = 32;
struct SmemAttrEntry {
    const void *kernel = nullptr;
    size_t bytes[kMaxDevices] = {};
};
static SmemAttrEntry g_smem_attr[48];
static std::mutex g_smem_attr_mutex;

template <typename Kern>
inline bool ensure_dynamic_smem(Kern kern, size_t bytes) {
    if (bytes <= 48 * 1024) return true;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return false;
    const void *key = reinterpret_cast<const void *>(kern);
    std::lock_guard<std::mutex> lock(g_smem_attr_mutex);
    SmemAttrEntry *slot = nullptr;
    for (auto &e : g_smem_attr) {
        if (e.kernel == key || e.kernel == nullptr) {
            slot = &e;
            break;
        }
    }
    if (slot != nullptr && slot->kernel == key && bytes <= slot->bytes[dev]) return true;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)) != cudaSuccess)
        return false;
    if (slot != nullptr) {
        slot->kernel = key;
        slot->bytes[dev] = bytes;
    }
    return true;
}

// Host-mapped pinned staging buffers for the B costs of a synchronous call: the lattice kernel writes them straight
// into host memory (128 bytes over PCIe for c2), so the call ends with the stream synchronisation alone -- no
// device-to-host copy is queued behind the gradient kernel.  Allocation of pinned memory is slow, so the buffers are
// pooled per module and reused; a batch that does not fit (or a failed allocation) falls back to cudaMemcpyAsync.
constexpr size_t kCostStageFloats = 16384;
struct CostStage {
    float *host = nullptr;  // == device address (unified addressing, cudaHostAllocPortable | cudaHostAllocMapped)
};
static std::vector<CostStage> g_cost_stages;
static std::mutex g_cost_stage_mutex;

inline CostStage acquire_cost_stage(size_t floats) {
    CostStage st;
    if (floats > kCostStageFloats) return st;
    {
        std::lock_guard<std::mutex> lock(g_cost_stage_mutex);
        if (!g_cost_stages.empty()) {
            st = g_cost_stages.back();
            g_cost_stages.pop_back();
            return st;
        }
    }
    void *p = nullptr;
    if (cudaHostAlloc(&p, kCostStageFloats * sizeof(float), cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
        (void)cudaGetLastError();
        return st;
    }
    st.host = static_cast<float *>(p);
    return st;
}
inline void release_cost_stage(const CostStage &st) {
    if (st.host == nullptr) return;
    std::lock_guard<std::mutex> lock(g_cost_stage_mutex);
    g_cost_stages.push_back(st);
}

// A side stream for the copy-engine part of an upload (Engine::upload_live_rows): forked off the caller's stream and
// joined back into it with two events.  One per device and module; the enqueue sequence holds the mutex, so two handles
// on two threads cannot interleave their records of the shared events.
struct UploadLane {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    bool ok = false;
};
static std::mutex g_upload_lane_mutex;
inline UploadLane &upload_lane(int device) {  // (call with g_upload_lane_mutex held)
    static UploadLane lanes[64];
    UploadLane &l = lanes[device & 63];
    if (!l.ok && l.stream == nullptr) {
        l.ok = cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking) == cudaSuccess &&
               cudaEventCreateWithFlags(&l.fork, cudaEventDisableTiming) == cudaSuccess &&
               cudaEventCreateWithFlags(&l.join, cudaEventDisableTiming) == cudaSuccess;
        if (!l.ok) (void)cudaGetLastError();
    }
    return l;
}

// T[] and S[] in one hop: a tiny kernel writes both into a host-mapped staging buffer (one launch + one
// synchronisation of the legacy stream, ~8 us) instead of two blocking cudaMemcpy calls (~16 us).
static __global__ void gather_lengths_kernel(const int *__restrict__ T, const int *__restrict__ S, int B,
                                             int *__restrict__ out) {
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        out[b] = T[b];
        out[B + b] = S[b];
    }
}

// Device blocks handed back by free_workspace() are kept for the next create_workspace() of this module instead of
// going through cudaFree / cudaMalloc: the reference's torch binding creates and frees the workspace on every loss
// call (pytorch_binding/monotonic_rnnt.cu:99-111), and the driver's allocator costs milliseconds for blocks of this
// size (measured: 2.9 ms per call against 0.36 ms with a workspace that stays).  At most kWorkspaceCacheBlocks
// blocks and g_ws_cache_limit bytes per module are kept (workspace_cache_set_limit; 0 turns the cache off: every
// free_workspace() is then a cudaFree, as in the reference); the smallest sufficient block is reused;
// workspace_cache_trim() gives everything back to the driver.
constexpr int kWorkspaceCacheBlocks = 4;
constexpr size_t kWorkspaceCacheDefaultLimit = size_t(1) << 30;  // 1 GiB (c3's workspace is 91 MB, c4's 56 MB)
struct CachedBlock {
    void *ptr = nullptr;
    size_t bytes = 0;
    int device = -1;
    cudaEvent_t busy = nullptr;  // behind the last kernels that use the block (nullptr: nothing can still touch it); the
                                 // cache owns the event, whoever takes the block orders its first launch behind it
};
static CachedBlock g_ws_cache[kWorkspaceCacheBlocks];
static size_t g_ws_cache_limit = kWorkspaceCacheDefaultLimit;
static std::mutex g_ws_cache_mutex;

inline void *workspace_cache_take(size_t bytes, int device, size_t *got, cudaEvent_t *busy) {
    std::lock_guard<std::mutex> lock(g_ws_cache_mutex);
    int best = -1;
    for (int i = 0; i < kWorkspaceCacheBlocks; ++i) {
        const CachedBlock &c = g_ws_cache[i];
        if (c.ptr != nullptr && c.device == device && c.bytes >= bytes && (best < 0 || c.bytes < g_ws_cache[best].bytes))
            best = i;
    }
    if (best < 0) return nullptr;
    void *p = g_ws_cache[best].ptr;
    *got = g_ws_cache[best].bytes;
    *busy = g_ws_cache[best].busy;
    g_ws_cache[best] = CachedBlock();
    return p;
}
// Free `ptr` on its own device, whatever the current device is.
inline void workspace_free_on(void *ptr, int device) {
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) cur = -1;
    if (cur != device && device >= 0) (void)cudaSetDevice(device);
    (void)cudaFree(ptr);
    if (cur != device && cur >= 0) (void)cudaSetDevice(cur);
}
inline void workspace_drop(const CachedBlock &c) {
    workspace_free_on(c.ptr, c.device);  // (cudaFree waits for the device: whatever `busy` stands for is over then)
    if (c.busy != nullptr) (void)cudaEventDestroy(c.busy);
}
// Keeps the block, or frees it (cache off, block above the byte limit); blocks the cache gives up for it are freed too.
// busy: see CachedBlock (ownership passes to the cache).
inline void workspace_cache_put(void *ptr, size_t bytes, int device, cudaEvent_t busy = nullptr) {
    CachedBlock drop[kWorkspaceCacheBlocks + 1];
    int ndrop = 0;
    {
        std::lock_guard<std::mutex> lock(g_ws_cache_mutex);
        if (bytes > g_ws_cache_limit) {
            drop[ndrop++] = CachedBlock{ptr, bytes, device, busy};
        } else {
            int slot = -1;
            for (int i = 0; i < kWorkspaceCacheBlocks && slot < 0; ++i)
                if (g_ws_cache[i].ptr == nullptr) slot = i;
            if (slot < 0) {  // full: the smallest cached block goes
                slot = 0;
                for (int i = 1; i < kWorkspaceCacheBlocks; ++i)
                    if (g_ws_cache[i].bytes < g_ws_cache[slot].bytes) slot = i;
                drop[ndrop++] = g_ws_cache[slot];
            }
            g_ws_cache[slot] = CachedBlock{ptr, bytes, device, busy};
            // the byte limit: largest blocks first out
            for (;;) {
                size_t total = 0;
                int big = -1;
                for (int i = 0; i < kWorkspaceCacheBlocks; ++i) {
                    if (g_ws_cache[i].ptr == nullptr) continue;
                    total += g_ws_cache[i].bytes;
                    if (big < 0 || g_ws_cache[i].bytes > g_ws_cache[big].bytes) big = i;
                }
                if (total <= g_ws_cache_limit || big < 0) break;
                drop[ndrop++] = g_ws_cache[big];
                g_ws_cache[big] = CachedBlock();
            }
        }
    }
    for (int i = 0; i < ndrop; ++i) workspace_drop(drop[i]);
}
inline void workspace_cache_trim() {
    CachedBlock drop[kWorkspaceCacheBlocks];
    {
        std::lock_guard<std::mutex> lock(g_ws_cache_mutex);
        for (int i = 0; i < kWorkspaceCacheBlocks; ++i) {
            drop[i] = g_ws_cache[i];
            g_ws_cache[i] = CachedBlock();
        }
    }
    for (const CachedBlock &c : drop)
        if (c.ptr != nullptr) workspace_drop(c);
}
inline void workspace_cache_set_limit(size_t bytes) {
    {
        std::lock_guard<std::mutex> lock(g_ws_cache_mutex);
        g_ws_cache_limit = bytes;
    }
    if (bytes == 0) workspace_cache_trim();
}

// Kernel launch, optionally as a programmatic dependent of the previous kernel in the stream (common.cuh).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, bool pdl,
                                 Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(static_cast<unsigned>(block));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Validation rules of the reference (gpu_workspace_manager.h:232-239, cpu twin :99-107).
// Padded layout: T_dim / U (= label positions + 1) are the tensor's own dimensions and must cover every
// utterance; label_stride is the width of the labels array (0: the reference's rule, max_b S_b).
inline RNNTStatus validate_lengths(const int *T_host, const int *S_host, int B, int V, Shape *shape, int T_dim = 0,
                                   int U = 0, int label_stride = 0) {
    if (B <= 0 || V <= 0 || T_host == nullptr || S_host == nullptr) return RNNT_STATUS_INVALID_VALUE;
    Shape sh;
    sh.B = B;
    sh.V = V;
    for (int b = 0; b < B; ++b) {
        const int t = T_host[b], s = S_host[b];
        if (t <= 0 || s < 0 || t < s) return RNNT_STATUS_INVALID_VALUE;
        sh.T_max = t > sh.T_max ? t : sh.T_max;
        sh.S_max = s > sh.S_max ? s : sh.S_max;
        sh.rows += static_cast<int64_t>(t) * (s + 1);
    }
    sh.label_stride = label_stride > 0 ? label_stride : sh.S_max;
    if (sh.label_stride < sh.S_max) return RNNT_STATUS_INVALID_VALUE;
    if (T_dim > 0 || U > 0) {
        if (T_dim < sh.T_max || U < sh.S_max + 1) return RNNT_STATUS_INVALID_VALUE;
        if (static_cast<int64_t>(T_dim) * U > 0x7fffffff) return RNNT_STATUS_INVALID_VALUE;  // per-utterance int index
        sh.T_dim = T_dim;
        sh.U = U;
        sh.rows = static_cast<int64_t>(B) * T_dim * U;
    }
    *shape = sh;
    return RNNT_STATUS_SUCCESS;
}

class Engine {
   public:
    Engine(const void *acts, const int *labels, int B, const int *T_dev, const int *S_dev, int V)
        : acts_(acts), labels_(labels), T_dev_(T_dev), S_dev_(S_dev), B_(B), V_(V) {}

    Engine(const Engine &) = delete;
    Engine &operator=(const Engine &) = delete;

    // Extension: acts and gradients are bfloat16 instead of float32; arithmetic stays float.
    void set_bf16(bool on) { bf16_ = on; }
    bool bf16() const { return bf16_; }

    // acts (and gradients) are a padded [B, T_dim, U, V] tensor, labels [B, label_stride] (see Shape).  Must be
    // called before the first size query.
    void set_padded_layout(int T_dim, int U, int label_stride) {
        pad_T_ = T_dim;
        pad_U_ = U;
        label_stride_ = label_stride;
        have_shape_ = false;
    }

    // Supply host copies of the length arrays (skips the one D2H fetch).
    RNNTStatus set_host_lengths(const int *T_host, const int *S_host) {
        if (B_ <= 0) return RNNT_STATUS_INVALID_VALUE;
        T_h_.assign(T_host, T_host + B_);
        S_h_.assign(S_host, S_host + B_);
        shape_status_ = validate_lengths(T_h_.data(), S_h_.data(), B_, V_, &shape_, pad_T_, pad_U_, label_stride_);
        have_shape_ = true;
        return shape_status_;
    }

    // One blocking D2H of 2*B ints, the first time the shape is needed.
    RNNTStatus ensure_shape() {
        if (have_shape_) return shape_status_;
        if (B_ <= 0 || V_ <= 0) {
            have_shape_ = true;
            return shape_status_ = RNNT_STATUS_INVALID_VALUE;
        }
        T_h_.resize(B_);
        S_h_.resize(B_);
        bool fetched = false;
        const CostStage stage = acquire_cost_stage(2 * static_cast<size_t>(B_));
        if (stage.host != nullptr) {
            // on the legacy default stream, like the reference's blocking copies (gpu_workspace_manager.h:87-95)
            int *mapped = reinterpret_cast<int *>(stage.host);
            gather_lengths_kernel<<<(B_ + 255) / 256, 256>>>(T_dev_, S_dev_, B_, mapped);
            if (cudaGetLastError() == cudaSuccess && cudaStreamSynchronize(nullptr) == cudaSuccess) {
                std::memcpy(T_h_.data(), mapped, sizeof(int) * B_);
                std::memcpy(S_h_.data(), mapped + B_, sizeof(int) * B_);
                fetched = true;
            }
            release_cost_stage(stage);
        }
        if (!fetched && (cudaMemcpy(T_h_.data(), T_dev_, sizeof(int) * B_, cudaMemcpyDeviceToHost) != cudaSuccess ||
                         cudaMemcpy(S_h_.data(), S_dev_, sizeof(int) * B_, cudaMemcpyDeviceToHost) != cudaSuccess)) {
            have_shape_ = true;
            return shape_status_ = RNNT_STATUS_MEMOPS_FAILED;
        }
        shape_status_ = validate_lengths(T_h_.data(), S_h_.data(), B_, V_, &shape_, pad_T_, pad_U_, label_stride_);
        have_shape_ = true;
        return shape_status_;
    }

    RNNTStatus workspace_size(size_t *bytes) {
        const RNNTStatus st = ensure_shape();
        if (st != RNNT_STATUS_SUCCESS) return st;
        *bytes = workspace_bytes(shape_);
        return RNNT_STATUS_SUCCESS;
    }

    // Hand over a caller-owned device buffer of at least workspace_size() bytes.
    RNNTStatus set_workspace(void *workspace) {
        const RNNTStatus st = ensure_shape();
        if (st != RNNT_STATUS_SUCCESS) return st;
        if (workspace == nullptr) return RNNT_STATUS_INVALID_VALUE;
        base_ = workspace;
        ws_ = carve_workspace(workspace, shape_);
        plan_dirty_ = true;
        band_dirty_ = true;
        alignment_ = nullptr;
        coef_blank_ = -1;
        return RNNT_STATUS_SUCCESS;
    }

    RNNTStatus create_workspace() {
        size_t bytes = 0;
        const RNNTStatus st = workspace_size(&bytes);
        if (st != RNNT_STATUS_SUCCESS) return st;
        release_owned();  // (a second create_workspace() without free_workspace(): the first block is not leaked)
        int device = 0;
        if (cudaGetDevice(&device) != cudaSuccess) return RNNT_STATUS_EXECUTION_FAILED;
        size_t got = bytes;
        cudaEvent_t busy = nullptr;
        void *p = workspace_cache_take(bytes, device, &got, &busy);
        if (p == nullptr && cudaMalloc(&p, bytes) != cudaSuccess) return RNNT_STATUS_MEMOPS_FAILED;
        if (block_busy_ != nullptr) (void)cudaEventDestroy(block_busy_);
        block_busy_ = busy;  // (the block's last user may still be running: setup() orders our first launch behind it)
        owned_ = p;
        owned_bytes_ = got;
        owned_device_ = device;
        return set_workspace(p);
    }

    void free_workspace() {
        release_owned();
        base_ = nullptr;
    }

    // Restrict the lattice to a band around `alignments` ([B, T_max] device ints).  Takes effect at the
    // next compute() on that call's stream; `alignments` must stay valid until then.
    // `stride`: ints per utterance in `alignments`; 0 = the reference's rule, max_b T_b (cpu_workspace_manager.h:208).
    // A stride below max_b T_b is refused by the next compute call (RNNT_STATUS_INVALID_VALUE).
    void restrict_to_alignment(const int *alignments, int max_shift, int blank_idx, int stride = 0) {
        alignment_ = alignments;
        max_shift_ = max_shift;
        align_blank_ = blank_idx;
        align_stride_ = stride;
        band_dirty_ = true;
    }

    // Fill the device logits from PINNED host memory (cudaHostAlloc / cudaHostRegister; `host_acts` has the layout and
    // type of acts): only the rows the lattice reads cross the bus, by a kernel that loads them straight from host
    // memory (plan.cuh::upload_live_rows_kernel).  Dead rows of the device array keep whatever they held -- no kernel
    // of a later call on this handle reads them.  Needs the workspace; call it after restrict_to_alignment (the band
    // decides which rows are live; widening the band later needs a new upload).  Asynchronous on `stream`.
    RNNTStatus upload_live_rows(const void *host_acts, cudaStream_t stream) {
        if (host_acts == nullptr) return RNNT_STATUS_INVALID_VALUE;
        RNNTStatus st = ensure_shape();
        if (st != RNNT_STATUS_SUCCESS) return st;
        if (base_ == nullptr) return RNNT_STATUS_INVALID_VALUE;
        const DeviceInfo &dev = device_info();
        if (!dev.ok) return RNNT_STATUS_EXECUTION_FAILED;
        void *src = nullptr;  // (the device's view of the pinned block; fails for pageable memory)
        if (cudaHostGetDevicePointer(&src, const_cast<void *>(host_acts), 0) != cudaSuccess) {
            (void)cudaGetLastError();
            return RNNT_STATUS_INVALID_VALUE;
        }
        st = setup(stream);
        if (st != RNNT_STATUS_SUCCESS) return st;
        void *dst = const_cast<void *>(acts_);
        const size_t row_bytes = static_cast<size_t>(V_) * elem_bytes();
        const uintptr_t bits = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | row_bytes;
        const int grid = dev.sm_count * 8;
        // The all-live block in the middle of every utterance (plan.cuh: upload_middle_block; packed layout, no band)
        // goes through the copy engine on a side stream WHILE the kernel brings the ragged frames around it: the copy
        // engine moves 55.6 GB/s over PCIe against the 51.5 of SM-issued reads (tools/h2d_probe.cu), and the tail of the
        // upload -- the block is most of an utterance -- runs at its rate.
        int64_t mid_min_rows = 0;
        std::unique_lock<std::mutex> lane_lock(g_upload_lane_mutex, std::defer_lock);
        UploadLane *lane = nullptr;
        if (upload_copy_min_bytes_ > 0 && alignment_ == nullptr && pad_T_ == 0 && static_cast<int>(T_h_.size()) == B_) {
            int cur = 0;
            if (cudaGetDevice(&cur) == cudaSuccess) {
                lane_lock.lock();
                lane = &upload_lane(cur);
                if (!lane->ok) {
                    lane = nullptr;
                    lane_lock.unlock();
                }
            }
        }
        if (lane != nullptr) {
            mid_min_rows = static_cast<int64_t>((upload_copy_min_bytes_ + row_bytes - 1) / row_bytes);
            if (cudaEventRecord(lane->fork, stream) != cudaSuccess || cudaStreamWaitEvent(lane->stream, lane->fork, 0) != cudaSuccess)
                return RNNT_STATUS_EXECUTION_FAILED;
            int64_t R = 0;
            bool any = false;
            for (int b = 0; b < B_; ++b) {
                const MiddleBlock m = upload_middle_block(T_h_[b], S_h_[b], mid_min_rows);
                if (m.rows > 0) {
                    const size_t off = static_cast<size_t>(R + m.first) * row_bytes;
                    if (cudaMemcpyAsync(static_cast<unsigned char *>(dst) + off, static_cast<const unsigned char *>(host_acts) + off,
                                        static_cast<size_t>(m.rows) * row_bytes, cudaMemcpyHostToDevice, lane->stream) != cudaSuccess)
                        return RNNT_STATUS_MEMOPS_FAILED;
                    any = true;
                }
                R += static_cast<int64_t>(T_h_[b]) * (S_h_[b] + 1);
            }
            if (!any) mid_min_rows = 0;
            if (cudaEventRecord(lane->join, lane->stream) != cudaSuccess) return RNNT_STATUS_EXECUTION_FAILED;
        }
        if ((bits & 15) == 0) {
            upload_live_rows_kernel<uint4><<<grid, kUploadThreads, 0, stream>>>(
                static_cast<const uint4 *>(src), static_cast<uint4 *>(dst), ws_.rowmeta, shape_.rows, static_cast<int>(row_bytes / 16),
                mid_min_rows, ws_.rowutt, ws_.row_start, T_dev_, S_dev_);
        } else if ((bits & 3) == 0) {
            upload_live_rows_kernel<uint32_t><<<grid, kUploadThreads, 0, stream>>>(
                static_cast<const uint32_t *>(src), static_cast<uint32_t *>(dst), ws_.rowmeta, shape_.rows, static_cast<int>(row_bytes / 4),
                mid_min_rows, ws_.rowutt, ws_.row_start, T_dev_, S_dev_);
        } else {
            upload_live_rows_kernel<uint16_t><<<grid, kUploadThreads, 0, stream>>>(
                static_cast<const uint16_t *>(src), static_cast<uint16_t *>(dst), ws_.rowmeta, shape_.rows, static_cast<int>(row_bytes / 2),
                mid_min_rows, ws_.rowutt, ws_.row_start, T_dev_, S_dev_);
        }
        const RNNTStatus kst = launched();
        if (lane != nullptr && cudaStreamWaitEvent(stream, lane->join, 0) != cudaSuccess) return RNNT_STATUS_EXECUTION_FAILED;
        return kst;
    }
    // The smallest all-live middle block (bytes) that goes through the copy engine next to the upload kernel (default
    // 1 MiB); 0: the kernel brings every live row.
    void set_upload_copy_engine(int min_bytes) { upload_copy_min_bytes_ = min_bytes < 0 ? kUploadMinCopyBytes : static_cast<size_t>(min_bytes); }

    const Shape &shape() const { return shape_; }
    const Workspace &workspace() const { return ws_; }
    bool has_workspace() const { return base_ != nullptr; }
    int B() const { return B_; }

    // costs_host: B floats on the host (valid on return).  grads_dev: packed like acts, or nullptr.
    // With gradients the call returns as soon as the costs have reached the host -- the first warp of the gradient
    // kernel sends them and then a sequence word, which this thread polls -- while the gradient kernel is still
    // running: the gradients are complete in STREAM ORDER, like the output of any kernel launch (both of the
    // reference's bindings consume them on the same stream), and the caller's next launches queue up behind the
    // gradient kernel instead of behind a host round trip.  set_return_early(false): wait for the whole stream, as the
    // reference's blocking copy at gpu_rnnt.h:229 does on the legacy stream.
    RNNTStatus compute(int blank, cudaStream_t stream, float *costs_host, void *grads_dev) {
        if (costs_host == nullptr) return RNNT_STATUS_INVALID_VALUE;
        RNNTStatus st = ensure_shape();
        if (st != RNNT_STATUS_SUCCESS) return st;
        if (base_ == nullptr || blank < 0 || blank >= V_) return RNNT_STATUS_INVALID_VALUE;
        if (peer_failed_) return RNNT_STATUS_EXECUTION_FAILED;  // an exchange gave up earlier: final (peer_reduce.cuh)
        // The staging buffer stays with the handle: after an early return the gradient kernel may still write the
        // exchange's "gave up" flag into it (one word behind the costs; behind that the sequence word).
        if (!stage_tried_) {
            stage_tried_ = true;
            stage_ = acquire_cost_stage(static_cast<size_t>(B_) + 2);
            if (stage_.host != nullptr) {
                reinterpret_cast<volatile unsigned *>(stage_.host)[B_] = 0u;
                reinterpret_cast<volatile unsigned *>(stage_.host)[B_ + 1] = 0u;
            }
        }
        const CostStage stage = stage_;
        volatile unsigned *words = stage.host != nullptr ? reinterpret_cast<volatile unsigned *>(stage.host) + B_ : nullptr;
        if (words != nullptr && words[0] != 0u) {  // (a collect that gave up behind the previous call's early return)
            peer_failed_ = true;
            return RNNT_STATUS_EXECUTION_FAILED;
        }
        // (with a peer reduce the caller is promised the world's sum and the exchange's verdict on return: the whole wait,
        // unless told that stream order will do for those too -- mode 2)
        const bool early = grads_dev != nullptr && stage.host != nullptr && !timing_ &&
                           (return_early_ >= 2 || (return_early_ == 1 && peer_.world <= 0));
        costs_mapped_ = stage.host;
        ready_seq_ = 0u;
        if (early) {
            if (++stage_seq_ == 0u) ++stage_seq_;
            ready_seq_ = stage_seq_;
        }
        st = enqueue(blank, stream, grads_dev);
        costs_mapped_ = nullptr;
        ready_seq_ = 0u;
        if (st == RNNT_STATUS_SUCCESS && stage.host == nullptr &&
            cudaMemcpyAsync(costs_host, ws_.costs, sizeof(float) * B_, cudaMemcpyDeviceToHost, stream) != cudaSuccess)
            st = RNNT_STATUS_MEMOPS_FAILED;
        if (st == RNNT_STATUS_SUCCESS && early) {
            // what remains in flight when we return is remembered by an event: the staging buffer and the workspace
            // must outlive it (~Engine, release_owned)
            if (inflight_ev_ == nullptr && cudaEventCreateWithFlags(&inflight_ev_, cudaEventDisableTiming) != cudaSuccess) {
                (void)cudaGetLastError();
                inflight_ev_ = nullptr;
            }
            const bool have_ev = inflight_ev_ != nullptr && cudaEventRecord(inflight_ev_, stream) == cudaSuccess;
            inflight_stream_ = stream;
            inflight_stream_set_ = have_ev;
            if (have_ev) unrecorded_work_ = false;
            stage_busy_ = peer_.world > 0;  // (the exchange writes its verdict into the staging buffer at K3's end)
            bool arrived = false;
            for (unsigned spins = 1; have_ev; ++spins) {
                if (words[1] == stage_seq_) {
                    arrived = true;
                    break;
                }
                if ((spins & 255u) == 0u) {
                    // the stream has run dry (or died) without the word: look once more, then give up on the short cut
                    const cudaError_t q = cudaEventQuery(inflight_ev_);
                    if (q != cudaErrorNotReady) {
                        arrived = q == cudaSuccess && words[1] == stage_seq_;
                        break;
                    }
                }
            }
            if (arrived) {
                std::atomic_thread_fence(std::memory_order_acquire);
            } else if (cudaStreamSynchronize(stream) != cudaSuccess) {
                st = RNNT_STATUS_EXECUTION_FAILED;
            }
        } else if (st == RNNT_STATUS_SUCCESS) {
            if (cudaStreamSynchronize(stream) != cudaSuccess) st = RNNT_STATUS_EXECUTION_FAILED;
            else stage_busy_ = false;
        }
        if (st == RNNT_STATUS_SUCCESS && stage.host != nullptr) {
            std::memcpy(costs_host, stage.host, sizeof(float) * B_);
            if (words[0] != 0u) {
                peer_failed_ = true;
                st = RNNT_STATUS_EXECUTION_FAILED;
            }
        }
        if (st != RNNT_STATUS_SUCCESS && stage.host != nullptr) (void)cudaStreamSynchronize(stream);  // nothing may still write into it
        return st;
    }
    // 0: a synchronous call returns only when everything it launched has completed.  1 (default): early return, except
    // with a peer reduce.  2: early return with a peer reduce as well -- *total_out is then valid in stream order, and an
    // exchange that gave up is reported by the NEXT call.
    void set_return_early(int mode) { return_early_ = mode < 0 ? 1 : (mode > 2 ? 2 : mode); }
    int return_early() const { return return_early_; }

    // Launch everything on `stream` without synchronising; costs stay in workspace().costs.
    RNNTStatus enqueue(int blank, cudaStream_t stream, void *grads_dev) {
        // Under an alignment band the zero rows are most of the call's traffic: the LSE kernel's and the gradient
        // kernel's zero-fill warps share one fill (SHARED protocol, zero_fill.cuh).  Two counters take turns; the
        // lattice kernel of this call clears the one the next call will use.
        shared_fill_ctr_ = shared_fill_clear_ = nullptr;
        zero_dst_ = grads_dev;
        ++tl_call_;  // (MRNNT_TIMELINE: this call's slot)
        if (grads_dev != nullptr && k3_zero_warp_wanted() && zero_fill_possible() && base_ != nullptr) {
            unsigned *pair = ws_.k2_flags + k2_zero_ctr_word(B_) + 2;
            ++shared_seq_;
            shared_fill_ctr_ = pair + (shared_seq_ & 1u);
            shared_fill_clear_ = pair + ((shared_seq_ + 1u) & 1u);
        }
        // the all-GPU sum of the costs (set_peer_reduce): inside the gradient kernel when there is one
        peer_in_k3_ = peer_.world > 0 && grads_dev != nullptr;
        RNNTStatus st = enqueue_forward(blank, stream, grads_dev != nullptr, grads_dev);
        if (st == RNNT_STATUS_SUCCESS && grads_dev != nullptr) {
            k3_follows_k2_ = !timing_;  // (the timing events between the kernels would break the dependent launch)
            st = enqueue_backward(stream, grads_dev, nullptr);
            k3_follows_k2_ = false;
        }
        peer_in_k3_ = false;
        shared_fill_ctr_ = shared_fill_clear_ = nullptr;
        return st;
    }

    // First half of a call: K1 and K2.  With want_grads the lattice kernel also leaves the per-row gradient
    // coefficients in the workspace, and enqueue_backward() may follow at any later time (same or another
    // stream-ordered point) as long as acts and the workspace are untouched: a training framework calls this
    // from its forward pass and enqueue_backward() from its backward pass.
    // zero_dst (enqueue() only): the gradient buffer the backward half is about to fill; the lattice kernel writes
    // its zero rows while the recursions leave the memory system idle.
    RNNTStatus enqueue_forward(int blank, cudaStream_t stream, bool want_grads, void *zero_dst = nullptr) {
        const DeviceInfo &dev = device_info();
        if (!dev.ok) return RNNT_STATUS_EXECUTION_FAILED;
        coef_blank_ = -1;
        zero_dst_ = want_grads ? zero_dst : nullptr;
        dead_rows_zeroed_ = nullptr;
        RNNTStatus st = setup(stream);
        if (st != RNNT_STATUS_SUCCESS) return st;
        mark(0, stream);
        st = launch_k1(blank, stream, dev);
        if (st != RNNT_STATUS_SUCCESS) return st;
        mark(1, stream);
        st = launch_k2(blank, stream, dev, want_grads);
        if (st != RNNT_STATUS_SUCCESS) return st;
        mark(2, stream);
        mark(3, stream);
        if (want_grads) coef_blank_ = blank;
        // a forward half or a cost-only call: no gradient kernel to carry the exchange, it gets a launch of its own
        if (peer_.world > 0 && !peer_in_k3_) {
            peer_reduce_kernel<<<1, kWarp, 0, stream>>>(peer_args(true));
            if (launched() != RNNT_STATUS_SUCCESS) return RNNT_STATUS_EXECUTION_FAILED;
        }
        return RNNT_STATUS_SUCCESS;
    }

    // The all-GPU sum of the summed cost without a collective library (peer_reduce.cuh): `boards[r]` is rank r's board
    // (peer_board_bytes(world) zeroed bytes of device memory) as mapped into this process, `total_out` (device or
    // host-mapped, may be nullptr) receives the world's sum with every one-shot call and every forward half from now
    // on.  All ranks must make the same sequence of such calls.  world == 0 turns it off.
    RNNTStatus set_peer_reduce(int rank, int world, void *const *boards, float *total_out) {
        peer_ = PeerReduce{};
        peer_failed_ = false;  // (the failure belongs to the boards that are being let go of)
        if (stage_.host != nullptr) {
            wait_inflight();
            reinterpret_cast<volatile unsigned *>(stage_.host)[B_] = 0u;
        }
        if (world == 0) return RNNT_STATUS_SUCCESS;
        if (world < 0 || world > kPeerMaxWorld || rank < 0 || rank >= world || boards == nullptr) return RNNT_STATUS_INVALID_VALUE;
        for (int r = 0; r < world; ++r) {
            if (boards[r] == nullptr || (reinterpret_cast<uintptr_t>(boards[r]) & 7)) return RNNT_STATUS_INVALID_VALUE;
            peer_.boards[r] = static_cast<unsigned long long *>(boards[r]);
        }
        peer_.rank = rank;
        peer_.world = world;
        peer_.total_out = total_out;
        peer_.epoch = 0u;
        peer_.timeout_ns = peer_timeout_ns_;
        peer_failed_ = false;
        return RNNT_STATUS_SUCCESS;
    }
    // How long the collect waits for the slowest peer before it gives up for good (0: without limit).
    void set_peer_timeout_ms(unsigned ms) {
        peer_timeout_ns_ = static_cast<unsigned long long>(ms) * 1000000ull;
        peer_.timeout_ns = peer_timeout_ns_;
    }
    bool peer_failed() const { return peer_failed_; }
    // Steps already exchanged through the boards by ANOTHER handle (the epoch belongs to the boards, not to the handle).
    void set_peer_epoch(unsigned epoch) { peer_.epoch = epoch; }
    unsigned peer_epoch() const { return peer_.epoch; }

    // Second half: K3, the gradient w.r.t. the logits.  scale_dev (optional, B floats on the device): utterance
    // b's gradient rows are multiplied by scale_dev[b] as they are written -- the upstream gradient of the
    // per-utterance costs (reference glue: pytorch_binding/monotonic_rnnt_op.py:97-118).
    RNNTStatus enqueue_backward(cudaStream_t stream, void *grads_dev, const float *scale_dev) {
        const DeviceInfo &dev = device_info();
        if (!dev.ok) return RNNT_STATUS_EXECUTION_FAILED;
        if (grads_dev == nullptr || coef_blank_ < 0) return RNNT_STATUS_INVALID_VALUE;
        if (order_behind_inflight(stream) != RNNT_STATUS_SUCCESS) return RNNT_STATUS_EXECUTION_FAILED;
        unrecorded_work_ = true;
        // who writes the plan's dead rows: the lattice kernel has (into this very buffer), or the gradient kernel's own
        // zero-fill warp will, or its consumer warps do
        zero_dst_ = grads_dev;
        k3_zero_warp_ = dead_rows_zeroed_ != grads_dev && k3_zero_warp_wanted() && zero_fill_possible();
        k3_write_dead_ = dead_rows_zeroed_ != grads_dev && !k3_zero_warp_;
        // the lattice kernel's fill stopped at a unit: this kernel's consumer warps write the dead rows behind it
        k3_fill_unit_begin_ = 0;
        if (dead_rows_zeroed_ == grads_dev && k2_fill_unit_end_ >= 0) k3_fill_unit_begin_ = k2_fill_unit_end_;
        k2_fill_unit_end_ = -1;
        dead_rows_zeroed_ = nullptr;  // (good for the one backward pass that follows directly)
        mark(2, stream);
        if (k3_zero_warp_) last_k2_zero_warps_ = 32;
        last_k2_fill_share_ = 100;
        if (k3_fill_unit_begin_ > 0)
            last_k2_fill_share_ = static_cast<int>(k3_fill_unit_begin_ * 100 / ((shape_.rows + kWarp - 1) / kWarp));
        const RNNTStatus st = launch_k3(coef_blank_, stream, dev, grads_dev, scale_dev);
        mark(3, stream);
        return st;
    }

    // Per-kernel device timing for the bench (CUDA events on the launch stream).  Off by default.
    RNNTStatus set_timing(bool on) {
        if (on && !timing_) {
            for (auto &e : ev_)
                if (cudaEventCreate(&e) != cudaSuccess) return RNNT_STATUS_EXECUTION_FAILED;
        } else if (!on && timing_) {
            for (auto &e : ev_) cudaEventDestroy(e);
        }
        timing_ = on;
        return RNNT_STATUS_SUCCESS;
    }

    // Durations (ms) of K1, K2, K3 of the last enqueue(); the stream must have been synchronised.
    RNNTStatus last_timings(float ms[3]) const {
        if (!timing_) return RNNT_STATUS_INVALID_VALUE;
        for (int i = 0; i < 3; ++i)
            if (cudaEventElapsedTime(&ms[i], ev_[i], ev_[i + 1]) != cudaSuccess) return RNNT_STATUS_EXECUTION_FAILED;
        return RNNT_STATUS_SUCCESS;
    }

    ~Engine() {
        if (timing_)
            for (auto &e : ev_) cudaEventDestroy(e);
        // (the gradient kernel of an early-return call may still be running: it has written all it ever writes into the
        // staging buffer -- the costs, then the word we saw -- unless it carries a peer exchange)
        if (stage_busy_) wait_inflight();
        release_owned();  // (mrnnt_destroy without mrnnt_free_workspace)
        if (inflight_ev_ != nullptr) cudaEventDestroy(inflight_ev_);
        if (block_busy_ != nullptr) cudaEventDestroy(block_busy_);
        release_cost_stage(stage_);
    }
    // Returns when nothing an early-return call left behind is still running.
    void wait_inflight() {
        if (inflight_ev_ != nullptr && cudaEventSynchronize(inflight_ev_) != cudaSuccess) (void)cudaGetLastError();
    }

    // Kernel launches this engine has made so far (set-up kernels included).
    unsigned long long launch_count() const { return launches_; }

    // Force the generic (non-TMA) streaming kernels; used by the tests to cross-check both variants.
    void set_force_generic(bool v) { force_generic_ = v; }
    void set_pdl(bool on) { pdl_ = on; }
    void set_k1_compact(int mode) { k1_compact_ = mode < 0 ? -1 : (mode != 0); }
    // SMs the gradient kernel leaves free for a concurrent collective (0: none).
    void set_reserved_sms(int n) { reserved_sms_ = n < 0 ? 0 : n; }
    int last_k2_zero_warps() const { return last_k2_zero_warps_; }
    int last_k2_fill_share() const { return last_k2_fill_share_; }  // percent of the fill's units the lattice kernel took
    int timeline_slot() const { return tl_call_; }
    void set_k2_zero_fill(int warps) { k2_zero_warps_ = warps; }
    // percent of the zero fill that stays in the lattice kernel (-1: automatic); see k2_fill_share()
    void set_k2_fill_share(int pct) { k2_fill_share_ = pct; }
    // the lattice kernel takes part in a fill that the LSE and gradient kernels' zero-fill warps share (default on)
    void set_k2_shared_fill(bool on) { k2_shared_fill_ = on; }
    // the plan in one launch (default) or as the three kernels it replaces (also used above kPlanFusedMaxB utterances)
    void set_fused_plan(bool on) { fused_plan_ = on; }
    // -1 automatic, 0 off, 1 on; 2..100: on, with that percentage of a CTA's share fixed before the counter takes over
    void set_dynamic_tiles(int mode) {
        dynamic_tiles_ = mode < 0 ? -1 : (mode != 0);
        dynamic_fixed_pct_ = mode >= 2 ? (mode > 100 ? 100 : mode) : kDynamicFixedPct;
    }
    // Upper limit for the CTAs per utterance of the lattice kernel's coefficient phase (0: automatic).
    void set_k2_parts(int parts) { k2_parts_ = parts < 0 ? 0 : parts; }
    // Consumer warps per CTA of the streaming kernels (8 or 16); tuning knob for the bench.
    void set_stream_warps(int k1, int k3) {
        if (k1 == 8 || k1 == 16 || k1 == 24) k1_warps_ = k1;
        if (k3 == 8 || k3 == 16 || k3 == 24) k3_warps_ = k3;
    }

   private:
    void mark(int i, cudaStream_t stream) {
        if (timing_) cudaEventRecord(ev_[i], stream);
    }

    RNNTStatus launched() {
        ++launches_;
        return cudaGetLastError() == cudaSuccess ? RNNT_STATUS_SUCCESS : RNNT_STATUS_EXECUTION_FAILED;
    }

    // Hand the block create_workspace() allocated back (to the module's cache, or to the driver).  Like cudaFree it
    // returns only when nothing on the block's device can still touch it.
    // When everything this handle has launched into the block lies in front of the event an early-return call has
    // recorded (the usual case for a caller that makes one synchronous call per manager, like the reference's torch
    // binding), the block goes back to the cache WITH that event and without a wait of the host: the next taker orders
    // its first launch behind it (stream order, a no-op on the same stream).
    void release_owned() {
        if (owned_ == nullptr) return;
        cudaEvent_t busy = nullptr;
        if (inflight_ev_ != nullptr && !unrecorded_work_ && block_busy_ == nullptr && !stage_busy_) {
            busy = inflight_ev_;   // (ownership passes to the cache)
            inflight_ev_ = nullptr;
            inflight_stream_set_ = false;
        } else {
            int cur = -1;
            if (cudaGetDevice(&cur) != cudaSuccess) cur = -1;
            if (cur != owned_device_) (void)cudaSetDevice(owned_device_);
            (void)cudaDeviceSynchronize();
            if (cur != owned_device_ && cur >= 0) (void)cudaSetDevice(cur);
        }
        if (block_busy_ != nullptr) {  // (behind the device synchronisation above)
            (void)cudaEventDestroy(block_busy_);
            block_busy_ = nullptr;
        }
        workspace_cache_put(owned_, owned_bytes_, owned_device_, busy);
        if (base_ == owned_) base_ = nullptr;
        owned_ = nullptr;
    }

    // The kernels an early-return call left behind use the workspace: work enqueued on ANOTHER stream goes behind them.
    RNNTStatus order_behind_inflight(cudaStream_t stream) {
        if (inflight_ev_ != nullptr && inflight_stream_set_ && stream != inflight_stream_ &&
            cudaStreamWaitEvent(stream, inflight_ev_, 0) != cudaSuccess)
            return RNNT_STATUS_EXECUTION_FAILED;
        return RNNT_STATUS_SUCCESS;
    }

    RNNTStatus setup(cudaStream_t stream) {
        if (order_behind_inflight(stream) != RNNT_STATUS_SUCCESS) return RNNT_STATUS_EXECUTION_FAILED;
        if (block_busy_ != nullptr) {  // the workspace block came out of the cache with its last user still running
            const cudaError_t e = cudaStreamWaitEvent(stream, block_busy_, 0);
            (void)cudaEventDestroy(block_busy_);
            block_busy_ = nullptr;
            if (e != cudaSuccess) return RNNT_STATUS_EXECUTION_FAILED;
        }
        unrecorded_work_ = true;  // (until an early-return call records its event behind everything)
        if ((plan_dirty_ || band_dirty_) && B_ <= kPlanFusedMaxB && fused_plan_) {
            // one launch for the whole plan (plan.cuh: plan_fused_kernel)
            size_t smem = alignment_ != nullptr ? (static_cast<size_t>(shape_.T_max) + 1) * sizeof(int) : 0;
            if (alignment_ != nullptr && align_stride_ != 0 && align_stride_ < shape_.T_max) return RNNT_STATUS_INVALID_VALUE;
            if (smem > 48 * 1024) {
                if (smem > static_cast<size_t>(device_info().max_smem_optin) - 2048) return RNNT_STATUS_INVALID_VALUE;
                if (!ensure_dynamic_smem(plan_fused_kernel, smem)) return RNNT_STATUS_EXECUTION_FAILED;
            }
            const int64_t block_rows = static_cast<int64_t>(shape_.T_dim) * shape_.U;  // (padded layout; 0: packed)
            const int64_t max_rows = block_rows > 0 ? block_rows : static_cast<int64_t>(shape_.T_max) * (shape_.S_max + 1);
            int64_t parts = (max_rows + kPlanFusedRowsPerCta - 1) / kPlanFusedRowsPerCta;
            parts = parts < 1 ? 1 : (parts > kPlanFusedMaxParts ? kPlanFusedMaxParts : parts);
            plan_fused_kernel<<<dim3(static_cast<unsigned>(parts), static_cast<unsigned>(B_)), kPlanFusedThreads, smem, stream>>>(
                T_dev_, S_dev_, B_, shape_.T_max, shape_.label_stride, shape_.U, block_rows, alignment_,
                align_stride_ > 0 ? align_stride_ : shape_.T_max, max_shift_, align_blank_, ws_.row_start, ws_.band, ws_.rowmeta,
                ws_.rowutt, ws_.k2_flags);
            if (launched() != RNNT_STATUS_SUCCESS) return RNNT_STATUS_EXECUTION_FAILED;
            plan_dirty_ = false;
            band_dirty_ = false;
        }
        if (plan_dirty_) {
            plan_row_start_kernel<<<1, kPlanThreads, 0, stream>>>(
                T_dev_, S_dev_, B_, ws_.row_start, ws_.k2_flags, static_cast<int64_t>(shape_.T_dim) * shape_.U);
            if (launched() != RNNT_STATUS_SUCCESS) return RNNT_STATUS_EXECUTION_FAILED;
            plan_dirty_ = false;  // (both counters of the shared zero fill are clear now: either may come first)
        }
        if (band_dirty_) {
            size_t smem = alignment_ != nullptr ? (static_cast<size_t>(shape_.T_max) + 1) * sizeof(int) : 0;
            if (smem > 48 * 1024) {
                if (smem > static_cast<size_t>(device_info().max_smem_optin) - 1024) return RNNT_STATUS_INVALID_VALUE;
                if (!ensure_dynamic_smem(band_kernel, smem)) return RNNT_STATUS_EXECUTION_FAILED;
            }
            if (alignment_ != nullptr && align_stride_ != 0 && align_stride_ < shape_.T_max) return RNNT_STATUS_INVALID_VALUE;
            band_kernel<<<B_, kBandThreads, smem, stream>>>(T_dev_, S_dev_, shape_.T_max, alignment_,
                                                            align_stride_ > 0 ? align_stride_ : shape_.T_max, max_shift_,
                                                            align_blank_, ws_.band);
            if (launched() != RNNT_STATUS_SUCCESS) return RNNT_STATUS_EXECUTION_FAILED;
            const int64_t blocks64 = (shape_.rows + 255) / 256;
            const int blocks = static_cast<int>(blocks64 > 65535 * 8 ? 65535 * 8 : blocks64);
            rowmeta_kernel<<<blocks, 256, 0, stream>>>(T_dev_, S_dev_, B_, shape_.T_max, shape_.label_stride, shape_.U,
                                                       ws_.row_start, ws_.band, ws_.rowmeta, ws_.rowutt);
            if (launched() != RNNT_STATUS_SUCCESS) return RNNT_STATUS_EXECUTION_FAILED;
            band_dirty_ = false;
        }
        return RNNT_STATUS_SUCCESS;
    }

    // Pick the streaming (TMA-staged) configuration: the preferred consumer-warp count if a safe ring exists
    // for it, else 8 warps, else none (generic kernels).
    bool can_stream(const void *p0, const void *p1, size_t extra_per_row, int preferred_warps, int tile_target,
                    bool whole_tiles, const DeviceInfo &dev, StreamTiling *tl) const {
        if (force_generic_) return false;
        if ((reinterpret_cast<uintptr_t>(p0) & 15) || (reinterpret_cast<uintptr_t>(p1) & 15)) return false;
        // rows that are not whole 16-byte vectors stream too (aligned windows, k1_lse.cuh), for float32 logits
        if (bf16_ && (static_cast<size_t>(V_) * elem_bytes()) % 16 != 0) return false;
        const int order[3] = {preferred_warps, 16, 8};
        for (int w : order) {
            if (stream_tiling(V_, elem_bytes(), extra_per_row, w, tile_target, whole_tiles, tl) &&
                tl->smem_bytes <= static_cast<size_t>(dev.max_smem_optin))
                return true;
        }
        return false;
    }

    int generic_grid(const DeviceInfo &dev) const {
        const int64_t want = (shape_.rows + kGenericWarps - 1) / kGenericWarps;
        const int64_t cap = static_cast<int64_t>(dev.sm_count) * 8;
        return static_cast<int>(want < cap ? (want < 1 ? 1 : want) : cap);
    }

    template <typename E, int NW, int C, bool COMPACT, bool UNALIGNED = false>
    RNNTStatus launch_k1_variant(int blank, cudaStream_t stream, const DeviceInfo &dev, const StreamTiling &tl) {
        auto kern = k1_lse_tma_kernel<E, NW, C, COMPACT, UNALIGNED>;
        // (the kernel's zero-fill warp: one more warp, 8 KB more shared memory; the gradient kernel continues the fill)
        ZeroFill zero{};
        const bool zero_warp = COMPACT && shared_fill_ctr_ != nullptr;
        if (zero_warp) {
            zero = zero_fill_args(zero_dst_, shared_fill_ctr_);
        }
        zero.tl_slot = tl_call_;
        const size_t smem = k1_smem_bytes(tl.smem_bytes, zero_warp);
        if (smem > static_cast<size_t>(dev.max_smem_optin) || !ensure_dynamic_smem(kern, smem)) return RNNT_STATUS_EXECUTION_FAILED;
        // (a programmatic dependent of whatever precedes it in the stream: behind the previous call's gradient kernel,
        // which lets its dependents go at once, this kernel's CTAs set themselves up on the SMs that kernel's CTAs leave
        // and wait there for the rest of them; behind anything else the attribute changes nothing)
        if (launch_kernel(kern, dev.sm_count, (NW + (zero_warp ? 2 : 1)) * kWarp, smem, stream, pdl_ && !timing_,
                          static_cast<const E *>(acts_), labels_, ws_.rowmeta, ws_.lp, shape_.rows, V_, blank, tl.G, tl.stages,
                          zero, tl.smem_bytes, tl.slot_bytes) != cudaSuccess)
            return RNNT_STATUS_EXECUTION_FAILED;
        return launched();
    }

    // Many dead tiles expected -- an alignment band, a padded tensor, or tiles of so few rows that the dead corners of
    // the lattice fill whole tiles (c4: tiles of 2 rows, 15 % of them dead, K1 2211 -> 2030 us): the variant of K1 that
    // gives them no ring slot.  On c2 / c3 (tiles of 8 / 16 rows) it costs 1-4 us, so it is not the default there.
    bool k1_compact(const StreamTiling &tl) const {
        if (k1_compact_ >= 0) return k1_compact_ != 0;
        return alignment_ != nullptr || shape_.U > 0 || tl.G <= 4;
    }

    // The gradient kernel's tiles handed out through a counter instead of round-robin by CTA index (kK3Dynamic,
    // k3_grad.cuh).  What it removes is the end of the kernel, where the CTAs of a fixed split finish up to ~15 us apart
    // (ncu: SMs active 91 % of the kernel's duration on c2); what it costs, measured, is a few percent of streaming
    // rate on long inputs.  tools/kernel_times.py --dyn 0,1: c2 K3 205 -> 191 us, bfloat16 c2 unchanged, c3 1290 ->
    // 1368 us, c4 4469 -> 4588 us, c5 (alignment band, most tiles dead: one request per dead tile) 521 -> 664 us.
    // Hence: dense inputs of up to kDynamicTilesPerCta tiles per CTA.
    static constexpr int64_t kDynamicTilesPerCta = 400;
    bool dynamic_tiles(const StreamTiling &tl) const {
        const int64_t ntiles = (shape_.rows + tl.G - 1) / tl.G;
        const int64_t sms = device_info().sm_count;
        if (ntiles + 16 * sms >= (int64_t(1) << 31)) return false;
        if (dynamic_tiles_ >= 0) return dynamic_tiles_ != 0;
        return alignment_ == nullptr && shape_.U == 0 && ntiles <= kDynamicTilesPerCta * sms;
    }

    template <typename E, int NW, int C>
    RNNTStatus launch_k1_tma(int blank, cudaStream_t stream, const DeviceInfo &dev, const StreamTiling &tl) {
        if (tl.unaligned) {
            // rows that are not whole 16-byte vectors (float32 only; k1_lse.cuh: StreamWindow): one variant, the compact one
            if constexpr (std::is_same<E, float>::value) return launch_k1_variant<E, NW, C, true, true>(blank, stream, dev, tl);
            else return RNNT_STATUS_EXECUTION_FAILED;  // (can_stream never chooses it for bfloat16)
        }
        return k1_compact(tl) ? launch_k1_variant<E, NW, C, true>(blank, stream, dev, tl)
                              : launch_k1_variant<E, NW, C, false>(blank, stream, dev, tl);
    }

    // C = 16-byte vectors a lane keeps in registers: a row in 32 floats per lane, in 64, or two passes over smem
    template <typename E, int NW>
    RNNTStatus launch_k1_nw(int blank, cudaStream_t stream, const DeviceInfo &dev, const StreamTiling &tl) {
        constexpr int NE = Elem<E>::kPerVec;
        const int NV = row_vectors<E>();
        if (NV <= (32 / NE) * kWarp) return launch_k1_tma<E, NW, 32 / NE>(blank, stream, dev, tl);

        if constexpr (NW < 24) {  // (launch_k1_typed never asks for 24 warps with 64 registers of row per lane)
            if (NV <= (64 / NE) * kWarp) return launch_k1_tma<E, NW, 64 / NE>(blank, stream, dev, tl);
        }
        return launch_k1_tma<E, NW, 0>(blank, stream, dev, tl);
    }

    template <typename E>
    RNNTStatus launch_k1_typed(int blank, cudaStream_t stream, const DeviceInfo &dev) {
        StreamTiling tl;
        // rows held in 64 registers per lane are too many for 25 warps on one SM
        constexpr int NE = Elem<E>::kPerVec;
        const int NV = row_vectors<E>();
        const bool wide_regs = NV > (32 / NE) * kWarp && NV <= (64 / NE) * kWarp;
        const int want = (k1_warps_ == 24 && wide_regs) ? 16 : k1_warps_;
        const size_t input_bytes = static_cast<size_t>(shape_.rows) * V_ * elem_bytes();
        const int tile_target = input_bytes < kK1SmallInputBytes ? kK1TileTargetSmall : kK1TileTarget;
        if (can_stream(acts_, acts_, sizeof(int), want, tile_target, false, dev, &tl)) {  // (8 bytes of slot metadata per row)
            return tl.warps == 8    ? launch_k1_nw<E, 8>(blank, stream, dev, tl)
                   : tl.warps == 16 ? launch_k1_nw<E, 16>(blank, stream, dev, tl)
                                    : launch_k1_nw<E, 24>(blank, stream, dev, tl);
        }
        k1_lse_generic_kernel<E><<<generic_grid(dev), kGenericWarps * kWarp, 0, stream>>>(
            static_cast<const E *>(acts_), labels_, ws_.rowmeta, ws_.lp, shape_.rows, V_, blank);
        return launched();
    }

    RNNTStatus launch_k1(int blank, cudaStream_t stream, const DeviceInfo &dev) {
        return bf16_ ? launch_k1_typed<__nv_bfloat16>(blank, stream, dev) : launch_k1_typed<float>(blank, stream, dev);
    }

    template <int K>
    RNNTStatus launch_k2_warp(K2Args args, cudaStream_t stream, const DeviceInfo &dev) {
        auto kern = k2_lattice_kernel<K>;
        const size_t smem = k2_smem_bytes(shape_.width(), args.row_warps);
        if (!ensure_dynamic_smem(kern, smem)) return RNNT_STATUS_EXECUTION_FAILED;
        if (args.parts > 1) {
            // The CTAs of an utterance wait for each other: the whole grid has to be co-resident.  How many CTAs fit
            // on an SM depends on the ring size of this launch; asked once per (kernel, shared-memory size).
            if (k2_occ_smem_ != smem || k2_occ_kernel_ != reinterpret_cast<const void *>(kern)) {
                int occ = 0;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kK2Threads, smem) != cudaSuccess || occ < 1)
                    return RNNT_STATUS_EXECUTION_FAILED;
                k2_occ_ = occ;
                k2_occ_smem_ = smem;
                k2_occ_kernel_ = reinterpret_cast<const void *>(kern);
            }
            const int fit = k2_occ_ * dev.sm_count / B_;
            if (args.parts > fit) args.parts = fit < 1 ? 1 : fit;
        }
        args.phase_ctas = B_ * args.parts;

        // with the zero fill: one CTA on every SM, the ones behind the utterances' CTAs only fill
        const int grid = args.zero_warps > 0 && args.phase_ctas < dev.sm_count ? dev.sm_count : args.phase_ctas;
        if (launch_kernel(kern, grid, kK2Threads, smem, stream, pdl_ && !timing_, args) != cudaSuccess)
            return RNNT_STATUS_EXECUTION_FAILED;
        return launched();
    }

    RNNTStatus launch_k2(int blank, cudaStream_t stream, const DeviceInfo &dev, bool need_beta) {
        K2Args a;
        a.T = T_dev_; a.S = S_dev_; a.labels = labels_; a.row_start = ws_.row_start; a.band = ws_.band;
        a.lp = ws_.lp; a.wts = ws_.wts; a.alpha = ws_.alpha; a.beta = ws_.beta; a.coef = ws_.coef; a.rowlab = ws_.rowlab;
        a.ll_fwd = ws_.ll_fwd; a.ll_bwd = ws_.ll_bwd; a.costs = ws_.costs;
        a.costs_mapped = need_beta ? nullptr : costs_mapped_;  // (with gradients: K3 mirrors the costs, see cost_mirror)
        a.T_max = shape_.T_max; a.S_max = shape_.S_max; a.V = V_; a.blank = blank;
        a.ld = shape_.U; a.T_dim = shape_.T_dim; a.label_stride = shape_.label_stride;
        a.need_beta = need_beta ? 1 : 0;
        a.chunk_frames = k2_chunk_frames(shape_.width());
        // Coefficient phase: spread every utterance over `parts` CTAs while the whole grid still fits on the
        // machine at one CTA per SM (the extra CTAs wait for their utterance's recursion on idle SMs).
        int parts = need_beta ? kK2MaxParts : 1;  // (launch_k2_warp lowers it to what is co-resident)
        if (k2_parts_ > 0 && k2_parts_ < parts) parts = k2_parts_;
        a.parts = parts;
        a.flags = ws_.k2_flags;
        if (++epoch_ == 0u) ++epoch_;
        a.epoch = epoch_;
        const int states = shape_.S_max + 1;
        const int K = k2_states_per_lane(states);
        a.zero_dst = nullptr;
        a.rowmeta = ws_.rowmeta;
        a.row_bytes = static_cast<unsigned>(static_cast<size_t>(V_) * elem_bytes());
        a.zero_warps = 0;
        a.zero_unit_end = -1;
        a.zero_shared_ctr = nullptr;
        k2_fill_unit_end_ = -1;
        a.zero_clear = shared_fill_clear_;
        last_k2_zero_warps_ = 0;
        a.rows = shape_.rows;
        a.B = B_;
        a.phase_ctas = 0;
        a.tl_slot = tl_call_;
        // (a padded tensor with many label positions per frame can ask for more shared memory than an SM has while its
        // longest label sequence still fits the row of warps: the block-wide kernel takes it then)
        const bool warp_kernel_fits =
            K > 0 && k2_smem_bytes(shape_.width(), k2_row_warps(states, K)) <= static_cast<size_t>(dev.max_smem_optin);
        if (warp_kernel_fits) {
            a.row_warps = k2_row_warps(states, K);
            a.chunk_bufs = k2_chunk_bufs(a.row_warps);
            // the zero fill needs bulk-copy granularity (16 bytes) and whole warps next to the chain warps and
            // their chunk issuers
            const int spare = kK2Warps - 1 - 2 * (a.row_warps + 1);
            int zw = k2_zero_warps_ < 0 ? (zero_fill_pays() ? 2 : 0) : (k2_zero_warps_ & 15);
            if (zw > spare) zw = spare;
            if (need_beta && zero_dst_ != nullptr && zw > 0 && !k3_zero_warp_wanted() && zero_fill_possible()) {
                a.zero_dst = static_cast<unsigned char *>(zero_dst_);
                a.zero_warps = zw;
                last_k2_zero_warps_ = zw;
                dead_rows_zeroed_ = zero_dst_;
                // the front of the batch here, the rest by the gradient kernel's zero-fill warp (k2_fill_share)
                const int share = k2_fill_share();
                if (share < 100) {
                    const int64_t nunits = (shape_.rows + kWarp - 1) / kWarp;
                    k2_fill_unit_end_ = nunits * share / 100;
                }
            } else if (need_beta && zero_dst_ != nullptr && shared_fill_ctr_ != nullptr && spare > 0 && k2_shared_fill_ &&
                       zero_fill_possible()) {
                // the fill the LSE kernel's zero-fill warp has begun and the gradient kernel's will finish (SHARED
                // protocol): two warps per CTA of this kernel take units from the same counter while the recursions run
                a.zero_dst = static_cast<unsigned char *>(zero_dst_);
                a.zero_warps = spare < 2 ? spare : 2;
                a.zero_shared_ctr = shared_fill_ctr_;
            }
            a.zero_unit_end = k2_fill_unit_end_;
            return K == 1   ? launch_k2_warp<1>(a, stream, dev)
                   : K == 2 ? launch_k2_warp<2>(a, stream, dev)
                            : launch_k2_warp<4>(a, stream, dev);
        }
        a.parts = 1;
        a.row_warps = 0;
        a.chunk_bufs = 0;
        const size_t smem = (static_cast<size_t>(shape_.S_max) + 2) * sizeof(Cell);
        if (smem > static_cast<size_t>(dev.max_smem_optin)) return RNNT_STATUS_INVALID_VALUE;
        if (!ensure_dynamic_smem(k2_lattice_wide_kernel, smem)) return RNNT_STATUS_EXECUTION_FAILED;
        k2_lattice_wide_kernel<<<B_, kK2Threads, smem, stream>>>(a);
        return launched();
    }

    ZeroFill zero_fill_args(void *dst, unsigned *ctr) const {
        ZeroFill z;
        z.dst = static_cast<unsigned char *>(dst);
        z.rowmeta = ws_.rowmeta;
        z.rows = shape_.rows;
        z.row_bytes = static_cast<unsigned>(static_cast<size_t>(V_) * elem_bytes());
        z.ctr = ctr;
        return z;
    }

    // the fill stores 4-byte words at the ragged ends of a run of rows and bulk copies in between: an aligned base,
    // rows of whole words
    bool zero_fill_possible() const {
        const size_t row_bytes = static_cast<size_t>(V_) * elem_bytes();
        return row_bytes % 4 == 0 && row_bytes >= 64 && reinterpret_cast<uintptr_t>(zero_dst_) % 16 == 0 &&
               row_bytes < (1ull << 31);
    }

    // The zero fill as one more warp of the gradient kernel instead (k3_grad.cuh) -- and, in a one-shot call, of the
    // LSE kernel before it, the two sharing one hand-out counter (enqueue()): where writing the zeros takes longer
    // than everything else -- an alignment band leaves a few percent of the rows alive (c5: 4.45 GB of zeros against
    // 0.7 GB of live traffic) -- it should overlap the live rows' work, not sit between the kernels.  Measured on c5
    // (tools/kernel_times.py --zero 0,2,32): 1.03 ms with the gradient kernel's consumers writing the zeros, 0.92 ms
    // with the lattice kernel's fill, 0.89 ms with the gradient kernel's zero-fill warp alone, 0.87 ms shared with the
    // LSE kernel's (5.15 GB of traffic in all: 5.9 TB/s over the whole call).  On c2 it loses against the lattice
    // kernel's fill (0.396 against 0.381 ms).
    bool k3_zero_warp_wanted() const {
        if (k2_zero_warps_ >= 0) return k2_zero_warps_ >= 32;
        // a band of 2 * max_shift + 1 states (at most) around the alignment, out of S + 1: "nearly all dead" = under half alive
        return alignment_ != nullptr && 2 * (static_cast<int64_t>(max_shift_) + 1) < (shape_.S_max + 1) / 2;
    }

    // Zero fill inside the lattice kernel (automatic choice).  What it wins is the duration of the recursions, during
    // which the memory system would idle (~0.17 us per frame); moving the zero rows out of the gradient kernel is
    // otherwise neutral at best (there they are written in the shadow of the live rows' arithmetic).  So: where many
    // rows are dead (alignment band, padded tensor: the gradient kernel then also skips whole tiles), and where the
    // recursions are a visible share of the call.  Measured (tools/kernel_times.py --zero 0,2): c2 -3 %, c3 -4 %,
    // c5 -12 %, c4 +1.5 % (off there).
    bool zero_fill_pays() const {
        if (alignment_ != nullptr || shape_.U > 0) return true;
        const double chain_us = 0.17 * shape_.T_max;
        const double stream_us = 3.0 * static_cast<double>(shape_.rows) * V_ * static_cast<double>(elem_bytes()) / 6.5e6;
        return chain_us >= 0.025 * stream_us;
    }

    // How much of the zero fill the lattice kernel takes (percent of the batch's units of 32 rows, from the front, 1..100);
    // the gradient kernel writes the dead rows behind that itself.  The idea: the lattice kernel alone ends when its fill
    // does (c2: ~50 us) although its recursions and coefficient phase are over after 26, so a fill that ends with the
    // phases might let the gradient kernel start earlier than the rest of the zeros costs it.  Measured
    // (tools/share_sweep.py, the whole call in the stream), two ways of writing the rest:
    //  * the gradient kernel's CONSUMER warps (what is built: k3_grad.cuh, own_dead_from): c2 322.8 us at 100 %, 325.0 at
    //    90 %, 327.4 at 80 %, 330.2 at 60 %; a quarter of c3 567 -> 572 at 65 %; an eighth 312.3 -> 312.8; the whole of
    //    c3 2113 -> 2101 at 55 % (-0.5 %) and back up to 2123 at 40 %;
    //  * one more warp per CTA of the gradient kernel issuing bulk stores (ranges in zero_fill.cuh): c2 321.7 -> 343.6 at
    //    80 %, 365.5 at 60 % -- its stores queue in the per-SM bulk-copy pipe in front of the producer's loads.
    // In the stream the lattice kernel's tail already overlaps the gradient kernel's set-up (dependent launch), and what
    // the fill moves costs the same bandwidth wherever it is written: nothing to win.  Hence 100 unless asked otherwise
    // (MRNNT_OPT_K2_FILL_SHARE).
    int k2_fill_share() const {
        if (k2_fill_share_ >= 0) return k2_fill_share_ > 100 ? 100 : (k2_fill_share_ < 1 ? 1 : k2_fill_share_);
        return 100;
    }

    // (dead tiles are skipped under the same conditions as in K1, and only when nobody has to zero them here)
    int k3_flags(const StreamTiling &tl) const {
        int dyn = 0;
        if (dynamic_tiles(tl)) {
            const int64_t share = (shape_.rows + tl.G - 1) / tl.G / device_info().sm_count;
            const int64_t fixed = share * dynamic_fixed_pct_ / 100;
            dyn = kK3Dynamic | (static_cast<int>(fixed < 2 ? 2 : fixed > 0x3fffff ? 0x3fffff : fixed) << kK3FixedShift);
        }
        if (k3_write_dead_) return kK3WriteDead | dyn;
        return dyn ? dyn : (k1_compact(tl) ? kK3Compact : 0);
    }

    template <typename E, int NW, bool SCALED>
    RNNTStatus launch_k3_tma(int blank, cudaStream_t stream, const DeviceInfo &dev, const StreamTiling &tl, void *grads,
                             const float *scale) {
        if (tl.unaligned) {
            if constexpr (std::is_same<E, float>::value) return launch_k3_tma_variant<E, NW, SCALED, true>(blank, stream, dev, tl, grads, scale);
            else return RNNT_STATUS_EXECUTION_FAILED;
        }
        return launch_k3_tma_variant<E, NW, SCALED, false>(blank, stream, dev, tl, grads, scale);
    }

    template <typename E, int NW, bool SCALED, bool UNALIGNED>
    RNNTStatus launch_k3_tma_variant(int blank, cudaStream_t stream, const DeviceInfo &dev, const StreamTiling &tl,
                                     void *grads, const float *scale) {
        auto kern = k3_grad_tma_kernel<E, NW, SCALED, UNALIGNED>;
        // the kernel's own zero-fill warp (one more warp, 8 KB more shared memory)
        ZeroFill zero{};
        const bool zero_warp = k3_zero_warp_ && !k3_write_dead_;
        int flags = k3_flags(tl);
        if (zero_warp) {
            zero = zero_fill_args(grads, shared_fill_ctr_ != nullptr ? shared_fill_ctr_ : ws_.k2_flags + k2_zero_ctr_word(B_));
            if (shared_fill_ctr_ != nullptr) flags |= kK3ZeroShared;
        }
        zero.tl_slot = tl_call_;
        const int64_t own_dead_from = (k3_fill_unit_begin_ > 0 && !k3_write_dead_) ? k3_fill_unit_begin_ * kWarp : INT64_MAX;
        const size_t smem = k3_smem_bytes(tl.smem_bytes, zero_warp);
        if (smem > static_cast<size_t>(dev.max_smem_optin) || !ensure_dynamic_smem(kern, smem)) return RNNT_STATUS_EXECUTION_FAILED;
        // K3 is persistent with one CTA per SM; `reserved_sms_` of them can be left to a collective that runs
        // concurrently on another stream (the all-reduce of the summed cost, which is final after K2)
        const int grid = dev.sm_count - reserved_sms_ > 0 ? dev.sm_count - reserved_sms_ : 1;
        // (a dependent launch only right behind the lattice kernel: a backward pass called on its own has no such
        // predecessor to wait for, and an ordinary launch orders it behind whatever precedes it in the stream)
        if (launch_kernel(kern, grid, (NW + (zero_warp ? 2 : 1)) * kWarp, smem, stream, pdl_ && k3_follows_k2_,
                          static_cast<const E *>(acts_), ws_.coef, ws_.rowlab, static_cast<E *>(grads), shape_.rows, V_, blank, tl.G,
                          tl.stages, ws_.rowutt, scale, cost_mirror(), flags, zero, tl.smem_bytes,
                          ws_.k2_flags + stream_ctr_word(B_), peer_args(peer_in_k3_), tl.slot_bytes, own_dead_from) != cudaSuccess)
            return RNNT_STATUS_EXECUTION_FAILED;
        return launched();
    }

    template <typename E>
    RNNTStatus launch_k3_typed(int blank, cudaStream_t stream, const DeviceInfo &dev, void *grads, const float *scale) {
        StreamTiling tl;
        if (can_stream(acts_, grads, sizeof(float4) + sizeof(int), k3_warps_, kK3TileTarget, true, dev, &tl)) {
            if (scale != nullptr)
                return tl.warps == 8    ? launch_k3_tma<E, 8, true>(blank, stream, dev, tl, grads, scale)
                       : tl.warps == 16 ? launch_k3_tma<E, 16, true>(blank, stream, dev, tl, grads, scale)
                                        : launch_k3_tma<E, 24, true>(blank, stream, dev, tl, grads, scale);
            return tl.warps == 8    ? launch_k3_tma<E, 8, false>(blank, stream, dev, tl, grads, scale)
                   : tl.warps == 16 ? launch_k3_tma<E, 16, false>(blank, stream, dev, tl, grads, scale)
                                    : launch_k3_tma<E, 24, false>(blank, stream, dev, tl, grads, scale);
        }
        k3_grad_generic_kernel<E><<<generic_grid(dev), kGenericWarps * kWarp, 0, stream>>>(
            static_cast<const E *>(acts_), ws_.coef, ws_.rowlab, static_cast<E *>(grads), shape_.rows, V_, blank, ws_.rowutt, scale,
            cost_mirror(), peer_args(peer_in_k3_));
        return launched();
    }

    RNNTStatus launch_k3(int blank, cudaStream_t stream, const DeviceInfo &dev, void *grads, const float *scale) {
        return bf16_ ? launch_k3_typed<__nv_bfloat16>(blank, stream, dev, grads, scale)
                     : launch_k3_typed<float>(blank, stream, dev, grads, scale);
    }

    size_t elem_bytes() const { return bf16_ ? 2 : 4; }

    // 16-byte vectors that lie wholly inside a row of logits, at most (what the consumer warps keep in registers: the up
    // to two edge vectors of a row that does not start / end on a 16-byte boundary are handled apart, k1_lse.cuh: RowEdges)
    template <typename E>
    int row_vectors() const {
        return V_ / Elem<E>::kPerVec;
    }

    // the exchange of this call (a new epoch), or none
    PeerReduce peer_args(bool active) {
        PeerReduce p{};
        if (!active || peer_.world <= 0) return p;
        if (++peer_.epoch == 0u) peer_.epoch = 2u;  // (0 is the boards' initial state; keep the parity sequence)
        p = peer_;
        p.costs = ws_.costs;
        p.B = B_;
        p.status_out = costs_mapped_ != nullptr ? reinterpret_cast<unsigned *>(costs_mapped_) + B_ : nullptr;
        return p;
    }

    // who copies the costs into a synchronous call's host-mapped staging buffer: the gradient kernel when there is
    // one (the write then overlaps it), else the lattice kernel
    CostMirror cost_mirror() const {
        CostMirror m;
        m.costs = ws_.costs;
        m.mapped = costs_mapped_;
        m.B = B_;
        if (costs_mapped_ != nullptr && ready_seq_ != 0u) {
            m.ready = reinterpret_cast<unsigned *>(costs_mapped_) + B_ + 1;
            m.seq = ready_seq_;
        }
        return m;
    }

    const void *acts_;   // float32 (the reference's type) or bfloat16 (set_bf16)
    bool bf16_ = false;
    const int *labels_;
    const int *T_dev_;
    const int *S_dev_;
    int B_, V_;

    int pad_T_ = 0, pad_U_ = 0, label_stride_ = 0;  // padded layout (0: packed)
    std::vector<int> T_h_, S_h_;
    Shape shape_;
    bool have_shape_ = false;
    RNNTStatus shape_status_ = RNNT_STATUS_UNKNOWN_ERROR;

    void *base_ = nullptr;
    void *owned_ = nullptr;
    size_t owned_bytes_ = 0;
    int owned_device_ = 0;
    Workspace ws_;
    bool plan_dirty_ = true;
    bool band_dirty_ = true;
    const int *alignment_ = nullptr;
    int align_stride_ = 0;  // ints per utterance in alignment_ (0: max_b T_b)
    unsigned long long launches_ = 0;
    int max_shift_ = 0;
    int align_blank_ = 0;
    bool force_generic_ = false;
    static constexpr size_t kUploadMinCopyBytes = size_t(1) << 20;  // a middle block below 1 MiB stays with the kernel
    size_t upload_copy_min_bytes_ = kUploadMinCopyBytes;            // 0: no copy-engine part at all
    int k1_warps_ = 24;
    int k1_compact_ = -1;  // -1: automatic (k1_compact()), 0 / 1: forced
    static constexpr int kDynamicFixedPct = 0;  // (0: two fixed tiles per CTA, everything else through the counter)
    int dynamic_fixed_pct_ = kDynamicFixedPct;
    int dynamic_tiles_ = -1;  // the gradient kernel's tiles through a counter: -1 automatic (dynamic_tiles()), 0 / 1 forced
    int k3_warps_ = 24;
    int k2_parts_ = 0;     // 0: automatic
    int k2_occ_ = 0;       // CTAs of the lattice kernel per SM for (k2_occ_kernel_, k2_occ_smem_)
    size_t k2_occ_smem_ = 0;
    const void *k2_occ_kernel_ = nullptr;
    int reserved_sms_ = 0;
    bool pdl_ = true;             // programmatic dependent launch of K2 behind K1 and of K3 behind K2
    void *zero_dst_ = nullptr;          // enqueue(): gradient buffer handed to the lattice kernel's zero fill
    void *dead_rows_zeroed_ = nullptr;  // the buffer whose dead rows the last lattice kernel zeroed (nullptr: none)
    bool k3_write_dead_ = true;         // the gradient kernel's consumer warps write the zero rows
    bool k3_zero_warp_ = false;         // ... or its own zero-fill warp does
    unsigned *shared_fill_ctr_ = nullptr;    // enqueue(): the counter K1's and K3's zero-fill warps share in this call
    unsigned *shared_fill_clear_ = nullptr;  // ... and the one this call's K2 clears for the next call
    unsigned shared_seq_ = 0u;
    int last_k2_zero_warps_ = 0;        // what the last lattice launch ran with
    int k2_zero_warps_ = -1;            // warps per lattice CTA for the zero fill: -1 automatic, 0 off
    int last_k2_fill_share_ = 100;
    bool k2_shared_fill_ = true;
    bool fused_plan_ = true;
    int tl_call_ = -1;                  // calls enqueued so far - 1 (MRNNT_TIMELINE)
    int k2_fill_share_ = -1;            // percent of the fill's units the lattice kernel takes: -1 automatic
    int64_t k2_fill_unit_end_ = -1;     // where the last lattice kernel's fill stops (-1: it takes everything)
    int64_t k3_fill_unit_begin_ = 0;    // ... and where the gradient kernel's zero-fill warp therefore starts
    PeerReduce peer_{};           // set_peer_reduce(): world > 0 when on; epoch = the last one used
    unsigned long long peer_timeout_ns_ = kPeerDefaultTimeoutNs;
    bool peer_failed_ = false;    // a collect of this handle gave up: every later synchronous call fails
    bool peer_in_k3_ = false;     // this call's exchange rides in the gradient kernel
    bool k3_follows_k2_ = false;  // K3 is being enqueued directly behind K2 (enqueue(), not a separate backward)
    unsigned epoch_ = 0u;  // launch counter published through Workspace::k2_flags
    float *costs_mapped_ = nullptr;  // set for the duration of a synchronous compute(): host-mapped copy of the costs
    CostStage stage_{};              // this handle's staging buffer (B costs, the exchange's flag, the sequence word)
    bool stage_tried_ = false;
    int return_early_ = 1;           // compute() with gradients returns when the costs are on the host (set_return_early)
    unsigned stage_seq_ = 0u;        // the last sequence number handed to a gradient kernel
    unsigned ready_seq_ = 0u;        // set for the duration of an early-return compute(): this call's sequence number
    cudaEvent_t inflight_ev_ = nullptr;  // behind the last early-return call's kernels
    cudaStream_t inflight_stream_ = nullptr;
    bool inflight_stream_set_ = false;
    bool unrecorded_work_ = false;   // something was enqueued that inflight_ev_ does not stand behind
    bool stage_busy_ = false;        // the last early-return call's kernels may still write into the staging buffer
    cudaEvent_t block_busy_ = nullptr;  // create_workspace(): the cached block's last user (see CachedBlock)
    int coef_blank_ = -1;  // blank label of the forward pass whose coefficients sit in the workspace (-1: none)
    bool timing_ = false;
    cudaEvent_t ev_[4] = {};
};

}  // namespace mrnnt
