// C++ harness for the drop-in C++ surface: compute_rnnt_loss (include/rnnt_entrypoint.h) and the
// GpuRNNTWorkspaceManager / GpuRNNTComputer classes, driven the way the reference's own GPU tests drive
// them (reference tests/test_gpu.cu: fwd_test :16-83, multibatch_test, align_restrict_test -- device
// buffers via cudaMalloc, a fresh stream, create_workspace / free_workspace, one manager reused across
// restrict_to_alignment + cost calls).  Expected values are the reference's (tests/test_cpu.cpp:57,
// :291-294, :399-430, :510-546) at its own tolerance (is_close = 1e-4, rnnt_helper.h:10-14; grads 1e-2).
// Links against libmonotonic_rnnt.so for compute_rnnt_loss; prints "ABI OK" and exits 0 on success.
#include <cmath>
#include <cstdio>
#include <vector>

#include <cuda_runtime.h>

#include "cpu_rnnt.h"  // the name-only CPU shells must compile next to the GPU classes (monotonic_rnnt.cu:12-13)
#include "gpu_rnnt.h"
#include "gpu_workspace_manager.h"
#include "rnnt_entrypoint.h"

static int failures = 0;
#define CHECK(cond, msg)                                               \
    do {                                                               \
        if (!(cond)) {                                                 \
            std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, msg);   \
            ++failures;                                                \
        }                                                              \
    } while (0)

template <typename T>
static T *to_gpu(const std::vector<T> &v) {
    T *p = nullptr;
    cudaMalloc(&p, v.size() * sizeof(T));
    cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    return p;
}
static bool close_to(float a, double b, double tol = 1e-4) { return std::fabs(a - b) < tol; }

static const float kReadme[36] = {0.6f, 0.3f, 0.1f, 0.7f, 0.1f, 0.2f, 0.5f, 0.1f, 0.4f, 0.5f, 0.4f, 0.1f,
                                  0.5f, 0.1f, 0.4f, 0.8f, 0.1f, 0.1f, 0.4f, 0.3f, 0.3f, 0.5f, 0.1f, 0.4f,
                                  0.7f, 0.2f, 0.1f, 0.8f, 0.1f, 0.1f, 0.3f, 0.1f, 0.6f, 0.8f, 0.1f, 0.1f};
static const float kReadmeGrads[36] = {0.04f, -0.14f, 0.1f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.13f, -0.19f, 0.06f,
                                       -0.04f, 0.04f, -0.01f, 0.f, 0.f, 0.f, 0.06f, -0.1f, 0.04f, 0.01f, 0.07f, -0.08f,
                                       -0.06f, 0.04f, 0.02f, 0.f, 0.f, 0.f, 0.14f, 0.05f, -0.19f, -0.11f, 0.05f, 0.05f};

static std::vector<float> readme_logits(int copies) {
    std::vector<float> l;
    for (int c = 0; c < copies; ++c)
        for (float p : kReadme) l.push_back(std::log(p));
    return l;
}

int main() {
    cudaStream_t stream;
    cudaStreamCreate(&stream);
    RNNTOptions opt{};
    opt.loc = RNNT_GPU;
    opt.blank_label = 0;
    opt.stream = stream;
    opt.num_threads = 0;

    {  // fwd / bwd / grads on the README example through the C entry point
        auto logits = readme_logits(1);
        float *acts = to_gpu(logits);
        int *labels = to_gpu(std::vector<int>{1, 2}), *T = to_gpu(std::vector<int>{4}), *S = to_gpu(std::vector<int>{2});
        float *grads = nullptr;
        cudaMalloc(&grads, 36 * sizeof(float));
        GpuRNNTWorkspaceManager<float> wm(acts, labels, 1, T, S, 3);
        CHECK(wm.create_workspace() == RNNT_STATUS_SUCCESS, "create_workspace");
        float cost_only = 0.f, cost = 0.f;
        CHECK(compute_rnnt_loss(wm, opt, &cost_only, nullptr) == RNNT_STATUS_SUCCESS, "compute_rnnt_loss cost");
        CHECK(compute_rnnt_loss(wm, opt, &cost, grads) == RNNT_STATUS_SUCCESS, "compute_rnnt_loss cost+grad");
        CHECK(close_to(cost_only, -std::log(0.363)), "fwd_test cost");
        CHECK(cost_only == cost, "bwd_test: cost() == cost_and_grad()");
        std::vector<float> g(36);
        cudaMemcpy(g.data(), grads, 36 * sizeof(float), cudaMemcpyDeviceToHost);
        for (int i = 0; i < 36; ++i) CHECK(std::fabs(g[i] - kReadmeGrads[i]) < 1e-2, "grads_test");
        // error conventions (src/rnnt_entrypoint.cpp:18-20,45-46)
        CHECK(compute_rnnt_loss(wm, opt, nullptr, nullptr) == RNNT_STATUS_INVALID_VALUE, "null costs");
        RNNTOptions cpu = opt;
        cpu.loc = RNNT_CPU;
        CHECK(compute_rnnt_loss(wm, cpu, &cost, nullptr) == RNNT_STATUS_EXECUTION_FAILED, "no CPU path");
        RNNTOptions bad = opt;
        bad.loc = static_cast<rnntComputeLocation>(7);
        CHECK(compute_rnnt_loss(wm, bad, &cost, nullptr) == RNNT_STATUS_INVALID_VALUE, "unknown loc");
        CpuRNNTWorkspaceManager<float> cpu_wm(nullptr, nullptr, 1, nullptr, nullptr, 3);
        CHECK(compute_rnnt_loss(cpu_wm, opt, &cost, nullptr) == RNNT_STATUS_INVALID_VALUE, "wrong manager type");
        CHECK(cpu_wm.create_workspace() == RNNT_STATUS_EXECUTION_FAILED, "CPU shell fails loudly");
        wm.free_workspace();
        // caller-owned workspace (TensorFlow path, tensorflow_binding/monotonic_rnnt_op.cu:103-125)
        size_t bytes = 0;
        GpuRNNTWorkspaceManager<float> wm2(acts, labels, 1, T, S, 3);
        CHECK(wm2.get_workspace_size(&bytes) == RNNT_STATUS_SUCCESS && bytes > 0, "get_workspace_size");
        void *ws = nullptr;
        cudaMalloc(&ws, bytes);
        wm2.set_workspace(ws);
        GpuRNNTComputer<float> computer(wm2, 0, stream);
        float c2 = 0.f;
        CHECK(computer.cost_and_grad(&c2, grads) == RNNT_STATUS_SUCCESS && c2 == cost, "set_workspace path");
        cudaFree(ws);
        cudaFree(acts); cudaFree(labels); cudaFree(T); cudaFree(S); cudaFree(grads);
    }
    {  // validation (gpu_workspace_manager.h:232-239)
        int *T = to_gpu(std::vector<int>{2}), *S = to_gpu(std::vector<int>{3});
        GpuRNNTWorkspaceManager<float> wm(nullptr, nullptr, 1, T, S, 3);
        size_t bytes = 0;
        CHECK(wm.get_workspace_size(&bytes) == RNNT_STATUS_INVALID_VALUE, "T < S rejected");
        CHECK(wm.create_workspace() == RNNT_STATUS_INVALID_VALUE, "create_workspace rejects too");
        GpuRNNTWorkspaceManager<float> wm0(nullptr, nullptr, 0, T, S, 3);
        CHECK(wm0.get_workspace_size(&bytes) == RNNT_STATUS_INVALID_VALUE, "B <= 0 rejected");
        cudaFree(T); cudaFree(S);
    }
    {  // multibatch_test: packed B=2, T={2,4}, S={1,2}, labels {1,0,1,2} (stride S_max=2)
        std::vector<float> probs = {0.6f, 0.3f, 0.1f, 0.7f, 0.1f, 0.2f, 0.5f, 0.4f, 0.1f, 0.5f, 0.1f, 0.4f};
        std::vector<float> logits;
        for (float p : probs) logits.push_back(std::log(p));
        auto r = readme_logits(1);
        logits.insert(logits.end(), r.begin(), r.end());
        float *acts = to_gpu(logits);
        int *labels = to_gpu(std::vector<int>{1, 0, 1, 2}), *T = to_gpu(std::vector<int>{2, 4}), *S = to_gpu(std::vector<int>{1, 2});
        GpuRNNTWorkspaceManager<float> wm(acts, labels, 2, T, S, 3);
        CHECK(wm.create_workspace() == RNNT_STATUS_SUCCESS, "create_workspace mb");
        float costs[2];
        CHECK(compute_rnnt_loss(wm, opt, costs, nullptr) == RNNT_STATUS_SUCCESS, "mb cost");
        CHECK(close_to(costs[0], -std::log(0.39)) && close_to(costs[1], -std::log(0.363)), "multibatch costs");
        wm.free_workspace();
        cudaFree(acts); cudaFree(labels); cudaFree(T); cudaFree(S);
    }
    {  // align_restrict_multibatch_test: one manager, restrict_to_alignment called repeatedly
        auto logits = readme_logits(2);
        float *acts = to_gpu(logits);
        int *labels = to_gpu(std::vector<int>{1, 2, 1, 2}), *T = to_gpu(std::vector<int>{4, 4}), *S = to_gpu(std::vector<int>{2, 2});
        int *al = to_gpu(std::vector<int>{0, 1, 0, 2, 1, 2, 0, 0});
        GpuRNNTWorkspaceManager<float> wm(acts, labels, 2, T, S, 3);
        CHECK(wm.create_workspace() == RNNT_STATUS_SUCCESS, "create_workspace align");
        GpuRNNTComputer<float> computer(wm, 0, stream);
        float c[2];
        computer.cost(c);
        CHECK(close_to(c[0], -std::log(0.363)) && close_to(c[1], -std::log(0.363)), "unrestricted");
        wm.restrict_to_alignment(al, 3, 0);
        computer.cost(c);
        CHECK(close_to(c[0], -std::log(0.363)) && close_to(c[1], -std::log(0.363)), "shift 3");
        wm.restrict_to_alignment(al, 0, 0);
        computer.cost(c);
        CHECK(close_to(c[0], -std::log(0.072)) && close_to(c[1], -std::log(0.0672)), "shift 0");
        wm.restrict_to_alignment(al, 1, 0);
        computer.cost(c);
        CHECK(close_to(c[0], -std::log(0.2958)) && close_to(c[1], -std::log(0.192)), "shift 1");
        wm.free_workspace();
        cudaFree(acts); cudaFree(labels); cudaFree(T); cudaFree(S); cudaFree(al);
    }
    cudaStreamDestroy(stream);
    CHECK(cudaGetLastError() == cudaSuccess, "no pending CUDA error");
    if (failures == 0) std::printf("ABI OK\n");
    return failures == 0 ? 0 : 1;
}
