// common.cuh -- shared declarations of libvlr_b200 (sm_100a only; no CPU fallback).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/vlr_gpu.h"

struct vlr_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;  // stream all work is enqueued on
    std::string err;
    int64_t launches = 0;
    int sm_count = 0;
    // device LUT: per base quality {eps, 1-eps, eps*0.3333} (linear) followed by the ln versions
    double *d_qlut = nullptr;
    // grow-only device / pinned scratch owned by the context (host-buffer entry points)
    struct Buf {
        void *p = nullptr;
        size_t cap = 0;
    };
    std::vector<Buf> dev_bufs;   // indexed by slot
    std::vector<Buf> pin_bufs;
    // host-buffer entry points overlap H2D of chunk c+1 with the kernels of chunk c
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    // flattened model of the last posterior launch (host copy) and where it sits on the device:
    // the same model is not uploaded again (the chunks of one batch, repeated calls of one scenario)
    std::vector<uint8_t> model_copy;
    const void *model_dev = nullptr;
};

namespace vlr {

int set_err(vlr_ctx *ctx, int code, const char *fmt, ...);
// returns device scratch of at least `bytes` in slot `slot` (grow-only), nullptr on failure
void *dev_scratch(vlr_ctx *ctx, int slot, size_t bytes);
void *pin_scratch(vlr_ctx *ctx, int slot, size_t bytes);
// creates the copy stream and its events on first use; VLR_OK or an error code
int ensure_copy_stream(vlr_ctx *ctx);

#define VLR_CUDA(ctx, call)                                                                  \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return vlr::set_err((ctx), VLR_ERR_CUDA, "%s failed: %s (%s:%d)", #call,         \
                                cudaGetErrorString(e__), __FILE__, __LINE__);                \
    } while (0)

#define VLR_LAUNCH_CHECK(ctx)                                                                \
    do {                                                                                     \
        cudaError_t e__ = cudaGetLastError();                                                \
        if (e__ != cudaSuccess)                                                              \
            return vlr::set_err((ctx), VLR_ERR_CUDA, "kernel launch failed: %s (%s:%d)",     \
                                cudaGetErrorString(e__), __FILE__, __LINE__);                \
        (ctx)->launches++;                                                                   \
    } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Host-buffer entry points enqueue asynchronous copies FROM / TO caller memory before they can fail
// (validation, allocation, a launch error).  Every error exit goes through this: the caller may free
// or reuse its buffers as soon as the call returns, so both streams are drained first.
static inline int drain_on_error(vlr_ctx *ctx, int rc) {
    if (rc != VLR_OK && ctx) {
        if (ctx->stream) cudaStreamSynchronize(ctx->stream);
        if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    }
    return rc;
}

// Host-side validation of large index / offset arrays runs at the memory bandwidth of ONE core (a few
// GB/s: 5 ms for the 2 M pairs of the C3 batch, serial in front of the first copy).  f(t, begin, end)
// is called for `n_threads` contiguous ranges of [0, n), range 0 on the calling thread.
template <class F>
static inline void parallel_ranges(int64_t n, int n_threads, F f) {
    if (n_threads < 1) n_threads = 1;
    if ((int64_t)n_threads > n) n_threads = n > 0 ? (int)n : 1;
    std::vector<std::thread> th;
    th.reserve((size_t)n_threads);
    int started = 1;  // ranges [1, started) run on their own threads
    try {
        for (; started < n_threads; ++started)
            th.emplace_back(f, started, n * started / n_threads, n * (started + 1) / n_threads);
    } catch (...) {  // no more threads to be had: the remaining ranges run here (nothing crosses the C ABI)
    }
    f(0, (int64_t)0, n / n_threads);
    for (int t = started; t < n_threads; ++t) f(t, n * t / n_threads, n * (t + 1) / n_threads);
    for (std::thread &x : th) x.join();
}
static inline int validation_threads(int64_t items) {
    if (items < 262144) return 1;
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::min<unsigned>(8u, std::max<unsigned>(1u, hw / 4));
}

// qlut layout: [256][4] doubles: eps, 1-eps, eps*0.3333, ln(1-eps)
constexpr int kQlutStride = 4;

}  // namespace vlr
