// ctx.cu -- context management of libvlr_b200 (C ABI: include/vlr_gpu.h).
#include <math.h>
#include <stdarg.h>

#include "common.cuh"
#include "detmath.cuh"

namespace vlr {

int set_err(vlr_ctx *ctx, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

static void *grow(std::vector<vlr_ctx::Buf> &bufs, int slot, size_t bytes, bool pinned) {
    if ((int)bufs.size() <= slot) bufs.resize(slot + 1);
    auto &b = bufs[slot];
    if (b.cap >= bytes && b.p) return b.p;
    if (b.p) {
        if (pinned) cudaFreeHost(b.p); else cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
    }
    size_t cap = align_up(bytes + bytes / 4 + 4096, 4096);
    cudaError_t e = pinned ? cudaMallocHost(&b.p, cap) : cudaMalloc(&b.p, cap);
    if (e != cudaSuccess) {
        b.p = nullptr;
        return nullptr;
    }
    b.cap = cap;
    return b.p;
}

void *dev_scratch(vlr_ctx *ctx, int slot, size_t bytes) {
    // a grow may free memory still in use by enqueued work: drain first
    if ((int)ctx->dev_bufs.size() <= slot || ctx->dev_bufs[slot].cap < bytes)
        cudaStreamSynchronize(ctx->stream);
    return grow(ctx->dev_bufs, slot, bytes, false);
}
void *pin_scratch(vlr_ctx *ctx, int slot, size_t bytes) {
    if ((int)ctx->pin_bufs.size() <= slot || ctx->pin_bufs[slot].cap < bytes)
        cudaStreamSynchronize(ctx->stream);
    return grow(ctx->pin_bufs, slot, bytes, true);
}

int ensure_copy_stream(vlr_ctx *ctx) {
    if (ctx->copy_stream) return VLR_OK;
    VLR_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
        VLR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_h2d[k], cudaEventDisableTiming));
        VLR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_done[k], cudaEventDisableTiming));
    }
    return VLR_OK;
}

__host__ __device__ static inline double ref_math_eval(int fn, double x) {
    switch (fn) {
    case 0: return dm::exp(x);
    case 1: return dm::expm1(x);
    case 2: return dm::log(x);
    default: return dm::log1p(x);
    }
}
__global__ void ref_math_kernel(int fn, int64_t n, const double *x, double *out) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
        out[k] = ref_math_eval(fn, x[k]);
}

}  // namespace vlr

extern "C" {

int vlr_version(void) { return 200; }
#ifndef VLR_BUILD_STAMP
#define VLR_BUILD_STAMP "unstamped"
#endif
const char *vlr_build_stamp(void) { return VLR_BUILD_STAMP; }

int vlr_ref_math(vlr_ctx *ctx, int fn, int64_t n, const double *x, double *out) {
    if (fn < 0 || fn > 3 || n < 0 || (n > 0 && (!x || !out))) return ctx ? vlr::set_err(ctx, VLR_ERR_INVALID, "bad argument") : VLR_ERR_INVALID;
    if (!ctx) {  // host evaluation
        for (int64_t k = 0; k < n; ++k) out[k] = vlr::ref_math_eval(fn, x[k]);
        return VLR_OK;
    }
    if (n == 0) return VLR_OK;
    VLR_CUDA(ctx, cudaSetDevice(ctx->device));
    double *d_x = (double *)vlr::dev_scratch(ctx, 92, (size_t)n * 8), *d_o = (double *)vlr::dev_scratch(ctx, 93, (size_t)n * 8);
    if (!d_x || !d_o) return vlr::set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
    VLR_CUDA(ctx, cudaMemcpyAsync(d_x, x, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    vlr::ref_math_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(fn, n, d_x, d_o);
    VLR_LAUNCH_CHECK(ctx);
    VLR_CUDA(ctx, cudaMemcpyAsync(out, d_o, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    VLR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VLR_OK;
}

int vlr_ctx_create(int device, vlr_ctx **out) {
    if (!out) return VLR_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || device < 0 || device >= n) return VLR_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return VLR_ERR_CUDA;
    vlr_ctx *ctx = new vlr_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return VLR_ERR_CUDA; }
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return VLR_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    // base-quality LUT, computed the way the reference does (bases.rs:23-48):
    // ln eps = q * (-ln10/10); ln(1-eps) = ln_one_minus_exp(ln eps); mismatch = ln eps + ln .3333
    // (the library's deterministic libm, detmath.cuh: the literal log-space kernel and a host
    // evaluation of the same formulas see identical table entries)
    std::vector<double> lut(256 * vlr::kQlutStride);
    const double f = -0.23025850929940456;
    const double ln_conf = vlr::dm::log(0.3333);
    for (int q = 0; q < 256; ++q) {
        double ln_eps = (double)q * f;
        double ln_call = vlr::dm::ln_one_minus_exp(ln_eps);
        lut[q * 4 + 0] = vlr::dm::exp(ln_eps);
        lut[q * 4 + 1] = vlr::dm::exp(ln_call);
        lut[q * 4 + 2] = vlr::dm::exp(ln_eps + ln_conf);
        lut[q * 4 + 3] = ln_call;
    }
    if (cudaMalloc(&ctx->d_qlut, lut.size() * sizeof(double)) != cudaSuccess ||
        cudaMemcpy(ctx->d_qlut, lut.data(), lut.size() * sizeof(double), cudaMemcpyHostToDevice) !=
            cudaSuccess) {
        vlr_ctx_destroy(ctx);
        return VLR_ERR_CUDA;
    }
    *out = ctx;
    return VLR_OK;
}

void vlr_ctx_destroy(vlr_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto &b : ctx->dev_bufs) if (b.p) cudaFree(b.p);
    for (auto &b : ctx->pin_bufs) if (b.p) cudaFreeHost(b.p);
    if (ctx->d_qlut) cudaFree(ctx->d_qlut);
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
        for (int k = 0; k < 2; ++k) {
            if (ctx->ev_h2d[k]) cudaEventDestroy(ctx->ev_h2d[k]);
            if (ctx->ev_done[k]) cudaEventDestroy(ctx->ev_done[k]);
        }
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char *vlr_last_error(const vlr_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int vlr_ctx_set_stream(vlr_ctx *ctx, void *cuda_stream) {
    if (!ctx) return VLR_ERR_INVALID;
    cudaStream_t ns = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    if (ns != ctx->stream) ctx->model_dev = nullptr;  // the cached model upload was ordered on the old stream
    ctx->stream = ns;
    return VLR_OK;
}

int vlr_ctx_synchronize(vlr_ctx *ctx) {
    if (!ctx) return VLR_ERR_INVALID;
    VLR_CUDA(ctx, cudaSetDevice(ctx->device));
    VLR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VLR_OK;
}

int64_t vlr_ctx_launch_count(const vlr_ctx *ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
