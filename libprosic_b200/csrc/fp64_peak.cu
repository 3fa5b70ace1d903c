// fp64_peak.cu -- register-resident DFMA micro-benchmark: the fp64 roofline denominator
// (SURVEY.md section 8d asks the builder to measure it; MEASURED_PEAKS.json has no fp64 entry).
#include "common.cuh"

namespace vlr {

template <int CHAINS>
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double a, double b) {
    double v[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) v[k] = (double)(threadIdx.x + k) * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int k = 0; k < CHAINS; ++k) v[k] = fma(v[k], a, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) s += v[k];
    if (s == 12345.678) out[0] = s;  // keep the chains alive
}

}  // namespace vlr

extern "C" int vlr_measure_fp64_peak(vlr_ctx *ctx, double *out_tflops) {
    if (!ctx || !out_tflops) return VLR_ERR_INVALID;
    VLR_CUDA(ctx, cudaSetDevice(ctx->device));
    constexpr int CH = 8;
    double *d = (double *)vlr::dev_scratch(ctx, 15, 64);
    if (!d) return vlr::set_err(ctx, VLR_ERR_NOMEM, "scratch");
    cudaEvent_t e0, e1;
    VLR_CUDA(ctx, cudaEventCreate(&e0));
    VLR_CUDA(ctx, cudaEventCreate(&e1));
    const int grid = ctx->sm_count * 8, block = 256, iters = 4096;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        VLR_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        vlr::dfma_peak_kernel<CH><<<grid, block, 0, ctx->stream>>>(d, iters, 0.999999, 1e-9);
        VLR_LAUNCH_CHECK(ctx);
        VLR_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        VLR_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0.f;
        VLR_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * CH * 8.0 * (double)iters * (double)grid * (double)block;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *out_tflops = best;
    return VLR_OK;
}
