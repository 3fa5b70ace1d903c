// stubs.cu -- entry points declared in include/vlr_gpu.h whose kernels are not built yet.
// They fail loudly (VLR_ERR_UNSUPPORTED); there is no CPU fallback in this library.
#include "common.cuh"

extern "C" {

int vlr_besthit_batch(vlr_ctx *ctx, int64_t, const uint8_t *, const int64_t *, int64_t,
                      const uint8_t *, const int64_t *, int64_t, const int32_t *, const int32_t *,
                      int32_t *, int32_t *, int32_t *, int32_t *) {
    return vlr::set_err(ctx, VLR_ERR_UNSUPPORTED, "vlr_besthit_batch: kernel not built yet");
}
int vlr_likelihood_batch(vlr_ctx *ctx, int64_t, int32_t, const vlr_obs_columns *, const int64_t *,
                         const vlr_biases *, int32_t, const vlr_lh_item *, int64_t, double *) {
    return vlr::set_err(ctx, VLR_ERR_UNSUPPORTED, "vlr_likelihood_batch: kernel not built yet");
}
int vlr_posterior_batch(vlr_ctx *ctx, int64_t, const vlr_model *, const vlr_obs_columns *,
                        const int64_t *, double *, double *, double *, int32_t *) {
    return vlr::set_err(ctx, VLR_ERR_UNSUPPORTED, "vlr_posterior_batch: kernel not built yet");
}
int vlr_posterior_batch_dev(vlr_ctx *ctx, int64_t, const vlr_model *, const vlr_obs_columns *,
                            const int64_t *, double *, double *, double *, int32_t *) {
    return vlr::set_err(ctx, VLR_ERR_UNSUPPORTED, "vlr_posterior_batch_dev: kernel not built yet");
}
}
