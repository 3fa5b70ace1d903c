// stubs.cu -- entry points declared in include/vlr_gpu.h whose kernels are not built yet.
// They fail loudly (VLR_ERR_UNSUPPORTED); there is no CPU fallback in this library.
#include "common.cuh"

extern "C" {

int vlr_besthit_batch(vlr_ctx *ctx, int64_t, const uint8_t *, const int64_t *, int64_t,
                      const uint8_t *, const int64_t *, int64_t, const int32_t *, const int32_t *,
                      int32_t *, int32_t *, int32_t *, int32_t *) {
    return vlr::set_err(ctx, VLR_ERR_UNSUPPORTED, "vlr_besthit_batch: kernel not built yet");
}
}
