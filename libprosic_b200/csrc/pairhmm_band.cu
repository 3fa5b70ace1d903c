// explicit instantiation of the band-only banded pair-HMM kernels (see pairhmm_band.cuh)
#include "pairhmm_band.cuh"
namespace vlr {
VLR_PAIRHMM_BAND_INSTANTIATE(true)
VLR_PAIRHMM_BAND_INSTANTIATE(false)
}
