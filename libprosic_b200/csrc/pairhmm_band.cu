// explicit instantiation of the band-only banded pair-HMM kernels (see pairhmm_band.cuh)
#include "pairhmm_band.cuh"
namespace vlr {
VLR_PAIRHMM_BAND_INSTANTIATE(0)
VLR_PAIRHMM_BAND_INSTANTIATE(1)
VLR_PAIRHMM_BAND_INSTANTIATE(2)
}
