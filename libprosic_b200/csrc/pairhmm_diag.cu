// explicit instantiation of the anti-diagonal log-space pair-HMM kernels (see pairhmm_diag.cuh)
#include "pairhmm_diag.cuh"
namespace vlr {
VLR_PAIRHMM_DIAG_INSTANTIATE(0)
VLR_PAIRHMM_DIAG_INSTANTIATE(1)
VLR_PAIRHMM_DIAG_INSTANTIATE(2)
}
