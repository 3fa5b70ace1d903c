// explicit instantiation of the pair-HMM kernels: BANDED=false, APPROX=1 (see pairhmm_kernel.cuh)
#include "pairhmm_kernel.cuh"
namespace vlr {
VLR_PAIRHMM_INSTANTIATE(false, 1)
}
