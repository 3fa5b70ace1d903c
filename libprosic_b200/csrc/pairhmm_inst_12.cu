// explicit instantiation of the pair-HMM kernels: BANDED=true, APPROX=2 (see pairhmm_kernel.cuh)
#include "pairhmm_kernel.cuh"
namespace vlr {
VLR_PAIRHMM_INSTANTIATE(true, 2)
}
