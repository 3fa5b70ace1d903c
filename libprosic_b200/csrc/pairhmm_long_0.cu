// explicit instantiation of the long-read pair-HMM kernels: APPROX=0 (see pairhmm_kernel.cuh)
#include "pairhmm_kernel.cuh"
namespace vlr {
VLR_PAIRHMM_LONG_INSTANTIATE(0)
}
