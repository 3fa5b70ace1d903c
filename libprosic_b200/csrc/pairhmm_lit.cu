// explicit instantiation of the literal log-space pair-HMM kernels (LogLit, see pairhmm_kernel.cuh)
#include "pairhmm_kernel.cuh"
#include "pairhmm_litband.cuh"
namespace vlr {
VLR_PAIRHMM_LITERAL_INSTANTIATE(0)
VLR_PAIRHMM_LITERAL_INSTANTIATE(1)
VLR_PAIRHMM_LITERAL_INSTANTIATE(2)
VLR_PAIRHMM_LITBAND_INSTANTIATE(0)
VLR_PAIRHMM_LITBAND_INSTANTIATE(1)
VLR_PAIRHMM_LITBAND_INSTANTIATE(2)
}
