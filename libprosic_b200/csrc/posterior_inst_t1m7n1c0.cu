// explicit instantiation of posterior_kernel<TEAM=1, MINB=7, NSC=1, CONT=0> (see posterior_kernel.cuh)
#include "posterior_kernel.cuh"
namespace vlr {
VLR_POSTERIOR_INSTANTIATE(1, 7, 1, 0)
}
