// explicit instantiation of posterior_kernel<TEAM=1, MINB=7, NSC=2, CONT=0> (see posterior_kernel.cuh)
#include "posterior_kernel.cuh"
namespace vlr {
VLR_POSTERIOR_INSTANTIATE(1, 7, 2, 0)
}
