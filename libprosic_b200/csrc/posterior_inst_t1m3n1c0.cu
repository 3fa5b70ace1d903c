// explicit instantiation of posterior_kernel<TEAM=1, MINB=3, NSC=1, CONT=0> (see posterior_kernel.cuh)
#include "posterior_kernel.cuh"
namespace vlr {
VLR_POSTERIOR_INSTANTIATE(1, 3, 1, 0)
}
