// explicit instantiation of posterior_kernel<TEAM=4, MINB=3, NSC=2, CONT=1> (see posterior_kernel.cuh)
#include "posterior_kernel.cuh"
namespace vlr {
VLR_POSTERIOR_INSTANTIATE(4, 3, 2, 1)
}
