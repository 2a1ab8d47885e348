// explicit instantiation of the junction-finding kernels for k-mers of 1 64-bit word(s)
#include "tpc_launch_impl.cuh"
template struct tpc::Launch<1>;
