// explicit instantiation of the junction-finding kernels for k-mers of 4 64-bit word(s)
#include "tpc_launch_impl.cuh"
template struct tpc::Launch<4>;
