// explicit instantiation of the junction-finding kernels for k-mers of 3 64-bit word(s)
#include "tpc_launch_impl.cuh"
template struct tpc::Launch<3>;
