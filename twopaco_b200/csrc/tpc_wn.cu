// explicit instantiation of the junction-finding kernels for long k-mers: TPC_W = 5..19 64-bit words (k = 129..603),
// one object per word count (Makefile: -DTPC_W=<n>)
#include "tpc_launch_impl.cuh"
#ifndef TPC_W
#error "compile with -DTPC_W=<words per k-mer>"
#endif
template struct tpc::Launch<TPC_W>;
