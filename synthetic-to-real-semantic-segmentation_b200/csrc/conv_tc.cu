// tcgen05 / TMA implicit-GEMM convolution (sm_100a).  Placeholder until the UMMA path lands:
// reports "not handled" so s2r_conv_fwd falls through to the mma.sync tap-GEMM.
#include "common.cuh"

int s2r_conv_fwd_tc(const s2r_conv_args* a, cudaStream_t st) {
  (void)a;
  (void)st;
  return 0;
}
