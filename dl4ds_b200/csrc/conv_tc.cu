// tcgen05 (5th-gen tensor core) implicit-GEMM convolution kernels -- placeholder dispatch until the
// kernels land: every shape reports DL4DS_E_UNSUPPORTED so api.cu routes to the CUDA-core fp32 path.
#include "common.cuh"

namespace dl4ds {

int conv2d_fwd_tc(const ConvArgs&, int, cudaStream_t) { return DL4DS_E_UNSUPPORTED; }
int conv2d_wgrad_tc(const WgradArgs&, void*, int, cudaStream_t) { return DL4DS_E_UNSUPPORTED; }
int64_t conv2d_wgrad_tc_workspace(int, int, int, int, int, int, int) { return 0; }

}  // namespace dl4ds
