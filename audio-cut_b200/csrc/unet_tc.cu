// tcgen05 (sm_100a tensor core) kernels of the U-Net - placeholder until the kernel lands.
#include "unet_kernels.cuh"
namespace ac {
int tc_conv3x3_supported(int, int, int) { return AC_E_INVALID; }
int tc_conv3x3_pack(const float*, int, TcConvWeights** out) { *out = nullptr; return AC_OK; }
void tc_conv3x3_free(TcConvWeights*) {}
int launch_tc_conv3x3(const TcConvArgs&, cudaStream_t) { set_error("tc conv not built"); return AC_E_INVALID; }
}  // namespace ac
