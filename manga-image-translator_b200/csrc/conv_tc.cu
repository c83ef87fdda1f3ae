// tcgen05 (5th-gen tensor core) implicit-GEMM convolution for sm_100a -- placeholder until the kernel lands.
#include "mitb_internal.h"
namespace mitb {
bool conv_tc_supported(const ConvOp&) { return false; }
void launch_conv_tc(const ConvOp&, cudaStream_t) { MITB_CHECK(false, "tcgen05 conv path not built"); }
}  // namespace mitb
