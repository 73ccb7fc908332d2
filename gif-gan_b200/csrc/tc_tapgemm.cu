// placeholder until the tcgen05 kernels land (next commit): the tensor-core path reports
// "unsupported" loudly instead of silently falling back.
#include "common.cuh"
namespace gg {
int tc_conv_down(const gg_conv_desc*, const void*, const void*, const float*, void*, cudaStream_t) { set_error("tcgen05 conv_down not built"); return GG_ERR_UNSUPPORTED; }
int tc_conv_up(const gg_conv_desc*, const void*, const void*, const float*, void*, cudaStream_t) { set_error("tcgen05 conv_up not built"); return GG_ERR_UNSUPPORTED; }
int tc_conv_wgrad(const gg_conv_desc*, const void*, const void*, float*, cudaStream_t) { set_error("tcgen05 conv_wgrad not built"); return GG_ERR_UNSUPPORTED; }
}
