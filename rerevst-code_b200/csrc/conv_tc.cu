// placeholder until the tcgen05 kernel lands (next commit)
#include "rrv_common.cuh"
namespace rrv {
int conv2d_tc(const rrv_conv*, cudaStream_t) { set_error("rrv_conv2d: tcgen05 path not built yet"); return 1; }
long long tc_weight_bytes(int, int, int, int) { return 0; }
int pack_weights_tc(const float*, int, int, int, int, void*, cudaStream_t) { set_error("rrv_pack_weights_tc: not built yet"); return 1; }
}
