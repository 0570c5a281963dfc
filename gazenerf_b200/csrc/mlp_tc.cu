// placeholder: tcgen05 fused MLP (implemented next)
#include "common.cuh"
using namespace gnrf;
extern "C" size_t gnrf_mlp_tc_packed_bytes(void) { return 0; }
extern "C" size_t gnrf_mlp_tc_bias_floats(void) { return 0; }
extern "C" int gnrf_mlp_tc_pack(const float* const*, void*, gnrf_stream_t) { return fail(GNRF_ERR_UNSUPPORTED, "tc path not built yet"); }
extern "C" int gnrf_mlp_tc_fold(const void*, const float*, const float*, int, float*, gnrf_stream_t) { return fail(GNRF_ERR_UNSUPPORTED, "tc path not built yet"); }
extern "C" size_t gnrf_mlp_tc_workspace_bytes(int, int, int) { return 0; }
extern "C" int gnrf_mlp_tc_fwd(int, const void* const*, const float* const*, const float*, const float*, const float*, int, int, int,
                               float* const*, float* const*, float* const*, void*, size_t, gnrf_stream_t) { return fail(GNRF_ERR_UNSUPPORTED, "tc path not built yet"); }
