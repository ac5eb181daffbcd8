// hp_tc.cu -- tensor-core (tcgen05 / TMEM / TMA) variant of the handposedd hot path.
// PLACEHOLDER while the FP32 path is brought up: every entry reports UNSUPPORTED.
#include "hp_common.cuh"
namespace hp {
int tc_init(Net &) { return 0; }
void tc_destroy(Net &) {}
int tc_refresh_weights(Net &, cudaStream_t) { return 0; }
int tc_forward(Net &, const float *, int64_t, float *, cudaStream_t)
{
    set_error("tensor-core path not built yet");
    return 2;
}
}  // namespace hp
