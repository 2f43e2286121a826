// K2 placeholder TU: replaced by the tcgen05 grouped DeepQN forward.
#include "common.cuh"
using namespace cev;
extern "C" int cev_deepqn_forward(cev_handle* h, const float* members, int P, int64_t pitch,
                                  const uint8_t* frames, int B, int c_in, int n_actions,
                                  float* logits, int32_t* actions, cev_stream stream) {
    (void)h; (void)members; (void)P; (void)pitch; (void)frames; (void)B; (void)c_in; (void)n_actions;
    (void)logits; (void)actions; (void)stream;
    set_error("cev_deepqn_forward: not built yet");
    return CEV_ERR_UNSUPPORTED;
}
