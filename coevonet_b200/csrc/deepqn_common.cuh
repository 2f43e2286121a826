// DeepQN row geometry shared by the K2 kernels (Atari/deepqn.py:7-36).
#pragma once

#include "common.cuh"

namespace cev {

constexpr float BN_EPS = 1e-5f;

struct DqnOffsets {
    int c1w, c1b, c2w, c2b, c3w, c3b, f1w, f1b, ow, ob, bn1g, bn1b, bn2g, bn2b, bn3g, bn3b, total;
};

__host__ __device__ inline DqnOffsets dqn_offsets(int c_in, int n_act) {
    DqnOffsets o;
    o.c1w = 0;
    o.c1b = o.c1w + 32 * c_in * 64;
    o.c2w = o.c1b + 32;
    o.c2b = o.c2w + 64 * 32 * 16;
    o.c3w = o.c2b + 64;
    o.c3b = o.c3w + 64 * 64 * 9;
    o.f1w = o.c3b + 64;
    o.f1b = o.f1w + 512 * 3136;
    o.ow = o.f1b + 512;
    o.ob = o.ow + n_act * 512;
    o.bn1g = o.ob + n_act;
    o.bn1b = o.bn1g + 32;
    o.bn2g = o.bn1b + 32;
    o.bn2b = o.bn2g + 64;
    o.bn3g = o.bn2b + 64;
    o.bn3b = o.bn3g + 64;
    o.total = o.bn3b + 64;
    return o;
}

// conv stack on the tensor cores (deepqn_conv_tc.cu): frames u8 [n_frames][C][84][84] -> act3 fp32
// [n_frames][3136] (post BatchNorm-3 + ReLU, flattened like torch's reshape); y1/y2/y3 are the pre-BatchNorm
// conv outputs [n_frames][positions][channels] (scratch).
int launch_deepqn_conv_tc(cev_handle* h, const float* members, int64_t pitch, int P, int B, int c_in, int n_act,
                          const uint8_t* frames, float* y1, float* y2, float* y3, float* act3, cudaStream_t stream);

}  // namespace cev
