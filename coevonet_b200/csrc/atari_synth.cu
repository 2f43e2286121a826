// N4: a synthetic Atari-like episode loop around K2.
//
// The reference's Atari path (utils/game_logic_functions.py:48-53,84-119, Atari/deepqn.py:50-62) is dead
// code (three TypeErrors before any work) and needs ALE ROMs this build has no
// access to.  What it evidently intends is kept: two agents (first_0, second_0) alternating in AEC order over
// an 84 x 84 grayscale emulator wrapped by frame_skip_v0(4), resize_v1(84, 84), frame_stack_v1(4) and
// agent_indicator_v0, i.e. observations of 4 stacked frames + 2 agent-indicator planes (C = 6) fed to
// DeepQN.determine_action (first-maximum argmax), rewards read from env.last() after env.step().  The
// emulator itself is SYNTHETIC and deterministic:
//   frame(e, 0)          = Philox bytes of (episode e, cycle 0, joint action 0)
//   frame(e, t), t >= 1  = Philox bytes of (e, t, joint action j_t = a_first + 32 a_second of cycle t)
//   r_first(e, t)        = ((word0 of block 0xFFFF) & 0xFF - 128) / 128,  r_second = -r_first  (zero sum)
// so every frame depends on the actions taken and an episode is a real feedback loop around the forward.
// Two kernels: the emulator step (next frame into a 4-slot ring + rewards) and the observation builder
// (frame_stack_v1 order, oldest first, zeros before the first frames; agent_indicator planes of 255 / 0).
#include "common.cuh"

namespace cev {

constexpr int AT_FRAME = 84 * 84;              // 7056 bytes = 441 Philox blocks
constexpr int AT_BLOCKS = AT_FRAME / 16;

__global__ void __launch_bounds__(256) atari_step_kernel(uint32_t k0, uint32_t k1, int64_t ep0, int64_t n, int t,
                                                         const int32_t* __restrict__ a_first,
                                                         const int32_t* __restrict__ a_second,
                                                         uint8_t* __restrict__ ring, float* __restrict__ r_first) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * (AT_BLOCKS + 1)) return;
    const int64_t ep = i / (AT_BLOCKS + 1);
    const int b = (int)(i % (AT_BLOCKS + 1));
    const uint32_t j = t == 0 ? 0u : (uint32_t)a_first[ep] + 32u * (uint32_t)a_second[ep];
    const uint32_t tag = noise_tag(CEV_KIND_FRAMES, 1);
    if (b == AT_BLOCKS) {                       // the reward block
        if (t > 0 && r_first) {
            const U4 w = philox4x32_10(U4{0xFFFFu | (j << 16), (uint32_t)(ep0 + ep), (uint32_t)t, tag}, k0, k1);
            r_first[ep] = (float)((int)(w.x & 0xFFu) - 128) * (1.0f / 128.0f);
        }
        return;
    }
    const U4 w = philox4x32_10(U4{(uint32_t)b | (j << 16), (uint32_t)(ep0 + ep), (uint32_t)t, tag}, k0, k1);
    uint4* dst = reinterpret_cast<uint4*>(ring + ((size_t)ep * 4 + (size_t)(t & 3)) * AT_FRAME) + b;
    *dst = make_uint4(w.x, w.y, w.z, w.w);
}

// obs [n][6][84 x 84] u8 for the agent `seat` (0 = first_0, 1 = second_0) after `t` emulator steps: planes 0..3 =
// frames t-3 .. t (zeros where the index is negative), plane 4 + s = 255 if s == seat else 0.
__global__ void __launch_bounds__(256) atari_observe_kernel(const uint8_t* __restrict__ ring, int64_t n, int t, int seat,
                                                            uint8_t* __restrict__ obs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 6 * AT_BLOCKS) return;
    const int64_t ep = i / (6 * AT_BLOCKS);
    const int pl = (int)((i / AT_BLOCKS) % 6), b = (int)(i % AT_BLOCKS);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (pl < 4) {
        const int ft = t - 3 + pl;
        if (ft >= 0) v = reinterpret_cast<const uint4*>(ring + ((size_t)ep * 4 + (size_t)(ft & 3)) * AT_FRAME)[b];
    } else if (pl - 4 == seat) {
        v = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    }
    reinterpret_cast<uint4*>(obs + ((size_t)ep * 6 + pl) * AT_FRAME)[b] = v;
}

}  // namespace cev

using namespace cev;

extern "C" {

int cev_atari_synth_step_u8(cev_handle* h, uint64_t seed, int64_t ep0, int64_t n, int t, const int32_t* a_first,
                            const int32_t* a_second, uint8_t* ring, float* r_first, cev_stream stream) {
    CEV_REQUIRE(h && ring, "atari_synth_step: null pointer");
    CEV_REQUIRE(n >= 0 && t >= 0 && ep0 >= 0 && ep0 + n <= 0xFFFFFFFFll, "atari_synth_step: bad range");
    CEV_REQUIRE(t == 0 || (a_first && a_second), "atari_synth_step: actions are needed for t >= 1");
    CEV_REQUIRE((reinterpret_cast<uintptr_t>(ring) & 15u) == 0, "atari_synth_step: ring must be 16B aligned");
    if (n == 0) return CEV_OK;
    CEV_GUARD(h);
    const int64_t total = n * (AT_BLOCKS + 1);
    atari_step_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (uint32_t)(seed & 0xFFFFFFFFull), (uint32_t)(seed >> 32), ep0, n, t, a_first, a_second, ring, r_first);
    return check_cuda(cudaGetLastError(), "atari_step_kernel");
}

int cev_atari_observe_u8(cev_handle* h, const uint8_t* ring, int64_t n, int t, int seat, uint8_t* obs,
                         cev_stream stream) {
    CEV_REQUIRE(h && ring && obs, "atari_observe: null pointer");
    CEV_REQUIRE(n >= 0 && t >= 0 && (seat == 0 || seat == 1), "atari_observe: bad arguments");
    CEV_REQUIRE((reinterpret_cast<uintptr_t>(ring) & 15u) == 0 && (reinterpret_cast<uintptr_t>(obs) & 15u) == 0,
                "atari_observe: 16B alignment");
    if (n == 0) return CEV_OK;
    CEV_GUARD(h);
    const int64_t total = n * 6 * AT_BLOCKS;
    atari_observe_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ring, n, t, seat, obs);
    return check_cuda(cudaGetLastError(), "atari_observe_kernel");
}

}  // extern "C"
