// K1 (lockstep form): the population evaluation as one pass per world step over ALL episodes.
//
// Replaces, for P members x K opponent sets x E env instances (same contract as
// rollout_cluster.cu):
//   play_game / play_MPE          utils/game_logic_functions.py:123-228
//   FCNetwork.forward / argmax    MPE/fcnetwork.py:37-90
//   simple_adversary_v3 world     SURVEY.md Appendix A (third-party pettingzoo)
//   the GA/ES evaluation loops    genetic_algorithm.py:125-217, evolutionary_strategy.py:236-251
//
// Why a second form: two of the three forwards of every world step use the OPPONENTS'
// weights, which are shared by every episode of the launch (the ES base agents,
// evolutionary_strategy.py:77/94/111; the GA Hall-of-Fame rows, genetic_algorithm.py:138-139).
// With all episodes advanced in lockstep that work is a dense [episodes x 512] . [512 x 256]
// GEMM and belongs on the tensor cores; only the member's own forward (unique weights per
// member) is a matrix that nobody else reads: its kernel is bound by the HBM stream of the member rows.  Per world step:
//
//   ls_member_tc_kernel (ls_member_tc.cuh) persistent, job = (member, 16 episodes): the member's fc2 matrix as the
//                     M-side operand of tcgen05 MMAs (3xTF32, both split parts in tensor memory), streamed once by
//                     TMA; layer 1 fused through closed-form LayerNorm statistics; LayerNorm-2 + output layer +
//                     first-max argmax in the epilogue warps.  Runs on a share of the SMs BESIDE ls_opp_kernel.
//   ls_opp_kernel     persistent, one CTA per SM, job = (opponent seat, opponent set, 128 episodes):
//                     8 producer warps (warp = 16-byte k chunk, lane = 4 episode rows) compute layer 1 +
//                     LayerNorm + ReLU and write the activations, scaled by a power of two and split into FP16
//                     hi + lo parts, straight into the 128B-swizzled K-major A tiles; one thread streams the
//                     pre-split (scaled FP16 hi + lo) opponent fc2 matrix with TMA; one warp issues
//                     tcgen05.mma kind::f16 (M=128, N=256, K=16) three times per k-step (lo.hi + hi.lo + hi.hi:
//                     22-23 significand bits per operand, fp32-level accuracy, at twice the TF32 rate)
//                     into a double-buffered TMEM accumulator; 4 epilogue warps read it back with
//                     tcgen05.ld and do bias + LayerNorm-2 + ReLU + output layer + argmax per row.
//                     Layer-1 LayerNorm statistics come from the closed form mean = wbar.x + bbar,
//                     var = z^T C z (z = [x; 1], C = row covariance of [W1 | b1], fp64), so a
//                     producer never needs the whole 512-vector at once.
//   ls_env_step_kernel one thread per episode: fp64 physics, rewards, next observations
//                     (bit-exact with oracle/mpe_env.py given equal actions).
//
// Arithmetic: member fc2 3xTF32, opponent fc2 the scaled two-term FP16 split, both with fp32 accumulation; layer 1 /
// LayerNorm / output fp32 (error ~1e-6 of the activations' scale, the same order as fp32 summation-order noise);
// environment fp64.
#include <stdlib.h>

#include <cuda_fp16.h>

#include "rollout_common.cuh"
#include "tc_common.cuh"

namespace cev {

// ---------------------------------------------------------------------------------------------
// workspace layout
// ---------------------------------------------------------------------------------------------
constexpr int LS_ST_FIELDS = 16;       // px[3] py[3] vx[3] vy[3] lx[2] ly[2]
constexpr int LS_OBS_PAD = 12;         // floats per (seat, episode) observation record
constexpr int LS_L1S = 132;            // doubles per opponent row: wbar[11] | C[11][11]

struct LsBuffers {
    double* st;        // [16][N]
    double* acc;       // [3][N]  sum_good, last_good, sum_adv
    float* min_gap;    // [N]
    int32_t* goal;     // [N]
    float* obs;        // [3][N][12]
    int32_t* act;      // [3][N]
    float* gap;        // [3][N]
    __half* w2split;   // [2 seats][K][hi, lo][256][512] fp16, scaled by 2^s_w
    uint32_t* wmax;    // [2 seats][K] bit pattern of max |fc2.W|
    float* oscale;     // [2 seats][K][2]: 2^s_x (activation scale), 2^-(s_w + s_x) (accumulator unscale)
    double* l1stats;   // [2 seats][K][132]
    float2* l1ms;      // [2 opponent seats][N]: LayerNorm-1 (mean, rstd) of every episode's current observation
    double* ml1stats;  // [P][132]  the members' own layer-1 statistics (tensor-core member form)
};

static size_t ls_align(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t ls_carve(void* base, int64_t N, int K, int P, LsBuffers* b) {
    size_t off = 0;
    char* p = static_cast<char*>(base);
    auto take = [&](size_t bytes) {
        void* r = p ? p + off : nullptr;
        off += ls_align(bytes);
        return r;
    };
    LsBuffers t;
    t.w2split = static_cast<__half*>(take((size_t)2 * K * 2 * H2 * H1 * sizeof(__half)));
    t.wmax = static_cast<uint32_t*>(take((size_t)2 * K * sizeof(uint32_t)));
    t.oscale = static_cast<float*>(take((size_t)2 * K * 2 * sizeof(float)));
    t.st = static_cast<double*>(take((size_t)LS_ST_FIELDS * N * sizeof(double)));
    t.acc = static_cast<double*>(take((size_t)3 * N * sizeof(double)));
    t.min_gap = static_cast<float*>(take((size_t)N * sizeof(float)));
    t.goal = static_cast<int32_t*>(take((size_t)N * sizeof(int32_t)));
    t.obs = static_cast<float*>(take((size_t)3 * N * LS_OBS_PAD * sizeof(float)));
    t.act = static_cast<int32_t*>(take((size_t)3 * N * sizeof(int32_t)));
    t.gap = static_cast<float*>(take((size_t)3 * N * sizeof(float)));
    t.l1stats = static_cast<double*>(take((size_t)2 * K * LS_L1S * sizeof(double)));
    t.l1ms = static_cast<float2*>(take((size_t)2 * N * sizeof(float2)));
    t.ml1stats = static_cast<double*>(take((size_t)P * LS_L1S * sizeof(double)));
    if (b) *b = t;
    return off;
}

// ---------------------------------------------------------------------------------------------
// per-episode state helpers (SoA in the workspace)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ls_store_state(const LsBuffers& b, int64_t N, int64_t ep, const EnvState& s) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        b.st[(0 + i) * N + ep] = s.px[i];
        b.st[(3 + i) * N + ep] = s.py[i];
        b.st[(6 + i) * N + ep] = s.vx[i];
        b.st[(9 + i) * N + ep] = s.vy[i];
    }
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        b.st[(12 + l) * N + ep] = s.lx[l];
        b.st[(14 + l) * N + ep] = s.ly[l];
    }
}
__device__ __forceinline__ void ls_load_state(const LsBuffers& b, int64_t N, int64_t ep, EnvState& s) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        s.px[i] = b.st[(0 + i) * N + ep];
        s.py[i] = b.st[(3 + i) * N + ep];
        s.vx[i] = b.st[(6 + i) * N + ep];
        s.vy[i] = b.st[(9 + i) * N + ep];
    }
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        s.lx[l] = b.st[(12 + l) * N + ep];
        s.ly[l] = b.st[(14 + l) * N + ep];
    }
    s.goal = b.goal[ep];
}
// LayerNorm-1 statistics of y = W1 x + b1 in closed form (fp64): mean = wbar . z, var = z^T C z with z = [x; 1] and
// S = wbar[11] | C[11][11] of the network (ls_l1stats_kernel).  Returns (mean, 1 / sqrt(var + eps)).
__device__ __forceinline__ float2 ls_row_stats(const double* __restrict__ S, const float* __restrict__ x, int in, bool& finite) {
    double z[11];
#pragma unroll
    for (int a = 0; a < 11; ++a) z[a] = a < in ? (double)x[a] : (a == in ? 1.0 : 0.0);
    double md = 0.0, vd = 0.0;
#pragma unroll
    for (int a = 0; a < 11; ++a) {
        md = fma(__ldg(S + a), z[a], md);
        double ra = 0.0;
#pragma unroll
        for (int b = 0; b < 11; ++b) ra = fma(__ldg(S + 11 + a * 11 + b), z[b], ra);
        vd = fma(ra, z[a], vd);
    }
    const float m1 = (float)md, var = (float)vd;
    finite = isfinite(m1) && isfinite(var);
    return make_float2(m1, 1.0f / sqrtf(var + LN_EPS));
}

// Observations of the three seats, and for the two OPPONENT seats the LayerNorm-1 statistics of the opponent
// network on that observation (one thread per episode here instead of a serial phase in front of every job of the
// opponent kernel, where four of its eight producer warps computed them while the other four waited).
__device__ __forceinline__ bool ls_store_obs(const LsBuffers& b, int64_t N, int64_t ep, const EnvState& s, int K, int E,
                                             const int (&opp_seat)[2]) {
    bool ok = true;
#pragma unroll
    for (int seat = 0; seat < 3; ++seat) {
        float o[LS_OBS_PAD];
#pragma unroll
        for (int j = 0; j < LS_OBS_PAD; ++j) o[j] = 0.f;
        env_observe(s, seat, o);
        float4* dst = reinterpret_cast<float4*>(b.obs + ((int64_t)seat * N + ep) * LS_OBS_PAD);
        dst[0] = make_float4(o[0], o[1], o[2], o[3]);
        dst[1] = make_float4(o[4], o[5], o[6], o[7]);
        dst[2] = make_float4(o[8], o[9], o[10], o[11]);
#pragma unroll
        for (int oi = 0; oi < 2; ++oi)
            if (seat == opp_seat[oi]) {
                const int k = (int)((ep / E) % K);
                bool fin;
                b.l1ms[(int64_t)oi * N + ep] = ls_row_stats(b.l1stats + (size_t)(oi * K + k) * LS_L1S, o, seat_in_dim(seat), fin);
                ok = ok && fin;
            }
    }
    return ok;
}

struct LsEnvParams {
    LsBuffers b;
    int64_t N;
    int K, E, init_shared, pos_first, last;
    const double* init;
    double* out;
    // parity instrumentation (cev_mpe_rollout_trace_f32; null in production): replay these actions
    // instead of the networks' (teacher forcing), and record the networks' own decisions
    const int32_t* forced;   // this cycle's [3][N]
    int32_t* act_out;        // this cycle's [3][N]
    int opp_seat[2];         // the two opponent seats (ascending)
    int32_t* status;
};

__global__ void __launch_bounds__(256) ls_init_kernel(const LsEnvParams p) {
    const int64_t ep = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ep >= p.N) return;
    const int e = (int)(ep % p.E), k = (int)((ep / p.E) % p.K);
    const int64_t rec = p.init_shared ? ((int64_t)k * p.E + e) : ep;
    EnvState s;
    env_load(s, p.init + rec * CEV_INIT_STATE_DIM);
    ls_store_state(p.b, p.N, ep, s);
    p.b.goal[ep] = s.goal;
    p.b.acc[ep] = 0.0;
    p.b.acc[p.N + ep] = 0.0;
    p.b.acc[2 * p.N + ep] = 0.0;
    p.b.min_gap[ep] = CUDART_INF_F;
    if (!ls_store_obs(p.b, p.N, ep, s, p.K, p.E, p.opp_seat) && p.status) atomicOr(p.status, CEV_STATUS_NONFINITE);
    if (p.last) {   // n_cycles == 0: nothing is played
        double* o = p.out + ep * CEV_ROLLOUT_OUT_DIM;
        o[0] = 0.0;
        o[1] = 0.0;
        o[2] = 0.0;
        o[3] = (double)CUDART_INF_F;
    }
}

__global__ void __launch_bounds__(256) ls_env_step_kernel(const LsEnvParams p) {
    const int64_t ep = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ep >= p.N) return;
    EnvState s;
    ls_load_state(p.b, p.N, ep, s);
    int act[3] = {p.b.act[ep], p.b.act[p.N + ep], p.b.act[2 * p.N + ep]};
    if (p.act_out) {
        p.act_out[ep] = act[0];
        p.act_out[p.N + ep] = act[1];
        p.act_out[2 * p.N + ep] = act[2];
    }
    if (p.forced) {
        act[0] = p.forced[ep];
        act[1] = p.forced[p.N + ep];
        act[2] = p.forced[2 * p.N + ep];
    }
    const float g = fminf(p.b.gap[ep], fminf(p.b.gap[p.N + ep], p.b.gap[2 * p.N + ep]));
    double rg, ra;
    env_step(s, act, p.pos_first != 0, rg, ra);
    const double sum_good = __dadd_rn(p.b.acc[ep], rg);
    const double sum_adv = __dadd_rn(p.b.acc[2 * p.N + ep], ra);
    const float min_gap = fminf(p.b.min_gap[ep], g);
    if (p.last) {
        double* o = p.out + ep * CEV_ROLLOUT_OUT_DIM;
        o[0] = sum_good;
        o[1] = rg;
        o[2] = sum_adv;
        o[3] = (double)min_gap;
        return;
    }
    p.b.acc[ep] = sum_good;
    p.b.acc[p.N + ep] = rg;
    p.b.acc[2 * p.N + ep] = sum_adv;
    p.b.min_gap[ep] = min_gap;
    ls_store_state(p.b, p.N, ep, s);
    if (!ls_store_obs(p.b, p.N, ep, s, p.K, p.E, p.opp_seat) && p.status) atomicOr(p.status, CEV_STATUS_NONFINITE);
}

// ---------------------------------------------------------------------------------------------
// opponent preparation (once per launch): scaled FP16 hi/lo split of fc2.W, layer-1 row statistics
//
// The opponents' fc2 runs as THREE kind::f16 MMAs per product, x1.w1 + x2.w1 + x1.w2 with v1 = fp16(v'),
// v2 = fp16(v' - v1) of the power-of-two scaled operands v' = 2^s v ("2xFP16": 22-23 significand bits per
// operand, the same as the TF32 hi/lo split of rounds 1-2, at twice the tensor rate and half the operand
// bytes).  FP16 has five exponent bits, so the scales keep both parts in the normal range:
//   s_w = 14 - ilogb(max |fc2.W|)            (|w'| < 2^15; the residuals sit 2^-12 below, still normal)
//   s_x = 13 - ilogb(22.63 max|ln1.g| + max|ln1.b|)   (a LayerNorm output over 512 values is at most sqrt(511))
// and the epilogue multiplies the fp32 accumulator by 2^-(s_w + s_x).  Values that would be subnormal after
// scaling lose RELATIVE precision only where they are 2^-24 of the largest weight: an absolute error of
// 2^-25 2^-s per element, far below fp32 summation noise of the dot product.
// ---------------------------------------------------------------------------------------------
struct LsPrepParams {
    const float* opp[2];
    int64_t opp_pitch[2];
    int seat[2];
    int K;
    __half* w2split;
    uint32_t* wmax;
    float* oscale;
    double* l1stats;
};

__device__ __forceinline__ int ls_scale_exp(float bound, int target) {
    // 2^s * bound < 2^(target + 1); clamped so that 2^s and its inverse stay far from fp32's range limits
    if (!(bound > 0.f) || !isfinite(bound)) return 0;
    const int s = target - ilogbf(bound);
    return s < -40 ? -40 : (s > 40 ? 40 : s);
}

__global__ void __launch_bounds__(256) ls_wmax_kernel(const LsPrepParams p) {
    const int oi = blockIdx.y / p.K, k = blockIdx.y % p.K;
    const FcOffsets o = fc_offsets(seat_in_dim(p.seat[oi]));
    const float4* src = reinterpret_cast<const float4*>(p.opp[oi] + (int64_t)k * p.opp_pitch[oi] + o.fc2w);
    float m = 0.f;
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < H2 * H1 / 4; f += gridDim.x * blockDim.x) {
        const float4 w = src[f];
        m = fmaxf(fmaxf(m, fmaxf(fabsf(w.x), fabsf(w.y))), fmaxf(fabsf(w.z), fabsf(w.w)));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) atomicMax(p.wmax + blockIdx.y, __float_as_uint(m));      // m >= 0: bit order = value order
}

__global__ void __launch_bounds__(256) ls_split_w2_kernel(const LsPrepParams p) {
    const int oi = blockIdx.y / p.K, k = blockIdx.y % p.K;
    const FcOffsets o = fc_offsets(seat_in_dim(p.seat[oi]));
    const float4* src = reinterpret_cast<const float4*>(p.opp[oi] + (int64_t)k * p.opp_pitch[oi] + o.fc2w);
    const float sw = ldexpf(1.0f, ls_scale_exp(__uint_as_float(p.wmax[blockIdx.y]), 14));
    uint2* hi = reinterpret_cast<uint2*>(p.w2split + (size_t)(blockIdx.y * 2 + 0) * H2 * H1);
    uint2* lo = reinterpret_cast<uint2*>(p.w2split + (size_t)(blockIdx.y * 2 + 1) * H2 * H1);
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < H2 * H1 / 4; f += gridDim.x * blockDim.x) {
        const float4 w = src[f];
        const float v[4] = {w.x * sw, w.y * sw, w.z * sw, w.w * sw};
        __half h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            h[i] = __float2half_rn(v[i]);
            l[i] = __float2half_rn(v[i] - __half2float(h[i]));
        }
        const __half2 h01 = __halves2half2(h[0], h[1]), h23 = __halves2half2(h[2], h[3]);
        const __half2 l01 = __halves2half2(l[0], l[1]), l23 = __halves2half2(l[2], l[3]);
        hi[f] = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
        lo[f] = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
    }
}

// wbar[j] = mean over the 512 rows of column j of [W1 | b1]; C[a][b] = row covariance (biased).
// LayerNorm-1 of y = W1 x + b1:  mean(y) = wbar . z,  var(y) = z^T C z  with z = [x; 1]
// (MPE/fcnetwork.py:44: nn.LayerNorm(512), biased variance).
__device__ __forceinline__ void ls_l1stats_row(const float* __restrict__ row, int in, double* __restrict__ out) {
    const int d = in + 1;
    const float* w1 = row;
    const float* b1 = row + H1 * in;
    __shared__ float cols[11][H1 + 1];      // [W1 | b1] transposed: column a of all 512 rows (+1: conflict-free transpose)
    __shared__ double wbar[11];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < H1 * in; i += 512) cols[i % in][i / in] = w1[i];
    cols[in][t] = b1[t];
    __syncthreads();
    for (int a = warp; a < 11; a += 16) {
        double s = 0.0;
        if (a < d)
            for (int r = lane; r < H1; r += 32) s += (double)cols[a][r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) wbar[a] = s / H1;
    }
    __syncthreads();
    if (t < 11) out[t] = wbar[t];
    // C is symmetric: the 66 pairs a <= b are summed (in the row order r = lane, lane + 32, ... of the full form, so
    // C[a][b] keeps its bits) and mirrored
    for (int pr = warp; pr < 66; pr += 16) {
        int a = 0, rem = pr;
        while (rem >= 11 - a) {
            rem -= 11 - a;
            ++a;
        }
        const int b = a + rem;
        double s = 0.0;
        if (a < d && b < d) {
            const double ma = wbar[a], mb = wbar[b];
            for (int r = lane; r < H1; r += 32) s += ((double)cols[a][r] - ma) * ((double)cols[b][r] - mb);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            out[11 + a * 11 + b] = s / H1;
            out[11 + b * 11 + a] = s / H1;
        }
    }
}

__global__ void __launch_bounds__(512) ls_l1stats_kernel(const LsPrepParams p) {
    const int oi = blockIdx.x / p.K, k = blockIdx.x % p.K;
    const int in = seat_in_dim(p.seat[oi]);
    const float* row = p.opp[oi] + (int64_t)k * p.opp_pitch[oi];
    ls_l1stats_row(row, in, p.l1stats + (size_t)blockIdx.x * LS_L1S);
    // FP16 operand scales of this opponent (see the note above ls_wmax_kernel)
    __shared__ float red[2][16];
    const int t = threadIdx.x;
    float g = fabsf(row[H1 * in + H1 + t]), be = fabsf(row[H1 * in + 2 * H1 + t]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        g = fmaxf(g, __shfl_xor_sync(0xffffffffu, g, off));
        be = fmaxf(be, __shfl_xor_sync(0xffffffffu, be, off));
    }
    if ((t & 31) == 0) {
        red[0][t >> 5] = g;
        red[1][t >> 5] = be;
    }
    __syncthreads();
    if (t == 0) {
        for (int w = 1; w < 16; ++w) {
            g = fmaxf(g, red[0][w]);
            be = fmaxf(be, red[1][w]);
        }
        const int sx = ls_scale_exp(22.63f * g + be, 13);
        const int sw = ls_scale_exp(__uint_as_float(p.wmax[blockIdx.x]), 14);
        p.oscale[blockIdx.x * 2 + 0] = ldexpf(1.0f, sx);
        p.oscale[blockIdx.x * 2 + 1] = ldexpf(1.0f, -(sx + sw));
    }
}

// the same statistics for every member row (tensor-core member form), once per rollout
__global__ void __launch_bounds__(512) ls_member_l1stats_kernel(const float* __restrict__ members, int64_t pitch, int in,
                                                                double* __restrict__ out) {
    ls_l1stats_row(members + (int64_t)blockIdx.x * pitch, in, out + (size_t)blockIdx.x * LS_L1S);
}

// ---------------------------------------------------------------------------------------------
// shared geometry of the member / opponent kernels
// ---------------------------------------------------------------------------------------------
#ifndef LS_GRID_OPP_DEFAULT
#define LS_GRID_OPP_DEFAULT 0                     // 0 = the library's split of the SMs (ls_split_sms)
#endif
#ifndef LS_GRID_MEM_DEFAULT
#define LS_GRID_MEM_DEFAULT 0
#endif
constexpr int LS_BT = 16;                         // episodes per member job
constexpr int LS_TAIL_FLOATS = 2056;              // fc2.b | ln2.g | ln2.b | out.W | out.b (+3 pad), contiguous in the row
constexpr int LS_W1A_FLOATS = H1 * IN_GOOD + 3 * H1;

// ---------------------------------------------------------------------------------------------
// opponent forward (tcgen05, scaled two-term FP16 split)
// ---------------------------------------------------------------------------------------------
constexpr int OP_THREADS = 448;                   // warps 0-3 epilogue, 4-11 A producers, 12 TMA, 13 MMA
constexpr int OP_PROD = 256;
constexpr int OP_BM = 128, OP_BN = 256, OP_BK = 64, OP_KT = H1 / OP_BK;   // 64 fp16 = one 128-byte swizzle row
constexpr uint32_t OP_A_BYTES = OP_BM * OP_BK * 2;          // 16 KB
constexpr uint32_t OP_B_BYTES = OP_BN * OP_BK * 2;          // 32 KB
constexpr uint32_t OP_STAGE_BYTES = 2 * OP_A_BYTES + 2 * OP_B_BYTES;   // A hi | A lo | B hi | B lo = 96 KB
constexpr int OP_STAGES = 2;
constexpr size_t OP_OFF_W1A = (size_t)OP_STAGES * OP_STAGE_BYTES;
constexpr size_t OP_OFF_TAIL = OP_OFF_W1A + (size_t)LS_W1A_FLOATS * 4;      // fc2.b | ln2.g | ln2.b | out.W | out.b
constexpr size_t OP_OFF_BAR = OP_OFF_TAIL + (size_t)LS_TAIL_FLOATS * 4;
constexpr size_t OP_SMEM_MAX = 232448;                                     // 227 KB opt-in limit per CTA
constexpr size_t OP_SLACK = OP_SMEM_MAX - (OP_OFF_BAR + 128);              // what is left for the 1024-byte alignment
constexpr size_t OP_SMEM = OP_SMEM_MAX;
// instruction descriptor: D = F32, A = B = F16 (format 0), both K-major, M = 128, N = 256
constexpr uint32_t OP_IDESC = (1u << 4) | ((uint32_t)(OP_BN >> 3) << 17) | ((uint32_t)(OP_BM >> 4) << 24);

struct LsOppParams {
    const float* opp[2];
    int64_t opp_pitch[2];
    int seat[2];
    int K, E, n_tiles, n_jobs;
    int64_t PE;              // P * E rows per (seat, k)
    int64_t N;
    const float* obs;        // [3][N][12]
    int32_t* act;            // [3][N]
    float* gap;
    const float2* l1ms;      // [2 opponent seats][N] LayerNorm-1 (mean, rstd) per episode (ls_store_obs)
    const float* oscale;     // [2 seats][K][2]: activation scale, accumulator unscale
    int32_t* status;
    float* logits;           // this cycle's [3][N][5] (parity instrumentation; null in production)
};

// One producer thread's share of an A tile: 16-byte chunk c (8 fp16 activations) of k-tile kt for its FOUR
// episode rows lane + 32 j: relu(LN1(W1 x + b1)) scaled by 2^s_x, split into FP16 hi + lo and stored K-major with
// the 128-byte swizzle the UMMA descriptor expects (chunk c of row r at c ^ (r & 7)).  The weight loads are
// warp-uniform (producer warp = chunk) and serve four rows each.  Rows of a quarter warp differ in (r & 7), so
// the tile stores are conflict-free.
template <int IN>
__device__ __forceinline__ void op_produce_chunk(const float* __restrict__ w1a, const float (&x)[4][IN_GOOD],
                                                 const float (&mean)[4], const float (&rstd)[4], float sx, int kt, int c,
                                                 int lane, unsigned char* __restrict__ a_hi,
                                                 unsigned char* __restrict__ a_lo) {
    const float* fc1b = w1a + H1 * IN;
    const float* ln1g = fc1b + H1;
    const float* ln1b = ln1g + H1;
    const int k0 = kt * OP_BK + c * 8;
    uint32_t hi[4][4], lo[4][4];      // [row j][k pair]
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const float4 bb = *reinterpret_cast<const float4*>(fc1b + k0 + 4 * half);
        const float4 gg = *reinterpret_cast<const float4*>(ln1g + k0 + 4 * half);
        const float4 ee = *reinterpret_cast<const float4*>(ln1b + k0 + 4 * half);
        // gamma and beta carry the activation scale 2^s_x: relu(fma(t, g, e)) * 2^s = relu(fma(t, g 2^s, e 2^s)) exactly
        const float bq[4] = {bb.x, bb.y, bb.z, bb.w}, gq[4] = {gg.x * sx, gg.y * sx, gg.z * sx, gg.w * sx},
                    eq[4] = {ee.x * sx, ee.y * sx, ee.z * sx, ee.w * sx};
        float hv[4][4];               // [row j][k]
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
            float w[IN];
            const float2* wr = reinterpret_cast<const float2*>(w1a + (k0 + 4 * half + qq) * IN);
#pragma unroll
            for (int i2 = 0; i2 < IN / 2; ++i2) {
                const float2 v = wr[i2];
                w[2 * i2] = v.x;
                w[2 * i2 + 1] = v.y;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float pre = 0.f;
#pragma unroll
                for (int i = 0; i < IN; ++i) pre = fmaf(w[i], x[j][i], pre);
                pre += bq[qq];
                hv[j][qq] = fmaxf(fmaf((pre - mean[j]) * rstd[j], gq[qq], eq[qq]), 0.f);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int pr = 0; pr < 2; ++pr) {
                // hi = fp16(h'), lo = fp16(h' - hi): the residual is exact in fp32
                const __half2 h2 = __floats2half2_rn(hv[j][2 * pr], hv[j][2 * pr + 1]);
                const float2 hf = __half22float2(h2);
                const __half2 l2 = __floats2half2_rn(hv[j][2 * pr] - hf.x, hv[j][2 * pr + 1] - hf.y);
                hi[j][2 * half + pr] = *reinterpret_cast<const uint32_t*>(&h2);
                lo[j][2 * half + pr] = *reinterpret_cast<const uint32_t*>(&l2);
            }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int r = lane + 32 * j;
        const int off = r * 128 + ((c ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[j][0], hi[j][1], hi[j][2], hi[j][3]);
        *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lo[j][0], lo[j][1], lo[j][2], lo[j][3]);
    }
}

// stage geometry: A hi | A lo | B hi | B lo
struct Geo {
    static constexpr uint32_t B_BYTES = OP_B_BYTES;
    static constexpr uint32_t STAGE_BYTES = OP_STAGE_BYTES;
    static constexpr int STAGES = OP_STAGES;
    static constexpr int ROWS_PER_JOB = OP_BM;
    static constexpr uint32_t N_FULL = OP_PROD / 32 + 1;          // producer warps + the TMA thread
    static constexpr uint32_t N_TEMPTY = 4;
};

__global__ void __launch_bounds__(OP_THREADS, 1)
ls_opp_kernel(const __grid_constant__ CUtensorMap map_b, const LsOppParams p) {
    const int job0 = (int)blockIdx.x;
    const int job_stride = (int)gridDim.x;
    extern __shared__ unsigned char op_raw[];
    const uint32_t pad = (1024u - (smem_u32(op_raw) & 1023u)) & 1023u;
    if (pad > OP_SLACK) __trap();          // dynamic shared memory starts 1 KB aligned on sm_100; checked, not assumed
    unsigned char* base = op_raw + pad;    // offset into the __shared__ array: keeps the shared address space
    unsigned char* stage_mem = base;
    float* w1a = reinterpret_cast<float*>(base + OP_OFF_W1A);
    float* tail = reinterpret_cast<float*>(base + OP_OFF_TAIL);
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(base + OP_OFF_BAR);   // [2] A written + B landed
    uint64_t* bar_empty = bar_full + Geo::STAGES;                           // [STAGES] MMAs retired
    uint64_t* bar_tfull = bar_empty + Geo::STAGES;                          // [2] accumulator ready
    uint64_t* bar_tempty = bar_tfull + 2;                                   // [2] accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 2);
    int* flag = reinterpret_cast<int*>(tmem_slot + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        *flag = 0;
        for (int i = 0; i < Geo::STAGES; ++i) {
            tc_mbar_init(bar_full + i, Geo::N_FULL);
            tc_mbar_init(bar_empty + i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc_mbar_init(bar_tfull + i, 1);
            tc_mbar_init(bar_tempty + i, Geo::N_TEMPTY);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 13) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                         tc_smem_u32(tmem_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int jobs_per_ok = p.n_tiles;

    if (warp < 4) {
        // ===================== epilogue: bias + LayerNorm-2 + ReLU + output layer + argmax =====
        const int q = warp;
        uint32_t pass = 0;
        int cur_ok = -1;
        const float* b2 = tail;
        const float* g2 = tail + H2;
        const float* be2 = tail + 2 * H2;
        const float* w3 = tail + 3 * H2;
        const float* b3 = tail + 3 * H2 + NACT * H2;
        for (int job = job0; job < p.n_jobs; job += job_stride, ++pass) {
            const int okey = job / jobs_per_ok, tile = job % jobs_per_ok;
            const int oi = okey / p.K, k = okey % p.K;
            const int seat = oi ? p.seat[1] : p.seat[0];
            if (okey != cur_ok) {
                // this opponent's epilogue parameters -> shared memory (only the epilogue warps touch them)
                const float* row = (oi ? p.opp[1] : p.opp[0]) + (int64_t)k * (oi ? p.opp_pitch[1] : p.opp_pitch[0]);
                const float4* src = reinterpret_cast<const float4*>(row + fc_offsets(seat_in_dim(seat)).fc2b);
                asm volatile("bar.sync 2, 128;\n" ::: "memory");
                for (int f = threadIdx.x; f < LS_TAIL_FLOATS / 4; f += 128)
                    reinterpret_cast<float4*>(tail)[f] = __ldg(src + f);
                asm volatile("bar.sync 2, 128;\n" ::: "memory");
                cur_ok = okey;
            }
            const float unscale = __ldg(p.oscale + okey * 2 + 1);
            const int64_t j = (int64_t)tile * Geo::ROWS_PER_JOB + q * 32 + lane;
            const bool valid = j < p.PE;
            const int64_t jj = valid ? j : p.PE - 1;
            const int64_t ep = ((jj / p.E) * p.K + k) * p.E + (jj % p.E);
            const uint32_t as = pass & 1, ause = pass >> 1;
            tc_mbar_wait(bar_tfull + as, ause & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * OP_BN;
            float sum = 0.f;
#pragma unroll 1
            for (int c0 = 0; c0 < OP_BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 bb = *reinterpret_cast<const float4*>(b2 + c0 + i);
                    sum += fmaf(__uint_as_float(v[i]), unscale, bb.x);          // 2^-(s_w + s_x): exact
                    sum += fmaf(__uint_as_float(v[i + 1]), unscale, bb.y);
                    sum += fmaf(__uint_as_float(v[i + 2]), unscale, bb.z);
                    sum += fmaf(__uint_as_float(v[i + 3]), unscale, bb.w);
                }
            }
            const float mean = sum * (1.0f / H2);
            float sq = 0.f;
#pragma unroll 1
            for (int c0 = 0; c0 < OP_BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 bb = *reinterpret_cast<const float4*>(b2 + c0 + i);
                    const float bq[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float d = fmaf(__uint_as_float(v[i + u]), unscale, bq[u]) - mean;
                        sq = fmaf(d, d, sq);
                    }
                }
            }
            const float var = sq * (1.0f / H2);
            const float rstd = 1.0f / sqrtf(var + LN_EPS);
            float lg[NACT] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
            for (int c0 = 0; c0 < OP_BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const int c = c0 + i;
                    const float4 bb = *reinterpret_cast<const float4*>(b2 + c);
                    const float4 gg = *reinterpret_cast<const float4*>(g2 + c);
                    const float4 ee = *reinterpret_cast<const float4*>(be2 + c);
                    const float bq[4] = {bb.x, bb.y, bb.z, bb.w}, gq[4] = {gg.x, gg.y, gg.z, gg.w},
                                eq[4] = {ee.x, ee.y, ee.z, ee.w};
                    float h[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float x = fmaf(__uint_as_float(v[i + u]), unscale, bq[u]) - mean;
                        h[u] = fmaxf(fmaf(x * rstd, gq[u], eq[u]), 0.f);
                    }
#pragma unroll
                    for (int a = 0; a < NACT; ++a) {
                        const float4 ww = *reinterpret_cast<const float4*>(w3 + a * H2 + c);
                        lg[a] = fmaf(ww.x, h[0], lg[a]);
                        lg[a] = fmaf(ww.y, h[1], lg[a]);
                        lg[a] = fmaf(ww.z, h[2], lg[a]);
                        lg[a] = fmaf(ww.w, h[3], lg[a]);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(bar_tempty + as);
            bool fin = isfinite(mean) && isfinite(var);
#pragma unroll
            for (int a = 0; a < NACT; ++a) {
                lg[a] += b3[a];
                fin = fin && isfinite(lg[a]);
            }
            float gap;
            const int a = argmax_first5(lg, gap);
            if (valid) {
                if (!fin) *flag = 1;
                p.act[(int64_t)seat * p.N + ep] = a;
                p.gap[(int64_t)seat * p.N + ep] = gap;
                if (p.logits) {
#pragma unroll
                    for (int u = 0; u < NACT; ++u) p.logits[((int64_t)seat * p.N + ep) * NACT + u] = lg[u];
                }
            }
        }
    } else if (warp < 12) {
        // ===================== A producers: layer 1 + LayerNorm + ReLU -> FP16 hi/lo tiles =====
        const int pt = threadIdx.x - 128;
        const int c = pt >> 5;                        // producer warp = 16-byte chunk (8 fp16 k) of every k-tile
        int cur_ok = -1;
        uint32_t it = 0;
        for (int job = job0; job < p.n_jobs; job += job_stride) {
            const int okey = job / jobs_per_ok, tile = job % jobs_per_ok;
            const int oi = okey / p.K, k = okey % p.K;
            const int seat = oi ? p.seat[1] : p.seat[0];
            const int in = seat_in_dim(seat);
            const float* row = (oi ? p.opp[1] : p.opp[0]) + (int64_t)k * (oi ? p.opp_pitch[1] : p.opp_pitch[0]);
            if (okey != cur_ok) {
                asm volatile("bar.sync 1, 256;\n" ::: "memory");       // producers done with the old block
                const int n4 = (H1 * in + 3 * H1) / 4;
                for (int f = pt; f < n4; f += OP_PROD)
                    reinterpret_cast<float4*>(w1a)[f] = __ldg(reinterpret_cast<const float4*>(row) + f);
                asm volatile("bar.sync 1, 256;\n" ::: "memory");
                cur_ok = okey;
            }
            const float sx = __ldg(p.oscale + okey * 2);               // 2^s_x of this opponent
            // observations of this thread's four episode rows lane + 32 j
            float x[4][IN_GOOD];
            int64_t epj[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t jrow = (int64_t)tile * Geo::ROWS_PER_JOB + lane + 32 * j;
                const int64_t jj = jrow < p.PE ? jrow : p.PE - 1;
                const int64_t ep = ((jj / p.E) * p.K + k) * p.E + (jj % p.E);
                epj[j] = ep;
                const float4* src = reinterpret_cast<const float4*>(p.obs + ((int64_t)seat * p.N + ep) * LS_OBS_PAD);
                const float4 v0 = __ldg(src), v1 = __ldg(src + 1), v2 = __ldg(src + 2);
                x[j][0] = v0.x; x[j][1] = v0.y; x[j][2] = v0.z; x[j][3] = v0.w;
                x[j][4] = v1.x; x[j][5] = v1.y; x[j][6] = v1.z; x[j][7] = v1.w;
                x[j][8] = v2.x; x[j][9] = v2.y;
            }
            float mean[4], rstd[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 v = __ldg(p.l1ms + (int64_t)oi * p.N + epj[j]);
                mean[j] = v.x;
                rstd[j] = v.y;
            }
            for (int kt = 0; kt < OP_KT; ++kt, ++it) {
                const uint32_t st = it % Geo::STAGES, use = it / Geo::STAGES;
                if (use > 0) tc_mbar_wait(bar_empty + st, (use - 1) & 1);
                unsigned char* a_hi = stage_mem + (size_t)st * Geo::STAGE_BYTES;
                unsigned char* a_lo = a_hi + OP_A_BYTES;
#if !(defined(CEV_EXP) && (CEV_EXP & 1))      // development experiment: bit 0 = producers write nothing
                if (in == IN_GOOD) {
                    op_produce_chunk<IN_GOOD>(w1a, x, mean, rstd, sx, kt, c, lane, a_hi, a_lo);
                } else {
                    op_produce_chunk<IN_ADV>(w1a, x, mean, rstd, sx, kt, c, lane, a_hi, a_lo);
                }
#endif
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic writes -> tensor core reads
                __syncwarp();
                if (lane == 0) tc_mbar_arrive(bar_full + st);      // one arrival per producer warp
            }
        }
    } else if (warp == 12) {
        // ===================== TMA: pre-split opponent fc2 (hi | lo), [256 x 32] boxes ==========
        if (lane == 0) {
            uint32_t it = 0;
            for (int job = job0; job < p.n_jobs; job += job_stride) {
                const int okey = job / jobs_per_ok;
                for (int kt = 0; kt < OP_KT; ++kt, ++it) {
                    const uint32_t st = it % Geo::STAGES, use = it / Geo::STAGES;
                    if (use > 0) tc_mbar_wait(bar_empty + st, (use - 1) & 1);
                    unsigned char* b_hi = stage_mem + (size_t)st * Geo::STAGE_BYTES + 2 * OP_A_BYTES;
#if defined(CEV_EXP) && (CEV_EXP & 2)         // development experiment: bit 1 = no opponent matrix stream
                    tc_mbar_arrive(bar_full + st);
                    (void)b_hi;
#else
                    tc_mbar_expect_tx(bar_full + st, 2 * OP_B_BYTES);
                    tma_load_2d(b_hi, &map_b, bar_full + st, kt * OP_BK, (okey * 2 + 0) * H2);
                    tma_load_2d(b_hi + OP_B_BYTES, &map_b, bar_full + st, kt * OP_BK, (okey * 2 + 1) * H2);
#endif
                }
            }
        }
    } else {
        // ===================== MMA issuer =====================
        const uint32_t stage_base = tc_smem_u32(stage_mem);
        uint32_t it = 0, pass = 0;
        for (int job = job0; job < p.n_jobs; job += job_stride, ++pass) {
            const uint32_t as = pass & 1, ause = pass >> 1;
            if (ause > 0) tc_mbar_wait(bar_tempty + as, (ause - 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t d_tmem = tmem_base + as * OP_BN;
            for (int kt = 0; kt < OP_KT; ++kt, ++it) {
                const uint32_t st = it % Geo::STAGES, use = it / Geo::STAGES;
                tc_mbar_wait(bar_full + st, use & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                {
                    // issued by one elected lane of the converged warp (tc_issue_ktile_3xf16, tc_common.cuh)
                    const uint32_t s_addr = stage_base + st * Geo::STAGE_BYTES;
                    tc_issue_ktile_3xf16(d_tmem, umma_desc_sw128(s_addr), umma_desc_sw128(s_addr + OP_A_BYTES),
                                          umma_desc_sw128(s_addr + 2 * OP_A_BYTES),
                                          umma_desc_sw128(s_addr + 2 * OP_A_BYTES + Geo::B_BYTES), OP_IDESC, kt ? 1u : 0u,
                                          tc_smem_u32(bar_empty + st));
                    if (kt == OP_KT - 1) tc_commit_elected(tc_smem_u32(bar_tfull + as));
                }
                __syncwarp();
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0 && *flag && p.status) atomicOr(p.status, CEV_STATUS_NONFINITE);
    if (warp == 13)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u) : "memory");
}

}  // namespace cev

#include "ls_member_tc.cuh"

namespace cev {

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// opponent preparation (3) + member layer-1 statistics + initial states + 3 kernels per world step
int rollout_lockstep_launches(int n_cycles) { return 5 + 3 * n_cycles; }

// How the SMs are shared out between the two persistent kernels of a world step (tensor-core member form).
// The opponent kernel is bound by the SMs it gets (~14 us per 128-episode job, ~12 us to get going); a member job
// streams 512 KB and takes ~8.6 us on its SM until the HBM rate (~6 TB/s over all member CTAs) becomes the limit.  Both end when their slowest CTA ends, so the split that minimises the later of the two is
// found by trying every one (148 candidates, host side).
static void ls_split_sms(int n_sm, int opp_jobs, int mem_jobs, int* g_opp, int* g_mem) {
    double best = 1e30;
    int bo = n_sm / 2;
    for (int g = 1; g < n_sm; ++g) {
        const int gm = n_sm - g;
        const int mg = mem_jobs < gm ? mem_jobs : gm, og = opp_jobs < g ? opp_jobs : g;
        const double t_job_mem = fmax(8.6, mg * 0.0873);
        const double t_opp = 12.0 + 14.0 * ((opp_jobs + og - 1) / og);
        const double t_mem = 4.0 + t_job_mem * ((mem_jobs + mg - 1) / mg);
        const double t = fmax(t_opp, t_mem);
        if (t < best - 1e-9) {
            best = t;
            bo = g;
        }
    }
    *g_opp = bo;
    *g_mem = n_sm - bo;
}

// Everything one role's rollout needs on the device: workspace views, tensor maps, kernel parameter blocks.
struct LsRoleCtx {
    LsBuffers b;
    CUtensorMap map_b, map_wtc;
    LsEnvParams ep;
    LsOppParams op;
    LsMemberTcParams tp;
    int env_blocks;
    int64_t member_ctas;
    int ms;
};

static int ls_configure(cev_handle* h) {
    static bool configured[16] = {};
    if (h->device < 16 && !configured[h->device]) {
        CEV_CUDA(cudaFuncSetAttribute(ls_opp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OP_SMEM));
        CEV_CUDA(cudaFuncSetAttribute(ls_member_tc_kernel<IN_ADV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MT_SMEM));
        CEV_CUDA(cudaFuncSetAttribute(ls_member_tc_kernel<IN_GOOD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MT_SMEM));
        configured[h->device] = true;
    }
    return CEV_OK;
}

static int ls_reserve(cev_handle* h, size_t need) {
    if (h->ls_workspace_bytes < need) {
        if (h->ls_workspace) CEV_CUDA(cudaFree(h->ls_workspace));
        h->ls_workspace = nullptr;
        h->ls_workspace_bytes = 0;
        CEV_CUDA(cudaMalloc(&h->ls_workspace, need));
        h->ls_workspace_bytes = need;
    }
    return CEV_OK;
}

// Launches the once-per-rollout preparation of one role (on `stream`: FP16 split of the opponents' fc2 and their
// layer-1 statistics; on `stream_m`: the layer-1 statistics of every member row) and fills the context.
static int ls_build_role(cev_handle* h, const ClusterParams& p, void* ws, cudaStream_t stream, cudaStream_t stream_m,
                         LsRoleCtx* c) {
    EncodeTiledFn encode = get_encode_fn();
    if (!encode) {
        set_error("rollout_lockstep: cuTensorMapEncodeTiled is not available from the driver");
        return CEV_ERR_UNSUPPORTED;
    }
    const int64_t N = (int64_t)p.P * p.K * p.E;
    LsBuffers& b = c->b;
    ls_carve(ws, N, p.K, p.P, &b);
    const int ms = p.member_seat;
    c->ms = ms;
    const int seat_of[2] = {ms == 0 ? 1 : 0, ms == 2 ? 1 : 2};

    // ---- opponents: scaled FP16 split + layer-1 statistics -----------------------------------
    LsPrepParams pp{};
    for (int oi = 0; oi < 2; ++oi) {
        pp.opp[oi] = p.opp[oi];
        pp.opp_pitch[oi] = p.opp_pitch[oi];
        pp.seat[oi] = seat_of[oi];
    }
    pp.K = p.K;
    pp.w2split = b.w2split;
    pp.wmax = b.wmax;
    pp.oscale = b.oscale;
    pp.l1stats = b.l1stats;
    CEV_CUDA(cudaMemsetAsync(b.wmax, 0, (size_t)2 * p.K * sizeof(uint32_t), stream));
    ls_wmax_kernel<<<dim3(32, 2 * p.K), 256, 0, stream>>>(pp);
    ls_split_w2_kernel<<<dim3(32, 2 * p.K), 256, 0, stream>>>(pp);
    ls_l1stats_kernel<<<2 * p.K, 512, 0, stream>>>(pp);
    ls_member_l1stats_kernel<<<p.P, 512, 0, stream_m>>>(p.members, p.member_pitch, seat_in_dim(ms), b.ml1stats);

    // ---- tensor maps ---------------------------------------------------------------------------
    const FcOffsets om = fc_offsets(seat_in_dim(ms));
    {
        cuuint64_t dims[3] = {(cuuint64_t)H1, (cuuint64_t)H2, (cuuint64_t)p.P};
        cuuint64_t strides[2] = {(cuuint64_t)H1 * 4, (cuuint64_t)p.member_pitch * 4};
        cuuint32_t box[3] = {MT_BK, H2, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&c->map_wtc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p.members + om.fc2w), dims,
                            strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("rollout_lockstep: cuTensorMapEncodeTiled(member fc2, tensor-core form) failed with %d", (int)r);
            return CEV_ERR_CUDA;
        }
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)H1, (cuuint64_t)2 * p.K * 2 * H2};
        cuuint64_t strides[1] = {(cuuint64_t)H1 * sizeof(__half)};
        cuuint32_t box[2] = {OP_BK, OP_BN};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&c->map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, b.w2split, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("rollout_lockstep: cuTensorMapEncodeTiled(opponent fc2) failed with %d", (int)r);
            return CEV_ERR_CUDA;
        }
    }

    // ---- per-step parameter blocks ---------------------------------------------------------------
    LsEnvParams& ep = c->ep;
    ep = LsEnvParams{};
    ep.b = b;
    ep.N = N;
    ep.K = p.K;
    ep.E = p.E;
    ep.init_shared = p.init_shared;
    ep.pos_first = p.pos_first;
    ep.init = p.init;
    ep.out = p.out;
    ep.opp_seat[0] = seat_of[0];
    ep.opp_seat[1] = seat_of[1];
    ep.status = p.status;
    c->env_blocks = (int)((N + 255) / 256);

    const int KE = p.K * p.E;
    const int n_chunks = (KE + LS_BT - 1) / LS_BT;
    c->member_ctas = (int64_t)p.P * n_chunks;
    CEV_REQUIRE(c->member_ctas < (1ll << 31), "rollout_lockstep: too many member tiles");

    LsOppParams& op = c->op;
    op = LsOppParams{};
    for (int oi = 0; oi < 2; ++oi) {
        op.opp[oi] = p.opp[oi];
        op.opp_pitch[oi] = p.opp_pitch[oi];
        op.seat[oi] = seat_of[oi];
    }
    op.K = p.K;
    op.E = p.E;
    op.PE = (int64_t)p.P * p.E;
    op.n_tiles = (int)((op.PE + OP_BM - 1) / OP_BM);
    op.n_jobs = 2 * p.K * op.n_tiles;
    op.N = N;
    op.obs = b.obs;
    op.act = b.act;
    op.gap = b.gap;
    op.l1ms = b.l1ms;
    op.oscale = b.oscale;
    op.status = p.status;

    LsMemberTcParams& tp = c->tp;
    tp = LsMemberTcParams{};
    tp.members = p.members;
    tp.pitch = p.member_pitch;
    tp.KE = KE;
    tp.n_chunks = n_chunks;
    tp.seat = ms;
    tp.n_jobs = (int)c->member_ctas;
    tp.N = N;
    tp.obs = b.obs + (int64_t)ms * N * LS_OBS_PAD;
    tp.act = b.act + (int64_t)ms * N;
    tp.gap = b.gap + (int64_t)ms * N;
    tp.l1stats = b.ml1stats;
    tp.status = p.status;
    return CEV_OK;
}

static void ls_launch_member_tc(const LsRoleCtx& c, int grid, cudaStream_t stream) {
    if (c.ms == 0) ls_member_tc_kernel<IN_ADV><<<grid, MT_THREADS, MT_SMEM, stream>>>(c.map_wtc, c.tp);
    else ls_member_tc_kernel<IN_GOOD><<<grid, MT_THREADS, MT_SMEM, stream>>>(c.map_wtc, c.tp);
}

int launch_rollout_lockstep(cev_handle* h, const ClusterParams& p, cudaStream_t stream) {
    if (p.P <= 0 || p.K <= 0 || p.E <= 0) return CEV_OK;
    const int64_t N = (int64_t)p.P * p.K * p.E;
    int rc = ls_reserve(h, ls_carve(nullptr, N, p.K, p.P, nullptr));
    if (rc) return rc;
    rc = ls_configure(h);
    if (rc) return rc;
    LsRoleCtx ctx;
    rc = ls_build_role(h, p, h->ls_workspace, stream, stream, &ctx);
    if (rc) return rc;
    LsEnvParams& ep = ctx.ep;
    LsOppParams& op = ctx.op;
    LsMemberTcParams& tp = ctx.tp;
    const int ms = ctx.ms;
    const int env_blocks = ctx.env_blocks;
    int opp_grid = op.n_jobs < h->n_sm ? op.n_jobs : h->n_sm;
    int mem_grid = tp.n_jobs < h->n_sm ? tp.n_jobs : h->n_sm;
    // The two forwards of a world step are independent (same observations): both kernels are persistent and get a
    // share of the SMs each (the opponent kernel on a side stream), so the HBM stream of the member rows runs under
    // the opponents' MMAs.  With the per-kernel timing hook on (cev_kernel_timing_enable) or CEV_LS_FORK=0 they run
    // one after the other on all SMs.
    static const int want_fork = getenv("CEV_LS_FORK") ? atoi(getenv("CEV_LS_FORK")) : 1;
    const bool fork = want_fork && !h->timing_on;
    if (fork) {
        // CEV_LS_GRID_OPP / CEV_LS_GRID_MEM override the split (development aid)
        static const int g_opp_env = getenv("CEV_LS_GRID_OPP") ? atoi(getenv("CEV_LS_GRID_OPP")) : LS_GRID_OPP_DEFAULT;
        static const int g_mem_env = getenv("CEV_LS_GRID_MEM") ? atoi(getenv("CEV_LS_GRID_MEM")) : LS_GRID_MEM_DEFAULT;
        int g_opp = 0, g_mem = 0;
        ls_split_sms(h->n_sm, op.n_jobs, tp.n_jobs, &g_opp, &g_mem);
        if (g_opp_env > 0) g_opp = g_opp_env;
        if (g_mem_env > 0) g_mem = g_mem_env;
        if (g_opp < opp_grid) opp_grid = g_opp;
        if (g_mem < mem_grid) mem_grid = g_mem;
    }

    // development aid: CEV_LS_SKIP bit 0 = no opponent kernel, bit 1 = no member kernel (timing only, results are
    // then invalid)
    static const int skip = getenv("CEV_LS_SKIP") ? atoi(getenv("CEV_LS_SKIP")) : 0;
    if (fork && !h->side_stream) {
        CEV_CUDA(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
        CEV_CUDA(cudaEventCreateWithFlags(&h->fork_ev, cudaEventDisableTiming));
        CEV_CUDA(cudaEventCreateWithFlags(&h->join_ev, cudaEventDisableTiming));
    }
    ep.last = p.n_cycles == 0;
    ls_init_kernel<<<env_blocks, 256, 0, stream>>>(ep);
    for (int c = 0; c < p.n_cycles; ++c) {
        if (p.trace_logits) {
            op.logits = p.trace_logits + (size_t)c * 3 * N * NACT;
            tp.logits = op.logits + (size_t)ms * N * NACT;
        }
        ep.forced = p.trace_forced ? p.trace_forced + (size_t)c * 3 * N : nullptr;
        ep.act_out = p.trace_actions ? p.trace_actions + (size_t)c * 3 * N : nullptr;
        // optional per-kernel CUDA-event timing on the launch stream (cev_kernel_timing_enable)
        auto tick = [&](int which, int end) {
            if (h->timing_on && h->timing_n[which] < CEV_TIMING_MAX) {
                cudaEventRecord(h->timing_ev[which][2 * h->timing_n[which] + end], stream);
                if (end) ++h->timing_n[which];
            }
        };
        if (!(skip & 1)) {
            if (fork) {
                CEV_CUDA(cudaEventRecord(h->fork_ev, stream));
                CEV_CUDA(cudaStreamWaitEvent(h->side_stream, h->fork_ev, 0));
                ls_opp_kernel<<<opp_grid, OP_THREADS, OP_SMEM, h->side_stream>>>(ctx.map_b, op);
                CEV_CUDA(cudaEventRecord(h->join_ev, h->side_stream));
            } else {
                tick(1, 0);
                ls_opp_kernel<<<opp_grid, OP_THREADS, OP_SMEM, stream>>>(ctx.map_b, op);
                tick(1, 1);
            }
        }
        if (!(skip & 2)) {
            tick(0, 0);
            ls_launch_member_tc(ctx, mem_grid, stream);
            tick(0, 1);
        }
        if (fork && !(skip & 1)) CEV_CUDA(cudaStreamWaitEvent(stream, h->join_ev, 0));
        ep.last = c == p.n_cycles - 1;
        ls_env_step_kernel<<<env_blocks, 256, 0, stream>>>(ep);
    }
    return check_cuda(cudaGetLastError(), "rollout_lockstep launch");
}

// The evaluations of several roles of one generation (the reference's three role loops, evolutionary_strategy.py:
// 236-251, genetic_algorithm.py:125-217) as ONE lockstep pass: per world step the roles' member kernels run back to
// back on the launch stream, their opponent kernels back to back on a second stream and their environment steps
// on a third (high priority), each waiting only for the two kernels of ITS role.  With one role per CUDA stream the
// block scheduler mixes the six persistent kernels at random (two member kernels sharing the HBM stream while the
// tensor pipes idle, then the reverse); here there is always exactly one member kernel on its share of the SMs and
// one opponent kernel on the rest, and a role's environment step hides under the next role's kernels.
// Same results as n_roles calls of launch_rollout_lockstep.  Every role has the same P, K, E.
int launch_rollout_lockstep_roles(cev_handle* h, const ClusterParams* ps, int n_roles, cudaStream_t stream) {
    if (n_roles <= 0) return CEV_OK;
    const ClusterParams& p0 = ps[0];
    if (p0.P <= 0 || p0.K <= 0 || p0.E <= 0) return CEV_OK;
    static const int want_roles = getenv("CEV_LS_ROLES") ? atoi(getenv("CEV_LS_ROLES")) : 1;
    if (h->timing_on || !want_roles || n_roles > CEV_MAX_ROLES) {
        for (int r = 0; r < n_roles; ++r) {
            const int rc = launch_rollout_lockstep(h, ps[r], stream);
            if (rc) return rc;
        }
        return CEV_OK;
    }
    const int64_t N = (int64_t)p0.P * p0.K * p0.E;
    const size_t per_role = ls_align(ls_carve(nullptr, N, p0.K, p0.P, nullptr));
    int rc = ls_reserve(h, per_role * n_roles);
    if (rc) return rc;
    rc = ls_configure(h);
    if (rc) return rc;
    if (!h->side_stream) {
        CEV_CUDA(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
        CEV_CUDA(cudaEventCreateWithFlags(&h->fork_ev, cudaEventDisableTiming));
        CEV_CUDA(cudaEventCreateWithFlags(&h->join_ev, cudaEventDisableTiming));
    }
    if (!h->env_stream) {
        // priorities: environment steps first (tiny, on the critical path of their role), then the opponent kernels
        // (they only ever ask for their share of the SMs), then the member kernels, which take what is left
        int lo = 0, hi = 0;
        CEV_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        const int mid = hi < lo ? lo - 1 : lo;
        static const int prio = getenv("CEV_LS_PRIO") ? atoi(getenv("CEV_LS_PRIO")) : 0;   // development aid
        const int p_opp = prio == 1 ? lo : mid, p_mem = prio == 2 ? (hi < lo - 1 ? lo - 2 : lo) : lo;
        CEV_CUDA(cudaStreamCreateWithPriority(&h->env_stream, cudaStreamNonBlocking, hi));
        for (int i = 0; i < 3; ++i) CEV_CUDA(cudaStreamCreateWithPriority(&h->opp_stream2[i], cudaStreamNonBlocking, p_opp));
        for (int i = 0; i < 2; ++i) CEV_CUDA(cudaStreamCreateWithPriority(&h->mem_stream2[i], cudaStreamNonBlocking, p_mem));
        for (int r = 0; r < CEV_MAX_ROLES; ++r) {
            CEV_CUDA(cudaEventCreateWithFlags(&h->ev_opp[r], cudaEventDisableTiming));
            CEV_CUDA(cudaEventCreateWithFlags(&h->ev_mem[r], cudaEventDisableTiming));
            CEV_CUDA(cudaEventCreateWithFlags(&h->ev_env[r], cudaEventDisableTiming));
        }
    }
    // Consecutive kernels of one kind alternate between two streams, so the next role's CTAs move in as the previous
    // role's CTAs retire (no idle tail at the end of every kernel, launch latency hidden).
    static const int alt = getenv("CEV_LS_ALT") ? atoi(getenv("CEV_LS_ALT")) : 1;
    static const int skip = getenv("CEV_LS_SKIP") ? atoi(getenv("CEV_LS_SKIP")) : 0;   // development aid, see the single-role path
    // ns streams per kind: 2 by default (alt = 0: one; CEV_LS_NSTREAMS=3: development aid)
    static const int ns_env = getenv("CEV_LS_NSTREAMS") ? atoi(getenv("CEV_LS_NSTREAMS")) : 2;
    const int ns = alt ? (ns_env >= 3 ? 3 : 2) : 1;
    cudaStream_t s_mem[3] = {stream, h->mem_stream2[0], h->mem_stream2[1]};
    cudaStream_t s_opp[3] = {h->opp_stream2[0], h->opp_stream2[1], h->opp_stream2[2]};
    cudaStream_t s_env = h->env_stream;
    CEV_CUDA(cudaEventRecord(h->fork_ev, stream));                  // everything before this call is on `stream`
    for (int k = 0; k < ns; ++k) {
        CEV_CUDA(cudaStreamWaitEvent(s_opp[k], h->fork_ev, 0));
        if (k > 0) CEV_CUDA(cudaStreamWaitEvent(s_mem[k], h->fork_ev, 0));
    }
    CEV_CUDA(cudaStreamWaitEvent(s_env, h->fork_ev, 0));
    // Every role's preparation goes on the streams of its first kernels: the opponents' split / statistics and the
    // initial states in front of its first opponent kernel, the members' statistics in front of its first member
    // kernel (ev_env[r] tells both streams that the initial observations exist), so the roles prepare side by side
    // and behind one another's first kernels instead of one after the other in front of the whole pass.
    LsRoleCtx ctx[CEV_MAX_ROLES];
    for (int r = 0; r < n_roles; ++r) {
        rc = ls_build_role(h, ps[r], static_cast<char*>(h->ls_workspace) + per_role * r, s_opp[r % ns], s_mem[r % ns], &ctx[r]);
        if (rc) return rc;
        ctx[r].ep.last = ps[r].n_cycles == 0;
        // the initial states go behind the opponents' preparation: ls_init_kernel evaluates the opponents'
        // LayerNorm-1 statistics on the first observations and reads what ls_l1stats_kernel wrote
        ls_init_kernel<<<ctx[r].env_blocks, 256, 0, s_opp[r % ns]>>>(ctx[r].ep);
        CEV_CUDA(cudaEventRecord(h->ev_env[r], s_opp[r % ns]));
    }
    int opp_grid = 0, mem_grid = 0;
    ls_split_sms(h->n_sm, ctx[0].op.n_jobs, ctx[0].tp.n_jobs, &opp_grid, &mem_grid);
    static const int g_opp_env = getenv("CEV_LS_GRID_OPP") ? atoi(getenv("CEV_LS_GRID_OPP")) : LS_GRID_OPP_DEFAULT;
    static const int g_mem_env = getenv("CEV_LS_GRID_MEM") ? atoi(getenv("CEV_LS_GRID_MEM")) : LS_GRID_MEM_DEFAULT;
    if (g_opp_env > 0) opp_grid = g_opp_env;
    if (g_mem_env > 0) mem_grid = g_mem_env;
    if (ctx[0].op.n_jobs < opp_grid) opp_grid = ctx[0].op.n_jobs;
    if (ctx[0].tp.n_jobs < mem_grid) mem_grid = ctx[0].tp.n_jobs;

    const int n_cycles = p0.n_cycles;
    int i = 0;
    for (int c = 0; c < n_cycles; ++c) {
        for (int r = 0; r < n_roles; ++r, ++i) {
            cudaStream_t so = s_opp[i % ns], sm = s_mem[i % ns];
            // this role's previous environment step (c = 0: its initial states)
            CEV_CUDA(cudaStreamWaitEvent(so, h->ev_env[r], 0));
            CEV_CUDA(cudaStreamWaitEvent(sm, h->ev_env[r], 0));
            if (!(skip & 1)) ls_opp_kernel<<<opp_grid, OP_THREADS, OP_SMEM, so>>>(ctx[r].map_b, ctx[r].op);
            CEV_CUDA(cudaEventRecord(h->ev_opp[r], so));
            if (!(skip & 2)) ls_launch_member_tc(ctx[r], mem_grid, sm);
            CEV_CUDA(cudaEventRecord(h->ev_mem[r], sm));
            CEV_CUDA(cudaStreamWaitEvent(s_env, h->ev_opp[r], 0));
            CEV_CUDA(cudaStreamWaitEvent(s_env, h->ev_mem[r], 0));
            ctx[r].ep.last = c == n_cycles - 1;
            ls_env_step_kernel<<<ctx[r].env_blocks, 256, 0, s_env>>>(ctx[r].ep);
            CEV_CUDA(cudaEventRecord(h->ev_env[r], s_env));
        }
    }
    // join: the environment steps wait for every member / opponent kernel of their role, and s_env is in order; the
    // other streams are joined too (with no world step to play they hold the preparation and the initial states)
    CEV_CUDA(cudaEventRecord(h->join_ev, s_env));
    CEV_CUDA(cudaStreamWaitEvent(stream, h->join_ev, 0));
    for (int k = 0; k < ns; ++k) {
        CEV_CUDA(cudaEventRecord(h->ev_opp[k], s_opp[k]));
        CEV_CUDA(cudaStreamWaitEvent(stream, h->ev_opp[k], 0));
        if (k > 0) {
            CEV_CUDA(cudaEventRecord(h->ev_mem[k], s_mem[k]));
            CEV_CUDA(cudaStreamWaitEvent(stream, h->ev_mem[k], 0));
        }
    }
    return check_cuda(cudaGetLastError(), "rollout_lockstep (roles) launch");
}

}  // namespace cev
