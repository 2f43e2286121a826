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
// member) stays on the FP32 pipe.  Per world step:
//
//   ls_member_kernel  one 4-warp CTA per (member, 16 episodes), two CTAs per SM: layer 1 + LayerNorm on CUDA
//                     cores, fc2 with packed FFMA2 (8 rows x 8 envs of accumulators per lane) from 8 KB TMA
//                     tensor tiles ([128 rows x 16 k], 64B swizzle) of the member's own row streamed from
//                     HBM (8 slots, each warp double-buffers its own tiles, no CTA-wide barrier in the
//                     loop), LayerNorm-2 + output layer + first-max argmax inside the CTA.
//   ls_opp_kernel     persistent, one CTA per SM, job = (opponent seat, opponent set, 128 episodes):
//                     8 producer warps (warp = 16-byte k chunk, lane = 4 episode rows) compute layer 1 +
//                     LayerNorm + ReLU and write the activations, split into TF32 hi + lo parts,
//                     straight into the 128B-swizzled K-major A tiles; one thread streams the pre-split opponent fc2
//                     matrix with TMA; one thread issues tcgen05.mma kind::tf32 (M=128, N=256, K=8)
//                     three times per k-step (hi.hi + lo.hi + hi.lo = "3xTF32", fp32-level accuracy)
//                     into a double-buffered TMEM accumulator; 4 epilogue warps read it back with
//                     tcgen05.ld and do bias + LayerNorm-2 + ReLU + output layer + argmax per row.
//                     Layer-1 LayerNorm statistics come from the closed form mean = wbar.x + bbar,
//                     var = z^T C z (z = [x; 1], C = row covariance of [W1 | b1], fp64), so a
//                     producer never needs the whole 512-vector at once.
//   ls_env_step_kernel one thread per episode: fp64 physics, rewards, next observations
//                     (bit-exact with oracle/mpe_env.py given equal actions).
//
// Arithmetic: member forward fp32 (FFMA); opponent fc2 3xTF32 with fp32 accumulation (error
// ~1e-6 of the activations' scale, the same order as fp32 summation-order noise); environment fp64.
#include <stdlib.h>

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
    float* w2split;    // [2 seats][K][hi, lo][256][512]
    double* l1stats;   // [2 seats][K][132]
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
    t.w2split = static_cast<float*>(take((size_t)2 * K * 2 * H2 * H1 * sizeof(float)));
    t.st = static_cast<double*>(take((size_t)LS_ST_FIELDS * N * sizeof(double)));
    t.acc = static_cast<double*>(take((size_t)3 * N * sizeof(double)));
    t.min_gap = static_cast<float*>(take((size_t)N * sizeof(float)));
    t.goal = static_cast<int32_t*>(take((size_t)N * sizeof(int32_t)));
    t.obs = static_cast<float*>(take((size_t)3 * N * LS_OBS_PAD * sizeof(float)));
    t.act = static_cast<int32_t*>(take((size_t)3 * N * sizeof(int32_t)));
    t.gap = static_cast<float*>(take((size_t)3 * N * sizeof(float)));
    t.l1stats = static_cast<double*>(take((size_t)2 * K * LS_L1S * sizeof(double)));
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
__device__ __forceinline__ void ls_store_obs(const LsBuffers& b, int64_t N, int64_t ep, const EnvState& s) {
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
    }
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
    ls_store_obs(p.b, p.N, ep, s);
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
    ls_store_obs(p.b, p.N, ep, s);
}

// ---------------------------------------------------------------------------------------------
// opponent preparation (once per launch): TF32 hi/lo split of fc2.W, layer-1 row statistics
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

struct LsPrepParams {
    const float* opp[2];
    int64_t opp_pitch[2];
    int seat[2];
    int K;
    float* w2split;
    double* l1stats;
};

__global__ void __launch_bounds__(256) ls_split_w2_kernel(const LsPrepParams p) {
    const int oi = blockIdx.y / p.K, k = blockIdx.y % p.K;
    const FcOffsets o = fc_offsets(seat_in_dim(p.seat[oi]));
    const float4* src = reinterpret_cast<const float4*>(p.opp[oi] + (int64_t)k * p.opp_pitch[oi] + o.fc2w);
    float4* hi = reinterpret_cast<float4*>(p.w2split + (size_t)(blockIdx.y * 2 + 0) * H2 * H1);
    float4* lo = reinterpret_cast<float4*>(p.w2split + (size_t)(blockIdx.y * 2 + 1) * H2 * H1);
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < H2 * H1 / 4; f += gridDim.x * blockDim.x) {
        const float4 w = src[f];
        float4 h, l;
        h.x = tf32_rna(w.x); l.x = tf32_rna(w.x - h.x);
        h.y = tf32_rna(w.y); l.y = tf32_rna(w.y - h.y);
        h.z = tf32_rna(w.z); l.z = tf32_rna(w.z - h.z);
        h.w = tf32_rna(w.w); l.w = tf32_rna(w.w - h.w);
        hi[f] = h;
        lo[f] = l;
    }
}

// wbar[j] = mean over the 512 rows of column j of [W1 | b1]; C[a][b] = row covariance (biased).
// LayerNorm-1 of y = W1 x + b1:  mean(y) = wbar . z,  var(y) = z^T C z  with z = [x; 1]
// (MPE/fcnetwork.py:44: nn.LayerNorm(512), biased variance).
__device__ __forceinline__ void ls_l1stats_row(const float* __restrict__ row, int in, double* __restrict__ out) {
    const int d = in + 1;
    const float* w1 = row;
    const float* b1 = row + H1 * in;
    __shared__ float cols[11][H1];          // [W1 | b1] transposed: column a of all 512 rows
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
    for (int pr = warp; pr < 121; pr += 16) {
        const int a = pr / 11, b = pr % 11;
        double s = 0.0;
        if (a < d && b < d) {
            const double ma = wbar[a], mb = wbar[b];
            for (int r = lane; r < H1; r += 32) s += ((double)cols[a][r] - ma) * ((double)cols[b][r] - mb);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) out[11 + pr] = s / H1;
    }
}

__global__ void __launch_bounds__(512) ls_l1stats_kernel(const LsPrepParams p) {
    const int oi = blockIdx.x / p.K, k = blockIdx.x % p.K;
    ls_l1stats_row(p.opp[oi] + (int64_t)k * p.opp_pitch[oi], seat_in_dim(p.seat[oi]),
                   p.l1stats + (size_t)blockIdx.x * LS_L1S);
}

// the same statistics for every member row (tensor-core member form), once per rollout
__global__ void __launch_bounds__(512) ls_member_l1stats_kernel(const float* __restrict__ members, int64_t pitch, int in,
                                                                double* __restrict__ out) {
    ls_l1stats_row(members + (int64_t)blockIdx.x * pitch, in, out + (size_t)blockIdx.x * LS_L1S);
}

// ---------------------------------------------------------------------------------------------
// member forward (FP32 pipe)
// ---------------------------------------------------------------------------------------------
#ifndef LS_MEMBER_TC_DEFAULT
#define LS_MEMBER_TC_DEFAULT 1                    // member forward: 1 = tensor cores beside the opponent kernel, 0 = FP32 pipe
#endif
#ifndef LS_GRID_OPP_DEFAULT
#define LS_GRID_OPP_DEFAULT 0                     // 0 = one CTA per SM
#endif
#ifndef LS_GRID_MEM_DEFAULT
#define LS_GRID_MEM_DEFAULT 0
#endif
constexpr int LS_BT = 16;                         // episodes per CTA
constexpr int LS_MT = 128;                        // threads per member CTA (4 warps)
constexpr int LS_MW = LS_MT / 32;
constexpr int LS_TILE_K = 16;                     // k per tile
constexpr int LS_TILE_ROWS = 128;
constexpr int LS_TILE_BYTES = LS_TILE_ROWS * LS_TILE_K * 4;   // [128 rows x 16 k] fp32 = 8 KB, 64B-swizzled
constexpr int LS_NSLOT = 8;
constexpr int LS_TPW = (H1 / LS_TILE_K) / 2;      // 16 tiles per warp: one row half, one k half
constexpr int LS_TAIL_FLOATS = 2056;              // fc2.b | ln2.g | ln2.b | out.W | out.b (+3 pad), contiguous in the row
constexpr int LS_W1A_FLOATS = H1 * IN_GOOD + 3 * H1;

// Two CTAs of four warps per SM (107 KB, up to 255 registers each): every lane owns an 8 rows x 8 envs
// accumulator tile, so a 4-k step is 16 LDS.128 for 128 FFMA2 (the 8 x 4 tile of an 8-warp CTA needed 12
// for 64 and left the LSU pipe as busy as the FMA pipe).  While one CTA is in its latency-bound phases
// (layer 1, LayerNorm, reductions, launch prologue) the other one keeps the FMA pipe and the HBM stream
// busy.  The W1 block is only needed by layer 1, so ring slots 4..7 alias it.
struct LsMemberSmem {
    static constexpr size_t off_ring = 0;
    static constexpr size_t off_w1a = off_ring + (size_t)(LS_NSLOT / 2) * LS_TILE_BYTES;   // = slots 4..7
    static constexpr size_t off_h1p = off_ring + (size_t)LS_NSLOT * LS_TILE_BYTES;
    static constexpr size_t off_tail = off_h1p + (size_t)H1 * LS_BT * 4;
    static constexpr size_t off_obs = off_tail + (size_t)LS_TAIL_FLOATS * 4;
    static constexpr size_t off_red1 = off_obs + (size_t)LS_BT * 12 * 4;
    static constexpr size_t off_red = off_red1 + (size_t)2 * LS_MW * LS_BT * 4;
    static constexpr size_t off_flag = off_red + (size_t)7 * LS_MW * LS_BT * 4;
    static constexpr size_t off_bar = off_flag + 16;
    static constexpr size_t total = off_bar + (size_t)(LS_NSLOT + 2) * 8 + 1024 /*alignment slack*/;
};
static_assert((size_t)LS_W1A_FLOATS * 4 <= (size_t)(LS_NSLOT / 2) * LS_TILE_BYTES, "W1 block must fit the aliased slots");
static_assert(2 * (LsMemberSmem::total + 1024) <= 233472, "two member CTAs must fit one SM");

struct LsMemberParams {
    const float* members;
    int64_t pitch;
    int n_chunks, KE, seat;
    int64_t N;
    const float* obs;     // this seat's [N][12]
    int32_t* act;         // this seat's [N]
    float* gap;
    int32_t* status;
    float* logits;        // this seat's [N][5] of this cycle (parity instrumentation; null in production)
};

// Layer 1 + LayerNorm + ReLU for the CTA's BT episodes with LS_MT threads (the 128-thread form of
// rollout_common.cuh's layer1): thread (ep = t % 8 env pair, g = t / 8) owns the row PAIRS g + 16 i;
// output in the k-pair interleaved layout h1p[(k >> 1) * (2 BT) + 2 e + (k & 1)].
template <int IN>
__device__ __forceinline__ void ls_layer1(const float* __restrict__ w1a, const float* __restrict__ obs_seat,
                                          float* __restrict__ h1p, float* __restrict__ red1, int* flag) {
    constexpr int BT = LS_BT;
    constexpr int NEP = BT / 2;             // env pairs
    constexpr int G = LS_MT / NEP;          // 16 row-pair groups
    constexpr int NP = (H1 / 2) / G;        // 16 row pairs per thread
    constexpr int NV = 2 * IN / 4;          // float4 per row pair of fc1.W
    const int t = threadIdx.x, ep = t % NEP, g = t / NEP, warp = t >> 5, lane = t & 31;
    const float* fc1w = w1a;
    const float* fc1b = w1a + H1 * IN;
    const float* ln1g = fc1b + H1;
    const float* ln1b = ln1g + H1;

    float2 ob[2][IN / 2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < IN / 2; ++k)
            ob[j][k] = *reinterpret_cast<const float2*>(obs_seat + (2 * ep + j) * 12 + 2 * k);

    float pre[NP][2][2];                    // [pair][row in pair][env in pair]
    float lsum[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const int rp = g + G * i;
        float w[2 * IN];
        const float4* wp = reinterpret_cast<const float4*>(fc1w + rp * 2 * IN);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const float4 x = wp[v];
            w[4 * v] = x.x; w[4 * v + 1] = x.y; w[4 * v + 2] = x.z; w[4 * v + 3] = x.w;
        }
        const float2 bb = *reinterpret_cast<const float2*>(fc1b + 2 * rp);
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                for (int k = 0; k < IN / 2; ++k)
                    acc = ffma2(make_float2(w[r * IN + 2 * k], w[r * IN + 2 * k + 1]), ob[j][k], acc);
                pre[i][r][j] = (acc.x + acc.y) + (r ? bb.y : bb.x);
                lsum[j] += pre[i][r][j];
            }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        lsum[j] = group_sum<NEP>(lsum[j]);
        if (lane < NEP) red1[warp * BT + 2 * ep + j] = lsum[j];
    }
    __syncthreads();
    float mean[2], lsq[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < LS_MW; ++w) tot += red1[w * BT + 2 * ep + j];
        mean[j] = tot * (1.0f / H1);
    }
#pragma unroll
    for (int i = 0; i < NP; ++i)
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                pre[i][r][j] -= mean[j];
                lsq[j] = fmaf(pre[i][r][j], pre[i][r][j], lsq[j]);
            }
    float* red1b = red1 + LS_MW * BT;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        lsq[j] = group_sum<NEP>(lsq[j]);
        if (lane < NEP) red1b[warp * BT + 2 * ep + j] = lsq[j];
    }
    __syncthreads();
    float rstd[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < LS_MW; ++w) tot += red1b[w * BT + 2 * ep + j];
        const float var = tot * (1.0f / H1);
        if (!isfinite(mean[j]) || !isfinite(var)) *flag = 1;
        rstd[j] = 1.0f / sqrtf(var + LN_EPS);
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const int rp = g + G * i;
        const float2 gg = *reinterpret_cast<const float2*>(ln1g + 2 * rp);
        const float2 be = *reinterpret_cast<const float2*>(ln1b + 2 * rp);
        float4 o;
        o.x = fmaxf(fmaf(pre[i][0][0] * rstd[0], gg.x, be.x), 0.f);
        o.y = fmaxf(fmaf(pre[i][1][0] * rstd[0], gg.y, be.y), 0.f);
        o.z = fmaxf(fmaf(pre[i][0][1] * rstd[1], gg.x, be.x), 0.f);
        o.w = fmaxf(fmaf(pre[i][1][1] * rstd[1], gg.y, be.y), 0.f);
        *reinterpret_cast<float4*>(h1p + rp * (2 * BT) + 4 * ep) = o;
    }
}

// One [128 rows x 16 k] tile of fc2 for one warp: lane = (eg = lane & 1: 8 envs, rl = lane >> 1: row lane),
// rows rl + 16 i (i < 8).  Tile rows are 64 bytes, 16-byte chunk c of row r stored at c ^ ((r >> 1) & 3)
// (TMA SWIZZLE_64B): the 16 row lanes of a load read 256 bytes in the minimal two wavefronts.
__device__ __forceinline__ void ls_fc2_tile(const float4* __restrict__ tile, const float* __restrict__ h1p, int kbase,
                                            float2 (&acc)[8][8]) {
    const int lane = threadIdx.x & 31;
    const int eg = lane & 1, rl = lane >> 1;
    const int sw = (rl >> 1) & 3;                 // (r >> 1) & 3 for r = rl + 16 i
    const float4* wrow0 = tile + rl * 4;
    const float* hbase = h1p + (kbase >> 1) * (2 * LS_BT) + eg * 16;
#pragma unroll
    for (int st = 0; st < LS_TILE_K / 4; ++st) {
        // activations: two k-pairs x four env-pairs ((k even, k odd) per env)
        float4 a[2][4];
#pragma unroll
        for (int kp = 0; kp < 2; ++kp)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                a[kp][j] = *reinterpret_cast<const float4*>(hbase + (2 * st + kp) * (2 * LS_BT) + j * 4);
        const int col = st ^ sw;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 w = wrow0[i * 64 + col];
            const float2 w0 = make_float2(w.x, w.y), w1 = make_float2(w.z, w.w);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[i][2 * j] = ffma2(w0, make_float2(a[0][j].x, a[0][j].y), acc[i][2 * j]);
                acc[i][2 * j] = ffma2(w1, make_float2(a[1][j].x, a[1][j].y), acc[i][2 * j]);
                acc[i][2 * j + 1] = ffma2(w0, make_float2(a[0][j].z, a[0][j].w), acc[i][2 * j + 1]);
                acc[i][2 * j + 1] = ffma2(w1, make_float2(a[1][j].z, a[1][j].w), acc[i][2 * j + 1]);
            }
        }
    }
}

__global__ void __launch_bounds__(LS_MT, 2)
ls_member_kernel(const __grid_constant__ CUtensorMap map_w2, const LsMemberParams p) {
    using L = LsMemberSmem;
    constexpr int BT = LS_BT;
    constexpr int G = LS_MT / BT;    // 8 row groups
    constexpr int RP = H2 / G;       // 32 fc2 rows per thread after the k-split reduce
    // 1024-byte alignment for the swizzled TMA tiles, computed as an OFFSET into the __shared__ array so
    // the compiler keeps the shared address space (LDS/STS, not generic LD/ST)
    extern __shared__ unsigned char ls_raw[];
    unsigned char* smem = ls_raw + ((1024u - (smem_u32(ls_raw) & 1023u)) & 1023u);
    unsigned char* ring = smem + L::off_ring;
    float* h1p = reinterpret_cast<float*>(smem + L::off_h1p);        // activations, later the k-split partials
    float* w1a = reinterpret_cast<float*>(smem + L::off_w1a);        // aliases ring slots 4..7
    float* tail = reinterpret_cast<float*>(smem + L::off_tail);
    float* obs = reinterpret_cast<float*>(smem + L::off_obs);
    float* red1 = reinterpret_cast<float*>(smem + L::off_red1);
    float* redA = reinterpret_cast<float*>(smem + L::off_red);       // [MW][BT]
    float* redB = redA + LS_MW * BT;                                 // [MW][BT]
    float* redC = redB + LS_MW * BT;                                 // [MW][5][BT]
    int* flag = reinterpret_cast<int*>(smem + L::off_flag);
    uint64_t* bar_tile = reinterpret_cast<uint64_t*>(smem + L::off_bar);   // [LS_NSLOT]
    uint64_t* bar_w1 = bar_tile + LS_NSLOT;
    uint64_t* bar_tail = bar_w1 + 1;

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int m = blockIdx.x / p.n_chunks, ch = blockIdx.x % p.n_chunks;
    const int n_env = min(BT, p.KE - ch * BT);
    const int64_t ep0 = (int64_t)m * p.KE + (int64_t)ch * BT;
    const float* mrow = p.members + (int64_t)m * p.pitch;
    const int in_dim = seat_in_dim(p.seat);
    const FcOffsets om = fc_offsets(in_dim);

    if (t == 0) {
        *flag = 0;
        for (int i = 0; i < LS_NSLOT + 2; ++i) mbar_init(bar_tile + i, 1);
        mbar_fence_init();
    }
    __syncthreads();
    // tile sequence s = i * 4 + w: the i-th tile of warp w = row half (w & 1), k-tile (w >> 1) * 16 + i;
    // it lives in slot s % 8, so warp w double-buffers its own stream in slots w and w + 4.
    auto issue_tile = [&](int s) {
        const int w = s & 3, i = s >> 2, slot = s & (LS_NSLOT - 1);
        const int kt = (w >> 1) * LS_TPW + i, rh = w & 1;
        mbar_arrive_expect_tx(bar_tile + slot, LS_TILE_BYTES);
        tma_load_3d(ring + (size_t)slot * LS_TILE_BYTES, &map_w2, bar_tile + slot, kt * LS_TILE_K, rh * LS_TILE_ROWS, m);
    };
    if (t == 0) {
        const uint32_t w1_bytes = (uint32_t)(H1 * in_dim + 3 * H1) * 4;
        mbar_arrive_expect_tx(bar_w1, w1_bytes);
        bulk_g2s(w1a, mrow, w1_bytes, bar_w1);
#pragma unroll 1
        for (int s = 0; s < LS_NSLOT / 2; ++s) issue_tile(s);      // every warp's first tile (slots 0..3)
        mbar_arrive_expect_tx(bar_tail, LS_TAIL_FLOATS * 4);
        bulk_g2s(tail, mrow + om.fc2b, LS_TAIL_FLOATS * 4, bar_tail);
    }
    if (t < BT * 3) {
        const int e = t / 3, v = t % 3;
        const int64_t ep = ep0 + (e < n_env ? e : 0);
        reinterpret_cast<float4*>(obs)[e * 3 + v] = __ldg(reinterpret_cast<const float4*>(p.obs) + ep * 3 + v);
    }
    __syncthreads();
    mbar_wait(bar_w1, 0);
    if (p.seat == 0) ls_layer1<IN_ADV>(w1a, obs, h1p, red1, flag);
    else ls_layer1<IN_GOOD>(w1a, obs, h1p, red1, flag);
    __syncthreads();          // h1p complete; the W1 block is dead, slots 4..7 are free
    if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic reads of W1 before the TMA writes
        issue_tile(4 + warp);
    }

    // ---- fc2: warp w = (row half w & 1, k half w >> 1), 16 tiles of [128 x 16] -----------------
    float2 acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int i = 0; i < LS_TPW; ++i) {
        const int slot = warp + 4 * (i & 1);
        mbar_wait(bar_tile + slot, (uint32_t)(i >> 1));
        ls_fc2_tile(reinterpret_cast<const float4*>(ring + (size_t)slot * LS_TILE_BYTES), h1p,
                    ((warp >> 1) * LS_TPW + i) * LS_TILE_K, acc);
        __syncwarp();
        if (lane == 0 && i + 2 < LS_TPW) issue_tile((i + 2) * 4 + warp);
    }
    __syncthreads();          // every warp is done reading h1p
    {
        float* part = h1p;    // [2 k-halves][256 rows][BT]
        const int eg = lane & 1, rl = lane >> 1;
        const int kh = warp >> 1, rh = warp & 1;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = rh * LS_TILE_ROWS + rl + 16 * i;
            float* dst = part + ((size_t)(kh * H2 + row)) * BT + eg * 8;
            float4 v0, v1;
            v0.x = acc[i][0].x + acc[i][0].y;
            v0.y = acc[i][1].x + acc[i][1].y;
            v0.z = acc[i][2].x + acc[i][2].y;
            v0.w = acc[i][3].x + acc[i][3].y;
            v1.x = acc[i][4].x + acc[i][4].y;
            v1.y = acc[i][5].x + acc[i][5].y;
            v1.z = acc[i][6].x + acc[i][6].y;
            v1.w = acc[i][7].x + acc[i][7].y;
            *reinterpret_cast<float4*>(dst) = v0;
            *reinterpret_cast<float4*>(dst + 4) = v1;
        }
    }
    __syncthreads();
    mbar_wait(bar_tail, 0);
    const float* b2 = tail;
    const float* g2 = tail + H2;
    const float* be2 = tail + 2 * H2;
    const float* w3 = tail + 3 * H2;
    const float* b3 = tail + 3 * H2 + NACT * H2;
    const int e1 = t % BT, g1 = t / BT;
    float pre2[RP];
    float lsum = 0.f;
#pragma unroll
    for (int j = 0; j < RP; ++j) {
        const int row = g1 + G * j;
        pre2[j] = (h1p[(size_t)row * BT + e1] + h1p[(size_t)(H2 + row) * BT + e1]) + b2[row];
        lsum += pre2[j];
    }
    lsum = group_sum<BT>(lsum);
    if (lane < BT) redA[warp * BT + e1] = lsum;
    __syncthreads();
    float mean = 0.f;
#pragma unroll
    for (int w = 0; w < LS_MW; ++w) mean += redA[w * BT + e1];
    mean *= (1.0f / H2);
    float lsq = 0.f;
#pragma unroll
    for (int j = 0; j < RP; ++j) {
        const float d = pre2[j] - mean;
        lsq = fmaf(d, d, lsq);
    }
    lsq = group_sum<BT>(lsq);
    if (lane < BT) redB[warp * BT + e1] = lsq;
    __syncthreads();
    float var = 0.f;
#pragma unroll
    for (int w = 0; w < LS_MW; ++w) var += redB[w * BT + e1];
    var *= (1.0f / H2);
    if (!isfinite(mean) || !isfinite(var)) *flag = 1;
    const float rstd = 1.0f / sqrtf(var + LN_EPS);
    float pl[NACT] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < RP; ++j) {
        const int row = g1 + G * j;
        const float h = fmaxf(fmaf((pre2[j] - mean) * rstd, g2[row], be2[row]), 0.f);
#pragma unroll
        for (int a = 0; a < NACT; ++a) pl[a] = fmaf(w3[a * H2 + row], h, pl[a]);
    }
#pragma unroll
    for (int a = 0; a < NACT; ++a) {
        const float v = group_sum<BT>(pl[a]);
        if (lane < BT) redC[(warp * NACT + a) * BT + e1] = v;
    }
    __syncthreads();
    if (t < BT) {
        float lg[NACT];
        bool fin = true;
#pragma unroll
        for (int a = 0; a < NACT; ++a) {
            float v = redC[a * BT + t];
#pragma unroll
            for (int w = 1; w < LS_MW; ++w) v += redC[(w * NACT + a) * BT + t];
            lg[a] = v + b3[a];
            fin = fin && isfinite(lg[a]);
        }
        if (!fin) *flag = 1;
        float gap;
        const int a = argmax_first5(lg, gap);
        if (t < n_env) {
            p.act[ep0 + t] = a;
            p.gap[ep0 + t] = gap;
            if (p.logits) {
#pragma unroll
                for (int q = 0; q < NACT; ++q) p.logits[(ep0 + t) * NACT + q] = lg[q];
            }
        }
    }
    __syncthreads();
    if (t == 0 && *flag && p.status) atomicOr(p.status, CEV_STATUS_NONFINITE);
}

// ---------------------------------------------------------------------------------------------
// opponent forward (tcgen05, 3xTF32)
// ---------------------------------------------------------------------------------------------
constexpr int OP_THREADS = 448;                   // warps 0-3 epilogue, 4-11 A producers, 12 TMA, 13 MMA
constexpr int OP_PROD = 256;
constexpr int OP_BM = 128, OP_BN = 256, OP_BK = 32, OP_KT = H1 / OP_BK;
constexpr uint32_t OP_A_BYTES = OP_BM * OP_BK * 4;          // 16 KB
constexpr uint32_t OP_B_BYTES = OP_BN * OP_BK * 4;          // 32 KB
constexpr uint32_t OP_STAGE_BYTES = 2 * OP_A_BYTES + 2 * OP_B_BYTES;   // A hi | A lo | B hi | B lo = 96 KB
constexpr int OP_STAGES = 2;
constexpr size_t OP_OFF_W1A = (size_t)OP_STAGES * OP_STAGE_BYTES;
constexpr size_t OP_OFF_TAIL = OP_OFF_W1A + (size_t)LS_W1A_FLOATS * 4;      // fc2.b | ln2.g | ln2.b | out.W | out.b
constexpr size_t OP_OFF_BAR = OP_OFF_TAIL + (size_t)LS_TAIL_FLOATS * 4;
constexpr size_t OP_SMEM_MAX = 232448;                                     // 227 KB opt-in limit per CTA
constexpr size_t OP_SLACK = OP_SMEM_MAX - (OP_OFF_BAR + 128);              // what is left for the 1024-byte alignment
constexpr size_t OP_SMEM = OP_SMEM_MAX;
constexpr uint32_t OP_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(OP_BN >> 3) << 17) |
                              ((uint32_t)(OP_BM >> 4) << 24);

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
    const double* l1stats;
    int32_t* status;
    float* logits;           // this cycle's [3][N][5] (parity instrumentation; null in production)
};

__device__ __forceinline__ void op_umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(OP_IDESC), "r"(accumulate)
        : "memory");
}

// The twelve MMAs of one k-tile (4 k-steps x {a_lo . b_hi, a_hi . b_lo, a_hi . b_hi}, small terms first) and the commit
// that releases the stage, issued by ONE elected lane of a CONVERGED warp: inside a divergent `if (lane == 0)` every
// tcgen05.mma costs an ELECT + R2UR.BROADCAST chain (~20 instructions, ~130 cycles of issue for 128 cycles of tensor
// pipe), which held the tensor pipe at ~50 %.
__device__ __forceinline__ void op_issue_ktile(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                               uint32_t accumulate, uint32_t bar_empty) {
    asm volatile(
        "{\n"
        ".reg .pred pe, pa, pt;\n"
        ".reg .b64 ah1, ah2, ah3, al1, al2, al3, bh1, bh2, bh3, bl1, bl2, bl3;\n"
        "elect.sync _|pe, 0xffffffff;\n"
        "setp.ne.b32 pa, %5, 0;\n"
        "setp.eq.b32 pt, 0, 0;\n"
        "add.s64 ah1, %1, 2;\n add.s64 ah2, %1, 4;\n add.s64 ah3, %1, 6;\n"
        "add.s64 al1, %2, 2;\n add.s64 al2, %2, 4;\n add.s64 al3, %2, 6;\n"
        "add.s64 bh1, %3, 2;\n add.s64 bh2, %3, 4;\n add.s64 bh3, %3, 6;\n"
        "add.s64 bl1, %4, 2;\n add.s64 bl2, %4, 4;\n add.s64 bl3, %4, 6;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %2, %3, %7, pa;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %4, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %3, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], al1, bh1, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], ah1, bl1, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], ah1, bh1, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], al2, bh2, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], ah2, bl2, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], ah2, bh2, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], al3, bh3, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], ah3, bl3, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], ah3, bh3, %7, pt;\n"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_hi), "l"(a_lo), "l"(b_hi), "l"(b_lo), "r"(accumulate), "r"(bar_empty), "r"(OP_IDESC)
        : "memory");
}
__device__ __forceinline__ void op_commit_elected(uint32_t bar) {
    asm volatile(
        "{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\n"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(bar)
        : "memory");
}

// One producer thread's share of an A tile: 16-byte chunk c (4 activations) of k-tile kt for its FOUR
// episode rows lane + 32 j: relu(LN1(W1 x + b1)), split into TF32 hi + lo and stored K-major with the
// 128-byte swizzle the UMMA descriptor expects (chunk c of row r at c ^ (r & 7)).  The weight loads are
// warp-uniform (producer warp = chunk) and serve four rows each: a shared-memory broadcast costs four
// wavefronts per LDS.128, and with one row per thread the LSU pipe bounded the producers.  Rows of a
// quarter warp differ in (r & 7), so the tile stores are conflict-free.
template <int IN>
__device__ __forceinline__ void op_produce_chunk(const float* __restrict__ w1a, const float (&x)[4][IN_GOOD],
                                                 const float (&mean)[4], const float (&rstd)[4], int kt, int c,
                                                 int lane, unsigned char* __restrict__ a_hi,
                                                 unsigned char* __restrict__ a_lo) {
    const float* fc1b = w1a + H1 * IN;
    const float* ln1g = fc1b + H1;
    const float* ln1b = ln1g + H1;
    const int k0 = kt * OP_BK + c * 4;
    const float4 bb = *reinterpret_cast<const float4*>(fc1b + k0);
    const float4 gg = *reinterpret_cast<const float4*>(ln1g + k0);
    const float4 ee = *reinterpret_cast<const float4*>(ln1b + k0);
    const float bq[4] = {bb.x, bb.y, bb.z, bb.w}, gq[4] = {gg.x, gg.y, gg.z, gg.w}, eq[4] = {ee.x, ee.y, ee.z, ee.w};
    float hi[4][4], lo[4][4];      // [row j][k]
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
        float w[IN];
        const float2* wr = reinterpret_cast<const float2*>(w1a + (k0 + qq) * IN);
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
            const float h = fmaxf(fmaf((pre - mean[j]) * rstd[j], gq[qq], eq[qq]), 0.f);
            // hi = h truncated to TF32 (exactly what the tensor core reads of an fp32 word); lo = h - hi is
            // exact in fp32 and is itself read truncated: relative error 2^-21, like the dropped lo.lo term
            hi[j][qq] = __uint_as_float(__float_as_uint(h) & 0xffffe000u);
            lo[j][qq] = h - hi[j][qq];
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int r = lane + 32 * j;
        const int off = r * 128 + ((c ^ (r & 7)) << 4);
        *reinterpret_cast<float4*>(a_hi + off) = make_float4(hi[j][0], hi[j][1], hi[j][2], hi[j][3]);
        *reinterpret_cast<float4*>(a_lo + off) = make_float4(lo[j][0], lo[j][1], lo[j][2], lo[j][3]);
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
                    sum += __uint_as_float(v[i]) + bb.x;
                    sum += __uint_as_float(v[i + 1]) + bb.y;
                    sum += __uint_as_float(v[i + 2]) + bb.z;
                    sum += __uint_as_float(v[i + 3]) + bb.w;
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
                        const float d = (__uint_as_float(v[i + u]) + bq[u]) - mean;
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
                        const float x = (__uint_as_float(v[i + u]) + bq[u]) - mean;
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
        // ===================== A producers: layer 1 + LayerNorm + ReLU -> TF32 hi/lo tiles =====
        const int pt = threadIdx.x - 128;
        const int c = pt >> 5;                        // producer warp = 16-byte chunk (4 k) of every k-tile
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
            // observations of this thread's four episode rows lane + 32 j
            float x[4][IN_GOOD];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t jrow = (int64_t)tile * Geo::ROWS_PER_JOB + lane + 32 * j;
                const int64_t jj = jrow < p.PE ? jrow : p.PE - 1;
                const int64_t ep = ((jj / p.E) * p.K + k) * p.E + (jj % p.E);
                const float4* src = reinterpret_cast<const float4*>(p.obs + ((int64_t)seat * p.N + ep) * LS_OBS_PAD);
                const float4 v0 = __ldg(src), v1 = __ldg(src + 1), v2 = __ldg(src + 2);
                x[j][0] = v0.x; x[j][1] = v0.y; x[j][2] = v0.z; x[j][3] = v0.w;
                x[j][4] = v1.x; x[j][5] = v1.y; x[j][6] = v1.z; x[j][7] = v1.w;
                x[j][8] = v2.x; x[j][9] = v2.y;
            }
            // the job's first stage: acquire it now, its A-hi tile doubles as the exchange buffer of the
            // LayerNorm-1 statistics (computed once per row by warps 0..3, read by all eight)
            {
                const uint32_t st = it % Geo::STAGES, use = it / Geo::STAGES;
                if (use > 0) tc_mbar_wait(bar_empty + st, (use - 1) & 1);
            }
            float2* xch = reinterpret_cast<float2*>(stage_mem + (size_t)(it % Geo::STAGES) * Geo::STAGE_BYTES);
            if (c < 4) {
                // closed form (fp64): mean = wbar . z, var = z^T C z, z = [x; 1], for row lane + 32 c
                const double* S = p.l1stats + (size_t)okey * LS_L1S;
                double z[11];
#pragma unroll
                for (int a = 0; a < 11; ++a) {
                    float xv = 0.f;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (j == c && a < IN_GOOD) xv = x[j][a];
                    z[a] = a < in ? (double)xv : (a == in ? 1.0 : 0.0);
                }
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
                if (!isfinite(m1) || !isfinite(var)) *flag = 1;
                xch[lane + 32 * c] = make_float2(m1, 1.0f / sqrtf(var + LN_EPS));
            }
            asm volatile("bar.sync 1, 256;\n" ::: "memory");
            float mean[4], rstd[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 v = xch[lane + 32 * j];
                mean[j] = v.x;
                rstd[j] = v.y;
            }
            asm volatile("bar.sync 1, 256;\n" ::: "memory");          // exchange buffer read: the tile may be written
            for (int kt = 0; kt < OP_KT; ++kt, ++it) {
                const uint32_t st = it % Geo::STAGES, use = it / Geo::STAGES;
                if (kt > 0 && use > 0) tc_mbar_wait(bar_empty + st, (use - 1) & 1);
                unsigned char* a_hi = stage_mem + (size_t)st * Geo::STAGE_BYTES;
                unsigned char* a_lo = a_hi + OP_A_BYTES;
#if !(defined(CEV_EXP) && (CEV_EXP & 1))      // development experiment: bit 0 = producers write nothing
                if (in == IN_GOOD) {
                    op_produce_chunk<IN_GOOD>(w1a, x, mean, rstd, kt, c, lane, a_hi, a_lo);
                } else {
                    op_produce_chunk<IN_ADV>(w1a, x, mean, rstd, kt, c, lane, a_hi, a_lo);
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
                    // issued by one elected lane of the converged warp (see op_issue_ktile)
                    const uint32_t s_addr = stage_base + st * Geo::STAGE_BYTES;
                    op_issue_ktile(d_tmem, umma_desc_sw128(s_addr), umma_desc_sw128(s_addr + OP_A_BYTES),
                                   umma_desc_sw128(s_addr + 2 * OP_A_BYTES),
                                   umma_desc_sw128(s_addr + 2 * OP_A_BYTES + Geo::B_BYTES), kt ? 1u : 0u,
                                   tc_smem_u32(bar_empty + st));
                    if (kt == OP_KT - 1) op_commit_elected(tc_smem_u32(bar_tfull + as));
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
static int ls_member_tc_enabled() {
    static const int v = getenv("CEV_LS_MEMBER_TC") ? atoi(getenv("CEV_LS_MEMBER_TC")) : LS_MEMBER_TC_DEFAULT;
    return v;
}

// opponent preparation (2) [+ member layer-1 statistics] + initial states + 3 kernels per world step
int rollout_lockstep_launches(int n_cycles) { return 3 + (ls_member_tc_enabled() ? 1 : 0) + 3 * n_cycles; }

// How the SMs are shared out between the two persistent kernels of a world step (tensor-core member form).
// The opponent kernel is bound by the tensor pipe of the SMs it gets (~20 us per 128-episode job, ~15 us to get
// going); a member job streams 512 KB and takes ~8.6 us on its SM until the HBM rate (~6 TB/s over all member CTAs)
// becomes the limit.  Both end when their slowest CTA ends, so the split that minimises the later of the two is
// found by trying every one (148 candidates, host side).
static void ls_split_sms(int n_sm, int opp_jobs, int mem_jobs, int* g_opp, int* g_mem) {
    double best = 1e30;
    int bo = n_sm / 2;
    for (int g = 1; g < n_sm; ++g) {
        const int gm = n_sm - g;
        const int mg = mem_jobs < gm ? mem_jobs : gm, og = opp_jobs < g ? opp_jobs : g;
        const double t_job_mem = fmax(8.6, mg * 0.0873);
        const double t_opp = 15.0 + 20.0 * ((opp_jobs + og - 1) / og);
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
    CUtensorMap map_w2, map_b, map_wtc;
    LsEnvParams ep;
    LsMemberParams mp;
    LsOppParams op;
    LsMemberTcParams tp;
    int env_blocks;
    int64_t member_ctas;
    int ms;
};

static int ls_configure(cev_handle* h) {
    static bool configured[16] = {};
    if (h->device < 16 && !configured[h->device]) {
        CEV_CUDA(cudaFuncSetAttribute(ls_member_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)LsMemberSmem::total));
        CEV_CUDA(cudaFuncSetAttribute(ls_member_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
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

// Launches the once-per-rollout preparation of one role on `stream` (TF32 split of the opponents' fc2, layer-1
// statistics of the opponents and, for the tensor-core member form, of every member row) and fills the context.
static int ls_build_role(cev_handle* h, const ClusterParams& p, void* ws, int member_tc, cudaStream_t stream, LsRoleCtx* c) {
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

    // ---- opponents: TF32 split + layer-1 statistics ------------------------------------------
    LsPrepParams pp{};
    for (int oi = 0; oi < 2; ++oi) {
        pp.opp[oi] = p.opp[oi];
        pp.opp_pitch[oi] = p.opp_pitch[oi];
        pp.seat[oi] = seat_of[oi];
    }
    pp.K = p.K;
    pp.w2split = b.w2split;
    pp.l1stats = b.l1stats;
    ls_split_w2_kernel<<<dim3(32, 2 * p.K), 256, 0, stream>>>(pp);
    ls_l1stats_kernel<<<2 * p.K, 512, 0, stream>>>(pp);
    if (member_tc)
        ls_member_l1stats_kernel<<<p.P, 512, 0, stream>>>(p.members, p.member_pitch, seat_in_dim(ms), b.ml1stats);

    // ---- tensor maps ---------------------------------------------------------------------------
    const FcOffsets om = fc_offsets(seat_in_dim(ms));
    if (member_tc) {
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
    } else {
        cuuint64_t dims[3] = {(cuuint64_t)H1, (cuuint64_t)H2, (cuuint64_t)p.P};
        cuuint64_t strides[2] = {(cuuint64_t)H1 * 4, (cuuint64_t)p.member_pitch * 4};
        cuuint32_t box[3] = {LS_TILE_K, LS_TILE_ROWS, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        static const int l2p = getenv("CEV_LS_L2P") ? atoi(getenv("CEV_LS_L2P")) : 128;
        CUresult r = encode(&c->map_w2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p.members + om.fc2w), dims,
                            strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                            l2p == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                       : (l2p == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B),
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("rollout_lockstep: cuTensorMapEncodeTiled(member fc2) failed with %d", (int)r);
            return CEV_ERR_CUDA;
        }
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)H1, (cuuint64_t)2 * p.K * 2 * H2};
        cuuint64_t strides[1] = {(cuuint64_t)H1 * 4};
        cuuint32_t box[2] = {OP_BK, OP_BN};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&c->map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, b.w2split, dims, strides, box, estr,
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
    c->env_blocks = (int)((N + 255) / 256);

    LsMemberParams& mp = c->mp;
    mp = LsMemberParams{};
    mp.members = p.members;
    mp.pitch = p.member_pitch;
    mp.KE = p.K * p.E;
    mp.n_chunks = (mp.KE + LS_BT - 1) / LS_BT;
    mp.seat = ms;
    mp.N = N;
    mp.obs = b.obs + (int64_t)ms * N * LS_OBS_PAD;
    mp.act = b.act + (int64_t)ms * N;
    mp.gap = b.gap + (int64_t)ms * N;
    mp.status = p.status;
    c->member_ctas = (int64_t)p.P * mp.n_chunks;
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
    op.l1stats = b.l1stats;
    op.status = p.status;

    LsMemberTcParams& tp = c->tp;
    tp = LsMemberTcParams{};
    tp.members = p.members;
    tp.pitch = p.member_pitch;
    tp.KE = mp.KE;
    tp.n_chunks = mp.n_chunks;
    tp.seat = ms;
    tp.n_jobs = (int)c->member_ctas;
    tp.N = N;
    tp.obs = mp.obs;
    tp.act = mp.act;
    tp.gap = mp.gap;
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
    const int member_tc = ls_member_tc_enabled();
    LsRoleCtx ctx;
    rc = ls_build_role(h, p, h->ls_workspace, member_tc, stream, &ctx);
    if (rc) return rc;
    LsEnvParams& ep = ctx.ep;
    LsMemberParams& mp = ctx.mp;
    LsOppParams& op = ctx.op;
    LsMemberTcParams& tp = ctx.tp;
    const int ms = ctx.ms;
    const int env_blocks = ctx.env_blocks;
    const int64_t member_ctas = ctx.member_ctas;
    int opp_grid = op.n_jobs < h->n_sm ? op.n_jobs : h->n_sm;
    int mem_grid = tp.n_jobs < h->n_sm ? tp.n_jobs : h->n_sm;
    static const int want_fork = getenv("CEV_LS_FORK") ? atoi(getenv("CEV_LS_FORK")) : (member_tc ? 1 : 0);
    // The two forwards of a world step are independent (same observations).  FP32 member form: two streams let the
    // block scheduler fill one kernel's tail with the other's CTAs (+4 % at 1024 members, -3 % at 4096: opt-in).
    // Tensor-core member form: both kernels are persistent and get a share of the SMs each, so the HBM stream of the
    // member rows runs under the opponents' MMAs (default).
    const bool fork = want_fork && !h->timing_on;
    if (member_tc && fork) {
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

    // development aids: CEV_LS_SKIP bit 0 = no opponent kernel, bit 1 = no member kernel (timing only,
    // results are then invalid); CEV_LS_FORK=1 = opponent kernel on a side stream beside the member kernel
    static const int skip = getenv("CEV_LS_SKIP") ? atoi(getenv("CEV_LS_SKIP")) : 0;
    // CEV_LS_MEMBER_PAD = extra dynamic shared memory (bytes) per member CTA: 16384 leaves room for ONE member CTA
    // per SM (timing experiment: what a member CTA delivers when it does not share the SM with a second one)
    static const int member_pad = getenv("CEV_LS_MEMBER_PAD") ? atoi(getenv("CEV_LS_MEMBER_PAD")) : 0;
    static bool pad_configured = false;
    if (member_pad > 0 && !pad_configured) {
        CEV_CUDA(cudaFuncSetAttribute(ls_member_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)LsMemberSmem::total + member_pad));
        pad_configured = true;
    }
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
            mp.logits = op.logits + (size_t)ms * N * NACT;
            tp.logits = mp.logits;
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
            if (member_tc) {
                ls_launch_member_tc(ctx, mem_grid, stream);
            } else {
                ls_member_kernel<<<(unsigned)member_ctas, LS_MT, LsMemberSmem::total + member_pad, stream>>>(ctx.map_w2, mp);
            }
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
    const int member_tc = ls_member_tc_enabled();
    static const int want_roles = getenv("CEV_LS_ROLES") ? atoi(getenv("CEV_LS_ROLES")) : 1;
    if (!member_tc || h->timing_on || !want_roles || n_roles > CEV_MAX_ROLES) {
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
        CEV_CUDA(cudaStreamCreateWithPriority(&h->env_stream, cudaStreamNonBlocking, hi));
        CEV_CUDA(cudaStreamCreateWithPriority(&h->opp_stream2[0], cudaStreamNonBlocking, mid));
        CEV_CUDA(cudaStreamCreateWithPriority(&h->opp_stream2[1], cudaStreamNonBlocking, mid));
        CEV_CUDA(cudaStreamCreateWithPriority(&h->mem_stream2, cudaStreamNonBlocking, lo));
        for (int r = 0; r < CEV_MAX_ROLES; ++r) {
            CEV_CUDA(cudaEventCreateWithFlags(&h->ev_opp[r], cudaEventDisableTiming));
            CEV_CUDA(cudaEventCreateWithFlags(&h->ev_mem[r], cudaEventDisableTiming));
            CEV_CUDA(cudaEventCreateWithFlags(&h->ev_env[r], cudaEventDisableTiming));
        }
    }
    LsRoleCtx ctx[CEV_MAX_ROLES];
    for (int r = 0; r < n_roles; ++r) {
        rc = ls_build_role(h, ps[r], static_cast<char*>(h->ls_workspace) + per_role * r, 1, stream, &ctx[r]);
        if (rc) return rc;
        ctx[r].ep.last = ps[r].n_cycles == 0;
        ls_init_kernel<<<ctx[r].env_blocks, 256, 0, stream>>>(ctx[r].ep);
    }
    int opp_grid = 0, mem_grid = 0;
    ls_split_sms(h->n_sm, ctx[0].op.n_jobs, ctx[0].tp.n_jobs, &opp_grid, &mem_grid);
    static const int g_opp_env = getenv("CEV_LS_GRID_OPP") ? atoi(getenv("CEV_LS_GRID_OPP")) : LS_GRID_OPP_DEFAULT;
    static const int g_mem_env = getenv("CEV_LS_GRID_MEM") ? atoi(getenv("CEV_LS_GRID_MEM")) : LS_GRID_MEM_DEFAULT;
    if (g_opp_env > 0) opp_grid = g_opp_env;
    if (g_mem_env > 0) mem_grid = g_mem_env;
    if (ctx[0].op.n_jobs < opp_grid) opp_grid = ctx[0].op.n_jobs;
    if (ctx[0].tp.n_jobs < mem_grid) mem_grid = ctx[0].tp.n_jobs;

    // Consecutive kernels of one kind alternate between two streams, so the next role's CTAs move in as the previous
    // role's CTAs retire (no idle tail at the end of every kernel, launch latency hidden).
    static const int alt = getenv("CEV_LS_ALT") ? atoi(getenv("CEV_LS_ALT")) : 1;
    cudaStream_t s_mem[2] = {stream, alt ? h->mem_stream2 : stream};
    cudaStream_t s_opp[2] = {h->opp_stream2[0], alt ? h->opp_stream2[1] : h->opp_stream2[0]};
    cudaStream_t s_env = h->env_stream;
    CEV_CUDA(cudaEventRecord(h->fork_ev, stream));                  // preparation + initial states are on `stream`
    CEV_CUDA(cudaStreamWaitEvent(s_opp[0], h->fork_ev, 0));
    CEV_CUDA(cudaStreamWaitEvent(s_opp[1], h->fork_ev, 0));
    CEV_CUDA(cudaStreamWaitEvent(s_mem[1], h->fork_ev, 0));
    CEV_CUDA(cudaStreamWaitEvent(s_env, h->fork_ev, 0));
    const int n_cycles = p0.n_cycles;
    int i = 0;
    for (int c = 0; c < n_cycles; ++c) {
        for (int r = 0; r < n_roles; ++r, ++i) {
            cudaStream_t so = s_opp[i & 1], sm = s_mem[i & 1];
            if (c > 0) {                                            // this role's previous environment step
                CEV_CUDA(cudaStreamWaitEvent(so, h->ev_env[r], 0));
                CEV_CUDA(cudaStreamWaitEvent(sm, h->ev_env[r], 0));
            }
            ls_opp_kernel<<<opp_grid, OP_THREADS, OP_SMEM, so>>>(ctx[r].map_b, ctx[r].op);
            CEV_CUDA(cudaEventRecord(h->ev_opp[r], so));
            ls_launch_member_tc(ctx[r], mem_grid, sm);
            CEV_CUDA(cudaEventRecord(h->ev_mem[r], sm));
            CEV_CUDA(cudaStreamWaitEvent(s_env, h->ev_opp[r], 0));
            CEV_CUDA(cudaStreamWaitEvent(s_env, h->ev_mem[r], 0));
            ctx[r].ep.last = c == n_cycles - 1;
            ls_env_step_kernel<<<ctx[r].env_blocks, 256, 0, s_env>>>(ctx[r].ep);
            CEV_CUDA(cudaEventRecord(h->ev_env[r], s_env));
        }
    }
    // join: the environment steps wait for every member / opponent kernel of their role, and s_env is in order
    CEV_CUDA(cudaEventRecord(h->join_ev, s_env));
    CEV_CUDA(cudaStreamWaitEvent(stream, h->join_ev, 0));
    return check_cuda(cudaGetLastError(), "rollout_lockstep (roles) launch");
}

}  // namespace cev
