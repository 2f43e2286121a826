// K1 (structured form): fused persistent MPE rollout, one 4-CTA thread-block
// cluster per (member, opponent set, env-chunk) tile.
//
// Replaces, for P members x K opponent sets x E env instances:
//   play_game / play_MPE          utils/game_logic_functions.py:123-228
//   FCNetwork.forward / argmax    MPE/fcnetwork.py:37-90
//   simple_adversary_v3 world     SURVEY.md Appendix A (third-party pettingzoo)
//   the GA/ES evaluation loops    genetic_algorithm.py:125-217,
//                                 evolutionary_strategy.py:236-251
//
// Design:
//  * A member's fc2 matrix (256x512 fp32 = 512 KB) does not fit one SM, so a
//    cluster of 4 CTAs each keeps a 64-row quarter (128 KB) RESIDENT in shared
//    memory for the whole tile: member weights leave HBM once per tile, not
//    once per step.
//  * Opponent weights (shared by every tile, L2-resident) are streamed through
//    a 3-stage cp.async ring, 64 rows x 32 k per stage.
//  * Per cycle the three seats' forwards are independent (same world state):
//    each CTA computes layer 1 + LayerNorm redundantly (K <= 10, cheap), its
//    quarter of fc2 for all BT env instances with packed FFMA2 on k-pairs,
//    then the quarters are combined across the cluster through distributed
//    shared memory: LayerNorm-2 statistics (Chan's parallel mean/M2) and the
//    partial logits.  Two cluster barriers per cycle.
//  * Physics, observations, rewards: fp64 in registers (one thread per env),
//    bit-exact with oracle/mpe_env.py given equal actions; fitness sums never
//    leave the chip until the tile ends.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

#ifndef CEV_USE_FFMA2
#define CEV_USE_FFMA2 1
#endif

namespace cev {

constexpr int CL = 4;                       // CTAs per cluster
constexpr int CT = 256;                     // threads per CTA
constexpr int NW = CT / 32;                 // warps per CTA
constexpr int ROWS_Q = H2 / CL;             // fc2 rows per CTA (64)
constexpr int KCH = 32;                     // k per streamed chunk
constexpr int NCHUNK = H1 / KCH;            // 16
constexpr int NSTAGE = 3;
constexpr int STAGE_F4 = ROWS_Q * KCH / 4;  // float4 per stage (512)
constexpr int W1A_FLOATS = H1 * IN_GOOD + 3 * H1;   // fc1.W | fc1.b | ln1.g | ln1.b
constexpr int SMALL_FLOATS = 512;           // b2[64] g2[64] be2[64] w3[5][64]

template <int BT>
struct SmemLayout {
    static constexpr size_t off_w2m = 0;
    static constexpr size_t off_w1a = off_w2m + (size_t)ROWS_Q * H1 * 4;
    static constexpr size_t off_stage = off_w1a + (size_t)W1A_FLOATS * 4;
    static constexpr size_t off_u = off_stage + (size_t)NSTAGE * STAGE_F4 * 16;
    static constexpr size_t off_small = off_u + (size_t)H1 * BT * 4;
    static constexpr size_t off_obs = off_small + (size_t)(3 * SMALL_FLOATS + 24) * 4;
    static constexpr size_t off_red1 = off_obs + (size_t)3 * BT * 12 * 4;
    static constexpr size_t off_lnx = off_red1 + (size_t)2 * NW * BT * 4;
    static constexpr size_t off_plog = off_lnx + (size_t)3 * CL * BT * 8;
    static constexpr size_t off_act = off_plog + (size_t)3 * CL * NACT * BT * 4;
    static constexpr size_t off_flag = off_act + (size_t)3 * BT * 8;
    static constexpr size_t total = off_flag + 16;
};

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
#if CEV_USE_FFMA2
    return __ffma2_rn(a, b, c);
#else
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// sum over the lanes that share (lane % BT); result valid in every lane
template <int BT>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = 16; o >= BT; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------
// Layer 1 + LayerNorm + ReLU for all BT env instances (every CTA computes all
// 512 rows: K <= 10 makes redundancy cheaper than a DSMEM all-gather).
// Thread (e = t % BT, g = t / BT) owns rows g + G*i, i < 2*BT.
// Output: h1p[(k>>1)*(2*BT) + 2*e + (k&1)]  (k-pair interleaved for FFMA2).
// ---------------------------------------------------------------------------
template <int BT, int IN>
__device__ __forceinline__ void layer1(const float* __restrict__ w1a, const float* __restrict__ obs_seat,
                                       float* __restrict__ h1p, float* __restrict__ red1,
                                       int* flag) {
    constexpr int G = CT / BT;          // row groups
    constexpr int R1 = H1 / G;          // rows per thread (2*BT)
    const int t = threadIdx.x, e = t % BT, g = t / BT, warp = t >> 5;
    const float* fc1w = w1a;
    const float* fc1b = w1a + H1 * IN;
    const float* ln1g = fc1b + H1;
    const float* ln1b = ln1g + H1;

    float2 ob[IN / 2];
#pragma unroll
    for (int k = 0; k < IN / 2; ++k) ob[k] = *reinterpret_cast<const float2*>(obs_seat + e * 12 + 2 * k);

    float pre[R1];
    float lsum = 0.f;
#pragma unroll
    for (int i = 0; i < R1; ++i) {
        const int row = g + G * i;
        const float2* w = reinterpret_cast<const float2*>(fc1w + row * IN);
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < IN / 2; ++k) acc = ffma2(w[k], ob[k], acc);
        pre[i] = (acc.x + acc.y) + fc1b[row];
        lsum += pre[i];
    }
    // mean over 512 rows of env e: lanes sharing e, then the 8 warps
    lsum = group_sum<BT>(lsum);
    if ((t & 31) < BT) red1[warp * BT + e] = lsum;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) tot += red1[w * BT + e];
    const float mean = tot * (1.0f / H1);
    float lsq = 0.f;
#pragma unroll
    for (int i = 0; i < R1; ++i) {
        pre[i] -= mean;
        lsq = fmaf(pre[i], pre[i], lsq);
    }
    lsq = group_sum<BT>(lsq);
    float* red1b = red1 + NW * BT;
    if ((t & 31) < BT) red1b[warp * BT + e] = lsq;
    __syncthreads();
    tot = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) tot += red1b[w * BT + e];
    const float var = tot * (1.0f / H1);
    if (!isfinite(mean) || !isfinite(var)) *flag = 1;
    const float rstd = 1.0f / sqrtf(var + LN_EPS);
#pragma unroll
    for (int i = 0; i < R1; ++i) {
        const int row = g + G * i;
        const float v = fmaxf(fmaf(pre[i] * rstd, ln1g[row], ln1b[row]), 0.f);
        h1p[(row >> 1) * (2 * BT) + 2 * e + (row & 1)] = v;
    }
}

// ---------------------------------------------------------------------------
// fc2 quarter: this CTA's 64 rows x 512 k x BT envs.
// lane = (eg = lane % GE env group of 4, rl = lane / GE row lane);
// rows rl + RL*i, i < RPL.  Warp w takes the w-th 4-k step of every 32-k
// chunk.  Accumulators are float2 (even-k, odd-k partial sums).
// ---------------------------------------------------------------------------
template <int BT>
struct Fc2Map {
    static constexpr int GE = BT / 4;
    static constexpr int RL = 32 / GE;
    static constexpr int RPL = ROWS_Q / RL;
};

template <int BT>
__device__ __forceinline__ void fc2_step(const float4* __restrict__ wbase, int row_stride_f4, int chunk16,
                                         const float* __restrict__ h1p, int kbase,
                                         float2 (&acc)[Fc2Map<BT>::RPL][4]) {
    using M = Fc2Map<BT>;
    const int lane = threadIdx.x & 31;
    const int eg = lane % M::GE, rl = lane / M::GE;
    // activations: two k-pairs x two env-pairs
    float4 a[2][2];
#pragma unroll
    for (int kp = 0; kp < 2; ++kp)
#pragma unroll
        for (int j = 0; j < 2; ++j)
            a[kp][j] = *reinterpret_cast<const float4*>(h1p + ((kbase >> 1) + kp) * (2 * BT) + eg * 8 + j * 4);
#pragma unroll
    for (int i = 0; i < M::RPL; ++i) {
        const int row = rl + M::RL * i;
        const float4 w = wbase[row * row_stride_f4 + (chunk16 ^ (row & 7))];
        const float2 w0 = make_float2(w.x, w.y), w1 = make_float2(w.z, w.w);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            acc[i][2 * j] = ffma2(w0, make_float2(a[0][j].x, a[0][j].y), acc[i][2 * j]);
            acc[i][2 * j] = ffma2(w1, make_float2(a[1][j].x, a[1][j].y), acc[i][2 * j]);
            acc[i][2 * j + 1] = ffma2(w0, make_float2(a[0][j].z, a[0][j].w), acc[i][2 * j + 1]);
            acc[i][2 * j + 1] = ffma2(w1, make_float2(a[1][j].z, a[1][j].w), acc[i][2 * j + 1]);
        }
    }
}

// issue the cp.async copies of one streamed chunk (64 rows x 32 k) into a stage
__device__ __forceinline__ void issue_chunk(float4* stage, const float* __restrict__ w2q, int chunk) {
#pragma unroll
    for (int i = 0; i < STAGE_F4 / CT; ++i) {
        const int f = threadIdx.x + i * CT;       // 0..511
        const int row = f >> 3, j16 = f & 7;
        cp_async16(stage + row * 8 + (j16 ^ (row & 7)), w2q + (size_t)row * H1 + chunk * KCH + j16 * 4);
    }
}

// issue the cp.async copies of a seat's W1 block (contiguous in the flat row)
__device__ __forceinline__ void issue_w1(float* w1a, const float* __restrict__ wrow, int in_dim) {
    const int n16 = (H1 * in_dim + 3 * H1) / 4;
    for (int f = threadIdx.x; f < n16; f += CT) cp_async16(w1a + f * 4, wrow + f * 4);
}

template <int BT>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(CT, 1)
rollout_cluster_kernel(const ClusterParams p) {
    using L = SmemLayout<BT>;
    using M = Fc2Map<BT>;
    constexpr int G = CT / BT;            // row groups of the (e, g) mapping
    constexpr int RP = ROWS_Q / G;        // fc2 rows per thread after the k-split reduce
    extern __shared__ __align__(128) unsigned char smem[];
    float4* w2m = reinterpret_cast<float4*>(smem + L::off_w2m);
    float* w1a = reinterpret_cast<float*>(smem + L::off_w1a);
    float4* stage = reinterpret_cast<float4*>(smem + L::off_stage);
    float* u = reinterpret_cast<float*>(smem + L::off_u);          // h1p | part | reductions
    float* small = reinterpret_cast<float*>(smem + L::off_small);
    float* small_b3 = small + 3 * SMALL_FLOATS;
    float* obs = reinterpret_cast<float*>(smem + L::off_obs);
    float* red1 = reinterpret_cast<float*>(smem + L::off_red1);
    float2* lnx = reinterpret_cast<float2*>(smem + L::off_lnx);    // [3][CL][BT]
    float* plog = reinterpret_cast<float*>(smem + L::off_plog);    // [3][CL][5][BT]
    int* act_s = reinterpret_cast<int*>(smem + L::off_act);        // [3][BT]
    float* gap_s = reinterpret_cast<float*>(act_s + 3 * BT);       // [3][BT]
    int* flag = reinterpret_cast<int*>(smem + L::off_flag);

    cg::cluster_group cluster = cg::this_cluster();
    const int q = (int)cluster.block_rank();
    const int cid = blockIdx.x / CL, ncl = gridDim.x / CL;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int e1 = t % BT, g1 = t / BT;
    const int ms = p.member_seat;
    const int n_chunks_e = (p.E + BT - 1) / BT;

    if (t == 0) *flag = 0;

    for (int m = cid; m < p.P; m += ncl) {
        const float* mrow = p.members + (int64_t)m * p.member_pitch;
        const FcOffsets om = fc_offsets(seat_in_dim(ms));
        __syncthreads();          // previous tile done with w2m / small
        // ---- member fc2 quarter -> resident smem (swizzled 16B chunks) --------
        {
            const float* w2q = mrow + om.fc2w + (size_t)q * ROWS_Q * H1;
            for (int f = t; f < ROWS_Q * H1 / 4; f += CT) {
                const int row = f >> 7, j16 = f & 127;
                cp_async16(w2m + row * 128 + (j16 ^ (row & 7)), w2q + (size_t)row * H1 + j16 * 4);
            }
            cp_async_commit();
        }
        for (int k = 0; k < p.K; ++k) {
            // per-seat flat rows for this (m, k)
            const float* wrow[3];
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                if (s == ms) wrow[s] = mrow;
                else {
                    const int oi = (s < ms) ? s : s - 1;
                    wrow[s] = p.opp[oi] + (int64_t)k * p.opp_pitch[oi];
                }
            }
            __syncthreads();      // previous (m,k) done with small[]
            // ---- small quarter blocks of the three seats ----------------------
            for (int s = 0; s < 3; ++s) {
                const FcOffsets o = fc_offsets(seat_in_dim(s));
                float* sm = small + s * SMALL_FLOATS;
                for (int i = t; i < SMALL_FLOATS; i += CT) {
                    float v;
                    if (i < 64) v = wrow[s][o.fc2b + q * ROWS_Q + i];
                    else if (i < 128) v = wrow[s][o.ln2g + q * ROWS_Q + (i - 64)];
                    else if (i < 192) v = wrow[s][o.ln2b + q * ROWS_Q + (i - 128)];
                    else {
                        const int a = (i - 192) / 64, r = (i - 192) % 64;
                        v = wrow[s][o.outw + a * H2 + q * ROWS_Q + r];
                    }
                    sm[i] = v;
                }
                if (t < NACT) small_b3[s * 8 + t] = wrow[s][o.outb + t];
            }
            for (int ec = 0; ec < n_chunks_e; ++ec) {
                const int e_base = ec * BT;
                const int n_env = min(BT, p.E - e_base);
                // ---- env state in registers of threads t < BT ------------------
                EnvState st;
                double sum_good = 0.0, last_good = 0.0, sum_adv = 0.0;
                float min_gap = CUDART_INF_F;
                int64_t ep_index = 0;
                if (t < BT) {
                    const int ei = e_base + (t < n_env ? t : 0);
                    ep_index = ((int64_t)m * p.K + k) * p.E + ei;
                    const int64_t rec = p.init_shared ? ((int64_t)k * p.E + ei) : ep_index;
                    env_load(st, p.init + rec * CEV_INIT_STATE_DIM);
#pragma unroll
                    for (int s = 0; s < 3; ++s) env_observe(st, s, obs + (s * BT + t) * 12);
                }
                // seat 0's W1 block
                issue_w1(w1a, wrow[0], seat_in_dim(0));
                cp_async_commit();
                cp_async_wait<0>();       // also covers the member fc2 quarter
                __syncthreads();

                for (int c = 0; c < p.n_cycles; ++c) {
                    float pre2[3][RP];
                    // =========== three independent forwards ====================
                    // (fully unrolled: pre2[s] must stay in registers)
#pragma unroll
                    for (int s = 0; s < 3; ++s) {
                        // ---- layer 1 (reads w1a, obs; writes h1p) --------------
                        if (s == 0) layer1<BT, IN_ADV>(w1a, obs + s * BT * 12, u, red1, flag);
                        else layer1<BT, IN_GOOD>(w1a, obs + s * BT * 12, u, red1, flag);
                        __syncthreads();          // h1p complete; w1a free
                        // ---- prefetch next seat's W1 block ---------------------
                        const int sn = (s + 1) % 3;
                        issue_w1(w1a, wrow[sn], seat_in_dim(sn));
                        // ---- fc2 quarter --------------------------------------
                        float2 acc[M::RPL][4];
#pragma unroll
                        for (int i = 0; i < M::RPL; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
                        if (s == ms) {
                            cp_async_commit();
#pragma unroll 2
                            for (int ch = 0; ch < NCHUNK; ++ch)
                                fc2_step<BT>(w2m, 128, ch * 8 + warp, u, ch * KCH + warp * 4, acc);
                            cp_async_wait<0>();
                        } else {
                            const FcOffsets o = fc_offsets(seat_in_dim(s));
                            const float* w2q = wrow[s] + o.fc2w + (size_t)q * ROWS_Q * H1;
                            issue_chunk(stage, w2q, 0);
                            cp_async_commit();                 // group: W1(next) + chunk 0
                            issue_chunk(stage + STAGE_F4, w2q, 1);
                            cp_async_commit();                 // group: chunk 1
#pragma unroll 1
                            for (int ch = 0; ch < NCHUNK; ++ch) {
                                cp_async_wait<NSTAGE - 2>();   // chunk ch has landed (this thread)
                                __syncthreads();               // ... for all threads; chunk ch-1 consumed
                                if (ch + 2 < NCHUNK)
                                    issue_chunk(stage + ((ch + 2) % NSTAGE) * STAGE_F4, w2q, ch + 2);
                                cp_async_commit();
                                fc2_step<BT>(stage + (ch % NSTAGE) * STAGE_F4, 8, warp, u,
                                             ch * KCH + warp * 4, acc);
                            }
                            cp_async_wait<0>();
                        }
                        __syncthreads();          // all warps done reading h1p; W1(next) landed
                        // ---- k-split partials -> part[warp][row][e] (aliases h1p)
                        {
                            const int eg = lane % M::GE, rl = lane / M::GE;
#pragma unroll
                            for (int i = 0; i < M::RPL; ++i) {
                                const int row = rl + M::RL * i;
                                float4 v;
                                v.x = acc[i][0].x + acc[i][0].y;
                                v.y = acc[i][1].x + acc[i][1].y;
                                v.z = acc[i][2].x + acc[i][2].y;
                                v.w = acc[i][3].x + acc[i][3].y;
                                *reinterpret_cast<float4*>(u + (warp * ROWS_Q + row) * BT + eg * 4) = v;
                            }
                        }
                        __syncthreads();
                        {
                            const float* b2 = small + s * SMALL_FLOATS;
#pragma unroll
                            for (int j = 0; j < RP; ++j) {
                                const int row = g1 + G * j;
                                float sacc = u[row * BT + e1];
#pragma unroll
                                for (int w = 1; w < NW; ++w) sacc += u[(w * ROWS_Q + row) * BT + e1];
                                pre2[s][j] = sacc + b2[row];
                            }
                        }
                        __syncthreads();          // part consumed; u free for the next seat
                    }
                    // =========== LayerNorm-2 statistics across the cluster =====
                    // local (64-row) mean and M2 per seat/env, then Chan combine.
                    float* redA = u;                       // [3][NW][BT]
                    float* redB = u + 3 * NW * BT;         // [3][NW][BT]
                    float lmean[3];
#pragma unroll
                    for (int s = 0; s < 3; ++s) {
                        float v = 0.f;
#pragma unroll
                        for (int j = 0; j < RP; ++j) v += pre2[s][j];
                        v = group_sum<BT>(v);
                        if (lane < BT) redA[(s * NW + warp) * BT + e1] = v;
                    }
                    __syncthreads();
#pragma unroll
                    for (int s = 0; s < 3; ++s) {
                        float tot = 0.f;
#pragma unroll
                        for (int w = 0; w < NW; ++w) tot += redA[(s * NW + w) * BT + e1];
                        lmean[s] = tot * (1.0f / ROWS_Q);
                        float v = 0.f;
#pragma unroll
                        for (int j = 0; j < RP; ++j) {
                            const float d = pre2[s][j] - lmean[s];
                            v = fmaf(d, d, v);
                        }
                        v = group_sum<BT>(v);
                        if (lane < BT) redB[(s * NW + warp) * BT + e1] = v;
                    }
                    __syncthreads();
                    if (t < 3 * BT) {
                        const int s = t / BT, e = t % BT;
                        float tm = 0.f, tq = 0.f;
#pragma unroll
                        for (int w = 0; w < NW; ++w) {
                            tm += redA[(s * NW + w) * BT + e];
                            tq += redB[(s * NW + w) * BT + e];
                        }
                        const float2 v = make_float2(tm * (1.0f / ROWS_Q), tq);
#pragma unroll
                        for (int r = 0; r < CL; ++r) {
                            float2* dst = cluster.map_shared_rank(lnx, r);
                            dst[(s * CL + q) * BT + e] = v;
                        }
                    }
                    cluster.sync();                        // barrier 1
                    // =========== normalise, partial logits ======================
                    float* redC = u + 6 * NW * BT;         // [3][NW][5][BT]
#pragma unroll
                    for (int s = 0; s < 3; ++s) {
                        float mc[CL], mu = 0.f, m2 = 0.f;
#pragma unroll
                        for (int r = 0; r < CL; ++r) {
                            const float2 v = lnx[(s * CL + r) * BT + e1];
                            mc[r] = v.x;
                            mu += v.x;
                            m2 += v.y;
                        }
                        mu *= (1.0f / CL);
#pragma unroll
                        for (int r = 0; r < CL; ++r) {
                            const float d = mc[r] - mu;
                            m2 = fmaf((float)ROWS_Q * d, d, m2);
                        }
                        const float var = m2 * (1.0f / H2);
                        if (!isfinite(mu) || !isfinite(var)) *flag = 1;
                        const float rstd = 1.0f / sqrtf(var + LN_EPS);
                        const float* sm = small + s * SMALL_FLOATS;
                        float pl[NACT] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int j = 0; j < RP; ++j) {
                            const int row = g1 + G * j;
                            const float h = fmaxf(fmaf((pre2[s][j] - mu) * rstd, sm[64 + row], sm[128 + row]), 0.f);
#pragma unroll
                            for (int a = 0; a < NACT; ++a) pl[a] = fmaf(sm[192 + a * 64 + row], h, pl[a]);
                        }
#pragma unroll
                        for (int a = 0; a < NACT; ++a) {
                            const float v = group_sum<BT>(pl[a]);
                            if (lane < BT) redC[((s * NW + warp) * NACT + a) * BT + e1] = v;
                        }
                    }
                    __syncthreads();
                    if (t < 3 * NACT * BT) {
                        const int s = t / (NACT * BT), a = (t / BT) % NACT, e = t % BT;
                        float tot = 0.f;
#pragma unroll
                        for (int w = 0; w < NW; ++w) tot += redC[((s * NW + w) * NACT + a) * BT + e];
#pragma unroll
                        for (int r = 0; r < CL; ++r) {
                            float* dst = cluster.map_shared_rank(plog, r);
                            dst[((s * CL + q) * NACT + a) * BT + e] = tot;
                        }
                    }
                    cluster.sync();                        // barrier 2
                    // =========== logits, argmax, physics ========================
                    if (t < 3 * BT) {
                        const int s = t / BT, e = t % BT;
                        float lg[NACT];
                        bool fin = true;
#pragma unroll
                        for (int a = 0; a < NACT; ++a) {
                            float v = plog[((s * CL + 0) * NACT + a) * BT + e];
#pragma unroll
                            for (int r = 1; r < CL; ++r) v += plog[((s * CL + r) * NACT + a) * BT + e];
                            lg[a] = v + small_b3[s * 8 + a];
                            fin = fin && isfinite(lg[a]);
                        }
                        if (!fin) *flag = 1;
                        float gap;
                        act_s[s * BT + e] = argmax_first5(lg, gap);
                        gap_s[s * BT + e] = gap;
                    }
                    __syncthreads();
                    if (t < BT) {
                        const int act[3] = {act_s[t], act_s[BT + t], act_s[2 * BT + t]};
                        min_gap = fminf(min_gap, fminf(gap_s[t], fminf(gap_s[BT + t], gap_s[2 * BT + t])));
                        double rg, ra;
                        env_step(st, act, p.pos_first != 0, rg, ra);
                        sum_good = __dadd_rn(sum_good, rg);
                        sum_adv = __dadd_rn(sum_adv, ra);
                        last_good = rg;
#pragma unroll
                        for (int s = 0; s < 3; ++s) env_observe(st, s, obs + (s * BT + t) * 12);
                    }
                    __syncthreads();
                }
                // ---- tile epilogue ---------------------------------------------
                if (q == 0 && t < n_env) {
                    double* o = p.out + ep_index * CEV_ROLLOUT_OUT_DIM;
                    o[0] = sum_good;
                    o[1] = last_good;
                    o[2] = sum_adv;
                    o[3] = (double)min_gap;
                }
                cp_async_wait<0>();
                __syncthreads();
            }
        }
    }
    __syncthreads();
    if (q == 0 && t == 0 && *flag && p.status) atomicOr(p.status, CEV_STATUS_NONFINITE);
    // no CTA may exit while a peer can still write its shared memory
    cluster.sync();
}

template <int BT>
static int launch_bt(cev_handle* h, const ClusterParams& p, cudaStream_t stream) {
    using L = SmemLayout<BT>;
    auto kern = rollout_cluster_kernel<BT>;
    static bool configured[16] = {};
    int dev = h->device;
    if (dev < 16 && !configured[dev]) {
        CEV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total));
        configured[dev] = true;
    }
    int ncl = h->n_clusters;
    if (ncl > p.P) ncl = p.P;
    if (ncl < 1) ncl = 1;
    kern<<<ncl * CL, CT, L::total, stream>>>(p);
    return check_cuda(cudaGetLastError(), "rollout_cluster_kernel launch");
}

int rollout_cluster_max_clusters(int device) {
    // co-resident 4-CTA clusters at the largest smem footprint
    using L = SmemLayout<16>;
    auto kern = rollout_cluster_kernel<16>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CL * 64);
    cfg.blockDim = dim3(CT);
    cfg.dynamicSmemBytes = L::total;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    (void)device;
    return n;
}

int launch_rollout_cluster(cev_handle* h, const ClusterParams& p, cudaStream_t stream) {
    if (p.P <= 0 || p.K <= 0 || p.E <= 0) return CEV_OK;
    if (p.E > 8) return launch_bt<16>(h, p, stream);
    if (p.E > 4) return launch_bt<8>(h, p, stream);
    return launch_bt<4>(h, p, stream);
}

}  // namespace cev
