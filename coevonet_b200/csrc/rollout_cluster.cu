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
//  * Opponent weights (shared by every tile, L2-resident) are repacked once per
//    launch into pre-swizzled 4 KB chunk images and streamed by TMA bulk copies
//    (cp.async.bulk + mbarrier complete_tx) into 12 slots (6 dedicated + 6 borrowed
//    from the idle W1 arena).  Each 16-k chunk has exactly one consumer warp, which
//    refills the slot it drained; two alternating "full" mbarriers per slot keep the
//    parity waits unambiguous.  No CTA-wide barrier inside the fc2 loop.
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

#include "rollout_common.cuh"

namespace cg = cooperative_groups;

#ifdef CEV_PROFILE
// development-only phase timers (thread 0 of the first CTA): cycles per phase
__device__ unsigned long long cev_prof_acc[16];
__device__ int cev_debug_flags;     // bit0: do not wait for streamed data; bit1: no copies at all
#define DBG_FLAG(b) (dbg_flags & (b))
#define DBG_LOAD const int dbg_flags = cev_debug_flags;
#define PROF_DECL long long _pt = clock64();
#define PROF(i)                                                     \
    do {                                                            \
        if (threadIdx.x == 0 && blockIdx.x == 0) {                  \
            const long long _n = clock64();                         \
            cev_prof_acc[i] += (unsigned long long)(_n - _pt);      \
            _pt = _n;                                               \
        }                                                           \
    } while (0)
#else
#define PROF_DECL
#define PROF(i)
#define DBG_FLAG(b) 0
#define DBG_LOAD
#endif

namespace cev {

constexpr int CL = 4;                       // CTAs per cluster
constexpr int ROWS_Q = H2 / CL;             // fc2 rows per CTA (64)
constexpr int KCH = 16;                     // k per chunk: one warp owns a whole chunk (4 steps of 4 k)
constexpr int NCHUNK = H1 / KCH;            // 32 chunks per fc2 pass, warp w owns chunks w, w+8, w+16, w+24
constexpr int NSLOT_D = 6;                  // dedicated 4 KB slots (region `stage`)
constexpr int NSLOT = 12;                   // + 6 slots borrowed from the W1 arena during a streamed pass
constexpr int STAGE_F4 = ROWS_Q * KCH / 4;  // float4 per slot (256 = 4 KB)
constexpr int W1A_FLOATS = H1 * IN_GOOD + 3 * H1;   // fc1.W | fc1.b | ln1.g | ln1.b
constexpr int SMALL_FLOATS = 512;           // b2[64] g2[64] be2[64] w3[5][64]

template <int BT>
struct SmemLayout {
    static constexpr size_t off_w2m = 0;
    static constexpr size_t off_w1a = off_w2m + (size_t)ROWS_Q * H1 * 4;
    static constexpr size_t off_stage = off_w1a + (size_t)W1A_FLOATS * 4;
    static constexpr size_t off_u = off_stage + (size_t)NSLOT_D * STAGE_F4 * 16;
    static constexpr size_t off_small = off_u + (size_t)H1 * BT * 4;
    static constexpr size_t off_obs = off_small + (size_t)(3 * SMALL_FLOATS + 24) * 4;
    static constexpr size_t off_red1 = off_obs + (size_t)3 * BT * 12 * 4;
    static constexpr size_t off_lnx = off_red1 + (size_t)2 * NW * BT * 4;
    static constexpr size_t off_plog = off_lnx + (size_t)3 * CL * BT * 8;
    static constexpr size_t off_act = off_plog + (size_t)3 * CL * NACT * BT * 4;
    static constexpr size_t off_flag = off_act + (size_t)3 * BT * 8;
    static constexpr size_t off_bar = off_flag + 16;             // full[NSLOT][2], w1_full
    static constexpr size_t total = off_bar + 256;
};

// ---------------------------------------------------------------------------
// fc2 quarter: this CTA's 64 rows x 512 k x BT envs.
// lane = (eg = lane % GE env group of 4, rl = lane / GE row lane);
// rows rl + RL*i, i < RPL.  Warp w takes the w-th 4-k step of every 32-k
// chunk.  Accumulators are float2 (even-k, odd-k partial sums).
// ---------------------------------------------------------------------------
template <int BT>
struct Fc2Map {
    static constexpr int GE = BT / 4;              // env groups of 4 per warp
    static constexpr int RL = 32 / GE;             // row lanes
    static constexpr int RPL = ROWS_Q / RL;        // rows per lane (all 64 rows of the CTA's quarter)
};

// One 16-k chunk (4 steps of 4 k) of the fc2 quarter for ONE warp: rows rl + RL*i,
// k = kbase .. kbase+15, all BT envs.  Accumulators are float2 (even-k, odd-k sums).
//  STAGED = false: resident operand, row stride 128 float4, 16-byte column c16_0 + step,
//                  swizzled by (row & 7);
//  STAGED = true : 4 KB slot image, row stride 4 float4, column = step, swizzled by
//                  ((row >> 1) & 3) (8 consecutive rows -> 8 distinct 16-byte bank groups).
template <int BT, bool STAGED>
__device__ __forceinline__ void fc2_chunk(const float4* __restrict__ wbase, int c16_0,
                                          const float* __restrict__ h1p, int kbase,
                                          float2 (&acc)[Fc2Map<BT>::RPL][4]) {
    using M = Fc2Map<BT>;
    constexpr int RS = STAGED ? 4 : 128;         // row stride in float4
    const int lane = threadIdx.x & 31;
    const int eg = lane % M::GE, rl = lane / M::GE;
    const int sw = STAGED ? ((rl >> 1) & 3) : (rl & 7);
    const float4* wrow0 = wbase + rl * RS;
    const float* hbase = h1p + (kbase >> 1) * (2 * BT) + eg * 8;
#pragma unroll
    for (int st = 0; st < KCH / 4; ++st) {
        // activations: two k-pairs x two env-pairs
        float4 a[2][2];
#pragma unroll
        for (int kp = 0; kp < 2; ++kp)
#pragma unroll
            for (int j = 0; j < 2; ++j)
                a[kp][j] = *reinterpret_cast<const float4*>(hbase + (2 * st + kp) * (2 * BT) + j * 4);
        const int col = (c16_0 + st) ^ sw;
#pragma unroll
        for (int i = 0; i < M::RPL; ++i) {
            const float4 w = wrow0[i * M::RL * RS + col];
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
}

constexpr uint32_t STAGE_BYTES = STAGE_F4 * 16;
constexpr int PACKED_F4_PER_ROW = CL * NCHUNK * STAGE_F4;     // float4 per packed fc2 matrix (32768)

// Opponent fc2 matrices repacked into the exact shared-memory stage images:
// packed[((q*NCHUNK + chunk)*64 + row)*4 + (j16 ^ ((row>>1)&3))] = W2[64q+row][16*chunk + 4*j16 .. +3]
__global__ void __launch_bounds__(256) pack_opponent_kernel(const float* __restrict__ rows, int64_t pitch,
                                                            int in_dim, int K, float4* __restrict__ packed) {
    const FcOffsets o = fc_offsets(in_dim);
    const int k = blockIdx.y;
    const float* w2 = rows + (int64_t)k * pitch + o.fc2w;
    float4* dst = packed + (int64_t)k * PACKED_F4_PER_ROW;
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < PACKED_F4_PER_ROW; f += gridDim.x * blockDim.x) {
        const int j16 = f & 3, row = (f >> 2) & 63, chunk = (f >> 8) & (NCHUNK - 1), q = f >> 13;
        const float4 v = *reinterpret_cast<const float4*>(w2 + (size_t)(q * ROWS_Q + row) * H1 + chunk * KCH + j16 * 4);
        dst[((q * NCHUNK + chunk) * ROWS_Q + row) * 4 + (j16 ^ ((row >> 1) & 3))] = v;
    }
    (void)K;
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
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + L::off_bar);   // [NSLOT][2]
    uint64_t* bar_w1 = bar_full + 2 * NSLOT;

    cg::cluster_group cluster = cg::this_cluster();
    const int q = (int)cluster.block_rank();
    const int cid = blockIdx.x / CL, ncl = gridDim.x / CL;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int e1 = t % BT, g1 = t / BT;
    const int ms = p.member_seat;
    const int n_chunks_e = (p.E + BT - 1) / BT;

    if (t == 0) {
        *flag = 0;
        for (int i = 0; i < 2 * NSLOT; ++i) mbar_init(bar_full + i, 1);   // one arrive.expect_tx per fill
        mbar_init(bar_w1, 1);
        mbar_fence_init();
    }
    __syncthreads();
    DBG_LOAD
    uint32_t sp = 0;       // streamed fc2 passes completed so far (identical in every thread)
    uint32_t w1n = 0;      // W1 blocks loaded so far

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
            const float4* wpack[3];     // this CTA's quarter of the packed opponent fc2 matrix
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                if (s == ms) {
                    wrow[s] = mrow;
                    wpack[s] = nullptr;
                } else {
                    const int oi = (s < ms) ? s : s - 1;
                    wrow[s] = p.opp[oi] + (int64_t)k * p.opp_pitch[oi];
                    wpack[s] = p.opp_packed[oi] + (int64_t)k * PACKED_F4_PER_ROW + (size_t)q * NCHUNK * STAGE_F4;
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
                // seat 0's W1 block (contiguous in the flat row): one bulk copy
                if (t == 0 && p.n_cycles > 0) {
                    const uint32_t bytes = (uint32_t)(H1 * seat_in_dim(0) + 3 * H1) * 4;
                    mbar_arrive_expect_tx(bar_w1, bytes);
                    bulk_g2s(w1a, wrow[0], bytes, bar_w1);
                }
                cp_async_wait<0>();       // the member fc2 quarter
                __syncthreads();

                PROF_DECL
                for (int c = 0; c < p.n_cycles; ++c) {
                    float pre2[3][RP];
                    PROF(0);
                    // =========== three independent forwards ====================
                    // (fully unrolled: pre2[s] must stay in registers)
#pragma unroll
                    for (int s = 0; s < 3; ++s) {
                        // ---- streamed seats: a slot is one 4 KB chunk image; chunk ch lives in slot
                        //      ch % 12 (6 dedicated + 6 borrowed from the W1 arena once layer 1 is done)
                        const float4* src = wpack[s];
                        // Successive uses of a slot have DIFFERENT consumer warps, and a parity wait can
                        // only tell "current phase done or not": with one barrier per slot, a warp running
                        // a round ahead would pass its wait for use u+1 while use u is still in flight.
                        // Two barriers per slot, alternating by use, make every wait unambiguous (the
                        // same warp consumes uses u and u+2 of a slot).
                        auto issue = [&](int ch) {
                            const int sl = ch % NSLOT;
                            const uint32_t use = sp * (sl < 8 ? 3u : 2u) + (uint32_t)(ch / NSLOT);
                            uint64_t* bar = bar_full + 2 * sl + (use & 1);
                            float4* dst = sl < NSLOT_D ? stage + sl * STAGE_F4
                                                       : reinterpret_cast<float4*>(w1a) + (sl - NSLOT_D) * STAGE_F4;
                            mbar_arrive_expect_tx(bar, STAGE_BYTES);
                            bulk_g2s(dst, src + (size_t)ch * STAGE_F4, STAGE_BYTES, bar);
                        };
                        auto issue_w1 = [&](int seat) {
                            const uint32_t bytes = (uint32_t)(H1 * seat_in_dim(seat) + 3 * H1) * 4;
                            mbar_arrive_expect_tx(bar_w1, bytes);
                            bulk_g2s(w1a, wrow[seat], bytes, bar_w1);
                        };
                        const int sn = (s + 1) % 3;
                        const bool more = !(s == 2 && c == p.n_cycles - 1);   // another forward follows
                        if (s != ms && t == 0 && !DBG_FLAG(2)) {
                            // chunks 0..5 land in the dedicated slots while layer 1 runs
#pragma unroll
                            for (int ch = 0; ch < NSLOT_D; ++ch) issue(ch);
                        }
                        // ---- layer 1 (reads w1a, obs; writes h1p) --------------
                        mbar_wait(bar_w1, w1n);
                        ++w1n;
                        PROF(1);
                        if (s == 0) layer1<BT, IN_ADV>(w1a, obs + s * BT * 12, u, red1, flag);
                        else layer1<BT, IN_GOOD>(w1a, obs + s * BT * 12, u, red1, flag);
                        __syncthreads();          // h1p complete; the W1 arena is free
                        PROF(2);
                        // ---- fc2 quarter: warp w owns the 16-k chunks w, w+8, w+16, w+24 ------
                        float2 acc[M::RPL][4];
#pragma unroll
                        for (int i = 0; i < M::RPL; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
                        if (s == ms) {
                            // resident operand; the arena is idle, prefetch the next seat's W1 block
                            if (t == 0 && more) issue_w1(sn);
#pragma unroll
                            for (int r = 0; r < NCHUNK / NW; ++r) {
                                const int ch = warp + NW * r;
                                fc2_chunk<BT, false>(w2m, ch * (KCH / 4), u, ch * KCH, acc);
                            }
                            PROF(3);
                        } else {
                            // Streamed operand through TMA bulk copies.  A chunk has exactly ONE
                            // consumer warp, so there are no "empty" barriers and no warp ever waits
                            // for a peer: after its 256 FFMA2 on chunk ch, the owning warp's lane 0
                            // refills the slot it just drained with chunk ch+12 (owned by another
                            // warp, which waits on the slot's "full" mbarrier).
                            if (t == 0 && !DBG_FLAG(2)) {
#pragma unroll
                                for (int ch = NSLOT_D; ch < NSLOT; ++ch) issue(ch);
                            }
#pragma unroll 1
                            for (int r = 0; r < NCHUNK / NW; ++r) {
                                const int ch = warp + NW * r;
                                const int sl = ch % NSLOT;
                                const uint32_t use = sp * (sl < 8 ? 3u : 2u) + (uint32_t)(ch / NSLOT);
                                if (!DBG_FLAG(3)) mbar_wait(bar_full + 2 * sl + (use & 1), use >> 1);
                                const float4* wst = sl < NSLOT_D ? stage + sl * STAGE_F4
                                                                 : reinterpret_cast<const float4*>(w1a) + (sl - NSLOT_D) * STAGE_F4;
                                fc2_chunk<BT, true>(wst, 0, u, ch * KCH, acc);
                                __syncwarp();
                                if (lane == 0 && ch + NSLOT < NCHUNK && !DBG_FLAG(2)) issue(ch + NSLOT);
                            }
                            ++sp;
                            PROF(4);
                        }
                        __syncthreads();          // all warps done reading h1p and the arena slots
                        if (s != ms && t == 0 && more) issue_w1(sn);
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
                        PROF(5);
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
                    PROF(6);
                    cluster.sync();                        // barrier 1
                    PROF(7);
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
                    PROF(8);
                    cluster.sync();                        // barrier 2
                    PROF(9);
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

int launch_rollout_cluster(cev_handle* h, const ClusterParams& p_in, cudaStream_t stream) {
    if (p_in.P <= 0 || p_in.K <= 0 || p_in.E <= 0) return CEV_OK;
    ClusterParams p = p_in;
    // repack the (few) opponent fc2 matrices into stage images: 512 KB per row and seat
    const size_t per_seat = (size_t)p.K * PACKED_F4_PER_ROW * sizeof(float4);
    if (h->opp_workspace_bytes < 2 * per_seat) {
        if (h->opp_workspace) CEV_CUDA(cudaFree(h->opp_workspace));
        h->opp_workspace = nullptr;
        h->opp_workspace_bytes = 0;
        CEV_CUDA(cudaMalloc(&h->opp_workspace, 2 * per_seat));
        h->opp_workspace_bytes = 2 * per_seat;
    }
    for (int oi = 0; oi < 2; ++oi) {
        const int seat = (oi < p.member_seat) ? oi : oi + 1;
        float4* dst = reinterpret_cast<float4*>(static_cast<char*>(h->opp_workspace) + oi * per_seat);
        pack_opponent_kernel<<<dim3(32, p.K), 256, 0, stream>>>(p.opp[oi], p.opp_pitch[oi], seat_in_dim(seat), p.K,
                                                               dst);
        p.opp_packed[oi] = dst;
    }
    CEV_CUDA(cudaGetLastError());
    if (p.E > 8) return launch_bt<16>(h, p, stream);
    if (p.E > 4) return launch_bt<8>(h, p, stream);
    return launch_bt<4>(h, p, stream);
}

}  // namespace cev

#ifdef CEV_PROFILE
// development-only: read (and optionally reset) the phase timers
extern "C" int cev_debug_profile(double* out, int n, int reset) {
    unsigned long long h[16];
    if (cudaMemcpyFromSymbol(h, cev_prof_acc, sizeof(h)) != cudaSuccess) return -1;
    for (int i = 0; i < n && i < 16; ++i) out[i] = (double)h[i];
    if (reset) {
        unsigned long long z[16] = {};
        cudaMemcpyToSymbol(cev_prof_acc, z, sizeof(z));
    }
    return 0;
}
extern "C" int cev_debug_set_flags(int flags) {
    return cudaMemcpyToSymbol(cev_debug_flags, &flags, sizeof(int)) == cudaSuccess ? 0 : -1;
}
#endif
