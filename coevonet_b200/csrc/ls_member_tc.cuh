// K1 (lockstep form), member forward on the tensor cores.  Included by rollout_lockstep.cu.
//
// The member's own forward (FCNetwork.forward, MPE/fcnetwork.py:37-70, with the member's OWN weights:
// evolutionary_strategy.py:63-116 mutate_weights, genetic_algorithm.py:125-217) is, per member and world step,
// H[256, 16 episodes] = W2_m[256, 512] . h1[16, 512]^T with 512 KB of W2_m that nobody else reads: the stage is
// bound by the HBM stream of the member rows, and a CTA that only streams does not need the FP32 pipe of its SM.
// This form therefore keeps the FMA pipe out of it (3xTF32 on tcgen05, the structure of the K2 fc stage,
// deepqn_tc.cu) and runs as a PERSISTENT kernel on a subset of the SMs, so that the opponent kernel
// (tensor-pipe bound, ls_opp_kernel) runs beside it on the other SMs instead of after it.
//
// One CTA per SM, job = (member, block of 16 episodes).  Per k-tile of 32:
//   warp 0      TMA: the raw W2 tile [256 rows x 32 k] (3-D tensor map (k, row, member), 128B swizzle, 256B L2
//               promotion) into a 5-slot ring, with the k-tile's 32 rows of the member's layer 1 (fc1.W | fc1.b |
//               ln1.g | ln1.b, 1.7 KB) behind it; per job the tail block (fc2.b | ln2.g | ln2.b | out.W | out.b)
//   warps 2-9   W2 lo = w - trunc_tf32(w) -> TENSOR MEMORY (tcgen05.st; the MMA takes it as a TMEM A operand), and
//               the activations of this k-tile: relu(LN1(W1 x + b1)) for the 16 episodes, split into hi | lo and
//               stored K-major behind the W2 tile, so [x_hi | x_lo] is ONE 32-row N operand.  LayerNorm-1 uses
//               the closed-form row statistics (mean = wbar . z, var = z^T C z, fp64; ls_l1stats_kernel over
//               the member rows, once per rollout), so a k-tile never needs the whole 512-vector
//   warps 1,14  16 x tcgen05.mma.cta_group::1.kind::tf32 per k-tile (2 row tiles, one per issuing warp, x 4 k-steps x
//               {w_hi . [x_hi | x_lo] (N = 32), w_lo (TMEM) . x_hi (N = 16)}) into ping-pong accumulators
//   warps 10-13 epilogue: tcgen05.ld, + fc2.b into shared memory, LayerNorm-2 + ReLU + output layer + first-max
//               argmax (MPE/fcnetwork.py:53-90); then the observations and LayerNorm-1 statistics of the job
//               after the next one, which the producers pick up from shared memory
#pragma once

namespace cev {

constexpr int MT_THREADS = 480;                               // warp 0 TMA, 1 + 14 MMA, 2-9 producers, 10-13 epilogue
constexpr int MT_PROD_WARPS = 8;
constexpr int MT_BK = 32, MT_KT = H1 / MT_BK;                 // 16 k-tiles of 128 bytes
constexpr int MT_NB = LS_BT;                                  // 16 episodes = MMA N
constexpr int MT_R = 5, MT_L = 4;                             // raw slots (shared memory), W2-lo slots (tensor memory)
constexpr uint32_t MT_A_BYTES = H2 * MT_BK * 4;               // 32 KB
constexpr uint32_t MT_X_BYTES = MT_NB * MT_BK * 4;            // 2 KB
constexpr uint32_t MT_SLOT = MT_A_BYTES + 2 * MT_X_BYTES;     // W2 raw | x hi | x lo = 36 KB
constexpr uint32_t MT_W1C_FLOATS = MT_BK * IN_GOOD + 3 * MT_BK;   // layer-1 chunk of a k-tile: fc1.W rows | fc1.b | ln1.g | ln1.b
constexpr uint32_t MT_W1C_BYTES = MT_W1C_FLOATS * 4;          // 1,664
constexpr uint32_t MT_TAIL_BYTES = LS_TAIL_FLOATS * 4;        // 8,224
constexpr size_t MT_OFF_W1C = (size_t)MT_R * MT_SLOT;         // [R] layer-1 chunks, one per raw slot
constexpr size_t MT_OFF_HS = MT_OFF_W1C + (size_t)MT_R * MT_W1C_BYTES;   // [256 rows][16 episodes] fp32
constexpr size_t MT_OFF_TAIL = MT_OFF_HS + (size_t)H2 * MT_NB * 4;
constexpr size_t MT_OFF_OBS = MT_OFF_TAIL + MT_TAIL_BYTES;    // [2][16][12] fp32
constexpr size_t MT_OFF_STAT = MT_OFF_OBS + 2 * MT_NB * LS_OBS_PAD * 4;   // [2][16] (mean, rstd)
constexpr size_t MT_OFF_RED = MT_OFF_STAT + 2 * MT_NB * 8;    // redA [4][16] | redB [4][16] | redC [4][5][16]
constexpr size_t MT_OFF_BAR = MT_OFF_RED + (size_t)(2 + NACT) * 4 * MT_NB * 4;
constexpr size_t MT_SMEM = MT_OFF_BAR + 256 + 1024 /*alignment*/;
constexpr int MT_ACC = 3 * MT_NB;                             // accumulator columns per row tile: hh | hl | lh
constexpr int MT_TM_LO = 2 * 2 * MT_ACC;                      // first TMEM column of the W2-lo slots
constexpr int MT_TM_LSLOT = 2 * MT_BK;                        // columns per lo slot: 2 row tiles x 32 k
static_assert(MT_TM_LO + MT_L * MT_TM_LSLOT <= 512, "tensor memory budget");
static_assert(MT_SLOT % 1024 == 0, "slots must keep the 1024-byte alignment of the swizzle atoms");
static_assert(MT_OFF_W1C % 16 == 0 && MT_W1C_BYTES % 16 == 0 && MT_OFF_TAIL % 16 == 0 && MT_OFF_BAR % 8 == 0,
              "bulk-copy alignment");
static_assert(MT_SMEM <= 232448, "member stage exceeds the 227 KB shared-memory limit");
constexpr uint32_t MT_IDESC32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * MT_NB) >> 3) << 17) |
                                ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t MT_IDESC16 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(MT_NB >> 3) << 17) |
                                ((uint32_t)(128 >> 4) << 24);

// The eight MMAs of one row tile and k-tile (4 k-steps x {w_hi . [x_hi | x_lo] (N = 32), w_lo (TMEM) . x_hi (N = 16)})
// plus the two commits that release the raw slot and the lo slot, issued by one elected lane of a converged warp.
__device__ __forceinline__ void mt_issue_ktile(uint32_t d_tmem, uint64_t w_hi, uint64_t x_hilo, uint32_t lo_tmem,
                                               uint32_t accumulate, uint32_t bar_raw, uint32_t bar_lo) {
    asm volatile(
        "{\n"
        ".reg .pred pe, pa, pt;\n"
        ".reg .b64 a1, a2, a3, b1, b2, b3;\n"
        ".reg .b32 l1, l2, l3, dl;\n"
        "elect.sync _|pe, 0xffffffff;\n"
        "setp.ne.b32 pa, %4, 0;\n"
        "setp.eq.b32 pt, 0, 0;\n"
        "add.s64 a1, %1, 2;\n add.s64 a2, %1, 4;\n add.s64 a3, %1, 6;\n"
        "add.s64 b1, %2, 2;\n add.s64 b2, %2, 4;\n add.s64 b3, %2, 6;\n"
        "add.u32 l1, %3, 8;\n add.u32 l2, %3, 16;\n add.u32 l3, %3, 24;\n"
        "add.u32 dl, %0, 32;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %7, pa;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [dl], [%3], %2, %8, pa;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], a1, b1, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [dl], [l1], b1, %8, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], a2, b2, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [dl], [l2], b2, %8, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], a3, b3, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [dl], [l3], b3, %8, pt;\n"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n"
        "}\n" ::"r"(d_tmem),
        "l"(w_hi), "l"(x_hilo), "r"(lo_tmem), "r"(accumulate), "r"(bar_raw), "r"(bar_lo), "r"(MT_IDESC32), "r"(MT_IDESC16)
        : "memory");
}
__device__ __forceinline__ void mt_commit_elected(uint32_t bar) {
    asm volatile(
        "{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\n"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(bar)
        : "memory");
}

struct LsMemberTcParams {
    const float* members;
    int64_t pitch;
    int n_chunks, KE, seat, n_jobs;
    int64_t N;
    const float* obs;          // this seat's [N][12]
    int32_t* act;              // this seat's [N]
    float* gap;
    const double* l1stats;     // [P][132]
    int32_t* status;
    float* logits;             // this seat's [N][5] of this cycle (parity instrumentation; null in production)
};

template <int IN>
__global__ void __launch_bounds__(MT_THREADS, 1)
ls_member_tc_kernel(const __grid_constant__ CUtensorMap map_w, const LsMemberTcParams p) {
    extern __shared__ unsigned char mt_raw[];
    unsigned char* base = mt_raw + ((1024u - (tc_smem_u32(mt_raw) & 1023u)) & 1023u);
    unsigned char* raw_mem = base;                                    // R x (W2 raw | x hi | x lo)
    float* hs = reinterpret_cast<float*>(base + MT_OFF_HS);
    unsigned char* w1c_mem = base + MT_OFF_W1C;                       // [R][fc1.W rows of the k-tile | fc1.b | ln1.g | ln1.b]
    float* tail = reinterpret_cast<float*>(base + MT_OFF_TAIL);
    float* obsbuf = reinterpret_cast<float*>(base + MT_OFF_OBS);
    float2* statbuf = reinterpret_cast<float2*>(base + MT_OFF_STAT);
    float* redA = reinterpret_cast<float*>(base + MT_OFF_RED);
    float* redB = redA + 4 * MT_NB;
    float* redC = redB + 4 * MT_NB;
    uint64_t* bar_raw_full = reinterpret_cast<uint64_t*>(base + MT_OFF_BAR);
    uint64_t* bar_raw_empty = bar_raw_full + MT_R;
    uint64_t* bar_lo_full = bar_raw_empty + MT_R;
    uint64_t* bar_lo_empty = bar_lo_full + MT_L;
    uint64_t* bar_tfull = bar_lo_empty + MT_L;        // [2] accumulators of a job ready
    uint64_t* bar_tempty = bar_tfull + 2;             // [2] accumulators drained
    uint64_t* bar_stat_full = bar_tempty + 2;         // [2] observations + LayerNorm-1 statistics of a job written
    uint64_t* bar_tail_full = bar_stat_full + 2;
    uint64_t* bar_tail_empty = bar_tail_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tail_empty + 1);
    int* flag = reinterpret_cast<int*>(tmem_slot + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int job0 = (int)blockIdx.x, job_stride = (int)gridDim.x;
    constexpr uint32_t W1ROWS_BYTES = (uint32_t)(MT_BK * IN) * 4;         // fc1.W rows of one k-tile
    const FcOffsets om = fc_offsets(IN);

    if (threadIdx.x == 0) {
        *flag = 0;
        for (int i = 0; i < MT_R; ++i) {
            tc_mbar_init(bar_raw_full + i, 1);
            tc_mbar_init(bar_raw_empty + i, 2);        // one commit per MMA warp
        }
        for (int i = 0; i < MT_L; ++i) {
            tc_mbar_init(bar_lo_full + i, MT_PROD_WARPS / 2);      // one arrive per producer warp of the k-tile's group
            tc_mbar_init(bar_lo_empty + i, 2);
        }
        for (int i = 0; i < 2; ++i) {
            tc_mbar_init(bar_tfull + i, 2);
            tc_mbar_init(bar_tempty + i, 4);           // one arrive per epilogue warp
            tc_mbar_init(bar_stat_full + i, 1);
        }
        tc_mbar_init(bar_tail_full, 1);
        tc_mbar_init(bar_tail_empty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) {
        // 2 (ping-pong) x 2 (row tiles) x 48 accumulator columns = 192, then 4 W2-lo slots x 64 columns -> 448 of 512
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tc_smem_u32(tmem_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA: layer-1 block, raw W2 tiles, tail block =====================
        if (lane == 0) {
            uint32_t it = 0, js = 0;
            for (int job = job0; job < p.n_jobs; job += job_stride, ++js) {
                const int m = job / p.n_chunks;
                const float* mrow = p.members + (int64_t)m * p.pitch;
                for (int kt = 0; kt < MT_KT; ++kt, ++it) {
                    const uint32_t rs = it % MT_R, use = it / MT_R;
                    if (use > 0) tc_mbar_wait(bar_raw_empty + rs, (use - 1) & 1);
                    tc_mbar_expect_tx(bar_raw_full + rs, MT_A_BYTES + W1ROWS_BYTES + 3 * MT_BK * 4);
                    tma_load_3d(raw_mem + (size_t)rs * MT_SLOT, &map_w, bar_raw_full + rs, kt * MT_BK, 0, m);
                    float* c = reinterpret_cast<float*>(w1c_mem + (size_t)rs * MT_W1C_BYTES);
                    bulk_g2s(c, mrow + kt * MT_BK * IN, W1ROWS_BYTES, bar_raw_full + rs);
                    bulk_g2s(c + MT_BK * IN, mrow + om.fc1b + kt * MT_BK, MT_BK * 4, bar_raw_full + rs);
                    bulk_g2s(c + MT_BK * IN + MT_BK, mrow + om.ln1g + kt * MT_BK, MT_BK * 4, bar_raw_full + rs);
                    bulk_g2s(c + MT_BK * IN + 2 * MT_BK, mrow + om.ln1b + kt * MT_BK, MT_BK * 4, bar_raw_full + rs);
                }
                if (js >= 1) tc_mbar_wait(bar_tail_empty, (js - 1) & 1);      // the previous job's epilogue has read it
                tc_mbar_expect_tx(bar_tail_full, MT_TAIL_BYTES);
                bulk_g2s(tail, mrow + om.fc2b, MT_TAIL_BYTES, bar_tail_full);
            }
        }
    } else if (warp == 1 || warp == 14) {
        // ===================== MMA issuers: warp 1 = W2 rows 0..127, warp 14 = rows 128..255 =====================
        // A tcgen05.mma of this size costs ~45 cycles of the tensor pipe whatever N <= 64 is (measured,
        // scripts/probe/mma_probe.cu) plus ~60 issue-side instructions, so the 16 instructions of a k-tile are
        // shared out between two issuing threads; every operand is computed warp-uniformly outside the elected branch.
        const int t = warp == 1 ? 0 : 1;
        const uint32_t raw_base = tc_smem_u32(raw_mem);
        uint32_t it = 0, js = 0;
        for (int job = job0; job < p.n_jobs; job += job_stride, ++js) {
            const uint32_t as = js & 1, ause = js >> 1;
            if (ause > 0) tc_mbar_wait(bar_tempty + as, (ause - 1) & 1);      // epilogue drained it
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t d_tmem = tmem_base + as * (2 * MT_ACC) + t * MT_ACC;
            for (int kt = 0; kt < MT_KT; ++kt, ++it) {
                const uint32_t rs = it % MT_R, ruse = it / MT_R, ls = it % MT_L, luse = it / MT_L;
                const uint32_t hi_addr = raw_base + rs * MT_SLOT;
                const uint64_t w_hi = umma_desc_sw128(hi_addr + t * (128 * MT_BK * 4));
                const uint64_t x_hilo = umma_desc_sw128(hi_addr + MT_A_BYTES);        // 32 rows: x hi | x lo
                const uint32_t lo_tmem = tmem_base + MT_TM_LO + ls * MT_TM_LSLOT + t * MT_BK;
                tc_mbar_wait(bar_raw_full + rs, ruse & 1);
                tc_mbar_wait(bar_lo_full + ls, luse & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                // one elected lane issues; the block is executed convergently so that every operand can live in the
                // uniform datapath (a divergent `if (lane == 0)` costs an ELECT + R2UR.BROADCAST chain per instruction)
                mt_issue_ktile(d_tmem, w_hi, x_hilo, lo_tmem, kt ? 1u : 0u, tc_smem_u32(bar_raw_empty + rs),
                               tc_smem_u32(bar_lo_empty + ls));
                if (kt == MT_KT - 1) mt_commit_elected(tc_smem_u32(bar_tfull + as));
                __syncwarp();
            }
        }
    } else if (warp < 2 + MT_PROD_WARPS) {
        // ===================== producers: W2 lo -> tensor memory, activations hi | lo -> behind the W2 tile ======
        // Two groups of four warps take alternate k-tiles, so the latency chain of one k-tile (LDS -> tcgen05.st ->
        // wait::st -> proxy fence -> arrive) runs beside the next one's.
        const int grp = (warp - 2) >> 2;          // k-tiles with (it & 1) == grp
        const int pt = (threadIdx.x - 64) & 127;  // 0..127 inside the group
        const int q = warp & 3;                   // TMEM lane quarter of this warp
        const int r = q * 32 + lane;              // row of the 128-row tiles this thread owns
        const int e = pt & 15, g = pt >> 4;       // activations: episode e, 16-byte chunk g (4 k) of the k-tile
        uint32_t it = 0, js = 0;
        for (int job = job0; job < p.n_jobs; job += job_stride, ++js) {
            const uint32_t b = js & 1, buse = js >> 1;
            tc_mbar_wait(bar_stat_full + b, buse & 1);
            float x[IN];
            {
                const float* ob = obsbuf + ((size_t)b * MT_NB + e) * LS_OBS_PAD;
#pragma unroll
                for (int i = 0; i < IN; i += 2) {
                    const float2 v = *reinterpret_cast<const float2*>(ob + i);
                    x[i] = v.x;
                    x[i + 1] = v.y;
                }
            }
            const float2 st = statbuf[b * MT_NB + e];
            const float mean = st.x, rstd = st.y;
            for (int kt = 0; kt < MT_KT; ++kt, ++it) {
                if ((int)(it & 1) != grp) continue;
                const uint32_t rs = it % MT_R, ruse = it / MT_R, ls = it % MT_L, luse = it / MT_L;
                unsigned char* slot = raw_mem + (size_t)rs * MT_SLOT;
                if (luse > 0) tc_mbar_wait(bar_lo_empty + ls, (luse - 1) & 1);
                tc_mbar_wait(bar_raw_full + rs, ruse & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    // row r of tile t: 128 bytes, 16-byte chunk c stored at c ^ (r & 7)
                    const float4* row = reinterpret_cast<const float4*>(slot + (size_t)(t * 128 + r) * 128);
                    float wl[32];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 w = row[c ^ (r & 7)];
                        wl[4 * c] = f3_lo(w.x);
                        wl[4 * c + 1] = f3_lo(w.y);
                        wl[4 * c + 2] = f3_lo(w.z);
                        wl[4 * c + 3] = f3_lo(w.w);
                    }
                    f3_tmem_st32(tmem_base + ((uint32_t)(q * 32) << 16) + MT_TM_LO + ls * MT_TM_LSLOT + t * MT_BK, wl);
                }
                {
                    // activations k = kt * 32 + g * 4 + {0..3} of episode e (while the stores above are in flight)
                    const float* w1c = reinterpret_cast<const float*>(w1c_mem + (size_t)rs * MT_W1C_BYTES);
                    const int kl = g * 4;
                    const float4 bb = *reinterpret_cast<const float4*>(w1c + MT_BK * IN + kl);
                    const float4 gg = *reinterpret_cast<const float4*>(w1c + MT_BK * IN + MT_BK + kl);
                    const float4 ee = *reinterpret_cast<const float4*>(w1c + MT_BK * IN + 2 * MT_BK + kl);
                    const float bq[4] = {bb.x, bb.y, bb.z, bb.w}, gq[4] = {gg.x, gg.y, gg.z, gg.w},
                                eq[4] = {ee.x, ee.y, ee.z, ee.w};
                    float hi[4], lo[4];
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        const float2* wr = reinterpret_cast<const float2*>(w1c + (kl + qq) * IN);
                        float pre = 0.f;
#pragma unroll
                        for (int i2 = 0; i2 < IN / 2; ++i2) {
                            const float2 v = wr[i2];
                            pre = fmaf(v.x, x[2 * i2], pre);
                            pre = fmaf(v.y, x[2 * i2 + 1], pre);
                        }
                        pre += bq[qq];
                        const float h = fmaxf(fmaf((pre - mean) * rstd, gq[qq], eq[qq]), 0.f);
                        hi[qq] = __uint_as_float(__float_as_uint(h) & 0xffffe000u);
                        lo[qq] = h - hi[qq];
                    }
                    const int off = e * 128 + ((g ^ (e & 7)) << 4);
                    *reinterpret_cast<float4*>(slot + MT_A_BYTES + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(slot + MT_A_BYTES + MT_X_BYTES + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic writes -> tensor core reads
                asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                __syncwarp();
                if (lane == 0) tc_mbar_arrive(bar_lo_full + ls);
            }
        }
    } else {
        // ===================== epilogue (warps 10..13) =====================
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        const int et = (warp - 10) * 32 + lane;          // 0..127
        const int we = warp - 10;
        const int e1 = et & 15, g1 = et >> 4;            // LayerNorm-2 stage: episode e1, rows g1 + 8 j
        constexpr int RP = H2 / 8;
        const float* b2 = tail;
        const float* g2 = tail + H2;
        const float* be2 = tail + 2 * H2;
        const float* w3 = tail + 3 * H2;
        const float* b3 = tail + 3 * H2 + NACT * H2;
        // observations + LayerNorm-1 statistics (closed form, fp64) of local job `js` -> shared memory
        auto stage_job = [&](int job, uint32_t js) {
            if (job < p.n_jobs && et < MT_NB) {
                const uint32_t b = js & 1;
                const int m = job / p.n_chunks, ch = job % p.n_chunks;
                const int n_env = min(MT_NB, p.KE - ch * MT_NB);
                const int64_t ep = (int64_t)m * p.KE + (int64_t)ch * MT_NB + (et < n_env ? et : 0);
                const float4* src = reinterpret_cast<const float4*>(p.obs + ep * LS_OBS_PAD);
                const float4 v0 = __ldg(src), v1 = __ldg(src + 1), v2 = __ldg(src + 2);
                float4* dst = reinterpret_cast<float4*>(obsbuf + ((size_t)b * MT_NB + et) * LS_OBS_PAD);
                dst[0] = v0;
                dst[1] = v1;
                dst[2] = v2;
                const float xo[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
                const double* S = p.l1stats + (size_t)m * LS_L1S;
                double z[11];
#pragma unroll
                for (int a = 0; a < 11; ++a) z[a] = a < IN ? (double)xo[a] : (a == IN ? 1.0 : 0.0);
                double md = 0.0, vd = 0.0;
#pragma unroll
                for (int a = 0; a < 11; ++a) {
                    md = fma(__ldg(S + a), z[a], md);
                    double ra = 0.0;
#pragma unroll
                    for (int c = 0; c < 11; ++c) ra = fma(__ldg(S + 11 + a * 11 + c), z[c], ra);
                    vd = fma(ra, z[a], vd);
                }
                const float m1 = (float)md, var = (float)vd;
                if (!isfinite(m1) || !isfinite(var)) *flag = 1;
                statbuf[b * MT_NB + et] = make_float2(m1, 1.0f / sqrtf(var + LN_EPS));
            }
            if (job < p.n_jobs && we == 0) {
                __syncwarp();
                if (lane == 0) tc_mbar_arrive(bar_stat_full + (js & 1));
            }
        };
        stage_job(job0, 0);
        stage_job(job0 + job_stride, 1);
        uint32_t js = 0;
        for (int job = job0; job < p.n_jobs; job += job_stride, ++js) {
            const int m = job / p.n_chunks, ch = job % p.n_chunks;
            const int n_env = min(MT_NB, p.KE - ch * MT_NB);
            const int64_t ep0 = (int64_t)m * p.KE + (int64_t)ch * MT_NB;
            const uint32_t as = js & 1, ause = js >> 1;
            tc_mbar_wait(bar_tail_full, js & 1);
            tc_mbar_wait(bar_tfull + as, ause & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                uint32_t hh[16], hl[16], lh[16];
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + as * (2 * MT_ACC) + t * MT_ACC;
                f3_tmem_ld16(ta, hh);
                f3_tmem_ld16(ta + MT_NB, hl);
                f3_tmem_ld16(ta + 2 * MT_NB, lh);
                const int row = t * 128 + q * 32 + lane;
                const float bias = b2[row];
                float4* dst = reinterpret_cast<float4*>(hs + row * MT_NB);
                const int sw = (row >> 1) & 3;
                float h[16];
#pragma unroll
                for (int n = 0; n < 16; ++n)           // small terms first, then the bias (like fc2(h) + b)
                    h[n] = ((__uint_as_float(lh[n]) + __uint_as_float(hl[n])) + __uint_as_float(hh[n])) + bias;
#pragma unroll
                for (int c = 0; c < 4; ++c) dst[c ^ sw] = make_float4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(bar_tempty + as);
            asm volatile("bar.sync 1, 128;\n" ::: "memory");         // hs complete
            float pre2[RP];
            float lsum = 0.f;
#pragma unroll
            for (int j = 0; j < RP; ++j) {
                const int row = g1 + 8 * j;
                pre2[j] = hs[row * MT_NB + ((((e1 >> 2) ^ ((row >> 1) & 3))) << 2) + (e1 & 3)];
                lsum += pre2[j];
            }
            lsum = group_sum<MT_NB>(lsum);
            if (lane < MT_NB) redA[we * MT_NB + e1] = lsum;
            asm volatile("bar.sync 1, 128;\n" ::: "memory");
            float mean = 0.f;
#pragma unroll
            for (int w = 0; w < 4; ++w) mean += redA[w * MT_NB + e1];
            mean *= (1.0f / H2);
            float lsq = 0.f;
#pragma unroll
            for (int j = 0; j < RP; ++j) {
                const float d = pre2[j] - mean;
                lsq = fmaf(d, d, lsq);
            }
            lsq = group_sum<MT_NB>(lsq);
            if (lane < MT_NB) redB[we * MT_NB + e1] = lsq;
            asm volatile("bar.sync 1, 128;\n" ::: "memory");
            float var = 0.f;
#pragma unroll
            for (int w = 0; w < 4; ++w) var += redB[w * MT_NB + e1];
            var *= (1.0f / H2);
            if (!isfinite(mean) || !isfinite(var)) *flag = 1;
            const float rstd = 1.0f / sqrtf(var + LN_EPS);
            float pl[NACT] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < RP; ++j) {
                const int row = g1 + 8 * j;
                const float h = fmaxf(fmaf((pre2[j] - mean) * rstd, g2[row], be2[row]), 0.f);
#pragma unroll
                for (int a = 0; a < NACT; ++a) pl[a] = fmaf(w3[a * H2 + row], h, pl[a]);
            }
#pragma unroll
            for (int a = 0; a < NACT; ++a) {
                const float v = group_sum<MT_NB>(pl[a]);
                if (lane < MT_NB) redC[(we * NACT + a) * MT_NB + e1] = v;
            }
            asm volatile("bar.sync 1, 128;\n" ::: "memory");
            if (et < MT_NB) {
                float lg[NACT];
                bool fin = true;
#pragma unroll
                for (int a = 0; a < NACT; ++a) {
                    float v = redC[a * MT_NB + et];
#pragma unroll
                    for (int w = 1; w < 4; ++w) v += redC[(w * NACT + a) * MT_NB + et];
                    lg[a] = v + b3[a];
                    fin = fin && isfinite(lg[a]);
                }
                float gap;
                const int a = argmax_first5(lg, gap);
                if (et < n_env) {
                    if (!fin) *flag = 1;
                    p.act[ep0 + et] = a;
                    p.gap[ep0 + et] = gap;
                    if (p.logits) {
#pragma unroll
                        for (int u = 0; u < NACT; ++u) p.logits[(ep0 + et) * NACT + u] = lg[u];
                    }
                }
            }
            asm volatile("bar.sync 1, 128;\n" ::: "memory");         // everyone is done with tail / hs / red
            if (et == 0) tc_mbar_arrive(bar_tail_empty);
            stage_job(job + 2 * job_stride, js + 2);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0 && *flag && p.status) atomicOr(p.status, CEV_STATUS_NONFINITE);
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u) : "memory");
}

}  // namespace cev
