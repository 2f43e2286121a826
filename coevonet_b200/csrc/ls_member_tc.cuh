// K1 (lockstep form), member forward on the tensor cores.  Included by rollout_lockstep.cu.
//
// The member's own forward (FCNetwork.forward, MPE/fcnetwork.py:37-70, with the member's OWN weights:
// evolutionary_strategy.py:63-116 mutate_weights, genetic_algorithm.py:125-217) is, per member and world step,
// H[256, 16 episodes] = W2_m[256, 512] . h1[16, 512]^T with 512 KB of W2_m that nobody else reads: the stage streams
// the member rows from HBM once per world step.  This form keeps the FMA pipe out of it (3xTF32 on tcgen05) and runs
// as a PERSISTENT kernel on a share of the SMs, so that the opponent kernel (tensor-pipe bound, ls_opp_kernel) runs
// beside it on the other SMs instead of after it.
//
// One CTA per SM, job = (member, block of 16 episodes).  Per k-tile of 32:
//   warp 0      TMA: the raw W2 tile [256 rows x 32 k] (3-D tensor map (k, row, member), 128B swizzle, 256B L2
//               promotion) into a 5-slot ring, with the k-tile's 32 rows of the member's layer 1 (fc1.W | fc1.b |
//               ln1.g | ln1.b, 1.7 KB) beside it; per job the tail block (fc2.b | ln2.g | ln2.b | out.W | out.b)
//   warps 2-9   two groups of four warps take alternate k-tiles: read the W2 tile ONCE from shared memory and write
//               both MMA A operands to TENSOR MEMORY (tcgen05.st, lane = W2 row): hi = the raw fp32 word (the tensor
//               core reads it truncated to TF32) and lo = w - trunc_tf32(w).  The raw slot is free again as soon as
//               it has been read (not when its MMAs retire), and the MMAs fetch nothing but the small activation
//               tile from shared memory: with W2 as a shared-memory operand the stage moved 115 KB per k-tile
//               through the 128 B/clk shared-memory port (TMA fill 34 + LDS 34 + operand fetch 44) and was bound by
//               it at ~1050 cycles per k-tile.  The same warps compute the activations of the k-tile,
//               relu(LN1(W1 x + b1)) for the 16 episodes, split into hi | lo and stored K-major as ONE 32-row N
//               operand [x_hi | x_lo].  LayerNorm-1 uses the closed-form row statistics (mean = wbar . z,
//               var = z^T C z, fp64; ls_member_l1stats_kernel, once per rollout)
//   warps 1,14  16 x tcgen05.mma.cta_group::1.kind::tf32 per k-tile, A from tensor memory (2 row tiles, one per
//               issuing warp, x 4 k-steps x {w_hi, w_lo} . [x_hi | x_lo], N = 32) into ping-pong accumulators; a
//               tcgen05.mma of this size occupies the tensor pipe ~45 cycles whatever N <= 64 is
//               (scripts/probe/mma_probe.cu), so 720 cycles per k-tile is the floor of this form
//   warps 10-13 epilogue: tcgen05.ld, (x_hi columns + x_lo columns) + fc2.b into shared memory, LayerNorm-2 + ReLU +
//               output layer + first-max argmax (MPE/fcnetwork.py:53-90); then the observations and LayerNorm-1
//               statistics of the job after the next one, which the producers pick up from shared memory
#pragma once

namespace cev {

constexpr int MT_THREADS = 480;                               // warp 0 TMA, 1 + 14 MMA, 2-9 producers, 10-13 epilogue
constexpr int MT_PROD_WARPS = 8;
constexpr int MT_BK = 32, MT_KT = H1 / MT_BK;                 // 16 k-tiles of 128 bytes
constexpr int MT_NB = LS_BT;                                  // 16 episodes
constexpr int MT_R = 5, MT_L = 3;                             // raw slots (shared memory), operand slots (tensor memory)
constexpr uint32_t MT_A_BYTES = H2 * MT_BK * 4;               // 32 KB
constexpr uint32_t MT_X_BYTES = MT_NB * MT_BK * 4;            // 2 KB
constexpr uint32_t MT_XSLOT = 2 * MT_X_BYTES;                 // x hi | x lo = 4 KB, one per operand slot
constexpr uint32_t MT_W1C_FLOATS = MT_BK * IN_GOOD + 3 * MT_BK;   // layer-1 chunk of a k-tile: fc1.W rows | fc1.b | ln1.g | ln1.b
constexpr uint32_t MT_W1C_BYTES = MT_W1C_FLOATS * 4;          // 1,664
constexpr uint32_t MT_TAIL_BYTES = LS_TAIL_FLOATS * 4;        // 8,224
constexpr size_t MT_OFF_X = (size_t)MT_R * MT_A_BYTES;        // [L] activation tiles
constexpr size_t MT_OFF_W1C = MT_OFF_X + (size_t)MT_L * MT_XSLOT;         // [R] layer-1 chunks, one per raw slot
constexpr size_t MT_OFF_HS = MT_OFF_W1C + (size_t)MT_R * MT_W1C_BYTES;    // [256 rows][16 episodes] fp32
constexpr size_t MT_OFF_TAIL = MT_OFF_HS + (size_t)H2 * MT_NB * 4;
constexpr size_t MT_OFF_OBS = MT_OFF_TAIL + MT_TAIL_BYTES;    // [2][16][12] fp32
constexpr size_t MT_OFF_STAT = MT_OFF_OBS + 2 * MT_NB * LS_OBS_PAD * 4;   // [2][16] (mean, rstd)
constexpr size_t MT_OFF_RED = MT_OFF_STAT + 2 * MT_NB * 8;    // redA [4][16] | redB [4][16] | redC [4][5][16]
constexpr size_t MT_OFF_BAR = MT_OFF_RED + (size_t)(2 + NACT) * 4 * MT_NB * 4;
constexpr size_t MT_SMEM = MT_OFF_BAR + 256 + 1024 /*alignment*/;
constexpr int MT_ACC = 2 * MT_NB;                             // accumulator columns per row tile: . x_hi | . x_lo
constexpr int MT_TM_OP = 2 * 2 * MT_ACC;                      // first TMEM column of the operand slots (after 2 x 2 accumulators)
constexpr int MT_TM_SLOT = 2 * 2 * MT_BK;                     // columns per operand slot: 2 row tiles x (hi 32 | lo 32)
static_assert(MT_TM_OP + MT_L * MT_TM_SLOT <= 512, "tensor memory budget");
static_assert(MT_OFF_X % 1024 == 0 && MT_XSLOT % 1024 == 0, "operand tiles must keep the 1024-byte alignment of the swizzle atoms");
static_assert(MT_OFF_W1C % 16 == 0 && MT_W1C_BYTES % 16 == 0 && MT_OFF_TAIL % 16 == 0 && MT_OFF_BAR % 8 == 0,
              "bulk-copy alignment");
static_assert(MT_SMEM <= 232448, "member stage exceeds the 227 KB shared-memory limit");
constexpr uint32_t MT_IDESC32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * MT_NB) >> 3) << 17) |
                                ((uint32_t)(128 >> 4) << 24);

// The eight MMAs of one row tile and k-tile (4 k-steps x {w_hi, w_lo} (tensor memory) . [x_hi | x_lo] (N = 32)) plus the
// commit that releases the operand slot, issued by one elected lane of a CONVERGED warp: inside a divergent
// `if (lane == 0)` every tcgen05.mma costs an ELECT + R2UR.BROADCAST chain (~20 instructions).
__device__ __forceinline__ void mt_issue_ktile(uint32_t d_tmem, uint32_t a_tmem, uint64_t x_hilo, uint32_t accumulate,
                                               uint32_t bar_op) {
    asm volatile(
        "{\n"
        ".reg .pred pe, pa, pt;\n"
        ".reg .b64 b1, b2, b3;\n"
        ".reg .b32 h1, h2, h3, l0, l1, l2, l3;\n"
        "elect.sync _|pe, 0xffffffff;\n"
        "setp.ne.b32 pa, %3, 0;\n"
        "setp.eq.b32 pt, 0, 0;\n"
        "add.s64 b1, %2, 2;\n add.s64 b2, %2, 4;\n add.s64 b3, %2, 6;\n"
        "add.u32 h1, %1, 8;\n add.u32 h2, %1, 16;\n add.u32 h3, %1, 24;\n"
        "add.u32 l0, %1, 32;\n add.u32 l1, %1, 40;\n add.u32 l2, %1, 48;\n add.u32 l3, %1, 56;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [l0], %2, %5, pa;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %5, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [l1], b1, %5, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [h1], b1, %5, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [l2], b2, %5, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [h2], b2, %5, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [l3], b3, %5, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [h3], b3, %5, pt;\n"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%4];\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(x_hilo), "r"(accumulate), "r"(bar_op), "r"(MT_IDESC32)
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
    unsigned char* raw_mem = base;                                    // [R] raw W2 tiles
    unsigned char* x_mem = base + MT_OFF_X;                           // [L] (x hi | x lo)
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
            tc_mbar_init(bar_raw_empty + i, MT_PROD_WARPS / 2);    // read by the producer warps of the k-tile's group
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
        // 2 (ping-pong) x 2 (row tiles) x 32 accumulator columns = 128, then 3 operand slots x 128 columns -> 512
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
                    tma_load_3d(raw_mem + (size_t)rs * MT_A_BYTES, &map_w, bar_raw_full + rs, kt * MT_BK, 0, m);
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
        const int t = warp == 1 ? 0 : 1;
        const uint32_t x_base = tc_smem_u32(x_mem);
        uint32_t it = 0, js = 0;
        for (int job = job0; job < p.n_jobs; job += job_stride, ++js) {
            const uint32_t as = js & 1, ause = js >> 1;
            if (ause > 0) tc_mbar_wait(bar_tempty + as, (ause - 1) & 1);      // epilogue drained it
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t d_tmem = tmem_base + as * (2 * MT_ACC) + t * MT_ACC;
            for (int kt = 0; kt < MT_KT; ++kt, ++it) {
                const uint32_t ls = it % MT_L, luse = it / MT_L;
                const uint64_t x_hilo = umma_desc_sw128(x_base + ls * MT_XSLOT);        // 32 rows: x hi | x lo
                const uint32_t a_tmem = tmem_base + MT_TM_OP + ls * MT_TM_SLOT + t * (2 * MT_BK);
                tc_mbar_wait(bar_lo_full + ls, luse & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                mt_issue_ktile(d_tmem, a_tmem, x_hilo, kt ? 1u : 0u, tc_smem_u32(bar_lo_empty + ls));
                if (kt == MT_KT - 1) tc_commit_elected(tc_smem_u32(bar_tfull + as));
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
                const unsigned char* slot = raw_mem + (size_t)rs * MT_A_BYTES;
                unsigned char* xs = x_mem + (size_t)ls * MT_XSLOT;
                if (luse > 0) tc_mbar_wait(bar_lo_empty + ls, (luse - 1) & 1);
                tc_mbar_wait(bar_raw_full + rs, ruse & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    // row r of tile t: 128 bytes, 16-byte chunk c stored at c ^ (r & 7)
                    const float4* row = reinterpret_cast<const float4*>(slot + (size_t)(t * 128 + r) * 128);
                    float w[32];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 v = row[c ^ (r & 7)];
                        w[4 * c] = v.x;
                        w[4 * c + 1] = v.y;
                        w[4 * c + 2] = v.z;
                        w[4 * c + 3] = v.w;
                    }
                    const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + MT_TM_OP + ls * MT_TM_SLOT + t * (2 * MT_BK);
                    f3_tmem_st32(ta, w);                       // hi: the raw word, read truncated to TF32 by the tensor core
#pragma unroll
                    for (int i = 0; i < 32; ++i) w[i] = f3_lo(w[i]);
                    f3_tmem_st32(ta + MT_BK, w);               // lo
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
                    *reinterpret_cast<float4*>(xs + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(xs + MT_X_BYTES + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
                __syncwarp();
                if (lane == 0) tc_mbar_arrive(bar_raw_empty + rs);           // the raw slot has been read: TMA may refill it
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
                uint32_t xh[16], xl[16];
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + as * (2 * MT_ACC) + t * MT_ACC;
                f3_tmem_ld16(ta, xh);
                f3_tmem_ld16(ta + MT_NB, xl);
                const int row = t * 128 + q * 32 + lane;
                const float bias = b2[row];
                float4* dst = reinterpret_cast<float4*>(hs + row * MT_NB);
                const int sw = (row >> 1) & 3;
                float h[16];
#pragma unroll
                for (int n = 0; n < 16; ++n)           // the x_lo columns (small) first, then the bias (like fc2(h) + b)
                    h[n] = (__uint_as_float(xl[n]) + __uint_as_float(xh[n])) + bias;
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
