// Device helpers shared by the K1 rollout kernels (rollout_cluster.cu, rollout_lockstep.cu):
// packed FFMA2, cp.async / mbarrier / bulk-copy plumbing and the layer-1 + LayerNorm stage of
// FCNetwork.forward (MPE/fcnetwork.py:37-52).
#pragma once

#include "common.cuh"

#ifndef CEV_USE_FFMA2
#define CEV_USE_FFMA2 1
#endif

namespace cev {

constexpr int CT = 256;                     // threads per CTA
constexpr int NW = CT / 32;                 // warps per CTA

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
// Thread (ep = t % (BT/2), g = t / (BT/2)) owns env pair (2ep, 2ep+1) and the
// row PAIRS g + G*i: two adjacent rows share 128-bit weight loads, two envs
// share every weight, and the output quad (2 rows x 2 envs) is one 128-bit store
// into the k-pair interleaved layout  h1p[(k>>1)*(2*BT) + 2*e + (k&1)].
// ---------------------------------------------------------------------------
template <int BT, int IN>
__device__ __forceinline__ void layer1(const float* __restrict__ w1a, const float* __restrict__ obs_seat,
                                       float* __restrict__ h1p, float* __restrict__ red1,
                                       int* flag) {
    constexpr int NEP = BT / 2;             // env pairs
    constexpr int G = CT / NEP;             // row-pair groups
    constexpr int NP = (H1 / 2) / G;        // row pairs per thread
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
    // mean over the 512 rows of each env: lanes sharing ep, then the warps
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
        for (int w = 0; w < NW; ++w) tot += red1[w * BT + 2 * ep + j];
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
    float* red1b = red1 + NW * BT;
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
        for (int w = 0; w < NW; ++w) tot += red1b[w * BT + 2 * ep + j];
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

// ---------------------------------------------------------------------------
// mbarrier + bulk-copy (TMA, non-tensor) plumbing for the streamed operands.
// A stage is one contiguous, pre-swizzled 8 KB block of the packed opponent
// matrix, so ONE elected lane moves it with one cp.async.bulk; consumers wait on
// the stage's "full" mbarrier (completed by the copy's byte count) and release
// it through the "empty" mbarrier -- no CTA-wide barrier inside the fc2 loop.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// wait for completion #n (n = 0, 1, ...) of the barrier; traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t n) {
    if (mbar_try_wait(bar, n & 1)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, n & 1)) {
        if (clock64() - t0 > 4000000000LL) {
#ifdef CEV_PROFILE
            return;   // development build: carry on (results are then invalid)
#else
            __trap();
#endif
        }
    }
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace cev
