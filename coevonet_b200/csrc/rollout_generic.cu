// K1 (generic form): one CTA per episode, arbitrary (adversary, agent_0,
// agent_1) row triple per episode.  Weights are read straight from L2/HBM each
// step, so this kernel is bandwidth-heavy by construction; it exists as the
// obviously-correct device path (play_game drop-in, odd shapes, cross-check of
// the cluster kernel).  The fast structured path is rollout_cluster.cu.
//
// Replaces: play_game/play_MPE (utils/game_logic_functions.py:123-228),
// FCNetwork.forward/determine_action (MPE/fcnetwork.py:37-90) and the
// simple_adversary_v3 world step.
#include "common.cuh"

namespace cev {

constexpr int GEN_THREADS = 256;
constexpr int GEN_WARPS = GEN_THREADS / 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum with a fixed reduction order (deterministic); all threads get it.
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();                     // protect red[] from the previous use
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < GEN_WARPS; ++w) t += red[w];
    return t;
}

// One FCNetwork forward for one observation.  obs/h1/h2 live in shared memory.
// Returns logits in lg[] (valid in every thread).
__device__ void fc_forward_one(const float* __restrict__ W, int in_dim, const float* obs,
                               float* h1, float* h2, float* red, float* lgs, int* nonfinite) {
    const FcOffsets o = fc_offsets(in_dim);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    // ---- layer 1: rows t, t+256 ------------------------------------------------
    float p[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int row = t + r * GEN_THREADS;
        const float* w = W + o.fc1w + row * in_dim;
        float acc = 0.f;
        for (int k = 0; k < in_dim; ++k) acc = fmaf(__ldg(w + k), obs[k], acc);
        p[r] = acc + __ldg(W + o.fc1b + row);
    }
    float mean = block_sum(p[0] + p[1], red) * (1.0f / H1);
    float d0 = p[0] - mean, d1 = p[1] - mean;
    float var = block_sum(d0 * d0 + d1 * d1, red) * (1.0f / H1);
    if (!isfinite(mean) || !isfinite(var)) *nonfinite = 1;
    float rstd = 1.0f / sqrtf(var + LN_EPS);
    {
        const int r0 = t, r1 = t + GEN_THREADS;
        h1[r0] = fmaxf(d0 * rstd * __ldg(W + o.ln1g + r0) + __ldg(W + o.ln1b + r0), 0.f);
        h1[r1] = fmaxf(d1 * rstd * __ldg(W + o.ln1g + r1) + __ldg(W + o.ln1b + r1), 0.f);
    }
    __syncthreads();
    // ---- layer 2: warp w owns rows w*32 .. w*32+31, lanes stride k -------------
    float mine = 0.f;   // lane r of warp w ends up with row w*32 + r
    for (int r = 0; r < 32; ++r) {
        const int row = warp * 32 + r;
        const float4* w4 = reinterpret_cast<const float4*>(W + o.fc2w + row * H1);
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 w = __ldg(w4 + lane + 32 * i);
            const float4 x = *reinterpret_cast<const float4*>(h1 + (lane + 32 * i) * 4);
            acc = fmaf(w.x, x.x, acc);
            acc = fmaf(w.y, x.y, acc);
            acc = fmaf(w.z, x.z, acc);
            acc = fmaf(w.w, x.w, acc);
        }
        acc = warp_sum(acc);
        if (lane == r) mine = acc + __ldg(W + o.fc2b + row);
    }
    mean = block_sum(mine, red) * (1.0f / H2);
    d0 = mine - mean;
    var = block_sum(d0 * d0, red) * (1.0f / H2);
    if (!isfinite(mean) || !isfinite(var)) *nonfinite = 1;
    rstd = 1.0f / sqrtf(var + LN_EPS);
    h2[t] = fmaxf(d0 * rstd * __ldg(W + o.ln2g + t) + __ldg(W + o.ln2b + t), 0.f);
    __syncthreads();
    // ---- output layer: warp a computes logit a ---------------------------------
    if (warp < NACT) {
        const float* w = W + o.outw + warp * H2;
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < H2 / 32; ++i) acc = fmaf(__ldg(w + lane + 32 * i), h2[lane + 32 * i], acc);
        acc = warp_sum(acc);
        if (lane == 0) lgs[warp] = acc + __ldg(W + o.outb + warp);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(GEN_THREADS) rollout_generic_kernel(GenericParams p) {
    __shared__ __align__(16) float h1[H1];
    __shared__ __align__(16) float h2[H2];
    __shared__ float red[GEN_WARPS];
    __shared__ float obs[3][12];
    __shared__ float lgs[NACT];
    __shared__ int act_s[3];
    __shared__ int nonfinite;

    for (int64_t e = blockIdx.x; e < p.N; e += gridDim.x) {
        int row[3];
        const double* rec;
        if (p.idx) {
            row[0] = p.idx[e * 3 + 0];
            row[1] = p.idx[e * 3 + 1];
            row[2] = p.idx[e * 3 + 2];
            rec = p.init + e * CEV_INIT_STATE_DIM;
        } else {
            const int64_t ke = (int64_t)p.K * p.E;
            const int m = (int)(e / ke);
            const int k = (int)((e % ke) / p.E);
            for (int s = 0; s < 3; ++s) row[s] = (s == p.member_seat) ? m : k;
            rec = p.init + (p.init_shared ? (e % ke) : e) * CEV_INIT_STATE_DIM;
        }
        EnvState st;
        env_load(st, rec);
        double sum_good = 0.0, last_good = 0.0, sum_adv = 0.0;
        float min_gap = CUDART_INF_F;
        if (threadIdx.x == 0) nonfinite = 0;
        __syncthreads();
        for (int c = 0; c < p.n_cycles; ++c) {
            if (threadIdx.x == 0) {
#pragma unroll
                for (int s = 0; s < 3; ++s) env_observe(st, s, obs[s]);
            }
            __syncthreads();
            for (int s = 0; s < 3; ++s) {
                const float* W = p.w[s] + (int64_t)row[s] * p.pitch[s];
                fc_forward_one(W, seat_in_dim(s), obs[s], h1, h2, red, lgs, &nonfinite);
                if (threadIdx.x == 0) {
                    float lg[NACT];
                    bool fin = true;
#pragma unroll
                    for (int a = 0; a < NACT; ++a) { lg[a] = lgs[a]; fin = fin && isfinite(lg[a]); }
                    if (!fin) nonfinite = 1;
                    float gap;
                    act_s[s] = argmax_first5(lg, gap);
                    min_gap = fminf(min_gap, gap);
                }
                __syncthreads();
            }
            const int act[3] = {act_s[0], act_s[1], act_s[2]};
            double rg, ra;
            env_step(st, act, p.pos_first != 0, rg, ra);   // every thread, identical
            sum_good = __dadd_rn(sum_good, rg);
            sum_adv = __dadd_rn(sum_adv, ra);
            last_good = rg;
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            double* o = p.out + e * CEV_ROLLOUT_OUT_DIM;
            o[0] = sum_good;
            o[1] = last_good;
            o[2] = sum_adv;
            o[3] = (double)min_gap;
            if (nonfinite && p.status) atomicOr(p.status, CEV_STATUS_NONFINITE);
        }
        __syncthreads();
    }
}

int launch_rollout_generic(cev_handle* h, const GenericParams& p, cudaStream_t stream) {
    if (p.N <= 0) return CEV_OK;
    const int64_t max_grid = (int64_t)h->n_sm * 8;
    const int grid = (int)(p.N < max_grid ? p.N : max_grid);
    rollout_generic_kernel<<<grid, GEN_THREADS, 0, stream>>>(p);
    return check_cuda(cudaGetLastError(), "rollout_generic_kernel launch");
}

// Batched FCNetwork.forward on arbitrary observations (function-level op for
// logits parity under teacher forcing and for the FCNetwork.forward drop-in):
// one CTA per (row idx[n], obs[n]) pair.
__global__ void __launch_bounds__(GEN_THREADS) fc_forward_kernel(const float* __restrict__ W, int64_t pitch,
                                                                int in_dim, const int32_t* __restrict__ idx,
                                                                const float* __restrict__ obs_in, int64_t N,
                                                                float* __restrict__ logits,
                                                                int32_t* __restrict__ actions,
                                                                int32_t* status) {
    __shared__ __align__(16) float h1[H1];
    __shared__ __align__(16) float h2[H2];
    __shared__ float red[GEN_WARPS];
    __shared__ float obs[12];
    __shared__ float lgs[NACT];
    __shared__ int nonfinite;
    for (int64_t n = blockIdx.x; n < N; n += gridDim.x) {
        if (threadIdx.x == 0) nonfinite = 0;
        if (threadIdx.x < in_dim) {
            const float v = obs_in[n * in_dim + threadIdx.x];
            obs[threadIdx.x] = v;
            if (!isfinite(v)) nonfinite = 1;
        }
        __syncthreads();
        const float* Wr = W + (int64_t)(idx ? idx[n] : 0) * pitch;
        fc_forward_one(Wr, in_dim, obs, h1, h2, red, lgs, &nonfinite);
        if (threadIdx.x == 0) {
            float lg[NACT];
            bool fin = true;
#pragma unroll
            for (int a = 0; a < NACT; ++a) {
                lg[a] = lgs[a];
                logits[n * NACT + a] = lg[a];
                fin = fin && isfinite(lg[a]);
            }
            float gap;
            const int act = argmax_first5(lg, gap);
            if (actions) actions[n] = act;
            if ((!fin || nonfinite) && status) atomicOr(status, CEV_STATUS_NONFINITE);
        }
        __syncthreads();
    }
}

int launch_fc_forward(cev_handle* h, const float* W, int64_t pitch, int in_dim, const int32_t* idx,
                      const float* obs, int64_t N, float* logits, int32_t* actions, int32_t* status,
                      cudaStream_t stream) {
    if (N <= 0) return CEV_OK;
    const int64_t max_grid = (int64_t)h->n_sm * 8;
    const int grid = (int)(N < max_grid ? N : max_grid);
    fc_forward_kernel<<<grid, GEN_THREADS, 0, stream>>>(W, pitch, in_dim, idx, obs, N, logits, actions, status);
    return check_cuda(cudaGetLastError(), "fc_forward_kernel launch");
}

}  // namespace cev
