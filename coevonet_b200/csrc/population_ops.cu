// K3-K7: population-level GA / ES kernels over flat parameter rows.
// All are streaming kernels (HBM-bound or Philox-ALU-bound): 128-bit accesses,
// one float4 of parameters per thread, grids sized from the data.
#include "common.cuh"

namespace cev {

constexpr int PT = 256;

// ---------------------------------------------------------------------------
// K3  GA re-population   (genetic_algorithm.py:32-48,255-290; agent.py:25-29)
// child c>=1 = elites[(c-1)%E] + sigma*z ; child 0 = elites[0] unmutated.
// Every parameter is mutated (LayerNorm gamma/beta too, Appendix C #8).
// mul and add are rounded separately, like `param.data += noise`.
// ---------------------------------------------------------------------------
// sigma of a launch: the by-value argument, or -- when the generation state lives on the device
// (cev_generation_end_f64 adapts it there) -- the fp64 scalar the caller points at,
// rounded to fp32 exactly like the host's float(sigma).
__device__ __forceinline__ float pick_sigma(float sigma, const double* __restrict__ sigma_dev) {
    return sigma_dev ? (float)__ldg(sigma_dev) : sigma;
}

// Optional uniform crossover (BASELINE north star item 3; the reference's README.md:47 promises crossover, its
// code has none -- SURVEY.md Appendix C #7 -- so this is an extension, off by default): with probability
// `xrate` child c takes every parameter from parent A = elites[(c-1) % E] or from a second, different elite B
// (one Philox bit per parameter), then mutates as usual.  Philox kind CEV_KIND_XOVER: block (0xFFFFFFFF, c, gen)
// decides (word 0 < xrate * 2^32) and picks the mate (word 1), block (j4, c, gen) holds the four mask bits.
__global__ void __launch_bounds__(PT) ga_repopulate_kernel(
    const float* __restrict__ elites, int E, int D, int64_t pitch, float sigma_arg,
    const double* __restrict__ sigma_dev,
    const PhiloxKeys keys, uint32_t tag, uint32_t gen, int64_t row0,
    float xrate, uint32_t xtag,
    float* __restrict__ out, float* __restrict__ noise_out) {
    const float sigma = pick_sigma(sigma_arg, sigma_dev);
    const int64_t r = blockIdx.x;
    const int64_t c = row0 + r;                       // global member id
    const int j4 = blockIdx.y * PT + threadIdx.x;
    __shared__ int64_t mate_s;
    int64_t mate = -1;
    if (xrate > 0.f && E >= 2) {                       // uniform per CTA: one child per blockIdx.x
        if (threadIdx.x == 0) {
            int64_t m = -1;
            if (c != 0) {
                const U4 y = philox4x32_10(U4{0xFFFFFFFFu, (uint32_t)c, gen, xtag}, keys);
                if ((double)y.x < (double)xrate * 4294967296.0)
                    m = ((c - 1) % E + 1 + (int64_t)(y.y % (uint32_t)(E - 1))) % E;
            }
            mate_s = m;
        }
        __syncthreads();
        mate = mate_s;
    }
    if ((int64_t)j4 * 4 >= pitch) return;
    const int64_t parent = (c == 0) ? 0 : (c - 1) % E;
    const float4 pv = *reinterpret_cast<const float4*>(elites + parent * pitch + (int64_t)j4 * 4);
    float p[4] = {pv.x, pv.y, pv.z, pv.w};
    if (mate >= 0) {
        const float4 qv = *reinterpret_cast<const float4*>(elites + mate * pitch + (int64_t)j4 * 4);
        const U4 x = philox4x32_10(U4{(uint32_t)j4, (uint32_t)c, gen, xtag}, keys);
        if (!(x.x & 1u)) p[0] = qv.x;
        if (!(x.y & 1u)) p[1] = qv.y;
        if (!(x.z & 1u)) p[2] = qv.z;
        if (!(x.w & 1u)) p[3] = qv.w;
    }
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    if (c != 0) {
        normal4(keys, tag, gen, (uint32_t)c, (uint32_t)j4, z);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int j = j4 * 4 + i;
            if (j < D) p[i] = __fadd_rn(p[i], __fmul_rn(sigma, z[i]));
            else { p[i] = 0.f; z[i] = 0.f; }
        }
    }
    *reinterpret_cast<float4*>(out + r * pitch + (int64_t)j4 * 4) = make_float4(p[0], p[1], p[2], p[3]);
    if (noise_out)
        *reinterpret_cast<float4*>(noise_out + r * pitch + (int64_t)j4 * 4) = make_float4(z[0], z[1], z[2], z[3]);
}

// ---------------------------------------------------------------------------
// K5  ES perturbation   (agent.py:31-70): Linear parameters only.
// ---------------------------------------------------------------------------
// CTAs of K5T threads at <= 48 registers: 6 K of the register file, so a perturb CTA of one role can be resident
// next to the rollout kernels' CTAs of another role (two 128-thread member CTAs at 230 registers leave 6.6 K,
// the 448-thread opponent CTA 9 K) and fill their idle issue slots instead of waiting for a free SM.
// Every segment boundary of the row is a multiple of 4 parameters except the end of the row, so a float4 is
// perturbable or not as a whole (the LayerNorm float4s skip the Philox / Box-Muller work altogether).
constexpr int K5T = 128;
__global__ void __launch_bounds__(K5T, 10) es_perturb_kernel(
    const float* __restrict__ theta, int in_dim, int64_t pitch, float sigma_arg,
    const double* __restrict__ sigma_dev,
    const PhiloxKeys keys, uint32_t tag, uint32_t gen, int64_t row0,
    float* __restrict__ out, float* __restrict__ noise_out) {
    const float sigma = pick_sigma(sigma_arg, sigma_dev);
    const FcOffsets o = fc_offsets(in_dim);
    const int64_t r = blockIdx.x;
    const int64_t c = row0 + r;
    const int j4 = blockIdx.y * K5T + threadIdx.x;
    const int j = j4 * 4;
    if ((int64_t)j >= pitch) return;
    const float4 pv = __ldg(reinterpret_cast<const float4*>(theta + j));
    float p[4] = {pv.x, pv.y, pv.z, pv.w};
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    if (j < o.total && fc_is_perturbable(o, j)) {
        normal4(keys, tag, gen, (uint32_t)c, (uint32_t)j4, z);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (j + i < o.total) p[i] = __fadd_rn(p[i], __fmul_rn(sigma, z[i]));
            else { z[i] = 0.f; p[i] = 0.f; }
        }
    } else if (j >= o.total) {
        p[0] = p[1] = p[2] = p[3] = 0.f;
    }
    __stcs(reinterpret_cast<float4*>(out + r * pitch + j), make_float4(p[0], p[1], p[2], p[3]));
    if (noise_out)
        *reinterpret_cast<float4*>(noise_out + r * pitch + j) = make_float4(z[0], z[1], z[2], z[3]);
}

// K5 for rows whose perturbable parameters are a PREFIX of the row: DeepQN rows hold the conv / Linear
// tensors first and the six BatchNorm vectors last (Atari/deepqn.py:158-172 get_perturbable_layers skips
// them), so parameter j is perturbed iff j < d_pert.
__global__ void __launch_bounds__(PT) es_perturb_prefix_kernel(
    const float* __restrict__ theta, int64_t d_pert, int64_t d_total, int64_t pitch, float sigma_arg,
    const double* __restrict__ sigma_dev,
    const PhiloxKeys keys, uint32_t tag, uint32_t gen, int64_t row0, float* __restrict__ out) {
    const float sigma = pick_sigma(sigma_arg, sigma_dev);
    const int64_t r = blockIdx.x;
    const int64_t j4 = (int64_t)blockIdx.y * PT + threadIdx.x;
    if (j4 * 4 >= pitch) return;
    const float4 pv = *reinterpret_cast<const float4*>(theta + j4 * 4);
    float p[4] = {pv.x, pv.y, pv.z, pv.w};
    if (j4 * 4 < d_pert) {
        float z[4];
        normal4(keys, tag, gen, (uint32_t)(row0 + r), (uint32_t)j4, z);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (j4 * 4 + i < d_pert) p[i] = __fadd_rn(p[i], __fmul_rn(sigma, z[i]));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (j4 * 4 + i >= d_total) p[i] = 0.f;
    __stcs(reinterpret_cast<float4*>(out + r * pitch + j4 * 4), make_float4(p[0], p[1], p[2], p[3]));
}

// ---------------------------------------------------------------------------
// K6  ES update   (evolutionary_strategy.py:120-148)
// delta = lr/(n*sigma) * sum_i (sigma*z_i) * F_i, noise regenerated from the
// Philox key.  Members are split over blockIdx.y; partial sums are reduced in a
// fixed order by the second kernel (deterministic, no atomics).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(PT) es_update_partial_kernel(
    const double* __restrict__ fitness, int in_dim, int64_t pitch, float sigma_arg,
    const double* __restrict__ sigma_dev,
    const PhiloxKeys keys, uint32_t tag, uint32_t gen, int64_t row0, int64_t n_rows,
    int n_split, float* __restrict__ partial) {
    const float sigma = pick_sigma(sigma_arg, sigma_dev);
    const FcOffsets o = fc_offsets(in_dim);
    const int j4 = blockIdx.x * PT + threadIdx.x;
    if ((int64_t)j4 * 4 >= pitch) return;
    const int sp = blockIdx.y;
    const int64_t per = (n_rows + n_split - 1) / n_split;
    const int64_t r_lo = sp * per, r_hi = min(n_rows, r_lo + per);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    bool pert[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int j = j4 * 4 + i;
        pert[i] = j < o.total && fc_is_perturbable(o, j);
    }
    if (pert[0] || pert[1] || pert[2] || pert[3]) {
        for (int64_t r = r_lo; r < r_hi; ++r) {
            const float f = (float)fitness[r];
            float z[4];
            normal4(keys, tag, gen, (uint32_t)(row0 + r), (uint32_t)j4, z);
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = fmaf(__fmul_rn(sigma, z[i]), f, acc[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (!pert[i]) acc[i] = 0.f;
    *reinterpret_cast<float4*>(partial + (int64_t)sp * pitch + (int64_t)j4 * 4) =
        make_float4(acc[0], acc[1], acc[2], acc[3]);
}

// The same partial sums from the MATERIALISED members: sigma z_i = member_i - theta up to one rounding
// of the perturbation's add (1e-7 relative), read back at HBM speed instead of regenerating P normals per
// parameter on the ALU.  LayerNorm entries are copies of theta, so they contribute exact zeros.
__global__ void __launch_bounds__(PT) es_update_members_partial_kernel(
    const double* __restrict__ fitness, const float* __restrict__ members, const float* __restrict__ theta,
    int64_t pitch, int64_t n_rows, int n_split, float* __restrict__ partial) {
    const int j4 = blockIdx.x * PT + threadIdx.x;
    if ((int64_t)j4 * 4 >= pitch) return;
    const int sp = blockIdx.y;
    const int64_t per = (n_rows + n_split - 1) / n_split;
    const int64_t r_lo = sp * per, r_hi = min(n_rows, r_lo + per);
    const float4 t4 = *reinterpret_cast<const float4*>(theta + (int64_t)j4 * 4);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float4* col = reinterpret_cast<const float4*>(members) + j4;
    const int64_t stride4 = pitch / 4;
#pragma unroll 8
    for (int64_t r = r_lo; r < r_hi; ++r) {
        const float f = (float)fitness[r];
        const float4 m = __ldcs(col + r * stride4);
        acc[0] = fmaf(__fsub_rn(m.x, t4.x), f, acc[0]);
        acc[1] = fmaf(__fsub_rn(m.y, t4.y), f, acc[1]);
        acc[2] = fmaf(__fsub_rn(m.z, t4.z), f, acc[2]);
        acc[3] = fmaf(__fsub_rn(m.w, t4.w), f, acc[3]);
    }
    *reinterpret_cast<float4*>(partial + (int64_t)sp * pitch + (int64_t)j4 * 4) =
        make_float4(acc[0], acc[1], acc[2], acc[3]);
}

// coefficient lr / (n_total * sigma) in fp64 like the reference's Python float, then one fp32 multiply
__global__ void __launch_bounds__(PT) es_update_finish_kernel(
    const float* __restrict__ partial, int n_split, int64_t pitch, double lr, double n_total, float sigma_arg,
    const double* __restrict__ sigma_dev, float* __restrict__ delta) {
    const int64_t j = (int64_t)blockIdx.x * PT + threadIdx.x;
    if (j >= pitch) return;
    const float coef = (float)(lr / (n_total * (double)pick_sigma(sigma_arg, sigma_dev)));
    float s = 0.f;
    for (int sp = 0; sp < n_split; ++sp) s += partial[(int64_t)sp * pitch + j];
    delta[j] = __fmul_rn(coef, s);
}

__global__ void __launch_bounds__(PT) axpy_kernel(float a, const float* __restrict__ x,
                                                  float* __restrict__ y, int64_t n) {
    const int64_t j = (int64_t)blockIdx.x * PT + threadIdx.x;
    if (j < n) y[j] = __fadd_rn(y[j], __fmul_rn(a, x[j]));
}

// ---------------------------------------------------------------------------
// K7  fitness-sharing distances   (utils/game_logic_functions.py:12-21)
// d[i] = || pop[i] - ref ||_2 over the Linear parameters; one CTA per row.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(PT) diversity_dist_kernel(
    const float* __restrict__ pop, int64_t pitch, const float* __restrict__ ref, int in_dim,
    float* __restrict__ dist) {
    const FcOffsets o = fc_offsets(in_dim);
    const int64_t r = blockIdx.x;
    const float4* row = reinterpret_cast<const float4*>(pop + r * pitch);
    const float4* rf = reinterpret_cast<const float4*>(ref);
    float acc = 0.f;
    // every segment boundary is a multiple of 4 parameters except the end of the row: a float4 is
    // perturbable or not as a whole; four independent 128-bit loads in flight per thread
    const int n4 = (int)(pitch / 4);
#pragma unroll 4
    for (int j4 = threadIdx.x; j4 < n4; j4 += PT) {
        const int j = j4 * 4;
        if (j >= o.total || !fc_is_perturbable(o, j)) continue;
        const float4 a = __ldcs(row + j4);
        const float4 b = __ldg(rf + j4);
        const float d[4] = {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (j + i < o.total) acc = fmaf(d[i], d[i], acc);
    }
    __shared__ float red[PT / 32];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < PT / 32; ++w) s += red[w];
        dist[r] = sqrtf(s);
    }
}

// ---------------------------------------------------------------------------
// K4  selection   (genetic_algorithm.py:223-234)
// k rounds of a block-wide arg-max over the not-yet-chosen members (a byte mask in the workspace, so
// k is unbounded).  Two total orders:
//   order 0: value descending, ties -> LOWER index, NaN ranks last  (= argsort(-f, kind="stable"))
//   order 1: the reference's expression np.argsort(f)[::-1] with a stable ascending sort (what
//            NumPy's small-array insertion sort does): value descending, ties -> HIGHER index, NaN
//            ranks FIRST (Appendix C #18).
// Single CTA: P <= a few 1e5 and a handful of elites is latency-bound.
// ---------------------------------------------------------------------------
constexpr int SEL_T = 1024;

__device__ __forceinline__ bool sel_better(double va, int64_t ia, double vb, int64_t ib, int order) {
    // is (va, ia) ranked before (vb, ib)?
    if (ib < 0) return ia >= 0;
    if (ia < 0) return false;
    const bool na = va != va, nb = vb != vb;
    if (na || nb) {
        if (na != nb) return order ? na : nb;            // order 1: NaN first; order 0: NaN last
        return order ? ia > ib : ia < ib;
    }
    if (va > vb) return true;
    if (va < vb) return false;
    return order ? ia > ib : ia < ib;
}

__global__ void __launch_bounds__(SEL_T) select_topk_kernel(const double* __restrict__ fitness, int64_t P,
                                                            int k, int order, uint8_t* __restrict__ taken,
                                                            int64_t* __restrict__ idx_out) {
    __shared__ double wv[SEL_T / 32];
    __shared__ int64_t wi[SEL_T / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t i = threadIdx.x; i < P; i += SEL_T) taken[i] = 0;
    __syncthreads();
    for (int round = 0; round < k; ++round) {
        double bv = 0.0;
        int64_t bi = -1;
        for (int64_t i = threadIdx.x; i < P; i += SEL_T) {
            if (taken[i]) continue;
            const double v = fitness[i];
            if (sel_better(v, i, bv, bi, order)) { bv = v; bi = i; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (sel_better(ov, oi, bv, bi, order)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { wv[warp] = bv; wi[warp] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double v = wv[0];
            int64_t ix = wi[0];
            for (int w = 1; w < SEL_T / 32; ++w)
                if (sel_better(wv[w], wi[w], v, ix, order)) { v = wv[w]; ix = wi[w]; }
            taken[ix] = 1;
            idx_out[round] = ix;
        }
        __syncthreads();
    }
}

// dst[r] = src[idx[r] - row0] when the GLOBAL row id idx[r] lies in this rank's block
// [row0, row0 + n_local), zeros otherwise: summed over ranks (all-reduce) the result is the
// elite / Hall-of-Fame broadcast with no host round trip for the indices.  n_local < 0: no range
// (plain gather of local ids).
__global__ void __launch_bounds__(PT) gather_rows_kernel(const float* __restrict__ src, int64_t pitch,
                                                         const int64_t* __restrict__ idx, int64_t row0,
                                                         int64_t n_local, float* __restrict__ dst) {
    const int64_t r = blockIdx.x;
    const int64_t s = idx[r] - row0;
    const int j4 = blockIdx.y * PT + threadIdx.x;
    if ((int64_t)j4 * 4 >= pitch) return;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n_local < 0 || (s >= 0 && s < n_local)) v = *reinterpret_cast<const float4*>(src + s * pitch + (int64_t)j4 * 4);
    *reinterpret_cast<float4*>(dst + r * pitch + (int64_t)j4 * 4) = v;
}

// ---------------------------------------------------------------------------
// Per-member statistics of the perturbable weights (MPEAgent.log_weight_statistics,
// MPE/mpe_agent.py:30-50: mean, min, max, population std of get_perturbable_weights()).
// One CTA per row, HBM-read bound (4 B/param/member); out fp32 [n_rows][4] = mean, min, max, std.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(PT) weight_stats_kernel(const float* __restrict__ rows, int64_t pitch, int in_dim,
                                                          float* __restrict__ out) {
    const FcOffsets o = fc_offsets(in_dim);
    const int64_t r = blockIdx.x;
    const float* row = rows + r * pitch;
    double sum = 0.0, sq = 0.0;
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    const int n4 = (int)(pitch / 4);
#pragma unroll 4
    for (int j4 = threadIdx.x; j4 < n4; j4 += PT) {
        const int j = j4 * 4;
        if (j >= o.total || !fc_is_perturbable(o, j)) continue;
        const float4 a = __ldcs(reinterpret_cast<const float4*>(row) + j4);
        const float v[4] = {a.x, a.y, a.z, a.w};
        // fp32 partial sums of one float4, fp64 across float4s (138 k terms per row)
        float s4 = 0.f, q4 = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (j + i < o.total) {
                s4 += v[i];
                q4 = fmaf(v[i], v[i], q4);
                mn = fminf(mn, v[i]);
                mx = fmaxf(mx, v[i]);
            }
        }
        sum += (double)s4;
        sq += (double)q4;
    }
    __shared__ double rs[PT / 32], rq[PT / 32];
    __shared__ float rmn[PT / 32], rmx[PT / 32];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, off);
        sq += __shfl_xor_sync(0xffffffffu, sq, off);
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    if ((threadIdx.x & 31) == 0) {
        rs[threadIdx.x >> 5] = sum; rq[threadIdx.x >> 5] = sq;
        rmn[threadIdx.x >> 5] = mn; rmx[threadIdx.x >> 5] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < PT / 32; ++w) {
            sum += rs[w]; sq += rq[w];
            mn = fminf(mn, rmn[w]); mx = fmaxf(mx, rmx[w]);
        }
        const double n = (double)(o.total - 2 * (H1 + H2));
        const double mean = sum / n;
        const double var = fmax(sq / n - mean * mean, 0.0);
        float4 res = make_float4((float)mean, mn, mx, (float)sqrt(var));
        *reinterpret_cast<float4*>(out + r * 4) = res;
    }
}

// ---------------------------------------------------------------------------
// End of a generation, on the device: the mean reward triple of the evaluation
// games (evaluate_current_weights, genetic_algorithm.py:12-29 == evolutionary_strategy.py:22-59),
// the reward history, the adaptive mutation power (genetic_algorithm.py:323-345 ==
// evolutionary_strategy.py:292-316, including agent_0 growing from sigma_agent_1 * 1.2, Appendix C #6)
// and the early-stopping counters (evolutionary_strategy.py:318-354).  One thread: it is ~100 scalar
// operations, and what matters is that nothing leaves the device between generations.
// gstate (fp64) layout: see CEV_GS_* in the header.
// ---------------------------------------------------------------------------
// NumPy's pairwise summation for n < 128 (np.mean of a list of Python floats): eight running sums
// for n >= 8, a plain loop below that.
__device__ double np_mean(const double* a, int n) {
    if (n <= 0) return CUDART_NAN;
    double res;
    if (n < 8) {
        res = 0.0;
        for (int i = 0; i < n; ++i) res += a[i];
    } else {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
    }
    return res / (double)n;
}

__global__ void generation_end_kernel(const double* __restrict__ eval_out, int n_games, int agent_step_limit,
                                      int reference_compat, double* __restrict__ gs, int hist_cap, int adaptive,
                                      double sigma_max, double sigma_min, int early_stopping, double min_delta,
                                      int patience) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int gen = (int)gs[CEV_GS_GEN];
    // ---- reward slots of play_MPE (utils/game_logic_functions.py:179-190, Appendix B) ----------
    const int L = agent_step_limit < 0 ? 75 : min(75, max(0, agent_step_limit));
    const int nc = L / 3;
    const int n_adv = (L + 2) / 3, n_a0 = (L + 1) / 3;     // ceil(L/3), ceil((L-1)/3) for L >= 1
    double tot[3] = {0.0, 0.0, 0.0};
    for (int g = 0; g < n_games; ++g) {
        const double sg = eval_out[g * CEV_ROLLOUT_OUT_DIM + 0], lg = eval_out[g * CEV_ROLLOUT_OUT_DIM + 1],
                     sa = eval_out[g * CEV_ROLLOUT_OUT_DIM + 2];
        double s0, s1, sadv;
        if (!reference_compat) { s0 = sg; s1 = sg; sadv = sa; }
        else if (nc == 0) { s0 = 0.0; s1 = sa; sadv = 0.0; }
        else {
            sadv = (n_adv - 1 >= nc) ? sg : sg - lg;
            s0 = (n_a0 - 1 >= nc) ? sg : sg - lg;
            s1 = sa;
        }
        tot[0] += s0; tot[1] += s1; tot[2] += sadv;       // the reference's running sums, game order
    }
    double ev[3];
    for (int r = 0; r < 3; ++r) ev[r] = tot[r] / (double)n_games;
    double* hist = gs + CEV_GS_HIST;                      // [hist_cap][3]
    double* shist = hist + (size_t)3 * hist_cap;          // [hist_cap + 1][3], entry 0 = initial sigma
    if (gen < hist_cap)
        for (int r = 0; r < 3; ++r) hist[(size_t)gen * 3 + r] = ev[r];
    for (int r = 0; r < 3; ++r) gs[CEV_GS_LAST_EVAL + r] = ev[r];
    // ---- adaptive mutation power ------------------------------------------------------------------
    if (adaptive) {
        double sig[3] = {gs[CEV_GS_SIGMA + 0], gs[CEV_GS_SIGMA + 1], gs[CEV_GS_SIGMA + 2]};
        const int n = gen + 1;                            // rewards recorded so far
        for (int r = 0; r < 3; ++r) {
            bool worse = false;
            if (gen > 10 && n <= hist_cap) {
                // lists sliced like Python: h[-10:] and h[-20:-10]
                double last[10], prev[10];
                const int n_last = min(n, 10);
                for (int i = 0; i < n_last; ++i) last[i] = hist[(size_t)(n - n_last + i) * 3 + r];
                const int lo = max(0, n - 20), hi = max(0, n - 10);
                const int n_prev = hi - lo;
                for (int i = 0; i < n_prev; ++i) prev[i] = hist[(size_t)(lo + i) * 3 + r];
                worse = np_mean(last, n_last) < np_mean(prev, n_prev);
            }
            // agent_0 (r = 0) grows from sigma_agent_1 (already this generation's old value: agent_0 is
            // updated first), Appendix C #6
            if (worse) sig[r] = fmin((r == 0 ? sig[1] : sig[r]) * 1.2, sigma_max);
            else sig[r] = fmax(sig[r] * 0.95, sigma_min);
        }
        for (int r = 0; r < 3; ++r) gs[CEV_GS_SIGMA + r] = sig[r];
    }
    if (gen + 1 <= hist_cap)
        for (int r = 0; r < 3; ++r) shist[(size_t)(gen + 1) * 3 + r] = gs[CEV_GS_SIGMA + r];
    // ---- early stopping ---------------------------------------------------------------------------
    if (early_stopping && gs[CEV_GS_STOP] == 0.0) {
        for (int r = 0; r < 3; ++r) {
            if (ev[r] > gs[CEV_GS_BEST + r] + min_delta) {
                gs[CEV_GS_BEST + r] = ev[r];
                gs[CEV_GS_STALE + r] = 0.0;
            } else {
                gs[CEV_GS_STALE + r] += 1.0;
            }
        }
        for (int r = 0; r < 3; ++r)
            if (gs[CEV_GS_STALE + r] >= (double)patience) {
                gs[CEV_GS_STOP] = (double)(1 + r);        // first role in agent_0, agent_1, adversary order
                gs[CEV_GS_STOP_GEN] = (double)gen;
                break;
            }
    }
    gs[CEV_GS_GEN] = (double)(gen + 1);
}

// ---------------------------------------------------------------------------
// synthetic inputs
// ---------------------------------------------------------------------------
__device__ __forceinline__ double u_pm1(uint32_t x) {
    // uniform in (-1, 1): ((x + 0.5) * 2^-32) * 2 - 1, exact in fp64
    return ((double)x + 0.5) * (2.0 / 4294967296.0) - 1.0;
}

__global__ void __launch_bounds__(PT) init_states_kernel(uint32_t k0, uint32_t k1, uint32_t stream_id,
                                                         int64_t rec0, int64_t n, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * PT + threadIdx.x;
    if (i >= n) return;
    const int64_t r = rec0 + i;
    const uint32_t tag = noise_tag(CEV_KIND_ENV, 0);
    uint32_t w[12];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        const U4 v = philox4x32_10(U4{(uint32_t)b, (uint32_t)r, stream_id, tag}, k0, k1);
        w[4 * b] = v.x; w[4 * b + 1] = v.y; w[4 * b + 2] = v.z; w[4 * b + 3] = v.w;
    }
    double* o = out + i * CEV_INIT_STATE_DIM;
    o[0] = (double)(w[0] & 1u);
#pragma unroll
    for (int i = 0; i < 10; ++i) o[1 + i] = u_pm1(w[1 + i]);
}

// Founder initialisation on the device (create_agent -> FCNetwork default init,
// MPE/fcnetwork.py:11-22): nn.Linear weight and bias ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)),
// LayerNorm gamma = 1, beta = 0.  Same distribution as PyTorch's default, Philox stream
// (kind 4) instead of torch's global generator; needed when P is too large to build
// 3*P nn.Modules on the host.
__global__ void __launch_bounds__(PT) fc_init_kernel(int in_dim, int64_t pitch, uint32_t k0, uint32_t k1,
                                                     uint32_t tag, int64_t row0, float* __restrict__ out) {
    const FcOffsets o = fc_offsets(in_dim);
    const int64_t r = blockIdx.x;
    const int j4 = blockIdx.y * PT + threadIdx.x;
    if ((int64_t)j4 * 4 >= pitch) return;
    const U4 w = philox4x32_10(U4{(uint32_t)j4, (uint32_t)(row0 + r), 0u, tag}, k0, k1);
    const uint32_t x[4] = {w.x, w.y, w.z, w.w};
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int j = j4 * 4 + i;
        if (j >= o.total) v[i] = 0.f;
        else if (j >= o.ln1g && j < o.ln1b) v[i] = 1.f;
        else if (j >= o.ln1b && j < o.fc2w) v[i] = 0.f;
        else if (j >= o.ln2g && j < o.ln2b) v[i] = 1.f;
        else if (j >= o.ln2b && j < o.outw) v[i] = 0.f;
        else {
            const int fan_in = j < o.ln1g ? in_dim : (j < o.ln2g ? H1 : H2);
            const double bound = 1.0 / sqrt((double)fan_in);
            const double u = ((double)x[i] + 0.5) * (2.0 / 4294967296.0) - 1.0;
            v[i] = (float)(u * bound);
        }
    }
    *reinterpret_cast<float4*>(out + r * pitch + (int64_t)j4 * 4) = make_float4(v[0], v[1], v[2], v[3]);
}

__global__ void __launch_bounds__(PT) random_frames_kernel(uint32_t k0, uint32_t k1, int64_t n16,
                                                           uint4* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * PT + threadIdx.x;
    if (i >= n16) return;
    const U4 v = philox4x32_10(U4{(uint32_t)i, (uint32_t)(i >> 32), 0u, noise_tag(CEV_KIND_FRAMES, 0)}, k0, k1);
    out[i] = make_uint4(v.x, v.y, v.z, v.w);
}

__global__ void __launch_bounds__(PT) philox_words_kernel(uint32_t k0, uint32_t k1, uint32_t tag, uint32_t gen,
                                                          int64_t member0, int64_t n_members, int64_t n4,
                                                          uint4* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * PT + threadIdx.x;
    if (i >= n_members * n4) return;
    const int64_t m = i / n4, j4 = i % n4;
    const U4 v = philox4x32_10(U4{(uint32_t)j4, (uint32_t)(member0 + m), gen, tag}, k0, k1);
    out[i] = make_uint4(v.x, v.y, v.z, v.w);
}

// ---------------------------------------------------------------------------
// FP32 FMA-pipe peak (roofline denominator for K1)
// ---------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, int iters) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, blockIdx.x * 1e-3f - i);
    const float2 b = make_float2(0.999f, 1.001f), c = make_float2(1e-3f, -1e-3f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 1) a[i] = __ffma2_rn(a[i], b, c);
            else { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    if (s == 123.456f) out[0] = s;       // keep the chain alive
}

}  // namespace cev

using namespace cev;

static inline void split_seed(uint64_t seed, uint32_t& k0, uint32_t& k1) {
    k0 = (uint32_t)(seed & 0xFFFFFFFFull);
    k1 = (uint32_t)(seed >> 32);
}

static int ensure_workspace(cev_handle* h, size_t bytes) {
    if (h->workspace_bytes >= bytes) return CEV_OK;
    if (h->workspace) cudaFree(h->workspace);
    h->workspace = nullptr;
    h->workspace_bytes = 0;
    CEV_CUDA(cudaMalloc(&h->workspace, bytes));
    h->workspace_bytes = bytes;
    return CEV_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" {

int cev_ga_repopulate_f32(cev_handle* h, const float* elites, int E, int D, int64_t pitch, float sigma,
                          const double* sigma_dev, float crossover_rate, uint64_t seed, int role, uint32_t gen,
                          int64_t row0, int64_t n_rows, float* out, float* noise_out, cev_stream stream) {
    CEV_REQUIRE(h != nullptr, "ga_repopulate: null handle");
    if (n_rows == 0) return CEV_OK;                   // an empty shard contributes nothing
    CEV_REQUIRE(elites && out, "ga_repopulate: null pointer");
    CEV_REQUIRE(E >= 1 && D >= 1 && pitch >= D && pitch % 4 == 0, "ga_repopulate: bad E/D/pitch");
    CEV_REQUIRE(crossover_rate >= 0.f && crossover_rate <= 1.f, "ga_repopulate: crossover_rate must be in [0, 1]");
    CEV_REQUIRE(aligned16(elites) && aligned16(out) && aligned16(noise_out), "ga_repopulate: 16B alignment");
    CEV_REQUIRE(row0 >= 0 && n_rows >= 0 && row0 + n_rows <= 0xFFFFFFFFll, "ga_repopulate: bad row range");
    CEV_GUARD(h);
    uint32_t k0, k1;
    split_seed(seed, k0, k1);
    dim3 grid((unsigned)n_rows, (unsigned)((pitch / 4 + PT - 1) / PT));
    ga_repopulate_kernel<<<grid, PT, 0, (cudaStream_t)stream>>>(elites, E, D, pitch, sigma, sigma_dev, philox_keys(k0, k1),
                                                               noise_tag(CEV_KIND_GA, role), gen, row0, crossover_rate,
                                                               noise_tag(CEV_KIND_XOVER, role), out, noise_out);
    return check_cuda(cudaGetLastError(), "ga_repopulate_kernel");
}

int cev_gather_rows_f32(cev_handle* h, const float* src, int64_t pitch, const int64_t* idx, int n, int64_t row0,
                        int64_t n_local, float* dst, cev_stream stream) {
    CEV_REQUIRE(h != nullptr, "gather_rows: null handle");
    if (n <= 0) return CEV_OK;
    // an empty shard (n_local == 0) has no rows to read: src may be null, dst is zero-filled
    CEV_REQUIRE((src || n_local == 0) && idx && dst, "gather_rows: null pointer");
    CEV_REQUIRE(pitch % 4 == 0 && aligned16(src) && aligned16(dst), "gather_rows: alignment");
    CEV_GUARD(h);
    dim3 grid((unsigned)n, (unsigned)((pitch / 4 + PT - 1) / PT));
    gather_rows_kernel<<<grid, PT, 0, (cudaStream_t)stream>>>(src, pitch, idx, row0, n_local, dst);
    return check_cuda(cudaGetLastError(), "gather_rows_kernel");
}

int cev_select_topk_f64(cev_handle* h, const double* fitness, int64_t P, int k, int order, int64_t* idx_out,
                        cev_stream stream) {
    CEV_REQUIRE(h && fitness && idx_out, "select_topk: null pointer");
    CEV_REQUIRE(k >= 1 && k <= P, "select_topk: need 1 <= k <= P (elites_number %d, population %lld)", k, (long long)P);
    CEV_REQUIRE(order == 0 || order == 1, "select_topk: order must be 0 (ties -> lower index) or 1 (reference)");
    CEV_GUARD(h);
    int rc = ensure_workspace(h, (size_t)P + 256);
    if (rc) return rc;
    select_topk_kernel<<<1, SEL_T, 0, (cudaStream_t)stream>>>(fitness, P, k, order, static_cast<uint8_t*>(h->workspace),
                                                              idx_out);
    return check_cuda(cudaGetLastError(), "select_topk_kernel");
}

int cev_es_perturb_f32(cev_handle* h, const float* theta, int in_dim, float sigma, const double* sigma_dev,
                       uint64_t seed, int role, uint32_t gen, int64_t row0, int64_t n_rows, int64_t pitch,
                       float* out, float* noise_out, cev_stream stream) {
    CEV_REQUIRE(h != nullptr, "es_perturb: null handle");
    if (n_rows == 0) return CEV_OK;                   // an empty shard contributes nothing
    CEV_REQUIRE(theta && out, "es_perturb: null pointer");
    CEV_REQUIRE(in_dim == IN_ADV || in_dim == IN_GOOD, "es_perturb: in_dim must be 8 or 10");
    CEV_REQUIRE(pitch >= fc_offsets(in_dim).total && pitch % 4 == 0, "es_perturb: bad pitch");
    CEV_REQUIRE(aligned16(theta) && aligned16(out) && aligned16(noise_out), "es_perturb: 16B alignment");
    CEV_REQUIRE(row0 >= 0 && n_rows >= 0 && row0 + n_rows <= 0xFFFFFFFFll, "es_perturb: bad row range");
    CEV_GUARD(h);
    uint32_t k0, k1;
    split_seed(seed, k0, k1);
    dim3 grid((unsigned)n_rows, (unsigned)((pitch / 4 + K5T - 1) / K5T));
    es_perturb_kernel<<<grid, K5T, 0, (cudaStream_t)stream>>>(theta, in_dim, pitch, sigma, sigma_dev, philox_keys(k0, k1),
                                                            noise_tag(CEV_KIND_ES, role), gen, row0, out,
                                                            noise_out);
    return check_cuda(cudaGetLastError(), "es_perturb_kernel");
}

int cev_es_update_f32(cev_handle* h, const double* fitness, int in_dim, float sigma, const double* sigma_dev,
                      float lr, int64_t n_total, uint64_t seed, int role, uint32_t gen, int64_t row0,
                      int64_t n_rows, float* delta, cev_stream stream) {
    // an empty shard (n_rows == 0, fitness may be null) still writes its all-zero partial delta
    CEV_REQUIRE(h && (fitness || n_rows == 0) && delta, "es_update: null pointer");
    CEV_REQUIRE(in_dim == IN_ADV || in_dim == IN_GOOD, "es_update: in_dim must be 8 or 10");
    CEV_REQUIRE(n_total >= 1 && n_rows >= 0 && (sigma_dev || sigma > 0.f), "es_update: bad n/sigma");
    CEV_REQUIRE(aligned16(delta), "es_update: 16B alignment");
    CEV_GUARD(h);
    const int64_t pitch = fc_pitch(in_dim);
    int n_split = (int)((n_rows + 63) / 64);
    if (n_split > 16) n_split = 16;
    if (n_split < 1) n_split = 1;
    int rc = ensure_workspace(h, (size_t)n_split * pitch * sizeof(float));
    if (rc) return rc;
    uint32_t k0, k1;
    split_seed(seed, k0, k1);
    float* partial = static_cast<float*>(h->workspace);
    dim3 grid((unsigned)((pitch / 4 + PT - 1) / PT), (unsigned)n_split);
    es_update_partial_kernel<<<grid, PT, 0, (cudaStream_t)stream>>>(fitness, in_dim, pitch, sigma, sigma_dev, philox_keys(k0, k1),
                                                                   noise_tag(CEV_KIND_ES, role), gen, row0,
                                                                   n_rows, n_split, partial);
    CEV_CUDA(cudaGetLastError());
    es_update_finish_kernel<<<(unsigned)((pitch + PT - 1) / PT), PT, 0, (cudaStream_t)stream>>>(
        partial, n_split, pitch, (double)lr, (double)n_total, sigma, sigma_dev, delta);
    return check_cuda(cudaGetLastError(), "es_update kernels");
}

int cev_es_perturb_prefix_f32(cev_handle* h, const float* theta, int64_t d_pert, int64_t d_total, float sigma,
                              const double* sigma_dev, uint64_t seed, int role, uint32_t gen, int64_t row0,
                              int64_t n_rows, int64_t pitch, float* out, cev_stream stream) {
    CEV_REQUIRE(h != nullptr, "es_perturb_prefix: null handle");
    if (n_rows == 0) return CEV_OK;
    CEV_REQUIRE(theta && out, "es_perturb_prefix: null pointer");
    CEV_REQUIRE(d_pert >= 0 && d_pert <= d_total && d_total <= pitch && pitch % 4 == 0 && pitch / 4 <= 0xFFFFFFFFll,
                "es_perturb_prefix: need 0 <= d_pert <= d_total <= pitch, pitch a multiple of 4");
    CEV_REQUIRE(aligned16(theta) && aligned16(out), "es_perturb_prefix: 16B alignment");
    CEV_REQUIRE(row0 >= 0 && n_rows >= 0 && row0 + n_rows <= 0xFFFFFFFFll, "es_perturb_prefix: bad row range");
    CEV_GUARD(h);
    uint32_t k0, k1;
    split_seed(seed, k0, k1);
    dim3 grid((unsigned)n_rows, (unsigned)((pitch / 4 + PT - 1) / PT));
    es_perturb_prefix_kernel<<<grid, PT, 0, (cudaStream_t)stream>>>(theta, d_pert, d_total, pitch, sigma, sigma_dev, philox_keys(k0, k1),
                                                                   noise_tag(CEV_KIND_ES, role), gen, row0, out);
    return check_cuda(cudaGetLastError(), "es_perturb_prefix_kernel");
}

int cev_es_update_members_f32(cev_handle* h, const double* fitness, const float* members, int64_t pitch,
                              const float* theta, int in_dim, float sigma, const double* sigma_dev, float lr,
                              int64_t n_total, int64_t n_rows, float* delta, cev_stream stream) {
    // an empty shard (n_rows == 0: fitness / members may be null) still writes its all-zero partial delta
    CEV_REQUIRE(h && ((fitness && members) || n_rows == 0) && theta && delta, "es_update_members: null pointer");
    CEV_REQUIRE(n_total >= 1 && n_rows >= 0 && (sigma_dev || sigma > 0.f), "es_update_members: bad n/sigma");
    // in_dim 8 / 10: FCNetwork rows (pitch checked); in_dim 0: any row layout of `pitch` floats (DeepQN):
    // the kernel is layout agnostic, unperturbed entries are copies of theta and contribute exact zeros
    CEV_REQUIRE(in_dim == 0 || in_dim == IN_ADV || in_dim == IN_GOOD, "es_update_members: in_dim must be 0, 8 or 10");
    CEV_REQUIRE(in_dim == 0 ? (pitch > 0 && pitch % 4 == 0) : pitch == fc_pitch(in_dim),
                "es_update_members: rows must use the padded pitch (cev_fc_pitch / cev_dqn_pitch)");
    CEV_REQUIRE(aligned16(delta) && aligned16(members) && aligned16(theta), "es_update_members: 16B alignment");
    CEV_GUARD(h);
    int n_split = (int)((n_rows + 31) / 32);
    if (n_split > 32) n_split = 32;
    if (n_split < 1) n_split = 1;
    int rc = ensure_workspace(h, (size_t)n_split * pitch * sizeof(float));
    if (rc) return rc;
    float* partial = static_cast<float*>(h->workspace);
    dim3 grid((unsigned)((pitch / 4 + PT - 1) / PT), (unsigned)n_split);
    es_update_members_partial_kernel<<<grid, PT, 0, (cudaStream_t)stream>>>(fitness, members, theta, pitch, n_rows,
                                                                           n_split, partial);
    CEV_CUDA(cudaGetLastError());
    es_update_finish_kernel<<<(unsigned)((pitch + PT - 1) / PT), PT, 0, (cudaStream_t)stream>>>(
        partial, n_split, pitch, (double)lr, (double)n_total, sigma, sigma_dev, delta);
    return check_cuda(cudaGetLastError(), "es_update_members kernels");
}

int cev_axpy_f32(cev_handle* h, float a, const float* x, float* y, int64_t n, cev_stream stream) {
    CEV_REQUIRE(h && x && y && n >= 0, "axpy: bad arguments");
    if (n == 0) return CEV_OK;
    CEV_GUARD(h);
    axpy_kernel<<<(unsigned)((n + PT - 1) / PT), PT, 0, (cudaStream_t)stream>>>(a, x, y, n);
    return check_cuda(cudaGetLastError(), "axpy_kernel");
}

int cev_diversity_dist_f32(cev_handle* h, const float* pop, int64_t n_rows, int64_t pitch, const float* ref,
                           int in_dim, float* dist, cev_stream stream) {
    CEV_REQUIRE(h != nullptr, "diversity_dist: null handle");
    if (n_rows <= 0) return CEV_OK;                   // an empty shard contributes nothing
    CEV_REQUIRE(pop && ref && dist, "diversity_dist: null pointer");
    CEV_REQUIRE(in_dim == IN_ADV || in_dim == IN_GOOD, "diversity_dist: in_dim must be 8 or 10");
    CEV_REQUIRE(pitch >= fc_offsets(in_dim).total && pitch % 4 == 0, "diversity_dist: bad pitch");
    CEV_REQUIRE(aligned16(pop) && aligned16(ref), "diversity_dist: 16B alignment");
    CEV_GUARD(h);
    diversity_dist_kernel<<<(unsigned)n_rows, PT, 0, (cudaStream_t)stream>>>(pop, pitch, ref, in_dim, dist);
    return check_cuda(cudaGetLastError(), "diversity_dist_kernel");
}

int cev_fc_init_f32(cev_handle* h, int in_dim, uint64_t seed, int role, int64_t row0, int64_t n_rows,
                    int64_t pitch, float* out, cev_stream stream) {
    CEV_REQUIRE(h && out, "fc_init: null pointer");
    CEV_REQUIRE(in_dim == IN_ADV || in_dim == IN_GOOD, "fc_init: in_dim must be 8 or 10");
    CEV_REQUIRE(pitch >= fc_offsets(in_dim).total && pitch % 4 == 0 && aligned16(out), "fc_init: bad pitch / alignment");
    CEV_REQUIRE(row0 >= 0 && n_rows >= 0 && row0 + n_rows <= 0xFFFFFFFFll, "fc_init: bad row range");
    if (n_rows == 0) return CEV_OK;
    CEV_GUARD(h);
    uint32_t k0, k1;
    split_seed(seed, k0, k1);
    dim3 grid((unsigned)n_rows, (unsigned)((pitch / 4 + PT - 1) / PT));
    fc_init_kernel<<<grid, PT, 0, (cudaStream_t)stream>>>(in_dim, pitch, k0, k1, noise_tag(4, role), row0, out);
    return check_cuda(cudaGetLastError(), "fc_init_kernel");
}

int cev_init_states_f64(cev_handle* h, uint64_t seed, uint32_t stream_id, int64_t rec0, int64_t n, double* out,
                        cev_stream stream) {
    CEV_REQUIRE(h && (out || n == 0) && n >= 0 && rec0 >= 0 && rec0 + n <= 0xFFFFFFFFll, "init_states: bad arguments");
    if (n == 0) return CEV_OK;
    CEV_GUARD(h);
    uint32_t k0, k1;
    split_seed(seed, k0, k1);
    init_states_kernel<<<(unsigned)((n + PT - 1) / PT), PT, 0, (cudaStream_t)stream>>>(k0, k1, stream_id, rec0, n,
                                                                                       out);
    return check_cuda(cudaGetLastError(), "init_states_kernel");
}

int cev_random_frames_u8(cev_handle* h, uint64_t seed, int64_t n_bytes, uint8_t* out, cev_stream stream) {
    CEV_REQUIRE(h && out && n_bytes >= 0 && n_bytes % 16 == 0 && aligned16(out), "random_frames: bad arguments");
    if (n_bytes == 0) return CEV_OK;
    CEV_GUARD(h);
    uint32_t k0, k1;
    split_seed(seed, k0, k1);
    const int64_t n16 = n_bytes / 16;
    random_frames_kernel<<<(unsigned)((n16 + PT - 1) / PT), PT, 0, (cudaStream_t)stream>>>(
        k0, k1, n16, reinterpret_cast<uint4*>(out));
    return check_cuda(cudaGetLastError(), "random_frames_kernel");
}

int cev_philox_words(cev_handle* h, uint64_t seed, int kind, int role, uint32_t gen, int64_t member0,
                     int64_t n_members, int64_t n4, uint32_t* out, cev_stream stream) {
    CEV_REQUIRE(h && out && n_members >= 0 && n4 >= 0 && aligned16(out), "philox_words: bad arguments");
    if (n_members * n4 == 0) return CEV_OK;
    CEV_GUARD(h);
    uint32_t k0, k1;
    split_seed(seed, k0, k1);
    const int64_t n = n_members * n4;
    philox_words_kernel<<<(unsigned)((n + PT - 1) / PT), PT, 0, (cudaStream_t)stream>>>(
        k0, k1, noise_tag(kind, role), gen, member0, n_members, n4, reinterpret_cast<uint4*>(out));
    return check_cuda(cudaGetLastError(), "philox_words_kernel");
}

int cev_weight_stats_f32(cev_handle* h, const float* rows, int64_t n_rows, int64_t pitch, int in_dim, float* out,
                         cev_stream stream) {
    CEV_REQUIRE(h != nullptr, "weight_stats: null handle");
    if (n_rows <= 0) return CEV_OK;
    CEV_REQUIRE(rows && out, "weight_stats: null pointer");
    CEV_REQUIRE(in_dim == IN_ADV || in_dim == IN_GOOD, "weight_stats: in_dim must be 8 or 10");
    CEV_REQUIRE(pitch >= fc_offsets(in_dim).total && pitch % 4 == 0 && aligned16(rows) && aligned16(out),
                "weight_stats: bad pitch / alignment");
    CEV_GUARD(h);
    weight_stats_kernel<<<(unsigned)n_rows, PT, 0, (cudaStream_t)stream>>>(rows, pitch, in_dim, out);
    return check_cuda(cudaGetLastError(), "weight_stats_kernel");
}

int cev_generation_state_doubles(int hist_capacity) {
    return hist_capacity < 0 ? -1 : CEV_GS_HIST + 3 * hist_capacity + 3 * (hist_capacity + 1);
}

int cev_generation_end_f64(cev_handle* h, const double* eval_out, int n_games, int agent_step_limit,
                           int reference_compat, double* gstate, int hist_capacity, int adaptive, double sigma_max,
                           double sigma_min, int early_stopping, double min_delta, int patience, cev_stream stream) {
    CEV_REQUIRE(h && eval_out && gstate, "generation_end: null pointer");
    CEV_REQUIRE(n_games >= 1 && hist_capacity >= 0, "generation_end: need n_games >= 1, hist_capacity >= 0");
    CEV_GUARD(h);
    generation_end_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(eval_out, n_games, agent_step_limit, reference_compat,
                                                             gstate, hist_capacity, adaptive, sigma_max, sigma_min,
                                                             early_stopping, min_delta, patience);
    return check_cuda(cudaGetLastError(), "generation_end_kernel");
}

int cev_fp32_peak(cev_handle* h, int mode, double* tflops, cev_stream stream) {
    CEV_REQUIRE(h && tflops, "fp32_peak: null pointer");
    CEV_GUARD(h);
    int rc = ensure_workspace(h, 256);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int iters = 4096, blocks = h->n_sm * 8;
    cudaEvent_t e0, e1;
    CEV_CUDA(cudaEventCreate(&e0));
    CEV_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CEV_CUDA(cudaEventRecord(e0, st));
        if (mode == 1) fp32_peak_kernel<1><<<blocks, 256, 0, st>>>((float*)h->workspace, iters);
        else fp32_peak_kernel<0><<<blocks, 256, 0, st>>>((float*)h->workspace, iters);
        CEV_CUDA(cudaEventRecord(e1, st));
        CEV_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        CEV_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double flops = (double)blocks * 256.0 * iters * 16.0 * 2.0;
    *tflops = flops / (best * 1e-3) / 1e12;
    return CEV_OK;
}

}  // extern "C"
