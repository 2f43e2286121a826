// Shared device/host helpers for the coevonet_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/coevonet_b200.h"

namespace cev {

// ---------------------------------------------------------------------------
// FCNetwork geometry (MPE/fcnetwork.py:11-22): in -> 512 -> LN -> ReLU -> 256
// -> LN -> ReLU -> 5.  Flat rows follow parameters() order (Appendix D).
// ---------------------------------------------------------------------------
constexpr int H1 = 512;
constexpr int H2 = 256;
constexpr int NACT = 5;
constexpr int IN_ADV = 8;
constexpr int IN_GOOD = 10;
constexpr float LN_EPS = 1e-5f;
constexpr int MAX_CYCLES = 25;

struct FcOffsets {
    int fc1w, fc1b, ln1g, ln1b, fc2w, fc2b, ln2g, ln2b, outw, outb, total;
};

__host__ __device__ constexpr FcOffsets fc_offsets(int in_dim) {
    FcOffsets o{};
    o.fc1w = 0;
    o.fc1b = H1 * in_dim;
    o.ln1g = o.fc1b + H1;
    o.ln1b = o.ln1g + H1;
    o.fc2w = o.ln1b + H1;
    o.fc2b = o.fc2w + H2 * H1;
    o.ln2g = o.fc2b + H2;
    o.ln2b = o.ln2g + H2;
    o.outw = o.ln2b + H2;
    o.outb = o.outw + NACT * H2;
    o.total = o.outb + NACT;
    return o;
}

__host__ __device__ constexpr int round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ constexpr int fc_pitch(int in_dim) { return round_up(fc_offsets(in_dim).total, 32); }

// true for Linear (perturbable) parameters, false for LayerNorm gamma/beta
// (MPE/fcnetwork.py:185-199 get_perturbable_layers)
__host__ __device__ inline bool fc_is_perturbable(const FcOffsets& o, int j) {
    return !((j >= o.ln1g && j < o.fc2w) || (j >= o.ln2g && j < o.outw));
}

__host__ __device__ constexpr int seat_in_dim(int seat) { return seat == 0 ? IN_ADV : IN_GOOD; }

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define CEV_CUDA(expr)                                        \
    do {                                                      \
        int _rc = ::cev::check_cuda((expr), #expr);           \
        if (_rc != 0) return _rc;                             \
    } while (0)

// Every entry point runs on the handle's device, whatever the caller's current device is, and leaves
// the caller's current device untouched (a process that drives several GPUs keeps torch's device).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define CEV_GUARD(h) ::cev::DeviceGuard _cev_device_guard((h)->device)

#define CEV_REQUIRE(cond, ...)                                \
    do {                                                      \
        if (!(cond)) {                                        \
            ::cev::set_error(__VA_ARGS__);                    \
            return CEV_ERR_ARG;                               \
        }                                                     \
    } while (0)

// ---------------------------------------------------------------------------
// simple_adversary_v3 world, fp64, op order identical
// to oracle/mpe_env.py so the environment is bit-exact given equal actions.
// Explicit _rn intrinsics keep nvcc from contracting mul+add into FMA.
// ---------------------------------------------------------------------------
struct EnvState {
    double px[3], py[3], vx[3], vy[3];   // adversary_0, agent_0, agent_1
    double lx[2], ly[2];                 // landmarks
    int goal;
};

__device__ __forceinline__ void env_load(EnvState& s, const double* rec) {
    s.goal = (int)rec[0];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        s.px[i] = rec[1 + 2 * i];
        s.py[i] = rec[2 + 2 * i];
        s.vx[i] = 0.0;
        s.vy[i] = 0.0;
    }
    s.lx[0] = rec[7];
    s.ly[0] = rec[8];
    s.lx[1] = rec[9];
    s.ly[1] = rec[10];
}

// Observation of seat `seat` into obs[0..in_dim) (Appendix A.6), fp64 diff -> fp32.
__device__ __forceinline__ void env_observe(const EnvState& s, int seat, float* obs) {
    const double mx = s.px[seat], my = s.py[seat];
    int o = 0;
    if (seat != 0) {
        const double gx = s.goal ? s.lx[1] : s.lx[0], gy = s.goal ? s.ly[1] : s.ly[0];
        obs[o++] = __double2float_rn(__dsub_rn(gx, mx));
        obs[o++] = __double2float_rn(__dsub_rn(gy, my));
    }
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        obs[o++] = __double2float_rn(__dsub_rn(s.lx[l], mx));
        obs[o++] = __double2float_rn(__dsub_rn(s.ly[l], my));
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        if (j == seat) continue;
        obs[o++] = __double2float_rn(__dsub_rn(s.px[j], mx));
        obs[o++] = __double2float_rn(__dsub_rn(s.py[j], my));
    }
}

__device__ __forceinline__ double env_dist_goal(const EnvState& s, int i) {
    const double gx = s.goal ? s.lx[1] : s.lx[0], gy = s.goal ? s.ly[1] : s.ly[0];
    const double dx = __dsub_rn(s.px[i], gx);
    const double dy = __dsub_rn(s.py[i], gy);
    return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

// One world step (Appendix A.2/A.4) + rewards (A.5).
__device__ __forceinline__ void env_step(EnvState& s, const int act[3], bool pos_first,
                                         double& r_good, double& r_adv) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int a = act[i];
        double ux = (a == 1) ? -1.0 : ((a == 2) ? 1.0 : 0.0);
        double uy = (a == 3) ? -1.0 : ((a == 4) ? 1.0 : 0.0);
        ux = __dmul_rn(ux, 5.0);
        uy = __dmul_rn(uy, 5.0);
        if (pos_first) {
            s.px[i] = __dadd_rn(s.px[i], __dmul_rn(s.vx[i], 0.1));
            s.py[i] = __dadd_rn(s.py[i], __dmul_rn(s.vy[i], 0.1));
        }
        s.vx[i] = __dmul_rn(s.vx[i], 0.75);
        s.vy[i] = __dmul_rn(s.vy[i], 0.75);
        s.vx[i] = __dadd_rn(s.vx[i], __dmul_rn(ux, 0.1));
        s.vy[i] = __dadd_rn(s.vy[i], __dmul_rn(uy, 0.1));
        if (!pos_first) {
            s.px[i] = __dadd_rn(s.px[i], __dmul_rn(s.vx[i], 0.1));
            s.py[i] = __dadd_rn(s.py[i], __dmul_rn(s.vy[i], 0.1));
        }
    }
    const double d0 = env_dist_goal(s, 0);
    const double d1 = env_dist_goal(s, 1);
    const double d2 = env_dist_goal(s, 2);
    r_adv = -d0;
    r_good = __dadd_rn(-fmin(d1, d2), d0);
}

// Strict-'>' scan from index 0 (MPE/fcnetwork.py:78-85): lowest index wins ties.
// Also returns the top-2 gap (decision margin).
__device__ __forceinline__ int argmax_first5(const float lg[NACT], float& gap) {
    int best = 0;
    float bv = lg[0];
#pragma unroll
    for (int a = 1; a < NACT; ++a)
        if (lg[a] > bv) { bv = lg[a]; best = a; }
    float second = -CUDART_INF_F;
#pragma unroll
    for (int a = 0; a < NACT; ++a)
        if (a != best && lg[a] > second) second = lg[a];
    gap = bv - second;
    return best;
}

// ---------------------------------------------------------------------------
// Philox4x32-10 counter RNG (restated in oracle/philox.py; Random123 KATs in
// tests/test_philox.py).  ctr = (j4, member, gen, role | kind << 8).
// ---------------------------------------------------------------------------
struct U4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
#else
        const uint64_t p0 = (uint64_t)M0 * c.x, p1 = (uint64_t)M1 * c.z;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        c = U4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
        k0 += W0;
        k1 += W1;
    }
    return c;
}

// The ten round keys of a launch, precomputed on the host: as a kernel parameter they sit in the constant
// bank and cost the noise kernels no per-thread key-schedule adds.
struct PhiloxKeys {
    uint32_t a[10], b[10];
};
__host__ __device__ __forceinline__ PhiloxKeys philox_keys(uint32_t k0, uint32_t k1) {
    PhiloxKeys k;
    for (int r = 0; r < 10; ++r) {
        k.a[r] = k0 + (uint32_t)r * 0x9E3779B9u;
        k.b[r] = k1 + (uint32_t)r * 0xBB67AE85u;
    }
    return k;
}
__device__ __forceinline__ U4 philox4x32_10(U4 c, const PhiloxKeys& k) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c.x, p1 = (uint64_t)M1 * c.z;
        c = U4{(uint32_t)(p1 >> 32) ^ c.y ^ k.a[r], (uint32_t)p1, (uint32_t)(p0 >> 32) ^ c.w ^ k.b[r], (uint32_t)p0};
    }
    return c;
}

__device__ __forceinline__ float u01(uint32_t x) {
    // (0,1]: x * 2^-32 + 2^-33 ; the scaling is exact so FMA contraction is harmless
    return __fmaf_rn(__uint2float_rn(x), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
}

// Box-Muller on the SFU.  The normals of K3/K5/K6 are ALU bound (141 M per ES role and generation), so every
// transcendental is one MUFU instruction with the accuracy argued below (absolute error of a normal vs the
// fp64-accurate value: <= 3e-6 for |z| <= 5; tests/test_gpu_population.py states the bound); every consumer
// regenerates noise with this same function, so the streams agree bit for bit.
//  * r^2 = -2 ln u1: MUFU.LG2 (__log2f: absolute error 2^-22 on [0.5, 2], 2 ulp elsewhere) times ln 2 -- except
//    near u1 = 1, where the result is near 0 and an absolute error is a large RELATIVE one (r would be off by
//    up to 5e-4): there t = 1 - u1 is exact (Sterbenz) and -ln(1 - t) = t + t^2/2 + ... + t^5/5 is good to
//    t^5/6 < 6e-9 relative for t < 1/32.  Outside that window |ln u1| >= 0.0317, so dr <= 2^-22 ln2 / r <= 7e-7.
//  * r = sqrt.approx (MUFU.SQRT, relative error 2^-23).
//  * sine / cosine: MUFU.SIN / MUFU.COS are accurate on [-pi, pi] (absolute error 2^-21.4): evaluate at
//    2 pi u2 - pi and flip both signs.
__device__ __forceinline__ float neg2_log_u(float u) {
    const float t = 1.0f - u;
    const float ser = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 0.2f, 0.25f), 0.33333334f), 0.5f), 1.0f);
    const float lg = -0.6931471805599453f * __log2f(u);
    return 2.0f * (t < 0.03125f ? ser : lg);
}

__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, float& z0, float& z1) {
    const float u1 = u01(xa), u2 = u01(xb);
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;\n" : "=f"(r) : "f"(neg2_log_u(u1)));
    float s, c;
    __sincosf(6.283185307179586f * (u2 - 0.5f), &s, &c);
    z0 = -(r * c);
    z1 = -(r * s);
}

// four standard normals for flat parameter indices 4*j4 .. 4*j4+3 of `member`
__device__ __forceinline__ void normal4(const PhiloxKeys& keys, uint32_t tag, uint32_t gen,
                                        uint32_t member, uint32_t j4, float z[4]) {
    const U4 r = philox4x32_10(U4{j4, member, gen, tag}, keys);
    box_muller(r.x, r.y, z[0], z[1]);
    box_muller(r.z, r.w, z[2], z[3]);
}

__host__ __device__ __forceinline__ uint32_t noise_tag(int kind, int role) {
    return (uint32_t)(role & 0xFF) | ((uint32_t)kind << 8);
}

}  // namespace cev

constexpr int CEV_TIMING_MAX = 4096;

#define CEV_MAX_ROLES 3

// per-handle state
struct cev_handle {
    int device;
    int n_sm;
    int n_clusters;        // co-resident 4-CTA clusters of the rollout kernel
    void* workspace;       // device scratch
    size_t workspace_bytes;
    void* opp_workspace;   // packed opponent fc2 matrices of the rollout kernel
    size_t opp_workspace_bytes;
    void* ls_workspace;    // episode state / split opponent weights of the lockstep rollout
    size_t ls_workspace_bytes;
    cudaStream_t side_stream;           // the opponent kernel of the lockstep rollout runs beside the member kernel
    cudaEvent_t fork_ev, join_ev;
    // several roles in one lockstep pass (launch_rollout_lockstep_roles): environment steps on a third stream
    cudaStream_t env_stream, opp_stream2[3], mem_stream2[2];
    cudaEvent_t ev_opp[CEV_MAX_ROLES], ev_mem[CEV_MAX_ROLES], ev_env[CEV_MAX_ROLES];
    // optional per-kernel timing of the lockstep rollout (cev_kernel_timing_*): CUDA events recorded
    // around every member / opponent kernel launch on the launch stream
    int timing_on;
    int timing_n[2];                    // events pairs used so far: 0 = member kernel, 1 = opponent kernel
    cudaEvent_t* timing_ev[2];          // [CEV_TIMING_MAX][2] each
};

// ---------------------------------------------------------------------------
// K1 launch parameter blocks (shared by api.cu and the kernel TUs)
// ---------------------------------------------------------------------------
namespace cev {
struct GenericParams {
    const float* w[3];          // per seat
    int64_t pitch[3];
    const int32_t* idx;         // [N,3] or null (structured indexing)
    // structured indexing: episode e = ((m*K)+k)*E + i ; member seat ms
    int member_seat, K, E, init_shared;
    const double* init;
    double* out;
    int32_t* status;
    int n_cycles, pos_first;
    int64_t N;
};
struct ClusterParams {
    const float* members;
    int64_t member_pitch;
    int P;
    const float* opp[2];          // the two non-member seats, ascending seat order
    int64_t opp_pitch[2];
    const float4* opp_packed[2];  // their fc2 matrices repacked into stage images (workspace)
    int K;
    int member_seat;
    const double* init;
    int init_shared;
    int E;
    double* out;
    int32_t* status;
    int n_cycles, pos_first;
    // parity instrumentation of the lockstep form (cev_mpe_rollout_trace_f32), null in production
    const int32_t* trace_forced;  // [n_cycles][3][N] actions to replay
    float* trace_logits;          // [n_cycles][3][N][5]
    int32_t* trace_actions;       // [n_cycles][3][N] the networks' own decisions
};
int launch_rollout_generic(cev_handle* h, const GenericParams& p, cudaStream_t stream);
int launch_rollout_cluster(cev_handle* h, const ClusterParams& p, cudaStream_t stream);
int rollout_cluster_max_clusters(int device);
int launch_rollout_lockstep(cev_handle* h, const ClusterParams& p, cudaStream_t stream);
int launch_rollout_lockstep_roles(cev_handle* h, const ClusterParams* ps, int n_roles, cudaStream_t stream);
int rollout_lockstep_launches(int n_cycles);
int launch_deepqn_fc_tc(cev_handle* h, const float* members, int64_t pitch, int P, int B, int n_act, int f1w_off,
                        int f1b_off, int ow_off, int ob_off, const float* act3, float* logits, int32_t* actions,
                        cudaStream_t stream);
int launch_fc_forward(cev_handle* h, const float* W, int64_t pitch, int in_dim, const int32_t* idx,
                      const float* obs, int64_t N, float* logits, int32_t* actions, int32_t* status,
                      cudaStream_t stream);
}  // namespace cev
