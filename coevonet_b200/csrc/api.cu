// C ABI: handle management, error text, K1 entry points (include/coevonet_b200.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace cev {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return CEV_OK;
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return CEV_ERR_CUDA;
}

}  // namespace cev

using namespace cev;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" {

int cev_version(void) { return 100; }

const char* cev_last_error(void) { return g_err; }

int cev_create(int device, cev_handle** out) {
    CEV_REQUIRE(out != nullptr, "cev_create: null out");
    int count = 0;
    CEV_CUDA(cudaGetDeviceCount(&count));
    CEV_REQUIRE(device >= 0 && device < count, "cev_create: device %d out of range (%d devices)", device, count);
    DeviceGuard guard(device);          // the caller's current device is restored on return
    cudaDeviceProp prop;
    CEV_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("cev_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                  prop.minor);
        return CEV_ERR_UNSUPPORTED;
    }
    cev_handle* h = new cev_handle();
    h->device = device;
    h->n_sm = prop.multiProcessorCount;
    h->workspace = nullptr;
    h->workspace_bytes = 0;
    h->opp_workspace = nullptr;
    h->opp_workspace_bytes = 0;
    h->ls_workspace = nullptr;
    h->ls_workspace_bytes = 0;
    h->side_stream = nullptr;
    h->fork_ev = h->join_ev = nullptr;
    h->env_stream = nullptr;
    h->timing_on = 0;
    h->timing_n[0] = h->timing_n[1] = 0;
    h->timing_ev[0] = h->timing_ev[1] = nullptr;
    int ncl = rollout_cluster_max_clusters(device);
    if (ncl <= 0) ncl = h->n_sm / 4 - 4;
    h->n_clusters = ncl;
    *out = h;
    return CEV_OK;
}

int cev_destroy(cev_handle* h) {
    if (!h) return CEV_OK;
    CEV_GUARD(h);
    if (h->workspace) cudaFree(h->workspace);
    if (h->opp_workspace) cudaFree(h->opp_workspace);
    if (h->ls_workspace) cudaFree(h->ls_workspace);
    if (h->side_stream) {
        cudaStreamDestroy(h->side_stream);
        cudaEventDestroy(h->fork_ev);
        cudaEventDestroy(h->join_ev);
    }
    if (h->env_stream) {
        cudaStreamDestroy(h->env_stream);
        for (int i = 0; i < 3; ++i) cudaStreamDestroy(h->opp_stream2[i]);
        for (int i = 0; i < 2; ++i) cudaStreamDestroy(h->mem_stream2[i]);
        for (int r = 0; r < CEV_MAX_ROLES; ++r) {
            cudaEventDestroy(h->ev_opp[r]);
            cudaEventDestroy(h->ev_mem[r]);
            cudaEventDestroy(h->ev_env[r]);
        }
    }
    for (int k = 0; k < 2; ++k)
        if (h->timing_ev[k]) {
            for (int i = 0; i < 2 * CEV_TIMING_MAX; ++i) cudaEventDestroy(h->timing_ev[k][i]);
            delete[] h->timing_ev[k];
        }
    delete h;
    return CEV_OK;
}

int cev_device_info(cev_handle* h, int* n_sm, int* n_clusters) {
    CEV_REQUIRE(h != nullptr, "cev_device_info: null handle");
    if (n_sm) *n_sm = h->n_sm;
    if (n_clusters) *n_clusters = h->n_clusters;
    return CEV_OK;
}

int cev_fc_dim(int in_dim) { return fc_offsets(in_dim).total; }
int cev_fc_pitch(int in_dim) { return fc_pitch(in_dim); }

int cev_dqn_dim(int c_in, int n_actions) {
    return 32 * c_in * 64 + 32 + 64 * 32 * 16 + 64 + 64 * 64 * 9 + 64 + 512 * 3136 + 512 + n_actions * 512 +
           n_actions + 2 * (32 + 64 + 64);
}
int cev_dqn_pitch(int c_in, int n_actions) { return round_up(cev_dqn_dim(c_in, n_actions), 32); }

// Which K1 kernel a structured rollout uses: 1 generic, 2 cluster (member weights resident in shared
// memory; small launches: evaluation games, tiny populations), 3 lockstep (opponent forwards on tcgen05,
// member rows streamed once per world step).  Lockstep wins whenever there are enough episodes to fill
// the opponent GEMM: 3.5x at the ES shape (16 episodes per member) and still 3.3x at the GA shape (one
// episode per member), where its member kernel is purely HBM bound (5.8 TB/s).
static int rollout_plan(const cev_handle* h, int P, int K, int E, int variant) {
    if (variant != 0) return variant;
    if ((int64_t)P * K * E >= 2048) return 3;
    return h->n_clusters > 0 ? 2 : 1;
}

static int check_cfg(const cev_rollout_cfg* cfg) {
    CEV_REQUIRE(cfg != nullptr, "rollout: null cfg");
    CEV_REQUIRE(cfg->n_cycles >= 0 && cfg->n_cycles <= MAX_CYCLES, "rollout: n_cycles must be in [0, %d]", MAX_CYCLES);
    CEV_REQUIRE(cfg->variant >= 0 && cfg->variant <= 3, "rollout: variant must be 0..3");
    return CEV_OK;
}

int cev_mpe_rollout_f32(cev_handle* h, int member_seat, const float* members, int P, int64_t member_pitch,
                        const float* opp_a, int64_t opp_a_pitch, const float* opp_b, int64_t opp_b_pitch, int K,
                        const double* init, int init_shared, int E, const cev_rollout_cfg* cfg, double* out,
                        int32_t* status, cev_stream stream) {
    CEV_REQUIRE(h && members && opp_a && opp_b && init && out, "mpe_rollout: null pointer");
    CEV_REQUIRE(member_seat >= 0 && member_seat <= 2, "mpe_rollout: member_seat must be 0..2");
    CEV_REQUIRE(P >= 0 && K >= 1 && E >= 1, "mpe_rollout: need P >= 0, K >= 1, E >= 1");
    int rc = check_cfg(cfg);
    if (rc) return rc;
    const int other[2] = {member_seat == 0 ? 1 : 0, member_seat == 2 ? 1 : 2};
    CEV_REQUIRE(member_pitch >= fc_offsets(seat_in_dim(member_seat)).total && member_pitch % 4 == 0,
                "mpe_rollout: member pitch too small / not a multiple of 4");
    CEV_REQUIRE(opp_a_pitch >= fc_offsets(seat_in_dim(other[0])).total && opp_a_pitch % 4 == 0 &&
                    opp_b_pitch >= fc_offsets(seat_in_dim(other[1])).total && opp_b_pitch % 4 == 0,
                "mpe_rollout: opponent pitch too small / not a multiple of 4");
    CEV_REQUIRE(aligned16(members) && aligned16(opp_a) && aligned16(opp_b), "mpe_rollout: rows must be 16B aligned");
    if (P == 0) return CEV_OK;
    CEV_GUARD(h);
    cudaStream_t st = (cudaStream_t)stream;
    const int variant = rollout_plan(h, P, K, E, cfg->variant);
    if (variant >= 2) {
        ClusterParams p{};
        p.members = members;
        p.member_pitch = member_pitch;
        p.P = P;
        p.opp[0] = opp_a;
        p.opp[1] = opp_b;
        p.opp_pitch[0] = opp_a_pitch;
        p.opp_pitch[1] = opp_b_pitch;
        p.K = K;
        p.member_seat = member_seat;
        p.init = init;
        p.init_shared = init_shared;
        p.E = E;
        p.out = out;
        p.status = status;
        p.n_cycles = cfg->n_cycles;
        p.pos_first = cfg->integrate_pos_first;
        return variant == 3 ? launch_rollout_lockstep(h, p, st) : launch_rollout_cluster(h, p, st);
    }
    GenericParams g{};
    g.w[member_seat] = members;
    g.pitch[member_seat] = member_pitch;
    g.w[other[0]] = opp_a;
    g.pitch[other[0]] = opp_a_pitch;
    g.w[other[1]] = opp_b;
    g.pitch[other[1]] = opp_b_pitch;
    g.idx = nullptr;
    g.member_seat = member_seat;
    g.K = K;
    g.E = E;
    g.init_shared = init_shared;
    g.init = init;
    g.out = out;
    g.status = status;
    g.n_cycles = cfg->n_cycles;
    g.pos_first = cfg->integrate_pos_first;
    g.N = (int64_t)P * K * E;
    return launch_rollout_generic(h, g, st);
}

int cev_mpe_rollout_roles_f32(cev_handle* h, int n_roles, const cev_rollout_role* roles, int P, int K, int init_shared,
                              int E, const cev_rollout_cfg* cfg, int32_t* status, cev_stream stream) {
    CEV_REQUIRE(h && roles, "mpe_rollout_roles: null pointer");
    CEV_REQUIRE(n_roles >= 1 && n_roles <= CEV_MAX_ROLES, "mpe_rollout_roles: n_roles must be 1..%d", CEV_MAX_ROLES);
    CEV_REQUIRE(P >= 0 && K >= 1 && E >= 1, "mpe_rollout_roles: need P >= 0, K >= 1, E >= 1");
    int rc = check_cfg(cfg);
    if (rc) return rc;
    for (int r = 0; r < n_roles; ++r)
        CEV_REQUIRE(roles[r].members && roles[r].opp_a && roles[r].opp_b && roles[r].init && roles[r].out,
                    "mpe_rollout_roles: null pointer in role %d", r);
    if (P == 0) return CEV_OK;
    const int variant = rollout_plan(h, P, K, E, cfg->variant);
    if (variant != 3) {
        for (int r = 0; r < n_roles; ++r) {
            rc = cev_mpe_rollout_f32(h, roles[r].member_seat, roles[r].members, P, roles[r].member_pitch, roles[r].opp_a,
                                     roles[r].opp_a_pitch, roles[r].opp_b, roles[r].opp_b_pitch, K, roles[r].init,
                                     init_shared, E, cfg, roles[r].out, status, stream);
            if (rc) return rc;
        }
        return CEV_OK;
    }
    ClusterParams ps[CEV_MAX_ROLES];
    for (int r = 0; r < n_roles; ++r) {
        const cev_rollout_role& q = roles[r];
        CEV_REQUIRE(q.member_seat >= 0 && q.member_seat <= 2, "mpe_rollout_roles: member_seat must be 0..2");
        const int other[2] = {q.member_seat == 0 ? 1 : 0, q.member_seat == 2 ? 1 : 2};
        CEV_REQUIRE(q.member_pitch >= fc_offsets(seat_in_dim(q.member_seat)).total && q.member_pitch % 4 == 0 &&
                        q.opp_a_pitch >= fc_offsets(seat_in_dim(other[0])).total && q.opp_a_pitch % 4 == 0 &&
                        q.opp_b_pitch >= fc_offsets(seat_in_dim(other[1])).total && q.opp_b_pitch % 4 == 0,
                    "mpe_rollout_roles: pitch too small / not a multiple of 4");
        CEV_REQUIRE(aligned16(q.members) && aligned16(q.opp_a) && aligned16(q.opp_b),
                    "mpe_rollout_roles: rows must be 16B aligned");
        ClusterParams& p = ps[r];
        p = ClusterParams{};
        p.members = q.members;
        p.member_pitch = q.member_pitch;
        p.P = P;
        p.opp[0] = q.opp_a;
        p.opp[1] = q.opp_b;
        p.opp_pitch[0] = q.opp_a_pitch;
        p.opp_pitch[1] = q.opp_b_pitch;
        p.K = K;
        p.member_seat = q.member_seat;
        p.init = q.init;
        p.init_shared = init_shared;
        p.E = E;
        p.out = q.out;
        p.status = status;
        p.n_cycles = cfg->n_cycles;
        p.pos_first = cfg->integrate_pos_first;
    }
    CEV_GUARD(h);
    return launch_rollout_lockstep_roles(h, ps, n_roles, (cudaStream_t)stream);
}

int cev_mpe_rollout_trace_f32(cev_handle* h, int member_seat, const float* members, int P, int64_t member_pitch,
                              const float* opp_a, int64_t opp_a_pitch, const float* opp_b, int64_t opp_b_pitch, int K,
                              const double* init, int init_shared, int E, const cev_rollout_cfg* cfg,
                              const int32_t* forced_actions, float* logits_out, int32_t* actions_out, double* out,
                              int32_t* status, cev_stream stream) {
    CEV_REQUIRE(h && members && opp_a && opp_b && init && out, "mpe_rollout_trace: null pointer");
    CEV_REQUIRE(member_seat >= 0 && member_seat <= 2, "mpe_rollout_trace: member_seat must be 0..2");
    CEV_REQUIRE(P >= 1 && K >= 1 && E >= 1, "mpe_rollout_trace: need P, K, E >= 1");
    int rc = check_cfg(cfg);
    if (rc) return rc;
    const int other[2] = {member_seat == 0 ? 1 : 0, member_seat == 2 ? 1 : 2};
    CEV_REQUIRE(member_pitch >= fc_offsets(seat_in_dim(member_seat)).total && member_pitch % 4 == 0 &&
                    opp_a_pitch >= fc_offsets(seat_in_dim(other[0])).total && opp_a_pitch % 4 == 0 &&
                    opp_b_pitch >= fc_offsets(seat_in_dim(other[1])).total && opp_b_pitch % 4 == 0,
                "mpe_rollout_trace: pitch too small / not a multiple of 4");
    CEV_REQUIRE(aligned16(members) && aligned16(opp_a) && aligned16(opp_b), "mpe_rollout_trace: rows must be 16B aligned");
    CEV_GUARD(h);
    ClusterParams p{};
    p.members = members;
    p.member_pitch = member_pitch;
    p.P = P;
    p.opp[0] = opp_a;
    p.opp[1] = opp_b;
    p.opp_pitch[0] = opp_a_pitch;
    p.opp_pitch[1] = opp_b_pitch;
    p.K = K;
    p.member_seat = member_seat;
    p.init = init;
    p.init_shared = init_shared;
    p.E = E;
    p.out = out;
    p.status = status;
    p.n_cycles = cfg->n_cycles;
    p.pos_first = cfg->integrate_pos_first;
    p.trace_forced = forced_actions;
    p.trace_logits = logits_out;
    p.trace_actions = actions_out;
    return launch_rollout_lockstep(h, p, (cudaStream_t)stream);
}

int cev_mpe_rollout_plan(cev_handle* h, int P, int K, int E, int n_cycles, int variant, int* variant_used,
                         int* n_launches) {
    CEV_REQUIRE(h != nullptr, "mpe_rollout_plan: null handle");
    CEV_REQUIRE(variant >= 0 && variant <= 3, "mpe_rollout_plan: variant must be 0..3");
    const int v = rollout_plan(h, P, K, E, variant);
    if (variant_used) *variant_used = v;
    if (n_launches) *n_launches = v == 3 ? rollout_lockstep_launches(n_cycles) : (v == 2 ? 3 : 1);
    return CEV_OK;
}

int cev_kernel_timing_enable(cev_handle* h, int on) {
    CEV_REQUIRE(h != nullptr, "kernel_timing_enable: null handle");
    CEV_GUARD(h);
    if (on && !h->timing_ev[0]) {
        for (int k = 0; k < 2; ++k) {
            h->timing_ev[k] = new cudaEvent_t[2 * CEV_TIMING_MAX];
            for (int i = 0; i < 2 * CEV_TIMING_MAX; ++i) CEV_CUDA(cudaEventCreate(&h->timing_ev[k][i]));
        }
    }
    h->timing_on = on ? 1 : 0;
    h->timing_n[0] = h->timing_n[1] = 0;
    return CEV_OK;
}

int cev_kernel_timing_read(cev_handle* h, int which, double* total_ms, int* n_launches) {
    CEV_REQUIRE(h != nullptr && (which == 0 || which == 1), "kernel_timing_read: bad arguments");
    CEV_GUARD(h);
    double tot = 0.0;
    const int n = h->timing_n[which];
    for (int i = 0; i < n; ++i) {
        CEV_CUDA(cudaEventSynchronize(h->timing_ev[which][2 * i + 1]));
        float ms = 0.f;
        CEV_CUDA(cudaEventElapsedTime(&ms, h->timing_ev[which][2 * i], h->timing_ev[which][2 * i + 1]));
        tot += ms;
    }
    if (total_ms) *total_ms = tot;
    if (n_launches) *n_launches = n;
    h->timing_n[which] = 0;
    return CEV_OK;
}

int cev_mpe_rollout_indexed_f32(cev_handle* h, const float* w_adv, int64_t adv_pitch, const float* w_a0,
                                int64_t a0_pitch, const float* w_a1, int64_t a1_pitch, const int32_t* idx,
                                const double* init, int N, const cev_rollout_cfg* cfg, double* out,
                                int32_t* status, cev_stream stream) {
    CEV_REQUIRE(h && w_adv && w_a0 && w_a1 && idx && init && out, "mpe_rollout_indexed: null pointer");
    CEV_REQUIRE(N >= 0, "mpe_rollout_indexed: N < 0");
    int rc = check_cfg(cfg);
    if (rc) return rc;
    CEV_REQUIRE(adv_pitch >= fc_offsets(IN_ADV).total && a0_pitch >= fc_offsets(IN_GOOD).total &&
                    a1_pitch >= fc_offsets(IN_GOOD).total && adv_pitch % 4 == 0 && a0_pitch % 4 == 0 &&
                    a1_pitch % 4 == 0,
                "mpe_rollout_indexed: pitch too small / not a multiple of 4");
    CEV_REQUIRE(aligned16(w_adv) && aligned16(w_a0) && aligned16(w_a1), "mpe_rollout_indexed: rows must be 16B aligned");
    CEV_GUARD(h);
    GenericParams g{};
    g.w[0] = w_adv;
    g.w[1] = w_a0;
    g.w[2] = w_a1;
    g.pitch[0] = adv_pitch;
    g.pitch[1] = a0_pitch;
    g.pitch[2] = a1_pitch;
    g.idx = idx;
    g.member_seat = 0;
    g.K = 1;
    g.E = 1;
    g.init_shared = 0;
    g.init = init;
    g.out = out;
    g.status = status;
    g.n_cycles = cfg->n_cycles;
    g.pos_first = cfg->integrate_pos_first;
    g.N = N;
    return launch_rollout_generic(h, g, (cudaStream_t)stream);
}

int cev_fc_forward_f32(cev_handle* h, const float* rows, int64_t pitch, int in_dim, const int32_t* idx,
                       const float* obs, int64_t N, float* logits, int32_t* actions, int32_t* status,
                       cev_stream stream) {
    CEV_REQUIRE(h && rows && obs && logits, "fc_forward: null pointer");
    CEV_REQUIRE(in_dim == IN_ADV || in_dim == IN_GOOD, "fc_forward: in_dim must be 8 or 10");
    CEV_REQUIRE(pitch >= fc_offsets(in_dim).total && pitch % 4 == 0 && aligned16(rows),
                "fc_forward: bad pitch / alignment");
    CEV_GUARD(h);
    return launch_fc_forward(h, rows, pitch, in_dim, idx, obs, N, logits, actions, status, (cudaStream_t)stream);
}

}  // extern "C"
