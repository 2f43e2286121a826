// K2, fully-connected stage on the 5th-generation tensor cores.
//
// fc1 of DeepQN (Atari/deepqn.py:27,46: 3136 -> 512) holds 95 % of a member's bytes and
// is a dense contraction: per member  H[frames, 512] = X[frames, 3136] . W1[512, 3136]^T.
// One persistent CTA per member streams W1 ONCE from HBM with TMA (128B-swizzled
// [256 x 32] fp32 boxes), feeds tcgen05.mma kind::tf32 (M = 128 frame rows, zero padded by
// TMA out-of-bounds fill, N = 256, K = 8 per instruction) with accumulators in TMEM, and the
// epilogue warps read the accumulators back with tcgen05.ld, apply bias + ReLU and fold
// the 512 -> A output layer and the first-max argmax in registers (no hidden-layer
// round trip).  Warp roles: 0 = TMA producer, 1 = MMA issuer (+ TMEM alloc), 2..5 = epilogue.
//
// Arithmetic: TF32 inputs (10-bit mantissa), fp32 accumulation; logits agree with the fp32
// reference to ~1e-3 (tests state the tolerance).  The fp32 CUDA-core stage in deepqn.cu
// stays selectable (COEVONET_DQN_FC=fp32) and is the 2e-5 parity path.
#include "tc_common.cuh"

namespace cev {

constexpr int TC_THREADS = 192;
constexpr int TC_BK = 32;                    // fp32 per k-tile = 128 bytes = one swizzle row
constexpr int TC_K = 3136;
constexpr int TC_KT = TC_K / TC_BK;          // 98 k-tiles
constexpr int TC_BM = 128;                   // frame rows per pass (TMEM lanes)
constexpr int TC_BN = 256;                   // fc1 outputs per pass (TMEM columns)
constexpr int TC_NH = 512 / TC_BN;           // 2 passes over N
constexpr int TC_STAGES = 4;
constexpr uint32_t TC_A_BYTES = TC_BM * TC_BK * 4;     // 16 KB
constexpr uint32_t TC_B_BYTES = TC_BN * TC_BK * 4;     // 32 KB
constexpr uint32_t TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
constexpr size_t TC_SMEM = (size_t)TC_STAGES * TC_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TC_MAX_ACT = 32;

// instruction descriptor: D = F32, A = B = TF32, both K-major, N = 256, M = 128
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) |
                              ((uint32_t)(TC_BM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(TC_IDESC), "r"(accumulate)
        : "memory");
}
struct TcParams {
    const float* members;
    int64_t pitch;
    int P, B, n_act;
    int f1b_off, ow_off, ob_off;
    float* logits;
    int32_t* actions;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
deepqn_fc_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                    const TcParams p) {
    extern __shared__ unsigned char tc_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* stage_mem = base;                                            // STAGES x (A | B), 1024B aligned
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(base + (size_t)TC_STAGES * TC_STAGE_BYTES);
    uint64_t* bar_empty = bar_full + TC_STAGES;
    uint64_t* bar_tfull = bar_empty + TC_STAGES;      // [2] accumulator ready
    uint64_t* bar_tempty = bar_tfull + 2;             // [2] accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_fblk = (p.B + TC_BM - 1) / TC_BM;

    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_STAGES; ++i) {
            tc_mbar_init(bar_full + i, 1);
            tc_mbar_init(bar_empty + i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc_mbar_init(bar_tfull + i, 1);
            tc_mbar_init(bar_tempty + i, 4);          // one arrive per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) {
        // 512 TMEM columns: two 256-column fp32 accumulators (ping-pong)
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
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int m = blockIdx.x; m < p.P; m += gridDim.x)
                for (int fb = 0; fb < n_fblk; ++fb)
                    for (int nh = 0; nh < TC_NH; ++nh)
                        for (int kt = 0; kt < TC_KT; ++kt, ++it) {
                            const uint32_t st = it % TC_STAGES, use = it / TC_STAGES;
                            if (use > 0) tc_mbar_wait(bar_empty + st, (use - 1) & 1);
                            unsigned char* a = stage_mem + (size_t)st * TC_STAGE_BYTES;
                            tc_mbar_expect_tx(bar_full + st, TC_STAGE_BYTES);
                            tma_load_2d(a, &map_x, bar_full + st, kt * TC_BK, m * p.B + fb * TC_BM);
                            tma_load_3d(a + TC_A_BYTES, &map_w, bar_full + st, kt * TC_BK, nh * TC_BN, m);
                        }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        uint32_t it = 0, pass = 0;
        for (int m = blockIdx.x; m < p.P; m += gridDim.x)
            for (int fb = 0; fb < n_fblk; ++fb)
                for (int nh = 0; nh < TC_NH; ++nh, ++pass) {
                    const uint32_t as = pass & 1, ause = pass >> 1;
                    if (ause > 0) tc_mbar_wait(bar_tempty + as, (ause - 1) & 1);      // epilogue drained it
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const uint32_t d_tmem = tmem_base + as * TC_BN;
                    for (int kt = 0; kt < TC_KT; ++kt, ++it) {
                        const uint32_t st = it % TC_STAGES, use = it / TC_STAGES;
                        tc_mbar_wait(bar_full + st, use & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                        if (lane == 0) {
                            const uint32_t a_addr = tc_smem_u32(stage_mem + (size_t)st * TC_STAGE_BYTES);
                            const uint64_t a_desc = umma_desc_sw128(a_addr);
                            const uint64_t b_desc = umma_desc_sw128(a_addr + TC_A_BYTES);
#pragma unroll
                            for (int k = 0; k < TC_BK / 8; ++k)       // 8 tf32 = 32 bytes per instruction
                                umma_tf32(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2),
                                          (kt | k) ? 1u : 0u);
                            umma_commit(bar_empty + st);             // frees the stage when the MMAs retire
                            if (kt == TC_KT - 1) umma_commit(bar_tfull + as);
                        }
                        __syncwarp();
                    }
                }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;                   // frame row inside the block
        uint32_t pass = 0;
        for (int m = blockIdx.x; m < p.P; m += gridDim.x) {
            const float* W = p.members + (int64_t)m * p.pitch;
            for (int fb = 0; fb < n_fblk; ++fb) {
                const int frame = fb * TC_BM + row;
                float lg[TC_MAX_ACT];
#pragma unroll
                for (int a = 0; a < TC_MAX_ACT; ++a) lg[a] = 0.f;
                for (int nh = 0; nh < TC_NH; ++nh, ++pass) {
                    const uint32_t as = pass & 1, ause = pass >> 1;
                    tc_mbar_wait(bar_tfull + as, ause & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * TC_BN;
                    for (int c0 = 0; c0 < TC_BN; c0 += 32) {
                        uint32_t v[32];
                        tmem_ld32(taddr + c0, v);
                        if (frame < p.B) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const int n = nh * TC_BN + c0 + j;
                                const float h = fmaxf(__uint_as_float(v[j]) + __ldg(W + p.f1b_off + n), 0.f);
#pragma unroll
                                for (int a = 0; a < TC_MAX_ACT; ++a)
                                    if (a < p.n_act) lg[a] = fmaf(__ldg(W + p.ow_off + a * 512 + n), h, lg[a]);
                            }
                        }
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                    __syncwarp();
                    if (lane == 0) tc_mbar_arrive(bar_tempty + as);
                }
                if (frame < p.B) {
                    float* out = p.logits + ((int64_t)m * p.B + frame) * p.n_act;
                    int best = 0;
                    float bv = -CUDART_INF_F;
#pragma unroll
                    for (int a = 0; a < TC_MAX_ACT; ++a)
                        if (a < p.n_act) {
                            const float v = lg[a] + __ldg(W + p.ob_off + a);
                            out[a] = v;
                            if (v > bv) { bv = v; best = a; }          // first maximum
                        }
                    if (p.actions) p.actions[(int64_t)m * p.B + frame] = best;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}


// fc1 + output layer + argmax for all members on the tensor cores.
// act3: fp32 [P*B, 3136] (conv3 output, post BN + ReLU, flattened like torch's reshape).
int launch_deepqn_fc_tc(cev_handle* h, const float* members, int64_t pitch, int P, int B, int n_act, int f1w_off,
                        int f1b_off, int ow_off, int ob_off, const float* act3, float* logits, int32_t* actions,
                        cudaStream_t stream) {
    EncodeTiledFn encode = get_encode_fn();
    if (!encode) {
        set_error("deepqn_fc_tc: cuTensorMapEncodeTiled is not available from the driver");
        return CEV_ERR_UNSUPPORTED;
    }
    CEV_REQUIRE(n_act <= TC_MAX_ACT, "deepqn_fc_tc: at most %d actions", TC_MAX_ACT);
    CUtensorMap map_x, map_w;
    {
        cuuint64_t dims[2] = {(cuuint64_t)TC_K, (cuuint64_t)P * B};
        cuuint64_t strides[1] = {(cuuint64_t)TC_K * 4};
        cuuint32_t box[2] = {TC_BK, TC_BM};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(act3), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("deepqn_fc_tc: cuTensorMapEncodeTiled(x) failed with %d", (int)r);
            return CEV_ERR_CUDA;
        }
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)TC_K, 512, (cuuint64_t)P};
        cuuint64_t strides[2] = {(cuuint64_t)TC_K * 4, (cuuint64_t)pitch * 4};
        cuuint32_t box[3] = {TC_BK, TC_BN, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&map_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(members + f1w_off), dims,
                            strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("deepqn_fc_tc: cuTensorMapEncodeTiled(w) failed with %d", (int)r);
            return CEV_ERR_CUDA;
        }
    }
    TcParams p;
    p.members = members;
    p.pitch = pitch;
    p.P = P;
    p.B = B;
    p.n_act = n_act;
    p.f1b_off = f1b_off;
    p.ow_off = ow_off;
    p.ob_off = ob_off;
    p.logits = logits;
    p.actions = actions;
    static bool configured = false;
    if (!configured) {
        CEV_CUDA(cudaFuncSetAttribute(deepqn_fc_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
        configured = true;
    }
    const int grid = P < h->n_sm ? P : h->n_sm;
    deepqn_fc_tc_kernel<<<grid, TC_THREADS, TC_SMEM, stream>>>(map_x, map_w, p);
    return check_cuda(cudaGetLastError(), "deepqn_fc_tc_kernel");
}

}  // namespace cev
