// tcgen05 / TMEM / TMA plumbing shared by the tensor-core kernels (deepqn_tc.cu, rollout_lockstep.cu).
// sm_100a only: inline PTX for mbarriers, cp.async.bulk.tensor loads, K-major 128B-swizzled UMMA
// shared-memory descriptors, tcgen05.commit / tcgen05.ld, and the driver's tensor-map encoder.
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace cev {

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(tc_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(tc_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool tc_mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(tc_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// wait for completion of the phase with the given parity; traps instead of hanging the GPU
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
    if (tc_mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!tc_mbar_try(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::
            "r"(tc_smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(tc_smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::
            "r"(tc_smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(tc_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// K-major, 128B-swizzled shared-memory matrix descriptor (rows 128 bytes apart, 8-row groups
// 1024 bytes apart): start >> 4 | LBO (unused for swizzled K-major) | SBO = 1024 >> 4 |
// version 1 (Blackwell) | layout 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                     tc_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// tcgen05.mma kind::tf32 with the instruction descriptor as an argument (A from shared memory / from tensor
// memory), tcgen05.st / tcgen05.ld of 32 / 16 columns, and the "lo" part of the 3xTF32 split: what is left
// of an fp32 word once the tensor core has read it truncated to TF32.
__device__ __forceinline__ void f3_umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand from tensor memory (lane = row of the 128-row tile, one 32-bit column per tf32 element)
__device__ __forceinline__ void f3_umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void f3_tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
        "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
        "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]), "f"(v[18]),
        "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]),
        "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31])
        : "memory");
}
__device__ __forceinline__ void f3_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ float f3_lo(float w) { return w - __uint_as_float(__float_as_uint(w) & 0xffffe000u); }

// The twelve MMAs of one k-tile (4 k-steps x {a_lo . b_hi, a_hi . b_lo, a_hi . b_hi}, small terms first) and the commit
// that releases the stage, issued by ONE elected lane of a CONVERGED warp: inside a divergent `if (lane == 0)` every
// tcgen05.mma costs an ELECT + R2UR.BROADCAST chain (~20 instructions, ~130 cycles of issue for 128 cycles of tensor
// pipe), which held the tensor pipe at ~50 %.
__device__ __forceinline__ void tc_issue_ktile_3xtf32(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                                      uint32_t idesc, uint32_t accumulate, uint32_t bar_empty) {
    asm volatile(
        "{\n"
        ".reg .pred pe, pa, pt;\n"
        ".reg .b64 ah1, ah2, ah3, al1, al2, al3, bh1, bh2, bh3, bl1, bl2, bl3;\n"
        "elect.sync _|pe, 0xffffffff;\n"
        "setp.ne.b32 pa, %5, 0;\n"
        "setp.eq.b32 pt, 0, 0;\n"
        "add.s64 ah1, %1, 2;\n add.s64 ah2, %1, 4;\n add.s64 ah3, %1, 6;\n"
        "add.s64 al1, %2, 2;\n add.s64 al2, %2, 4;\n add.s64 al3, %2, 6;\n"
        "add.s64 bh1, %3, 2;\n add.s64 bh2, %3, 4;\n add.s64 bh3, %3, 6;\n"
        "add.s64 bl1, %4, 2;\n add.s64 bl2, %4, 4;\n add.s64 bl3, %4, 6;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %2, %3, %7, pa;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %4, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %3, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], al1, bh1, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], ah1, bl1, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], ah1, bh1, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], al2, bh2, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], ah2, bl2, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], ah2, bh2, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], al3, bh3, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], ah3, bl3, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], ah3, bh3, %7, pt;\n"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_hi), "l"(a_lo), "l"(b_hi), "l"(b_lo), "r"(accumulate), "r"(bar_empty), "r"(idesc)
        : "memory");
}
// the same twelve MMAs with kind::f16 operands (K = 16 per instruction: a 128-byte swizzle row holds 64 k)
__device__ __forceinline__ void tc_issue_ktile_3xf16(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                                      uint32_t idesc, uint32_t accumulate, uint32_t bar_empty) {
    asm volatile(
        "{\n"
        ".reg .pred pe, pa, pt;\n"
        ".reg .b64 ah1, ah2, ah3, al1, al2, al3, bh1, bh2, bh3, bl1, bl2, bl3;\n"
        "elect.sync _|pe, 0xffffffff;\n"
        "setp.ne.b32 pa, %5, 0;\n"
        "setp.eq.b32 pt, 0, 0;\n"
        "add.s64 ah1, %1, 2;\n add.s64 ah2, %1, 4;\n add.s64 ah3, %1, 6;\n"
        "add.s64 al1, %2, 2;\n add.s64 al2, %2, 4;\n add.s64 al3, %2, 6;\n"
        "add.s64 bh1, %3, 2;\n add.s64 bh2, %3, 4;\n add.s64 bh3, %3, 6;\n"
        "add.s64 bl1, %4, 2;\n add.s64 bl2, %4, 4;\n add.s64 bl3, %4, 6;\n"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %2, %3, %7, pa;\n"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %4, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %3, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], al1, bh1, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], ah1, bl1, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], ah1, bh1, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], al2, bh2, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], ah2, bl2, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], ah2, bh2, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], al3, bh3, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], ah3, bl3, %7, pt;\n"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], ah3, bh3, %7, pt;\n"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_hi), "l"(a_lo), "l"(b_hi), "l"(b_lo), "r"(accumulate), "r"(bar_empty), "r"(idesc)
        : "memory");
}
__device__ __forceinline__ void tc_commit_elected(uint32_t bar) {
    asm volatile(
        "{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\n"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(bar)
        : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

}  // namespace cev
