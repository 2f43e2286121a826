// tcgen05.mma kind::tf32 cost probe: cycles per instruction for small N, A from shared memory (SS) or tensor memory (TS),
// one accumulator or several (development aid, not part of the library).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../coevonet_b200/csrc -o mma_probe mma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
    return (uint64_t)((a >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(su32(bar)) : "memory");
}
__device__ __forceinline__ void mwait(uint64_t* bar, uint32_t par) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(su32(bar)), "r"(par) : "memory");
}

// MODE: 0 = SS, 1 = TS; N = MMA N; NACC = accumulators used round robin; 64 unrolled MMAs x outer reps
template <int MODE, int N, int NACC>
__global__ void __launch_bounds__(128, 1) probe(int outer, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(su32(&tslot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tb = tslot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t a = desc_sw128(su32(smem)), b = desc_sw128(su32(smem) + 16384);
        const long long t0 = clock64();
        for (int o = 0; o < outer; ++o) {
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                const uint32_t d = tb + (uint32_t)((i % NACC) * N);
                const uint64_t ko = (uint64_t)((i & 3) * 2);
                if (MODE == 0) mma_ss(d, a + ko, b + ko, idesc, 1u);
                else mma_ts(d, tb + 384 + (i & 3) * 8, b + ko, idesc, 1u);
            }
        }
        const long long t1 = clock64();
        commit(&bar);
        mwait(&bar, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tb), "r"(512u) : "memory");
}

template <int MODE, int N, int NACC>
void run(long long* out) {
    CK(cudaFuncSetAttribute(probe<MODE, N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    const int outer = 8;
    for (int grid : {1, 148}) {
        for (int rep = 0; rep < 2; ++rep) { probe<MODE, N, NACC><<<grid, 128, 64 * 1024>>>(outer, out); CK(cudaDeviceSynchronize()); }
        long long h[2]; CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost));
        printf("grid %3d %s N=%3d acc=%d: issue %6.1f clk/MMA, complete %6.1f clk/MMA\n", grid, MODE ? "TS" : "SS", N, NACC,
               (double)h[0] / (64 * outer), (double)h[1] / (64 * outer)); fflush(stdout);
    }
}

int main() {
    long long* out; CK(cudaMalloc(&out, 16));
    run<0, 16, 1>(out); run<0, 32, 1>(out); run<0, 32, 4>(out); run<0, 64, 1>(out); run<0, 128, 1>(out); run<0, 256, 1>(out);
    run<1, 16, 1>(out); run<1, 16, 4>(out); run<1, 32, 1>(out); run<1, 64, 1>(out); run<1, 256, 1>(out);
    return 0;
}
