// HBM streaming probe: persistent CTAs pull a [P x pitch] fp32 array through shared memory with different copy shapes
// (development aid, not part of the library).  nvcc -arch=sm_100a -O3 -o bw_probe bw_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(su32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void tma3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(su32(dst)), "l"((uint64_t)m), "r"(su32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(su32(dst)), "l"(src), "r"(bytes), "r"(su32(bar)) : "memory");
}

// mode 0: 3-D tensor tiles (box given by the map), mode 1: contiguous bulk copies of `chunk` bytes
__global__ void __launch_bounds__(128, 2) probe(const __grid_constant__ CUtensorMap map, const float* base, int64_t pitch, int P, int mode,
                                                int box_k, int box_rows, int chunk, int slots, int off_floats, float* sink) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw;
    unsigned char* smem_all = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_all + (size_t)slots * chunk);
    if (threadIdx.x == 0) {
        for (int i = 0; i < slots; ++i) mbar_init(bars + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if ((threadIdx.x & 31) != 0) return;
    const int W = blockDim.x / 32, w = threadIdx.x / 32;
    const int tiles_k = 512 / box_k, tiles_r = 256 / box_rows;
    const int per_member_all = mode == 0 ? tiles_k * tiles_r : (512 * 256 * 4) / chunk;
    const int per_member = per_member_all / W;             // this warp's share
    slots /= W;
    smem += (size_t)w * slots * chunk;
    bars += w * slots;
    int64_t issued = 0, done = 0;
    float acc = 0.f;
    int64_t total = 0;
    for (int m = blockIdx.x; m < P; m += gridDim.x) total += per_member;
    int m_i = blockIdx.x, t_i = 0;
    while (done < total) {
        while (issued < total && issued - done < slots) {
            const int s = (int)(issued % slots);
            mbar_expect(bars + s, chunk);
            if (mode == 0) {
                const int tt = t_i * W + w;
                const int kt = tt / tiles_r, rt = tt % tiles_r;          // rows fastest, like the kernels
                tma3d(smem + (size_t)s * chunk, &map, bars + s, kt * box_k, rt * box_rows, m_i);
            } else {
                bulk1d(smem + (size_t)s * chunk, base + (int64_t)m_i * pitch + off_floats + (int64_t)(t_i * W + w) * (chunk / 4), chunk, bars + s);
            }
            ++issued;
            if (++t_i == per_member) { t_i = 0; m_i += gridDim.x; }
        }
        const int s = (int)(done % slots);
        mbar_wait(bars + s, (uint32_t)((done / slots) & 1));
        acc += reinterpret_cast<volatile float*>(smem + (size_t)s * chunk)[0];
        ++done;
    }
    if (acc == 123.f) sink[0] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int P = argc > 1 ? atoi(argv[1]) : 1024;
    const int NW = argc > 2 ? atoi(argv[2]) : 4;
    const int64_t pitch = 139808;
    const int off = 6656;
    float* d; float* sink;
    CK(cudaMalloc(&d, (size_t)P * pitch * 4)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(d, 0, (size_t)P * pitch * 4));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fn;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    struct Cfg { int mode, bk, br, chunk, slots, sw, l2; const char* name; };
    Cfg cfgs[] = {
        {0, 16, 128, 8192, 8, 2, 128, "tma [128 rows x 16 k] 64B rows, 8 slots x 2 CTAs (FP32 member form)"},
        {0, 32, 64, 8192, 8, 3, 128, "tma [64 rows x 32 k] 128B rows, 8 slots x 2 CTAs"},
        {0, 32, 64, 8192, 8, 3, 256, "tma [64 rows x 32 k] 128B rows, 8 slots x 2 CTAs, L2 256B"},
        {0, 16, 128, 8192, 16, 2, 128, "tma [128 rows x 16 k] 64B rows, 16 slots"},
        {0, 32, 256, 32768, 5, 3, 128, "tma [256 rows x 32 k] 128B rows, 5 slots"},
        {0, 32, 256, 32768, 5, 3, 256, "tma [256 rows x 32 k] 128B rows, 5 slots, L2 256B"},
        {0, 32, 128, 16384, 10, 3, 256, "tma [128 rows x 32 k] 128B rows, 10 slots, L2 256B"},
        {0, 32, 64, 8192, 20, 3, 256, "tma [64 rows x 32 k] 128B rows, 20 slots, L2 256B"},
        {1, 0, 0, 32768, 5, 0, 0, "bulk 32 KB contiguous, 5 slots"},
        {1, 0, 0, 8192, 20, 0, 0, "bulk 8 KB contiguous, 20 slots"},
        {1, 0, 0, 2048, 64, 0, 0, "bulk 2 KB contiguous (one fc2 row), 64 slots"},
    };
    for (auto& c : cfgs) {
        CUtensorMap map; memset(&map, 0, sizeof(map));
        if (c.mode == 0) {
            cuuint64_t dims[3] = {512, 256, (cuuint64_t)P};
            cuuint64_t strides[2] = {512 * 4, (cuuint64_t)pitch * 4};
            cuuint32_t box[3] = {(cuuint32_t)c.bk, (cuuint32_t)c.br, 1};
            cuuint32_t es[3] = {1, 1, 1};
            CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d + off, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             c.sw == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                             c.l2 == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
        }
        const size_t smem = (size_t)c.slots * c.chunk + 1024;
        CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int grid : {37, 74, 111, 148, 296}) {
            if (grid == 296 && smem > 110 * 1024) continue;
            float best = 1e9f;
            for (int rep = 0; rep < 4; ++rep) {
                CK(cudaEventRecord(e0));
                probe<<<grid, 32 * NW, smem>>>(map, d, pitch, P, c.mode, c.bk, c.br, c.chunk, c.slots, off, sink);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (rep > 0 && ms < best) best = ms;
            }
            CK(cudaGetLastError());
            const double bytes = (double)P * 512 * 256 * 4;
            printf("%-76s grid %3d: %8.1f us  %6.2f TB/s\n", c.name, grid, best * 1e3, bytes / (best * 1e-3) / 1e12);
        }
    }
    return 0;
}
