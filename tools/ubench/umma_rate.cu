// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16) for different shared-memory layouts and N.
// Operands are whatever is in shared memory (zeros); only the issue/execute rate matters.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../radar_depth_b200/csrc/rd_common.cuh"
using namespace rd;

__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}

// mode 0: K-major no swizzle (A: LBO=plane, SBO=128)   mode 1: K-major SW128 (SBO=1024)
// mode 2: MN-major no swizzle (LBO=128, SBO=plane)      mode 3: MN-major SW128 (LBO=plane, SBO=1024)
__global__ void __launch_bounds__(128, 1) k(int mode, int N, int iters, int nacc, int shift_units, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    fence_proxy_async_smem();
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x < 32) {
        const bool leader = elect_one_sync();
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
        const uint32_t plane = 4112;     // bytes between chunk planes (513 slots: odd)
        uint64_t da, db; uint32_t idesc;
        if (mode == 0) { da = desc(a0, plane, 128, 0); db = desc(b0, (uint32_t)N * 16, 128, 0); idesc = make_idesc_bf16(128, N, 0, 0); }
        else if (mode == 1) { da = desc(a0, 16, 1024, 2); db = desc(b0, 16, 1024, 2); idesc = make_idesc_bf16(128, N, 0, 0); }
        else if (mode == 2) { da = desc(a0, 128, plane, 0); db = desc(b0, 128, plane, 0); idesc = make_idesc_bf16(128, N, 1, 1); }
        else { da = desc(a0, 16384, 1024, 2); db = desc(b0, 16384, 1024, 2); idesc = make_idesc_bf16(128, N, 1, 1); }
        __syncwarp();
        long long t0 = clock64();
        // tight issue loop: nacc accumulators in rotation (like one k-group of the weight-gradient kernel: same A,
        // shifted B, a different accumulator per tap), operands precomputed and uniform
        for (int a = 0; a < nacc; ++a)
            if (leader) umma_bf16(tm + (uint32_t)(a * N), da, db, idesc, 0u);
        for (int i = 0; i < iters; i += nacc) {
            uint32_t d = tm;
            uint64_t dbi = db;
            for (int a = 0; a < nacc; ++a, d += (uint32_t)N, dbi += (uint32_t)shift_units) {
                if (leader) umma_bf16(d, da, dbi, idesc, 1u);
            }
        }
        if (leader) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0, 0x1);
        long long t1 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) *out = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

int main() {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const char* names[4] = {"K-major none ", "K-major SW128", "MN-major none", "MN-major SW128"};
    for (int grid : {148})
        for (int mode = 0; mode < 4; ++mode)
            for (int N : {32, 64, 128})
              if (!(mode >= 2 && N > 64))
                for (int nacc : {1, 2, 3, 5, 8}) {
                    if (nacc * N > 512) continue;
                    for (int shift : {0, 1}) {
                        const int iters = 4000;
                        k<<<grid, 128, 200 * 1024>>>(mode, N, iters, nacc, shift, d);
                        cudaError_t e = cudaDeviceSynchronize();
                        long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                        printf("grid %3d %s N=%3d nacc=%d shift=%d : %7.1f cycles/MMA (ideal %5.1f) %s\n", grid, names[mode], N, nacc, shift,
                               (double)c / iters, N / 2.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
                    }
                }
    return 0;
}
