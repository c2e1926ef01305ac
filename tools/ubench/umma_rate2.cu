// Micro-benchmark 2: cycles per tcgen05.mma (M=128, K=16, bf16, K-major SWIZZLE_NONE) when the A operand CHANGES from one
// instruction to the next, as in the convolution kernels (every tap / accumulator block reads a different shared-memory
// window), against the same-A loop of umma_rate.cu.  One elected thread issues; operands are zeros.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../radar_depth_b200/csrc/rd_common.cuh"
using namespace rd;

__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// pattern 0: same A, same B        1: A advances by 128 rows per MMA over MB blocks (conv accumulator blocks), same B
// pattern 2: like 1, and a new tap (A shifted by one slot, B advanced by one tap) after every MB MMAs: the conv loop
template <int MB>
__global__ void __launch_bounds__(128, 1) k(int pattern, int N, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    fence_proxy_async_smem();
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x < 32) {
        if (elect_one_sync()) {
            const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
            const uint32_t plane = 36 * 1024;                // bytes between the two 8-channel chunk planes of A
            const uint64_t da0 = desc(a0, plane, 128), db0 = desc(b0, (uint32_t)N * 16, 128);
            const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
            const uint32_t tap_units = (uint32_t)N * 2;
            long long t0 = clock64();
            for (int i = 0; i < iters; i += 9 * MB) {
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const uint64_t da = pattern == 2 ? da0 + (uint32_t)((t / 3) * 64 + (t % 3)) : da0;
                    const uint64_t db = pattern == 2 ? db0 + (uint32_t)t * tap_units : db0;
#pragma unroll
                    for (int mb = 0; mb < MB; ++mb)
                        umma(tm + (uint32_t)(mb * N), pattern == 0 ? da : da + (uint32_t)mb * 128u, db, idesc, (i | t) ? 1u : 0u);
                }
            }
            umma_commit(&bar);
            mbar_wait(&bar, 0, 0x1);
            long long t1 = clock64();
            if (blockIdx.x == 0) *out = t1 - t0;
        }
        __syncwarp();
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

template <int MB>
void run(int N, long long* d) {
    if (MB * N > 512) return;
    cudaFuncSetAttribute(k<MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int pattern = 0; pattern < 3; ++pattern) {
        const int iters = 9 * MB * 200;
        k<MB><<<148, 128, 200 * 1024>>>(pattern, N, iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        const char* names[3] = {"same A, same B          ", "A per accumulator block ", "conv loop (taps x blocks)"};
        printf("N=%3d MB=%d %s : %7.1f cycles/MMA (math %5.1f, smem (A+B)/128 %5.1f) %s\n", N, MB, names[pattern], (double)c / iters, N / 2.0,
               (4096.0 + 32.0 * N) / 128.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
}

int main() {
    long long* d; cudaMalloc(&d, 8);
    for (int N : {16, 32, 64, 128, 256}) { run<1>(N, d); run<2>(N, d); run<4>(N, d); run<8>(N, d); }
    return 0;
}
