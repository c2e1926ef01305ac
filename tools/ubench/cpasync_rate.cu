// Micro-benchmark: bytes/clk/SM of cp.async (LDGSTS) 16-byte copies from an L2-resident NHWC tensor into shared
// memory, for the access patterns of the tile loaders (nchunks x 16 B per pixel at a pixel pitch of `pitch` bytes),
// as a function of loader warps per CTA.  One CTA per SM, no consumers: only the copy rate matters.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>   // 0 = cp.async.cg 16 B, 1 = cp.async.cg with zero-fill (src size 0), 2 = ld.global.v4 + st.shared.v4, 3 = cp.async.ca
__global__ void __launch_bounds__(512, 1) k(const uint8_t* __restrict__ src, int pitch, int nchunks, int rows, int Wl, int iters,
                                            long long* out, size_t img_bytes) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int row_items = Wl * nchunks;
    const int cs = rows * Wl + 1;                   // chunk stride in slots (odd)
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint8_t* tile = src + ((size_t)(blockIdx.x * 37 + it * 11) % 64) * (size_t)pitch * 97;   // wander inside the buffer
        for (int r = warp; r < rows; r += nwarps) {
            const uint8_t* rowp = tile + (size_t)r * 304 * pitch;
            const uint32_t srow = smem_u32(smem) + ((uint32_t)(r * Wl) << 4);
#pragma unroll 4
            for (int i = lane; i < row_items; i += 32) {
                const int cx = i / nchunks, j = i - cx * nchunks;
                const uint8_t* g = rowp + (size_t)cx * pitch + j * 16;
                const uint32_t d = srow + ((uint32_t)(j * cs + cx) << 4);
                if (MODE == 0) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(g) : "memory");
                else if (MODE == 1) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(g), "r"(0) : "memory");
                else if (MODE == 3) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(g) : "memory");
                else { uint4 v = *reinterpret_cast<const uint4*>(g); asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(d), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
            }
        }
        if (MODE != 2) { asm volatile("cp.async.commit_group;\n" ::: "memory"); asm volatile("cp.async.wait_group 2;\n" ::: "memory"); }
    }
    if (MODE != 2) asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

int main(int argc, char** argv) {
    const int grid = argc > 1 ? atoi(argv[1]) : 148;
    const size_t bytes = 64ull << 20;          // 64 MiB: L2 resident after the first pass
    uint8_t* src; cudaMalloc(&src, bytes + (8 << 20)); cudaMemset(src, 1, bytes + (8 << 20));
    long long* d; cudaMalloc(&d, 148 * 8);
    const int rows = 10, Wl = 64, iters = 200;
    for (int mode = 0; mode < 1; ++mode)
        for (int pitch : {128})
            for (int nchunks : {2, 4, 8})
                for (int warps : {4, 8, 16}) {
                    if (nchunks * 16 > pitch) continue;
                    const size_t smem = (size_t)nchunks * (rows * Wl + 1) * 16 + 256;
                    if (smem > 200 * 1024) continue;
                    auto launch = [&](int it) {
                        switch (mode) {
                            case 0: cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); k<0><<<grid, warps * 32, smem>>>(src, pitch, nchunks, rows, Wl, it, d, bytes); break;
                            case 1: cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); k<1><<<grid, warps * 32, smem>>>(src, pitch, nchunks, rows, Wl, it, d, bytes); break;
                            case 2: cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); k<2><<<grid, warps * 32, smem>>>(src, pitch, nchunks, rows, Wl, it, d, bytes); break;
                            default: cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); k<3><<<grid, warps * 32, smem>>>(src, pitch, nchunks, rows, Wl, it, d, bytes); break;
                        }
                    };
                    launch(20);
                    cudaDeviceSynchronize();
                    launch(iters);
                    cudaError_t e = cudaDeviceSynchronize();
                    long long c[148]; cudaMemcpy(c, d, sizeof(c), cudaMemcpyDeviceToHost);
                    double avg = 0; for (int i = 0; i < grid; ++i) avg += c[i]; avg /= grid;
                    const double b = (double)iters * rows * Wl * nchunks * 16;
                    printf("grid %3d mode %d pitch %4d nchunks %d warps %2d : %6.2f B/clk/SM  (%s)\n", grid, mode, pitch, nchunks, warps, b / avg, e == cudaSuccess ? "ok" : cudaGetErrorString(e));
                }
    return 0;
}
