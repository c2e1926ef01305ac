// Micro-benchmark: bytes/clk/SM of TMA tensor loads (cp.async.bulk.tensor) of NHWC halo tiles into the chunk-planar,
// pixel-linear shared-memory layout of the convolution kernels: the activation [B][H][W][C] bf16 is described to TMA as
// a 5-D tensor (8 channels, W, H, C/8, B) so that one box (8, Wl, rows, nch, 1) lands as [chunk][row][col] x 16 B.
// Also measured: the same bytes as a chunk-planar GLOBAL layout [B][C/8][H][W][8] (box rows are then contiguous KBs).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap map, int box_bytes, int tiles_x, int tiles_y, int B,
                                            int Wt, int Ht, int nchk, int nch, int iters, int depth, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[8];
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        int issued = 0, done = 0;
        const int total = iters;
        uint32_t phase_bits = 0;
        while (done < total) {
            while (issued < total && issued - done < depth) {
                const int s = issued % depth;
                const int t = (blockIdx.x * 131 + issued) % (tiles_x * tiles_y * B * nchk);
                int r = t;
                const int ck = r % nchk; r /= nchk;
                const int tx = r % tiles_x; r /= tiles_x;
                const int ty = r % tiles_y; r /= tiles_y;
                const int b = r;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar[s])), "r"(box_bytes) : "memory");
                asm volatile(
                    "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2,%3,%4,%5,%6}], [%7];\n" ::"r"(
                        smem_u32(smem + (size_t)s * box_bytes)),
                    "l"(&map), "r"(0), "r"(tx * Wt - 1), "r"(ty * Ht - 1), "r"(ck * nch), "r"(b), "r"(smem_u32(&bar[s]))
                    : "memory");
                ++issued;
            }
            const int s = done % depth;
            const uint32_t par = (phase_bits >> s) & 1u;
            uint32_t ok = 0;
            while (!ok) {
                asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n"
                             : "=r"(ok) : "r"(smem_u32(&bar[s])), "r"(par) : "memory");
            }
            phase_bits ^= (1u << s);
            ++done;
        }
        out[blockIdx.x] = clock64() - t0;
    }
}

__global__ void __launch_bounds__(128, 1) k4(const __grid_constant__ CUtensorMap map, int box_bytes, int tiles_x, int tiles_y, int B,
                                             int Wt, int Ht, int nchk, int nch, int iters, int depth, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[8];
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        int issued = 0, done = 0;
        uint32_t phase_bits = 0;
        while (done < iters) {
            while (issued < iters && issued - done < depth) {
                const int s = issued % depth;
                const int t = (blockIdx.x * 131 + issued) % (tiles_x * tiles_y * B * nchk);
                int r = t;
                const int ck = r % nchk; r /= nchk;
                const int tx = r % tiles_x; r /= tiles_x;
                const int ty = r % tiles_y; r /= tiles_y;
                const int b = r;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar[s])), "r"(box_bytes) : "memory");
                asm volatile(
                    "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2,%3,%4,%5}], [%6];\n" ::"r"(
                        smem_u32(smem + (size_t)s * box_bytes)),
                    "l"(&map), "r"((tx * Wt - 1) * 2), "r"(ty * Ht - 1), "r"(ck * nch), "r"(b), "r"(smem_u32(&bar[s]))
                    : "memory");
                ++issued;
            }
            const int s = done % depth;
            const uint32_t par = (phase_bits >> s) & 1u;
            uint32_t ok = 0;
            while (!ok) {
                asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n"
                             : "=r"(ok) : "r"(smem_u32(&bar[s])), "r"(par) : "memory");
            }
            phase_bits ^= (1u << s);
            ++done;
        }
        out[blockIdx.x] = clock64() - t0;
    }
}

int main(int argc, char** argv) {
    const int grid = argc > 1 ? atoi(argv[1]) : 148;
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres);
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    const int B = 16, H = 88, W = 304, C = 64;
    void* src; cudaMalloc(&src, (size_t)B * H * W * C * 2); cudaMemset(src, 1, (size_t)B * H * W * C * 2);
    long long* d; cudaMalloc(&d, 148 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int Wl = 64, rows = 10, Wt = 62, Ht = 8;
    for (int layout = 0; layout < 2; ++layout)          // 0 = NHWC source, 1 = chunk-planar source [B][C/8][H][W][8]
        for (int nch : {1, 2, 4, 8})
            for (int depth : {2, 4, 8}) {
                CUtensorMap map;
                cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(C / 8), (cuuint64_t)B};
                cuuint64_t strides[4];
                if (layout == 0) { strides[0] = (cuuint64_t)C * 2; strides[1] = (cuuint64_t)W * C * 2; strides[2] = 16; strides[3] = (cuuint64_t)H * W * C * 2; }
                else { strides[0] = 16; strides[1] = (cuuint64_t)W * 16; strides[2] = (cuuint64_t)H * W * 16; strides[3] = (cuuint64_t)H * W * C * 2; }
                cuuint32_t box[5] = {8, (cuuint32_t)Wl, (cuuint32_t)rows, (cuuint32_t)nch, 1};
                cuuint32_t es[5] = {1, 1, 1, 1, 1};
                CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
                const int box_bytes = Wl * rows * nch * 16;
                if ((size_t)box_bytes * depth > 190 * 1024) continue;
                const int iters = 400;
                const int tiles_x = (W + Wt - 1) / Wt, tiles_y = (H + Ht - 1) / Ht;
                for (int rep = 0; rep < 2; ++rep)
                    k<<<grid, 128, (size_t)box_bytes * depth>>>(map, box_bytes, tiles_x, tiles_y, B, Wt, Ht, (C / 8) / nch, nch, iters, depth, d);
                cudaError_t e = cudaDeviceSynchronize();
                long long c[148]; cudaMemcpy(c, d, grid * 8, cudaMemcpyDeviceToHost);
                double avg = 0; for (int i = 0; i < grid; ++i) avg += c[i]; avg /= grid;
                printf("grid %3d %s box (8ch x %d x %d x %d chunks) = %6d B, depth %d : %7.2f B/clk/SM (%s)\n", grid,
                       layout ? "chunk-planar src" : "NHWC src        ", Wl, rows, nch, box_bytes, depth, (double)iters * box_bytes / avg,
                       e == cudaSuccess ? "ok" : cudaGetErrorString(e));
            }
    // layout 2: chunk-planar source described with MERGED (8 channels x W) rows as 8-byte elements: box rows are Wl*16 B
    cudaFuncSetAttribute(k4, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int nch : {1, 2, 4, 8})
        for (int depth : {2, 4}) {
            CUtensorMap map;
            cuuint64_t dims[4] = {(cuuint64_t)(2 * W), (cuuint64_t)H, (cuuint64_t)(C / 8), (cuuint64_t)B};
            cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)H * W * C * 2};
            cuuint32_t box[4] = {(cuuint32_t)(2 * Wl), (cuuint32_t)rows, (cuuint32_t)nch, 1};
            cuuint32_t es[4] = {1, 1, 1, 1};
            CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
            const int box_bytes = Wl * rows * nch * 16;
            if ((size_t)box_bytes * depth > 190 * 1024) continue;
            const int iters = 400;
            const int tiles_x = (W + Wt - 1) / Wt, tiles_y = (H + Ht - 1) / Ht;
            for (int rep = 0; rep < 2; ++rep)
                k4<<<grid, 128, (size_t)box_bytes * depth>>>(map, box_bytes, tiles_x, tiles_y, B, Wt, Ht, (C / 8) / nch, nch, iters, depth, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long c[148]; cudaMemcpy(c, d, grid * 8, cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < grid; ++i) avg += c[i]; avg /= grid;
            printf("grid %3d chunk-planar src, merged rows: box (%d B x %d x %d chunks) = %6d B, depth %d : %7.2f B/clk/SM (%s)\n", grid,
                   Wl * 16, rows, nch, box_bytes, depth, (double)iters * box_bytes / avg, e == cudaSuccess ? "ok" : cudaGetErrorString(e));
        }
    return 0;
}
