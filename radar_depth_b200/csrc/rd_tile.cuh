// rd_tile.cuh -- staging of an NHWC halo tile into the chunk-planar, pixel-linear shared-memory layout
// consumed by tcgen05.mma through SWIZZLE_NONE descriptors (see rd_conv_fprop.cuh / rd_conv_wgrad.cuh).
//
//   smem[part][chunk j][plane q][slot] : 16 bytes = 8 consecutive channels (bf16) of one pixel
//   slot = r*Wl + cx  <->  source pixel ((y0+oy0+r)*S + q/S, (x0+ox0+cx)*S + q%S)
//
// Out-of-image pixels, rows >= plane_rows, and (for gradient tiles) columns >= vcols / rows >= vrows are
// written as zeros, which is exactly the convolution's zero padding.  The optional per-channel affine +
// leaky-ReLU is the producer's BatchNorm+activation fused into the consumer's operand path
// (reference models.py:99-101: relu(bn1(conv1(x))) is never materialised here).
#pragma once
#include "rd_common.cuh"

namespace rd {

struct TileSrc {
    const void* ptr;
    int pitch, coff;        // NHWC view
    int H, W;               // image size of the view
    int S;                  // 1, or 2 = four parity planes
    int plane_slots, plane_rows, Wl;
    int oy0, ox0;           // plane-coordinate offset of slot 0 relative to the tile origin
    int vrows, vcols;       // rows/cols (local) beyond which zeros are written
    const float* sc;        // shared-memory copies of the fused BN scale/shift (indexed by channel), or nullptr
    const float* sh;
    float slope;
};

template <typename T, int SPLIT>
__device__ __forceinline__ void stage_tile(const TileSrc& t, uint8_t* dst, int img, int y0, int x0, int c0,
                                           int nchunks, int tid, int nthreads) {
    const int planes = t.S * t.S;
    const int PS = planes * t.plane_slots;
    const int items = PS * nchunks;
    const T* src = reinterpret_cast<const T*>(t.ptr);
    for (int it = tid; it < items; it += nthreads) {
        const int s = it / nchunks, j = it - s * nchunks;
        const int q = s / t.plane_slots;
        const int rs = s - q * t.plane_slots;
        const int r = rs / t.Wl;
        const int cx = rs - r * t.Wl;
        const int py = q / t.S, px = q - py * t.S;
        const int iy = (y0 + t.oy0 + r) * t.S + py;
        const int ix = (x0 + t.ox0 + cx) * t.S + px;
        float v[8];
        const bool ok = (r < t.plane_rows) && (r < t.vrows) && (cx < t.vcols) && iy >= 0 && iy < t.H && ix >= 0 && ix < t.W;
        if (ok) {
            const int c = c0 + j * 8;
            const size_t off = (((size_t)img * t.H + iy) * t.W + ix) * t.pitch + t.coff + c;
            Act<T>::load8(src + off, v);
            if (t.sc != nullptr) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float y = fmaf(v[k], t.sc[c + k], t.sh[c + k]);
                    v[k] = y > 0.f ? y : y * t.slope;
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = 0.f;
        }
        uint4 hi;
        hi.x = pack_bf16x2(v[0], v[1]); hi.y = pack_bf16x2(v[2], v[3]);
        hi.z = pack_bf16x2(v[4], v[5]); hi.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(dst + ((size_t)(j * PS + s) << 4)) = hi;
        if (SPLIT == 3) {
            uint4 lo;
            lo.x = pack_bf16x2(v[0] - bf16lo(hi.x), v[1] - bf16hi(hi.x));
            lo.y = pack_bf16x2(v[2] - bf16lo(hi.y), v[3] - bf16hi(hi.y));
            lo.z = pack_bf16x2(v[4] - bf16lo(hi.z), v[5] - bf16hi(hi.z));
            lo.w = pack_bf16x2(v[6] - bf16lo(hi.w), v[7] - bf16hi(hi.w));
            *reinterpret_cast<uint4*>(dst + ((size_t)((nchunks + j) * PS + s) << 4)) = lo;
        }
    }
}

struct PipeState {
    int stage; uint32_t phase; int depth;
    __device__ __forceinline__ PipeState(int d) : stage(0), phase(0), depth(d) {}
    __device__ __forceinline__ void advance() { if (++stage == depth) { stage = 0; phase ^= 1; } }
};

}  // namespace rd
