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

// Exact n / d for n * d < 2^32 with one multiply-high (d is a runtime constant of the launch).
struct FastDiv {
    uint32_t m, d;
    __device__ __forceinline__ FastDiv() : m(0), d(1) {}
    __device__ __forceinline__ explicit FastDiv(uint32_t d_) : m(d_ > 1 ? 0xFFFFFFFFu / d_ + 1u : 0u), d(d_) {}
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return d > 1 ? __umulhi(n, m) : n; }
};

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
    FastDiv fd_slots, fd_wl;   // divisions by plane_slots / Wl (filled by prepare())
    __device__ __forceinline__ void prepare() { fd_slots = FastDiv((uint32_t)plane_slots); fd_wl = FastDiv((uint32_t)Wl); }
};

// Raw 16-byte-granular loads of 8 consecutive channels (kept as raw registers so that a batch of independent loads
// can be in flight before any of them is consumed).
template <typename T> struct Raw8;
template <> struct Raw8<bf16> {
    uint4 u;
    __device__ __forceinline__ void load(const bf16* p) { u = __ldg(reinterpret_cast<const uint4*>(p)); }
    __device__ __forceinline__ void zero() { u = make_uint4(0, 0, 0, 0); }
    __device__ __forceinline__ void unpack(float* v) const {
        v[0] = bf16lo(u.x); v[1] = bf16hi(u.x); v[2] = bf16lo(u.y); v[3] = bf16hi(u.y);
        v[4] = bf16lo(u.z); v[5] = bf16hi(u.z); v[6] = bf16lo(u.w); v[7] = bf16hi(u.w);
    }
};
template <> struct Raw8<float> {
    float4 a, b;
    __device__ __forceinline__ void load(const float* p) {
        a = __ldg(reinterpret_cast<const float4*>(p));
        b = __ldg(reinterpret_cast<const float4*>(p + 4));
    }
    __device__ __forceinline__ void zero() { a = make_float4(0.f, 0.f, 0.f, 0.f); b = a; }
    __device__ __forceinline__ void unpack(float* v) const {
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
};

// Stages `nchunks` 8-channel chunks (first channel c0) of the tile.  Work items (slot, chunk) are processed in
// batches of kBatch per thread: all global loads of a batch are issued before the first is consumed, so each
// loader thread keeps kBatch x 16 B (32 B in fp32 mode) in flight instead of one dependent load at a time.
template <typename T, int SPLIT>
__device__ __forceinline__ void stage_tile(const TileSrc& t, uint8_t* dst, int img, int y0, int x0, int c0,
                                           int nchunks, int tid, int nthreads) {
    constexpr int kBatch = (sizeof(T) == 2) ? 8 : 4;
    const int planes = t.S * t.S;
    const int PS = planes * t.plane_slots;
    const int items = PS * nchunks;
    const T* src = reinterpret_cast<const T*>(t.ptr);
    const size_t img_base = (size_t)img * t.H * t.W;
    const FastDiv fd_ch((uint32_t)nchunks);
    const int sh_s = t.S >> 1;                         // S is 1 or 2
    for (int it0 = tid; it0 < items; it0 += nthreads * kBatch) {
        Raw8<T> raw[kBatch];
        int sj[kBatch];            // (slot << 8) | chunk, or -1 when out of range
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int it = it0 + u * nthreads;
            sj[u] = -1;
            raw[u].zero();
            if (it < items) {
                const int s = (int)fd_ch.div((uint32_t)it), j = it - s * nchunks;
                const int q = (int)t.fd_slots.div((uint32_t)s);
                const int rs = s - q * t.plane_slots;
                const int r = (int)t.fd_wl.div((uint32_t)rs);
                const int cx = rs - r * t.Wl;
                const int py = q >> sh_s, px = q - (py << sh_s);
                const int iy = (y0 + t.oy0 + r) * t.S + py;
                const int ix = (x0 + t.ox0 + cx) * t.S + px;
                const bool ok = (r < t.plane_rows) && (r < t.vrows) && (cx < t.vcols) && iy >= 0 && iy < t.H && ix >= 0 && ix < t.W;
                sj[u] = (s << 8) | j | (ok ? 0 : 0x80);
                if (ok) raw[u].load(src + (img_base + (size_t)iy * t.W + ix) * t.pitch + t.coff + c0 + j * 8);
            }
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            if (sj[u] < 0) continue;
            const int s = sj[u] >> 8, j = sj[u] & 0x7F;
            const bool ok = !(sj[u] & 0x80);
            float v[8];
            raw[u].unpack(v);
            if (ok && t.sc != nullptr) {
                const int c = c0 + j * 8;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float y = fmaf(v[k], t.sc[c + k], t.sh[c + k]);
                    v[k] = y > 0.f ? y : y * t.slope;
                }
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
}

struct PipeState {
    int stage; uint32_t phase; int depth;
    __device__ __forceinline__ PipeState(int d) : stage(0), phase(0), depth(d) {}
    __device__ __forceinline__ void advance() { if (++stage == depth) { stage = 0; phase ^= 1; } }
};

}  // namespace rd
