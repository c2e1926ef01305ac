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
    int nplanes = 0;        // planes to stage, from plane 0 on (0 = S*S)
    int plane_slots, plane_rows, Wl;
    int oy0, ox0;           // plane-coordinate offset of slot 0 relative to the tile origin
    int vrows, vcols;       // rows/cols (local) beyond which zeros are written
    const float* sc;        // shared-memory copies of the fused BN scale/shift (indexed by channel), or nullptr
    const float* sh;
    float slope;
    FastDivS fd_slots, fd_wl;   // divisions by plane_slots / Wl (filled by prepare())
    __device__ __forceinline__ void prepare() { fd_slots = FastDivS((uint32_t)plane_slots); fd_wl = FastDivS((uint32_t)Wl); }
};

// Raw 16-byte-granular loads of 8 consecutive channels (kept as raw registers so that a batch of independent loads
// can be in flight before any of them is consumed).
template <typename T> struct Raw8;
template <> struct Raw8<bf16> {
    uint4 u;
    __device__ __forceinline__ void load(const bf16* p) { u = __ldg(reinterpret_cast<const uint4*>(p)); }
    __device__ __forceinline__ void zero() { u = make_uint4(0, 0, 0, 0); }
    __device__ __forceinline__ uint4 bits() const { return u; }
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
    __device__ __forceinline__ uint4 bits() const { return make_uint4(0, 0, 0, 0); }   // (never used: fp32 always converts)
    __device__ __forceinline__ void unpack(float* v) const {
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
};

// Stages `nchunks` 8-channel chunks (first channel c0) of the tile.  Loader warp w owns tile rows w, w+nwarps, ...;
// its (row, column, chunk) items are flattened (chunk fastest, so a warp reads contiguous nchunks*16 B per pixel)
// and processed in batches of kBatch per lane with every global load of a batch issued before any is consumed,
// so one DRAM/L2 latency is paid per batch, not per row.  A raw bf16 source with no fused transform is moved as
// 16-byte words without unpacking.  `cs` = chunk stride in slots (>= planes*plane_slots, padded by the host so
// that the 16-byte stores of one quarter-warp fall into distinct shared-memory banks).
// Slots past plane_rows*Wl are never written here: the caller zero-fills the ring once at kernel start.
template <typename T, int SPLIT>
__device__ __forceinline__ void stage_tile(const TileSrc& t, uint8_t* dst, int cs, int img, int y0, int x0, int c0,
                                           int nchunks, int warp_idx, int nwarps, int lane) {
    constexpr int kBatch = (sizeof(T) == 2) ? 8 : 4;
    const int planes = t.nplanes > 0 ? t.nplanes : t.S * t.S;
    const int sh_s = t.S >> 1;                         // S is 1 or 2
    const T* src = reinterpret_cast<const T*>(t.ptr);
    const int row_items = t.Wl * nchunks;
    const FastDivS fd_ch((uint32_t)nchunks), fd_row((uint32_t)row_items), fd_pr((uint32_t)t.plane_rows);
    const bool raw_copy = (SPLIT == 1) && (sizeof(T) == 2) && (t.sc == nullptr);
    const int nrows = planes * t.plane_rows;
    const int my_rows = nrows > warp_idx ? (nrows - warp_idx + nwarps - 1) / nwarps : 0;
    const int my_items = my_rows * row_items;
    const size_t img_off = (size_t)img * t.H;
    for (int f0 = lane; f0 < my_items; f0 += 32 * kBatch) {
        Raw8<T> raw[kBatch];
        int sj[kBatch];                                // (slot << 8) | j | 0x80 if zero, -1 past the end
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int f = f0 + u * 32;
            sj[u] = -1;
            raw[u].zero();
            if (f < my_items) {
                const int k = (int)fd_row.div((uint32_t)f);
                const int it = f - k * row_items;
                const int rr = warp_idx + k * nwarps;
                const int q = (int)fd_pr.div((uint32_t)rr), r = rr - q * t.plane_rows;
                const int py = q >> sh_s, px = q - (py << sh_s);
                const int cx = (int)fd_ch.div((uint32_t)it), j = it - cx * nchunks;
                const int iy = (y0 + t.oy0 + r) * t.S + py;
                const int ix = (x0 + t.ox0 + cx) * t.S + px;
                const bool ok = (r < t.vrows) && (cx < t.vcols) && iy >= 0 && iy < t.H && ix >= 0 && ix < t.W;
                sj[u] = ((q * t.plane_slots + r * t.Wl + cx) << 8) | j | (ok ? 0 : 0x80);
                if (ok) raw[u].load(src + ((img_off + iy) * t.W + ix) * t.pitch + t.coff + c0 + j * 8);
            }
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            if (sj[u] < 0) continue;
            const int s = sj[u] >> 8, j = sj[u] & 0x7F;
            uint8_t* d = dst + ((size_t)(j * cs + s) << 4);
            if (raw_copy) {
                *reinterpret_cast<uint4*>(d) = raw[u].bits();
                continue;
            }
            float v[8];
            raw[u].unpack(v);
            if (!(sj[u] & 0x80) && t.sc != nullptr) {
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
            *reinterpret_cast<uint4*>(d) = hi;
            if (SPLIT == 3) {
                uint4 lo;
                lo.x = pack_bf16x2(v[0] - bf16lo(hi.x), v[1] - bf16hi(hi.x));
                lo.y = pack_bf16x2(v[2] - bf16lo(hi.y), v[3] - bf16hi(hi.y));
                lo.z = pack_bf16x2(v[4] - bf16lo(hi.z), v[5] - bf16hi(hi.z));
                lo.w = pack_bf16x2(v[6] - bf16lo(hi.w), v[7] - bf16hi(hi.w));
                *reinterpret_cast<uint4*>(d + ((size_t)(nchunks * cs) << 4)) = lo;
            }
        }
    }
}

// True when the tile can be staged with fire-and-forget cp.async copies: bf16 storage, no hi/lo split, no fused
// BatchNorm/activation on load.
template <typename T, int SPLIT>
__device__ __forceinline__ bool tile_is_raw(const TileSrc& t) {
    return (SPLIT == 1) && (sizeof(T) == 2) && (t.sc == nullptr);
}

// Asynchronous variant of stage_tile for raw tiles: every (row, column, chunk) item becomes one 16-byte cp.async
// (zero-filling out-of-range items), nothing is held in registers and the caller overlaps the copies with other
// work; completion is observed with cp.async.wait_group by the issuing thread.
template <typename T>
__device__ __forceinline__ void stage_tile_async(const TileSrc& t, uint8_t* dst, int cs, int img, int y0, int x0, int c0,
                                                 int nchunks, int warp_idx, int nwarps, int lane) {
    const int planes = t.nplanes > 0 ? t.nplanes : t.S * t.S;
    const int sh_s = t.S >> 1;
    const T* src = reinterpret_cast<const T*>(t.ptr);
    const int row_items = t.Wl * nchunks;
    const FastDivS fd_ch((uint32_t)nchunks), fd_row((uint32_t)row_items), fd_pr((uint32_t)t.plane_rows);
    const int nrows = planes * t.plane_rows;
    const int my_rows = nrows > warp_idx ? (nrows - warp_idx + nwarps - 1) / nwarps : 0;
    const int my_items = my_rows * row_items;
    const size_t img_off = (size_t)img * t.H;
#pragma unroll 4
    for (int f = lane; f < my_items; f += 32) {
        const int k = (int)fd_row.div((uint32_t)f);
        const int it = f - k * row_items;
        const int rr = warp_idx + k * nwarps;
        const int q = (int)fd_pr.div((uint32_t)rr), r = rr - q * t.plane_rows;
        const int py = q >> sh_s, px = q - (py << sh_s);
        const int cx = (int)fd_ch.div((uint32_t)it), j = it - cx * nchunks;
        const int iy = (y0 + t.oy0 + r) * t.S + py;
        const int ix = (x0 + t.ox0 + cx) * t.S + px;
        const bool ok = (r < t.vrows) && (cx < t.vcols) && iy >= 0 && iy < t.H && ix >= 0 && ix < t.W;
        const int s = q * t.plane_slots + r * t.Wl + cx;
        const T* g = ok ? src + ((img_off + iy) * t.W + ix) * t.pitch + t.coff + c0 + j * 8 : src;
        cp_async16(dst + ((size_t)(j * cs + s) << 4), g, ok ? 16u : 0u);
    }
}

// Zero-fills a shared-memory region (16-byte granular) cooperatively; used once per CTA for the operand rings.
__device__ __forceinline__ void zero_smem(uint8_t* base, size_t bytes, int tid, int nthreads) {
    uint4* p = reinterpret_cast<uint4*>(base);
    const size_t n = bytes >> 4;
    for (size_t i = tid; i < n; i += nthreads) p[i] = make_uint4(0, 0, 0, 0);
}

struct PipeState {
    int stage; uint32_t phase; int depth;
    __device__ __forceinline__ PipeState(int d) : stage(0), phase(0), depth(d) {}
    __device__ __forceinline__ void advance() { if (++stage == depth) { stage = 0; phase ^= 1; } }
};

}  // namespace rd
