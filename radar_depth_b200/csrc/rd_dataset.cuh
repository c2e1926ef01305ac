// rd_dataset.cuh -- the reference's per-sample input pipeline as batch kernels.
//
// Replaces, for a whole batch resident in HBM as RAW exported data (uint8 HWC image, int16 x256 lidar / radar depth),
// what dataset/nuscenes_dataset_torch_new.py:237-412 (transform_train) and :415-560 (transform_val) do per sample on
// CPU workers with scipy.ndimage / PIL: depth decode (/256), nearest rotation (scipy.ndimage.rotate, order 0, constant 0),
// resize by the random scale (scipy.misc.imresize = min-max "bytescale" to uint8 + PIL bilinear for the image, PIL
// nearest in mode 'F' for depth), crop, horizontal flip, ColorJitter (PIL ImageEnhance Brightness / Contrast / Color in
// a random order), /255, the max_depth filter on the radar channel and the 4-channel concatenation.  Every step is
// integer / IEEE arithmetic restated exactly (the tests demand bit equality with PIL + scipy):
//   * rotation: input coordinate c = (y*m0 + x*m1) + off in double (no FMA contraction: __dmul_rn/__dadd_rn), valid iff
//     0 <= c <= len-1, nearest index floor(c + 0.5)                         (scipy ni_interpolation.c, NI_GeometricTransform);
//   * PIL nearest: source index table built on the host by PIL's running sum xo += in/out (ImagingScaleAffine);
//   * PIL bilinear on 8-bit images: two passes (horizontal, then vertical) with 22-bit fixed-point coefficients built on
//     the host (precompute_coeffs / normalize_coeffs_8bpc of Resample.c), each pass rounded to uint8 (clip8);
//   * ImageEnhance: out = (uint8)(d + alpha*(v - d)) in float32, alpha = (float)factor, clamped when alpha is outside
//     [0,1] (ImagingBlend); d = 0 (Brightness), the rounded mean of L (Contrast), L of the pixel (Color);
//     L = (R*19595 + G*38470 + B*7471 + 0x8000) >> 16 (ImagingConvert rgb2l).
#pragma once
#include "rd_common.cuh"
#include "../../include/radar_depth_b200.h"

namespace rd {

__device__ __forceinline__ bool aug_rot_src(const rd_aug_sample& s, int H, int W, int y, int x, int* iy, int* ix) {
    if (s.identity_rot) { *iy = y; *ix = x; return true; }
    const double cy = __dadd_rn(__dadd_rn(__dmul_rn((double)y, s.m00), __dmul_rn((double)x, s.m01)), s.off0);
    const double cx = __dadd_rn(__dadd_rn(__dmul_rn((double)y, s.m10), __dmul_rn((double)x, s.m11)), s.off1);
    if (cy < 0.0 || cy > (double)(H - 1) || cx < 0.0 || cx > (double)(W - 1)) return false;
    int a = (int)floor(__dadd_rn(cy, 0.5)), b = (int)floor(__dadd_rn(cx, 0.5));
    *iy = min(max(a, 0), H - 1);
    *ix = min(max(b, 0), W - 1);
    return true;
}

// value of the rotated image at (y, x), channel c (0 where the rotation reads outside the source)
__device__ __forceinline__ int aug_rot_pix(const uint8_t* __restrict__ img, const rd_aug_sample& s, int H, int W, int y, int x, int c) {
    int iy, ix;
    if (!aug_rot_src(s, H, W, y, x, &iy, &ix)) return 0;
    return (int)img[((size_t)iy * W + ix) * 3 + c];
}

// ---- K1: min / max of the rotated float image (scipy bytescale's cmin / cmax).  mm[b] = {min, max}, preset to {255, 0}.
__global__ void aug_minmax_kernel(const uint8_t* __restrict__ images, const rd_aug_sample* __restrict__ samples, int H, int W, int* mm) {
    pdl_enter();
    const int b = blockIdx.y;
    const rd_aug_sample s = samples[b];
    const uint8_t* img = images + (size_t)b * H * W * 3;
    int lo = 255, hi = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
        const int y = i / W, x = i - y * W;
        int iy, ix;
        if (aug_rot_src(s, H, W, y, x, &iy, &ix)) {
            const uint8_t* p = img + ((size_t)iy * W + ix) * 3;
            lo = min(lo, min((int)p[0], min((int)p[1], (int)p[2])));
            hi = max(hi, max((int)p[0], max((int)p[1], (int)p[2])));
        } else {
            lo = 0;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&mm[2 * b], lo);
        atomicMax(&mm[2 * b + 1], hi);
    }
}

// scipy bytescale of one value of the rotated image: (v - cmin) * scale, clipped to [0, 255], + 0.5, truncated
__device__ __forceinline__ int aug_bytescale(int v, float cmin, float scale) {
    float t = __fmul_rn(__fsub_rn((float)v, cmin), scale);
    t = fminf(fmaxf(t, 0.f), 255.f);
    return (int)__fadd_rn(t, 0.5f);
}

// ---- K2: rotate -> bytescale -> PIL bilinear resize -> crop -> flip, one thread per output pixel.  tab[b] holds, for the
// crop window only, rows then columns: {first source index, taps, k0, k1, k2} (22-bit fixed point).
__global__ void aug_rgb_kernel(const uint8_t* __restrict__ images, const rd_aug_sample* __restrict__ samples, const int* __restrict__ tab,
                               const int* __restrict__ mm, int H, int W, int ch, int cw, uint8_t* __restrict__ out) {
    pdl_enter();
    const int b = blockIdx.y;
    const rd_aug_sample s = samples[b];
    const uint8_t* img = images + (size_t)b * H * W * 3;
    const int* rows = tab + (size_t)b * (ch + cw) * 5;
    const int* cols = rows + (size_t)ch * 5;
    const int cmin_i = mm[2 * b], cmax_i = mm[2 * b + 1];
    const float cmin = (float)cmin_i;
    const int cscale = (cmax_i - cmin_i) == 0 ? 1 : (cmax_i - cmin_i);
    const float scale = (float)(255.0 / (double)cscale);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ch * cw; i += gridDim.x * blockDim.x) {
        const int y = i / cw, x = i - y * cw;
        const int xs = s.flip ? (cw - 1 - x) : x;                 // column of the un-flipped crop
        const int* R = rows + y * 5;
        const int* Cc = cols + xs * 5;
        int acc_v[3] = {1 << 21, 1 << 21, 1 << 21};
        for (int ky = 0; ky < R[1]; ++ky) {
            const int yy = R[0] + ky;
            int acc_h[3] = {1 << 21, 1 << 21, 1 << 21};
            for (int kx = 0; kx < Cc[1]; ++kx) {
                const int xx = Cc[0] + kx;
                int iy, ix;
                int v0 = 0, v1 = 0, v2 = 0;
                if (aug_rot_src(s, H, W, yy, xx, &iy, &ix)) {
                    const uint8_t* p = img + ((size_t)iy * W + ix) * 3;
                    v0 = p[0]; v1 = p[1]; v2 = p[2];
                }
                const int k = Cc[2 + kx];
                acc_h[0] += aug_bytescale(v0, cmin, scale) * k;
                acc_h[1] += aug_bytescale(v1, cmin, scale) * k;
                acc_h[2] += aug_bytescale(v2, cmin, scale) * k;
            }
            const int k = R[2 + ky];
#pragma unroll
            for (int c = 0; c < 3; ++c) acc_v[c] += min(max(acc_h[c] >> 22, 0), 255) * k;
        }
        uint8_t* o = out + (((size_t)b * ch + y) * cw + x) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) o[c] = (uint8_t)min(max(acc_v[c] >> 22, 0), 255);
    }
}

__device__ __forceinline__ int aug_L(int r, int g, int b) { return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16; }

// ---- K3: sum of L over every image (ImageStat.Stat(image.convert("L")).mean of ImageEnhance.Contrast)
__global__ void aug_lsum_kernel(const uint8_t* __restrict__ img8, int npix, unsigned long long* __restrict__ lsum) {
    pdl_enter();
    const int b = blockIdx.y;
    const uint8_t* img = img8 + (size_t)b * npix * 3;
    unsigned int acc = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x)
        acc += (unsigned int)aug_L(img[3 * i], img[3 * i + 1], img[3 * i + 2]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&lsum[b], (unsigned long long)acc);
}

__device__ __forceinline__ uint8_t aug_blend(int d, int v, float alpha, bool inside) {
    const float t = __fadd_rn((float)d, __fmul_rn(alpha, (float)(v - d)));
    if (inside) return (uint8_t)(int)t;
    if (t <= 0.f) return 0;
    if (t >= 255.f) return 255;
    return (uint8_t)(int)t;
}

// ---- K4: round r of ColorJitter: every sample applies ITS r-th operation in place (0 brightness, 1 contrast, 2 colour)
__global__ void aug_jitter_kernel(uint8_t* __restrict__ img8, const rd_aug_sample* __restrict__ samples, int round, int npix,
                                  const unsigned long long* __restrict__ lsum) {
    pdl_enter();
    const int b = blockIdx.y;
    const rd_aug_sample s = samples[b];
    const int op = s.op[round];
    const double f = s.factor[round];
    if (f == 1.0) return;                                       // Image.blend returns a copy of the image
    const float alpha = (float)f;
    const bool inside = f >= 0.0 && f <= 1.0;
    uint8_t* img = img8 + (size_t)b * npix * 3;
    int mean = 0;
    if (op == 1) mean = (int)((double)lsum[b] / (double)npix + 0.5);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        uint8_t* p = img + 3 * (size_t)i;
        const int r = p[0], g = p[1], bl = p[2];
        if (f == 0.0) {                                          // Image.blend returns a copy of the degenerate image
            const int d = op == 0 ? 0 : (op == 1 ? mean : aug_L(r, g, bl));
            p[0] = p[1] = p[2] = (uint8_t)d;
            continue;
        }
        const int d = op == 0 ? 0 : (op == 1 ? mean : aug_L(r, g, bl));
        p[0] = aug_blend(d, r, alpha, inside);
        p[1] = aug_blend(d, g, alpha, inside);
        p[2] = aug_blend(d, bl, alpha, inside);
    }
}

// ---- K5: final assembly.  mode 0 (train): rgb from the jittered crop, depth through flip -> crop -> PIL-nearest table ->
// rotation; mode 1 (val): centre crop of the raw data.  inputs [B][3+has_radar][ch][cw], labels / radar_out [B][1][ch][cw].
__device__ __forceinline__ float aug_depth_value(const int16_t* __restrict__ d, const rd_aug_sample& s, int H, int W, int ry, int rx, bool rot) {
    int iy = ry, ix = rx;
    if (rot && !aug_rot_src(s, H, W, ry, rx, &iy, &ix)) return 0.f;
    const float v = (float)d[(size_t)iy * W + ix] * (1.0f / 256.0f);         // exact: |raw| < 2^15
    return __fdiv_rn(v, s.depth_div);
}

__global__ void aug_pack_kernel(const uint8_t* __restrict__ img8, const uint8_t* __restrict__ images, const int16_t* __restrict__ lidar,
                                const int16_t* __restrict__ radar, const rd_aug_sample* __restrict__ samples, const int* __restrict__ ntab,
                                int H, int W, int ch, int cw, int mode, int has_radar, float max_depth, float* __restrict__ inputs,
                                float* __restrict__ labels, float* __restrict__ radar_out) {
    pdl_enter();
    const int b = blockIdx.y;
    const rd_aug_sample s = samples[b];
    const int Cc = 3 + (has_radar ? 1 : 0);
    const size_t plane = (size_t)ch * cw;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ch * cw; i += gridDim.x * blockDim.x) {
        const int y = i / cw, x = i - y * cw;
        int r, g, bl, sy, sx;
        if (mode == 0) {
            const uint8_t* p = img8 + (((size_t)b * ch + y) * cw + x) * 3;
            r = p[0]; g = p[1]; bl = p[2];
            const int xs = s.flip ? (cw - 1 - x) : x;
            const int* nt = ntab + (size_t)b * (ch + cw);
            sy = nt[y];                                          // row / column of the ROTATED image PIL's nearest resize reads
            sx = nt[ch + xs];
        } else {
            sy = s.crop_i + y; sx = s.crop_j + x;
            const uint8_t* p = images + (((size_t)b * H + sy) * W + sx) * 3;
            r = p[0]; g = p[1]; bl = p[2];
        }
        float* in = inputs + (size_t)b * Cc * plane + i;
        in[0] = (float)((double)r / 255.0);
        in[plane] = (float)((double)g / 255.0);
        in[2 * plane] = (float)((double)bl / 255.0);
        const float lv = aug_depth_value(lidar + (size_t)b * H * W, s, H, W, sy, sx, mode == 0);
        labels[(size_t)b * plane + i] = lv;
        float rv = aug_depth_value(radar + (size_t)b * H * W, s, H, W, sy, sx, mode == 0);
        if (rv > max_depth) rv = 0.f;
        radar_out[(size_t)b * plane + i] = rv;
        if (has_radar) in[3 * plane] = rv;
    }
}

}  // namespace rd
