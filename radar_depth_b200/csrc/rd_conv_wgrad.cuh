// rd_conv_wgrad.cuh -- convolution weight gradient on tcgen05.
//
// dw[tap][co][ci] = sum over pixels of gy[pixel][co] * x[pixel + tap][ci].  The contraction index is the pixel,
// so the gradient tile (A, M = output channels) and the source halo tile (B, N = input channels) are both
// staged in the same chunk-planar pixel-linear layout as the forward kernel and handed to tcgen05.mma as
// MN-major SWIZZLE_NONE operands: the 8 channels of a 16-byte unit are 8 consecutive M (or N) indices, the
// slot index is K.  A tap is again only a start-address shift of the B descriptor.  Every tap owns a TMEM
// accumulator [128 x Nc]; a CTA walks its share of the pixel tiles with the accumulators resident and
// finally adds them into dw with vector fp32 reductions (REDG.128).
//
// Warp roles (416 threads): warps 0-3 epilogue, warps 4-11 loaders (gy raw; x with fused BN+activation),
// warp 12 lane 0 UMMA issuer (also owns TMEM alloc/dealloc).
#pragma once
#include "rd_common.cuh"
#include "rd_tile.cuh"
#include "rd_conv_fprop.cuh"      // umma_bf16_elected
#include "../../include/radar_depth_b200.h"

namespace rd {

constexpr int kWgLoaderWarps = 12;
constexpr int kWgradThreads = (4 + kWgLoaderWarps + 1) * 32;   // epilogue x4, loaders, UMMA issuer
constexpr int kMaxTapsPerCta = 16;       // taps (TMEM accumulators) handled by one CTA
constexpr int kWgSmemHeader = 10240;     // barriers + tmem slot + BN scale/shift (2 x 1024 floats)
constexpr int kWgOffTmemSlot = 256;
constexpr int kWgOffTaps = 512;          // int[2][32]: gradient-plane offset, source shift (16-byte units)
constexpr int kWgOffLdScale = 1024;
constexpr int kWgOffLdShift = 5120;

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}
// Deterministic mode: the pixel-split CTA blockIdx.x stores its accumulators to its own copy of dw (det = true) and
// wgrad_reduce_kernel adds the copies in index order; otherwise vector fp32 reductions straight into dw.
__device__ __forceinline__ void acc_out_v4(bool det, float* p, float a, float b, float c, float d) {
    if (det) *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
    else red_add_v4(p, a, b, c, d);
}
// dw[i] += part[0][i] + part[1][i] + ... (fixed order)
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, int nparts, size_t n, float* __restrict__ dw) {
    pdl_enter();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < nparts; ++k) s += part[(size_t)k * n + i];
        dw[i] += s;
    }
}

template <typename T, int SPLIT>
__global__ void __launch_bounds__(kWgradThreads, 1) conv_wgrad_kernel(const __grid_constant__ rd_wgrad_params p,
                                                                      const __grid_constant__ CUtensorMap g_map,
                                                                      const __grid_constant__ CUtensorMap x_map, const int tma_mode,
                                                                      float* const det_part) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint64_t* full = bars;
    uint64_t* empty = bars + 8;
    uint64_t* tmem_full = bars + 16;
    uint64_t* tma_full = bars + 24;                // [8] TMA landing barriers (gradient tiles need a fix-up hop, see below)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kWgOffTmemSlot);
    float* ld_sc = reinterpret_cast<float*>(smem + kWgOffLdScale);
    float* ld_sh = reinterpret_cast<float*>(smem + kWgOffLdShift);
    uint8_t* ring = smem + kWgSmemHeader;
    int* tap_g = reinterpret_cast<int*>(smem + kWgOffTaps);
    int* tap_x = tap_g + 32;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int kWarpMma = 4 + kWgLoaderWarps;
    // dbg_flags & 8: the four counters become a timeline (cycles since kernel entry): first stage ready, last UMMA issued,
    // accumulators complete (epilogue starts), epilogue done
    const bool tl_mode = p.dbg && (p.dbg_flags & 8);
    const long long t_entry = p.dbg ? clock64() : 0;
    const size_t dbg_ncta = (size_t)gridDim.x * gridDim.y * gridDim.z;
    const size_t dbg_cta = blockIdx.x + (size_t)gridDim.x * (blockIdx.y + (size_t)gridDim.y * blockIdx.z);
    const int cob = blockIdx.y;
    const int cib = blockIdx.z / p.ntg;
    const int tg = blockIdx.z - cib * p.ntg;
    const int t0 = tg * p.tg_size;
    const int R = p.gcopies > 1 ? p.gcopies : 1;               // gradient copies / planes stacked in M (rd_wgrad_params.gcopies)
    const int Rload = (p.Sg == 1) ? R : 1;                     // Sg = 2: the R rows blocks are the parity planes staged anyway
    const int T_n = min(p.tg_size, (R > 1 ? p.njobs : p.ntaps) - t0);
    const int co0 = cob * p.Mc, ci0 = cib * p.Nc;
    const int tiles_per_img = p.tiles_y * p.tiles_x;
    const int ntiles = tiles_per_img * p.B;
    const int g_chunks = p.Mc >> 3, x_chunks = p.Nc >> 3;
    const int GPS = p.g_chunk_stride;              // chunk strides in slots (>= planes * plane slots)
    const int XPS = p.x_chunk_stride;

    // raw bf16 tiles are staged with cp.async (one arrival per loader thread), transformed tiles through registers
    // (one arrival per loader warp)
    // Staging modes per operand.  TMA (raw bf16, stride 1; decided by the launcher, which also encodes the tensor maps and
    // hands the kernel the dense chunk strides TMA writes): ONE thread issues a box load per stage.  Otherwise raw bf16
    // tiles are staged with cp.async (one arrival per loader thread) and transformed tiles through registers (one arrival
    // per loader warp).  A TMA box of the gradient tile is Wl columns wide, so its Wl-Wt junk columns hold the neighbour
    // tile's pixels and must be cleared before the contraction: the TMA lands on tma_full[stage], the worker warps then
    // zero those columns and arrive on full[stage].
    const bool g_tma = (tma_mode & 1) != 0, x_tma = (tma_mode & 2) != 0, any_tma = g_tma || x_tma;
    // tma_mode & 4 (with 1 and 2): both tiles land as 128-byte-swizzled [slot][64 channels] blocks -- the TMA box's inner
    // dimension is a whole 128-byte run of the NHWC pixel instead of a 16-byte granule (4-8x the box rate, no half-used
    // sectors), and the blocks ARE the SWIZZLE_128B MN-major canonical layout: slot = k-row, a tap is still a start-address
    // shift (whole 128-byte rows).  Block strides: KS slots (gradient: copies, then 64-channel blocks) and x_plane_slots.
    const bool sw = (tma_mode & 4) != 0, sw_bo = (tma_mode & 8) != 0;
    const int g_blocks = p.Mc >> 6, x_blocks = p.Nc >> 6;
    // Parity planes (stride-2 gradient of the sub-pixel programs / stride-2 source) are boxes with element stride 2 that land
    // one after the other: gradient [plane][copy][block], source [block][plane].
    const int x_npl = (p.Sx == 1) ? 1 : (p.x_planes > 0 ? p.x_planes : p.Sx * p.Sx);
    const uint32_t GB = (uint32_t)p.KS * 128u, XPL = (uint32_t)p.x_plane_slots * 128u, XB = (uint32_t)x_npl * XPL;
    const bool g_async = (SPLIT == 1) && (sizeof(T) == 2) && !g_tma;
    const bool x_async = (SPLIT == 1) && (sizeof(T) == 2) && (p.ld_scale == nullptr) && !x_tma;
    const bool g_reg = !g_tma && !g_async, x_reg = !x_tma && !x_async;
    const bool x_bn = x_tma && (p.ld_scale != nullptr);   // raw box by TMA, BatchNorm+activation applied in place by the workers
    const bool hop = g_tma || x_bn;                       // TMA lands on tma_full, the workers fix up and arrive on full
    const int nworkers = any_tma ? kWgLoaderWarps - 1 : kWgLoaderWarps;      // warp 4 drives the TMA unit
    const bool warp_arrives = g_reg || x_reg || hop;
    const uint32_t full_count = (uint32_t)((warp_arrives ? nworkers : 0) + ((g_async || x_async) ? 32 * nworkers : 0) +
                                           ((any_tma && !hop) ? 1 : 0));
    if (tid == 0) {
        for (int i = 0; i < p.NS; ++i) { mbar_init(&full[i], full_count); mbar_init(&empty[i], 1); mbar_init(&tma_full[i], 1); }
        mbar_init(tmem_full, 1);
        fence_mbar_init();
    }
    if (tid < T_n) {
        tap_g[tid] = p.taps[t0 + tid].g_off;
        tap_x[tid] = p.taps[t0 + tid].x_shift;
    }
    zero_smem(ring, (size_t)p.NS * p.stage_bytes, tid, kWgradThreads);   // tile tails stay zero forever
    fence_proxy_async_smem();
    if (warp == kWarpMma) tmem_alloc<512>(tmem_slot);
    pdl_enter();      // parameter-only setup above overlaps the preceding kernel's tail (rd_common.cuh)
    if (p.ld_scale) {
        for (int i = tid; i < p.Cin; i += kWgradThreads) { ld_sc[i] = p.ld_scale[i]; ld_sh[i] = p.ld_shift[i]; }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 4 && warp < kWarpMma) {
        // ================= loaders =================
        PipeState st(p.NS);
        TileSrc tg_, tx_;
        tg_.ptr = p.gy.ptr; tg_.pitch = p.gy.pitch; tg_.coff = p.gy.coff; tg_.H = p.gH; tg_.W = p.gW; tg_.S = p.Sg;
        tg_.plane_slots = p.KS; tg_.plane_rows = p.Ht; tg_.Wl = p.Wl; tg_.oy0 = 0; tg_.ox0 = 0;
        tg_.vrows = p.Ht; tg_.vcols = p.Wt; tg_.sc = nullptr; tg_.sh = nullptr; tg_.slope = 1.f;
        tx_.ptr = p.x.ptr; tx_.pitch = p.x.pitch; tx_.coff = p.x.coff; tx_.H = p.xH; tx_.W = p.xW; tx_.S = p.Sx; tx_.nplanes = p.x_planes;
        tx_.plane_slots = p.x_plane_slots; tx_.plane_rows = p.x_plane_rows; tx_.Wl = p.Wl; tx_.oy0 = p.sy_min; tx_.ox0 = p.sx_min;
        tx_.vrows = p.x_plane_rows; tx_.vcols = p.Wl;
        tx_.sc = p.ld_scale ? ld_sc : nullptr; tx_.sh = ld_sh; tx_.slope = p.ld_slope;
        tg_.prepare(); tx_.prepare();
        // Raw tiles go through fire-and-forget cp.async: every loader thread arrives on full[stage] through
        // cp.async.mbarrier.arrive.noinc when ITS copies have landed, so the loaders run ahead as far as the ring
        // allows and never wait for memory.  Tiles with a fused transform are staged through registers and
        // signalled with one ordinary arrival per warp.  The generic->async proxy fence is executed by the consumer
        // (UMMA warp) after it has observed full[stage]: a producer-side fence would have to drain the copies.
        const int widx = any_tma ? warp - 5 : warp - 4;
        const int g_planes = p.Sg * p.Sg;
        const uint32_t g_tx = (uint32_t)(g_planes * p.Wl * p.Ht * g_chunks * 16), x_tx = (uint32_t)((sw ? x_npl : 1) * p.Wl * p.x_plane_rows * x_chunks * 16);
        // in-place BatchNorm transform (x_bn): this thread's chunk, first slot and slot stride
        int bn_j = -1, bn_s0 = 0, bn_tpc = 1, bn_r0 = 0, bn_c0 = 0, bn_dr = 0, bn_dc = 0;
        float bn_sc[8], bn_sh[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { bn_sc[k] = 1.f; bn_sh[k] = 0.f; }
        if (x_bn && widx >= 0) {
            const int t = widx * 32 + lane;
            bn_tpc = (nworkers * 32) / x_chunks;
            if (t < bn_tpc * x_chunks) {
                bn_j = t % x_chunks;
                bn_s0 = t / x_chunks;
                bn_r0 = bn_s0 / p.Wl; bn_c0 = bn_s0 - bn_r0 * p.Wl;
                bn_dr = bn_tpc / p.Wl; bn_dc = bn_tpc - bn_dr * p.Wl;
#pragma unroll
                for (int k = 0; k < 8; ++k) { bn_sc[k] = ld_sc[ci0 + bn_j * 8 + k]; bn_sh[k] = ld_sh[ci0 + bn_j * 8 + k]; }
            }
        }
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int img = tile / tiles_per_img;
            const int trem = tile - img * tiles_per_img;
            const int ty = trem / p.tiles_x, tx = trem - ty * p.tiles_x;
            const int y0 = ty * p.Ht - p.tile_oy, x0 = tx * p.Wt - p.tile_ox;
            const long long tw0 = p.dbg ? clock64() : 0;
            mbar_wait(&empty[st.stage], st.phase ^ 1, 0x500 + st.stage);
            const long long tw1 = p.dbg ? clock64() : 0;
            uint8_t* sbase = ring + (size_t)st.stage * p.stage_bytes;
            if (any_tma && warp == 4) {
                // ---- TMA issuer
                if (lane == 0) {
                    uint64_t* bar = hop ? &tma_full[st.stage] : &full[st.stage];
                    if (!(p.dbg_flags & 2)) {
                        mbar_arrive_expect_tx(bar, (g_tma ? g_tx * (uint32_t)Rload : 0u) + (x_tma ? x_tx : 0u));
                        if (sw) {
                            for (int q = 0; q < g_planes; ++q)
                                for (int r = 0; r < Rload; ++r)
                                    for (int b = 0; b < g_blocks; ++b)
                                        tma_load_4d(sbase + (size_t)((q * Rload + r) * g_blocks + b) * GB, &g_map, co0 + b * 64,
                                                    x0 * p.Sg + (q & (p.Sg - 1)) + (Rload > 1 ? p.gcopy_dx[r] : 0),
                                                    y0 * p.Sg + (q >> (p.Sg >> 1)) + (Rload > 1 ? p.gcopy_dy[r] : 0), img, bar);
                            for (int b = 0; b < x_blocks; ++b)
                                for (int q = 0; q < x_npl; ++q)
                                    tma_load_4d(sbase + p.g_bytes + (size_t)b * XB + (size_t)q * XPL, &x_map, ci0 + b * 64,
                                                (x0 + p.sx_min) * p.Sx + (q & (p.Sx - 1)), (y0 + p.sy_min) * p.Sx + (q >> (p.Sx >> 1)), img, bar);
                        } else
                        if (g_tma && Rload > 1) {
                            for (int r = 0; r < Rload; ++r)          // copy r: the same box, (dy, dx) pixels further
                                tma_load_5d(sbase + (size_t)r * g_chunks * p.KS * 16, &g_map, 0, x0 + p.gcopy_dx[r], y0 + p.gcopy_dy[r],
                                            co0 >> 3, img, bar);
                        } else if (g_tma) {
                            for (int q = 0; q < g_planes; ++q)       // parity plane (py, px) of a stride-2 gradient: every 2nd pixel
                                tma_load_5d(sbase + (size_t)q * g_chunks * p.KS * 16, &g_map, 0, x0 * p.Sg + (q & (p.Sg - 1)),
                                            y0 * p.Sg + (q >> (p.Sg >> 1)), co0 >> 3, img, bar);
                        }
                        if (x_tma && !sw) tma_load_5d(sbase + p.g_bytes, &x_map, 0, x0 + p.sx_min, y0 + p.sy_min, ci0 >> 3, img, bar);
                    } else {
                        mbar_arrive(bar);
                    }
                }
                st.advance();
                if (p.dbg && !tl_mode && lane == 0) p.dbg[0 * dbg_ncta + dbg_cta] += tw1 - tw0;
                continue;
            }
            // ---- worker warps
            if (!(p.dbg_flags & 2)) {
                if (g_async) stage_tile_async<T>(tg_, sbase, GPS, img, y0, x0, co0, g_chunks, widx, nworkers, lane);
                if (x_async) stage_tile_async<T>(tx_, sbase + p.g_bytes, XPS, img, y0, x0, ci0, x_chunks, widx, nworkers, lane);
            }
            if (g_async || x_async) cp_async_mbar_arrive_noinc(&full[st.stage]);
            if (!(p.dbg_flags & 2)) {
                if (g_reg) stage_tile<T, SPLIT>(tg_, sbase, GPS, img, y0, x0, co0, g_chunks, widx, nworkers, lane);
                if (x_reg) stage_tile<T, SPLIT>(tx_, sbase + p.g_bytes, XPS, img, y0, x0, ci0, x_chunks, widx, nworkers, lane);
            }
            if (hop) mbar_wait(&tma_full[st.stage], st.phase, 0x520 + st.stage);
            if (x_bn && !(p.dbg_flags & 2)) {
                // fused BatchNorm + activation of the source tile, in place (zero padding stays zero).  A worker thread owns
                // ONE 8-channel chunk for the whole kernel (its scale / shift live in registers) and walks the tile's slots
                // with a fixed stride: row / column follow incrementally, no per-item division or coefficient load
                // (the per-item form cost 30 us of 106 on the layer1 shape, tools/check_wgrad_sw128.py bn=1 vs bn=0).
                if (bn_j >= 0) {
                    uint8_t* xbase = sbase + p.g_bytes;
                    const uint32_t xaddr = smem_u32(xbase);
                    int r = bn_r0, cx = bn_c0;
                    for (int sl = bn_s0; sl < XPS; sl += bn_tpc) {
                        const int iy = y0 + p.sy_min + r, ix = x0 + p.sx_min + cx;
                        if (iy >= 0 && iy < p.xH && ix >= 0 && ix < p.xW) {
                            uint32_t off;
                            if (sw) {
                                // swizzled blocks: unit (j & 7) of a slot sits at unit (j & 7) ^ (address bits [7,10)) of its row
                                const uint32_t row = (uint32_t)(bn_j >> 3) * XB + (uint32_t)sl * 128u;
                                off = row + ((((uint32_t)bn_j & 7u) ^ (((xaddr + row) >> 7) & 7u)) << 4);
                            } else {
                                off = (uint32_t)(bn_j * XPS + sl) << 4;
                            }
                            uint4* xp = reinterpret_cast<uint4*>(xbase + off);
                            *xp = bn_act8(*xp, bn_sc, bn_sh, p.ld_slope, p.ld_slope == 0.f);
                        }
                        r += bn_dr; cx += bn_dc;
                        if (cx >= p.Wl) { cx -= p.Wl; ++r; }
                    }
                }
            }
            if (g_tma) {
                // clear the junk columns [Wt, Wl) of every row and chunk plane of the gradient tile
                const int jc = p.Wl - p.Wt;
                const int items = p.Ht * jc * g_chunks * g_planes * Rload;   // planes, copies and chunk planes are all KS slots apart
                const FastDivS fd_jc((uint32_t)jc), fd_ht((uint32_t)p.Ht);
                for (int it = widx * 32 + lane; it < items; it += nworkers * 32) {
                    const int q = (int)fd_jc.div((uint32_t)it), c = it - q * jc;
                    const int j = (int)fd_ht.div((uint32_t)q), r = q - j * p.Ht;
                    // (swizzled blocks: the 8 units of a junk slot are cleared by the 8 values j & 7 -- any order)
                    const size_t off = sw ? (size_t)(j >> 3) * GB + (size_t)(r * p.Wl + p.Wt + c) * 128 + (size_t)(j & 7) * 16
                                          : ((size_t)(j * GPS + r * p.Wl + p.Wt + c) << 4);
                    *reinterpret_cast<uint4*>(sbase + off) = make_uint4(0, 0, 0, 0);
                }
            }
            if (warp_arrives) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[st.stage]);
            }
            st.advance();
            if (p.dbg && !tl_mode && !any_tma && warp == 4 && lane == 0) {
                const long long tw2 = clock64();
                const size_t ncta = (size_t)gridDim.x * gridDim.y * gridDim.z;
                const size_t cta = blockIdx.x + (size_t)gridDim.x * (blockIdx.y + (size_t)gridDim.y * blockIdx.z);
                p.dbg[0 * ncta + cta] += tw1 - tw0;
                p.dbg[1 * ncta + cta] += tw2 - tw1;
            }
        }
    } else if (warp == kWarpMma) {
        // ================= UMMA issuer: ONE elected thread runs the whole loop (waits, UMMAs, commits).  Inside the election
        // branch ptxas keeps descriptors in uniform registers and emits bare UTCHMMAs (rd_conv_fprop.cuh, fprop_issue); with
        // the whole warp walking the loops and a per-UMMA `if (leader)` the loop alone cost ~40 cycles per UMMA slot
        // (tools/bench_wgrad.py, "neither" mode: 67 of 79 us on the layer1 shape with loads AND UMMAs switched off).
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        if (elect_one_sync()) {
            PipeState st(p.NS);
            // tap-row folding (see rd_wgrad_params.fold_rows): a job = one row of taps x one 8-channel source chunk
            const bool fold = (SPLIT == 1) && p.fold_len > 0;
            const int njobs = fold ? p.fold_rows * x_chunks : T_n;
            const uint32_t idesc = make_idesc_bf16(128, fold ? 32 : p.Nc, 1, 1);
            const uint32_t g_sbo = (uint32_t)GPS * 16u, x_sbo = fold ? 16u : (uint32_t)XPS * 16u;
            const int KG = p.KS >> 4;
            const bool do_issue = !(p.dbg_flags & 1);
            const uint32_t kstep = sw ? 128u : 16u;            // 16 slots in 16-byte units: 128-byte rows when swizzled
            const uint32_t lo_b = (uint32_t)x_chunks * (uint32_t)XPS, lo_a = (uint32_t)g_chunks * (uint32_t)GPS;   // fp32 split: low parts
            uint32_t acc0 = 0u;                                // the first tile overwrites the accumulators
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const long long tm0 = p.dbg ? clock64() : 0;
                mbar_wait(&full[st.stage], st.phase, 0x510 + st.stage);
                const long long tm1 = p.dbg ? clock64() : 0;
                fence_proxy_async_smem();          // consumer-side: loaders' generic-proxy writes -> async proxy (UMMA)
                tc_fence_after();
                const uint32_t g_base = smem_u32(ring + (size_t)st.stage * p.stage_bytes);
                const uint64_t da0 = sw ? make_smem_desc_sw128(g_base, GB, 1024u, 0u) : make_smem_desc(g_base, 128, g_sbo);
                const uint64_t db0 = sw ? make_smem_desc_sw128(g_base + (uint32_t)p.g_bytes, XB, 1024u, 0u)
                                        : make_smem_desc(g_base + (uint32_t)p.g_bytes, 128, x_sbo);
                // tap-outer / k-group-inner: inside the inner loop the descriptors only advance by 16 slots, so one
                // UMMA costs two uniform adds; the tap offsets come from the (uniform) parameter bank once per tap
                if (do_issue)
                for (int jb = 0; jb < njobs; ++jb) {
                    const int tl = fold ? (R > 1 ? (jb >> 1) : (jb >> 1) * p.fold_len) : jb;   // x_chunks == 2 when folding
                    const uint32_t d = tmem_u + (uint32_t)(fold ? jb * 32 : jb * p.Nc);
                    uint64_t da = da0 + (uint32_t)p.taps[t0 + tl].g_off;
                    // swizzled blocks: a shift of s slots = s 128-byte rows.  The hardware swizzles on absolute shared-memory
                    // address bits, so a start address at any row phase of the 1024-byte pattern reads what TMA wrote
                    // (base_offset stays 0; setting it to the phase, tma_mode & 8, was measured WRONG)
                    uint64_t db = db0 + (sw ? (uint32_t)p.taps[t0 + tl].x_shift * 8u
                                            : (uint32_t)p.taps[t0 + tl].x_shift + (uint32_t)(fold ? (jb & 1) * XPS : 0));
                    if (sw && sw_bo) db |= (uint64_t)(((g_base + (uint32_t)p.g_bytes + (uint32_t)p.taps[t0 + tl].x_shift * 128u) >> 7) & 7u) << 49;
                    if (SPLIT == 1) {
                        umma_bf16_elected(d, da, db, idesc, acc0);
#pragma unroll 4
                        for (int kg = 1; kg < KG; ++kg) {
                            da += kstep;
                            db += kstep;
                            umma_bf16_elected(d, da, db, idesc, 1u);
                        }
                    } else {
                        for (int kg = 0; kg < KG; ++kg, da += 16u, db += 16u) {
                            umma_bf16_elected(d, da, db, idesc, kg == 0 ? acc0 : 1u);
                            umma_bf16_elected(d, da, db + lo_b, idesc, 1u);
                            umma_bf16_elected(d, da + lo_a, db, idesc, 1u);
                        }
                    }
                }
                umma_commit(&empty[st.stage]);
                st.advance();
                acc0 = 1u;
                if (p.dbg) {
                    const long long tm2 = clock64();
                    if (tl_mode) {
                        if (tile == (int)blockIdx.x) p.dbg[0 * dbg_ncta + dbg_cta] = tm1 - t_entry;
                        p.dbg[1 * dbg_ncta + dbg_cta] = tm2 - t_entry;
                    } else {
                        p.dbg[2 * dbg_ncta + dbg_cta] += tm1 - tm0;
                        p.dbg[3 * dbg_ncta + dbg_cta] += tm2 - tm1;
                    }
                }
            }
            umma_commit(tmem_full);
        }
        __syncwarp();
    }
    if (warp < 4 + kWgLoaderWarps && (warp < 4 || det_part == nullptr)) {
        // ================= epilogue: TMEM -> fp32 reductions into dw.  Warps 0-3 and, once their tile loop has ended, the
        // loader warps 4-15 (TMEM lane quarter = warp & 3): four groups that take every fourth accumulator (non-deterministic
        // mode; the ordered per-CTA copies of the deterministic mode are written by warps 0-3 alone, as before)
        const bool has_work = (int)blockIdx.x < ntiles;
        const bool det_mode = det_part != nullptr;
        const int eg = warp >> 2, neg = det_mode ? 1 : (4 + kWgLoaderWarps) / 4;
        mbar_wait(tmem_full, 0, 0x600);
        tc_fence_after();
        if (tl_mode && tid == 0) p.dbg[2 * dbg_ncta + dbg_cta] = clock64() - t_entry;
        const int arow = (warp & 3) * 32 + lane;          // accumulator row: copy * Mc + output channel within the block
        const int copy = R > 1 ? arow / p.Mc : 0;
        const int row = arow - copy * p.Mc;
        const bool valid = has_work && copy < R && row < p.Mc && (co0 + row) < p.Cout;
        const bool det = det_part != nullptr;
        float* const dwo = det ? det_part + (size_t)blockIdx.x * ((size_t)p.ntaps * p.Cout * p.Cin) : p.dw;
        if ((SPLIT == 1) && p.fold_len > 0) {
            // folded accumulators: job (row, chunk jx) holds columns [tap-in-row][8 channels of chunk jx]
            const int njobs = p.fold_rows * x_chunks;
            for (int jb = eg; jb < njobs; jb += neg) {
                const int jx = jb & 1;
                // first tap of the tap row this accumulator row belongs to (gradient copies: one job row covers several)
                const int tap0 = R > 1 ? (copy < R ? (int)p.job_tap[jb >> 1][copy] : -1) : (jb >> 1) * p.fold_len;
                for (int cc = 0; cc < 2; ++cc) {
                    float v[16];
                    tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(jb * 32 + cc * 16), v);
                    if (valid && tap0 >= 0) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int tx = cc * 2 + h;
                            if (tx < p.fold_len) {
                                float* out = dwo + ((size_t)(tap0 + tx) * p.Cout + (co0 + row)) * p.Cin + ci0 + jx * 8;
                                acc_out_v4(det, out, v[h * 8 + 0], v[h * 8 + 1], v[h * 8 + 2], v[h * 8 + 3]);
                                acc_out_v4(det, out + 4, v[h * 8 + 4], v[h * 8 + 5], v[h * 8 + 6], v[h * 8 + 7]);
                            }
                        }
                    }
                }
            }
        } else
        for (int tl = eg; tl < T_n; tl += neg) {
            const int tap = R > 1 ? (copy < R ? (int)p.job_tap[t0 + tl][copy] : -1) : t0 + tl;
            float* out = dwo + ((size_t)(tap < 0 ? 0 : tap) * p.Cout + (co0 + row)) * p.Cin + ci0;
            for (int cc = 0; cc < (p.Nc >> 4); ++cc) {
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(tl * p.Nc + cc * 16), v);
                if (valid && tap >= 0) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4) acc_out_v4(det, out + cc * 16 + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (tl_mode && tid == 0) p.dbg[3 * dbg_ncta + dbg_cta] = clock64() - t_entry;
    if (warp == kWarpMma) tmem_dealloc<512>(tmem_base);
}

}  // namespace rd
