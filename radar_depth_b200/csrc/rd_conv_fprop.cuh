// rd_conv_fprop.cuh -- implicit-GEMM convolution on tcgen05 (forward convs, data gradients, sub-pixel UpProj).
//
// One CTA computes, for a tile of Ht x Wt base output pixels of one image, P output phases x N output channels.
// The source tile (with halo) is staged ONCE per 16-channel block into shared memory in a "chunk-planar,
// pixel-linear" layout   [part][8-channel chunk][slot] x 16 B   (slot = row*Wl + col of the halo tile).
// In that layout every filter tap is a pure linear shift of the slot index, so the A operand of each
// tcgen05.mma is just a shared-memory matrix descriptor (SWIZZLE_NONE, K-major) whose start address is
// offset by the tap -- no im2col expansion, no per-tap re-gather: shared-memory fill traffic is 1x instead
// of taps x.  Accumulator rows are the tile's slots (including Wl-Wt junk columns that are never stored).
// Stride-2 convolutions stage the source as 4 parity planes (same trick per plane).  The 5x5 convs on the
// zero-stuffed Unpool output (models.py:13-27,191,198) and stride-2 data gradients run as 4 output phases,
// each with its own tap list, on the UN-stuffed source.
//
// Warp roles (512 threads; warps 14-15 are extra transform workers, see bn_epi2): warps 0-3 epilogue (TMEM lane quarter = warp id), warps 4-11 source-tile loaders
// (fused BatchNorm affine + ReLU/LeakyReLU + bf16 cast, optional hi/lo split for parity mode), warp 12 lane 0
// UMMA issuer, warp 13 lane 0 weight bulk-copy issuer.  Two mbarrier rings (source tile stages, weight
// stages) plus a TMEM full/empty pair; the kernel is persistent over tiles.
#pragma once
#include "rd_common.cuh"
#include "rd_tile.cuh"
#include "../../include/radar_depth_b200.h"

namespace rd {

constexpr int kFpropLoaderWarps = 8;
constexpr int kFpropThreads = (4 + kFpropLoaderWarps + 2 + 2) * 32;   // epilogue x4, loaders, UMMA issuer, weight copier, 2 extra transform warps (512 x 128 registers = the register file)
constexpr int kSmemHeader = 16384;      // barriers, tmem slot, stats, BN vectors
constexpr int kOffTmemSlot = 384;      // (the barrier block below occupies bytes [0, 352))
constexpr int kOffTapTable = 512;       // int[3][32]: a_shift, accumulator column, first-of-phase flag
constexpr int kOffStats = 1024;         // float[512]
constexpr int kOffEpScale = 3072;       // float[256]
constexpr int kOffEpShift = 4096;       // float[256]
constexpr int kOffLdScale = 5120;       // float[1024]
constexpr int kOffLdShift = 9216;       // float[1024]
constexpr int kMaxStages = 8;

// Butterfly sum over the 32 lanes of 16 per-lane values.  Afterwards lane L holds the full sum of element
// ((L>>4)&1)*8 + ((L>>3)&1)*4 + ((L>>2)&1)*2 + ((L>>1)&1)  (both lanes of an even/odd pair hold the same).
__device__ __forceinline__ float reduce16_lanes(float* v, int lane) {
    {
        const bool up = lane & 16;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float send = up ? v[i] : v[i + 8];
            float recv = __shfl_xor_sync(0xffffffffu, send, 16);
            v[i] = (up ? v[i + 8] : v[i]) + recv;
        }
    }
    {
        const bool up = lane & 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float send = up ? v[i] : v[i + 4];
            float recv = __shfl_xor_sync(0xffffffffu, send, 8);
            v[i] = (up ? v[i + 4] : v[i]) + recv;
        }
    }
    {
        const bool up = lane & 4;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float send = up ? v[i] : v[i + 2];
            float recv = __shfl_xor_sync(0xffffffffu, send, 4);
            v[i] = (up ? v[i + 2] : v[i]) + recv;
        }
    }
    {
        const bool up = lane & 2;
        float send = up ? v[0] : v[1];
        float recv = __shfl_xor_sync(0xffffffffu, send, 2);
        v[0] = (up ? v[1] : v[0]) + recv;
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    return v[0];
}
__device__ __forceinline__ int reduce16_owner_channel(int lane) {
    return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

// ---------------------------------------------------------------- UMMA issue loop
// Measured on B200 (tools/bench_fprop.py, dbg_flags): with loads, stores AND the tcgen05.mma instructions themselves
// switched off, the layer1-4 forward programs still took 80-90 % of their full time -- the single issuing warp, not the
// tensor pipe or shared memory, set the pace: ~60 SASS instructions per tap of four UMMAs (the issue predicate passed
// as an asm operand, so every UTCHMMA carried VOTEU / UMOV pairs; the B descriptor lived in vector registers and crossed
// to the uniform file with R2UR every tap; a 4-unrolled loop plus remainder branches), i.e. 69-108 cycles per UMMA
// against a floor of 48 (N = 64) / 64 (N = 128).  Now ONE elected thread runs the whole loop inside the election
// branch, the accumulator-block loop is unrolled at compile time (MBT = MB for MB <= 4, MBT = 0: run-time loop), and
// the taps of a channel block are numbered consecutively across the weight groups (no wrap-around bookkeeping):
// ~9 instructions per UMMA, the UTCHMMAs of a tap back to back.
// tcgen05.mma issued from inside an `if (elect_one_sync())` branch: in that form ptxas emits one bare UTCHMMA per call
// (operands in uniform registers, descriptor arithmetic as UIADD3.64); with the election passed as a predicate OPERAND
// it cannot prove that at most one lane is active and wraps every UTCHMMA in an ELECT / BRA.U.ANY replay loop.
__device__ __forceinline__ void umma_bf16_elected(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                  uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

template <int SPLIT, int MBT, int NT>
__device__ __forceinline__ void fprop_issue(const rd_conv_params& p, uint8_t* smem, uint32_t tmem_base, bool dbuf, int acc_cols,
                                            int ntiles, int ncblk, int lane, long long t_entry, bool tl_mode) {
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint64_t* in_full = bars;
    uint64_t* in_empty = bars + kMaxStages;
    uint64_t* w_full = bars + 2 * kMaxStages;
    uint64_t* w_empty = bars + 3 * kMaxStages;
    uint64_t* tmem_full = bars + 4 * kMaxStages;
    uint64_t* tmem_empty = bars + 4 * kMaxStages + 2;
    uint8_t* a_ring = smem + kSmemHeader;
    uint8_t* w_ring = a_ring + (size_t)p.IS * p.istage_bytes;
    constexpr int PARTS = (SPLIT == 3) ? 2 : 1;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    PipeState si(p.IS), sw(p.WS);
    const uint32_t N = (uint32_t)p.N;
    const uint32_t idesc = make_idesc_bf16(128, p.N, 0, 0);
    const uint32_t PS = (uint32_t)p.chunk_stride;                    // chunk stride in slots (= 16-byte units)
    const uint32_t tap_units = (uint32_t)PARTS * N * 2u;            // [part][2][N][8] bf16 per tap, in 16-byte units
    const bool do_issue = !(p.dbg_flags & 1);
    const int MB = MBT > 0 ? MBT : p.MB;
    const uint32_t mbn = (uint32_t)MB * N;
    // ONE elected thread runs the whole issue loop (waits, UMMAs, commits): inside the election branch ptxas keeps
    // descriptors and tap data in uniform registers and emits bare UTCHMMAs back to back.
    if (elect_one_sync()) {
        uint32_t tile_iter = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tile_iter) {
            const uint32_t ab = dbuf ? (tile_iter & 1u) : 0u;
            const uint32_t use = dbuf ? (tile_iter >> 1) : tile_iter;
            mbar_wait(&tmem_empty[ab], (use & 1u) ^ 1u, 0x300);
            tc_fence_after();
            const uint32_t d_tile = tmem_u + ab * (uint32_t)acc_cols;
            for (int c = 0; c < ncblk; ++c) {
                const long long t0_ = p.dbg ? clock64() : 0;
                mbar_wait(&in_full[si.stage], si.phase, 0x310 + si.stage);
                const long long t1_ = p.dbg ? clock64() : 0;
                long long tw_ = 0;
                fence_proxy_async_smem();      // consumer-side: loaders' generic-proxy writes -> async proxy (UMMA)
                tc_fence_after();
                const uint64_t da0 = make_smem_desc(smem_u32(a_ring + (size_t)si.stage * p.istage_bytes), PS * 16u, 128);
                const uint32_t keep = c == 0 ? 0u : 1u;                  // channel blocks after the first always accumulate
                int t = 0;
                if (NT > 0) {
                    // NT taps, one weight group, one output phase (every 3x3 / 1x1 forward conv and stride-1 data gradient with
                    // N <= 128): straight-line code, the tap shifts are immediate-offset reads of the parameter bank
                    const long long t2_ = p.dbg ? clock64() : 0;
                    mbar_wait(&w_full[sw.stage], sw.phase, 0x320 + sw.stage);
                    if (p.dbg) tw_ += clock64() - t2_;
                    tc_fence_after();
                    const uint64_t db0 = make_smem_desc(smem_u32(w_ring + (size_t)sw.stage * p.wstage_bytes), N * 16u, 128);
                    if (do_issue) {
#pragma unroll
                        for (int tt = 0; tt < NT; ++tt) {
                            const uint64_t da = da0 + (uint32_t)p.taps[tt].a_shift;
                            const uint64_t db = db0 + (uint32_t)tt * tap_units;
                            const uint32_t acc = tt == 0 ? keep : 1u;
#pragma unroll
                            for (int mb = 0; mb < MBT; ++mb) {
                                umma_bf16_elected(d_tile + (uint32_t)mb * N, da + (uint32_t)mb * 128u, db, idesc, acc);
                                if (SPLIT == 3) {
                                    umma_bf16_elected(d_tile + (uint32_t)mb * N, da + (uint32_t)mb * 128u, db + 2u * N, idesc, 1u);
                                    umma_bf16_elected(d_tile + (uint32_t)mb * N, da + (uint32_t)mb * 128u + 2u * PS, db, idesc, 1u);
                                }
                            }
                        }
                    }
                    umma_commit(&w_empty[sw.stage]);
                    sw.advance();
                } else
                for (int g = 0; g < p.ngroups; ++g) {
                    const long long t2_ = p.dbg ? clock64() : 0;
                    mbar_wait(&w_full[sw.stage], sw.phase, 0x320 + sw.stage);
                    if (p.dbg) tw_ += clock64() - t2_;
                    tc_fence_after();
                    const uint64_t db0 = make_smem_desc(smem_u32(w_ring + (size_t)sw.stage * p.wstage_bytes), N * 16u, 128);
                    const int gn = p.grp_n[g];
                    for (int tl = 0; tl < gn; ++tl, ++t) {
                        const uint64_t db = db0 + (uint32_t)tl * tap_units;
                        const uint64_t da = da0 + (uint32_t)p.taps[t].a_shift;
                        const uint32_t d = d_tile + (uint32_t)p.taps[t].phase * mbn;
                        const uint32_t acc = keep | ((uint32_t)p.taps[t].first ^ 1u);
                        if (do_issue) {
                            if (MBT > 0) {
#pragma unroll
                                for (int mb = 0; mb < MBT; ++mb) {
                                    umma_bf16_elected(d + (uint32_t)mb * N, da + (uint32_t)mb * 128u, db, idesc, acc);
                                    if (SPLIT == 3) {
                                        umma_bf16_elected(d + (uint32_t)mb * N, da + (uint32_t)mb * 128u, db + 2u * N, idesc, 1u);
                                        umma_bf16_elected(d + (uint32_t)mb * N, da + (uint32_t)mb * 128u + 2u * PS, db, idesc, 1u);
                                    }
                                }
                            } else {
                                uint64_t a = da;
                                uint32_t dd = d;
#pragma unroll 2
                                for (int mb = 0; mb < MB; ++mb, a += 128u, dd += N) {
                                    umma_bf16_elected(dd, a, db, idesc, acc);
                                    if (SPLIT == 3) {
                                        umma_bf16_elected(dd, a, db + 2u * N, idesc, 1u);
                                        umma_bf16_elected(dd, a + 2u * PS, db, idesc, 1u);
                                    }
                                }
                            }
                        }
                    }
                    umma_commit(&w_empty[sw.stage]);
                    sw.advance();
                }
                umma_commit(&in_empty[si.stage]);
                si.advance();
                if (p.dbg) {
                    const size_t ncta = (size_t)gridDim.x * gridDim.y, cta = blockIdx.x + (size_t)gridDim.x * blockIdx.y;
                    if (tl_mode) { if (tile_iter == 0 && c == 0) p.dbg[2 * ncta + cta] = t1_ - t_entry; p.dbg[3 * ncta + cta] = clock64() - t_entry; }
                    else { p.dbg[2 * ncta + cta] += (t1_ - t0_) + tw_; p.dbg[3 * ncta + cta] += clock64() - t1_ - tw_; }
                }
            }
            umma_commit(&tmem_full[ab]);
        }
    }
}

template <typename T, int SPLIT>
__global__ void __launch_bounds__(kFpropThreads, 1) conv_fprop_kernel(const __grid_constant__ rd_conv_params p,
                                                                      const __grid_constant__ CUtensorMap src_map, const int use_tma,
                                                                      float* const det_part) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint64_t* in_full = bars;                       // [kMaxStages]
    uint64_t* in_empty = bars + kMaxStages;
    uint64_t* w_full = bars + 2 * kMaxStages;
    uint64_t* w_empty = bars + 3 * kMaxStages;
    uint64_t* tmem_full = bars + 4 * kMaxStages;        // [2]
    uint64_t* tmem_empty = bars + 4 * kMaxStages + 2;   // [2]
    uint64_t* tma_full = bars + 4 * kMaxStages + 4;     // [kMaxStages] TMA landing barriers of the transform-in-place mode
    int* tap_a = reinterpret_cast<int*>(smem + kOffTapTable);
    int* tap_d = tap_a + 32;
    int* tap_f = tap_a + 64;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffTmemSlot);
    float* stats_s = reinterpret_cast<float*>(smem + kOffStats);
    float* ep_sc = reinterpret_cast<float*>(smem + kOffEpScale);
    float* ep_sh = reinterpret_cast<float*>(smem + kOffEpShift);
    float* ld_sc = reinterpret_cast<float*>(smem + kOffLdScale);
    float* ld_sh = reinterpret_cast<float*>(smem + kOffLdShift);
    uint8_t* a_ring = smem + kSmemHeader;
    uint8_t* w_ring = a_ring + (size_t)p.IS * p.istage_bytes;
    // Deterministic mode (det_part != nullptr, rd_set_deterministic): every epilogue warp sums its statistics into a
    // private [2][N] array behind the rings (plain adds, fixed tile order), the CTA adds those arrays in warp order, stores
    // its partial to det_part[cta][2][N], and the last CTA adds the partials of all CTAs in index order -- no
    // floating-point atomics anywhere, so two runs give bit-identical BatchNorm statistics.
    float* det_s = reinterpret_cast<float*>(w_ring + (size_t)p.WS * p.wstage_bytes);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int kWarpMma = 4 + kFpropLoaderWarps, kWarpW = kWarpMma + 1;
    // dbg_flags & 8: the six counters become a timeline (cycles since kernel entry): prologue done, loaders done,
    // first operand stage consumed, last UMMA committed, first accumulator ready, epilogue done
    const bool tl_mode = p.dbg && (p.dbg_flags & 8);
    const long long t_entry = p.dbg ? clock64() : 0;
    // dbg_flags & 16 (with 8): rows 0 / 1 of the timeline hold the GLOBAL timer (ns) at CTA entry / exit instead, so that the
    // host can separate the kernel's own duration from launch gaps (max exit of launch i -> min entry of launch i+1)
    const bool gt_mode = tl_mode && (p.dbg_flags & 16);
    if (gt_mode && tid == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        p.dbg[0 * ((size_t)gridDim.x * gridDim.y) + blockIdx.x + (size_t)gridDim.x * blockIdx.y] = (long long)gt;
    }
    const int nb = blockIdx.y;                       // N block
    constexpr int PARTS = (SPLIT == 3) ? 2 : 1;
    const int ncblk = p.Cin >> 4;
    const int tiles_per_img = p.tiles_y * p.tiles_x;
    const int ntiles = tiles_per_img * p.B;

    // ---- one-time setup
    // raw bf16 source tiles are staged with cp.async (one arrival per loader thread), others through registers
    const bool src_async = (SPLIT == 1) && (sizeof(T) == 2) && (p.ld_scale == nullptr);
    // raw stride-1 tiles: ONE elected thread issues a TMA box load per stage (use_tma is decided by the launcher, which
    // also encodes src_map and sets p.chunk_stride = plane_rows * Wl, the chunk pitch TMA writes)
    const bool src_tma = src_async && use_tma;
    // bf16 stride-1 tiles with a fused BatchNorm+activation: the raw box lands by TMA on tma_full[stage], seven worker
    // warps apply the transform IN PLACE in shared memory (no global latency on their path) and arrive on in_full[stage]
    const bool src_tma_bn = (SPLIT == 1) && (sizeof(T) == 2) && (p.ld_scale != nullptr) && use_tma;
    // Transform mode with >= 4 taps per stage: warps 8-11 form the second epilogue group as in the raw-TMA mode and the
    // transform runs on warps 5-7 + 14-15 (five workers; three were measured too few: 74.5 vs 67.6 us on layer1).  With four
    // epilogue warps that program spent 74.9 kcycles per CTA in its epilogue against 58.7 in its UMMAs (tools/bench_fprop.py
    // l1 bn).  1-tap programs keep seven workers (5-11).
    const bool bn_epi2 = src_tma_bn && p.ntaps >= 4 && !(p.dbg_flags & 64);
    const int bn_workers = bn_epi2 ? 5 : kFpropLoaderWarps - 1;
    // cp.async / register staging (stride-2 parity planes, fp32 parity mode): six loader warps (4-7, 14-15) and the second
    // epilogue group instead of eight loaders and one group
    const bool ld_epi2 = !src_tma && !src_tma_bn && !(p.dbg_flags & 64);
    const bool epi3_pre = src_tma && det_part == nullptr && !(p.dbg_flags & 64);   // three epilogue groups in the raw-TMA mode (see below;
                                                                                       // the deterministic mode's per-warp arrays are sized for two)
    const int ld_warps = ld_epi2 ? 6 : kFpropLoaderWarps;
    if (tid == 0) {
        for (int i = 0; i < p.IS; ++i) {
            mbar_init(&in_full[i], src_tma ? 1 : (src_tma_bn ? bn_workers : (src_async ? 32 * ld_warps : ld_warps)));
            mbar_init(&in_empty[i], 1);
            mbar_init(&tma_full[i], 1);
        }
        for (int i = 0; i < p.WS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], epi3_pre ? 12 : ((src_tma || bn_epi2 || ld_epi2) ? 8 : 4)); }
        fence_mbar_init();
    }
    if (tid < p.ntaps) {
        tap_a[tid] = p.taps[tid].a_shift;
        tap_d[tid] = p.taps[tid].phase * p.MB * p.N;
        tap_f[tid] = p.taps[tid].first;
    }
    // two accumulator sets when they fit: the epilogue of tile i overlaps the MMAs of tile i+1
    const int acc_cols = p.P * p.MB * p.N;
    const bool dbuf = 2 * acc_cols <= 512;
    for (int i = tid; i < 512; i += kFpropThreads) stats_s[i] = 0.f;
    zero_smem(a_ring, (size_t)p.IS * p.istage_bytes, tid, kFpropThreads);   // tile tails / junk rows stay zero forever
    fence_proxy_async_smem();
    if (warp == kWarpW) tmem_alloc<512>(tmem_slot);
    // Everything above depends on the kernel parameters only and overlaps the tail of the preceding kernel; from here on
    // the kernel touches global memory (the BatchNorm vectors below were written by the producer's rd_bn_tail).
    pdl_enter();
    if (p.ld_scale) {
        for (int i = tid; i < p.Cin; i += kFpropThreads) { ld_sc[i] = p.ld_scale[i]; ld_sh[i] = p.ld_shift[i]; }
    }
    if (p.epi != 0) {
        for (int i = tid; i < p.N; i += kFpropThreads) { ep_sc[i] = p.ep_scale[nb * p.N + i]; ep_sh[i] = p.ep_shift[nb * p.N + i]; }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // With raw TMA staging seven of the eight loader warps have nothing to load: warps 8-11 (TMEM lane quarters 0-3 again)
    // become a second epilogue group that takes the odd 16-column chunks.
    const bool epi2 = src_tma || bn_epi2 || ld_epi2;
    // Raw-TMA mode: the box loads are issued by warp 14 and warps 4-7 (TMEM lane quarters 0-3 once more) are a THIRD epilogue
    // group: the 16-column chunks are dealt to three groups (5 chunks of the stem: 2/2/1 instead of 3/2; 8 chunks: 3/3/2).
    const bool epi3 = epi3_pre;
    const int tma_warp = epi3 ? 14 : 4;
    if ((warp >= 4 && warp < kWarpMma && !(epi2 && warp >= 8 && warp < 12) && !(epi3 && warp < 8)) ||
        ((bn_epi2 || ld_epi2 || epi3) && warp == 14) || ((bn_epi2 || ld_epi2) && warp == 15)) {
        // ================= source tile loaders =================
        PipeState st(p.IS);
        TileSrc ts;
        ts.ptr = p.src.ptr; ts.pitch = p.src.pitch; ts.coff = p.src.coff; ts.H = p.srcH; ts.W = p.srcW; ts.S = p.S; ts.nplanes = p.src_planes;
        ts.plane_slots = p.plane_slots; ts.plane_rows = p.plane_rows; ts.Wl = p.Wl; ts.oy0 = p.sy_min; ts.ox0 = p.sx_min;
        ts.vrows = p.plane_rows; ts.vcols = p.Wl;
        ts.sc = p.ld_scale ? ld_sc : nullptr; ts.sh = ld_sh; ts.slope = p.ld_slope;
        ts.prepare();
        if (tl_mode && !gt_mode && warp == tma_warp && lane == 0) p.dbg[0 * ((size_t)gridDim.x * gridDim.y) + blockIdx.x + (size_t)gridDim.x * blockIdx.y] = clock64() - t_entry;
        // Raw tiles: fire-and-forget cp.async, each loader thread arrives through cp.async.mbarrier.arrive.noinc
        // when its copies have landed (see rd_conv_wgrad.cuh); transformed tiles: registers + one arrival per warp.
        const int ld_idx = warp >= 14 ? warp - 10 : warp - 4;          // loader index in the cp.async / register modes
        if (src_tma && !(warp == tma_warp && lane == 0)) { /* one thread drives the TMA unit; the other loader threads have nothing to do */ }
        else if (src_tma_bn && warp != 4) {
            // ---- transform workers (warps 5..11)
            // A worker thread owns ONE of the two 8-channel chunks of every stage and walks the tile's slots with a fixed
            // stride: the chunk's scale / shift are read once per stage (not per item), row / column follow incrementally
            const int t = (warp >= 14 ? warp - 11 : warp - 5) * 32 + lane, nthr = bn_workers * 32;      // workers 5-7, 14-15 or 5-11
            const int cs = p.chunk_stride;                 // = plane_rows * Wl in this mode
            const int nslots = p.plane_rows * p.Wl;
            const int j = t & 1, s0 = t >> 1, tpc = nthr >> 1;
            const int r0 = s0 / p.Wl, c0 = s0 - r0 * p.Wl, dr = tpc / p.Wl, dc = tpc - dr * p.Wl;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int img = tile / tiles_per_img;
                const int trem = tile - img * tiles_per_img;
                const int ty = trem / p.tiles_x, tx = trem - ty * p.tiles_x;
                const int yb = ty * p.Ht + p.sy_min, xb = tx * p.Wt + p.sx_min;
                (void)img;
                for (int c = 0; c < ncblk; ++c) {
                    float scv[8], shv[8];
                    lds8(ld_sc + c * 16 + j * 8, scv);
                    lds8(ld_sh + c * 16 + j * 8, shv);
                    mbar_wait(&tma_full[st.stage], st.phase, 0x120 + st.stage);
                    uint4* sbase = reinterpret_cast<uint4*>(a_ring + (size_t)st.stage * p.istage_bytes) + j * cs;
                    if (!(p.dbg_flags & 2)) {
                        int r = r0, cx = c0;
                        for (int sl = s0; sl < nslots; sl += tpc) {
                            const int iy = yb + r, ix = xb + cx;
                            if (iy >= 0 && iy < p.srcH && ix >= 0 && ix < p.srcW) {          // zero padding stays zero
                                sbase[sl] = bn_act8(sbase[sl], scv, shv, p.ld_slope, p.ld_slope == 0.f);
                            }
                            r += dr; cx += dc;
                            if (cx >= p.Wl) { cx -= p.Wl; ++r; }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&in_full[st.stage]);
                    st.advance();
                }
            }
        }
        else if (src_tma_bn && lane != 0) { /* warp 4: only lane 0 issues */ }
        else
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int img = tile / tiles_per_img;
            const int trem = tile - img * tiles_per_img;
            const int ty = trem / p.tiles_x, tx = trem - ty * p.tiles_x;
            const int y0 = ty * p.Ht, x0 = tx * p.Wt;
            for (int c = 0; c < ncblk; ++c) {
                const long long t0_ = p.dbg ? clock64() : 0;
                mbar_wait(&in_empty[st.stage], st.phase ^ 1, 0x100 + st.stage);
                if (src_tma || src_tma_bn) {
                    uint64_t* bar = src_tma ? &in_full[st.stage] : &tma_full[st.stage];
                    if (!(p.dbg_flags & 2)) {
                        mbar_arrive_expect_tx(bar, (uint32_t)(p.plane_rows * p.Wl * 32));
                        tma_load_5d(a_ring + (size_t)st.stage * p.istage_bytes, &src_map, 0, x0 + p.sx_min, y0 + p.sy_min, c * 2, img, bar);
                    } else {
                        mbar_arrive(bar);
                    }
                    st.advance();
                    if (p.dbg) {
                        const size_t ncta = (size_t)gridDim.x * gridDim.y, cta = blockIdx.x + (size_t)gridDim.x * blockIdx.y;
                        if (tl_mode) { if (!gt_mode) p.dbg[1 * ncta + cta] = clock64() - t_entry; }
                        else { const long long t1_ = clock64(); p.dbg[0 * ncta + cta] += t1_ - t0_; }
                    }
                    continue;
                }
                const long long t1_ = p.dbg ? clock64() : 0;
                uint8_t* sbase = a_ring + (size_t)st.stage * p.istage_bytes;
                if (p.dbg_flags & 2) {
                    if (src_async) cp_async_mbar_arrive_noinc(&in_full[st.stage]);
                    else { __syncwarp(); if (lane == 0) mbar_arrive(&in_full[st.stage]); }
                } else
                if (src_async) {
                    stage_tile_async<T>(ts, sbase, p.chunk_stride, img, y0, x0, c * 16, 2, ld_idx, ld_warps, lane);
                    cp_async_mbar_arrive_noinc(&in_full[st.stage]);
                } else {
                    stage_tile<T, SPLIT>(ts, sbase, p.chunk_stride, img, y0, x0, c * 16, 2, ld_idx, ld_warps, lane);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&in_full[st.stage]);
                }
                st.advance();
                if (p.dbg && warp == 4 && lane == 0) {
                    const size_t ncta = (size_t)gridDim.x * gridDim.y, cta = blockIdx.x + (size_t)gridDim.x * blockIdx.y;
                    if (tl_mode) { if (!gt_mode) p.dbg[1 * ncta + cta] = clock64() - t_entry; }
                    else { p.dbg[0 * ncta + cta] += t1_ - t0_; p.dbg[1 * ncta + cta] += clock64() - t1_; }
                }
            }
        }
    } else if (warp == kWarpW) {
        // ================= weight bulk-copy issuer =================
        if (lane == 0) {
            PipeState st(p.WS);
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpk);
            const size_t tap_bytes = (size_t)PARTS * p.N * 32;                      // [part][2][N][8] bf16
            const size_t cblk_bytes = tap_bytes * p.ntaps;
            const size_t nblk_bytes = cblk_bytes * ncblk;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int c = 0; c < ncblk; ++c) {
                    for (int g = 0; g < p.ngroups; ++g) {
                        mbar_wait(&w_empty[st.stage], st.phase ^ 1, 0x200 + st.stage);
                        const uint32_t bytes = (uint32_t)(tap_bytes * p.grp_n[g]);
                        mbar_arrive_expect_tx(&w_full[st.stage], bytes);
                        bulk_g2s(w_ring + (size_t)st.stage * p.wstage_bytes,
                                 wsrc + nb * nblk_bytes + c * cblk_bytes + p.grp_first[g] * tap_bytes, bytes,
                                 &w_full[st.stage]);
                        st.advance();
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == kWarpMma) {
        // ================= UMMA issuer: the whole warp walks the loops (uniform values), one elected lane issues
        {
            // straight-line tap loops for the two common programs (first tap of phase 0 clears, one weight group)
            const bool simple = p.P == 1 && p.ngroups == 1 && p.MB <= 4 && p.taps[0].first == 1;
            const int nt = simple ? (p.ntaps == 9 ? 9 : (p.ntaps == 1 ? 1 : 0)) : 0;
#define RD_ISSUE(MBV, NTV) fprop_issue<SPLIT, MBV, NTV>(p, smem, tmem_base, dbuf, acc_cols, ntiles, ncblk, lane, t_entry, tl_mode)
            if (nt == 9) {
                switch (p.MB) { case 1: RD_ISSUE(1, 9); break; case 2: RD_ISSUE(2, 9); break; case 3: RD_ISSUE(3, 9); break; default: RD_ISSUE(4, 9); break; }
            } else if (nt == 1) {
                switch (p.MB) { case 1: RD_ISSUE(1, 1); break; case 2: RD_ISSUE(2, 1); break; case 3: RD_ISSUE(3, 1); break; default: RD_ISSUE(4, 1); break; }
            } else {
                switch (p.MB) {
                    case 1: RD_ISSUE(1, 0); break;
                    case 2: RD_ISSUE(2, 0); break;
                    case 3: RD_ISSUE(3, 0); break;
                    case 4: RD_ISSUE(4, 0); break;
                    default: RD_ISSUE(0, 0); break;
                }
            }
#undef RD_ISSUE
        }
        __syncwarp();
    } else if (warp >= 14) {
        // (extra transform warps: idle outside the five-worker transform mode)
    } else {
        // ================= epilogue (warps 0-3, and warps 8-11 in the raw-TMA and five-worker transform modes) =================
        const int wq = warp & 3, eg = warp < 4 ? 0 : (warp >= 8 ? 1 : 2), neg = epi3 ? 3 : (epi2 ? 2 : 1);
        T* dst = reinterpret_cast<T*>(p.dst.ptr);
        const T* addend = reinterpret_cast<const T*>(p.addend.ptr);
        const T* zsrc = reinterpret_cast<const T*>(p.zsrc.ptr);
        const bool want_stats = p.stats != nullptr;
        const FastDivS fd_wl((uint32_t)p.Wl);
        // 16 channels of a pixel move as one 256-bit request when every address of the view is 32-byte aligned
        const int al = 32 / (int)sizeof(T);
        auto wide_ok = [&](const rd_view& v) { return v.ptr == nullptr || ((((uintptr_t)v.ptr) & 31) == 0 && v.pitch % al == 0 && v.coff % al == 0); };
        const bool wide = wide_ok(p.dst) && wide_ok(p.addend) && wide_ok(p.zsrc) && !(p.dbg_flags & 32);
        float* det_mine = det_part ? det_s + (size_t)(eg * 4 + wq) * 2 * p.N : nullptr;
        if (det_mine) {
            for (int i = lane; i < 2 * p.N; i += 32) det_mine[i] = 0.f;
            __syncwarp();
        }
        uint32_t tile_iter = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tile_iter) {
            const int img = tile / tiles_per_img;
            const int trem = tile - img * tiles_per_img;
            const int ty = trem / p.tiles_x, tx = trem - ty * p.tiles_x;
            const int y0 = ty * p.Ht, x0 = tx * p.Wt;
            const uint32_t ab = dbuf ? (tile_iter & 1u) : 0u;
            const uint32_t use = dbuf ? (tile_iter >> 1) : tile_iter;
            const long long t0_ = p.dbg ? clock64() : 0;
            mbar_wait(&tmem_full[ab], use & 1u, 0x400);
            const long long t1_ = p.dbg ? clock64() : 0;
            tc_fence_after();
            const uint32_t t_tile = tmem_base + ab * (uint32_t)acc_cols + ((uint32_t)(wq * 32) << 16);
            // two groups: split the 16-column chunks between them, or (single-chunk layers) the 128-row blocks
            const bool split_cc = (p.N >> 4) >= 2;
            const int cc0 = split_cc ? eg : 0, ccs = split_cc ? neg : 1;
            const int mb0 = split_cc ? 0 : eg, mbs = split_cc ? 1 : neg;
            for (int cc = cc0; cc < (p.N >> 4); cc += ccs) {
                float s1[16], s2[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
                for (int ph = 0; ph < p.P; ++ph) {
                    for (int mb = mb0; mb < p.MB; mb += mbs) {
                        const int m = mb * 128 + wq * 32 + lane;
                        const int ly = (int)fd_wl.div((uint32_t)m), lx = m - ly * p.Wl;
                        const int oy = y0 + ly, ox = x0 + lx;
                        const int fy = oy * p.OS + p.phase_y[ph], fx = ox * p.OS + p.phase_x[ph];
                        const bool valid = ly < p.Ht && lx < p.Wt && oy < p.Hb && ox < p.Wb && fy < p.dstH && fx < p.dstW;
                        float v[16];
                        tmem_ld16(t_tile + (uint32_t)((ph * p.MB + mb) * p.N + cc * 16), v);
                        if (valid && !(p.dbg_flags & 4)) {
                            const size_t pix = ((size_t)img * p.dstH + fy) * p.dstW + fx;
                            const int ch = nb * p.N + cc * 16;
                            if (p.epi == 2) {        // inference: the following BatchNorm is a fixed affine map
#pragma unroll
                                for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], ep_sc[cc * 16 + i], ep_sh[cc * 16 + i]);
                            }
                            if (addend) {
                                float a[16];
                                const T* ap = addend + pix * p.addend.pitch + p.addend.coff + ch;
                                load16(ap, a, wide);
#pragma unroll
                                for (int i = 0; i < 16; ++i) v[i] += a[i];
                            }
                            if (p.epi == 2) {
                                const float sl = ch < p.ep_split ? p.ep_slope : p.ep_slope_b;
#pragma unroll
                                for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * sl;
                            } else
                            if (p.epi == 1) {
                                float z[16];
                                const T* zp = zsrc + pix * p.zsrc.pitch + p.zsrc.coff + ch;
                                load16(zp, z, wide);
#pragma unroll
                                for (int i = 0; i < 16; ++i) {
                                    const float y = fmaf(z[i], ep_sc[cc * 16 + i], ep_sh[cc * 16 + i]);
                                    const float g = y > 0.f ? v[i] : v[i] * p.ep_slope;
                                    v[i] = g;
                                    s1[i] += g;
                                    s2[i] += g * z[i];
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < 16; ++i) { s1[i] += v[i]; s2[i] += v[i] * v[i]; }
                            }
                            T* dp = dst + pix * p.dst.pitch + p.dst.coff + ch;
                            store16(dp, v, wide);
                        }
                    }
                }
                if (want_stats) {
                    const float r1 = reduce16_lanes(s1, lane);
                    const float r2 = reduce16_lanes(s2, lane);
                    if ((lane & 1) == 0) {
                        const int chn = cc * 16 + reduce16_owner_channel(lane);
                        if (det_mine) {
                            det_mine[chn] += r1;                 // this lane pair is the only writer of the entry
                            det_mine[p.N + chn] += r2;
                        } else {
                            atomicAdd(&stats_s[chn], r1);
                            atomicAdd(&stats_s[256 + chn], r2);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[ab]);
            if (p.dbg && warp == 0 && lane == 0) {
                const size_t ncta = (size_t)gridDim.x * gridDim.y, cta = blockIdx.x + (size_t)gridDim.x * blockIdx.y;
                if (tl_mode) { if (tile_iter == 0) p.dbg[4 * ncta + cta] = t1_ - t_entry; p.dbg[5 * ncta + cta] = clock64() - t_entry; }
                else { p.dbg[4 * ncta + cta] += t1_ - t0_; p.dbg[5 * ncta + cta] += clock64() - t1_; }
            }
        }
        if (want_stats) {
            if (neg == 3) asm volatile("bar.sync 1, 384;\n" ::: "memory");      // all epilogue groups have added their partials
            else if (neg == 2) asm volatile("bar.sync 1, 256;\n" ::: "memory");
            else asm volatile("bar.sync 1, 128;\n" ::: "memory");
            if (eg == 0) {
                if (det_part) {
                    float* mypart = det_part + (size_t)(blockIdx.x + gridDim.x * blockIdx.y) * 2 * p.N;
                    for (int i = tid; i < 2 * p.N; i += 128) {
                        float s = 0.f;
                        for (int w = 0; w < 4 * neg; ++w) s += det_s[(size_t)w * 2 * p.N + i];
                        mypart[i] = s;
                    }
                } else {
                    double* sdst = p.stats + (size_t)((blockIdx.x + gridDim.x * blockIdx.y) % (unsigned)tail_slots(p.tail)) * p.tail.slot_stride;
                    for (int i = tid; i < p.N; i += 128) {
                        atomicAdd(&sdst[nb * p.N + i], (double)stats_s[i]);
                        atomicAdd(&sdst[p.stats_stride + nb * p.N + i], (double)stats_s[256 + i]);
                    }
                }
                if (p.tail.counter) {
                    // last CTA to get here finalises the BatchNorm(s) fed by these statistics (rd_bn_tail)
                    __threadfence();
                    asm volatile("bar.sync 2, 128;\n" ::: "memory");
                    if (tid == 0) tmem_slot[1] = (atomicAdd(p.tail.counter, 1u) == gridDim.x * gridDim.y - 1u) ? 1u : 0u;
                    asm volatile("bar.sync 2, 128;\n" ::: "memory");
                    if (tmem_slot[1]) {
                        __threadfence();
                        if (det_part) {
                            // channel ch belongs to N block ch / N: add the partials of that block's CTAs in x order
                            for (int ch = tid; ch < (int)gridDim.y * p.N; ch += 128) {
                                const int y = ch / p.N, i = ch - y * p.N;
                                double a = 0.0, b = 0.0;
                                for (unsigned x = 0; x < gridDim.x; ++x) {
                                    const float* q = det_part + (size_t)(x + gridDim.x * y) * 2 * p.N;
                                    a += (double)__ldcg(q + i);
                                    b += (double)__ldcg(q + p.N + i);
                                }
                                p.stats[ch] = a;
                                p.stats[p.stats_stride + ch] = b;
                            }
                            __threadfence();
                            asm volatile("bar.sync 2, 128;\n" ::: "memory");
                        }
                        bn_tail_run(p.tail, tid, 128);
                    }
                    asm volatile("bar.sync 2, 128;\n" ::: "memory");
                }
            }
        }
    }

    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == kWarpW) tmem_dealloc<512>(tmem_base);
    if (gt_mode && tid == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        p.dbg[1 * ((size_t)gridDim.x * gridDim.y) + blockIdx.x + (size_t)gridDim.x * blockIdx.y] = (long long)gt;
        p.dbg[2 * ((size_t)gridDim.x * gridDim.y) + blockIdx.x + (size_t)gridDim.x * blockIdx.y] = clock64() - t_entry;
    }
}

}  // namespace rd
