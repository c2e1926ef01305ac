// rd_common.cuh -- sm_100a device primitives shared by the radar_depth_b200 kernels.
//
// Thin inline-PTX wrappers for mbarrier, bulk async copy (UBLKCP), tcgen05 (TMEM alloc, UMMA issue,
// commit, TMEM load) and the shared-memory matrix descriptor used by tcgen05.mma.  All waits are
// bounded: a barrier that does not flip within RD_WAIT_SPINS polls records a code in a global
// error word and traps, so a protocol bug surfaces as a CUDA error instead of hanging the GPU.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#ifndef RD_WAIT_SPINS
#define RD_WAIT_SPINS (1u << 24)
#endif

namespace rd {

typedef __nv_bfloat16 bf16;

__device__ unsigned int g_rd_device_error = 0;   // last device-side protocol error (0 = none)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- programmatic dependent launch
// With RD_PDL=1 every kernel of the library is launched with cudaLaunchAttributeProgrammaticStreamSerialization (rd_launch
// in radar_depth_b200.cu): the next kernel of the stream may be scheduled onto SMs as soon as this kernel's CTAs have
// all passed pdl_enter() (and SM resources free up), so its launch latency and its prologue (barrier init, shared-memory
// zero fill, TMEM allocation) overlap the tail of this kernel.  pdl_enter() = wait until the preceding kernel has COMPLETED
// and its memory is visible (griddepcontrol.wait), then allow the next kernel to start launching: at most one kernel is
// ever resident ahead of the running one.  Nothing produced by an earlier kernel may be read, and nothing it might still
// read may be written, before pdl_enter().  Without the launch attribute both instructions are no-ops.
// Measured on B200 (round 2): eager back-to-back launches have 3.2 us between the last CTA's exit and the next kernel's
// first CTA (tools/launch_gap.py) and PDL makes the entries overlap by 8 us, but the early CTAs only wait: 59.3 vs 56.9 us
// per launch; under CUDA-graph replay (gap 1.7 us) the training step is 9.234 (on) vs 9.238 ms (off), multistage 12.96 vs
// 12.84.  Hence OFF by default; the hooks stay because they cost nothing.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait.  `code` identifies the waiting role in g_rd_device_error if the wait times out.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t code) {
#pragma unroll 1
    for (uint32_t spin = 0; spin < RD_WAIT_SPINS; ++spin) {
        if (mbar_try_wait(bar, parity)) return;
    }
    atomicExch(&g_rd_device_error, code);
    __threadfence_system();
    asm volatile("trap;\n");
}

// generic-proxy writes (st.shared) -> visible to the async proxy (tcgen05.mma / bulk copies)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}

// ---------------------------------------------------------------- cp.async (LDGSTS) 16-byte copies with zero fill
// src_bytes = 16 copies, src_bytes = 0 writes 16 zero bytes (the source address must still be valid).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(src_bytes)
                 : "memory");
}
// 8-byte variant (.ca: the 8-byte form has no .cg), same zero-fill rule
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
// Arrive on `bar` once every cp.async previously issued by this thread has landed (no wait, no pending-count increment:
// the barrier's expected count must include one arrival per issuing thread).
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------- bulk async copy global -> shared
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---------------------------------------------------------------- TMA tensor load (5-D tile) global -> shared
// Box of the NHWC activation seen as (8 channels, W, H, C/8, B): lands as [chunk][row][col] x 16 B, out-of-image
// coordinates are zero-filled by the hardware (= the convolution's zero padding).
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2,%3,%4,%5,%6}], [%7];\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2,%3,%4,%5}], [%6];\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}

// eight consecutive floats of a 16-byte aligned shared-memory vector as two LDS.128
__device__ __forceinline__ void lds8(const float* p, float* v) {
    const float4 t0 = *reinterpret_cast<const float4*>(p), t1 = *reinterpret_cast<const float4*>(p + 4);
    v[0] = t0.x; v[1] = t0.y; v[2] = t0.z; v[3] = t0.w; v[4] = t1.x; v[5] = t1.y; v[6] = t1.z; v[7] = t1.w;
}

// ---------------------------------------------------------------- tcgen05 / TMEM
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_slot)),
                 "n"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {      // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// tcgen05.commit: arrive on `bar` when all previously issued UMMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
                 : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, single CTA.  Issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Same, with the issue predicate INSIDE the asm block: the C++ side executes it unconditionally (all operands stay in
// uniform registers, no predicated operand copies) and only the tcgen05.mma itself is guarded.
__device__ __forceinline__ void umma_bf16_if(uint32_t issue, uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(issue)
        : "memory");
}

// Instruction descriptor for kind::f16, bf16 operands, fp32 accumulate (cute::UMMA::InstrDescriptor):
//   [4,6) c_format=1 (F32)  [7,10) a_format=1 (BF16)  [10,13) b_format=1 (BF16)
//   [15] a_major  [16] b_major (0 = K-major, 1 = MN-major)  [17,23) N>>3  [24,29) M>>4
__host__ __device__ __forceinline__ uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Shared-memory matrix descriptor, SWIZZLE_NONE ("interleave") canonical layouts, in 16-byte units:
//   K-major : ((8,m),2) : ((1,SBO),LBO)   8 rows 16 B apart, row groups SBO apart, the two K halves LBO apart
//   MN-major: ((1,n),(8,2)) : ((-,SBO),(1,LBO))   8-element MN chunks SBO apart, 8 k-rows 16 B apart, k groups LBO apart
// bits [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1 (sm_100), [61,64) layout=0.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// SWIZZLE_128B MN-major canonical layout (cute: Swizzle<3,4,3> o ((8,8,m),(8,k)) : ((1,8,LBO),(64,SBO)) in elements):
// a k-row is 128 bytes = 64 consecutive M/N indices, 8 k-rows form a 1024-byte swizzle atom, atoms of the next 8 k-rows
// are SBO apart, the next 64 M/N indices LBO apart.  layout type 2 in bits [61,64); base_offset (bits [49,52)) = the
// phase of the start address inside the 1024-byte swizzle pattern when the start is not atom-aligned.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t base_offset) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)(base_offset & 7u) << 49) | (2ull << 61);
}

// TMEM -> registers: 32 lanes (this warp's quarter) x 16 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Warp-uniform leader election (elect.sync).  The whole warp runs the issue loops with identical (uniform) values
// and only the tcgen05 instructions are predicated on the elected lane: descriptors then live in uniform registers.
// (Issuing from inside an `if (lane == 0)` branch makes every operand divergent and costs a R2UR "waterfall" of
// ~15 instructions per tcgen05.mma.)
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\t"
        "elect.sync %%rx|%%px, %1;\n\t"
        "@%%px mov.s32 %0, 1;\n\t}\n"
        : "+r"(pred)
        : "r"(0xFFFFFFFFu));
    return pred != 0;
}

// ---------------------------------------------------------------- small helpers
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
// BatchNorm + (leaky) ReLU of 8 packed bf16 channels: fp32 fma, round to bf16, then -- for slope 0 -- the ReLU as a packed
// bf16x2 max (max commutes with the rounding, so the bits equal the fp32 select except for the sign of a zero; NaN propagates); the
// leaky form keeps the fp32 select.  `relu` must be warp-uniform.
__device__ __forceinline__ uint4 bn_act8(uint4 u, const float* sc, const float* sh, float slope, bool relu) {
    float v[8] = {bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y), bf16lo(u.z), bf16hi(u.z), bf16lo(u.w), bf16hi(u.w)};
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], sc[k], sh[k]);
    uint4 o;
    if (relu) {
        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
        const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f);
        uint32_t* w = &o.x;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const __nv_bfloat162 m = __hmax2_nan(*reinterpret_cast<const __nv_bfloat162*>(&w[q]), zero);
            w[q] = *reinterpret_cast<const uint32_t*>(&m);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = v[k] > 0.f ? v[k] : v[k] * slope;
        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    }
    return o;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// n / d with one multiply-high (d is a runtime constant of the launch), m = floor((2^32-1)/d) + 1.
// FastDivS: exact only while n * d < 2^32 -- for tile-local indices (n < 2^16) in the convolution loaders.
struct FastDivS {
    uint32_t m, d;
    __device__ __forceinline__ FastDivS() : m(0), d(1) {}
    __device__ __forceinline__ explicit FastDivS(uint32_t d_) : m(d_ > 1 ? 0xFFFFFFFFu / d_ + 1u : 0u), d(d_) {}
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return d > 1 ? __umulhi(n, m) : n; }
};
// FastDiv: exact for every n < 2^31 (flat pixel indices of whole tensors).  The multiply-high estimate is the true
// quotient or one above it (m*d - 2^32 lies in (0, d]), so one multiply-compare corrects it.  The uncorrected form
// returned row+1 for the last pixels of a B=16 352x1216 prediction (n*d > 2^32).
struct FastDiv {
    uint32_t m, d;
    __device__ __forceinline__ FastDiv() : m(0), d(1) {}
    __device__ __forceinline__ explicit FastDiv(uint32_t d_) : m(d_ > 1 ? 0xFFFFFFFFu / d_ + 1u : 0u), d(d_) {}
    __device__ __forceinline__ uint32_t div(uint32_t n) const {
        if (d <= 1) return n;
        uint32_t q = __umulhi(n, m);
        if (q * d > n) --q;
        return q;
    }
};

// ---------------------------------------------------------------- fused BatchNorm finalisation (rd_bn_tail)
// Executed by every thread of the LAST block of the statistics-producing kernel (after the ticket); the fp64 sums
// were accumulated with L2 atomics by all blocks, hence the .cg loads.
template <typename TAIL>
__device__ __forceinline__ int tail_slots(const TAIL& t) { return (t.counter && t.slots > 1) ? t.slots : 1; }
__device__ __forceinline__ double slot_sum(const double* p, int c, int slots, long long stride) {
    // eight independent L2 loads in flight per round (a dependent load-add chain would pay the L2 latency per copy)
    double s = 0.0;
    for (int k0 = 0; k0 < slots; k0 += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (k0 + u < slots) ? __ldcg(p + (size_t)(k0 + u) * stride + c) : 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) s += v[u];
    }
    return s;
}
template <typename TAIL>
__device__ __forceinline__ void bn_tail_run(const TAIL& t, int tid, int nthreads) {
    const int slots = tail_slots(t);
    const long long sstr = t.slot_stride;
    for (int j = 0; j < t.njobs; ++j) {
        const auto& J = t.job[j];
        if (J.kind == 1) {
            if (tid == 0 && J.nbt) *J.nbt += 1;
            for (int c = tid; c < J.C; c += nthreads) {
                const double m = slot_sum(J.sum_a, c, slots, sstr) / J.count;
                double var = slot_sum(J.sum_b, c, slots, sstr) / J.count - m * m;
                if (var < 0.0) var = 0.0;
                const float mean = (float)m;
                const float invstd = (float)(1.0 / sqrt(var + (double)J.eps));
                const double unbiased = J.count > 1.0 ? var * J.count / (J.count - 1.0) : var;
                J.running_mean[c] = (1.f - J.momentum) * J.running_mean[c] + J.momentum * mean;
                J.running_var[c] = (1.f - J.momentum) * J.running_var[c] + J.momentum * (float)unbiased;
                const float sc = J.gamma[c] * invstd;
                J.v0[c] = sc;
                J.v1[c] = J.beta[c] - mean * sc;
                J.v2[c] = mean;
                J.v3[c] = invstd;
            }
        } else if (J.kind == 2) {
            for (int c = tid; c < J.C; c += nthreads) {
                const double sg = slot_sum(J.sum_a, c, slots, sstr), sgz = slot_sum(J.sum_b, c, slots, sstr);
                const double mu = J.v0[c], r = J.v1[c], g = J.gamma[c];
                const double dg = r * (sgz - mu * sg);
                J.v2[c] += (float)dg;
                J.v3[c] += (float)sg;
                J.cA[c] = (float)(g * r);
                J.cB[c] = (float)(-g * r * r * dg / J.count);
                J.cC[c] = (float)(-g * r * sg / J.count + g * r * r * mu * dg / J.count);
            }
        }
    }
}

// Activation storage abstraction: bf16 (throughput mode) or fp32 (parity mode).
template <typename T> struct Act;
template <> struct Act<bf16> {
    static constexpr int kBytes = 2;
    // load 8 consecutive channels
    static __device__ __forceinline__ void load8(const bf16* p, float* v) {
        uint4 u = *reinterpret_cast<const uint4*>(p);
        v[0] = bf16lo(u.x); v[1] = bf16hi(u.x); v[2] = bf16lo(u.y); v[3] = bf16hi(u.y);
        v[4] = bf16lo(u.z); v[5] = bf16hi(u.z); v[6] = bf16lo(u.w); v[7] = bf16hi(u.w);
    }
    // the same load split in two, so that several loads can be in flight while their data stays packed (4 registers)
    typedef uint4 Raw;
    static __device__ __forceinline__ Raw load_raw(const bf16* p) { return *reinterpret_cast<const uint4*>(p); }
    static __device__ __forceinline__ void unpack(const Raw& u, float* v) {
        v[0] = bf16lo(u.x); v[1] = bf16hi(u.x); v[2] = bf16lo(u.y); v[3] = bf16hi(u.y);
        v[4] = bf16lo(u.z); v[5] = bf16hi(u.z); v[6] = bf16lo(u.w); v[7] = bf16hi(u.w);
    }
    static __device__ __forceinline__ void store8(bf16* p, const float* v) {
        uint4 u;
        u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
        u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(p) = u;
    }
    static __device__ __forceinline__ float ld(const bf16* p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void st(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
    static __device__ __forceinline__ float round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
};
template <> struct Act<float> {
    static constexpr int kBytes = 4;
    static __device__ __forceinline__ void load8(const float* p, float* v) {
        float4 a = *reinterpret_cast<const float4*>(p);
        float4 b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    struct Raw { float4 a, b; };
    static __device__ __forceinline__ Raw load_raw(const float* p) {
        Raw r;
        r.a = *reinterpret_cast<const float4*>(p);
        r.b = *reinterpret_cast<const float4*>(p + 4);
        return r;
    }
    static __device__ __forceinline__ void unpack(const Raw& r, float* v) {
        v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
    }
    static __device__ __forceinline__ void store8(float* p, const float* v) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    static __device__ __forceinline__ float ld(const float* p) { return *p; }
    static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
    static __device__ __forceinline__ float round(float v) { return v; }
};

// 256-bit global accesses (LDG/STG.E.ENL2.256 on sm_100): 16 bf16 channels of one pixel in ONE request per lane.  The conv
// epilogue's lanes each own a different pixel (= a different 128-byte line), so every per-lane request is its own L1
// wavefront on the data path the tensor core's shared-memory operand fetch also uses; two 128-bit halves were two.
__device__ __forceinline__ void ldg256(const void* p, uint32_t* r) {
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t* r) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
                 "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
// 16 consecutive channels; `wide` = the address is 32-byte aligned (checked once per launch by the caller)
__device__ __forceinline__ void load16(const bf16* p, float* v, bool wide) {
    if (wide) {
        uint32_t r[8];
        ldg256(p, r);
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[2 * i] = bf16lo(r[i]); v[2 * i + 1] = bf16hi(r[i]); }
    } else {
        Act<bf16>::load8(p, v); Act<bf16>::load8(p + 8, v + 8);
    }
}
__device__ __forceinline__ void store16(bf16* p, const float* v, bool wide) {
    if (wide) {
        uint32_t r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
        stg256(p, r);
    } else {
        Act<bf16>::store8(p, v); Act<bf16>::store8(p + 8, v + 8);
    }
}
__device__ __forceinline__ void load16(const float* p, float* v, bool wide) {
    if (wide) {
        uint32_t r[8];
        ldg256(p, r);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
        ldg256(p + 8, r);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[8 + i] = __uint_as_float(r[i]);
    } else {
        Act<float>::load8(p, v); Act<float>::load8(p + 8, v + 8);
    }
}
__device__ __forceinline__ void store16(float* p, const float* v, bool wide) {
    if (wide) {
        uint32_t r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(v[i]);
        stg256(p, r);
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(v[8 + i]);
        stg256(p + 8, r);
    } else {
        Act<float>::store8(p, v); Act<float>::store8(p + 8, v + 8);
    }
}


}  // namespace rd
