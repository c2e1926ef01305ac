// radar_depth_b200.cu -- the single translation unit behind libradar_depth_b200.so: extern "C" launchers
// (declared in include/radar_depth_b200.h) around the sm_100a kernels in rd_*.cuh.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cstdint>
#include <mutex>
#include <string>
#include <algorithm>

#include "rd_common.cuh"
#include "rd_conv_fprop.cuh"
#include "rd_conv_wgrad.cuh"
#include "rd_elementwise.cuh"
#include "rd_dataset.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return RD_OK;
    return fail(RD_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define RD_REQUIRE(cond, msg) do { if (!(cond)) return fail(RD_EINVAL, std::string("rd: ") + (msg) + " [" #cond "]"); } while (0)

constexpr int kMaxSmem = 232448;   // 227 KB opt-in per CTA on sm_100
constexpr int kMinSmemExclusive = 117 * 1024;   // > (228 KB - 2 x 1 KB reserved) / 2: at most one such CTA per SM

constexpr int kMaxDevices = 64;
int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute: remembered per device, so a
// process that drives several GPUs opts in on each of them.  `done` is the caller's per-kernel table.
template <typename K>
int set_smem(K kernel, int bytes, bool* done) {
    const int dev = current_device();
    if (done[dev]) return RD_OK;
    int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes), "cudaFuncSetAttribute");
    if (rc == RD_OK) done[dev] = true;
    return rc;
}

// ---- launches: every kernel goes through rd_launch, which sets the programmatic-stream-serialization attribute (see
// pdl_enter() in rd_common.cuh) when RD_PDL=1.  Default off: measured neutral under CUDA-graph replay, see rd_common.cuh
bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("RD_PDL");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v != 0;
}
template <typename... KArgs, typename... Args>
void rd_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);      // errors are read with cudaGetLastError by the caller
}

// ---- deterministic mode (rd_set_deterministic): thread-local, read by the launchers at launch time
thread_local int g_det = 0;
thread_local char* g_scratch = nullptr;
thread_local long long g_scratch_bytes = 0;

// ---- TMA tensor maps (driver entry point resolved through the runtime: no link-time dependency on libcuda)
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TmapEncodeFn tmap_encoder() {
    static TmapEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (TmapEncodeFn)p;
        (void)cudaGetLastError();
    }
    return fn;
}
bool tma_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("RD_TMA");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}
// bf16 NHWC view [B][H][W][pitch] (channels coff .. coff+C) as the 5-D tensor (8, W, H, C/8, B); box (8, box_w, box_rows, 2, 1)
// `step` = 2 loads every second pixel in both directions (one parity plane of a stride-2 operand): the box then spans
// 2*box_w x 2*box_rows pixels of the tensor and still lands as box_w x box_rows slots.
bool encode_nhwc_map(CUtensorMap* map, const rd_view& v, int B, int H, int W, int C, int box_w, int box_rows, int box_chunks = 2, int step = 1) {
    TmapEncodeFn enc = tmap_encoder();
    if (!enc || box_w * step > 256 || box_rows * step > 256 || box_chunks > 256 || C % 8) return false;
    const cuuint64_t pitch_b = (cuuint64_t)v.pitch * 2;
    cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(C / 8), (cuuint64_t)B};
    cuuint64_t strides[4] = {pitch_b, (cuuint64_t)W * pitch_b, 16, (cuuint64_t)H * W * pitch_b};
    cuuint32_t box[5] = {8, (cuuint32_t)(box_w * step), (cuuint32_t)(box_rows * step), (cuuint32_t)box_chunks, 1};
    cuuint32_t es[5] = {1, (cuuint32_t)step, (cuuint32_t)step, 1, 1};
    void* base = (void*)((char*)v.ptr + (size_t)v.coff * 2);
    if (((uintptr_t)base & 15) || (pitch_b & 15)) return false;
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// The same view as the 4-D tensor (C, W, H, B) with box (64 channels, box_w, box_rows, 1) and 128-byte swizzle: a box row
// is ONE 128-byte run per pixel and lands as the SWIZZLE_128B MN-major operand layout of the weight-gradient kernel.
bool encode_nhwc_map_sw128(CUtensorMap* map, const rd_view& v, int B, int H, int W, int C, int box_w, int box_rows, int step = 1) {
    TmapEncodeFn enc = tmap_encoder();
    if (!enc || box_w * step > 256 || box_rows * step > 256 || C % 64) return false;
    const cuuint64_t pitch_b = (cuuint64_t)v.pitch * 2;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {pitch_b, (cuuint64_t)W * pitch_b, (cuuint64_t)H * W * pitch_b};
    cuuint32_t box[4] = {64, (cuuint32_t)(box_w * step), (cuuint32_t)(box_rows * step), 1};
    cuuint32_t es[4] = {1, (cuuint32_t)step, (cuuint32_t)step, 1};
    void* base = (void*)((char*)v.ptr + (size_t)v.coff * 2);
    if (((uintptr_t)base & 15) || (pitch_b & 15)) return false;
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int num_sms() {
    static int n[kMaxDevices] = {0};
    const int dev = current_device();
    if (n[dev] == 0) {
        cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
        if (n[dev] <= 0) n[dev] = 148;
    }
    return n[dev];
}

}  // namespace

extern "C" {

const char* rd_last_error(void) { return g_err.c_str(); }
int rd_version(void) { return 2; }

int rd_set_deterministic(int on, void* scratch, long long scratch_bytes) {
    if (on && (!scratch || scratch_bytes < (1ll << 20) || ((uintptr_t)scratch & 255))) {
        g_det = 0;
        return fail(RD_EINVAL, "rd: deterministic mode needs a 256-byte aligned device scratch buffer of at least 1 MiB");
    }
    g_det = on ? 1 : 0;
    g_scratch = on ? (char*)scratch : nullptr;
    g_scratch_bytes = on ? scratch_bytes : 0;
    return RD_OK;
}
int rd_get_deterministic(void) { return g_det; }

int rd_sizeof(int which) {
    switch (which) {
        case 0: return (int)sizeof(rd_conv_params);
        case 1: return (int)sizeof(rd_wgrad_params);
        case 2: return (int)sizeof(rd_bn_tail);
        case 3: return (int)sizeof(rd_aug_sample);
        default: return -1;
    }
}

int rd_device_error(void* stream) {
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    unsigned int code = 0;
    if (e == cudaSuccess) {
        cudaMemcpyFromSymbol(&code, rd::g_rd_device_error, sizeof(code));
        unsigned int zero = 0;
        cudaMemcpyToSymbol(rd::g_rd_device_error, &zero, sizeof(zero));
    } else {
        g_err = std::string("stream sync: ") + cudaGetErrorString(e);
        return RD_ECUDA;
    }
    return (int)code;
}

static const char* check_tail(const rd_bn_tail& t);

int rd_conv_fprop(const rd_conv_params* p, void* stream) {
    RD_REQUIRE(p != nullptr, "null params");
    RD_REQUIRE(p->Cin > 0 && p->Cin % 16 == 0 && p->Cin <= 1024, "Cin must be a multiple of 16 (<= 1024)");
    RD_REQUIRE(p->N >= 16 && p->N <= 256 && p->N % 16 == 0, "N must be a multiple of 16 in [16,256]");
    RD_REQUIRE(p->S == 1 || p->S == 2, "S must be 1 or 2");
    RD_REQUIRE(p->P >= 1 && p->P <= RD_MAX_PHASES && p->MB >= 1, "bad phase / block count");
    RD_REQUIRE(p->P * p->MB * p->N <= 512, "accumulators exceed 512 TMEM columns");
    RD_REQUIRE(p->ntaps >= 1 && p->ntaps <= RD_MAX_TAPS && p->ngroups >= 1 && p->ngroups <= RD_MAX_GROUPS, "bad tap program");
    RD_REQUIRE(p->IS >= 1 && p->IS <= rd::kMaxStages && p->WS >= 1 && p->WS <= rd::kMaxStages, "bad ring depth");
    RD_REQUIRE(p->src.pitch % 8 == 0 && p->src.coff % 8 == 0 && p->dst.pitch % 8 == 0 && p->dst.coff % 8 == 0, "views must be 8-channel aligned");
    RD_REQUIRE(p->Wl >= p->Wt && p->Ht * p->Wl <= p->MB * 128, "tile does not fit its accumulator blocks");
    const int parts = (p->act_dtype == RD_F32) ? 2 : 1;
    const int PS = p->S * p->S * p->plane_slots;
    RD_REQUIRE(p->chunk_stride >= PS, "chunk_stride smaller than the plane set");
    RD_REQUIRE(p->istage_bytes >= parts * 2 * p->chunk_stride * 16 && p->istage_bytes % 128 == 0, "istage_bytes too small / misaligned");
    int max_shift = 0, max_grp = 0;
    for (int t = 0; t < p->ntaps; ++t) {
        RD_REQUIRE(p->taps[t].a_shift >= 0 && p->taps[t].phase >= 0 && p->taps[t].phase < p->P, "bad tap");
        if (p->taps[t].a_shift > max_shift) max_shift = p->taps[t].a_shift;
    }
    RD_REQUIRE(max_shift + p->MB * 128 <= PS, "tap shift reads past the staged tile");
    int covered = 0;
    for (int g = 0; g < p->ngroups; ++g) {
        RD_REQUIRE(p->grp_first[g] == covered && p->grp_n[g] >= 1, "tap groups must partition the tap list in order");
        covered += p->grp_n[g];
        if (p->grp_n[g] > max_grp) max_grp = p->grp_n[g];
    }
    RD_REQUIRE(covered == p->ntaps, "tap groups must cover all taps");
    RD_REQUIRE(p->wstage_bytes >= max_grp * parts * p->N * 32 && p->wstage_bytes % 128 == 0, "wstage_bytes too small / misaligned");
    RD_REQUIRE(p->epi == 0 || (p->epi == 1 && p->zsrc.ptr && p->ep_scale && p->ep_shift) ||
               (p->epi == 2 && p->ep_scale && p->ep_shift && p->ep_split % 16 == 0 && p->stats == nullptr),
               "epi 1 needs zsrc / scale / shift; epi 2 needs scale / shift, a 16-aligned split and no statistics");
    RD_REQUIRE(p->stats == nullptr || p->stats_stride >= p->nblk * p->N, "stats_stride too small");
    RD_REQUIRE(p->tail.counter == nullptr || p->stats != nullptr, "a fused BatchNorm finalisation needs the statistics epilogue");
    { const char* e = check_tail(p->tail); if (e) return fail(RD_EINVAL, e); }
    const int ntiles = p->tiles_y * p->tiles_x * p->B;
    RD_REQUIRE(ntiles > 0 && p->nblk >= 1, "empty problem");
    int gx = ntiles;
    if (p->max_ctas > 0 && gx > p->max_ctas) gx = p->max_ctas;
    // deterministic statistics: per-warp arrays behind the rings, per-CTA partials in the registered scratch buffer
    float* det_part = nullptr;
    long long det_smem = 0;
    if (g_det && p->stats != nullptr) {
        RD_REQUIRE(p->tail.counter != nullptr, "deterministic statistics need the fused BatchNorm finalisation (rd_bn_tail)");
        det_smem = 8ll * 2 * p->N * 4;
        RD_REQUIRE((long long)gx * p->nblk * 2 * p->N * 4 <= g_scratch_bytes, "deterministic mode: scratch buffer too small");
        det_part = (float*)g_scratch;
    }
    long long smem = (long long)rd::kSmemHeader + (long long)p->IS * p->istage_bytes + (long long)p->WS * p->wstage_bytes + det_smem;
    RD_REQUIRE(smem <= kMaxSmem, "shared memory budget exceeded");
    // Every CTA allocates all 512 TMEM columns: two convolution CTAs (of concurrent launches on different streams) must
    // never share an SM, or the second would sit in tcgen05.alloc until the first retires.  More than half of the SM's
    // shared memory per CTA guarantees that.
    if (smem < kMinSmemExclusive) smem = kMinSmemExclusive;
    dim3 grid(gx, p->nblk, 1), block(rd::kFpropThreads, 1, 1);
    cudaStream_t st = (cudaStream_t)stream;
    // Raw bf16 stride-1 source tiles go through TMA: one box (8 ch, Wl, plane_rows, 2 chunks) per 16-channel stage.  TMA
    // writes the two chunk planes plane_rows*Wl slots apart, so the kernel gets that chunk stride; the slots a tap shift
    // reads past a plane (junk accumulator rows, never stored) must still lie inside the stage.
    rd_conv_params q = *p;
    if (det_part) q.tail.slots = 1;
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    int use_tma = 0;
    if (tma_enabled() && p->act_dtype == RD_BF16 && p->S == 1) {
        const long long cs = (long long)p->plane_rows * p->Wl;
        if ((cs + max_shift + (long long)p->MB * 128) * 16 <= p->istage_bytes &&
            encode_nhwc_map(&map, p->src, p->B, p->srcH, p->srcW, p->Cin, p->Wl, p->plane_rows)) {
            use_tma = 1;
            q.chunk_stride = (int32_t)cs;
        }
    }
    if (p->act_dtype == RD_BF16) {
        static bool done[kMaxDevices] = {false};
        { int rc = set_smem(rd::conv_fprop_kernel<rd::bf16, 1>, kMaxSmem, done); if (rc) return rc; }
        rd_launch(rd::conv_fprop_kernel<rd::bf16, 1>, dim3(grid), dim3(block), (size_t)((size_t)smem), st, q, map, use_tma, det_part);
    } else if (p->act_dtype == RD_F32) {
        static bool done[kMaxDevices] = {false};
        { int rc = set_smem(rd::conv_fprop_kernel<float, 3>, kMaxSmem, done); if (rc) return rc; }
        rd_launch(rd::conv_fprop_kernel<float, 3>, dim3(grid), dim3(block), (size_t)((size_t)smem), st, q, map, 0, det_part);
    } else {
        return fail(RD_EINVAL, "rd: bad act_dtype");
    }
    return check_cuda(cudaGetLastError(), "conv_fprop launch");
}

#include "rd_api_rest.inc"

}  // extern "C"
