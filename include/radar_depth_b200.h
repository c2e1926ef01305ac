/* radar_depth_b200.h -- C ABI of libradar_depth_b200.so (sm_100a).
 *
 * This library is the device side of the radar_depth hot path:
 * ResNet_latefusion / ResNet_latefusion2 forward+backward (reference model/models.py:519-664,
 * model/multistage_model.py:123-276), Filter_layer (multistage_model.py:87-119) and the
 * MaskedL1 / Smoothness losses (evaluation/criteria_new.py:8-54).  The reference owns no kernels:
 * every entry point below replaces a torch.nn / ATen call the reference issues (the call site is
 * cited next to each function).  The host side (radar_depth_b200/model/*.py) mirrors the reference's
 * nn.Module constructors and sequences these calls.
 *
 * Conventions
 *  - plain C: raw DEVICE pointers, ints, a cudaStream_t passed as void*; no torch types.
 *  - every function returns 0 on success or a negative RD_E* code; rd_last_error() gives the text.
 *  - nothing allocates, frees or synchronises the device; the caller owns every buffer.
 *  - activations are NHWC "views": base pointer + pixel pitch (elements) + channel offset, so that
 *    concatenations (reference torch.cat, models.py:652) are just adjacent channel ranges.
 *  - act_dtype: RD_BF16 (throughput mode) or RD_F32 (parity mode: fp32 activations and a 3-term
 *    bf16 split on the tensor cores, ~fp32 accuracy).
 */
#ifndef RADAR_DEPTH_B200_H_
#define RADAR_DEPTH_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RD_OK 0
#define RD_EINVAL (-1)   /* bad argument / unsupported shape */
#define RD_ECUDA (-2)    /* CUDA runtime error (text in rd_last_error) */
#define RD_EDEVICE (-3)  /* device-side protocol error (barrier timeout) */

#define RD_BF16 0
#define RD_F32 1

#define RD_MAX_TAPS 32
#define RD_MAX_GROUPS 16
#define RD_MAX_PHASES 4

const char* rd_last_error(void);
int rd_version(void);
/* sizeof() of the parameter blocks below, for binding self-checks. */
int rd_sizeof(int which); /* 0: rd_conv_params, 1: rd_wgrad_params, 2: rd_bn_tail, 3: rd_aug_sample */
/* reads and clears the device-side error word (0 = none).  Synchronises the stream. */
int rd_device_error(void* stream);

/* Deterministic mode.  The reference's training step is reproducible on its CPU path; its cuDNN path is not
 * (torch.backends.cudnn.deterministic is never set, main.py:11,47).  With on != 0 every launch issued by the CALLING THREAD
 * afterwards (the flag is thread-local and read at launch time, so it is baked into a captured CUDA graph) replaces its
 * floating-point atomics by fixed-order sums: BatchNorm statistics (conv epilogues, rd_join_bwd, rd_maxpool_bwd), weight
 * gradients (rd_conv_wgrad, rd_head_conv_bwd) and the loss sums (rd_l1_fwd, rd_smoothness_*) become bit-identical from run
 * to run.  scratch = caller-owned device buffer (256-byte aligned, >= 1 MiB and >= max over the weight-gradient launches of
 * max_ctas * ntaps * Cout * Cin * 4 bytes) that holds the per-CTA partials; launches that use it must be stream-ordered.
 * The parity mode (act_dtype RD_F32) of the Python layer switches it on by default. */
int rd_set_deterministic(int on, void* scratch, long long scratch_bytes);
int rd_get_deterministic(void);

/* NHWC view of a [B, H, W, C] slice living inside a [B, H, W, pitch] buffer. */
typedef struct rd_view {
    void* ptr;
    int32_t pitch; /* elements per pixel in the underlying buffer (multiple of 8) */
    int32_t coff;  /* first channel of the slice (multiple of 8) */
} rd_view;

/* BatchNorm finalisation fused into the tail of the kernel that produces its statistics: the LAST CTA to finish
 * (ticket counter) turns the per-channel fp64 sums into the vectors the next kernel consumes, so that no separate
 * one-block launch sits between two convolutions.  Same arithmetic as rd_bn_finalize / rd_bn_bwd_finalize.
 *   kind 1 (forward, training):  sum_a = sum z, sum_b = sum z^2 -> v0 = scale, v1 = shift, v2 = save_mean,
 *                                v3 = save_invstd; running_mean/var (unbiased) and num_batches_tracked updated.
 *   kind 2 (backward):           sum_a = sum g, sum_b = sum g*z, v0 = save_mean (read), v1 = save_invstd (read),
 *                                v2 = dgamma (+=), v3 = dbeta (+=), cA/cB/cC = coefficients of dz = A*g + B*z + C. */
typedef struct rd_bn_job {
    int32_t kind; /* 0 = none */
    int32_t C;
    const double* sum_a;
    const double* sum_b;
    double count;
    const float* gamma;
    const float* beta;
    float* running_mean;
    float* running_var;
    long long* nbt;
    float* v0;
    float* v1;
    float* v2;
    float* v3;
    float* cA;
    float* cB;
    float* cC;
    float momentum, eps;
} rd_bn_job;

#define RD_MAX_BN_JOBS 2
typedef struct rd_bn_tail {
    uint32_t* counter; /* zero before the launch; NULL = no fused finalisation */
    int32_t njobs;
    int32_t slots;     /* >1: the statistics are accumulated into `slots` copies (block b adds to copy b % slots, copy k
                          of an array lives slot_stride doubles after copy k-1) and summed by the finalising CTA: same-
                          address fp64 atomics serialise in L2, this divides that tail by `slots`.  0/1 = one copy. */
    int64_t slot_stride;
    rd_bn_job job[RD_MAX_BN_JOBS];
} rd_bn_tail;

typedef struct rd_tap {
    int32_t a_shift; /* slot offset of this tap inside the staged input tile (includes the parity-plane base) */
    int32_t phase;   /* output phase this tap accumulates into */
    int32_t first;   /* 1 = first tap of its phase (clears the accumulator on the first channel block) */
    int32_t pad_;
} rd_tap;

/* Implicit-GEMM convolution on tcgen05 (forward convs, data gradients, sub-pixel UpProj convs).
 * Replaces nn.Conv2d / F.conv_transpose2d calls: models.py:539,559,573,582,191,194,198,27 and torchvision
 * conv3x3/conv1x1 (models.py:8,88,91,605-608) and their autograd data-gradient counterparts.
 *
 * out[b, oy*OS+phy, ox*OS+phx, n] = sum_taps sum_c  T(src)[b, (oy+sy)*S+py, (ox+sx)*S+px, c] * W[tap][c][n]
 * where T is the optional fused per-channel affine+activation of the producer's BatchNorm
 * (models.py:540,546 etc.: relu(bn(z)) is never materialised).                                   */
typedef struct rd_conv_params {
    /* source */
    rd_view src;
    int32_t srcH, srcW, Cin; /* Cin multiple of 16 */
    int32_t S;               /* 1, or 2 = source staged as 4 parity planes (stride-2 convs) */
    const float* ld_scale;   /* [Cin] or NULL: fused y = act(x*scale+shift) on load */
    const float* ld_shift;
    float ld_slope;          /* negative-side slope of act: 0 relu, 0.2 leaky, 1 identity */
    int32_t B;
    /* tile geometry in base output pixels */
    int32_t Hb, Wb;          /* base output size (before phase interleave) */
    int32_t Ht, Wt, Wl;      /* tile rows, valid cols, local row width in slots */
    int32_t plane_rows, plane_slots;
    int32_t chunk_stride;    /* slots between consecutive 8-channel chunk planes (>= S*S*plane_slots; padded against bank conflicts) */
    int32_t sy_min, sx_min;  /* plane-coordinate offset of slot 0 relative to the tile origin */
    int32_t MB;              /* 128-row accumulator blocks per tile */
    int32_t tiles_y, tiles_x;
    /* tap program */
    int32_t P, OS, ntaps, ngroups;
    int32_t phase_y[RD_MAX_PHASES], phase_x[RD_MAX_PHASES];
    rd_tap taps[RD_MAX_TAPS];
    int32_t grp_first[RD_MAX_GROUPS], grp_n[RD_MAX_GROUPS];
    /* weights, packed by rd_pack_weights: [nblk][Cin/16][tap][part][2][N][8] bf16 */
    const void* wpk;
    int32_t N;               /* output channels per CTA (multiple of 16, <= 256) */
    int32_t nblk;            /* number of N blocks (grid.y) */
    /* destination */
    rd_view dst;
    int32_t dstH, dstW;
    /* epilogue */
    int32_t epi;             /* 0 store(+addend)+sum/sumsq stats; 1 activation-gradient mask + sum g / sum g*z stats;
                                2 folded BatchNorm + addend + activation (inference, see ep_split below) */
    rd_view addend;          /* ptr NULL = none; same spatial size as dst */
    rd_view zsrc;            /* epi 1: producer's pre-BN output z, same spatial size as dst */
    const float* ep_scale;   /* epi 1: [nblk*N] BN scale/shift of the layer being differentiated */
    const float* ep_shift;
    float ep_slope;
    double* stats;           /* [2][stats_stride] or NULL */
    int32_t stats_stride;
    rd_bn_tail tail;         /* BatchNorm finalisation(s) run by the last CTA (needs stats) */
    /* pipeline */
    int32_t IS, WS;          /* input / weight ring depths */
    int32_t istage_bytes, wstage_bytes;
    int32_t act_dtype;       /* RD_BF16 / RD_F32 */
    int32_t max_ctas;        /* persistent grid cap (0 = one CTA per tile) */
    long long* dbg;          /* optional [6][gridDim.x*gridDim.y] cycle counters: loader wait/fill, issuer wait/issue, epilogue wait/work */
    int32_t dbg_flags;       /* diagnostics only: 1 = skip UMMA issue, 2 = skip tile staging, 4 = skip epilogue stores */
    int32_t src_planes;      /* S = 2: number of parity planes the taps actually read, staged from plane 0 on (0 = all four);
                                a 1x1 stride-2 convolution only ever reads plane (0,0) */
    /* epi 2 (inference, main.py:584-595 under model.eval() + torch.no_grad()): the BatchNorm that follows the convolution is
     * a fixed per-channel affine map (running statistics), so it is folded into the epilogue together with the residual
     * add and the activation: out = act(acc * ep_scale[n] + ep_shift[n] + addend), slope ep_slope for output channels
     * < ep_split and ep_slope_b from ep_split on (UpProj: ReLU on the upper branch, identity on the bottom branch; stems:
     * ReLU on the RGB channels, LeakyReLU(0.2) on the depth channels).  No statistics, no separate join kernel. */
    int32_t ep_split;        /* multiple of 16 */
    float ep_slope_b;
} rd_conv_params;

int rd_conv_fprop(const rd_conv_params* p, void* stream);

/* Weight gradient of the same convolutions on tcgen05 (replaces autograd's conv weight-gradient kernels for
 * every nn.Conv2d above).  Contraction runs over pixels: both operands are staged pixel-linear and fed
 * to the tensor core as MN-major matrices; each tap has its own TMEM accumulator [co x ci]; CTAs split the
 * pixel tiles and reduce with vector fp32 atomics into dw[tap][Cout][Cin].
 *   dw[tap][co][ci] += sum_pixels gy[b, oy*Sg+gpy, ox*Sg+gpx, co] * T(x)[b, (oy+sy)*Sx+py, (ox+sx)*Sx+px, ci]  */
typedef struct rd_wtap {
    int32_t g_off;   /* slot offset of the gradient parity plane of this tap */
    int32_t x_shift; /* slot offset into the staged source tile (includes its parity-plane base) */
} rd_wtap;

typedef struct rd_wgrad_params {
    rd_view gy;
    int32_t gH, gW, Cout, Sg;
    rd_view x;
    int32_t xH, xW, Cin, Sx;
    const float* ld_scale; /* fused BN+activation on x (NULL = raw) */
    const float* ld_shift;
    float ld_slope;
    int32_t B;
    int32_t Hb, Wb, Ht, Wt, Wl;
    int32_t KS;                 /* gradient slots per plane (multiple of 16, >= Ht*Wl) */
    int32_t g_chunk_stride, x_chunk_stride; /* slots between chunk planes of the two staged tiles */
    int32_t x_plane_rows, x_plane_slots, sy_min, sx_min;
    int32_t tiles_y, tiles_x;
    int32_t ntaps, tg_size, ntg;  /* taps per CTA and number of tap groups */
    rd_wtap taps[RD_MAX_TAPS];
    int32_t Mc, ncob;            /* output-channel rows per CTA (<=128, multiple of 8) and block count */
    int32_t Nc, ncib;            /* input-channel columns per CTA (multiple of 16) and block count */
    float* dw;                   /* [ntaps][Cout][Cin] fp32, accumulated with atomics (caller zeroes) */
    int32_t NS, stage_bytes, g_bytes;
    int32_t act_dtype;
    int32_t max_ctas;            /* pixel-split CTAs (grid.x) */
    long long* dbg;              /* optional [4][gridDim.x*y*z] cycle counters (loader wait/fill, issuer wait/issue); NULL = off */
    int32_t dbg_flags;           /* diagnostics only: 1 = skip UMMA issue, 2 = skip tile staging */
    int32_t x_planes;            /* Sx = 2: parity planes of the source the taps read, from plane 0 on (0 = all four) */
    /* Tap-row folding for 16-channel sources (Nc = Cin = 16, Sx = 1, one tap group): the taps form fold_rows rows of
     * fold_len (<= 4) horizontally adjacent taps (taps[r*fold_len + i].x_shift = taps[r*fold_len].x_shift + i).  One UMMA
     * then covers a whole row for one 8-channel chunk: its B operand is that chunk plane with an MN-chunk stride of ONE
     * slot, i.e. N = 4 taps x 8 channels = 32 (a 4th tap of a 3-tap row is junk and ignored).  fold_rows*2 UMMAs of N=32
     * per 16 pixels instead of ntaps UMMAs of N=16 -- small-N UMMAs cost ~40 cycles whatever N is.  0 = off. */
    int32_t fold_rows, fold_len;
    int32_t pad2_;
    /* Gradient copies stacked in M (stride-1 convs with Cout <= 64, gradient tile staged by TMA): the M = 128 operand of the
     * UMMA spans 16 chunk planes of which a Cout-channel gradient tile fills only Cout/8, so `gcopies` copies of the tile,
     * copy r shifted by (gcopy_dy[r], gcopy_dx[r]) pixels (a second TMA box at shifted coordinates), are laid side by side:
     * rows [r*Mc, (r+1)*Mc) of accumulator j then hold the tap whose source shift is taps[j].x_shift MINUS copy r's shift.
     * One UMMA covers up to 128/Mc taps: 9 -> 6 accumulators (one pass instead of two) for 64-channel layers, 9 -> 3 (or, with
     * tap-row folding, 6 -> 2) for 16/32-channel layers.  The pixel-tile grid starts at (-tile_oy, -tile_ox) so that every
     * gradient pixel meets every copy exactly once.  njobs accumulators (taps[0..njobs) are their descriptors); job_tap[j][r]
     * = index into dw of the tap produced by copy r of job j (with tap-row folding: of the first tap of its row), -1 = unused.
     * gcopies <= 1: off (one accumulator per tap, as above). */
    int32_t gcopies, njobs;
    int32_t gcopy_dy[8], gcopy_dx[8];
    int32_t tile_oy, tile_ox;
    int8_t job_tap[RD_MAX_TAPS][8];
} rd_wgrad_params;

int rd_conv_wgrad(const rd_wgrad_params* p, void* stream);

/* ---- HBM-bound kernels (rd_elementwise.cuh).  npix = B*H*W of the tensors involved. ---- */

/* NCHW fp32 [B,C,H,W] -> space-to-depth NHWC [B,ceil(H/2),ceil(W/2),4*Cs] (channel = parity*Cs + c).  Replaces the
 * x[:, :3] / x[:, 3:] slicing at models.py:633,643 and multistage_model.py:236-241. */
int rd_input_pack(const float* x, void* out, int B, int C, int H, int W, int Cs, int act_dtype, void* stream);

/* The same packing with channel c read from planes[c] (fp32 [H][W] plane of image 0, image b at + b * batch_strides[c]
 * elements): replaces torch.cat((x_img, x_d_filtered, depth_stage1), dim=1) at multistage_model.py:78 -- the stage-2 input
 * is never materialised.  planes / batch_strides are HOST arrays of C entries. */
int rd_input_pack_parts(const float* const* planes, const long long* batch_strides, void* out, int B, int C, int H, int W, int Cs,
                        int act_dtype, void* stream);

/* Channel c of d(loss)/d(input), fp32 [B,1,H,W], from the stem's space-to-depth data gradient [B,ceil(H/2),ceil(W/2),4*Cs]:
 * autograd's slice of the stage-2 input gradient that flows into stage 1 (multistage_model.py:75,78). */
int rd_input_grad_channel(const void* dxs, float* out, int B, int H, int W, int Cs, int c, int act_dtype, void* stream);

/* nn.BatchNorm2d statistics -> fused scale/shift (+ running-stat update in training).  models.py:540 etc. */
int rd_bn_finalize(const double* sum, const double* sumsq, double count, const float* gamma, const float* beta,
                   float* running_mean, float* running_var, long long* num_batches_tracked, int C, int training,
                   float momentum, float eps, float* scale, float* shift, float* save_mean, float* save_invstd,
                   void* stream);
/* Eval-mode finalisation of MANY BatchNorm layers in one launch (inference: scale/shift only depend on the parameters and
 * the running statistics).  table = n rows of 9 device words (int64): gamma, beta, running_mean, running_var, scale,
 * shift, save_mean, save_invstd (pointers) and C. */
int rd_bn_finalize_eval_multi(const long long* table, int n, float eps, void* stream);
/* BatchNorm2d backward reductions -> dgamma/dbeta (accumulated) and dz = A*g + B*z + C coefficients. */
int rd_bn_bwd_finalize(const double* sum_g, const double* sum_gz, double count, const float* gamma,
                       const float* save_mean, const float* save_invstd, int C, int training, float* dgamma,
                       float* dbeta, float* coefA, float* coefB, float* coefC, void* stream);
/* out = act(z*sc+sh + identity): residual join of BasicBlock (models.py:104-110) / UpProjModule (models.py:205-208). */
int rd_bn_add_act(rd_view z, const float* sc, const float* sh, rd_view idv, const float* id_sc, const float* id_sh,
                  rd_view out, long long npix, int C, float slope, int act_dtype, void* stream);
/* g = dout*act'(out) + BN-backward statistics of the joined branches. */
int rd_join_bwd(rd_view dout, rd_view out, rd_view z, rd_view zid, rd_view g, long long npix, int C, float slope,
                double* sum_g, double* sum_gz, double* sum_gzid, const rd_bn_tail* tail /* may be NULL */, int act_dtype,
                void* stream);
/* dz = A*g + B*z + C per channel: the elementwise half of autograd's nn.BatchNorm2d backward (every BatchNorm of
 * models.py:539-594; coefficients from rd_bn_bwd_finalize or a rd_bn_tail job).  dz may alias g. */
int rd_bn_bwd_apply(rd_view g, rd_view z, rd_view dz, const float* coefA, const float* coefB, const float* coefC,
                    long long npix, int C, int act_dtype, void* stream);
/* sum g, sum g*z over a tensor whose gradient is already final (BatchNorm backward reductions without an activation
 * mask, e.g. bn_fusion / bn2 of models.py:652-657 when their gradient comes from an elementwise producer). */
int rd_grad_stats(rd_view g, rd_view z, long long npix, int C, double* sum_g, double* sum_gz, int act_dtype, void* stream);

/* nn.MaxPool2d(3,2,1) over act(bn(z)) (models.py:546-547,564-565), both stems at once; stores the arg-max.
 * zarg (optional, bf16 only): [B][Ho][Wo][C] bf16, the PRE-activation value z of every window's winner, for
 * rd_maxpool_bwd_stats. */
int rd_maxpool_fwd(rd_view z, const float* sc, const float* sh, int B, int H, int W, int C, int split, float slope_a,
                   float slope_b, rd_view outa, rd_view outb, uint8_t* amax, int Ho, int Wo, void* zarg, int act_dtype,
                   void* stream);
/* The backward of the same max-pool + activation + (training-mode) BatchNorm in two passes that never materialise the
 * gradient of the BatchNorm output (bf16 path; replaces rd_maxpool_bwd + rd_bn_bwd_apply, i.e. autograd of
 * models.py:540-547,560-565):
 *   rd_maxpool_bwd_stats: sum g and sum g*z over the POOLED elements (a pooled element sends its gradient to exactly one
 *     input pixel) from the pooled gradient and zarg; `tail` finalises dgamma / dbeta and the coefficients of
 *     dz = A*g + B*z + C like every other statistics producer;
 *   rd_maxpool_bwd_apply: dz for every input pixel: gathers the gradients of the <= 4 windows whose arg-max is this pixel,
 *     applies the activation derivative and the three coefficients, writes dz. */
int rd_maxpool_bwd_stats(rd_view dpa, rd_view dpb, const void* zarg, const float* sc, const float* sh, int B, int Ho, int Wo,
                         int C, int split, float slope_a, float slope_b, double* sum_g, double* sum_gz,
                         const rd_bn_tail* tail /* may be NULL */, void* stream);
int rd_maxpool_bwd_apply(rd_view dpa, rd_view dpb, const uint8_t* amax, rd_view z, const float* sc, const float* sh,
                         const float* coefA, const float* coefB, const float* coefC, int B, int H, int W, int C, int split,
                         float slope_a, float slope_b, int Ho, int Wo, rd_view dz, void* stream);
int rd_maxpool_bwd(rd_view dpa, rd_view dpb, const uint8_t* amax, rd_view z, const float* sc, const float* sh, int B,
                   int H, int W, int C, int split, float slope_a, float slope_b, int Ho, int Wo, rd_view g,
                   double* sum_g, double* sum_gz, const rd_bn_tail* tail /* may be NULL */, int act_dtype, void* stream);

/* conv3 (3x3, 16->1, models.py:587,661) and nn.Upsample(bilinear, align_corners=True) (models.py:588,662). */
int rd_head_conv_fwd(rd_view x, const float* w, int B, int H, int W, float* out, int act_dtype, void* stream);
int rd_head_conv_bwd(const float* dc3, rd_view x, const float* w, int B, int H, int W, rd_view dx, float* dw,
                     int act_dtype, void* stream);
int rd_bilinear_fwd(const float* in, int B, int Hi, int Wi, float* out, int Ho, int Wo, void* stream);
int rd_bilinear_bwd(const float* dout, int B, int Hi, int Wi, float* din, int Ho, int Wo, void* stream);

/* MaskedL1Loss (criteria_new.py:44-54).  acc = 2 doubles of scratch (sum, count), zeroed by the call. */
int rd_l1_fwd(const float* pred, const float* target, long long n, double* acc, float* loss, void* stream);
int rd_l1_bwd(const float* pred, const float* target, long long n, const double* acc, const float* gout,
              float* gpred, int accumulate, void* stream);

/* SmoothnessLoss (criteria_new.py:8-28); image = the C-channel tensor main.py:422 passes (all 4 network inputs).
 * scratch = 2*B+2 doubles (per-image sums, the two loss terms, per-image sum of G*d); fwd zeroes it, bwd reuses it. */
int rd_smoothness_fwd(const float* pred, const float* image, int B, int C, int H, int W, double* scratch, float* loss,
                      void* stream);
int rd_smoothness_bwd(const float* pred, const float* image, int B, int C, int H, int W, double* scratch,
                      const float* gout, float* gpred, int accumulate, void* stream);

/* Result.evaluate / Result_multidist.evaluate (evaluation/metrics.py:34-58,91-140): all masked error statistics of one
 * prediction in ONE pass and one device->host copy (the reference does ~10 boolean gathers and host syncs per call,
 * every training iteration, main.py:450-458).  Pixels with target > 0 and lo <= target <= hi count (lo = 0, hi = +inf for
 * Result).  acc = RD_METRIC_SLOTS doubles, zeroed by the call: count, sum d^2, sum |d|, sum |log10 o - log10 t|,
 * sum |d|/t, #(max(o/t,t/o) < 1.25), # < 1.25^2, # < 1.25^3, sum (1/o - 1/t)^2, sum |1/o - 1/t|. */
#define RD_METRIC_SLOTS 10
int rd_depth_metrics(const float* output, const float* target, long long n, float lo, float hi, double* acc, void* stream);

/* Filter_layer (multistage_model.py:87-119). */
int rd_sid_filter(const float* radar, const float* depth, long long n, float* radar_f, float* mask, void* stream);

/* Packed bf16 weights from the flat fp32 parameter arena via a gather table; gradient scatter back; fused SGD
 * (torch.optim.SGD as configured at main.py:285-290). */
int rd_pack_weights(const float* src, const int32_t* idx, void* out, long long n, void* stream);

/* Inference program (model.eval() + torch.no_grad(), main.py:584-595: the weights do not change between forwards):
 * rd_weights_hash compares a content hash of the fp32 parameter arena (nchunks 64-bit values in `state`, zero-initialised by the
 * caller; splitmix64 of every word and its index, summed per chunk) with the stored one, stores the new hash and sets
 * *dirty = 1 if any chunk changed (0 otherwise); rd_pack_weights_if is rd_pack_weights that returns at once when *dirty == 0.
 * No host-side bookkeeping of "who touched the weights" is involved: in-place updates through .data are seen as well. */
int rd_weights_hash(const float* w, long long n, unsigned long long* state, int nchunks, int* dirty, void* stream);

/* rd_pack_weights with a compact table: groups[g] = (base, stride) describes outputs 8g .. 8g+7 = src[base + i*stride] (bit 30 of
 * base = lo part of the bf16 split, base < 0 = eight zeros); stride < 0 marks a group with holes whose eight rd_pack_weights
 * indices are fallback[8*base .. 8*base+7].  dirty: NULL, or the flag of rd_weights_hash (returns at once when *dirty == 0). */
int rd_pack_weights_g8(const float* src, const int32_t* groups, const int32_t* fallback, void* out, long long ngroups, const int* dirty,
                       void* stream);
int rd_pack_weights_if(const float* src, const int32_t* idx, void* out, long long n, const int* dirty, void* stream);
int rd_unpack_grads(const float* dw, const int32_t* idx, float* grad, long long n, void* stream);
int rd_sgd(float* p, const float* g, float* mom, long long n, float lr, float momentum, float wd, int first, void* stream);
/* Same with the gradient multiplied by grad_scale first: the 1/world of a data-parallel SUM all-reduce folded into
 * the update instead of a separate pass over the 58.8 MB gradient arena. */
int rd_sgd_scaled(float* p, const float* g, float* mom, long long n, float lr, float momentum, float wd, int first,
                  float grad_scale, void* stream);

/* Graph cut at the bottleneck: ResNet_latefusion.pnp_forward_front returns bn2's output, pnp_forward_rear consumes it
 * (models.py:669-707).  export: NHWC activation slice -> NCHW fp32 through an optional per-channel affine (sc/sh both
 * NULL = plain copy; used for the bottleneck feature and for its gradient); import: NCHW fp32 -> NHWC slice. */
int rd_feature_export(rd_view z, const float* sc, const float* sh, float* out_nchw, int B, int H, int W, int C, int act_dtype,
                      void* stream);
int rd_feature_import(const float* x_nchw, rd_view z, int B, int H, int W, int C, int act_dtype, void* stream);

/* ---- Input pipeline on the GPU: the reference's per-sample CPU transforms for a whole batch of RAW exported
 * samples resident in device memory -- uint8 image [B][H][W][3], int16 lidar / radar depth [B][H][W] in 1/256 m
 * (dataset/nuscenes_export.py:24-27).  Replaces dataset/nuscenes_dataset_torch_new.py:190-195 (h5 depth decode), :237-412
 * (transform_train: transforms.Rotate -> Resize -> Crop -> HorizontalFlip, ColorJitter, /255, depth / scale, the
 * max_depth filter of the radar channel :373-374, torch.cat :375) and :415-560 (transform_val: CenterCrop, /255).
 * Bit-exact with scipy.ndimage.rotate(order 0) / PIL Image.resize / PIL ImageEnhance; the random draws and the small
 * per-sample tables (PIL's resampling coefficients and nearest-index tables for the crop window) are made on the host
 * (radar_depth_b200/dataset/gpu_pipeline.py). */
typedef struct rd_aug_sample {
    double m00, m01, m10, m11, off0, off1; /* scipy.ndimage.rotate: input (y, x) = M * output (y, x) + off */
    double factor[3];                      /* ColorJitter factors, in the order the operations are applied */
    float depth_div;                       /* (float) scale factor: depth /= scale (nuscenes_dataset_torch_new.py:311,325) */
    int32_t identity_rot;                  /* 1 = no rotation at all (validation) */
    int32_t flip;                          /* HorizontalFlip */
    int32_t crop_i, crop_j;                /* upper / left corner of the crop (in the resized image; val: in the raw image) */
    int32_t op[3];                         /* 0 brightness, 1 contrast, 2 saturation ("Color"), in application order */
    int32_t pad_;
} rd_aug_sample;

/* rotate -> bytescale (scipy.misc.imresize's toimage) -> PIL bilinear resize -> crop -> flip -> ColorJitter.
 * bil_tab [B][ch + cw][5] int32: for every output row then column of the crop window {first source index, taps (<= 3),
 * k0, k1, k2} in PIL's 22-bit fixed point; scratch: >= B * 32 bytes; img8 [B][ch][cw][3] uint8 (the jittered crop). */
int rd_aug_rgb(const void* images, const rd_aug_sample* samples, const int32_t* bil_tab, void* scratch, int B, int H, int W,
               int ch, int cw, void* img8, void* stream);
/* Final assembly into the network's NCHW fp32 input.  mode 0 (train): rgb from img8, depth through flip -> crop -> PIL
 * nearest (near_tab [B][ch + cw] int32: source row / column in the rotated image) -> rotation; mode 1 (val): centre crop of
 * the raw data at (samples[b].crop_i, crop_j).  inputs [B][3 + has_radar][ch][cw], labels / radar_out [B][1][ch][cw]. */
int rd_aug_pack(const void* img8, const void* images, const void* lidar, const void* radar, const rd_aug_sample* samples,
                const int32_t* near_tab, int B, int H, int W, int ch, int cw, int mode, int has_radar, float max_depth,
                float* inputs, float* labels, float* radar_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RADAR_DEPTH_B200_H_ */
