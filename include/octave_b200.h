/*
 * octave_b200 — C-ABI of the B200-native kernels behind the OCTAve training step.
 *
 * The reference (IoBT-VISTEC/OCTAve, /root/reference) has no FFI of its own: its boundary is the
 * Python class surface.  This header is the C-ABI that sits directly beneath that
 * surface; each entry point names the reference function whose arithmetic it replaces.  The Python
 * host (the octave_b200 Python package, mirroring the reference's architectures package) binds these symbols with
 * ctypes.
 *
 * Conventions
 *  - plain pointers + sizes, no torch types; `stream` is a cudaStream_t passed as void*.
 *  - returns 0 on success, negative OCT_ERR_* otherwise; never throws, never allocates, never
 *    synchronises.  Workspaces are passed in; their size comes from the *_bytes() queries.
 *  - activations of the network are NHWC (channels innermost); loss maps are NCHW planar, exactly
 *    the layout of the reference tensors, so the loss kernels read the caller's tensors in place.
 *  - dtype: 0 = fp32, 1 = bf16 (storage type; accumulation is always fp32 or wider).
 */
#ifndef OCTAVE_B200_H_
#define OCTAVE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OCTAVE_ABI_VERSION 4   /* 4: octave_set_stats_prezeroed, octave_stream_capture_id; 2: OctaveLossDesc.jsd_eps; extra outputs of head_bwd / bn_bwd_apply / space_to_depth */
int octave_abi_version(void);
/* number of SMs of the current device (grid sizing); <0 on error */
int octave_sm_count(void);
/* number of kernels this library has launched in this process (bench.py reports the delta as gpu_launches) */
unsigned long long octave_launch_count(void);
/* Deterministic mode (process-wide, default off): reductions that are normally split across CTAs and merged with fp32
 * atomics (split-K weight gradients, K-split linears, head parameter gradients) run with ONE writer per output element,
 * so two runs on the same inputs give bit-identical results.  Slower for the small layers; off for benchmarks. */
void octave_set_deterministic(int on);
int octave_get_deterministic(void);
/* Statistics contract (process-wide, default off).  Off: every entry point that ACCUMULATES into a double-precision
 * statistics output (`stats` of the conv forward, `sums` / `sums2` of the BatchNorm passes, `chan_sum` of
 * octave_space_to_depth) zeroes it first with its own memset.  On: the caller guarantees those outputs are zero on entry
 * (the Python host hands out slices of one pre-zeroed arena: ~250 memset nodes fewer per training step). */
void octave_set_stats_prezeroed(int on);
int octave_get_stats_prezeroed(void);
/* id of the stream capture `stream` is part of, 0 if it is not capturing (an arena zeroed outside a capture must not be
 * handed out inside it) */
unsigned long long octave_stream_capture_id(void* stream);

/* ------------------------------------------------------------------------------------------------
 * K9 — fused loss kernel (forward statistics pass + gradient pass).
 * Replaces, in one pass over the maps:
 *   WeightedPartialCE.forward   architectures/segmentor/losses.py:26-61 (manual branch :51-55)
 *   DiceLoss.forward            architectures/segmentor/losses.py:70-74
 *   InterlayerDivergence.forward (KLD, mode='mean')  architectures/segmentor/losses.py:111-147
 *   LSGeneratorLoss.forward / LSDiscriminatorialLoss.forward  architectures/discriminator/losses.py:11-24
 * ---------------------------------------------------------------------------------------------- */
#define OCT_LOSS_WPCE        (1 << 0)
#define OCT_LOSS_DICE        (1 << 1)
#define OCT_LOSS_KLD         (1 << 2)
#define OCT_LOSS_LSG         (1 << 3)
#define OCT_LOSS_LSD         (1 << 4)
#define OCT_LOSS_FROM_LOGITS (1 << 5) /* yhat holds logits; softmax(dim=1) is fused (compose.py:191-192) */
#define OCT_LOSS_WPCE_FULL   (1 << 6) /* kwargs['full']: do not mask yhat by ys (losses.py:31-32) */
#define OCT_LOSS_KLD_STOPGRAD (1 << 7) /* stop_gradient=True: basis gets no gradient (losses.py:114) */
#define OCT_LOSS_JSD         (1 << 8) /* with OCT_LOSS_KLD: divergence='JSD' (losses.py:154-169): mean_q = mean_k(w_k*up(att_k)),
                                         M = (b + mean_q)/2, 0.5*KL(b||M) + 0.5*KL(mean_q||M); generic (fp32) kernel */

#define OCT_LOSS_MAX_CLASSES 8
#define OCT_LOSS_MAX_ATT 5

typedef struct OctaveLossDesc {
  int32_t dtype;            /* storage type of yhat / ys / att and of their gradients */
  int32_t B, C, H, W;       /* full-resolution maps are [B,C,H,W] planar */
  int32_t flags;            /* OCT_LOSS_* */
  int32_t n_att;            /* attention maps incl. the basis att[0]; <2 disables KLD */
  int32_t att_h[OCT_LOSS_MAX_ATT];
  int32_t att_w[OCT_LOSS_MAX_ATT];
  float att_weight[OCT_LOSS_MAX_ATT - 1]; /* weight of att[1..]; 0 skips the level (losses.py:124-126) */
  float sum_weights;        /* divisor of the summed log-posterior (losses.py:135) */
  float wpce_scale;         /* 1/(B*H*W) for reduction='mean', 1 for 'sum' (losses.py:55) */
  float dice_eps;           /* DiceLoss.eps (losses.py:66) */
  int32_t n_real, n_fake;   /* number of discriminator logits in d_real / d_fake */
  float jsd_eps;            /* InterlayerDivergence.eps inside log(M + eps) of the JSD branch (losses.py:159) */
} OctaveLossDesc;

/* out[] slots written by octave_loss_fwd */
#define OCT_LOSS_OUT_WPCE 0
#define OCT_LOSS_OUT_DICE 1
#define OCT_LOSS_OUT_KLD 2
#define OCT_LOSS_OUT_LSG 3
#define OCT_LOSS_OUT_LSD 4
#define OCT_LOSS_OUT_NANFLAG 5 /* 1.0 when the divergence is NaN (losses.py:140-142) */
#define OCT_LOSS_OUT_SLOTS 8

/* bytes of the statistics workspace shared by fwd and bwd */
size_t octave_loss_stats_bytes(const OctaveLossDesc* d);
/* 1 when the vectorised C=2 pyramid kernel will be used, 0 when the generic kernel will (fp32 only) */
int octave_loss_uses_fast_path(const OctaveLossDesc* d);

int octave_loss_fwd(const OctaveLossDesc* d, const void* yhat, const void* ys,
                    const void* const* att /* [n_att] */, const float* d_real, const float* d_fake,
                    void* stats, float* out /* [OCT_LOSS_OUT_SLOTS] device */, void* stream);

/* gscale: device float[5], upstream gradient of each loss term in OCT_LOSS_OUT_* order.
 * Gradient buffers whose term is disabled may be NULL.  d_att[k] is written for every k < n_att
 * (zeros for skipped levels).  */
int octave_loss_bwd(const OctaveLossDesc* d, const void* yhat, const void* ys,
                    const void* const* att, const float* d_real, const float* d_fake,
                    const void* stats, const float* gscale, void* g_yhat, void* const* g_att,
                    float* g_real, float* g_fake, void* stream);

/* The two branches of WeightedPartialCE.forward that OctaScribbleNet does not take (segmentor/losses.py:40-49,56-59), fp32:
 *   mode 0  manual=False, C == 2: nn.CrossEntropyLoss()(z, ys[:,1].long()), z = y_hat*ys (y_hat when full) as logits
 *   mode 1  num_classes == 1    : nn.BCEWithLogitsLoss()(z, ys)
 * mean over B*H*W; scratch2: 2 doubles of device scratch; out: float[1]; gscale: device float[1] upstream gradient. */
int octave_wpce_alt_fwd(int32_t mode, const float* yhat, const float* ys, int32_t B, int32_t C, int32_t H, int32_t W,
                        int32_t full, double* scratch2, float* out, void* stream);
int octave_wpce_alt_bwd(int32_t mode, const float* yhat, const float* ys, int32_t B, int32_t C, int32_t H, int32_t W,
                        int32_t full, const float* gscale, float* g_yhat, void* stream);

/* Fused single pass of the G-step (training loop): loss VALUES and GRADIENTS from one sweep over the maps.
 * total = lambdas[0]*WPCE + lambdas[1]*KLD + lambdas[2]*LSG (host float[3]); the gradients of `total` w.r.t. yhat,
 * att[k] and d_fake are written in the same pass (a labels-only pre-pass supplies the class counts of losses.py:34-38),
 * out[0..5] as octave_loss_fwd, out[6] = total.  Supported: C == 2 pyramids (octave_loss_fused_supported), flags within
 * WPCE | KLD | LSG | FROM_LOGITS | WPCE_FULL.  Reductions are in a fixed order (bit-reproducible).
 * `stats` (octave_loss_fused_stats_bytes) must be ZERO before the first call; every evaluation leaves it zero again, so
 * consecutive evaluations need no memset between them (one workspace per stream).
 * octave_loss_scale_grads multiplies the written gradients by *g_total (device) and is a no-op when it is 1. */
int octave_loss_fused_supported(const OctaveLossDesc* d);
size_t octave_loss_fused_stats_bytes(const OctaveLossDesc* d);
int octave_loss_fused(const OctaveLossDesc* d, const void* yhat, const void* ys, const void* const* att,
                      const float* d_fake, const float* lambdas /* host [3] */, void* stats, float* out,
                      void* g_yhat, void* const* g_att, float* g_fake, void* stream);
int octave_loss_scale_grads(const OctaveLossDesc* d, const float* g_total /* device */, void* g_yhat,
                            void* const* g_att, float* g_fake, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K1/K2/K3 — convolutions.  Activations NHWC; a tensor argument is a *channel view*: base pointer,
 * `ld` = elements per pixel of the underlying buffer, `coff` = first channel of the view (this is how the
 * skip-concat of compose.py:141,155,162,169 is fused: producers write straight into channel slices).
 * Replaces nn.Conv2d / nn.ConvTranspose2d inside
 *   ResNet deep stem, Bottleneck, SplAtConv2d, ResNestDecoder, Upsampling  architectures/extra/resnest.py:18-138,170-267,326-334
 *   DiscriminatorBlock convs                                              architectures/discriminator/blocks.py:46-50,91-109
 * ---------------------------------------------------------------------------------------------- */
#define OCT_CONV_MODE_CONV 0   /* Conv2d, stride 1 (tc path) or general (direct path) */
#define OCT_CONV_MODE_CONVT 1  /* ConvTranspose2d k=2 s=2: GEMM + 2x2 pixel-shuffle store */

typedef struct OctaveConvDesc {
  int32_t B, H, W;        /* input pixel grid */
  int32_t cin, cout, groups;
  int32_t ksize;          /* square kernel (tc path: 1 or 3, padding ksize/2, stride 1) */
  int32_t stride, pad;    /* direct path only; the tc path requires stride 1, pad ksize/2 */
  int32_t x_ld, x_coff;   /* input view */
  int32_t y_ld, y_coff;   /* output view */
  int32_t Hout, Wout;     /* output pixel grid (ConvT: 2H,2W or cropped, compose.py:142-147) */
  int32_t mode;           /* OCT_CONV_MODE_* */
  int32_t relu;           /* fused activation: 0 none, 1 ReLU, 2 LeakyReLU(0.2), 3 sigmoid, 4 tanh */
  int32_t in_dtype;       /* direct path: storage type of x / w */
  int32_t out_dtype;      /* storage type of y */
  int32_t accumulate;     /* y += result (fwd/dgrad); dw += result (wgrad) instead of overwrite */
  int32_t real_groups;    /* wgrad: groups of the torch weight when `groups` was merged into dense groups (0 = groups) */
  int32_t out_s2d_qs;     /* tc fwd: >0 stores the output space-to-depth: pixel (h>>1,w>>1), channel ((h&1)*2+(w&1))*qs + c */
} OctaveConvDesc;

/* tcgen05 path (bf16 in, fp32 accumulate).  `wpack` is bf16 [taps][Cout][Cin/groups] (ConvT: [4*Cout][Cin],
 * row = (i*2+j)*Cout + co).  Data gradient = the same entry point with the dgrad pack and cin/cout swapped. */
int octave_conv_tc_supported(const OctaveConvDesc* d);
/* stats (nullable, fp64 [2*cout], overwritten): per-channel sum and sum of squares of the stored outputs — the batch
 * statistics of the BatchNorm that follows (resnest.py:25,86,182,224,338), fused into the conv epilogue. */
int octave_conv_tc_fwd(const OctaveConvDesc* d, const void* x, const void* wpack, const float* bias, void* y,
                       double* stats, void* stream);
/* Narrow-layer variant of octave_conv_tc_fwd (3x3 s1 p1, Cin/groups and Cout/groups in {32, 64}, all taps of all groups
 * <= 72 KB): weights resident in shared memory, one halo TMA box per output tile, tap shifts as descriptor offsets.
 * octave_conv_tc_fwd forwards to it when octave_conv_halo_supported(); same arguments and semantics.
 * octave_conv_halo_config(enabled, base_off_mode) overrides the OCTAVE_HALO / OCTAVE_HALO_BASEOFF environment switches. */
int octave_conv_halo_supported(const OctaveConvDesc* d);
int octave_conv_halo_fwd(const OctaveConvDesc* d, const void* x, const void* wpack, const float* bias, void* y,
                         double* stats, void* stream);
void octave_conv_halo_config(int32_t enabled, int32_t base_off_mode);
/* Narrow-layer variant of octave_conv_tc_wgrad (3x3 s1 p1, channels per group in {32, 64} with one side 32): one CTA
 * keeps all nine tap accumulators in TMEM and loads every dy / x patch once.  octave_conv_tc_wgrad forwards to it. */
int octave_conv_halo_wgrad_supported(const OctaveConvDesc* d);
int octave_conv_halo_wgrad(const OctaveConvDesc* d, const void* x, const void* dy, float* dw, void* stream);
/* dw: fp32 gradient in the torch parameter layout ([Cout][Cin/real_groups][k][k]; ConvT: [Cin][Cout][2][2]),
 * overwritten unless `accumulate`.  x = forward input view, dy = output-gradient view (x_* / y_* of the descriptor;
 * for ConvT dy is the space-to-depth view with 4*cout channels). */
int octave_conv_tc_wgrad_supported(const OctaveConvDesc* d);
int octave_conv_tc_wgrad(const OctaveConvDesc* d, const void* x, const void* dy, float* dw, void* stream);

/* CUDA-core direct convolution (any k / stride / pad / groups, any channel count; fp32 or bf16 storage, fp32
 * accumulate; weights fp32 in the torch layout).  Used for the fp32 mode, the 3-channel stem conv
 * (resnest.py:327) and the narrow discriminator convs (discriminator/blocks.py:46-50,91-109).
 * d->relu selects the fused activation: 0 none, 1 ReLU, 2 LeakyReLU(0.2), 3 sigmoid, 4 tanh. */
int octave_conv_direct_fwd(const OctaveConvDesc* d, const void* x, const float* w, const float* bias, void* y,
                           void* stream);
int octave_conv_direct_dgrad(const OctaveConvDesc* d, const void* dy, const float* w, void* dx, void* stream);
/* dw overwritten unless d->accumulate; dbias nullable */
int octave_conv_direct_wgrad(const OctaveConvDesc* d, const void* x, const void* dy, float* dw, float* dbias,
                             void* stream);

/* ------------------------------------------------------------------------------------------------
 * K4-K8 — bandwidth-bound glue.  An OctaveAct is an NHWC channel view: pixel p, channel c lives at
 * data + (p*ld + coff + c) elements.  C must be a multiple of 8 (16-byte vectors) unless stated.
 * ---------------------------------------------------------------------------------------------- */
typedef struct OctaveAct {
  void* data;
  int32_t B, H, W, C;
  int32_t ld, coff;
  int32_t dtype; /* OCT_DTYPE_* */
} OctaveAct;

/* dy <- dy * act'(y) for the activations fused by conv_direct_fwd (y is the stored activation output) */
int octave_act_bwd(const OctaveAct* y, const OctaveAct* dy, int32_t act, const OctaveAct* dz, void* stream);

/* K4 BatchNorm2d (train: batch statistics + running-stat update; eval: running stats) and K7 residual.
 *   nn.BatchNorm2d at resnest.py:25,35,86,101-105,182,224,338,394; residual add + ReLU resnest.py:42,264-265 */
int octave_chan_stats(const OctaveAct* x, double* sums /* [2C]: sum, sum of squares; overwritten */, void* stream);
int octave_bn_prepare(int32_t C, double count, const double* sums, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, int64_t* num_batches_tracked, float eps, float momentum,
                      int32_t training, float* ab /* [2C] scale, shift */, float* mean_invstd /* [2C] */, void* stream);
/* y = act(x*a[c] + b[c] + res); ab NULL => identity; gap (nullable, fp32 [B][C/2], overwritten):
 * gap[b][c % (C/2)] = sum_pixels y  — the radix-sum + global-average-pool of resnest.py:106-116, reduced in a fixed
 * order (bit-reproducible, no floating-point atomics) through gap_ws: octave_affine_gap_ws_bytes(x) bytes whose first
 * 65536 32-bit words (one counter per image) are zero before the first launch (the kernel leaves them zero). */
size_t octave_affine_gap_ws_bytes(const OctaveAct* x);
int octave_affine_act(const OctaveAct* x, const float* ab, const OctaveAct* res, int32_t relu, const OctaveAct* y,
                      float* gap, void* gap_ws, void* stream);
/* dz = dy * (mask > 0) (mask nullable); or, with mask NULL and relu_ab != NULL (the [2C] scale/shift of this very BN),
 * the ReLU mask is recomputed from x as (x*a+b > 0) instead of being read.  sums2[c] = sum dz, sums2[C+c] = sum dz * xhat.
 * dmasked (nullable): dz itself is stored too — it is the gradient of the residual branch added before the ReLU
 * (resnest.py:42,264-265), and octave_bn_bwd_apply can then be given dz as its dy with no mask (one tensor pass less). */
int octave_bn_bwd_reduce(const OctaveAct* dy, const OctaveAct* mask, const float* relu_ab, const OctaveAct* x,
                         const float* mean_invstd, double* sums2, const OctaveAct* dmasked, void* stream);
/* dx = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat)) (training) or gamma*invstd*dz (eval, mean_invstd then
 * holds the running statistics).  dgamma = sum dz*xhat, dbeta = sum dz (nullable) are overwritten.
 * dmasked (nullable): also store dz = dy * (mask > 0) itself — the gradient of the residual branch that was added
 * before the ReLU (resnest.py:42,264-265), saving the separate octave_relu_bwd pass over dy and the mask. */
int octave_bn_bwd_apply(const OctaveAct* dy, const OctaveAct* mask, const float* relu_ab, const OctaveAct* x,
                        const float* mean_invstd, const float* gamma, const double* sums2, int32_t training,
                        const OctaveAct* dx, float* dgamma, float* dbeta, const OctaveAct* dmasked, void* stream);
/* dst += src (same shape); used where gradients of two branches merge. */
int octave_add_inplace(const OctaveAct* dst, const OctaveAct* src, void* stream);
/* dst = src * (mask > 0) */
int octave_relu_bwd(const OctaveAct* dy, const OctaveAct* mask, const OctaveAct* dx, void* stream);

/* K5 split-attention (SplAtConv2d.forward resnest.py:106-138), radix 2.  U has 2C channels, out has C.
 *   out[p][c] = att[b][c]*U[p][c] + att[b][C+c]*U[p][C+c]  (+ReLU for the decoder, resnest.py:29) */
int octave_splat_combine(const OctaveAct* U, const float* att /* [B][2C] */, int32_t relu, const OctaveAct* out,
                         void* stream);
/* datt[b][r*C+c] = sum_p dout[p][c]*(mask[p][c]>0)*U[p][r*C+c]  (overwritten) */
int octave_splat_bwd_reduce(const OctaveAct* dout, const OctaveAct* mask, const OctaveAct* U, float* datt, void* stream);
/* dU[p][r*C+c] = att[b][r*C+c]*dout[p][c]*(mask>0) + dgap[b][c]*gap_scale */
int octave_splat_bwd_du(const OctaveAct* dout, const OctaveAct* mask, const float* att, const float* dgap,
                        float gap_scale, const OctaveAct* dU, void* stream);

/* Backward of SplAtConv2d's combine + bn0 + ReLU in two passes over z (resnest.py:101-105,133-135): the gradient of U,
 *   dU[p][r*C+c] = att[b][r*C+c] * dout[p][c] * (omask[p][c] > 0) + dgap[b][c] * gap_scale,
 * is rebuilt on the fly from dout instead of being written and re-read.  z: bn0 input [B,H,W,2C]; ab / mean_invstd: the
 * [2*2C] scale-shift and statistics octave_bn_prepare produced in the forward pass (the ReLU mask is z*a+b > 0).
 * Outputs: dz (gradient w.r.t. z), dgamma / dbeta [2C], sums2 (fp64 [2*2C] scratch, overwritten). */
int octave_splat_bn_bwd(const OctaveAct* dout, const OctaveAct* omask, const float* att, const float* dgap, float gap_scale,
                        const OctaveAct* z, const float* ab, const float* mean_invstd, const float* gamma, int32_t training,
                        double* sums2, const OctaveAct* dz, float* dgamma, float* dbeta, void* stream);

/* K6 pools (NHWC).  kind 0: MaxPool2d; 1: AvgPool2d.  Exact torch semantics incl. ceil_mode and
 * count_include_pad (resnest.py:340 maxpool 3/2/1; :189 avd AvgPool 3/s/1; :383 shortcut AvgPool s/s ceil, no pad count). */
typedef struct OctavePoolDesc {
  int32_t kind, k, stride, pad, ceil_mode, count_include_pad;
} OctavePoolDesc;
int octave_pool_out_size(const OctavePoolDesc* p, int32_t in);
/* argmax: uint8 [B][Ho][Wo][C] (max pool only; nullable for avg) */
int octave_pool_fwd(const OctavePoolDesc* p, const OctaveAct* x, const OctaveAct* y, uint8_t* argmax, void* stream);
int octave_pool_bwd(const OctavePoolDesc* p, const OctaveAct* dy, const uint8_t* argmax, const OctaveAct* dx,
                    void* stream);

/* K8 pointwise heads: 1x1 conv C -> K (K <= 8) with bias, emitting NCHW-planar fp32 maps.
 *   mode 0 (linear): out = W x + b                      — ResnestUNet.fc, compose.py:79,181
 *   mode 1 (gate)  : y_hat = softmax(W x + b); gated = x * sum_{k>=1} y_hat_k; out = y_hat
 *                                                       — AdversarialAttentionGate.forward, segmentor/blocks.py:38-46 */
int octave_head_fwd(const OctaveAct* x, const float* w /* [K][C] */, const float* b /* [K] */, int32_t K, int32_t mode,
                    float* out /* [B][K][H][W] */, const OctaveAct* gated /* mode 1 */, void* stream);
/* dout nullable (no gradient reached the map).  dw [K][C] / db [K] (both or neither, nullable): parameter gradients,
 * overwritten.  K == 2 with dlogits == NULL takes the fused kernel (dx, dw, db in one pass over x).  Otherwise dlogits
 * (fp32 [B][K][H][W] scratch, overwritten) is required and dw/db are produced through octave_head_wgrad. */
int octave_head_bwd(const OctaveAct* x, const float* w, const float* b, int32_t K, int32_t mode, const float* dout,
                    const OctaveAct* dgated, const OctaveAct* dx, float* dlogits, float* dw, float* db, void* stream);
int octave_head_wgrad(const OctaveAct* x, const float* dlogits, int32_t K, float* dw /* [K][C] */, float* db /* [K] */,
                      void* stream);

/* Layout plumbing. */
int octave_nchw_to_nhwc(const float* src, int32_t C_src, const OctaveAct* dst /* C >= C_src, extra channels zeroed */,
                        void* stream);
int octave_nhwc_to_nchw(const OctaveAct* src, float* dst, int32_t accumulate, void* stream);
/* InstanceNoise fused into the layout change (discriminator/blocks.py:149-154): dst = clip(src + noise[h][w], 0, 1);
 * noise (fp32 [H][W], one plane broadcast over batch and channel) and clip are optional. */
int octave_nchw_to_nhwc_noise(const float* src, int32_t C_src, const float* noise, int32_t clip, const OctaveAct* dst,
                              void* stream);
/* its backward: dst = src * 1[0 <= x + noise <= 1] (x: the forward input, NCHW fp32) */
int octave_nhwc_to_nchw_clipmask(const OctaveAct* src, const float* x, const float* noise, int32_t clip, float* dst,
                                 void* stream);
/* dst(h,w) = [accumulate ? dst : 0] + (h < src.H && w < src.W ? src(h,w) : 0): zero-pad (compose.py:125-130) and crop. */
int octave_copy_window(const OctaveAct* src, const OctaveAct* dst, int32_t accumulate, void* stream);
/* space-to-depth by 2: dst[h][w][(i*2+j)*C + c] = src[2h+i][2w+j][c] (0 outside src) — data-gradient view of ConvT k2s2.
 * chan_sum (nullable, fp64 [C], overwritten): per-channel sum of src = the ConvTranspose2d bias gradient
 * (Upsampling.up.bias, resnest.py:50), fused because this pass reads every element of the output gradient once. */
int octave_space_to_depth(const OctaveAct* src, const OctaveAct* dst, double* chan_sum, void* stream);
/* inverse: dst[2h+i][2w+j][c] = src[h][w][(i*2+j)*C + c] for the pixels that exist in dst (fp32-mode ConvT k2s2) */
int octave_depth_to_space(const OctaveAct* src, const OctaveAct* dst, void* stream);

/* K5 (small part): the per-sample attention branch of SplAtConv2d on [B][C] vectors, fp32
 * (resnest.py:116-127: GAP -> fc1 (grouped 1x1) -> BatchNorm over the batch -> ReLU -> fc2 -> view(B,radix,C) -> softmax(dim=1)). */
/* out[b][j] = bias[j] + in_scale * sum_i in[b][g(j)*Kg + i] * w[j][i],  g(j) = j / (N/groups) */
int octave_glinear_fwd(const float* in, const float* w, const float* bias, int32_t B, int32_t K_total, int32_t N,
                       int32_t groups, float in_scale, float* out, void* stream);
/* din[b][i] = in_scale * sum_{j in group(i)} dout[b][j] * w[j][i - g*Kg]   (overwritten) */
int octave_glinear_bwd_data(const float* dout, const float* w, int32_t B, int32_t K_total, int32_t N, int32_t groups,
                            float in_scale, float* din, void* stream);
/* dw[j][i] = in_scale * sum_b dout[b][j] * in[b][g*Kg+i]; dbias[j] = sum_b dout[b][j]   (overwritten) */
int octave_glinear_bwd_weight(const float* dout, const float* in, int32_t B, int32_t K_total, int32_t N, int32_t groups,
                              float in_scale, float* dw, float* dbias, void* stream);
/* BatchNorm over the batch of a [B][C] matrix followed by ReLU (SplAtConv2d.bn1 + relu, resnest.py:120-122) */
int octave_bn1d_relu_fwd(const float* x, int32_t B, int32_t C, const float* gamma, const float* beta, float* running_mean,
                         float* running_var, int64_t* num_batches_tracked, float eps, float momentum, int32_t training,
                         float* y, float* mean_invstd /* [2C] */, void* stream);
int octave_bn1d_relu_bwd(const float* dy, const float* x, const float* y, int32_t B, int32_t C, const float* gamma,
                         const float* mean_invstd, int32_t training, float* dx, float* dgamma, float* dbeta, void* stream);
/* The attention branch of SplAtConv2d (resnest.py:116-127) in two launches per direction instead of four, for a batch that
 * fits one 32-row slab (octave_attn_fused_supported: B <= 32, one group, radix 2):
 *   octave_glinear_bn_relu_fwd      x_out = fc1(in * in_scale), y = relu(BatchNorm1d(x_out)) (+ running statistics), mean_invstd
 *   octave_glinear_rsoftmax_fwd     att = softmax over the radix pair (c, C + c) of fc2(in); the logits are not stored
 *   octave_rsoftmax_glinear_bn_bwd  dlogits = r-softmax backward of datt; dh = dlogits . W2; dx = BatchNorm1d+ReLU backward of
 *                                   dh (+ dgamma, dbeta).  W2 is [2C][Kt]; x / y are fc1's output before / after bn1 + relu.
 * Same summation orders as the unfused entry points: results are bit-identical to them. */
int octave_attn_fused_supported(int32_t B, int32_t C, int32_t inter, int32_t groups, int32_t radix);
int octave_glinear_bn_relu_fwd(const float* in, const float* w, const float* bias, int32_t B, int32_t K_total, int32_t N,
                               float in_scale, const float* gamma, const float* beta, float* running_mean, float* running_var,
                               int64_t* num_batches_tracked, float eps, float momentum, int32_t training, float* x_out, float* y,
                               float* mean_invstd /* [2N] */, void* stream);
int octave_glinear_rsoftmax_fwd(const float* in, const float* w, const float* bias, int32_t B, int32_t K_total, int32_t C,
                                float* att /* [B][2C] */, void* stream);
int octave_rsoftmax_glinear_bn_bwd(const float* datt, const float* att, const float* w, int32_t B, int32_t K_total, int32_t C,
                                   const float* x, const float* y, const float* gamma, const float* mean_invstd, int32_t training,
                                   float* dlogits, float* dx, float* dgamma, float* dbeta, void* stream);
/* att[b][r*C+c] = softmax_r(logits[b][r*C+c]), radix R */
int octave_rsoftmax_fwd(const float* logits, int32_t B, int32_t R, int32_t C, float* att, void* stream);
int octave_rsoftmax_bwd(const float* datt, const float* att, int32_t B, int32_t R, int32_t C, float* dlogits, void* stream);

/* Space-to-depth formulation of the discriminator's 4x4 stride-2 pad-1 convs (discriminator/blocks.py:46-50,91-109):
 * with X'[h'][w'][(i*2+j)*qs + c] = X[2h'+i][2w'+j][c] the conv becomes a 3x3 stride-1 pad-1 conv over 4*qs channels
 * (tap dh' in {-1,0,1}, parity i: kh = 2dh'+i+1 if 0<=kh<4), which runs on the tcgen05 kernel. */
/* Writes channels [coff, qs) of every quadrant of dst: the C source channels (+ noise[h][w], clipped to [0,1] if clip),
 * then zeros; quadrant pixels outside the source (odd H / W) are written as zeros.  With coff = 0 and dst->C = 4*qs the
 * whole tensor is written: no memset needed. */
int octave_nchw_to_s2d(const float* src, int32_t B, int32_t C, int32_t H, int32_t W, const float* noise, int32_t clip,
                       const OctaveAct* dst /* [B][ceil(H/2)][ceil(W/2)][4*qs] */, int32_t qs, int32_t coff, void* stream);
/* dst[b][c][h][w] = src quadrant channel coff+c (optionally masked by 1[0 <= x+noise <= 1], the clip backward) */
int octave_s2d_to_nchw(const OctaveAct* src, int32_t qs, int32_t coff, int32_t C, int32_t H, int32_t W, const float* x,
                       const float* noise, int32_t clip, float* dst, void* stream);
/* The same remap serves the stem's 3x3 stride-2 pad-1 conv (resnest.py:327): ksize 3 (valid kh in 0..2) or 4.
 * w fp32 [cout][cin][k][k] (* scale[0] if scale != NULL) -> bf16 operand: mode 0 [9][cout][4*qs]; mode 1 (dgrad) [9][4*qs][cout] */
int octave_pack_weight_s2d(const float* w, const float* scale, int32_t mode, int32_t cout, int32_t cin, int32_t qs,
                           int32_t ksize, void* out_bf16, void* stream);
/* dw3 fp32 [cout][4*qs][3][3] (gradient of the remapped weight) -> dw fp32 [cout][cin][k][k] */
int octave_unpack_wgrad_s2d(const float* dw3, int32_t cout, int32_t cin, int32_t qs, int32_t ksize, float* dw, void* stream);
/* Full-extent output conv of the critic (blocks.py:68-71) as one dot product per sample:
 * out[b] = bias + sum_i x[b][i] * w[i]   (x: NHWC activation flattened per sample, w fp32 in the same order) */
int octave_rowdot_fwd(const OctaveAct* x, const float* w, const float* bias, float* out /* [B] */, void* stream);
/* dx[b][i] = g[b] * w[i];  dw[i] = sum_b g[b] * x[b][i];  dbias = sum_b g[b]   (dw/dbias nullable) */
int octave_rowdot_bwd(const OctaveAct* x, const float* w, const float* g, const OctaveAct* dx, float* dw, float* dbias,
                      void* stream);

/* Weight re-packing fp32 [Cout][Cin/groups][k][k] (torch layout) -> bf16 operand of the tcgen05 kernels.
 * `dense_groups` <= groups: groups are merged into block-diagonal dense groups (zeros off the diagonal) when the
 * per-group channel count is too small for a UMMA tile. */
#define OCT_PACK_FWD 0          /* [taps][Cout][Cin/dense_groups] */
#define OCT_PACK_DGRAD 1        /* [taps][Cin][Cout/dense_groups], taps flipped */
#define OCT_PACK_CONVT_FWD 2    /* w [Cin][Cout][2][2] -> [4*Cout][Cin] */
#define OCT_PACK_CONVT_DGRAD 3  /* w [Cin][Cout][2][2] -> [Cin][4*Cout] */
int octave_pack_weight(const float* w, int32_t mode, int32_t cout, int32_t cin, int32_t groups, int32_t dense_groups,
                       int32_t ksize, void* out_bf16, void* stream);
/* Every operand pack of a network in ONE launch (the per-step re-pack after the optimiser update).
 * jobs_device: device array sorted by block_start; job i owns the blocks [block_start_i, block_start_i +
 * octave_pack_job_blocks(...)); total_blocks = their sum. */
typedef struct OctavePackJob {
  const float* w;        /* fp32 weight, torch layout */
  void* out;             /* bf16 operand pack */
  int32_t mode, cout, cin, groups, dense_groups, ksize;
  int64_t block_start;
} OctavePackJob;
int64_t octave_pack_job_blocks(int32_t mode, int32_t cout, int32_t cin, int32_t dense_groups, int32_t ksize);
int octave_pack_weight_multi(const OctavePackJob* jobs_device, int32_t n_jobs, int64_t total_blocks, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-tensor optimiser step: every parameter of a module in ONE launch.  The reference ships no
 * optimiser (README.md:39-47: the training script lives on another branch); the update rules are torch.optim.SGD
 * (momentum, weight decay; dampening 0, no Nesterov) and torch.optim.AdamW (no amsgrad).  jobs_device: device array
 * sorted by block_start; job i owns octave_optim_job_blocks(n_i) blocks.  m / v: fp32 state of the same size as w
 * (v unused by SGD; with first_step != 0 SGD initialises m with the gradient, as torch does).  bias_correction{1,2} =
 * 1 - beta^t are supplied by the host. */
#define OCT_OPT_SGD 0
#define OCT_OPT_ADAMW 1
typedef struct OctaveOptJob {
  float* w;
  const float* g;
  float* m;
  float* v;
  int64_t n;
  int64_t block_start;
} OctaveOptJob;
typedef struct OctaveOptHyper {
  int32_t algo, first_step;
  float lr, momentum, weight_decay, beta1, beta2, eps, bias_correction1, bias_correction2;
} OctaveOptHyper;
int64_t octave_optim_job_blocks(int64_t n);
int octave_optim_multi(const OctaveOptJob* jobs_device, int32_t n_jobs, int64_t total_blocks, const OctaveOptHyper* hyper,
                       void* stream);

/* Data-parallel gradient buckets in bf16 (half the all-reduce bytes).  pack: flat[flat_off + i] = bf16(g[i]) for every
 * job — the gather-and-cast of a bucket's fp32 gradient tensors into one flat bf16 buffer; unpack: g[i] = float(flat[..]),
 * the averaged values written back into the gradient tensors.  jobs: HOST array, at most OCTAVE_GRAD_MAX_JOBS per call
 * (passed to the kernel by value: no table upload, capture-safe); flat_off in elements, multiples of 8;
 * block_start = sum of octave_optim_job_blocks(n) of the jobs before. */
#define OCTAVE_GRAD_MAX_JOBS 128
typedef struct OctaveGradJob {
  float* g;
  int64_t flat_off;
  int32_t n;
  int32_t block_start;
} OctaveGradJob;
int octave_grad_pack_bf16(const OctaveGradJob* jobs, int32_t n_jobs, void* flat_bf16, void* stream);
int octave_grad_unpack_bf16(const OctaveGradJob* jobs, int32_t n_jobs, const void* flat_bf16, void* stream);

/* ------------------------------------------------------------------------------------------------
 * On-GPU input pipeline.  The reference ships no data loader (README.md:39-47); the synthetic OCTA model
 * is the one of SURVEY.md 8d.  Counter-based random numbers: a batch is a pure function of `seed`.
 *   octave_synth_octa          x [B,3,H,W] fp32 in [0,1] (one plane replicated), ys [B,2,H,W] one-hot scribbles / all-zero,
 *                              vessel (nullable) [B,H,W] u8 ground truth
 *   octave_synth_mask_pyramid  out[k] = [B,2,ceil(H/2^k),ceil(W/2^k)] one-hot "real" masks (unpaired draw), k < levels
 *   octave_augment             per-sample flips / 90-degree rotations applied to x AND ys, photometric jitter on x only;
 *                              out-of-place (x_out != x) */
#define OCT_AUG_FLIP_H (1 << 0)
#define OCT_AUG_FLIP_V (1 << 1)
#define OCT_AUG_ROT90  (1 << 2)
#define OCT_AUG_PHOTO  (1 << 3)
int octave_synth_octa(uint64_t seed, int32_t B, int32_t H, int32_t W, int32_t n_ridges, float* x, float* ys, uint8_t* vessel,
                      void* stream);
int octave_synth_mask_pyramid(uint64_t seed, int32_t B, int32_t H, int32_t W, int32_t n_ridges, int32_t levels,
                              float* const* out /* host array of device pointers */, void* stream);
int octave_augment(uint64_t seed, int32_t B, int32_t C, int32_t Cy, int32_t H, int32_t W, int32_t flags, const float* x,
                   const float* ys, float* x_out, float* ys_out, void* stream);

/* Spectral norm of the critic's 4x4 convs (torch.nn.utils.spectral_norm legacy semantics, discriminator/blocks.py:101-104):
 * one power iteration (training != 0: u, v updated in place) and sigma = u . (W v) in one launch, fixed summation order.
 * W: weight_orig as [rows][cols] fp32, rows <= 1024; out2: device float[2] = sigma, 1/sigma. */
int octave_spectral_sigma(const float* W, int32_t rows, int32_t cols, float* u, float* v, int32_t training, float eps,
                          float* out2, void* stream);
/* The same for every spectral-norm layer of one critic call in ONE launch (the layers' power iterations depend on the
 * weights only).  jobs: host array; out: device float[2 + rows + cols] = sigma, 1/sigma, then the u and the v this call
 * leaves behind (a snapshot for the backward pass: later calls keep iterating u / v in place). */
#define OCTAVE_SN_MAX_JOBS 8
typedef struct OctaveSnJob {
  const float* W;
  float* u;
  float* v;
  float* out;
  int32_t rows, cols;
} OctaveSnJob;
int octave_spectral_sigma_multi(const OctaveSnJob* jobs, int32_t n_jobs, int32_t training, float eps, void* stream);
/* Weight gradient through the normalisation W = W_orig / sigma, sigma = u^T W_orig v with u, v constants (what autograd
 * does for torch.nn.utils.spectral_norm): out (+)= dW / sigma - <dW, W_orig> / sigma^2 * u v^T.  dw, w_orig, out:
 * [rows][cols] fp32; sigma: device float; parts: device float[64] scratch.  Two launches, fixed summation order. */
int octave_spectral_wgrad(const float* dw, const float* w_orig, const float* u, const float* v, const float* sigma,
                          int32_t rows, int32_t cols, float* parts, float* out, int32_t accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OCTAVE_B200_H_ */
