/*
 * octave_b200 — C-ABI of the B200-native kernels behind the OCTAve training step.
 *
 * The reference (IoBT-VISTEC/OCTAve, /root/reference) has no FFI of its own: its boundary is the
 * Python class surface.  This header is the C-ABI that sits directly beneath that
 * surface; each entry point names the reference function whose arithmetic it replaces.  The Python
 * host (octave_b200/*.py, mirroring architectures/*.py of the reference) binds these symbols with
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

#define OCTAVE_ABI_VERSION 1
int octave_abi_version(void);
/* number of SMs of the current device (grid sizing); <0 on error */
int octave_sm_count(void);

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
  int32_t relu;           /* fuse ReLU into the epilogue */
  int32_t in_dtype;       /* direct path: storage type of x / w */
  int32_t out_dtype;      /* storage type of y */
} OctaveConvDesc;

/* tcgen05 path (bf16 in, fp32 accumulate).  `wpack` is bf16 [taps][Cout][Cin/groups] (ConvT: [4*Cout][Cin],
 * row = (i*2+j)*Cout + co).  Data gradient = the same entry point with the dgrad pack and cin/cout swapped. */
int octave_conv_tc_supported(const OctaveConvDesc* d);
int octave_conv_tc_fwd(const OctaveConvDesc* d, const void* x, const void* wpack, const float* bias, void* y,
                       void* stream);
/* dwpack: fp32 [taps][Cout][Cin/groups], overwritten.  x = forward input view, dy = output-gradient view
 * (x_* / y_* of the descriptor). */
int octave_conv_tc_wgrad_supported(const OctaveConvDesc* d);
int octave_conv_tc_wgrad(const OctaveConvDesc* d, const void* x, const void* dy, float* dwpack, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OCTAVE_B200_H_ */
