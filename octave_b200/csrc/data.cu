// On-GPU input pipeline: synthetic OCTA-like batches and augmentation written straight into device
// buffers, so that a multi-GPU run is not bound by host data generation / H2D.  The reference ships no data loader (its
// dataset preparation lives on an unmounted branch, README.md:39-47); the synthetic model is the one of SURVEY.md §8d and
// octave_b200/synth.py: en face image = 0.3*uniform background + max over ~n sinusoidal ridges exp(-(d/width)^2), replicated
// to 3 channels; scribbles = sparse one-hot strokes on ridge cores / off ridges, all-zero elsewhere (unlabelled); "real"
// masks = thresholded ridges of an unpaired draw, nearest-downsampled.  Random numbers are counter-based (a hash of seed,
// sample, index), so a batch is a pure function of its seed — reproducible and independent of the launch geometry.
#include "common.cuh"
#include "../../include/octave_b200.h"

namespace {

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ float urand(uint64_t seed, uint32_t a, uint32_t b, uint32_t c) {
  uint32_t h = hash32((uint32_t)seed ^ hash32(a + 0x9e3779b9u * (uint32_t)(seed >> 32)));
  h = hash32(h ^ hash32(b + 0x85ebca6bu));
  h = hash32(h ^ hash32(c + 0xc2b2ae35u));
  return (float)(h >> 8) * (1.0f / 16777216.0f);
}

constexpr int kMaxRidges = 64;
struct Ridge { float c, s, off, amp, freq, inv_w, phase; };

__device__ __forceinline__ void make_ridges(uint64_t seed, int b, int n, float extent, Ridge* r) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float theta = urand(seed, b, i, 0) * 3.14159265f;
    Ridge k;
    k.c = cosf(theta); k.s = sinf(theta);
    k.off = urand(seed, b, i, 1) * extent;
    k.amp = 4.f + 20.f * urand(seed, b, i, 2);
    k.freq = 0.01f + 0.04f * urand(seed, b, i, 3);
    k.inv_w = 1.f / (0.8f + 2.5f * urand(seed, b, i, 4));
    k.phase = 6.28f * urand(seed, b, i, 5);
    r[i] = k;
  }
}
__device__ __forceinline__ float ridge_field(const Ridge* r, int n, float x, float y) {
  float acc = 0.f;
  for (int i = 0; i < n; ++i) {
    const Ridge k = r[i];
    const float u = x * k.c + y * k.s, v = y * k.c - x * k.s;
    const float d = (u - k.off - k.amp * __sinf(k.freq * v + k.phase)) * k.inv_w;
    acc = fmaxf(acc, __expf(-d * d));
  }
  return acc;
}

// grid (ceil(H*W/256), B)
__global__ void __launch_bounds__(256) synth_octa_kernel(uint64_t seed, int H, int W, int n_ridges, float* __restrict__ x,
                                                         float* __restrict__ ys, uint8_t* __restrict__ vessel) {
  __shared__ Ridge rd[kMaxRidges];
  const int b = blockIdx.y;
  make_ridges(seed, b, n_ridges, (float)max(H, W), rd);
  __syncthreads();
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= H * W) return;
  const int py = i / W, px = i - py * W;
  const float rf = ridge_field(rd, n_ridges, (float)px, (float)py);
  const float bg = 0.3f * urand(seed, b, i, 101);
  const float gain = 0.5f + 0.5f * urand(seed, b, 0xffffffu, 102);
  const float img = fminf(fmaxf(bg + rf * gain, 0.f), 1.f);
  const size_t plane = (size_t)H * W;
  float* xb = x + (size_t)b * 3 * plane + i;
  xb[0] = img; xb[plane] = img; xb[2 * plane] = img;
  const bool ves = rf > 0.5f;
  const float pick = urand(seed, b, i, 103);
  const bool fg = ves && rf > 0.9f && pick < 0.35f;
  const bool bgs = !ves && rf < 0.05f && pick < 0.04f;
  float* yb = ys + (size_t)b * 2 * plane + i;
  yb[0] = bgs ? 1.f : 0.f;
  yb[plane] = fg ? 1.f : 0.f;
  if (vessel) vessel[(size_t)b * plane + i] = ves ? 1 : 0;
}

// one launch per pyramid level: out [B,2,h,w], pixel (y,x) samples the full-resolution field at (y<<k, x<<k)
__global__ void __launch_bounds__(256) synth_mask_kernel(uint64_t seed, int h, int w, int shift, int n_ridges, float extent,
                                                         float* __restrict__ out) {
  __shared__ Ridge rd[kMaxRidges];
  const int b = blockIdx.y;
  make_ridges(seed, b, n_ridges, extent, rd);
  __syncthreads();
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= h * w) return;
  const int py = i / w, px = i - py * w;
  const bool ves = ridge_field(rd, n_ridges, (float)(px << shift), (float)(py << shift)) > 0.5f;
  const size_t plane = (size_t)h * w;
  float* o = out + (size_t)b * 2 * plane + i;
  o[0] = ves ? 0.f : 1.f;
  o[plane] = ves ? 1.f : 0.f;
}

// per-sample geometric transform (flips, 90-degree rotations of square maps) applied to image AND labels, photometric
// jitter (gain, gamma, additive noise) applied to the image only.  grid (ceil(H*W/256), B)
__global__ void __launch_bounds__(256) augment_kernel(uint64_t seed, int C, int Cy, int H, int W, int flags, const float* __restrict__ x,
                                                      const float* __restrict__ ys, float* __restrict__ xo, float* __restrict__ yo) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= H * W) return;
  const bool fx = (flags & OCT_AUG_FLIP_H) && urand(seed, b, 1, 201) < 0.5f;
  const bool fy = (flags & OCT_AUG_FLIP_V) && urand(seed, b, 2, 201) < 0.5f;
  const int rot = ((flags & OCT_AUG_ROT90) && H == W) ? (int)(urand(seed, b, 3, 201) * 4.f) & 3 : 0;
  const float gain = (flags & OCT_AUG_PHOTO) ? 0.8f + 0.4f * urand(seed, b, 4, 201) : 1.f;
  const float gamma = (flags & OCT_AUG_PHOTO) ? 0.8f + 0.45f * urand(seed, b, 5, 201) : 1.f;
  const float sigma = (flags & OCT_AUG_PHOTO) ? 0.02f : 0.f;
  int py = i / W, px = i - py * W;
  // destination (py, px) <- source (sy, sx)
  int sy = py, sx = px;
  for (int r = 0; r < rot; ++r) { const int t = sy; sy = sx; sx = W - 1 - t; }
  if (fx) sx = W - 1 - sx;
  if (fy) sy = H - 1 - sy;
  const size_t plane = (size_t)H * W, so = (size_t)sy * W + sx;
  // Box-Muller from two counter-based uniforms (one draw per pixel, shared by the replicated channels)
  const float u1 = fmaxf(urand(seed, b, i, 202), 1e-7f), u2 = urand(seed, b, i, 203);
  const float nz = sigma * sqrtf(-2.f * __logf(u1)) * __cosf(6.2831853f * u2);
  for (int c = 0; c < C; ++c) {
    const float v = x[((size_t)b * C + c) * plane + so];
    xo[((size_t)b * C + c) * plane + i] = (flags & OCT_AUG_PHOTO) ? fminf(fmaxf(gain * __powf(fmaxf(v, 0.f), gamma) + nz, 0.f), 1.f) : v;
  }
  for (int c = 0; c < Cy; ++c) yo[((size_t)b * Cy + c) * plane + i] = ys[((size_t)b * Cy + c) * plane + so];
}

}  // namespace

extern "C" int octave_synth_octa(uint64_t seed, int32_t B, int32_t H, int32_t W, int32_t n_ridges, float* x, float* ys,
                                 uint8_t* vessel, void* stream) {
  if (!x || !ys || B <= 0 || H <= 0 || W <= 0 || n_ridges <= 0 || n_ridges > kMaxRidges || B > 65535) return OCT_ERR_INVALID;
  synth_octa_kernel<<<dim3((H * W + 255) / 256, B), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(seed, H, W, n_ridges, x, ys, vessel);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_synth_mask_pyramid(uint64_t seed, int32_t B, int32_t H, int32_t W, int32_t n_ridges, int32_t levels,
                                         float* const* out, void* stream) {
  if (!out || B <= 0 || H <= 0 || W <= 0 || n_ridges <= 0 || n_ridges > kMaxRidges || levels <= 0 || levels > 8 || B > 65535)
    return OCT_ERR_INVALID;
  for (int k = 0; k < levels; ++k) {
    if (!out[k]) return OCT_ERR_INVALID;
    const int h = (H + (1 << k) - 1) >> k, w = (W + (1 << k) - 1) >> k;     // == len(range(0, H, 2^k)): strided slicing
    synth_mask_kernel<<<dim3((h * w + 255) / 256, B), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(seed, h, w, k, n_ridges,
                                                                                                   (float)max(H, W), out[k]);
    OCT_CHECK_LAUNCH();
  }
  return OCT_OK;
}

extern "C" int octave_augment(uint64_t seed, int32_t B, int32_t C, int32_t Cy, int32_t H, int32_t W, int32_t flags, const float* x,
                              const float* ys, float* x_out, float* ys_out, void* stream) {
  if (!x || !ys || !x_out || !ys_out || x == x_out || ys == ys_out || B <= 0 || C <= 0 || Cy <= 0 || H <= 0 || W <= 0 || B > 65535)
    return OCT_ERR_INVALID;
  augment_kernel<<<dim3((H * W + 255) / 256, B), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(seed, C, Cy, H, W, flags, x, ys, x_out, ys_out);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}
