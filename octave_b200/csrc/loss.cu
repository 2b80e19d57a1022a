// K9 — fused loss kernels.
//
// One statistics pass + one gradient pass over the loss maps replace the reference's chains of
// elementwise/reduce kernels:
//   WeightedPartialCE   /root/reference/architectures/segmentor/losses.py:26-61
//   DiceLoss            .../segmentor/losses.py:70-74
//   InterlayerDivergence (KLD/mean)  .../segmentor/losses.py:111-147
//   LSGeneratorLoss / LSDiscriminatorialLoss  .../discriminator/losses.py:11-24
//
// Layout: maps are NCHW planar exactly as the reference holds them.  Fast path: C == 2, H,W % 16 == 0,
// attention pyramid att[k] = [B,2,H>>k,W>>k]; one warp owns a 16x16 full-resolution cell (lane ->
// row l>>1, 8 consecutive pixels), so every load is a 16/32-byte vector, the nearest-neighbour
// upsample (losses.py:126) is an index shift, and the coarse-level gradients are 2^k x 2^k box sums
// done with warp shuffles — no atomics on the maps.  Generic path (any C <= 8, any map sizes, fp32):
// one thread per pixel, fp32 atomics for coarse gradients.
//
// Algorithmic traffic (C=2): forward reads 6.664 elements/pixel, backward re-reads them and writes
// 4.664 => 17.99 elements/pixel.
#include <cooperative_groups.h>

#include "common.cuh"
#include "../../include/octave_b200.h"

namespace {

constexpr int ST_N = 0;        // n_c      [8]
constexpr int ST_S = 8;        // S_c      [8]  sum ys_c * log(arg_c + eps)
constexpr int ST_KLD = 16;
constexpr int ST_LSG = 17;     // sum (f-1)^2
constexpr int ST_LSDR = 18;    // sum (r-1)^2
constexpr int ST_LSDF = 19;    // sum (f+1)^2
constexpr int ST_COUNTER = 20; // unsigned long long
constexpr int ST_W = 24;       // class weights w_c [8], written by the finalising block
constexpr int ST_DICE = 32;    // I_b [B], Card_b [B]

constexpr float kEps = 1e-12f;  // literal in losses.py:37,52,112,135 (not self.eps)

struct LossArgs {
  const void* yhat;
  const void* ys;
  const void* att[OCT_LOSS_MAX_ATT];
  const float* d_real;
  const float* d_fake;
  double* stats;
  float* out;
  // backward only
  const float* gscale;
  void* g_yhat;
  void* g_att[OCT_LOSS_MAX_ATT];
  float* g_real;
  float* g_fake;
  int B, C, H, W, flags, n_att;
  int ah[OCT_LOSS_MAX_ATT], aw[OCT_LOSS_MAX_ATT];
  float aw8[OCT_LOSS_MAX_ATT - 1];
  float inv_sumw, wpce_scale, dice_eps, jsd_eps;
  int n_real, n_fake;
  float lam_wpce, lam_kld, lam_lsg;   // fused single pass: weights of the terms inside `total` (= upstream gradients)
};

__device__ __forceinline__ void atomic_add_f64(double* p, float v) { atomicAdd(p, (double)v); }

// Sum of LS-GAN squared errors, computed by one block (strided over threads).
__device__ __forceinline__ void ls_partial(const LossArgs& a, float& lsg, float& lsdr, float& lsdf) {
  lsg = lsdr = lsdf = 0.f;
  if (a.flags & (OCT_LOSS_LSG | OCT_LOSS_LSD)) {
    for (int i = threadIdx.x; i < a.n_fake; i += blockDim.x) {
      float f = a.d_fake[i];
      lsg += (f - 1.f) * (f - 1.f);
      lsdf += (f + 1.f) * (f + 1.f);
    }
  }
  if (a.flags & OCT_LOSS_LSD) {
    for (int i = threadIdx.x; i < a.n_real; i += blockDim.x) {
      float r = a.d_real[i];
      lsdr += (r - 1.f) * (r - 1.f);
    }
  }
}

// Executed by the last block to retire: turns the accumulated sums into the five loss scalars.
__device__ void finalize(const LossArgs& a) {
  volatile double* st = a.stats;
  float* out = a.out;
  const double npix = (double)a.B * a.H * a.W;
  double wpce = 0.0, dice = 0.0, kld = 0.0, lsg = 0.0, lsd = 0.0;
  if (a.flags & OCT_LOSS_WPCE) {
    // ni, n_tot, weights: losses.py:34-38 (fp32 in the reference; the sums are exact integers there)
    float ntot = 0.f;
    for (int c = 0; c < a.C; ++c) ntot += (float)st[ST_N + c];
    double acc = 0.0;
    for (int c = 0; c < a.C; ++c) {
      float w = ntot / ((float)st[ST_N + c] + kEps);
      st[ST_W + c] = (double)w;
      acc += (double)w * st[ST_S + c];
    }
    wpce = -acc * (double)a.wpce_scale;
  }
  if (a.flags & OCT_LOSS_DICE) {
    for (int b = 0; b < a.B; ++b) {
      double I = st[ST_DICE + b], card = st[ST_DICE + a.B + b];
      dice += 1.0 - 2.0 * I / (card + (double)a.dice_eps);
    }
    dice /= (double)a.B;
  }
  if (a.flags & OCT_LOSS_KLD) kld = st[ST_KLD] / npix;
  if (a.flags & OCT_LOSS_LSG) lsg = 0.5 * st[ST_LSG] / (double)a.n_fake;
  if (a.flags & OCT_LOSS_LSD)
    lsd = 0.5 * st[ST_LSDR] / (double)a.n_real + 0.5 * st[ST_LSDF] / (double)a.n_fake;
  out[OCT_LOSS_OUT_WPCE] = (float)wpce;
  out[OCT_LOSS_OUT_DICE] = (float)dice;
  out[OCT_LOSS_OUT_KLD] = (float)kld;
  out[OCT_LOSS_OUT_LSG] = (float)lsg;
  out[OCT_LOSS_OUT_LSD] = (float)lsd;
  out[OCT_LOSS_OUT_NANFLAG] = (kld != kld) ? 1.f : 0.f;
  out[6] = 0.f;
  out[7] = 0.f;
}

__device__ __forceinline__ void retire_block(const LossArgs& a) {
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned long long total = (unsigned long long)gridDim.x * gridDim.y;
    unsigned long long prev =
        atomicAdd(reinterpret_cast<unsigned long long*>(a.stats + ST_COUNTER), 1ULL);
    s_last = (prev == total - 1ULL);
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    finalize(a);
  }
}

// ------------------------------------------------------------------------------------------------
// Fast path, C == 2.
// ------------------------------------------------------------------------------------------------
template <typename T>
struct CellLoad {
  float p0[8], p1[8];  // yhat as probabilities
  float t0[8], t1[8];  // ys
  float b0[8], b1[8];  // basis attention
  float q1[2][4], q2[2][2], q3[2], q4[2];  // coarse attentions
  float m0[8], m1[8];  // sum_k log(w_k q_k + eps)
};

__device__ __forceinline__ void softmax2(float& z0, float& z1) {
  // softmax over dim=1 with C=2 in the max-subtracted form torch uses: the larger logit maps to exp(0) = 1, so one
  // exponential and one reciprocal per pixel suffice (the kernel is MUFU-bound next to HBM-bound otherwise)
  const float d = z1 - z0;
  const float e = __expf(-fabsf(d));
  const float inv = __fdividef(1.f, 1.f + e);
  const float big = inv, small = e * inv;
  z0 = d > 0.f ? small : big;
  z1 = d > 0.f ? big : small;
}

#define FLG(a) (FL >= 0 ? FL : (a).flags)

// Raw (storage-form) holders so that every global load of a cell is issued before the first dependent instruction:
// a warp then has ~10 independent requests in flight instead of paying one DRAM round trip per map.
template <typename T, int N> struct RawN;
template <> struct RawN<bf16, 4> {
  uint2 r;
  __device__ __forceinline__ void ld(const bf16* p) { r = __ldg(reinterpret_cast<const uint2*>(p)); }
  __device__ __forceinline__ void get(float* v) const { bf16x2_unpack(r.x, v[0], v[1]); bf16x2_unpack(r.y, v[2], v[3]); }
};
template <> struct RawN<bf16, 2> {
  uint32_t r;
  __device__ __forceinline__ void ld(const bf16* p) { r = __ldg(reinterpret_cast<const uint32_t*>(p)); }
  __device__ __forceinline__ void get(float* v) const { bf16x2_unpack(r, v[0], v[1]); }
};
template <> struct RawN<bf16, 1> {
  unsigned short r;
  __device__ __forceinline__ void ld(const bf16* p) { r = __ldg(reinterpret_cast<const unsigned short*>(p)); }
  __device__ __forceinline__ void get(float* v) const { v[0] = __uint_as_float((uint32_t)r << 16); }
};
template <> struct RawN<float, 4> {
  float4 r;
  __device__ __forceinline__ void ld(const float* p) { r = __ldg(reinterpret_cast<const float4*>(p)); }
  __device__ __forceinline__ void get(float* v) const { v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w; }
};
template <> struct RawN<float, 2> {
  float2 r;
  __device__ __forceinline__ void ld(const float* p) { r = __ldg(reinterpret_cast<const float2*>(p)); }
  __device__ __forceinline__ void get(float* v) const { v[0] = r.x; v[1] = r.y; }
};
template <> struct RawN<float, 1> {
  float r;
  __device__ __forceinline__ void ld(const float* p) { r = __ldg(p); }
  __device__ __forceinline__ void get(float* v) const { v[0] = r; }
};

template <typename T>
struct CellRaw {
  Raw8<T> p0, p1, t0, t1, b0, b1;
  RawN<T, 4> q1[2];
  RawN<T, 2> q2[2];
  RawN<T, 1> q3[2], q4[2];
};

// Touch the six full-resolution vectors of the warp's NEXT cell: by the time the grid-stride loop reaches it the lines
// sit in L2 and the loads cost an L2 hit instead of a DRAM round trip (no registers are held across the iteration).
template <typename T, int FL>
__device__ __forceinline__ void prefetch_cell(const LossArgs& a, int b, int y, int x0) {
  const size_t plane = (size_t)a.H * a.W;
  const size_t off = (size_t)b * 2 * plane + (size_t)y * a.W + x0;
  if (FLG(a) & (OCT_LOSS_WPCE | OCT_LOSS_DICE)) {
    const T* yh = reinterpret_cast<const T*>(a.yhat);
    const T* ys = reinterpret_cast<const T*>(a.ys);
    asm volatile("prefetch.global.L2 [%0];" ::"l"(yh + off));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(yh + off + plane));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(ys + off));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(ys + off + plane));
  }
  if (FLG(a) & OCT_LOSS_KLD) {
    const T* a0 = reinterpret_cast<const T*>(a.att[0]);
    asm volatile("prefetch.global.L2 [%0];" ::"l"(a0 + off));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(a0 + off + plane));
  }
}

// phase 1: nothing but loads (uniform flag tests only)
template <typename T, int FL, bool PYR>
__device__ __forceinline__ void issue_cell(const LossArgs& a, int b, int y, int x0, CellRaw<T>& r) {
  const size_t plane = (size_t)a.H * a.W;
  const size_t off = (size_t)b * 2 * plane + (size_t)y * a.W + x0;
  if (FLG(a) & (OCT_LOSS_WPCE | OCT_LOSS_DICE)) {
    const T* yh = reinterpret_cast<const T*>(a.yhat);
    const T* ys = reinterpret_cast<const T*>(a.ys);
    r.p0.ld(yh + off);
    r.p1.ld(yh + off + plane);
    r.t0.ld(ys + off);
    r.t1.ld(ys + off + plane);
  }
  if (FLG(a) & OCT_LOSS_KLD) {
    const T* a0 = reinterpret_cast<const T*>(a.att[0]);
    r.b0.ld(a0 + off);
    r.b1.ld(a0 + off + plane);
    if (PYR || ((PYR || a.n_att > 1) && a.aw8[0] != 0.f)) {
      const size_t pl = plane >> 2;
      const T* q = reinterpret_cast<const T*>(a.att[1]) + (size_t)b * 2 * pl + (size_t)(y >> 1) * (a.W >> 1) + (x0 >> 1);
      r.q1[0].ld(q);
      r.q1[1].ld(q + pl);
    }
    if (PYR || ((PYR || a.n_att > 2) && a.aw8[1] != 0.f)) {
      const size_t pl = plane >> 4;
      const T* q = reinterpret_cast<const T*>(a.att[2]) + (size_t)b * 2 * pl + (size_t)(y >> 2) * (a.W >> 2) + (x0 >> 2);
      r.q2[0].ld(q);
      r.q2[1].ld(q + pl);
    }
    if (PYR || ((PYR || a.n_att > 3) && a.aw8[2] != 0.f)) {
      const size_t pl = plane >> 6;
      const T* q = reinterpret_cast<const T*>(a.att[3]) + (size_t)b * 2 * pl + (size_t)(y >> 3) * (a.W >> 3) + (x0 >> 3);
      r.q3[0].ld(q);
      r.q3[1].ld(q + pl);
    }
    if (PYR || ((PYR || a.n_att > 4) && a.aw8[3] != 0.f)) {
      const size_t pl = plane >> 8;
      const T* q = reinterpret_cast<const T*>(a.att[4]) + (size_t)b * 2 * pl + (size_t)(y >> 4) * (a.W >> 4) + (x0 >> 4);
      r.q4[0].ld(q);
      r.q4[1].ld(q + pl);
    }
  }
}

// phase 2: conversions, softmax and the coarse-level logarithms
template <typename T, int FL, bool PYR>
__device__ __forceinline__ void finish_cell(const LossArgs& a, const CellRaw<T>& r, CellLoad<T>& c) {
  if (FLG(a) & (OCT_LOSS_WPCE | OCT_LOSS_DICE)) {
    // c.p0 / c.p1 stay logits here: cell_probs() turns them into probabilities only for cells that need them
    r.p0.get(c.p0); r.p1.get(c.p1); r.t0.get(c.t0); r.t1.get(c.t1);
  }
  if (FLG(a) & OCT_LOSS_KLD) {
    r.b0.get(c.b0); r.b1.get(c.b1);
#pragma unroll
    for (int j = 0; j < 8; ++j) c.m0[j] = c.m1[j] = 0.f;
    if (PYR || ((PYR || a.n_att > 1) && a.aw8[0] != 0.f)) {
      const float w = a.aw8[0];
      r.q1[0].get(c.q1[0]); r.q1[1].get(c.q1[1]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float l0 = __logf(w * c.q1[0][j] + kEps), l1 = __logf(w * c.q1[1][j] + kEps);
        c.m0[2 * j] += l0; c.m0[2 * j + 1] += l0;
        c.m1[2 * j] += l1; c.m1[2 * j + 1] += l1;
      }
    }
    if (PYR || ((PYR || a.n_att > 2) && a.aw8[1] != 0.f)) {
      const float w = a.aw8[1];
      r.q2[0].get(c.q2[0]); r.q2[1].get(c.q2[1]);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float l0 = __logf(w * c.q2[0][j] + kEps), l1 = __logf(w * c.q2[1][j] + kEps);
#pragma unroll
        for (int i = 0; i < 4; ++i) { c.m0[4 * j + i] += l0; c.m1[4 * j + i] += l1; }
      }
    }
    if (PYR || ((PYR || a.n_att > 3) && a.aw8[2] != 0.f)) {
      const float w = a.aw8[2];
      r.q3[0].get(&c.q3[0]); r.q3[1].get(&c.q3[1]);
      float l0 = __logf(w * c.q3[0] + kEps), l1 = __logf(w * c.q3[1] + kEps);
#pragma unroll
      for (int i = 0; i < 8; ++i) { c.m0[i] += l0; c.m1[i] += l1; }
    }
    if (PYR || ((PYR || a.n_att > 4) && a.aw8[3] != 0.f)) {
      const float w = a.aw8[3];
      r.q4[0].get(&c.q4[0]); r.q4[1].get(&c.q4[1]);
      float l0 = __logf(w * c.q4[0] + kEps), l1 = __logf(w * c.q4[1] + kEps);
#pragma unroll
      for (int i = 0; i < 8; ++i) { c.m0[i] += l0; c.m1[i] += l1; }
    }
  }
}

template <typename T, int FL, bool PYR>
__device__ __forceinline__ void load_cell(const LossArgs& a, int b, int y, int x0, CellLoad<T>& c) {
  CellRaw<T> r;
  issue_cell<T, FL, PYR>(a, b, y, x0, r);
  finish_cell<T, FL, PYR>(a, r, c);
}

template <typename T, int FL>
__device__ __forceinline__ void cell_probs(const LossArgs& a, CellLoad<T>& c) {
  if (FLG(a) & OCT_LOSS_FROM_LOGITS) {
#pragma unroll
    for (int j = 0; j < 8; ++j) softmax2(c.p0[j], c.p1[j]);
  }
}

// true when some pixel of the warp's cell carries a label (any non-zero ys)
template <typename T>
__device__ __forceinline__ bool cell_labelled(const CellLoad<T>& c) {
  float tsum = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) tsum += fabsf(c.t0[j]) + fabsf(c.t1[j]);
  return __any_sync(0xffffffffu, tsum != 0.f);
}

// Persistent: one wave of blocks, every warp walks 16x16 cells (all images) with a grid stride; sums stay in registers
// until one block reduction + one set of fp64 atomics per block (the Dice sums, which are per image, go out per cell).
template <typename T, int FL, bool PYR>
__global__ void __launch_bounds__(256, 3) loss_fast_fwd_kernel(const LossArgs a) {
  __shared__ float red[10 * 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cw = a.W >> 4, cells = cw * (a.H >> 4);
  const long long total = (long long)cells * a.B;
  // acc: n0 n1 S0 S1 kld I Card lsg lsdr lsdf
  float acc[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) acc[i] = 0.f;

  if (FLG(a) & (OCT_LOSS_WPCE | OCT_LOSS_DICE | OCT_LOSS_KLD)) {
    for (long long id = (long long)blockIdx.x * 8 + warp; id < total; id += (long long)gridDim.x * 8) {
      const int b = (int)(id / cells), cell = (int)(id - (long long)b * cells);
      const int cy = cell / cw, cx = cell - cy * cw;
      const int y = cy * 16 + (lane >> 1), x0 = cx * 16 + (lane & 1) * 8;
      {
        const long long nid = id + (long long)gridDim.x * 8;
        if (nid < total) {
          const int nb = (int)(nid / cells), nc = (int)(nid - (long long)nb * cells);
          const int ncy = nc / cw, ncx = nc - ncy * cw;
          prefetch_cell<T, FL>(a, nb, ncy * 16 + (lane >> 1), ncx * 16 + (lane & 1) * 8);
        }
      }
      CellLoad<T> c;
      load_cell<T, FL, PYR>(a, b, y, x0, c);
      // unlabelled cells (most of a scribble mask) contribute ys * log(.) = 0 exactly: no softmax, no logarithms
      const bool need_p = (FLG(a) & OCT_LOSS_DICE) || ((FLG(a) & OCT_LOSS_WPCE) && ((FLG(a) & OCT_LOSS_WPCE_FULL) || cell_labelled(c)));
      if (need_p) cell_probs<T, FL>(a, c);
      if (FLG(a) & OCT_LOSS_WPCE) {
        const bool full = FLG(a) & OCT_LOSS_WPCE_FULL;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[0] += c.t0[j];
          acc[1] += c.t1[j];
        }
        if (need_p) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float a0 = full ? c.p0[j] : c.p0[j] * c.t0[j];
            float a1 = full ? c.p1[j] : c.p1[j] * c.t1[j];
            acc[2] += c.t0[j] * __logf(a0 + kEps);
            acc[3] += c.t1[j] * __logf(a1 + kEps);
          }
        }
      }
      if (FLG(a) & OCT_LOSS_DICE) {
        float di = 0.f, dc = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          di += c.p0[j] * c.t0[j] + c.p1[j] * c.t1[j];
          dc += (c.p0[j] + c.t0[j]) + (c.p1[j] + c.t1[j]);
        }
        di = warp_sum(di);
        dc = warp_sum(dc);
        if (lane == 0) {
          atomic_add_f64(a.stats + ST_DICE + b, di);
          atomic_add_f64(a.stats + ST_DICE + a.B + b, dc);
        }
      }
      if (FLG(a) & OCT_LOSS_KLD) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[4] += c.b0[j] * (__logf(c.b0[j] + kEps) - c.m0[j] * a.inv_sumw) +
                    c.b1[j] * (__logf(c.b1[j] + kEps) - c.m1[j] * a.inv_sumw);
        }
      }
    }
  }
  if (blockIdx.x == 0) ls_partial(a, acc[7], acc[8], acc[9]);

  block_sum<10>(acc, red);
  if (threadIdx.x == 0) {
    double* st = a.stats;
    if (FLG(a) & OCT_LOSS_WPCE) {
      atomic_add_f64(st + ST_N + 0, acc[0]);
      atomic_add_f64(st + ST_N + 1, acc[1]);
      atomic_add_f64(st + ST_S + 0, acc[2]);
      atomic_add_f64(st + ST_S + 1, acc[3]);
    }
    if (FLG(a) & OCT_LOSS_KLD) atomic_add_f64(st + ST_KLD, acc[4]);
    if (blockIdx.x == 0) {
      atomic_add_f64(st + ST_LSG, acc[7]);
      atomic_add_f64(st + ST_LSDR, acc[8]);
      atomic_add_f64(st + ST_LSDF, acc[9]);
    }
  }
  retire_block(a);
}

// LS-GAN gradients, one block.
__device__ __forceinline__ void ls_backward(const LossArgs& a) {
  if (a.flags & OCT_LOSS_LSD) {
    // 0.5*mean((r-1)^2) + 0.5*mean((f+1)^2): discriminator/losses.py:11-14
    const float g = a.gscale[OCT_LOSS_OUT_LSD];
    for (int i = threadIdx.x; i < a.n_real; i += blockDim.x)
      a.g_real[i] = g * (a.d_real[i] - 1.f) / (float)a.n_real;
    for (int i = threadIdx.x; i < a.n_fake; i += blockDim.x) {
      float v = g * (a.d_fake[i] + 1.f) / (float)a.n_fake;
      if (a.flags & OCT_LOSS_LSG)
        v += a.gscale[OCT_LOSS_OUT_LSG] * (a.d_fake[i] - 1.f) / (float)a.n_fake;
      a.g_fake[i] = v;
    }
  } else if (a.flags & OCT_LOSS_LSG) {
    // 0.5*mean((f-1)^2): discriminator/losses.py:22-24
    const float g = a.gscale[OCT_LOSS_OUT_LSG];
    for (int i = threadIdx.x; i < a.n_fake; i += blockDim.x)
      a.g_fake[i] = g * (a.d_fake[i] - 1.f) / (float)a.n_fake;
  }
}

// upstream gradient of each loss term (device array for the two-pass path, host constants for the fused single pass)
struct GradScale {
  float wpce, dice, kld;
};

// kld_sum != nullptr: also accumulate the cell's divergence terms (the fused single pass needs values and gradients)
template <typename T, int FL, bool PYR>
__device__ __forceinline__ void loss_bwd_cell(const LossArgs& a, int b, int cell, int cw, int lane, const GradScale gs,
                                              float* kld_sum = nullptr) {
  const int cy = cell / cw, cx = cell - cy * cw;
  const int r = lane >> 1;
  const int y = cy * 16 + r, x0 = cx * 16 + (lane & 1) * 8;
  const size_t plane = (size_t)a.H * a.W;
  const size_t off = (size_t)b * 2 * plane + (size_t)y * a.W + x0;
  CellLoad<T> c;
  load_cell<T, FL, PYR>(a, b, y, x0, c);

  if (FLG(a) & (OCT_LOSS_WPCE | OCT_LOSS_DICE)) {
    float g0[8], g1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) g0[j] = g1[j] = 0.f;
    // an unlabelled cell has an all-zero WPCE gradient (with WPCE_FULL the log argument is not masked, but ys = 0 still
    // zeroes the numerator): only Dice needs the probabilities there
    const bool need_p = (FLG(a) & OCT_LOSS_DICE) || cell_labelled(c);
    if (need_p) cell_probs<T, FL>(a, c);
    if ((FLG(a) & OCT_LOSS_WPCE) && need_p) {
      const bool full = FLG(a) & OCT_LOSS_WPCE_FULL;
      const float k = -gs.wpce * a.wpce_scale;
      const float w0 = k * (float)a.stats[ST_W + 0], w1 = k * (float)a.stats[ST_W + 1];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        // d/dyhat [ w * ys * log(yhat*ys + eps) ] = w * ys * ys / (yhat*ys + eps)
        float a0 = full ? c.p0[j] : c.p0[j] * c.t0[j];
        float a1 = full ? c.p1[j] : c.p1[j] * c.t1[j];
        float m0 = full ? c.t0[j] : c.t0[j] * c.t0[j];
        float m1 = full ? c.t1[j] : c.t1[j] * c.t1[j];
        g0[j] += __fdividef(w0 * m0, a0 + kEps);
        g1[j] += __fdividef(w1 * m1, a1 + kEps);
      }
    }
    if (FLG(a) & OCT_LOSS_DICE) {
      // L = mean_b(1 - 2 I/(Card+eps)); dL/dp = (-2 t/(Card+eps) + 2 I/(Card+eps)^2)/B
      const float I = (float)a.stats[ST_DICE + b];
      const float card = (float)a.stats[ST_DICE + a.B + b] + a.dice_eps;
      const float k = gs.dice / (float)a.B;
      const float ka = -2.f * k / card, kb = 2.f * k * I / (card * card);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        g0[j] += ka * c.t0[j] + kb;
        g1[j] += ka * c.t1[j] + kb;
      }
    }
    if ((FLG(a) & OCT_LOSS_FROM_LOGITS) && need_p) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float dot = g0[j] * c.p0[j] + g1[j] * c.p1[j];
        g0[j] = c.p0[j] * (g0[j] - dot);
        g1[j] = c.p1[j] * (g1[j] - dot);
      }
    }
    T* gy = reinterpret_cast<T*>(a.g_yhat);
    VecIO<T, 8>::st(gy + off, g0);
    VecIO<T, 8>::st(gy + off + plane, g1);
  }

  if (FLG(a) & OCT_LOSS_KLD) {
    const float gk = gs.kld / ((float)a.B * (float)a.H * (float)a.W);
    {
      float g0[8], g1[8];
      const bool stop = FLG(a) & OCT_LOSS_KLD_STOPGRAD;
      float ksum = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        // d/db [ b (log(b+eps) - m) ] = log(b+eps) - m + b/(b+eps)
        const float d0 = __logf(c.b0[j] + kEps) - c.m0[j] * a.inv_sumw, d1 = __logf(c.b1[j] + kEps) - c.m1[j] * a.inv_sumw;
        ksum += c.b0[j] * d0 + c.b1[j] * d1;
        g0[j] = stop ? 0.f : gk * (d0 + __fdividef(c.b0[j], c.b0[j] + kEps));
        g1[j] = stop ? 0.f : gk * (d1 + __fdividef(c.b1[j], c.b1[j] + kEps));
      }
      if (kld_sum) *kld_sum += ksum;
      T* g = reinterpret_cast<T*>(a.g_att[0]);
      VecIO<T, 8>::st(g + off, g0);
      VecIO<T, 8>::st(g + off + plane, g1);
    }
    // Box sums of the basis over 2^k x 2^k blocks; d/dq_k = -gk/sum_w * w_k/(w_k q + eps) * boxsum(b)
    const float kq = -gk * a.inv_sumw;
    float v1[2][4], v2[2][2], v3[2], v4[2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v1[0][j] = c.b0[2 * j] + c.b0[2 * j + 1];
      v1[1][j] = c.b1[2 * j] + c.b1[2 * j + 1];
    }
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v1[ch][j] += __shfl_xor_sync(0xffffffffu, v1[ch][j], 2);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        v2[ch][j] = v1[ch][2 * j] + v1[ch][2 * j + 1];
        v2[ch][j] += __shfl_xor_sync(0xffffffffu, v2[ch][j], 4);
      }
      v3[ch] = v2[ch][0] + v2[ch][1];
      v3[ch] += __shfl_xor_sync(0xffffffffu, v3[ch], 8);
      v4[ch] = v3[ch] + __shfl_xor_sync(0xffffffffu, v3[ch], 1);
      v4[ch] += __shfl_xor_sync(0xffffffffu, v4[ch], 16);
    }
    if ((PYR || a.n_att > 1) && (r & 1) == 0) {
      const size_t pl = plane >> 2;
      T* g = reinterpret_cast<T*>(a.g_att[1]) + (size_t)b * 2 * pl + (size_t)(y >> 1) * (a.W >> 1) + (x0 >> 1);
      const float w = a.aw8[0];
      float o[2][4];
#pragma unroll
      for (int ch = 0; ch < 2; ++ch)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[ch][j] = (PYR || w != 0.f) ? __fdividef(kq * w * v1[ch][j], w * c.q1[ch][j] + kEps) : 0.f;
      VecIO<T, 4>::st(g, o[0]);
      VecIO<T, 4>::st(g + pl, o[1]);
    }
    if ((PYR || a.n_att > 2) && (r & 3) == 0) {
      const size_t pl = plane >> 4;
      T* g = reinterpret_cast<T*>(a.g_att[2]) + (size_t)b * 2 * pl + (size_t)(y >> 2) * (a.W >> 2) + (x0 >> 2);
      const float w = a.aw8[1];
      float o[2][2];
#pragma unroll
      for (int ch = 0; ch < 2; ++ch)
#pragma unroll
        for (int j = 0; j < 2; ++j) o[ch][j] = (PYR || w != 0.f) ? __fdividef(kq * w * v2[ch][j], w * c.q2[ch][j] + kEps) : 0.f;
      VecIO<T, 2>::st(g, o[0]);
      VecIO<T, 2>::st(g + pl, o[1]);
    }
    if ((PYR || a.n_att > 3) && (r & 7) == 0) {
      const size_t pl = plane >> 6;
      T* g = reinterpret_cast<T*>(a.g_att[3]) + (size_t)b * 2 * pl + (size_t)(y >> 3) * (a.W >> 3) + (x0 >> 3);
      const float w = a.aw8[2];
      float o0 = (PYR || w != 0.f) ? __fdividef(kq * w * v3[0], w * c.q3[0] + kEps) : 0.f;
      float o1 = (PYR || w != 0.f) ? __fdividef(kq * w * v3[1], w * c.q3[1] + kEps) : 0.f;
      VecIO<T, 1>::st(g, &o0);
      VecIO<T, 1>::st(g + pl, &o1);
    }
    if ((PYR || a.n_att > 4) && lane == 0) {
      const size_t pl = plane >> 8;
      T* g = reinterpret_cast<T*>(a.g_att[4]) + (size_t)b * 2 * pl + (size_t)(y >> 4) * (a.W >> 4) + (x0 >> 4);
      const float w = a.aw8[3];
      float o0 = (PYR || w != 0.f) ? __fdividef(kq * w * v4[0], w * c.q4[0] + kEps) : 0.f;
      float o1 = (PYR || w != 0.f) ? __fdividef(kq * w * v4[1], w * c.q4[1] + kEps) : 0.f;
      VecIO<T, 1>::st(g, &o0);
      VecIO<T, 1>::st(g + pl, &o1);
    }
  }
}


template <typename T, int FL, bool PYR>
__global__ void __launch_bounds__(256, 3) loss_fast_bwd_kernel(const LossArgs a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cw = a.W >> 4, cells = cw * (a.H >> 4);
  const long long total = (long long)cells * a.B;
  if (blockIdx.x == 0) ls_backward(a);
  if (!(FLG(a) & (OCT_LOSS_WPCE | OCT_LOSS_DICE | OCT_LOSS_KLD))) return;
  const GradScale gs{a.gscale[OCT_LOSS_OUT_WPCE], a.gscale[OCT_LOSS_OUT_DICE], a.gscale[OCT_LOSS_OUT_KLD]};
  for (long long id = (long long)blockIdx.x * 8 + warp; id < total; id += (long long)gridDim.x * 8) {
    const int b = (int)(id / cells), cell = (int)(id - (long long)b * cells);
    {
      const long long nid = id + (long long)gridDim.x * 8;
      if (nid < total) {
        const int nb = (int)(nid / cells), nc = (int)(nid - (long long)nb * cells);
        const int ncy = nc / cw, ncx = nc - ncy * cw;
        prefetch_cell<T, FL>(a, nb, ncy * 16 + (lane >> 1), ncx * 16 + (lane & 1) * 8);
      }
    }
    loss_bwd_cell<T, FL, PYR>(a, b, cell, cw, lane, gs);
  }
}

// ------------------------------------------------------------------------------------------------
// Fused single pass (the G-step of the training loop): loss values AND gradients from ONE sweep over the maps.
//
// The WPCE gradient needs the class weights w_c = sum(n)/(n_c + eps) (losses.py:34-38), which depend on the labels only:
// a labels-only pre-pass (loss_label_count_kernel, 2 elements/pixel) counts them, then loss_fused_kernel reads every
// map once and writes every gradient once: 2 + 6.664 + 4.664 = 13.33 elements/pixel instead of 17.99 for the
// statistics pass + gradient pass pair.  The upstream gradients are the term weights of  total = lam_wpce * WPCE +
// lam_kld * KLD + lam_lsg * LSG, known when the forward runs; a different upstream gradient of `total` is applied
// afterwards by loss_scale_kernel, which exits at once when that gradient is 1 (the `total.backward()` case).
//
// WPCE (agg, ys -> g_agg) is purely elementwise and KLD (att pyramid -> g_att) works on 16x16 cells; the two parts
// share nothing, so a block walks WPCE chunks first and KLD cells afterwards (fewer live registers than one merged
// body).  Reductions are deterministic: every block stores its partial sums, the last block adds them in block order.
// ------------------------------------------------------------------------------------------------
constexpr int kFusedMaxBlocks = 4 * 256;   // >= 4 resident blocks x SM count
constexpr int kFusedPartials = 4;          // S0 S1 kld pad

template <typename T>
__global__ void __launch_bounds__(256) loss_label_count_kernel(const T* __restrict__ ys, long long plane8, long long groups,
                                                              double* stats) {
  __shared__ float red[2 * 8];
  float n[2] = {0.f, 0.f};
  constexpr int U = 8;
  for (long long g0 = (long long)blockIdx.x * 256 * U + threadIdx.x; g0 < groups; g0 += (long long)gridDim.x * 256 * U) {
    Raw8<T> r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long g = g0 + (long long)u * 256;
      if (g < groups) r[u].ld(ys + g * 8); else r[u].zero();
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long g = g0 + (long long)u * 256;
      float v[8];
      r[u].get(v);
      const float s = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
      const int c = (int)((g / plane8) & 1);      // [B][2][plane]: channel of this 8-pixel group
      n[0] += c ? 0.f : s;
      n[1] += c ? s : 0.f;
    }
  }
  block_sum<2>(n, red);
  if (threadIdx.x == 0) {
    // per-block sums of 0/1 labels are exact integers, and fp64 sums of integers do not depend on the order
    atomic_add_f64(stats + ST_N + 0, n[0]);
    atomic_add_f64(stats + ST_N + 1, n[1]);
  }
}

// ---- KLD part of the fused pass: the work unit is ONE class plane of one 16x16 cell (the divergence and its gradients are
// separable per class), so a lane carries 8 pixels of one plane instead of two: half the live registers, twice the units.
template <typename T>
struct PlaneRaw {
  Raw8<T> b;
  RawN<T, 4> q1;
  RawN<T, 2> q2;
  RawN<T, 1> q3, q4;
  unsigned off[5];               // element offset of this lane's first value in att[k] / g_att[k]
};

// exact unsigned division by a run-time constant (round-up multiplier; valid for every 32-bit n)
struct FastDiv {
  unsigned d, m, s;
  __host__ void init(unsigned div) {
    d = div;
    unsigned l = 0;
    while ((1ull << l) < div) ++l;
    m = (unsigned)((((1ull << l) - div) << 32) / div + 1);
    s = l;
  }
  __device__ __forceinline__ unsigned div(unsigned n) const {
    const unsigned t = __umulhi(m, n);
    return s == 0 ? n : (t + ((n - t) >> 1)) >> (s - 1);
  }
};

struct PlaneGeo {
  FastDiv cells, cw;             // cells per plane / per row
  unsigned units;                // B * 2 * cells
  unsigned W, plane;
  bool use[4];                   // level k+1 takes part (present and non-zero weight)
};

template <typename T, bool PYR>
__device__ __forceinline__ void plane_issue(const LossArgs& a, const PlaneGeo& g, unsigned u, int lane, PlaneRaw<T>& r) {
  const unsigned pl = g.cells.div(u), cell = u - pl * g.cells.d;
  const unsigned cy = g.cw.div(cell), cx = cell - cy * g.cw.d;
  const unsigned y = cy * 16 + (lane >> 1), x0 = cx * 16 + (lane & 1) * 8;
  r.off[0] = pl * g.plane + y * g.W + x0;
  r.off[1] = pl * (g.plane >> 2) + (y >> 1) * (g.W >> 1) + (x0 >> 1);
  r.off[2] = pl * (g.plane >> 4) + (y >> 2) * (g.W >> 2) + (x0 >> 2);
  r.off[3] = pl * (g.plane >> 6) + (y >> 3) * (g.W >> 3) + (x0 >> 3);
  r.off[4] = pl * (g.plane >> 8) + (y >> 4) * (g.W >> 4) + (x0 >> 4);
  r.b.ld(reinterpret_cast<const T*>(a.att[0]) + r.off[0]);
  if (PYR || g.use[0]) r.q1.ld(reinterpret_cast<const T*>(a.att[1]) + r.off[1]);
  if (PYR || g.use[1]) r.q2.ld(reinterpret_cast<const T*>(a.att[2]) + r.off[2]);
  if (PYR || g.use[2]) r.q3.ld(reinterpret_cast<const T*>(a.att[3]) + r.off[3]);
  if (PYR || g.use[3]) r.q4.ld(reinterpret_cast<const T*>(a.att[4]) + r.off[4]);
}

// single-instruction base-2 logarithm / reciprocal: every argument here is >= 1e-12 (eps is added first), far above the
// denormal range whose handling makes __log2f four instructions
__device__ __forceinline__ float lg2_fast(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_fast(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// values (sum into ksum) and gradients of one plane unit: the arithmetic of loss_bwd_cell term by term, except that the
// logarithms stay in base 2 until one multiplication by ln 2 per pixel, the four coarse-level logarithms are summed once per
// pixel PAIR (levels >= 1 are constant over a pair), the arguments w*q + eps are computed once for the logarithm and the
// gradient, and b/(b + 1e-12) is taken as exactly 1 where fp32 rounds b + 1e-12 to b (b > 2e-5; warp-uniform test)
template <typename T, bool PYR>
__device__ __forceinline__ void plane_finish(const LossArgs& a, const PlaneGeo& g, int lane, const PlaneRaw<T>& r, float gk, float& ksum) {
  constexpr float kLn2 = 0.6931471805599453f;
  const int row = lane >> 1;
  float b[8], q1[4], q2[2], q3 = 1.f, q4 = 1.f;   // q*: w*q + eps of each level (1 -> log 0 for an absent level)
  r.b.get(b);
#pragma unroll
  for (int j = 0; j < 4; ++j) q1[j] = 1.f;
  q2[0] = q2[1] = 1.f;
  if (PYR || g.use[0]) {
    const float w = a.aw8[0];
    r.q1.get(q1);
#pragma unroll
    for (int j = 0; j < 4; ++j) q1[j] = w * q1[j] + kEps;
  }
  if (PYR || g.use[1]) {
    const float w = a.aw8[1];
    r.q2.get(q2);
#pragma unroll
    for (int j = 0; j < 2; ++j) q2[j] = w * q2[j] + kEps;
  }
  if (PYR || g.use[2]) {
    r.q3.get(&q3);
    q3 = a.aw8[2] * q3 + kEps;
  }
  if (PYR || g.use[3]) {
    r.q4.get(&q4);
    q4 = a.aw8[3] * q4 + kEps;
  }
  // m2[p] = (sum over the levels of log2(w_k q_k + eps)) / sum_w for the pixel pair p (levels >= 1 are constant over a pair)
  float m2[4];
  {
    const float s34 = lg2_fast(q3) + lg2_fast(q4);
    const float s2a = lg2_fast(q2[0]) + s34, s2b = lg2_fast(q2[1]) + s34;
    m2[0] = (lg2_fast(q1[0]) + s2a) * a.inv_sumw;
    m2[1] = (lg2_fast(q1[1]) + s2a) * a.inv_sumw;
    m2[2] = (lg2_fast(q1[2]) + s2b) * a.inv_sumw;
    m2[3] = (lg2_fast(q1[3]) + s2b) * a.inv_sumw;
  }
  {
    float gb[8];
    float ks = 0.f, bmin = b[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) bmin = fminf(bmin, b[j]);
    const bool unit_ratio = __all_sync(0xffffffffu, bmin > 2e-5f);
    const float gkl = gk * kLn2;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // d/db [ b (log(b+eps) - m) ] = log(b+eps) - m + b/(b+eps);  t = the bracket in base 2
      const float t = lg2_fast(b[j] + kEps) - m2[j >> 1];
      ks += b[j] * t;
      gb[j] = unit_ratio ? fmaf(gkl, t, gk) : fmaf(gkl, t, gk * b[j] * rcp_fast(b[j] + kEps));
    }
    ksum += kLn2 * ks;
    VecIO<T, 8>::st(reinterpret_cast<T*>(a.g_att[0]) + r.off[0], gb);
  }
  // box sums of the basis over 2^k x 2^k blocks; d/dq_k = -gk/sum_w * w_k/(w_k q + eps) * boxsum(b)
  const float kq = -gk * a.inv_sumw;
  float v1[4], v2[2], v3, v4;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v1[j] = b[2 * j] + b[2 * j + 1];
    v1[j] += __shfl_xor_sync(0xffffffffu, v1[j], 2);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    v2[j] = v1[2 * j] + v1[2 * j + 1];
    v2[j] += __shfl_xor_sync(0xffffffffu, v2[j], 4);
  }
  v3 = v2[0] + v2[1];
  v3 += __shfl_xor_sync(0xffffffffu, v3, 8);
  v4 = v3 + __shfl_xor_sync(0xffffffffu, v3, 1);
  v4 += __shfl_xor_sync(0xffffffffu, v4, 16);
  if ((PYR || a.n_att > 1) && (row & 1) == 0) {
    const float kw = kq * a.aw8[0];
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = (PYR || g.use[0]) ? kw * v1[j] * rcp_fast(q1[j]) : 0.f;
    VecIO<T, 4>::st(reinterpret_cast<T*>(a.g_att[1]) + r.off[1], o);
  }
  if ((PYR || a.n_att > 2) && (row & 3) == 0) {
    const float kw = kq * a.aw8[1];
    float o[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) o[j] = (PYR || g.use[1]) ? kw * v2[j] * rcp_fast(q2[j]) : 0.f;
    VecIO<T, 2>::st(reinterpret_cast<T*>(a.g_att[2]) + r.off[2], o);
  }
  if ((PYR || a.n_att > 3) && (row & 7) == 0) {
    float o = (PYR || g.use[2]) ? kq * a.aw8[2] * v3 * rcp_fast(q3) : 0.f;
    VecIO<T, 1>::st(reinterpret_cast<T*>(a.g_att[3]) + r.off[3], &o);
  }
  if ((PYR || a.n_att > 4) && lane == 0) {
    float o = (PYR || g.use[3]) ? kq * a.aw8[3] * v4 * rcp_fast(q4) : 0.f;
    VecIO<T, 1>::st(reinterpret_cast<T*>(a.g_att[4]) + r.off[4], &o);
  }
}

// COOP (cooperative launch, every block resident): the labels-only count runs as phase 0 of this very kernel — each block
// counts a slice of ys and publishes it, does its KLD units, and only then (when every block's count has long arrived)
// the WPCE part, which re-reads ys largely from L2 — instead of a separate pre-pass launch.
template <typename T, bool PYR, bool COOP>
__global__ void __launch_bounds__(256, 4) loss_fused_kernel(const LossArgs a, const PlaneGeo g) {
  __shared__ float red[3 * 8];
  __shared__ int s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[3] = {0.f, 0.f, 0.f};   // S0 S1 kld
  const unsigned plane = (unsigned)a.H * (unsigned)a.W;
  if (COOP) {
    if (a.flags & OCT_LOSS_WPCE) {
      const T* ys = reinterpret_cast<const T*>(a.ys);
      const unsigned plane8 = plane >> 3, groups = plane8 * 2u * (unsigned)a.B;    // 8-element groups of the whole [B][2][plane] tensor
      float n[2] = {0.f, 0.f};
      for (unsigned g0 = blockIdx.x * 256u + threadIdx.x; g0 < groups; g0 += gridDim.x * 256u) {
        Raw8<T> r;
        r.ld(ys + (size_t)g0 * 8);
        float v[8];
        r.get(v);
        const float sm = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
        const bool c1 = (g0 / plane8) & 1u;
        n[0] += c1 ? 0.f : sm;
        n[1] += c1 ? sm : 0.f;
      }
      block_sum<2>(n, red);
      if (threadIdx.x == 0) {   // exact integer sums: the order of the fp64 atomics does not matter
        atomic_add_f64(a.stats + ST_N + 0, n[0]);
        atomic_add_f64(a.stats + ST_N + 1, n[1]);
      }
    }
    // publish this block's counts: the WPCE part below waits until every block has done so.  The KLD part runs in between,
    // so in practice nobody waits (all blocks are resident: cooperative launch), and the count costs no separate launch.
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(reinterpret_cast<unsigned long long*>(a.stats + ST_COUNTER + 1), 1ULL);
    }
  }
  if (a.flags & OCT_LOSS_KLD) {
    const float gk = a.lam_kld / ((float)a.B * (float)a.H * (float)a.W);
    const unsigned stride = gridDim.x * 8u;
    // blocks are walked in reverse here: the ones that drew an extra WPCE round above draw one KLD round less
    unsigned u = (gridDim.x - 1 - blockIdx.x) * 8u + warp;
    PlaneRaw<T> nxt;
    if (u < g.units) plane_issue<T, PYR>(a, g, u, lane, nxt);
    while (u < g.units) {
      const PlaneRaw<T> cur = nxt;
      u += stride;
      if (u < g.units) plane_issue<T, PYR>(a, g, u, lane, nxt);   // next unit's loads fly while this one is computed
      plane_finish<T, PYR>(a, g, lane, cur, gk, acc[2]);
    }
  }

  if (COOP && (a.flags & OCT_LOSS_WPCE)) {
    if (threadIdx.x == 0) {
      volatile unsigned long long* c = reinterpret_cast<volatile unsigned long long*>(a.stats + ST_COUNTER + 1);
      while (*c < (unsigned long long)gridDim.x) __nanosleep(64);
      __threadfence();
    }
    __syncthreads();
  }
  if (a.flags & OCT_LOSS_WPCE) {
    // class weights exactly as finalize() forms them
    const float n0 = (float)__ldcg(a.stats + ST_N + 0), n1 = (float)__ldcg(a.stats + ST_N + 1);
    const float ntot = n0 + n1;
    const float k = -a.lam_wpce * a.wpce_scale;
    const float w0 = k * (ntot / (n0 + kEps)), w1 = k * (ntot / (n1 + kEps));
    const bool full = a.flags & OCT_LOSS_WPCE_FULL, logits = a.flags & OCT_LOSS_FROM_LOGITS;
    const T* yh = reinterpret_cast<const T*>(a.yhat);
    const T* ys = reinterpret_cast<const T*>(a.ys);
    T* gy = reinterpret_cast<T*>(a.g_yhat);
    const unsigned plane8 = plane >> 3, groups = plane8 * (unsigned)a.B;   // 8-pixel groups of one class plane
    const unsigned stride = gridDim.x * 256u;
    unsigned g = blockIdx.x * 256u + threadIdx.x;
    // software pipeline: the four 16-byte loads of the NEXT group are in flight while the current one is processed
    Raw8<T> nz0, nz1, nt0, nt1;
    size_t noff = 0;
    if (g < groups) {
      const unsigned b = g / plane8;
      noff = (size_t)b * 2 * plane + (size_t)(g - b * plane8) * 8;
      nz0.ld(yh + noff); nz1.ld(yh + noff + plane); nt0.ld(ys + noff); nt1.ld(ys + noff + plane);
    }
    while (g < groups) {
      const Raw8<T> rz0 = nz0, rz1 = nz1, rt0 = nt0, rt1 = nt1;
      const size_t off = noff;
      g += stride;
      if (g < groups) {
        const unsigned b = g / plane8;
        noff = (size_t)b * 2 * plane + (size_t)(g - b * plane8) * 8;
        nz0.ld(yh + noff); nz1.ld(yh + noff + plane); nt0.ld(ys + noff); nt1.ld(ys + noff + plane);
      }
      float q0[8], q1[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) q0[j] = q1[j] = 0.f;
      if (rt0.any_nonzero() || rt1.any_nonzero()) {   // an unlabelled group contributes ys * log(.) = 0 and an all-zero gradient
        float p0[8], p1[8], t0[8], t1[8];
        rt0.get(t0); rt1.get(t1);
        rz0.get(p0); rz1.get(p1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (logits) softmax2(p0[j], p1[j]);
          const float a0 = full ? p0[j] : p0[j] * t0[j], a1 = full ? p1[j] : p1[j] * t1[j];
          const float m0 = full ? t0[j] : t0[j] * t0[j], m1 = full ? t1[j] : t1[j] * t1[j];
          acc[0] += t0[j] * __logf(a0 + kEps);
          acc[1] += t1[j] * __logf(a1 + kEps);
          float e0 = __fdividef(w0 * m0, a0 + kEps), e1 = __fdividef(w1 * m1, a1 + kEps);
          if (logits) {
            const float dot = e0 * p0[j] + e1 * p1[j];
            e0 = p0[j] * (e0 - dot);
            e1 = p1[j] * (e1 - dot);
          }
          q0[j] = e0; q1[j] = e1;
        }
      }
      VecIO<T, 8>::st(gy + off, q0);
      VecIO<T, 8>::st(gy + off + plane, q1);
    }
  }

  float ls[3] = {0.f, 0.f, 0.f};
  if (blockIdx.x == 0 && (a.flags & OCT_LOSS_LSG)) {
    // 0.5*mean((f-1)^2) and its gradient: discriminator/losses.py:22-24
    for (int i = threadIdx.x; i < a.n_fake; i += blockDim.x) {
      const float f = a.d_fake[i];
      ls[0] += (f - 1.f) * (f - 1.f);
      a.g_fake[i] = a.lam_lsg * (f - 1.f) / (float)a.n_fake;
    }
  }
  block_sum<3>(acc, red);
  if (blockIdx.x == 0) block_sum<3>(ls, red);
  double* part = a.stats + ST_DICE + 2 * a.B;
  if (threadIdx.x == 0) {
    part[blockIdx.x * kFusedPartials + 0] = (double)acc[0];
    part[blockIdx.x * kFusedPartials + 1] = (double)acc[1];
    part[blockIdx.x * kFusedPartials + 2] = (double)acc[2];
    if (blockIdx.x == 0) a.stats[ST_LSG] = (double)ls[0];
    __threadfence();
    const unsigned long long prev = atomicAdd(reinterpret_cast<unsigned long long*>(a.stats + ST_COUNTER), 1ULL);
    s_last = (prev == (unsigned long long)gridDim.x - 1ULL);
  }
  __syncthreads();
  __shared__ double fin[3][256];
  if (s_last) {
    // fixed-order (bit-reproducible) sum of the per-block partials by the whole block: thread t adds blocks t, t+256, ...
    // (independent L2 loads in flight), then a fixed shared-memory tree
    __threadfence();
    double s0 = 0.0, s1 = 0.0, kl = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += 256) {
      s0 += __ldcg(part + i * kFusedPartials + 0);
      s1 += __ldcg(part + i * kFusedPartials + 1);
      kl += __ldcg(part + i * kFusedPartials + 2);
    }
    fin[0][threadIdx.x] = s0; fin[1][threadIdx.x] = s1; fin[2][threadIdx.x] = kl;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
      if ((int)threadIdx.x < w) {
#pragma unroll
        for (int q = 0; q < 3; ++q) fin[q][threadIdx.x] += fin[q][threadIdx.x + w];
      }
      __syncthreads();
    }
  }
  if (s_last && threadIdx.x == 0) {
    volatile double* st = a.stats;
    st[ST_S + 0] = fin[0][0]; st[ST_S + 1] = fin[1][0]; st[ST_KLD] = fin[2][0];
    finalize(a);
    a.out[6] = a.lam_wpce * a.out[OCT_LOSS_OUT_WPCE] + a.lam_kld * a.out[OCT_LOSS_OUT_KLD] + a.lam_lsg * a.out[OCT_LOSS_OUT_LSG];
    // leave the accumulators and counters zero for the next evaluation: no memset node between two launches
    for (int i = 0; i < ST_DICE; ++i) st[i] = 0.0;
  }
}

// In-place scaling of the gradients written by the fused pass by the upstream gradient of `total`; a gradient of
// exactly 1 (the usual `total.backward()`) returns at once.
struct ScaleArgs {
  void* ptr[8];
  long long n[8];   // elements, multiples of 8 for the maps
  int count, dtype;
  const float* g;
};
__global__ void __launch_bounds__(256) loss_scale_kernel(const ScaleArgs sa) {
  const float g = *sa.g;
  if (g == 1.f) return;
  for (int t = 0; t < sa.count; ++t) {
    const long long n = sa.n[t];
    if (sa.dtype == OCT_DTYPE_BF16 && t < sa.count - 1) {
      bf16* p = reinterpret_cast<bf16*>(sa.ptr[t]);
      for (long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 8; i + 8 <= n; i += (long long)gridDim.x * 256 * 8) {
        float v[8];
        VecIO<bf16, 8>::ld(p + i, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= g;
        VecIO<bf16, 8>::st(p + i, v);
      }
      for (long long i = (n & ~7LL) + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
        p[i] = __float2bfloat16_rn(__bfloat162float(p[i]) * g);
    } else {
      float* p = reinterpret_cast<float*>(sa.ptr[t]);
      for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) p[i] *= g;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Generic path: any C <= 8, any attention sizes (torch 'nearest': src = min(floor(dst*in/out), in-1)),
// fp32 storage.  One thread per full-resolution pixel; blockIdx.y = sample.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int nearest_src(int dst, int in, int out) {
  float scale = (float)in / (float)out;
  int s = (int)floorf((float)dst * scale);
  return s < in - 1 ? s : in - 1;
}

struct GenericPixel {
  float p[OCT_LOSS_MAX_CLASSES];
  float t[OCT_LOSS_MAX_CLASSES];
  float bs[OCT_LOSS_MAX_CLASSES];
  float m[OCT_LOSS_MAX_CLASSES];
};

__device__ __forceinline__ void generic_load(const LossArgs& a, int b, int y, int x, GenericPixel& px) {
  const int C = a.C;
  const size_t plane = (size_t)a.H * a.W;
  const size_t off = (size_t)b * C * plane + (size_t)y * a.W + x;
  if (a.flags & (OCT_LOSS_WPCE | OCT_LOSS_DICE)) {
    const float* yh = reinterpret_cast<const float*>(a.yhat);
    const float* ys = reinterpret_cast<const float*>(a.ys);
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) {
      px.p[c] = yh[off + c * plane];
      px.t[c] = ys[off + c * plane];
      mx = fmaxf(mx, px.p[c]);
    }
    if (a.flags & OCT_LOSS_FROM_LOGITS) {
      float s = 0.f;
      for (int c = 0; c < C; ++c) { px.p[c] = __expf(px.p[c] - mx); s += px.p[c]; }
      float inv = 1.f / s;
      for (int c = 0; c < C; ++c) px.p[c] *= inv;
    }
  }
  if (a.flags & OCT_LOSS_KLD) {
    const float* a0 = reinterpret_cast<const float*>(a.att[0]);
    for (int c = 0; c < C; ++c) { px.bs[c] = a0[off + c * plane]; px.m[c] = 0.f; }
    int used = 0;
    for (int k = 1; k < a.n_att; ++k) {
      const float w = a.aw8[k - 1];
      if (w == 0.f) continue;
      ++used;
      const int sy = nearest_src(y, a.ah[k], a.H), sx = nearest_src(x, a.aw[k], a.W);
      const size_t pl = (size_t)a.ah[k] * a.aw[k];
      const float* q = reinterpret_cast<const float*>(a.att[k]) + (size_t)b * C * pl + (size_t)sy * a.aw[k] + sx;
      if (a.flags & OCT_LOSS_JSD) {
        for (int c = 0; c < C; ++c) px.m[c] += w * q[c * pl];               // sum of the weighted posteriors (losses.py:126)
      } else {
        for (int c = 0; c < C; ++c) px.m[c] += __logf(w * q[c * pl] + kEps);
      }
    }
    if (a.flags & OCT_LOSS_JSD) {
      const float inv = used > 0 ? 1.f / (float)used : 0.f;                  // mean over the levels (losses.py:156-157)
      for (int c = 0; c < C; ++c) px.m[c] *= inv;
    }
  }
}

// number of posterior levels that take part (non-zero weight)
__device__ __forceinline__ int used_levels(const LossArgs& a) {
  int used = 0;
  for (int k = 1; k < a.n_att; ++k) used += a.aw8[k - 1] != 0.f;
  return used;
}

__global__ void __launch_bounds__(256) loss_generic_fwd_kernel(const LossArgs a) {
  constexpr int NV = 2 * OCT_LOSS_MAX_CLASSES + 6;
  __shared__ float red[NV * 8];
  const int b = blockIdx.y, C = a.C;
  const int hw = a.H * a.W;
  float acc[NV];  // n[8] S[8] kld I Card lsg lsdr lsdf
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;
  const bool maps = a.flags & (OCT_LOSS_WPCE | OCT_LOSS_DICE | OCT_LOSS_KLD);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; maps && i < hw; i += gridDim.x * blockDim.x) {
    const int y = i / a.W, x = i - y * a.W;
    GenericPixel px;
    generic_load(a, b, y, x, px);
    const bool full = a.flags & OCT_LOSS_WPCE_FULL;
#pragma unroll
    for (int c = 0; c < OCT_LOSS_MAX_CLASSES; ++c) {
      if (c < C) {
        if (a.flags & OCT_LOSS_WPCE) {
          float arg = full ? px.p[c] : px.p[c] * px.t[c];
          acc[c] += px.t[c];
          acc[8 + c] += px.t[c] * __logf(arg + kEps);
        }
        if (a.flags & OCT_LOSS_DICE) {
          acc[17] += px.p[c] * px.t[c];
          acc[18] += px.p[c] + px.t[c];
        }
        if (a.flags & OCT_LOSS_KLD) {
          if (a.flags & OCT_LOSS_JSD) {
            // 0.5 * b * (log b - log M) + 0.5 * mq * (log mq - log M),  M = (b + mq) / 2   (losses.py:158-168)
            const float lm = logf(0.5f * (px.bs[c] + px.m[c]) + a.jsd_eps);
            acc[16] += 0.5f * px.bs[c] * (logf(px.bs[c] + kEps) - lm) + 0.5f * px.m[c] * (logf(px.m[c] + kEps) - lm);
          } else {
            acc[16] += px.bs[c] * (__logf(px.bs[c] + kEps) - px.m[c] * a.inv_sumw);
          }
        }
      }
    }
  }
  if (blockIdx.x == 0 && blockIdx.y == 0) ls_partial(a, acc[19], acc[20], acc[21]);
  block_sum<NV>(acc, red);
  if (threadIdx.x == 0) {
    double* st = a.stats;
    if (a.flags & OCT_LOSS_WPCE)
      for (int c = 0; c < C; ++c) {
        atomic_add_f64(st + ST_N + c, acc[c]);
        atomic_add_f64(st + ST_S + c, acc[8 + c]);
      }
    if (a.flags & OCT_LOSS_KLD) atomic_add_f64(st + ST_KLD, acc[16]);
    if (a.flags & OCT_LOSS_DICE) {
      atomic_add_f64(st + ST_DICE + b, acc[17]);
      atomic_add_f64(st + ST_DICE + a.B + b, acc[18]);
    }
    if (blockIdx.x == 0 && blockIdx.y == 0) {
      atomic_add_f64(st + ST_LSG, acc[19]);
      atomic_add_f64(st + ST_LSDR, acc[20]);
      atomic_add_f64(st + ST_LSDF, acc[21]);
    }
  }
  retire_block(a);
}

__global__ void __launch_bounds__(256) loss_generic_bwd_kernel(const LossArgs a) {
  const int b = blockIdx.y, C = a.C;
  const int hw = a.H * a.W;
  if (blockIdx.x == 0 && blockIdx.y == 0) ls_backward(a);
  if (!(a.flags & (OCT_LOSS_WPCE | OCT_LOSS_DICE | OCT_LOSS_KLD))) return;
  const size_t plane = (size_t)hw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
    const int y = i / a.W, x = i - y * a.W;
    const size_t off = (size_t)b * C * plane + (size_t)i;
    GenericPixel px;
    generic_load(a, b, y, x, px);
    if (a.flags & (OCT_LOSS_WPCE | OCT_LOSS_DICE)) {
      float g[OCT_LOSS_MAX_CLASSES];
      const bool full = a.flags & OCT_LOSS_WPCE_FULL;
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < OCT_LOSS_MAX_CLASSES; ++c) {
        g[c] = 0.f;
        if (c < C) {
          if (a.flags & OCT_LOSS_WPCE) {
            float arg = full ? px.p[c] : px.p[c] * px.t[c];
            float mm = full ? px.t[c] : px.t[c] * px.t[c];
            g[c] += -a.gscale[OCT_LOSS_OUT_WPCE] * a.wpce_scale * (float)a.stats[ST_W + c] * mm / (arg + kEps);
          }
          if (a.flags & OCT_LOSS_DICE) {
            const float I = (float)a.stats[ST_DICE + b];
            const float card = (float)a.stats[ST_DICE + a.B + b] + a.dice_eps;
            const float k = a.gscale[OCT_LOSS_OUT_DICE] / (float)a.B;
            g[c] += -2.f * k / card * px.t[c] + 2.f * k * I / (card * card);
          }
          dot += g[c] * px.p[c];
        }
      }
      float* gy = reinterpret_cast<float*>(a.g_yhat);
#pragma unroll
      for (int c = 0; c < OCT_LOSS_MAX_CLASSES; ++c)
        if (c < C) gy[off + c * plane] = (a.flags & OCT_LOSS_FROM_LOGITS) ? px.p[c] * (g[c] - dot) : g[c];
    }
    if (a.flags & OCT_LOSS_KLD) {
      const float gk = a.gscale[OCT_LOSS_OUT_KLD] / ((float)a.B * (float)hw);
      const bool stop = a.flags & OCT_LOSS_KLD_STOPGRAD;
      const bool jsd = a.flags & OCT_LOSS_JSD;
      float* g0 = reinterpret_cast<float*>(a.g_att[0]);
      float gq[OCT_LOSS_MAX_CLASSES];   // JSD: gradient w.r.t. mean_q
#pragma unroll
      for (int c = 0; c < OCT_LOSS_MAX_CLASSES; ++c) {
        gq[c] = 0.f;
        if (c < C) {
          if (jsd) {
            const float bsum = px.bs[c] + px.m[c];
            const float mix = 0.5f * bsum + a.jsd_eps;
            const float lm = logf(mix), common = 0.25f * bsum / mix;
            const float gb = 0.5f * (logf(px.bs[c] + kEps) - lm + px.bs[c] / (px.bs[c] + kEps)) - common;
            gq[c] = gk * (0.5f * (logf(px.m[c] + kEps) - lm + px.m[c] / (px.m[c] + kEps)) - common);
            g0[off + c * plane] = stop ? 0.f : gk * gb;
          } else {
            g0[off + c * plane] =
                stop ? 0.f : gk * (__logf(px.bs[c] + kEps) - px.m[c] * a.inv_sumw + px.bs[c] / (px.bs[c] + kEps));
          }
        }
      }
      const int used = jsd ? used_levels(a) : 0;
      for (int k = 1; k < a.n_att; ++k) {
        const float w = a.aw8[k - 1];
        if (w == 0.f) continue;
        const int sy = nearest_src(y, a.ah[k], a.H), sx = nearest_src(x, a.aw[k], a.W);
        const size_t pl = (size_t)a.ah[k] * a.aw[k];
        const size_t qo = (size_t)b * C * pl + (size_t)sy * a.aw[k] + sx;
        const float* q = reinterpret_cast<const float*>(a.att[k]) + qo;
        float* g = reinterpret_cast<float*>(a.g_att[k]) + qo;
        for (int c = 0; c < C; ++c) {
          if (jsd) atomicAdd(g + c * pl, gq[c] * w / (float)used);
          else atomicAdd(g + c * pl, -gk * a.inv_sumw * w / (w * q[c * pl] + kEps) * px.bs[c]);
        }
      }
    }
  }
}

bool fast_ok(const OctaveLossDesc* d) {
  const bool maps = d->flags & (OCT_LOSS_WPCE | OCT_LOSS_DICE | OCT_LOSS_KLD);
  if (!maps) return true;  // LS-only launches take the fast kernel with a 1x1 grid
  if (d->flags & OCT_LOSS_JSD) return false;   // Jensen-Shannon variant: generic kernel
  if (d->C != 2 || (d->H & 15) || (d->W & 15)) return false;
  if (d->flags & OCT_LOSS_KLD) {
    if (d->n_att < 2 || d->n_att > OCT_LOSS_MAX_ATT) return false;
    for (int k = 0; k < d->n_att; ++k)
      if (d->att_h[k] != (d->H >> k) || d->att_w[k] != (d->W >> k)) return false;
  }
  return true;
}

int validate(const OctaveLossDesc* d) {
  if (!d) return OCT_ERR_INVALID;
  if (d->dtype != OCT_DTYPE_F32 && d->dtype != OCT_DTYPE_BF16) return OCT_ERR_INVALID;
  const bool maps = d->flags & (OCT_LOSS_WPCE | OCT_LOSS_DICE | OCT_LOSS_KLD);
  if (maps) {
    if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->C <= 0 || d->C > OCT_LOSS_MAX_CLASSES) return OCT_ERR_INVALID;
    if (d->B > 65535) return OCT_ERR_UNSUPPORTED;
  }
  if ((d->flags & OCT_LOSS_KLD) && (d->n_att < 2 || d->n_att > OCT_LOSS_MAX_ATT)) return OCT_ERR_INVALID;
  if ((d->flags & (OCT_LOSS_LSG | OCT_LOSS_LSD)) && d->n_fake <= 0) return OCT_ERR_INVALID;
  if ((d->flags & OCT_LOSS_LSD) && d->n_real <= 0) return OCT_ERR_INVALID;
  if (!fast_ok(d) && d->dtype != OCT_DTYPE_F32) return OCT_ERR_UNSUPPORTED;
  return OCT_OK;
}

constexpr int kGStepFlags = OCT_LOSS_WPCE | OCT_LOSS_KLD | OCT_LOSS_FROM_LOGITS;
// map flags of the G-step (LS-G rides along: it is handled by block 0 outside the specialised code)
bool gstep_config(const OctaveLossDesc* d) {
  if ((d->flags & ~(OCT_LOSS_LSG)) != kGStepFlags) return false;
  if (d->n_att != OCT_LOSS_MAX_ATT) return false;
  for (int k = 0; k < OCT_LOSS_MAX_ATT - 1; ++k)
    if (d->att_weight[k] == 0.f) return false;
  return true;
}

// one wave of 8-warp blocks (2 resident blocks per SM by the launch bounds), never more blocks than cells / 8
int fast_grid(const OctaveLossDesc* d, bool maps) {
  if (!maps) return 1;
  static int sms = 0;
  if (!sms) {
    sms = octave_sm_count();
    if (sms <= 0) sms = 148;
  }
  const long long total = (long long)(d->H >> 4) * (d->W >> 4) * d->B;
  long long need = (total + 7) / 8;
  const long long wave = 3LL * sms;
  if (need > wave) need = wave;
  return (int)(need < 1 ? 1 : need);
}

void fill_args(LossArgs& a, const OctaveLossDesc* d, const void* yhat, const void* ys, const void* const* att,
               const float* d_real, const float* d_fake, void* stats) {
  a.yhat = yhat; a.ys = ys; a.d_real = d_real; a.d_fake = d_fake;
  a.stats = reinterpret_cast<double*>(stats);
  a.B = d->B; a.C = d->C; a.H = d->H; a.W = d->W; a.flags = d->flags;
  a.n_att = (d->flags & OCT_LOSS_KLD) ? d->n_att : 0;
  for (int k = 0; k < OCT_LOSS_MAX_ATT; ++k) {
    a.att[k] = (att && k < a.n_att) ? att[k] : nullptr;
    a.ah[k] = d->att_h[k]; a.aw[k] = d->att_w[k];
    a.g_att[k] = nullptr;
  }
  for (int k = 0; k < OCT_LOSS_MAX_ATT - 1; ++k) a.aw8[k] = (k + 1 < a.n_att) ? d->att_weight[k] : 0.f;
  a.inv_sumw = d->sum_weights != 0.f ? 1.f / d->sum_weights : 0.f;
  a.wpce_scale = d->wpce_scale; a.dice_eps = d->dice_eps; a.jsd_eps = d->jsd_eps;
  a.n_real = d->n_real; a.n_fake = d->n_fake;
  a.out = nullptr; a.gscale = nullptr; a.g_yhat = nullptr; a.g_real = nullptr; a.g_fake = nullptr;
}

}  // namespace

extern "C" size_t octave_loss_stats_bytes(const OctaveLossDesc* d) {
  if (!d) return 0;
  int B = d->B > 0 ? d->B : 0;
  return (size_t)(ST_DICE + 2 * B) * sizeof(double);
}

extern "C" int octave_loss_uses_fast_path(const OctaveLossDesc* d) { return d && fast_ok(d) ? 1 : 0; }

extern "C" int octave_loss_fwd(const OctaveLossDesc* d, const void* yhat, const void* ys, const void* const* att,
                               const float* d_real, const float* d_fake, void* stats, float* out, void* stream) {
  int rc = validate(d);
  if (rc != OCT_OK) return rc;
  if (!stats || !out) return OCT_ERR_INVALID;
  if ((d->flags & (OCT_LOSS_WPCE | OCT_LOSS_DICE)) && (!yhat || !ys)) return OCT_ERR_INVALID;
  if ((d->flags & OCT_LOSS_KLD) && !att) return OCT_ERR_INVALID;
  if ((d->flags & (OCT_LOSS_LSG | OCT_LOSS_LSD)) && !d_fake) return OCT_ERR_INVALID;
  if ((d->flags & OCT_LOSS_LSD) && !d_real) return OCT_ERR_INVALID;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  LossArgs a;
  fill_args(a, d, yhat, ys, att, d_real, d_fake, stats);
  a.out = out;
  if (cudaMemsetAsync(stats, 0, octave_loss_stats_bytes(d), s) != cudaSuccess) return OCT_ERR_LAUNCH;
  const bool maps = d->flags & (OCT_LOSS_WPCE | OCT_LOSS_DICE | OCT_LOSS_KLD);
  if (fast_ok(d)) {
    const int grid = fast_grid(d, maps);
    // the G-step configuration (WPCE + KLD on logits, full 5-level pyramid) runs a kernel specialised at compile time
    if (gstep_config(d)) {
      if (d->dtype == OCT_DTYPE_F32) loss_fast_fwd_kernel<float, kGStepFlags, true><<<grid, 256, 0, s>>>(a);
      else loss_fast_fwd_kernel<bf16, kGStepFlags, true><<<grid, 256, 0, s>>>(a);
    } else {
      if (d->dtype == OCT_DTYPE_F32) loss_fast_fwd_kernel<float, -1, false><<<grid, 256, 0, s>>>(a);
      else loss_fast_fwd_kernel<bf16, -1, false><<<grid, 256, 0, s>>>(a);
    }
  } else {
    int gx = (d->H * d->W + 255) / 256;
    if (gx > 1024) gx = 1024;
    loss_generic_fwd_kernel<<<dim3(gx, d->B), 256, 0, s>>>(a);
  }
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_loss_bwd(const OctaveLossDesc* d, const void* yhat, const void* ys, const void* const* att,
                               const float* d_real, const float* d_fake, const void* stats, const float* gscale,
                               void* g_yhat, void* const* g_att, float* g_real, float* g_fake, void* stream) {
  int rc = validate(d);
  if (rc != OCT_OK) return rc;
  if (!stats || !gscale) return OCT_ERR_INVALID;
  if ((d->flags & (OCT_LOSS_WPCE | OCT_LOSS_DICE)) && (!yhat || !ys || !g_yhat)) return OCT_ERR_INVALID;
  if ((d->flags & OCT_LOSS_KLD) && (!att || !g_att)) return OCT_ERR_INVALID;
  if ((d->flags & (OCT_LOSS_LSG | OCT_LOSS_LSD)) && (!d_fake || !g_fake)) return OCT_ERR_INVALID;
  if ((d->flags & OCT_LOSS_LSD) && (!d_real || !g_real)) return OCT_ERR_INVALID;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  LossArgs a;
  fill_args(a, d, yhat, ys, att, d_real, d_fake, const_cast<void*>(stats));
  a.gscale = gscale; a.g_yhat = g_yhat; a.g_real = g_real; a.g_fake = g_fake;
  if (d->flags & OCT_LOSS_KLD)
    for (int k = 0; k < a.n_att; ++k) {
      if (!g_att[k]) return OCT_ERR_INVALID;
      a.g_att[k] = g_att[k];
    }
  const bool maps = d->flags & (OCT_LOSS_WPCE | OCT_LOSS_DICE | OCT_LOSS_KLD);
  if (fast_ok(d)) {
    const int grid = fast_grid(d, maps);
    if (gstep_config(d)) {
      if (d->dtype == OCT_DTYPE_F32) loss_fast_bwd_kernel<float, kGStepFlags, true><<<grid, 256, 0, s>>>(a);
      else loss_fast_bwd_kernel<bf16, kGStepFlags, true><<<grid, 256, 0, s>>>(a);
    } else {
      if (d->dtype == OCT_DTYPE_F32) loss_fast_bwd_kernel<float, -1, false><<<grid, 256, 0, s>>>(a);
      else loss_fast_bwd_kernel<bf16, -1, false><<<grid, 256, 0, s>>>(a);
    }
  } else {
    // coarse-level gradients are accumulated with atomics: zero them first
    if (d->flags & OCT_LOSS_KLD)
      for (int k = 1; k < a.n_att; ++k) {
        size_t bytes = (size_t)d->B * d->C * d->att_h[k] * d->att_w[k] * sizeof(float);
        if (cudaMemsetAsync(a.g_att[k], 0, bytes, s) != cudaSuccess) return OCT_ERR_LAUNCH;
      }
    int gx = (d->H * d->W + 255) / 256;
    if (gx > 1024) gx = 1024;
    loss_generic_bwd_kernel<<<dim3(gx, d->B), 256, 0, s>>>(a);
  }
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

// ---- fused single pass (G-step) ---------------------------------------------------------------------------------
extern "C" int octave_loss_fused_supported(const OctaveLossDesc* d) {
  if (validate(d) != OCT_OK || !fast_ok(d)) return 0;
  const int allowed = OCT_LOSS_WPCE | OCT_LOSS_KLD | OCT_LOSS_LSG | OCT_LOSS_FROM_LOGITS | OCT_LOSS_WPCE_FULL;
  if (d->flags & ~allowed) return 0;
  if (!(d->flags & (OCT_LOSS_WPCE | OCT_LOSS_KLD))) return 0;
  if (d->C != 2 || (d->H & 15) || (d->W & 15)) return 0;
  return 1;
}

extern "C" size_t octave_loss_fused_stats_bytes(const OctaveLossDesc* d) {
  if (!d) return 0;
  const int B = d->B > 0 ? d->B : 0;
  return (size_t)(ST_DICE + 2 * B + kFusedMaxBlocks * kFusedPartials) * sizeof(double);
}

extern "C" int octave_loss_fused(const OctaveLossDesc* d, const void* yhat, const void* ys, const void* const* att,
                                 const float* d_fake, const float* lambdas, void* stats, float* out, void* g_yhat,
                                 void* const* g_att, float* g_fake, void* stream) {
  if (!octave_loss_fused_supported(d)) return OCT_ERR_UNSUPPORTED;
  if (!stats || !out || !lambdas) return OCT_ERR_INVALID;
  if ((d->flags & OCT_LOSS_WPCE) && (!yhat || !ys || !g_yhat)) return OCT_ERR_INVALID;
  if ((d->flags & OCT_LOSS_KLD) && (!att || !g_att)) return OCT_ERR_INVALID;
  if ((d->flags & OCT_LOSS_LSG) && (!d_fake || !g_fake)) return OCT_ERR_INVALID;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  LossArgs a;
  fill_args(a, d, yhat, ys, att, nullptr, d_fake, stats);
  a.out = out; a.g_yhat = g_yhat; a.g_fake = g_fake;
  a.lam_wpce = lambdas[0]; a.lam_kld = lambdas[1]; a.lam_lsg = lambdas[2];
  if (d->flags & OCT_LOSS_KLD) {
    for (int k = 0; k < a.n_att; ++k) {
      if (!att[k] || !g_att[k]) return OCT_ERR_INVALID;
      a.g_att[k] = g_att[k];
    }
  }
  // (no memset: the first ST_DICE doubles of `stats` are zero on entry by contract and the kernel leaves them zero)
  int sms = octave_sm_count();
  if (sms <= 0) sms = 148;
  const long long plane = (long long)d->H * d->W;
  long long grid = 4LL * sms;
  const long long items = (plane * d->B + 2047) / 2048;
  if (grid > items) grid = items;
  if (grid > kFusedMaxBlocks) grid = kFusedMaxBlocks;
  if (grid < 1) grid = 1;
  // Cooperative single launch (count phase, KLD units, then WPCE once every block's count is in) when every block of the grid
  // is resident at once; OCTAVE_LOSS_COOP=0 selects the two-launch form (labels pre-pass kernel + fused kernel).  (A first
  // version with a grid-wide barrier right after the count phase was 1.5 us slower than the two launches.)
  static const int coop_env = [] { const char* e = getenv("OCTAVE_LOSS_COOP"); return e ? atoi(e) : 1; }();
  bool coop = false;
  if (coop_env && (d->flags & OCT_LOSS_WPCE)) {
    static int per_sm_cache[2] = {-1, -1};     // resident blocks per SM of the fp32 / bf16 kernel (0: no cooperative launch)
    int& per_sm = per_sm_cache[d->dtype == OCT_DTYPE_F32 ? 0 : 1];
    if (per_sm < 0) {
      int dev = 0, can = 0, n = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&can, cudaDevAttrCooperativeLaunch, dev);
      const void* fn = d->dtype == OCT_DTYPE_F32 ? (const void*)loss_fused_kernel<float, false, true> : (const void*)loss_fused_kernel<bf16, false, true>;
      per_sm = (can && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, 256, 0) == cudaSuccess) ? n : 0;
    }
    if ((long long)per_sm * sms >= grid) coop = true;
  }
  if ((d->flags & OCT_LOSS_WPCE) && !coop) {
    const long long groups = plane / 8 * 2 * d->B;
    long long gx = (groups + 256 * 8 - 1) / (256 * 8);
    if (gx > 6LL * sms) gx = 6LL * sms;
    if (d->dtype == OCT_DTYPE_F32)
      loss_label_count_kernel<float><<<(int)gx, 256, 0, s>>>(reinterpret_cast<const float*>(ys), plane / 8, groups, a.stats);
    else
      loss_label_count_kernel<bf16><<<(int)gx, 256, 0, s>>>(reinterpret_cast<const bf16*>(ys), plane / 8, groups, a.stats);
    OCT_CHECK_LAUNCH();
  }
  PlaneGeo g{};
  g.cw.init((unsigned)d->W >> 4);
  g.cells.init(((unsigned)d->W >> 4) * ((unsigned)d->H >> 4));
  g.units = g.cells.d * 2u * (unsigned)d->B;
  g.W = (unsigned)d->W; g.plane = (unsigned)plane;
  bool pyr = (d->flags & OCT_LOSS_KLD) && d->n_att == OCT_LOSS_MAX_ATT;
  for (int k = 0; k < 4; ++k) {
    g.use[k] = a.n_att > k + 1 && a.aw8[k] != 0.f;
    pyr = pyr && g.use[k];
  }
  if ((long long)d->B * 2 * plane >= (1LL << 31)) return OCT_ERR_UNSUPPORTED;   // 32-bit element offsets
  if (coop) {
    void* args[2] = {&a, &g};
    const void* fn;
    if (d->dtype == OCT_DTYPE_F32) fn = pyr ? (const void*)loss_fused_kernel<float, true, true> : (const void*)loss_fused_kernel<float, false, true>;
    else fn = pyr ? (const void*)loss_fused_kernel<bf16, true, true> : (const void*)loss_fused_kernel<bf16, false, true>;
    if (cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(256), args, 0, s) != cudaSuccess) return OCT_ERR_LAUNCH;
  } else if (d->dtype == OCT_DTYPE_F32) {
    if (pyr) loss_fused_kernel<float, true, false><<<(int)grid, 256, 0, s>>>(a, g);
    else loss_fused_kernel<float, false, false><<<(int)grid, 256, 0, s>>>(a, g);
  } else {
    if (pyr) loss_fused_kernel<bf16, true, false><<<(int)grid, 256, 0, s>>>(a, g);
    else loss_fused_kernel<bf16, false, false><<<(int)grid, 256, 0, s>>>(a, g);
  }
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_loss_scale_grads(const OctaveLossDesc* d, const float* g_total, void* g_yhat, void* const* g_att,
                                       float* g_fake, void* stream) {
  if (!d || !g_total) return OCT_ERR_INVALID;
  ScaleArgs sa{};
  sa.g = g_total; sa.dtype = d->dtype;
  int c = 0;
  if ((d->flags & OCT_LOSS_WPCE) && g_yhat) { sa.ptr[c] = g_yhat; sa.n[c] = (long long)d->B * d->C * d->H * d->W; ++c; }
  if ((d->flags & OCT_LOSS_KLD) && g_att)
    for (int k = 0; k < d->n_att && k < OCT_LOSS_MAX_ATT; ++k)
      if (g_att[k]) { sa.ptr[c] = g_att[k]; sa.n[c] = (long long)d->B * d->C * d->att_h[k] * d->att_w[k]; ++c; }
  // the last entry is always the fp32 critic-logit gradient (possibly empty)
  sa.ptr[c] = g_fake; sa.n[c] = ((d->flags & OCT_LOSS_LSG) && g_fake) ? d->n_fake : 0; ++c;
  sa.count = c;
  int sms = octave_sm_count();
  if (sms <= 0) sms = 148;
  loss_scale_kernel<<<2 * sms, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(sa);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

// ---- the two remaining branches of WeightedPartialCE.forward (segmentor/losses.py:40-49,56-59), fp32 maps ------------------
//   mode 0: manual=False, C == 2   nn.CrossEntropyLoss()(z, ys[:,1]) with z = y_hat * ys (y_hat when `full`) taken as logits
//   mode 1: num_classes == 1       nn.BCEWithLogitsLoss()(z, ys), same z
// Both are means over all B*H*W pixels (`reduction` and the class weights are ignored by the reference there).
namespace {
__global__ void __launch_bounds__(256) wpce_alt_fwd_kernel(const float* __restrict__ yh, const float* __restrict__ ys, int mode, int full,
                                                           long long npix, long long plane, double* acc, float* out) {
  __shared__ float red[8];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < npix; i += (long long)gridDim.x * 256) {
    if (mode == 0) {
      const long long b = i / plane, o = b * 2 * plane + (i - b * plane);
      const float t0 = ys[o], t1 = ys[o + plane];
      const float z0 = full ? yh[o] : yh[o] * t0, z1 = full ? yh[o + plane] : yh[o + plane] * t1;
      const float mx = fmaxf(z0, z1);
      const float lse = mx + logf(expf(z0 - mx) + expf(z1 - mx));
      s += lse - (t1 != 0.f ? z1 : z0);            // target class index = (long) ys[:,1]
    } else {
      const float t = ys[i];
      const float z = full ? yh[i] : yh[i] * t;
      s += fmaxf(z, 0.f) - z * t + log1pf(expf(-fabsf(z)));
    }
  }
  float v[1] = {s};
  block_sum<1>(v, red);
  if (threadIdx.x == 0) {
    atomicAdd(acc, (double)v[0]);
    __threadfence();
    const unsigned long long prev = atomicAdd(reinterpret_cast<unsigned long long*>(acc + 1), 1ULL);
    if (prev == (unsigned long long)gridDim.x - 1ULL) {
      __threadfence();
      out[0] = (float)(*reinterpret_cast<volatile double*>(acc) / (double)npix);
    }
  }
}

__global__ void __launch_bounds__(256) wpce_alt_bwd_kernel(const float* __restrict__ yh, const float* __restrict__ ys, int mode, int full,
                                                           long long npix, long long plane, const float* __restrict__ gscale,
                                                           float* __restrict__ g) {
  const float k = gscale[0] / (float)npix;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < npix; i += (long long)gridDim.x * 256) {
    if (mode == 0) {
      const long long b = i / plane, o = b * 2 * plane + (i - b * plane);
      const float t0 = ys[o], t1 = ys[o + plane];
      const float m0 = full ? 1.f : t0, m1 = full ? 1.f : t1;
      const float z0 = yh[o] * m0, z1 = yh[o + plane] * m1;
      const float mx = fmaxf(z0, z1);
      const float e0 = expf(z0 - mx), e1 = expf(z1 - mx), inv = 1.f / (e0 + e1);
      const bool c1 = t1 != 0.f;
      g[o] = k * m0 * (e0 * inv - (c1 ? 0.f : 1.f));
      g[o + plane] = k * m1 * (e1 * inv - (c1 ? 1.f : 0.f));
    } else {
      const float t = ys[i];
      const float m = full ? 1.f : t;
      const float z = yh[i] * m;
      g[i] = k * m * (1.f / (1.f + expf(-z)) - t);
    }
  }
}
}  // namespace

extern "C" int octave_wpce_alt_fwd(int32_t mode, const float* yhat, const float* ys, int32_t B, int32_t C, int32_t H, int32_t W,
                                   int32_t full, double* scratch2, float* out, void* stream) {
  if (!yhat || !ys || !scratch2 || !out || B <= 0 || H <= 0 || W <= 0) return OCT_ERR_INVALID;
  if ((mode == 0 && C != 2) || (mode == 1 && C != 1) || (mode != 0 && mode != 1)) return OCT_ERR_UNSUPPORTED;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(scratch2, 0, 2 * sizeof(double), s) != cudaSuccess) return OCT_ERR_LAUNCH;
  const long long plane = (long long)H * W, npix = plane * B;
  long long gx = (npix + 255) / 256;
  if (gx > 1184) gx = 1184;
  wpce_alt_fwd_kernel<<<(int)gx, 256, 0, s>>>(yhat, ys, mode, full, npix, plane, scratch2, out);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_wpce_alt_bwd(int32_t mode, const float* yhat, const float* ys, int32_t B, int32_t C, int32_t H, int32_t W,
                                   int32_t full, const float* gscale, float* g_yhat, void* stream) {
  if (!yhat || !ys || !gscale || !g_yhat || B <= 0 || H <= 0 || W <= 0) return OCT_ERR_INVALID;
  if ((mode == 0 && C != 2) || (mode == 1 && C != 1) || (mode != 0 && mode != 1)) return OCT_ERR_UNSUPPORTED;
  const long long plane = (long long)H * W, npix = plane * B;
  long long gx = (npix + 255) / 256;
  if (gx > 2368) gx = 2368;
  wpce_alt_bwd_kernel<<<(int)gx, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(yhat, ys, mode, full, npix, plane, gscale, g_yhat);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}
