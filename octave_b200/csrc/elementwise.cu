// K4/K5/K7 — bandwidth-bound glue around the convolutions: BatchNorm (train/eval, forward/backward),
// split-attention combine, residual add + ReLU.  NHWC, 8 channels (16 B of bf16 / 32 B of fp32) per thread.
//
// Thread mapping shared by every kernel here: a block of `bs` threads is `bs/G` pixel lanes x G channel
// groups (G = C/8); a thread keeps the SAME 8 channels for its whole pixel loop, so per-channel constants
// (scale/shift, mean/invstd, attention) are loaded once and per-channel reductions stay in registers until
// one shared-memory fold + one atomic per channel per block.  blockIdx.y = image.
//
// Reference arithmetic: nn.BatchNorm2d / ReLU / SplAtConv2d.forward (/root/reference/architectures/extra/resnest.py:97-138),
// residual adds at resnest.py:42,264-265.
#include <stdlib.h>

#include "common.cuh"
#include "../../include/octave_b200.h"

namespace {

// Geometry shared by every kernel here.  G = C/8 channel groups; a block covers Gb = min(G, 256) of them (blockIdx.z
// selects the 256-group slab when G > 256) with ppb = bs/Gb pixel lanes.  Every loop iteration of a block handles U*ppb
// consecutive pixels: all U 16-byte loads of a thread are issued before the first use (memory-level parallelism), and a
// block reads U*ppb*C*2 contiguous bytes per tensor.
struct Geo {
  int G, Gb, bs, ppb;
  dim3 grid;
};

// Resident blocks per SM of a kernel (cached): grids are sized to ONE wave so that reductions flush their atomics once
// per resident block and no tail wave runs at partial occupancy.
int resident_blocks(const void* fn, int bs, size_t smem) {
  struct Entry { const void* fn; int bs; size_t smem; int n; };
  static Entry cache[64];
  static int used = 0;
  for (int i = 0; i < used; ++i)
    if (cache[i].fn == fn && cache[i].bs == bs && cache[i].smem == smem) return cache[i].n;
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, bs, smem) != cudaSuccess || n < 1) n = 1;
  if (used < 64) cache[used++] = Entry{fn, bs, smem, n};
  return n;
}

int sm_count() {
  static int sms = 0;
  if (!sms) {
    sms = octave_sm_count();
    if (sms <= 0) sms = 148;
  }
  return sms;
}

// per_image: blockIdx.y = image and the pixel domain of a block is one image; otherwise the whole batch is one domain.
// total_blocks: upper bound of the grid (all of x, y, z together).
bool make_geo(const OctaveAct* a, Geo* g, int U, bool per_image, int total_blocks, int min_iters = 1) {
  if (a->C % 8) return false;
  g->G = a->C / 8;
  int gz = 1;
  if (g->G <= 256) {
    g->Gb = g->G;
    g->bs = (256 / g->G) * g->G;
  } else {
    if (g->G % 256) return false;
    g->Gb = 256;
    g->bs = 256;
    gz = g->G / 256;
  }
  g->ppb = g->bs / g->Gb;
  const long long hw = (long long)a->H * a->W;
  const long long n = per_image ? hw : hw * a->B;
  const int gy = per_image ? a->B : 1;
  long long bx = (n + (long long)g->ppb * U - 1) / ((long long)g->ppb * U);
  if (min_iters > 1) bx = (bx + min_iters - 1) / min_iters;  // reductions: amortise the per-block fold + atomics
  long long cap = total_blocks / ((long long)gy * gz);
  if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  g->grid = dim3((unsigned)bx, (unsigned)gy, (unsigned)gz);
  return true;
}

bool same_shape(const OctaveAct* a, const OctaveAct* b) {
  return a->B == b->B && a->H == b->H && a->W == b->W && a->C == b->C && a->dtype == b->dtype;
}
bool view_ok(const OctaveAct* a) {
  if (!a || !a->data) return false;
  if (a->dtype != OCT_DTYPE_F32 && a->dtype != OCT_DTYPE_BF16) return false;
  if (a->C % 8 || a->ld % 8 || a->coff % 8) return false;
  return a->B > 0 && a->H > 0 && a->W > 0 && a->B < 65536;
}

template <typename T>
__device__ __forceinline__ T* at(const OctaveAct& a, long long pix, int c) {
  return reinterpret_cast<T*>(a.data) + pix * a.ld + a.coff + c;
}

struct Tix {
  int cg, lane, ppb;
  long long base, n, first, stride;
};
template <int U>
__device__ __forceinline__ Tix make_tix(const OctaveAct& x, int Gb) {
  Tix t;
  t.cg = blockIdx.z * Gb + threadIdx.x % Gb;
  t.lane = threadIdx.x / Gb;
  t.ppb = blockDim.x / Gb;
  const long long hw = (long long)x.H * x.W;
  if (gridDim.y > 1 || x.B == 1) { t.base = (long long)blockIdx.y * hw; t.n = hw; }
  else { t.base = 0; t.n = hw * x.B; }
  t.first = (long long)blockIdx.x * t.ppb * U + t.lane;
  t.stride = (long long)gridDim.x * t.ppb * U;
  return t;
}

// ---------------------------------------------------------------------------------------------------
constexpr int U_STATS = 8;
template <typename T>
__global__ void __launch_bounds__(256, 4) chan_stats_kernel(const OctaveAct x, int Gb, double* sums) {
  extern __shared__ float sm[];
  const Tix t = make_tix<U_STATS>(x, Gb);
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.f;
  for (long long p = t.first; p < t.n; p += t.stride) {
    Raw8<T> r[U_STATS];
#pragma unroll
    for (int u = 0; u < U_STATS; ++u) {
      const long long q = p + (long long)u * t.ppb;
      if (q < t.n) r[u].ld(at<T>(x, t.base + q, t.cg * 8));
    }
#pragma unroll
    for (int u = 0; u < U_STATS; ++u) {
      const long long q = p + (long long)u * t.ppb;
      if (q < t.n) {
        float f[8];
        r[u].get(f);
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[i] += f[i]; v[8 + i] += f[i] * f[i]; }
      }
    }
  }
  fold_lanes<16>(v, sm, Gb);
  if (threadIdx.x < Gb) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(sums + t.cg * 8 + i, (double)v[i]);
      atomicAdd(sums + x.C + t.cg * 8 + i, (double)v[8 + i]);
    }
  }
}

__global__ void bn_prepare_kernel(int C, double count, const double* sums, const float* gamma, const float* beta,
                                  float* rm, float* rv, long long* nbt, float eps, float mom, int training, float* ab,
                                  float* mi) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, invstd;
  if (training) {
    const double m = sums[c] / count;
    double var = sums[C + c] / count - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    invstd = (float)(1.0 / sqrt(var + (double)eps));
    if (rm) rm[c] = (1.f - mom) * rm[c] + mom * mean;
    if (rv) {
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      rv[c] = (1.f - mom) * rv[c] + mom * (float)unbiased;
    }
    if (c == 0 && nbt) *nbt += 1;
  } else {
    mean = rm[c];
    invstd = 1.f / sqrtf(rv[c] + eps);
  }
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  ab[c] = g * invstd;
  ab[C + c] = b - mean * g * invstd;
  mi[c] = mean;
  mi[C + c] = invstd;
}

constexpr int U_AFF = 4;
constexpr int kGapMaxBlocks = 64;   // blocks per image of a launch with the fused global average pool
constexpr int kGapCounterWords = 65536;   // fixed-size counter region (one word per image, B < 65536) ahead of the partial sums
template <typename T, bool HAS_RES, bool HAS_GAP>
__global__ void __launch_bounds__(256) affine_act_kernel(const OctaveAct x, int Gb, const float* ab, const OctaveAct res,
                                                         int relu, const OctaveAct y, float* gap, unsigned* gap_ws) {
  extern __shared__ float sm[];
  const Tix t = make_tix<U_AFF>(x, Gb);
  float a[8], b[8], acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = ab ? ab[t.cg * 8 + i] : 1.f;
    b[i] = ab ? ab[x.C + t.cg * 8 + i] : 0.f;
    acc[i] = 0.f;
  }
  for (long long p = t.first; p < t.n; p += t.stride) {
    Raw8<T> rx[U_AFF], rr[HAS_RES ? U_AFF : 1];
#pragma unroll
    for (int u = 0; u < U_AFF; ++u) {
      const long long q = p + (long long)u * t.ppb;
      if (q < t.n) {
        rx[u].ld(at<T>(x, t.base + q, t.cg * 8));
        if (HAS_RES) rr[u].ld(at<T>(res, t.base + q, t.cg * 8));
      }
    }
#pragma unroll
    for (int u = 0; u < U_AFF; ++u) {
      const long long q = p + (long long)u * t.ppb;
      if (q < t.n) {
        float f[8];
        rx[u].get(f);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = f[i] * a[i] + b[i];
        if (HAS_RES) {
          float r[8];
          rr[u].get(r);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] += r[i];
        }
        if (relu) {
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
        }
        if (HAS_GAP) {
          // accumulate what the consumer will read back (storage-rounded values)
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += to_f(from_f<T>(f[i]));
        }
        VecIO<T, 8>::st(at<T>(y, t.base + q, t.cg * 8), f);
      }
    }
  }
  if (HAS_GAP) {
    // Deterministic two-stage reduction (no floating-point atomics): every block stores its per-channel partial sums
    // (radix halves already added) into the workspace; the LAST block of an image to finish adds the partials in block
    // order.  gap_ws = [kGapCounterWords counters (zero between launches: the last block resets its own)] [B][gridDim.x][C/2] floats.
    __shared__ int s_last;
    fold_lanes<8>(acc, sm, Gb);
    __syncthreads();
    if (threadIdx.x < Gb) {
#pragma unroll
      for (int i = 0; i < 8; ++i) sm[threadIdx.x * 8 + i] = acc[i];
    }
    __syncthreads();
    const int half = x.C >> 1, img = blockIdx.y;
    float* part = reinterpret_cast<float*>(gap_ws + kGapCounterWords) + ((long long)img * gridDim.x + blockIdx.x) * half;
    for (int c = threadIdx.x; c < half; c += blockDim.x) part[c] = sm[c] + sm[c + half];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(gap_ws + img, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last) {
      __threadfence();
      const float* all = reinterpret_cast<const float*>(gap_ws + kGapCounterWords) + (long long)img * gridDim.x * half;
      for (int c = threadIdx.x; c < half; c += blockDim.x) {
        float s = 0.f;
        unsigned bx = 0;
        for (; bx + 8 <= gridDim.x; bx += 8) {   // 8 independent L2 loads in flight, added in block order
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __ldcg(all + (long long)(bx + i) * half + c);
#pragma unroll
          for (int i = 0; i < 8; ++i) s += v[i];
        }
        for (; bx < gridDim.x; ++bx) s += __ldcg(all + (long long)bx * half + c);
        gap[(long long)img * half + c] = s;
      }
      if (threadIdx.x == 0) gap_ws[img] = 0u;
    }
  }
}

// MASK: 0 none, 1 mask tensor (dz = dy * (mask > 0)), 2 recompute the ReLU mask of this very BN from x: (x*a+b > 0)
constexpr int U_BNR = 4;
template <typename T, int MASK>
__global__ void __launch_bounds__(256, 2) bn_bwd_reduce_kernel(const OctaveAct dy, const OctaveAct mask, int Gb,
                                                               const float* ab, const OctaveAct x, const float* mi,
                                                               double* sums2, const OctaveAct dmasked, int has_dm) {
  extern __shared__ float sm[];
  const Tix t = make_tix<U_BNR>(x, Gb);
  float mean[8], v[16], aa[8], bb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mean[i] = mi[t.cg * 8 + i];
    aa[i] = MASK == 2 ? ab[t.cg * 8 + i] : 0.f;
    bb[i] = MASK == 2 ? ab[x.C + t.cg * 8 + i] : 0.f;
    v[i] = v[8 + i] = 0.f;
  }
  for (long long p = t.first; p < t.n; p += t.stride) {
    Raw8<T> rd[U_BNR], rx[U_BNR], rm[MASK == 1 ? U_BNR : 1];
#pragma unroll
    for (int u = 0; u < U_BNR; ++u) {
      const long long q = p + (long long)u * t.ppb;
      if (q < t.n) {
        rd[u].ld(at<T>(dy, t.base + q, t.cg * 8));
        rx[u].ld(at<T>(x, t.base + q, t.cg * 8));
        if (MASK == 1) rm[u].ld(at<T>(mask, t.base + q, t.cg * 8));
      } else {
        rd[u].zero();
        rx[u].zero();
        if (MASK == 1) rm[u].zero();
      }
    }
#pragma unroll
    for (int u = 0; u < U_BNR; ++u) {
      float d[8], f[8];
      rd[u].get(d);
      rx[u].get(f);
      if (MASK == 1) {
        float m[8];
        rm[u].get(m);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = m[i] > 0.f ? d[i] : 0.f;
      } else if (MASK == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = to_f(from_f<T>(f[i] * aa[i] + bb[i])) > 0.f ? d[i] : 0.f;
      }
      // dy * (mask > 0) leaves with this pass (it is the residual branch's gradient, and the apply pass then reads it
      // instead of dy AND the mask: one tensor pass less per BatchNorm backward)
      if (has_dm && p + (long long)u * t.ppb < t.n) VecIO<T, 8>::st(at<T>(dmasked, t.base + p + (long long)u * t.ppb, t.cg * 8), d);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i] += d[i];
        v[8 + i] += d[i] * (f[i] - mean[i]);
      }
    }
  }
  fold_lanes<16>(v, sm, Gb);
  if (threadIdx.x < Gb) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(sums2 + t.cg * 8 + i, (double)v[i]);
      atomicAdd(sums2 + x.C + t.cg * 8 + i, (double)v[8 + i] * (double)mi[x.C + t.cg * 8 + i]);
    }
  }
}

// dx = P*dz + Q*x + R with P = gamma*invstd, Q = -P*invstd*mean(dz*xhat), R = -P*mean(dz) - Q*mean
constexpr int U_BNA = 2;
template <typename T, int MASK>
__global__ void __launch_bounds__(256, sizeof(T) == 2 ? 3 : 2) bn_bwd_apply_kernel(const OctaveAct dy, const OctaveAct mask, int Gb,
                                                              const float* ab, const OctaveAct x, const float* mi,
                                                              const float* gamma, const double* sums2, int training,
                                                              const OctaveAct dx, float* dgamma, float* dbeta,
                                                              const OctaveAct dmasked, int has_dm) {
  const Tix t = make_tix<U_BNA>(x, Gb);
  const float inv_n = 1.f / ((float)x.B * (float)x.H * (float)x.W);
  float P[8], Q[8], R[8], aa[8], bb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = t.cg * 8 + i;
    const float mean = mi[c], inv = mi[x.C + c];
    aa[i] = MASK == 2 ? ab[c] : 0.f;
    bb[i] = MASK == 2 ? ab[x.C + c] : 0.f;
    const float sd = (float)sums2[c], sdx = (float)sums2[x.C + c];
    const float k1 = training ? sd * inv_n : 0.f;
    const float k2 = training ? sdx * inv_n : 0.f;
    P[i] = (gamma ? gamma[c] : 1.f) * inv;
    Q[i] = -P[i] * inv * k2;
    R[i] = -P[i] * k1 - Q[i] * mean;
    if (blockIdx.x == 0 && blockIdx.y == 0 && t.lane == 0) {
      if (dgamma) dgamma[c] = sdx;
      if (dbeta) dbeta[c] = sd;
    }
  }
  for (long long p = t.first; p < t.n; p += t.stride) {
    Raw8<T> rd[U_BNA], rx[U_BNA], rm[MASK == 1 ? U_BNA : 1];
#pragma unroll
    for (int u = 0; u < U_BNA; ++u) {
      const long long q = p + (long long)u * t.ppb;
      if (q < t.n) {
        rd[u].ld(at<T>(dy, t.base + q, t.cg * 8));
        rx[u].ld(at<T>(x, t.base + q, t.cg * 8));
        if (MASK == 1) rm[u].ld(at<T>(mask, t.base + q, t.cg * 8));
      }
    }
#pragma unroll
    for (int u = 0; u < U_BNA; ++u) {
      const long long q = p + (long long)u * t.ppb;
      if (q < t.n) {
        float d[8], f[8];
        rd[u].get(d);
        rx[u].get(f);
        if (MASK == 1) {
          float m[8];
          rm[u].get(m);
#pragma unroll
          for (int i = 0; i < 8; ++i) d[i] = m[i] > 0.f ? d[i] : 0.f;
        } else if (MASK == 2) {
#pragma unroll
          for (int i = 0; i < 8; ++i) d[i] = to_f(from_f<T>(f[i] * aa[i] + bb[i])) > 0.f ? d[i] : 0.f;
        }
        if (has_dm) VecIO<T, 8>::st(at<T>(dmasked, t.base + q, t.cg * 8), d);   // dy * (mask > 0): the residual branch's gradient
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = P[i] * d[i] + (Q[i] * f[i] + R[i]);
        VecIO<T, 8>::st(at<T>(dx, t.base + q, t.cg * 8), d);
      }
    }
  }
}

// OP 0: dst += src; OP 1: dst = src * (mask > 0)  (a = dst/dy, b = src/mask);
// OP 2..5: activation backward, a = dy, b = y (stored activation output): ReLU, LeakyReLU(0.2), sigmoid, tanh
constexpr int U_BIN = 4;
template <typename T, int OP>
__global__ void __launch_bounds__(256) binary_kernel(const OctaveAct a, const OctaveAct b, int Gb, const OctaveAct out) {
  const Tix t = make_tix<U_BIN>(a, Gb);
  for (long long p = t.first; p < t.n; p += t.stride) {
    Raw8<T> ra[U_BIN], rb[U_BIN];
#pragma unroll
    for (int u = 0; u < U_BIN; ++u) {
      const long long q = p + (long long)u * t.ppb;
      if (q < t.n) {
        ra[u].ld(at<T>(a, t.base + q, t.cg * 8));
        rb[u].ld(at<T>(b, t.base + q, t.cg * 8));
      }
    }
#pragma unroll
    for (int u = 0; u < U_BIN; ++u) {
      const long long q = p + (long long)u * t.ppb;
      if (q < t.n) {
        float f[8], g[8];
        ra[u].get(f);
        rb[u].get(g);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (OP == 0) f[i] = f[i] + g[i];
          else if (OP == 1 || OP == 2) f[i] = g[i] > 0.f ? f[i] : 0.f;
          else if (OP == 3) f[i] = g[i] > 0.f ? f[i] : 0.2f * f[i];
          else if (OP == 4) f[i] = f[i] * g[i] * (1.f - g[i]);
          else f[i] = f[i] * (1.f - g[i] * g[i]);
        }
        VecIO<T, 8>::st(at<T>(out, t.base + q, t.cg * 8), f);
      }
    }
  }
}

// ---- split attention (per image: blockIdx.y = image) -------------------------------------------------
constexpr int U_SPL = 2;
template <typename T>
__global__ void __launch_bounds__(256) splat_combine_kernel(const OctaveAct U, const float* att, int Gb, int relu,
                                                            const OctaveAct out) {
  const int C = out.C;
  const Tix t = make_tix<U_SPL>(out, Gb);
  float a0[8], a1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a0[i] = att[(long long)blockIdx.y * 2 * C + t.cg * 8 + i];
    a1[i] = att[(long long)blockIdx.y * 2 * C + C + t.cg * 8 + i];
  }
  for (long long p = t.first; p < t.n; p += t.stride) {
    Raw8<T> r0[U_SPL], r1[U_SPL];
#pragma unroll
    for (int u = 0; u < U_SPL; ++u) {
      const long long q = p + (long long)u * t.ppb;
      if (q < t.n) {
        r0[u].ld(at<T>(U, t.base + q, t.cg * 8));
        r1[u].ld(at<T>(U, t.base + q, C + t.cg * 8));
      }
    }
#pragma unroll
    for (int u = 0; u < U_SPL; ++u) {
      const long long q = p + (long long)u * t.ppb;
      if (q < t.n) {
        float u0[8], u1[8];
        r0[u].get(u0);
        r1[u].get(u1);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          u0[i] = a0[i] * u0[i] + a1[i] * u1[i];
          if (relu) u0[i] = fmaxf(u0[i], 0.f);
        }
        VecIO<T, 8>::st(at<T>(out, t.base + q, t.cg * 8), u0);
      }
    }
  }
}

template <typename T, bool HAS_MASK>
__global__ void __launch_bounds__(256, 2) splat_bwd_reduce_kernel(const OctaveAct dout, const OctaveAct mask, int Gb,
                                                                  const OctaveAct U, float* datt) {
  extern __shared__ float sm[];
  const int C = dout.C;
  const Tix t = make_tix<U_SPL>(dout, Gb);
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.f;
  for (long long p = t.first; p < t.n; p += t.stride) {
    Raw8<T> rd[U_SPL], r0[U_SPL], r1[U_SPL], rm[HAS_MASK ? U_SPL : 1];
#pragma unroll
    for (int u = 0; u < U_SPL; ++u) {
      const long long q = p + (long long)u * t.ppb;
      if (q < t.n) {
        rd[u].ld(at<T>(dout, t.base + q, t.cg * 8));
        r0[u].ld(at<T>(U, t.base + q, t.cg * 8));
        r1[u].ld(at<T>(U, t.base + q, C + t.cg * 8));
        if (HAS_MASK) rm[u].ld(at<T>(mask, t.base + q, t.cg * 8));
      } else {
        rd[u].zero(); r0[u].zero(); r1[u].zero();
        if (HAS_MASK) rm[u].zero();
      }
    }
#pragma unroll
    for (int u = 0; u < U_SPL; ++u) {
      float d[8], u0[8], u1[8];
      rd[u].get(d);
      r0[u].get(u0);
      r1[u].get(u1);
      if (HAS_MASK) {
        float m[8];
        rm[u].get(m);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = m[i] > 0.f ? d[i] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) { v[i] += d[i] * u0[i]; v[8 + i] += d[i] * u1[i]; }
    }
  }
  fold_lanes<16>(v, sm, Gb);
  if (threadIdx.x < Gb) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(datt + (long long)blockIdx.y * 2 * C + t.cg * 8 + i, v[i]);
      atomicAdd(datt + (long long)blockIdx.y * 2 * C + C + t.cg * 8 + i, v[8 + i]);
    }
  }
}

template <typename T, bool HAS_MASK>
__global__ void __launch_bounds__(256) splat_bwd_du_kernel(const OctaveAct dout, const OctaveAct mask, int Gb,
                                                           const float* att, const float* dgap, float gap_scale,
                                                           const OctaveAct dU) {
  const int C = dout.C;
  const Tix t = make_tix<U_SPL>(dout, Gb);
  float a0[8], a1[8], gg[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a0[i] = att[(long long)blockIdx.y * 2 * C + t.cg * 8 + i];
    a1[i] = att[(long long)blockIdx.y * 2 * C + C + t.cg * 8 + i];
    gg[i] = dgap ? dgap[(long long)blockIdx.y * C + t.cg * 8 + i] * gap_scale : 0.f;
  }
  for (long long p = t.first; p < t.n; p += t.stride) {
    Raw8<T> rd[U_SPL], rm[HAS_MASK ? U_SPL : 1];
#pragma unroll
    for (int u = 0; u < U_SPL; ++u) {
      const long long q = p + (long long)u * t.ppb;
      if (q < t.n) {
        rd[u].ld(at<T>(dout, t.base + q, t.cg * 8));
        if (HAS_MASK) rm[u].ld(at<T>(mask, t.base + q, t.cg * 8));
      }
    }
#pragma unroll
    for (int u = 0; u < U_SPL; ++u) {
      const long long q = p + (long long)u * t.ppb;
      if (q < t.n) {
        float d[8], o0[8], o1[8];
        rd[u].get(d);
        if (HAS_MASK) {
          float m[8];
          rm[u].get(m);
#pragma unroll
          for (int i = 0; i < 8; ++i) d[i] = m[i] > 0.f ? d[i] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { o0[i] = a0[i] * d[i] + gg[i]; o1[i] = a1[i] * d[i] + gg[i]; }
        VecIO<T, 8>::st(at<T>(dU, t.base + q, t.cg * 8), o0);
        VecIO<T, 8>::st(at<T>(dU, t.base + q, C + t.cg * 8), o1);
      }
    }
  }
}

// ---- split-attention backward fused with the backward of its BatchNorm (bn0) -----------------------------------------
// dU[p][r*C+c] = att[b][r*C+c] * dout[p][c] * (omask > 0) + dgap[b][c] * gap_scale   (splat_bwd_du) is never written:
// both BatchNorm-backward passes rebuild it from dout (C channels) while they stream z (2C channels), so the backward of
// SplAtConv2d.bn0 + relu (resnest.py:101-105) costs z twice + dz once instead of seven passes over 2C-wide tensors.
// blockIdx.y = image, blockIdx.z = radix half r.  The ReLU mask of bn0 is recomputed from z (y = z*a + b > 0).
constexpr int U_SBR = 4;   // reduce pass: loads per thread and tensor in flight
constexpr int U_SBA = 3;   // apply pass
template <typename T, bool HAS_OMASK>
__global__ void __launch_bounds__(256, 2) splat_bn_bwd_reduce_kernel(const OctaveAct dout, const OctaveAct omask, int Gb,
                                                                     const float* att, const float* dgap, float gap_scale,
                                                                     const OctaveAct z, const float* ab, const float* mi,
                                                                     double* sums2) {
  extern __shared__ float sm[];
  const int C = dout.C, C2 = 2 * C, r = blockIdx.z;
  Tix t = make_tix<U_SBR>(dout, Gb);
  t.cg = threadIdx.x % Gb;                   // blockIdx.z is the radix half here, not a channel slab
  const int c0 = r * C + t.cg * 8;           // first of this thread's 8 channels in the 2C-wide tensors
  float mean[8], v[16], aa[8], bb[8], at8[8], gg[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mean[i] = mi[c0 + i];
    aa[i] = ab[c0 + i];
    bb[i] = ab[C2 + c0 + i];
    at8[i] = att[(long long)blockIdx.y * C2 + c0 + i];
    gg[i] = dgap ? dgap[(long long)blockIdx.y * C + t.cg * 8 + i] * gap_scale : 0.f;
    v[i] = v[8 + i] = 0.f;
  }
  for (long long p = t.first; p < t.n; p += t.stride) {
    Raw8<T> rd[U_SBR], rz[U_SBR], rm[HAS_OMASK ? U_SBR : 1];
#pragma unroll
    for (int u = 0; u < U_SBR; ++u) {
      const long long q = p + (long long)u * t.ppb;
      if (q < t.n) {
        rd[u].ld(at<T>(dout, t.base + q, t.cg * 8));
        rz[u].ld(at<T>(z, t.base + q, c0));
        if (HAS_OMASK) rm[u].ld(at<T>(omask, t.base + q, t.cg * 8));
      }
    }
#pragma unroll
    for (int u = 0; u < U_SBR; ++u) {
      const long long q = p + (long long)u * t.ppb;
      if (q < t.n) {
        float d[8], f[8];
        rd[u].get(d);
        rz[u].get(f);
        if (HAS_OMASK) {
          float m[8];
          rm[u].get(m);
#pragma unroll
          for (int i = 0; i < 8; ++i) d[i] = m[i] > 0.f ? d[i] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          // dU rounded to storage precision like the unfused path stored it, then gated by bn0's own ReLU
          float du = to_f(from_f<T>(at8[i] * d[i] + gg[i]));
          du = to_f(from_f<T>(f[i] * aa[i] + bb[i])) > 0.f ? du : 0.f;
          v[i] += du;
          v[8 + i] += du * (f[i] - mean[i]);
        }
      }
    }
  }
  fold_lanes<16>(v, sm, Gb);
  if (threadIdx.x < Gb) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(sums2 + c0 + i, (double)v[i]);
      atomicAdd(sums2 + C2 + c0 + i, (double)v[8 + i] * (double)mi[C2 + c0 + i]);
    }
  }
}

template <typename T, bool HAS_OMASK>
__global__ void __launch_bounds__(256, 2) splat_bn_bwd_apply_kernel(const OctaveAct dout, const OctaveAct omask, int Gb,
                                                                    const float* att, const float* dgap, float gap_scale,
                                                                    const OctaveAct z, const float* ab, const float* mi,
                                                                    const float* gamma, const double* sums2, int training,
                                                                    const OctaveAct dz, float* dgamma, float* dbeta) {
  const int C = dout.C, C2 = 2 * C, r = blockIdx.z;
  Tix t = make_tix<U_SBA>(dout, Gb);
  t.cg = threadIdx.x % Gb;                   // blockIdx.z is the radix half here, not a channel slab
  const int c0 = r * C + t.cg * 8;
  const float inv_n = 1.f / ((float)z.B * (float)z.H * (float)z.W);
  float P[8], Q[8], R[8], aa[8], bb[8], at8[8], gg[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + i;
    const float mean = mi[c], inv = mi[C2 + c];
    aa[i] = ab[c];
    bb[i] = ab[C2 + c];
    at8[i] = att[(long long)blockIdx.y * C2 + c];
    gg[i] = dgap ? dgap[(long long)blockIdx.y * C + t.cg * 8 + i] * gap_scale : 0.f;
    const float sd = (float)sums2[c], sdx = (float)sums2[C2 + c];
    const float k1 = training ? sd * inv_n : 0.f;
    const float k2 = training ? sdx * inv_n : 0.f;
    P[i] = (gamma ? gamma[c] : 1.f) * inv;
    Q[i] = -P[i] * inv * k2;
    R[i] = -P[i] * k1 - Q[i] * mean;
    if (blockIdx.x == 0 && blockIdx.y == 0 && t.lane == 0) {
      if (dgamma) dgamma[c] = sdx;
      if (dbeta) dbeta[c] = sd;
    }
  }
  for (long long p = t.first; p < t.n; p += t.stride) {
    Raw8<T> rd[U_SBA], rz[U_SBA], rm[HAS_OMASK ? U_SBA : 1];
#pragma unroll
    for (int u = 0; u < U_SBA; ++u) {
      const long long q = p + (long long)u * t.ppb;
      if (q < t.n) {
        rd[u].ld(at<T>(dout, t.base + q, t.cg * 8));
        rz[u].ld(at<T>(z, t.base + q, c0));
        if (HAS_OMASK) rm[u].ld(at<T>(omask, t.base + q, t.cg * 8));
      }
    }
#pragma unroll
    for (int u = 0; u < U_SBA; ++u) {
      const long long q = p + (long long)u * t.ppb;
      if (q < t.n) {
        float d[8], f[8];
        rd[u].get(d);
        rz[u].get(f);
        if (HAS_OMASK) {
          float m[8];
          rm[u].get(m);
#pragma unroll
          for (int i = 0; i < 8; ++i) d[i] = m[i] > 0.f ? d[i] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float du = to_f(from_f<T>(at8[i] * d[i] + gg[i]));
          du = to_f(from_f<T>(f[i] * aa[i] + bb[i])) > 0.f ? du : 0.f;
          d[i] = P[i] * du + (Q[i] * f[i] + R[i]);
        }
        VecIO<T, 8>::st(at<T>(dz, t.base + q, c0), d);
      }
    }
  }
}

#define DISPATCH_T(dtype, ...)                         \
  do {                                                 \
    if ((dtype) == OCT_DTYPE_F32) { using T = float; __VA_ARGS__; } \
    else { using T = bf16; __VA_ARGS__; }              \
  } while (0)

// one wave of `fn`: resident blocks per SM x SMs
#define ONE_WAVE(fn, bs, smem) (resident_blocks(reinterpret_cast<const void*>(fn), (bs), (smem)) * sm_count())

}  // namespace

extern "C" int octave_chan_stats(const OctaveAct* x, double* sums, void* stream) {
  if (!view_ok(x) || !sums) return OCT_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  if (!g_octave_stats_prezeroed && cudaMemsetAsync(sums, 0, sizeof(double) * 2 * x->C, s) != cudaSuccess) return OCT_ERR_LAUNCH;
  Geo g;
  DISPATCH_T(x->dtype, {
    auto fn = chan_stats_kernel<T>;
    if (!make_geo(x, &g, U_STATS, false, ONE_WAVE(fn, 256, 16 * 256 * sizeof(float)), 4)) return OCT_ERR_UNSUPPORTED;
    fn<<<g.grid, g.bs, 16 * g.bs * sizeof(float), s>>>(*x, g.Gb, sums);
  });
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_bn_prepare(int32_t C, double count, const double* sums, const float* gamma, const float* beta,
                                 float* running_mean, float* running_var, int64_t* nbt, float eps, float momentum,
                                 int32_t training, float* ab, float* mean_invstd, void* stream) {
  if (C <= 0 || !ab || !mean_invstd) return OCT_ERR_INVALID;
  if (training && !sums) return OCT_ERR_INVALID;
  if (!training && (!running_mean || !running_var)) return OCT_ERR_INVALID;
  bn_prepare_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(C, count, sums, gamma, beta, running_mean,
                                                                     running_var, (long long*)nbt, eps, momentum,
                                                                     training, ab, mean_invstd);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

template <typename T, bool HAS_RES, bool HAS_GAP>
static int launch_affine(const OctaveAct* x, const float* ab, const OctaveAct* res, int relu, const OctaveAct* y,
                         float* gap, unsigned* gap_ws, cudaStream_t s) {
  auto fn = affine_act_kernel<T, HAS_RES, HAS_GAP>;
  const size_t smem = HAS_GAP ? 8 * 256 * sizeof(float) : 0;
  Geo g;
  if (!make_geo(x, &g, U_AFF, HAS_GAP, ONE_WAVE(fn, 256, smem), HAS_GAP ? 4 : 1)) return OCT_ERR_UNSUPPORTED;
  if (HAS_GAP && (g.grid.z != 1 || (int)g.grid.x > kGapMaxBlocks)) g.grid.x = kGapMaxBlocks;
  if (HAS_GAP && g.grid.z != 1) return OCT_ERR_UNSUPPORTED;   // C <= 2048: one block covers every channel group
  OctaveAct r = res ? *res : *x;
  fn<<<g.grid, g.bs, HAS_GAP ? 8 * g.bs * sizeof(float) : 0, s>>>(*x, g.Gb, ab, r, relu, *y, gap, gap_ws);
  return OCT_OK;
}

extern "C" size_t octave_affine_gap_ws_bytes(const OctaveAct* x) {
  if (!x || x->B <= 0 || x->C <= 0) return 0;
  return sizeof(unsigned) * kGapCounterWords + sizeof(float) * (size_t)x->B * kGapMaxBlocks * (x->C / 2);
}

extern "C" int octave_affine_act(const OctaveAct* x, const float* ab, const OctaveAct* res, int32_t relu,
                                 const OctaveAct* y, float* gap, void* gap_ws, void* stream) {
  if (!view_ok(x) || !view_ok(y) || !same_shape(x, y)) return OCT_ERR_INVALID;
  if (res && (!view_ok(res) || !same_shape(x, res))) return OCT_ERR_INVALID;
  if (gap && ((x->C % 16) || !gap_ws)) return OCT_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned* ws = reinterpret_cast<unsigned*>(gap_ws);
  int rc = OCT_OK;
  DISPATCH_T(x->dtype, {
    if (gap) rc = res ? launch_affine<T, true, true>(x, ab, res, relu, y, gap, ws, s) : launch_affine<T, false, true>(x, ab, res, relu, y, gap, ws, s);
    else rc = res ? launch_affine<T, true, false>(x, ab, res, relu, y, gap, ws, s) : launch_affine<T, false, false>(x, ab, res, relu, y, gap, ws, s);
  });
  if (rc != OCT_OK) return rc;
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

template <typename T, int MASK>
static int launch_bn_bwd_reduce(const OctaveAct* dy, const OctaveAct* m, const float* relu_ab, const OctaveAct* x,
                                const float* mi, double* sums2, const OctaveAct* dmasked, cudaStream_t s) {
  auto fn = bn_bwd_reduce_kernel<T, MASK>;
  Geo g;
  if (!make_geo(x, &g, U_BNR, false, ONE_WAVE(fn, 256, 16 * 256 * sizeof(float)), 4)) return OCT_ERR_UNSUPPORTED;
  fn<<<g.grid, g.bs, 16 * g.bs * sizeof(float), s>>>(*dy, *m, g.Gb, relu_ab, *x, mi, sums2, dmasked ? *dmasked : *dy, dmasked != nullptr);
  return OCT_OK;
}

extern "C" int octave_bn_bwd_reduce(const OctaveAct* dy, const OctaveAct* mask, const float* relu_ab, const OctaveAct* x,
                                    const float* mean_invstd, double* sums2, const OctaveAct* dmasked, void* stream) {
  if (!view_ok(dy) || !view_ok(x) || !same_shape(dy, x) || !mean_invstd || !sums2) return OCT_ERR_INVALID;
  if (mask && (!view_ok(mask) || !same_shape(mask, x))) return OCT_ERR_INVALID;
  if (dmasked && (!view_ok(dmasked) || !same_shape(dmasked, x))) return OCT_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  if (!g_octave_stats_prezeroed && cudaMemsetAsync(sums2, 0, sizeof(double) * 2 * x->C, s) != cudaSuccess) return OCT_ERR_LAUNCH;
  const OctaveAct* m = mask ? mask : x;
  const int mmode = mask ? 1 : (relu_ab ? 2 : 0);
  int rc = OCT_OK;
  DISPATCH_T(x->dtype, {
    if (mmode == 0) rc = launch_bn_bwd_reduce<T, 0>(dy, m, relu_ab, x, mean_invstd, sums2, dmasked, s);
    else if (mmode == 1) rc = launch_bn_bwd_reduce<T, 1>(dy, m, relu_ab, x, mean_invstd, sums2, dmasked, s);
    else rc = launch_bn_bwd_reduce<T, 2>(dy, m, relu_ab, x, mean_invstd, sums2, dmasked, s);
  });
  if (rc != OCT_OK) return rc;
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

template <typename T, int MASK>
static int launch_bn_bwd_apply(const OctaveAct* dy, const OctaveAct* m, const float* relu_ab, const OctaveAct* x,
                               const float* mi, const float* gamma, const double* sums2, int training,
                               const OctaveAct* dx, float* dgamma, float* dbeta, const OctaveAct* dmasked, cudaStream_t s) {
  auto fn = bn_bwd_apply_kernel<T, MASK>;
  Geo g;
  if (!make_geo(x, &g, U_BNA, false, ONE_WAVE(fn, 256, 0))) return OCT_ERR_UNSUPPORTED;
  fn<<<g.grid, g.bs, 0, s>>>(*dy, *m, g.Gb, relu_ab, *x, mi, gamma, sums2, training, *dx, dgamma, dbeta,
                             dmasked ? *dmasked : *dx, dmasked != nullptr);
  return OCT_OK;
}

extern "C" int octave_bn_bwd_apply(const OctaveAct* dy, const OctaveAct* mask, const float* relu_ab, const OctaveAct* x,
                                   const float* mean_invstd, const float* gamma, const double* sums2, int32_t training,
                                   const OctaveAct* dx, float* dgamma, float* dbeta, const OctaveAct* dmasked, void* stream) {
  if (!view_ok(dy) || !view_ok(x) || !view_ok(dx) || !same_shape(dy, x) || !same_shape(dx, x)) return OCT_ERR_INVALID;
  if (dmasked && (!view_ok(dmasked) || !same_shape(dmasked, x))) return OCT_ERR_INVALID;
  if (!mean_invstd || !sums2) return OCT_ERR_INVALID;
  if (mask && (!view_ok(mask) || !same_shape(mask, x))) return OCT_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  const OctaveAct* m = mask ? mask : x;
  const int mmode = mask ? 1 : (relu_ab ? 2 : 0);
  int rc = OCT_OK;
  DISPATCH_T(x->dtype, {
    if (mmode == 0) rc = launch_bn_bwd_apply<T, 0>(dy, m, relu_ab, x, mean_invstd, gamma, sums2, training, dx, dgamma, dbeta, dmasked, s);
    else if (mmode == 1) rc = launch_bn_bwd_apply<T, 1>(dy, m, relu_ab, x, mean_invstd, gamma, sums2, training, dx, dgamma, dbeta, dmasked, s);
    else rc = launch_bn_bwd_apply<T, 2>(dy, m, relu_ab, x, mean_invstd, gamma, sums2, training, dx, dgamma, dbeta, dmasked, s);
  });
  if (rc != OCT_OK) return rc;
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

template <typename T, int OP>
static int launch_binary(const OctaveAct* a, const OctaveAct* b, const OctaveAct* out, cudaStream_t s) {
  auto fn = binary_kernel<T, OP>;
  Geo g;
  if (!make_geo(a, &g, U_BIN, false, ONE_WAVE(fn, 256, 0))) return OCT_ERR_UNSUPPORTED;
  fn<<<g.grid, g.bs, 0, s>>>(*a, *b, g.Gb, *out);
  return OCT_OK;
}

// vectorised path of octave_act_bwd (conv_direct.cu) for views with 8-channel granularity
int oct_act_bwd_vec(const OctaveAct* y, const OctaveAct* dy, int act, const OctaveAct* dz, cudaStream_t s) {
  if (!view_ok(y) || !view_ok(dy) || !view_ok(dz) || !same_shape(y, dy) || !same_shape(y, dz)) return OCT_ERR_UNSUPPORTED;
  int rc = OCT_OK;
  DISPATCH_T(y->dtype, {
    switch (act) {
      case 1: rc = launch_binary<T, 2>(dy, y, dz, s); break;
      case 2: rc = launch_binary<T, 3>(dy, y, dz, s); break;
      case 3: rc = launch_binary<T, 4>(dy, y, dz, s); break;
      case 4: rc = launch_binary<T, 5>(dy, y, dz, s); break;
      default: rc = OCT_ERR_UNSUPPORTED;
    }
  });
  return rc;
}

extern "C" int octave_add_inplace(const OctaveAct* dst, const OctaveAct* src, void* stream) {
  if (!view_ok(dst) || !view_ok(src) || !same_shape(dst, src)) return OCT_ERR_INVALID;
  int rc = OCT_OK;
  DISPATCH_T(dst->dtype, rc = (launch_binary<T, 0>(dst, src, dst, (cudaStream_t)stream)));
  if (rc != OCT_OK) return rc;
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_relu_bwd(const OctaveAct* dy, const OctaveAct* mask, const OctaveAct* dx, void* stream) {
  if (!view_ok(dy) || !view_ok(mask) || !view_ok(dx) || !same_shape(dy, mask) || !same_shape(dy, dx)) return OCT_ERR_INVALID;
  int rc = OCT_OK;
  DISPATCH_T(dy->dtype, rc = (launch_binary<T, 1>(dy, mask, dx, (cudaStream_t)stream)));
  if (rc != OCT_OK) return rc;
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_splat_combine(const OctaveAct* U, const float* att, int32_t relu, const OctaveAct* out,
                                    void* stream) {
  if (!view_ok(U) || !view_ok(out) || !att) return OCT_ERR_INVALID;
  if (U->C != 2 * out->C || U->B != out->B || U->H != out->H || U->W != out->W || U->dtype != out->dtype) return OCT_ERR_INVALID;
  Geo g;
  DISPATCH_T(out->dtype, {
    auto fn = splat_combine_kernel<T>;
    if (!make_geo(out, &g, U_SPL, true, 2 * ONE_WAVE(fn, 256, 0))) return OCT_ERR_UNSUPPORTED;
    fn<<<g.grid, g.bs, 0, (cudaStream_t)stream>>>(*U, att, g.Gb, relu, *out);
  });
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_splat_bwd_reduce(const OctaveAct* dout, const OctaveAct* mask, const OctaveAct* U, float* datt,
                                       void* stream) {
  if (!view_ok(dout) || !view_ok(U) || !datt) return OCT_ERR_INVALID;
  if (U->C != 2 * dout->C || U->B != dout->B || U->H != dout->H || U->W != dout->W || U->dtype != dout->dtype) return OCT_ERR_INVALID;
  if (mask && (!view_ok(mask) || !same_shape(mask, dout))) return OCT_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(datt, 0, sizeof(float) * dout->B * 2 * dout->C, s) != cudaSuccess) return OCT_ERR_LAUNCH;
  OctaveAct m = mask ? *mask : *dout;
  Geo g;
  DISPATCH_T(dout->dtype, {
    if (mask) {
      auto fn = splat_bwd_reduce_kernel<T, true>;
      if (!make_geo(dout, &g, U_SPL, true, ONE_WAVE(fn, 256, 16 * 256 * sizeof(float)), 4)) return OCT_ERR_UNSUPPORTED;
      fn<<<g.grid, g.bs, 16 * g.bs * sizeof(float), s>>>(*dout, m, g.Gb, *U, datt);
    } else {
      auto fn = splat_bwd_reduce_kernel<T, false>;
      if (!make_geo(dout, &g, U_SPL, true, ONE_WAVE(fn, 256, 16 * 256 * sizeof(float)), 4)) return OCT_ERR_UNSUPPORTED;
      fn<<<g.grid, g.bs, 16 * g.bs * sizeof(float), s>>>(*dout, m, g.Gb, *U, datt);
    }
  });
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_splat_bwd_du(const OctaveAct* dout, const OctaveAct* mask, const float* att, const float* dgap,
                                   float gap_scale, const OctaveAct* dU, void* stream) {
  if (!view_ok(dout) || !view_ok(dU) || !att) return OCT_ERR_INVALID;
  if (dU->C != 2 * dout->C || dU->B != dout->B || dU->H != dout->H || dU->W != dout->W || dU->dtype != dout->dtype) return OCT_ERR_INVALID;
  if (mask && (!view_ok(mask) || !same_shape(mask, dout))) return OCT_ERR_INVALID;
  OctaveAct m = mask ? *mask : *dout;
  Geo g;
  DISPATCH_T(dout->dtype, {
    if (mask) {
      auto fn = splat_bwd_du_kernel<T, true>;
      if (!make_geo(dout, &g, U_SPL, true, 2 * ONE_WAVE(fn, 256, 0))) return OCT_ERR_UNSUPPORTED;
      fn<<<g.grid, g.bs, 0, (cudaStream_t)stream>>>(*dout, m, g.Gb, att, dgap, gap_scale, *dU);
    } else {
      auto fn = splat_bwd_du_kernel<T, false>;
      if (!make_geo(dout, &g, U_SPL, true, 2 * ONE_WAVE(fn, 256, 0))) return OCT_ERR_UNSUPPORTED;
      fn<<<g.grid, g.bs, 0, (cudaStream_t)stream>>>(*dout, m, g.Gb, att, dgap, gap_scale, *dU);
    }
  });
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_splat_bn_bwd(const OctaveAct* dout, const OctaveAct* omask, const float* att, const float* dgap,
                                   float gap_scale, const OctaveAct* z, const float* ab, const float* mean_invstd,
                                   const float* gamma, int32_t training, double* sums2, const OctaveAct* dz, float* dgamma,
                                   float* dbeta, void* stream) {
  if (!view_ok(dout) || !view_ok(z) || !view_ok(dz) || !att || !ab || !mean_invstd || !sums2) return OCT_ERR_INVALID;
  if (z->C != 2 * dout->C || z->B != dout->B || z->H != dout->H || z->W != dout->W || z->dtype != dout->dtype) return OCT_ERR_INVALID;
  if (!same_shape(z, dz)) return OCT_ERR_INVALID;
  if (omask && (!view_ok(omask) || !same_shape(omask, dout))) return OCT_ERR_INVALID;
  if (dout->C / 8 > 256) return OCT_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  if (!g_octave_stats_prezeroed && cudaMemsetAsync(sums2, 0, sizeof(double) * 2 * z->C, s) != cudaSuccess) return OCT_ERR_LAUNCH;
  OctaveAct m = omask ? *omask : *dout;
  Geo g;
  DISPATCH_T(dout->dtype, {
    if (omask) {
      auto fr = splat_bn_bwd_reduce_kernel<T, true>;
      auto fa = splat_bn_bwd_apply_kernel<T, true>;
      if (!make_geo(dout, &g, U_SBR, true, ONE_WAVE(fr, 256, 16 * 256 * sizeof(float)) / 2, 4)) return OCT_ERR_UNSUPPORTED;
      g.grid.z = 2;
      fr<<<g.grid, g.bs, 16 * g.bs * sizeof(float), s>>>(*dout, m, g.Gb, att, dgap, gap_scale, *z, ab, mean_invstd, sums2);
      if (!make_geo(dout, &g, U_SBA, true, ONE_WAVE(fa, 256, 0) / 2)) return OCT_ERR_UNSUPPORTED;
      g.grid.z = 2;
      fa<<<g.grid, g.bs, 0, s>>>(*dout, m, g.Gb, att, dgap, gap_scale, *z, ab, mean_invstd, gamma, sums2, training, *dz, dgamma, dbeta);
    } else {
      auto fr = splat_bn_bwd_reduce_kernel<T, false>;
      auto fa = splat_bn_bwd_apply_kernel<T, false>;
      if (!make_geo(dout, &g, U_SBR, true, ONE_WAVE(fr, 256, 16 * 256 * sizeof(float)) / 2, 4)) return OCT_ERR_UNSUPPORTED;
      g.grid.z = 2;
      fr<<<g.grid, g.bs, 16 * g.bs * sizeof(float), s>>>(*dout, m, g.Gb, att, dgap, gap_scale, *z, ab, mean_invstd, sums2);
      if (!make_geo(dout, &g, U_SBA, true, ONE_WAVE(fa, 256, 0) / 2)) return OCT_ERR_UNSUPPORTED;
      g.grid.z = 2;
      fa<<<g.grid, g.bs, 0, s>>>(*dout, m, g.Gb, att, dgap, gap_scale, *z, ab, mean_invstd, gamma, sums2, training, *dz, dgamma, dbeta);
    }
  });
  ++g_octave_launches;    // two kernels: the macro below counts one
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}
