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

struct Geo {
  int G, bs, ppb;  // channel groups, block size, pixels per block pass
  dim3 grid;
};

static int reduce_blocks() {
  static int v = 0;
  if (!v) {
    const char* e = getenv("OCTAVE_REDUCE_BLOCKS");
    v = e ? atoi(e) : 148 * 4;
    if (v < 1) v = 148 * 4;
  }
  return v;
}
#define kReduceBlocks reduce_blocks()   // total blocks of a reduction kernel: bounds the atomics per channel
constexpr int kStreamBlocks = 148 * 16;

bool make_geo(const OctaveAct* a, Geo* g, int total_blocks = kStreamBlocks) {
  if (a->C % 8) return false;
  g->G = a->C / 8;
  if (g->G > 1024) return false;
  if (g->G <= 256) g->bs = (256 / g->G) * g->G; else g->bs = g->G;
  g->ppb = g->bs / g->G;
  const long long hw = (long long)a->H * a->W;
  long long bx = (hw + g->ppb - 1) / g->ppb;
  long long cap = (long long)(total_blocks + a->B - 1) / a->B;
  if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  g->grid = dim3((unsigned)bx, (unsigned)a->B);
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

// Fold NV per-thread values over the pixel lanes of a block (threads with equal tid % G).
// Result valid for threads tid < G.  sm must hold NV * blockDim.x floats.
template <int NV>
__device__ __forceinline__ void fold_lanes(float (&v)[NV], float* sm, int G) {
  const int tid = threadIdx.x, bs = blockDim.x;
#pragma unroll
  for (int i = 0; i < NV; ++i) sm[i * bs + tid] = v[i];
  __syncthreads();
  if (tid < G) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float s = 0.f;
      for (int l = tid; l < bs; l += G) s += sm[i * bs + l];
      v[i] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void chan_stats_kernel(const OctaveAct x, double* sums) {
  extern __shared__ float sm[];
  const int G = x.C >> 3, cg = threadIdx.x % G, lane = threadIdx.x / G, ppb = blockDim.x / G;
  const long long hw = (long long)x.H * x.W, base = (long long)blockIdx.y * hw;
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.f;
  const long long step = (long long)gridDim.x * ppb;
  for (long long p = (long long)blockIdx.x * ppb + lane; p < hw; p += 4 * step) {
    float f[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (p + u * step < hw) VecIO<T, 8>::ld(at<T>(x, base + p + u * step, cg * 8), f[u]);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[u][i] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { v[i] += f[u][i]; v[8 + i] += f[u][i] * f[u][i]; }
    }
  }
  fold_lanes<16>(v, sm, G);
  if (threadIdx.x < G) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(sums + cg * 8 + i, (double)v[i]);
      atomicAdd(sums + x.C + cg * 8 + i, (double)v[8 + i]);
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

template <typename T>
__global__ void affine_act_kernel(const OctaveAct x, const float* ab, const OctaveAct res, int has_res, int relu,
                                  const OctaveAct y, float* gap) {
  extern __shared__ float sm[];
  const int G = x.C >> 3, cg = threadIdx.x % G, lane = threadIdx.x / G, ppb = blockDim.x / G;
  const long long hw = (long long)x.H * x.W, base = (long long)blockIdx.y * hw;
  float a[8], b[8], acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = ab ? ab[cg * 8 + i] : 1.f;
    b[i] = ab ? ab[x.C + cg * 8 + i] : 0.f;
    acc[i] = 0.f;
  }
  for (long long p = (long long)blockIdx.x * ppb + lane; p < hw; p += (long long)gridDim.x * ppb) {
    float f[8];
    VecIO<T, 8>::ld(at<T>(x, base + p, cg * 8), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = f[i] * a[i] + b[i];
    if (has_res) {
      float r[8];
      VecIO<T, 8>::ld(at<T>(res, base + p, cg * 8), r);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += r[i];
    }
    if (relu) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
    }
    if (gap) {
      // accumulate what the consumer will read back (storage-rounded values)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += to_f(from_f<T>(f[i]));
    }
    VecIO<T, 8>::st(at<T>(y, base + p, cg * 8), f);
  }
  if (gap) {
    fold_lanes<8>(acc, sm, G);
    if (threadIdx.x < G) {
      const int half = x.C >> 1;
      const int c = (cg * 8) % half;
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(gap + (long long)blockIdx.y * half + c + i, acc[i]);
    }
  }
}

// has_mask: 0 none, 1 mask tensor (dz = dy * (mask > 0)), 2 recompute the ReLU mask of this very BN from x: (x*a+b > 0)
template <typename T>
__global__ void bn_bwd_reduce_kernel(const OctaveAct dy, const OctaveAct mask, int has_mask, const float* ab, const OctaveAct x,
                                     const float* mi, double* sums2) {
  extern __shared__ float sm[];
  const int G = x.C >> 3, cg = threadIdx.x % G, lane = threadIdx.x / G, ppb = blockDim.x / G;
  const long long hw = (long long)x.H * x.W, base = (long long)blockIdx.y * hw;
  float mean[8], inv[8], v[16], aa[8], bb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mean[i] = mi[cg * 8 + i];
    inv[i] = mi[x.C + cg * 8 + i];
    aa[i] = has_mask == 2 ? ab[cg * 8 + i] : 0.f;
    bb[i] = has_mask == 2 ? ab[x.C + cg * 8 + i] : 0.f;
    v[i] = v[8 + i] = 0.f;
  }
  const long long step = (long long)gridDim.x * ppb;
  for (long long p = (long long)blockIdx.x * ppb + lane; p < hw; p += 2 * step) {
    float d[2][8], f[2][8], m[2][8];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const bool ok = p + u * step < hw;
      const long long q = ok ? p + u * step : p;
      VecIO<T, 8>::ld(at<T>(dy, base + q, cg * 8), d[u]);
      VecIO<T, 8>::ld(at<T>(x, base + q, cg * 8), f[u]);
      if (has_mask == 1) VecIO<T, 8>::ld(at<T>(mask, base + q, cg * 8), m[u]);
      if (has_mask == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) m[u][i] = to_f(from_f<T>(f[u][i] * aa[i] + bb[i]));
      }
      if (!ok) {
#pragma unroll
        for (int i = 0; i < 8; ++i) d[u][i] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float dd = (has_mask && !(m[u][i] > 0.f)) ? 0.f : d[u][i];
        v[i] += dd;
        v[8 + i] += dd * (f[u][i] - mean[i]) * inv[i];
      }
    }
  }
  fold_lanes<16>(v, sm, G);
  if (threadIdx.x < G) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(sums2 + cg * 8 + i, (double)v[i]);
      atomicAdd(sums2 + x.C + cg * 8 + i, (double)v[8 + i]);
    }
  }
}

template <typename T>
__global__ void bn_bwd_apply_kernel(const OctaveAct dy, const OctaveAct mask, int has_mask, const float* ab, const OctaveAct x,
                                    const float* mi, const float* gamma, const double* sums2, int training,
                                    const OctaveAct dx, float* dgamma, float* dbeta) {
  const int G = x.C >> 3, cg = threadIdx.x % G, lane = threadIdx.x / G, ppb = blockDim.x / G;
  const long long hw = (long long)x.H * x.W, base = (long long)blockIdx.y * hw;
  const float inv_n = 1.f / ((float)x.B * (float)hw);
  float mean[8], inv[8], k1[8], k2[8], ag[8], aa[8], bb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = cg * 8 + i;
    mean[i] = mi[c];
    inv[i] = mi[x.C + c];
    aa[i] = has_mask == 2 ? ab[c] : 0.f;
    bb[i] = has_mask == 2 ? ab[x.C + c] : 0.f;
    const float sd = (float)sums2[c], sdx = (float)sums2[x.C + c];
    k1[i] = training ? sd * inv_n : 0.f;
    k2[i] = training ? sdx * inv_n : 0.f;
    ag[i] = (gamma ? gamma[c] : 1.f) * inv[i];
    if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) {
      if (dgamma) dgamma[c] = sdx;
      if (dbeta) dbeta[c] = sd;
    }
  }
  for (long long p = (long long)blockIdx.x * ppb + lane; p < hw; p += (long long)gridDim.x * ppb) {
    float d[8], f[8];
    VecIO<T, 8>::ld(at<T>(dy, base + p, cg * 8), d);
    VecIO<T, 8>::ld(at<T>(x, base + p, cg * 8), f);
    if (has_mask == 1) {
      float m[8];
      VecIO<T, 8>::ld(at<T>(mask, base + p, cg * 8), m);
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = m[i] > 0.f ? d[i] : 0.f;
    } else if (has_mask == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = to_f(from_f<T>(f[i] * aa[i] + bb[i])) > 0.f ? d[i] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = ag[i] * (d[i] - k1[i] - (f[i] - mean[i]) * inv[i] * k2[i]);
    VecIO<T, 8>::st(at<T>(dx, base + p, cg * 8), d);
  }
}

template <typename T>
__global__ void add_inplace_kernel(const OctaveAct dst, const OctaveAct src) {
  const int G = dst.C >> 3, cg = threadIdx.x % G, lane = threadIdx.x / G, ppb = blockDim.x / G;
  const long long hw = (long long)dst.H * dst.W, base = (long long)blockIdx.y * hw;
  for (long long p = (long long)blockIdx.x * ppb + lane; p < hw; p += (long long)gridDim.x * ppb) {
    float a[8], b[8];
    VecIO<T, 8>::ld(at<T>(dst, base + p, cg * 8), a);
    VecIO<T, 8>::ld(at<T>(src, base + p, cg * 8), b);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] += b[i];
    VecIO<T, 8>::st(at<T>(dst, base + p, cg * 8), a);
  }
}

template <typename T>
__global__ void relu_bwd_kernel(const OctaveAct dy, const OctaveAct mask, const OctaveAct dx) {
  const int G = dy.C >> 3, cg = threadIdx.x % G, lane = threadIdx.x / G, ppb = blockDim.x / G;
  const long long hw = (long long)dy.H * dy.W, base = (long long)blockIdx.y * hw;
  for (long long p = (long long)blockIdx.x * ppb + lane; p < hw; p += (long long)gridDim.x * ppb) {
    float a[8], m[8];
    VecIO<T, 8>::ld(at<T>(dy, base + p, cg * 8), a);
    VecIO<T, 8>::ld(at<T>(mask, base + p, cg * 8), m);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = m[i] > 0.f ? a[i] : 0.f;
    VecIO<T, 8>::st(at<T>(dx, base + p, cg * 8), a);
  }
}

// ---- split attention ------------------------------------------------------------------------------
template <typename T>
__global__ void splat_combine_kernel(const OctaveAct U, const float* att, int relu, const OctaveAct out) {
  const int C = out.C, G = C >> 3, cg = threadIdx.x % G, lane = threadIdx.x / G, ppb = blockDim.x / G;
  const long long hw = (long long)out.H * out.W, base = (long long)blockIdx.y * hw;
  float a0[8], a1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a0[i] = att[(long long)blockIdx.y * 2 * C + cg * 8 + i];
    a1[i] = att[(long long)blockIdx.y * 2 * C + C + cg * 8 + i];
  }
  for (long long p = (long long)blockIdx.x * ppb + lane; p < hw; p += (long long)gridDim.x * ppb) {
    float u0[8], u1[8];
    VecIO<T, 8>::ld(at<T>(U, base + p, cg * 8), u0);
    VecIO<T, 8>::ld(at<T>(U, base + p, C + cg * 8), u1);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      u0[i] = a0[i] * u0[i] + a1[i] * u1[i];
      if (relu) u0[i] = fmaxf(u0[i], 0.f);
    }
    VecIO<T, 8>::st(at<T>(out, base + p, cg * 8), u0);
  }
}

template <typename T>
__global__ void splat_bwd_reduce_kernel(const OctaveAct dout, const OctaveAct mask, int has_mask, const OctaveAct U,
                                        float* datt) {
  extern __shared__ float sm[];
  const int C = dout.C, G = C >> 3, cg = threadIdx.x % G, lane = threadIdx.x / G, ppb = blockDim.x / G;
  const long long hw = (long long)dout.H * dout.W, base = (long long)blockIdx.y * hw;
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.f;
  for (long long p = (long long)blockIdx.x * ppb + lane; p < hw; p += (long long)gridDim.x * ppb) {
    float d[8], u0[8], u1[8];
    VecIO<T, 8>::ld(at<T>(dout, base + p, cg * 8), d);
    VecIO<T, 8>::ld(at<T>(U, base + p, cg * 8), u0);
    VecIO<T, 8>::ld(at<T>(U, base + p, C + cg * 8), u1);
    if (has_mask) {
      float m[8];
      VecIO<T, 8>::ld(at<T>(mask, base + p, cg * 8), m);
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = m[i] > 0.f ? d[i] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] += d[i] * u0[i]; v[8 + i] += d[i] * u1[i]; }
  }
  fold_lanes<16>(v, sm, G);
  if (threadIdx.x < G) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(datt + (long long)blockIdx.y * 2 * C + cg * 8 + i, v[i]);
      atomicAdd(datt + (long long)blockIdx.y * 2 * C + C + cg * 8 + i, v[8 + i]);
    }
  }
}

template <typename T>
__global__ void splat_bwd_du_kernel(const OctaveAct dout, const OctaveAct mask, int has_mask, const float* att,
                                    const float* dgap, float gap_scale, const OctaveAct dU) {
  const int C = dout.C, G = C >> 3, cg = threadIdx.x % G, lane = threadIdx.x / G, ppb = blockDim.x / G;
  const long long hw = (long long)dout.H * dout.W, base = (long long)blockIdx.y * hw;
  float a0[8], a1[8], gg[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a0[i] = att[(long long)blockIdx.y * 2 * C + cg * 8 + i];
    a1[i] = att[(long long)blockIdx.y * 2 * C + C + cg * 8 + i];
    gg[i] = dgap ? dgap[(long long)blockIdx.y * C + cg * 8 + i] * gap_scale : 0.f;
  }
  for (long long p = (long long)blockIdx.x * ppb + lane; p < hw; p += (long long)gridDim.x * ppb) {
    float d[8], o0[8], o1[8];
    VecIO<T, 8>::ld(at<T>(dout, base + p, cg * 8), d);
    if (has_mask) {
      float m[8];
      VecIO<T, 8>::ld(at<T>(mask, base + p, cg * 8), m);
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = m[i] > 0.f ? d[i] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { o0[i] = a0[i] * d[i] + gg[i]; o1[i] = a1[i] * d[i] + gg[i]; }
    VecIO<T, 8>::st(at<T>(dU, base + p, cg * 8), o0);
    VecIO<T, 8>::st(at<T>(dU, base + p, C + cg * 8), o1);
  }
}

#define DISPATCH_T(dtype, ...)                         \
  do {                                                 \
    if ((dtype) == OCT_DTYPE_F32) { using T = float; __VA_ARGS__; } \
    else { using T = bf16; __VA_ARGS__; }              \
  } while (0)

}  // namespace

extern "C" int octave_chan_stats(const OctaveAct* x, double* sums, void* stream) {
  if (!view_ok(x) || !sums) return OCT_ERR_INVALID;
  Geo g;
  if (!make_geo(x, &g, kReduceBlocks)) return OCT_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(sums, 0, sizeof(double) * 2 * x->C, s) != cudaSuccess) return OCT_ERR_LAUNCH;
  DISPATCH_T(x->dtype, (chan_stats_kernel<T><<<g.grid, g.bs, 16 * g.bs * sizeof(float), s>>>(*x, sums)));
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

extern "C" int octave_affine_act(const OctaveAct* x, const float* ab, const OctaveAct* res, int32_t relu,
                                 const OctaveAct* y, float* gap, void* stream) {
  if (!view_ok(x) || !view_ok(y) || !same_shape(x, y)) return OCT_ERR_INVALID;
  if (res && (!view_ok(res) || !same_shape(x, res))) return OCT_ERR_INVALID;
  if (gap && (x->C % 16)) return OCT_ERR_INVALID;
  Geo g;
  if (!make_geo(x, &g, gap ? kReduceBlocks * 2 : kStreamBlocks)) return OCT_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  if (gap && cudaMemsetAsync(gap, 0, sizeof(float) * x->B * (x->C / 2), s) != cudaSuccess) return OCT_ERR_LAUNCH;
  OctaveAct r = res ? *res : *x;
  DISPATCH_T(x->dtype, (affine_act_kernel<T><<<g.grid, g.bs, gap ? 8 * g.bs * sizeof(float) : 0, s>>>(
                           *x, ab, r, res != nullptr, relu, *y, gap)));
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_bn_bwd_reduce(const OctaveAct* dy, const OctaveAct* mask, const float* relu_ab, const OctaveAct* x,
                                    const float* mean_invstd, double* sums2, void* stream) {
  if (!view_ok(dy) || !view_ok(x) || !same_shape(dy, x) || !mean_invstd || !sums2) return OCT_ERR_INVALID;
  if (mask && (!view_ok(mask) || !same_shape(mask, x))) return OCT_ERR_INVALID;
  Geo g;
  if (!make_geo(x, &g, kReduceBlocks)) return OCT_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(sums2, 0, sizeof(double) * 2 * x->C, s) != cudaSuccess) return OCT_ERR_LAUNCH;
  OctaveAct m = mask ? *mask : *x;
  const int mmode = mask ? 1 : (relu_ab ? 2 : 0);
  DISPATCH_T(x->dtype, (bn_bwd_reduce_kernel<T><<<g.grid, g.bs, 16 * g.bs * sizeof(float), s>>>(
                           *dy, m, mmode, relu_ab, *x, mean_invstd, sums2)));
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_bn_bwd_apply(const OctaveAct* dy, const OctaveAct* mask, const float* relu_ab, const OctaveAct* x,
                                   const float* mean_invstd, const float* gamma, const double* sums2, int32_t training,
                                   const OctaveAct* dx, float* dgamma, float* dbeta, void* stream) {
  if (!view_ok(dy) || !view_ok(x) || !view_ok(dx) || !same_shape(dy, x) || !same_shape(dx, x)) return OCT_ERR_INVALID;
  if (!mean_invstd || !sums2) return OCT_ERR_INVALID;
  if (mask && (!view_ok(mask) || !same_shape(mask, x))) return OCT_ERR_INVALID;
  Geo g;
  if (!make_geo(x, &g)) return OCT_ERR_UNSUPPORTED;
  OctaveAct m = mask ? *mask : *x;
  const int mmode = mask ? 1 : (relu_ab ? 2 : 0);
  DISPATCH_T(x->dtype, (bn_bwd_apply_kernel<T><<<g.grid, g.bs, 0, (cudaStream_t)stream>>>(
                           *dy, m, mmode, relu_ab, *x, mean_invstd, gamma, sums2, training, *dx, dgamma, dbeta)));
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_add_inplace(const OctaveAct* dst, const OctaveAct* src, void* stream) {
  if (!view_ok(dst) || !view_ok(src) || !same_shape(dst, src)) return OCT_ERR_INVALID;
  Geo g;
  if (!make_geo(dst, &g)) return OCT_ERR_UNSUPPORTED;
  DISPATCH_T(dst->dtype, (add_inplace_kernel<T><<<g.grid, g.bs, 0, (cudaStream_t)stream>>>(*dst, *src)));
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_relu_bwd(const OctaveAct* dy, const OctaveAct* mask, const OctaveAct* dx, void* stream) {
  if (!view_ok(dy) || !view_ok(mask) || !view_ok(dx) || !same_shape(dy, mask) || !same_shape(dy, dx)) return OCT_ERR_INVALID;
  Geo g;
  if (!make_geo(dy, &g)) return OCT_ERR_UNSUPPORTED;
  DISPATCH_T(dy->dtype, (relu_bwd_kernel<T><<<g.grid, g.bs, 0, (cudaStream_t)stream>>>(*dy, *mask, *dx)));
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_splat_combine(const OctaveAct* U, const float* att, int32_t relu, const OctaveAct* out,
                                    void* stream) {
  if (!view_ok(U) || !view_ok(out) || !att) return OCT_ERR_INVALID;
  if (U->C != 2 * out->C || U->B != out->B || U->H != out->H || U->W != out->W || U->dtype != out->dtype) return OCT_ERR_INVALID;
  Geo g;
  if (!make_geo(out, &g)) return OCT_ERR_UNSUPPORTED;
  DISPATCH_T(out->dtype, (splat_combine_kernel<T><<<g.grid, g.bs, 0, (cudaStream_t)stream>>>(*U, att, relu, *out)));
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_splat_bwd_reduce(const OctaveAct* dout, const OctaveAct* mask, const OctaveAct* U, float* datt,
                                       void* stream) {
  if (!view_ok(dout) || !view_ok(U) || !datt) return OCT_ERR_INVALID;
  if (U->C != 2 * dout->C || U->B != dout->B || U->H != dout->H || U->W != dout->W || U->dtype != dout->dtype) return OCT_ERR_INVALID;
  if (mask && (!view_ok(mask) || !same_shape(mask, dout))) return OCT_ERR_INVALID;
  Geo g;
  if (!make_geo(dout, &g, kReduceBlocks)) return OCT_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(datt, 0, sizeof(float) * dout->B * 2 * dout->C, s) != cudaSuccess) return OCT_ERR_LAUNCH;
  OctaveAct m = mask ? *mask : *dout;
  DISPATCH_T(dout->dtype, (splat_bwd_reduce_kernel<T><<<g.grid, g.bs, 16 * g.bs * sizeof(float), s>>>(
                              *dout, m, mask != nullptr, *U, datt)));
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_splat_bwd_du(const OctaveAct* dout, const OctaveAct* mask, const float* att, const float* dgap,
                                   float gap_scale, const OctaveAct* dU, void* stream) {
  if (!view_ok(dout) || !view_ok(dU) || !att) return OCT_ERR_INVALID;
  if (dU->C != 2 * dout->C || dU->B != dout->B || dU->H != dout->H || dU->W != dout->W || dU->dtype != dout->dtype) return OCT_ERR_INVALID;
  if (mask && (!view_ok(mask) || !same_shape(mask, dout))) return OCT_ERR_INVALID;
  Geo g;
  if (!make_geo(dout, &g)) return OCT_ERR_UNSUPPORTED;
  OctaveAct m = mask ? *mask : *dout;
  DISPATCH_T(dout->dtype, (splat_bwd_du_kernel<T><<<g.grid, g.bs, 0, (cudaStream_t)stream>>>(
                              *dout, m, mask != nullptr, att, dgap, gap_scale, *dU)));
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}
