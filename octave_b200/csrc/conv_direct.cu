// CUDA-core direct convolution: the general path (fp32 mode, 3-channel stem conv, discriminator convs).
// Tiled 32 pixels x 32 channels per block, K loop over (tap, 32-channel chunk) staged in shared memory.
// Reference arithmetic: nn.Conv2d as used in /root/reference/architectures/extra/resnest.py and
// /root/reference/architectures/discriminator/blocks.py:46-50,91-109 (fp32 accumulate).
#include "common.cuh"
#include "../../include/octave_b200.h"

int oct_act_bwd_vec(const OctaveAct* y, const OctaveAct* dy, int act, const OctaveAct* dz, cudaStream_t s);

namespace {

constexpr int TM = 32, TN = 32, TK = 32;

template <typename T> __device__ __forceinline__ float ldf(const T* p) { return to_f(*p); }

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case 1: return fmaxf(v, 0.f);
    case 2: return v > 0.f ? v : 0.2f * v;
    case 3: return 1.f / (1.f + __expf(-v));
    case 4: return tanhf(v);
    default: return v;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) direct_fwd_kernel(const OctaveConvDesc d, const T* x, const float* w, const float* bias, T* y) {
  __shared__ float xs[TM][TK + 1];
  __shared__ float ws[TN][TK + 1];
  const int cin_g = d.cin / d.groups, cout_g = d.cout / d.groups, kk = d.ksize * d.ksize;
  const int g = blockIdx.z;
  const long long npix = (long long)d.B * d.Hout * d.Wout;
  const long long p0 = (long long)blockIdx.x * TM;
  const int co0 = blockIdx.y * TN;
  const int t = threadIdx.x;
  const int px = t >> 3, c4 = (t & 7) * 4;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  // loader coordinates: this thread loads xs[lp][lc..lc+3] and ws[lp][lc..lc+3]
  const int lp = t >> 3, lc = (t & 7) * 4;
  const long long lpix = p0 + lp;
  int ln = 0, loh = 0, low = 0;
  const bool lvalid = lpix < npix;
  if (lvalid) {
    low = (int)(lpix % d.Wout);
    loh = (int)((lpix / d.Wout) % d.Hout);
    ln = (int)(lpix / ((long long)d.Wout * d.Hout));
  }
  for (int tap = 0; tap < kk; ++tap) {
    const int kh = tap / d.ksize, kw = tap - kh * d.ksize;
    const int ih = loh * d.stride - d.pad + kh, iw = low * d.stride - d.pad + kw;
    const bool inb = lvalid && ih >= 0 && ih < d.H && iw >= 0 && iw < d.W;
    const T* xrow = x + (((long long)ln * d.H + ih) * d.W + iw) * d.x_ld + d.x_coff + g * cin_g;
    for (int ci0 = 0; ci0 < cin_g; ci0 += TK) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ci = ci0 + lc + j;
        xs[lp][lc + j] = (inb && ci < cin_g) ? ldf(xrow + ci) : 0.f;
        const int co = co0 + lp;
        ws[lp][lc + j] = (co < cout_g && ci < cin_g) ? w[((long long)(g * cout_g + co) * cin_g + ci) * kk + tap] : 0.f;
      }
      __syncthreads();
      float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < TK; ++k) {
        const float a = xs[px][k];
#pragma unroll
        for (int j = 0; j < 4; ++j) part[j] += a * ws[c4 + j][k];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] += part[j];
      __syncthreads();
    }
  }
  const long long pix = p0 + px;
  if (pix < npix) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + c4 + j;
      if (co < cout_g) {
        float v = acc[j] + (bias ? bias[g * cout_g + co] : 0.f);
        T* o = y + pix * d.y_ld + d.y_coff + g * cout_g + co;
        if (d.accumulate) v += to_f(*o);
        *o = from_f<T>(apply_act(v, d.relu));
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) direct_dgrad_kernel(const OctaveConvDesc d, const T* dy, const float* w, T* dx) {
  __shared__ float ds[TM][TK + 1];
  __shared__ float ws[TN][TK + 1];  // [ci][co]
  const int cin_g = d.cin / d.groups, cout_g = d.cout / d.groups, kk = d.ksize * d.ksize;
  const int g = blockIdx.z;
  const long long npix = (long long)d.B * d.H * d.W;
  const long long p0 = (long long)blockIdx.x * TM;
  const int ci0 = blockIdx.y * TN;
  const int t = threadIdx.x;
  const int px = t >> 3, c4 = (t & 7) * 4;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int lp = t >> 3, lc = (t & 7) * 4;
  const long long lpix = p0 + lp;
  int ln = 0, lih = 0, liw = 0;
  const bool lvalid = lpix < npix;
  if (lvalid) {
    liw = (int)(lpix % d.W);
    lih = (int)((lpix / d.W) % d.H);
    ln = (int)(lpix / ((long long)d.W * d.H));
  }
  for (int tap = 0; tap < kk; ++tap) {
    const int kh = tap / d.ksize, kw = tap - kh * d.ksize;
    const int nh = lih + d.pad - kh, nw = liw + d.pad - kw;
    const bool div = nh >= 0 && nw >= 0 && (nh % d.stride) == 0 && (nw % d.stride) == 0;
    const int oh = nh / d.stride, ow = nw / d.stride;
    const bool inb = lvalid && div && oh < d.Hout && ow < d.Wout;
    const T* drow = dy + (((long long)ln * d.Hout + oh) * d.Wout + ow) * d.y_ld + d.y_coff + g * cout_g;
    for (int cb = 0; cb < cout_g; cb += TK) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int co = cb + lc + j;
        ds[lp][lc + j] = (inb && co < cout_g) ? ldf(drow + co) : 0.f;
        const int ci = ci0 + lp;
        ws[lp][lc + j] = (ci < cin_g && co < cout_g) ? w[((long long)(g * cout_g + co) * cin_g + ci) * kk + tap] : 0.f;
      }
      __syncthreads();
      float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < TK; ++k) {
        const float a = ds[px][k];
#pragma unroll
        for (int j = 0; j < 4; ++j) part[j] += a * ws[c4 + j][k];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] += part[j];
      __syncthreads();
    }
  }
  const long long pix = p0 + px;
  if (pix < npix) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + c4 + j;
      if (ci < cin_g) {
        T* o = dx + pix * d.x_ld + d.x_coff + g * cin_g + ci;
        float v = acc[j];
        if (d.accumulate) v += to_f(*o);
        *o = from_f<T>(v);
      }
    }
  }
}

// block = (pixel slice, (co tile, ci tile), tap*groups); thread -> 1 co x 4 ci
template <typename T>
__global__ void __launch_bounds__(256) direct_wgrad_kernel(const OctaveConvDesc d, const T* x, const T* dy, float* dw, float* dbias,
                                                           int chunks_per_block) {
  __shared__ float ds[TK][TN + 1];  // [pixel][co]
  __shared__ float xs[TK][TN + 1];  // [pixel][ci]
  const int cin_g = d.cin / d.groups, cout_g = d.cout / d.groups, kk = d.ksize * d.ksize;
  const int tap = blockIdx.z % kk, g = blockIdx.z / kk;
  const int n_ci_tiles = (cin_g + TN - 1) / TN;
  const int co0 = (blockIdx.y / n_ci_tiles) * TN, ci0 = (blockIdx.y % n_ci_tiles) * TN;
  const int kh = tap / d.ksize, kw = tap - kh * d.ksize;
  const long long npix = (long long)d.B * d.Hout * d.Wout;
  const int t = threadIdx.x;
  const int co_l = t >> 3, c4 = (t & 7) * 4;
  const int lp = t >> 3, lc = (t & 7) * 4;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float bacc = 0.f;
  const long long chunk0 = (long long)blockIdx.x * chunks_per_block;
  for (int ch = 0; ch < chunks_per_block; ++ch) {
    const long long p0 = (chunk0 + ch) * TK;
    if (p0 >= npix) break;
    const long long lpix = p0 + lp;
    const bool lvalid = lpix < npix;
    int ln = 0, loh = 0, low = 0;
    if (lvalid) {
      low = (int)(lpix % d.Wout);
      loh = (int)((lpix / d.Wout) % d.Hout);
      ln = (int)(lpix / ((long long)d.Wout * d.Hout));
    }
    const int ih = loh * d.stride - d.pad + kh, iw = low * d.stride - d.pad + kw;
    const bool inb = lvalid && ih >= 0 && ih < d.H && iw >= 0 && iw < d.W;
    const T* xrow = x + (((long long)ln * d.H + ih) * d.W + iw) * d.x_ld + d.x_coff + g * cin_g;
    const T* drow = dy + lpix * d.y_ld + d.y_coff + g * cout_g;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + lc + j, ci = ci0 + lc + j;
      ds[lp][lc + j] = (lvalid && co < cout_g) ? ldf(drow + co) : 0.f;
      xs[lp][lc + j] = (inb && ci < cin_g) ? ldf(xrow + ci) : 0.f;
    }
    __syncthreads();
    float part[4] = {0.f, 0.f, 0.f, 0.f};
    float bpart = 0.f;
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float a = ds[k][co_l];
#pragma unroll
      for (int j = 0; j < 4; ++j) part[j] += a * xs[k][c4 + j];
      bpart += a;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] += part[j];
    if (c4 == 0) bacc += bpart;
    __syncthreads();
  }
  const int co = co0 + co_l;
  if (co < cout_g) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + c4 + j;
      if (ci < cin_g) atomicAdd(dw + ((long long)(g * cout_g + co) * cin_g + ci) * kk + tap, acc[j]);
    }
    if (dbias && c4 == 0 && tap == 0 && ci0 == 0) atomicAdd(dbias + g * cout_g + co, bacc);
  }
}

template <typename T>
__global__ void act_bwd_kernel(const OctaveAct y, const OctaveAct dy, int act, const OctaveAct dz) {
  const long long total = (long long)y.B * y.H * y.W * y.C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % y.C);
    const long long pix = idx / y.C;
    const float yv = to_f(reinterpret_cast<const T*>(y.data)[pix * y.ld + y.coff + c]);
    const float d = to_f(reinterpret_cast<const T*>(dy.data)[pix * dy.ld + dy.coff + c]);
    float r;
    switch (act) {
      case 1: r = yv > 0.f ? d : 0.f; break;
      case 2: r = yv > 0.f ? d : 0.2f * d; break;
      case 3: r = d * yv * (1.f - yv); break;
      case 4: r = d * (1.f - yv * yv); break;
      default: r = d;
    }
    reinterpret_cast<T*>(dz.data)[pix * dz.ld + dz.coff + c] = from_f<T>(r);
  }
}

int check(const OctaveConvDesc* d) {
  if (!d) return OCT_ERR_INVALID;
  if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->Hout <= 0 || d->Wout <= 0) return OCT_ERR_INVALID;
  if (d->cin <= 0 || d->cout <= 0 || d->groups <= 0 || d->cin % d->groups || d->cout % d->groups) return OCT_ERR_INVALID;
  if (d->ksize <= 0 || d->stride <= 0 || d->pad < 0) return OCT_ERR_INVALID;
  if (d->mode != OCT_CONV_MODE_CONV) return OCT_ERR_UNSUPPORTED;
  if (d->in_dtype != d->out_dtype) return OCT_ERR_UNSUPPORTED;
  if (d->in_dtype != OCT_DTYPE_F32 && d->in_dtype != OCT_DTYPE_BF16) return OCT_ERR_INVALID;
  return OCT_OK;
}

}  // namespace

extern "C" int octave_conv_direct_fwd(const OctaveConvDesc* d, const void* x, const float* w, const float* bias, void* y, void* stream) {
  int rc = check(d);
  if (rc != OCT_OK) return rc;
  if (!x || !w || !y) return OCT_ERR_INVALID;
  const long long npix = (long long)d->B * d->Hout * d->Wout;
  const int cout_g = d->cout / d->groups;
  dim3 grid((unsigned)((npix + TM - 1) / TM), (unsigned)((cout_g + TN - 1) / TN), (unsigned)d->groups);
  if (d->in_dtype == OCT_DTYPE_F32)
    direct_fwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(*d, (const float*)x, w, bias, (float*)y);
  else
    direct_fwd_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>(*d, (const bf16*)x, w, bias, (bf16*)y);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_conv_direct_dgrad(const OctaveConvDesc* d, const void* dy, const float* w, void* dx, void* stream) {
  int rc = check(d);
  if (rc != OCT_OK) return rc;
  if (!dy || !w || !dx) return OCT_ERR_INVALID;
  const long long npix = (long long)d->B * d->H * d->W;
  const int cin_g = d->cin / d->groups;
  dim3 grid((unsigned)((npix + TM - 1) / TM), (unsigned)((cin_g + TN - 1) / TN), (unsigned)d->groups);
  if (d->in_dtype == OCT_DTYPE_F32)
    direct_dgrad_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(*d, (const float*)dy, w, (float*)dx);
  else
    direct_dgrad_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>(*d, (const bf16*)dy, w, (bf16*)dx);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_conv_direct_wgrad(const OctaveConvDesc* d, const void* x, const void* dy, float* dw, float* dbias, void* stream) {
  int rc = check(d);
  if (rc != OCT_OK) return rc;
  if (!x || !dy || !dw) return OCT_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  const int cin_g = d->cin / d->groups, cout_g = d->cout / d->groups, kk = d->ksize * d->ksize;
  if (!d->accumulate) {
    if (cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d->cout * cin_g * kk, s) != cudaSuccess) return OCT_ERR_LAUNCH;
    if (dbias && cudaMemsetAsync(dbias, 0, sizeof(float) * d->cout, s) != cudaSuccess) return OCT_ERR_LAUNCH;
  }
  const long long npix = (long long)d->B * d->Hout * d->Wout;
  const long long chunks = (npix + TK - 1) / TK;
  const int out_tiles = ((cout_g + TN - 1) / TN) * ((cin_g + TN - 1) / TN) * kk * d->groups;
  long long split = (148LL * 8 + out_tiles - 1) / out_tiles;
  if (split > chunks) split = chunks;
  if (split < 1) split = 1;
  const int cpb = (int)((chunks + split - 1) / split);
  split = (chunks + cpb - 1) / cpb;
  dim3 grid((unsigned)split, (unsigned)(((cout_g + TN - 1) / TN) * ((cin_g + TN - 1) / TN)), (unsigned)(kk * d->groups));
  if (d->in_dtype == OCT_DTYPE_F32)
    direct_wgrad_kernel<float><<<grid, 256, 0, s>>>(*d, (const float*)x, (const float*)dy, dw, dbias, cpb);
  else
    direct_wgrad_kernel<bf16><<<grid, 256, 0, s>>>(*d, (const bf16*)x, (const bf16*)dy, dw, dbias, cpb);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_act_bwd(const OctaveAct* y, const OctaveAct* dy, int32_t act, const OctaveAct* dz, void* stream) {
  if (!y || !dy || !dz || !y->data || !dy->data || !dz->data) return OCT_ERR_INVALID;
  if (y->dtype != dy->dtype || y->dtype != dz->dtype || y->C != dy->C || y->C != dz->C) return OCT_ERR_INVALID;
  const long long total = (long long)y->B * y->H * y->W * y->C;
  if (act >= 1 && act <= 4) {
    const int rc = oct_act_bwd_vec(y, dy, act, dz, (cudaStream_t)stream);
    if (rc == OCT_OK) {
      OCT_CHECK_LAUNCH();
      return OCT_OK;
    }
  }
  long long g = (total + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  if (y->dtype == OCT_DTYPE_F32) act_bwd_kernel<float><<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(*y, *dy, act, *dz);
  else act_bwd_kernel<bf16><<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(*y, *dy, act, *dz);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}
