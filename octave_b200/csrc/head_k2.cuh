// K8 fast path for the two-class heads (num_classes = 2, the OCTAve configuration): AdversarialAttentionGate
// (/root/reference/architectures/segmentor/blocks.py:38-46) and ResnestUNet.fc (segmentor/compose.py:181).
//
// LP lanes cooperate on one pixel (lane `sub` owns channel chunks {sub, sub+LP, ...} of 8 channels); a block of 256
// threads covers PPB = 256/LP pixels per slot and U slots per loop iteration, with every 16-byte load of the iteration
// issued before the first use.  The backward kernel also accumulates the weight / bias gradient in registers (no second
// pass over x, no dlogits round trip) and flushes it with one fold + fp32 atomics per resident block.
#pragma once
#include "common.cuh"
#include "../../include/octave_b200.h"

namespace head_k2 {

template <typename T>
__device__ __forceinline__ T* at(const OctaveAct& a, long long pix, int c) {
  return reinterpret_cast<T*>(a.data) + pix * a.ld + a.coff + c;
}

template <int LP>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// MODE 0: out = W x + b.  MODE 1: y = softmax(W x + b); gated = x * y_1; out = y.
template <typename T, int LP, int NCH, int U, int MODE>
__global__ void __launch_bounds__(256) fwd_kernel(const OctaveAct x, const float* __restrict__ w, const float* __restrict__ b,
                                                  float* __restrict__ out, const OctaveAct gated) {
  constexpr int PPB = 256 / LP;
  const int C = x.C;
  const int sub = threadIdx.x % LP, slot = threadIdx.x / LP;
  const long long hw = (long long)x.H * x.W, npix = (long long)x.B * hw;
  float w0[NCH][8], w1[NCH][8];
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    const int c = (j * LP + sub) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) { w0[j][i] = w[c + i]; w1[j][i] = w[C + c + i]; }
  }
  const float b0 = b[0], b1 = b[1];
  const long long stride = (long long)gridDim.x * PPB * U;
  for (long long p0 = (long long)blockIdx.x * PPB * U + slot; p0 < npix; p0 += stride) {
    Raw8<T> rx[U][NCH];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long q = p0 + (long long)u * PPB;
      const long long qq = q < npix ? q : npix - 1;   // every lane takes part in the shuffles below
#pragma unroll
      for (int j = 0; j < NCH; ++j) rx[u][j].ld(at<T>(x, qq, (j * LP + sub) * 8));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long q = p0 + (long long)u * PPB;
      const bool ok = q < npix;
      float f[NCH][8];
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        rx[u][j].get(f[j]);
#pragma unroll
        for (int i = 0; i < 8; ++i) { l0 += f[j][i] * w0[j][i]; l1 += f[j][i] * w1[j][i]; }
      }
      l0 = group_sum<LP>(l0) + b0;
      l1 = group_sum<LP>(l1) + b1;
      const long long n = q / hw, r = q - n * hw;
      if (MODE == 0) {
        if (ok && sub == 0) {
          out[(n * 2) * hw + r] = l0;
          out[(n * 2 + 1) * hw + r] = l1;
        }
      } else {
        const float mx = fmaxf(l0, l1);
        const float e0 = __expf(l0 - mx), e1 = __expf(l1 - mx);
        const float inv = 1.f / (e0 + e1);
        const float y0 = e0 * inv, y1 = e1 * inv;
        if (ok) {
          if (sub == 0) {
            out[(n * 2) * hw + r] = y0;
            out[(n * 2 + 1) * hw + r] = y1;
          }
#pragma unroll
          for (int j = 0; j < NCH; ++j) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[j][i] *= y1;
            VecIO<T, 8>::st(at<T>(gated, q, (j * LP + sub) * 8), f[j]);
          }
        }
      }
    }
  }
}

// dx, and (dw, db nullable, pre-zeroed) the parameter gradients.  dout nullable.
template <typename T, int LP, int NCH, int U, int MODE>
__global__ void __launch_bounds__(256) bwd_kernel(const OctaveAct x, const float* __restrict__ w, const float* __restrict__ b,
                                                  const float* __restrict__ dout, const OctaveAct dgated, const OctaveAct dx,
                                                  float* __restrict__ dw, float* __restrict__ db) {
  extern __shared__ float sm[];
  constexpr int PPB = 256 / LP;
  const int C = x.C;
  const int sub = threadIdx.x % LP, slot = threadIdx.x / LP;
  const long long hw = (long long)x.H * x.W, npix = (long long)x.B * hw;
  float w0[NCH][8], w1[NCH][8], g0[NCH][8], g1[NCH][8];
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    const int c = (j * LP + sub) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) { w0[j][i] = w[c + i]; w1[j][i] = w[C + c + i]; g0[j][i] = g1[j][i] = 0.f; }
  }
  const float b0 = b[0], b1 = b[1];
  float gb0 = 0.f, gb1 = 0.f;
  const long long stride = (long long)gridDim.x * PPB * U;
  for (long long p0 = (long long)blockIdx.x * PPB * U + slot; p0 < npix; p0 += stride) {
    Raw8<T> rx[U][NCH], rg[MODE == 1 ? U : 1][NCH];
    float d0[U], d1[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long q = p0 + (long long)u * PPB;
      const long long qq = q < npix ? q : npix - 1;
      const long long n = qq / hw, r = qq - n * hw;
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        rx[u][j].ld(at<T>(x, qq, (j * LP + sub) * 8));
        if (MODE == 1) rg[u][j].ld(at<T>(dgated, qq, (j * LP + sub) * 8));
      }
      d0[u] = dout ? __ldg(dout + (n * 2) * hw + r) : 0.f;
      d1[u] = dout ? __ldg(dout + (n * 2 + 1) * hw + r) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long q = p0 + (long long)u * PPB;
      const bool ok = q < npix;
      float f[NCH][8], dg[NCH][8];
      float dl0, dl1, y1 = 0.f;
#pragma unroll
      for (int j = 0; j < NCH; ++j) rx[u][j].get(f[j]);
      if (MODE == 0) {
        dl0 = d0[u];
        dl1 = d1[u];
      } else {
        float l0 = 0.f, l1 = 0.f, dmask = 0.f;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          rg[u][j].get(dg[j]);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            l0 += f[j][i] * w0[j][i];
            l1 += f[j][i] * w1[j][i];
            dmask += f[j][i] * dg[j][i];
          }
        }
        l0 = group_sum<LP>(l0) + b0;
        l1 = group_sum<LP>(l1) + b1;
        dmask = group_sum<LP>(dmask);
        const float mx = fmaxf(l0, l1);
        const float e0 = __expf(l0 - mx), e1 = __expf(l1 - mx);
        const float inv = 1.f / (e0 + e1);
        const float y0 = e0 * inv;
        y1 = e1 * inv;
        const float p0g = d0[u], p1g = d1[u] + dmask;   // dL/dy_k
        const float dot = p0g * y0 + p1g * y1;
        dl0 = y0 * (p0g - dot);
        dl1 = y1 * (p1g - dot);
      }
      if (!ok) { dl0 = 0.f; dl1 = 0.f; }
      if (sub == 0) { gb0 += dl0; gb1 += dl1; }
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          g0[j][i] += dl0 * f[j][i];
          g1[j][i] += dl1 * f[j][i];
          o[i] = dl0 * w0[j][i] + dl1 * w1[j][i];
          if (MODE == 1) o[i] += dg[j][i] * y1;
        }
        if (ok) VecIO<T, 8>::st(at<T>(dx, q, (j * LP + sub) * 8), o);
      }
    }
  }
  if (dw) {
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      float v[18];
#pragma unroll
      for (int i = 0; i < 8; ++i) { v[i] = g0[j][i]; v[8 + i] = g1[j][i]; }
      v[16] = j == 0 ? gb0 : 0.f;
      v[17] = j == 0 ? gb1 : 0.f;
      if (j) __syncthreads();
      fold_lanes<18>(v, sm, LP);
      if (threadIdx.x < LP) {
        const int c = (j * LP + sub) * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          atomicAdd(dw + c + i, v[i]);
          atomicAdd(dw + C + c + i, v[8 + i]);
        }
        if (j == 0 && threadIdx.x == 0 && db) {
          atomicAdd(db, v[16]);
          atomicAdd(db + 1, v[17]);
        }
      }
    }
  }
}

}  // namespace head_k2
