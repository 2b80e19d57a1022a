// K6 pools, K8 pointwise heads (attention gate / final 1x1 classifier) and layout plumbing.  NHWC.
// Reference arithmetic: nn.MaxPool2d / nn.AvgPool2d as configured at /root/reference/architectures/extra/resnest.py:189,340,383;
// AdversarialAttentionGate.forward /root/reference/architectures/segmentor/blocks.py:38-46; ResnestUNet.fc compose.py:181.
#include "common.cuh"
#include "head_k2.cuh"
#include "../../include/octave_b200.h"

namespace {

bool view_ok(const OctaveAct* a, int cmul = 8) {
  if (!a || !a->data) return false;
  if (a->dtype != OCT_DTYPE_F32 && a->dtype != OCT_DTYPE_BF16) return false;
  if (a->C % cmul || a->ld % cmul || a->coff % cmul) return false;
  return a->B > 0 && a->H > 0 && a->W > 0;
}

template <typename T>
__device__ __forceinline__ T* at(const OctaveAct& a, long long pix, int c) {
  return reinterpret_cast<T*>(a.data) + pix * a.ld + a.coff + c;
}

__host__ __device__ inline int pool_out(int in, int k, int s, int pad, int ceil_mode) {
  int num = in + 2 * pad - k;
  int o = (ceil_mode ? (num + s - 1) / s : num / s) + 1;
  if (ceil_mode && (o - 1) * s >= in + pad) --o;
  return o;
}

#define DISPATCH_T(dtype, ...)                                     \
  do {                                                             \
    if ((dtype) == OCT_DTYPE_F32) { using T = float; __VA_ARGS__; } \
    else { using T = bf16; __VA_ARGS__; }                          \
  } while (0)

// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void pool_fwd_kernel(const OctavePoolDesc pd, const OctaveAct x, const OctaveAct y, uint8_t* argmax) {
  const int G = x.C >> 3;
  const long long total = (long long)y.B * y.H * y.W * G;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(idx % G);
    long long pix = idx / G;
    const int ow = (int)(pix % y.W);
    const int oh = (int)((pix / y.W) % y.H);
    const int n = (int)(pix / ((long long)y.W * y.H));
    int hs = oh * pd.stride - pd.pad, ws = ow * pd.stride - pd.pad;
    int he = min(hs + pd.k, x.H + pd.pad), we = min(ws + pd.k, x.W + pd.pad);
    const int pool_size = (he - hs) * (we - ws);
    const int hs0 = hs, ws0 = ws;
    hs = max(hs, 0); ws = max(ws, 0); he = min(he, x.H); we = min(we, x.W);
    float acc[8];
    int am[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i] = pd.kind == 0 ? -INFINITY : 0.f; am[i] = 0; }
    for (int ih = hs; ih < he; ++ih)
      for (int iw = ws; iw < we; ++iw) {
        float f[8];
        VecIO<T, 8>::ld(at<T>(x, ((long long)n * x.H + ih) * x.W + iw, cg * 8), f);
        const int pos = (ih - hs0) * pd.k + (iw - ws0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (pd.kind == 0) {
            if (f[i] > acc[i] || f[i] != f[i]) { acc[i] = f[i]; am[i] = pos; }
          } else {
            acc[i] += f[i];
          }
        }
      }
    if (pd.kind == 1) {
      const int div = pd.count_include_pad ? pool_size : (he - hs) * (we - ws);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] /= (float)div;
    } else if (argmax) {
      uint2 pk;
      pk.x = am[0] | (am[1] << 8) | (am[2] << 16) | (am[3] << 24);
      pk.y = am[4] | (am[5] << 8) | (am[6] << 16) | (am[7] << 24);
      *reinterpret_cast<uint2*>(argmax + pix * x.C + cg * 8) = pk;
    }
    VecIO<T, 8>::st(at<T>(y, pix, cg * 8), acc);
  }
}

// gather form: every input pixel sums the contributions of the (few) windows that contain it
template <typename T>
__global__ void pool_bwd_kernel(const OctavePoolDesc pd, const OctaveAct dy, const uint8_t* argmax, const OctaveAct dx) {
  const int G = dx.C >> 3;
  const long long total = (long long)dx.B * dx.H * dx.W * G;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(idx % G);
    long long pix = idx / G;
    const int iw = (int)(pix % dx.W);
    const int ih = (int)((pix / dx.W) % dx.H);
    const int n = (int)(pix / ((long long)dx.W * dx.H));
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    // windows oh with oh*s - pad <= ih < oh*s - pad + k
    int oh_lo = ih + pd.pad - pd.k + 1;
    oh_lo = oh_lo <= 0 ? 0 : (oh_lo + pd.stride - 1) / pd.stride;
    int oh_hi = min((ih + pd.pad) / pd.stride, dy.H - 1);
    int ow_lo = iw + pd.pad - pd.k + 1;
    ow_lo = ow_lo <= 0 ? 0 : (ow_lo + pd.stride - 1) / pd.stride;
    int ow_hi = min((iw + pd.pad) / pd.stride, dy.W - 1);
    for (int oh = oh_lo; oh <= oh_hi; ++oh)
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        const long long op = ((long long)n * dy.H + oh) * dy.W + ow;
        float d[8];
        VecIO<T, 8>::ld(at<T>(dy, op, cg * 8), d);
        const int hs0 = oh * pd.stride - pd.pad, ws0 = ow * pd.stride - pd.pad;
        if (pd.kind == 0) {
          const uint2 pk = *reinterpret_cast<const uint2*>(argmax + op * dx.C + cg * 8);
          const int pos = (ih - hs0) * pd.k + (iw - ws0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int a = ((i < 4 ? pk.x : pk.y) >> (8 * (i & 3))) & 0xff;
            if (a == pos) acc[i] += d[i];
          }
        } else {
          int he = min(hs0 + pd.k, dx.H + pd.pad), we = min(ws0 + pd.k, dx.W + pd.pad);
          const int pool_size = (he - hs0) * (we - ws0);
          const int hs = max(hs0, 0), ws = max(ws0, 0);
          he = min(he, dx.H); we = min(we, dx.W);
          const float inv = 1.f / (float)(pd.count_include_pad ? pool_size : (he - hs) * (we - ws));
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += d[i] * inv;
        }
      }
    VecIO<T, 8>::st(at<T>(dx, pix, cg * 8), acc);
  }
}

// Fast path of the gather form for k <= 2*stride (every pool of the network: 3/2/1 max and average, 2/2 average): at most
// NW x NW windows contain an input pixel, so all their dy / argmax loads — for U items — are issued before the first use
// (one dependent 16-byte load per loop trip kept this kernel at 1.9 TB/s).  32-bit index arithmetic.
template <typename T, int NW, int U>
__global__ void __launch_bounds__(256) pool_bwd_win_kernel(const OctavePoolDesc pd, const OctaveAct dy, const uint8_t* __restrict__ argmax,
                                                          const OctaveAct dx) {
  constexpr int NWW = NW * NW;
  const unsigned G = dx.C >> 3;
  const unsigned total = (unsigned)dx.B * dx.H * dx.W * G;
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned base = blockIdx.x * blockDim.x + threadIdx.x; base < total; base += stride * U) {
    Raw8<T> raw[U][NWW];
    uint2 am[U][NWW];
    int ihs[U], iws[U], ohl[U], owl[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned idx = base + u * stride;
      ihs[u] = -1;
#pragma unroll
      for (int j = 0; j < NWW; ++j) { raw[u][j].zero(); am[u][j] = make_uint2(0xffffffffu, 0xffffffffu); }
      if (idx < total) {
        const unsigned cg = idx % G, pix = idx / G;
        const int iw = (int)(pix % dx.W), ih = (int)((pix / dx.W) % dx.H);
        const unsigned n = pix / ((unsigned)dx.W * dx.H);
        int oh_lo = ih + pd.pad - pd.k + 1;
        oh_lo = oh_lo <= 0 ? 0 : (oh_lo + pd.stride - 1) / pd.stride;
        const int oh_hi = min((ih + pd.pad) / pd.stride, dy.H - 1);
        int ow_lo = iw + pd.pad - pd.k + 1;
        ow_lo = ow_lo <= 0 ? 0 : (ow_lo + pd.stride - 1) / pd.stride;
        const int ow_hi = min((iw + pd.pad) / pd.stride, dy.W - 1);
        ihs[u] = ih; iws[u] = iw; ohl[u] = oh_lo; owl[u] = ow_lo;
#pragma unroll
        for (int a = 0; a < NW; ++a)
#pragma unroll
          for (int b = 0; b < NW; ++b) {
            const int oh = oh_lo + a, ow = ow_lo + b;
            if (oh <= oh_hi && ow <= ow_hi) {
              const long long op = ((long long)n * dy.H + oh) * dy.W + ow;
              raw[u][a * NW + b].ld(at<T>(dy, op, cg * 8));
              if (pd.kind == 0) am[u][a * NW + b] = __ldg(reinterpret_cast<const uint2*>(argmax + op * dx.C + cg * 8));
            }
          }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (ihs[u] < 0) continue;
      const unsigned idx = base + u * stride;
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
      for (int a = 0; a < NW; ++a)
#pragma unroll
        for (int b = 0; b < NW; ++b) {
          float d[8];
          raw[u][a * NW + b].get(d);      // zeros for a window that does not exist
          const int hs0 = (ohl[u] + a) * pd.stride - pd.pad, ws0 = (owl[u] + b) * pd.stride - pd.pad;
          if (pd.kind == 0) {
            const uint2 pk = am[u][a * NW + b];   // 0xff never equals a window position (k <= 15)
            const int pos = (ihs[u] - hs0) * pd.k + (iws[u] - ws0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int am_i = ((i < 4 ? pk.x : pk.y) >> (8 * (i & 3))) & 0xff;
              if (am_i == pos) acc[i] += d[i];
            }
          } else {
            int he = min(hs0 + pd.k, dx.H + pd.pad), we = min(ws0 + pd.k, dx.W + pd.pad);
            const int pool_size = (he - hs0) * (we - ws0);
            const int hs = max(hs0, 0), ws = max(ws0, 0);
            he = min(he, dx.H); we = min(we, dx.W);
            const int div = pd.count_include_pad ? pool_size : (he - hs) * (we - ws);
            const float inv = div > 0 ? 1.f / (float)div : 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += d[i] * inv;
          }
        }
      VecIO<T, 8>::st(at<T>(dx, idx / G, (int)(idx % G) * 8), acc);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Heads.  LP lanes cooperate on one pixel; lane l owns channel chunks {l, l+LP, ...} of 8 channels.
// ---------------------------------------------------------------------------------------------------
constexpr int KMAX = 8;
constexpr int MAXCH = 4;  // chunks per lane => C <= 32 lanes * 4 * 8 = 1024

template <int LP>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T, int LP>
__global__ void __launch_bounds__(256) head_fwd_kernel(const OctaveAct x, const float* w, const float* b, int K, int mode,
                                                       float* out, const OctaveAct gated) {
  const int C = x.C, nch = C / (8 * LP);
  const int sub = threadIdx.x % LP;
  const long long hw = (long long)x.H * x.W, npix = (long long)x.B * hw;
  const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LP;
  const long long gstride = ((long long)gridDim.x * blockDim.x) / LP;
  for (long long p0 = gid; p0 < ((npix + gstride - 1) / gstride) * gstride; p0 += gstride) {
    const bool act = p0 < npix;
    const long long p = act ? p0 : npix - 1;
    float f[MAXCH][8];
    float logit[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) logit[k] = 0.f;
#pragma unroll
    for (int j = 0; j < MAXCH; ++j) {
      if (j < nch) {
        const int c = (j * LP + sub) * 8;
        VecIO<T, 8>::ld(at<T>(x, p, c), f[j]);
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          if (k < K) {
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + (long long)k * C + c));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + (long long)k * C + c) + 1);
            logit[k] += f[j][0] * w0.x + f[j][1] * w0.y + f[j][2] * w0.z + f[j][3] * w0.w + f[j][4] * w1.x +
                        f[j][5] * w1.y + f[j][6] * w1.z + f[j][7] * w1.w;
          }
        }
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k < K) {
        logit[k] = group_sum<LP>(logit[k]) + b[k];
        mx = fmaxf(mx, logit[k]);
      }
    }
    const long long n = p / hw, q = p - n * hw;
    if (mode == 0) {
      if (act && sub == 0) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K) out[(n * K + k) * hw + q] = logit[k];
      }
    } else {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) { logit[k] = __expf(logit[k] - mx); s += logit[k]; }
      const float inv = 1.f / s;
      float mask = 0.f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) { logit[k] *= inv; if (k >= 1) mask += logit[k]; }
      if (act) {
        if (sub == 0) {
#pragma unroll
          for (int k = 0; k < KMAX; ++k)
            if (k < K) out[(n * K + k) * hw + q] = logit[k];
        }
#pragma unroll
        for (int j = 0; j < MAXCH; ++j) {
          if (j < nch) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[j][i] *= mask;
            VecIO<T, 8>::st(at<T>(gated, p, (j * LP + sub) * 8), f[j]);
          }
        }
      }
    }
  }
}

template <typename T, int LP>
__global__ void __launch_bounds__(256) head_bwd_kernel(const OctaveAct x, const float* w, const float* b, int K, int mode,
                                                       const float* dout, const OctaveAct dgated, const OctaveAct dx,
                                                       float* dlogits) {
  const int C = x.C, nch = C / (8 * LP);
  const int sub = threadIdx.x % LP;
  const long long hw = (long long)x.H * x.W, npix = (long long)x.B * hw;
  const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LP;
  const long long gstride = ((long long)gridDim.x * blockDim.x) / LP;
  for (long long p0 = gid; p0 < ((npix + gstride - 1) / gstride) * gstride; p0 += gstride) {
    const bool act = p0 < npix;
    const long long p = act ? p0 : npix - 1;
    const long long n = p / hw, q = p - n * hw;
    float dl[KMAX];
    float mask = 1.f;
    float dg[MAXCH][8];
    if (mode == 0) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k) dl[k] = (k < K && dout) ? dout[(n * K + k) * hw + q] : 0.f;
    } else {
      float f[MAXCH][8];
      float logit[KMAX];
#pragma unroll
      for (int k = 0; k < KMAX; ++k) logit[k] = 0.f;
      float dmask = 0.f;
#pragma unroll
      for (int j = 0; j < MAXCH; ++j) {
        if (j < nch) {
          const int c = (j * LP + sub) * 8;
          VecIO<T, 8>::ld(at<T>(x, p, c), f[j]);
          VecIO<T, 8>::ld(at<T>(dgated, p, c), dg[j]);
#pragma unroll
          for (int i = 0; i < 8; ++i) dmask += f[j][i] * dg[j][i];
#pragma unroll
          for (int k = 0; k < KMAX; ++k) {
            if (k < K) {
              const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + (long long)k * C + c));
              const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + (long long)k * C + c) + 1);
              logit[k] += f[j][0] * w0.x + f[j][1] * w0.y + f[j][2] * w0.z + f[j][3] * w0.w + f[j][4] * w1.x +
                          f[j][5] * w1.y + f[j][6] * w1.z + f[j][7] * w1.w;
            }
          }
        }
      }
      dmask = group_sum<LP>(dmask);
      float mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) { logit[k] = group_sum<LP>(logit[k]) + b[k]; mx = fmaxf(mx, logit[k]); }
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) { logit[k] = __expf(logit[k] - mx); s += logit[k]; }
      const float inv = 1.f / s;
      float dot = 0.f;
      mask = 0.f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if (k < K) {
          logit[k] *= inv;  // p_k
          if (k >= 1) mask += logit[k];
          dl[k] = (dout ? dout[(n * K + k) * hw + q] : 0.f) + (k >= 1 ? dmask : 0.f);  // dL/dp_k
          dot += dl[k] * logit[k];
        }
      }
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) dl[k] = logit[k] * (dl[k] - dot);  // dL/dlogit_k
    }
    if (act && sub == 0 && dlogits) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) dlogits[(n * K + k) * hw + q] = dl[k];
    }
    if (act) {
#pragma unroll
      for (int j = 0; j < MAXCH; ++j) {
        if (j < nch) {
          const int c = (j * LP + sub) * 8;
          float o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = mode == 1 ? dg[j][i] * mask : 0.f;
#pragma unroll
          for (int k = 0; k < KMAX; ++k) {
            if (k < K) {
              const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + (long long)k * C + c));
              const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + (long long)k * C + c) + 1);
              o[0] += dl[k] * w0.x; o[1] += dl[k] * w0.y; o[2] += dl[k] * w0.z; o[3] += dl[k] * w0.w;
              o[4] += dl[k] * w1.x; o[5] += dl[k] * w1.y; o[6] += dl[k] * w1.z; o[7] += dl[k] * w1.w;
            }
          }
          VecIO<T, 8>::st(at<T>(dx, p, c), o);
        }
      }
    }
  }
}

// dW[k][c] = sum_p dlogits[k][p] * x[p][c]; db[k] = sum_p dlogits[k][p].  Thread keeps 8 channels.
template <typename T>
__global__ void head_wgrad_kernel(const OctaveAct x, const float* dlogits, int K, float* dw, float* db) {
  extern __shared__ float sm[];
  const int G = x.C >> 3, cg = threadIdx.x % G, lane = threadIdx.x / G, ppb = blockDim.x / G;
  const long long hw = (long long)x.H * x.W, base = (long long)blockIdx.y * hw;
  for (int k = 0; k < K; ++k) {
    float v[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) v[i] = 0.f;
    const float* dl = dlogits + ((long long)blockIdx.y * K + k) * hw;
    for (long long p = (long long)blockIdx.x * ppb + lane; p < hw; p += (long long)gridDim.x * ppb) {
      float f[8];
      VecIO<T, 8>::ld(at<T>(x, base + p, cg * 8), f);
      const float d = dl[p];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += d * f[i];
      v[8] += d;
    }
    // fold over pixel lanes
    const int bs = blockDim.x;
#pragma unroll
    for (int i = 0; i < 9; ++i) sm[i * bs + threadIdx.x] = v[i];
    __syncthreads();
    if (threadIdx.x < G) {
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        float s = 0.f;
        for (int l = threadIdx.x; l < bs; l += G) s += sm[i * bs + l];
        if (i < 8) atomicAdd(dw + (long long)k * x.C + cg * 8 + i, s);
        else if (cg == 0) atomicAdd(db + k, s);
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* src, int Cs, const OctaveAct dst, const float* noise, int clip) {
  const long long hw = (long long)dst.H * dst.W, total = (long long)dst.B * hw * dst.C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % dst.C);
    const long long pix = idx / dst.C;
    const long long n = pix / hw, q = pix - n * hw;
    float v = 0.f;
    if (c < Cs) {
      v = src[(n * Cs + c) * hw + q];
      if (noise) v += noise[q];
      if (clip) v = fminf(fmaxf(v, 0.f), 1.f);
    }
    *at<T>(dst, pix, c) = from_f<T>(v);
  }
}

template <typename T>
__global__ void nhwc_to_nchw_clipmask_kernel(const OctaveAct src, const float* x, const float* noise, int clip, float* dst) {
  const long long hw = (long long)src.H * src.W, total = (long long)src.B * hw * src.C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long q = idx % hw;
    const int c = (int)((idx / hw) % src.C);
    const long long n = idx / (hw * src.C);
    float g = to_f(*at<T>(src, n * hw + q, c));
    if (clip) {
      const float v = x[idx] + (noise ? noise[q] : 0.f);
      if (!(v >= 0.f && v <= 1.f)) g = 0.f;
    }
    dst[idx] = g;
  }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const OctaveAct src, float* dst, int accumulate) {
  // 32x32 tile transpose through shared memory: reads coalesced over channels, writes coalesced over pixels
  __shared__ float tile[32][33];
  const long long hw = (long long)src.H * src.W;
  const int n = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long long p = p0 + r;
    const int c = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (p < hw && c < src.C) ? to_f(*at<T>(src, n * hw + p, c)) : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r;
    const long long p = p0 + threadIdx.x;
    if (p < hw && c < src.C) {
      float* o = dst + ((long long)n * src.C + c) * hw + p;
      *o = accumulate ? *o + tile[threadIdx.x][r] : tile[threadIdx.x][r];
    }
  }
}

template <typename T>
__global__ void copy_window_kernel(const OctaveAct src, const OctaveAct dst, int accumulate) {
  const int G = dst.C >> 3;
  const long long total = (long long)dst.B * dst.H * dst.W * G;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(idx % G);
    const long long pix = idx / G;
    const int w = (int)(pix % dst.W), h = (int)((pix / dst.W) % dst.H);
    const long long n = pix / ((long long)dst.W * dst.H);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    if (h < src.H && w < src.W) VecIO<T, 8>::ld(at<T>(src, (n * src.H + h) * src.W + w, cg * 8), v);
    if (accumulate) {
      float o[8];
      VecIO<T, 8>::ld(at<T>(dst, pix, cg * 8), o);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += o[i];
    }
    VecIO<T, 8>::st(at<T>(dst, pix, cg * 8), v);
  }
}

// chan_sum (nullable, fp64 [C], pre-zeroed): per-channel sum of src over all pixels — the bias gradient of the
// ConvTranspose2d whose output gradient is being rearranged (every src element is read exactly once here).
template <typename T, typename I>
__global__ void __launch_bounds__(256) space_to_depth_kernel(const OctaveAct src, const OctaveAct dst, double* chan_sum) {
  __shared__ float sm[8 * 256];
  constexpr int U = 4;   // independent 16-byte loads in flight per thread
  const I G = src.C >> 3;
  const I total = (I)dst.B * dst.H * dst.W * 4 * G;
  const I stride = (I)gridDim.x * blockDim.x;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (I base = (I)blockIdx.x * blockDim.x + threadIdx.x; base < total; base += stride * U) {
    Raw8<T> raw[U];
    T* out[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const I idx = base + (I)u * stride;
      raw[u].zero();
      out[u] = nullptr;
      if (idx < total) {
        const int cg = (int)(idx % G);
        const I r = idx / G;
        const int t = (int)(r % 4);
        const I pix = r / 4;
        const int w = (int)(pix % dst.W), h = (int)((pix / dst.W) % dst.H);
        const I n = pix / ((I)dst.W * dst.H);
        const int sh = 2 * h + (t >> 1), sw = 2 * w + (t & 1);
        out[u] = at<T>(dst, (long long)pix, t * src.C + cg * 8);
        if (sh < src.H && sw < src.W) raw[u].ld(at<T>(src, ((long long)n * src.H + sh) * src.W + sw, cg * 8));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (out[u]) {
        float v[8];
        raw[u].get(v);
        VecIO<T, 8>::st(out[u], v);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += v[i];
      }
    }
  }
  if (chan_sum) {
    // launcher guarantees 256 % G == 0: a thread keeps the channel group tid % G for the whole grid-stride loop
    fold_lanes<8>(acc, sm, (int)G);
    if (threadIdx.x < G) {
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(chan_sum + threadIdx.x * 8 + i, (double)acc[i]);
    }
  }
}

template <typename T>
__global__ void depth_to_space_kernel(const OctaveAct src, const OctaveAct dst) {
  const int G = dst.C >> 3;
  const long long total = (long long)dst.B * dst.H * dst.W * G;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(idx % G);
    const long long pix = idx / G;
    const int w = (int)(pix % dst.W), h = (int)((pix / dst.W) % dst.H);
    const long long n = pix / ((long long)dst.W * dst.H);
    const int t = (h & 1) * 2 + (w & 1);
    float v[8];
    VecIO<T, 8>::ld(at<T>(src, (n * src.H + (h >> 1)) * src.W + (w >> 1), t * dst.C + cg * 8), v);
    VecIO<T, 8>::st(at<T>(dst, pix, cg * 8), v);
  }
}

int grid_for(long long total, int bs) {
  long long g = (total + bs - 1) / bs;
  const long long cap = 148LL * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

int pick_lp(int C) {
  int lp = C / 8;
  if (lp > 32) lp = 32;
  return lp;
}

}  // namespace

extern "C" int octave_pool_out_size(const OctavePoolDesc* p, int32_t in) {
  if (!p || p->k <= 0 || p->stride <= 0) return OCT_ERR_INVALID;
  return pool_out(in, p->k, p->stride, p->pad, p->ceil_mode);
}

extern "C" int octave_pool_fwd(const OctavePoolDesc* p, const OctaveAct* x, const OctaveAct* y, uint8_t* argmax, void* stream) {
  if (!p || !view_ok(x) || !view_ok(y) || x->C != y->C || x->B != y->B || x->dtype != y->dtype) return OCT_ERR_INVALID;
  if (p->k > 15 || y->H != pool_out(x->H, p->k, p->stride, p->pad, p->ceil_mode) ||
      y->W != pool_out(x->W, p->k, p->stride, p->pad, p->ceil_mode))
    return OCT_ERR_INVALID;
  const long long total = (long long)y->B * y->H * y->W * (x->C / 8);
  DISPATCH_T(x->dtype, (pool_fwd_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*p, *x, *y, argmax)));
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_pool_bwd(const OctavePoolDesc* p, const OctaveAct* dy, const uint8_t* argmax, const OctaveAct* dx,
                               void* stream) {
  if (!p || !view_ok(dy) || !view_ok(dx) || dy->C != dx->C || dy->B != dx->B || dy->dtype != dx->dtype) return OCT_ERR_INVALID;
  if (p->kind == 0 && !argmax) return OCT_ERR_INVALID;
  const long long total = (long long)dx->B * dx->H * dx->W * (dx->C / 8);
  const int grid = grid_for(total, 256);
  const bool idx32 = total + 4LL * grid * 256 < (1LL << 32) && (long long)dy->B * dy->H * dy->W < (1LL << 31);
  if (idx32 && p->k <= 15 && p->k <= p->stride) {
    DISPATCH_T(dx->dtype, (pool_bwd_win_kernel<T, 1, 4><<<grid_for((total + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(*p, *dy, argmax, *dx)));
  } else if (idx32 && p->k <= 15 && p->k <= 2 * p->stride) {
    DISPATCH_T(dx->dtype, (pool_bwd_win_kernel<T, 2, 1><<<grid, 256, 0, (cudaStream_t)stream>>>(*p, *dy, argmax, *dx)));
  } else {
    DISPATCH_T(dx->dtype, (pool_bwd_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(*p, *dy, argmax, *dx)));
  }
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

#define DISPATCH_LP(lp, ...)                         \
  switch (lp) {                                      \
    case 1: { constexpr int LP = 1; __VA_ARGS__; break; }   \
    case 2: { constexpr int LP = 2; __VA_ARGS__; break; }   \
    case 4: { constexpr int LP = 4; __VA_ARGS__; break; }   \
    case 8: { constexpr int LP = 8; __VA_ARGS__; break; }   \
    case 16: { constexpr int LP = 16; __VA_ARGS__; break; } \
    default: { constexpr int LP = 32; __VA_ARGS__; break; } \
  }

static int head_check(const OctaveAct* x, int K) {
  if (!view_ok(x) || K <= 0 || K > KMAX) return OCT_ERR_INVALID;
  const int lp = pick_lp(x->C);
  if ((lp & (lp - 1)) || x->C % (8 * lp) || x->C / (8 * lp) > MAXCH) return OCT_ERR_UNSUPPORTED;
  return OCT_OK;
}

// one wave of resident blocks for the persistent K=2 head kernels
template <typename F>
static int head_wave(F fn, size_t smem) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, 256, smem) != cudaSuccess || n < 1) n = 1;
  int sms = octave_sm_count();
  if (sms <= 0) sms = 148;
  return n * sms;
}

template <typename T, int LP, int NCH, int MODE>
static void launch_head_k2_fwd(const OctaveAct* x, const float* w, const float* b, float* out, const OctaveAct& g,
                               cudaStream_t s) {
  constexpr int U = 4 / NCH;
  auto fn = head_k2::fwd_kernel<T, LP, NCH, U, MODE>;
  static int wave = 0;
  if (!wave) wave = head_wave(fn, 0);
  const long long npix = (long long)x->B * x->H * x->W;
  long long need = (npix + (256 / LP) * U - 1) / ((256 / LP) * U);
  const int grid = (int)(need < wave ? (need < 1 ? 1 : need) : wave);
  fn<<<grid, 256, 0, s>>>(*x, w, b, out, g);
}

template <typename T, int LP, int NCH, int MODE>
static void launch_head_k2_bwd(const OctaveAct* x, const float* w, const float* b, const float* dout, const OctaveAct& g,
                               const OctaveAct* dx, float* dw, float* db, cudaStream_t s) {
  constexpr int U = 4 / NCH;
  auto fn = head_k2::bwd_kernel<T, LP, NCH, U, MODE>;
  constexpr size_t smem = 18 * 256 * sizeof(float);
  static int wave = 0;
  if (!wave) wave = head_wave(fn, smem);
  const long long npix = (long long)x->B * x->H * x->W;
  long long need = (npix + (256 / LP) * U * 4 - 1) / ((256 / LP) * U * 4);   // >= 4 iterations per block: amortise the fold
  const int grid = (int)(need < wave ? (need < 1 ? 1 : need) : wave);
  fn<<<grid, 256, smem, s>>>(*x, w, b, dout, g, *dx, dw, db);
}

#define DISPATCH_K2(lp, nch, ...)                                                        \
  switch (lp) {                                                                          \
    case 4: { constexpr int LP = 4; constexpr int NCH = 1; __VA_ARGS__; break; }         \
    case 8: { constexpr int LP = 8; constexpr int NCH = 1; __VA_ARGS__; break; }         \
    case 16: { constexpr int LP = 16; constexpr int NCH = 1; __VA_ARGS__; break; }       \
    default:                                                                             \
      if (nch == 1) { constexpr int LP = 32; constexpr int NCH = 1; __VA_ARGS__; }       \
      else if (nch == 2) { constexpr int LP = 32; constexpr int NCH = 2; __VA_ARGS__; }  \
      else { constexpr int LP = 32; constexpr int NCH = 4; __VA_ARGS__; }                \
      break;                                                                             \
  }

static bool head_k2_ok(const OctaveAct* x, int K) {
  const int lp = pick_lp(x->C), nch = x->C / (8 * lp);
  return K == 2 && lp >= 4 && (nch == 1 || nch == 2 || nch == 4);
}

extern "C" int octave_head_fwd(const OctaveAct* x, const float* w, const float* b, int32_t K, int32_t mode, float* out,
                               const OctaveAct* gated, void* stream) {
  int rc = head_check(x, K);
  if (rc != OCT_OK) return rc;
  if (!w || !b || !out) return OCT_ERR_INVALID;
  if (mode == 1 && (!view_ok(gated) || gated->C != x->C)) return OCT_ERR_INVALID;
  const int lp = pick_lp(x->C);
  const long long npix = (long long)x->B * x->H * x->W;
  OctaveAct g = gated ? *gated : *x;
  if (head_k2_ok(x, K)) {
    const int nch = x->C / (8 * lp);
    cudaStream_t s = (cudaStream_t)stream;
    DISPATCH_T(x->dtype, DISPATCH_K2(lp, nch, {
      if (mode == 1) launch_head_k2_fwd<T, LP, NCH, 1>(x, w, b, out, g, s);
      else launch_head_k2_fwd<T, LP, NCH, 0>(x, w, b, out, g, s);
    }));
    OCT_CHECK_LAUNCH();
    return OCT_OK;
  }
  const int grid = grid_for(npix * lp, 256);
  DISPATCH_T(x->dtype, DISPATCH_LP(lp, (head_fwd_kernel<T, LP><<<grid, 256, 0, (cudaStream_t)stream>>>(*x, w, b, K, mode, out, g))));
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_head_bwd(const OctaveAct* x, const float* w, const float* b, int32_t K, int32_t mode, const float* dout,
                               const OctaveAct* dgated, const OctaveAct* dx, float* dlogits, float* dw, float* db,
                               void* stream) {
  int rc = head_check(x, K);
  if (rc != OCT_OK) return rc;
  if (!w || !b || !view_ok(dx) || dx->C != x->C) return OCT_ERR_INVALID;
  if (mode == 1 && (!view_ok(dgated) || dgated->C != x->C)) return OCT_ERR_INVALID;
  if ((dw != nullptr) != (db != nullptr)) return OCT_ERR_INVALID;
  const int lp = pick_lp(x->C);
  const long long npix = (long long)x->B * x->H * x->W;
  OctaveAct g = dgated ? *dgated : *x;
  cudaStream_t s = (cudaStream_t)stream;
  if (head_k2_ok(x, K) && !dlogits) {
    if (dw) {
      if (cudaMemsetAsync(dw, 0, sizeof(float) * K * x->C, s) != cudaSuccess) return OCT_ERR_LAUNCH;
      if (cudaMemsetAsync(db, 0, sizeof(float) * K, s) != cudaSuccess) return OCT_ERR_LAUNCH;
    }
    const int nch = x->C / (8 * lp);
    DISPATCH_T(x->dtype, DISPATCH_K2(lp, nch, {
      if (mode == 1) launch_head_k2_bwd<T, LP, NCH, 1>(x, w, b, dout, g, dx, dw, db, s);
      else launch_head_k2_bwd<T, LP, NCH, 0>(x, w, b, dout, g, dx, dw, db, s);
    }));
    OCT_CHECK_LAUNCH();
    return OCT_OK;
  }
  // generic path: dlogits scratch is required; parameter gradients come from octave_head_wgrad
  if (!dlogits) return OCT_ERR_INVALID;
  const int grid = grid_for(npix * lp, 256);
  DISPATCH_T(x->dtype, DISPATCH_LP(lp, (head_bwd_kernel<T, LP><<<grid, 256, 0, s>>>(*x, w, b, K, mode, dout, g, *dx, dlogits))));
  OCT_CHECK_LAUNCH();
  if (dw) return octave_head_wgrad(x, dlogits, K, dw, db, stream);
  return OCT_OK;
}

extern "C" int octave_head_wgrad(const OctaveAct* x, const float* dlogits, int32_t K, float* dw, float* db, void* stream) {
  if (!view_ok(x) || !dlogits || !dw || !db || K <= 0 || K > KMAX) return OCT_ERR_INVALID;
  const int G = x->C / 8;
  if (G > 256) return OCT_ERR_UNSUPPORTED;
  const int bs = (256 / G) * G, ppb = bs / G;
  const long long hw = (long long)x->H * x->W;
  long long bx = (hw + ppb - 1) / ppb;
  long long cap = (148 * 8 + x->B - 1) / x->B;
  if (bx > cap) bx = cap;
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(dw, 0, sizeof(float) * K * x->C, s) != cudaSuccess) return OCT_ERR_LAUNCH;
  if (cudaMemsetAsync(db, 0, sizeof(float) * K, s) != cudaSuccess) return OCT_ERR_LAUNCH;
  DISPATCH_T(x->dtype, (head_wgrad_kernel<T><<<dim3((unsigned)bx, x->B), bs, 9 * bs * sizeof(float), s>>>(*x, dlogits, K, dw, db)));
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_nchw_to_nhwc(const float* src, int32_t C_src, const OctaveAct* dst, void* stream) {
  if (!src || !view_ok(dst, 1) || C_src > dst->C) return OCT_ERR_INVALID;
  const long long total = (long long)dst->B * dst->H * dst->W * dst->C;
  DISPATCH_T(dst->dtype, (nchw_to_nhwc_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, C_src, *dst, nullptr, 0)));
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_nchw_to_nhwc_noise(const float* src, int32_t C_src, const float* noise, int32_t clip, const OctaveAct* dst,
                                         void* stream) {
  if (!src || !view_ok(dst, 1) || C_src > dst->C) return OCT_ERR_INVALID;
  const long long total = (long long)dst->B * dst->H * dst->W * dst->C;
  DISPATCH_T(dst->dtype, (nchw_to_nhwc_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, C_src, *dst, noise, clip)));
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_nhwc_to_nchw_clipmask(const OctaveAct* src, const float* x, const float* noise, int32_t clip, float* dst,
                                            void* stream) {
  if (!view_ok(src, 1) || !dst || (clip && !x)) return OCT_ERR_INVALID;
  const long long total = (long long)src->B * src->H * src->W * src->C;
  DISPATCH_T(src->dtype, (nhwc_to_nchw_clipmask_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*src, x, noise, clip, dst)));
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_nhwc_to_nchw(const OctaveAct* src, float* dst, int32_t accumulate, void* stream) {
  if (!view_ok(src, 1) || !dst) return OCT_ERR_INVALID;
  const long long hw = (long long)src->H * src->W;
  dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((src->C + 31) / 32), (unsigned)src->B);
  DISPATCH_T(src->dtype, (nhwc_to_nchw_kernel<T><<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(*src, dst, accumulate)));
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_copy_window(const OctaveAct* src, const OctaveAct* dst, int32_t accumulate, void* stream) {
  if (!view_ok(src) || !view_ok(dst) || src->C != dst->C || src->B != dst->B || src->dtype != dst->dtype) return OCT_ERR_INVALID;
  const long long total = (long long)dst->B * dst->H * dst->W * (dst->C / 8);
  DISPATCH_T(dst->dtype, (copy_window_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*src, *dst, accumulate)));
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_space_to_depth(const OctaveAct* src, const OctaveAct* dst, double* chan_sum, void* stream) {
  if (!view_ok(src) || !view_ok(dst) || dst->C != 4 * src->C || src->B != dst->B || src->dtype != dst->dtype) return OCT_ERR_INVALID;
  const int G = src->C / 8;
  if (chan_sum && (G > 256 || 256 % G)) return OCT_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  if (chan_sum && !g_octave_stats_prezeroed && cudaMemsetAsync(chan_sum, 0, sizeof(double) * src->C, s) != cudaSuccess) return OCT_ERR_LAUNCH;
  const long long total = (long long)dst->B * dst->H * dst->W * 4 * G;
  int grid = grid_for((total + 3) / 4, 256);
  if (chan_sum && grid > 148 * 8) grid = 148 * 8;    // bounds the atomics per channel
  if (total + 4LL * grid * 256 < (1LL << 32)) {
    DISPATCH_T(dst->dtype, (space_to_depth_kernel<T, unsigned><<<grid, 256, 0, s>>>(*src, *dst, chan_sum)));
  } else {
    DISPATCH_T(dst->dtype, (space_to_depth_kernel<T, unsigned long long><<<grid, 256, 0, s>>>(*src, *dst, chan_sum)));
  }
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_depth_to_space(const OctaveAct* src, const OctaveAct* dst, void* stream) {
  if (!view_ok(src) || !view_ok(dst) || src->C != 4 * dst->C || src->B != dst->B || src->dtype != dst->dtype) return OCT_ERR_INVALID;
  if (dst->H > 2 * src->H || dst->W > 2 * src->W) return OCT_ERR_INVALID;
  const long long total = (long long)dst->B * dst->H * dst->W * (dst->C / 8);
  DISPATCH_T(dst->dtype, (depth_to_space_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*src, *dst)));
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}
