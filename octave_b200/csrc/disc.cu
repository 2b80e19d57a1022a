// Discriminator-specific plumbing: space-to-depth layout changes (with InstanceNoise + clip fused), weight remapping of
// the 4x4 stride-2 convs onto 3x3 stride-1 convs over space-to-depth inputs (so they run on the tcgen05 kernel), and the
// full-extent output conv as a per-sample dot product.
// Reference: /root/reference/architectures/discriminator/blocks.py:46-50,68-71,91-109,149-154.
#include "common.cuh"
#include "../../include/octave_b200.h"

namespace {

// One thread per (destination pixel, quadrant): writes channels [coff, qs) of its quadrant — the C source channels (+ noise,
// clipped), then zeros — so the caller's pad channels need no memset; quadrant pixels outside the source (odd H / W) are
// written as zeros.  Four neighbouring threads fill one destination row: coalesced stores, source rows read with stride 2.
template <typename T>
__global__ void nchw_to_s2d_kernel(const float* src, int B, int C, int H, int W, const float* noise, int clip, OctaveAct dst,
                                   int qs, int coff) {
  const long long total = (long long)B * dst.H * dst.W * 4;
  const int nw = qs - coff;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(idx & 3);
    const long long pix = idx >> 2;
    const int w2 = (int)(pix % dst.W);
    const int h2 = (int)((pix / dst.W) % dst.H);
    const long long n = pix / ((long long)dst.W * dst.H);
    const int h = 2 * h2 + (q >> 1), w = 2 * w2 + (q & 1);
    const bool in = h < H && w < W;
    const float nz = (in && noise) ? noise[(long long)h * W + w] : 0.f;
    T* o = reinterpret_cast<T*>(dst.data) + pix * dst.ld + dst.coff + q * qs + coff;
    const float* sp = src + ((n * C) * H + h) * (long long)W + w;
    if (sizeof(T) == 2 && nw == 8 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
      float f[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float v = 0.f;
        if (in && c < C) {
          v = sp[(long long)c * H * W] + nz;
          if (clip) v = fminf(fmaxf(v, 0.f), 1.f);
        }
        f[c] = v;
      }
      uint4 pk;
      pk.x = bf16x2_pack(f[0], f[1]); pk.y = bf16x2_pack(f[2], f[3]);
      pk.z = bf16x2_pack(f[4], f[5]); pk.w = bf16x2_pack(f[6], f[7]);
      *reinterpret_cast<uint4*>(o) = pk;
    } else {
      for (int c = 0; c < nw; ++c) {
        float v = 0.f;
        if (in && c < C) {
          v = sp[(long long)c * H * W] + nz;
          if (clip) v = fminf(fmaxf(v, 0.f), 1.f);
        }
        o[c] = from_f<T>(v);
      }
    }
  }
}

template <typename T>
__global__ void s2d_to_nchw_kernel(OctaveAct src, int qs, int coff, int C, int H, int W, const float* x, const float* noise,
                                   int clip, float* dst) {
  const long long hw = (long long)H * W, total = (long long)src.B * C * hw;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long q = idx % hw;
    const int c = (int)((idx / hw) % C);
    const long long n = idx / (hw * C);
    const int h = (int)(q / W), w = (int)(q - (long long)h * W);
    const long long pix = (n * src.H + (h >> 1)) * src.W + (w >> 1);
    float g = to_f(reinterpret_cast<const T*>(src.data)[pix * src.ld + src.coff + ((h & 1) * 2 + (w & 1)) * qs + coff + c]);
    if (clip) {
      const float v = x[idx] + (noise ? noise[q] : 0.f);
      if (!(v >= 0.f && v <= 1.f)) g = 0.f;
    }
    dst[idx] = g;
  }
}

// remapped weight W3[co][q*qs + c][tap] = W[co][c][kh][kw], kh = 2*(tap/3 - 1) + (q>>1) + 1, kw likewise
__device__ __forceinline__ float w3_at(const float* w, int cin, int qs, int ksz, int co, int k, int tap) {
  const int q = k / qs, c = k - q * qs;
  if (c >= cin) return 0.f;
  const int kh = 2 * (tap / 3 - 1) + (q >> 1) + 1, kw = 2 * (tap % 3 - 1) + (q & 1) + 1;
  if (kh < 0 || kh >= ksz || kw < 0 || kw >= ksz) return 0.f;
  return w[(((long long)co * cin + c) * ksz + kh) * ksz + kw];
}

__global__ void pack_s2d_kernel(const float* w, const float* scale, int mode, int cout, int cin, int qs, int ksz, bf16* out) {
  const int K = 4 * qs;
  const long long total = 9LL * cout * K;
  const float sc = scale ? scale[0] : 1.f;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    float v;
    if (mode == 0) {  // [tap][co][k]
      const int k = (int)(idx % K);
      const int co = (int)((idx / K) % cout);
      const int tap = (int)(idx / ((long long)K * cout));
      v = w3_at(w, cin, qs, ksz, co, k, tap);
    } else {          // [tap'][k][co], tap flipped
      const int co = (int)(idx % cout);
      const int k = (int)((idx / cout) % K);
      const int tap = (int)(idx / ((long long)K * cout));
      v = w3_at(w, cin, qs, ksz, co, k, 8 - tap);
    }
    out[idx] = __float2bfloat16_rn(v * sc);
  }
}

__global__ void unpack_wgrad_s2d_kernel(const float* dw3, int cout, int cin, int qs, int ksz, float* dw) {
  const int kk = ksz * ksz;
  const long long total = (long long)cout * cin * kk;
  const int K = 4 * qs;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int kw = (int)(idx % ksz), kh = (int)((idx / ksz) % ksz);
    const int c = (int)((idx / kk) % cin);
    const int co = (int)(idx / ((long long)kk * cin));
    const int i = (kh + 1) & 1, j = (kw + 1) & 1;
    const int th = ((kh + 1) >> 1), tw = ((kw + 1) >> 1);  // dh'+1, dw'+1
    const int k = (i * 2 + j) * qs + c;
    dw[idx] = dw3[((long long)co * K + k) * 9 + th * 3 + tw];
  }
}

template <typename T>
__global__ void __launch_bounds__(1024) rowdot_fwd_kernel(OctaveAct x, const float* w, const float* bias, float* out) {
  // grid (1, B); x is dense per sample: n = H*W*C elements
  __shared__ float red[32];
  const long long n = (long long)x.H * x.W * x.C;
  const T* xp = reinterpret_cast<const T*>(x.data) + (long long)blockIdx.y * n;
  float acc = 0.f;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += (long long)gridDim.x * blockDim.x * 8) {
    float f[8];
    VecIO<T, 8>::ld(xp + i, f);
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + i)), w1 = __ldg(reinterpret_cast<const float4*>(w + i) + 1);
    acc += f[0] * w0.x + f[1] * w0.y + f[2] * w0.z + f[3] * w0.w + f[4] * w1.x + f[5] * w1.y + f[6] * w1.z + f[7] * w1.w;
  }
  float v[1] = {acc};
  block_sum<1>(v, red);
  // ONE block per sample (gridDim.x == 1): a single writer, so the critic logit is bit-reproducible
  if (threadIdx.x == 0) out[blockIdx.y] = v[0] + (bias ? bias[0] : 0.f);
}

template <typename T>
__global__ void rowdot_bwd_kernel(OctaveAct x, const float* w, const float* g, OctaveAct dx, float* dw, float* dbias) {
  const long long n = (long long)x.H * x.W * x.C;
  const int B = x.B;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += (long long)gridDim.x * blockDim.x * 8) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + i)), w1 = __ldg(reinterpret_cast<const float4*>(w + i) + 1);
    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int b = 0; b < B; ++b) {
      const float gb = g[b];
      float f[8], o[8];
      if (dw) VecIO<T, 8>::ld(reinterpret_cast<const T*>(x.data) + (long long)b * n + i, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        o[k] = gb * wv[k];
        if (dw) acc[k] += gb * f[k];
      }
      VecIO<T, 8>::st(reinterpret_cast<T*>(dx.data) + (long long)b * n + i, o);
    }
    if (dw) {
#pragma unroll
      for (int k = 0; k < 8; ++k) dw[i + k] = acc[k];
    }
  }
  if (dbias && blockIdx.x == 0 && threadIdx.x == 0) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += g[b];
    dbias[0] = s;
  }
}

int grid_for(long long total, int bs) {
  long long g = (total + bs - 1) / bs;
  const long long cap = 148LL * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

bool dense(const OctaveAct* a) { return a && a->data && a->coff == 0 && a->ld == a->C && (a->C % 8) == 0; }

}  // namespace

extern "C" int octave_nchw_to_s2d(const float* src, int32_t B, int32_t C, int32_t H, int32_t W, const float* noise, int32_t clip,
                                  const OctaveAct* dst, int32_t qs, int32_t coff, void* stream) {
  if (!src || !dst || !dst->data || B <= 0 || C <= 0 || qs <= 0 || coff + C > qs) return OCT_ERR_INVALID;
  if (dst->H != (H + 1) / 2 || dst->W != (W + 1) / 2 || dst->C < 4 * qs || dst->B != B) return OCT_ERR_INVALID;
  const long long total = (long long)B * dst->H * dst->W * 4;
  if (dst->dtype == OCT_DTYPE_F32)
    nchw_to_s2d_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, B, C, H, W, noise, clip, *dst, qs, coff);
  else
    nchw_to_s2d_kernel<bf16><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, B, C, H, W, noise, clip, *dst, qs, coff);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_s2d_to_nchw(const OctaveAct* src, int32_t qs, int32_t coff, int32_t C, int32_t H, int32_t W, const float* x,
                                  const float* noise, int32_t clip, float* dst, void* stream) {
  if (!src || !src->data || !dst || qs <= 0 || coff + C > qs || (clip && !x)) return OCT_ERR_INVALID;
  if (src->H != (H + 1) / 2 || src->W != (W + 1) / 2 || src->C < 4 * qs) return OCT_ERR_INVALID;
  const long long total = (long long)src->B * C * H * W;
  if (src->dtype == OCT_DTYPE_F32)
    s2d_to_nchw_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*src, qs, coff, C, H, W, x, noise, clip, dst);
  else
    s2d_to_nchw_kernel<bf16><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*src, qs, coff, C, H, W, x, noise, clip, dst);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_pack_weight_s2d(const float* w, const float* scale, int32_t mode, int32_t cout, int32_t cin, int32_t qs,
                                      int32_t ksize, void* out, void* stream) {
  if (!w || !out || cout <= 0 || cin <= 0 || cin > qs || (mode != 0 && mode != 1) || (ksize != 3 && ksize != 4)) return OCT_ERR_INVALID;
  const long long total = 9LL * cout * 4 * qs;
  pack_s2d_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(w, scale, mode, cout, cin, qs, ksize, (bf16*)out);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_unpack_wgrad_s2d(const float* dw3, int32_t cout, int32_t cin, int32_t qs, int32_t ksize, float* dw, void* stream) {
  if (!dw3 || !dw || cin > qs || (ksize != 3 && ksize != 4)) return OCT_ERR_INVALID;
  const long long total = (long long)cout * cin * ksize * ksize;
  unpack_wgrad_s2d_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(dw3, cout, cin, qs, ksize, dw);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_rowdot_fwd(const OctaveAct* x, const float* w, const float* bias, float* out, void* stream) {
  if (!dense(x) || !w || !out) return OCT_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  if (x->dtype == OCT_DTYPE_F32) rowdot_fwd_kernel<float><<<dim3(1, x->B), 1024, 0, s>>>(*x, w, bias, out);
  else rowdot_fwd_kernel<bf16><<<dim3(1, x->B), 1024, 0, s>>>(*x, w, bias, out);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_rowdot_bwd(const OctaveAct* x, const float* w, const float* g, const OctaveAct* dx, float* dw, float* dbias,
                                 void* stream) {
  if (!dense(x) || !dense(dx) || !w || !g || dx->C != x->C || dx->dtype != x->dtype) return OCT_ERR_INVALID;
  const long long n = (long long)x->H * x->W * x->C;
  const int gx = grid_for(n / 8, 256);
  if (x->dtype == OCT_DTYPE_F32) rowdot_bwd_kernel<float><<<gx, 256, 0, (cudaStream_t)stream>>>(*x, w, g, *dx, dw, dbias);
  else rowdot_bwd_kernel<bf16><<<gx, 256, 0, (cudaStream_t)stream>>>(*x, w, g, *dx, dw, dbias);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

// ---- spectral norm: one power iteration and sigma in ONE launch (torch.nn.utils.spectral_norm, legacy hook semantics used by
// DiscriminatorBlock, discriminator/blocks.py:101-104: n_power_iterations = 1, eps = 1e-12):
//   training:  v <- normalize(W^T u);  u <- normalize(W v);  sigma = u . (W v)      (u, v updated in place)
//   eval:      sigma = u . (W v)
// W is [rows][cols] fp32 (weight_orig viewed as a matrix).  One block: the matrices here are at most 1024 x 240.  Fixed
// summation order (no atomics): the critic's weights are bit-reproducible from run to run.
namespace {
constexpr int kSnThreads = 1024;
__device__ __forceinline__ void spectral_sigma_body(const float* __restrict__ W, int rows, int cols, float* __restrict__ u,
                                                    float* __restrict__ v, int training, float eps,
                                                    float* __restrict__ out /* sigma, 1/sigma */,
                                                    float* __restrict__ snap_u, float* __restrict__ snap_v) {
  extern __shared__ float sm[];
  float* us = sm;                 // [rows]
  float* vs = sm + rows;          // [cols]
  float* part = vs + cols;        // [4][256]
  __shared__ float red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int j = tid; j < rows; j += kSnThreads) us[j] = u[j];
  for (int k = tid; k < cols; k += kSnThreads) vs[k] = v[k];
  __syncthreads();
  auto block_total = [&](float x) {
    x = warp_sum(x);
    __syncthreads();
    if (lane == 0) red[warp] = x;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < kSnThreads / 32; ++w) t += red[w];     // every thread adds the 32 warp sums in the same order
    return t;
  };
  if (training) {
    // v = W^T u: thread (c, rg) sums rows rg, rg+4, ... of column k0 + c
    const int c = tid & 255, rg = tid >> 8;
    for (int k0 = 0; k0 < cols; k0 += 256) {
      const int k = k0 + c;
      float acc = 0.f;
      if (k < cols) {
#pragma unroll 8
        for (int j = rg; j < rows; j += 4) acc += __ldg(W + (size_t)j * cols + k) * us[j];
      }
      part[rg * 256 + c] = acc;
      __syncthreads();
      if (rg == 0 && k < cols) vs[k] = (part[c] + part[256 + c]) + (part[512 + c] + part[768 + c]);
      __syncthreads();
    }
    float sq = 0.f;
    for (int k = tid; k < cols; k += kSnThreads) sq += vs[k] * vs[k];
    const float nv = fmaxf(sqrtf(block_total(sq)), eps);
    for (int k = tid; k < cols; k += kSnThreads) { vs[k] /= nv; v[k] = vs[k]; }
    __syncthreads();
  }
  // wv = W v: one warp per row
  float dot_uw = 0.f, sq_w = 0.f;
  for (int j = warp; j < rows; j += kSnThreads / 32) {
    float acc = 0.f;
    for (int k = lane; k < cols; k += 32) acc += __ldg(W + (size_t)j * cols + k) * vs[k];
    acc = warp_sum(acc);
    if (lane == 0) {
      sq_w += acc * acc;
      dot_uw += us[j] * acc;
      part[j & 1023] = acc;          // rows <= 1024 fit; larger matrices recompute below
    }
  }
  const float nw2 = block_total(sq_w);
  float sigma;
  if (training) {
    const float nu = fmaxf(sqrtf(nw2), eps);
    if (rows <= 1024) {
      for (int j = tid; j < rows; j += kSnThreads) u[j] = part[j] / nu;
    }
    sigma = nw2 / nu;                // u . (W v) with u = W v / nu
    if (snap_u) {
      for (int j = tid; j < rows; j += kSnThreads) snap_u[j] = part[j] / nu;
    }
  } else {
    sigma = block_total(dot_uw);
    if (snap_u) {
      for (int j = tid; j < rows; j += kSnThreads) snap_u[j] = us[j];
    }
  }
  if (snap_v) {
    for (int k = tid; k < cols; k += kSnThreads) snap_v[k] = vs[k];
  }
  if (tid == 0) { out[0] = sigma; out[1] = 1.f / sigma; }
}

__global__ void __launch_bounds__(kSnThreads) spectral_sigma_kernel(const float* __restrict__ W, int rows, int cols, float* __restrict__ u,
                                                                    float* __restrict__ v, int training, float eps,
                                                                    float* __restrict__ out) {
  spectral_sigma_body(W, rows, cols, u, v, training, eps, out, nullptr, nullptr);
}

// every spectral-norm layer of one critic call in ONE launch (one block per layer; the layers' power iterations are
// independent of each other and of the activations), leaving a snapshot of the u / v this call used for the backward pass
struct SnJobs { OctaveSnJob j[OCTAVE_SN_MAX_JOBS]; };
__global__ void __launch_bounds__(kSnThreads) spectral_sigma_multi_kernel(const SnJobs jobs, int training, float eps) {
  const OctaveSnJob& jb = jobs.j[blockIdx.x];
  spectral_sigma_body(jb.W, jb.rows, jb.cols, jb.u, jb.v, training, eps, jb.out, jb.out + 2, jb.out + 2 + jb.rows);
}

// ---- spectral norm, weight gradient.  W = W_orig / sigma with sigma = u^T W_orig v (u, v constants of the backward pass):
//   dW_orig = dW / sigma - <dW, W_orig> / sigma^2 * u v^T
// Two launches, no atomics: fixed-order partial dot products, then every block adds the partials in the same order.
constexpr int kSwParts = 64;
__global__ void __launch_bounds__(256) spectral_wgrad_dot_kernel(const float* __restrict__ dw, const float* __restrict__ wo, long long n,
                                                                 float* __restrict__ parts) {
  __shared__ float red[8];
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  const long long lo = per * blockIdx.x, hi = lo + per < n ? lo + per : n;
  float acc = 0.f;
  for (long long i = lo + threadIdx.x; i < hi; i += 256) acc += dw[i] * wo[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    parts[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256) spectral_wgrad_apply_kernel(const float* __restrict__ dw, const float* __restrict__ u,
                                                                   const float* __restrict__ v, const float* __restrict__ sigma1,
                                                                   const float* __restrict__ parts, int nparts, int rows, int cols,
                                                                   float* __restrict__ out, int accumulate) {
  float dot = 0.f;
  for (int i = 0; i < nparts; ++i) dot += parts[i];
  const float sigma = sigma1[0];
  const float coef = dot / (sigma * sigma);
  const long long n = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    const float g = dw[i] / sigma - coef * (u[r] * v[c]);
    out[i] = accumulate ? out[i] + g : g;
  }
}
}  // namespace

extern "C" int octave_spectral_sigma_multi(const OctaveSnJob* jobs, int32_t n_jobs, int32_t training, float eps, void* stream) {
  if (!jobs || n_jobs <= 0 || n_jobs > OCTAVE_SN_MAX_JOBS) return OCT_ERR_INVALID;
  SnJobs js;
  size_t smem = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const OctaveSnJob& jb = jobs[i];
    if (!jb.W || !jb.u || !jb.v || !jb.out || jb.rows <= 0 || jb.cols <= 0) return OCT_ERR_INVALID;
    if (jb.rows > 1024 || jb.cols > 8192) return OCT_ERR_UNSUPPORTED;
    js.j[i] = jb;
    const size_t need = (size_t)(jb.rows + jb.cols + 1024) * sizeof(float);
    smem = need > smem ? need : smem;
  }
  spectral_sigma_multi_kernel<<<n_jobs, kSnThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(js, training, eps);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_spectral_wgrad(const float* dw, const float* w_orig, const float* u, const float* v, const float* sigma,
                                     int32_t rows, int32_t cols, float* parts, float* out, int32_t accumulate, void* stream) {
  if (!dw || !w_orig || !u || !v || !sigma || !parts || !out || rows <= 0 || cols <= 0) return OCT_ERR_INVALID;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long n = (long long)rows * cols;
  int nparts = (int)((n + 4095) / 4096);
  nparts = nparts < 1 ? 1 : (nparts > kSwParts ? kSwParts : nparts);
  spectral_wgrad_dot_kernel<<<nparts, 256, 0, s>>>(dw, w_orig, n, parts);
  OCT_CHECK_LAUNCH();
  spectral_wgrad_apply_kernel<<<grid_for(n, 256), 256, 0, s>>>(dw, u, v, sigma, parts, nparts, rows, cols, out, accumulate);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_spectral_sigma(const float* W, int32_t rows, int32_t cols, float* u, float* v, int32_t training, float eps,
                                     float* out2, void* stream) {
  if (!W || !u || !v || !out2 || rows <= 0 || cols <= 0) return OCT_ERR_INVALID;
  if (rows > 1024 || cols > 8192) return OCT_ERR_UNSUPPORTED;
  const size_t smem = (size_t)(rows + cols + 1024) * sizeof(float);
  spectral_sigma_kernel<<<1, kSnThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(W, rows, cols, u, v, training, eps, out2);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}
