// Shared device helpers for the octave_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define OCT_OK 0
#define OCT_ERR_INVALID (-1)      // bad argument (null pointer, shape out of contract)
#define OCT_ERR_UNSUPPORTED (-2)  // shape/dtype outside what this entry point handles
#define OCT_ERR_LAUNCH (-3)       // cudaGetLastError() != success after the launch

#define OCT_DTYPE_F32 0
#define OCT_DTYPE_BF16 1

// every kernel launch site is followed by exactly one OCT_CHECK_LAUNCH(): it also counts launches
extern unsigned long long g_octave_launches;
extern int g_octave_deterministic;
extern int g_octave_stats_prezeroed;   // octave_set_deterministic(): single-writer reductions, no fp32 atomics
#define OCT_CHECK_LAUNCH()                                   \
  do {                                                       \
    ++g_octave_launches;                                     \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) return OCT_ERR_LAUNCH;           \
  } while (0)

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum of NV values per thread. `red` must hold NV * (blockDim.x/32) floats.
// Result valid in warp 0 lane 0 (returned to every thread of warp 0).
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) red[i * nw + warp] = v[i];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float t = lane < nw ? red[i * nw + lane] : 0.f;
      v[i] = warp_sum(t);
    }
  }
  __syncthreads();
}

// ---- vector loads/stores of N consecutive elements, converted to/from float ----
template <typename T, int N> struct VecIO;

template <> struct VecIO<float, 8> {
  static __device__ __forceinline__ void ld(const float* p, float* v) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void st(float* p, const float* v) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct VecIO<float, 4> {
  static __device__ __forceinline__ void ld(const float* p, float* v) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  }
  static __device__ __forceinline__ void st(float* p, const float* v) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct VecIO<float, 2> {
  static __device__ __forceinline__ void ld(const float* p, float* v) {
    float2 a = __ldg(reinterpret_cast<const float2*>(p));
    v[0] = a.x; v[1] = a.y;
  }
  static __device__ __forceinline__ void st(float* p, const float* v) {
    reinterpret_cast<float2*>(p)[0] = make_float2(v[0], v[1]);
  }
};
template <> struct VecIO<float, 1> {
  static __device__ __forceinline__ void ld(const float* p, float* v) { v[0] = __ldg(p); }
  static __device__ __forceinline__ void st(float* p, const float* v) { p[0] = v[0]; }
};

__device__ __forceinline__ void bf16x2_unpack(uint32_t u, float& lo, float& hi) {
  lo = __uint_as_float(u << 16);
  hi = __uint_as_float(u & 0xffff0000u);
}
// Column statistics on the packed-fp32 pipe (sm_100: add / fma .f32x2): (s.x, s.y) += (a, b); (q.x, q.y) += (a*a, b*b) for the
// bf16 pair u = (a, b).  Four instructions per pair instead of six; same roundings as the scalar FADD / FFMA form.
__device__ __forceinline__ void stat_acc_bf16x2(uint32_t u, unsigned long long& s, unsigned long long& q) {
  const uint32_t lo = u << 16, hi = u & 0xffff0000u;
  unsigned long long ab;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ab) : "r"(lo), "r"(hi));
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(s) : "l"(ab));
  asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(q) : "l"(ab));
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float& lo, float& hi) {
  uint32_t a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(v));
  lo = __uint_as_float(a);
  hi = __uint_as_float(b);
}

__device__ __forceinline__ uint32_t bf16x2_pack(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <> struct VecIO<bf16, 8> {
  static __device__ __forceinline__ void ld(const bf16* p, float* v) {
    uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
    bf16x2_unpack(a.x, v[0], v[1]); bf16x2_unpack(a.y, v[2], v[3]);
    bf16x2_unpack(a.z, v[4], v[5]); bf16x2_unpack(a.w, v[6], v[7]);
  }
  static __device__ __forceinline__ void st(bf16* p, const float* v) {
    uint4 a;
    a.x = bf16x2_pack(v[0], v[1]); a.y = bf16x2_pack(v[2], v[3]);
    a.z = bf16x2_pack(v[4], v[5]); a.w = bf16x2_pack(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = a;
  }
};
template <> struct VecIO<bf16, 4> {
  static __device__ __forceinline__ void ld(const bf16* p, float* v) {
    uint2 a = __ldg(reinterpret_cast<const uint2*>(p));
    bf16x2_unpack(a.x, v[0], v[1]); bf16x2_unpack(a.y, v[2], v[3]);
  }
  static __device__ __forceinline__ void st(bf16* p, const float* v) {
    uint2 a;
    a.x = bf16x2_pack(v[0], v[1]); a.y = bf16x2_pack(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = a;
  }
};
template <> struct VecIO<bf16, 2> {
  static __device__ __forceinline__ void ld(const bf16* p, float* v) {
    uint32_t a = __ldg(reinterpret_cast<const uint32_t*>(p));
    bf16x2_unpack(a, v[0], v[1]);
  }
  static __device__ __forceinline__ void st(bf16* p, const float* v) {
    *reinterpret_cast<uint32_t*>(p) = bf16x2_pack(v[0], v[1]);
  }
};
template <> struct VecIO<bf16, 1> {
  static __device__ __forceinline__ void ld(const bf16* p, float* v) { v[0] = __bfloat162float(*p); }
  static __device__ __forceinline__ void st(bf16* p, const float* v) { p[0] = __float2bfloat16_rn(v[0]); }
};

// 8 consecutive elements held in their storage form (4 registers for bf16) until they are consumed.
template <typename T> struct Raw8;
template <> struct Raw8<bf16> {
  uint4 r;
  __device__ __forceinline__ void ld(const bf16* p) { r = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void zero() { r = make_uint4(0u, 0u, 0u, 0u); }
  __device__ __forceinline__ bool any_nonzero() const { return ((r.x | r.y | r.z | r.w) & 0x7fff7fffu) != 0u; }   // +-0 excluded
  __device__ __forceinline__ void get(float* v) const {
    bf16x2_unpack(r.x, v[0], v[1]); bf16x2_unpack(r.y, v[2], v[3]);
    bf16x2_unpack(r.z, v[4], v[5]); bf16x2_unpack(r.w, v[6], v[7]);
  }
};
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void ld(const float* p) {
    a = __ldg(reinterpret_cast<const float4*>(p));
    b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  }
  __device__ __forceinline__ void zero() { a = b = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ bool any_nonzero() const {
    return ((__float_as_uint(a.x) | __float_as_uint(a.y) | __float_as_uint(a.z) | __float_as_uint(a.w) | __float_as_uint(b.x) |
             __float_as_uint(b.y) | __float_as_uint(b.z) | __float_as_uint(b.w)) & 0x7fffffffu) != 0u;
  }
  __device__ __forceinline__ void get(float* v) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};


// Fold NV per-thread values over the pixel lanes of a block (threads with equal tid % Gb).
// Result valid for threads tid < Gb.  sm must hold NV * blockDim.x floats.
template <int NV>
__device__ __forceinline__ void fold_lanes(float (&v)[NV], float* sm, int Gb) {
  const int tid = threadIdx.x, bs = blockDim.x;
  if (Gb < 32 && (32 % Gb) == 0 && (bs & 31) == 0) {
    // lanes of a warp that share a channel group are Gb apart: butterfly over them, then one row per warp
    for (int off = 16; off >= Gb; off >>= 1) {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
    }
    const int lane = tid & 31, warp = tid >> 5, nw = bs >> 5;
    if (lane < Gb) {
#pragma unroll
      for (int i = 0; i < NV; ++i) sm[(i * nw + warp) * Gb + lane] = v[i];
    }
    __syncthreads();
    if (tid < Gb) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float s = 0.f;
        for (int w = 0; w < nw; ++w) s += sm[(i * nw + w) * Gb + tid];
        v[i] = s;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) sm[i * bs + tid] = v[i];
    __syncthreads();
    if (tid < Gb) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float s = 0.f;
        for (int l = tid; l < bs; l += Gb) s += sm[i * bs + l];
        v[i] = s;
      }
    }
  }
}


// Per-column sum / sum of squares of a warp's staged 32 x COLS_W bf16 tile (row pitch PITCH bytes), for the BatchNorm
// statistics fused into the conv epilogues.  A lane owns the column PAIR cp = lane % (COLS_W/2) (one 32-bit shared load
// covers both columns) and every (32 / (COLS_W/2))-th row; the row groups are folded with shuffles.  acc = {sum a, sum b,
// sum sq a, sum sq b} accumulates across tiles and is meaningful on lanes < COLS_W/2.
template <int COLS_W, int PITCH>
__device__ __forceinline__ void tile_col_stats(const uint8_t* stage, int lane, float (&acc)[4]) {
  constexpr int CP = COLS_W / 2, RG = 32 / CP, RPG = 32 / RG;
  const int cp = lane % CP, rg = lane / CP;
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
  for (int i = 0; i < RPG; ++i) {
    const int row = i * RG + rg;
    const uint32_t u = *reinterpret_cast<const uint32_t*>(stage + row * PITCH + cp * 4);
    float a, b;
    bf16x2_unpack(u, a, b);
    s0 += a; s1 += b;
    q0 += a * a; q1 += b * b;
  }
#pragma unroll
  for (int off = 16; off >= CP; off >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, off);
    s1 += __shfl_xor_sync(0xffffffffu, s1, off);
    q0 += __shfl_xor_sync(0xffffffffu, q0, off);
    q1 += __shfl_xor_sync(0xffffffffu, q1, off);
  }
  acc[0] += s0; acc[1] += s1; acc[2] += q0; acc[3] += q1;
}
