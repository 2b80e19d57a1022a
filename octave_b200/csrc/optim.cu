// Multi-tensor optimiser step: every parameter of a module in ONE launch (the reference ships no
// optimiser — its training script lives on an unmounted branch, README.md:39-47 — so the update rules are those of
// torch.optim.SGD / torch.optim.AdamW, which the parity tests compare against).
//   SGD   : g' = g + wd*w;  m = first ? g' : mu*m + g';  w -= lr*m
//   AdamW : w *= 1 - lr*wd;  m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;  w -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
// A block owns kOptElemsPerBlock consecutive elements of one tensor (block -> job by binary search over block_start);
// 16-byte loads and stores, every operand read once and written once.
#include "common.cuh"
#include "../../include/octave_b200.h"

namespace {

constexpr int kOptElemsPerBlock = 4096;   // 256 threads x 4 float4

__device__ __forceinline__ const OctaveOptJob& find_job(const OctaveOptJob* __restrict__ jobs, int n_jobs, long long b) {
  int lo = 0, hi = n_jobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].block_start <= b) lo = mid; else hi = mid - 1;
  }
  return jobs[lo];
}

template <int ALGO>   // 0 SGD, 1 AdamW
__global__ void __launch_bounds__(256) optim_multi_kernel(const OctaveOptJob* __restrict__ jobs, int n_jobs, const OctaveOptHyper h) {
  const OctaveOptJob j = find_job(jobs, n_jobs, blockIdx.x);
  const long long base = ((long long)blockIdx.x - j.block_start) * kOptElemsPerBlock;
  float* __restrict__ w = j.w;
  const float* __restrict__ g = j.g;
  float* __restrict__ m = j.m;
  float* __restrict__ v = j.v;
  const bool vec = (j.n % 4 == 0) && ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                                       reinterpret_cast<uintptr_t>(v)) % 16 == 0);
  const float step_size = ALGO == 1 ? h.lr / h.bias_correction1 : h.lr;
  const float inv_sqrt_bc2 = ALGO == 1 ? rsqrtf(h.bias_correction2) : 1.f;
  auto upd = [&](float& wi, float gi, float& mi, float& vi) {
    if (ALGO == 0) {
      gi = h.weight_decay != 0.f ? gi + h.weight_decay * wi : gi;
      mi = h.first_step ? gi : __fadd_rn(__fmul_rn(h.momentum, mi), gi);
      wi = __fadd_rn(wi, __fmul_rn(-h.lr, h.momentum != 0.f ? mi : gi));
    } else {
      wi = wi * (1.f - h.lr * h.weight_decay);
      mi = h.beta1 * mi + (1.f - h.beta1) * gi;
      vi = h.beta2 * vi + (1.f - h.beta2) * gi * gi;
      wi = wi - step_size * (mi / (sqrtf(vi) * inv_sqrt_bc2 + h.eps));
    }
  };
  if (vec) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = base + ((long long)u * 256 + threadIdx.x) * 4;
      if (i >= j.n) break;
      float4 wv = *reinterpret_cast<float4*>(w + i);
      const float4 gv = __ldg(reinterpret_cast<const float4*>(g + i));
      float4 mv = (ALGO == 0 && (h.first_step || h.momentum == 0.f)) ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<float4*>(m + i);
      float4 vv = ALGO == 1 ? *reinterpret_cast<float4*>(v + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      upd(wv.x, gv.x, mv.x, vv.x); upd(wv.y, gv.y, mv.y, vv.y); upd(wv.z, gv.z, mv.z, vv.z); upd(wv.w, gv.w, mv.w, vv.w);
      *reinterpret_cast<float4*>(w + i) = wv;
      if (ALGO == 1 || h.momentum != 0.f) *reinterpret_cast<float4*>(m + i) = mv;
      if (ALGO == 1) *reinterpret_cast<float4*>(v + i) = vv;
    }
  } else {
    for (long long i = base + threadIdx.x; i < j.n && i < base + kOptElemsPerBlock; i += 256) {
      float wi = w[i], mi = (ALGO == 0 && (h.first_step || h.momentum == 0.f)) ? 0.f : m[i], vi = ALGO == 1 ? v[i] : 0.f;
      upd(wi, g[i], mi, vi);
      w[i] = wi;
      if (ALGO == 1 || h.momentum != 0.f) m[i] = mi;
      if (ALGO == 1) v[i] = vi;
    }
  }
}

// ---- bf16 gradient buckets: gather-and-cast / scatter-and-cast of a bucket's tensors, the job table passed by value
struct GradJobs { OctaveGradJob j[OCTAVE_GRAD_MAX_JOBS]; int n; };

template <bool PACK>
__global__ void __launch_bounds__(256) grad_bucket_kernel(const __grid_constant__ GradJobs jobs, bf16* __restrict__ flat) {
  int lo = 0, hi = jobs.n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs.j[mid].block_start <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const OctaveGradJob& j = jobs.j[lo];
  const long long base = ((long long)blockIdx.x - j.block_start) * kOptElemsPerBlock;
  float* __restrict__ g = j.g;
  bf16* __restrict__ f = flat + j.flat_off;
  const int n = j.n;
  if ((n % 8 == 0) && (reinterpret_cast<uintptr_t>(g) % 16 == 0)) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long i = base + ((long long)u * 256 + threadIdx.x) * 8;
      if (i >= n) break;
      if (PACK) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(g + i)), b = __ldg(reinterpret_cast<const float4*>(g + i + 4));
        uint4 pk;
        pk.x = bf16x2_pack(a.x, a.y); pk.y = bf16x2_pack(a.z, a.w); pk.z = bf16x2_pack(b.x, b.y); pk.w = bf16x2_pack(b.z, b.w);
        *reinterpret_cast<uint4*>(f + i) = pk;
      } else {
        const uint4 pk = *reinterpret_cast<const uint4*>(f + i);
        float4 a, b;
        bf16x2_unpack(pk.x, a.x, a.y); bf16x2_unpack(pk.y, a.z, a.w); bf16x2_unpack(pk.z, b.x, b.y); bf16x2_unpack(pk.w, b.z, b.w);
        *reinterpret_cast<float4*>(g + i) = a;
        *reinterpret_cast<float4*>(g + i + 4) = b;
      }
    }
  } else {
    for (long long i = base + threadIdx.x; i < n && i < base + kOptElemsPerBlock; i += 256) {
      if (PACK) f[i] = __float2bfloat16_rn(g[i]); else g[i] = __bfloat162float(f[i]);
    }
  }
}

template <bool PACK>
int launch_grad_bucket(const OctaveGradJob* jobs, int32_t n_jobs, void* flat, void* stream) {
  if (!jobs || !flat || n_jobs <= 0 || n_jobs > OCTAVE_GRAD_MAX_JOBS) return OCT_ERR_INVALID;
  GradJobs js;
  js.n = n_jobs;
  long long blocks = 0;
  for (int i = 0; i < n_jobs; ++i) {
    if (!jobs[i].g || jobs[i].n <= 0 || (jobs[i].flat_off & 7) || jobs[i].block_start != blocks) return OCT_ERR_INVALID;
    js.j[i] = jobs[i];
    blocks += (jobs[i].n + kOptElemsPerBlock - 1) / kOptElemsPerBlock;
  }
  if (blocks > 0x7fffffffLL) return OCT_ERR_INVALID;
  grad_bucket_kernel<PACK><<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(js, reinterpret_cast<bf16*>(flat));
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

}  // namespace

extern "C" int octave_grad_pack_bf16(const OctaveGradJob* jobs, int32_t n_jobs, void* flat_bf16, void* stream) {
  return launch_grad_bucket<true>(jobs, n_jobs, flat_bf16, stream);
}
extern "C" int octave_grad_unpack_bf16(const OctaveGradJob* jobs, int32_t n_jobs, const void* flat_bf16, void* stream) {
  return launch_grad_bucket<false>(jobs, n_jobs, const_cast<void*>(flat_bf16), stream);
}

extern "C" int64_t octave_optim_job_blocks(int64_t n) { return n <= 0 ? 0 : (n + kOptElemsPerBlock - 1) / kOptElemsPerBlock; }

extern "C" int octave_optim_multi(const OctaveOptJob* jobs_device, int32_t n_jobs, int64_t total_blocks, const OctaveOptHyper* h,
                                  void* stream) {
  if (!jobs_device || !h || n_jobs <= 0 || total_blocks <= 0 || total_blocks > 0x7fffffffLL) return OCT_ERR_INVALID;
  if (h->algo != OCT_OPT_SGD && h->algo != OCT_OPT_ADAMW) return OCT_ERR_UNSUPPORTED;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (h->algo == OCT_OPT_SGD) optim_multi_kernel<0><<<(unsigned)total_blocks, 256, 0, s>>>(jobs_device, n_jobs, *h);
  else optim_multi_kernel<1><<<(unsigned)total_blocks, 256, 0, s>>>(jobs_device, n_jobs, *h);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}
