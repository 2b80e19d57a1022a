// Small fp32 kernels: the per-sample attention branch of SplAtConv2d
// (/root/reference/architectures/extra/resnest.py:116-127) and the weight re-packing for the tcgen05 kernels.
#include "common.cuh"
#include "../../include/octave_b200.h"

namespace {

// ---- grouped linear layers on [B][K] vectors (fc1 / fc2 of the split-attention branch) --------------------------------
// These GEMMs are tiny (B = batch rows, <= 17 MMAC) and latency-bound: the kernels below keep many independent 16-byte
// loads in flight, split the contraction over the warps of a block and reuse the activation rows across output columns.
constexpr int kJB = 8;   // output columns per block of glinear_fwd (= warps per block)

// Epilogues fused into glinear_fwd_kernel (the whole batch column of an output is in one warp when B <= 32):
//   1  BatchNorm1d over the batch + ReLU  (fc1 -> bn1 -> relu, resnest.py:118-122): out = pre-BN values (kept for backward),
//      e.y = activations, e.mi = mean / invstd, running statistics updated; same summation order as bn1d_relu_fwd_kernel
//   2  r-softmax over radix 2 (fc2 -> view(B,2,C) softmax, resnest.py:125-127): the block's 8 columns are the 4 channels
//      c0..c0+3 of both radix halves (j = r*C + c); out = attention weights, the logits are not stored
struct GlinearEpi {
  const float* gamma; const float* beta; float* rm; float* rv; long long* nbt; float eps, mom; int training;
  float* y; float* mi; int C;
};

// block: kJB columns of one group; lane = batch row (blockIdx.y selects the 32-row slab); warp w = K slice w
template <int EPI>
__global__ void __launch_bounds__(256) glinear_fwd_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                          const float* __restrict__ bias, int B, int Kt, int N, int groups,
                                                          float scale, float* __restrict__ out, const GlinearEpi e) {
  __shared__ float red[8][kJB][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Kg = Kt / groups, Ng = N / groups;
  const int j0 = blockIdx.x * kJB, g = EPI == 2 ? 0 : j0 / Ng;
  auto col = [&](int jj) { return EPI == 2 ? (jj >> 2) * e.C + (int)blockIdx.x * 4 + (jj & 3) : j0 + jj; };
  const int b = blockIdx.y * 32 + lane;
  const int bb = b < B ? b : B - 1;
  const int Ks = Kg / 8;                    // floats of this warp's K slice (multiple of 4)
  const float* arow = in + (long long)bb * Kt + g * Kg + warp * Ks;
  const float* wbase = w + warp * Ks;
  float acc[kJB];
#pragma unroll
  for (int jj = 0; jj < kJB; ++jj) acc[jj] = 0.f;
#pragma unroll 2
  for (int i = 0; i < Ks; i += 4) {
    const float4 a = *reinterpret_cast<const float4*>(arow + i);
#pragma unroll
    for (int jj = 0; jj < kJB; ++jj) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(wbase + (long long)col(jj) * Kg + i));
      acc[jj] += a.x * wv.x + a.y * wv.y + a.z * wv.z + a.w * wv.w;
    }
  }
#pragma unroll
  for (int jj = 0; jj < kJB; ++jj) red[warp][jj][lane] = acc[jj];
  __syncthreads();
  // warp w finalises column col(w)
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) sum += red[k][warp][lane];
  const int j = col(warp);
  const float val = sum * scale + (bias ? bias[j] : 0.f);
  if (EPI == 0) {
    if (b < B) out[(long long)b * N + j] = val;
    return;
  }
  __syncthreads();                          // every warp has read its partial sums: red is reused for the column values
  float* fin = &red[0][0][0];               // [8 columns][32 batch rows]
  fin[warp * 32 + lane] = val;
  __syncthreads();
  if (EPI == 1) {
    if (b < B) out[(long long)b * N + j] = val;
    float mean, invstd;
    if (e.training) {
      double sd = 0.0, qd = 0.0;
      for (int r = 0; r < B; ++r) { const double v = fin[warp * 32 + r]; sd += v; qd += v * v; }
      const double m = sd / B;
      double var = qd / B - m * m;
      if (var < 0.0) var = 0.0;
      mean = (float)m;
      invstd = (float)(1.0 / sqrt(var + (double)e.eps));
      if (lane == 0) {
        if (e.rm) e.rm[j] = (1.f - e.mom) * e.rm[j] + e.mom * mean;
        if (e.rv) e.rv[j] = (1.f - e.mom) * e.rv[j] + e.mom * (float)(B > 1 ? var * B / (B - 1.0) : var);
        if (j == 0 && e.nbt) *e.nbt += 1;
      }
    } else {
      mean = e.rm[j];
      invstd = 1.f / sqrtf(e.rv[j] + e.eps);
    }
    if (lane == 0) { e.mi[j] = mean; e.mi[N + j] = invstd; }
    if (b < B) e.y[(long long)b * N + j] = fmaxf((val - mean) * invstd * e.gamma[j] + e.beta[j], 0.f);
  } else {
    if (warp < 4 && b < B) {
      const float l0 = fin[warp * 32 + lane], l1 = fin[(4 + warp) * 32 + lane];
      const float mx = fmaxf(fmaxf(-INFINITY, l0), l1);
      float sm = 0.f;
      sm += expf(l0 - mx);
      sm += expf(l1 - mx);
      const int c = blockIdx.x * 4 + warp;
      out[(long long)b * N + c] = expf(l0 - mx) / sm;
      out[(long long)b * N + e.C + c] = expf(l1 - mx) / sm;
    }
  }
}

// din[b][i] = scale * sum_{j in group(i)} dout[b][j] * w[j][i_local]
// block: 32 consecutive inputs i (lane) x 8 j slices (warp); all 32 batch rows of the slab are accumulators of a thread
// FUSED (the fc2 -> bn1 -> relu part of the split-attention backward, B <= 32, one group, gridDim.z = 1):
//   prologue  dout is not read: dlogits = att * (datt - sum_r datt*att) (r-softmax backward, radix 2) is computed while
//             staging, and written out by block column 0 for the weight-gradient kernel;
//   epilogue  din is not written: the BatchNorm1d + ReLU backward of bn1d_relu_bwd_kernel (same summation order) is applied
//             to the block's 32 columns -> f.dx, f.dgamma, f.dbeta.
struct GlinearBwdFuse {
  const float* att; const float* datt; float* dlogits; int C;
  const float* x; const float* y; const float* gamma; const float* mi; int training;
  float* dx; float* dgamma; float* dbeta;
};
constexpr int kJC = 64;   // outputs staged per pass
template <bool FUSED>
__global__ void __launch_bounds__(256) glinear_bwd_data_kernel(const float* __restrict__ dout, const float* __restrict__ w,
                                                               int B, int Kt, int N, int groups, float scale,
                                                               float* __restrict__ din, const GlinearBwdFuse f) {
  __shared__ __align__(16) float tile[kJC][32];     // dout^T of the current pass: [j][b]
  __shared__ float red[8][32][33];                  // [warp][b][i]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Kg = Kt / groups, Ng = N / groups;
  const int i = blockIdx.x * 32 + lane;             // global input index (Kt % 32 == 0)
  const int g = i / Kg, il = i - g * Kg;
  const int b0 = blockIdx.y * 32;
  float acc[32];
#pragma unroll
  for (int r = 0; r < 32; ++r) acc[r] = 0.f;
  // blockIdx.z splits the outputs j of the group: partial sums meet in din through atomics (din zeroed by the launcher)
  const int jper = ((Ng + gridDim.z - 1) / gridDim.z + kJC - 1) / kJC * kJC;
  const int jbeg = blockIdx.z * jper, jend = min(Ng, jbeg + jper);
  for (int jc = jbeg; jc < jend; jc += kJC) {
    __syncthreads();
    // stage dout[b0 + r][g*Ng + jc + jj] -> tile[jj][r]   (lanes run over r: conflict-free shared stores)
    for (int e = threadIdx.x; e < kJC * 32; e += 256) {
      const int jj = e >> 5, r = e & 31;
      const int bq = b0 + r;
      float dv = 0.f;
      if (bq < B && jc + jj < jend) {
        if (FUSED) {
          const int j = jc + jj, c = j < f.C ? j : j - f.C;
          const long long base = (long long)bq * N + c;
          const float a0 = f.att[base], a1 = f.att[base + f.C], d0 = f.datt[base], d1 = f.datt[base + f.C];
          float dot = 0.f;
          dot += d0 * a0;
          dot += d1 * a1;
          dv = j < f.C ? a0 * (d0 - dot) : a1 * (d1 - dot);
          if (blockIdx.x == 0) f.dlogits[(long long)bq * N + j] = dv;
        } else {
          dv = __ldg(dout + (long long)bq * N + g * Ng + jc + jj);
        }
      }
      tile[jj][r] = dv;
    }
    __syncthreads();
    const int jn = min(kJC, jend - jc);
#pragma unroll 2
    for (int jj = warp; jj < jn; jj += 8) {
      const float wv = __ldg(w + (long long)(g * Ng + jc + jj) * Kg + il);
#pragma unroll
      for (int r4 = 0; r4 < 8; ++r4) {
        const float4 d = *reinterpret_cast<const float4*>(&tile[jj][r4 * 4]);
        acc[r4 * 4 + 0] += d.x * wv;
        acc[r4 * 4 + 1] += d.y * wv;
        acc[r4 * 4 + 2] += d.z * wv;
        acc[r4 * 4 + 3] += d.w * wv;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 32; ++r) red[warp][r][lane] = acc[r];
  __syncthreads();
  if (FUSED) {
    // dh[r][lane] of the block's 32 columns -> shared memory (the staging tile is free now), then bn1d_relu_bwd per column
    float* fin = &tile[0][0];                 // [32 rows][32 columns]
    __shared__ float kk[2][32];
    for (int r = warp; r < 32; r += 8) {
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) sum += red[k][r][lane];
      fin[r * 32 + lane] = sum * scale;
    }
    __syncthreads();
    const float mean = f.mi[i], invstd = f.mi[Kt + i];
    if (warp == 0) {
      float sd = 0.f, sdx = 0.f;
      for (int b = 0; b < B; ++b) {
        const long long o = (long long)b * Kt + i;
        const float d = f.y[o] > 0.f ? fin[b * 32 + lane] : 0.f;
        sd += d;
        sdx += d * (f.x[o] - mean) * invstd;
      }
      if (f.dgamma) f.dgamma[i] = sdx;
      if (f.dbeta) f.dbeta[i] = sd;
      kk[0][lane] = f.training ? sd / B : 0.f;
      kk[1][lane] = f.training ? sdx / B : 0.f;
    }
    __syncthreads();
    const float k1 = kk[0][lane], k2 = kk[1][lane], ag = f.gamma[i] * invstd;
    for (int r = warp; r < 32 && r < B; r += 8) {
      const long long o = (long long)r * Kt + i;
      const float d = f.y[o] > 0.f ? fin[r * 32 + lane] : 0.f;
      f.dx[o] = ag * (d - k1 - (f.x[o] - mean) * invstd * k2);
    }
    return;
  }
  // thread (warp, lane): rows r = warp, warp + 8, ... ; column lane
  for (int r = warp; r < 32; r += 8) {
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += red[k][r][lane];
    if (b0 + r < B) {
      float* o = din + (long long)(b0 + r) * Kt + i;
      if (gridDim.z > 1) atomicAdd(o, sum * scale); else *o = sum * scale;
    }
  }
}

// one thread per (j, i)
__global__ void glinear_bwd_weight_kernel(const float* __restrict__ dout, const float* __restrict__ in, int B, int Kt, int N,
                                          int groups, float scale, float* __restrict__ dw, float* __restrict__ dbias) {
  const int Kg = Kt / groups, Ng = N / groups;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * Kg) return;
  const int j = (int)(idx / Kg), il = (int)(idx % Kg), g = j / Ng;
  float acc = 0.f, accb = 0.f;
  const float* dp = dout + j;
  const float* ip = in + g * Kg + il;
#pragma unroll 8
  for (int b = 0; b < B; ++b) {
    const float d = __ldg(dp + (long long)b * N);
    acc += d * __ldg(ip + (long long)b * Kt);
    accb += d;
  }
  dw[idx] = acc * scale;
  if (il == 0 && dbias) dbias[j] = accb;
}

__global__ void bn1d_relu_fwd_kernel(const float* x, int B, int C, const float* gamma, const float* beta, float* rm,
                                     float* rv, long long* nbt, float eps, float mom, int training, float* y, float* mi) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, invstd;
  if (training) {
    double s = 0.0, q = 0.0;
    for (int b = 0; b < B; ++b) { const double v = x[(long long)b * C + c]; s += v; q += v * v; }
    const double m = s / B;
    double var = q / B - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    invstd = (float)(1.0 / sqrt(var + (double)eps));
    if (rm) rm[c] = (1.f - mom) * rm[c] + mom * mean;
    if (rv) rv[c] = (1.f - mom) * rv[c] + mom * (float)(B > 1 ? var * B / (B - 1.0) : var);
    if (c == 0 && nbt) *nbt += 1;
  } else {
    mean = rm[c];
    invstd = 1.f / sqrtf(rv[c] + eps);
  }
  mi[c] = mean;
  mi[C + c] = invstd;
  const float g = gamma[c], bt = beta[c];
  for (int b = 0; b < B; ++b) {
    const float v = (x[(long long)b * C + c] - mean) * invstd * g + bt;
    y[(long long)b * C + c] = fmaxf(v, 0.f);
  }
}

__global__ void bn1d_relu_bwd_kernel(const float* dy, const float* x, const float* y, int B, int C, const float* gamma,
                                     const float* mi, int training, float* dx, float* dgamma, float* dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mean = mi[c], invstd = mi[C + c];
  float sd = 0.f, sdx = 0.f;
  for (int b = 0; b < B; ++b) {
    const long long o = (long long)b * C + c;
    const float d = y[o] > 0.f ? dy[o] : 0.f;
    sd += d;
    sdx += d * (x[o] - mean) * invstd;
  }
  if (dgamma) dgamma[c] = sdx;
  if (dbeta) dbeta[c] = sd;
  const float k1 = training ? sd / B : 0.f, k2 = training ? sdx / B : 0.f, ag = gamma[c] * invstd;
  for (int b = 0; b < B; ++b) {
    const long long o = (long long)b * C + c;
    const float d = y[o] > 0.f ? dy[o] : 0.f;
    dx[o] = ag * (d - k1 - (x[o] - mean) * invstd * k2);
  }
}

__global__ void rsoftmax_fwd_kernel(const float* logits, int B, int R, int C, float* att) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * C) return;
  const int b = (int)(idx / C), c = (int)(idx % C);
  const float* l = logits + (long long)b * R * C + c;
  float mx = -INFINITY;
  for (int r = 0; r < R; ++r) mx = fmaxf(mx, l[r * C]);
  float s = 0.f;
  for (int r = 0; r < R; ++r) s += expf(l[r * C] - mx);
  for (int r = 0; r < R; ++r) att[(long long)b * R * C + r * C + c] = expf(l[r * C] - mx) / s;
}

__global__ void rsoftmax_bwd_kernel(const float* datt, const float* att, int B, int R, int C, float* dlogits) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * C) return;
  const int b = (int)(idx / C), c = (int)(idx % C);
  const long long base = (long long)b * R * C + c;
  float dot = 0.f;
  for (int r = 0; r < R; ++r) dot += datt[base + r * C] * att[base + r * C];
  for (int r = 0; r < R; ++r) dlogits[base + r * C] = att[base + r * C] * (datt[base + r * C] - dot);
}

// value of element idx of the packed operand (see OCT_PACK_* in the header)
__device__ __forceinline__ float pack_value(const float* __restrict__ w, int mode, int cout, int cin, int groups, int dg,
                                            int k, long long idx) {
  const int taps = k * k;
  const int cin_g = cin / groups, cout_g = cout / groups;
  const int cin_d = cin / dg, cout_d = cout / dg;
  float v = 0.f;
  if (mode == OCT_PACK_FWD) {
    const int ci_d = (int)(idx % cin_d);
    const int co = (int)((idx / cin_d) % cout);
    const int tap = (int)(idx / ((long long)cin_d * cout));
    const int ci = (co / cout_d) * cin_d + ci_d;  // global input channel
    if (ci / cin_g == co / cout_g) v = w[((long long)co * cin_g + (ci % cin_g)) * taps + tap];
  } else if (mode == OCT_PACK_DGRAD) {
    const int co_d = (int)(idx % cout_d);
    const int ci = (int)((idx / cout_d) % cin);
    const int tap = (int)(idx / ((long long)cout_d * cin));
    const int co = (ci / cin_d) * cout_d + co_d;
    if (ci / cin_g == co / cout_g) v = w[((long long)co * cin_g + (ci % cin_g)) * taps + (taps - 1 - tap)];
  } else if (mode == OCT_PACK_CONVT_FWD) {
    // out[(t*cout + co)][ci] = w[ci][co][t]
    const int ci = (int)(idx % cin);
    const int co = (int)((idx / cin) % cout);
    const int t = (int)(idx / ((long long)cin * cout));
    v = w[((long long)ci * cout + co) * 4 + t];
  } else {
    // out[ci][t*cout + co] = w[ci][co][t]
    const int co = (int)(idx % cout);
    const int t = (int)((idx / cout) % 4);
    const int ci = (int)(idx / ((long long)4 * cout));
    v = w[((long long)ci * cout + co) * 4 + t];
  }
  return v;
}

__device__ __host__ inline long long pack_total(int mode, int cout, int cin, int dg, int k) {
  const int taps = k * k;
  if (mode == OCT_PACK_FWD) return (long long)taps * cout * (cin / dg);
  if (mode == OCT_PACK_DGRAD) return (long long)taps * cin * (cout / dg);
  return 4LL * cout * cin;
}

__global__ void pack_weight_kernel(const float* w, int mode, int cout, int cin, int groups, int dg, int k, bf16* out) {
  const long long total = pack_total(mode, cout, cin, dg, k);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
    out[idx] = __float2bfloat16_rn(pack_value(w, mode, cout, cin, groups, dg, k, idx));
}

// Every operand pack of a network in ONE launch: block -> job by binary search over the jobs' first block, then
// kItemsPerBlock consecutive ITEMS of that job.  An item is one (output row, K index) pair and carries all taps: a thread
// reads the taps of its weight (contiguous in the torch layout) and writes one element into every tap plane, with the
// lane-fastest item index chosen so that the 2-byte stores of a warp are contiguous.
constexpr int kItemsPerBlock = 1024;
__device__ __host__ inline long long pack_items(int mode, int cout, int cin, int dg) {
  if (mode == OCT_PACK_FWD) return (long long)cout * (cin / dg);
  if (mode == OCT_PACK_DGRAD) return (long long)cin * (cout / dg);
  return (long long)cout * cin;
}

__global__ void __launch_bounds__(256) pack_weight_multi_kernel(const OctavePackJob* __restrict__ jobs, int n_jobs) {
  int lo = 0, hi = n_jobs - 1;
  const long long b = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].block_start <= b) lo = mid; else hi = mid - 1;
  }
  const OctavePackJob j = jobs[lo];
  const float* __restrict__ w = j.w;
  bf16* out = reinterpret_cast<bf16*>(j.out);
  const int cout = j.cout, cin = j.cin, dg = j.dense_groups;
  const int taps = j.mode <= OCT_PACK_DGRAD ? j.ksize * j.ksize : 4;
  const int cin_g = cin / j.groups, cout_g = cout / j.groups;
  const int cin_d = cin / dg, cout_d = cout / dg;
  const long long items = pack_items(j.mode, cout, cin, dg);
  const long long base = (b - j.block_start) * kItemsPerBlock;
  for (int i = threadIdx.x; i < kItemsPerBlock; i += 256) {
    const long long it = base + i;
    if (it >= items) break;
    if (j.mode == OCT_PACK_FWD) {
      const int ci_d = (int)(it % cin_d), co = (int)(it / cin_d);
      const int ci = (co / cout_d) * cin_d + ci_d;
      const bool on = ci / cin_g == co / cout_g;
      const float* src = w + ((long long)co * cin_g + (ci % cin_g)) * taps;
      for (int t = 0; t < taps; ++t)
        out[((long long)t * cout + co) * cin_d + ci_d] = __float2bfloat16_rn(on ? src[t] : 0.f);
    } else if (j.mode == OCT_PACK_DGRAD) {
      const int co_d = (int)(it % cout_d), ci = (int)(it / cout_d);
      const int co = (ci / cin_d) * cout_d + co_d;
      const bool on = ci / cin_g == co / cout_g;
      const float* src = w + ((long long)co * cin_g + (ci % cin_g)) * taps;
      for (int t = 0; t < taps; ++t)
        out[((long long)t * cin + ci) * cout_d + co_d] = __float2bfloat16_rn(on ? src[taps - 1 - t] : 0.f);
    } else if (j.mode == OCT_PACK_CONVT_FWD) {
      const int ci = (int)(it % cin), co = (int)(it / cin);
      const float4 v = *reinterpret_cast<const float4*>(w + ((long long)ci * cout + co) * 4);
      out[((long long)0 * cout + co) * cin + ci] = __float2bfloat16_rn(v.x);
      out[((long long)1 * cout + co) * cin + ci] = __float2bfloat16_rn(v.y);
      out[((long long)2 * cout + co) * cin + ci] = __float2bfloat16_rn(v.z);
      out[((long long)3 * cout + co) * cin + ci] = __float2bfloat16_rn(v.w);
    } else {
      const int co = (int)(it % cout), ci = (int)(it / cout);
      const float4 v = *reinterpret_cast<const float4*>(w + ((long long)ci * cout + co) * 4);
      bf16* o = out + (long long)ci * 4 * cout + co;
      o[0] = __float2bfloat16_rn(v.x);
      o[cout] = __float2bfloat16_rn(v.y);
      o[2 * cout] = __float2bfloat16_rn(v.z);
      o[3 * cout] = __float2bfloat16_rn(v.w);
    }
  }
}

}  // namespace

// shapes the tiled kernels take: 8-column blocks inside one group, 8 K slices of whole float4s, 32-input blocks
static bool glinear_tiled_ok(int Kt, int N, int groups) {
  const int Kg = Kt / groups, Ng = N / groups;
  return Ng % kJB == 0 && Kg % 32 == 0;
}

// fallback (any shape): one warp per output element (b, j), lanes stride over K
__global__ void glinear_fwd_ref_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                       int B, int Kt, int N, int groups, float scale, float* __restrict__ out) {
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= (long long)B * N) return;
  const int b = (int)(wid / N), j = (int)(wid % N);
  const int Kg = Kt / groups, g = j / (N / groups);
  const float* ip = in + (long long)b * Kt + g * Kg;
  const float* wp = w + (long long)j * Kg;
  float acc = 0.f;
#pragma unroll 4
  for (int i = lane; i < Kg; i += 32) acc += __ldg(ip + i) * __ldg(wp + i);
  acc = warp_sum(acc);
  if (lane == 0) out[wid] = acc * scale + (bias ? bias[j] : 0.f);
}

// fallback (any Kg): thread per input i, blockIdx.y = slice of the outputs j, blockIdx.z = batch row; partial sums meet
// in the zeroed din through atomics.  Serves the B = 1 power-iteration matvecs of the critic's spectral norm.
__global__ void glinear_bwd_data_ref_kernel(const float* __restrict__ dout, const float* __restrict__ w, int B, int Kt, int N,
                                            int groups, float scale, float* __restrict__ din) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Kt) return;
  const int b = blockIdx.z;
  const int Kg = Kt / groups, Ng = N / groups, g = i / Kg, il = i - g * Kg;
  const int chunk = (Ng + gridDim.y - 1) / gridDim.y;
  const int j0 = blockIdx.y * chunk, j1 = min(j0 + chunk, Ng);
  const float* dp = dout + (long long)b * N + g * Ng;
  const float* wp = w + (long long)g * Ng * Kg + il;
  float acc = 0.f;
#pragma unroll 8
  for (int j = j0; j < j1; ++j) acc += __ldg(dp + j) * __ldg(wp + (long long)j * Kg);
  atomicAdd(din + (long long)b * Kt + i, acc * scale);
}

extern "C" int octave_glinear_fwd(const float* in, const float* w, const float* bias, int32_t B, int32_t Kt, int32_t N,
                                  int32_t groups, float in_scale, float* out, void* stream) {
  if (!in || !w || !out || B <= 0 || Kt <= 0 || N <= 0 || groups <= 0 || Kt % groups || N % groups) return OCT_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  if (glinear_tiled_ok(Kt, N, groups))
    glinear_fwd_kernel<0><<<dim3(N / kJB, (B + 31) / 32), 256, 0, s>>>(in, w, bias, B, Kt, N, groups, in_scale, out, GlinearEpi{});
  else
    glinear_fwd_ref_kernel<<<(unsigned)(((long long)B * N * 32 + 255) / 256), 256, 0, s>>>(in, w, bias, B, Kt, N, groups, in_scale, out);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_attn_fused_supported(int32_t B, int32_t C, int32_t inter, int32_t groups, int32_t radix) {
  // whole batch in one 32-row slab; one group; radix pairs (c, C + c); the tiled kernels' shape rules for both linears
  return B > 0 && B <= 32 && groups == 1 && radix == 2 && C % 32 == 0 && inter % 32 == 0 && glinear_tiled_ok(C, inter, 1) &&
         glinear_tiled_ok(inter, 2 * C, 1);
}

extern "C" int octave_glinear_bn_relu_fwd(const float* in, const float* w, const float* bias, int32_t B, int32_t Kt, int32_t N,
                                          float in_scale, const float* gamma, const float* beta, float* rm, float* rv, int64_t* nbt,
                                          float eps, float momentum, int32_t training, float* x_out, float* y, float* mean_invstd,
                                          void* stream) {
  if (!in || !w || !gamma || !beta || !x_out || !y || !mean_invstd || B <= 0 || Kt <= 0 || N <= 0) return OCT_ERR_INVALID;
  if (!training && (!rm || !rv)) return OCT_ERR_INVALID;
  if (B > 32 || !glinear_tiled_ok(Kt, N, 1)) return OCT_ERR_UNSUPPORTED;
  GlinearEpi e{gamma, beta, rm, rv, (long long*)nbt, eps, momentum, training, y, mean_invstd, 0};
  glinear_fwd_kernel<1><<<dim3(N / kJB, 1), 256, 0, (cudaStream_t)stream>>>(in, w, bias, B, Kt, N, 1, in_scale, x_out, e);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_glinear_rsoftmax_fwd(const float* in, const float* w, const float* bias, int32_t B, int32_t Kt, int32_t C,
                                           float* att, void* stream) {
  if (!in || !w || !att || B <= 0 || Kt <= 0 || C <= 0) return OCT_ERR_INVALID;
  if (B > 32 || C % 4 || !glinear_tiled_ok(Kt, 2 * C, 1)) return OCT_ERR_UNSUPPORTED;
  GlinearEpi e{};
  e.C = C;
  glinear_fwd_kernel<2><<<dim3(2 * C / kJB, 1), 256, 0, (cudaStream_t)stream>>>(in, w, bias, B, Kt, 2 * C, 1, 1.f, att, e);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_rsoftmax_glinear_bn_bwd(const float* datt, const float* att, const float* w, int32_t B, int32_t Kt, int32_t C,
                                              const float* x, const float* y, const float* gamma, const float* mean_invstd,
                                              int32_t training, float* dlogits, float* dx, float* dgamma, float* dbeta, void* stream) {
  if (!datt || !att || !w || !x || !y || !gamma || !mean_invstd || !dlogits || !dx || B <= 0 || Kt <= 0 || C <= 0) return OCT_ERR_INVALID;
  if (B > 32 || Kt % 32) return OCT_ERR_UNSUPPORTED;
  GlinearBwdFuse f{att, datt, dlogits, C, x, y, gamma, mean_invstd, training, dx, dgamma, dbeta};
  glinear_bwd_data_kernel<true><<<dim3(Kt / 32, 1, 1), 256, 0, (cudaStream_t)stream>>>(nullptr, w, B, Kt, 2 * C, 1, 1.f, nullptr, f);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_glinear_bwd_data(const float* dout, const float* w, int32_t B, int32_t Kt, int32_t N, int32_t groups,
                                       float in_scale, float* din, void* stream) {
  if (!dout || !w || !din || B <= 0 || Kt <= 0 || N <= 0 || groups <= 0 || Kt % groups || N % groups) return OCT_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  if ((Kt / groups) % 32 == 0) {
    // enough blocks to occupy the machine: split the contraction when there are few 32-input column blocks
    const int Ng = N / groups;
    int split = 1;
    while (!g_octave_deterministic && split < 16 && (Kt / 32) * split < 96 && Ng / (split * 2) >= kJC) split *= 2;
    if (split > 1 && cudaMemsetAsync(din, 0, sizeof(float) * (size_t)B * Kt, s) != cudaSuccess) return OCT_ERR_LAUNCH;
    glinear_bwd_data_kernel<false><<<dim3(Kt / 32, (B + 31) / 32, split), 256, 0, s>>>(dout, w, B, Kt, N, groups, in_scale, din, GlinearBwdFuse{});
  }
  else {
    if (B > 65535) return OCT_ERR_UNSUPPORTED;
    if (cudaMemsetAsync(din, 0, sizeof(float) * (size_t)B * Kt, s) != cudaSuccess) return OCT_ERR_LAUNCH;
    const int Ng = N / groups;
    int split = (Ng + 15) / 16;
    if (split > 128) split = 128;
    if (g_octave_deterministic) split = 1;   // one writer per din element
    glinear_bwd_data_ref_kernel<<<dim3((Kt + 127) / 128, split, B), 128, 0, s>>>(dout, w, B, Kt, N, groups, in_scale, din);
  }
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_glinear_bwd_weight(const float* dout, const float* in, int32_t B, int32_t Kt, int32_t N, int32_t groups,
                                         float in_scale, float* dw, float* dbias, void* stream) {
  if (!dout || !in || !dw || B <= 0 || Kt % groups || N % groups) return OCT_ERR_INVALID;
  const long long total = (long long)N * (Kt / groups);
  glinear_bwd_weight_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dout, in, B, Kt, N, groups, in_scale, dw, dbias);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_bn1d_relu_fwd(const float* x, int32_t B, int32_t C, const float* gamma, const float* beta, float* rm,
                                    float* rv, int64_t* nbt, float eps, float momentum, int32_t training, float* y,
                                    float* mean_invstd, void* stream) {
  if (!x || !gamma || !beta || !y || !mean_invstd || B <= 0 || C <= 0) return OCT_ERR_INVALID;
  if (!training && (!rm || !rv)) return OCT_ERR_INVALID;
  bn1d_relu_fwd_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(x, B, C, gamma, beta, rm, rv, (long long*)nbt, eps,
                                                                        momentum, training, y, mean_invstd);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_bn1d_relu_bwd(const float* dy, const float* x, const float* y, int32_t B, int32_t C, const float* gamma,
                                    const float* mean_invstd, int32_t training, float* dx, float* dgamma, float* dbeta,
                                    void* stream) {
  if (!dy || !x || !y || !gamma || !mean_invstd || !dx) return OCT_ERR_INVALID;
  bn1d_relu_bwd_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(dy, x, y, B, C, gamma, mean_invstd, training, dx, dgamma, dbeta);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_rsoftmax_fwd(const float* logits, int32_t B, int32_t R, int32_t C, float* att, void* stream) {
  if (!logits || !att || B <= 0 || R <= 0 || C <= 0) return OCT_ERR_INVALID;
  const long long total = (long long)B * C;
  rsoftmax_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(logits, B, R, C, att);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_rsoftmax_bwd(const float* datt, const float* att, int32_t B, int32_t R, int32_t C, float* dlogits, void* stream) {
  if (!datt || !att || !dlogits) return OCT_ERR_INVALID;
  const long long total = (long long)B * C;
  rsoftmax_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(datt, att, B, R, C, dlogits);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int octave_pack_weight(const float* w, int32_t mode, int32_t cout, int32_t cin, int32_t groups, int32_t dense_groups,
                                  int32_t ksize, void* out, void* stream) {
  if (!w || !out || cout <= 0 || cin <= 0 || groups <= 0 || dense_groups <= 0) return OCT_ERR_INVALID;
  if (mode < 0 || mode > 3) return OCT_ERR_INVALID;
  if (mode <= OCT_PACK_DGRAD && (cin % groups || cout % groups || groups % dense_groups)) return OCT_ERR_INVALID;
  const long long total = pack_total(mode, cout, cin, dense_groups, ksize);
  long long g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  pack_weight_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(w, mode, cout, cin, groups, dense_groups, ksize, (bf16*)out);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

extern "C" int64_t octave_pack_job_blocks(int32_t mode, int32_t cout, int32_t cin, int32_t dense_groups, int32_t ksize) {
  if (mode < 0 || mode > 3 || cout <= 0 || cin <= 0 || dense_groups <= 0 || ksize <= 0) return 0;
  (void)ksize;
  const long long items = pack_items(mode, cout, cin, dense_groups);
  return (items + kItemsPerBlock - 1) / kItemsPerBlock;
}

extern "C" int octave_pack_weight_multi(const OctavePackJob* jobs_device, int32_t n_jobs, int64_t total_blocks, void* stream) {
  if (!jobs_device || n_jobs <= 0 || total_blocks <= 0 || total_blocks > 0x7fffffffLL) return OCT_ERR_INVALID;
  pack_weight_multi_kernel<<<(unsigned)total_blocks, 256, 0, (cudaStream_t)stream>>>(jobs_device, n_jobs);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}
