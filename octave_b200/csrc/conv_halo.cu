// K1b — 3x3 stride-1 convolution for NARROW layers (Cin/group in {32, 64}, Cout/group in {32, 64}) on tcgen05.
//
// The generic implicit-GEMM kernel (conv_tc.cu) fetches one shifted A patch and one weight tile per filter tap: for a
// narrow layer that is 18 small TMA requests and 9x the input bytes over the L2 fabric per output tile, and ncu shows
// the MMA lane starving on the full barrier while HBM idles (profiles/conv_3x3_narrow_r01.txt).  These layers are
// HBM-bound (decoder_0/1, stem, critic: 400^2 and 200^2 maps), so this kernel moves every byte once:
//
//  * ALL filter taps of every group stay resident in shared memory for the life of the persistent CTA
//    (9 * Cout_g * Cin_g * 2 B per group, <= 72 KB) — one TMA burst at start-up.
//  * One TMA box per output tile brings the (TH+2) x TWp input patch INCLUDING its halo (hardware zero-fill outside
//    the image = the conv padding).  The GEMM M index runs over the flattened patch, m = oh * TWp + ow, so the A operand
//    of tap (dh, dw) is the very same shared-memory tile starting (dh * TWp + dw) rows further down: the tap shift is a
//    byte offset in the UMMA shared-memory descriptor (the swizzle is a function of the absolute address, which TMA and
//    UMMA share).  Rows with ow >= TW (two per patch row) are computed but not stored.
//  * 9 * Cin_g/16 MMAs (128 x Cout_g x 16) per tile into one of two TMEM accumulators; the epilogue (bias / activation,
//    bf16, fused BatchNorm statistics, coalesced 16-byte stores, optional accumulate) overlaps the next tile.
//
// Serves forward and data-gradient (the latter through the flipped/transposed weight pack, like conv_tc.cu) of
//   ResNet deep stem, ResNestDecoder / SplAtConv2d at the two finest levels   /root/reference/architectures/extra/resnest.py:18-138,326-334
//   DiscriminatorBlock 4x4 s2 convs in space-to-depth form                    /root/reference/architectures/discriminator/blocks.py:46-50,91-109
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "../../include/octave_b200.h"

namespace {

constexpr int kMaxStages = 8;
constexpr int kHaloThreads = 10 * 32;  // warp 0 TMA producer, warp 1 MMA, warps 2..9 epilogue

struct HaloParams {
  int TWp, TW, TH;     // patch row pitch (TW + 2), stored patch width / height
  int tiles_w, tiles_h;
  int groups;
  int H, W;            // output pixel grid
  void* out;
  long long ldc;
  int c_off, Hout, Wout;
  const float* bias;
  int act, accumulate;
  long long total_tiles;   // B * tiles_h * tiles_w * groups (group fastest)
  double* stats;
  int stats_stride;
  int a_rows;              // rows of one halo box = (TH + 2) * TWp
  int a_stage_bytes;       // shared-memory bytes reserved per stage (1024-aligned, >= (a_rows + 8) rows)
  int stages;
  int base_off_mode;       // 1: also fill the descriptor's matrix-base-offset field with (addr >> 7) & 7
};

__device__ __forceinline__ float halo_act(float v, int act) {
  switch (act) {
    case 1: return fmaxf(v, 0.f);
    case 2: return v > 0.f ? v : 0.2f * v;
    case 3: return 1.f / (1.f + __expf(-v));
    case 4: return tanhf(v);
    default: return v;
  }
}

// EPI = 1 (plain outputs: no bias / activation): TMA-store epilogue as in conv_tc_kernel — the live rows of a tile
// (ow < TW) are staged COMPACTLY in the TMA box layout [TH][TW][BN] (16-byte units XOR-swizzled), one cp.async.bulk.tensor
// store (or reduce-add, for accumulating data gradients) per tile, two groups of four epilogue warps on alternate tiles.
template <int BN, int CIN, int EPI = 0>
__global__ void __launch_bounds__(kHaloThreads, 1) conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                       const __grid_constant__ CUtensorMap tmW,
                                                                       const __grid_constant__ CUtensorMap tmC,
                                                                       const HaloParams p) {
  constexpr int ROWB = CIN * 2;          // bytes per pixel row of the A tile = swizzle span
  constexpr int TAPB = BN * ROWB;        // bytes of one tap's weight tile
  constexpr int ACC_COLS = BN < 32 ? 32 : BN;
  constexpr int TMEM_COLS = 2 * ACC_COLS;
  constexpr int COLS_W = BN / 2;         // columns handled by one epilogue warp
  constexpr int NCHUNK_W = COLS_W / 16;
  constexpr int PITCH = COLS_W * 2 + 16;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t acc_full[2];
  __shared__ __align__(8) uint64_t acc_empty[2];
  __shared__ __align__(8) uint64_t w_bar;
  __shared__ uint32_t tmem_slot;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const smem_gen = smem_raw + (smem_base - tc::smem_u32(smem_raw));
  const uint32_t w_smem = smem_base;
  const uint32_t w_bytes = (uint32_t)p.groups * 9u * TAPB;
  const uint32_t a_smem = smem_base + ((w_bytes + 1023u) & ~1023u);
  uint8_t* const stage_gen = smem_gen + ((w_bytes + 1023u) & ~1023u) + (size_t)p.stages * p.a_stage_bytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      tc::mbar_init(tc::smem_u32(&full_bar[s]), 1);
      tc::mbar_init(tc::smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(tc::smem_u32(&acc_full[b]), 1);
      tc::mbar_init(tc::smem_u32(&acc_empty[b]), EPI == 1 ? 4 : 8);
    }
    tc::mbar_init(tc::smem_u32(&w_bar), 1);
    tc::fence_barrier_init();
    tc::fence_proxy_async();
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmW);
    if (EPI == 1) tc::tma_prefetch_desc(&tmC);
  }
  if (warp == 1) tc::tmem_alloc<TMEM_COLS>(tc::smem_u32(&tmem_slot));
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  if (warp == 0) {
    // ---- TMA producer: warp-uniform loop, one elected lane issues
    const bool leader = tc::elect_one();
    const uint32_t wb = tc::smem_u32(&w_bar);
    if (leader) {
      // resident weights: every tap of every group, once
      tc::mbar_arrive_expect_tx(wb, w_bytes);
      for (int g = 0; g < p.groups; ++g)
        for (int tap = 0; tap < 9; ++tap) tc::tma_load_3d(w_smem + (g * 9 + tap) * TAPB, &tmW, wb, 0, g * BN, tap);
    }
    const uint32_t tx_bytes = (uint32_t)p.a_rows * ROWB;
    int s = 0;
    uint32_t ph = 0;
    for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const int tt = (int)t;
      const int g = tt % p.groups;
      const int m_tile = tt / p.groups;
      const int img = m_tile / tiles_per_img;
      const int trem = m_tile - img * tiles_per_img;
      const int th_i = trem / p.tiles_w;
      const int h0 = th_i * p.TH, w0 = (trem - th_i * p.tiles_w) * p.TW;
      tc::mbar_wait(tc::smem_u32(&empty_bar[s]), ph ^ 1u);
      const uint32_t fb = tc::smem_u32(&full_bar[s]);
      if (leader) {
        tc::mbar_arrive_expect_tx(fb, tx_bytes);
        tc::tma_load_4d(a_smem + s * p.a_stage_bytes, &tmA, fb, g * CIN, w0 - 1, h0 - 1, img);
      }
      if (++s == p.stages) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: warp-uniform loop; descriptors = per-tile base + per-tap constant (only the 14-bit address field moves)
    const bool leader = tc::elect_one();
    constexpr uint32_t idesc = tc::umma_idesc_bf16(128, BN, 0, 0);
    uint32_t tap_off[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) tap_off[tap] = ((uint32_t)((tap / 3) * p.TWp + (tap % 3)) * ROWB) >> 4;
    tc::mbar_wait(tc::smem_u32(&w_bar), 0);
    tc::fence_after_sync();
    int s = 0;
    uint32_t ph = 0, it = 0;
    for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      const int g = (int)t % p.groups;
      const uint32_t buf = it & 1u;
      tc::mbar_wait(tc::smem_u32(&acc_empty[buf]), ((it >> 1) & 1u) ^ 1u);
      tc::mbar_wait(tc::smem_u32(&full_bar[s]), ph);
      tc::fence_after_sync();
      const uint32_t d_tmem = tmem_base + buf * ACC_COLS;
      const uint64_t da0 = tc::umma_smem_desc(a_smem + s * p.a_stage_bytes, ROWB, 16);
      const uint64_t db0 = tc::umma_smem_desc(w_smem + g * 9 * TAPB, ROWB, 16);
      if (leader) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
          for (int k = 0; k < CIN / 16; ++k)
            tc::umma_bf16(d_tmem, da0 + tap_off[tap] + k * 2, db0 + (uint32_t)((tap * TAPB) >> 4) + k * 2, idesc, (tap | k) != 0);
        }
        tc::umma_commit(tc::smem_u32(&empty_bar[s]));
        tc::umma_commit(tc::smem_u32(&acc_full[buf]));
      }
      __syncwarp();
      if (++s == p.stages) { s = 0; ph ^= 1u; }
    }
  } else if (EPI == 1) {
    // ---- TMA-store epilogue: two groups of four warps on alternate tiles (group = accumulator buffer); a warp owns the 32
    // GEMM rows of its TMEM lane quarter over all BN columns
    constexpr int RB = BN * 2;                      // bytes per staged row = swizzle span (128 or 64)
    constexpr int CHS = BN / 16;
    constexpr int CP = BN / 2, RG = 32 / CP > 0 ? 32 / CP : 1, RPG = 32 / RG;
    const int q = warp & 3;
    const int group = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int th = r / p.TWp, tw = r - th * p.TWp;
    const bool live = tw < p.TW && th < p.TH;
    const int rc = th * p.TW + tw;                  // compact row of this lane's pixel inside the staged [TH][TW] box
    uint8_t* const sbuf = stage_gen + group * (128 * RB);
    const uint32_t sbuf_u32 = tc::smem_u32(sbuf);
    const bool issuer = q == 0 && lane == 0;
    const int bar_id = 1 + group;
    auto swz = [](int row) { return RB == 128 ? (row & 7) : ((row >> 1) & 3); };
    float sacc[4] = {0.f, 0.f, 0.f, 0.f};
    int stat_col0 = -1;
    const uint32_t buf = (uint32_t)group;
    uint32_t use = 0;
    for (long long t = blockIdx.x + (long long)group * gridDim.x; t < p.total_tiles; t += 2LL * gridDim.x, ++use) {
      const int tt = (int)t;
      const int g = tt % p.groups;
      const int m_tile = tt / p.groups;
      const int img = m_tile / tiles_per_img;
      const int trem = m_tile - img * tiles_per_img;
      const int th_i = trem / p.tiles_w;
      const int h0 = th_i * p.TH, w0 = (trem - th_i * p.tiles_w) * p.TW;
      const bool valid = live && (h0 + th < p.Hout) && (w0 + tw < p.Wout);
      stat_col0 = g * BN;
      tc::mbar_wait(tc::smem_u32(&acc_full[buf]), use & 1u);
      tc::fence_after_sync();
      const uint32_t taddr = tmem_base + buf * ACC_COLS + ((uint32_t)(q * 32) << 16);
      uint32_t v[CHS][16];
#pragma unroll
      for (int c = 0; c < CHS; ++c) tc::tmem_ld16(taddr + c * 16, v[c]);
      tc::tmem_ld_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tc::smem_u32(&acc_empty[buf]));
      // (A) the group's previous store has finished reading the staging tile
      if (issuer) tc::tma_store_wait_read();
      tc::bar_sync_named(bar_id, 128);
      if (live) {
        uint8_t* rowp = sbuf + rc * RB;
#pragma unroll
        for (int c = 0; c < CHS; ++c) {
          uint4 u0, u1;
          u0.x = bf16x2_pack(__uint_as_float(v[c][0]), __uint_as_float(v[c][1]));   u0.y = bf16x2_pack(__uint_as_float(v[c][2]), __uint_as_float(v[c][3]));
          u0.z = bf16x2_pack(__uint_as_float(v[c][4]), __uint_as_float(v[c][5]));   u0.w = bf16x2_pack(__uint_as_float(v[c][6]), __uint_as_float(v[c][7]));
          u1.x = bf16x2_pack(__uint_as_float(v[c][8]), __uint_as_float(v[c][9]));   u1.y = bf16x2_pack(__uint_as_float(v[c][10]), __uint_as_float(v[c][11]));
          u1.z = bf16x2_pack(__uint_as_float(v[c][12]), __uint_as_float(v[c][13])); u1.w = bf16x2_pack(__uint_as_float(v[c][14]), __uint_as_float(v[c][15]));
          if (!valid) { u0 = make_uint4(0, 0, 0, 0); u1 = u0; }   // outside the image: clipped by the store, kept out of the statistics
          *reinterpret_cast<uint4*>(rowp + (((2 * c) ^ swz(rc)) << 4)) = u0;
          *reinterpret_cast<uint4*>(rowp + (((2 * c + 1) ^ swz(rc)) << 4)) = u1;
        }
      }
      __syncwarp();
      if (p.stats) {
        // lane owns the column pair cp over the warp's rows rg, rg + RG, ...; dead rows (ow >= TW) are skipped
        const int cp = lane % CP, rg = lane / CP;
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
        for (int i = 0; i < RPG; ++i) {
          const int rr = q * 32 + i * RG + rg;
          const int th2 = rr / p.TWp, tw2 = rr - th2 * p.TWp;
          if (tw2 < p.TW && th2 < p.TH) {
            const int row = th2 * p.TW + tw2;
            const uint32_t u = *reinterpret_cast<const uint32_t*>(sbuf + row * RB + (((cp >> 2) ^ swz(row)) << 4) + (cp & 3) * 4);
            float a, b;
            bf16x2_unpack(u, a, b);
            s0 += a; s1 += b; q0 += a * a; q1 += b * b;
          }
        }
#pragma unroll
        for (int off = 16; off >= CP; off >>= 1) {
          s0 += __shfl_xor_sync(0xffffffffu, s0, off); s1 += __shfl_xor_sync(0xffffffffu, s1, off);
          q0 += __shfl_xor_sync(0xffffffffu, q0, off); q1 += __shfl_xor_sync(0xffffffffu, q1, off);
        }
        sacc[0] += s0; sacc[1] += s1; sacc[2] += q0; sacc[3] += q1;
      }
      tc::fence_proxy_async();
      tc::bar_sync_named(bar_id, 128);               // (B) every live row of the tile is staged
      if (issuer) {
        if (p.accumulate) tc::tma_reduce_add_4d(&tmC, sbuf_u32, g * BN, w0, h0, img);
        else tc::tma_store_4d(&tmC, sbuf_u32, g * BN, w0, h0, img);
        tc::tma_store_commit();
      }
    }
    if (issuer) tc::tma_store_wait_all();
    if (p.stats && stat_col0 >= 0 && lane < CP) {
      atomicAdd(p.stats + stat_col0 + 2 * lane, (double)sacc[0]);
      atomicAdd(p.stats + stat_col0 + 2 * lane + 1, (double)sacc[1]);
      atomicAdd(p.stats + p.stats_stride + stat_col0 + 2 * lane, (double)sacc[2]);
      atomicAdd(p.stats + p.stats_stride + stat_col0 + 2 * lane + 1, (double)sacc[3]);
    }
  } else {
    // ---- 8 epilogue warps: TMEM lane quarter q <-> GEMM rows [32q, 32q+32); column half hsel
    const int q = warp & 3;
    const int hsel = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int th = r / p.TWp, tw = r - th * p.TWp;
    uint8_t* const stage = stage_gen + (warp - 2) * (32 * PITCH);
    float sacc[4] = {0.f, 0.f, 0.f, 0.f};   // statistics of this lane's column pair (tile_col_stats)
    const bool plain = p.bias == nullptr && p.act == 0;
    int stat_col0 = -1;
    uint32_t it = 0;
    for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      const uint32_t buf = it & 1u;
      const int tt = (int)t;
      const int g = tt % p.groups;
      const int m_tile = tt / p.groups;
      const int img = m_tile / tiles_per_img;
      const int trem = m_tile - img * tiles_per_img;
      const int th_i = trem / p.tiles_w;
      const int h = th_i * p.TH + th, w = (trem - th_i * p.tiles_w) * p.TW + tw;
      const bool valid = (tw < p.TW) && (th < p.TH) && (h < p.H) && (w < p.W) && (h < p.Hout) && (w < p.Wout);
      const int cbase = g * BN + hsel * COLS_W;
      const long long row_off = valid ? (((long long)img * p.Hout + h) * p.Wout + w) * p.ldc + p.c_off + cbase : -1;
      stat_col0 = cbase;
      tc::mbar_wait(tc::smem_u32(&acc_full[buf]), (it >> 1) & 1u);
      tc::fence_after_sync();
      const uint32_t taddr = tmem_base + buf * ACC_COLS + hsel * COLS_W + ((uint32_t)(q * 32) << 16);
      // all TMEM loads in flight, one wait, accumulator handed back before the bias / activation / packing work
      uint32_t v[NCHUNK_W][16];
#pragma unroll
      for (int c = 0; c < NCHUNK_W; ++c) tc::tmem_ld16(taddr + c * 16, v[c]);
      tc::tmem_ld_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tc::smem_u32(&acc_empty[buf]));
      if (plain && __all_sync(0xffffffffu, valid)) {
        // common case (conv -> BatchNorm, interior patch): straight pack, no per-element predicated bias loads / selects
#pragma unroll
        for (int c = 0; c < NCHUNK_W; ++c) {
          uint4 u0, u1;
          u0.x = bf16x2_pack(__uint_as_float(v[c][0]), __uint_as_float(v[c][1]));   u0.y = bf16x2_pack(__uint_as_float(v[c][2]), __uint_as_float(v[c][3]));
          u0.z = bf16x2_pack(__uint_as_float(v[c][4]), __uint_as_float(v[c][5]));   u0.w = bf16x2_pack(__uint_as_float(v[c][6]), __uint_as_float(v[c][7]));
          u1.x = bf16x2_pack(__uint_as_float(v[c][8]), __uint_as_float(v[c][9]));   u1.y = bf16x2_pack(__uint_as_float(v[c][10]), __uint_as_float(v[c][11]));
          u1.z = bf16x2_pack(__uint_as_float(v[c][12]), __uint_as_float(v[c][13])); u1.w = bf16x2_pack(__uint_as_float(v[c][14]), __uint_as_float(v[c][15]));
          *reinterpret_cast<uint4*>(stage + lane * PITCH + c * 32) = u0;
          *reinterpret_cast<uint4*>(stage + lane * PITCH + c * 32 + 16) = u1;
        }
      } else {
#pragma unroll
        for (int c = 0; c < NCHUNK_W; ++c) {
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            f[i] = __uint_as_float(v[c][i]);
            if (p.bias) f[i] += __ldg(p.bias + cbase + c * 16 + i);
          }
          if (p.act) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = halo_act(f[i], p.act);
          }
          if (!valid) {   // rows that are not stored stage zeros: they must not enter the statistics (and may be NaN junk)
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = 0.f;
          }
          uint4 u0, u1;
          u0.x = bf16x2_pack(f[0], f[1]); u0.y = bf16x2_pack(f[2], f[3]);
          u0.z = bf16x2_pack(f[4], f[5]); u0.w = bf16x2_pack(f[6], f[7]);
          u1.x = bf16x2_pack(f[8], f[9]); u1.y = bf16x2_pack(f[10], f[11]);
          u1.z = bf16x2_pack(f[12], f[13]); u1.w = bf16x2_pack(f[14], f[15]);
          *reinterpret_cast<uint4*>(stage + lane * PITCH + c * 32) = u0;
          *reinterpret_cast<uint4*>(stage + lane * PITCH + c * 32 + 16) = u1;
        }
      }
      __syncwarp();   // staged rows are read by other lanes below
      if (p.stats) tile_col_stats<COLS_W, PITCH>(stage, lane, sacc);
      constexpr int LPR = (COLS_W * 2) / 16;   // lanes per row (16 B each): 2 or 4
      constexpr int RPI = 32 / LPR;
      constexpr int NIT = 32 / RPI;
      const int sub = lane % LPR, rsel = lane / LPR;
      if (p.accumulate) {
        // y += result: all old values are fetched first (independent loads in flight), then added and stored
        uint4 old[NIT];
        long long offs[NIT];
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
          const int row = i * RPI + rsel;
          offs[i] = __shfl_sync(0xffffffffu, row_off, row);
          if (offs[i] >= 0) old[i] = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.out) + offs[i] + sub * 8);
        }
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
          if (offs[i] >= 0) {
            const int row = i * RPI + rsel;
            uint4 val = *reinterpret_cast<const uint4*>(stage + row * PITCH + sub * 16);
            const uint4 e = old[i];
            float x0, x1, y0, y1;
            bf16x2_unpack(val.x, x0, x1); bf16x2_unpack(e.x, y0, y1); val.x = bf16x2_pack(x0 + y0, x1 + y1);
            bf16x2_unpack(val.y, x0, x1); bf16x2_unpack(e.y, y0, y1); val.y = bf16x2_pack(x0 + y0, x1 + y1);
            bf16x2_unpack(val.z, x0, x1); bf16x2_unpack(e.z, y0, y1); val.z = bf16x2_pack(x0 + y0, x1 + y1);
            bf16x2_unpack(val.w, x0, x1); bf16x2_unpack(e.w, y0, y1); val.w = bf16x2_pack(x0 + y0, x1 + y1);
            *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + offs[i] + sub * 8) = val;
          }
        }
      } else {
#pragma unroll
        for (int r0 = 0; r0 < 32; r0 += RPI) {
          const int row = r0 + rsel;
          const long long off = __shfl_sync(0xffffffffu, row_off, row);
          if (off >= 0) {
            const uint4 val = *reinterpret_cast<const uint4*>(stage + row * PITCH + sub * 16);
            *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + off + sub * 8) = val;
          }
        }
      }
      __syncwarp();
    }
    if (p.stats && stat_col0 >= 0 && lane < COLS_W / 2) {
      atomicAdd(p.stats + stat_col0 + 2 * lane, (double)sacc[0]);
      atomicAdd(p.stats + stat_col0 + 2 * lane + 1, (double)sacc[1]);
      atomicAdd(p.stats + p.stats_stride + stat_col0 + 2 * lane, (double)sacc[2]);
      atomicAdd(p.stats + p.stats_stride + stat_col0 + 2 * lane + 1, (double)sacc[3]);
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------
// Weight gradient of the same narrow 3x3 layers.  dW[tap][co][ci] = sum_pixels dy[p][co] * x[p + tap][ci]: the GEMM K
// dimension is the pixel index, so with the flattened-patch indexing of the forward kernel (row i = oh * 32 + ow of a
// 4 x 30 output patch, pitch 32) BOTH operands are MN-major views of two TMA boxes — the dy patch (128 rows) and the x
// patch with halo (192 rows) — and the nine taps are nine row offsets into the SAME x tile.  One CTA owns a contiguous
// range of patches and keeps all nine tap accumulators in TMEM (9 x 32 columns): per patch it loads dy once and x once
// (the generic kernel loads both nine times, from nine CTAs) and issues 9 taps x 8 K-steps of 128 x 32 x 16 MMAs.
// The two junk columns of every patch row (ow = 30, 31: real pixels of the neighbouring patch) are zeroed in the dy
// tile by the MMA warp before it issues, so they contribute nothing.  The side with 32 channels is the N operand; the
// other side (32 or 64 channels) is the M operand, padded to 128 rows by aliasing (those accumulator rows are not read).
// ---------------------------------------------------------------------------------------------------
struct HaloWgradParams {
  int tiles_w, tiles_h, B;
  int groups;                 // dense groups (blockIdx.y)
  int cin_g, cout_g;          // per dense group
  int real_cin_g, real_cout_g;
  int x_is_n;                 // 1: x (Cin_g = 32) is the N operand, dy the M operand; 0: dy (Cout_g = 32) is N, x is M
  int total_tiles;            // B * tiles_h * tiles_w
  int tiles_per_cta;
  float* dw;                  // [Cout][Cin/real_groups][3][3], pre-zeroed
};

constexpr int kHwThreads = 6 * 32;   // warp 0 TMA producer, warp 1 MMA (+ junk-row zeroing), warps 2..5 epilogue
constexpr int kHwTWp = 32, kHwTW = 30, kHwTH = 4;

template <int CY, int CX, int STAGES>    // channels per group of dy / x (32 or 64)
__global__ void __launch_bounds__(kHwThreads, 1) conv3x3_halo_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY,
                                                                           const __grid_constant__ CUtensorMap tmX,
                                                                           const HaloWgradParams p) {
  constexpr int ROWY = CY * 2, ROWX = CX * 2;                 // bytes per pixel row = swizzle span
  constexpr int Y_ROWS = kHwTH * kHwTWp;                      // 128
  constexpr int X_ROWS = (kHwTH + 2) * kHwTWp;                // 192
  constexpr int Y_BYTES = Y_ROWS * ROWY;
  constexpr int X_BYTES = ((X_ROWS + 8) * ROWX + 1023) & ~1023;  // + rows read past the box by the last taps
  constexpr int STAGE_BYTES = Y_BYTES + X_BYTES;
  constexpr int TMEM_COLS = 512;                              // 9 taps x 32 columns
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t acc_bar;
  __shared__ uint32_t tmem_slot;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const smem_gen = smem_raw + (smem_base - tc::smem_u32(smem_raw));
  const int g = blockIdx.y;
  const int t_begin = blockIdx.x * p.tiles_per_cta;
  const int t_end = min(t_begin + p.tiles_per_cta, p.total_tiles);
  const int n_iter = t_end - t_begin;

  // rows [X_ROWS, X_ROWS + 8) of every x stage are read (against zeroed dy rows) but never written by TMA: keep them zero
  for (int st = 0; st < STAGES; ++st) {
    uint8_t* q = smem_gen + st * STAGE_BYTES + Y_BYTES + X_ROWS * ROWX;
    for (int i = threadIdx.x * 16; i < 8 * ROWX; i += kHwThreads * 16) *reinterpret_cast<uint4*>(q + i) = make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(tc::smem_u32(&full_bar[s]), 1);
      tc::mbar_init(tc::smem_u32(&empty_bar[s]), 1);
    }
    tc::mbar_init(tc::smem_u32(&acc_bar), 1);
    tc::fence_barrier_init();
    tc::tma_prefetch_desc(&tmDY);
    tc::tma_prefetch_desc(&tmX);
  }
  tc::fence_proxy_async();
  if (warp == 1) tc::tmem_alloc<TMEM_COLS>(tc::smem_u32(&tmem_slot));
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  if (n_iter <= 0) {  // uniform across the CTA
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<TMEM_COLS>(tmem_base);
    return;
  }
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  if (warp == 0) {
    const bool leader = tc::elect_one();
    int s = 0;
    uint32_t ph = 0;
    int img = t_begin / tiles_per_img;
    int trem = t_begin - img * tiles_per_img;
    for (int it = 0; it < n_iter; ++it) {
      tc::mbar_wait(tc::smem_u32(&empty_bar[s]), ph ^ 1u);
      const uint32_t fb = tc::smem_u32(&full_bar[s]);
      const int th_i = trem / p.tiles_w;
      const int h0 = th_i * kHwTH, w0 = (trem - th_i * p.tiles_w) * kHwTW;
      const uint32_t sa = smem_base + s * STAGE_BYTES;
      if (leader) {
        tc::mbar_arrive_expect_tx(fb, (uint32_t)(Y_ROWS * ROWY + X_ROWS * ROWX));
        tc::tma_load_4d(sa, &tmDY, fb, g * CY, w0, h0, img);
        tc::tma_load_4d(sa + Y_BYTES, &tmX, fb, g * CX, w0 - 1, h0 - 1, img);
      }
      if (++trem == tiles_per_img) { trem = 0; ++img; }
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    const bool leader = tc::elect_one();
    constexpr uint32_t idesc = tc::umma_idesc_bf16(128, 32, 1, 1);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < n_iter; ++it) {
      tc::mbar_wait(tc::smem_u32(&full_bar[s]), ph);
      // zero the junk rows of the dy tile (ow = 30, 31 of each of the 4 patch rows): 8 rows x ROWY bytes
      {
        uint8_t* yb = smem_gen + s * STAGE_BYTES;
        constexpr int V = ROWY / 16;                       // uint4 per row
        for (int e = lane; e < 8 * V; e += 32) {
          const int jr = e / V, v = e - jr * V;
          const int row = (jr >> 1) * kHwTWp + kHwTW + (jr & 1);
          *reinterpret_cast<uint4*>(yb + row * ROWY + v * 16) = make_uint4(0, 0, 0, 0);
        }
      }
      tc::fence_proxy_async();
      __syncwarp();
      tc::fence_after_sync();
      const uint32_t sy = smem_base + s * STAGE_BYTES, sx = sy + Y_BYTES;
      // MN-major operands: LBO (distance between 64-element MN chunks) = 0 aliases the padding chunks onto the real one
      const uint64_t dY = tc::umma_smem_desc(sy, ROWY, 0);
      const uint64_t dX = tc::umma_smem_desc(sx, ROWX, 0);
      if (leader) {
        if (p.x_is_n) {
          // x is the N operand (32 channels per pixel row, one 64-byte MN run): the three taps of a filter row are three
          // MN runs ONE pixel row (64 B) apart — leading-dimension byte offset = ROWX — so one 128 x 96 x 16 MMA covers
          // dw = 0, 1, 2 at once: 24 MMAs per patch instead of 72 (a small-N tcgen05.mma costs ~64 cycles regardless of N)
          constexpr uint32_t idesc3 = tc::umma_idesc_bf16(128, 96, 1, 1);
          const uint64_t dX3 = tc::umma_smem_desc(sx, ROWX, ROWX);
#pragma unroll
          for (int dh = 0; dh < 3; ++dh) {
            const uint32_t xoff = (uint32_t)(dh * kHwTWp * ROWX) >> 4;
#pragma unroll
            for (int k = 0; k < Y_ROWS / 16; ++k)
              tc::umma_bf16(tmem_base + dh * 96, dY + (uint32_t)(k * ROWY), dX3 + xoff + (uint32_t)(k * ROWX), idesc3, (it | k) != 0);
          }
        } else {
          // x is the M operand (64 channels = one 128-byte MN run per pixel row; M = 128 takes two runs): two taps per MMA,
          // the second run being the first one shifted by the tap distance (leading-dimension byte offset).  Groups:
          // (dh,0)+(dh,1) for dh = 0..2 [one row apart], (0,2)+(1,2) [one patch row apart], (2,2) alone [aliased].
#pragma unroll
          for (int gi = 0; gi < 5; ++gi) {
            const int first = gi < 3 ? gi * 3 : (gi == 3 ? 2 : 8);
            const uint32_t lbo = gi < 3 ? (uint32_t)ROWX : (gi == 3 ? (uint32_t)(kHwTWp * ROWX) : 0u);
            const uint64_t dXg = tc::umma_smem_desc(sx, ROWX, lbo);
            const uint32_t xoff = (uint32_t)(((first / 3) * kHwTWp + (first % 3)) * ROWX) >> 4;
#pragma unroll
            for (int k = 0; k < Y_ROWS / 16; ++k)
              tc::umma_bf16(tmem_base + gi * 32, dXg + xoff + (uint32_t)(k * ROWX), dY + (uint32_t)(k * ROWY), idesc, (it | k) != 0);
          }
        }
        tc::umma_commit(tc::smem_u32(&empty_bar[s]));
      }
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }
    if (leader) tc::umma_commit(tc::smem_u32(&acc_bar));
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;   // accumulator row = channel of the M operand
    const int m_ch = p.x_is_n ? p.cout_g : p.cin_g;
    tc::mbar_wait(tc::smem_u32(&acc_bar), 0);
    tc::fence_after_sync();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    if (p.x_is_n) {
      // accumulator column = tap * 32 + ci (dh * 96 + dw * 32 + ci); row = output channel
#pragma unroll 1
      for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
        for (int c0 = 0; c0 < 32; c0 += 16) {
          uint32_t v[16];
          tc::tmem_ld16(taddr + tap * 32 + c0, v);
          tc::tmem_ld_wait();
          if (r < m_ch) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int co = g * p.cout_g + r, ci = g * p.cin_g + c0 + i;
              if (ci / p.real_cin_g == co / p.real_cout_g)
                atomicAdd(p.dw + ((long long)co * p.real_cin_g + (ci % p.real_cin_g)) * 9 + tap, __uint_as_float(v[i]));
            }
          }
        }
      }
    } else {
      // five tap groups of 32 columns (output channels); rows 0..63 = first tap of the group, 64..127 = second tap
#pragma unroll 1
      for (int gi = 0; gi < 5; ++gi) {
        const int first = gi < 3 ? gi * 3 : (gi == 3 ? 2 : 8);
        const int second = gi < 3 ? first + 1 : (gi == 3 ? 5 : -1);
        const int tap = r < 64 ? first : second;
        const int ci = g * p.cin_g + (r & 63);
#pragma unroll
        for (int c0 = 0; c0 < 32; c0 += 16) {
          uint32_t v[16];
          tc::tmem_ld16(taddr + gi * 32 + c0, v);
          tc::tmem_ld_wait();
          if (tap >= 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int co = g * p.cout_g + c0 + i;
              if (ci / p.real_cin_g == co / p.real_cout_g)
                atomicAdd(p.dw + ((long long)co * p.real_cin_g + (ci % p.real_cin_g)) * 9 + tap, __uint_as_float(v[i]));
            }
          }
        }
      }
    }
    tc::fence_before_sync();
  }
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

PFN_cuTensorMapEncodeTiled_v12000 halo_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

CUtensorMapSwizzle swz_of(int bytes) {
  return bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

template <int BN, int CIN, int EPI = 0>
int launch_halo(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmC, HaloParams& p, int grid, cudaStream_t s) {
  constexpr int ROWB = CIN * 2;
  constexpr int PITCH = (BN / 2) * 2 + 16;
  const int w_bytes = ((p.groups * 9 * BN * ROWB) + 1023) & ~1023;
  p.a_stage_bytes = (((p.a_rows + 8) * ROWB) + 1023) & ~1023;
  const int fixed = w_bytes + (EPI == 1 ? 2 * 128 * BN * 2 : 8 * 32 * PITCH) + 1024;   // EPI 1: one staged tile per epilogue group
  int stages = (225 * 1024 - fixed) / p.a_stage_bytes;   // static barriers + alignment slack stay below 227 KB
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return OCT_ERR_UNSUPPORTED;
  p.stages = stages;
  const int smem = fixed + stages * p.a_stage_bytes;
  static unsigned long long attr_devs = 0;   // the attribute is per device: one bit per device ordinal
  int dev__ = 0;
  cudaGetDevice(&dev__);
  if (!(attr_devs >> (dev__ & 63) & 1ull)) {
    cudaError_t ea = cudaFuncSetAttribute(conv3x3_halo_kernel<BN, CIN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (ea != cudaSuccess) {
      if (getenv("OCTAVE_DEBUG")) fprintf(stderr, "[octave] halo: smem attribute: %s\n", cudaGetErrorString(ea));
      return OCT_ERR_LAUNCH;
    }
    attr_devs |= 1ull << (dev__ & 63);
  }
  conv3x3_halo_kernel<BN, CIN, EPI><<<grid, kHaloThreads, smem, s>>>(tmA, tmW, tmC, p);
  if (getenv("OCTAVE_DEBUG")) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) fprintf(stderr, "[octave] halo launch<%d,%d> grid %d smem %d stages %d: %s\n", BN, CIN, grid, smem, stages, cudaGetErrorString(e));
  }
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

}  // namespace

static int g_halo_enabled = -1, g_halo_baseoff = -1;

extern "C" void octave_conv_halo_config(int32_t enabled, int32_t base_off_mode) {
  g_halo_enabled = enabled ? 1 : 0;
  g_halo_baseoff = base_off_mode;
}

// 1 when the narrow-layer kernel takes this descriptor (see the header comment for the shape class)
extern "C" int octave_conv_halo_supported(const OctaveConvDesc* d) {
  if (g_halo_enabled < 0) {
    const char* e = getenv("OCTAVE_HALO");
    g_halo_enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (!g_halo_enabled || !d) return 0;
  if (d->ksize != 3 || d->mode != OCT_CONV_MODE_CONV || d->out_s2d_qs > 0) return 0;
  if (d->groups <= 0 || d->cin % d->groups || d->cout % d->groups) return 0;
  const int cin_g = d->cin / d->groups, cout_g = d->cout / d->groups;
  if (cin_g != 32 && cin_g != 64) return 0;
  if (cout_g != 32 && cout_g != 64) return 0;
  if ((long long)d->groups * 9 * cout_g * cin_g * 2 > 72 * 1024) return 0;
  if (d->x_ld % 8 || d->x_coff % 8 || d->y_ld % 8 || d->y_coff % 8) return 0;
  if (d->out_dtype == OCT_DTYPE_F32) return 0;
  if (d->Hout > d->H || d->Wout > d->W) return 0;
  return 1;
}

extern "C" int octave_conv_halo_fwd(const OctaveConvDesc* d, const void* x, const void* wpack, const float* bias, void* y,
                                    double* stats, void* stream) {
  if (!octave_conv_halo_supported(d)) return OCT_ERR_UNSUPPORTED;
  if (!x || !wpack || !y) return OCT_ERR_INVALID;
  auto enc = halo_encode();
  if (!enc) return OCT_ERR_LAUNCH;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int cin_g = d->cin / d->groups, cout_g = d->cout / d->groups;
  HaloParams p{};
  // patch pitch: the candidate that wastes the fewest of the 128 GEMM rows (ties: the smaller halo overhead)
  double best = -1.0;
  const int cand[4] = {32, 64, 16, 128};
  for (int i = 0; i < 4; ++i) {
    const int twp = cand[i], tw = twp - 2, th = (128 - tw) / twp + 1;
    const long long tiles = (long long)((d->Wout + tw - 1) / tw) * ((d->Hout + th - 1) / th);
    const double eff = (double)d->Hout * d->Wout / ((double)tiles * 128.0);
    if (eff > best + 0.02) { best = eff; p.TWp = twp; p.TW = tw; p.TH = th; }
  }
  p.tiles_w = (d->Wout + p.TW - 1) / p.TW;
  p.tiles_h = (d->Hout + p.TH - 1) / p.TH;
  p.groups = d->groups;
  p.H = d->Hout; p.W = d->Wout;
  p.out = y; p.ldc = d->y_ld; p.c_off = d->y_coff; p.Hout = d->Hout; p.Wout = d->Wout;
  p.bias = bias; p.act = d->relu; p.accumulate = d->accumulate;
  p.total_tiles = (long long)p.tiles_w * p.tiles_h * d->B * d->groups;
  p.stats = stats; p.stats_stride = d->cout;
  p.a_rows = (p.TH + 2) * p.TWp;
  if (g_halo_baseoff < 0) {
    const char* e = getenv("OCTAVE_HALO_BASEOFF");
    g_halo_baseoff = e ? atoi(e) : 0;
  }
  p.base_off_mode = g_halo_baseoff;
  const int rowb = cin_g * 2;
  CUtensorMap tmA, tmW;
  {
    const bf16* xb = reinterpret_cast<const bf16*>(x) + d->x_coff;
    cuuint64_t dims[4] = {(cuuint64_t)d->cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->x_ld * 2, (cuuint64_t)d->x_ld * 2 * d->W, (cuuint64_t)d->x_ld * 2 * d->W * d->H};
    cuuint32_t box[4] = {(cuuint32_t)cin_g, (cuuint32_t)p.TWp, (cuuint32_t)(p.TH + 2), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(xb), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swz_of(rowb), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      if (getenv("OCTAVE_DEBUG")) fprintf(stderr, "[octave] halo: A tensor map rejected (box %d x %d x %d)\n", cin_g, p.TWp, p.TH + 2);
      return OCT_ERR_LAUNCH;
    }
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)cin_g, (cuuint64_t)d->cout, 9};
    cuuint64_t strides[2] = {(cuuint64_t)cin_g * 2, (cuuint64_t)cin_g * 2 * d->cout};
    cuuint32_t box[3] = {(cuuint32_t)cin_g, (cuuint32_t)cout_g, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(wpack), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swz_of(rowb), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      if (getenv("OCTAVE_DEBUG")) fprintf(stderr, "[octave] halo: W tensor map rejected\n");
      return OCT_ERR_LAUNCH;
    }
  }
  int sms = octave_sm_count();
  if (sms <= 0) sms = 148;
  long long grid = sms - sms % d->groups;   // a CTA keeps one group: its statistics columns never change
  if (grid > p.total_tiles) grid = p.total_tiles;
  if (stats && !g_octave_stats_prezeroed && cudaMemsetAsync(stats, 0, sizeof(double) * 2 * d->cout, s) != cudaSuccess) return OCT_ERR_LAUNCH;
  // OCTAVE_HALO_TMA_EPI=1: plain outputs take the TMA-store epilogue.  Off by default: measured 5-25 % SLOWER than the
  // per-lane stores on the five halo shapes of the c2 step (the larger staging tile costs an input stage and the
  // K = 9 * cin_g mainloop is too short to hide a four-warp epilogue); kept as the measured alternative.
  static const int tma_epi = [] { const char* e = getenv("OCTAVE_HALO_TMA_EPI"); return e ? atoi(e) : 0; }();
  const bool epi1 = tma_epi && !bias && d->relu == 0;
  CUtensorMap tmC = tmA;
  if (epi1) {
    const bf16* yb = reinterpret_cast<const bf16*>(y) + d->y_coff;
    cuuint64_t dims[4] = {(cuuint64_t)d->cout, (cuuint64_t)d->Wout, (cuuint64_t)d->Hout, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->y_ld * 2, (cuuint64_t)d->y_ld * 2 * d->Wout, (cuuint64_t)d->y_ld * 2 * d->Wout * d->Hout};
    cuuint32_t box[4] = {(cuuint32_t)cout_g, (cuuint32_t)p.TW, (cuuint32_t)p.TH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(yb), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            swz_of(cout_g * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return OCT_ERR_LAUNCH;
  }
  if (cin_g == 64) {
    if (cout_g == 64) return epi1 ? launch_halo<64, 64, 1>(tmA, tmW, tmC, p, (int)grid, s) : launch_halo<64, 64>(tmA, tmW, tmC, p, (int)grid, s);
    return epi1 ? launch_halo<32, 64, 1>(tmA, tmW, tmC, p, (int)grid, s) : launch_halo<32, 64>(tmA, tmW, tmC, p, (int)grid, s);
  }
  if (cout_g == 64) return epi1 ? launch_halo<64, 32, 1>(tmA, tmW, tmC, p, (int)grid, s) : launch_halo<64, 32>(tmA, tmW, tmC, p, (int)grid, s);
  return epi1 ? launch_halo<32, 32, 1>(tmA, tmW, tmC, p, (int)grid, s) : launch_halo<32, 32>(tmA, tmW, tmC, p, (int)grid, s);
}

template <int CY, int CX, int STAGES>
static int launch_halo_wgrad(const CUtensorMap& tmDY, const CUtensorMap& tmX, const HaloWgradParams& p, dim3 grid, cudaStream_t s) {
  constexpr int ROWY = CY * 2, ROWX = CX * 2;
  constexpr int stage = 128 * ROWY + (((192 + 8) * ROWX + 1023) & ~1023);
  constexpr int smem = STAGES * stage + 1024;
  static_assert(smem <= 226 * 1024, "shared memory budget");
  static unsigned long long attr_devs = 0;   // the attribute is per device: one bit per device ordinal
  int dev__ = 0;
  cudaGetDevice(&dev__);
  if (!(attr_devs >> (dev__ & 63) & 1ull)) {
    if (cudaFuncSetAttribute(conv3x3_halo_wgrad_kernel<CY, CX, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return OCT_ERR_LAUNCH;
    attr_devs |= 1ull << (dev__ & 63);
  }
  conv3x3_halo_wgrad_kernel<CY, CX, STAGES><<<grid, kHwThreads, smem, s>>>(tmDY, tmX, p);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

// 1 when the halo weight-gradient kernel takes this descriptor: 3x3 s1 p1, channels per (dense) group in {32, 64} with
// at least one side 32, not accumulate
extern "C" int octave_conv_halo_wgrad_supported(const OctaveConvDesc* d) {
  if (g_halo_enabled < 0) {
    const char* e = getenv("OCTAVE_HALO");
    g_halo_enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (!g_halo_enabled || !d) return 0;
  // Measured on B200 (tools/probe_halo.py, B=32).  With nine separate taps this kernel cut the L2->SM traffic ~5x but
  // not the time (774 vs 751 us on 400^2 64->32): a small-N tcgen05.mma costs ~64 cycles whatever its operands, and both
  // kernels issued 72 of them per 128-pixel patch.  Pairing taps inside one MMA (24 / 40 MMAs per patch) is what pays:
  // 400^2 32->64 790 -> 329 us, 200^2 32->64 208 -> 108 us, 200^2 64->128 g2 401 -> 171 us.  OCTAVE_HALO_WGRAD=0 disables.
  static int wg_on = -1;
  if (wg_on < 0) {
    const char* e = getenv("OCTAVE_HALO_WGRAD");
    wg_on = e ? atoi(e) : 1;       // 0: always the generic weight-gradient kernel
  }
  if (!wg_on) return 0;
  if (d->ksize != 3 || d->mode != OCT_CONV_MODE_CONV || d->accumulate) return 0;
  if (d->groups <= 0 || d->cin % d->groups || d->cout % d->groups) return 0;
  const int cin_g = d->cin / d->groups, cout_g = d->cout / d->groups;
  if ((cin_g != 32 && cin_g != 64) || (cout_g != 32 && cout_g != 64)) return 0;
  if (cin_g != 32 && cout_g != 32) return 0;
  if (d->x_ld % 8 || d->x_coff % 8 || d->y_ld % 8 || d->y_coff % 8) return 0;
  if (d->Hout != d->H || d->Wout != d->W) return 0;
  if ((long long)d->B * d->H * d->W < 4096) return 0;   // tiny maps: the generic split-K kernel is fine
  return 1;
}

extern "C" int octave_conv_halo_wgrad(const OctaveConvDesc* d, const void* x, const void* dy, float* dw, void* stream) {
  if (!octave_conv_halo_wgrad_supported(d)) return OCT_ERR_UNSUPPORTED;
  if (!x || !dy || !dw) return OCT_ERR_INVALID;
  auto enc = halo_encode();
  if (!enc) return OCT_ERR_LAUNCH;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int cin_g = d->cin / d->groups, cout_g = d->cout / d->groups;
  const int real_groups = d->real_groups > 0 ? d->real_groups : d->groups;
  HaloWgradParams p{};
  p.tiles_w = (d->W + kHwTW - 1) / kHwTW;
  p.tiles_h = (d->H + kHwTH - 1) / kHwTH;
  p.B = d->B;
  p.groups = d->groups;
  p.cin_g = cin_g; p.cout_g = cout_g;
  p.real_cin_g = d->cin / real_groups; p.real_cout_g = d->cout / real_groups;
  p.x_is_n = cin_g == 32 ? 1 : 0;
  p.total_tiles = d->B * p.tiles_w * p.tiles_h;
  int sms = octave_sm_count();
  if (sms <= 0) sms = 148;
  int ctas = sms / d->groups;
  if (ctas < 1) ctas = 1;
  if (ctas > p.total_tiles) ctas = p.total_tiles;
  p.tiles_per_cta = (p.total_tiles + ctas - 1) / ctas;
  ctas = (p.total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  p.dw = dw;
  CUtensorMap tmDY, tmX;
  {
    const bf16* yb = reinterpret_cast<const bf16*>(dy) + d->y_coff;
    cuuint64_t dims[4] = {(cuuint64_t)d->cout, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->y_ld * 2, (cuuint64_t)d->y_ld * 2 * d->W, (cuuint64_t)d->y_ld * 2 * d->W * d->H};
    cuuint32_t box[4] = {(cuuint32_t)cout_g, (cuuint32_t)kHwTWp, (cuuint32_t)kHwTH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&tmDY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(yb), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swz_of(cout_g * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return OCT_ERR_LAUNCH;
  }
  {
    const bf16* xb = reinterpret_cast<const bf16*>(x) + d->x_coff;
    cuuint64_t dims[4] = {(cuuint64_t)d->cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->x_ld * 2, (cuuint64_t)d->x_ld * 2 * d->W, (cuuint64_t)d->x_ld * 2 * d->W * d->H};
    cuuint32_t box[4] = {(cuuint32_t)cin_g, (cuuint32_t)kHwTWp, (cuuint32_t)(kHwTH + 2), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(xb), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swz_of(cin_g * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return OCT_ERR_LAUNCH;
  }
  const size_t wbytes = (size_t)d->cout * p.real_cin_g * 9 * sizeof(float);
  if (cudaMemsetAsync(dw, 0, wbytes, s) != cudaSuccess) return OCT_ERR_LAUNCH;
  dim3 grid(ctas, d->groups);
  if (cout_g == 64) return launch_halo_wgrad<64, 32, 5>(tmDY, tmX, p, grid, s);
  if (cin_g == 64) return launch_halo_wgrad<32, 64, 5>(tmDY, tmX, p, grid, s);
  return launch_halo_wgrad<32, 32, 6>(tmDY, tmX, p, grid, s);
}
