// K1/K2/K3 — implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), bf16 in,
// fp32 accumulate.  Replaces the stride-1 3x3 / 1x1 Conv2d and the k2s2 ConvTranspose2d of
//   ResNestDecoder / SplAtConv2d / Bottleneck / Upsampling   /root/reference/architectures/extra/resnest.py:18-138,170-267
// (forward, data-gradient and weight-gradient).
//
// Formulation.  Activations are NHWC bf16.  The GEMM M dimension runs over output pixels: one CTA owns a
// TW x TH spatial patch of one image (TW*TH <= 128 rows of a 128-row UMMA tile).  For each filter tap the A
// operand is the same patch shifted by the tap offset, fetched by ONE 4-D TMA box load (channels, w, h, n)
// whose out-of-bounds elements are zero-filled by the hardware => padding costs nothing and there is no
// im2col buffer.  A box lands in shared memory as [pixel][BK channels] rows of 128 B (64 B for BK=32) with
// the matching TMA/UMMA swizzle, i.e. exactly the canonical K-major UMMA operand.  Weights are pre-packed
// [tap][Cout][Cin_g] (K-major B operand) and fetched by a 3-D TMA box.  Accumulators live in TMEM.
//
// Warp roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer (one
// lane), warps 2..5 = epilogue (tcgen05.ld -> bias/ReLU -> bf16/fp32 -> global).  A STAGES-deep mbarrier
// ring couples producer and MMA; tcgen05.commit releases stages and publishes the accumulator.
//
// Weight gradient: dW[tap][co][ci] = sum_pixels dy[p][co] * x[p + tap][ci] is a GEMM whose K dimension is
// the pixel index, so both operands are MN-major views of the very same TMA boxes; partial sums over
// pixel-tile slices are reduced with fp32 red.global.add.
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "../../include/octave_b200.h"

namespace {

struct ConvTcParams {
  int taps;
  int tap_dh[9], tap_dw[9];
  int kchunks;   // Cin_g / BK
  int cin_g, cout_g;
  int tiles_w, tiles_h, TW, TH;
  int n_tiles, groups;
  int H, W;      // extent of the pixel grid the GEMM M dimension runs over
  // output addressing: pixel (n, oh, ow) -> out + ((n*Hout + oh)*Wout + ow)*ldc + c_off + channel
  void* out;
  long long ldc;
  int c_off, Hout, Wout;
  int scatter;       // 1: ConvTranspose k2s2 — GEMM column n = tap*cout_total + co, pixel (2h+i, 2w+j)
                     // 2: space-to-depth store — pixel (h>>1, w>>1), channel += ((h&1)*2 + (w&1)) * s2d_qs
  int cout_total;    // channels per tap in scatter mode 1
  int s2d_qs;        // quadrant stride in scatter mode 2
  const float* bias;
  int act, out_f32, accumulate;
  long long total_tiles;
  double* stats;        // nullable: [2][stats_stride] per-channel sum, sum of squares of the stored outputs
  int stats_stride;
};

__device__ __forceinline__ float tc_act(float v, int act) {
  switch (act) {
    case 1: return fmaxf(v, 0.f);
    case 2: return v > 0.f ? v : 0.2f * v;
    case 3: return 1.f / (1.f + __expf(-v));
    case 4: return tanhf(v);
    default: return v;
  }
}


// 16 fp32 accumulator columns -> two 16-byte vectors of bf16
__device__ __forceinline__ void pack16(const uint32_t (&v)[16], uint4& u0, uint4& u1) {
  u0.x = bf16x2_pack(__uint_as_float(v[0]), __uint_as_float(v[1]));   u0.y = bf16x2_pack(__uint_as_float(v[2]), __uint_as_float(v[3]));
  u0.z = bf16x2_pack(__uint_as_float(v[4]), __uint_as_float(v[5]));   u0.w = bf16x2_pack(__uint_as_float(v[6]), __uint_as_float(v[7]));
  u1.x = bf16x2_pack(__uint_as_float(v[8]), __uint_as_float(v[9]));   u1.y = bf16x2_pack(__uint_as_float(v[10]), __uint_as_float(v[11]));
  u1.z = bf16x2_pack(__uint_as_float(v[12]), __uint_as_float(v[13])); u1.w = bf16x2_pack(__uint_as_float(v[14]), __uint_as_float(v[15]));
}

// ---------------------------------------------------------------------------------------------------
// forward / dgrad kernel — persistent: one CTA per SM loops over output tiles (N tile fastest, so the CTAs that
// share an A tile run side by side and hit L2).  Two TMEM accumulators: the epilogue of tile i (TMEM -> registers ->
// bias/activation -> bf16 -> warp-private padded smem -> 16-byte coalesced global stores) overlaps the TMA/MMA main
// loop of tile i+1.  Optional fused BatchNorm statistics: per-channel sum / sum of squares of the stored (bf16-rounded)
// outputs are reduced across the 32 rows of a warp with a shuffle butterfly, kept in registers across the CTA's tiles
// (the grid is a multiple of n_tiles*groups, so a CTA always owns the same channels) and flushed with one fp64 atomic
// per channel per warp at the end.
// ---------------------------------------------------------------------------------------------------
constexpr int kProducers = 4;                         // TMA producer warps (one issuing lane each)
constexpr int kFwdThreads = (10 + kProducers - 1) * 32;  // warp 0 + warps 10.. producers, warp 1 MMA, warps 2..9 epilogue

// EPI = 1 (plain outputs: no bias / activation / accumulate / scatter, BN <= 128): the epilogue stages the bf16 tile in the
// TMA box layout (128 pixel rows x 64-column slabs, 16-byte units XOR-swizzled like SWIZZLE_128B/64B/32B) and ONE thread
// per slab issues a TMA store (cp.async.bulk.tensor ... bulk_group) — the hardware clips rows / columns outside the tensor,
// so there are no per-row offsets, shuffles or per-lane global stores left in the epilogue (ncu on the 1x1 layers: the
// per-lane store loop, the predicated bias / activation code and the barrier spin loops were > 2/3 of the issued
// instructions of these HBM-bound launches).
template <int BN, int BK, int STAGES, int EPI = 0>
__global__ void __launch_bounds__(kFwdThreads, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                             const __grid_constant__ CUtensorMap tmB,
                                                             const __grid_constant__ CUtensorMap tmC,
                                                             const ConvTcParams p) {
  constexpr int A_BYTES = 128 * BK * 2;
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int SWZ = BK * 2;  // bytes per smem row = swizzle span
  constexpr int ACC_COLS = BN < 32 ? 32 : BN;
  constexpr int TMEM_COLS = 2 * ACC_COLS;
  constexpr int COLS_W = BN >= 32 ? BN / 2 : BN;   // columns handled by one epilogue warp
  constexpr int NCHUNK_W = COLS_W / 16;
  // SLAB (256-column tiles with a 4-deep ring): the epilogue staging shrinks from 32 x 272 B to 32 x 128 B per warp - one
  // 64-column slab at a time, 16-byte units XOR-swizzled by the row instead of padded - which frees the 34 KB the fourth
  // 48 KB stage needs.
  // The same trade for the 128-column kernel (6 x 32 KB stages, one slab per warp) is compiled but opt-in
  // (OCTAVE_FWD128_STAGES=6): it has not been A/B-ed on the GPU yet.
  constexpr bool SLAB = (BN == 256 && STAGES >= 4) || (BN == 128 && STAGES >= 6);
  constexpr int NSLAB = COLS_W >= 64 ? COLS_W / 64 : 1;   // 64-column slabs per epilogue warp
  constexpr int PITCH = SLAB ? 128 : COLS_W * 2 + 16;   // bytes per staged row (+16 spreads the banks)
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t acc_full[2];
  __shared__ __align__(8) uint64_t acc_empty[2];
  __shared__ uint32_t tmem_slot;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const smem_gen = smem_raw + (smem_base - tc::smem_u32(smem_raw));

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(tc::smem_u32(&full_bar[s]), 1);
      tc::mbar_init(tc::smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(tc::smem_u32(&acc_full[b]), 1);
      tc::mbar_init(tc::smem_u32(&acc_empty[b]), EPI == 1 ? 4 : 8);  // one arrival per epilogue warp that reads this accumulator
    }
    tc::fence_barrier_init();
    tc::fence_proxy_async();
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmB);
    if (EPI == 1) tc::tma_prefetch_desc(&tmC);
  }
  if (EPI == 1) {
    // staging rows beyond the patch (TW*TH < 128) are never written by the epilogue: zero them once (they feed the statistics)
    uint8_t* st0 = smem_gen + STAGES * STAGE_BYTES;
    for (int i = threadIdx.x * 16; i < 2 * 128 * BN * 2; i += kFwdThreads * 16) *reinterpret_cast<uint4*>(st0 + i) = make_uint4(0, 0, 0, 0);
  }
  if (warp == 1) tc::tmem_alloc<TMEM_COLS>(tc::smem_u32(&tmem_slot));
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  const int total_k = p.taps * p.kchunks;
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int ng = p.n_tiles * p.groups;

  if (warp == 0 || warp >= 10) {
    // kProducers TMA producer warps.  Even roles load the A tile of their k-iterations and arm the barrier with the
    // stage's byte count, odd roles load the B tile.  The loops are warp-uniform (every lane walks the tiles and waits on
    // the barriers) and one elected lane issues, so coordinates and addresses stay in uniform registers.
    const bool leader = tc::elect_one();
    const int role = warp == 0 ? 0 : warp - 9;       // 0 .. kProducers-1
    const bool is_a = (role & 1) == 0;                // even roles load A (and arm the barrier), odd roles load B
    const uint32_t it_sel = (uint32_t)(role >> 1);    // this producer serves k-iterations it % (kProducers/2) == it_sel
    const uint32_t tx_bytes = (uint32_t)(p.TW * p.TH * SWZ + B_BYTES);
    uint32_t it = 0, ph = 0;
    int s = 0;
    for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const int tt = (int)t;
      const int n_tile = tt % p.n_tiles;
      const int g = (tt / p.n_tiles) % p.groups;
      const int m_tile = tt / ng;
      const int img = m_tile / tiles_per_img;
      const int trem = m_tile - img * tiles_per_img;
      const int th_i = trem / p.tiles_w;
      const int h0 = th_i * p.TH, w0 = (trem - th_i * p.tiles_w) * p.TW;
      const int n0 = n_tile * BN;
      for (int tap = 0; tap < p.taps; ++tap) {
        const int cw = w0 + p.tap_dw[tap], ch = h0 + p.tap_dh[tap];
        for (int kc = 0; kc < p.kchunks; ++kc, ++it) {
          if ((it % (kProducers / 2)) == it_sel) {
            tc::mbar_wait(tc::smem_u32(&empty_bar[s]), ph ^ 1u);
            const uint32_t fb = tc::smem_u32(&full_bar[s]);
            const uint32_t sa = smem_base + s * STAGE_BYTES;
            if (leader) {
              if (is_a) {
                tc::mbar_arrive_expect_tx(fb, tx_bytes);
                tc::tma_load_4d(sa, &tmA, fb, g * p.cin_g + kc * BK, cw, ch, img);
              } else {
                tc::tma_load_3d(sa + A_BYTES, &tmB, fb, kc * BK, g * p.cout_g + n0, tap);
              }
            }
          }
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: warp-uniform loop, one elected lane issues (see tc::elect_one)
    const bool leader = tc::elect_one();
    constexpr uint32_t idesc = tc::umma_idesc_bf16(128, BN, 0, 0);
    uint32_t li = 0, ph = 0;
    int s = 0;
    for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++li) {
      const uint32_t buf = li & 1u;
      tc::mbar_wait(tc::smem_u32(&acc_empty[buf]), ((li >> 1) & 1u) ^ 1u);  // epilogue drained this accumulator
      tc::fence_after_sync();
      const uint32_t d_tmem = tmem_base + buf * ACC_COLS;
      for (int kk = 0; kk < total_k; ++kk) {
        tc::mbar_wait(tc::smem_u32(&full_bar[s]), ph);
        tc::fence_after_sync();
        const uint32_t sa = smem_base + s * STAGE_BYTES;
        const uint64_t da = tc::umma_smem_desc(sa, SWZ, 16);
        const uint64_t db = tc::umma_smem_desc(sa + A_BYTES, SWZ, 16);
        if (leader) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) tc::umma_bf16(d_tmem, da + k * 2, db + k * 2, idesc, (kk | k) != 0);
          tc::umma_commit(tc::smem_u32(&empty_bar[s]));
        }
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
      if (leader) tc::umma_commit(tc::smem_u32(&acc_full[buf]));
    }
  } else if (EPI == 1) {
    // ---- TMA-store epilogue (see the kernel comment): TWO groups of four warps take alternate tiles (group = accumulator
    // buffer), so the TMEM -> staging -> statistics -> store chain of one tile overlaps the next tile's — these launches
    // are bound by that per-tile latency chain, not by issue slots or HBM.  A warp owns the 32 rows of its TMEM lane
    // quarter over ALL BN columns; a slab = up to 64 columns x 128 rows in the TMA box layout.
    constexpr int SWC = BN < 64 ? BN : 64;          // columns of one slab
    constexpr int RB = SWC * 2;                     // bytes per staged row = swizzle span
    constexpr int SLAB_BYTES = 128 * RB;
    constexpr int NSL = BN / SWC;                   // slabs per tile (2 for BN = 128)
    constexpr int CHS = SWC / 16;                   // 16-column TMEM chunks per slab
    const int q = warp & 3;
    const int group = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int th = r / p.TW, tw = r - th * p.TW;
    uint8_t* const sbuf = smem_gen + STAGES * STAGE_BYTES + group * (NSL * SLAB_BYTES);
    const uint32_t sbuf_u32 = smem_base + STAGES * STAGE_BYTES + group * (NSL * SLAB_BYTES);
    const bool issuer = q == 0 && lane == 0;
    const int bar_id = 1 + group;
    auto swz = [](int row) { return RB == 128 ? (row & 7) : (RB == 64 ? ((row >> 1) & 3) : ((row >> 2) & 1)); };
    constexpr int CP = SWC / 2, RG = 32 / CP, RPG = 32 / RG;   // statistics of one slab: column pairs x row groups
    float sacc[NSL][4];
#pragma unroll
    for (int i = 0; i < NSL; ++i) sacc[i][0] = sacc[i][1] = sacc[i][2] = sacc[i][3] = 0.f;
    int stat_col0 = -1;
    const bool row_live = r < p.TW * p.TH;
    const uint32_t buf = (uint32_t)group;
    uint32_t use = 0;                                // how often this group has used its accumulator
    for (long long t = blockIdx.x + (long long)group * gridDim.x; t < p.total_tiles; t += 2LL * gridDim.x, ++use) {
      const int tt = (int)t;
      const int n_tile = tt % p.n_tiles;
      const int g = (tt / p.n_tiles) % p.groups;
      const int m_tile = tt / ng;
      const int img = m_tile / tiles_per_img;
      const int trem = m_tile - img * tiles_per_img;
      const int th_i = trem / p.tiles_w;
      const int h0 = th_i * p.TH, w0 = (trem - th_i * p.tiles_w) * p.TW;
      const bool valid = row_live && (h0 + th < p.H) && (w0 + tw < p.W);
      const bool keep = p.taps == 1 || valid;   // 3x3: a row outside the image still sums in-image taps: keep it out of the statistics
      const int cbase = g * p.cout_g + n_tile * BN;            // first channel of the tile
      stat_col0 = cbase;
      tc::mbar_wait(tc::smem_u32(&acc_full[buf]), use & 1u);
      tc::fence_after_sync();
      const uint32_t taddr = tmem_base + buf * ACC_COLS + ((uint32_t)(q * 32) << 16);
      // (A) the group's previous store has finished READING the staging slabs
      if (issuer) tc::tma_store_wait_read();
      tc::bar_sync_named(bar_id, 128);
#pragma unroll
      for (int sl = 0; sl < NSL; ++sl) {
        uint32_t v[CHS][16];
#pragma unroll
        for (int c = 0; c < CHS; ++c) tc::tmem_ld16(taddr + (sl * CHS + c) * 16, v[c]);
        tc::tmem_ld_wait();
        if (sl == NSL - 1) {
          // every column is in registers: hand the accumulator back to the MMA warp before packing
          tc::fence_before_sync();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(tc::smem_u32(&acc_empty[buf]));
        }
        if (row_live) {
          uint8_t* rowp = sbuf + sl * SLAB_BYTES + r * RB;
#pragma unroll
          for (int c = 0; c < CHS; ++c) {
            uint4 u0, u1;
            pack16(v[c], u0, u1);
            if (!keep) { u0 = make_uint4(0, 0, 0, 0); u1 = u0; }
            *reinterpret_cast<uint4*>(rowp + (((2 * c) ^ swz(r)) << 4)) = u0;
            *reinterpret_cast<uint4*>(rowp + (((2 * c + 1) ^ swz(r)) << 4)) = u1;
          }
        }
      }
      __syncwarp();
      if (p.stats) {
        // lane owns the column pair cp of each slab over the rows rg, rg + RG, ... of the warp's 32 rows
        const int cp = lane % CP, rg = lane / CP;
#pragma unroll
        for (int sl = 0; sl < NSL; ++sl) {
          unsigned long long s01 = 0ull, q01 = 0ull;
#pragma unroll 8
          for (int i = 0; i < RPG; ++i) {
            const int row = q * 32 + i * RG + rg;
            const uint32_t u = *reinterpret_cast<const uint32_t*>(sbuf + sl * SLAB_BYTES + row * RB + ((((cp >> 2)) ^ swz(row)) << 4) + (cp & 3) * 4);
            stat_acc_bf16x2(u, s01, q01);
          }
          float s0, s1, q0, q1;
          unpack_f32x2(s01, s0, s1);
          unpack_f32x2(q01, q0, q1);
#pragma unroll
          for (int off = 16; off >= CP; off >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, off); s1 += __shfl_xor_sync(0xffffffffu, s1, off);
            q0 += __shfl_xor_sync(0xffffffffu, q0, off); q1 += __shfl_xor_sync(0xffffffffu, q1, off);
          }
          sacc[sl][0] += s0; sacc[sl][1] += s1; sacc[sl][2] += q0; sacc[sl][3] += q1;
        }
      }
      tc::fence_proxy_async();                       // generic-proxy writes of this thread -> visible to the TMA (async proxy)
      tc::bar_sync_named(bar_id, 128);               // (B) every row of the tile is staged
      if (issuer) {
#pragma unroll
        for (int sl = 0; sl < NSL; ++sl) {
          if (p.accumulate) tc::tma_reduce_add_4d(&tmC, sbuf_u32 + sl * SLAB_BYTES, cbase + sl * SWC, w0, h0, img);   // y += tile
          else tc::tma_store_4d(&tmC, sbuf_u32 + sl * SLAB_BYTES, cbase + sl * SWC, w0, h0, img);
        }
        tc::tma_store_commit();
      }
    }
    if (issuer) tc::tma_store_wait_all();
    if (p.stats && stat_col0 >= 0 && lane < CP) {
#pragma unroll
      for (int sl = 0; sl < NSL; ++sl) {
        atomicAdd(p.stats + stat_col0 + sl * SWC + 2 * lane, (double)sacc[sl][0]);
        atomicAdd(p.stats + stat_col0 + sl * SWC + 2 * lane + 1, (double)sacc[sl][1]);
        atomicAdd(p.stats + p.stats_stride + stat_col0 + sl * SWC + 2 * lane, (double)sacc[sl][2]);
        atomicAdd(p.stats + p.stats_stride + stat_col0 + sl * SWC + 2 * lane + 1, (double)sacc[sl][3]);
      }
    }
  } else {
    // ---- 8 epilogue warps: TMEM lane quarter q <-> GEMM rows [32q, 32q+32); column half hsel
    const int q = warp & 3;
    const int hsel = (warp - 2) >> 2;
    const bool works = (hsel == 0) || (BN >= 32);
    const int r = q * 32 + lane;
    const int th = r / p.TW, tw = r - th * p.TW;
    uint8_t* const stage = smem_gen + STAGES * STAGE_BYTES + (warp - 2) * (32 * PITCH);
    constexpr int NS = COLS_W > 64 ? COLS_W / 64 : 1;   // 64-column statistic slabs per warp
    constexpr int SW = COLS_W > 64 ? 64 : COLS_W;       // columns per slab
    float sacc[NS][4];                                  // statistics of this lane's column pair(s) (tile_col_stats)
#pragma unroll
    for (int h2 = 0; h2 < NS; ++h2) sacc[h2][0] = sacc[h2][1] = sacc[h2][2] = sacc[h2][3] = 0.f;
    int stat_col0 = -1;  // global channel of this warp's column 0 (fixed across the CTA's tiles when stats are on)
    const bool plain = p.bias == nullptr && p.act == 0;
    uint32_t li = 0;
    for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++li) {
      const uint32_t buf = li & 1u;
      const int tt = (int)t;
      const int n_tile = tt % p.n_tiles;
      const int g = (tt / p.n_tiles) % p.groups;
      const int m_tile = tt / ng;
      const int img = m_tile / tiles_per_img;
      const int trem = m_tile - img * tiles_per_img;
      const int th_i = trem / p.tiles_w;
      const int h = th_i * p.TH + th, w = (trem - th_i * p.tiles_w) * p.TW + tw;
      const int n0 = n_tile * BN + hsel * COLS_W;   // first column of this warp inside the group
      bool valid = (r < p.TW * p.TH) && (h < p.H) && (w < p.W);
      int oh = h, ow = w, cbase = g * p.cout_g + n0, cshift = 0;
      if (p.scatter == 1) {
        const int tap = cbase / p.cout_total;
        cbase -= tap * p.cout_total;
        oh = 2 * h + (tap >> 1);
        ow = 2 * w + (tap & 1);
      } else if (p.scatter == 2) {
        oh = h >> 1;
        ow = w >> 1;
        cshift = ((h & 1) * 2 + (w & 1)) * p.s2d_qs;
      }
      valid = valid && (oh < p.Hout) && (ow < p.Wout);
      // element offset of this row's first output channel; -1 marks rows that are not stored
      const long long row_off = valid ? (((long long)img * p.Hout + oh) * p.Wout + ow) * p.ldc + p.c_off + cshift + cbase : -1;
      stat_col0 = cbase;
      const int nvalid = min(COLS_W, p.cout_g - n0);
      // warp-uniform: nothing to add or clamp and no row of this warp falls outside the image
      const bool plain_tile = plain && __all_sync(0xffffffffu, valid);
      tc::mbar_wait(tc::smem_u32(&acc_full[buf]), (li >> 1) & 1u);
      tc::fence_after_sync();
      if constexpr (SLAB) {
        // 64-column slabs (two per warp for BN = 256): TMEM -> registers -> swizzled staging -> statistics -> stores
        const uint32_t taddr = tmem_base + buf * ACC_COLS + hsel * COLS_W + ((uint32_t)(q * 32) << 16);
        const int rx = lane & 7;
#pragma unroll
        for (int h2 = 0; h2 < NSLAB; ++h2) {
          uint32_t v[4][16];
#pragma unroll
          for (int c = 0; c < 4; ++c) tc::tmem_ld16(taddr + (h2 * 4 + c) * 16, v[c]);
          tc::tmem_ld_wait();
          if (h2 == NSLAB - 1) {
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(tc::smem_u32(&acc_empty[buf]));
          }
          if (plain_tile) {
            // common case (conv -> BatchNorm: no bias, no activation, every row inside the image): straight pack, no
            // per-element predicated bias loads / selects
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              uint4 u0, u1;
              pack16(v[cc], u0, u1);
              *reinterpret_cast<uint4*>(stage + lane * 128 + (((2 * cc) ^ rx) << 4)) = u0;
              *reinterpret_cast<uint4*>(stage + lane * 128 + (((2 * cc + 1) ^ rx) << 4)) = u1;
            }
          } else {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              float f[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                f[i] = __uint_as_float(v[cc][i]);
                if (p.bias) f[i] += __ldg(p.bias + cbase + (h2 * 4 + cc) * 16 + i);
              }
              if (p.act) {
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = tc_act(f[i], p.act);
              }
              if (!valid) {
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = 0.f;
              }
              uint4 u0, u1;
              u0.x = bf16x2_pack(f[0], f[1]); u0.y = bf16x2_pack(f[2], f[3]);
              u0.z = bf16x2_pack(f[4], f[5]); u0.w = bf16x2_pack(f[6], f[7]);
              u1.x = bf16x2_pack(f[8], f[9]); u1.y = bf16x2_pack(f[10], f[11]);
              u1.z = bf16x2_pack(f[12], f[13]); u1.w = bf16x2_pack(f[14], f[15]);
              *reinterpret_cast<uint4*>(stage + lane * 128 + (((2 * cc) ^ rx) << 4)) = u0;
              *reinterpret_cast<uint4*>(stage + lane * 128 + (((2 * cc + 1) ^ rx) << 4)) = u1;
            }
          }
          __syncwarp();
          if (p.stats) {
            // lane owns the column pair `lane` of this slab (unit lane/4, word lane%4) over the 32 rows
            unsigned long long s01 = 0ull, q01 = 0ull;
#pragma unroll 8
            for (int row = 0; row < 32; ++row) {
              const uint32_t u = *reinterpret_cast<const uint32_t*>(stage + row * 128 + ((((lane >> 2) ^ (row & 7))) << 4) + (lane & 3) * 4);
              stat_acc_bf16x2(u, s01, q01);
            }
            float s0, s1, q0, q1;
            unpack_f32x2(s01, s0, s1);
            unpack_f32x2(q01, q0, q1);
            sacc[h2][0] += s0; sacc[h2][1] += s1; sacc[h2][2] += q0; sacc[h2][3] += q1;
          }
          {
            // 8 lanes cover one 128-byte row, 4 rows per instruction
            const int sub = lane & 7, rsel = lane >> 3;
#pragma unroll
            for (int r0 = 0; r0 < 32; r0 += 4) {
              const int row = r0 + rsel;
              const long long off = __shfl_sync(0xffffffffu, row_off, row);
              if (off >= 0 && h2 * 64 + sub * 8 < nvalid) {
                const uint4 val = *reinterpret_cast<const uint4*>(stage + row * 128 + ((sub ^ (row & 7)) << 4));
                *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + off + h2 * 64 + sub * 8) = val;
              }
            }
          }
          __syncwarp();   // the staging rows are rewritten by the next slab / tile
        }
        continue;
      }
      if (works) {
        const uint32_t taddr = tmem_base + buf * ACC_COLS + hsel * COLS_W + ((uint32_t)(q * 32) << 16);
        // TMEM loads of up to 64 columns in flight before ONE wait; the accumulator goes back to the MMA warp as soon as
        // the last load has landed, before the bias / activation / packing work of that slab
        constexpr int LDG = NCHUNK_W > 4 ? 4 : NCHUNK_W;
#pragma unroll
        for (int c0 = 0; c0 < NCHUNK_W; c0 += LDG) {
          uint32_t v[LDG][16];
#pragma unroll
          for (int c = 0; c < LDG; ++c) tc::tmem_ld16(taddr + (c0 + c) * 16, v[c]);
          tc::tmem_ld_wait();
          if (c0 + LDG >= NCHUNK_W) {
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(tc::smem_u32(&acc_empty[buf]));
          }
          if (plain_tile) {
            // common case (conv -> BatchNorm: no bias, no activation, every row inside the image): straight pack
#pragma unroll
            for (int cc = 0; cc < LDG; ++cc) {
              uint4 u0, u1;
              pack16(v[cc], u0, u1);
              *reinterpret_cast<uint4*>(stage + lane * PITCH + (c0 + cc) * 32) = u0;
              *reinterpret_cast<uint4*>(stage + lane * PITCH + (c0 + cc) * 32 + 16) = u1;
            }
          } else {
#pragma unroll
            for (int cc = 0; cc < LDG; ++cc) {
              const int c = c0 + cc;
              float f[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                f[i] = __uint_as_float(v[cc][i]);
                if (p.bias) f[i] += __ldg(p.bias + cbase + c * 16 + i);
              }
              if (p.act) {
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = tc_act(f[i], p.act);
              }
              if (!valid) {   // rows outside the image stage zeros: they are not stored and must not enter the statistics
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
        }
        __syncwarp();   // staged rows are read by other lanes below
      } else {
        // idle column half (BN < 32): still hand the accumulator back
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(tc::smem_u32(&acc_empty[buf]));
      }
      if (works) {
        // per-channel sum / sum of squares of the staged (bf16-rounded) values
        if (p.stats) {
#pragma unroll
          for (int h2 = 0; h2 < NS; ++h2) tile_col_stats<SW, PITCH>(stage + h2 * 128, lane, sacc[h2]);
        }
        // coalesced write-out of the warp's 32 staged rows: LPR lanes cover one row (16 B each)
        constexpr int LPR = (COLS_W * 2) / 16;   // lanes per row: 8 (64 cols), 4, 2
        constexpr int RPI = 32 / LPR;            // rows per instruction
        constexpr int NIT = 32 / RPI;            // store instructions per lane
        const int sub = lane % LPR, rsel = lane / LPR;
        if (p.accumulate) {
          // y += result: fetch the old values first (up to 8 independent 16-byte loads in flight), then add and store
          constexpr int NB = NIT > 8 ? 8 : NIT;
#pragma unroll 1
          for (int i0 = 0; i0 < NIT; i0 += NB) {
            uint4 old[NB];
            long long offs[NB];
#pragma unroll
            for (int i = 0; i < NB; ++i) {
              const int row = (i0 + i) * RPI + rsel;
              long long off = __shfl_sync(0xffffffffu, row_off, row);
              if (sub * 8 >= nvalid) off = -1;
              offs[i] = off;
              if (off >= 0) old[i] = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.out) + off + sub * 8);
            }
#pragma unroll
            for (int i = 0; i < NB; ++i) {
              if (offs[i] >= 0) {
                const int row = (i0 + i) * RPI + rsel;
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
          }
        } else {
          // the row offsets travel by shuffle first (independent), then the loads from the staging rows, then the stores
          constexpr int SB = NIT > 8 ? 8 : NIT;
#pragma unroll 1
          for (int i0 = 0; i0 < NIT; i0 += SB) {
            long long offs[SB];
            uint4 vals[SB];
#pragma unroll
            for (int i = 0; i < SB; ++i) offs[i] = __shfl_sync(0xffffffffu, row_off, (i0 + i) * RPI + rsel);
            const bool col_ok = sub * 8 < nvalid;
#pragma unroll
            for (int i = 0; i < SB; ++i) vals[i] = *reinterpret_cast<const uint4*>(stage + ((i0 + i) * RPI + rsel) * PITCH + sub * 16);
#pragma unroll
            for (int i = 0; i < SB; ++i)
              if (offs[i] >= 0 && col_ok) *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + offs[i] + sub * 8) = vals[i];
          }
        }
        __syncwarp();  // staging rows are rewritten by the next tile
      }
    }
    if (p.stats && works && stat_col0 >= 0 && lane < SW / 2) {
      const int nvalid_cols = p.cout_g - (stat_col0 % p.cout_g);
#pragma unroll
      for (int h2 = 0; h2 < NS; ++h2) {
        const int col = h2 * 64 + 2 * lane;
        if (col < nvalid_cols) {
          atomicAdd(p.stats + stat_col0 + col, (double)sacc[h2][0]);
          atomicAdd(p.stats + p.stats_stride + stat_col0 + col, (double)sacc[h2][2]);
        }
        if (col + 1 < nvalid_cols) {
          atomicAdd(p.stats + stat_col0 + col + 1, (double)sacc[h2][1]);
          atomicAdd(p.stats + p.stats_stride + stat_col0 + col + 1, (double)sacc[h2][3]);
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------
// weight-gradient kernel.  D[128 x BN] (rows = output channels, cols = input channels of one tap)
// accumulates over the pixel tiles of this CTA's slice; operands are MN-major.
// ---------------------------------------------------------------------------------------------------
struct WgradParams {
  int taps;
  int tap_dhs[9], tap_dws[9];
  int cin_g, cout_g;          // per (dense) group
  int real_cin_g, real_cout_g;  // per real group (<= dense group): only diagonal blocks are stored
  int tiles_w, tiles_h, TW, TH, B;
  int n_ci_tiles;    // Cin_g tiles of BN
  int tiles_per_cta; // pixel tiles handled by one CTA (split-K slice)
  int total_tiles;   // B * tiles_w * tiles_h
  float* dw;         // torch layout [Cout][Cin/groups][k][k] (conv) or [Cin][Cout][2][2] (convT), pre-zeroed
  int cout_total;
  int convt;         // 1: GEMM rows are n = t*Cout + co of a ConvTranspose k2s2
  int convt_cout;
  int single_writer; // 1: no split-K (gridDim.z == 1): plain stores, the output needs no zero fill
};

constexpr int kWgProducers = 4;
constexpr int kWgThreads = (6 + kWgProducers - 1) * 32;  // warp 0 + warps 6.. producers, warp 1 MMA, warps 2..5 epilogue

template <int BN, int CWA, int CWB, int NA, int STAGES, int ROWS>
__global__ void __launch_bounds__(kWgThreads) conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY,
                                                                const __grid_constant__ CUtensorMap tmX,
                                                                const WgradParams p) {
  constexpr int PITCH_A = CWA * 2, PITCH_B = CWB * 2;   // bytes per pixel row of one box
  // ROWS = pixel rows of one box region (>= 16 * ksteps): 128, or fewer rows per stage and a deeper ring where the
  // stage is so large that two of them cannot hide the L2 latency (BN = 256)
  constexpr int CHUNK_A = ROWS * PITCH_A, CHUNK_B = ROWS * PITCH_B;
  // NA = A chunks that really exist (cout_g <= NA*CWA): only they get shared memory; the 128-row UMMA still strides
  // over 128/CWA chunks, the missing ones alias whatever follows (their accumulator rows are never stored)
  constexpr int NCH_A = NA, NCH_B = (BN + CWB - 1) / CWB;
  constexpr int A_BYTES = NCH_A * CHUNK_A;
  constexpr int B_BYTES = NCH_B * CHUNK_B;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t acc_bar;
  __shared__ uint32_t tmem_slot;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;

  // blockIdx.x = tap + taps * (output tile): the CTAs that read the same pixel slice (all taps, all channel tiles)
  // are adjacent in launch order, so x / dy are fetched from HBM once and then served by L2.
  const int tap = blockIdx.x % p.taps;
  const int otile = blockIdx.x / p.taps;
  const int co_tile = otile / p.n_ci_tiles, ci_tile = otile - co_tile * p.n_ci_tiles;
  const int co0 = co_tile * 128, ci0 = ci_tile * BN;
  const int g = blockIdx.y;
  const int t_begin = blockIdx.z * p.tiles_per_cta;
  const int t_end = min(t_begin + p.tiles_per_cta, p.total_tiles);
  const int n_iter = t_end - t_begin;
  const int rows = p.TW * p.TH;            // pixel rows written by one box
  const int ksteps = (rows + 15) >> 4;     // UMMA K = 16 pixels

  // Rows [rows, 16*ksteps) of every box region are read by the MMA but never written by TMA: zero them.
  {
    uint8_t* base = smem_raw + (smem_base - tc::smem_u32(smem_raw));
    const int tail_rows = ksteps * 16 - rows;
    for (int st = 0; st < STAGES; ++st) {
      for (int rgn = 0; rgn < NCH_A; ++rgn) {
        uint8_t* q = base + st * STAGE_BYTES + rgn * CHUNK_A + rows * PITCH_A;
        for (int i = threadIdx.x * 16; i < tail_rows * PITCH_A; i += kWgThreads * 16) *reinterpret_cast<uint4*>(q + i) = make_uint4(0, 0, 0, 0);
      }
      for (int rgn = 0; rgn < NCH_B; ++rgn) {
        uint8_t* q = base + st * STAGE_BYTES + A_BYTES + rgn * CHUNK_B + rows * PITCH_B;
        for (int i = threadIdx.x * 16; i < tail_rows * PITCH_B; i += kWgThreads * 16) *reinterpret_cast<uint4*>(q + i) = make_uint4(0, 0, 0, 0);
      }
    }
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
  tc::fence_proxy_async();  // generic-proxy zero fill above must be visible to the tensor-core (async) proxy
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
  // number of channel boxes that exist on each side (the rest of the 128 x BN tile is never stored)
  const int a_boxes = min(NCH_A, (p.cout_g - co0 + CWA - 1) / CWA);
  const int b_boxes = min(NCH_B, (p.cin_g - ci0 + CWB - 1) / CWB);

  if (warp == 0 || warp >= 6) {
    // kWgProducers TMA producer warps (warp-uniform loops, one elected lane issues); the boxes of a pixel tile are dealt
    // round-robin, producer 0 also arms the barrier with the byte count
    const bool leader = tc::elect_one();
    const int role = warp == 0 ? 0 : warp - 5;
    const uint32_t tx_bytes = (uint32_t)(rows * (PITCH_A * a_boxes + PITCH_B * b_boxes));
    const int dh = p.tap_dhs[tap], dw = p.tap_dws[tap];
    const int nbox = a_boxes + b_boxes;
    if (role < nbox || role == 0) {
      int s = 0;
      uint32_t ph = 0;
      int img = t_begin / tiles_per_img;
      int trem = t_begin - img * tiles_per_img;
      for (int it = 0; it < n_iter; ++it) {
        tc::mbar_wait(tc::smem_u32(&empty_bar[s]), ph ^ 1u);
        const uint32_t fb = tc::smem_u32(&full_bar[s]);
        const int th_i = trem / p.tiles_w;
        const int h0 = th_i * p.TH, w0 = (trem - th_i * p.tiles_w) * p.TW;
        const uint32_t sa = smem_base + s * STAGE_BYTES;
        if (leader) {
          if (role == 0) tc::mbar_arrive_expect_tx(fb, tx_bytes);
          for (int j = role; j < nbox; j += kWgProducers) {
            if (j < a_boxes)
              tc::tma_load_4d(sa + j * CHUNK_A, &tmDY, fb, g * p.cout_g + co0 + CWA * j, w0, h0, img);
            else
              tc::tma_load_4d(sa + A_BYTES + (j - a_boxes) * CHUNK_B, &tmX, fb, g * p.cin_g + ci0 + CWB * (j - a_boxes), w0 + dw,
                              h0 + dh, img);
          }
        }
        if (++trem == tiles_per_img) { trem = 0; ++img; }
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    const bool leader = tc::elect_one();
    constexpr uint32_t idesc = tc::umma_idesc_bf16(128, BN, 1, 1);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < n_iter; ++it) {
      tc::mbar_wait(tc::smem_u32(&full_bar[s]), ph);
      tc::fence_after_sync();
      const uint32_t sa = smem_base + s * STAGE_BYTES;
      const uint64_t da = tc::umma_smem_desc(sa, PITCH_A, CHUNK_A);
      const uint64_t db = tc::umma_smem_desc(sa + A_BYTES, PITCH_B, CHUNK_B);
      if (leader) {
        for (int k = 0; k < ksteps; ++k)
          tc::umma_bf16(tmem_base, da + (uint32_t)(k * PITCH_A), db + (uint32_t)(k * PITCH_B), idesc, (it | k) != 0);
        tc::umma_commit(tc::smem_u32(&empty_bar[s]));
      }
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }
    if (leader) tc::umma_commit(tc::smem_u32(&acc_bar));
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;  // output channel within the tile
    const bool valid = (co0 + r) < p.cout_g;
    tc::mbar_wait(tc::smem_u32(&acc_bar), 0);
    tc::fence_after_sync();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int co = g * p.cout_g + co0 + r;           // global GEMM row
    const int nvalid = min(BN, p.cin_g - ci0);
    // 1x1 layers without merged groups: dW[co][ci] is contiguous in ci, so a lane's 16 columns are four 16-byte vectors —
    // one red.global.add.v4.f32 (or one 16-byte store when this CTA is the only writer) instead of four scalar atomics
    const bool vec = !p.convt && p.taps == 1 && p.real_cin_g == p.cin_g && p.real_cout_g == p.cout_g && (p.cin_g & 3) == 0;
    const bool single = p.single_writer != 0;        // no split-K: every element is written exactly once (no memset, no atomics)
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t v[16];
      tc::tmem_ld16(taddr + c0, v);
      tc::tmem_ld_wait();
      if (valid) {
        if (vec) {
          float* dst = p.dw + (long long)co * p.cin_g + ci0 + c0;      // (group offset is inside co: co * cin_g indexes [Cout][Cin/g])
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            if (c0 + i < nvalid) {
              if (single) {
                *reinterpret_cast<float4*>(dst + i) = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
              } else {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(__uint_as_float(v[i])), "f"(__uint_as_float(v[i + 1])),
                             "f"(__uint_as_float(v[i + 2])), "f"(__uint_as_float(v[i + 3])) : "memory");
              }
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (c0 + i < nvalid) {
              const int ci = g * p.cin_g + ci0 + c0 + i;  // global input channel
              float* dst;
              if (p.convt) {
                const int t = co / p.convt_cout, cc = co - t * p.convt_cout;
                dst = p.dw + ((long long)ci * p.convt_cout + cc) * 4 + t;
              } else if (ci / p.real_cin_g == co / p.real_cout_g) {
                dst = p.dw + ((long long)co * p.real_cin_g + (ci % p.real_cin_g)) * p.taps + tap;
              } else {
                continue;
              }
              if (single) *dst = __uint_as_float(v[i]); else atomicAdd(dst, __uint_as_float(v[i]));
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

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
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

// NHWC activation view -> 4-D map (C, W, H, B); box (box_c, TW, TH, 1).
bool make_act_map(CUtensorMap* m, const void* base, int C_extent, int W, int H, int B, long long ld, int box_c, int TW,
                  int TH, int swizzle_bytes) {
  auto enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[4] = {(cuuint64_t)C_extent, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * W, (cuuint64_t)ld * 2 * W * H};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)TW, (cuuint32_t)TH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// packed weights [taps][rows][K] -> 3-D map (K, rows, taps); box (BK, BN, 1).
bool make_w_map(CUtensorMap* m, const void* base, int K, int rows, int taps, int BK, int BN, int swizzle_bytes) {
  auto enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)taps};
  cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)K * 2 * rows};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)BN, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// Pick the spatial patch TW x TH (<= max_px pixels, TW,TH <= 256) that wastes the fewest UMMA rows.  quantum = the row
// granularity the kernel pays for: the whole 128-row M tile (forward), or one 16-pixel K step (weight gradient).
void pick_patch(int H, int W, int* TW, int* TH, int max_px = 128, int quantum = 128) {
  double best = -1.0;
  int bw = 1, bh = 1;
  for (int tw = 1; tw <= (W < max_px ? W : max_px); ++tw) {
    int th = max_px / tw;
    if (th > H) th = H;
    if (th > 256) th = 256;
    long long tiles = (long long)((W + tw - 1) / tw) * ((H + th - 1) / th);
    double eff = (double)H * W / ((double)tiles * (double)((tw * th + quantum - 1) / quantum * quantum));
    if (eff > best + 1e-9 || (eff > best - 1e-9 && tw > bw)) {
      best = eff; bw = tw; bh = th;
    }
  }
  *TW = bw; *TH = bh;
}

template <int BN, int BK, int STAGES, int EPI = 0>
int launch_fwd(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const ConvTcParams& p, int grid, cudaStream_t s) {
  constexpr int pitch = ((BN == 256 && STAGES >= 4) || (BN == 128 && STAGES >= 6)) ? 128
                                                                                    : (BN >= 32 ? BN / 2 : BN) * 2 + 16;   // see SLAB
  constexpr int staging = EPI == 1 ? 2 * 128 * BN * 2 : 8 * 32 * pitch;   // EPI 1: one staged tile per epilogue group
  constexpr int smem = STAGES * (128 * BK * 2 + BN * BK * 2) + staging + 1024;
  static_assert(smem <= 226 * 1024, "shared memory budget");
  // (the attribute is per device: set on every launch, it is cheap)
  if (cudaFuncSetAttribute(conv_tc_kernel<BN, BK, STAGES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
    return OCT_ERR_LAUNCH;
  conv_tc_kernel<BN, BK, STAGES, EPI><<<grid, kFwdThreads, smem, s>>>(tmA, tmB, tmC, p);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

template <int BN, int CWA, int CWB, int NA, int ROWS>
constexpr int wgrad_stage_bytes() { return NA * ROWS * CWA * 2 + ((BN + CWB - 1) / CWB) * ROWS * CWB * 2; }

template <int BN, int CWA, int CWB, int NA, int STAGES, int ROWS = 128>
int launch_wgrad(const CUtensorMap& tmDY, const CUtensorMap& tmX, const WgradParams& p, dim3 grid, cudaStream_t s) {
  // +32 KB slack: the chunks a 128-row UMMA strides over beyond NA must stay inside the allocation
  constexpr int smem = STAGES * wgrad_stage_bytes<BN, CWA, CWB, NA, ROWS>() + (128 / CWA - NA) * ROWS * CWA * 2 + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static unsigned long long attr_devs = 0;   // the attribute is per device: one bit per device ordinal
  int dev__ = 0;
  cudaGetDevice(&dev__);
  if (!(attr_devs >> (dev__ & 63) & 1ull)) {
    if (cudaFuncSetAttribute(conv_tc_wgrad_kernel<BN, CWA, CWB, NA, STAGES, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return OCT_ERR_LAUNCH;
    attr_devs |= 1ull << (dev__ & 63);
  }
  conv_tc_wgrad_kernel<BN, CWA, CWB, NA, STAGES, ROWS><<<grid, kWgThreads, smem, s>>>(tmDY, tmX, p);
  OCT_CHECK_LAUNCH();
  return OCT_OK;
}

int validate_conv(const OctaveConvDesc* d) {
  if (!d) return OCT_ERR_INVALID;
  if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->cin <= 0 || d->cout <= 0 || d->groups <= 0) return OCT_ERR_INVALID;
  if (d->cin % d->groups || d->cout % d->groups) return OCT_ERR_INVALID;
  if (d->ksize != 1 && d->ksize != 3) return OCT_ERR_UNSUPPORTED;
  if (d->x_ld % 8 || d->x_coff % 8 || d->y_ld % 8 || d->y_coff % 8) return OCT_ERR_UNSUPPORTED;
  return OCT_OK;
}

}  // namespace

extern "C" int octave_conv_tc_supported(const OctaveConvDesc* d) {
  if (validate_conv(d) != OCT_OK) return 0;
  const int cin_g = d->cin / d->groups, cout_g = d->cout / d->groups;
  if (d->mode == OCT_CONV_MODE_CONVT && (d->groups != 1 || d->ksize != 1)) return 0;
  if (cin_g % 32) return 0;                 // K chunk of 64 (SW128) or 32 (SW64)
  if (cout_g % 16) return 0;                // epilogue vector width
  if (d->groups > 1 && cout_g % 32) return 0;
  return 1;
}

extern "C" int octave_conv_tc_fwd(const OctaveConvDesc* d, const void* x, const void* wpack, const float* bias, void* y,
                                  double* stats, void* stream) {
  int rc = validate_conv(d);
  if (rc != OCT_OK) return rc;
  if (!octave_conv_tc_supported(d)) return OCT_ERR_UNSUPPORTED;
  if (!x || !wpack || !y) return OCT_ERR_INVALID;
  if (octave_conv_halo_supported(d)) return octave_conv_halo_fwd(d, x, wpack, bias, y, stats, stream);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int cin_g = d->cin / d->groups, cout_g = d->cout / d->groups;
  const bool convt = d->mode == OCT_CONV_MODE_CONVT;
  const int n_cols_g = convt ? 4 * d->cout : cout_g;  // GEMM N per group
  ConvTcParams p{};
  p.taps = d->ksize * d->ksize;
  for (int t = 0; t < p.taps; ++t) {
    p.tap_dh[t] = d->ksize == 3 ? t / 3 - 1 : 0;
    p.tap_dw[t] = d->ksize == 3 ? t % 3 - 1 : 0;
  }
  const int BK = (cin_g % 64 == 0) ? 64 : 32;
  p.kchunks = cin_g / BK;
  p.cin_g = cin_g;
  p.cout_g = n_cols_g;
  // A operand dims (Ha, Wa) = the input tensor; GEMM M grid (Hm, Wm) = the output pixel grid (input grid for ConvT).
  int Ha = d->H, Wa = d->W, B = d->B;
  int Hm = convt ? d->H : d->Hout, Wm = convt ? d->W : d->Wout;
  const bool s2d = d->out_s2d_qs > 0;
  // 1x1: flatten all pixels into one long row so that every tile is full
  const bool flat = d->ksize == 1 && !convt && !s2d && d->H == d->Hout && d->W == d->Wout;
  if (flat) { Wa = Wm = d->B * d->H * d->W; Ha = Hm = 1; B = 1; }
  pick_patch(Hm, Wm, &p.TW, &p.TH);
  p.tiles_w = (Wm + p.TW - 1) / p.TW;
  p.tiles_h = (Hm + p.TH - 1) / p.TH;
  p.H = Hm; p.W = Wm;
  p.out = y; p.ldc = d->y_ld; p.c_off = d->y_coff;
  if (flat) { p.Hout = 1; p.Wout = Wm; }
  else if (s2d) { p.Hout = (Hm + 1) / 2; p.Wout = (Wm + 1) / 2; }
  else { p.Hout = d->Hout; p.Wout = d->Wout; }
  p.scatter = convt ? 1 : (s2d ? 2 : 0); p.cout_total = d->cout; p.s2d_qs = d->out_s2d_qs;
  p.bias = bias; p.act = d->relu; p.out_f32 = d->out_dtype == OCT_DTYPE_F32; p.accumulate = d->accumulate;
  int BN = 128;
  if (n_cols_g % 128) BN = (n_cols_g % 64 == 0) ? 64 : ((n_cols_g % 32 == 0) ? 32 : 16);
  {
    // 256-column tiles for the long-K, many-tile convolutions (the fat decoder layers): the A tile is fetched once per
    // 256 output channels, which takes a quarter of the L2->SM traffic off the layers that sit on that limit
    static int bn256 = -1;
    if (bn256 < 0) { const char* e = getenv("OCTAVE_BN256"); bn256 = e ? atoi(e) : 1; }
    int sms0 = octave_sm_count();
    if (sms0 <= 0) sms0 = 148;
    const long long tiles256 = (long long)p.tiles_w * p.tiles_h * B * (n_cols_g / 256) * d->groups;
    if (bn256 && BK == 64 && n_cols_g % 256 == 0 && !convt && !s2d && !d->accumulate && p.taps * p.kchunks >= 16 &&
        tiles256 >= 4LL * sms0)
      BN = 256;
  }
  if (convt && d->cout % BN) BN = (d->cout % 64 == 0) ? 64 : 32;  // a column tile must stay inside one tap
  if (convt && d->cout % BN) return OCT_ERR_UNSUPPORTED;
  CUtensorMap tmA, tmB;
  const bf16* xb = reinterpret_cast<const bf16*>(x) + d->x_coff;
  if (!make_act_map(&tmA, xb, d->cin, Wa, Ha, B, d->x_ld, BK, p.TW, p.TH, BK * 2)) return OCT_ERR_LAUNCH;
  if (!make_w_map(&tmB, wpack, cin_g, d->groups * n_cols_g, p.taps, BK, BN, BK * 2)) return OCT_ERR_LAUNCH;
  p.n_tiles = (n_cols_g + BN - 1) / BN;
  p.groups = d->groups;
  p.total_tiles = (long long)p.tiles_w * p.tiles_h * B * p.n_tiles * d->groups;
  p.stats = stats; p.stats_stride = d->cout;
  if (p.out_f32) return OCT_ERR_UNSUPPORTED;
  if (stats && (convt || s2d)) return OCT_ERR_UNSUPPORTED;
  int sms = octave_sm_count();
  if (sms <= 0) sms = 148;
  long long grid = sms;
  const int ng = p.n_tiles * d->groups;
  if (stats) {
    // a CTA must always own the same output channels: grid is a multiple of n_tiles * groups
    if (ng > sms) return OCT_ERR_UNSUPPORTED;
    grid = (sms / ng) * ng;
    if (!g_octave_stats_prezeroed && cudaMemsetAsync(stats, 0, sizeof(double) * 2 * d->cout, s) != cudaSuccess) return OCT_ERR_LAUNCH;
  }
  if (grid > p.total_tiles) grid = stats ? ((p.total_tiles + ng - 1) / ng) * ng : p.total_tiles;
  if (stats && grid > p.total_tiles) grid = p.total_tiles;  // total_tiles is itself a multiple of ng
  // plain outputs of <= 128-column tiles take the TMA-store epilogue (OCTAVE_TMA_EPI=0 restores the per-lane stores)
  static const int tma_epi = [] { const char* e = getenv("OCTAVE_TMA_EPI"); return e ? atoi(e) : 1; }();
  const bool epi1 = tma_epi && BN <= 128 && !bias && d->relu == 0 && !convt && !s2d && !p.out_f32;   // accumulate: TMA reduce-add
  CUtensorMap tmC = tmA;
  if (epi1) {
    const int swc = BN < 64 ? BN : 64;
    const bf16* yb = reinterpret_cast<const bf16*>(y) + d->y_coff;
    if (!make_act_map(&tmC, yb, d->cout, Wm, Hm, B, d->y_ld, swc, p.TW, p.TH, swc * 2)) return OCT_ERR_LAUNCH;
  }
  if (BK == 64) {
    switch (BN) {
      // pipeline depth sized to ~160-190 KB in flight per SM: the persistent CTA is alone on its SM, so the ring must
      // cover the HBM bandwidth-delay product by itself, also for the small stages of narrow layers
      case 256: {
        // four 48 KB stages with the slab epilogue (see SLAB): bit-exact with the 3-stage kernel and 2.6-3.8 % faster
        // (profiles/slab_epilogue_ab_r01.log); OCTAVE_FWD_STAGES=3 selects the padded-staging kernel
        static const int st2 = [] { const char* e = getenv("OCTAVE_FWD_STAGES"); return e ? atoi(e) : 4; }();
        if (st2 == 3) return launch_fwd<256, 64, 3>(tmA, tmB, tmC, p, (int)grid, s);
        return launch_fwd<256, 64, 4>(tmA, tmB, tmC, p, (int)grid, s);
      }
      case 128: {
        // short-K tiles (1x1 layers) are bound by the epilogue chain: 5 x 32 KB ring + 2 x 32 KB staged tiles; long-K tiles
        // (3x3) hide the epilogue behind the MMAs and want the deeper ring of the slab kernel below
        if (epi1 && p.taps * p.kchunks <= 4) return launch_fwd<128, 64, 5, 1>(tmA, tmB, tmC, p, (int)grid, s);
        // six stages with the slab epilogue (plain stores only); five with the padded staging (accumulating dgrads)
        static const int st1 = [] { const char* e = getenv("OCTAVE_FWD128_STAGES"); return e ? atoi(e) : 6; }();
        if (st1 == 6 && !p.accumulate) return launch_fwd<128, 64, 6>(tmA, tmB, tmC, p, (int)grid, s);
        return launch_fwd<128, 64, 5>(tmA, tmB, tmC, p, (int)grid, s);
      }
      case 64: return epi1 ? launch_fwd<64, 64, 7, 1>(tmA, tmB, tmC, p, (int)grid, s) : launch_fwd<64, 64, 7>(tmA, tmB, tmC, p, (int)grid, s);
      case 32: return epi1 ? launch_fwd<32, 64, 9, 1>(tmA, tmB, tmC, p, (int)grid, s) : launch_fwd<32, 64, 9>(tmA, tmB, tmC, p, (int)grid, s);
      default: return epi1 ? launch_fwd<16, 64, 10, 1>(tmA, tmB, tmC, p, (int)grid, s) : launch_fwd<16, 64, 10>(tmA, tmB, tmC, p, (int)grid, s);
    }
  } else {
    switch (BN) {
      case 128: return epi1 ? launch_fwd<128, 32, 10, 1>(tmA, tmB, tmC, p, (int)grid, s) : launch_fwd<128, 32, 10>(tmA, tmB, tmC, p, (int)grid, s);
      case 64: return epi1 ? launch_fwd<64, 32, 14, 1>(tmA, tmB, tmC, p, (int)grid, s) : launch_fwd<64, 32, 14>(tmA, tmB, tmC, p, (int)grid, s);
      case 32: return epi1 ? launch_fwd<32, 32, 16, 1>(tmA, tmB, tmC, p, (int)grid, s) : launch_fwd<32, 32, 16>(tmA, tmB, tmC, p, (int)grid, s);
      default: return epi1 ? launch_fwd<16, 32, 18, 1>(tmA, tmB, tmC, p, (int)grid, s) : launch_fwd<16, 32, 18>(tmA, tmB, tmC, p, (int)grid, s);
    }
  }
}

extern "C" int octave_conv_tc_wgrad_supported(const OctaveConvDesc* d) {
  if (validate_conv(d) != OCT_OK) return 0;
  const int cin_g = d->cin / d->groups;
  const int cout_g = (d->mode == OCT_CONV_MODE_CONVT ? 4 * d->cout : d->cout) / d->groups;
  if (d->mode == OCT_CONV_MODE_CONVT && (d->groups != 1 || d->ksize != 1)) return 0;
  // 64- or 32-channel TMA boxes on both operands
  if (cin_g % 32 || cout_g % 32) return 0;
  return 1;
}

extern "C" int octave_conv_tc_wgrad(const OctaveConvDesc* d, const void* x, const void* dy, float* dw, void* stream) {
  int rc = validate_conv(d);
  if (rc != OCT_OK) return rc;
  if (!octave_conv_tc_wgrad_supported(d)) return OCT_ERR_UNSUPPORTED;
  if (!x || !dy || !dw) return OCT_ERR_INVALID;
  if (octave_conv_halo_wgrad_supported(d)) return octave_conv_halo_wgrad(d, x, dy, dw, stream);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bool convt = d->mode == OCT_CONV_MODE_CONVT;
  const int cout_all = convt ? 4 * d->cout : d->cout;
  const int cin_g = d->cin / d->groups, cout_g = cout_all / d->groups;
  const int real_groups = d->real_groups > 0 ? d->real_groups : d->groups;
  WgradParams p{};
  p.taps = d->ksize * d->ksize;
  for (int t = 0; t < p.taps; ++t) {
    p.tap_dhs[t] = d->ksize == 3 ? t / 3 - 1 : 0;
    p.tap_dws[t] = d->ksize == 3 ? t % 3 - 1 : 0;
  }
  p.cin_g = cin_g; p.cout_g = cout_g; p.cout_total = cout_all;
  p.real_cin_g = d->cin / real_groups; p.real_cout_g = cout_all / real_groups;
  p.convt = convt ? 1 : 0; p.convt_cout = d->cout;
  int H = d->H, W = d->W, B = d->B;
  if (d->ksize == 1) { W = d->B * d->H * d->W; H = 1; B = 1; }
  const int CWA = (cout_g % 64 == 0) ? 64 : 32;
  const int CWB = (cin_g % 64 == 0) ? 64 : 32;
  int BN;
  if (CWB == 32) BN = 32;
  else if (CWA == 32) BN = 64;
  else BN = (cin_g % 256 == 0) ? 256 : ((cin_g % 128 == 0) ? 128 : 64);
  // 128 x 256 tiles: a 128-pixel stage is 96 KB, and a ring of two cannot cover the L2 latency (ncu: tensor pipe 56 %
  // busy at 13 TB/s of L2->SM traffic, far below what L2 delivers).  Four 64-pixel stages do, even though the smaller
  // patches waste more K rows (measured 1.15 -> 1.25-1.35 PFLOP/s; 96 x 3 is between, 48 x 6 and 32 x 9 are slower:
  // the TMA boxes get too small).  Only for dense 3x3 layers: in the step's launch list the 1x1 and grouped layers at
  // 25x25 with few split-K slices lost 20-50 % with the short stages.  OCTAVE_WGRAD_ROWS=128 restores the two-stage
  // ring everywhere, OCTAVE_WGRAD_ROWS=-64 forces the short stages on every 128 x 256 tile.
  static const int rows_env = [] { const char* e = getenv("OCTAVE_WGRAD_ROWS"); return e ? atoi(e) : 64; }();
  int stage_rows = 128;
  if (BN == 256 && cout_g > CWA && ((rows_env == 64 && d->ksize == 3 && d->groups == 1) || rows_env == -64)) stage_rows = 64;
  pick_patch(H, W, &p.TW, &p.TH, stage_rows, 16);
  p.tiles_w = (W + p.TW - 1) / p.TW;
  p.tiles_h = (H + p.TH - 1) / p.TH;
  p.B = B;
  p.total_tiles = B * p.tiles_w * p.tiles_h;
  p.n_ci_tiles = (cin_g + BN - 1) / BN;
  const int n_co_tiles = (cout_g + 127) / 128;
  const int out_tiles = n_co_tiles * p.n_ci_tiles * p.taps * d->groups;
  int sms = octave_sm_count();
  if (sms <= 0) sms = 148;
  // one wave of CTAs (1 CTA per SM: the TMA ring takes most of the shared memory): no tail wave
  int split = sms / out_tiles;
  if (split < 1 || g_octave_deterministic) split = 1;   // deterministic mode: no split-K, one CTA owns an output tile
  if (split > p.total_tiles) split = p.total_tiles;
  p.tiles_per_cta = (p.total_tiles + split - 1) / split;
  split = (p.total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  p.dw = dw;
  CUtensorMap tmDY, tmX;
  const bf16* xb = reinterpret_cast<const bf16*>(x) + d->x_coff;
  const bf16* dyb = reinterpret_cast<const bf16*>(dy) + d->y_coff;
  if (!make_act_map(&tmX, xb, d->cin, W, H, B, d->x_ld, CWB, p.TW, p.TH, CWB * 2)) return OCT_ERR_LAUNCH;
  if (!make_act_map(&tmDY, dyb, cout_all, d->ksize == 1 ? W : d->Wout, d->ksize == 1 ? H : d->Hout, B, d->y_ld, CWA, p.TW, p.TH, CWA * 2)) return OCT_ERR_LAUNCH;
  const size_t wbytes = (size_t)cout_all * p.real_cin_g * p.taps * sizeof(float);
  // one writer per element (no split-K) and every element of dW covered by a tile: plain stores, no zero fill
  p.single_writer = (split == 1 && !d->accumulate) ? 1 : 0;
  if (!d->accumulate && !p.single_writer && cudaMemsetAsync(dw, 0, wbytes, s) != cudaSuccess) return OCT_ERR_LAUNCH;
  if (split > 65535) return OCT_ERR_UNSUPPORTED;
  dim3 grid(p.taps * n_co_tiles * p.n_ci_tiles, d->groups, split);
  const bool one_a = cout_g <= CWA;   // a single A box per pixel tile
  if (CWA == 64 && CWB == 64) {
    switch (BN) {
      case 256:
        if (one_a) return launch_wgrad<256, 64, 64, 1, 2>(tmDY, tmX, p, grid, s);
        if (stage_rows == 64) {
          // several waves of CTAs and no split-K: two co-resident CTAs per SM (2 x 2 stages, 2 x 256 TMEM columns) hide
          // each other's prologue and reduction epilogue: 780 -> 697 us on decoder_4.conv.0 (OCTAVE_WGRAD_2CTA=0: off)
          static const int two = [] { const char* e = getenv("OCTAVE_WGRAD_2CTA"); return e ? atoi(e) : 0; }();
          if (two && split == 1 && out_tiles >= 2 * sms) return launch_wgrad<256, 64, 64, 2, 2, 64>(tmDY, tmX, p, grid, s);
          return launch_wgrad<256, 64, 64, 2, 4, 64>(tmDY, tmX, p, grid, s);
        }
        return launch_wgrad<256, 64, 64, 2, 2>(tmDY, tmX, p, grid, s);
      case 128: return one_a ? launch_wgrad<128, 64, 64, 1, 4>(tmDY, tmX, p, grid, s) : launch_wgrad<128, 64, 64, 2, 3>(tmDY, tmX, p, grid, s);
      default: return one_a ? launch_wgrad<64, 64, 64, 1, 6>(tmDY, tmX, p, grid, s) : launch_wgrad<64, 64, 64, 2, 4>(tmDY, tmX, p, grid, s);
    }
  }
  if (CWA == 32 && CWB == 64)
    return one_a ? launch_wgrad<64, 32, 64, 1, 8>(tmDY, tmX, p, grid, s) : launch_wgrad<64, 32, 64, 4, 4>(tmDY, tmX, p, grid, s);
  if (CWA == 64 && CWB == 32)
    return one_a ? launch_wgrad<32, 64, 32, 1, 8>(tmDY, tmX, p, grid, s) : launch_wgrad<32, 64, 32, 2, 5>(tmDY, tmX, p, grid, s);
  return one_a ? launch_wgrad<32, 32, 32, 1, 12>(tmDY, tmX, p, grid, s) : launch_wgrad<32, 32, 32, 4, 5>(tmDY, tmX, p, grid, s);
}
