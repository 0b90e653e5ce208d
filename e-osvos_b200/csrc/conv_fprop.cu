// Persistent tcgen05 implicit-GEMM forward kernel: every conv / Linear / deconv forward and every data
// gradient of the e-OSVOS hot path (reference call site src/networks/mask_rcnn.py:716 and,
// for dgrad, src/meta_optim/meta_optim.py:202-204).
//
// One CTA per SM slot loops over output tiles (tile = blockIdx.x + i * gridDim.x, column tile fastest so CTAs
// sharing an A tile run together).  Three pipelines overlap across tiles:
//   TMA producer warp  -> smem ring (STAGES x {A 128x64, B BNx64}, SWIZZLE_128B, K-major)   full/empty mbarriers
//   MMA issuer warp    -> tcgen05.mma kind::f16 M=128 N=BN into one of TWO TMEM accumulators  tmem_full/empty
//   4 epilogue warps   -> tcgen05.ld, bias / residual / ReLU / GroupNorm partial statistics, 128-bit stores
// so the epilogue of tile i runs under the loads and MMAs of tile i+1 (the HBM-bound 1x1 layers have a single
// K block per tile and would otherwise pay load latency + MMA + epilogue serially per tile).
#include "conv_gemm.cuh"
#include "ptx.cuh"
#include "common.h"

// Timing-only diagnostic builds (tools/diag_kblock.py; never the product library): bit 0 drops the A-operand TMA loads,
// bit 1 the B-operand loads, bit 2 the MMAs of the single-CTA kernel, to see which of them bounds a K block.
#ifndef EOSVOS_DIAG
#define EOSVOS_DIAG 0
#endif

namespace eosvos {

// DEEP: launches whose tiles fit in ONE wave of one CTA per SM (most layers at batch 1, layer3/4 + the small pyramid
// levels at batch 3).  A CTA then sees a single tile, so nothing overlaps its K loop except its own ring: measured, a
// 36-block K loop takes ~18 us whatever the tile (0.5 us per 64-wide K block = load latency / ring depth).  These
// launches get the whole shared memory of the SM as ONE deep ring (8 / 6 / 4 stages) instead of two CTAs with 4 / 3.
template <int BN, bool DEEP = false>
struct FpropCfg {
  static constexpr int STAGES = DEEP ? (BN == 256 ? 4 : (BN == 128 ? 6 : 8)) : (BN == 256 ? 4 : (BN == 128 ? 3 : 4));
  static constexpr int CTAS_PER_SM = (DEEP || BN == 256) ? 1 : 2;
  static constexpr uint32_t TMEM_COLS = 2 * BN;      // two accumulators: 128, 256 or 512 columns
  static constexpr int SMEM = 1024 + STAGES * (128 * 128 + BN * 128) + 256;
};

// GroupNorm partial statistics are accumulated per CTA in shared memory across ALL of its tiles and flushed with
// one global atomic per (image, group) at the end: per-tile global atomics on the 64 * N hot addresses serialise
// in L2 (measured: 1.07 ms instead of ~60 us for the stem GEMM).
constexpr int GN_MAXN = 8;   // images kept in the shared accumulator; larger batches fall back to global atomics

// Output coordinates of one accumulator row (shared by the 1-CTA and the CTA-pair kernels).
struct RowCoord {
  bool valid;
  long long off, roff;
  int gn_n, n_first;
  bool all_same;
};

__device__ __forceinline__ RowCoord fprop_row_coord(const FpropParams& p, int mt, int r) {
  RowCoord rc;
  int rr = r;
  bool valid = true;
  long long off = 0, roff = 0;
  int gn_n = 0;
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    const int td = mt % p.ntiles[d];
    mt /= p.ntiles[d];
    const int i = rr % p.rows_box[d];
    rr /= p.rows_box[d];
    const int coord = td * p.rows_box[d] + i;
    valid = valid && (coord < p.odim[d]);
    off += (long long)coord * p.ostride[d];
    roff += (long long)(coord >> p.rshift[d]) * p.rstride[d];
    if (d == p.gn_dim) gn_n = coord;
  }
  valid = valid && (rr == 0) && (mt == 0);
  rc.valid = valid;
  rc.off = off;
  rc.roff = roff;
  rc.gn_n = gn_n;
  rc.n_first = 0;
  rc.all_same = true;
  if (p.gn_sum) {
    rc.n_first = __shfl_sync(0xffffffffu, gn_n, 0);
    rc.all_same = __all_sync(0xffffffffu, (gn_n == rc.n_first) || !valid);
  }
  return rc;
}

// Epilogue of one 32-row x BN-column accumulator slice held by this warp: tcgen05.ld, bias / residual / ReLU /
// GroupNorm partial statistics, 128-bit stores.  `acc` = TMEM address of the warp's first lane and column.
template <int BN>
__device__ __forceinline__ void fprop_epilogue_tile(const FpropParams& p, const RowCoord& rc, int n0, uint32_t acc,
                                                    int lane, float* gn_acc) {
  const bool valid = rc.valid;
  const long long off = rc.off, roff = rc.roff;
  const int gn_n = rc.gn_n, n_first = rc.n_first;
  const bool all_same = rc.all_same;
#pragma unroll 1
  for (int c = 0; c < BN; c += 32) {
    const int col0 = n0 + c;
    if (col0 >= p.n_valid) break;  // uniform across the CTA
    uint32_t v[32];
    tmem_ld_32x32(acc + (uint32_t)c, v);
    tmem_ld_wait();
    long long o = off + col0;
    if (p.ogroup) {
      const int g = col0 / p.ogroup;
      o = off + p.ogroup_off[g] + (col0 - g * p.ogroup);
    }
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
    if (p.bias) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.n_valid) f[j] += __ldg(p.bias + col0 + j);
    }
    if (p.gn_sum) {
      // GroupNorm partial statistics on the fp32 accumulators (pre-rounding).  The warp holds a 32-row x
      // 32-column block (lane = row).  A butterfly "transpose-reduce" (16+8+4+2+1 shuffles) leaves lane l with
      // the sum over all 32 rows of column l; lanes of one group are then combined (cpg is a power of two) and
      // one lane per group adds into the CTA's shared accumulator.
      if (all_same) {
        float a1[32], a2[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = valid ? f[j] : 0.f;
          a1[j] = x;
          a2[j] = x * x;
        }
#pragma unroll
        for (int offx = 16, n = 32; offx >= 1; offx >>= 1, n >>= 1) {
          const bool upper = (lane & offx) != 0;
#pragma unroll
          for (int i = 0; i < n / 2; ++i) {
            const float s1 = upper ? a1[i] : a1[i + n / 2];
            const float k1 = upper ? a1[i + n / 2] : a1[i];
            a1[i] = k1 + __shfl_xor_sync(0xffffffffu, s1, offx);
            const float s2 = upper ? a2[i] : a2[i + n / 2];
            const float k2 = upper ? a2[i + n / 2] : a2[i];
            a2[i] = k2 + __shfl_xor_sync(0xffffffffu, s2, offx);
          }
        }
        float c1 = a1[0], c2 = a2[0];            // column (col0 + lane): sums over the warp's 32 rows
        const int gl = p.gn_cpg < 32 ? p.gn_cpg : 32;   // lanes per group inside this chunk
        for (int m = 1; m < gl; m <<= 1) {
          c1 += __shfl_xor_sync(0xffffffffu, c1, m);
          c2 += __shfl_xor_sync(0xffffffffu, c2, m);
        }
        if ((lane & (gl - 1)) == 0 && col0 + lane < p.n_valid && (c1 != 0.f || c2 != 0.f)) {
          const int grp = (col0 + lane) / p.gn_cpg;
          float* d = gn_acc ? gn_acc + (n_first * 32 + grp) * 2 : p.gn_sum + ((long long)n_first * 32 + grp) * 2;
          atomicAdd(d, c1);
          atomicAdd(d + 1, c2);
        }
      } else {
        // rows of this warp straddle two images (flat tiles only): slow per-lane path
        const int cmask = p.gn_cpg - 1;
        float s1 = 0.f, s2 = 0.f;
        for (int j = 0; j < 32; ++j) {
          const float x = valid ? f[j] : 0.f;
          s1 += x;
          s2 += x * x;
          if ((((j + 1) & cmask) == 0) || j == 31) {
            if (valid && col0 + j < p.n_valid) {
              const int grp = (col0 + j) / p.gn_cpg;
              float* d = gn_acc ? gn_acc + (gn_n * 32 + grp) * 2 : p.gn_sum + ((long long)gn_n * 32 + grp) * 2;
              atomicAdd(d, s1);
              atomicAdd(d + 1, s2);
            }
            s1 = 0.f;
            s2 = 0.f;
          }
        }
      }
    }
    if (valid) {
      if (p.res) {
        const act_t* rp = p.res + roff + col0;
        uint4 rv[4];
        if ((reinterpret_cast<uintptr_t>(rp) & 31) == 0 && col0 + 32 <= p.n_valid) {
          ld_global_nc_256(rp, rv[0], rv[1]);
          ld_global_nc_256(rp + 16, rv[2], rv[3]);
        } else {
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8)
            rv[j8] = (col0 + j8 * 8 < p.n_valid) ? __ldg(reinterpret_cast<const uint4*>(rp + j8 * 8))
                                                 : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          const act2_t* rh = reinterpret_cast<const act2_t*>(&rv[j8]);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 rf = act22float2(rh[k]);
            f[j8 * 8 + 2 * k] += rf.x;
            f[j8 * 8 + 2 * k + 1] += rf.y;
          }
        }
      }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
      }
      if (p.out_fp32) {
        float* op = reinterpret_cast<float*>(p.out) + o;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          if (col0 + j4 * 4 < p.n_valid)
            *reinterpret_cast<float4*>(op + j4 * 4) =
                make_float4(f[j4 * 4], f[j4 * 4 + 1], f[j4 * 4 + 2], f[j4 * 4 + 3]);
        }
      } else {
        act_t* op = reinterpret_cast<act_t*>(p.out) + o;
        uint4 w[4];
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          w[j8].x = pack_act2(f[j8 * 8 + 0], f[j8 * 8 + 1]);
          w[j8].y = pack_act2(f[j8 * 8 + 2], f[j8 * 8 + 3]);
          w[j8].z = pack_act2(f[j8 * 8 + 4], f[j8 * 8 + 5]);
          w[j8].w = pack_act2(f[j8 * 8 + 6], f[j8 * 8 + 7]);
        }
        // each lane owns 64 contiguous bytes of its output row: two 256-bit stores (whole 32-byte sectors) when the
        // row is 32-byte aligned, else four 128-bit ones
        if ((reinterpret_cast<uintptr_t>(op) & 31) == 0 && col0 + 32 <= p.n_valid) {
          st_global_256(op, w[0], w[1]);
          st_global_256(op + 16, w[2], w[3]);
        } else {
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8)
            if (col0 + j8 * 8 < p.n_valid) *reinterpret_cast<uint4*>(op + j8 * 8) = w[j8];
        }
      }
    }
  }
}

template <int BN, bool DEEP>
__global__ void __launch_bounds__(192, FpropCfg<BN, DEEP>::CTAS_PER_SM)
conv_fprop_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const FpropParams p) {
  constexpr int STAGES = FpropCfg<BN, DEEP>::STAGES;
  constexpr int A_STAGE = 128 * 128;
  constexpr int B_STAGE = BN * 128;
  constexpr uint32_t TMEM_COLS = FpropCfg<BN, DEEP>::TMEM_COLS;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_STAGE);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;    // [2]
  uint64_t* tempty = tfull + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  __shared__ float gn_acc_s[GN_MAXN * 32 * 2];
  float* gn_acc = (p.gn_sum && p.gn_nimg <= GN_MAXN) ? gn_acc_s : nullptr;
  if (gn_acc)
    for (int i = threadIdx.x; i < p.gn_nimg * 64; i += blockDim.x) gn_acc_s[i] = 0.f;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles_n;
  const int num_it = p.num_taps * p.kchunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 4);   // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // everything above is private to the CTA: it overlaps the tail of the previous kernel; from here on global memory
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const uint32_t tx = ((EOSVOS_DIAG & 1) ? 0u : (uint32_t)p.a_bytes) + ((EOSVOS_DIAG & 2) ? 0u : (uint32_t)B_STAGE);
      int gi = 0;  // global k-block counter across tiles -> ring stage / phase
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n0 = (tile % p.n_tiles_n) * BN;
        int mt = tile / p.n_tiles_n;
        int base[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          base[d] = (mt % p.ntiles[d]) * p.a_tile_step[d];
          mt /= p.ntiles[d];
        }
        for (int it = 0; it < num_it; ++it, ++gi) {
          const int s = gi % STAGES;
          const uint32_t ph = (uint32_t)(gi / STAGES) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          mbar_arrive_expect_tx(&full[s], tx);
          const int tap = it / p.kchunks;
          const int kc = it - tap * p.kchunks;
          if (!(EOSVOS_DIAG & 1))
            tma_load_5d(sA + s * A_STAGE, &tmA, &full[s], p.tap_delta[tap][0] + kc * 64, base[0] + p.tap_delta[tap][1],
                        base[1] + p.tap_delta[tap][2], base[2] + p.tap_delta[tap][3], base[3] + p.tap_delta[tap][4]);
          if (!(EOSVOS_DIAG & 2)) tma_load_2d(sB + s * B_STAGE, &tmB, &full[s], p.tap_bk[tap] + kc * 64, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_act(128, BN, 0, 0);
      int gi = 0, lt = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
        const int buf = lt & 1;
        mbar_wait(&tempty[buf], ((uint32_t)(lt >> 1) & 1u) ^ 1u);   // epilogue has drained this accumulator
        tc_fence_after_sync();
        const uint32_t acc = tmem_base + (uint32_t)(buf * BN);
        for (int it = 0; it < num_it; ++it, ++gi) {
          const int s = gi % STAGES;
          const uint32_t ph = (uint32_t)(gi / STAGES) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(sA + s * A_STAGE);
          const uint32_t b_addr = smem_u32(sB + s * B_STAGE);
#pragma unroll
          for (int k = 0; k < ((EOSVOS_DIAG & 4) ? 0 : 4); ++k) {
            const uint64_t ad = make_smem_desc_sw128(a_addr + k * 32, 0, 1024);
            const uint64_t bd = make_smem_desc_sw128(b_addr + k * 32, 0, 1024);
            umma_f16kind(acc, ad, bd, idesc, (uint32_t)((it | k) != 0));
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&tfull[buf]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    int lt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      const int n0 = (tile % p.n_tiles_n) * BN;
      const RowCoord rc = fprop_row_coord(p, tile / p.n_tiles_n, r);
      mbar_wait(&tfull[buf], (uint32_t)(lt >> 1) & 1u);
      tc_fence_after_sync();
      fprop_epilogue_tile<BN>(p, rc, n0, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN), lane, gn_acc);
      // hand the accumulator back to the MMA warp
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
    }
    if (gn_acc) {
      // all four epilogue warps are done with every tile: publish this CTA's partial statistics
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int i = (warp - 2) * 32 + lane; i < p.gn_nimg * 64; i += 128) {
        const float v = gn_acc_s[i];
        if (v != 0.f) atomicAdd(p.gn_sum + i, v);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// =============================================================================================
// CTA-pair variant (cta_group::2) for the tensor-bound layers (Cout % 256 == 0, many tiles).
//
// A cluster of two CTAs (one TPC) owns a 256-pixel x 256-channel output tile.  Each CTA stages its own 128 pixel
// rows of A and its own 128-channel HALF of B; the leader issues one M=256, N=256 tcgen05.mma per 16-deep k slice
// that reads both CTAs' shared memory and writes each CTA's 128 x 256 accumulator half into that CTA's TMEM.  Per SM
// and k block this moves 32 KB (16 A + 16 B) instead of 48 KB, which is what bounds the single-CTA kernel (ncu: tensor
// pipe 74 %, SM clock power-capped).  The smaller stage also buys a 6-deep ring.  Epilogues are per CTA and identical
// to the single-CTA kernel's.
//   full[s]   (leader)    : leader arrive.expect_tx(both CTAs' bytes) + peer's remote arrive; all TMA loads of the pair
//   empty[s]  (both)      : multicast tcgen05.commit
//   tfull[b]  (both)      : multicast tcgen05.commit after the last k block of a tile
//   tempty[b] (leader)    : 4 epilogue warps x 2 CTAs (peer arrives remotely)
// =============================================================================================
constexpr int PAIR_STAGES = 6;
constexpr int PAIR_BN = 256;
constexpr int PAIR_SMEM = 1024 + PAIR_STAGES * (128 * 128 + 128 * 128) + 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
conv_fprop_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const FpropParams p) {
  constexpr int STAGES = PAIR_STAGES;
  constexpr int BN = PAIR_BN;
  constexpr int A_STAGE = 128 * 128;
  constexpr int B_STAGE = 128 * 128;        // this CTA's 128-channel half
  constexpr uint32_t TMEM_COLS = 512;       // two 256-column accumulators

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_STAGE);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;    // [2]
  uint64_t* tempty = tfull + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  __shared__ float gn_acc_s[GN_MAXN * 32 * 2];
  float* gn_acc = (p.gn_sum && p.gn_nimg <= GN_MAXN) ? gn_acc_s : nullptr;
  if (gn_acc)
    for (int i = threadIdx.x; i < p.gn_nimg * 64; i += blockDim.x) gn_acc_s[i] = 0.f;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int m_pairs = (p.m_tiles + 1) >> 1;
  const int total_pt = m_pairs * p.n_tiles_n;     // pair tiles
  const int num_it = p.num_taps * p.kchunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 2);    // leader's expect_tx arrive + peer's remote arrive (only the leader's copy is used)
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 8);  // 4 epilogue warps of each CTA (only the leader's copy is used)
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, TMEM_COLS);
  tc_fence_before_sync();
  cluster_sync_all();            // barrier inits + TMEM allocation visible to both CTAs
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      const uint32_t tx = 2u * ((uint32_t)p.a_bytes + (uint32_t)B_STAGE);
      int gi = 0;
      for (int pt = pair; pt < total_pt; pt += num_pairs) {
        const int n0 = (pt % p.n_tiles_n) * BN + (int)rank * 128;
        int mt = (pt / p.n_tiles_n) * 2 + (int)rank;
        int base[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          base[d] = (mt % p.ntiles[d]) * p.a_tile_step[d];
          mt /= p.ntiles[d];
        }
        // an odd trailing M tile has no partner: push its coordinates out of range (TMA zero-fills, nothing stored)
        if (mt != 0) base[3] = p.ntiles[3] * p.a_tile_step[3] + 1;
        for (int it = 0; it < num_it; ++it, ++gi) {
          const int s = gi % STAGES;
          const uint32_t ph = (uint32_t)(gi / STAGES) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          const uint32_t fb = mapa_shared(smem_u32(&full[s]), 0);
          if (rank == 0) mbar_arrive_expect_tx(&full[s], tx);
          const int tap = it / p.kchunks;
          const int kc = it - tap * p.kchunks;
          tma_load_5d_pair(sA + s * A_STAGE, &tmA, fb, p.tap_delta[tap][0] + kc * 64, base[0] + p.tap_delta[tap][1],
                           base[1] + p.tap_delta[tap][2], base[2] + p.tap_delta[tap][3],
                           base[3] + p.tap_delta[tap][4]);
          tma_load_2d_pair(sB + s * B_STAGE, &tmB, fb, p.tap_bk[tap] + kc * 64, n0);
          if (rank != 0) mbar_arrive_cluster(fb);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_act(256, BN, 0, 0);
      int gi = 0, lt = 0;
      for (int pt = pair; pt < total_pt; pt += num_pairs, ++lt) {
        const int buf = lt & 1;
        mbar_wait(&tempty[buf], ((uint32_t)(lt >> 1) & 1u) ^ 1u);   // both epilogues drained this accumulator
        tc_fence_after_sync();
        const uint32_t acc = tmem_base + (uint32_t)(buf * BN);
        for (int it = 0; it < num_it; ++it, ++gi) {
          const int s = gi % STAGES;
          const uint32_t ph = (uint32_t)(gi / STAGES) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(sA + s * A_STAGE);
          const uint32_t b_addr = smem_u32(sB + s * B_STAGE);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = make_smem_desc_sw128(a_addr + k * 32, 0, 1024);
            const uint64_t bd = make_smem_desc_sw128(b_addr + k * 32, 0, 1024);
            umma_f16kind_pair(acc, ad, bd, idesc, (uint32_t)((it | k) != 0));
          }
          umma_commit_pair(&empty[s]);
        }
        umma_commit_pair(&tfull[buf]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5, both CTAs)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    int lt = 0;
    for (int pt = pair; pt < total_pt; pt += num_pairs, ++lt) {
      const int buf = lt & 1;
      const int n0 = (pt % p.n_tiles_n) * BN;
      const RowCoord rc = fprop_row_coord(p, (pt / p.n_tiles_n) * 2 + (int)rank, r);
      mbar_wait(&tfull[buf], (uint32_t)(lt >> 1) & 1u);
      tc_fence_after_sync();
      fprop_epilogue_tile<BN>(p, rc, n0, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN), lane, gn_acc);
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty[buf]), 0));
    }
    if (gn_acc) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int i = (warp - 2) * 32 + lane; i < p.gn_nimg * 64; i += 128) {
        const float v = gn_acc_s[i];
        if (v != 0.f) atomicAdd(p.gn_sum + i, v);
      }
    }
  }
  // neither CTA may exit (or free TMEM) while its partner can still signal its barriers or read its shared memory
  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

static int launch_fprop_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const FpropParams& p,
                             cudaStream_t stream) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e =
        cudaFuncSetAttribute(conv_fprop_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(conv_fprop_pair)");
    attr_done = true;
  }
  const long long total = (long long)((p.m_tiles + 1) / 2) * p.n_tiles_n;
  const long long pairs = num_sms() / 2;
  dim3 grid((unsigned)(2 * (total < pairs ? total : pairs)));
  conv_fprop_pair_kernel<<<grid, 192, PAIR_SMEM, stream>>>(tmA, tmB, p);
  return check_launch("conv_fprop_pair_kernel");
}

template <int BN, bool DEEP>
static int launch_fprop_td(const CUtensorMap& tmA, const CUtensorMap& tmB, const FpropParams& p, cudaStream_t stream) {
  constexpr int SMEM = FpropCfg<BN, DEEP>::SMEM;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_fprop_kernel<BN, DEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(conv_fprop)");
    attr_done = true;
  }
  const long long total = (long long)p.m_tiles * p.n_tiles_n;
  const long long slots = (long long)num_sms() * FpropCfg<BN, DEEP>::CTAS_PER_SM;
  dim3 grid((unsigned)(total < slots ? total : slots));
  cudaError_t le = launch_pdl(conv_fprop_kernel<BN, DEEP>, grid, dim3(192), SMEM, stream, tmA, tmB, p);
  if (le != cudaSuccess) return set_cuda_error(le, "conv_fprop_kernel launch");
  return check_launch("conv_fprop_kernel");
}

template <int BN>
static int launch_fprop_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const FpropParams& p, cudaStream_t stream) {
  static const bool deep_ok = [] {
    const char* e = getenv("EOSVOS_FPROP_DEEP");
    return !(e && e[0] == '0');
  }();
  const long long total = (long long)p.m_tiles * p.n_tiles_n;
  if (deep_ok && BN != 256 && total <= (long long)num_sms()) return launch_fprop_td<BN, true>(tmA, tmB, p, stream);
  return launch_fprop_td<BN, false>(tmA, tmB, p, stream);
}

int launch_fprop(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const FpropParams& p, int m_tiles,
                 cudaStream_t stream) {
  (void)m_tiles;
  switch (bn) {
    case 512: return launch_fprop_pair(tmA, tmB, p, stream);   // CTA pair: 256 x 256 tile, B box = 128 rows
    case 64: return launch_fprop_t<64>(tmA, tmB, p, stream);
    case 128: return launch_fprop_t<128>(tmA, tmB, p, stream);
    case 256: return launch_fprop_t<256>(tmA, tmB, p, stream);
  }
  return set_error(EOSVOS_ERR_ARG, "launch_fprop: unsupported BN");
}

}  // namespace eosvos
