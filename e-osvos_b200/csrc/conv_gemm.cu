// tcgen05 / TMEM / TMA implicit-GEMM kernels for every dense contraction on the e-OSVOS hot path
// (SURVEY.md §2.2 K1): conv fprop, dgrad (fprop on dY with transposed weights), wgrad, Linear,
// 2x2 deconv.  They replace the cuDNN calls the reference reaches through
// torchvision's GeneralizedRCNN.forward (reference: src/networks/mask_rcnn.py:716).
//
// Data layout: activations NHWC bf16, weights [Cout][taps][Cin] bf16 (K-major), fp32 accumulate
// in TMEM.  One CTA = one 128-row x BN-column output tile.  Warp roles: warp 0 TMA producer,
// warp 1 MMA issuer + TMEM owner, warps 2..5 epilogue (one TMEM lane quadrant each).
#include "conv_gemm.cuh"
#include "ptx.cuh"
#include "common.h"

namespace eosvos {

// =============================================================================================
// wgrad: dW[m, n] += sum over pixel tiles of A[pix, m] * B[pix, n]; both operands MN-major SW128.
// A = dY boxes (2 x 64 channels), B = X boxes (BN/64 x 64 channels), K = pixels (<= 64 per stage,
// zero-padded to a multiple of 16).  Split over pixel tiles (grid.x), fp32 red.add epilogue.
// =============================================================================================
template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const WgradParams p) {
  constexpr int SLOT = 64 * 128;  // one [64 pixels][64 channels] bf16 box
  constexpr int A_STAGE = 2 * SLOT;
  constexpr int NB = BN / 64;
  constexpr int B_STAGE = NB * SLOT;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_STAGE);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tap = blockIdx.z;
  const int n_tile = blockIdx.y % p.n_tiles_n;
  const int m_tile = blockIdx.y / p.n_tiles_n;
  const int m0 = m_tile * 128, n0 = n_tile * BN;
  const int tile_begin = blockIdx.x * p.tiles_per_split;
  int tile_end = tile_begin + p.tiles_per_split;
  if (tile_end > p.total_tiles) tile_end = p.total_tiles;
  const int num_it = tile_end > tile_begin ? tile_end - tile_begin : 0;

  // rows [rows, kpad) of every box slot are never written by TMA: zero them once so the padded
  // K range contributes exactly 0.
  if (p.kpad > p.rows) {
    const int nz = (p.kpad - p.rows) * 128 / 16;
    for (int s = 0; s < STAGES; ++s) {
      for (int b = 0; b < 2 + NB; ++b) {
        uint8_t* base = (b < 2 ? sA + s * A_STAGE + b * SLOT : sB + s * B_STAGE + (b - 2) * SLOT) + p.rows * 128;
        for (int i = threadIdx.x; i < nz; i += blockDim.x) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
      }
    }
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (num_it > 0) {
    if (warp == 0) {
      if (lane == 0) {
        const uint32_t tx = (uint32_t)(2 + NB) * (uint32_t)p.rows * 128u;
        for (int it = 0; it < num_it; ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          mbar_arrive_expect_tx(&full[s], tx);
          int tt = tile_begin + it;
          int ca[4], cb[4];
#pragma unroll
          for (int d = 0; d < 4; ++d) {
            const int td = tt % p.ntiles[d];
            tt /= p.ntiles[d];
            ca[d] = td * p.a_tile_step[d] + p.tap_delta_a[tap][d + 1];
            cb[d] = td * p.b_tile_step[d] + p.tap_delta_b[tap][d + 1];
          }
#pragma unroll
          for (int b = 0; b < 2; ++b)
            tma_load_5d(sA + s * A_STAGE + b * SLOT, &tmA, &full[s], p.tap_delta_a[tap][0] + m0 + b * 64, ca[0],
                        ca[1], ca[2], ca[3]);
#pragma unroll
          for (int b = 0; b < NB; ++b)
            tma_load_5d(sB + s * B_STAGE + b * SLOT, &tmB, &full[s], p.tap_delta_b[tap][0] + n0 + b * 64, cb[0],
                        cb[1], cb[2], cb[3]);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = make_idesc_act(128, BN, 1, 1);
        const int ksteps = p.kpad >> 4;
        for (int it = 0; it < num_it; ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(sA + s * A_STAGE);
          const uint32_t b_addr = smem_u32(sB + s * B_STAGE);
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t ad = make_smem_desc_sw128(a_addr + k * 2048, SLOT, 1024);
            const uint64_t bd = make_smem_desc_sw128(b_addr + k * 2048, SLOT, 1024);
            umma_f16kind(tmem_base, ad, bd, idesc, (uint32_t)((it | k) != 0));
          }
          umma_commit(&empty[s]);
        }
        umma_commit(tmem_full);
      }
    } else {
      const int q = warp & 3;
      const int m = m0 + q * 32 + lane;
      mbar_wait(tmem_full, 0);
      tc_fence_after_sync();
      float* dst = p.dw + (long long)m * p.dw_m_stride + (long long)tap * p.dw_tap_stride;
      // columns contiguous in the destination (1x1 convs, Linear, channels-last filter gradients): one 16-byte
      // vector RED per 4 columns instead of 4 scalar ones -- the scattered scalar REDs, not the MMAs, bound the
      // short-reduction layers (layer3/4: 6..24 k-blocks per CTA against 32768 REDs)
      const bool vec = p.n_inner_stride == 1 && p.n_inner >= p.n_valid && (p.n_valid & 3) == 0 &&
                       (((p.dw_m_stride | p.dw_tap_stride) & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.dw) & 15) == 0);
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        if (n0 + c >= p.n_valid) break;  // uniform
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
        tmem_ld_wait();
        if (m < p.m_valid) {
          if (vec) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const int n = n0 + c + j;
              if (n < p.n_valid)
                red_add_v4(dst + n, __uint_as_float(v[j]) * p.alpha, __uint_as_float(v[j + 1]) * p.alpha,
                           __uint_as_float(v[j + 2]) * p.alpha, __uint_as_float(v[j + 3]) * p.alpha);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = n0 + c + j;
              if (n < p.n_valid) {
                const int no = n / p.n_inner;
                atomicAdd(dst + (long long)no * p.n_outer_stride + (long long)(n - no * p.n_inner) * p.n_inner_stride,
                          __uint_as_float(v[j]) * p.alpha);
              }
            }
          }
        }
      }
      tc_fence_before_sync();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// =============================================================================================
// host side
// =============================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// dims/box/estride are given innermost-first; strides_bytes has rank-1 entries (dims 1..rank-1).
int make_tensor_map_act(CUtensorMap* m, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estride) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(EOSVOS_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = estride ? estride[i] : 1;
  }
  for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
  CUresult r = fn(m, EOSVOS_TMA_DTYPE, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf,
             "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u]", (int)r,
             rank, (unsigned long long)gd[0], (unsigned long long)(rank > 1 ? gd[1] : 0),
             (unsigned long long)(rank > 2 ? gd[2] : 0), (unsigned long long)(rank > 3 ? gd[3] : 0),
             (unsigned long long)(rank > 4 ? gd[4] : 0), bx[0], rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0,
             rank > 3 ? bx[3] : 0, rank > 4 ? bx[4] : 0);
    return set_error(EOSVOS_ERR_CUDA, buf);
  }
  return 0;
}

template <int BN, int STAGES>
static int launch_wgrad_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const WgradParams& p, dim3 grid,
                          cudaStream_t stream) {
  constexpr int SMEM = 1024 + STAGES * (2 + BN / 64) * 8192 + 256;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(conv_wgrad)");
    attr_done = true;
  }
  conv_wgrad_kernel<BN, STAGES><<<grid, 192, SMEM, stream>>>(tmA, tmB, p);
  return check_launch("conv_wgrad_kernel");
}

int launch_wgrad(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const WgradParams& p, dim3 grid,
                 cudaStream_t stream) {
  switch (bn) {
    // 64-channel inputs: a 64-pixel K step carries a quarter of the MMA work of the wide tiles, so the CTA is bound by
    // the latency of its own load -> MMA -> commit chain; three shallow CTAs per SM (3 stages, 75 KB) hide it
    case 63: return launch_wgrad_t<64, 3>(tmA, tmB, p, grid, stream);
    case 64: return launch_wgrad_t<64, 6>(tmA, tmB, p, grid, stream);
    case 128: return launch_wgrad_t<128, 5>(tmA, tmB, p, grid, stream);
    case 256: return launch_wgrad_t<256, 4>(tmA, tmB, p, grid, stream);
  }
  return set_error(EOSVOS_ERR_ARG, "launch_wgrad: unsupported BN");
}

}  // namespace eosvos
