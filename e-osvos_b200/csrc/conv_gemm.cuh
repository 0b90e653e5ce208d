// Device-side parameter blocks for the two tcgen05 implicit-GEMM kernels (conv_gemm.cu).
#pragma once
#include <stdint.h>
#include "act.cuh"

namespace eosvos {

constexpr int MAX_TAPS = 9;

// D[rows, cols] = sum_taps A_tap[rows, K] * B_tap[cols, K]^T   (both K-major, bf16, fp32 accumulate)
// "rows" are output pixels arranged as a 4-D box (b1 x b2 x b3 x b4 <= 128) of the A tensor view.
struct FpropParams {
  int rows_box[4];      // logical rows per tiled dim (dims 1..4 of the A view)
  int ntiles[4];        // tiles per dim
  int a_tile_step[4];   // A-coordinate step per tile (rows_box * conv stride)
  int num_taps, kchunks;          // K loop = num_taps * kchunks blocks of 64
  int tap_delta[MAX_TAPS][5];     // per-tap A coordinate offsets (dim 0 = channel offset)
  int tap_bk[MAX_TAPS];           // per-tap B K-coordinate base
  int a_bytes;                    // bytes one A box deposits (rows * 128)
  int n_tiles_n;                  // column tiles
  int m_tiles;                    // row tiles (product of ntiles[])
  // epilogue
  void* out;
  const float* bias;
  int relu, out_fp32, n_valid;
  long long ostride[4];
  int odim[4];
  int ogroup;                     // 0, or columns per output group (deconv sub-pixel groups)
  long long ogroup_off[4];
  const act_t* res;       // optional residual: out = act(acc + bias + res[(coord >> rshift) . rstride + col])
  long long rstride[4];
  int rshift[4];
  float* gn_sum;                  // optional GroupNorm partial statistics: [gn_n][32][2] fp32 (sum, sumsq)
  int gn_cpg;                     // channels per group
  int gn_dim;                     // which tiled dim (0..3) indexes the image n
  int gn_nimg;                    // number of images (rows of gn_sum)
};

// dW[m, n] (+)= sum_pixels A[pixel, m] * B[pixel, n]   (both MN-major), split over pixel tiles.
struct WgradParams {
  int rows_box[4];
  int ntiles[4];
  int a_tile_step[4];
  int b_tile_step[4];
  int num_taps;
  int tap_delta_a[MAX_TAPS][5];
  int tap_delta_b[MAX_TAPS][5];
  int rows, kpad;
  int total_tiles, tiles_per_split;
  int m_valid, n_valid;
  int n_tiles_n;
  float* dw;
  // dw[m * m_stride + tap * tap_stride + (n / n_inner) * n_outer_stride + (n % n_inner) * n_inner_stride]
  long long dw_m_stride, dw_tap_stride;
  int n_inner;
  long long n_inner_stride, n_outer_stride;
  float alpha;   // every accumulated value is multiplied by alpha (1 / loss-scale of the fp16 backward)
};

}  // namespace eosvos
