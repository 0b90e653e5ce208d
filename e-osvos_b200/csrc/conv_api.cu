// Host side of the dense-contraction entry points (C ABI declared in include/eosvos_b200.h).
// Each op is expressed as one or more launches of the two generic tcgen05 kernels in
// conv_gemm.cu by describing (a) a rank-5 TMA view of the activation operand, (b) the list of
// filter taps as coordinate offsets into that view and (c) the output addressing.
//
// Reference call sites these replace (all reach cuDNN/ATen through torchvision):
//   conv / linear / deconv forward  : src/networks/mask_rcnn.py:716 (GeneralizedRCNN.forward)
//   their backward (dgrad + wgrad)  : src/meta_optim/meta_optim.py:202-204 (torch.autograd.grad)
#include <algorithm>
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "common.h"
#include "conv_gemm.cuh"
#include "../../include/eosvos_b200.h"

namespace eosvos {
int make_tensor_map_act(CUtensorMap* m, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estride);
int launch_fprop(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const FpropParams& p, int m_tiles,
                 cudaStream_t stream);
int launch_wgrad(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const WgradParams& p, dim3 grid,
                 cudaStream_t stream);

// Pick the pixel-tile rectangle (tw x th <= max_rows) that minimises padded MMA work.
static void choose_tile(int Wo, int Ho, int max_rows, int max_tw, int max_th, bool pad16, int* TW, int* TH) {
  long long best_cost = -1;
  int btw = 1, bth = 1;
  const int tw_hi = std::min(std::min(Wo, max_rows), max_tw);
  for (int tw = 1; tw <= tw_hi; ++tw) {
    const int th_hi = std::min(std::min(max_rows / tw, Ho), max_th);
    for (int th = 1; th <= th_hi; ++th) {
      const long long tiles = (long long)((Wo + tw - 1) / tw) * ((Ho + th - 1) / th);
      const int rows = tw * th;
      const long long cost = tiles * (pad16 ? ((rows + 15) / 16) * 16 + 8 /*per-iteration overhead*/ : 1);
      if (best_cost < 0 || cost < best_cost || (cost == best_cost && tw > btw)) {
        best_cost = cost;
        btw = tw;
        bth = th;
      }
    }
  }
  *TW = btw;
  *TH = bth;
}

struct AView {
  const void* base;
  uint64_t dims[5];
  uint64_t strides[4];   // bytes, dims 1..4
  uint32_t estride[5];
};

struct Epilogue {
  void* out = nullptr;
  const float* bias = nullptr;
  int relu = 0, out_fp32 = 0;
  long long ostride[4] = {0, 0, 0, 0};
  int odim[4] = {1, 1, 1, 1};
  int ogroup = 0;
  long long ogroup_off[4] = {0, 0, 0, 0};
  const void* res = nullptr;
  long long rstride[4] = {0, 0, 0, 0};
  int rshift[4] = {0, 0, 0, 0};
  float* gn_sum = nullptr;
  int gn_cpg = 1, gn_dim = 3;
};

static int pick_bn(int n_cols, long long m_tiles, int bn_hint) {
  if (bn_hint == 64 || bn_hint == 128 || bn_hint == 256) return bn_hint;
  if (bn_hint == 512) return 256;   // CTA-pair kernel forced
  if (n_cols <= 64) return 64;
  const long long sms = num_sms();
  // launches that cannot fill the chip even with 128-column tiles are latency-bound (one tile per CTA): the narrowest
  // tile gives the most CTAs and the deepest ring (conv_fprop.cu FpropCfg<BN, DEEP>); measured 1.1-1.25x on layer3/4,
  // the mask head and the coarse pyramid levels, more at batch 1
  if (m_tiles * ((n_cols + 127) / 128) <= sms && m_tiles * ((n_cols + 63) / 64) <= 2 * sms) return 64;
  if (n_cols <= 128) return 128;
  // wide outputs: 256-column tiles halve A re-reads, but only when the grid still fills the chip
  const long long ctas256 = m_tiles * ((n_cols + 255) / 256);
  if (n_cols % 256 == 0 && ctas256 >= sms) return 256;
  return 128;
}

// Generic fprop-style launch.  extent[d] = number of output rows along A-view dim d+1;
// conv_stride[d] = A coordinate step per output row; sp_w / sp_h = which of those dims are the
// two spatial ones the pixel tile spans (others get box 1).
static int run_fprop(const AView& av, const int extent[4], const int conv_stride[4], int sp_w, int sp_h,
                     int num_taps, const int (*tap_delta)[5], const int* tap_bk, int k_per_tap,
                     const void* b_base, uint64_t b_rows, uint64_t b_k, int n_valid, const Epilogue& ep, int bn_hint,
                     cudaStream_t stream) {
  EOSVOS_REQUIRE(k_per_tap % 64 == 0, "fprop: reduction length per tap must be a multiple of 64");
  EOSVOS_REQUIRE(num_taps >= 1 && num_taps <= MAX_TAPS, "fprop: 1..9 taps supported");
  EOSVOS_REQUIRE(n_valid % 8 == 0, "fprop: output columns must be a multiple of 8");
  FpropParams p;
  memset(&p, 0, sizeof p);
  int tw = 1, th = 1;
  if (sp_w >= 0 && sp_h >= 0) {
    choose_tile(extent[sp_w], extent[sp_h], 128, 256 / conv_stride[sp_w], 256 / conv_stride[sp_h], false, &tw, &th);
  } else if (sp_w >= 0) {
    tw = std::min(128, extent[sp_w]);
  }
  long long m_tiles = 1;
  uint32_t box[5] = {64, 1, 1, 1, 1};
  for (int d = 0; d < 4; ++d) {
    int b = 1;
    if (d == sp_w) b = tw;
    if (d == sp_h) b = th;
    p.rows_box[d] = b;
    p.ntiles[d] = (extent[d] + b - 1) / b;
    p.a_tile_step[d] = b * conv_stride[d];
    box[d + 1] = (uint32_t)(b * conv_stride[d]);
    m_tiles *= p.ntiles[d];
    p.ostride[d] = ep.ostride[d];
    p.odim[d] = ep.odim[d];
    p.rstride[d] = ep.rstride[d];
    p.rshift[d] = ep.rshift[d];
    p.ogroup_off[d] = ep.ogroup_off[d];
  }
  EOSVOS_REQUIRE(m_tiles > 0 && m_tiles < (1LL << 30), "fprop: bad tile count");
  const int rows = p.rows_box[0] * p.rows_box[1] * p.rows_box[2] * p.rows_box[3];
  p.a_bytes = rows * 128;
  p.num_taps = num_taps;
  p.kchunks = k_per_tap / 64;
  for (int t = 0; t < num_taps; ++t) {
    for (int d = 0; d < 5; ++d) p.tap_delta[t][d] = tap_delta[t][d];
    p.tap_bk[t] = tap_bk[t];
  }
  const int bn = pick_bn(n_valid, m_tiles, bn_hint);
  p.n_tiles_n = (n_valid + bn - 1) / bn;
  p.m_tiles = (int)m_tiles;
  p.out = ep.out;
  p.bias = ep.bias;
  p.relu = ep.relu;
  p.out_fp32 = ep.out_fp32;
  p.n_valid = n_valid;
  p.ogroup = ep.ogroup;
  p.res = reinterpret_cast<const act_t*>(ep.res);
  p.gn_sum = ep.gn_sum;
  p.gn_cpg = ep.gn_cpg;
  p.gn_dim = ep.gn_dim;
  p.gn_nimg = ep.gn_sum ? ep.odim[ep.gn_dim] : 0;
  if (ep.gn_sum) EOSVOS_REQUIRE((ep.gn_cpg & (ep.gn_cpg - 1)) == 0, "fprop: GroupNorm channels per group must be a power of two");
  if (ep.ogroup) EOSVOS_REQUIRE(ep.ogroup % bn == 0, "fprop: output group must be a multiple of the column tile");

  CUtensorMap tmA, tmB;
  EOSVOS_TRY(make_tensor_map_act(&tmA, av.base, 5, av.dims, av.strides, box, av.estride));
  const uint64_t bdims[2] = {b_k, b_rows};
  const uint64_t bstr[1] = {b_k * 2};
  // tensor-bound shapes (256-wide column tiles, at least one full wave of CTA pairs) run on CTA pairs: each CTA
  // stages half of B, see conv_fprop_pair_kernel.  bn_hint 512 forces, EOSVOS_FPROP_PAIR=0 disables.
  static const bool pair_ok = [] {
    const char* e = getenv("EOSVOS_FPROP_PAIR");
    return !(e && e[0] == '0');
  }();
  const bool pair = bn == 256 && n_valid % 256 == 0 && rows == 128 &&
                    (bn_hint == 512 || (pair_ok && bn_hint == 0 && p.num_taps * p.kchunks >= 8 &&
                                        m_tiles * p.n_tiles_n >= 4LL * num_sms()));
  EOSVOS_REQUIRE(pair || bn_hint != 512, "fprop: CTA-pair kernel needs Cout % 256 == 0 and full 128-pixel tiles");
  const uint32_t bbox[2] = {64, (uint32_t)(pair ? 128 : bn)};
  EOSVOS_TRY(make_tensor_map_act(&tmB, b_base, 2, bdims, bstr, bbox, nullptr));
  return launch_fprop(pair ? 512 : bn, tmA, tmB, p, (int)m_tiles, stream);
}

static inline AView nhwc_view(const void* base, int N, int H, int W, int C, int estride_hw) {
  AView v;
  v.base = base;
  v.dims[0] = (uint64_t)C;
  v.dims[1] = (uint64_t)W;
  v.dims[2] = (uint64_t)H;
  v.dims[3] = (uint64_t)N;
  v.dims[4] = 1;
  v.strides[0] = (uint64_t)C * 2;
  v.strides[1] = (uint64_t)W * C * 2;
  v.strides[2] = (uint64_t)H * W * C * 2;
  v.strides[3] = (uint64_t)N * H * W * C * 2;
  v.estride[0] = 1;
  v.estride[1] = v.estride[2] = (uint32_t)estride_hw;
  v.estride[3] = v.estride[4] = 1;
  return v;
}

}  // namespace eosvos

using namespace eosvos;

// ---------------------------------------------------------------------------------------------
// conv2d forward.  x [N,H,W,Cin] bf16, w [Cout,KH,KW,Cin] bf16, y [N,Ho,Wo,Cout] bf16|fp32.
// ---------------------------------------------------------------------------------------------
extern "C" int eosvos_conv2d_fprop(const void* x, const void* w, const float* bias, const void* res, void* y,
                                   float* gn_sum, int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride,
                                   int pad, int flags, int bn_hint, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(x && w && y, "conv2d_fprop: null pointer");
  EOSVOS_REQUIRE(Cin % 64 == 0, "conv2d_fprop: Cin must be a multiple of 64");
  EOSVOS_REQUIRE(Cout % 8 == 0, "conv2d_fprop: Cout must be a multiple of 8");
  EOSVOS_REQUIRE(KH * KW <= MAX_TAPS && (stride == 1 || stride == 2), "conv2d_fprop: unsupported filter");
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  Epilogue ep;
  ep.out = y;
  ep.bias = bias;
  ep.relu = flags & EOSVOS_FLAG_RELU;
  ep.out_fp32 = (flags & EOSVOS_FLAG_OUT_FP32) ? 1 : 0;
  ep.gn_sum = gn_sum;
  ep.gn_cpg = Cout / 32 > 0 ? Cout / 32 : 1;
  int taps[MAX_TAPS][5], bk[MAX_TAPS];
  int nt = 0;
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw) {
      taps[nt][0] = 0;
      taps[nt][1] = kw - pad;
      taps[nt][2] = kh - pad;
      taps[nt][3] = 0;
      taps[nt][4] = 0;
      bk[nt] = (kh * KW + kw) * Cin;
      ++nt;
    }
  const bool flat = (KH == 1 && KW == 1 && stride == 1 && pad == 0 && !(flags & EOSVOS_FLAG_RES_HALF));
  if (flat) {
    // 1x1 / Linear: rows are a flat pixel index
    const long long M = (long long)N * H * W;
    AView av = nhwc_view(x, 1, 1, (int)M, Cin, 1);
    const int extent[4] = {(int)M, 1, 1, 1};
    const int cs[4] = {1, 1, 1, 1};
    ep.ostride[0] = Cout;
    ep.odim[0] = (int)M;
    if (res) {
      ep.res = res;
      ep.rstride[0] = Cout;
    }
    if (gn_sum) {
      // image index of a flat row = row / (H*W): expressed by tiling n separately when possible
      if (N > 1) {
        AView av4 = nhwc_view(x, N, 1, H * W, Cin, 1);
        const int ext4[4] = {H * W, 1, N, 1};
        ep.ostride[0] = Cout;
        ep.ostride[2] = (long long)H * W * Cout;
        ep.odim[0] = H * W;
        ep.odim[2] = N;
        ep.gn_dim = 2;
        if (res) {
          ep.rstride[0] = Cout;
          ep.rstride[2] = (long long)H * W * Cout;
        }
        return run_fprop(av4, ext4, cs, 0, -1, 1, taps, bk, Cin, w, (uint64_t)Cout, (uint64_t)Cin, Cout, ep, bn_hint,
                         stream);
      }
      ep.gn_dim = 1;  // coord is always 0
    }
    return run_fprop(av, extent, cs, 0, -1, 1, taps, bk, Cin, w, (uint64_t)Cout, (uint64_t)Cin, Cout, ep, bn_hint,
                     stream);
  }
  AView av = nhwc_view(x, N, H, W, Cin, stride);
  const int extent[4] = {Wo, Ho, N, 1};
  const int cs[4] = {stride, stride, 1, 1};
  ep.ostride[0] = Cout;
  ep.ostride[1] = (long long)Wo * Cout;
  ep.ostride[2] = (long long)Ho * Wo * Cout;
  ep.odim[0] = Wo;
  ep.odim[1] = Ho;
  ep.odim[2] = N;
  ep.gn_dim = 2;
  if (res) {
    ep.res = res;
    if (flags & EOSVOS_FLAG_RES_HALF) {  // residual lives on the 2x coarser grid (FPN top-down path)
      const int Hr = Ho / 2, Wr = Wo / 2;
      ep.rstride[0] = Cout;
      ep.rstride[1] = (long long)Wr * Cout;
      ep.rstride[2] = (long long)Hr * Wr * Cout;
      ep.rshift[0] = ep.rshift[1] = 1;
    } else {
      ep.rstride[0] = Cout;
      ep.rstride[1] = (long long)Wo * Cout;
      ep.rstride[2] = (long long)Ho * Wo * Cout;
    }
  }
  return run_fprop(av, extent, cs, 0, 1, nt, taps, bk, Cin, w, (uint64_t)Cout, (uint64_t)KH * KW * Cin, Cout, ep,
                   bn_hint, stream);
}

// ---------------------------------------------------------------------------------------------
// conv2d data gradient.  dy [N,Ho,Wo,Cout] bf16, wt [Cin,KH,KW,Cout] bf16, dx [N,H,W,Cin] bf16.
// stride 1: one fprop-style launch with mirrored tap offsets.  stride 2: one launch per input
// parity class (each class sees only the taps that reach it), no atomics, no zero-insertion.
// ---------------------------------------------------------------------------------------------
extern "C" int eosvos_conv2d_dgrad(const void* dy, const void* wt, void* dx, const void* acc, int N, int H, int W,
                                   int Cin, int Cout, int KH, int KW, int stride, int pad, int flags, int bn_hint,
                                   eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(dy && wt && dx, "conv2d_dgrad: null pointer");
  EOSVOS_REQUIRE(!acc || stride == 1, "conv2d_dgrad: the accumulate operand needs stride 1");
  EOSVOS_REQUIRE(Cout % 64 == 0, "conv2d_dgrad: Cout must be a multiple of 64");
  EOSVOS_REQUIRE(Cin % 8 == 0, "conv2d_dgrad: Cin must be a multiple of 8");
  EOSVOS_REQUIRE(KH * KW <= MAX_TAPS && (stride == 1 || stride == 2), "conv2d_dgrad: unsupported filter");
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  const int out_fp32 = (flags & EOSVOS_FLAG_OUT_FP32) ? 1 : 0;
  AView av = nhwc_view(dy, N, Ho, Wo, Cout, 1);
  const int cs[4] = {1, 1, 1, 1};
  if (stride == 1) {
    int taps[MAX_TAPS][5], bk[MAX_TAPS];
    int nt = 0;
    for (int kh = 0; kh < KH; ++kh)
      for (int kw = 0; kw < KW; ++kw) {
        taps[nt][0] = 0;
        taps[nt][1] = pad - kw;
        taps[nt][2] = pad - kh;
        taps[nt][3] = taps[nt][4] = 0;
        bk[nt] = (kh * KW + kw) * Cout;
        ++nt;
      }
    Epilogue ep;
    ep.out = dx;
    ep.out_fp32 = out_fp32;
    ep.ostride[0] = Cin;
    ep.ostride[1] = (long long)W * Cin;
    ep.ostride[2] = (long long)H * W * Cin;
    ep.odim[0] = W;
    ep.odim[1] = H;
    ep.odim[2] = N;
    if (acc) {   // dx = dgrad + acc (gradient that reached the same tensor through another branch), added in fp32
      ep.res = acc;
      ep.rstride[0] = Cin;
      ep.rstride[1] = (long long)W * Cin;
      ep.rstride[2] = (long long)H * W * Cin;
    }
    const int extent[4] = {W, H, N, 1};
    const bool flat = (KH == 1 && KW == 1 && pad == 0);
    if (flat) {
      const long long M = (long long)N * H * W;
      AView avf = nhwc_view(dy, 1, 1, (int)M, Cout, 1);
      const int ext[4] = {(int)M, 1, 1, 1};
      Epilogue ef;
      ef.out = dx;
      ef.out_fp32 = out_fp32;
      ef.ostride[0] = Cin;
      ef.odim[0] = (int)M;
      if (acc) {
        ef.res = acc;
        ef.rstride[0] = Cin;
      }
      return run_fprop(avf, ext, cs, 0, -1, 1, taps, bk, Cout, wt, (uint64_t)Cin, (uint64_t)Cout, Cin, ef, bn_hint,
                       stream);
    }
    return run_fprop(av, extent, cs, 0, 1, nt, taps, bk, Cout, wt, (uint64_t)Cin, (uint64_t)KH * KW * Cout, Cin, ep,
                     bn_hint, stream);
  }
  // stride 2
  bool need_zero = false;
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      int cnt = 0;
      for (int kh = 0; kh < KH; ++kh)
        for (int kw = 0; kw < KW; ++kw)
          if (((ph + pad - kh) & 1) == 0 && ((pw + pad - kw) & 1) == 0) ++cnt;
      if (cnt == 0) need_zero = true;
    }
  if (need_zero) {
    cudaError_t e = cudaMemsetAsync(dx, 0, (size_t)N * H * W * Cin * (out_fp32 ? 4 : 2), stream);
    if (e != cudaSuccess) return set_cuda_error(e, "conv2d_dgrad memset");
  }
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      int taps[MAX_TAPS][5], bk[MAX_TAPS];
      int nt = 0;
      for (int kh = 0; kh < KH; ++kh)
        for (int kw = 0; kw < KW; ++kw) {
          const int nh = ph + pad - kh, nw = pw + pad - kw;
          if ((nh & 1) || (nw & 1)) continue;
          taps[nt][0] = 0;
          taps[nt][1] = nw >= 0 ? nw / 2 : -((-nw) / 2);
          taps[nt][2] = nh >= 0 ? nh / 2 : -((-nh) / 2);
          taps[nt][3] = taps[nt][4] = 0;
          bk[nt] = (kh * KW + kw) * Cout;
          ++nt;
        }
      if (nt == 0) continue;
      const int Hp = (H - ph + 1) / 2, Wp = (W - pw + 1) / 2;
      if (Hp <= 0 || Wp <= 0) continue;
      Epilogue ep;
      const size_t esz = out_fp32 ? 4 : 2;
      ep.out = reinterpret_cast<uint8_t*>(dx) + ((size_t)ph * W + pw) * Cin * esz;
      ep.out_fp32 = out_fp32;
      ep.ostride[0] = 2LL * Cin;
      ep.ostride[1] = 2LL * W * Cin;
      ep.ostride[2] = (long long)H * W * Cin;
      ep.odim[0] = Wp;
      ep.odim[1] = Hp;
      ep.odim[2] = N;
      const int extent[4] = {Wp, Hp, N, 1};
      EOSVOS_TRY(run_fprop(av, extent, cs, 0, 1, nt, taps, bk, Cout, wt, (uint64_t)Cin, (uint64_t)KH * KW * Cout, Cin,
                           ep, bn_hint, stream));
    }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// conv2d weight gradient.  x [N,H,W,Cin] bf16, dy [N,Ho,Wo,Cout] bf16 -> dw fp32 [Cout,Cin,KH,KW]
// (torch OIHW), ACCUMULATED with red.add: the caller zeroes dw (or lets it accumulate).
// ---------------------------------------------------------------------------------------------
namespace eosvos {
static int run_wgrad(const AView& a_view, const AView& b_view, const int extent[4], const int b_stride[4], int sp_w,
                     int sp_h, int num_taps, const int (*tda)[5], const int (*tdb)[5], int m_valid, int n_valid,
                     float* dw, long long s_m, long long s_tap, int n_inner, long long s_n_inner,
                     long long s_n_outer, float alpha, int bn_hint, int split_hint, cudaStream_t stream) {
  WgradParams p;
  memset(&p, 0, sizeof p);
  int tw = 1, th = 1;
  if (sp_w >= 0 && sp_h >= 0)
    choose_tile(extent[sp_w], extent[sp_h], 64, 256 / b_stride[sp_w], 256 / b_stride[sp_h], true, &tw, &th);
  else if (sp_w >= 0)
    tw = std::min(64, extent[sp_w]);
  uint32_t boxa[5] = {64, 1, 1, 1, 1}, boxb[5] = {64, 1, 1, 1, 1};
  long long tiles = 1;
  for (int d = 0; d < 4; ++d) {
    int b = 1;
    if (d == sp_w) b = tw;
    if (d == sp_h) b = th;
    p.rows_box[d] = b;
    p.ntiles[d] = (extent[d] + b - 1) / b;
    p.a_tile_step[d] = b;
    p.b_tile_step[d] = b * b_stride[d];
    boxa[d + 1] = (uint32_t)b;
    boxb[d + 1] = (uint32_t)(b * b_stride[d]);
    tiles *= p.ntiles[d];
  }
  p.rows = p.rows_box[0] * p.rows_box[1] * p.rows_box[2] * p.rows_box[3];
  p.kpad = ((p.rows + 15) / 16) * 16;
  p.num_taps = num_taps;
  for (int t = 0; t < num_taps; ++t)
    for (int d = 0; d < 5; ++d) {
      p.tap_delta_a[t][d] = tda[t][d];
      p.tap_delta_b[t][d] = tdb[t][d];
    }
  int bn = n_valid <= 64 ? 64 : (n_valid <= 128 ? 128 : 256);
  if (bn_hint == 64 || bn_hint == 128 || bn_hint == 256) bn = bn_hint;
  p.n_tiles_n = (n_valid + bn - 1) / bn;
  const int m_tiles = (m_valid + 127) / 128;
  const long long base_ctas = (long long)m_tiles * p.n_tiles_n * num_taps;
  // narrow (64-column) tiles run three shallow-ring CTAs per SM (conv_gemm.cu::launch_wgrad case 63)
  static const bool light_ok = [] {
    const char* e = getenv("EOSVOS_WGRAD_LIGHT");
    return !(e && e[0] == '0');
  }();
  // (only where every CTA keeps a long K loop: >= 32 pixel tiles per CTA slot; short launches lose to the extra
  // prologues and RED epilogues -- 1x1 64->256 at 3x192x336: 34.8 us deep vs 45.1 us light; 3x3 64->64: 100 vs 47)
  const bool light = light_ok && bn == 64 && tiles * base_ctas >= 32LL * 3 * num_sms();
  const long long per_sm = light ? 3 : 1;
  // split the pixel reduction so that the grid is a whole number of waves (1 CTA / SM, long-running CTAs: a
  // partial trailing wave costs a full wave of time); prefer the fewest waves among near-equal efficiencies
  long long split = 1;
  if (split_hint > 0) {
    split = split_hint;
  } else {
    const long long sms = num_sms() * per_sm;
    double best_eff = -1.0;
    for (int k = 1; k <= 4; ++k) {
      long long sp = (k * sms) / base_ctas;
      if (sp < 1) sp = 1;
      if (sp > tiles) sp = tiles;
      const long long ctas = base_ctas * sp;
      const long long waves = (ctas + sms - 1) / sms;
      const double eff = (double)ctas / (double)(waves * sms);
      if (eff > best_eff + 0.03) {
        best_eff = eff;
        split = sp;
      }
    }
  }
  split = std::max(1LL, std::min(split, tiles));
  p.tiles_per_split = (int)((tiles + split - 1) / split);
  split = (tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.total_tiles = (int)tiles;
  p.m_valid = m_valid;
  p.n_valid = n_valid;
  p.dw = dw;
  p.dw_m_stride = s_m;
  p.dw_tap_stride = s_tap;
  p.n_inner = n_inner > 0 ? n_inner : 0x7fffffff;
  p.n_inner_stride = s_n_inner;
  p.n_outer_stride = s_n_outer;
  p.alpha = alpha;
  CUtensorMap tmA, tmB;
  EOSVOS_TRY(make_tensor_map_act(&tmA, a_view.base, 5, a_view.dims, a_view.strides, boxa, a_view.estride));
  EOSVOS_TRY(make_tensor_map_act(&tmB, b_view.base, 5, b_view.dims, b_view.strides, boxb, b_view.estride));
  dim3 grid((unsigned)split, (unsigned)(m_tiles * p.n_tiles_n), (unsigned)num_taps);
  return launch_wgrad(light ? 63 : bn, tmA, tmB, p, grid, stream);
}
}  // namespace eosvos

extern "C" int eosvos_conv2d_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout,
                                   int KH, int KW, int stride, int pad, float alpha, int bn_hint, int split_hint,
                                   int dw_layout, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(x && dy && dw, "conv2d_wgrad: null pointer");
  EOSVOS_REQUIRE(Cin % 8 == 0 && Cout % 8 == 0, "conv2d_wgrad: channels must be multiples of 8");
  EOSVOS_REQUIRE(KH * KW <= MAX_TAPS && (stride == 1 || stride == 2), "conv2d_wgrad: unsupported filter");
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  int tda[MAX_TAPS][5], tdb[MAX_TAPS][5];
  int nt = 0;
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw) {
      for (int d = 0; d < 5; ++d) tda[nt][d] = tdb[nt][d] = 0;
      tdb[nt][1] = kw - pad;
      tdb[nt][2] = kh - pad;
      ++nt;
    }
  const int T = KH * KW;
  // output scattered straight into torch's OIHW layout: dw[co][ci][tap]
  const bool flat = (KH == 1 && KW == 1 && stride == 1 && pad == 0);
  if (flat) {
    const long long M = (long long)N * H * W;
    AView a = nhwc_view(dy, 1, 1, (int)M, Cout, 1);
    AView b = nhwc_view(x, 1, 1, (int)M, Cin, 1);
    const int ext[4] = {(int)M, 1, 1, 1};
    const int bs[4] = {1, 1, 1, 1};
    return run_wgrad(a, b, ext, bs, 0, -1, 1, tda, tdb, Cout, Cin, dw, (long long)Cin, 0, 0, 1, 0, alpha, bn_hint,
                     split_hint, stream);
  }
  AView a = nhwc_view(dy, N, Ho, Wo, Cout, 1);
  AView b = nhwc_view(x, N, H, W, Cin, stride);
  const int ext[4] = {Wo, Ho, N, 1};
  const int bs[4] = {stride, stride, 1, 1};
  if (dw_layout == 1)   // [Cout][tap][Cin]
    return run_wgrad(a, b, ext, bs, 0, 1, nt, tda, tdb, Cout, Cin, dw, (long long)T * Cin, (long long)Cin, 0, 1, 0,
                     alpha, bn_hint, split_hint, stream);
  return run_wgrad(a, b, ext, bs, 0, 1, nt, tda, tdb, Cout, Cin, dw, (long long)T * Cin, 1, 0, (long long)T, 0,
                   alpha, bn_hint, split_hint, stream);
}

// ---------------------------------------------------------------------------------------------
// 2x2 stride-2 transposed conv (mask head, torchvision MaskRCNNPredictor.conv5_mask).
//   fprop : x [N,h,w,Cin], wd [(dy,dx,co)][Cin]           -> y [N,2h,2w,Cout] (+bias, relu)
//   dgrad : dy [N,2h,2w,Cout], wdt [Cin][(dy,dx,co)]      -> dx [N,h,w,Cin]
//   wgrad : dw[ci][co][dy][dx] += sum_pix dy[pix@(dy,dx)][co] * x[pix][ci]   (torch layout)
// ---------------------------------------------------------------------------------------------
extern "C" int eosvos_deconv2x2_fprop(const void* x, const void* wd, const float* bias4, void* y, int N, int h, int w,
                                      int Cin, int Cout, int flags, int bn_hint, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(x && wd && y, "deconv2x2_fprop: null pointer");
  EOSVOS_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "deconv2x2_fprop: channels must be multiples of 64");
  AView av = nhwc_view(x, N, h, w, Cin, 1);
  const int extent[4] = {w, h, N, 1};
  const int cs[4] = {1, 1, 1, 1};
  int taps[1][5] = {{0, 0, 0, 0, 0}};
  int bk[1] = {0};
  Epilogue ep;
  ep.out = y;
  ep.bias = bias4;  // bias replicated per (dy,dx) group by the caller: [4*Cout]
  ep.relu = flags & EOSVOS_FLAG_RELU;
  ep.out_fp32 = (flags & EOSVOS_FLAG_OUT_FP32) ? 1 : 0;
  const long long W2 = 2LL * w;
  ep.ostride[0] = 2LL * Cout;
  ep.ostride[1] = 2LL * W2 * Cout;
  ep.ostride[2] = 4LL * h * w * Cout;
  ep.odim[0] = w;
  ep.odim[1] = h;
  ep.odim[2] = N;
  ep.ogroup = Cout;
  for (int g = 0; g < 4; ++g) ep.ogroup_off[g] = ((long long)(g >> 1) * W2 + (g & 1)) * Cout;
  int bn = bn_hint;
  if (bn != 64 && bn != 128 && bn != 256) bn = (Cout % 128 == 0) ? 128 : 64;
  if (Cout % bn != 0) bn = 64;
  return run_fprop(av, extent, cs, 0, 1, 1, taps, bk, Cin, wd, (uint64_t)4 * Cout, (uint64_t)Cin, 4 * Cout, ep, bn,
                   stream);
}

namespace eosvos {
// rank-5 view of a [N,2h,2w,C] tensor as (2C [dx,c], w, dy, h, N)
static inline AView subpixel_view(const void* base, int N, int h, int w, int C) {
  AView v;
  v.base = base;
  v.dims[0] = 2ULL * C;
  v.dims[1] = (uint64_t)w;
  v.dims[2] = 2;
  v.dims[3] = (uint64_t)h;
  v.dims[4] = (uint64_t)N;
  v.strides[0] = 2ULL * C * 2;
  v.strides[1] = 2ULL * w * C * 2;
  v.strides[2] = 4ULL * w * C * 2;
  v.strides[3] = 4ULL * h * w * C * 2;
  for (int d = 0; d < 5; ++d) v.estride[d] = 1;
  return v;
}
}  // namespace eosvos

extern "C" int eosvos_deconv2x2_dgrad(const void* dy, const void* wdt, void* dx, int N, int h, int w, int Cin, int Cout,
                                      int flags, int bn_hint, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(dy && wdt && dx, "deconv2x2_dgrad: null pointer");
  EOSVOS_REQUIRE(Cin % 8 == 0 && Cout % 64 == 0, "deconv2x2_dgrad: bad channel counts");
  AView av = subpixel_view(dy, N, h, w, Cout);
  const int extent[4] = {w, 1, h, N};
  const int cs[4] = {1, 1, 1, 1};
  int taps[4][5], bk[4];
  for (int g = 0; g < 4; ++g) {
    taps[g][0] = (g & 1) * Cout;
    taps[g][1] = 0;
    taps[g][2] = g >> 1;
    taps[g][3] = 0;
    taps[g][4] = 0;
    bk[g] = g * Cout;
  }
  Epilogue ep;
  ep.out = dx;
  ep.out_fp32 = (flags & EOSVOS_FLAG_OUT_FP32) ? 1 : 0;
  ep.ostride[0] = Cin;
  ep.ostride[1] = 0;
  ep.ostride[2] = (long long)w * Cin;
  ep.ostride[3] = (long long)h * w * Cin;
  ep.odim[0] = w;
  ep.odim[1] = 1;
  ep.odim[2] = h;
  ep.odim[3] = N;
  return run_fprop(av, extent, cs, 0, 2, 4, taps, bk, Cout, wdt, (uint64_t)Cin, (uint64_t)4 * Cout, Cin, ep, bn_hint,
                   stream);
}

extern "C" int eosvos_deconv2x2_wgrad(const void* x, const void* dy, float* dw, int N, int h, int w, int Cin, int Cout,
                                      float alpha, int bn_hint, int split_hint, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(x && dy && dw, "deconv2x2_wgrad: null pointer");
  EOSVOS_REQUIRE(Cin % 8 == 0 && Cout % 8 == 0, "deconv2x2_wgrad: bad channel counts");
  AView a = subpixel_view(dy, N, h, w, Cout);
  // x viewed with a unit "dy" dim so both operands share the (w, dy, h, N) tiling
  AView b;
  b.base = x;
  b.dims[0] = (uint64_t)Cin;
  b.dims[1] = (uint64_t)w;
  b.dims[2] = 1;
  b.dims[3] = (uint64_t)h;
  b.dims[4] = (uint64_t)N;
  b.strides[0] = (uint64_t)Cin * 2;
  b.strides[1] = (uint64_t)w * Cin * 2;
  b.strides[2] = (uint64_t)w * Cin * 2;
  b.strides[3] = (uint64_t)h * w * Cin * 2;
  for (int d = 0; d < 5; ++d) b.estride[d] = 1;
  int tda[4][5], tdb[4][5];
  for (int g = 0; g < 4; ++g) {
    for (int d = 0; d < 5; ++d) tda[g][d] = tdb[g][d] = 0;
    tda[g][0] = (g & 1) * Cout;
    tda[g][2] = g >> 1;
  }
  const int ext[4] = {w, 1, h, N};
  const int bs[4] = {1, 1, 1, 1};
  // torch ConvTranspose2d layout dw[ci][co][dy][dx]
  return run_wgrad(a, b, ext, bs, 0, 2, 4, tda, tdb, Cout, Cin, dw, 4LL, 1, 0, 4LL * Cout, 0, alpha, bn_hint,
                   split_hint, stream);
}

// ---------------------------------------------------------------------------------------------
// flat weight gradient with a caller-defined destination layout (Linear, fc6, stem im2col).
// ---------------------------------------------------------------------------------------------
extern "C" int eosvos_gemm_wgrad(const void* x, const void* dy, float* dw, long long rows, int n_cols, int m_cols,
                                 long long s_m, int n_inner, long long s_n_inner, long long s_n_outer, float alpha,
                                 int bn_hint, int split_hint, int n_valid, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(x && dy && dw, "gemm_wgrad: null pointer");
  EOSVOS_REQUIRE(n_valid >= 0 && n_valid <= n_cols, "gemm_wgrad: n_valid must be within the x columns");
  EOSVOS_REQUIRE(n_cols % 8 == 0 && m_cols % 8 == 0, "gemm_wgrad: column counts must be multiples of 8");
  EOSVOS_REQUIRE(rows > 0 && rows < (1LL << 31), "gemm_wgrad: bad row count");
  AView a = nhwc_view(dy, 1, 1, (int)rows, m_cols, 1);
  AView b = nhwc_view(x, 1, 1, (int)rows, n_cols, 1);
  int tda[1][5] = {{0, 0, 0, 0, 0}}, tdb[1][5] = {{0, 0, 0, 0, 0}};
  const int ext[4] = {(int)rows, 1, 1, 1};
  const int bs[4] = {1, 1, 1, 1};
  // columns >= n_valid (zero padding of x, e.g. the stem's im2col K 147 -> 192) are never written: their destination
  // offsets would fall outside dw
  return run_wgrad(a, b, ext, bs, 0, -1, 1, tda, tdb, m_cols, n_valid > 0 ? n_valid : n_cols, dw, s_m, 0, n_inner,
                   s_n_inner, s_n_outer, alpha, bn_hint, split_hint, stream);
}
