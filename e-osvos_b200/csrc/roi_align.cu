// Multi-scale RoIAlign forward / backward on NHWC bf16 FPN levels, and the adaptive-sampling
// single-channel variant that builds 56x56 mask targets.
// Replaces torchvision::roi_align / _roi_align_backward reached from the reference's
// box_roi_pool / mask_roi_pool calls (src/networks/mask_rcnn.py:113,147; pooler built at :435-442,
// sampling_ratio=2, aligned=False) and project_masks_on_boxes (:38,:70; sampling_ratio=-1).
// Level assignment follows torchvision.ops.poolers.LevelMapper: floor(4 + log2(sqrt(area)/224) + 1e-6)
// clamped to [2,5]  (SURVEY.md App. B).
#include "common.h"
#include "../../include/eosvos_b200.h"
#include "act.cuh"
#include "ptx.cuh"

namespace eosvos {

struct RoiLevels {
  const act_t* feat[4];
  float* dfeat[4];
  int H[4], W[4];
  float scale[4];
};

__device__ __forceinline__ int roi_level(float x1, float y1, float x2, float y2) {
  const float area = (x2 - x1) * (y2 - y1);
  const float s = sqrtf(area);
  float lvl = floorf(4.0f + log2f(s / 224.0f) + 1e-6f);
  lvl = fminf(fmaxf(lvl, 2.0f), 5.0f);
  return (int)lvl - 2;
}

struct Bilin {
  int y0, y1, x0, x1;
  float w00, w01, w10, w11;
  bool ok;
};

__device__ __forceinline__ Bilin bilin_setup(float y, float x, int H, int W) {
  Bilin b;
  b.ok = !(y < -1.0f || y > (float)H || x < -1.0f || x > (float)W);
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  b.y0 = (int)y;
  b.x0 = (int)x;
  if (b.y0 >= H - 1) {
    b.y1 = b.y0 = H - 1;
    y = (float)b.y0;
  } else {
    b.y1 = b.y0 + 1;
  }
  if (b.x0 >= W - 1) {
    b.x1 = b.x0 = W - 1;
    x = (float)b.x0;
  } else {
    b.x1 = b.x0 + 1;
  }
  const float ly = y - (float)b.y0, lx = x - (float)b.x0;
  const float hy = 1.f - ly, hx = 1.f - lx;
  b.w00 = hy * hx;
  b.w01 = hy * lx;
  b.w10 = ly * hx;
  b.w11 = ly * lx;
  return b;
}


__device__ __forceinline__ void ld8(const act_t* p, float (&f)[8]) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  const act2_t* h = reinterpret_cast<const act2_t*>(&v);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 t = act22float2(h[k]);
    f[2 * k] = t.x;
    f[2 * k + 1] = t.y;
  }
}

// one thread = one (roi, ph, pw, 8-channel vector); rois: [R][5] = (batch, x1, y1, x2, y2)
template <bool BWD>
__global__ void __launch_bounds__(256)
roi_align_kernel(const RoiLevels lv, const float* __restrict__ rois, act_t* __restrict__ out,
                 const act_t* __restrict__ dout, int R, int P, int C, int sampling) {
  const int cv = C >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)R * P * P * cv;
  if (idx >= total) return;
  const int c8 = (int)(idx % cv);
  long long t = idx / cv;
  const int pw = (int)(t % P);
  t /= P;
  const int ph = (int)(t % P);
  const int r = (int)(t / P);
  const float* roi = rois + (size_t)r * 5;
  const int b = (int)roi[0];
  if (b < 0) {   // padding row of a fixed-size RoI list (CUDA-graphed mask branch): zeros forward, nothing backward
    if (!BWD) *reinterpret_cast<uint4*>(out + (size_t)idx * 8) = make_uint4(0, 0, 0, 0);
    return;
  }
  const float bx1 = roi[1], by1 = roi[2], bx2 = roi[3], by2 = roi[4];
  const int l = roi_level(bx1, by1, bx2, by2);
  const int H = lv.H[l], W = lv.W[l];
  const float sc = lv.scale[l];
  const float x1 = bx1 * sc, y1 = by1 * sc, x2 = bx2 * sc, y2 = by2 * sc;
  const float rw = fmaxf(x2 - x1, 1.f), rh = fmaxf(y2 - y1, 1.f);
  const float bin_h = rh / (float)P, bin_w = rw / (float)P;
  const int gh = sampling, gw = sampling;
  const float inv_count = 1.f / (float)(gh * gw);
  const size_t img_off = (size_t)b * H * W * C + (size_t)c8 * 8;
  float acc[8];
  float g[8];
  if (BWD) {
    ld8(dout + (size_t)idx * 8, g);
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] *= inv_count;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  }
  for (int iy = 0; iy < gh; ++iy) {
    const float y = y1 + ph * bin_h + (iy + 0.5f) * bin_h / (float)gh;
    for (int ix = 0; ix < gw; ++ix) {
      const float x = x1 + pw * bin_w + (ix + 0.5f) * bin_w / (float)gw;
      const Bilin bl = bilin_setup(y, x, H, W);
      if (!bl.ok) continue;
      const size_t o00 = img_off + ((size_t)bl.y0 * W + bl.x0) * C, o01 = img_off + ((size_t)bl.y0 * W + bl.x1) * C;
      const size_t o10 = img_off + ((size_t)bl.y1 * W + bl.x0) * C, o11 = img_off + ((size_t)bl.y1 * W + bl.x1) * C;
      if (BWD) {
        float* d = lv.dfeat[l];
        red_add_v4(d + o00, g[0] * bl.w00, g[1] * bl.w00, g[2] * bl.w00, g[3] * bl.w00);
        red_add_v4(d + o00 + 4, g[4] * bl.w00, g[5] * bl.w00, g[6] * bl.w00, g[7] * bl.w00);
        red_add_v4(d + o01, g[0] * bl.w01, g[1] * bl.w01, g[2] * bl.w01, g[3] * bl.w01);
        red_add_v4(d + o01 + 4, g[4] * bl.w01, g[5] * bl.w01, g[6] * bl.w01, g[7] * bl.w01);
        red_add_v4(d + o10, g[0] * bl.w10, g[1] * bl.w10, g[2] * bl.w10, g[3] * bl.w10);
        red_add_v4(d + o10 + 4, g[4] * bl.w10, g[5] * bl.w10, g[6] * bl.w10, g[7] * bl.w10);
        red_add_v4(d + o11, g[0] * bl.w11, g[1] * bl.w11, g[2] * bl.w11, g[3] * bl.w11);
        red_add_v4(d + o11 + 4, g[4] * bl.w11, g[5] * bl.w11, g[6] * bl.w11, g[7] * bl.w11);
      } else {
        const act_t* f = lv.feat[l];
        float v00[8], v01[8], v10[8], v11[8];
        ld8(f + o00, v00);
        ld8(f + o01, v01);
        ld8(f + o10, v10);
        ld8(f + o11, v11);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          acc[k] += bl.w00 * v00[k] + bl.w01 * v01[k] + bl.w10 * v10[k] + bl.w11 * v11[k];
      }
    }
  }
  if (!BWD) {
    uint4 v;
    act2_t* h = reinterpret_cast<act2_t*>(&v);
#pragma unroll
    for (int k = 0; k < 4; ++k) h[k] = floats2act2(acc[2 * k] * inv_count, acc[2 * k + 1] * inv_count);
    *reinterpret_cast<uint4*>(out + (size_t)idx * 8) = v;
  }
}

// Mask targets: roi_align(gt_masks[matched], boxes, M, spatial_scale 1, sampling_ratio -1 (adaptive)).
// masks: uint8 [G][H][W]; rois [R][5] = (mask index, x1, y1, x2, y2); out fp32 [R][M][M].
__global__ void __launch_bounds__(256)
mask_target_kernel(const uint8_t* __restrict__ masks, const float* __restrict__ rois, float* __restrict__ out, int R,
                   int M, int H, int W) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)R * M * M) return;
  const int pw = (int)(idx % M);
  const int ph = (int)((idx / M) % M);
  const int r = (int)(idx / ((long long)M * M));
  const float* roi = rois + (size_t)r * 5;
  if (roi[0] < 0.f) {   // padding row
    out[idx] = 0.f;
    return;
  }
  const uint8_t* m = masks + (size_t)((int)roi[0]) * H * W;
  const float x1 = roi[1], y1 = roi[2], x2 = roi[3], y2 = roi[4];
  const float rw = fmaxf(x2 - x1, 1.f), rh = fmaxf(y2 - y1, 1.f);
  const float bin_h = rh / (float)M, bin_w = rw / (float)M;
  const int gh = (int)ceilf(rh / (float)M), gw = (int)ceilf(rw / (float)M);
  const float count = fmaxf((float)(gh * gw), 1.f);
  float acc = 0.f;
  for (int iy = 0; iy < gh; ++iy) {
    const float y = y1 + ph * bin_h + (iy + 0.5f) * bin_h / (float)gh;
    for (int ix = 0; ix < gw; ++ix) {
      const float x = x1 + pw * bin_w + (ix + 0.5f) * bin_w / (float)gw;
      const Bilin bl = bilin_setup(y, x, H, W);
      if (!bl.ok) continue;
      acc += bl.w00 * (float)m[(size_t)bl.y0 * W + bl.x0] + bl.w01 * (float)m[(size_t)bl.y0 * W + bl.x1] +
             bl.w10 * (float)m[(size_t)bl.y1 * W + bl.x0] + bl.w11 * (float)m[(size_t)bl.y1 * W + bl.x1];
    }
  }
  out[idx] = acc / count;
}

}  // namespace eosvos

using namespace eosvos;

static int fill_levels(RoiLevels* lv, const void* const* feats, float* const* dfeats, const int* Hs, const int* Ws,
                       const float* scales) {
  for (int l = 0; l < 4; ++l) {
    lv->feat[l] = feats ? reinterpret_cast<const act_t*>(feats[l]) : nullptr;
    lv->dfeat[l] = dfeats ? dfeats[l] : nullptr;
    lv->H[l] = Hs[l];
    lv->W[l] = Ws[l];
    lv->scale[l] = scales[l];
  }
  return 0;
}

extern "C" int eosvos_roi_align_fwd(const void* const* feats, const int* Hs, const int* Ws, const float* scales,
                                    const float* rois, void* out, int R, int P, int C, int sampling,
                                    eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (R == 0) return 0;
  EOSVOS_REQUIRE(feats && rois && out, "roi_align_fwd: null pointer");
  EOSVOS_REQUIRE(C % 8 == 0 && sampling > 0, "roi_align_fwd: C must be a multiple of 8, sampling_ratio > 0");
  RoiLevels lv;
  fill_levels(&lv, feats, nullptr, Hs, Ws, scales);
  const long long total = (long long)R * P * P * (C >> 3);
  roi_align_kernel<false><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
      lv, rois, reinterpret_cast<act_t*>(out), nullptr, R, P, C, sampling);
  return check_launch("roi_align_kernel<fwd>");
}

// dfeats: fp32 [N][H_l][W_l][C] per level, ACCUMULATED into (caller zeroes).
extern "C" int eosvos_roi_align_bwd(float* const* dfeats, const int* Hs, const int* Ws, const float* scales,
                                    const float* rois, const void* dout, int R, int P, int C, int sampling,
                                    eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (R == 0) return 0;
  EOSVOS_REQUIRE(dfeats && rois && dout, "roi_align_bwd: null pointer");
  EOSVOS_REQUIRE(C % 8 == 0 && sampling > 0, "roi_align_bwd: C must be a multiple of 8, sampling_ratio > 0");
  RoiLevels lv;
  fill_levels(&lv, nullptr, dfeats, Hs, Ws, scales);
  const long long total = (long long)R * P * P * (C >> 3);
  roi_align_kernel<true><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
      lv, rois, nullptr, reinterpret_cast<const act_t*>(dout), R, P, C, sampling);
  return check_launch("roi_align_kernel<bwd>");
}

extern "C" int eosvos_mask_targets(const uint8_t* masks, const float* rois, float* out, int R, int M, int H, int W,
                                   eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (R == 0) return 0;
  EOSVOS_REQUIRE(masks && rois && out, "mask_targets: null pointer");
  const long long total = (long long)R * M * M;
  mask_target_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(masks, rois, out, R, M, H, W);
  return check_launch("mask_target_kernel");
}
