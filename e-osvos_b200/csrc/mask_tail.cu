// Inference tail: sigmoid -> pad(1) -> bilinear paste into the frame -> (max<0.5 -> bg, argmax+1)
// -> bounding box / pixel count of the propagated target, in ONE kernel per frame.
// Replaces, per object-frame, in the reference:
//   maskrcnn_inference (sigmoid + class select)            tv roi_heads.py:56-82
//   paste_masks_in_image (per-detection Python loop)       tv roi_heads.py:378-502
//   threshold / argmax of run_loader                       src/util/helper_func.py:113-121
//   mask -> box via host np.where                          src/networks/mask_rcnn.py:626-632
// and removes their host syncs.  mask_to_bbox alone serves MaskRCNN.forward on arbitrary targets.
#include "common.h"
#include "../../include/eosvos_b200.h"
#include <limits.h>

namespace eosvos {

constexpr int TAIL_MAXK = 8;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// padded (M+2)x(M+2) probability map lookup
__device__ __forceinline__ float padded_prob(const float* __restrict__ lg, int M, int py, int px) {
  if (py <= 0 || px <= 0 || py > M || px > M) return 0.f;
  return sigmoidf_(lg[(py - 1) * M + (px - 1)]);
}

// F.interpolate(mode='bilinear', align_corners=False) source index / weight
__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int* i0, int* i1, float* l1) {
  float s = scale * ((float)dst + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  int a = (int)s;
  if (a > in_size - 1) a = in_size - 1;
  *i0 = a;
  *i1 = a + ((a < in_size - 1) ? 1 : 0);
  *l1 = fminf(fmaxf(s - (float)a, 0.f), 1.f);
}

struct BoxStat {  // per (image, id): xmin, ymin, xmax, ymax (inclusive pixel extents), count
  int v[5];
};

__device__ __forceinline__ void warp_box_update(int* stat, bool on, int x, int y, bool count_it = true) {
  const unsigned m = __ballot_sync(0xffffffffu, on);
  const unsigned mc = __ballot_sync(0xffffffffu, on && count_it);
  if (m == 0) return;
  int xmin = on ? x : INT_MAX, ymin = on ? y : INT_MAX, xmax = on ? x : -1, ymax = on ? y : -1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
    ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
    xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
    ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(stat + 0, xmin);
    atomicMin(stat + 1, ymin);
    atomicMax(stat + 2, xmax);
    atomicMax(stat + 3, ymax);
    if (mc) atomicAdd(stat + 4, __popc(mc));
  }
}

// logits [D][Cc][M][M] fp32; det_of_chan [B*K] (detection index or -1); det_label [D]; det_box [D][4]
// (image coords).  probs [B][K][H][W]; target [B][H][W] (0 = bg, k+1 = object); stats [B][K][5].
__global__ void __launch_bounds__(256)
paste_threshold_kernel(const float* __restrict__ logits, const int* __restrict__ det_of_chan,
                       const long long* __restrict__ det_label, const float* __restrict__ det_box,
                       float* __restrict__ probs, float* __restrict__ target, int* __restrict__ stats, int B, int K,
                       int H, int W, int M, int Cc, float thresh) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * H * W;
  const bool inb = idx < total;
  const long long cidx = inb ? idx : total - 1;
  const int x = (int)(cidx % W);
  const int y = (int)((cidx / W) % H);
  const int b = (int)(cidx / ((long long)W * H));
  float best = -1.f;
  int besti = 0;
  const int Mp = M + 2;
  const float escale = (float)Mp / (float)M;
  for (int k = 0; k < K; ++k) {
    const int d = det_of_chan[b * K + k];
    float p = 0.f;
    if (d >= 0) {
      const float* bx = det_box + (size_t)d * 4;
      // expand_boxes(...).to(int64): truncation toward zero
      const float wh = (bx[2] - bx[0]) * 0.5f * escale, hh = (bx[3] - bx[1]) * 0.5f * escale;
      const float xc = (bx[2] + bx[0]) * 0.5f, yc = (bx[3] + bx[1]) * 0.5f;
      const int bx0 = (int)(xc - wh), bx1 = (int)(xc + wh), by0 = (int)(yc - hh), by1 = (int)(yc + hh);
      const int w = max(bx1 - bx0 + 1, 1), h = max(by1 - by0 + 1, 1);
      const int x0 = max(bx0, 0), x1 = min(bx1 + 1, W), y0 = max(by0, 0), y1 = min(by1 + 1, H);
      if (x >= x0 && x < x1 && y >= y0 && y < y1) {
        const int mx = x - bx0, my = y - by0;
        if (mx < w && my < h) {
          int ya, yb, xa, xb;
          float ly, lx;
          src_index(my, (float)Mp / (float)h, Mp, &ya, &yb, &ly);
          src_index(mx, (float)Mp / (float)w, Mp, &xa, &xb, &lx);
          const float* lg = logits + ((size_t)d * Cc + (int)det_label[d]) * M * M;
          const float v00 = padded_prob(lg, M, ya, xa), v01 = padded_prob(lg, M, ya, xb);
          const float v10 = padded_prob(lg, M, yb, xa), v11 = padded_prob(lg, M, yb, xb);
          p = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
        }
      }
    }
    if (inb) probs[((size_t)b * K + k) * H * W + (size_t)y * W + x] = p;
    if (p > best) {  // first maximum wins, as torch.argmax
      best = p;
      besti = k;
    }
  }
  const bool fg = inb && !(best < thresh);
  const int id = fg ? besti + 1 : 0;
  if (inb && target) target[idx] = (float)id;
  if (stats) {
    // lanes of one warp share b except at image boundaries: fall back to per-lane atomics there
    const int b0 = __shfl_sync(0xffffffffu, b, 0);
    const bool uniform = __all_sync(0xffffffffu, b == b0);
    for (int k = 0; k < K; ++k) {
      const bool on = (id == k + 1);
      if (uniform) {
        warp_box_update(stats + ((size_t)b0 * K + k) * 5, on, x, y);
      } else if (on) {
        int* s = stats + ((size_t)b * K + k) * 5;
        atomicMin(s + 0, x);
        atomicMin(s + 1, y);
        atomicMax(s + 2, x);
        atomicMax(s + 3, y);
        atomicAdd(s + 4, 1);
      }
    }
  }
}

// targets [B][H][W] float ids; id k+1 -> stats[b][k]; value 255 (ignore) belongs to every id
// (reference src/networks/mask_rcnn.py:602-604).
__global__ void __launch_bounds__(256)
mask_to_bbox_kernel(const float* __restrict__ target, int* __restrict__ stats, int B, int K, int H, int W) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * H * W;
  const bool inb = idx < total;
  const long long cidx = inb ? idx : total - 1;
  const int x = (int)(cidx % W);
  const int y = (int)((cidx / W) % H);
  const int b = (int)(cidx / ((long long)W * H));
  const float v = inb ? target[idx] : 0.f;
  const int b0 = __shfl_sync(0xffffffffu, b, 0);
  const bool uniform = __all_sync(0xffffffffu, b == b0);
  for (int k = 0; k < K; ++k) {
    const bool exact = inb && v == (float)(k + 1);
    const bool on = exact || (inb && v == 255.0f);   // box covers ignore pixels, count only the id itself
    if (uniform) {
      warp_box_update(stats + ((size_t)b0 * K + k) * 5, on, x, y, exact);
    } else if (on) {
      int* s = stats + ((size_t)b * K + k) * 5;
      atomicMin(s + 0, x);
      atomicMin(s + 1, y);
      atomicMax(s + 2, x);
      atomicMax(s + 3, y);
      if (exact) atomicAdd(s + 4, 1);
    }
  }
}

__global__ void init_stats_kernel(int* stats, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int f = i % 5;
  stats[i] = (f < 2) ? INT_MAX : (f < 4 ? -1 : 0);
}

}  // namespace eosvos

using namespace eosvos;

extern "C" int eosvos_mask_paste_threshold(const float* logits, const int* det_of_chan, const long long* det_label,
                                           const float* det_box, float* probs, float* target, int* stats, int B, int K,
                                           int H, int W, int M, int Cc, float thresh, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(det_of_chan && probs, "mask_paste_threshold: null pointer");
  EOSVOS_REQUIRE(K >= 1 && K <= TAIL_MAXK, "mask_paste_threshold: 1..8 object channels");
  EOSVOS_REQUIRE(B > 0 && H > 0 && W > 0, "mask_paste_threshold: empty frame");
  if (stats) {
    init_stats_kernel<<<(B * K * 5 + 127) / 128, 128, 0, stream>>>(stats, B * K * 5);
    EOSVOS_TRY(check_launch("init_stats_kernel"));
  }
  const long long total = (long long)B * H * W;
  paste_threshold_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
      logits, det_of_chan, det_label, det_box, probs, target, stats, B, K, H, W, M, Cc, thresh);
  return check_launch("paste_threshold_kernel");
}

extern "C" int eosvos_mask_to_bbox(const float* target, int* stats, int B, int K, int H, int W,
                                   eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(target && stats, "mask_to_bbox: null pointer");
  EOSVOS_REQUIRE(K >= 1 && K <= 64, "mask_to_bbox: 1..64 ids");
  init_stats_kernel<<<(B * K * 5 + 127) / 128, 128, 0, stream>>>(stats, B * K * 5);
  EOSVOS_TRY(check_launch("init_stats_kernel"));
  const long long total = (long long)B * H * W;
  mask_to_bbox_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(target, stats, B, K, H, W);
  return check_launch("mask_to_bbox_kernel");
}
