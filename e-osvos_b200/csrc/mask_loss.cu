// Fused mask losses (forward + gradient in one launch), one CTA per positive RoI.
//   Lovasz hinge : reference src/networks/loss_lovasz.py:18-30 (lovasz_grad), :94-111
//                  (lovasz_hinge_flat), :114-126 (flatten_binary_scores, ignore=255) called per RoI
//                  from src/networks/mask_rcnn.py:56-92 (maskrcnn_loss_lovasz; targets > 1 -> 255).
//   BCE          : reference src/networks/mask_rcnn.py:24-53 (mean BCE-with-logits over all pixels).
// The reference spends ~30 tiny launches per RoI (sort, gather, 2 cumsums, relu, dot); here the
// 3136 errors of a RoI are sorted in shared memory (bitonic, 4096 slots), scanned with warp
// shuffles and reduced to the loss and its gradient without leaving the SM.
#include "common.h"
#include "../../include/eosvos_b200.h"
#include <math_constants.h>

namespace eosvos {

constexpr int LV_THREADS = 1024;
constexpr int LV_SLOTS = 4096;

__device__ __forceinline__ float warp_incl_scan(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// logits: [R][Cc][P] fp32; labels: [R] int64 (class channel); targets: [R][P] fp32
// loss_out: scalar, += loss_r / R.   dlogits: [R][Cc][P] fp32 = d(mean loss)/d logits.
__global__ void __launch_bounds__(LV_THREADS)
lovasz_hinge_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                    const float* __restrict__ targets, float* __restrict__ loss_out, float* __restrict__ loss_per_roi,
                    float* __restrict__ dlogits, int R, int Cc, int P) {
  __shared__ float key[LV_SLOTS];
  __shared__ unsigned short pidx[LV_SLOTS];
  __shared__ float ytab[LV_SLOTS];
  __shared__ float wsum0[32], wsum1[32], wred[32];
  __shared__ float s_gts;
  __shared__ int s_nvalid;

  const int r = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cls = (int)labels[r];
  const float* x = logits + ((size_t)r * Cc + cls) * P;
  const float* y = targets + (size_t)r * P;
  float* dl = dlogits + (size_t)r * Cc * P;

  if (tid == 0) s_nvalid = 0;
  __syncthreads();
  // zero the gradient of the channels that do not belong to this RoI's class
  for (int c = 0; c < Cc; ++c) {
    if (c == cls) continue;
    for (int i = tid; i < P; i += LV_THREADS) dl[(size_t)c * P + i] = 0.f;
  }
  int nv = 0;
  for (int i = tid; i < LV_SLOTS; i += LV_THREADS) {
    float k = -CUDART_INF_F;
    float yy = 0.f;
    if (i < P) {
      yy = y[i];
      if (!(yy > 1.0f)) {  // mask_rcnn.py:86: targets > 1 become 255 = ignored
        k = 1.0f - x[i] * (2.0f * yy - 1.0f);
        ++nv;
      }
    }
    key[i] = k;
    ytab[i] = yy;
    pidx[i] = (unsigned short)i;
  }
  if (nv) atomicAdd(&s_nvalid, nv);
  __syncthreads();
  const int n = s_nvalid;

  // bitonic sort, descending by key (invalid = -inf sink to the end)
  for (int k = 2; k <= LV_SLOTS; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < LV_SLOTS / 2; t += LV_THREADS) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // index with bit j clear
        const int l = i | j;
        const bool desc = ((i & k) == 0);
        const float a = key[i], b = key[l];
        if (desc ? (a < b) : (a > b)) {
          key[i] = b;
          key[l] = a;
          const unsigned short pa = pidx[i];
          pidx[i] = pidx[l];
          pidx[l] = pa;
        }
      }
      __syncthreads();
    }
  }

  // scans over the sorted order: each thread owns 4 consecutive slots
  float g[4], ng[4];
  float loc1 = 0.f, loc0 = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int j = tid * 4 + q;
    const float gt = (j < n) ? ytab[pidx[j]] : 0.f;
    g[q] = gt;
    ng[q] = (j < n) ? 1.0f - gt : 0.f;
    loc1 += g[q];
    loc0 += ng[q];
  }
  const float inc1 = warp_incl_scan(loc1, lane), inc0 = warp_incl_scan(loc0, lane);
  if (lane == 31) {
    wsum1[warp] = inc1;
    wsum0[warp] = inc0;
  }
  __syncthreads();
  if (warp == 0) {
    const float a = wsum1[lane], b = wsum0[lane];
    const float ia = warp_incl_scan(a, lane), ib = warp_incl_scan(b, lane);
    wsum1[lane] = ia - a;  // exclusive warp offsets
    wsum0[lane] = ib - b;
    if (lane == 31) s_gts = ia;
  }
  __syncthreads();
  const float gts = s_gts;
  float c1 = wsum1[warp] + inc1 - loc1;  // exclusive prefix before this thread's 4 slots
  float c0 = wsum0[warp] + inc0 - loc0;
  // jaccard at the slot just before this thread's first slot
  float jprev = 0.f;
  if (tid > 0) jprev = 1.0f - (gts - c1) / (gts + c0);
  float part = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int j = tid * 4 + q;
    c1 += g[q];
    c0 += ng[q];
    if (j < n) {
      const float jac = 1.0f - (gts - c1) / (gts + c0);
      const float gr = (j == 0) ? jac : jac - jprev;
      jprev = jac;
      const float e = key[j];
      const int src = pidx[j];
      const float active = e > 0.f ? 1.f : 0.f;
      part += fmaxf(e, 0.f) * gr;
      const float sign = 2.0f * ytab[src] - 1.0f;
      dl[(size_t)cls * P + src] = -sign * gr * active / (float)R;
    } else if (j < LV_SLOTS) {
      const int src = pidx[j];
      if (src < P) dl[(size_t)cls * P + src] = 0.f;  // ignored pixel
    }
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) part += __shfl_xor_sync(0xffffffffu, part, m);
  if (lane == 0) wred[warp] = part;
  __syncthreads();
  if (warp == 0) {
    float v = wred[lane];
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if (lane == 0) {
      if (loss_per_roi) loss_per_roi[r] = v;
      atomicAdd(loss_out, v / (float)R);
    }
  }
}

// BCE-with-logits, mean over R*P.  One CTA per RoI as well (keeps the launch shape identical).
__global__ void __launch_bounds__(256)
bce_mask_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                const float* __restrict__ targets, float* __restrict__ loss_out, float* __restrict__ dlogits, int R,
                int Cc, int P) {
  __shared__ float wred[8];
  const int r = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cls = (int)labels[r];
  const float* x = logits + ((size_t)r * Cc + cls) * P;
  const float* y = targets + (size_t)r * P;
  float* dl = dlogits + (size_t)r * Cc * P;
  const float inv = 1.0f / ((float)R * (float)P);
  for (int c = 0; c < Cc; ++c) {
    if (c == cls) continue;
    for (int i = tid; i < P; i += 256) dl[(size_t)c * P + i] = 0.f;
  }
  float part = 0.f;
  for (int i = tid; i < P; i += 256) {
    const float xv = x[i], yv = y[i];
    part += fmaxf(xv, 0.f) - xv * yv + log1pf(expf(-fabsf(xv)));
    const float s = 1.0f / (1.0f + expf(-xv));
    dl[(size_t)cls * P + i] = (s - yv) * inv;
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) part += __shfl_xor_sync(0xffffffffu, part, m);
  if (lane == 0) wred[warp] = part;
  __syncthreads();
  if (warp == 0) {
    float v = lane < 8 ? wred[lane] : 0.f;
#pragma unroll
    for (int m = 4; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if (lane == 0) atomicAdd(loss_out, v * inv);
  }
}

}  // namespace eosvos

using namespace eosvos;

extern "C" int eosvos_mask_loss_lovasz(const float* logits, const long long* labels, const float* targets,
                                       float* loss_out, float* loss_per_roi, float* dlogits, int R, int Cc, int P,
                                       eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(loss_out, "mask_loss_lovasz: null loss_out");
  cudaError_t e = cudaMemsetAsync(loss_out, 0, sizeof(float), stream);
  if (e != cudaSuccess) return set_cuda_error(e, "mask_loss memset");
  if (R == 0) return 0;
  EOSVOS_REQUIRE(logits && labels && targets && dlogits, "mask_loss_lovasz: null pointer");
  EOSVOS_REQUIRE(P > 0 && P <= LV_SLOTS, "mask_loss_lovasz: at most 4096 pixels per RoI");
  lovasz_hinge_kernel<<<R, LV_THREADS, 0, stream>>>(logits, labels, targets, loss_out, loss_per_roi, dlogits, R, Cc, P);
  return check_launch("lovasz_hinge_kernel");
}

extern "C" int eosvos_mask_loss_bce(const float* logits, const long long* labels, const float* targets,
                                    float* loss_out, float* dlogits, int R, int Cc, int P, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(loss_out, "mask_loss_bce: null loss_out");
  cudaError_t e = cudaMemsetAsync(loss_out, 0, sizeof(float), stream);
  if (e != cudaSuccess) return set_cuda_error(e, "mask_loss memset");
  if (R == 0) return 0;
  EOSVOS_REQUIRE(logits && labels && targets && dlogits, "mask_loss_bce: null pointer");
  bce_mask_kernel<<<R, 256, 0, stream>>>(logits, labels, targets, loss_out, dlogits, R, Cc, P);
  return check_launch("bce_mask_kernel");
}
