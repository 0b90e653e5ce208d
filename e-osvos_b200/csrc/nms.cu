// Segmented non-maximum suppression for the RPN proposals and the detections.
// Replaces torchvision::nms as reached from filter_proposals (reference src/networks/mask_rcnn.py:249 ->
// tv rpn.py filter_proposals -> batched_nms) and postprocess_detections (mask_rcnn.py:392).
// torchvision resolves the IoU bit-matrix with a serial single-block pass (gather_keep_from_mask: ~3 ms per
// image at 8.7k boxes, 28 % of a fine-tune iteration); here every (image, FPN level) is its own segment --
// boxes of different levels never suppress each other under batched_nms -- and all segments of the batch are
// resolved concurrently, 64 boxes per step, in one launch.
// Input boxes are sorted by descending score inside each segment (top-k order).  IoU arithmetic is
// torchvision's devIoU / CPU nms (no coordinate offsets, like the CPU batched_nms "vanilla" path).
#include "common.h"
#include "../../include/eosvos_b200.h"

namespace eosvos {

__device__ __forceinline__ bool iou_gt(const float4 a, const float4 b, float thr) {
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float width = fmaxf(right - left, 0.f), height = fmaxf(bottom - top, 0.f);
  const float inter = width * height;
  const float sa = (a.z - a.x) * (a.w - a.y);
  const float sb = (b.z - b.x) * (b.w - b.y);
  return (inter / (sa + sb - inter)) > thr;
}

// grid (col blocks, row blocks, segments), 64 threads: mask[seg][row][colword] bit j = IoU(row, col*64+j) > thr
__global__ void __launch_bounds__(64)
nms_mask_kernel(const float4* __restrict__ boxes, const int* __restrict__ seg_off, unsigned long long* __restrict__ mask,
                int max_seg, int W, float thr) {
  const int seg = blockIdx.z, rb = blockIdx.y, cb = blockIdx.x;
  if (cb < rb) return;
  const int base = seg_off[seg];
  const int n = seg_off[seg + 1] - base;
  if (rb * 64 >= n || cb * 64 >= n) return;
  __shared__ float4 cbox[64];
  const int ccount = min(64, n - cb * 64);
  if ((int)threadIdx.x < ccount) cbox[threadIdx.x] = boxes[base + cb * 64 + threadIdx.x];
  __syncthreads();
  const int row = rb * 64 + threadIdx.x;
  if (row >= n) return;
  const float4 a = boxes[base + row];
  unsigned long long bits = 0;
  const int start = (rb == cb) ? threadIdx.x + 1 : 0;
  for (int j = start; j < ccount; ++j)
    if (iou_gt(a, cbox[j], thr)) bits |= 1ULL << j;
  mask[((size_t)seg * max_seg + row) * W + cb] = bits;
}

// one CTA per segment
__global__ void __launch_bounds__(1024)
nms_reduce_kernel(const int* __restrict__ seg_off, const unsigned long long* __restrict__ mask,
                  unsigned char* __restrict__ keep, int max_seg, int W) {
  extern __shared__ unsigned long long removed[];   // [W]
  __shared__ unsigned long long diag[64];
  __shared__ unsigned long long s_kept;
  const int seg = blockIdx.x;
  const int base = seg_off[seg];
  const int n = seg_off[seg + 1] - base;
  const int tid = threadIdx.x, lane = tid & 31;
  const int nb = (n + 63) >> 6;
  const unsigned long long* m = mask + (size_t)seg * max_seg * W;
  for (int i = tid; i < W; i += blockDim.x) removed[i] = 0;
  __syncthreads();
  for (int b = 0; b < nb; ++b) {
    if (tid < 64) {
      const int row = b * 64 + tid;
      diag[tid] = row < n ? m[(size_t)row * W + b] : 0ULL;
    }
    __syncthreads();
    if (tid == 0) {
      unsigned long long word = removed[b], kept = 0;
      const int cnt = min(64, n - b * 64);
      for (int i = 0; i < cnt; ++i) {
        if (!((word >> i) & 1ULL)) {
          kept |= 1ULL << i;
          word |= diag[i];
        }
      }
      s_kept = kept;
    }
    __syncthreads();
    const unsigned long long kept = s_kept;
    if (tid < 64 && b * 64 + tid < n) keep[base + b * 64 + tid] = (unsigned char)((kept >> tid) & 1ULL);
    const unsigned gmask = 0xFFu << (lane & ~7);
    for (int w = b + 1 + (tid >> 3); w < nb; w += 128) {
      unsigned long long acc = 0;
      for (int i = tid & 7; i < 64; i += 8)
        if ((kept >> i) & 1ULL) acc |= m[(size_t)(b * 64 + i) * W + w];
      acc |= __shfl_xor_sync(gmask, acc, 1);
      acc |= __shfl_xor_sync(gmask, acc, 2);
      acc |= __shfl_xor_sync(gmask, acc, 4);
      if ((tid & 7) == 0) removed[w] |= acc;
    }
    __syncthreads();
  }
}

}  // namespace eosvos

using namespace eosvos;

extern "C" long long eosvos_nms_scratch_bytes(int num_segments, int max_seg) {
  const long long W = (max_seg + 63) / 64;
  return (long long)num_segments * max_seg * W * 8;
}

// boxes [n][4] fp32 (x1,y1,x2,y2), sorted by descending score inside each segment; seg_off [S+1] int32 (device);
// keep [n] uint8 out; scratch: eosvos_nms_scratch_bytes(S, max_seg) bytes.
extern "C" int eosvos_nms_segments(const float* boxes, const int* seg_off, int num_segments, int max_seg, float thresh,
                                   void* scratch, unsigned char* keep, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (num_segments == 0 || max_seg == 0) return 0;
  EOSVOS_REQUIRE(boxes && seg_off && scratch && keep, "nms_segments: null pointer");
  EOSVOS_REQUIRE(max_seg <= 16384, "nms_segments: at most 16384 boxes per segment");
  const int W = (max_seg + 63) / 64;
  dim3 grid(W, W, num_segments);
  nms_mask_kernel<<<grid, 64, 0, stream>>>(reinterpret_cast<const float4*>(boxes), seg_off,
                                          reinterpret_cast<unsigned long long*>(scratch), max_seg, W, thresh);
  EOSVOS_TRY(check_launch("nms_mask_kernel"));
  nms_reduce_kernel<<<num_segments, 1024, W * sizeof(unsigned long long), stream>>>(
      seg_off, reinterpret_cast<const unsigned long long*>(scratch), keep, max_seg, W);
  return check_launch("nms_reduce_kernel");
}
