// DAVIS region similarity J and contour accuracy F of a whole sequence on the device: two launches produce, per
// (frame, object), the six integer counts the measures are ratios of.  Stage after the hot path:
// the reference reaches the external `davis` package from src/util/helper_func.py:444-458
// (db_eval_sequence(segmentations, annotations, measure)); the host restatement is util/metrics.py.
//   J = inter / union                                  counts[.,0], counts[.,1]
//   F = 2PR/(P+R), P = match_fg / n_fg, R = match_gt / n_gt   counts[.,2..5]
// Boundary maps follow seg2bmap (east / south / south-east neighbour differs, with its last-row / last-column
// rules); a boundary pixel is matched when the other mask has a boundary pixel within the disk x^2+y^2 <= r^2.
// All objects of a frame are handled together: a pixel's boundary membership is a K-bit mask (bit k-1 = object k).
#include "common.h"
#include "../../include/eosvos_b200.h"

namespace eosvos {

constexpr int JF_MAXK = 8;

// objects id0+1 .. id0+K of this call map to bits 0 .. K-1 (a sequence with more than 8 objects takes several calls)
__device__ __forceinline__ unsigned id_bit(int id, int id0, int K) {
  id -= id0;
  return (id >= 1 && id <= K) ? (1u << (id - 1)) : 0u;
}

// boundary bits of pixel (y, x) of an id map: object k is on the boundary here when exactly one of the pixel and a
// considered neighbour carries id k
__device__ __forceinline__ unsigned boundary_bits(const unsigned char* __restrict__ m, int y, int x, int H, int W,
                                                  int id0, int K) {
  const int p = m[(size_t)y * W + x];
  const bool last_row = (y == H - 1), last_col = (x == W - 1);
  if (last_row && last_col) return 0u;
  unsigned bits = 0u;
  const unsigned bp = id_bit(p, id0, K);
  if (!last_col) {  // east neighbour (the only one in the last row)
    const int e = m[(size_t)y * W + x + 1];
    if (e != p) bits |= bp | id_bit(e, id0, K);
  }
  if (!last_row) {  // south neighbour (the only one in the last column)
    const int s = m[(size_t)(y + 1) * W + x];
    if (s != p) bits |= bp | id_bit(s, id0, K);
  }
  if (!last_row && !last_col) {
    const int se = m[(size_t)(y + 1) * W + x + 1];
    if (se != p) bits |= bp | id_bit(se, id0, K);
  }
  return bits;
}

// counts[f][k][0..3] += inter, union, n_fg, n_gt; bmap[f][0|1][H][W] = boundary bits of pred | gt
__global__ void __launch_bounds__(256) jf_bmap_kernel(const unsigned char* __restrict__ pred,
                                                      const unsigned char* __restrict__ gt,
                                                      unsigned char* __restrict__ bmap, int* __restrict__ counts, int id0,
                                                      int K, int H, int W) {
  __shared__ int acc[JF_MAXK * 4];
  const int f = blockIdx.y;
  for (int i = threadIdx.x; i < K * 4; i += blockDim.x) acc[i] = 0;
  __syncthreads();
  const size_t hw = (size_t)H * W;
  const unsigned char* pf = pred + f * hw;
  const unsigned char* gf = gt + f * hw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (size_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
    const unsigned bp = boundary_bits(pf, y, x, H, W, id0, K), bg = boundary_bits(gf, y, x, H, W, id0, K);
    bmap[(f * 2 + 0) * hw + i] = (unsigned char)bp;
    bmap[(f * 2 + 1) * hw + i] = (unsigned char)bg;
    const unsigned ip = id_bit(pf[i], id0, K), ig = id_bit(gf[i], id0, K);
    if (ip | ig | bp | bg) {
      for (int k = 0; k < K; ++k) {
        const unsigned b = 1u << k;
        if (ip & ig & b) atomicAdd(&acc[k * 4 + 0], 1);
        if ((ip | ig) & b) atomicAdd(&acc[k * 4 + 1], 1);
        if (bp & b) atomicAdd(&acc[k * 4 + 2], 1);
        if (bg & b) atomicAdd(&acc[k * 4 + 3], 1);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * 4; i += blockDim.x)
    if (acc[i]) atomicAdd(&counts[(f * K + i / 4) * 6 + (i & 3)], acc[i]);
}

// counts[f][k][4] += pred boundary pixels of k with a gt boundary pixel of k inside the disk; [5] the converse
__global__ void __launch_bounds__(256) jf_match_kernel(const unsigned char* __restrict__ bmap, int* __restrict__ counts,
                                                       int K, int H, int W, int radius) {
  __shared__ int acc[JF_MAXK * 2];
  const int f = blockIdx.y;
  for (int i = threadIdx.x; i < K * 2; i += blockDim.x) acc[i] = 0;
  __syncthreads();
  const size_t hw = (size_t)H * W;
  const unsigned char* bp = bmap + (f * 2 + 0) * hw;
  const unsigned char* bg = bmap + (f * 2 + 1) * hw;
  const int r2 = radius * radius;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned mp = bp[i], mg = bg[i];
    if (!(mp | mg)) continue;
    const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
    unsigned near_p = 0u, near_g = 0u;  // boundary bits of pred / gt found inside the disk
    for (int dy = -radius; dy <= radius; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= H) continue;
      for (int dx = -radius; dx <= radius; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= W || dx * dx + dy * dy > r2) continue;
        const size_t j = (size_t)yy * W + xx;
        if (mg) near_p |= bp[j];
        if (mp) near_g |= bg[j];
      }
    }
    const unsigned hit_fg = mp & near_g, hit_gt = mg & near_p;
    for (int k = 0; k < K; ++k) {
      if (hit_fg & (1u << k)) atomicAdd(&acc[k * 2 + 0], 1);
      if (hit_gt & (1u << k)) atomicAdd(&acc[k * 2 + 1], 1);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * 2; i += blockDim.x)
    if (acc[i]) atomicAdd(&counts[(f * K + i / 2) * 6 + 4 + (i & 1)], acc[i]);
}

}  // namespace eosvos

using namespace eosvos;

extern "C" int eosvos_jf_counts(const unsigned char* pred, const unsigned char* gt, unsigned char* bmap, int* counts,
                                int T, int id0, int K, int H, int W, int radius, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(pred && gt && bmap && counts, "jf_counts: null pointer");
  EOSVOS_REQUIRE(K >= 1 && K <= JF_MAXK && id0 >= 0 && id0 + K <= 255, "jf_counts: 1..8 objects per call, ids <= 255");
  EOSVOS_REQUIRE(T > 0 && T <= 65535 && H > 0 && W > 0, "jf_counts: empty sequence or more than 65535 frames");
  EOSVOS_REQUIRE(radius >= 0 && radius <= 64, "jf_counts: radius out of range");
  cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)T * K * 6, stream);
  if (e != cudaSuccess) return set_cuda_error(e, "jf_counts memset");
  const size_t hw = (size_t)H * W;
  const unsigned gx = (unsigned)((hw + 255) / 256 < 1184 ? (hw + 255) / 256 : 1184);
  dim3 grid(gx, (unsigned)T);
  jf_bmap_kernel<<<grid, 256, 0, stream>>>(pred, gt, bmap, counts, id0, K, H, W);
  EOSVOS_TRY(check_launch("jf_bmap_kernel"));
  jf_match_kernel<<<grid, 256, 0, stream>>>(bmap, counts, K, H, W, radius);
  return check_launch("jf_match_kernel");
}
