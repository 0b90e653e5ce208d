// Region-proposal and detection post-processing as fixed-shape device kernels.
//
// Reference call sites: src/networks/mask_rcnn.py:237-249 (BoxCoder.decode of ALL 257,796 x N anchors, then tv rpn.py
// filter_proposals: per-level top-k, clip, small-box / score filter, per-level NMS, post-NMS top-n), :251-332 (EXTEND /
// REPLACE proposal augmentation from the previous frame's box), :347-420 (postprocess_detections) and tv
// roi_heads.py:642-678 (assign_targets_to_proposals).  The reference reaches ~150 small ATen launches and several host
// synchronisations per image through these; here every stage is one launch over padded, statically shaped buffers:
//
//   rpn_keys_kernel / rpn_hist2_kernel / rpn_compact_kernel   two-level (11 + 11 bit) radix select of the top-k
//                                                              objectness logits of every (image, level) segment
//   rpn_sort_decode_kernel   bitonic sort of the <= 4096 survivors, box decoding of ONLY the selected anchors,
//                            clip, validity (min size, score threshold), sigmoid
//   (nms.cu)                 per-segment NMS on the decoded boxes
//   rpn_postnms_kernel       rank of every kept box among all levels of its image by merge-counting the sorted
//                            segments (binary searches over prefix sums of the keep flags) -> first post_nms_top_n
//   extend_boxes_kernel      jittered copies of the target box (CPU random numbers uploaded by the caller, same
//                            arithmetic order as the reference so the boxes are bit-equal)
//   det_top1_kernel          detections_per_img == 1 (multi_object 'single_id', evaluate.py:106-107): softmax, box
//                            decode, clip, score / size filter and arg-max in one launch (the best-scoring valid
//                            candidate always survives NMS)
//   roi_match_kernel         IoU matching of proposals (+ appended ground-truth boxes) with class labels and the
//                            foreground / background counts the sampler needs
//   roi_encode_kernel        regression targets + (image, box) rows of the sampled RoIs
// Arithmetic that feeds discrete decisions uses explicit round-to-nearest multiplies / adds (no FMA contraction), i.e.
// the operation sequence of the ATen elementwise kernels the reference runs.
#include "common.h"
#include "act.cuh"
#include "../../include/eosvos_b200.h"

namespace eosvos {

constexpr int RPN_MAX_LEVELS = 8;
constexpr int RPN_BINS = 2048;       // 11-bit digits
constexpr int RPN_CAP = 4096;        // survivors per segment handed to the sorter
constexpr int RPN_CHUNK = 4096;      // anchors per CTA in the streaming passes

struct RpnLevels {
  const float* head[RPN_MAX_LEVELS];   // fp32 [N * hw][16]: A objectness logits, then A x 4 box deltas
  int hw[RPN_MAX_LEVELS];              // pixels per image
  int anchor_off[RPN_MAX_LEVELS];      // first anchor of the level in the per-image anchor list
  int out_off[RPN_MAX_LEVELS];         // first slot of the level in the per-image candidate list
  int k[RPN_MAX_LEVELS];               // min(pre_nms_top_n, hw * A)
  int num_levels, A, anchors_per_image, cand_per_image;
};

__device__ __forceinline__ unsigned ord_key(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);     // ascending unsigned order == ascending float order
}

// Largest digit d with  #(keys whose digit > d) < need <= #(keys whose digit >= d); also returns that first count.
// hist: RPN_BINS counters (global or shared).  Called by all threads of a 256-thread CTA; result broadcast.
__device__ void find_digit(const unsigned* __restrict__ hist, int need, int* s_tmp /* >= 12 ints of smem */, int& digit,
                           int& above) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int PER = RPN_BINS / 256;
  unsigned local[PER];
  unsigned sum = 0;
#pragma unroll
  for (int j = 0; j < PER; ++j) {            // thread t owns bins (2047 - 8t) ... (2047 - 8t - 7), descending
    local[j] = hist[RPN_BINS - 1 - (tid * PER + j)];
    sum += local[j];
  }
  unsigned inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) s_tmp[warp] = (int)inc;
  if (tid == 0) s_tmp[10] = -1;
  __syncthreads();
  unsigned warp_base = 0;
  for (int w = 0; w < warp; ++w) warp_base += (unsigned)s_tmp[w];
  const unsigned before = warp_base + inc - sum;       // keys in bins above this thread's bins
  if ((int)before < need && need <= (int)(before + sum)) {
    unsigned acc = before;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      if ((int)acc < need && need <= (int)(acc + local[j])) {
        s_tmp[10] = RPN_BINS - 1 - (tid * PER + j);
        s_tmp[11] = (int)acc;
      }
      acc += local[j];
    }
  }
  __syncthreads();
  digit = s_tmp[10];
  above = s_tmp[11];
  __syncthreads();
}

// grid (chunks, levels, images) x 256
__global__ void __launch_bounds__(256)
rpn_keys_kernel(const RpnLevels lv, unsigned* __restrict__ keys, unsigned* __restrict__ hist1) {
  const int l = blockIdx.y, n = blockIdx.z;
  const int count = lv.hw[l] * lv.A;
  const int base = blockIdx.x * RPN_CHUNK;
  if (base >= count) return;
  __shared__ unsigned h[RPN_BINS];
  for (int b = threadIdx.x; b < RPN_BINS; b += 256) h[b] = 0;
  __syncthreads();
  const float* src = lv.head[l] + (size_t)n * lv.hw[l] * 16;
  unsigned* dst = keys + (size_t)n * lv.anchors_per_image + lv.anchor_off[l];
  const int end = min(base + RPN_CHUNK, count);
  for (int i = base + threadIdx.x; i < end; i += 256) {
    const int row = i / lv.A, a = i - row * lv.A;
    const unsigned key = ord_key(src[(size_t)row * 16 + a]);
    dst[i] = key;
    atomicAdd(&h[key >> 21], 1u);
  }
  __syncthreads();
  unsigned* g = hist1 + (size_t)(n * lv.num_levels + l) * RPN_BINS;
  for (int b = threadIdx.x; b < RPN_BINS; b += 256)
    if (h[b]) atomicAdd(&g[b], h[b]);
}

__global__ void __launch_bounds__(256)
rpn_hist2_kernel(const RpnLevels lv, const unsigned* __restrict__ keys, const unsigned* __restrict__ hist1,
                 unsigned* __restrict__ hist2) {
  const int l = blockIdx.y, n = blockIdx.z;
  const int count = lv.hw[l] * lv.A;
  const int base = blockIdx.x * RPN_CHUNK;
  if (base >= count) return;
  __shared__ unsigned h[RPN_BINS];
  __shared__ int s_tmp[12];
  const int seg = n * lv.num_levels + l;
  int d1, above1;
  find_digit(hist1 + (size_t)seg * RPN_BINS, lv.k[l], s_tmp, d1, above1);
  for (int b = threadIdx.x; b < RPN_BINS; b += 256) h[b] = 0;
  __syncthreads();
  const unsigned* src = keys + (size_t)n * lv.anchors_per_image + lv.anchor_off[l];
  const int end = min(base + RPN_CHUNK, count);
  for (int i = base + threadIdx.x; i < end; i += 256) {
    const unsigned key = src[i];
    if ((int)(key >> 21) == d1) atomicAdd(&h[(key >> 10) & (RPN_BINS - 1)], 1u);
  }
  __syncthreads();
  unsigned* g = hist2 + (size_t)seg * RPN_BINS;
  for (int b = threadIdx.x; b < RPN_BINS; b += 256)
    if (h[b]) atomicAdd(&g[b], h[b]);
}

__global__ void __launch_bounds__(256)
rpn_compact_kernel(const RpnLevels lv, const unsigned* __restrict__ keys, const unsigned* __restrict__ hist1,
                   const unsigned* __restrict__ hist2, unsigned* __restrict__ counters,
                   unsigned long long* __restrict__ list) {
  const int l = blockIdx.y, n = blockIdx.z;
  const int count = lv.hw[l] * lv.A;
  const int base = blockIdx.x * RPN_CHUNK;
  if (base >= count) return;
  __shared__ int s_tmp[12];
  const int seg = n * lv.num_levels + l;
  int d1, above1, d2, above2;
  find_digit(hist1 + (size_t)seg * RPN_BINS, lv.k[l], s_tmp, d1, above1);
  find_digit(hist2 + (size_t)seg * RPN_BINS, lv.k[l] - above1, s_tmp, d2, above2);
  const unsigned prefix22 = ((unsigned)d1 << 11) | (unsigned)d2;
  const unsigned* src = keys + (size_t)n * lv.anchors_per_image + lv.anchor_off[l];
  const int end = min(base + RPN_CHUNK, count);
  for (int i = base + threadIdx.x; i < end; i += 256) {
    const unsigned key = src[i];
    if ((key >> 10) >= prefix22) {
      const unsigned slot = atomicAdd(&counters[seg], 1u);
      if (slot < RPN_CAP) list[(size_t)seg * RPN_CAP + slot] = ((unsigned long long)key << 32) | (unsigned)(~(unsigned)i);
    }
  }
}

__device__ __forceinline__ void bitonic_desc(unsigned long long* s, int n /* power of two */) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < n / 2; t += blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int p = i | j;
        const bool desc = (i & k) == 0;
        const unsigned long long a = s[i], b = s[p];
        if ((a < b) == desc) {
          s[i] = b;
          s[p] = a;
        }
      }
      __syncthreads();
    }
  }
}

struct RpnDecodeArgs {
  float img_h[16], img_w[16];          // per image (resized, un-padded) size; N <= 16
  float clip, min_size, score_thresh;  // bbox_xform_clip, rpn.min_size, rpn.score_thresh
};

// grid (levels, images) x 1024
__global__ void __launch_bounds__(1024)
rpn_sort_decode_kernel(const RpnLevels lv, const RpnDecodeArgs args, const unsigned* __restrict__ counters,
                       const unsigned long long* __restrict__ list, const float4* __restrict__ anchors,
                       float4* __restrict__ boxes, float* __restrict__ scores, unsigned char* __restrict__ valid) {
  __shared__ unsigned long long s[RPN_CAP];
  const int l = blockIdx.x, n = blockIdx.y;
  const int seg = n * lv.num_levels + l;
  const int n_sel = min((int)counters[seg], RPN_CAP);
  int n_sort = 1;
  while (n_sort < n_sel) n_sort <<= 1;
  n_sort = max(n_sort, 2);
  for (int i = threadIdx.x; i < n_sort; i += blockDim.x) s[i] = i < n_sel ? list[(size_t)seg * RPN_CAP + i] : 0ULL;
  __syncthreads();
  bitonic_desc(s, n_sort);
  const int k = lv.k[l];
  const float* src = lv.head[l] + (size_t)n * lv.hw[l] * 16;
  const float4* anc = anchors + lv.anchor_off[l];
  const size_t out = (size_t)n * lv.cand_per_image + lv.out_off[l];
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
    float prob = 0.f;
    bool ok = false;
    if (j < n_sel) {
      const int idx = (int)(~(unsigned)(s[j] & 0xffffffffULL));
      const int row = idx / lv.A, a = idx - row * lv.A;
      const float* r = src + (size_t)row * 16;
      const float logit = r[a];
      const float dx = r[lv.A + a * 4 + 0], dy = r[lv.A + a * 4 + 1];
      const float dw = fminf(r[lv.A + a * 4 + 2], args.clip), dh = fminf(r[lv.A + a * 4 + 3], args.clip);
      const float4 an = anc[idx];
      // tv _utils.py BoxCoder.decode_single (weights 1, 1, 1, 1), one rounding per ATen elementwise op
      const float w = __fsub_rn(an.z, an.x), h = __fsub_rn(an.w, an.y);
      const float cx = __fadd_rn(an.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(an.y, __fmul_rn(0.5f, h));
      const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
      const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
      const float hw_ = __fmul_rn(0.5f, pw), hh_ = __fmul_rn(0.5f, ph);
      float x1 = __fsub_rn(pcx, hw_), y1 = __fsub_rn(pcy, hh_), x2 = __fadd_rn(pcx, hw_), y2 = __fadd_rn(pcy, hh_);
      const float W = args.img_w[n], H = args.img_h[n];
      x1 = fminf(fmaxf(x1, 0.f), W);
      x2 = fminf(fmaxf(x2, 0.f), W);
      y1 = fminf(fmaxf(y1, 0.f), H);
      y2 = fminf(fmaxf(y2, 0.f), H);
      prob = 1.f / (1.f + expf(-logit));
      ok = (x2 - x1) >= args.min_size && (y2 - y1) >= args.min_size && prob >= args.score_thresh;
      if (ok) box = make_float4(x1, y1, x2, y2);
    }
    boxes[out + j] = box;             // invalid boxes are zero: zero area, IoU 0 with everything, suppress nothing
    scores[out + j] = prob;
    valid[out + j] = ok ? 1 : 0;
  }
}

// One CTA per image.  Candidates are sorted by descending score inside every level segment; `keep` are the NMS flags.
// Final order = descending score over all kept boxes, ties by candidate index (== torch's stable argsort of the
// concatenated list): rank(i) = #kept before i in its own segment + sum over other segments of #kept with a larger
// score (or an equal score when the segment comes first).
__global__ void __launch_bounds__(1024)
rpn_postnms_kernel(const RpnLevels lv, const float4* __restrict__ boxes, const float* __restrict__ scores,
                   const unsigned char* __restrict__ valid, const unsigned char* __restrict__ keep, int post_n,
                   int out_stride, int out_offset, float4* __restrict__ out_boxes, float* __restrict__ out_scores,
                   int* __restrict__ out_count) {
  extern __shared__ unsigned char smem_raw[];
  const int n = blockIdx.x;
  const int C = lv.cand_per_image, L = lv.num_levels;
  float* s_score = reinterpret_cast<float*>(smem_raw);                 // [C]
  int* s_pre = reinterpret_cast<int*>(s_score + C);                    // [C + L]: exclusive prefix per segment (+ total)
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const size_t base = (size_t)n * C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < C; i += blockDim.x) s_score[i] = scores[base + i];
  // exclusive scan of the flags, segment by segment (block-wide, 1024 per round)
  for (int l = 0; l < L; ++l) {
    const int off = lv.out_off[l], k = lv.k[l];
    int* pre = s_pre + off + l;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int b = 0; b < k; b += blockDim.x) {
      const int i = b + tid;
      const int f = (i < k && keep[base + off + i] && valid[base + off + i]) ? 1 : 0;
      int inc = f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      if (lane == 31) s_warp[warp] = inc;
      __syncthreads();
      int wb = 0;
      for (int w = 0; w < warp; ++w) wb += s_warp[w];
      const int carry = s_carry;
      if (i < k) pre[i] = carry + wb + inc - f;
      __syncthreads();
      if (tid == blockDim.x - 1) s_carry = carry + wb + inc;
      __syncthreads();
    }
    if (tid == 0) pre[k] = s_carry;
    __syncthreads();
  }
  int total = 0;
  for (int l = 0; l < L; ++l) total += s_pre[lv.out_off[l] + l + lv.k[l]];
  const int n_out = min(total, post_n);
  if (tid == 0) out_count[n] = n_out;
  float4* ob = out_boxes + (size_t)n * out_stride + out_offset;
  float* os = out_scores ? out_scores + (size_t)n * out_stride + out_offset : nullptr;
  for (int i = tid; i < post_n; i += blockDim.x) {       // padding rows: zero boxes
    if (i >= n_out) {
      ob[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (os) os[i] = 0.f;
    }
  }
  for (int la = 0; la < L; ++la) {
    const int offa = lv.out_off[la], ka = lv.k[la];
    const int* prea = s_pre + offa + la;
    for (int q = tid; q < ka; q += blockDim.x) {
      if (prea[q + 1] == prea[q]) continue;              // not kept
      const float sc = s_score[offa + q];
      int rank = prea[q];
      for (int lb = 0; lb < L; ++lb) {
        if (lb == la) continue;
        const int offb = lv.out_off[lb], kb = lv.k[lb];
        const float* sb = s_score + offb;
        int lo = 0, hi = kb;                              // first position whose score is NOT "ahead of" sc
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          const bool ahead = lb < la ? (sb[mid] >= sc) : (sb[mid] > sc);
          if (ahead) lo = mid + 1; else hi = mid;
        }
        rank += s_pre[offb + lb + lo];
      }
      if (rank < post_n) {
        ob[rank] = boxes[base + offa + q];
        if (os) os[rank] = sc;
      }
    }
  }
}

// jittered copies of the (resized) target boxes: reference mask_rcnn.py:262-285.  rand: [B][G][4][n_aug] uniform
// numbers drawn on the host in the reference's order (x_min, y_min, x_max, y_max draws per box).  stats: int32
// [B][G][5] = (xmin, ymin, xmax, ymax, count) in input-frame pixels (mask_to_bbox / paste tail); a target without
// pixels falls back to `fallback` (the start target, helper_func.py:124-126).
__global__ void extend_boxes_kernel(const int* __restrict__ stats, const int* __restrict__ fallback,
                                    const float* __restrict__ rnd, float rw, float rh, float img_w, float img_h,
                                    float share, int G, int n_aug, int out_stride, int out_offset,
                                    float4* __restrict__ out) {
  const int b = blockIdx.y, g = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_aug) return;
  const int* st = stats + ((size_t)b * G + g) * 5;
  if (st[4] <= 0 && fallback) st = fallback + ((size_t)b * G + g) * 5;
  // resize_boxes (tv transform.py): integer pixel box (+1 on the max side, mask_rcnn.py:626-632) times the fp32 ratios
  const float x0 = __fmul_rn((float)st[0], rw), y0 = __fmul_rn((float)st[1], rh);
  const float x1 = __fmul_rn((float)(st[2] + 1), rw), y1 = __fmul_rn((float)(st[3] + 1), rh);
  const float bw = __fsub_rn(x1, x0), bh = __fsub_rn(y1, y0);
  const float* r = rnd + ((size_t)b * G + g) * 4 * n_aug;
  float4 o;
  o.x = __fsub_rn(x0, __fmul_rn(__fmul_rn(r[j], bw), share));
  o.y = __fsub_rn(y0, __fmul_rn(__fmul_rn(r[n_aug + j], bh), share));
  o.z = __fadd_rn(x1, __fmul_rn(__fmul_rn(r[2 * n_aug + j], bw), share));
  o.w = __fadd_rn(y1, __fmul_rn(__fmul_rn(r[3 * n_aug + j], bh), share));
  o.x = fminf(fmaxf(o.x, 0.f), img_w);
  o.y = fminf(fmaxf(o.y, 0.f), img_h);
  o.z = fminf(fmaxf(o.z, 0.f), img_w);
  o.w = fminf(fmaxf(o.w, 0.f), img_h);
  out[(size_t)b * out_stride + out_offset + g * n_aug + j] = o;
}

struct DetArgs {
  float wx, wy, ww, wh, clip;          // box coder weights (10, 10, 5, 5), bbox_xform_clip
  float score_thresh, min_size;        // roi_heads.score_thresh, 1e-2
  float img_w, img_h;                  // resized image size (clip)
  float back_w, back_h;                // ratios back to the input frame (tv transform.py postprocess)
};

// One CTA per image: best-scoring valid (row, class) candidate.  head: fp32 [B * R][16] = class logits [ncls], then
// box deltas [ncls][4].  Outputs per image: det_box [4] (input-frame coordinates), det_score, det_label (int64),
// det_row (int32: row * (ncls - 1) + class - 1, -1 when nothing passed), det_roi [5] = (image, box in resized
// coordinates) or (-1, 0, 0, 0, 0) -- the RoI of the mask branch --, chan [ncls - 1] (int32: index of the detection
// that fills class channel c, -1 = none).
__global__ void __launch_bounds__(1024)
det_top1_kernel(const float* __restrict__ head, const float4* __restrict__ props, const DetArgs a, int R, int ncls,
                float* __restrict__ det_box, float* __restrict__ det_score, long long* __restrict__ det_label,
                int* __restrict__ det_row, float* __restrict__ det_roi, int* __restrict__ chan) {
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float best = -1.f;
  int best_c = 0x7fffffff;
  float4 best_box = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = tid; r < R; r += blockDim.x) {
    const float* h = head + ((size_t)b * R + r) * 16;
    float m = h[0];
    for (int c = 1; c < ncls; ++c) m = fmaxf(m, h[c]);
    float den = 0.f;
    for (int c = 0; c < ncls; ++c) den += expf(h[c] - m);
    const float4 p = props[(size_t)b * R + r];
    const float w = __fsub_rn(p.z, p.x), hh = __fsub_rn(p.w, p.y);
    const float cx = __fadd_rn(p.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(p.y, __fmul_rn(0.5f, hh));
    for (int c = 1; c < ncls; ++c) {
      const float score = expf(h[c] - m) / den;
      if (!(score > a.score_thresh)) continue;
      const float* d = h + ncls + c * 4;
      const float dx = d[0] / a.wx, dy = d[1] / a.wy;
      const float dw = fminf(d[2] / a.ww, a.clip), dh = fminf(d[3] / a.wh, a.clip);
      const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, hh), cy);
      const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), hh);
      const float hw_ = __fmul_rn(0.5f, pw), hh_ = __fmul_rn(0.5f, ph);
      float x1 = __fsub_rn(pcx, hw_), y1 = __fsub_rn(pcy, hh_), x2 = __fadd_rn(pcx, hw_), y2 = __fadd_rn(pcy, hh_);
      x1 = fminf(fmaxf(x1, 0.f), a.img_w);
      x2 = fminf(fmaxf(x2, 0.f), a.img_w);
      y1 = fminf(fmaxf(y1, 0.f), a.img_h);
      y2 = fminf(fmaxf(y2, 0.f), a.img_h);
      if (!((x2 - x1) >= a.min_size && (y2 - y1) >= a.min_size)) continue;
      const int cand = r * (ncls - 1) + (c - 1);
      if (score > best || (score == best && cand < best_c)) {
        best = score;
        best_c = cand;
        best_box = make_float4(x1, y1, x2, y2);
      }
    }
  }
  // block arg-max: larger score wins, ties by the smaller candidate index (stable descending sort order)
  __shared__ float s_sc[32];
  __shared__ int s_c[32];
  __shared__ float4 s_b[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float os = __shfl_xor_sync(0xffffffffu, best, o);
    const int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
    float4 ob;
    ob.x = __shfl_xor_sync(0xffffffffu, best_box.x, o);
    ob.y = __shfl_xor_sync(0xffffffffu, best_box.y, o);
    ob.z = __shfl_xor_sync(0xffffffffu, best_box.z, o);
    ob.w = __shfl_xor_sync(0xffffffffu, best_box.w, o);
    if (os > best || (os == best && oc < best_c)) {
      best = os;
      best_c = oc;
      best_box = ob;
    }
  }
  if (lane == 0) {
    s_sc[warp] = best;
    s_c[warp] = best_c;
    s_b[warp] = best_box;
  }
  __syncthreads();
  if (tid == 0) {
    const int nw = blockDim.x >> 5;
    for (int w = 1; w < nw; ++w)
      if (s_sc[w] > best || (s_sc[w] == best && s_c[w] < best_c)) {
        best = s_sc[w];
        best_c = s_c[w];
        best_box = s_b[w];
      }
    const bool found = best >= 0.f;
    const int cls = found ? best_c % (ncls - 1) + 1 : 0;
    det_box[b * 4 + 0] = found ? __fmul_rn(best_box.x, a.back_w) : 0.f;
    det_box[b * 4 + 1] = found ? __fmul_rn(best_box.y, a.back_h) : 0.f;
    det_box[b * 4 + 2] = found ? __fmul_rn(best_box.z, a.back_w) : 0.f;
    det_box[b * 4 + 3] = found ? __fmul_rn(best_box.w, a.back_h) : 0.f;
    det_score[b] = found ? best : 0.f;
    det_label[b] = cls;
    det_row[b] = found ? best_c : -1;
    det_roi[b * 5 + 0] = found ? (float)b : -1.f;
    det_roi[b * 5 + 1] = found ? best_box.x : 0.f;
    det_roi[b * 5 + 2] = found ? best_box.y : 0.f;
    det_roi[b * 5 + 3] = found ? best_box.z : 0.f;
    det_roi[b * 5 + 4] = found ? best_box.w : 0.f;
    for (int c = 1; c < ncls; ++c) chan[b * (ncls - 1) + c - 1] = (found && c == cls) ? b : -1;
  }
}

// tv roi_heads.py assign_targets_to_proposals for a padded proposal list.  Rows of image b: P proposal slots (the
// first count[b] are real) followed by the image's ground-truth boxes (add_gt_proposals); gt_off [B + 1].
// labels int64: class of the matched ground truth (IoU >= thr), 0 background, -1 padding;  matched int64: arg-max
// ground truth (first maximum), clamped at 0;  counts int32 [B][2] = (#foreground, #background), caller-zeroed.
__global__ void __launch_bounds__(256)
roi_match_kernel(const float4* __restrict__ props, const int* __restrict__ count, const float4* __restrict__ gt_boxes,
                 const long long* __restrict__ gt_labels, const int* __restrict__ gt_off, int P, int rows_per_image,
                 float thr, float4* __restrict__ all_boxes, long long* __restrict__ labels,
                 long long* __restrict__ matched, int* __restrict__ counts) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int g0 = gt_off[b], G = gt_off[b + 1] - g0;
  int fg = 0, bg = 0;
  if (i < rows_per_image) {
    const bool is_gt = i >= P;
    const bool real = is_gt ? (i - P) < G : i < count[b];
    float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
    long long lab = -1, mi = 0;
    if (real) {
      box = is_gt ? gt_boxes[g0 + i - P] : props[(size_t)b * P + i];
      const float area = __fmul_rn(__fsub_rn(box.z, box.x), __fsub_rn(box.w, box.y));
      float best = -1.f;
      for (int g = 0; g < G; ++g) {
        const float4 t = gt_boxes[g0 + g];
        const float ta = __fmul_rn(__fsub_rn(t.z, t.x), __fsub_rn(t.w, t.y));
        const float w = fmaxf(__fsub_rn(fminf(box.z, t.z), fmaxf(box.x, t.x)), 0.f);
        const float h = fmaxf(__fsub_rn(fminf(box.w, t.w), fmaxf(box.y, t.y)), 0.f);
        const float inter = __fmul_rn(w, h);
        const float iou = inter / __fsub_rn(__fadd_rn(ta, area), inter);
        if (iou > best) {
          best = iou;
          mi = g;
        }
      }
      if (G == 0) {
        lab = 0;
      } else if (best >= thr) {
        lab = gt_labels[g0 + mi];
      } else {
        lab = 0;
      }
      fg = lab >= 1;
      bg = lab == 0;
    }
    const size_t o = (size_t)b * rows_per_image + i;
    all_boxes[o] = box;
    labels[o] = lab;
    matched[o] = mi;
  }
  fg = __reduce_add_sync(0xffffffffu, fg);
  bg = __reduce_add_sync(0xffffffffu, bg);
  if ((threadIdx.x & 31) == 0) {
    if (fg) atomicAdd(&counts[b * 2 + 0], fg);
    if (bg) atomicAdd(&counts[b * 2 + 1], bg);
  }
}

// tv _utils.py BalancedPositiveNegativeSampler given the two `torch.randperm` draws the reference makes per image
// (perm_pos over the foreground candidates, perm_neg over the background candidates, both in row order):
// selected = { fg[perm_pos[j]] : j < num_pos } U { bg[perm_neg[j]] : j < num_neg }, emitted in ascending row order
// (== torch.where(mask) of the reference), plus the positions of the foreground rows inside that list.
// table int64 [B][4] = (perm_pos ptr, perm_neg ptr, num_pos, num_neg).  Three passes (the RPN runs this over 257,796
// anchors per image): per-chunk candidate counts -> ranks + selection bitmap -> compaction by one CTA per image.
constexpr int SMP_CHUNK = 4096;

__global__ void __launch_bounds__(256)
sample_count_kernel(const long long* __restrict__ labels, int rows, int nchunks, int* __restrict__ chunk_cnt) {
  const int c = blockIdx.x, b = blockIdx.y;
  const long long* lab = labels + (size_t)b * rows;
  int fg = 0, bg = 0;
  const int end = min((c + 1) * SMP_CHUNK, rows);
  for (int i = c * SMP_CHUNK + threadIdx.x; i < end; i += 256) {
    const long long l = lab[i];
    fg += l >= 1;
    bg += l == 0;
  }
  fg = __reduce_add_sync(0xffffffffu, fg);
  bg = __reduce_add_sync(0xffffffffu, bg);
  __shared__ int s_f[8], s_b[8];
  if ((threadIdx.x & 31) == 0) {
    s_f[threadIdx.x >> 5] = fg;
    s_b[threadIdx.x >> 5] = bg;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int f = 0, g = 0;
    for (int w = 0; w < 8; ++w) {
      f += s_f[w];
      g += s_b[w];
    }
    chunk_cnt[((size_t)b * nchunks + c) * 2 + 0] = f;
    chunk_cnt[((size_t)b * nchunks + c) * 2 + 1] = g;
  }
}

__global__ void __launch_bounds__(256)
sample_select_kernel(const long long* __restrict__ labels, const long long* __restrict__ table, int rows, int nchunks,
                     const int* __restrict__ chunk_cnt, unsigned* __restrict__ selbits) {
  const int c = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long* lab = labels + (size_t)b * rows;
  __shared__ unsigned fgsel[SMP_CHUNK / 32], bgsel[SMP_CHUNK / 32];   // drawn ranks that fall into this chunk
  __shared__ int s_f[8], s_b[8];
  for (int i = tid; i < SMP_CHUNK / 32; i += 256) {
    fgsel[i] = 0;
    bgsel[i] = 0;
  }
  int pre_f = 0, pre_b = 0;
  for (int k = 0; k < c; ++k) {
    pre_f += chunk_cnt[((size_t)b * nchunks + k) * 2 + 0];
    pre_b += chunk_cnt[((size_t)b * nchunks + k) * 2 + 1];
  }
  __syncthreads();
  const long long* perm_pos = reinterpret_cast<const long long*>(table[b * 4 + 0]);
  const long long* perm_neg = reinterpret_cast<const long long*>(table[b * 4 + 1]);
  const int num_pos = (int)table[b * 4 + 2], num_neg = (int)table[b * 4 + 3];
  for (int j = tid; j < num_pos; j += 256) {
    const long long r = perm_pos[j] - pre_f;
    if (r >= 0 && r < SMP_CHUNK) atomicOr(&fgsel[r >> 5], 1u << (r & 31));
  }
  for (int j = tid; j < num_neg; j += 256) {
    const long long r = perm_neg[j] - pre_b;
    if (r >= 0 && r < SMP_CHUNK) atomicOr(&bgsel[r >> 5], 1u << (r & 31));
  }
  // every warp owns SMP_CHUNK / 8 consecutive rows: first its candidate counts, then ranks by ballots
  constexpr int PER_WARP = SMP_CHUNK / 8;
  const int w0 = c * SMP_CHUNK + warp * PER_WARP;
  int cf = 0, cb = 0;
  for (int o = 0; o < PER_WARP; o += 32) {
    const int i = w0 + o + lane;
    const long long l = i < rows ? lab[i] : -1;
    cf += __popc(__ballot_sync(0xffffffffu, l >= 1));
    cb += __popc(__ballot_sync(0xffffffffu, l == 0));
  }
  if (lane == 0) {
    s_f[warp] = cf;
    s_b[warp] = cb;
  }
  __syncthreads();
  int rf = 0, rb = 0;                                   // chunk-local rank of the warp's first candidate
  for (int w = 0; w < warp; ++w) {
    rf += s_f[w];
    rb += s_b[w];
  }
  const unsigned below = (1u << lane) - 1u;
  for (int o = 0; o < PER_WARP; o += 32) {
    const int i = w0 + o + lane;
    const long long l = i < rows ? lab[i] : -1;
    const unsigned mf = __ballot_sync(0xffffffffu, l >= 1), mb = __ballot_sync(0xffffffffu, l == 0);
    bool sel = false;
    if (l >= 1) {
      const int r = rf + __popc(mf & below);
      sel = (fgsel[r >> 5] >> (r & 31)) & 1u;
    } else if (l == 0) {
      const int r = rb + __popc(mb & below);
      sel = (bgsel[r >> 5] >> (r & 31)) & 1u;
    }
    const unsigned word = __ballot_sync(0xffffffffu, sel);
    if (lane == 0 && w0 + o < rows) selbits[(size_t)b * ((rows + 31) / 32) + (w0 + o) / 32] = word;
    rf += __popc(mf);
    rb += __popc(mb);
  }
}

// one CTA per image: ascending rows of the set bits -> inds [S] (-1 padded); foreground positions -> pos_in [Pmax]
__global__ void __launch_bounds__(1024)
sample_compact_kernel(const long long* __restrict__ labels, const unsigned* __restrict__ selbits, int rows, int S,
                      int Pmax, long long* __restrict__ inds, long long* __restrict__ pos_in) {
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int words = (rows + 31) / 32;
  const int per = (words + blockDim.x - 1) / blockDim.x;
  const unsigned* bits = selbits + (size_t)b * words;
  const long long* lab = labels + (size_t)b * rows;
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  for (int i = tid; i < S; i += blockDim.x) inds[(size_t)b * S + i] = -1;
  for (int i = tid; i < Pmax; i += blockDim.x) pos_in[(size_t)b * Pmax + i] = -1;
  int cnt = 0;
  for (int k = 0; k < per; ++k) {
    const int wd = tid * per + k;
    if (wd < words) cnt += __popc(bits[wd]);
  }
  int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  int slot = inc - cnt;
  for (int w = 0; w < warp; ++w) slot += s_warp[w];
  for (int k = 0; k < per; ++k) {
    const int wd = tid * per + k;
    if (wd >= words) break;
    unsigned m = bits[wd];
    while (m) {
      const int bit = __ffs(m) - 1;
      m &= m - 1;
      if (slot < S) inds[(size_t)b * S + slot] = (long long)wd * 32 + bit;
      ++slot;
    }
  }
  __syncthreads();
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < S; base += blockDim.x) {
    const int j = base + tid;
    const long long r = j < S ? inds[(size_t)b * S + j] : -1;
    const int f = r >= 0 && lab[r] >= 1;
    int in2 = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, in2, o);
      if (lane >= o) in2 += t;
    }
    if (lane == 31) s_warp[warp] = in2;
    __syncthreads();
    int wb = 0;
    for (int w = 0; w < warp; ++w) wb += s_warp[w];
    const int sl = s_carry + wb + in2 - f;
    __syncthreads();
    if (tid == blockDim.x - 1) s_carry = sl + f;
    if (f && sl < Pmax) pos_in[(size_t)b * Pmax + sl] = j;
    __syncthreads();
  }
}

// Sampled RoIs -> (image, box) rows, labels, regression targets (tv _utils.py encode_boxes, weights wx..wh).
// inds int64 [B][S]: row inside the image's candidate list (roi_match_kernel layout).
__global__ void roi_encode_kernel(const float4* __restrict__ all_boxes, const long long* __restrict__ labels,
                                  const long long* __restrict__ matched, const float4* __restrict__ gt_boxes,
                                  const int* __restrict__ gt_off, const long long* __restrict__ inds, int S,
                                  int rows_per_image, float wx, float wy, float ww, float wh, float* __restrict__ rois5,
                                  long long* __restrict__ out_labels, long long* __restrict__ out_matched,
                                  float4* __restrict__ reg) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= S) return;
  const long long r = inds[(size_t)b * S + j];
  const size_t src = (size_t)b * rows_per_image + r;
  const float4 p = all_boxes[src];
  const long long mi = matched[src];
  const int G = gt_off[b + 1] - gt_off[b];
  const float4 t = G > 0 ? gt_boxes[gt_off[b] + mi] : make_float4(0.f, 0.f, 0.f, 0.f);
  const size_t o = (size_t)b * S + j;
  rois5[o * 5 + 0] = (float)b;
  rois5[o * 5 + 1] = p.x;
  rois5[o * 5 + 2] = p.y;
  rois5[o * 5 + 3] = p.z;
  rois5[o * 5 + 4] = p.w;
  out_labels[o] = labels[src];
  out_matched[o] = mi;
  const float ex_w = __fsub_rn(p.z, p.x), ex_h = __fsub_rn(p.w, p.y);
  const float ex_cx = __fadd_rn(p.x, __fmul_rn(0.5f, ex_w)), ex_cy = __fadd_rn(p.y, __fmul_rn(0.5f, ex_h));
  const float gt_w = __fsub_rn(t.z, t.x), gt_h = __fsub_rn(t.w, t.y);
  const float gt_cx = __fadd_rn(t.x, __fmul_rn(0.5f, gt_w)), gt_cy = __fadd_rn(t.y, __fmul_rn(0.5f, gt_h));
  float4 e;
  e.x = __fmul_rn(wx, __fsub_rn(gt_cx, ex_cx)) / ex_w;
  e.y = __fmul_rn(wy, __fsub_rn(gt_cy, ex_cy)) / ex_h;
  e.z = __fmul_rn(ww, logf(gt_w / ex_w));
  e.w = __fmul_rn(wh, logf(gt_h / ex_h));
  reg[o] = e;
}

// tv rpn.py assign_targets_to_anchors: box_iou(gt, anchors) + Matcher(0.7, 0.3, allow_low_quality_matches=True).
// Pass 1: per ground-truth box the highest IoU over all anchors (IoU >= 0: the float bit pattern orders like an
// unsigned int, so atomicMax works).  Pass 2: labels (1 foreground / 0 background / -1 between the thresholds, as int64
// for the sampler), arg-max ground truth, per-image counts.  An anchor that attains a box's highest IoU keeps its
// arg-max match even below the thresholds (tv _utils.py Matcher.set_low_quality_matches_).
__device__ __forceinline__ float box_iou_tv(const float4 g, const float ga, const float4 a, const float aa) {
  const float w = fmaxf(__fsub_rn(fminf(g.z, a.z), fmaxf(g.x, a.x)), 0.f);
  const float h = fmaxf(__fsub_rn(fminf(g.w, a.w), fmaxf(g.y, a.y)), 0.f);
  const float inter = __fmul_rn(w, h);
  return inter / __fsub_rn(__fadd_rn(ga, aa), inter);
}

template <int PASS>
__global__ void __launch_bounds__(256)
anchor_match_kernel(const float4* __restrict__ anchors, int num_anchors, const float4* __restrict__ gt,
                    const int* __restrict__ gt_off, float hi, float lo, unsigned* __restrict__ gt_best,
                    long long* __restrict__ labels, int* __restrict__ matched, int* __restrict__ counts) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int g0 = gt_off[b], G = gt_off[b + 1] - g0;
  int fg = 0, bg = 0;
  if (i < num_anchors) {
    const float4 a = anchors[i];
    const float aa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    float best = -1.f;
    int mi = 0;
    bool low_quality = false;
    for (int g = 0; g < G; ++g) {
      const float4 t = gt[g0 + g];
      const float ta = __fmul_rn(__fsub_rn(t.z, t.x), __fsub_rn(t.w, t.y));
      const float iou = box_iou_tv(t, ta, a, aa);
      if (PASS == 1) {
        atomicMax(&gt_best[g0 + g], __float_as_uint(iou));
      } else {
        if (iou > best) {
          best = iou;
          mi = g;
        }
        low_quality |= __float_as_uint(iou) == gt_best[g0 + g];
      }
    }
    if (PASS == 2) {
      long long lab;
      if (G == 0) lab = 0;
      else if (best >= hi || low_quality) lab = 1;
      else if (best < lo) lab = 0;
      else lab = -1;
      labels[(size_t)b * num_anchors + i] = lab;
      matched[(size_t)b * num_anchors + i] = mi;
      fg = lab == 1;
      bg = lab == 0;
    }
  }
  if (PASS == 2) {
    fg = __reduce_add_sync(0xffffffffu, fg);
    bg = __reduce_add_sync(0xffffffffu, bg);
    if ((threadIdx.x & 31) == 0) {
      if (fg) atomicAdd(&counts[b * 2 + 0], fg);
      if (bg) atomicAdd(&counts[b * 2 + 1], bg);
    }
  }
}

// tv rpn.py compute_loss on the sampled anchors, read straight from the per-level head outputs (no concatenated
// [N * 257,796] tensors, no index / cat autograd nodes).  idx: positions in the reference's flattened order
// (image, level, pixel, anchor); the first n_pos entries are the foreground samples.
// out[0] = BCE-with-logits mean over the n_pos + n_neg samples, out[1] = smooth-L1(beta) sum over the foreground / M.
struct RpnLossLevels {
  const float* head[RPN_MAX_LEVELS];
  float* dy[RPN_MAX_LEVELS];
  int hw[RPN_MAX_LEVELS];
  int anchor_off[RPN_MAX_LEVELS];
  int num_levels, A, anchors_per_image;
};

__device__ __forceinline__ size_t rpn_locate(const RpnLossLevels& lv, long long idx, int& l, int& a) {
  const int n = (int)(idx / lv.anchors_per_image);
  const int r = (int)(idx - (long long)n * lv.anchors_per_image);
  l = 0;
  while (l + 1 < lv.num_levels && r >= lv.anchor_off[l + 1]) ++l;
  const int i = r - lv.anchor_off[l];
  const int row = i / lv.A;
  a = i - row * lv.A;
  return ((size_t)n * lv.hw[l] + row) * 16;
}

template <bool BWD>
__global__ void __launch_bounds__(1024)
rpn_loss_kernel(const RpnLossLevels lv, const long long* __restrict__ sampled, int M, const long long* __restrict__ labels,
                const int* __restrict__ matched, const float4* __restrict__ anchors, const float4* __restrict__ gt,
                const int* __restrict__ gt_off, float beta, float* __restrict__ out, const float* __restrict__ g_obj,
                const float* __restrict__ g_box) {
  const float inv_m = 1.f / (float)M;
  float bce = 0.f, box = 0.f;
  const float go = BWD ? g_obj[0] * inv_m : 0.f, gb = BWD ? g_box[0] * inv_m : 0.f;
  for (int s = threadIdx.x; s < M; s += blockDim.x) {
    const long long idx = sampled[s];
    if (idx < 0) continue;                               // padding of a short sample
    int l, a;
    const size_t base = rpn_locate(lv, idx, l, a);
    const float* h = lv.head[l] + base;
    const long long lab = labels[idx];
    const float x = h[a], y = lab >= 1 ? 1.f : 0.f;
    if (BWD) {
      lv.dy[l][base + a] = go * (1.f / (1.f + expf(-x)) - y);
    } else {
      bce += fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));
    }
    if (lab >= 1) {
      // regression target of this anchor (tv _utils.py encode_boxes, weights 1): only the sampled foreground needs it
      const int n = (int)(idx / lv.anchors_per_image);
      const float4 an = anchors[idx - (long long)n * lv.anchors_per_image];
      const float4 t = gt[gt_off[n] + matched[idx]];
      const float ew = __fsub_rn(an.z, an.x), eh = __fsub_rn(an.w, an.y);
      const float ecx = __fadd_rn(an.x, __fmul_rn(0.5f, ew)), ecy = __fadd_rn(an.y, __fmul_rn(0.5f, eh));
      const float gw = __fsub_rn(t.z, t.x), gh = __fsub_rn(t.w, t.y);
      const float gcx = __fadd_rn(t.x, __fmul_rn(0.5f, gw)), gcy = __fadd_rn(t.y, __fmul_rn(0.5f, gh));
      const float tv[4] = {__fsub_rn(gcx, ecx) / ew, __fsub_rn(gcy, ecy) / eh, logf(gw / ew), logf(gh / eh)};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float d = h[lv.A + a * 4 + c] - tv[c];
        const float ad = fabsf(d);
        if (BWD) {
          lv.dy[l][base + lv.A + a * 4 + c] = gb * (ad < beta ? d / beta : (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)));
        } else {
          box += ad < beta ? 0.5f * d * d / beta : ad - 0.5f * beta;
        }
      }
    }
  }
  if (!BWD) {
    __shared__ float s_a[32], s_b[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      bce += __shfl_xor_sync(0xffffffffu, bce, o);
      box += __shfl_xor_sync(0xffffffffu, box, o);
    }
    if ((threadIdx.x & 31) == 0) {
      s_a[threadIdx.x >> 5] = bce;
      s_b[threadIdx.x >> 5] = box;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float ta = 0.f, tb = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
        ta += s_a[w];
        tb += s_b[w];
      }
      out[0] = ta * inv_m;
      out[1] = tb * inv_m;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Sparse backward of the RPN head.  The RPN losses touch only the <= 256 sampled anchors per image, so the gradient
// that enters the head (1x1 objectness / box convs on the shared 3x3 conv's ReLU output t) is non-zero on <= 768
// pixel rows out of 257,796 x N / 3: instead of dense dgrad + wgrad over every pyramid level (2 x 102 GFLOP per
// image in the reference's cuDNN path) the gradient is propagated for those rows only:
//   rpn_sparse_head_kernel   per sampled anchor ("event"): d loss / d head (as rpn_loss_kernel<true>), the 1x1 heads'
//                            weight / bias gradients, dt = W_head^T dy masked by t > 0, event pixel coordinates
//   rpn_sparse_gather_kernel X_g[e][tap][ci] = f[pixel_e + tap][ci]   (the 3x3 conv's input rows, zero outside)
//   (tensor cores)           dW_conv = dt^T X_g  (eosvos_gemm_wgrad),  G = dt W_conv  (eosvos_conv2d_fprop on rows)
//   rpn_sparse_scatter_kernel df[pixel_e + tap][ci] += G[e][ci][tap]  (16-bit atomics into the zeroed level maps)
// Mathematically identical to the dense backward (the skipped terms are exact zeros).
// ---------------------------------------------------------------------------------------------------------------
struct RpnSparseLevels {
  const float* head[RPN_MAX_LEVELS];    // fp32 [N * hw][16]
  const act_t* t[RPN_MAX_LEVELS];       // RPN conv output after ReLU, [N * hw][C]
  const act_t* f[RPN_MAX_LEVELS];       // pyramid level (conv input), [N][H][W][C]
  act_t* df[RPN_MAX_LEVELS];            // gradient of the level, zeroed by the caller
  int H[RPN_MAX_LEVELS], W[RPN_MAX_LEVELS];
  int anchor_off[RPN_MAX_LEVELS];
  int num_levels, A, anchors_per_image, C;
};

// grid = events, block = C threads (C <= 1024, multiple of 32)
__global__ void rpn_sparse_head_kernel(const RpnSparseLevels lv, const long long* __restrict__ sampled, int M,
                                       const long long* __restrict__ labels, const int* __restrict__ matched,
                                       const float4* __restrict__ anchors, const float4* __restrict__ gt,
                                       const int* __restrict__ gt_off, float beta, const float* __restrict__ g_obj,
                                       const float* __restrict__ g_box, const float* __restrict__ w_cls,
                                       const float* __restrict__ w_box, float scale, act_t* __restrict__ dt,
                                       int4* __restrict__ ev_pix, float* __restrict__ dw_cls, float* __restrict__ db_cls,
                                       float* __restrict__ dw_box, float* __restrict__ db_box) {
  const int e = blockIdx.x, c = threadIdx.x;
  const long long idx = sampled[e];
  const int C = lv.C, A = lv.A;
  if (idx < 0) {                                   // padding event: contributes nothing
    dt[(size_t)e * C + c] = float2act(0.f);
    if (c == 0) ev_pix[e] = make_int4(-1, 0, 0, 0);
    return;
  }
  const int n = (int)(idx / lv.anchors_per_image);
  const int r = (int)(idx - (long long)n * lv.anchors_per_image);
  int l = 0;
  while (l + 1 < lv.num_levels && r >= lv.anchor_off[l + 1]) ++l;
  const int i = r - lv.anchor_off[l];
  const int row = i / A, a = i - row * A;
  const size_t prow = (size_t)n * lv.H[l] * lv.W[l] + row;
  const float* h = lv.head[l] + prow * 16;
  const long long lab = labels[idx];
  const float inv_m = 1.f / (float)M;
  const float go = g_obj[0] * inv_m, gb = g_box[0] * inv_m;
  // d loss / d head for this anchor: 1 objectness column, 4 box columns when foreground
  float dy[5];
  const float x = h[a], y = lab >= 1 ? 1.f : 0.f;
  dy[0] = go * (1.f / (1.f + expf(-x)) - y);
  dy[1] = dy[2] = dy[3] = dy[4] = 0.f;
  if (lab >= 1) {
    const float4 an = anchors[r];
    const float4 t4 = gt[gt_off[n] + matched[idx]];
    const float ew = __fsub_rn(an.z, an.x), eh = __fsub_rn(an.w, an.y);
    const float ecx = __fadd_rn(an.x, __fmul_rn(0.5f, ew)), ecy = __fadd_rn(an.y, __fmul_rn(0.5f, eh));
    const float gw = __fsub_rn(t4.z, t4.x), gh = __fsub_rn(t4.w, t4.y);
    const float gcx = __fadd_rn(t4.x, __fmul_rn(0.5f, gw)), gcy = __fadd_rn(t4.y, __fmul_rn(0.5f, gh));
    const float tv[4] = {__fsub_rn(gcx, ecx) / ew, __fsub_rn(gcy, ecy) / eh, logf(gw / ew), logf(gh / eh)};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float d = h[A + a * 4 + k] - tv[k];
      const float ad = fabsf(d);
      dy[1 + k] = gb * (ad < beta ? d / beta : (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)));
    }
  }
  const float tc = (float)lv.t[l][prow * C + c];
  // heads' parameter gradients (unscaled fp32): dW[col][c] += dy[col] * t[c], db[col] += dy[col]
  atomicAdd(&dw_cls[(size_t)a * C + c], dy[0] * tc);
  float acc = dy[0] * w_cls[(size_t)a * C + c];
  if (lab >= 1) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      atomicAdd(&dw_box[(size_t)(a * 4 + k) * C + c], dy[1 + k] * tc);
      acc += dy[1 + k] * w_box[(size_t)(a * 4 + k) * C + c];
    }
  }
  if (c == 0) {
    atomicAdd(&db_cls[a], dy[0]);
    if (lab >= 1) {
#pragma unroll
      for (int k = 0; k < 4; ++k) atomicAdd(&db_box[a * 4 + k], dy[1 + k]);
    }
    const int py = row / lv.W[l];
    ev_pix[e] = make_int4(l, n, py, row - py * lv.W[l]);
  }
  // gradient w.r.t. the conv output, through the ReLU, entering the 16-bit domain under the loss scale
  dt[(size_t)e * C + c] = float2act(tc > 0.f ? acc * scale : 0.f);
}

// grid (events, 9 taps), block C / 2 threads (two channels each)
__global__ void rpn_sparse_gather_kernel(const RpnSparseLevels lv, const int4* __restrict__ ev_pix,
                                         act2_t* __restrict__ xg) {
  const int e = blockIdx.x, tap = blockIdx.y, c2 = threadIdx.x;
  const int C2 = lv.C / 2;
  const int4 p = ev_pix[e];
  act2_t v = floats2act2(0.f, 0.f);
  if (p.x >= 0) {
    const int y = p.z + tap / 3 - 1, x = p.w + tap % 3 - 1;
    if (y >= 0 && y < lv.H[p.x] && x >= 0 && x < lv.W[p.x])
      v = reinterpret_cast<const act2_t*>(lv.f[p.x])[(((size_t)p.y * lv.H[p.x] + y) * lv.W[p.x] + x) * C2 + c2];
  }
  xg[((size_t)e * 9 + tap) * C2 + c2] = v;
}

// G [events][C * 9] with column = ci * 9 + tap (rows of the [Cin][KH][KW][Cout] dgrad operand).
// grid (events, 9 taps), block C threads
__global__ void rpn_sparse_scatter_kernel(const RpnSparseLevels lv, const int4* __restrict__ ev_pix,
                                          const act_t* __restrict__ G) {
  const int e = blockIdx.x, tap = blockIdx.y, ci = threadIdx.x;
  const int4 p = ev_pix[e];
  if (p.x < 0) return;
  const int y = p.z + tap / 3 - 1, x = p.w + tap % 3 - 1;
  if (y < 0 || y >= lv.H[p.x] || x < 0 || x >= lv.W[p.x]) return;
  const act_t g = G[(size_t)e * lv.C * 9 + (size_t)ci * 9 + tap];
  atomicAdd(&lv.df[p.x][(((size_t)p.y * lv.H[p.x] + y) * lv.W[p.x] + x) * lv.C + ci], g);
}

}  // namespace eosvos

using namespace eosvos;

static int fill_levels(RpnLevels& lv, const void* const* heads, const int* hw, int num_levels, int A, int pre_nms_top_n) {
  if (num_levels < 1 || num_levels > RPN_MAX_LEVELS) return -1;
  lv.num_levels = num_levels;
  lv.A = A;
  int aoff = 0, ooff = 0;
  for (int l = 0; l < num_levels; ++l) {
    lv.head[l] = reinterpret_cast<const float*>(heads[l]);
    lv.hw[l] = hw[l];
    lv.anchor_off[l] = aoff;
    lv.out_off[l] = ooff;
    lv.k[l] = hw[l] * A < pre_nms_top_n ? hw[l] * A : pre_nms_top_n;
    aoff += hw[l] * A;
    ooff += lv.k[l];
  }
  lv.anchors_per_image = aoff;
  lv.cand_per_image = ooff;
  return 0;
}

extern "C" long long eosvos_rpn_scratch_bytes(int N, const int* hw, int num_levels, int A) {
  long long anchors = 0;
  for (int l = 0; l < num_levels; ++l) anchors += (long long)hw[l] * A;
  const long long segs = (long long)N * num_levels;
  // [zeroed by the caller: hist1, hist2, counters] [keys] [survivor lists]
  return segs * (2 * RPN_BINS + 1) * 4 + N * anchors * 4 + segs * RPN_CAP * 8 + 64;
}

extern "C" long long eosvos_rpn_scratch_zero_bytes(int N, int num_levels) {
  return (long long)N * num_levels * (2 * RPN_BINS + 1) * 4;
}

// Per-level top-k + decode + clip + validity of the RPN head outputs.  heads[l]: fp32 [N * hw[l]][16]; anchors: fp32
// [sum hw*A][4] (one image's anchors, identical for every image of the batch); image_hw: host [N][2] (h, w) of the
// resized images.  Candidate list per image: level after level, min(pre_nms_top_n, hw*A) slots each, sorted by
// descending objectness.  scratch: eosvos_rpn_scratch_bytes(...) bytes whose first eosvos_rpn_scratch_zero_bytes(...)
// bytes are zero.
extern "C" int eosvos_rpn_select(const void* const* heads, const int* hw, int num_levels, int A, int N,
                                 const float* anchors, const float* image_hw, int pre_nms_top_n, float bbox_clip,
                                 float min_size, float score_thresh, void* scratch, float* boxes, float* scores,
                                 unsigned char* valid, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RpnLevels lv;
  EOSVOS_REQUIRE(fill_levels(lv, heads, hw, num_levels, A, pre_nms_top_n) == 0, "rpn_select: 1..8 pyramid levels");
  EOSVOS_REQUIRE(N >= 1 && N <= 16, "rpn_select: 1..16 images per call");
  EOSVOS_REQUIRE(pre_nms_top_n >= 1 && pre_nms_top_n <= 2048, "rpn_select: pre_nms_top_n must be in 1..2048");
  EOSVOS_REQUIRE(anchors && image_hw && scratch && boxes && scores && valid, "rpn_select: null pointer");
  const long long segs = (long long)N * num_levels;
  unsigned* hist1 = reinterpret_cast<unsigned*>(scratch);
  unsigned* hist2 = hist1 + segs * RPN_BINS;
  unsigned* counters = hist2 + segs * RPN_BINS;
  unsigned* keys = counters + segs;
  size_t off = (size_t)(segs * (2 * RPN_BINS + 1) + (long long)N * lv.anchors_per_image) * 4;
  off = (off + 15) / 16 * 16;
  unsigned long long* list = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(scratch) + off);
  int max_count = 0;
  for (int l = 0; l < num_levels; ++l) max_count = hw[l] * A > max_count ? hw[l] * A : max_count;
  dim3 grid((max_count + RPN_CHUNK - 1) / RPN_CHUNK, num_levels, N);
  rpn_keys_kernel<<<grid, 256, 0, stream>>>(lv, keys, hist1);
  EOSVOS_TRY(check_launch("rpn_keys_kernel"));
  rpn_hist2_kernel<<<grid, 256, 0, stream>>>(lv, keys, hist1, hist2);
  EOSVOS_TRY(check_launch("rpn_hist2_kernel"));
  rpn_compact_kernel<<<grid, 256, 0, stream>>>(lv, keys, hist1, hist2, counters, list);
  EOSVOS_TRY(check_launch("rpn_compact_kernel"));
  RpnDecodeArgs da;
  for (int n = 0; n < N; ++n) {
    da.img_h[n] = image_hw[2 * n];
    da.img_w[n] = image_hw[2 * n + 1];
  }
  da.clip = bbox_clip;
  da.min_size = min_size;
  da.score_thresh = score_thresh;
  rpn_sort_decode_kernel<<<dim3(num_levels, N), 1024, 0, stream>>>(lv, da, counters, list,
                                                                  reinterpret_cast<const float4*>(anchors),
                                                                  reinterpret_cast<float4*>(boxes), scores, valid);
  return check_launch("rpn_sort_decode_kernel");
}

// Post-NMS selection: the first post_n kept candidates of every image in descending score order, written to
// out_boxes[n * out_stride + out_offset + rank] (padding rows zeroed), out_count[n] = number of real rows.
extern "C" int eosvos_rpn_postnms(const int* hw, int num_levels, int A, int N, int pre_nms_top_n, const float* boxes,
                                  const float* scores, const unsigned char* valid, const unsigned char* keep,
                                  int post_n, int out_stride, int out_offset, float* out_boxes, float* out_scores,
                                  int* out_count, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RpnLevels lv;
  const void* none[RPN_MAX_LEVELS] = {};
  EOSVOS_REQUIRE(fill_levels(lv, none, hw, num_levels, A, pre_nms_top_n) == 0, "rpn_postnms: 1..8 pyramid levels");
  EOSVOS_REQUIRE(boxes && scores && valid && keep && out_boxes && out_count, "rpn_postnms: null pointer");
  EOSVOS_REQUIRE(out_offset >= 0 && out_offset + post_n <= out_stride, "rpn_postnms: output window out of range");
  const size_t smem = (size_t)lv.cand_per_image * 4 + (size_t)(lv.cand_per_image + num_levels) * 4;
  EOSVOS_REQUIRE(smem <= 200 * 1024, "rpn_postnms: too many candidates per image for shared memory");
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(rpn_postnms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  rpn_postnms_kernel<<<N, 1024, smem, stream>>>(lv, reinterpret_cast<const float4*>(boxes), scores, valid, keep, post_n,
                                                out_stride, out_offset, reinterpret_cast<float4*>(out_boxes),
                                                out_scores, out_count);
  return check_launch("rpn_postnms_kernel");
}

extern "C" int eosvos_extend_boxes(const int* stats, const int* fallback_stats, const float* rnd, int B, int G,
                                   int n_aug, float ratio_w, float ratio_h, float img_w, float img_h, float share,
                                   int out_stride, int out_offset, float* out_boxes, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(stats && rnd && out_boxes, "extend_boxes: null pointer");
  EOSVOS_REQUIRE(out_offset >= 0 && out_offset + G * n_aug <= out_stride, "extend_boxes: output window out of range");
  dim3 grid((n_aug + 255) / 256, B, G);
  extend_boxes_kernel<<<grid, 256, 0, stream>>>(stats, fallback_stats, rnd, ratio_w, ratio_h, img_w, img_h, share, G,
                                                n_aug, out_stride, out_offset, reinterpret_cast<float4*>(out_boxes));
  return check_launch("extend_boxes_kernel");
}

extern "C" int eosvos_det_top1(const float* head, const float* proposals, int B, int R, int num_classes,
                               const float* coder_weights4, float bbox_clip, float score_thresh, float min_size,
                               float img_w, float img_h, float back_w, float back_h, float* det_box, float* det_score,
                               long long* det_label, int* det_row, float* det_roi, int* chan, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(head && proposals && det_box && det_score && det_label && det_row && det_roi && chan,
                 "det_top1: null pointer");
  EOSVOS_REQUIRE(num_classes >= 2 && num_classes * 5 <= 16, "det_top1: the fused head row holds at most 3 classes");
  DetArgs a;
  a.wx = coder_weights4[0];
  a.wy = coder_weights4[1];
  a.ww = coder_weights4[2];
  a.wh = coder_weights4[3];
  a.clip = bbox_clip;
  a.score_thresh = score_thresh;
  a.min_size = min_size;
  a.img_w = img_w;
  a.img_h = img_h;
  a.back_w = back_w;
  a.back_h = back_h;
  det_top1_kernel<<<B, 1024, 0, stream>>>(head, reinterpret_cast<const float4*>(proposals), a, R, num_classes, det_box,
                                          det_score, det_label, det_row, det_roi, chan);
  return check_launch("det_top1_kernel");
}

extern "C" int eosvos_roi_match(const float* proposals, const int* count, const float* gt_boxes,
                                const long long* gt_labels, const int* gt_off, int B, int P, int max_gt, float iou_thresh,
                                float* all_boxes, long long* labels, long long* matched, int* counts,
                                eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(proposals && count && gt_boxes && gt_labels && gt_off && all_boxes && labels && matched && counts,
                 "roi_match: null pointer");
  const int rows = P + max_gt;
  dim3 grid((rows + 255) / 256, B);
  roi_match_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(proposals), count,
                                             reinterpret_cast<const float4*>(gt_boxes), gt_labels, gt_off, P, rows,
                                             iou_thresh, reinterpret_cast<float4*>(all_boxes), labels, matched, counts);
  return check_launch("roi_match_kernel");
}

// Anchor labelling for the RPN losses.  anchors fp32 [A_total][4] (one image's anchors); gt boxes concatenated over the
// images with gt_off [N + 1]; gt_best: scratch, one unsigned per ground-truth box, ZEROED by the caller; counts int32
// [N][2] = (#foreground, #background), ZEROED by the caller.
extern "C" int eosvos_rpn_anchor_match(const float* anchors, int num_anchors, const float* gt_boxes, const int* gt_off,
                                       int N, float fg_iou, float bg_iou, unsigned* gt_best, long long* labels,
                                       int* matched, int* counts, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(anchors && gt_boxes && gt_off && gt_best && labels && matched && counts, "rpn_anchor_match: null pointer");
  dim3 grid((num_anchors + 255) / 256, N);
  anchor_match_kernel<1><<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(anchors), num_anchors,
                                                   reinterpret_cast<const float4*>(gt_boxes), gt_off, fg_iou, bg_iou,
                                                   gt_best, labels, matched, counts);
  EOSVOS_TRY(check_launch("anchor_match_kernel<1>"));
  anchor_match_kernel<2><<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(anchors), num_anchors,
                                                   reinterpret_cast<const float4*>(gt_boxes), gt_off, fg_iou, bg_iou,
                                                   gt_best, labels, matched, counts);
  return check_launch("anchor_match_kernel<2>");
}

// mode 0: out[2] = (loss_objectness, loss_rpn_box_reg) over the M sampled anchors (entries < 0 are padding; M counts
// them out: pass the number of REAL samples as `m_norm`).  mode 1: scatter d loss / d head into the ZEROED fp32 buffers
// dys[l] ([N * hw[l]][16], same layout as heads[l]) scaled by the upstream gradients *g_obj / *g_box (device scalars).
extern "C" int eosvos_rpn_loss(const void* const* heads, void* const* dys, const int* hw, int num_levels, int A,
                               const long long* sampled, int num_sampled, const long long* labels, const int* matched,
                               const float* anchors, const float* gt_boxes, const int* gt_off, float beta, int mode,
                               float* out, const float* g_obj, const float* g_box, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(num_levels >= 1 && num_levels <= RPN_MAX_LEVELS, "rpn_loss: 1..8 pyramid levels");
  EOSVOS_REQUIRE(sampled && labels && matched && anchors && gt_boxes && gt_off, "rpn_loss: null pointer");
  EOSVOS_REQUIRE(num_sampled > 0, "rpn_loss: no sampled anchors");
  RpnLossLevels lv;
  lv.num_levels = num_levels;
  lv.A = A;
  int off = 0;
  for (int l = 0; l < num_levels; ++l) {
    lv.head[l] = reinterpret_cast<const float*>(heads[l]);
    lv.dy[l] = dys ? reinterpret_cast<float*>(dys[l]) : nullptr;
    lv.hw[l] = hw[l];
    lv.anchor_off[l] = off;
    off += hw[l] * A;
  }
  lv.anchors_per_image = off;
  const float4* an = reinterpret_cast<const float4*>(anchors);
  const float4* gt = reinterpret_cast<const float4*>(gt_boxes);
  if (mode == 0) {
    EOSVOS_REQUIRE(out, "rpn_loss: null output");
    rpn_loss_kernel<false><<<1, 1024, 0, stream>>>(lv, sampled, num_sampled, labels, matched, an, gt, gt_off, beta, out,
                                                   nullptr, nullptr);
    return check_launch("rpn_loss_kernel<fwd>");
  }
  EOSVOS_REQUIRE(dys && g_obj && g_box, "rpn_loss: backward needs gradient buffers and upstream gradients");
  rpn_loss_kernel<true><<<1, 1024, 0, stream>>>(lv, sampled, num_sampled, labels, matched, an, gt, gt_off, beta, nullptr,
                                                g_obj, g_box);
  return check_launch("rpn_loss_kernel<bwd>");
}

static int fill_sparse(RpnSparseLevels& lv, const void* const* heads, const void* const* ts, const void* const* fs,
                       void* const* dfs, const int* Hs, const int* Ws, int num_levels, int A, int C) {
  if (num_levels < 1 || num_levels > RPN_MAX_LEVELS) return -1;
  lv.num_levels = num_levels;
  lv.A = A;
  lv.C = C;
  int off = 0;
  for (int l = 0; l < num_levels; ++l) {
    lv.head[l] = heads ? reinterpret_cast<const float*>(heads[l]) : nullptr;
    lv.t[l] = ts ? reinterpret_cast<const act_t*>(ts[l]) : nullptr;
    lv.f[l] = fs ? reinterpret_cast<const act_t*>(fs[l]) : nullptr;
    lv.df[l] = dfs ? reinterpret_cast<act_t*>(dfs[l]) : nullptr;
    lv.H[l] = Hs[l];
    lv.W[l] = Ws[l];
    lv.anchor_off[l] = off;
    off += Hs[l] * Ws[l] * A;
  }
  lv.anchors_per_image = off;
  return 0;
}

// Stage 1 of the sparse RPN-head backward (see above).  Outputs: dt [M][C] 16-bit (x grad_scale), ev_pix int32 [M][4],
// X_g [M][9][C] 16-bit; accumulates into the ZEROED fp32 dw_cls [A][C], db_cls [A], dw_box [4A][C], db_box [4A].
extern "C" int eosvos_rpn_sparse_head(const void* const* heads, const void* const* ts, const void* const* fs,
                                      const int* Hs, const int* Ws, int num_levels, int A, int C,
                                      const long long* sampled, int M, const long long* labels, const int* matched,
                                      const float* anchors, const float* gt_boxes, const int* gt_off, float beta,
                                      const float* g_obj, const float* g_box, const float* w_cls, const float* w_box,
                                      float grad_scale, void* dt, int* ev_pix, void* xg, float* dw_cls, float* db_cls,
                                      float* dw_box, float* db_box, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RpnSparseLevels lv;
  EOSVOS_REQUIRE(fill_sparse(lv, heads, ts, fs, nullptr, Hs, Ws, num_levels, A, C) == 0, "rpn_sparse_head: 1..8 levels");
  EOSVOS_REQUIRE(C % 64 == 0 && C <= 1024, "rpn_sparse_head: channel count must be a multiple of 64, at most 1024");
  EOSVOS_REQUIRE(sampled && labels && matched && anchors && gt_boxes && gt_off && g_obj && g_box && w_cls && w_box && dt &&
                     ev_pix && xg && dw_cls && db_cls && dw_box && db_box,
                 "rpn_sparse_head: null pointer");
  EOSVOS_REQUIRE(M > 0, "rpn_sparse_head: no sampled anchors");
  rpn_sparse_head_kernel<<<M, C, 0, stream>>>(lv, sampled, M, labels, matched, reinterpret_cast<const float4*>(anchors),
                                              reinterpret_cast<const float4*>(gt_boxes), gt_off, beta, g_obj, g_box, w_cls,
                                              w_box, grad_scale, reinterpret_cast<act_t*>(dt),
                                              reinterpret_cast<int4*>(ev_pix), dw_cls, db_cls, dw_box, db_box);
  EOSVOS_TRY(check_launch("rpn_sparse_head_kernel"));
  rpn_sparse_gather_kernel<<<dim3(M, 9), C / 2, 0, stream>>>(lv, reinterpret_cast<const int4*>(ev_pix),
                                                             reinterpret_cast<act2_t*>(xg));
  return check_launch("rpn_sparse_gather_kernel");
}

// Stage 3: df[level][pixel + tap][ci] += G[e][ci * 9 + tap] into the ZEROED 16-bit level gradient maps.
extern "C" int eosvos_rpn_sparse_scatter(void* const* dfs, const int* Hs, const int* Ws, int num_levels, int C,
                                         const int* ev_pix, int M, const void* G, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RpnSparseLevels lv;
  EOSVOS_REQUIRE(fill_sparse(lv, nullptr, nullptr, nullptr, dfs, Hs, Ws, num_levels, 1, C) == 0, "rpn_sparse_scatter: levels");
  EOSVOS_REQUIRE(dfs && ev_pix && G && C <= 1024, "rpn_sparse_scatter: bad argument");
  rpn_sparse_scatter_kernel<<<dim3(M, 9), C, 0, stream>>>(lv, reinterpret_cast<const int4*>(ev_pix),
                                                          reinterpret_cast<const act_t*>(G));
  return check_launch("rpn_sparse_scatter_kernel");
}

extern "C" long long eosvos_roi_sample_scratch_bytes(int B, int rows) {
  const long long nchunks = (rows + SMP_CHUNK - 1) / SMP_CHUNK;
  return (long long)B * nchunks * 2 * 4 + (long long)B * ((rows + 31) / 32) * 4 + 64;
}

extern "C" int eosvos_roi_sample(const long long* labels, const long long* table, int B, int rows, int S, int Pmax,
                                 void* scratch, long long* inds, long long* pos_in, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(labels && table && inds && pos_in && scratch, "roi_sample: null pointer");
  EOSVOS_REQUIRE(rows >= 1, "roi_sample: no candidate rows");
  const int nchunks = (rows + SMP_CHUNK - 1) / SMP_CHUNK;
  int* chunk_cnt = reinterpret_cast<int*>(scratch);
  unsigned* selbits = reinterpret_cast<unsigned*>(chunk_cnt + (size_t)B * nchunks * 2);
  dim3 grid(nchunks, B);
  sample_count_kernel<<<grid, 256, 0, stream>>>(labels, rows, nchunks, chunk_cnt);
  EOSVOS_TRY(check_launch("sample_count_kernel"));
  sample_select_kernel<<<grid, 256, 0, stream>>>(labels, table, rows, nchunks, chunk_cnt, selbits);
  EOSVOS_TRY(check_launch("sample_select_kernel"));
  sample_compact_kernel<<<B, 1024, 0, stream>>>(labels, selbits, rows, S, Pmax, inds, pos_in);
  return check_launch("sample_compact_kernel");
}

extern "C" int eosvos_roi_encode(const float* all_boxes, const long long* labels, const long long* matched,
                                 const float* gt_boxes, const int* gt_off, const long long* inds, int B, int S,
                                 int rows_per_image, const float* coder_weights4, float* rois5, long long* out_labels,
                                 long long* out_matched, float* reg_targets, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(all_boxes && labels && matched && gt_boxes && gt_off && inds && rois5 && out_labels && out_matched &&
                     reg_targets,
                 "roi_encode: null pointer");
  dim3 grid((S + 255) / 256, B);
  roi_encode_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(all_boxes), labels, matched,
                                              reinterpret_cast<const float4*>(gt_boxes), gt_off, inds, S, rows_per_image,
                                              coder_weights4[0], coder_weights4[1], coder_weights4[2], coder_weights4[3],
                                              rois5, out_labels, out_matched, reinterpret_cast<float4*>(reg_targets));
  return check_launch("roi_encode_kernel");
}
