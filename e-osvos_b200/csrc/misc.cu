// Small HBM-bound helpers around the tensor-core kernels: layout/precision conversion, the
// GeneralizedRCNNTransform (normalise + bilinear resize + pad), stem im2col, max-pool, FPN
// sub-sampling, ReLU backward and bias gradients.  All NHWC, 128-bit accesses where shapes allow.
//   transform : tv transform.py:119-160, 25-84, 237-255 as invoked from src/networks/mask_rcnn.py:716
//   max-pool  : torchvision ResNet stem (3x3 / s2 / p1) and FPN LastLevelMaxPool (1x1 / s2)
#include "common.h"
#include "../../include/eosvos_b200.h"
#include "act.cuh"

namespace eosvos {

__device__ __forceinline__ void ld8f(const act_t* p, float (&f)[8]) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  const act2_t* h = reinterpret_cast<const act2_t*>(&v);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 t = act22float2(h[k]);
    f[2 * k] = t.x;
    f[2 * k + 1] = t.y;
  }
}
__device__ __forceinline__ void st8f(act_t* p, const float (&f)[8]) {
  uint4 v;
  act2_t* h = reinterpret_cast<act2_t*>(&v);
#pragma unroll
  for (int k = 0; k < 4; ++k) h[k] = floats2act2(f[2 * k], f[2 * k + 1]);
  *reinterpret_cast<uint4*>(p) = v;
}

// ---------------------------------------------------------------------------------------------
// generic 4-D strided gather with conversion: dst[i0][i1][i2][i3] (given dst strides) = src[...]
// ---------------------------------------------------------------------------------------------
struct Permute4 {
  long long dims[4];
  long long sstride[4];
  long long dstride[4];
};

template <typename TS, typename TD>
__global__ void __launch_bounds__(256) permute_cast_kernel(const TS* __restrict__ src, TD* __restrict__ dst, Permute4 pm,
                                                          long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long i3 = i % pm.dims[3];
  i /= pm.dims[3];
  const long long i2 = i % pm.dims[2];
  i /= pm.dims[2];
  const long long i1 = i % pm.dims[1];
  const long long i0 = i / pm.dims[1];
  const float v = (float)src[i0 * pm.sstride[0] + i1 * pm.sstride[1] + i2 * pm.sstride[2] + i3 * pm.sstride[3]];
  dst[i0 * pm.dstride[0] + i1 * pm.dstride[1] + i2 * pm.dstride[2] + i3 * pm.dstride[3]] = (TD)v;
}

// ---------------------------------------------------------------------------------------------
// transform: NCHW fp32 image in [0,1] -> normalised, bilinearly resized (align_corners=False,
// scale = in/out), zero-padded NHWC bf16 with Cs stored channels (3 used).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
transform_kernel(const float* __restrict__ img, act_t* __restrict__ out, int B, int h, int w, int oh, int ow,
                 int Hp, int Wp, int Cs, float m0, float m1, float m2, float is0, float is1, float is2) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * Hp * Wp;
  if (idx >= total) return;
  const int x = (int)(idx % Wp);
  const int y = (int)((idx / Wp) % Hp);
  const int b = (int)(idx / ((long long)Wp * Hp));
  float v[3] = {0.f, 0.f, 0.f};
  if (y < oh && x < ow) {
    const float sh = (float)h / (float)oh, sw = (float)w / (float)ow;
    float sy = sh * ((float)y + 0.5f) - 0.5f;
    float sx = sw * ((float)x + 0.5f) - 0.5f;
    if (sy < 0.f) sy = 0.f;
    if (sx < 0.f) sx = 0.f;
    int y0 = min((int)sy, h - 1), x0 = min((int)sx, w - 1);
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly = fminf(fmaxf(sy - (float)y0, 0.f), 1.f), lx = fminf(fmaxf(sx - (float)x0, 0.f), 1.f);
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float mean[3] = {m0, m1, m2}, istd[3] = {is0, is1, is2};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* pl = img + ((size_t)b * 3 + c) * h * w;
      const float v00 = (pl[(size_t)y0 * w + x0] - mean[c]) * istd[c], v01 = (pl[(size_t)y0 * w + x1] - mean[c]) * istd[c];
      const float v10 = (pl[(size_t)y1 * w + x0] - mean[c]) * istd[c], v11 = (pl[(size_t)y1 * w + x1] - mean[c]) * istd[c];
      v[c] = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
    }
  }
  act_t* o = out + (size_t)idx * Cs;
  for (int c = 0; c < Cs; ++c) o[c] = float2act(c < 3 ? v[c] : 0.f);
}

// nearest resize of id masks (legacy 'nearest': src = floor(dst * in/out)), fp32 -> uint8, padded
__global__ void __launch_bounds__(256)
mask_resize_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int G, int h, int w, int oh, int ow) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)G * oh * ow) return;
  const int x = (int)(idx % ow);
  const int y = (int)((idx / ow) % oh);
  const int g = (int)(idx / ((long long)ow * oh));
  const float sh = (float)h / (float)oh, sw = (float)w / (float)ow;
  const int sy = min((int)floorf((float)y * sh), h - 1), sx = min((int)floorf((float)x * sw), w - 1);
  dst[idx] = src[((size_t)g * h + sy) * w + sx];
}

// ---------------------------------------------------------------------------------------------
// stem im2col: x [N,H,W,Cs] bf16 (3 used) -> col [N*Ho*Wo][Kp] bf16, k = (kh*KW + kw)*3 + c
// ---------------------------------------------------------------------------------------------
// One CTA = IM2COL_ROWS consecutive output pixels: every (pixel, tap) item is ONE 16-byte load of the stored
// 8-channel input pixel (3 used), the Kp-wide rows are assembled in shared memory and leave as coalesced 16-byte
// stores (the per-element version issued 8 scalar gathers with a div/mod each per 16 output bytes: 400 us for the
// 297 MB column matrix of a 3 x 768 x 1344 batch; HBM floor ~55 us).
constexpr int IM2COL_ROWS = 32;
__global__ void __launch_bounds__(256)
im2col_stem_kernel(const act_t* __restrict__ x, act_t* __restrict__ col, int N, int H, int W, int Cs,
                   int Ho, int Wo, int KH, int KW, int stride, int pad, int Kp) {
  extern __shared__ __align__(16) unsigned char im2col_smem[];
  act_t* tile = reinterpret_cast<act_t*>(im2col_smem);          // [IM2COL_ROWS][Kp]
  const long long rows_total = (long long)N * Ho * Wo;
  const long long row0 = (long long)blockIdx.x * IM2COL_ROWS;
  const int taps = KH * KW;
  // zero the K padding (columns taps*3 .. Kp) once
  const int padw = Kp - taps * 3;
  for (int i = threadIdx.x; i < IM2COL_ROWS * padw; i += 256)
    tile[(i / padw) * Kp + taps * 3 + (i % padw)] = float2act(0.f);
  // (image, top-left input row / column) of each of the CTA's output pixels, computed once
  __shared__ int rn[IM2COL_ROWS], rh[IM2COL_ROWS], rw[IM2COL_ROWS];
  if (threadIdx.x < IM2COL_ROWS) {
    const long long row = row0 + threadIdx.x;
    if (row < rows_total) {
      const unsigned urow = (unsigned)row;           // N*Ho*Wo < 2^31 (checked on the host)
      const unsigned wo = urow % (unsigned)Wo, q = urow / (unsigned)Wo;
      rn[threadIdx.x] = (int)(q / (unsigned)Ho);
      rh[threadIdx.x] = (int)(q % (unsigned)Ho) * stride - pad;
      rw[threadIdx.x] = (int)wo * stride - pad;
    } else {
      rn[threadIdx.x] = -1;
      rh[threadIdx.x] = rw[threadIdx.x] = 0;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < IM2COL_ROWS * taps; i += 256) {
    const int r = i / taps, t = i - r * taps;
    uint4 v = make_uint4(0, 0, 0, 0);
    const int n = rn[r];
    if (n >= 0) {
      const int kh = t / KW, kw = t - kh * KW;
      const int hi = rh[r] + kh, wi = rw[r] + kw;
      if (hi >= 0 && hi < H && wi >= 0 && wi < W)
        v = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)n * H + hi) * W + wi) * Cs));
    }
    const act_t* h = reinterpret_cast<const act_t*>(&v);
    act_t* d = tile + r * Kp + t * 3;
    d[0] = h[0];
    d[1] = h[1];
    d[2] = h[2];
  }
  __syncthreads();
  const int kv = Kp >> 3;
  for (int i = threadIdx.x; i < IM2COL_ROWS * kv; i += 256) {
    const int r = i / kv, k8 = i - r * kv;
    if (row0 + r < rows_total)
      *reinterpret_cast<uint4*>(col + (size_t)(row0 + r) * Kp + (size_t)k8 * 8) =
          *reinterpret_cast<const uint4*>(tile + r * Kp + k8 * 8);
  }
}

// ---------------------------------------------------------------------------------------------
// max-pool k x k / stride s / pad p, NHWC 16-bit.  Forward also records, per output element, which window
// position (kh*k + kw, first maximum in scan order, as ATen) won; backward gathers from those indices.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(const act_t* __restrict__ x, act_t* __restrict__ y, uint8_t* __restrict__ arg, int N, int H, int W,
                   int C, int Ho, int Wo, int ksz, int stride, int pad) {
  const int cv = C >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * Ho * Wo * cv) return;
  const int c8 = (int)(idx % cv);
  long long t = idx / cv;
  const int wo = (int)(t % Wo);
  const int ho = (int)((t / Wo) % Ho);
  const int n = (int)(t / ((long long)Wo * Ho));
  float m[8];
  uint8_t am[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    m[k] = -3.0e38f;
    am[k] = 255;
  }
  for (int kh = 0; kh < ksz; ++kh) {
    const int hi = ho * stride - pad + kh;
    if (hi < 0 || hi >= H) continue;
    for (int kw = 0; kw < ksz; ++kw) {
      const int wi = wo * stride - pad + kw;
      if (wi < 0 || wi >= W) continue;
      float f[8];
      ld8f(x + (((size_t)n * H + hi) * W + wi) * C + (size_t)c8 * 8, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (f[k] > m[k]) {
          m[k] = f[k];
          am[k] = (uint8_t)(kh * ksz + kw);
        }
      }
    }
  }
  st8f(y + (size_t)idx * 8, m);
  if (arg) {
    uint2 pk;
    pk.x = am[0] | (am[1] << 8) | (am[2] << 16) | ((uint32_t)am[3] << 24);
    pk.y = am[4] | (am[5] << 8) | (am[6] << 16) | ((uint32_t)am[7] << 24);
    *reinterpret_cast<uint2*>(arg + (size_t)idx * 8) = pk;
  }
}

// dx[hi,wi] = sum over windows whose recorded arg-max is (hi,wi) of dy[window]
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const uint8_t* __restrict__ arg, const act_t* __restrict__ dy, act_t* __restrict__ dx, int N, int H,
                   int W, int C, int Ho, int Wo, int ksz, int stride, int pad) {
  const int cv = C >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * H * W * cv) return;
  const int c8 = (int)(idx % cv);
  long long t = idx / cv;
  const int wi = (int)(t % W);
  const int hi = (int)((t / W) % H);
  const int n = (int)(t / ((long long)W * H));
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  const int ho_lo = max(0, (hi + pad - ksz + stride) / stride), ho_hi = min(Ho - 1, (hi + pad) / stride);
  const int wo_lo = max(0, (wi + pad - ksz + stride) / stride), wo_hi = min(Wo - 1, (wi + pad) / stride);
  for (int ho = ho_lo; ho <= ho_hi; ++ho)
    for (int wo = wo_lo; wo <= wo_hi; ++wo) {
      const int pos = (hi - (ho * stride - pad)) * ksz + (wi - (wo * stride - pad));   // my position in that window
      const size_t o = (((size_t)n * Ho + ho) * Wo + wo) * C + (size_t)c8 * 8;
      const uint2 pk = *reinterpret_cast<const uint2*>(arg + o);
      float dv[8];
      ld8f(dy + o, dv);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t a = ((k < 4 ? pk.x : pk.y) >> (8 * (k & 3))) & 0xffu;
        acc[k] += (a == (uint32_t)pos) ? dv[k] : 0.f;
      }
    }
  st8f(dx + (size_t)idx * 8, acc);
}

// FPN LastLevelMaxPool (kernel 1, stride 2) forward and backward
__global__ void __launch_bounds__(256)
subsample2_kernel(const act_t* __restrict__ x, act_t* __restrict__ y, int N, int H, int W, int C, int Ho,
                  int Wo, int backward) {
  const int cv = C >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (backward) {  // x = dy [N,Ho,Wo,C] -> y = dx [N,H,W,C]
    if (idx >= (long long)N * H * W * cv) return;
    const int c8 = (int)(idx % cv);
    long long t = idx / cv;
    const int wi = (int)(t % W), hi = (int)((t / W) % H), n = (int)(t / ((long long)W * H));
    uint4 v = make_uint4(0, 0, 0, 0);
    if (!(hi & 1) && !(wi & 1) && (hi >> 1) < Ho && (wi >> 1) < Wo)
      v = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)n * Ho + (hi >> 1)) * Wo + (wi >> 1)) * C + (size_t)c8 * 8));
    *reinterpret_cast<uint4*>(y + (size_t)idx * 8) = v;
  } else {
    if (idx >= (long long)N * Ho * Wo * cv) return;
    const int c8 = (int)(idx % cv);
    long long t = idx / cv;
    const int wo = (int)(t % Wo), ho = (int)((t / Wo) % Ho), n = (int)(t / ((long long)Wo * Ho));
    *reinterpret_cast<uint4*>(y + (size_t)idx * 8) =
        __ldg(reinterpret_cast<const uint4*>(x + (((size_t)n * H + 2 * ho) * W + 2 * wo) * C + (size_t)c8 * 8));
  }
}

// backward of nearest 2x upsampling: dcoarse[h,w] = sum of the 2x2 fine gradients
__global__ void __launch_bounds__(256)
sum2x2_kernel(const act_t* __restrict__ dfine, act_t* __restrict__ dcoarse, int N, int Hc, int Wc, int C) {
  const int cv = C >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * Hc * Wc * cv) return;
  const int c8 = (int)(idx % cv);
  long long t = idx / cv;
  const int w = (int)(t % Wc), h = (int)((t / Wc) % Hc), n = (int)(t / ((long long)Wc * Hc));
  const int Hf = 2 * Hc, Wf = 2 * Wc;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      float f[8];
      ld8f(dfine + (((size_t)n * Hf + 2 * h + dy) * Wf + 2 * w + dx) * C + (size_t)c8 * 8, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += f[k];
    }
  st8f(dcoarse + (size_t)idx * 8, acc);
}

// dy_eff = dy * (y > 0)   (fused conv+bias+ReLU backward entry)
__global__ void __launch_bounds__(256)
relu_bwd_kernel(const act_t* __restrict__ dy, const act_t* __restrict__ y, act_t* __restrict__ out,
                long long nvec) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nvec) return;
  float d[8], yy[8];
  ld8f(dy + i * 8, d);
  ld8f(y + i * 8, yy);
#pragma unroll
  for (int k = 0; k < 8; ++k) d[k] = yy[k] > 0.f ? d[k] : 0.f;
  st8f(out + i * 8, d);
}

// bias gradient: out[c] += sum_rows dy[row][c]; dy bf16 [M][C]
__global__ void __launch_bounds__(256)
colsum_kernel(const act_t* __restrict__ dy, float* __restrict__ out, long long M, int C, int rows_per_block,
              float alpha) {
  extern __shared__ float sm[];  // [C]
  const int cv = C >> 3;
  const int my_cv = threadIdx.x % cv;
  const int rpp = 256 / cv;
  for (int i = threadIdx.x; i < C; i += 256) sm[i] = 0.f;
  __syncthreads();
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(M, r0 + rows_per_block);
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (long long r = r0 + threadIdx.x / cv; r < r1; r += rpp) {
    float f[8];
    ld8f(dy + r * C + (size_t)my_cv * 8, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += f[k];
  }
  if (threadIdx.x / cv < rpp) {
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&sm[my_cv * 8 + k], acc[k]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) atomicAdd(&out[i], sm[i] * alpha);
}

static inline unsigned blocks_for(long long total) { return (unsigned)((total + 255) / 256); }

}  // namespace eosvos

using namespace eosvos;

// dtype codes: 0 = fp32, 1 = activation type (act_t)
extern "C" int eosvos_permute_cast(const void* src, void* dst, const long long* dims, const long long* sstride,
                                   const long long* dstride, int src_dtype, int dst_dtype, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(src && dst && dims && sstride && dstride, "permute_cast: null pointer");
  Permute4 pm;
  long long total = 1;
  for (int i = 0; i < 4; ++i) {
    pm.dims[i] = dims[i];
    pm.sstride[i] = sstride[i];
    pm.dstride[i] = dstride[i];
    total *= dims[i];
  }
  if (total == 0) return 0;
  const unsigned nb = blocks_for(total);
  if (src_dtype == 0 && dst_dtype == 1)
    permute_cast_kernel<float, act_t><<<nb, 256, 0, stream>>>(reinterpret_cast<const float*>(src),
                                                                     reinterpret_cast<act_t*>(dst), pm, total);
  else if (src_dtype == 1 && dst_dtype == 0)
    permute_cast_kernel<act_t, float><<<nb, 256, 0, stream>>>(reinterpret_cast<const act_t*>(src),
                                                                     reinterpret_cast<float*>(dst), pm, total);
  else if (src_dtype == 0 && dst_dtype == 0)
    permute_cast_kernel<float, float><<<nb, 256, 0, stream>>>(reinterpret_cast<const float*>(src),
                                                             reinterpret_cast<float*>(dst), pm, total);
  else if (src_dtype == 1 && dst_dtype == 1)
    permute_cast_kernel<act_t, act_t><<<nb, 256, 0, stream>>>(
        reinterpret_cast<const act_t*>(src), reinterpret_cast<act_t*>(dst), pm, total);
  else
    return set_error(EOSVOS_ERR_ARG, "permute_cast: unknown dtype code");
  return check_launch("permute_cast_kernel");
}

extern "C" int eosvos_transform(const float* img, void* out, int B, int h, int w, int oh, int ow, int Hp, int Wp, int Cs,
                                const float* mean3, const float* std3, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(img && out && mean3 && std3, "transform: null pointer");
  EOSVOS_REQUIRE(Cs >= 3 && oh <= Hp && ow <= Wp, "transform: bad geometry");
  const long long total = (long long)B * Hp * Wp;
  transform_kernel<<<blocks_for(total), 256, 0, stream>>>(img, reinterpret_cast<act_t*>(out), B, h, w, oh, ow, Hp,
                                                        Wp, Cs, mean3[0], mean3[1], mean3[2], 1.f / std3[0],
                                                        1.f / std3[1], 1.f / std3[2]);
  return check_launch("transform_kernel");
}

extern "C" int eosvos_mask_resize_nearest(const uint8_t* src, uint8_t* dst, int G, int h, int w, int oh, int ow,
                                          eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (G == 0) return 0;
  EOSVOS_REQUIRE(src && dst, "mask_resize_nearest: null pointer");
  mask_resize_kernel<<<blocks_for((long long)G * oh * ow), 256, 0, stream>>>(src, dst, G, h, w, oh, ow);
  return check_launch("mask_resize_kernel");
}

extern "C" int eosvos_im2col_stem(const void* x, void* col, int N, int H, int W, int Cs, int KH, int KW, int stride,
                                  int pad, int Kp, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(x && col, "im2col_stem: null pointer");
  EOSVOS_REQUIRE(Kp % 64 == 0 && Kp >= KH * KW * 3, "im2col_stem: Kp must be a multiple of 64 covering KH*KW*3");
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  EOSVOS_REQUIRE(Cs == 8, "im2col_stem: the input must store 8 channels per pixel (16-byte pixel loads)");
  EOSVOS_REQUIRE(Kp <= 512, "im2col_stem: Kp too large for the shared-memory row tile");
  const long long rows = (long long)N * Ho * Wo;
  EOSVOS_REQUIRE(rows < (1LL << 31), "im2col_stem: too many output pixels");
  const unsigned nb = (unsigned)((rows + eosvos::IM2COL_ROWS - 1) / eosvos::IM2COL_ROWS);
  im2col_stem_kernel<<<nb, 256, (size_t)eosvos::IM2COL_ROWS * Kp * sizeof(act_t), stream>>>(
      reinterpret_cast<const act_t*>(x), reinterpret_cast<act_t*>(col), N, H, W, Cs, Ho, Wo, KH, KW, stride, pad, Kp);
  return check_launch("im2col_stem_kernel");
}

extern "C" int eosvos_maxpool_fwd(const void* x, void* y, uint8_t* argmax, int N, int H, int W, int C, int ksz,
                                  int stride, int pad, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(x && y && C % 8 == 0, "maxpool_fwd: bad arguments");
  const int Ho = (H + 2 * pad - ksz) / stride + 1, Wo = (W + 2 * pad - ksz) / stride + 1;
  maxpool_fwd_kernel<<<blocks_for((long long)N * Ho * Wo * (C >> 3)), 256, 0, stream>>>(
      reinterpret_cast<const act_t*>(x), reinterpret_cast<act_t*>(y), argmax, N, H, W, C, Ho, Wo, ksz, stride, pad);
  return check_launch("maxpool_fwd_kernel");
}

extern "C" int eosvos_maxpool_bwd(const uint8_t* argmax, const void* dy, void* dx, int N, int H, int W, int C, int ksz,
                                  int stride, int pad, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(argmax && dy && dx && C % 8 == 0, "maxpool_bwd: bad arguments");
  const int Ho = (H + 2 * pad - ksz) / stride + 1, Wo = (W + 2 * pad - ksz) / stride + 1;
  maxpool_bwd_kernel<<<blocks_for((long long)N * H * W * (C >> 3)), 256, 0, stream>>>(
      argmax, reinterpret_cast<const act_t*>(dy), reinterpret_cast<act_t*>(dx), N, H, W, C, Ho, Wo, ksz, stride, pad);
  return check_launch("maxpool_bwd_kernel");
}

extern "C" int eosvos_subsample2(const void* x, void* y, int N, int H, int W, int C, int backward,
                                 eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(x && y && C % 8 == 0, "subsample2: bad arguments");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long total = backward ? (long long)N * H * W * (C >> 3) : (long long)N * Ho * Wo * (C >> 3);
  subsample2_kernel<<<blocks_for(total), 256, 0, stream>>>(reinterpret_cast<const act_t*>(x),
                                                         reinterpret_cast<act_t*>(y), N, H, W, C, Ho, Wo,
                                                         backward);
  return check_launch("subsample2_kernel");
}

extern "C" int eosvos_sum2x2(const void* dfine, void* dcoarse, int N, int Hc, int Wc, int C, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(dfine && dcoarse && C % 8 == 0, "sum2x2: bad arguments");
  sum2x2_kernel<<<blocks_for((long long)N * Hc * Wc * (C >> 3)), 256, 0, stream>>>(
      reinterpret_cast<const act_t*>(dfine), reinterpret_cast<act_t*>(dcoarse), N, Hc, Wc, C);
  return check_launch("sum2x2_kernel");
}

extern "C" int eosvos_relu_bwd(const void* dy, const void* y, void* out, long long numel, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (numel == 0) return 0;
  EOSVOS_REQUIRE(dy && y && out && numel % 8 == 0, "relu_bwd: numel must be a multiple of 8");
  relu_bwd_kernel<<<blocks_for(numel / 8), 256, 0, stream>>>(reinterpret_cast<const act_t*>(dy),
                                                           reinterpret_cast<const act_t*>(y),
                                                           reinterpret_cast<act_t*>(out), numel / 8);
  return check_launch("relu_bwd_kernel");
}

// out[c] (fp32, ACCUMULATED: caller zeroes) += column sums of dy [M][C] bf16
extern "C" int eosvos_colsum(const void* dy, float* out, long long M, int C, float alpha, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (M == 0) return 0;
  EOSVOS_REQUIRE(dy && out, "colsum: null pointer");
  EOSVOS_REQUIRE(C % 8 == 0 && C <= 2048 && 256 % (C >> 3) == 0, "colsum: C/8 must divide 256");
  const int rpp = 256 / (C >> 3);
  long long want_blocks = 4LL * num_sms();
  long long rows_per_block = (M + want_blocks - 1) / want_blocks;
  rows_per_block = ((rows_per_block + rpp - 1) / rpp) * rpp;
  const unsigned nb = (unsigned)((M + rows_per_block - 1) / rows_per_block);
  colsum_kernel<<<nb, 256, C * sizeof(float), stream>>>(reinterpret_cast<const act_t*>(dy), out, M, C,
                                                        (int)rows_per_block, alpha);
  return check_launch("colsum_kernel");
}

// ---------------------------------------------------------------------------------------------
// multi-tensor permute + cast (fp32 -> act_t): all tensor-core operand layouts of one iteration in ONE launch.
// table: int64 [T][14] = (src, dst, dims[4], sstride[4], dstride[4]); chunks: int32 [n][2] = (tensor, chunk)
// ---------------------------------------------------------------------------------------------
namespace eosvos {
constexpr int PM_CHUNK = 8192;
__global__ void __launch_bounds__(256)
permute_cast_multi_kernel(const long long* __restrict__ table, const int* __restrict__ chunks) {
  const int t = chunks[blockIdx.x * 2];
  const long long start = (long long)chunks[blockIdx.x * 2 + 1] * PM_CHUNK;
  const long long* e = table + (size_t)t * 14;
  const float* __restrict__ src = reinterpret_cast<const float*>(e[0]);
  act_t* __restrict__ dst = reinterpret_cast<act_t*>(e[1]);
  const long long d1 = e[3], d2 = e[4], d3 = e[5];
  const long long total = e[2] * d1 * d2 * d3;
  long long end = start + PM_CHUNK;
  if (end > total) end = total;
  for (long long i = start + threadIdx.x; i < end; i += 256) {
    long long r = i;
    const long long i3 = r % d3;
    r /= d3;
    const long long i2 = r % d2;
    r /= d2;
    const long long i1 = r % d1;
    const long long i0 = r / d1;
    const float v = src[i0 * e[6] + i1 * e[7] + i2 * e[8] + i3 * e[9]];
    dst[i0 * e[10] + i1 * e[11] + i2 * e[12] + i3 * e[13]] = float2act(v);
  }
}
}  // namespace eosvos

extern "C" int eosvos_permute_cast_multi_chunk_elems(void) { return eosvos::PM_CHUNK; }

extern "C" int eosvos_permute_cast_multi(const long long* table_dev, const int* chunks_dev, int num_chunks,
                                         eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (num_chunks == 0) return 0;
  EOSVOS_REQUIRE(table_dev && chunks_dev, "permute_cast_multi: null table");
  eosvos::permute_cast_multi_kernel<<<num_chunks, 256, 0, stream>>>(table_dev, chunks_dev);
  return eosvos::check_launch("permute_cast_multi_kernel");
}

// ---------------------------------------------------------------------------------------------
// Tensor-core operand layouts of MANY parameter tensors in one launch, tiled through shared memory so that both the
// fp32 reads and the 16-bit writes are coalesced (the element-wise kernels above read with a stride of KH*KW or
// Cin*KH*KW floats).  Every layout the path needs is a permutation of a contiguous fp32 [X][Y][Z] tensor:
//     dst[x * dx + y * dy + z * dz] = (act) src[(x * Y + y) * Z + z]          with dx == 1 or dy == 1
//   conv fprop operand  [O][I][T] -> [O][T][I]   (dy = 1)        conv dgrad operand [O][I][T] -> [I][T][O]   (dx = 1)
//   fc6 (NHWC pooling)  [O][C][S] -> [O][S][C]   (dy = 1)        its transpose      [O][C][S] -> [S][C][O]   (dx = 1)
// table: int64 [n][10] = (src, dst, X, Y, Z, dx, dy, dz, TX, TY); tiles: int32 [m][2] = (tensor, tile index).
// One CTA = one TX x TY x Z tile (<= WP_TILE floats).
// ---------------------------------------------------------------------------------------------
namespace eosvos {
constexpr int WP_TILE = 8192;
__device__ __forceinline__ uint32_t pack_act2_f(float a, float b) {
  act2_t h = floats2act2(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__global__ void __launch_bounds__(256)
weight_prep_kernel(const long long* __restrict__ table, const int* __restrict__ tiles) {
  extern __shared__ float wp_smem[];
  const int t = tiles[blockIdx.x * 2];
  const int tile = tiles[blockIdx.x * 2 + 1];
  const long long* e = table + (size_t)t * 10;
  const float* __restrict__ src = reinterpret_cast<const float*>(e[0]);
  act_t* __restrict__ dst = reinterpret_cast<act_t*>(e[1]);
  const int X = (int)e[2], Y = (int)e[3], Z = (int)e[4];
  const long long dx = e[5], dy = e[6], dz = e[7];
  const int TX = (int)e[8], TY = (int)e[9];
  const int tiles_y = (Y + TY - 1) / TY;
  const int x0 = (tile / tiles_y) * TX, y0 = (tile % tiles_y) * TY;
  const int nx = min(TX, X - x0), ny = min(TY, Y - y0);
  const int RL = ny * Z;                 // contiguous floats per x row of this tile
  const int RLP = (TY * Z) | 1;          // odd row pitch: conflict-free column reads
  const bool src4 = (RL & 3) == 0 && ((((size_t)Y * Z) | ((size_t)y0 * Z)) & 3) == 0 && ((e[0] & 15) == 0);
  if (src4) {   // 16-byte global loads (the shared row pitch is odd, so the tile is filled with scalar stores)
    const int RL4 = RL >> 2;
    for (int i = threadIdx.x; i < nx * RL4; i += 256) {
      const int x = i / RL4, r = (i - x * RL4) << 2;
      const float4 v = __ldg(reinterpret_cast<const float4*>(src + ((size_t)(x0 + x) * Y + y0) * Z + r));
      float* d = wp_smem + x * RLP + r;
      d[0] = v.x;
      d[1] = v.y;
      d[2] = v.z;
      d[3] = v.w;
    }
  } else {
    for (int i = threadIdx.x; i < nx * RL; i += 256) {
      const int x = i / RL, r = i - x * RL;
      wp_smem[x * RLP + r] = __ldg(src + ((size_t)(x0 + x) * Y + y0) * Z + r);
    }
  }
  __syncthreads();
  const bool dst16 = (e[1] & 15) == 0;     // + per-mode: the non-unit strides and the tile origin are multiples of 8
  if (dy == 1 && dst16 && (ny & 7) == 0 && (y0 & 7) == 0 && (dx & 7) == 0 && (dz & 7) == 0) {
    // 8 consecutive y per thread -> one 16-byte store
    const int ny8 = ny >> 3;
    for (int i = threadIdx.x; i < nx * Z * ny8; i += 256) {
      const int yi = (i % ny8) << 3;
      const int rem = i / ny8;
      const int z = rem % Z, x = rem / Z;
      const float* s = wp_smem + x * RLP + yi * Z + z;
      uint4 w;
      w.x = pack_act2_f(s[0], s[Z]);
      w.y = pack_act2_f(s[2 * Z], s[3 * Z]);
      w.z = pack_act2_f(s[4 * Z], s[5 * Z]);
      w.w = pack_act2_f(s[6 * Z], s[7 * Z]);
      *reinterpret_cast<uint4*>(dst + (long long)(x0 + x) * dx + (y0 + yi) + (long long)z * dz) = w;
    }
  } else if (dx == 1 && dst16 && (nx & 7) == 0 && (x0 & 7) == 0 && (dy & 7) == 0 && (dz & 7) == 0) {
    const int nx8 = nx >> 3;
    for (int i = threadIdx.x; i < RL * nx8; i += 256) {
      const int xi = (i % nx8) << 3;
      const int r = i / nx8;
      const int y = r / Z, z = r - y * Z;
      const float* s = wp_smem + xi * RLP + r;
      uint4 w;
      w.x = pack_act2_f(s[0], s[RLP]);
      w.y = pack_act2_f(s[2 * RLP], s[3 * RLP]);
      w.z = pack_act2_f(s[4 * RLP], s[5 * RLP]);
      w.w = pack_act2_f(s[6 * RLP], s[7 * RLP]);
      *reinterpret_cast<uint4*>(dst + (long long)(x0 + xi) + (long long)(y0 + y) * dy + (long long)z * dz) = w;
    }
  } else if (dy == 1) {
    for (int i = threadIdx.x; i < nx * RL; i += 256) {
      const int yi = i % ny;
      const int rem = i / ny;
      const int z = rem % Z, x = rem / Z;
      dst[(long long)(x0 + x) * dx + (y0 + yi) + (long long)z * dz] = float2act(wp_smem[x * RLP + yi * Z + z]);
    }
  } else {
    for (int i = threadIdx.x; i < nx * RL; i += 256) {
      const int xi = i % nx;
      const int r = i / nx;
      const int y = r / Z, z = r - y * Z;
      dst[(long long)(x0 + xi) + (long long)(y0 + y) * dy + (long long)z * dz] = float2act(wp_smem[xi * RLP + r]);
    }
  }
}
}  // namespace eosvos

extern "C" int eosvos_weight_prep_tile_elems(void) { return eosvos::WP_TILE; }

extern "C" int eosvos_weight_prep_multi(const long long* table_dev, const int* tiles_dev, int num_tiles,
                                        eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (num_tiles == 0) return 0;
  EOSVOS_REQUIRE(table_dev && tiles_dev, "weight_prep_multi: null table");
  constexpr int SMEM = (eosvos::WP_TILE + 64) * (int)sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e2 = cudaFuncSetAttribute(eosvos::weight_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e2 != cudaSuccess) return eosvos::set_cuda_error(e2, "cudaFuncSetAttribute(weight_prep)");
    attr_done = true;
  }
  eosvos::weight_prep_kernel<<<num_tiles, 256, SMEM, stream>>>(table_dev, tiles_dev);
  return eosvos::check_launch("weight_prep_kernel");
}

// ---------------------------------------------------------------------------------------------
// First-frame augmentation on the device: horizontal flip + rotate/scale about the centre with bicubic
// (a = -0.75, constant-0 border) sampling -- the image half of the reference's RandomHorizontalFlip +
// RandomScaleNRotate (src/data/custom_transforms.py:40-51,188-211; cv2.warpAffine INTER_CUBIC).  The label half
// (nearest) and the random draws / rejection loop stay on the host (util/augment.py).  cv2 quantises source
// coordinates to 1/32 px; this kernel uses exact float coordinates.
// src [3][H][W] fp32, minv [B][6] (dst -> src affine), flip [B]; dst [B][3][H][W] fp32.
// ---------------------------------------------------------------------------------------------
namespace eosvos {
__device__ __forceinline__ void cubic_w(float t, float (&w)[4]) {
  const float A = -0.75f;
  w[0] = ((A * (t + 1.f) - 5.f * A) * (t + 1.f) + 8.f * A) * (t + 1.f) - 4.f * A;
  w[1] = ((A + 2.f) * t - (A + 3.f)) * t * t + 1.f;
  w[2] = ((A + 2.f) * (1.f - t) - (A + 3.f)) * (1.f - t) * (1.f - t) + 1.f;
  w[3] = 1.f - w[0] - w[1] - w[2];
}

__global__ void __launch_bounds__(256)
affine_warp_cubic_kernel(const float* __restrict__ src, const float* __restrict__ minv, const int* __restrict__ flip,
                         float* __restrict__ dst, int B, int H, int W) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H * W) return;
  const int x = (int)(idx % W);
  const int y = (int)((idx / W) % H);
  const int b = (int)(idx / ((long long)W * H));
  const float* m = minv + b * 6;
  const float sx = m[0] * x + m[1] * y + m[2];
  const float sy = m[3] * x + m[4] * y + m[5];
  const int ix = (int)floorf(sx), iy = (int)floorf(sy);
  float wx[4], wy[4];
  cubic_w(sx - (float)ix, wx);
  cubic_w(sy - (float)iy, wy);
  const bool fl = flip[b] != 0;
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int yy = iy - 1 + j;
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int xx = ix - 1 + i;
      if (xx < 0 || xx >= W) continue;
      if (fl) xx = W - 1 - xx;
      const float w = wy[j] * wx[i];
      const size_t o = (size_t)yy * W + xx;
      acc[0] += w * src[o];
      acc[1] += w * src[(size_t)H * W + o];
      acc[2] += w * src[2 * (size_t)H * W + o];
    }
  }
  const size_t po = (size_t)y * W + x;
#pragma unroll
  for (int c = 0; c < 3; ++c) dst[((size_t)b * 3 + c) * H * W + po] = acc[c];
}

// cv2.warpAffine(label, M, flags=INTER_NEAREST) (+ an optional horizontal flip of the source first), bit for bit:
// OpenCV walks the destination with 10-bit fixed-point source coordinates -- per column rint(m0*x*1024) and
// rint(m3*x*1024), per row rint((m1*y + m2)*1024) + 512 and rint((m4*y + m5)*1024) + 512, summed and shifted right by
// 10, saturated to int16 -- and copies the source pixel or writes 0 outside (imgwarp.cpp WarpAffineInvoker +
// remapNearest, BORDER_CONSTANT).  minv is the INVERTED matrix in double, as OpenCV forms it.  Explicit
// round-to-nearest multiplies / adds: no FMA contraction, the host code has none either.
__global__ void __launch_bounds__(256)
label_warp_nearest_kernel(const float* __restrict__ src, const double* __restrict__ minv, const int* __restrict__ flip,
                          float* __restrict__ dst, int B, int H, int W) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H * W) return;
  const int x = (int)(idx % W);
  const int y = (int)((idx / W) % H);
  const int b = (int)(idx / ((long long)W * H));
  const double* m = minv + b * 6;
  const double xd = (double)x, yd = (double)y;
  const int ad = __double2int_rn(__dmul_rn(__dmul_rn(m[0], xd), 1024.0));
  const int bd = __double2int_rn(__dmul_rn(__dmul_rn(m[3], xd), 1024.0));
  const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[1], yd), m[2]), 1024.0)) + 512;
  const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[4], yd), m[5]), 1024.0)) + 512;
  int X = (X0 + ad) >> 10, Y = (Y0 + bd) >> 10;
  X = min(max(X, -32768), 32767);
  Y = min(max(Y, -32768), 32767);
  float v = 0.f;
  if (X >= 0 && X < W && Y >= 0 && Y < H) v = src[(size_t)Y * W + (flip[b] ? W - 1 - X : X)];
  dst[idx] = v;
}
}  // namespace eosvos

extern "C" int eosvos_label_warp_nearest(const float* src, const double* minv, const int* flip, float* dst, int B, int H,
                                         int W, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(src && minv && flip && dst && B > 0 && H > 0 && W > 0, "label_warp_nearest: bad arguments");
  EOSVOS_REQUIRE(H <= 32767 && W <= 32767, "label_warp_nearest: image larger than OpenCV's int16 coordinate range");
  const long long total = (long long)B * H * W;
  eosvos::label_warp_nearest_kernel<<<eosvos::blocks_for(total), 256, 0, stream>>>(src, minv, flip, dst, B, H, W);
  return eosvos::check_launch("label_warp_nearest_kernel");
}

extern "C" int eosvos_affine_warp_cubic(const float* src, const float* minv, const int* flip, float* dst, int B, int H,
                                        int W, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_REQUIRE(src && minv && flip && dst && B > 0, "affine_warp_cubic: bad arguments");
  const long long total = (long long)B * H * W;
  eosvos::affine_warp_cubic_kernel<<<eosvos::blocks_for(total), 256, 0, stream>>>(src, minv, flip, dst, B, H, W);
  return eosvos::check_launch("affine_warp_cubic_kernel");
}
