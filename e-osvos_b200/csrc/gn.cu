// GroupNorm(32, C) (+ residual add) (+ ReLU) forward / backward on NHWC bf16 activations.
// Replaces the 53 ATen native_group_norm calls the reference makes after its BN->GN swap
// (reference: src/networks/mask_rcnn.py:523-534).
//
// HBM-bound.  Forward = statistics (sum, sumsq per (n, group); skipped when the producing conv's
// epilogue already emitted them) + one normalise/affine/add/ReLU pass (1 read [+1 residual read]
// + 1 write).  Backward = one reduction pass + one apply pass.
#include "common.h"
#include "../../include/eosvos_b200.h"
#include "act.cuh"

namespace eosvos {

constexpr int GN_GROUPS = 32;
constexpr int GN_THREADS = 256;

__device__ __forceinline__ void load8(const act_t* p, float (&f)[8]) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  const act2_t* h = reinterpret_cast<const act2_t*>(&v);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 t = act22float2(h[k]);
    f[2 * k] = t.x;
    f[2 * k + 1] = t.y;
  }
}
__device__ __forceinline__ void store8(act_t* p, const float (&f)[8]) {
  uint4 v;
  act2_t* h = reinterpret_cast<act2_t*>(&v);
#pragma unroll
  for (int k = 0; k < 4; ++k) h[k] = floats2act2(f[2 * k], f[2 * k + 1]);
  *reinterpret_cast<uint4*>(p) = v;
}

// ---------------------------------------------------------------------------------------------
// statistics: sums[n][g] = (sum x, sum x^2).  grid = (chunks, N); C/8 must divide GN_THREADS.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GN_THREADS) gn_stats_kernel(const act_t* __restrict__ x, float* __restrict__ sums,
                                                             int HW, int C, int chunk_pixels) {
  __shared__ float sm[GN_GROUPS * 2];
  const int n = blockIdx.y;
  const int cv = C >> 3;                       // channel vectors per pixel
  const int my_cv = threadIdx.x % cv;
  const int pix_per_pass = GN_THREADS / cv;
  const int p0 = blockIdx.x * chunk_pixels;
  const int p1 = min(HW, p0 + chunk_pixels);
  if (threadIdx.x < GN_GROUPS * 2) sm[threadIdx.x] = 0.f;
  __syncthreads();
  float s1[8], s2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s1[k] = s2[k] = 0.f;
  const act_t* base = x + (size_t)n * HW * C + (size_t)my_cv * 8;
  for (int p = p0 + threadIdx.x / cv; p < p1; p += pix_per_pass) {
    float f[8];
    load8(base + (size_t)p * C, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s1[k] += f[k];
      s2[k] += f[k] * f[k];
    }
  }
  const int cpg = C / GN_GROUPS;
  if (cpg >= 8) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      a += s1[k];
      b += s2[k];
    }
    const int g = (my_cv * 8) / cpg;
    atomicAdd(&sm[g * 2], a);
    atomicAdd(&sm[g * 2 + 1], b);
  } else {
    for (int k0 = 0; k0 < 8; k0 += cpg) {
      float a = 0.f, b = 0.f;
      for (int k = k0; k < k0 + cpg; ++k) {
        a += s1[k];
        b += s2[k];
      }
      const int g = (my_cv * 8 + k0) / cpg;
      atomicAdd(&sm[g * 2], a);
      atomicAdd(&sm[g * 2 + 1], b);
    }
  }
  __syncthreads();
  if (threadIdx.x < GN_GROUPS * 2) atomicAdd(&sums[(size_t)n * GN_GROUPS * 2 + threadIdx.x], sm[threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------
// apply: y = relu?( (x - mean) * rstd * gamma + beta  (+ res) )
//
// All three streaming kernels below keep GN_UNROLL independent 16-byte loads per operand in flight per thread
// (HBM latency x bandwidth needs ~45 KB in flight per SM) and fold the per-channel constants into one multiply-add
// so the register budget allows 4 CTAs of 256 threads per SM.  The apply kernels walk the tensor in REVERSE block
// order: the producer (conv epilogue / reduce pass) touched the high addresses last, so those are still in L2.
// ---------------------------------------------------------------------------------------------
constexpr int GN_UNROLL = 4;

__device__ __forceinline__ uint4 ldg16(const act_t* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const act2_t* h = reinterpret_cast<const act2_t*>(&v);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 t = act22float2(h[k]);
    f[2 * k] = t.x;
    f[2 * k + 1] = t.y;
  }
}

template <bool RES>
__global__ void __launch_bounds__(GN_THREADS, 3)
gn_apply_kernel(const act_t* __restrict__ x, const float* __restrict__ sums, const float* __restrict__ gamma,
                const float* __restrict__ beta, const act_t* __restrict__ res, act_t* __restrict__ y,
                int HW, int C, int chunk_pixels, float eps, int relu) {
  const int n = gridDim.y - 1 - blockIdx.y;
  const int bx = gridDim.x - 1 - blockIdx.x;
  const int cv = C >> 3;
  const int my_cv = threadIdx.x % cv;
  const int pix_per_pass = GN_THREADS / cv;
  const int cpg = C / GN_GROUPS;
  const float inv_m = 1.f / ((float)cpg * (float)HW);
  float a[8], b[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = my_cv * 8 + k;
    const int g = c / cpg;
    const float mean = sums[((size_t)n * GN_GROUPS + g) * 2] * inv_m;
    const float var = fmaxf(sums[((size_t)n * GN_GROUPS + g) * 2 + 1] * inv_m - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    a[k] = rstd * gamma[c];
    b[k] = beta[c] - mean * a[k];
  }
  const int p0 = bx * chunk_pixels;
  const int p1 = min(HW, p0 + chunk_pixels);
  const size_t base = (size_t)n * HW * C + (size_t)my_cv * 8;
  const size_t step = (size_t)pix_per_pass * C;
  auto one = [&](const uint4& xv, const uint4& rv, size_t o) {
    float f[8];
    unpack8(xv, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = f[k] * a[k] + b[k];
    if (RES) {
      float r[8];
      unpack8(rv, r);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] += r[k];
    }
    if (relu) {
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.f);
    }
    store8(y + o, f);
  };
  int p = p0 + threadIdx.x / cv;
  size_t o = base + (size_t)p * C;
  for (; p + (GN_UNROLL - 1) * pix_per_pass < p1; p += GN_UNROLL * pix_per_pass, o += GN_UNROLL * step) {
    uint4 xv[GN_UNROLL], rv[GN_UNROLL];
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u) {
      xv[u] = ldg16(x + o + u * step);
      if (RES) rv[u] = ldg16(res + o + u * step);
    }
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u) one(xv[u], rv[u], o + u * step);
  }
  for (; p < p1; p += pix_per_pass, o += step) {
    const uint4 xv = ldg16(x + o);
    uint4 rv = xv;
    if (RES) rv = ldg16(res + o);
    one(xv, rv, o);
  }
}

// ---------------------------------------------------------------------------------------------
// backward pass 1: per (n, c): S1 = sum dy_eff, S2 = sum dy_eff * xhat  ->  part[n][c][2]
//   MODE 0: dy_eff = dy;  1: dy_eff = dy * (xhat*gamma+beta > 0);  2: dy_eff = dy * (yout > 0)
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(GN_THREADS, 3)
gn_bwd_reduce_kernel(const act_t* __restrict__ x, const float* __restrict__ sums,
                     const float* __restrict__ gamma, const float* __restrict__ beta,
                     const act_t* __restrict__ dy, const act_t* __restrict__ yout,
                     float* __restrict__ part, int HW, int C, int chunk_pixels, float eps) {
  extern __shared__ float smp[];  // [C][2]
  const int n = blockIdx.y;
  const int cv = C >> 3;
  const int my_cv = threadIdx.x % cv;
  const int pix_per_pass = GN_THREADS / cv;
  const int cpg = C / GN_GROUPS;
  const float inv_m = 1.f / ((float)cpg * (float)HW);
  for (int i = threadIdx.x; i < C * 2; i += GN_THREADS) smp[i] = 0.f;
  // mask test (MODE 1): x*a + b > 0.  The loop accumulates S1 = sum dy_eff and R = sum dy_eff * x; the statistics
  // enter once at the end, S2 = rstd * (R - mean * S1), which keeps the loop at two coefficient and two accumulator
  // registers per channel (3 CTAs of 256 threads per SM instead of 2)
  float a[8], b[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = my_cv * 8 + k;
    const int g = c / cpg;
    if (MODE == 1) {
      const float mean = sums[((size_t)n * GN_GROUPS + g) * 2] * inv_m;
      const float var = fmaxf(sums[((size_t)n * GN_GROUPS + g) * 2 + 1] * inv_m - mean * mean, 0.f);
      a[k] = rsqrtf(var + eps) * gamma[c];
      b[k] = beta[c] - mean * a[k];
    }
  }
  __syncthreads();
  float s1[8], s2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s1[k] = s2[k] = 0.f;
  const int p0 = blockIdx.x * chunk_pixels;
  const int p1 = min(HW, p0 + chunk_pixels);
  const size_t base = (size_t)n * HW * C + (size_t)my_cv * 8;
  const size_t step = (size_t)pix_per_pass * C;
  auto one = [&](const uint4& xv, const uint4& dv, const uint4& yv) {
    float f[8], d[8];
    unpack8(xv, f);
    unpack8(dv, d);
    if (MODE == 2) {
      float yo[8];
      unpack8(yv, yo);
#pragma unroll
      for (int k = 0; k < 8; ++k) d[k] = (yo[k] > 0.f) ? d[k] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float de = d[k];
      if (MODE == 1) de = (f[k] * a[k] + b[k] > 0.f) ? de : 0.f;
      s1[k] += de;
      s2[k] += de * f[k];
    }
  };
  int p = p0 + threadIdx.x / cv;
  size_t o = base + (size_t)p * C;
  for (; p + (GN_UNROLL - 1) * pix_per_pass < p1; p += GN_UNROLL * pix_per_pass, o += GN_UNROLL * step) {
    uint4 xv[GN_UNROLL], dv[GN_UNROLL], yv[GN_UNROLL];
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u) {
      xv[u] = ldg16(x + o + u * step);
      dv[u] = ldg16(dy + o + u * step);
      if (MODE == 2) yv[u] = ldg16(yout + o + u * step);
    }
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u) one(xv[u], dv[u], yv[u]);
  }
  for (; p < p1; p += pix_per_pass, o += step) {
    const uint4 xv = ldg16(x + o), dv = ldg16(dy + o);
    uint4 yv = xv;
    if (MODE == 2) yv = ldg16(yout + o);
    one(xv, dv, yv);
  }
  // narrow layers (C/8 < 32): several lanes of a warp own the same channels -- combine them with shuffles first, so
  // the shared-memory atomics below see 8-way instead of 32-way same-address conflicts (C = 64: 30 -> 15 us)
  for (int off = cv; off < 32; off <<= 1) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], off);
      s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], off);
    }
  }
  if (cv >= 32 || (threadIdx.x & 31) < cv) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = my_cv * 8 + k;
      const int g = c / cpg;
      const float mean = sums[((size_t)n * GN_GROUPS + g) * 2] * inv_m;
      const float var = fmaxf(sums[((size_t)n * GN_GROUPS + g) * 2 + 1] * inv_m - mean * mean, 0.f);
      atomicAdd(&smp[c * 2], s1[k]);
      atomicAdd(&smp[c * 2 + 1], (s2[k] - mean * s1[k]) * rsqrtf(var + eps));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 2; i += GN_THREADS) {
    const float v = smp[i];
    if (v != 0.f) atomicAdd(&part[(size_t)n * C * 2 + i], v);
  }
}

// backward pass 2: dx = rstd * (gamma*dy_eff - (A_g + xhat*B_g)/m) = dy_eff*k1 + x*k2 + k3;  optional d_res = dy_eff
template <int MODE, bool DRES>
__global__ void __launch_bounds__(GN_THREADS, 3)
gn_bwd_apply_kernel(const act_t* __restrict__ x, const float* __restrict__ sums,
                    const float* __restrict__ gamma, const float* __restrict__ beta,
                    const act_t* __restrict__ dy, const act_t* __restrict__ yout,
                    const float* __restrict__ part, act_t* __restrict__ dx, act_t* __restrict__ dres,
                    int HW, int C, int chunk_pixels, float eps) {
  __shared__ float gA[GN_GROUPS], gB[GN_GROUPS];
  const int n = gridDim.y - 1 - blockIdx.y;
  const int bx = gridDim.x - 1 - blockIdx.x;
  const int cv = C >> 3;
  const int my_cv = threadIdx.x % cv;
  const int pix_per_pass = GN_THREADS / cv;
  const int cpg = C / GN_GROUPS;
  const float inv_m = 1.f / ((float)cpg * (float)HW);
  if (threadIdx.x < GN_GROUPS) {
    float A = 0.f, B = 0.f;
    for (int c = threadIdx.x * cpg; c < (threadIdx.x + 1) * cpg; ++c) {
      A += gamma[c] * part[((size_t)n * C + c) * 2];
      B += gamma[c] * part[((size_t)n * C + c) * 2 + 1];
    }
    gA[threadIdx.x] = A * inv_m;
    gB[threadIdx.x] = B * inv_m;
  }
  __syncthreads();
  float k1[8], k2[8], k3[8], b[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = my_cv * 8 + k;
    const int g = c / cpg;
    const float mean = sums[((size_t)n * GN_GROUPS + g) * 2] * inv_m;
    const float var = fmaxf(sums[((size_t)n * GN_GROUPS + g) * 2 + 1] * inv_m - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    k1[k] = rstd * gamma[c];
    k2[k] = -rstd * rstd * gB[g];
    k3[k] = -rstd * gA[g] - mean * k2[k];
    if (MODE == 1) b[k] = beta[c] - mean * k1[k];
  }
  const int p0 = bx * chunk_pixels;
  const int p1 = min(HW, p0 + chunk_pixels);
  const size_t base = (size_t)n * HW * C + (size_t)my_cv * 8;
  const size_t step = (size_t)pix_per_pass * C;
  auto one = [&](const uint4& xv, const uint4& dv, const uint4& yv, size_t o) {
    float f[8], d[8];
    unpack8(xv, f);
    unpack8(dv, d);
    if (MODE == 2) {
      float yo[8];
      unpack8(yv, yo);
#pragma unroll
      for (int k = 0; k < 8; ++k) d[k] = (yo[k] > 0.f) ? d[k] : 0.f;
    }
    float r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (MODE == 1) d[k] = (f[k] * k1[k] + b[k] > 0.f) ? d[k] : 0.f;
      r[k] = d[k] * k1[k] + (f[k] * k2[k] + k3[k]);
    }
    store8(dx + o, r);
    if (DRES) store8(dres + o, d);
  };
  int p = p0 + threadIdx.x / cv;
  size_t o = base + (size_t)p * C;
  for (; p + (GN_UNROLL - 1) * pix_per_pass < p1; p += GN_UNROLL * pix_per_pass, o += GN_UNROLL * step) {
    uint4 xv[GN_UNROLL], dv[GN_UNROLL], yv[GN_UNROLL];
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u) {
      xv[u] = ldg16(x + o + u * step);
      dv[u] = ldg16(dy + o + u * step);
      if (MODE == 2) yv[u] = ldg16(yout + o + u * step);
    }
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u) one(xv[u], dv[u], yv[u], o + u * step);
  }
  for (; p < p1; p += pix_per_pass, o += step) {
    const uint4 xv = ldg16(x + o), dv = ldg16(dy + o);
    uint4 yv = xv;
    if (MODE == 2) yv = ldg16(yout + o);
    one(xv, dv, yv, o);
  }
}

// dgamma[c] = sum_n part[n][c][1], dbeta[c] = sum_n part[n][c][0]
__global__ void gn_bwd_param_kernel(const float* __restrict__ part, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, int N, int C, float alpha) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, b = 0.f;
  for (int n = 0; n < N; ++n) {
    b += part[((size_t)n * C + c) * 2];
    a += part[((size_t)n * C + c) * 2 + 1];
  }
  dgamma[c] = a * alpha;
  dbeta[c] = b * alpha;
}

static int gn_chunk(int HW, int N, int C, int* chunks) {
  // exactly one wave of 3 resident CTAs per SM with equal work: the per-CTA prologue (32 dependent coefficient loads
  // per thread) and the atomic epilogue are paid once per SM slot; each CTA covers whole unrolled passes of pixels
  const int pix_per_pass = GN_THREADS / (C >> 3) * GN_UNROLL;
  int want = (3 * num_sms()) / N;
  if (want < 1) want = 1;
  int chunk = (HW + want - 1) / want;
  chunk = ((chunk + pix_per_pass - 1) / pix_per_pass) * pix_per_pass;
  if (chunk < pix_per_pass) chunk = pix_per_pass;
  *chunks = (HW + chunk - 1) / chunk;
  return chunk;
}

}  // namespace eosvos

using namespace eosvos;

static int gn_check(int N, int HW, int C) {
  EOSVOS_REQUIRE(N > 0 && HW > 0, "groupnorm: empty tensor");
  EOSVOS_REQUIRE(C % 32 == 0 && C >= 64 && C <= 2048 && (GN_THREADS % (C >> 3)) == 0,
                 "groupnorm: C must be 64..2048 with C/8 dividing 256");
  return 0;
}

extern "C" int eosvos_gn_stats(const void* x, float* sums, int N, int HW, int C, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_TRY(gn_check(N, HW, C));
  cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)N * GN_GROUPS * 2 * sizeof(float), stream);
  if (e != cudaSuccess) return set_cuda_error(e, "gn_stats memset");
  int chunks;
  const int chunk = gn_chunk(HW, N, C, &chunks);
  gn_stats_kernel<<<dim3(chunks, N), GN_THREADS, 0, stream>>>(reinterpret_cast<const act_t*>(x), sums, HW, C,
                                                              chunk);
  return check_launch("gn_stats_kernel");
}

extern "C" int eosvos_gn_apply(const void* x, const float* sums, const float* gamma, const float* beta, const void* res,
                               void* y, int N, int HW, int C, float eps, int relu, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_TRY(gn_check(N, HW, C));
  int chunks;
  const int chunk = gn_chunk(HW, N, C, &chunks);
  if (res)
    gn_apply_kernel<true><<<dim3(chunks, N), GN_THREADS, 0, stream>>>(
        reinterpret_cast<const act_t*>(x), sums, gamma, beta, reinterpret_cast<const act_t*>(res),
        reinterpret_cast<act_t*>(y), HW, C, chunk, eps, relu);
  else
    gn_apply_kernel<false><<<dim3(chunks, N), GN_THREADS, 0, stream>>>(
        reinterpret_cast<const act_t*>(x), sums, gamma, beta, nullptr, reinterpret_cast<act_t*>(y), HW, C, chunk, eps,
        relu);
  return check_launch("gn_apply_kernel");
}

// part: scratch [N][C][2] fp32 (zeroed here).  mask_mode: 0 none, 1 recompute ReLU mask, 2 mask from yout.
extern "C" int eosvos_gn_backward(const void* x, const float* sums, const float* gamma, const float* beta,
                                  const void* dy, const void* yout, float* part, void* dx, void* dres, float* dgamma,
                                  float* dbeta, int N, int HW, int C, float eps, int mask_mode, float alpha,
                                  eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EOSVOS_TRY(gn_check(N, HW, C));
  EOSVOS_REQUIRE(mask_mode != 2 || yout, "groupnorm backward: mask_mode 2 needs yout");
  cudaError_t e = cudaMemsetAsync(part, 0, (size_t)N * C * 2 * sizeof(float), stream);
  if (e != cudaSuccess) return set_cuda_error(e, "gn_backward memset");
  int chunks;
  const int chunk = gn_chunk(HW, N, C, &chunks);
  const act_t* xb = reinterpret_cast<const act_t*>(x);
  const act_t* dyb = reinterpret_cast<const act_t*>(dy);
  const act_t* yb = reinterpret_cast<const act_t*>(yout);
  EOSVOS_REQUIRE(mask_mode >= 0 && mask_mode <= 2, "groupnorm backward: mask_mode must be 0, 1 or 2");
  const dim3 grid(chunks, N);
  const size_t sm = (size_t)C * 2 * sizeof(float);
  act_t* dxb = reinterpret_cast<act_t*>(dx);
  act_t* drb = reinterpret_cast<act_t*>(dres);
#define GN_BWD(MODE)                                                                                                  \
  do {                                                                                                                \
    gn_bwd_reduce_kernel<MODE><<<grid, GN_THREADS, sm, stream>>>(xb, sums, gamma, beta, dyb, yb, part, HW, C, chunk,  \
                                                                 eps);                                                \
    EOSVOS_TRY(check_launch("gn_bwd_reduce_kernel"));                                                                 \
    if (drb)                                                                                                          \
      gn_bwd_apply_kernel<MODE, true><<<grid, GN_THREADS, 0, stream>>>(xb, sums, gamma, beta, dyb, yb, part, dxb,     \
                                                                       drb, HW, C, chunk, eps);                       \
    else                                                                                                              \
      gn_bwd_apply_kernel<MODE, false><<<grid, GN_THREADS, 0, stream>>>(xb, sums, gamma, beta, dyb, yb, part, dxb,    \
                                                                        nullptr, HW, C, chunk, eps);                  \
    EOSVOS_TRY(check_launch("gn_bwd_apply_kernel"));                                                                  \
  } while (0)
  if (mask_mode == 0) GN_BWD(0);
  else if (mask_mode == 1) GN_BWD(1);
  else GN_BWD(2);
#undef GN_BWD
  gn_bwd_param_kernel<<<(C + 127) / 128, 128, 0, stream>>>(part, dgamma, dbeta, N, C, alpha);
  return check_launch("gn_bwd_param_kernel");
}
