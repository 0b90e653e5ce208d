// MetaOptimizer parameter update, fused over all parameter tensors in one launch:
//     p'[i] = p[i] - lr[row(i)] * g[i]        (lr = exp(log_lr) in log mode)
// Reference: src/meta_optim/meta_optim.py:177-214 (step: exp, grad * lr per tensor) and
// src/meta_optim/meta_model.py:78-80 (p - step per tensor) -- >= 402 elementwise launches there,
// one here.  HBM-bound: 12 B / parameter (+ 4 B per learning-rate scalar).
//
// Also: the outer RAdam step of meta-training (src/util/radam.py:28-94 as used from
// src/train_meta.py:361-373: grad / meta_batch, clamp, RAdam, lr clamp) as one flat kernel.
#include "common.h"
#include "../../include/eosvos_b200.h"

namespace eosvos {

constexpr int MU_THREADS = 256;
constexpr int MU_CHUNK = 8192;  // elements per CTA

// table: int64 [T][8] = (p, g, lr, out, numel, row_len, g_taps, g_cin) ; chunks: int32 [n][2] = (tensor, chunk index)
// g_taps > 1: the gradient of a [Cout][Cin][taps] filter is stored as [Cout][taps][Cin] (channels_last, what the wgrad
// kernel's vector-RED epilogue produces); element i = (co, ci, t) of p pairs with g[(co * taps + t) * Cin + ci].  The
// permuted reads stay inside one filter's Cin*taps window, so they are served by L1 / L2, not HBM.
__global__ void __launch_bounds__(MU_THREADS)
meta_update_kernel(const long long* __restrict__ table, const int* __restrict__ chunks, int use_log,
                   int* __restrict__ nonfinite) {
  // `nonfinite` (optional): set to 1 when an updated value is Inf / NaN, i.e. when a gradient of the scaled 16-bit
  // backward overflowed -- checked by the host at its next synchronisation point (MetaOptimizer.check_finite)
  bool bad = false;
  const int t = chunks[blockIdx.x * 2];
  const long long start = (long long)chunks[blockIdx.x * 2 + 1] * (long long)MU_CHUNK;
  const long long* e = table + (size_t)t * 8;
  const float* p = reinterpret_cast<const float*>(e[0]);     // may alias `out` (in-place update, same index)
  const float* __restrict__ g = reinterpret_cast<const float*>(e[1]);
  const float* __restrict__ lr = reinterpret_cast<const float*>(e[2]);
  float* out = reinterpret_cast<float*>(e[3]);
  const long long numel = e[4];
  const long long row_len = e[5];
  const int g_taps = (int)e[6];
  const int g_cin = (int)e[7];
  long long n = numel - start;
  if (n > MU_CHUNK) n = MU_CHUNK;
  if (g_taps > 1) {
    // 32-bit index arithmetic: a single filter tensor has < 2^31 elements
    const unsigned filt = (unsigned)g_taps * (unsigned)g_cin;
    const unsigned rl = (unsigned)row_len;
    const unsigned end = (unsigned)(start + n);
    const bool al16 = (((e[0] | e[3]) & 15) == 0);
    const unsigned nv = al16 ? (unsigned)(n >> 2) : 0u;       // 4 consecutive parameters per thread, 128-bit p / out
    for (unsigned v = threadIdx.x; v < nv; v += MU_THREADS) {
      const unsigned i = (unsigned)start + (v << 2);
      unsigned co = i / filt;
      const unsigned rem = i - co * filt;
      unsigned ci = rem / (unsigned)g_taps;
      unsigned tp = rem - ci * (unsigned)g_taps;
      const float4 pv = *reinterpret_cast<const float4*>(p + i);
      float gq[4], lq[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        gq[k] = __ldg(g + (size_t)co * filt + tp * (unsigned)g_cin + ci);
        const float lv = __ldg(lr + (i + k) / rl);
        lq[k] = use_log ? expf(lv) : lv;
        if (++tp == (unsigned)g_taps) {
          tp = 0;
          if (++ci == (unsigned)g_cin) {
            ci = 0;
            ++co;
          }
        }
      }
      float4 o;
      o.x = __fsub_rn(pv.x, __fmul_rn(gq[0], lq[0]));
      o.y = __fsub_rn(pv.y, __fmul_rn(gq[1], lq[1]));
      o.z = __fsub_rn(pv.z, __fmul_rn(gq[2], lq[2]));
      o.w = __fsub_rn(pv.w, __fmul_rn(gq[3], lq[3]));
      bad |= !(isfinite(o.x) && isfinite(o.y) && isfinite(o.z) && isfinite(o.w));
      *reinterpret_cast<float4*>(out + i) = o;
    }
    for (unsigned i = (unsigned)start + (nv << 2) + threadIdx.x; i < end; i += MU_THREADS) {
      const unsigned co = i / filt;
      const unsigned rem = i - co * filt;
      const unsigned ci = rem / (unsigned)g_taps;
      const unsigned tp = rem - ci * (unsigned)g_taps;
      float lv = __ldg(lr + i / rl);
      if (use_log) lv = expf(lv);
      const float o = __fsub_rn(p[i], __fmul_rn(__ldg(g + (size_t)co * filt + tp * (unsigned)g_cin + ci), lv));
      bad |= !isfinite(o);
      out[i] = o;
    }
    if (bad && nonfinite) atomicOr(nonfinite, 1);
    return;
  }
  const bool aligned = (((e[0] | e[1] | e[3]) & 15) == 0);
  if (aligned) {
    const long long nv = n >> 2;
    for (long long v = threadIdx.x; v < nv; v += MU_THREADS) {
      const long long i = start + (v << 2);
      const float4 pv = *reinterpret_cast<const float4*>(p + i);
      const float4 gv = __ldcs(reinterpret_cast<const float4*>(g + i));
      long long row = i / row_len;
      long long col = i - row * row_len;
      float l[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float lv = __ldg(lr + row);
        l[k] = use_log ? expf(lv) : lv;
        if (++col == row_len) {
          col = 0;
          ++row;
        }
      }
      float4 o;
      o.x = __fsub_rn(pv.x, __fmul_rn(gv.x, l[0]));  // no FMA contraction: bit-exact with p - (g * lr)
      o.y = __fsub_rn(pv.y, __fmul_rn(gv.y, l[1]));
      o.z = __fsub_rn(pv.z, __fmul_rn(gv.z, l[2]));
      o.w = __fsub_rn(pv.w, __fmul_rn(gv.w, l[3]));
      bad |= !(isfinite(o.x) && isfinite(o.y) && isfinite(o.z) && isfinite(o.w));
      *reinterpret_cast<float4*>(out + i) = o;
    }
    for (long long i = start + (nv << 2) + threadIdx.x; i < start + n; i += MU_THREADS) {
      float lv = __ldg(lr + i / row_len);
      if (use_log) lv = expf(lv);
      const float o = __fsub_rn(p[i], __fmul_rn(g[i], lv));
      bad |= !isfinite(o);
      out[i] = o;
    }
  } else {
    for (long long i = start + threadIdx.x; i < start + n; i += MU_THREADS) {
      float lv = __ldg(lr + i / row_len);
      if (use_log) lv = expf(lv);
      const float o = __fsub_rn(p[i], __fmul_rn(g[i], lv));
      bad |= !isfinite(o);
      out[i] = o;
    }
  }
  if (bad && nonfinite) atomicOr(nonfinite, 1);
}

// Meta-gradient of the learning rates through one fused update (first-order BPTT, reference
// src/util/meta_run.py:124-214 via autograd of meta_model.py:78-80):  theta' = theta - lr (.) g  =>
//     d L / d lr[row] = - sum_{j in row} d[j] * g[j]      (x exp(log_lr[row]) in log mode),
// d = d L / d theta'.  One launch over all 201 tensors instead of 201 x (mul, sum, neg).
// table: int64 [T][10] = (d, g, lr, dl, numel, row_len, g_taps, g_cin, d_taps, d_cin): *_taps > 1 marks a K x K filter
// stored channels-last ([Cout][taps][Cin], what the wgrad epilogue emits) instead of [Cout][Cin][taps].
// work: int32 [n][3] = (tensor, first row, kind): kind 0 = one CTA reduces ONE row (row_len >= 256), kind 1 = 256
// consecutive rows, one thread each (short rows).  Deterministic: no atomics.
__global__ void __launch_bounds__(MU_THREADS)
lr_grad_kernel(const long long* __restrict__ table, const int* __restrict__ work, int use_log) {
  const int t = work[blockIdx.x * 3], row0 = work[blockIdx.x * 3 + 1], kind = work[blockIdx.x * 3 + 2];
  const long long* e = table + (size_t)t * 10;
  const float* __restrict__ d = reinterpret_cast<const float*>(e[0]);
  const float* __restrict__ g = reinterpret_cast<const float*>(e[1]);
  const float* __restrict__ lr = reinterpret_cast<const float*>(e[2]);
  float* __restrict__ dl = reinterpret_cast<float*>(e[3]);
  const long long numel = e[4], row_len = e[5];
  const int g_taps = (int)e[6], g_cin = (int)e[7], d_taps = (int)e[8], d_cin = (int)e[9];
  const long long rows = numel / row_len;
  // offset inside a filter (logical element j = ci * taps + tp) for a channels-last operand
  auto off = [](long long j, int taps, int cin) -> long long {
    if (taps <= 1) return j;
    const long long ci = j / taps, tp = j - ci * taps;
    return tp * cin + ci;
  };
  if (kind == 0) {
    const long long base = (long long)row0 * row_len;
    float acc = 0.f;
    if (g_taps <= 1 && d_taps <= 1) {
      for (long long j = threadIdx.x; j < row_len; j += MU_THREADS) acc += __ldcs(d + base + j) * __ldcs(g + base + j);
    } else {
      // a row of a K x K filter tensor is one filter (NEURON level) or the whole tensor: permute inside each filter
      const long long filt = (long long)(g_taps > 1 ? g_taps : d_taps) * (g_taps > 1 ? g_cin : d_cin);
      for (long long j = threadIdx.x; j < row_len; j += MU_THREADS) {
        const long long f = j / filt, r = j - f * filt;
        acc += d[base + f * filt + off(r, d_taps, d_cin)] * g[base + f * filt + off(r, g_taps, g_cin)];
      }
    }
    __shared__ float s_w[MU_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int w = 0; w < MU_THREADS / 32; ++w) tot += s_w[w];
      dl[row0] = -tot * (use_log ? expf(lr[row0]) : 1.f);
    }
    return;
  }
  const long long row = (long long)row0 + threadIdx.x;
  if (row >= rows) return;
  const long long base = row * row_len;
  float acc = 0.f;
  if (g_taps <= 1 && d_taps <= 1) {
    for (long long j = 0; j < row_len; ++j) acc += d[base + j] * g[base + j];
  } else {
    for (long long j = 0; j < row_len; ++j) acc += d[base + off(j, d_taps, d_cin)] * g[base + off(j, g_taps, g_cin)];
  }
  dl[row] = -acc * (use_log ? expf(lr[row]) : 1.f);
}

// Flat RAdam (radam.py:28-94).  Per-element group id selects (lr, weight_decay); the
// rectification terms depend only on the step count and are computed on the host.
//   g' = clamp(g * gscale, -clip, clip) ; m = b1 m + (1-b1) g' ; v = b2 v + (1-b2) g'^2
//   p -= wd*lr*p ; p -= step_size * m / (sqrt(v) + eps)   [n_sma >= 5]   or   p -= step_size * m
__global__ void __launch_bounds__(256)
radam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             long long n, float gscale, float clip, float beta1, float beta2, float omb1, float omb2, float eps, float lr,
             float wd, float step_size, int rectified, float clamp_lo, float clamp_hi, int do_clamp) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gg = g[i] * gscale;
  if (clip > 0.f) gg = fminf(fmaxf(gg, -clip), clip);
  // omb = 1 - beta rounded from double on the host, as the reference's Python scalars are (radam.py:60-61)
  const float mm = __fadd_rn(__fmul_rn(beta1, m[i]), __fmul_rn(omb1, gg));
  const float vv = __fadd_rn(__fmul_rn(beta2, v[i]), __fmul_rn(__fmul_rn(omb2, gg), gg));
  m[i] = mm;
  v[i] = vv;
  float pp = p[i];
  if (wd != 0.f) pp += -wd * lr * pp;
  if (rectified)
    pp += -step_size * mm / (sqrtf(vv) + eps);
  else
    pp += -step_size * mm;
  if (do_clamp) pp = fminf(fmaxf(pp, clamp_lo), clamp_hi);
  p[i] = pp;
}

}  // namespace eosvos

using namespace eosvos;

extern "C" int eosvos_meta_update_chunk_elems(void) { return MU_CHUNK; }

extern "C" int eosvos_meta_update(const long long* table_dev, const int* chunks_dev, int num_chunks, int use_log,
                                  int* nonfinite_flag, eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (num_chunks == 0) return 0;
  EOSVOS_REQUIRE(table_dev && chunks_dev && num_chunks > 0, "meta_update: null table");
  meta_update_kernel<<<num_chunks, MU_THREADS, 0, stream>>>(table_dev, chunks_dev, use_log, nonfinite_flag);
  return check_launch("meta_update_kernel");
}

extern "C" int eosvos_lr_grad(const long long* table_dev, const int* work_dev, int num_work, int use_log,
                              eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (num_work == 0) return 0;
  EOSVOS_REQUIRE(table_dev && work_dev && num_work > 0, "lr_grad: null table");
  lr_grad_kernel<<<num_work, MU_THREADS, 0, stream>>>(table_dev, work_dev, use_log);
  return check_launch("lr_grad_kernel");
}

extern "C" int eosvos_radam_step(float* p, const float* g, float* m, float* v, long long n, float gscale, float clip,
                                 float beta1, float beta2, float omb1, float omb2, float eps, float lr, float wd,
                                 float step_size, int rectified, float clamp_lo, float clamp_hi, int do_clamp,
                                 eosvos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (n == 0) return 0;
  EOSVOS_REQUIRE(p && g && m && v, "radam_step: null pointer");
  radam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p, g, m, v, n, gscale, clip, beta1, beta2, omb1, omb2, eps,
                                                               lr, wd, step_size, rectified, clamp_lo, clamp_hi,
                                                               do_clamp);
  return check_launch("radam_kernel");
}
