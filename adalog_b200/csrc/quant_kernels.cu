// Elementwise fake-quant forwards, self-error sweeps and bf16 operand generators.
// Compiled with -fmad=false: every multiply/add below rounds exactly like the separate eager torch
// kernels of the reference (no FMA contraction), which is what makes the integer codes bit-exact.
// HBM-bound kernels: 128-bit loads/stores, grid sized in multiples of the SM count.
#include "common.cuh"
#include "../../include/adalog_b200.h"
#include <stdarg.h>

namespace adalog {

static thread_local char g_err[512];
char* err_buf() { return g_err; }
int fail(int code, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
  return code;
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(-100, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// K1: uniform fake-quant forward (quantizers/uniform.py:25-36)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void uq_fwd(float x, float s, float z, int nl, bool sym, float& y, float& c) {
  float r = rintf(__fdiv_rn(x, s));
  if (sym) {
    c = fminf(fmaxf(r, -(float)nl), (float)(nl - 1));
    y = __fmul_rn(c, s);
  } else {
    c = fminf(fmaxf(__fadd_rn(r, z), 0.0f), (float)(2 * nl - 1));
    y = __fmul_rn(__fsub_rn(c, z), s);
  }
}

// VEC4: n % 4 == 0, pointers 16B aligned and (ngroups == 1 or inner % 4 == 0) so a float4 never straddles groups
template <bool VEC4>
__global__ void __launch_bounds__(256) uniform_fakequant_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                               int16_t* __restrict__ codes, int64_t n,
                                                               const float* __restrict__ scale,
                                                               const float* __restrict__ zp, int64_t inner,
                                                               int64_t ngroups, int nl, int sym) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool single = ngroups == 1;
  float s0 = scale[0], z0 = (zp != nullptr) ? zp[0] : 0.0f;
  if (VEC4) {
    const int64_t n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (; i < n4; i += stride) {
      float4 v = __ldg(x4 + i);
      float s = s0, z = z0;
      if (!single) {
        int64_t g = ((i << 2) / inner) % ngroups;
        s = __ldg(scale + g);
        z = zp ? __ldg(zp + g) : 0.0f;
      }
      float4 o; float c0, c1, c2, c3;
      uq_fwd(v.x, s, z, nl, sym, o.x, c0);
      uq_fwd(v.y, s, z, nl, sym, o.y, c1);
      uq_fwd(v.z, s, z, nl, sym, o.z, c2);
      uq_fwd(v.w, s, z, nl, sym, o.w, c3);
      if (y) reinterpret_cast<float4*>(y)[i] = o;
      if (codes) {
        short4 cc = make_short4((short)c0, (short)c1, (short)c2, (short)c3);
        reinterpret_cast<short4*>(codes)[i] = cc;
      }
    }
  } else {
    for (; i < n; i += stride) {
      float s = s0, z = z0;
      if (!single) {
        int64_t g = (i / inner) % ngroups;
        s = __ldg(scale + g);
        z = zp ? __ldg(zp + g) : 0.0f;
      }
      float o, c;
      uq_fwd(__ldg(x + i), s, z, nl, sym, o, c);
      if (y) y[i] = o;
      if (codes) codes[i] = (int16_t)c;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K2: log-family fake-quant forward (quantizers/logarithm.py)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) log_fakequant_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                           int16_t* __restrict__ codes, int64_t n,
                                                           const float* __restrict__ scale, int kind, int nl,
                                                           const long long* __restrict__ q,
                                                           const float* __restrict__ table1,
                                                           const float* __restrict__ table2,
                                                           const float* __restrict__ shift, int sub_shift) {
  __shared__ float t1s[256], t2s[256];
  const int ncode = 2 * nl;
  if (kind == 2) {
    for (int i = threadIdx.x; i < ncode && i < 256; i += blockDim.x) { t1s[i] = table1[i]; t2s[i] = table2[i]; }
  }
  __syncthreads();
  const float s = scale[0];
  const float sh = shift ? shift[0] : 0.0f;
  const float qf = (kind == 2) ? (float)q[0] : 1.0f;
  const float top = (float)(ncode - 1);
  const float kSqrt2m1 = (float)(1.4142135623730951 - 1.0);  // math.sqrt(2) - 1 cast to FP32 by torch
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float xv = __ldg(x + i);
    if (shift) xv = __fadd_rn(xv, sh);
    float v = fminf(fmaxf(__fdiv_rn(xv, s), 1e-15f), 1.0f);
    float nlg = -log2f(v);
    float c;
    if (kind == 0)      c = rintf(nlg);
    else if (kind == 1) c = rintf(__fmul_rn(nlg, 2.0f));
    else                c = rintf(__fdiv_rn(__fmul_rn(nlg, 37.0f), qf));
    const float mask = (c < (float)ncode) ? 1.0f : 0.0f;
    c = fminf(fmaxf(c, 0.0f), top);
    float d;
    if (kind == 0) {
      d = __fmul_rn(ldexpf(1.0f, -(int)c), s);
    } else if (kind == 1) {
      float odd = __fadd_rn(__fmul_rn(fmodf(c, 2.0f), kSqrt2m1), 1.0f);
      d = __fmul_rn(__fmul_rn(ldexpf(1.0f, -(int)ceilf(__fdiv_rn(c, 2.0f))), odd), s);
    } else {
      int ci = (int)c;
      d = __fmul_rn(__fmul_rn(ldexpf(1.0f, -(int)t1s[ci]), t2s[ci]), s);
    }
    d = __fmul_rn(d, mask);
    if (sub_shift) d = __fsub_rn(d, sh);
    if (y) y[i] = d;
    if (codes) codes[i] = (int16_t)c;
  }
}

__global__ void __launch_bounds__(256) twin_fakequant_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                            int64_t n, const float* __restrict__ scale2, int nl) {
  const float sp = scale2[0], sn = scale2[1];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float xv = __ldg(x + i);
    float p = __fmul_rn(fminf(fmaxf(rintf(__fdiv_rn(xv, sp)), 0.0f), (float)(nl - 1)), sp);
    float m = __fmul_rn(fminf(fmaxf(rintf(__fdiv_rn(xv, sn)), -(float)nl), 0.0f), sn);
    y[i] = __fadd_rn(p, m);
  }
}

// ------------------------------------------------------------------------------------------------
// K3: weight self-error sweep (linear.py:296-309).  One CTA per weight row, one thread per candidate.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) sweep_err_w_self_kernel(const float* __restrict__ W, int R, int K,
                                                              const float* __restrict__ cs,
                                                              const float* __restrict__ cz, int P, int nl,
                                                              double* __restrict__ err_sum) {
  extern __shared__ float wrow[];
  const int r = blockIdx.x;
  for (int k = threadIdx.x; k < K; k += blockDim.x) wrow[k] = W[(int64_t)r * K + k];
  __syncthreads();
  const int p = threadIdx.x;
  if (p >= P) return;
  const float s = cs[(int64_t)p * R + r], z = cz[(int64_t)p * R + r];
  const float L = (float)(2 * nl - 1);
  double acc = 0.0;
  for (int k0 = 0; k0 < K; k0 += 64) {
    float a = 0.0f;
    const int k1 = min(K, k0 + 64);
    for (int k = k0; k < k1; ++k) {
      float w = wrow[k];
      float d = __fsub_rn(w, __fmul_rn(uq_int(w, s, z, L), s));
      a = __fadd_rn(a, __fmul_rn(d, d));
    }
    acc += (double)a;
  }
  err_sum[(int64_t)p * R + r] = acc;
}

// ------------------------------------------------------------------------------------------------
// K4: activation self-error sweep (linear.py:320-345).  blockDim (32, 8): lane = column of a 32-wide
// column tile, y = group of 16 candidates kept in registers; rows are streamed once.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sweep_err_a_self_kernel(const float* __restrict__ x, int64_t n_total,
                                                              int Cw, int per_channel,
                                                              const float* __restrict__ cs,
                                                              const float* __restrict__ cz, int P, int nl,
                                                              double* __restrict__ partial, int nsplit) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int p0 = threadIdx.y * 16;
  const bool col_ok = c < Cw;
  float s[16], z[16];
  double acc[16];
  float a32[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    int p = min(p0 + j, P - 1);
    int64_t ci = per_channel ? ((int64_t)min(c, Cw - 1) * P + p) : p;
    s[j] = __ldg(cs + ci);
    z[j] = __ldg(cz + ci);
    acc[j] = 0.0;
    a32[j] = 0.0f;
  }
  const float L = (float)(2 * nl - 1);
  const int64_t M = (n_total + Cw - 1) / Cw;
  const int64_t rps = (M + nsplit - 1) / nsplit;
  const int64_t m0 = (int64_t)blockIdx.y * rps;
  const int64_t m1 = min(M, m0 + rps);
  int cnt = 0;
  for (int64_t m = m0; m < m1; ++m) {
    const int64_t idx = m * Cw + c;
    if (col_ok && idx < n_total) {
      const float xv = __ldg(x + idx);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float d = __fsub_rn(xv, __fmul_rn(uq_int(xv, s[j], z[j], L), s[j]));
        a32[j] = __fadd_rn(a32[j], __fmul_rn(d, d));
      }
    }
    if (++cnt == 32) {
      cnt = 0;
#pragma unroll
      for (int j = 0; j < 16; ++j) { acc[j] += (double)a32[j]; a32[j] = 0.0f; }
    }
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] += (double)a32[j];
  if (per_channel) {
    if (col_ok) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (p0 + j < P) partial[((int64_t)blockIdx.y * Cw + c) * P + p0 + j] = acc[j];
    }
  } else {
    // per-tensor: fixed-order butterfly over the 32 lanes, one output per (split, candidate)
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      double v = acc[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (threadIdx.x == 0 && p0 + j < P) partial[(int64_t)blockIdx.y * P + p0 + j] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// operand generators (bf16 integer parts; K-major rows of pitch kpad, zero padded)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store8(uint16_t* dst, const float (&v)[8]) {
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
  o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(dst) = o;
}

__global__ void __launch_bounds__(256) gen_uniform_fixed_kernel(const float* __restrict__ x, int64_t R, int K,
                                                               int64_t ldx, const float* __restrict__ scale,
                                                               const float* __restrict__ zp, int64_t g_div,
                                                               int64_t g_mod, int nl, uint16_t* __restrict__ out,
                                                               int kpad, float* __restrict__ rowsum) {
  const int cpr = kpad >> 3;
  const int64_t total = R * cpr;
  const float L = (float)(2 * nl - 1);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / cpr;
    const int kc = (int)(idx - r * cpr) << 3;
    const int64_t g = (r / g_div) % g_mod;
    const float s = __ldg(scale + g), z = __ldg(zp + g);
    float v[8];
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = kc + j;
      v[j] = (k < K) ? uq_int(__ldg(x + r * ldx + k), s, z, L) : 0.0f;
      sum += v[j];
    }
    store8(out + r * kpad + kc, v);
    if (rowsum && kc < K) atomicAdd(rowsum + r, sum);  // integers: exact, order independent
  }
}

// ------------------------------------------------------------------------------------------------
// Candidate expansion: one CTA per unit.  Each thread keeps an 8-element K chunk of the FP32 row in registers and
// loops over the candidates, so the 128x expansion costs ~8 ALU instructions per generated element.
//
// EXACT FAST PATH for rint(x / s) (the reference's torch.round(x / s), linear.py:304/:409):
//   r = fl(1/s) (IEEE), q = fl(x*r)  =>  |q - x/s| <= 2 ulp(q), and the reference value fl(x/s) is within 0.5 ulp of
//   x/s.  For |q| <= 256, ulp <= 2^-16, so |q - fl(x/s)| <= 2.5*2^-16 < 2^-14.  Hence if q is farther than 2^-14 from
//   every half-integer, rint(q) == rint(fl(x/s)) (no tie can be involved either).  Elements closer than that
//   (probability ~1.2e-4) take the IEEE-division path element-wise.  Values more than 1 outside the clamp range
//   [lo, hi] = [-zp, L-zp] are decided by the clamp whatever the rounding.  rint is done with the 1.5*2^23 trick.
// ------------------------------------------------------------------------------------------------
constexpr float kMagic = 12582912.0f;                 // 1.5 * 2^23
constexpr float kFracSafe = 0.5f - 6.103515625e-05f;  // 0.5 - 2^-14

__device__ __forceinline__ float rint_magic(float q) { return __fsub_rn(__fadd_rn(q, kMagic), kMagic); }

// c = {1/s, lo = -zp, hi = L - zp, s}.  Returns the clamped integer; sets `unsafe` when the element sits within
// 2^-14 of a rounding boundary AND the rounding can change the clamped result (lo <= rint <= hi): one step beyond
// the clamp range the neighbouring integer clamps to the same value.  NaN compares false everywhere -> unsafe.
__device__ __forceinline__ float uq_int_fast(float x, const float4 c, bool& unsafe) {
  const float q = __fmul_rn(x, c.x);
  const float t = rint_magic(q);
  const float f = fabsf(__fsub_rn(q, t));
  const float tc = fminf(fmaxf(t, c.y), c.z);
  unsafe |= !(f <= kFracSafe) && !(tc != t);
  return tc;
}

__global__ void __launch_bounds__(256) gen_uniform_cand_kernel(const float* __restrict__ x, int K, int64_t ldx,
                                                              const float* __restrict__ cs,
                                                              const float* __restrict__ cz, int P, int64_t pstride,
                                                              int64_t gstride, int64_t g_div, int64_t g_mod,
                                                              int64_t u_base, int nl, uint16_t* __restrict__ out,
                                                              int kpad, int krep, float* __restrict__ rowsum,
                                                              int tpc) {
  __shared__ float4 cand[ADALOG_P];
  __shared__ float rsum[ADALOG_P];
  const int64_t u = blockIdx.x;
  const int64_t g = ((u_base + u) / g_div) % g_mod;
  const float L = (float)(2 * nl - 1);
  for (int p = threadIdx.x; p < ADALOG_P; p += blockDim.x) {
    const int pp = min(p, P - 1);          // pad rows repeat the last candidate
    const float s = __ldg(cs + pp * pstride + g * gstride);
    const float z = __ldg(cz + pp * pstride + g * gstride);
    float r = __fdiv_rn(1.0f, s);
    if (z != rintf(z)) r = __int_as_float(0x7fc00000);   // non-integer zero point: always take the IEEE path
    cand[p] = make_float4(r, -z, L - z, s);
    rsum[p] = 0.0f;
  }
  __syncthreads();
  const int cpr = kpad >> 3;
  const int npg = blockDim.x / tpc;                      // host guarantees blockDim.x == tpc * npg
  const int lane_chunk = threadIdx.x % tpc, pg = threadIdx.x / tpc;
  const int per = ADALOG_P / gridDim.y;                  // candidates of this CTA: [p_lo, p_lo + per)
  const int p_lo = blockIdx.y * per;
  const int64_t pitch = (int64_t)krep * kpad;
  uint16_t* obase = out + u * ADALOG_P * pitch;
  const float* xrow = x + u * ldx;
  for (int ch = lane_chunk; ch < cpr; ch += tpc) {
    const int kc = ch << 3;
    float xv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) xv[j] = (kc + j < K) ? __ldg(xrow + kc + j) : 0.0f;
    for (int p = p_lo + pg; p < p_lo + per; p += npg) {
      const float4 c = cand[p];
      float v[8];
      bool unsafe = false;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = uq_int_fast(xv[j], c, unsafe);
      if (unsafe) {                                      // ~3% of warps: redo the chunk on the IEEE path
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = uq_int(xv[j], c.w, -c.y, c.z - c.y);
      }
      if (kc + 8 > K) {
#pragma unroll
        for (int j = 0; j < 8; ++j) if (kc + j >= K) v[j] = 0.0f;
      }
      uint16_t* dst = obase + p * pitch + kc;
      for (int rep = 0; rep < krep; ++rep) store8(dst + (int64_t)rep * kpad, v);
      if (rowsum) {
        float sum = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += v[j];
        atomicAdd(rsum + p, sum);
      }
    }
  }
  if (rowsum) {
    __syncthreads();
    for (int p = p_lo + threadIdx.x; p < p_lo + per; p += blockDim.x) rowsum[u * ADALOG_P + p] = rsum[p];
  }
}

// ------------------------------------------------------------------------------------------------
// AdaLog search form (linear.py:872-878 / :913-919, matmul.py:337-342): code from (scale_p, q_p), value m*2^-e.
//
// Reference chain per (element, candidate): v = clamp((x+shift)/s, 1e-15, 1); T = fl(fl(-log2f(v)*37)/q);
// c = rint(T).  EXACT FAST PATH: lx = -log2f(x+shift) once per element, ls = -log2f(s), kq = fl(37/q) once per
// candidate, t = fma(lx, kq, -ls*kq).  With log2f <= 1 ulp and |lx - ls| <= 49 (outside: IEEE path) one gets
// |t - T| < 6e-5, so when t is farther than 2.5e-4 from every half-integer rint(t) == c.
// Unscaled form (post-softmax: no division, no clamp): l37 = fl(lx*37) is the reference's own intermediate and
// t = fl(l37 * fl(1/q)) is within 2.5 ulp of T: margin 2^-14 as for the uniform case.
// Codes >= 2n are masked to zero whatever their exact value, so t >= 2n - 0.5 + margin needs no check (covers +inf).
// e = floor(c*q/37) and (c*q) mod 37 are taken in exact float integer arithmetic (c*q < 2^22).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float log_value_slow(float xs, float lx, bool scaled, float s, float qf,
                                                const float* mt, float ncode) {
  float nlg = lx;
  if (scaled) nlg = -log2f(fminf(fmaxf(__fdiv_rn(xs, s), 1e-15f), 1.0f));
  float c = rintf(__fdiv_rn(__fmul_rn(nlg, 37.0f), qf));
  if (!(c < ncode)) return 0.0f;                        // +inf / NaN are masked like the reference
  c = fmaxf(c, 0.0f);
  const int cqi = (int)c * (int)qf;
  if (cqi / 37 > 120) return 0.0f;                      // below 2^-113: flushed on both paths
  return ldexpf(mt[cqi % 37], -(cqi / 37));
}

__global__ void __launch_bounds__(256) gen_log_cand_kernel(const float* __restrict__ x, int K, int64_t ldx,
                                                          const float* __restrict__ cs,
                                                          const long long* __restrict__ cq, int P,
                                                          const float* __restrict__ shift,
                                                          const float* __restrict__ mtab, int nl,
                                                          uint16_t* __restrict__ out, int kpad, int tpc) {
  __shared__ float4 cand[ADALOG_P];        // {mul, off, lim, q}
  __shared__ float cscale[ADALOG_P];
  __shared__ float mt[40];
  const int64_t u = blockIdx.x;
  const bool scaled = cs != nullptr;
  const float sh = shift ? shift[0] : 0.0f;
  const float margin = scaled ? 2.5e-4f : 6.103515625e-05f;
  for (int p = threadIdx.x; p < ADALOG_P; p += blockDim.x) {
    const int pp = min(p, P - 1);
    const float qf = (float)cq[pp];
    const float s = scaled ? __ldg(cs + pp) : 1.0f;
    float mul, off, lim;
    if (scaled) {
      const float ls = -log2f(s);
      mul = __fdiv_rn(37.0f, qf);
      off = -__fmul_rn(ls, mul);
      lim = __fadd_rn(ls, 49.0f);                       // lx - ls <= 49  <=>  v comfortably above the 1e-15 clamp
    } else {
      mul = __fdiv_rn(1.0f, qf);
      off = 0.0f;
      lim = __int_as_float(0x7f800000);                 // +inf: no clamp in the post-softmax form
    }
    cand[p] = make_float4(mul, off, lim, qf);
    cscale[p] = s;
  }
  for (int j = threadIdx.x; j < 37; j += blockDim.x) mt[j] = mtab[j];
  __syncthreads();
  const float ncode = (float)(2 * nl);
  const float t_masked = ncode - 0.5f + margin;
  const float frac_safe = 0.5f - margin;
  const float inv37 = 1.0f / 37.0f;
  const int cpr = kpad >> 3;
  const int npg = blockDim.x / tpc;
  const int lane_chunk = threadIdx.x % tpc, pg = threadIdx.x / tpc;
  const int per = ADALOG_P / gridDim.y;
  const int p_lo = blockIdx.y * per;
  uint16_t* obase = out + u * ADALOG_P * (int64_t)kpad;
  const float* xrow = x + u * ldx;
  for (int ch = lane_chunk; ch < cpr; ch += tpc) {
    const int kc = ch << 3;
    float xs[8], lx[8], e1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = (kc + j < K) ? __ldg(xrow + kc + j) : 1.0f;
      if (shift) v = __fadd_rn(v, sh);
      xs[j] = v;
      lx[j] = -log2f(v);                                // x <= 0 gives +inf / NaN -> IEEE path below
      e1[j] = scaled ? lx[j] : __fmul_rn(lx[j], 37.0f);
    }
    for (int p = p_lo + pg; p < p_lo + per; p += npg) {
      const float4 c = cand[p];
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float val;
        const float t = fmaxf(fmaf(e1[j], c.x, c.y), 0.0f);
        const float cr = rint_magic(t);
        const float f = fabsf(__fsub_rn(t, cr));
        if (!(lx[j] <= c.z)) {
          // inside (or near) the 1e-15 clamp of the reference, or x <= 0: the closed form does not apply
          val = log_value_slow(xs[j], lx[j], scaled, cscale[p], c.w, mt, ncode);
        } else if (t >= t_masked) {
          val = 0.0f;                                   // code >= 2n whatever the rounding: masked (covers +inf)
        } else if (f <= frac_safe) {
          const float cqf = __fmul_rn(cr, c.w);                         // exact integer c*q
          const float ef = rint_magic(fmaf(__fadd_rn(cqf, 0.5f), inv37, -0.5f));   // floor(c*q/37)
          const float idxf = fmaf(-37.0f, ef, cqf);                     // (c*q) mod 37, exact
          const int idx = __float_as_int(__fadd_rn(idxf, kMagic)) & 0x3f;
          const int ei = __float_as_int(__fadd_rn(ef, kMagic)) & 0xfff;
          val = (ei > 120) ? 0.0f : __int_as_float(__float_as_int(mt[idx]) - (ei << 23));
        } else {
          val = log_value_slow(xs[j], lx[j], scaled, cscale[p], c.w, mt, ncode);
        }
        v[j] = (kc + j < K) ? val : 0.0f;
      }
      store8(obase + p * (int64_t)kpad + kc, v);
    }
  }
}

// AdaLog inference form with the quantizer's own LUTs (logarithm.py:87-99)
__global__ void __launch_bounds__(256) gen_log_fixed_kernel(const float* __restrict__ x, int64_t R, int K,
                                                           int64_t ldx, const float* __restrict__ scale,
                                                           const long long* __restrict__ q,
                                                           const float* __restrict__ shift,
                                                           const float* __restrict__ table1,
                                                           const float* __restrict__ m2, int nl,
                                                           uint16_t* __restrict__ out, int kpad) {
  __shared__ float t1s[256], m2s[256];
  const int ncode = 2 * nl;
  for (int i = threadIdx.x; i < ncode && i < 256; i += blockDim.x) { t1s[i] = table1[i]; m2s[i] = m2[i]; }
  __syncthreads();
  const float s = scale[0], qf = (float)q[0];
  const float sh = shift ? shift[0] : 0.0f;
  const int cpr = kpad >> 3;
  const int64_t total = R * cpr;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / cpr;
    const int kc = (int)(idx - r * cpr) << 3;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = kc + j;
      float val = 0.0f;
      if (k < K) {
        float xv = __ldg(x + r * ldx + k);
        if (shift) xv = __fadd_rn(xv, sh);
        float vv = fminf(fmaxf(__fdiv_rn(xv, s), 1e-15f), 1.0f);
        float c = rintf(__fdiv_rn(__fmul_rn(-log2f(vv), 37.0f), qf));
        if (c < (float)ncode) {
          const int ci = (int)fmaxf(c, 0.0f);
          val = ldexpf(m2s[ci], -(int)t1s[ci]);
        }
      }
      v[j] = val;
    }
    store8(out + r * kpad + kc, v);
  }
}

__global__ void __launch_bounds__(256) gen_split3_kernel(const float* __restrict__ x, int64_t R, int K, int64_t ldx,
                                                        uint16_t* __restrict__ out, int kpad) {
  const int cpr = kpad >> 3;
  const int64_t total = R * cpr;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / cpr;
    const int kc = (int)(idx - r * cpr) << 3;
    float h[8], m[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = kc + j;
      const float xv = (k < K) ? __ldg(x + r * ldx + k) : 0.0f;
      h[j] = __bfloat162float(__float2bfloat16_rn(xv));
      const float r1 = __fsub_rn(xv, h[j]);
      m[j] = __bfloat162float(__float2bfloat16_rn(r1));
      l[j] = __bfloat162float(__float2bfloat16_rn(__fsub_rn(r1, m[j])));
    }
    uint16_t* dst = out + r * (3 * (int64_t)kpad) + kc;
    store8(dst, h);
    store8(dst + kpad, m);
    store8(dst + 2 * (int64_t)kpad, l);
  }
}

// candidate split (gridDim.y) of the expansion kernels: enough CTAs to fill the chip even for a short unit list,
// while every CTA keeps at least `npg` candidates per pass
static inline int cand_split(int64_t U, int npg) {
  int ps = 1;
  while (ps < 32 && U * ps < (int64_t)kNumSMs * 8 && (ADALOG_P / (ps * 2)) >= npg) ps *= 2;
  return ps;
}

static inline int grid_for(int64_t work_items, int threads, int per_sm = 8) {
  int64_t need = (work_items + threads - 1) / threads;
  int64_t cap = (int64_t)kNumSMs * per_sm;
  return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

}  // namespace adalog

using namespace adalog;

extern "C" {

int adalog_version(void) { return 100; }
const char* adalog_last_error(void) { return err_buf(); }

int adalog_uniform_fakequant_f32(const float* x, float* y, int16_t* codes, int64_t n, const float* scale,
                                 const float* zp, int64_t inner, int64_t ngroups, int n_levels, int symmetric,
                                 void* stream) {
  if (n == 0) return 0;
  ADALOG_REQUIRE(x && scale && n > 0 && inner > 0 && ngroups > 0, -1, "uniform_fakequant: bad arguments");
  ADALOG_REQUIRE(symmetric || zp, -1, "uniform_fakequant: zp required for the asymmetric form");
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(codes) & 7) == 0;
  const bool vec = aligned && (n % 4 == 0) && (ngroups == 1 || inner % 4 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  if (vec)
    uniform_fakequant_kernel<true><<<grid_for(n / 4, 256), 256, 0, st>>>(x, y, codes, n, scale, zp, inner, ngroups,
                                                                        n_levels, symmetric);
  else
    uniform_fakequant_kernel<false><<<grid_for(n, 256), 256, 0, st>>>(x, y, codes, n, scale, zp, inner, ngroups,
                                                                     n_levels, symmetric);
  return check_launch("uniform_fakequant");
}

int adalog_log_fakequant_f32(const float* x, float* y, int16_t* codes, int64_t n, const float* scale, int kind,
                             int n_levels, const long long* q, const float* table1, const float* table2,
                             const float* shift, int sub_shift, void* stream) {
  if (n == 0) return 0;
  ADALOG_REQUIRE(x && scale && n > 0 && kind >= 0 && kind <= 2, -1, "log_fakequant: bad arguments");
  ADALOG_REQUIRE(kind != 2 || (q && table1 && table2), -1, "log_fakequant: adalog needs q/table1/table2");
  ADALOG_REQUIRE(n_levels <= 128, -1, "log_fakequant: n_levels > 128 unsupported");
  if (n == 0) return 0;
  log_fakequant_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, y, codes, n, scale, kind, n_levels, q,
                                                                          table1, table2, shift, sub_shift);
  return check_launch("log_fakequant");
}

int adalog_twin_fakequant_f32(const float* x, float* y, int64_t n, const float* scale2, int n_levels, void* stream) {
  if (n == 0) return 0;
  ADALOG_REQUIRE(x && y && scale2 && n > 0, -1, "twin_fakequant: bad arguments");
  twin_fakequant_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, y, n, scale2, n_levels);
  return check_launch("twin_fakequant");
}

int adalog_sweep_err_w_self(const float* W, int R, int K, const float* cs, const float* cz, int P, int n_levels,
                            double* err_sum, void* stream) {
  ADALOG_REQUIRE(W && cs && cz && err_sum && R > 0 && K > 0 && P > 0 && P <= ADALOG_P, -1,
                 "sweep_err_w_self: bad arguments (P must be <= 128)");
  ADALOG_REQUIRE((size_t)K * 4 <= 200 * 1024, -1, "sweep_err_w_self: K too large for shared memory");
  size_t smem = (size_t)K * sizeof(float);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(sweep_err_w_self_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  sweep_err_w_self_kernel<<<R, 128, smem, (cudaStream_t)stream>>>(W, R, K, cs, cz, P, n_levels, err_sum);
  return check_launch("sweep_err_w_self");
}

int adalog_sweep_err_a_self(const float* x, int64_t n_total, int C, int per_channel, const float* cs,
                            const float* cz, int P, int n_levels, double* partial, int nsplit, void* stream) {
  ADALOG_REQUIRE(x && cs && cz && partial && n_total > 0 && C > 0 && P > 0 && P <= ADALOG_P && nsplit > 0, -1,
                 "sweep_err_a_self: bad arguments (P must be <= 128)");
  const int Cw = per_channel ? C : 32;
  dim3 grid((Cw + 31) / 32, nsplit), block(32, 8);
  sweep_err_a_self_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, n_total, Cw, per_channel, cs, cz, P, n_levels,
                                                                   partial, nsplit);
  return check_launch("sweep_err_a_self");
}

int adalog_gen_uniform_fixed(const float* x, int64_t R, int K, int64_t ldx, const float* scale, const float* zp,
                             int64_t g_div, int64_t g_mod, int n_levels, uint16_t* out, int kpad, float* rowsum,
                             void* stream) {
  ADALOG_REQUIRE(x && scale && zp && out && R > 0 && K > 0 && kpad % ADALOG_BK == 0 && kpad >= K && g_div > 0 &&
                     g_mod > 0, -1, "gen_uniform_fixed: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (rowsum) cudaMemsetAsync(rowsum, 0, (size_t)R * sizeof(float), st);
  gen_uniform_fixed_kernel<<<grid_for(R * (kpad / 8), 256), 256, 0, st>>>(x, R, K, ldx, scale, zp, g_div, g_mod,
                                                                         n_levels, out, kpad, rowsum);
  return check_launch("gen_uniform_fixed");
}

int adalog_gen_uniform_cand(const float* x, int64_t U, int K, int64_t ldx, const float* cs, const float* cz, int P,
                            int64_t pstride, int64_t gstride, int64_t g_div, int64_t g_mod, int64_t u_base,
                            int n_levels, uint16_t* out, int kpad, int krep, float* rowsum, void* stream) {
  ADALOG_REQUIRE(x && cs && cz && out && U > 0 && U < (1ll << 31) && K > 0 && kpad % ADALOG_BK == 0 && kpad >= K &&
                     P > 0 && P <= ADALOG_P && (krep == 1 || krep == 3) && g_div > 0 && g_mod > 0, -1,
                 "gen_uniform_cand: bad arguments");
  const int cpr = kpad / 8;
  const int tpc = cpr < 256 ? cpr : 256;                 // threads along K (8 elements each)
  const int npg = 256 / tpc;                             // candidate groups per CTA
  dim3 grid((unsigned)U, (unsigned)cand_split(U, npg));
  gen_uniform_cand_kernel<<<grid, tpc * npg, 0, (cudaStream_t)stream>>>(x, K, ldx, cs, cz, P, pstride, gstride,
                                                                              g_div, g_mod, u_base, n_levels, out,
                                                                              kpad, krep, rowsum, tpc);
  return check_launch("gen_uniform_cand");
}

int adalog_gen_log_cand(const float* x, int64_t U, int K, int64_t ldx, const float* cs, const long long* cq, int P,
                        const float* shift, const float* mtab, int n_levels, uint16_t* out, int kpad, void* stream) {
  ADALOG_REQUIRE(x && cq && mtab && out && U > 0 && U < (1ll << 31) && K > 0 && kpad % ADALOG_BK == 0 && kpad >= K &&
                     P > 0 && P <= ADALOG_P, -1, "gen_log_cand: bad arguments");
  const int cpr = kpad / 8;
  const int tpc = cpr < 256 ? cpr : 256;
  const int npg = 256 / tpc;
  dim3 grid((unsigned)U, (unsigned)cand_split(U, npg));
  gen_log_cand_kernel<<<grid, tpc * npg, 0, (cudaStream_t)stream>>>(x, K, ldx, cs, cq, P, shift, mtab, n_levels,
                                                                          out, kpad, tpc);
  return check_launch("gen_log_cand");
}

int adalog_gen_log_fixed(const float* x, int64_t R, int K, int64_t ldx, const float* scale, const long long* q,
                         const float* shift, const float* table1, const float* m2, int n_levels, uint16_t* out,
                         int kpad, void* stream) {
  ADALOG_REQUIRE(x && scale && q && table1 && m2 && out && R > 0 && K > 0 && kpad % ADALOG_BK == 0 && kpad >= K &&
                     n_levels <= 128, -1, "gen_log_fixed: bad arguments");
  gen_log_fixed_kernel<<<grid_for(R * (kpad / 8), 256), 256, 0, (cudaStream_t)stream>>>(x, R, K, ldx, scale, q, shift,
                                                                                       table1, m2, n_levels, out, kpad);
  return check_launch("gen_log_fixed");
}

int adalog_gen_split3(const float* x, int64_t R, int K, int64_t ldx, uint16_t* out, int kpad, void* stream) {
  ADALOG_REQUIRE(x && out && R > 0 && K > 0 && kpad % ADALOG_BK == 0 && kpad >= K, -1, "gen_split3: bad arguments");
  gen_split3_kernel<<<grid_for(R * (kpad / 8), 256), 256, 0, (cudaStream_t)stream>>>(x, R, K, ldx, out, kpad);
  return check_launch("gen_split3");
}

}  // extern "C"
