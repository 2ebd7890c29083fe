// Elementwise fake-quant forwards, self-error sweeps and bf16 operand generators.
// Compiled with -fmad=false: every multiply/add below rounds exactly like the separate eager torch
// kernels of the reference (no FMA contraction), which is what makes the integer codes bit-exact.
// HBM-bound kernels: 128-bit loads/stores, grid sized in multiples of the SM count.
#include "common.cuh"
#include "quant_device.cuh"
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

// VEC4: n % 4 == 0, pointers 16B aligned and (ngroups == 1 or inner % 4 == 0) so a float4 never straddles groups.
// Each thread keeps UNR independent 16-byte loads in flight (memory-level parallelism is what this HBM-bound kernel
// needs); IdxT = uint32_t whenever n < 2^31 so the per-vector group index costs a 32-bit divide.
template <bool VEC4, typename IdxT>
__global__ void __launch_bounds__(256) uniform_fakequant_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                               int16_t* __restrict__ codes, int64_t n,
                                                               const float* __restrict__ scale,
                                                               const float* __restrict__ zp, int64_t inner,
                                                               int64_t ngroups, int nl, int sym) {
  const bool single = ngroups == 1;
  const float s0 = scale[0], z0 = (zp != nullptr) ? zp[0] : 0.0f;
  const float two_n = (float)(2 * nl), inv2n = 1.0f / two_n, L = (float)(2 * nl - 1);
  const bool pow2 = !sym && nl > 0 && (nl & (nl - 1)) == 0;     // r/2n etc. are exact scalings only then
  if (VEC4) {
    constexpr int UNR = 4;
    const IdxT n4 = (IdxT)(n >> 2);
    const IdxT inner4 = (IdxT)(inner >> 2), ng = (IdxT)ngroups;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const IdxT stride = (IdxT)gridDim.x * blockDim.x;
    // software pipelined: the loads of iteration k+1 are in flight while iteration k is computed, so a thread always
    // has UNR 16-byte loads outstanding (with the loads issued only between compute phases the kernel sat at 75% of
    // the bandwidth a plain copy kernel reaches: not enough bytes in flight per SM)
    auto load_set = [&](IdxT base, float4 (&v)[UNR], float (&sv)[UNR], float (&zv)[UNR]) {
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const IdxT i = base + (IdxT)u * stride;
        sv[u] = s0; zv[u] = z0;
        if (i < n4) {
          v[u] = __ldg(x4 + i);
          if (!single) {
            const IdxT g = (i / inner4) % ng;
            sv[u] = __ldg(scale + g);
            zv[u] = zp ? __ldg(zp + g) : 0.0f;
          }
        }
      }
    };
    float4 v[UNR], vn[UNR];
    float sv[UNR], zv[UNR], svn[UNR], zvn[UNR];
    const IdxT base0 = (IdxT)blockIdx.x * blockDim.x + threadIdx.x;
    if (base0 < n4) load_set(base0, v, sv, zv);
    for (IdxT base = base0; base < n4; base += stride * UNR) {
      const IdxT nbase = base + stride * UNR;
      if (nbase < n4 && nbase > base) load_set(nbase, vn, svn, zvn);
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const IdxT i = base + (IdxT)u * stride;
        if (i >= n4) break;
        const float s = sv[u], z = zv[u];
        float4 o; float c0, c1, c2, c3;
        // exact fast path of the generators (quant_device.cuh: uq_code_fast), four elements branch-free, one rare
        // IEEE redo per float4: r = fl(1/s), clamp by the FMA's saturation, rint by the 1.5*2^23 trick
        const float r = __fdiv_rn(1.0f, s);
        const bool fast = pow2 && z == rintf(z) && z >= 0.0f && z <= L && fabsf(r) <= 3.0e38f;
        const float4 cc = make_float4(r * inv2n, z * inv2n, L * inv2n, kMagic);
        bool unsafe = !fast;
        const float t0 = uq_code_fast(v[u].x, cc, two_n, kFracSafe, unsafe);
        const float t1 = uq_code_fast(v[u].y, cc, two_n, kFracSafe, unsafe);
        const float t2 = uq_code_fast(v[u].z, cc, two_n, kFracSafe, unsafe);
        const float t3 = uq_code_fast(v[u].w, cc, two_n, kFracSafe, unsafe);
        unsafe |= !(v[u].x == v[u].x && v[u].y == v[u].y && v[u].z == v[u].z && v[u].w == v[u].w);   // NaN
        if (!unsafe) {
          c0 = __fsub_rn(t0, kMagic); c1 = __fsub_rn(t1, kMagic); c2 = __fsub_rn(t2, kMagic); c3 = __fsub_rn(t3, kMagic);
          o.x = __fmul_rn(__fsub_rn(c0, z), s); o.y = __fmul_rn(__fsub_rn(c1, z), s);
          o.z = __fmul_rn(__fsub_rn(c2, z), s); o.w = __fmul_rn(__fsub_rn(c3, z), s);
        } else {
          uq_fwd(v[u].x, s, z, nl, sym, o.x, c0);
          uq_fwd(v[u].y, s, z, nl, sym, o.y, c1);
          uq_fwd(v[u].z, s, z, nl, sym, o.z, c2);
          uq_fwd(v[u].w, s, z, nl, sym, o.w, c3);
        }
        if (y) reinterpret_cast<float4*>(y)[i] = o;
        if (codes) reinterpret_cast<short4*>(codes)[i] = make_short4((short)c0, (short)c1, (short)c2, (short)c3);
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) { v[u] = vn[u]; sv[u] = svn[u]; zv[u] = zvn[u]; }
    }
  } else {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
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
// Reference chain (logarithm.py:87-99):  v = clamp(fl((x+shift)/s), 1e-15, 1);  T = fl(fl(-log2f(v)*m1)/m2) with
// (m1, m2) = (1,1) log2, (2,1) logsqrt2, (37,q) adalog;  c = rint(T);  y = fl(fl(2^-e(c) * f(c)) * s) * [c < 2n] - shift.
// The dequantised value depends only on the code, so it is tabulated once per CTA (lutv[c], lutv[2n] = 0) with the
// reference's exact operation order.  EXACT FAST PATH for the code: r = fl(1/s), v' = clamp(x'*r) (within 2 ulp of
// v), L' = -__log2f(v') (MUFU.LG2: abs error <= 2^-22 on [0.5,2], <= 2 ulp elsewhere), t = L' * fl(m1/m2).  Then
//   |t - T| <= 2.2e-6 + 6.6e-7 * t   (v' vs v: 3.4e-7*k; LG2: k*2.4e-7 + 2.4e-7 t; log2f 1 ulp: 1.2e-7 t; roundings 3e-7 t)
// and the check uses m(t) = 5e-6 + 1.5e-6*t: if t is farther than m(t) from every half-integer, rint(t) == c.
// Otherwise (probability ~1e-4) the element is recomputed with the IEEE chain.  Codes >= 2n read lutv[2n] = 0.
__device__ __forceinline__ float log_code_exact(float xv, float s, int kind, float qf) {
  const float v = fminf(fmaxf(__fdiv_rn(xv, s), 1e-15f), 1.0f);
  const float nlg = -log2f(v);
  if (kind == 0) return rintf(nlg);
  if (kind == 1) return rintf(__fmul_rn(nlg, 2.0f));
  return rintf(__fdiv_rn(__fmul_rn(nlg, 37.0f), qf));
}

template <bool VEC4>
__global__ void __launch_bounds__(256) log_fakequant_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                           int16_t* __restrict__ codes, int64_t n,
                                                           const float* __restrict__ scale, int kind, int nl,
                                                           const long long* __restrict__ q,
                                                           const float* __restrict__ table1,
                                                           const float* __restrict__ table2,
                                                           const float* __restrict__ shift, int sub_shift,
                                                           uint32_t magic4) {
  __shared__ float lutv[260];
  const int ncode = 2 * nl;
  const float s = scale[0];
  const float sh = shift ? shift[0] : 0.0f;
  const float qf = (kind == 2) ? (float)q[0] : 1.0f;
  const float kSqrt2m1 = (float)(1.4142135623730951 - 1.0);  // math.sqrt(2) - 1 cast to FP32 by torch
  for (int c = threadIdx.x; c <= ncode && c < 260; c += blockDim.x) {
    float d = 0.0f;
    if (c < ncode) {
      if (kind == 0) {
        d = __fmul_rn(ldexpf(1.0f, -c), s);
      } else if (kind == 1) {
        const float cf = (float)c;
        const float odd = __fadd_rn(__fmul_rn(fmodf(cf, 2.0f), kSqrt2m1), 1.0f);
        d = __fmul_rn(__fmul_rn(ldexpf(1.0f, -(int)ceilf(__fdiv_rn(cf, 2.0f))), odd), s);
      } else {
        d = __fmul_rn(__fmul_rn(ldexpf(1.0f, -(int)table1[c]), table2[c]), s);
      }
    }
    lutv[c] = d;                                             // entry 2n: masked codes
  }
  __syncthreads();
  const float r = __fdiv_rn(1.0f, s);
  const float kq = (kind == 0) ? 1.0f : (kind == 1 ? 2.0f : __fdiv_rn(37.0f, qf));
  const float ncf = (float)ncode;
  // fast part: t = -lg2(clamp((x+shift)*r, 1e-15, 1)) * k (>= 0, so only the upper clamp is needed), tm = t + 1.5*2^23
  // (the code in the low mantissa bits), `unsafe` when t is within m(t) of a half-integer or NaN
  auto fast = [&](float xin, float& xv, bool& unsafe) -> float {
    xv = shift ? __fadd_rn(xin, sh) : xin;
    const float vp = fminf(fmaxf(__fmul_rn(xv, r), 1e-15f), 1.0f);
    const float t = fminf(__fmul_rn(-__log2f(vp), kq), ncf);
    const float tm = __fadd_rn(t, kMagic);
    const float f = fabsf(__fsub_rn(t, __fsub_rn(tm, kMagic)));
    unsafe |= !(f <= fmaf(-1.5e-6f, t, 0.5f - 5e-6f));
    return tm;
  };
  // IEEE chain for one element; returns the LUT index as tm
  auto exact = [&](float xv) -> float {
    float c = log_code_exact(xv, s, kind, qf);
    c = (c < ncf) ? fmaxf(c, 0.0f) : ncf;                      // >= 2n (incl. +inf, NaN): masked entry
    if (!(c >= 0.0f)) c = ncf;
    return __fadd_rn(c, kMagic);
  };
  const uint32_t lut_bias = (uint32_t)__cvta_generic_to_shared(lutv) - magic4;
  auto finish = [&](float tm, float& yo, float& co) {
    float d;
    asm("ld.shared.f32 %0, [%1];" : "=f"(d) : "r"(__float_as_uint(tm) * 4u + lut_bias));
    if (sub_shift) d = __fsub_rn(d, sh);
    yo = d;
    co = fminf(__fsub_rn(tm, kMagic), ncf - 1.0f);             // the reference clamps the stored code to 2n-1
  };
  auto one = [&](float xin, float& yo, float& co) {
    float xv; bool unsafe = false;
    float tm = fast(xin, xv, unsafe);
    if (unsafe) tm = exact(xv);
    finish(tm, yo, co);
  };
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (VEC4) {
    constexpr int UNR = 4;
    const int64_t n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    // software pipelined like uniform_fakequant_kernel: iteration k+1's loads are in flight during iteration k's math
    float4 v[UNR], vn[UNR];
    const int64_t base0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int u = 0; u < UNR; ++u) if (base0 + u * stride < n4) v[u] = __ldg(x4 + base0 + u * stride);
    for (int64_t base = base0; base < n4; base += stride * UNR) {
      const int64_t nbase = base + stride * UNR;
#pragma unroll
      for (int u = 0; u < UNR; ++u) if (nbase + u * stride < n4) vn[u] = __ldg(x4 + nbase + u * stride);
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int64_t i = base + u * stride;
        if (i >= n4) break;
        float4 o; float c0, c1, c2, c3;
        // four elements branch-free (their dependency chains interleave), one rare IEEE redo per float4
        float x0, x1, x2, x3; bool unsafe = false;
        float t0 = fast(v[u].x, x0, unsafe), t1 = fast(v[u].y, x1, unsafe);
        float t2 = fast(v[u].z, x2, unsafe), t3 = fast(v[u].w, x3, unsafe);
        if (unsafe) { t0 = exact(x0); t1 = exact(x1); t2 = exact(x2); t3 = exact(x3); }
        finish(t0, o.x, c0); finish(t1, o.y, c1); finish(t2, o.z, c2); finish(t3, o.w, c3);
        if (y) reinterpret_cast<float4*>(y)[i] = o;
        if (codes) reinterpret_cast<short4*>(codes)[i] = make_short4((short)c0, (short)c1, (short)c2, (short)c3);
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) v[u] = vn[u];
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      float o, c;
      one(__ldg(x + i), o, c);
      if (y) y[i] = o;
      if (codes) codes[i] = (int16_t)c;
    }
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
  // the element error d is the reference's FP32 value; its square and the sum over the row are taken in FP64 (the
  // kernel is tiny: out x in elements per evaluation), so this sweep sits at the exact value and every difference
  // from the reference's FP32 mean is the reference's own summation noise (tests/gpu_parity.py)
  double acc = 0.0;
  for (int k = 0; k < K; ++k) {
    const float w = wrow[k];
    const double d = (double)__fsub_rn(w, __fmul_rn(uq_int(w, s, z, L), s));
    acc = fma(d, d, acc);
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
  // Per thread: one column, 16 candidates handled as 8 PAIRS with packed FP32 (FFMA2 / FADD2 / FMUL2: two IEEE-exact
  // lanes per instruction, i.e. per FMA-pipe slot -- the kernel is bound by that pipe, 8 -> 4.5 slots per (element,
  // candidate)).  The running sums are error-free FP32 hi/lo pairs (two-sum) in shared memory ([candidate][thread]:
  // conflict free), updated once per 32 rows; FP64 instructions are kept out of the loop (narrow FP64 pipe).
  __shared__ float2 acc_s[16][256];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int p0 = threadIdx.y * 16;
  const bool col_ok = c < Cw;
  const float L = (float)(2 * nl - 1);
  const float two_n = (float)(2 * nl), inv2n = 1.0f / two_n;
  float cx[16], cy[16], cw[16], s[16];
  float a32[16];
  bool all_fast = nl > 0 && (nl & (nl - 1)) == 0;          // r/2n etc. are exact scalings only for a power of two
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int p = min(p0 + j, P - 1);
    const int64_t ci = per_channel ? ((int64_t)min(c, Cw - 1) * P + p) : p;
    s[j] = __ldg(cs + ci);
    const float z = __ldg(cz + ci);
    const float r = __fdiv_rn(1.0f, s[j]);
    all_fast = all_fast && z == rintf(z) && z >= 0.0f && z <= L && fabsf(r) <= 3.0e38f;
    cx[j] = r * inv2n; cy[j] = z * inv2n; cw[j] = __fsub_rn(kMagic, z);
    a32[j] = 0.0f;
    acc_s[j][tid] = make_float2(0.0f, 0.0f);
  }
  const float thr = all_fast ? kFracSafe : -1.0f;          // a thread with any irregular candidate always takes the IEEE path
  const float Lq = L * inv2n;
  const int64_t M = (n_total + Cw - 1) / Cw;
  // rows per split: a multiple of 32, so that the 32-row groups whose FP32 partial is folded into the hi/lo sums sit at
  // absolute multiples of 32 rows -- the FP32 roundings are then the same however the rows are split over CTAs or
  // sharded over GPUs (shards of a multiple of 32 rows)
  const int64_t rps = (((M + nsplit - 1) / nsplit + 31) / 32) * 32;
  const int64_t m0 = (int64_t)blockIdx.y * rps;
  const int64_t m1 = min(M, m0 + rps);
  constexpr int RU = 4;                                    // rows per iteration: four independent loads in flight
  int cnt = 0;
  auto fold = [&]() {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      TwoSumF t;
      const float2 v = acc_s[j][tid];
      t.hi = v.x; t.lo = v.y;
      t.add(a32[j]);
      acc_s[j][tid] = make_float2(t.hi, t.lo);
      a32[j] = 0.0f;
    }
  };
  for (int64_t m = m0; m < m1; m += RU) {
    float xv[RU];
    bool ok[RU];
#pragma unroll
    for (int rr = 0; rr < RU; ++rr) {
      const int64_t idx = (m + rr) * Cw + c;
      ok[rr] = col_ok && m + rr < m1 && idx < n_total;
      xv[rr] = ok[rr] ? __ldg(x + idx) : 0.0f;
    }
#pragma unroll
    for (int rr = 0; rr < RU; ++rr) {
      if (!ok[rr]) continue;
      const float xr = xv[rr];
      const bool x_nan = xr != xr;                         // FFMA.SAT maps NaN to 0: send it down the IEEE path
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        // exact fast path of the generators (quant_device.cuh uq_code_fast, same proof), two candidates per instruction
        float ts0 = fminf(__saturatef(fmaf(xr, cx[j], cy[j])), Lq);
        float ts1 = fminf(__saturatef(fmaf(xr, cx[j + 1], cy[j + 1])), Lq);
        float tm0, tm1, n0, n1, d0, d1, q0, q1;
        ffma2(tm0, tm1, ts0, ts1, two_n, two_n, cw[j], cw[j + 1]);          // (code - zp) + 1.5*2^23
        fadd2(n0, n1, cw[j], cw[j + 1], -tm0, -tm1);                        // -(rounded clamped value)
        ffma2(d0, d1, ts0, ts1, two_n, two_n, n0, n1);                      // clamped - rint(clamped)
        fadd2(q0, q1, tm0, tm1, -kMagic, -kMagic);                          // code - zp
        if (!(fmaxf(fabsf(d0), fabsf(d1)) <= thr) || x_nan) {               // rare: IEEE path for the pair
          q0 = uq_int(xr, s[j], __fsub_rn(kMagic, cw[j]), L);
          q1 = uq_int(xr, s[j + 1], __fsub_rn(kMagic, cw[j + 1]), L);
        }
        float pr0, pr1, e0, e1;
        fmul2(pr0, pr1, q0, q1, s[j], s[j + 1]);
        fadd2(e0, e1, xr, xr, -pr0, -pr1);
        ffma2(a32[j], a32[j + 1], e0, e1, e0, e1, a32[j], a32[j + 1]);
      }
    }
    cnt += RU;
    if (cnt >= 32) { cnt = 0; fold(); }
  }
  fold();
  double acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) { const float2 v = acc_s[j][tid]; acc[j] = (double)v.x + (double)v.y; }
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

template <bool I8>
__global__ void __launch_bounds__(256) gen_uniform_fixed_kernel(const float* __restrict__ x, int64_t R, int K,
                                                               int64_t ldx, const float* __restrict__ scale,
                                                               const float* __restrict__ zp, int64_t g_div,
                                                               int64_t g_mod, int nl, void* __restrict__ out,
                                                               int kpad, float* __restrict__ rowsum) {
  constexpr int EPT = I8 ? 16 : 8;         // elements per 16-byte store
  const int cpr = kpad / EPT;
  const int64_t total = R * cpr;
  const float L = (float)(2 * nl - 1);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / cpr;
    const int kc = (int)(idx - r * cpr) * EPT;
    const int64_t g = (r / g_div) % g_mod;
    const float s = __ldg(scale + g), z = __ldg(zp + g);
    float v[EPT];
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      const int k = kc + j;
      v[j] = (k < K) ? uq_int(__ldg(x + r * ldx + k), s, z, L) : 0.0f;
      sum += v[j];
    }
    uint4 o;
    if (I8) {
      o.x = pack_i8x4(v[0], v[1], v[2], v[3]);   o.y = pack_i8x4(v[4], v[5], v[6], v[7]);
      o.z = pack_i8x4(v[8 % EPT], v[9 % EPT], v[10 % EPT], v[11 % EPT]);
      o.w = pack_i8x4(v[12 % EPT], v[13 % EPT], v[14 % EPT], v[15 % EPT]);
      *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(out) + r * kpad + kc) = o;
    } else {
      o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
      o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
      *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(out) + r * kpad + kc) = o;
    }
    if (rowsum && kc < K) atomicAdd(rowsum + r, sum);  // integers: exact, order independent
  }
}

template <int KREP, bool ROWSUM, bool I8>
__global__ void __launch_bounds__(256, 4) gen_uniform_cand_kernel(const float* __restrict__ x, int K, int64_t ldx,
                                                              const float* __restrict__ cs,
                                                              const float* __restrict__ cz, int P, int64_t pstride,
                                                              int64_t gstride, int64_t g_div, int64_t g_mod,
                                                              int64_t u_base, int nl, void* __restrict__ out_v,
                                                              int kpad, float* __restrict__ rowsum, int tpc) {
  constexpr int EPT = I8 ? 16 : 8;         // elements per thread chunk = one 16-byte store
  constexpr int ESZ = I8 ? 1 : 2;
  uint8_t* out = reinterpret_cast<uint8_t*>(out_v);
  __shared__ float4 cand[ADALOG_P];        // {r/2n, zp/2n, L/2n, 1.5*2^23 - zp}
  __shared__ float2 cand_sz[ADALOG_P];     // {s, zp}: IEEE path
  __shared__ float cthr[ADALOG_P];         // 0.5 - 2^-14, or -1 when the candidate must take the IEEE path
  __shared__ float rsum[ADALOG_P];
  const int64_t u = blockIdx.x;
  const int64_t g = ((u_base + u) / g_div) % g_mod;
  const float L = (float)(2 * nl - 1);
  const float two_n = (float)(2 * nl);
  for (int p = threadIdx.x; p < ADALOG_P; p += blockDim.x) {
    const int pp = min(p, P - 1);          // pad rows repeat the last candidate
    const float s = __ldg(cs + pp * pstride + g * gstride);
    const float z = __ldg(cz + pp * pstride + g * gstride);
    const float r = __fdiv_rn(1.0f, s);
    const bool fast = z == rintf(z) && z >= 0.0f && z <= L && r == r && fabsf(r) <= 3.0e38f;
    cand[p] = make_float4(r / two_n, z / two_n, L / two_n, kMagic - z);
    cand_sz[p] = make_float2(s, z);
    cthr[p] = fast ? kFracSafe : -1.0f;
    rsum[p] = 0.0f;
  }
  __syncthreads();
  const int cpr = kpad / EPT;
  const int npg = blockDim.x / tpc;                      // host guarantees blockDim.x == tpc * npg
  const int lane_chunk = threadIdx.x % tpc, pg = threadIdx.x / tpc;
  const int per = ADALOG_P / gridDim.y;                  // candidates of this CTA: [p_lo, p_lo + per)
  const int p_lo = blockIdx.y * per;
  const int64_t pitch = (int64_t)KREP * kpad * ESZ;      // bytes
  const int64_t dstep = (int64_t)npg * pitch;
  const int iters = (per - pg + npg - 1) / npg;
  const float* xrow = x + u * ldx;
  for (int ch = lane_chunk; ch < cpr; ch += tpc) {
    const int kc = ch * EPT;
    const bool tail = kc + EPT > K;
    float xv[EPT];
    int nan_flag = 0;
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      xv[j] = (kc + j < K) ? __ldg(xrow + kc + j) : 0.0f;
      nan_flag |= (xv[j] != xv[j]) ? 1 : 0;
    }
    asm volatile("" : "+r"(nan_flag));     // one register test per candidate, not EPT re-compares
    uint8_t* dst = out + (u * ADALOG_P + p_lo + pg) * pitch + (int64_t)kc * ESZ;
    int p = p_lo + pg;
    for (int it = 0; it < iters; ++it, dst += dstep, p += npg) {
      const float4 c = cand[p];
      const float thr = cthr[p];
      float tm[EPT], v[EPT], dm[EPT];
      // uq_code_fast (quant_device.cuh) on two elements per packed FP32 instruction; the distances to the rounded
      // values are reduced by a max tree and tested once per (candidate, chunk)
#pragma unroll
      for (int j = 0; j < EPT; j += 2) {
        const float ts0 = fminf(__saturatef(fmaf(xv[j], c.x, c.y)), c.z);
        const float ts1 = fminf(__saturatef(fmaf(xv[j + 1], c.x, c.y)), c.z);
        float n0, n1;
        ffma2(tm[j], tm[j + 1], ts0, ts1, two_n, two_n, c.w, c.w);           // (code - zp) + 1.5*2^23
        fadd2(n0, n1, c.w, c.w, -tm[j], -tm[j + 1]);                         // -(rounded clamped value)
        ffma2(dm[j], dm[j + 1], ts0, ts1, two_n, two_n, n0, n1);             // clamped - rint(clamped)
      }
#pragma unroll
      for (int w2 = EPT / 2; w2 > 0; w2 >>= 1) {
#pragma unroll
        for (int j = 0; j < w2; ++j) dm[j] = fmaxf(fabsf(dm[j]), fabsf(dm[j + w2]));
      }
      const bool unsafe = nan_flag != 0 || !(dm[0] <= thr);
      if (unsafe) {                                      // rare: redo the chunk on the IEEE path
        const float2 sz = cand_sz[p];
#pragma unroll
        for (int j = 0; j < EPT; ++j) tm[j] = __fadd_rn(uq_int(xv[j], sz.x, sz.y, L), kMagic);
      }
      if (tail) {
#pragma unroll
        for (int j = 0; j < EPT; ++j) if (kc + j >= K) tm[j] = kMagic;
      }
      uint4 o;
      if (I8) {                                          // low mantissa byte of (v + 1.5*2^23) = v as int8
        o.x = pack_i8x4_bits(tm[0], tm[1], tm[2], tm[3]);   o.y = pack_i8x4_bits(tm[4], tm[5], tm[6], tm[7]);
        o.z = pack_i8x4_bits(tm[8 % EPT], tm[9 % EPT], tm[10 % EPT], tm[11 % EPT]);
        o.w = pack_i8x4_bits(tm[12 % EPT], tm[13 % EPT], tm[14 % EPT], tm[15 % EPT]);
      }
      if (!I8 || ROWSUM) {
#pragma unroll
        for (int j = 0; j < EPT; ++j) v[j] = __fsub_rn(tm[j], kMagic);
      }
      if (!I8) {
        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
        o.z = pack_bf16x2(v[4 % EPT], v[5 % EPT]); o.w = pack_bf16x2(v[6 % EPT], v[7 % EPT]);
      }
#pragma unroll
      for (int rep = 0; rep < KREP; ++rep) *reinterpret_cast<uint4*>(dst + (int64_t)rep * kpad * ESZ) = o;
      if (ROWSUM) {
        float sum = 0.0f;
#pragma unroll
        for (int j = 0; j < EPT; ++j) sum += v[j];
        atomicAdd(rsum + p_lo + pg + it * npg, sum);
      }
    }
  }
  if (ROWSUM) {
    __syncthreads();
    for (int p = p_lo + threadIdx.x; p < p_lo + per; p += blockDim.x) rowsum[u * ADALOG_P + p] = rsum[p];
  }
}

// ------------------------------------------------------------------------------------------------
// AdaLog search form (linear.py:872-878 / :913-919, matmul.py:337-342): code from (scale_p, q_p), value m*2^-e.
//
// Reference chain per (element, candidate): v = clamp((x+shift)/s, 1e-15, 1); T = fl(fl(-log2f(v)*37)/q);
// c = rint(T).  EXACT FAST PATH: lx = -log2f(x+shift) once per element, ls = -log2f(s), kq = fl(37/q) once per
// candidate, t = fma(lx, kq, -ls*kq).  With log2f accurate to 1 ulp and |lx - ls| <= 49 (outside: IEEE path)
//   |t - T| <= kq*(3.0e-7*|lx| + 3.8e-7*|ls| + 8.6e-8) + 1.8e-7*t        
// and the check uses twice that bound, evaluated per (element, candidate) with one FFMA:
//   margin = kq*(6e-7*|lx| + 2e-7) + [kq*7.6e-7*|ls| + 4e-7*2n];  rint(t) == c whenever |t - rint(t)| <= 0.5 - margin.
// Unscaled form (post-softmax: no division, no clamp): l37 = fl(lx*37) is the reference's own intermediate and
// t = fl(l37 * fl(1/q)) is within 2.5 ulp of T: margin 2^-14 as for the uniform case.
// An 8-element chunk with any element near a rounding boundary (or in the reference's 1e-15 clamp region) is redone
// element by element on the IEEE path.  Supported for n_bits <= 6: the numerators m <= 4n-2 = 126 of the dequantised
// values are then exact in bf16 (8-bit AdaLog would need 9 significant bits; the reference's configs use 3/4/6).
// ------------------------------------------------------------------------------------------------
// The dequantised value of candidate p at code c depends only on (p, c), so each
// CTA tabulates lut[p][c] = mtab[(c*q_p) % 37] * 2^-floor(c*q_p/37) (0 for c = 2n, the masked code) once and the inner
// loop is: t (FFMA) -> clamp to [0, 2n] (2 FMNMX) -> rint + code bits (2 FADD) -> boundary check (FADD, FFMA, FSETP)
// -> one shared load whose address is formed from the magic-number bits.  The profile of the arithmetic version showed
// the ALU pipe (compares / logic / shifts) at 83% with the FMA pipe at 26%; this version is balanced.
template <bool SCALED>
__global__ void __launch_bounds__(256) gen_log_cand_lut_kernel(const float* __restrict__ x, int K, int64_t ldx,
                                                              const float* __restrict__ cs,
                                                              const long long* __restrict__ cq, int P,
                                                              const float* __restrict__ shift,
                                                              const float* __restrict__ mtab, int nl,
                                                              uint16_t* __restrict__ out, int kpad, int tpc, uint32_t magic4,
                                                              int64_t U) {
  extern __shared__ float lut[];           // [per][2n + 1]
  __shared__ float4 cand[ADALOG_P];        // {mul / 2n, off / 2n, lim, mul}
  __shared__ float candq[ADALOG_P];
  __shared__ float chalf[ADALOG_P];        // 0.5 - candidate part of the rounding margin
  __shared__ float cscale[ADALOG_P];
  __shared__ float mt[64];
  __shared__ float lim_min_s;
  const float sh = shift ? shift[0] : 0.0f;
  const int ncode_i = 2 * nl;
  const float ncode = (float)ncode_i;
  const int per = ADALOG_P / gridDim.y;
  const int p_lo = blockIdx.y * per;
  for (int j = threadIdx.x; j < 37; j += blockDim.x) mt[j] = mtab[j];
  for (int i = threadIdx.x; i < per; i += blockDim.x) {
    const int p = p_lo + i;
    const int pp = min(p, P - 1);
    const float qf = (float)cq[pp];
    const float s = SCALED ? __ldg(cs + pp) : 1.0f;
    float mul, off, lim, hp;
    if (SCALED) {
      const float ls = -log2f(s);
      mul = __fdiv_rn(37.0f, qf);
      off = -__fmul_rn(ls, mul);
      lim = __fadd_rn(ls, 49.0f);
      hp = mul * 7.6e-7f * fabsf(ls) + 4e-7f * ncode;
    } else {
      mul = __fdiv_rn(1.0f, qf);
      off = 0.0f;
      lim = __int_as_float(0x7f800000);
      hp = 6.103515625e-05f;
    }
    cand[p] = make_float4(mul / ncode, off / ncode, lim, mul);
    candq[p] = qf;
    chalf[p] = 0.5f - hp;
    cscale[p] = s;
  }
  __syncthreads();
  const int lw = ncode_i + 1;
  for (int i = threadIdx.x; i < per * lw; i += blockDim.x) {
    const int pi = i / lw, c = i - pi * lw;
    float val = 0.0f;
    if (c < ncode_i) {
      const int cqi = c * (int)candq[p_lo + pi];
      const int e = cqi / 37;
      if (e <= 120) val = ldexpf(mt[cqi - e * 37], -e);
    }
    lut[i] = val;
  }
  if (threadIdx.x == 0) {
    float m = __int_as_float(0x7f800000);
    for (int i = 0; i < per; ++i) m = fminf(m, cand[p_lo + i].z);
    lim_min_s = m;
  }
  __syncthreads();
  const float lim_min = lim_min_s;
  const int cpr = kpad >> 3;
  const int npg = blockDim.x / tpc;
  const int lane_chunk = threadIdx.x % tpc, pg = threadIdx.x / tpc;
  const int64_t dstep = (int64_t)npg * kpad;
  const int iters = (per - pg + npg - 1) / npg;
  // shared address of lut[0][0] minus the magic-number offset: addr = bits(t + 1.5*2^23) * 4 + lut_bias (mod 2^32)
  // (magic4 = bits(1.5*2^23) * 4 mod 2^32 arrives as a kernel argument: as a literal, ptxas re-splits it out of the
  // row address and spends an extra add per element on it)
  const uint32_t lut_bias = (uint32_t)__cvta_generic_to_shared(lut) - magic4;
  // (grid-stride over units; the host launches one CTA per unit, see adalog_gen_log_cand)
  for (int64_t u = blockIdx.x; u < U; u += gridDim.x) {
  const float* xrow = x + u * ldx;
  for (int ch = lane_chunk; ch < cpr; ch += tpc) {
    const int kc = ch << 3;
    const bool tail = kc + 8 > K;
    float xs[8], lx[8], e1[8];
    // element part of the rounding margin, 6e-7 |lx| + 2e-7: the chunk's LARGEST is used for all of its elements
    // (conservative: a few more chunks take the IEEE path, still ~1e-4 of them), so the check is ONE threshold per
    // (candidate, chunk) against the max of |d| instead of an FFMA + FSETP per element; +inf (an element near the
    // reference's 1e-15 clamp, x <= 0 or NaN) routes the whole chunk to the IEEE path
    float gmax = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = (kc + j < K) ? __ldg(xrow + kc + j) : 1.0f;
      if (shift) v = __fadd_rn(v, sh);
      xs[j] = v;
      lx[j] = -log2f(v);
      e1[j] = SCALED ? lx[j] : __fmul_rn(lx[j], 37.0f);
      gmax = fmaxf(gmax, !(lx[j] <= lim_min) ? __int_as_float(0x7f800000) : fabsf(lx[j]));
    }
    gmax = SCALED ? fmaf(6e-7f, gmax, 2e-7f) : (gmax <= 3.0e38f ? 0.0f : gmax);
    asm volatile("" : "+f"(gmax));       // keep it in a register: the compiler otherwise re-derives it per candidate
    uint16_t* dst = out + (u * ADALOG_P + p_lo + pg) * (int64_t)kpad + kc;
    int p = p_lo + pg;
    for (int it = 0; it < iters; ++it, dst += dstep, p += npg) {
      const float4 c = cand[p];          // {mul / 2n, off / 2n, lim, mul}: 2n is a power of two, the scaling is exact
      const float lim_d = fmaf(-gmax, c.w, chalf[p]);
      uint32_t row = lut_bias + (uint32_t)((p - p_lo) * lw) * 4u;
      asm volatile("" : "+r"(row));      // one add per element below, not a re-derivation of the row address
      float v[8], dm[8];
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        // t/2n clamped to [0,1] by the FMA's own saturation (codes >= 2n all read lut[2n] = 0), then
        // tm = t + 1.5*2^23 and d = t - rint(t), two elements per packed FP32 instruction (same IEEE results per lane)
        const float ts0 = __saturatef(fmaf(e1[j], c.x, c.y));
        const float ts1 = __saturatef(fmaf(e1[j + 1], c.x, c.y));
        float tm0, tm1, n0, n1;
        ffma2(tm0, tm1, ts0, ts1, ncode, ncode, kMagic, kMagic);
        fadd2(n0, n1, kMagic, kMagic, -tm0, -tm1);
        ffma2(dm[j], dm[j + 1], ts0, ts1, ncode, ncode, n0, n1);
        float val0, val1;
        asm("ld.shared.f32 %0, [%1];" : "=f"(val0) : "r"(__float_as_uint(tm0) * 4u + row));
        asm("ld.shared.f32 %0, [%1];" : "=f"(val1) : "r"(__float_as_uint(tm1) * 4u + row));
        v[j] = val0; v[j + 1] = val1;
      }
#pragma unroll
      for (int w2 = 4; w2 > 0; w2 >>= 1) {
#pragma unroll
        for (int j = 0; j < w2; ++j) dm[j] = fmaxf(fabsf(dm[j]), fabsf(dm[j + w2]));
      }
      const bool unsafe = !(dm[0] <= lim_d);
      if (unsafe) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = log_value_slow(xs[j], lx[j], SCALED, cscale[p], candq[p], mt, ncode);
      }
      if (tail) {
#pragma unroll
        for (int j = 0; j < 8; ++j) if (kc + j >= K) v[j] = 0.0f;
      }
      store8(dst, v);
    }
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
    // K order [low | mid | high]: the MMA walks K upwards and the tensor core's FP32 accumulator TRUNCATES addends
    // that fall below its last bit.  Smallest pieces first, the accumulator is still small when they arrive and
    // they are kept (measured: high-first, the 768 + 768 late small addends each lost up to an ulp of the large
    // running sum -- 1.3e-5 relative on the patch-embedding scores against 3e-7 for the reference's own FP32)
    uint16_t* dst = out + r * (3 * (int64_t)kpad) + kc;
    store8(dst, l);
    store8(dst + kpad, m);
    store8(dst + 2 * (int64_t)kpad, h);
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

// grid of a grid-stride streaming kernel: exactly the number of CTAs that are resident at once (occupancy x 148), so all
// CTAs run in ONE wave and do equal shares.  (148 x 8 CTAs of 256 threads at 48 registers were 1.6 waves: the second,
// partly filled wave cost ~20% of the HBM bandwidth.)
template <typename K>
static int resident_grid(K kernel, int64_t work_items, int threads, size_t smem = 0) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem) != cudaSuccess || occ < 1) occ = 4;
  const int64_t need = (work_items + threads - 1) / threads;
  const int64_t cap = (int64_t)kNumSMs * occ;
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
  if (vec && n < (1ll << 31))
    uniform_fakequant_kernel<true, uint32_t><<<resident_grid(uniform_fakequant_kernel<true, uint32_t>, n / 16, 256), 256, 0, st>>>(x, y, codes, n, scale, zp, inner,
                                                                                    ngroups, n_levels, symmetric);
  else if (vec)
    uniform_fakequant_kernel<true, uint64_t><<<resident_grid(uniform_fakequant_kernel<true, uint64_t>, n / 16, 256), 256, 0, st>>>(x, y, codes, n, scale, zp, inner,
                                                                                    ngroups, n_levels, symmetric);
  else
    uniform_fakequant_kernel<false, uint64_t><<<grid_for(n, 256), 256, 0, st>>>(x, y, codes, n, scale, zp, inner,
                                                                                ngroups, n_levels, symmetric);
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
  const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(codes) & 7) == 0 && (n % 4 == 0);
  if (vec)
    log_fakequant_kernel<true><<<resident_grid(log_fakequant_kernel<true>, n / 16, 256), 256, 0, (cudaStream_t)stream>>>(
        x, y, codes, n, scale, kind, n_levels, q, table1, table2, shift, sub_shift, 0x2D000000u);
  else
    log_fakequant_kernel<false><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
        x, y, codes, n, scale, kind, n_levels, q, table1, table2, shift, sub_shift, 0x2D000000u);
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
                             int64_t g_div, int64_t g_mod, int n_levels, void* out, int kpad, float* rowsum,
                             int dtype, void* stream) {
  const int kb = dtype == ADALOG_I8 ? 128 : ADALOG_BK;
  ADALOG_REQUIRE(x && scale && zp && out && R > 0 && K > 0 && kpad % kb == 0 && kpad >= K && g_div > 0 && g_mod > 0 &&
                     (dtype == ADALOG_BF16 || dtype == ADALOG_I8), -1, "gen_uniform_fixed: bad arguments");
  ADALOG_REQUIRE(dtype != ADALOG_I8 || n_levels <= 64, -2, "gen_uniform_fixed: int8 operands need n_bits <= 7");
  cudaStream_t st = (cudaStream_t)stream;
  if (rowsum) cudaMemsetAsync(rowsum, 0, (size_t)R * sizeof(float), st);
  if (dtype == ADALOG_I8)
    gen_uniform_fixed_kernel<true><<<grid_for(R * (kpad / 16), 256), 256, 0, st>>>(x, R, K, ldx, scale, zp, g_div, g_mod,
                                                                                  n_levels, out, kpad, rowsum);
  else
    gen_uniform_fixed_kernel<false><<<grid_for(R * (kpad / 8), 256), 256, 0, st>>>(x, R, K, ldx, scale, zp, g_div, g_mod,
                                                                                  n_levels, out, kpad, rowsum);
  return check_launch("gen_uniform_fixed");
}

int adalog_gen_uniform_cand(const float* x, int64_t U, int K, int64_t ldx, const float* cs, const float* cz, int P,
                            int64_t pstride, int64_t gstride, int64_t g_div, int64_t g_mod, int64_t u_base,
                            int n_levels, void* out, int kpad, int krep, float* rowsum, int dtype, void* stream) {
  const int kb = dtype == ADALOG_I8 ? 128 : ADALOG_BK;
  ADALOG_REQUIRE(x && cs && cz && out && U > 0 && U < (1ll << 31) && K > 0 && kpad % kb == 0 && kpad >= K && P > 0 &&
                     P <= ADALOG_P && (krep == 1 || krep == 3) && g_div > 0 && g_mod > 0 &&
                     (dtype == ADALOG_BF16 || dtype == ADALOG_I8), -1, "gen_uniform_cand: bad arguments");
  ADALOG_REQUIRE(dtype != ADALOG_I8 || (n_levels <= 64 && krep == 1 && !rowsum), -2,
                 "gen_uniform_cand: int8 operands need n_bits <= 7, krep == 1 and no rowsum");
  const int cpr = kpad / (dtype == ADALOG_I8 ? 16 : 8);
  const int tpc = cpr < 256 ? cpr : 256;                 // threads along K (one 16-byte store each)
  const int npg = 256 / tpc;                             // candidate groups per CTA
  dim3 grid((unsigned)U, (unsigned)cand_split(U, npg));
  cudaStream_t st = (cudaStream_t)stream;
  // max-shared carve-out like the GEMM kernel: CTAs of kernels that want different L1/shared splits do not share an SM,
  // and the generator of chunk i+1 is meant to run in the registers the GEMM of chunk i leaves free
#define ADALOG_LAUNCH_UCAND(KR, RS, I8)                                                                              \
  do {                                                                                                               \
    cudaFuncSetAttribute(gen_uniform_cand_kernel<KR, RS, I8>, cudaFuncAttributePreferredSharedMemoryCarveout,       \
                         cudaSharedmemCarveoutMaxShared);                                                            \
    gen_uniform_cand_kernel<KR, RS, I8><<<grid, tpc * npg, 0, st>>>(x, K, ldx, cs, cz, P, pstride, gstride, g_div,   \
                                                                     g_mod, u_base, n_levels, out, kpad, rowsum,     \
                                                                     tpc);                                           \
  } while (0)
  if (dtype == ADALOG_I8) ADALOG_LAUNCH_UCAND(1, false, true);
  else if (krep == 1) { if (rowsum) ADALOG_LAUNCH_UCAND(1, true, false); else ADALOG_LAUNCH_UCAND(1, false, false); }
  else                { if (rowsum) ADALOG_LAUNCH_UCAND(3, true, false); else ADALOG_LAUNCH_UCAND(3, false, false); }
#undef ADALOG_LAUNCH_UCAND
  return check_launch("gen_uniform_cand");
}

int adalog_gen_log_cand(const float* x, int64_t U, int K, int64_t ldx, const float* cs, const long long* cq, int P,
                        const float* shift, const float* mtab, int n_levels, uint16_t* out, int kpad, void* stream) {
  ADALOG_REQUIRE(x && cq && mtab && out && U > 0 && U < (1ll << 31) && K > 0 && kpad % ADALOG_BK == 0 && kpad >= K &&
                     P > 0 && P <= ADALOG_P, -1, "gen_log_cand: bad arguments");
  const int cpr = kpad / 8;
  const int tpc = cpr < 256 ? cpr : 256;
  const int npg = 256 / tpc;
  // one CTA per unit: capping the grid and looping over units (to build the LUT once per CTA) measured 2x slower --
  // uneven unit counts per CTA and fewer CTAs to fill the SMs beside the GEMM -- so the kernel's unit loop runs once
  dim3 grid((unsigned)U, (unsigned)cand_split(U, npg));
  cudaStream_t st = (cudaStream_t)stream;
  ADALOG_REQUIRE(2 * n_levels <= 64, -2, "gen_log_cand: AdaLog sweeps support n_bits <= 6 (bf16-exact numerators)");
  const size_t lut_bytes = (size_t)(ADALOG_P / grid.y) * (2 * n_levels + 1) * sizeof(float);
  cudaFuncSetAttribute(gen_log_cand_lut_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                       cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(gen_log_cand_lut_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                       cudaSharedmemCarveoutMaxShared);
  if (cs)
    gen_log_cand_lut_kernel<true><<<grid, tpc * npg, lut_bytes, st>>>(x, K, ldx, cs, cq, P, shift, mtab, n_levels,
                                                                      out, kpad, tpc, 0x2D000000u, U);
  else
    gen_log_cand_lut_kernel<false><<<grid, tpc * npg, lut_bytes, st>>>(x, K, ldx, cs, cq, P, shift, mtab, n_levels,
                                                                       out, kpad, tpc, 0x2D000000u, U);
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
