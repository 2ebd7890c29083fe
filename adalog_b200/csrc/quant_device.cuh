// Device helpers shared by the operand generators (quant_kernels.cu) and the fused generator + GEMM kernel
// (fused_gemm_err.cu): the exact fast paths of rint(x/s) and of the AdaLog code, and the operand packers.
// Every arithmetic step that must round like the reference's eager kernels uses an explicit _rn intrinsic or fmaf, so
// the results do not depend on the translation unit's -fmad setting.
#pragma once
#include "common.cuh"

namespace adalog {

constexpr float kMagic = 12582912.0f;                 // 1.5 * 2^23
constexpr float kFracSafe = 0.5f - 6.103515625e-05f;  // 0.5 - 2^-14

// ---- packed FP32 (Blackwell FFMA2 / FADD2 / FMUL2): two IEEE-exact lanes per instruction and per FMA-pipe slot
__device__ __forceinline__ void fmul2(float& o0, float& o1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "=f"(o0), "=f"(o1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fadd2(float& o0, float& o1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "=f"(o0), "=f"(o1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void ffma2(float& o0, float& o1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(o0), "=f"(o1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}

// Error-free FP32 accumulation (Knuth two-sum into a hi/lo pair): hi + lo carries the running sum to ~2^-48 relative,
// i.e. as well as an FP64 accumulator for FP32-valued terms, in 7 FP32 adds (the FP64 pipe of this part is narrow:
// ncu showed F2F.F64 + DADD holding 12-15% of an epilogue warp's time when used once per 32-column slab).
struct TwoSumF {
  float hi = 0.0f, lo = 0.0f;
  __device__ __forceinline__ void add(float x) {
    const float s = __fadd_rn(hi, x);
    const float bb = __fsub_rn(s, hi);
    const float err = __fadd_rn(__fsub_rn(hi, __fsub_rn(s, bb)), __fsub_rn(x, bb));
    lo = __fadd_rn(lo, err);
    hi = s;
  }
  __device__ __forceinline__ double value() const { return (double)hi + (double)lo; }
};

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

// Generator form of the same fast path, arranged for the FMA pipe (the generators were ALU-pipe bound on compares
// and min/max).  c = {r/2n, zp/2n, L/2n, 1.5*2^23 - zp} with r = fl(1/s), L = 2n-1; 2n is a power of two, so
//   ts = sat(fma(x, r/2n, zp/2n))          == clamp(fl(x*r + zp), 0, 2n) / 2n      (FFMA.SAT: the clamp is free)
//   ts = min(ts, L/2n)                                                               (the one FMNMX left)
//   tm = fma(ts, 2n, 1.5*2^23 - zp)        == rint(clamped) - zp + 1.5*2^23         (code - zp in the low mantissa bits)
//   d  = fma(ts, 2n, -(tm - (1.5*2^23 - zp))) == clamped - rint(clamped)
// fl(x*r + zp) is within 0.5 ulp(256) + |x/s| 2^-24 <= 2.3e-5 of x/s + zp and the reference's fl(x/s) within 1.6e-5 of
// x/s (0 <= zp <= L <= 255, so |x/s| <= 255 wherever the clamp does not decide), so with |d| <= 0.5 - 2^-14 both round
// to the same integer; where the clamp decides d == 0.  Candidates with a zero point that is not an integer in [0, L]
// carry thr < 0 and always take the IEEE path; +-inf saturate like the reference's clamp; NaN (FFMA.SAT returns 0 for
// it) is caught per chunk by the caller.
__device__ __forceinline__ float uq_code_fast(float x, const float4 c, float two_n, float thr, bool& unsafe) {
  float ts = __saturatef(fmaf(x, c.x, c.y));
  ts = fminf(ts, c.z);
  const float tm = fmaf(ts, two_n, c.w);
  const float d = fmaf(ts, two_n, -__fsub_rn(tm, c.w));
  unsafe |= !(fabsf(d) <= thr);
  return tm;                               // (code - zp) + 1.5*2^23
}

// four small-integer floats -> four int8 (two's complement) in one word: the low mantissa byte of v + 1.5*2^23
__device__ __forceinline__ uint32_t pack_i8x4(float a, float b, float c, float d) {
  const uint32_t ua = __float_as_uint(__fadd_rn(a, 12582912.0f)), ub = __float_as_uint(__fadd_rn(b, 12582912.0f));
  const uint32_t uc = __float_as_uint(__fadd_rn(c, 12582912.0f)), ud = __float_as_uint(__fadd_rn(d, 12582912.0f));
  return __byte_perm(__byte_perm(ua, ub, 0x0040), __byte_perm(uc, ud, 0x0040), 0x5410);
}

// same, for values that already carry the 1.5*2^23 offset
__device__ __forceinline__ uint32_t pack_i8x4_bits(float a, float b, float c, float d) {
  return __byte_perm(__byte_perm(__float_as_uint(a), __float_as_uint(b), 0x0040),
                     __byte_perm(__float_as_uint(c), __float_as_uint(d), 0x0040), 0x5410);
}

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


}  // namespace adalog
