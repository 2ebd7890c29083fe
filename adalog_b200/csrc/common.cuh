// Shared host/device helpers for libadalog_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace adalog {

// thread-local last-error buffer (returned by adalog_last_error)
char* err_buf();
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);

#define ADALOG_REQUIRE(cond, code, ...) \
  do { if (!(cond)) return ::adalog::fail(code, __VA_ARGS__); } while (0)

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// asymmetric uniform quantisation to the integer (code - zp); bit-exact w.r.t. torch:
// ((x / s).round() + zp).clamp(0, L) - zp     (linear.py:304-305, uniform.py:29,34-35)
__device__ __forceinline__ float uq_int(float x, float s, float z, float L) {
  float r = rintf(__fdiv_rn(x, s));
  float c = fminf(fmaxf(__fadd_rn(r, z), 0.0f), L);
  return __fsub_rn(c, z);
}

}  // namespace adalog
