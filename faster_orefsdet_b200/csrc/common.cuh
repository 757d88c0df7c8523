// Shared helpers for the fod_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/fod_b200.h"

namespace fod {

constexpr int kC = FOD_CHANNELS;  // 128 channels everywhere on this path

void set_error(const char* fmt, ...);

#define FOD_REQUIRE(cond, ...)      \
  do {                              \
    if (!(cond)) {                  \
      fod::set_error(__VA_ARGS__);  \
      return FOD_ERR_BAD_ARG;       \
    }                               \
  } while (0)

#define FOD_CUDA_LAUNCH_CHECK(name)                                                   \
  do {                                                                                \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess) {                                                         \
      fod::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));         \
      return FOD_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)

#define FOD_CUDA_CALL(expr)                                                           \
  do {                                                                                \
    cudaError_t e__ = (expr);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      fod::set_error("%s failed: %s", #expr, cudaGetErrorString(e__));                \
      return FOD_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)

inline cudaStream_t as_stream(fod_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Order-preserving map float -> uint32 (ascending float == ascending uint).
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// IoU test with the exact arithmetic of torchvision's CPU nms kernel
// (torchvision/csrc/ops/cpu/nms_kernel.cpp): every operation a separately rounded
// fp32 op (no FMA contraction), the ratio compared against the double threshold.
// thr_f is the smallest float whose double value exceeds the double threshold, so
// `ovr >= thr_f`  <=>  `(double)ovr > thr`.
__device__ __forceinline__ bool iou_exceeds(const float4 a, float area_a, const float4 b, float area_b, float thr_f) {
  float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  float w = fmaxf(0.f, __fsub_rn(xx2, xx1));
  float h = fmaxf(0.f, __fsub_rn(yy2, yy1));
  float inter = __fmul_rn(w, h);
  float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
  // Same decision as the exact quotient below, without the division in the clear cases: for uni > 0 and thr > 0
  // the rounded quotient is >= thr iff inter/uni is, up to half an ulp; a 1e-6 relative margin around thr*uni is
  // hundreds of ulps wide, everything inside it takes the exact path.  (Most pairs do not overlap at all.)
  if (uni > 0.f && thr_f > 0.f) {
    const float t = __fmul_rn(thr_f, uni);
    if (inter > __fmul_rn(t, 1.000001f)) return true;
    if (inter < __fmul_rn(t, 0.999999f)) return false;
  }
  float ovr = __fdiv_rn(inter, uni);
  return ovr >= thr_f;  // false for NaN, like `ovr > thr`
}

__device__ __forceinline__ float box_area(const float4 b) { return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y)); }

float iou_threshold_as_float(double thr);

}  // namespace fod
