// Shared device/host helpers for libssdhead (sm_100a only).
// The whole library is compiled with -fmad=false: every index decision (match mask, NMS keep list,
// TP flags) is taken on fp32 values computed with exactly the reference's sequence of IEEE add / sub /
// mul / div / min / max, so those results are bit-identical to torch's.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ssdhead.h"

#ifndef __CUDA_ARCH__
#define SSDH_HOST_ONLY 1
#endif

namespace ssdh {

constexpr int kMaxGT = 64;
constexpr int kMaxClasses = 64;

void set_error(const char* fmt, ...);
int cuda_status(const char* what);   // cudaGetLastError -> return code (0 or cudaError_t), sets message
// cudaFuncAttributeMaxDynamicSharedMemorySize is per (kernel, device): set it once for each pair.  0 = OK.
int ensure_dyn_smem(const void* kernel, int bytes, const char* what);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------------------------------------
// Boxes.  The reference keeps boxes in centre form and recomputes the corners inside every IoU
// (src/model/ssd.py:247-248, src/utils.py:74-75).  Corners and area depend on one box only, so we
// compute them once per box with the same fp32 operations: c -/+ (s / 2) and w * h.
// ---------------------------------------------------------------------------------------------
struct Corners {
  float x1, x2, y1, y2, area;
};

__device__ __forceinline__ Corners make_corners(float cx, float cy, float w, float h) {
  Corners c;
  const float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);   // w / 2 is exact either way
  c.x1 = __fsub_rn(cx, hw);
  c.x2 = __fadd_rn(cx, hw);
  c.y1 = __fsub_rn(cy, hh);
  c.y2 = __fadd_rn(cy, hh);
  c.area = __fmul_rn(w, h);
  return c;
}

// torch.min / torch.max / clamp(min=0) on finite inputs.
__device__ __forceinline__ float overlap_1d(float a_lo, float a_hi, float b_lo, float b_hi) {
  return fmaxf(__fsub_rn(fminf(a_hi, b_hi), fmaxf(a_lo, b_lo)), 0.0f);
}

__device__ __forceinline__ float intersection(const Corners& a, const Corners& b) {
  return __fmul_rn(overlap_1d(a.x1, a.x2, b.x1, b.x2), overlap_1d(a.y1, a.y2, b.y1, b.y2));
}

// (area_a + area_b) - inter, in the reference's order (the sum of the two areas is commutative).
__device__ __forceinline__ float union_area(const Corners& a, const Corners& b, float inter) {
  return __fsub_rn(__fadd_rn(a.area, b.area), inter);
}

// Exact value of  fl(inter / uni) > thr  without the division in the common case.
// thr_lo = thr * (1 - 2^-20), thr_hi = thr * (1 + 2^-20) are precomputed by ThrBand below.  For a
// normal positive `uni` the rounded products bracket uni*thr tightly enough that anything outside the
// band is decided with certainty; the (rare) pairs inside it take the IEEE division.
struct ThrBand {
  float thr, lo, hi;
  bool usable;   // thr large enough for the relative-error argument to hold
};

__host__ __device__ inline ThrBand make_band(float thr) {
  ThrBand b;
  b.thr = thr;
  b.lo = thr * (1.0f - 9.5367431640625e-07f);
  b.hi = thr * (1.0f + 9.5367431640625e-07f);
  b.usable = (thr >= 1e-6f) && (thr <= 1e6f);
  return b;
}

__device__ __forceinline__ bool quotient_gt(float inter, float uni, const ThrBand& b) {
  if (b.usable && uni >= 1e-30f) {
    if (inter > __fmul_rn(uni, b.hi)) return true;
    if (inter < __fmul_rn(uni, b.lo)) return false;
  }
  return __fdiv_rn(inter, uni) > b.thr;
}

// src/utils.py:77:  where(w*h > 0, w*h / union, w*h)
__device__ __forceinline__ float iou_value(const Corners& a, const Corners& b) {
  const float inter = intersection(a, b);
  return inter > 0.0f ? __fdiv_rn(inter, union_area(a, b, inter)) : inter;
}
__device__ __forceinline__ bool iou_gt(const Corners& a, const Corners& b, const ThrBand& band) {
  const float inter = intersection(a, b);
  if (!(inter > 0.0f)) return inter > band.thr;
  return quotient_gt(inter, union_area(a, b, inter), band);
}

// ---------------------------------------------------------------------------------------------
// Order-preserving float <-> uint key (handles negatives; CE and scores are >= 0 in practice).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_key(float f) {
  uint32_t u = __float_as_uint(f);
  if (u == 0x80000000u) u = 0;                      // -0.0 sorts with +0.0
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ---------------------------------------------------------------------------------------------
// Warp helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

}  // namespace ssdh
