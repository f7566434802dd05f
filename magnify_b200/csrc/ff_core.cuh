// Per-pixel flat-field arithmetic shared by the device kernels and the host fuzz test
// (tests/csrc/ff_fuzz.cpp compiles this header with g++).
//
// Reference: src/magnify/preprocess.py:83-87
//     t = clip(float64(x) - dark, 0);  u = t / flat;  v = (u * M) / M2;  out = (dtype) v
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MGB_HD __host__ __device__ __forceinline__
#else
#define MGB_HD inline
#endif

namespace mgb {

// s = (2^20 + x) * gain + bias lands in [2^20, 2^21) when v' = (x - dark) * gain is in
// [-2^19, 2^19): there ulp(s) = 2^-32, so the high word's low 20 bits are floor(v' + G) + 2^19
// and the low word is frac(v' + G) * 2^32.  G = 2^-24 recentres the guard band so that one
// unsigned compare (lo < 2G) catches "v' within G of an integer".
constexpr double kFFOffset = 1572864.0;                 // 2^20 + 2^19
constexpr double kFFGuard = 1.0 / 16777216.0;           // G = 2^-24
constexpr unsigned kFFGuardLo = 512u;                   // 2G in units of 2^-32
constexpr unsigned kFFHiBase = 0x41300000u;             // high word of 2^20
constexpr double kFFMaxGain = 16.0;
constexpr double kFFMaxDark = 1048576.0;

// The reference's operation order, one IEEE rounding per operation.
MGB_HD double ff_exact_value(double x, double flat, double dark, double M, double M2) {
#if defined(__CUDA_ARCH__)
  double t = __dsub_rn(x, dark);
  t = t < 0.0 ? 0.0 : t;
  double u = __ddiv_rn(t, flat);
  double w = __dmul_rn(u, M);
  return __ddiv_rn(w, M2);
#else
  volatile double t = x - dark;
  if (t < 0.0) t = 0.0;
  volatile double u = t / flat;
  volatile double w = u * M;
  volatile double v = w / M2;
  return v;
#endif
}

MGB_HD uint16_t ff_exact_u16(uint16_t x, double flat, double dark, double M, double M2) {
  double v = ff_exact_value((double)x, flat, dark, M, M2);
  // C cast of the reference's astype (preprocess.py:87): truncation; out-of-range values wrap
  // like x86 (only reachable with a negative darkfield).
  return (uint16_t)(long long)v;
}

MGB_HD double ff_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return fma(a, b, c);
#endif
}

MGB_HD double ff_from_hi(unsigned hi_word) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double((int)hi_word, 0);
#else
  union { double d; uint64_t u; } cv;
  cv.u = (uint64_t)hi_word << 32;
  return cv.d;
#endif
}

// Fast-path coefficients for one position.  Coefficients are accepted only if s(x) stays in
// [2^20, 2^21) for the two extreme pixel values (s is monotone in x because gain > 0), so the
// per-pixel code needs no range check.  Rejected positions get gain = 0, bias = 2^20: then
// s = 2^20 exactly, its low word is 0 < 2G and every pixel there takes the exact path.
MGB_HD void ff_make_coeffs(double flat, double dark, double M, double M2, double* gain,
                           double* bias) {
  *gain = 0.0;
  *bias = 1048576.0;
  double k = M / M2;
  double g = (1.0 / flat) * k;
  bool ok = (flat > 0.0) && (M2 > 0.0) && (g > 0.0) && (g <= kFFMaxGain) &&
            (fabs(dark) <= kFFMaxDark) && (k == k);
  if (!ok) return;
  double b = ff_fma(-(dark + 1048576.0), g, kFFOffset + kFFGuard);
  double s_lo = ff_fma(ff_from_hi(kFFHiBase), g, b);             // x = 0
  double s_hi = ff_fma(ff_from_hi(kFFHiBase | 0xffffu), g, b);   // x = 65535
  if (!(s_lo >= 1048576.0) || !(s_hi < 2097152.0)) return;
  *gain = g;
  *bias = b;
}

// One pixel of the fast path.  Returns the candidate output; *slow is OR-ed with "recompute
// exactly".  hi_word = 0x41300000 | x, i.e. the double 2^20 + x with a zero low word.
MGB_HD int ff_fast_px(unsigned hi_word, double gain, double bias, bool* slow) {
  double s = ff_fma(ff_from_hi(hi_word), gain, bias);
#if defined(__CUDA_ARCH__)
  unsigned hi = (unsigned)__double2hiint(s);
  unsigned lo = (unsigned)__double2loint(s);
#else
  union { double d; uint64_t u; } cv;
  cv.d = s;
  unsigned hi = (unsigned)(cv.u >> 32);
  unsigned lo = (unsigned)cv.u;
#endif
  *slow = *slow || (lo < kFFGuardLo);
  int o = (int)(hi - kFFHiBase) - 0x80000;
  return o < 0 ? 0 : o;
}

}  // namespace mgb
