// sin / cos used by every kernel (generic and model-specialised): split out of rbd_device.cuh so that the NVRTC translation
// units of rbd_jit.cpp can include exactly the same code.
#pragma once
#if !defined(RBD_HD)
#if defined(__CUDACC__)
#define RBD_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#include <cstring>
#define RBD_HD inline
#endif
#endif

namespace rbd {

// fp32 sin/cos in ~25 instructions: 3-term Cody-Waite reduction by pi/2 (exact with FMA for |x| <= 1e4), degree-7 / degree-8
// minimax polynomials on [-pi/4, pi/4], quadrant fix-up.  Max abs error 9.2e-8 on |x| <= 1e4 (tools/check_sincos.py), the
// same class as CUDA's sincosf (2 ulp) at a third of its instruction count; joint angles beyond 1e4 rad take the libm path.
RBD_HD float fma_f(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}
RBD_HD int float_bits(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(x);
#else
  int i; std::memcpy(&i, &x, sizeof(i)); return i;
#endif
}
// Fast path: accurate for |x| <= 1e4 (see above); callers are responsible for larger / non-finite arguments.
RBD_HD void sincos_fast(float x, float& s, float& c) {
  const float t = fma_f(x, 0.6366197466850281f, 12582912.0f);   // 1.5 * 2^23: low mantissa bits = round(x * 2/pi)
  const int n = float_bits(t);
  const float j = t - 12582912.0f;
  float r = fma_f(-j, 1.5707963705062866f, x);
  r = fma_f(-j, -4.371138828673793e-08f, r);
  r = fma_f(-j, -1.7151245100e-15f, r);
  const float z = r * r;
  float sp = fma_f(-1.9515295891e-4f, z, 8.3321608736e-3f);
  sp = fma_f(sp, z, -1.6666654611e-1f);
  const float sr = fma_f(sp * z, r, r);
  float cp = fma_f(2.443315711809948e-5f, z, -1.388731625493765e-3f);
  cp = fma_f(cp, z, 4.166664568298827e-2f);
  const float cr = fma_f(cp * z, z, fma_f(-0.5f, z, 1.0f));
  const float ss = (n & 1) ? cr : sr;
  const float cc = (n & 1) ? sr : cr;
  s = (n & 2) ? -ss : ss;
  c = ((n + 1) & 2) ? -cc : cc;
}
// huge / non-finite angles: library path, kept OUT of line (~150 instructions of Payne-Hanek reduction per call site otherwise)
#if defined(__CUDACC__)
__device__ __noinline__ float2 sincos_slow(float x) { float2 r; sincosf(x, &r.x, &r.y); return r; }
#endif
RBD_HD void sincos_t(float x, float& s, float& c) {
  if (!(x >= -1.0e4f && x <= 1.0e4f)) {
#if defined(__CUDA_ARCH__)
    const float2 r = sincos_slow(x);      // by value: s and c stay in registers at the call site
    s = r.x; c = r.y;
#else
    s = std::sin(x); c = std::cos(x);
#endif
    return;
  }
  sincos_fast(x, s, c);
}
RBD_HD void sincos_t(double x, double& s, double& c) {
#if defined(__CUDA_ARCH__)
  sincos(x, &s, &c);
#else
  s = std::sin(x); c = std::cos(x);
#endif
}

}  // namespace rbd
