// fp32 twin of common.cuh for the generated single-precision kernels (csrc/gen_f32/, written by build.py from the fp64
// sources: double -> float outside the lines tagged "f64", which keep the phase / acceleration arithmetic in fp64).
// The reference itself never runs in fp32 (adept/_base_.py:287-292 switches x64 on); the _f32 entry points are the
// explicit extra SURVEY.md 8b names, held to <= 1e-5 relative L2 against the fp64 oracle.
#pragma once
#include "common.cuh"

namespace adept32 {
using namespace adept;  // error codes, set_last_error, check_launch, ProfileScope, accel_of (fp64)

typedef float2 cplx;

__device__ __forceinline__ cplx cmake(float x, float y) { return make_float2(x, y); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ cplx cconj(cplx a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ cplx cmul_mi(cplx a) { return make_float2(a.y, -a.x); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// collide_core.cuh: reciprocal and the cyclic-reduction cut-off (couplings below 2^-27 cannot change an fp32 solution)
__device__ __forceinline__ float fast_rcp(float x) { return __frcp_rn(x); }
#define PCR_TOL 7.4505806e-9f

// fp32 twiddle tables (f32.cu): same layout as adept::get_twiddles, entries rounded once from long double
const cplx* get_twiddles(int logn);

// transform lengths that are not powers of two are not offered in fp32
inline bool bluestein_supported(int) { return false; }
template <class... A>
inline int bluestein_push_f32(A...) {
  return ADEPT_ERR_UNSUPPORTED;
}

}  // namespace adept32
