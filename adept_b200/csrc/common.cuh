// Shared helpers for the adept_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace adept {

typedef double2 cplx;

__device__ __forceinline__ cplx cmake(double x, double y) { return make_double2(x, y); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
// a * (-i)
__device__ __forceinline__ cplx cmul_mi(cplx a) { return make_double2(a.y, -a.x); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// device-wide barrier among the co-resident CTAs of this launch: tickets on a monotonic counter (target = tickets of
// all CTAs up to and including this barrier)
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    while (*reinterpret_cast<volatile unsigned int*>(counter) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

// the same barrier in two halves, so that work that does not depend on the other CTAs can run between them
__device__ __forceinline__ void grid_arrive(unsigned int* counter) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
  }
}
__device__ __forceinline__ void grid_wait(unsigned int* counter, unsigned int target) {
  if (threadIdx.x == 0) {
    while (*reinterpret_cast<volatile unsigned int*>(counter) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

// accel = (q*e + (q^2/m)*pond)/m with the reference's rounding sequence (no FMA contraction):
// adept/_vlasov1d/solvers/pushers/vlasov.py:83-84
__device__ __forceinline__ double accel_of(double e, double pond, double q, double q2m, double m) {
  const double force = __dadd_rn(__dmul_rn(q, e), __dmul_rn(q2m, pond));
  return __ddiv_rn(force, m);
}

// Device-resident time row (adept_b200_step::time_row): when a step descriptor carries one, the kernels take the O(1)
// time factors of the drivers and collision profiles from it instead of the by-value fields, so that a captured CUDA
// graph can be replayed for later steps.  Layout (doubles): tenv[s][d] at 8 s + d, wt[s][d] at 48 + 8 s + d, nu_fp_time
// at 96, nu_K_time at 97, ex_t[s] at 98 + s.
enum { TROW_TENV = 0, TROW_WT = 48, TROW_NU_FP = 96, TROW_NU_K = 97, TROW_EX_T = 98, TROW_LEN = 104 };
// the row of the step being enqueued by this host thread (step.cu sets it around adept_b200_step_f64); launchers copy
// it into their argument structs
const double* current_time_row();
void set_current_time_row(const double* row);

// error codes returned across the C ABI (include/adept_b200.h)
enum {
  ADEPT_OK = 0,
  ADEPT_ERR_BAD_SHAPE = -1,
  ADEPT_ERR_UNSUPPORTED = -2,
  ADEPT_ERR_CUDA = -3,
  ADEPT_ERR_BAD_ARG = -4,
};

void set_last_error(const char* fmt, ...);
int check_launch(const char* what);

// Optional per-kernel timing (api.cu): when adept_b200_profile(1) is on, every launch site brackets its kernel with
// CUDA events on the launching stream; adept_b200_profile_report() sums them per kernel name.  Off by default.
struct ProfileScope {
  ProfileScope(const char* name, cudaStream_t stream);
  ~ProfileScope();
  int slot;
  cudaStream_t stream;
};

// twiddle-table cache (api.cu): per-pass Stockham tables for a complex FFT of length 2^logn on the current device
const cplx* get_twiddles(int logn);
// Bluestein tables for a transform of any length 2 <= n <= 4096 (api.cu): M = 2^logm >= 2n - 1
int get_bluestein(int n, int* logm_out, const cplx** chirp_out, const cplx** bhat_out);

}  // namespace adept
