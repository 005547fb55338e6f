// Single-precision entry points (adept_b200_*_f32): host side of the generated fp32 kernels (csrc/gen_f32/, see
// common32.cuh) -- the fp32 twiddle tables and the extern "C" wrappers.  f is float; the velocity grid, the fields and
// every scalar stay double, so phases and accelerations are formed exactly as in the fp64 path and rounded once.
#include <math.h>

#include <mutex>
#include <vector>

#include "../../include/adept_b200.h"
#include "common32.cuh"

namespace adept32 {

int vdfdx_f32(const float* fin, float* fout, int batch, int nx, int nv, const double* v, double dt,
              const double* k1_batch, double k1, cudaStream_t stream, const float* filt);
int edfdv_exp_f32(const float* fin, float* fout, int batch, int nx, int nv, const double* e, const double* dex,
                  const double* pond, double q, double m, double dt, double k1, cudaStream_t stream);

int collide_f32(const float* fin, float* fout, int batch, int nx, int nv, const double* v, double dv, double dt,
                const double* nu_fp, const double* nu_K, const double* f_mx, int model, int scheme, int nodrag,
                double sg_m, double sg_ratio, float* n_out, double nu_fp_scale, double nu_K_scale, cudaStream_t stream,
                int sc_steps, double sc_rtol, double sc_atol, const double* coef_in, double* coef_out, int coef_div);

static std::mutex g_tw_mutex;
static cplx* g_tw[64][16] = {};

// per-pass Stockham tables exp(-2 pi i k r / (Ns R)), the layout fft_core.cuh documents (radix-16 passes, then one
// radix-2/4/8 pass), evaluated in long double and rounded to float once
const cplx* get_twiddles(int logn) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || logn < 1 || logn > 13) {
    set_last_error("get_twiddles(f32): no CUDA device or unsupported log2(n)=%d", logn);
    (void)cudaGetLastError();
    return nullptr;
  }
  std::lock_guard<std::mutex> lock(g_tw_mutex);
  if (g_tw[dev][logn]) return g_tw[dev][logn];
  std::vector<int> rad;
  if (logn < 4) {
    rad.push_back(1 << logn);
  } else {
    for (int i = 0; i < logn / 4; i++) rad.push_back(16);
    if (logn % 4) rad.push_back(1 << (logn % 4));
  }
  std::vector<cplx> host;
  int ns = rad[0];
  const long double two_pi = 6.283185307179586476925286766559005768L;
  for (size_t p = 1; p < rad.size(); p++) {
    const int R = rad[p];
    for (int r = 1; r < R; r++)
      for (int k = 0; k < ns; k++) {
        const long double ang = two_pi * (long double)((long long)k * r) / (long double)((long long)ns * R);
        host.push_back(make_float2((float)cosl(ang), (float)(-sinl(ang))));
      }
    ns *= R;
  }
  if (host.empty()) host.push_back(make_float2(1.0f, 0.0f));
  cplx* d = nullptr;
  cudaError_t err = cudaMalloc(&d, host.size() * sizeof(cplx));
  if (err == cudaSuccess) err = cudaMemcpy(d, host.data(), host.size() * sizeof(cplx), cudaMemcpyHostToDevice);
  if (err != cudaSuccess) {
    set_last_error("get_twiddles(f32, 2^%d): %s", logn, cudaGetErrorString(err));
    (void)cudaGetLastError();
    if (d) cudaFree(d);
    return nullptr;
  }
  g_tw[dev][logn] = d;
  return d;
}

}  // namespace adept32

#define F32_REQUIRE(ptr, name)                                              \
  if (!(ptr)) {                                                             \
    adept::set_last_error("%s: null pointer argument '%s'", __func__, name); \
    return adept::ADEPT_ERR_BAD_ARG;                                        \
  }

extern "C" int adept_b200_vdfdx_f32(const float* f_in, float* f_out, int batch, int nx, int nv, const double* v,
                                    double dt, double k1x, const double* k1x_batch, void* stream) {
  F32_REQUIRE(f_in, "f_in") F32_REQUIRE(f_out, "f_out") F32_REQUIRE(v, "v")
  return adept32::vdfdx_f32(f_in, f_out, batch, nx, nv, v, dt, k1x_batch, k1x, (cudaStream_t)stream, nullptr);
}

extern "C" int adept_b200_edfdv_exp_f32(const float* f_in, float* f_out, int batch, int nx, int nv, const double* e,
                                        const double* dex, const double* pond, double charge, double mass, double dt,
                                        double k1v, void* stream) {
  F32_REQUIRE(f_in, "f_in") F32_REQUIRE(f_out, "f_out") F32_REQUIRE(e, "e")
  return adept32::edfdv_exp_f32(f_in, f_out, batch, nx, nv, e, dex, pond, charge, mass, dt, k1v, (cudaStream_t)stream);
}

extern "C" int adept_b200_collide_f32(const float* f_in, float* f_out, int batch, int nx, int nv, const double* v,
                                      double dv, double dt, const double* nu_fp, const double* nu_K,
                                      const double* f_mx, int model, int scheme, float* n_out, void* stream) {
  F32_REQUIRE(f_in, "f_in") F32_REQUIRE(f_out, "f_out") F32_REQUIRE(v, "v")
  return adept32::collide_f32(f_in, f_out, batch, nx, nv, v, dv, dt, nu_fp, nu_K, f_mx, model, scheme, 0, 2.0, 0.5,
                              n_out, 1.0, 1.0, (cudaStream_t)stream, 0, 1e-8, 1e-12, nullptr, nullptr, 1);
}
