// Spectral advection pushers: f <- irfft( exp(-i * m * alpha_seq) * rfft(f) ) along x (strided axis) or
// along v (contiguous axis), one fused kernel: one HBM read + one HBM write of f per call.
//
// Reference semantics (file:line relative to /root/reference):
//   x-advection  adept/_vlasov1d/solvers/pushers/vlasov.py:234-251  (SpaceExponential.push / __call__)
//   v-advection  adept/_vlasov1d/solvers/pushers/vlasov.py:74-91    (VelocityExponential.push)
//
// Two real sequences (two neighbouring v-columns, or two neighbouring x-rows) are transformed as one
// complex FFT z = a + i b.  The spectra are separated with the k <-> N-k symmetry, multiplied by their own
// phase factors (Im of DC/Nyquist dropped exactly like irfft does), recombined, and transformed back with the
// same forward kernel (swap trick).  Phase factors exp(-i m alpha) are built from two small per-sequence tables
// (m = 64*hi + lo) so no sincos is evaluated per element.
#include <stdlib.h>

#include "internal.h"
#include "push_core.cuh"

namespace adept {

enum { AXIS_X = 0, AXIS_V = 1 };

template <int LOGN, int AXIS>
struct PushCfg {
  static constexpr int T = FftCfg<LOGN>::T;
  // x-axis at N=4096: two column pairs per CTA so that the 16-byte column-pair loads of both groups fall in the
  // same 32-byte sectors at the same time (L1 merges them).
  static constexpr int THREADS_ = (T >= 512) ? T : ((AXIS == AXIS_X && T == 256) ? 512 : (16 * T > 256 ? 256 : 16 * T));
  static constexpr int THREADS = THREADS_ < 32 ? 32 : THREADS_;
  static constexpr int F = THREADS / T;  // FFT groups (sequence pairs) per CTA
  static constexpr size_t SMEM = (size_t)F * (FftCfg<LOGN>::BUF + 2 * PhaseCfg<LOGN>::PER_SEQ) * sizeof(cplx);
};

struct PushArgs {
  const double* fin;
  double* fout;
  int batch, nx, nv;
  long long npairs;  // total sequence pairs over the batch
  // (lines tagged "f64" keep double precision in the generated fp32 build: phases are formed in fp64)
  // AXIS_X: alpha_j = k1[b] * (v[j] * dt)
  const double* v;         // f64
  const double* k1_batch;  // f64  nullable -> k1
  double k1;               // f64
  double dt;               // f64
  // AXIS_V: accel_i = (q*(e_i+dex_i) + (q*q/m)*pond_i)/m ; alpha_i = k1 * (dt * accel_i)
  const double* e;     // f64
  const double* dex;   // f64  nullable
  const double* pond;  // f64  nullable
  double q, m;         // f64
  const cplx* tw;
  int zero;  // always 0; opaque to ptxas (see FftPass)
  const double* filt;  // nullable [N/2+1]: real multiplier per mode (Hou-Li filter)
};

template <int LOGN, int AXIS, int TW = 0>
__global__ void __launch_bounds__(PushCfg<LOGN, AXIS>::THREADS, (PushCfg<LOGN, AXIS>::THREADS <= 256 ? 2 : 1))
    spectral_push_kernel(PushArgs p) {
  using C = FftCfg<LOGN>;
  using PC = PhaseCfg<LOGN>;
  using K = PushCfg<LOGN, AXIS>;
  constexpr int N = C::N, E = C::E, T = C::T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* smem = reinterpret_cast<cplx*>(smem_raw);

  const int g = threadIdx.x / T;  // FFT group in this CTA
  const int t = threadIdx.x % T;
  cplx* buf = smem + (size_t)g * C::BUF;
  cplx* ph = smem + (size_t)K::F * C::BUF + (size_t)g * 2 * PC::PER_SEQ;  // [seq][hi.. lo..]

  const long long G = (long long)blockIdx.x * K::F + g;
  const bool active = G < p.npairs;

  // ---- addressing + per-sequence phase increments ------------------------------------------------------
  const double* src = p.fin;
  double* dst = p.fout;
  long long stride = 1;  // element stride along the transformed axis
  double alpha_a = 0.0, alpha_b = 0.0;  // f64
  if (active) {
    if (AXIS == AXIS_V) {
      const long long row0 = 2 * G;  // global row over [batch, nx]
      src += row0 * p.nv;
      dst += row0 * p.nv;
      const int b = (int)(row0 / p.nx);
      const double k1 = p.k1_batch ? p.k1_batch[b] : p.k1;  // f64
      const double q2m = p.q * p.q / p.m;                     // f64
      double acc[2];                                          // f64
#pragma unroll
      for (int s = 0; s < 2; s++) {
        double ee = p.e[row0 + s];                           // f64
        if (p.dex) ee = __dadd_rn(ee, p.dex[row0 + s]);      // f64
        const double pd = p.pond ? p.pond[row0 + s] : 0.0;   // f64
        acc[s] = accel_of(ee, pd, p.q, q2m, p.m);
      }
      alpha_a = k1 * (p.dt * acc[0]);
      alpha_b = k1 * (p.dt * acc[1]);
    } else {
      const int half = p.nv / 2;
      const int b = (int)(G / half);
      const int cp = (int)(G % half);
      src += (long long)b * p.nx * p.nv + 2 * cp;
      dst += (long long)b * p.nx * p.nv + 2 * cp;
      stride = p.nv;
      const double k1 = p.k1_batch ? p.k1_batch[b] : p.k1;  // f64
      alpha_a = k1 * (p.v[2 * cp] * p.dt);
      alpha_b = k1 * (p.v[2 * cp + 1] * p.dt);
    }
  }

  // ---- load: x[m] = a[e] + i b[e], e = t + T m -----------------------------------------------------------
  cplx x[E];
  if (active) {
    if (AXIS == AXIS_V) {
      const double* a = src;
      const double* bq = src + p.nv;
#pragma unroll
      for (int m = 0; m < E; m++) {
        const int e = t + T * m;
        x[m] = cmake(__ldcs(a + e), __ldcs(bq + e));  // streaming: keep L1 for the twiddle table
      }
    } else {
#pragma unroll
      for (int m = 0; m < E; m++) {
        const long long e = t + T * m;
        x[m] = *reinterpret_cast<const double2*>(src + e * stride);
      }
    }
  } else {
#pragma unroll
    for (int m = 0; m < E; m++) x[m] = cmake(0.0, 0.0);
  }

  // after the loads are in flight: the sincos latency of the filling warps hides behind the HBM latency of the loads
  phase_table_fill<LOGN>(ph, alpha_a, alpha_b, t, T);

  fft_forward<LOGN, 1, TW>(x, buf, p.tw, t, p.zero);

  half_spectrum_update<LOGN, 1>(x, buf, ph, t, p.filt);

  fft_forward<LOGN, 1, TW>(x, buf, p.tw + p.zero, t, p.zero);  // opaque offset: no CSE of twiddle loads with the first FFT

  // ---- store: a'[e] = Im, b'[e] = Re of the swapped result ---------------------------------------------
  if (active) {
    if (AXIS == AXIS_V) {
      double* a = dst;
      double* bq = dst + p.nv;
#pragma unroll
      for (int m = 0; m < E; m++) {
        const int e = t + T * m;
        __stcs(a + e, x[m].y);
        __stcs(bq + e, x[m].x);
      }
    } else {
#pragma unroll
      for (int m = 0; m < E; m++) {
        const long long e = t + T * m;
        *reinterpret_cast<double2*>(dst + e * stride) = make_double2(x[m].y, x[m].x);
      }
    }
  }
}

template <int LOGN, int AXIS, int TW = 0>
static int launch_push(const PushArgs& p, cudaStream_t stream) {
  using K = PushCfg<LOGN, AXIS>;
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = spectral_push_kernel<LOGN, AXIS, TW>;
  if (dev < 64 && !configured[dev]) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM);
    if (err != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(push, smem=%zu): %s", K::SMEM, cudaGetErrorString(err));
      return ADEPT_ERR_CUDA;
    }
    configured[dev] = true;
  }
  const long long blocks = (p.npairs + K::F - 1) / K::F;
  ProfileScope prof(AXIS == AXIS_X ? "vdfdx" : "edfdv_exp", stream);
  kern<<<(unsigned)blocks, K::THREADS, K::SMEM, stream>>>(p);
  return check_launch("spectral_push_kernel");
}

// The nv = 4096 v-advection runs with two twiddle loads per pass (fft_core.cuh TW = 1): 98.3 -> 96.3 us (r02s);
// ADEPT_B200_PTW=0 selects the six-load passes for A/B timing
static bool push_tw1() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("ADEPT_B200_PTW");
    mode = (e && atoi(e) == 0) ? 0 : 1;
  }
  return mode == 1;
}

template <int AXIS>
static int dispatch_push(int logn, const PushArgs& p, cudaStream_t stream) {
  if (AXIS == AXIS_V && logn == 12 && push_tw1()) return launch_push<12, AXIS, 1>(p, stream);
  switch (logn) {
#define ADEPT_CASE(L) \
  case L:             \
    return launch_push<L, AXIS>(p, stream);
    ADEPT_CASE(1)
    ADEPT_CASE(2)
    ADEPT_CASE(3)
    ADEPT_CASE(4)
    ADEPT_CASE(5)
    ADEPT_CASE(6)
    ADEPT_CASE(7)
    ADEPT_CASE(8)
    ADEPT_CASE(9)
    ADEPT_CASE(10)
    ADEPT_CASE(11)
    ADEPT_CASE(12)
    ADEPT_CASE(13)
#undef ADEPT_CASE
    default:
      set_last_error("spectral push: unsupported transform length 2^%d", logn);
      return ADEPT_ERR_UNSUPPORTED;
  }
}

static int ilog2_exact(int n) {
  if (n < 2 || (n & (n - 1))) return -1;
  int l = 0;
  while ((1 << l) < n) l++;
  return l;
}

int vdfdx_f64(const double* fin, double* fout,
              int batch, int nx, int nv, const double* v, double dt, const double* k1_batch, double k1,  // f64
              cudaStream_t stream, const double* filt) {
  if (batch < 1 || nx < 2 || nv < 2 || (nv & 1)) {
    set_last_error("vdfdx: bad shape batch=%d nx=%d nv=%d (nv must be even)", batch, nx, nv);
    return ADEPT_ERR_BAD_SHAPE;
  }
  const int logn = ilog2_exact(nx);
  if (logn < 0 && bluestein_supported(nx))  // any even length <= 4096: chirp-z on the power-of-two core
    return bluestein_push_f64(0, fin, fout, batch, nx, nv, v, dt, k1_batch, k1, nullptr, nullptr, nullptr, 0.0, 1.0,
                              filt, stream);
  if (logn < 1 || logn > 13) {
    set_last_error("vdfdx: nx=%d must be a power of two <= 8192 or an even number <= 4096", nx);
    return ADEPT_ERR_UNSUPPORTED;
  }
  if ((reinterpret_cast<uintptr_t>(fin) | reinterpret_cast<uintptr_t>(fout)) & 15) {
    set_last_error("vdfdx: f buffers must be 16-byte aligned");
    return ADEPT_ERR_BAD_ARG;
  }
  PushArgs p = {};
  p.fin = fin, p.fout = fout, p.batch = batch, p.nx = nx, p.nv = nv;
  p.npairs = (long long)batch * (nv / 2);
  p.v = v, p.k1_batch = k1_batch, p.k1 = k1, p.dt = dt, p.filt = filt;
  p.tw = get_twiddles(logn);
  if (!p.tw) return ADEPT_ERR_CUDA;
  return dispatch_push<AXIS_X>(logn, p, stream);
}

int edfdv_exp_f64(const double* fin, double* fout,
                  int batch, int nx, int nv, const double* e, const double* dex, const double* pond,  // f64
                  double q, double m, double dt, double k1,                                           // f64
                  cudaStream_t stream) {
  if (batch < 1 || nx < 2 || nv < 2 || (nx & 1)) {
    set_last_error("edfdv_exp: bad shape batch=%d nx=%d nv=%d (nx must be even)", batch, nx, nv);
    return ADEPT_ERR_BAD_SHAPE;
  }
  const int logn = ilog2_exact(nv);
  if (logn < 0 && bluestein_supported(nv))
    return bluestein_push_f64(1, fin, fout, batch, nx, nv, nullptr, dt, nullptr, k1, e, dex, pond, q, m, nullptr, stream);
  if (logn < 1 || logn > 13) {
    set_last_error("edfdv_exp: nv=%d must be a power of two <= 8192 or an even number <= 4096", nv);
    return ADEPT_ERR_UNSUPPORTED;
  }
  PushArgs p = {};
  p.fin = fin, p.fout = fout, p.batch = batch, p.nx = nx, p.nv = nv;
  p.npairs = (long long)batch * (nx / 2);
  p.e = e, p.dex = dex, p.pond = pond, p.q = q, p.m = m, p.dt = dt, p.k1 = k1;
  p.tw = get_twiddles(logn);
  if (!p.tw) return ADEPT_ERR_CUDA;
  return dispatch_push<AXIS_V>(logn, p, stream);
}

// ---- |rfft_x f|: the spectrum save of get_dist_save_func's {t, kx, v} block (adept/_vlasov1d/storage.py:183-190) ------
// One complex FFT per column pair like the x-advection; the two real spectra are separated with the k <-> N - k
// symmetry and their moduli written to out[k, col], k = 0 .. N/2.
template <int LOGN>
__global__ void __launch_bounds__(PushCfg<LOGN, AXIS_X>::THREADS, (PushCfg<LOGN, AXIS_X>::THREADS <= 256 ? 2 : 1))
    abs_rfft_x_kernel(PushArgs p) {
  using C = FftCfg<LOGN>;
  using K = PushCfg<LOGN, AXIS_X>;
  constexpr int N = C::N, E = C::E, T = C::T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* smem = reinterpret_cast<cplx*>(smem_raw);
  const int g = threadIdx.x / T, t = threadIdx.x % T;
  cplx* buf = smem + (size_t)g * C::BUF;
  const long long G = (long long)blockIdx.x * K::F + g;
  const bool active = G < p.npairs;
  const int half = p.nv / 2;
  const int b = active ? (int)(G / half) : 0, cp = active ? (int)(G % half) : 0;
  const double* src = p.fin + (long long)b * p.nx * p.nv + 2 * cp;
  double* dst = p.fout + (long long)b * (N / 2 + 1) * p.nv + 2 * cp;
  cplx x[E];
#pragma unroll
  for (int m = 0; m < E; m++)
    x[m] = active ? *reinterpret_cast<const double2*>(src + (long long)(t + T * m) * p.nv) : cmake(0.0, 0.0);
  fft_forward<LOGN>(x, buf, p.tw, t, p.zero);
  __syncthreads();
#pragma unroll
  for (int m = 0; m < E; m++) buf[fft_pad(t + T * m)] = x[m];
  __syncthreads();
  if (active) {
#pragma unroll
    for (int m = 0; m < E; m++) {
      const int k = t + T * m;
      if (k > N / 2) continue;
      const cplx zq = buf[fft_pad((N - k) & (N - 1))];
      // A_k = (Z_k + conj(Z_{N-k})) / 2,  B_k = (Z_k - conj(Z_{N-k})) / (2 i)
      const double ar = 0.5 * (x[m].x + zq.x), ai = 0.5 * (x[m].y - zq.y);
      const double br = 0.5 * (x[m].y + zq.y), bi = 0.5 * (zq.x - x[m].x);
      *reinterpret_cast<double2*>(dst + (long long)k * p.nv) = make_double2(hypot(ar, ai), hypot(br, bi));
    }
  }
}

template <int LOGN>
static int launch_abs_rfft_x(const PushArgs& p, cudaStream_t stream) {
  using K = PushCfg<LOGN, AXIS_X>;
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = abs_rfft_x_kernel<LOGN>;
  if (dev < 64 && !configured[dev]) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM);
    if (err != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(abs_rfft_x, smem=%zu): %s", K::SMEM, cudaGetErrorString(err));
      return ADEPT_ERR_CUDA;
    }
    configured[dev] = true;
  }
  const long long blocks = (p.npairs + K::F - 1) / K::F;
  ProfileScope prof("abs_rfft_x", stream);
  kern<<<(unsigned)blocks, K::THREADS, K::SMEM, stream>>>(p);
  return check_launch("abs_rfft_x_kernel");
}

int abs_rfft_x_f64(const double* fin, double* fout, int batch, int nx, int nv, cudaStream_t stream) {
  if (batch < 1 || nx < 2 || nv < 2 || (nv & 1)) {
    set_last_error("abs_rfft_x: bad shape batch=%d nx=%d nv=%d (nv must be even)", batch, nx, nv);
    return ADEPT_ERR_BAD_SHAPE;
  }
  const int logn = ilog2_exact(nx);
  if (logn < 1 || logn > 13) {
    set_last_error("abs_rfft_x: nx=%d must be a power of two <= 8192", nx);
    return ADEPT_ERR_UNSUPPORTED;
  }
  if ((reinterpret_cast<uintptr_t>(fin) | reinterpret_cast<uintptr_t>(fout)) & 15) {
    set_last_error("abs_rfft_x: buffers must be 16-byte aligned");
    return ADEPT_ERR_BAD_ARG;
  }
  PushArgs p = {};
  p.fin = fin, p.fout = fout, p.batch = batch, p.nx = nx, p.nv = nv;
  p.npairs = (long long)batch * (nv / 2);
  p.tw = get_twiddles(logn);
  if (!p.tw) return ADEPT_ERR_CUDA;
  switch (logn) {
#define ADEPT_CASE(L) \
  case L:             \
    return launch_abs_rfft_x<L>(p, stream);
    ADEPT_CASE(1) ADEPT_CASE(2) ADEPT_CASE(3) ADEPT_CASE(4) ADEPT_CASE(5) ADEPT_CASE(6) ADEPT_CASE(7)
    ADEPT_CASE(8) ADEPT_CASE(9) ADEPT_CASE(10) ADEPT_CASE(11) ADEPT_CASE(12) ADEPT_CASE(13)
#undef ADEPT_CASE
  }
  return ADEPT_ERR_UNSUPPORTED;
}

}  // namespace adept
