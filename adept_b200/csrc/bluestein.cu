// Spectral advection and field solve for transform lengths that are NOT powers of two (any even n <= 4096), by
// Bluestein's chirp-z algorithm on top of the power-of-two FFT core: DFT_n = chirp . (circular convolution of length
// M = 2^k >= 2n - 1 with the conjugate chirp) . chirp, i.e. two FFTs of length M per DFT of length n.  jnp.fft in the
// reference takes any length (stock decks use nx = 1028, 1728, 3456); this is the general path behind the same entry
// points, four to eight times the work of the power-of-two kernels but the same semantics:
//   x-advection  adept/_vlasov1d/solvers/pushers/vlasov.py:234-251, v-advection :74-91 (rfft . phase . irfft),
//   field solve  adept/_vlasov1d/solvers/pushers/field.py:210-224, 282-298.
// Tables (chirp, FFT of the conjugate chirp with 1/M folded in) come from get_bluestein() in api.cu.
#include "fft_core.cuh"
#include "internal.h"

namespace adept {

// x[m] = z[e], e = t + T m (anything for e >= n)  ->  x[m] = DFT_n(z)[e] for e < n, 0 for e >= n.
// All threads of the CTA (exactly T of them) must call it.
template <int LOGM>
__device__ __forceinline__ void bluestein_dft(cplx (&x)[FftCfg<LOGM>::E], cplx* buf, const cplx* tw, int zero, int t,
                                              int n, const cplx* __restrict__ chirp, const cplx* __restrict__ bhat) {
  using C = FftCfg<LOGM>;
  constexpr int E = C::E, T = C::T;
#pragma unroll
  for (int m = 0; m < E; m++) {
    const int e = t + T * m;
    x[m] = e < n ? cmul(x[m], __ldg(chirp + e)) : cmake(0.0, 0.0);
  }
  fft_forward<LOGM>(x, buf, tw, t, zero);
#pragma unroll
  for (int m = 0; m < E; m++) x[m] = cconj(cmul(x[m], __ldg(bhat + t + T * m)));  // conj: inverse FFT by a forward one
  fft_forward<LOGM>(x, buf, tw + zero, t, zero);
#pragma unroll
  for (int m = 0; m < E; m++) {
    const int e = t + T * m;
    x[m] = e < n ? cmul(cconj(x[m]), __ldg(chirp + e)) : cmake(0.0, 0.0);
  }
}

struct BsPushArgs {
  const double* fin;
  double* fout;
  int batch, nx, nv;
  long long npairs;
  int n;     // transform length (nx or nv)
  int axis;  // 0: along x (pairs of columns), 1: along v (pairs of rows)
  const double* v;
  const double* k1_batch;
  double k1, dt;
  const double* e;
  const double* dex;
  const double* pond;
  double q, m;
  const double* filt;
  const cplx *tw, *chirp, *bhat;
  int zero;
};

template <int LOGM>
__global__ void __launch_bounds__(FftCfg<LOGM>::T) bluestein_push_kernel(BsPushArgs p) {
  using C = FftCfg<LOGM>;
  constexpr int E = C::E, T = C::T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* buf = reinterpret_cast<cplx*>(smem_raw);
  const int t = threadIdx.x, n = p.n;
  const long long G = blockIdx.x;
  const double* src = p.fin;
  double* dst = p.fout;
  long long stride = 1, pair_off = 1;  // element stride along the transform, offset between the two sequences
  double alpha_a, alpha_b;
  if (p.axis == 1) {
    const long long row0 = 2 * G;
    src += row0 * p.nv, dst += row0 * p.nv;
    pair_off = p.nv;
    const int b = (int)(row0 / p.nx);
    const double k1 = p.k1_batch ? p.k1_batch[b] : p.k1;
    const double q2m = p.q * p.q / p.m;
    double acc[2];
#pragma unroll
    for (int s = 0; s < 2; s++) {
      double ee = p.e[row0 + s];
      if (p.dex) ee = __dadd_rn(ee, p.dex[row0 + s]);
      const double pd = p.pond ? p.pond[row0 + s] : 0.0;
      acc[s] = accel_of(ee, pd, p.q, q2m, p.m);
    }
    alpha_a = k1 * (p.dt * acc[0]), alpha_b = k1 * (p.dt * acc[1]);
  } else {
    const int half = p.nv / 2;
    const int b = (int)(G / half), cp = (int)(G % half);
    src += (long long)b * p.nx * p.nv + 2 * cp, dst += (long long)b * p.nx * p.nv + 2 * cp;
    stride = p.nv;
    const double k1 = p.k1_batch ? p.k1_batch[b] : p.k1;
    alpha_a = k1 * (p.v[2 * cp] * p.dt), alpha_b = k1 * (p.v[2 * cp + 1] * p.dt);
  }

  cplx x[E];
#pragma unroll
  for (int m = 0; m < E; m++) {
    const int e = t + T * m;
    x[m] = e < n ? cmake(src[e * stride], src[e * stride + pair_off]) : cmake(0.0, 0.0);
  }
  bluestein_dft<LOGM>(x, buf, p.tw, p.zero, t, n, p.chirp, p.bhat);

  // spectra of the two real sequences from Z[k] and Z[n - k], their own phase factors, recombination
  __syncthreads();
#pragma unroll
  for (int m = 0; m < E; m++) {
    const int e = t + T * m;
    if (e < n) buf[e] = x[m];
  }
  __syncthreads();
#pragma unroll
  for (int m = 0; m < E; m++) {
    const int k = t + T * m;
    if (k < n) {
      const cplx zk = x[m], zq = buf[k ? n - k : 0];
      const cplx A = cmake(0.5 * (zk.x + zq.x), 0.5 * (zk.y - zq.y));   // (Z[k] + conj Z[n-k]) / 2
      const cplx B = cmake(0.5 * (zk.y + zq.y), 0.5 * (zq.x - zk.x));   // (Z[k] - conj Z[n-k]) / (2 i)
      const int keff = (2 * k <= n) ? k : k - n;
      double sa, ca, sb, cb;
      sincos(alpha_a * (double)keff, &sa, &ca);
      sincos(alpha_b * (double)keff, &sb, &cb);
      cplx pa = cmake(ca, -sa), pb = cmake(cb, -sb);          // exp(-i alpha keff)
      if (2 * k == n) pa.y = 0.0, pb.y = 0.0;                  // Nyquist: irfft keeps the real part only
      cplx Ap = cmul(A, pa), Bp = cmul(B, pb);
      if (p.filt) {
        const double s = __ldg(p.filt + ((2 * k <= n) ? k : n - k));
        Ap.x *= s, Ap.y *= s, Bp.x *= s, Bp.y *= s;
      }
      // Z' = A' + i B', conjugated for the inverse transform
      x[m] = cmake(Ap.x - Bp.y, -(Ap.y + Bp.x));
    }
  }
  bluestein_dft<LOGM>(x, buf, p.tw, p.zero, t, n, p.chirp, p.bhat);
  const double inv_n = 1.0 / (double)n;
#pragma unroll
  for (int m = 0; m < E; m++) {
    const int e = t + T * m;
    if (e < n) {
      dst[e * stride] = x[m].x * inv_n;                  // z' = conj(DFT(conj Z')) / n
      dst[e * stride + pair_off] = -x[m].y * inv_n;
    }
  }
}

struct BsPoissonArgs {
  const double* rho;
  const double* kmul;
  long long kmul_stride;
  double* e;
  int n, mode;
  double Te, lambda_De;
  const cplx *tw, *chirp, *bhat;
  int zero;
};

template <int LOGM>
__global__ void __launch_bounds__(FftCfg<LOGM>::T) bluestein_poisson_kernel(BsPoissonArgs p) {
  using C = FftCfg<LOGM>;
  constexpr int E = C::E, T = C::T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* buf = reinterpret_cast<cplx*>(smem_raw);
  __shared__ double rho0_s;
  const int t = threadIdx.x, n = p.n;
  const double* rho = p.rho + (long long)blockIdx.x * n;
  const double* kmul = p.kmul + (long long)blockIdx.x * p.kmul_stride;
  double* eo = p.e + (long long)blockIdx.x * n;
  cplx x[E];
#pragma unroll
  for (int m = 0; m < E; m++) {
    const int e = t + T * m;
    x[m] = cmake(e < n ? rho[e] : 0.0, 0.0);
  }
  bluestein_dft<LOGM>(x, buf, p.tw, p.zero, t, n, p.chirp, p.bhat);
  if (t == 0) rho0_s = x[0].x / (double)n;
  __syncthreads();
  const double rho0 = rho0_s;
#pragma unroll
  for (int m = 0; m < E; m++) {
    const int k = t + T * m;
    if (k < n) {
      double mult;
      if (p.mode == 0) {
        mult = kmul[k];
      } else {
        const double kx = kmul[k];
        const double lam_sq = p.lambda_De < 0.0 ? p.Te / rho0 : p.lambda_De * p.lambda_De;
        mult = kx * (p.Te / rho0) / (1.0 + lam_sq * kx * kx);
      }
      // Y = -i mult X = (mult Xi, -mult Xr); conjugated for the inverse transform
      x[m] = cmake(mult * x[m].y, mult * x[m].x);
    }
  }
  bluestein_dft<LOGM>(x, buf, p.tw, p.zero, t, n, p.chirp, p.bhat);
#pragma unroll
  for (int m = 0; m < E; m++) {
    const int e = t + T * m;
    if (e < n) eo[e] = x[m].x / (double)n;  // Re(conj(.)) = Re(.)
  }
}

template <int LOGM, class KERN, class ARGS>
static int launch_bs(KERN kern, const ARGS& p, long long blocks, const char* name, cudaStream_t stream) {
  const size_t smem = (size_t)FftCfg<LOGM>::BUF * sizeof(cplx);
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) {
    set_last_error("cudaFuncSetAttribute(%s, smem=%zu): %s", name, smem, cudaGetErrorString(err));
    return ADEPT_ERR_CUDA;
  }
  ProfileScope prof(name, stream);
  kern<<<(unsigned)blocks, FftCfg<LOGM>::T, smem, stream>>>(p);
  return check_launch(name);
}

bool bluestein_supported(int n) { return n >= 2 && n <= 4096 && (n & 1) == 0 && (n & (n - 1)) != 0; }

#define ADEPT_BS_SWITCH(LOGM_VAR, CALL)         \
  switch (LOGM_VAR) {                           \
    case 3: { constexpr int L = 3; return CALL; }   \
    case 4: { constexpr int L = 4; return CALL; }   \
    case 5: { constexpr int L = 5; return CALL; }   \
    case 6: { constexpr int L = 6; return CALL; }   \
    case 7: { constexpr int L = 7; return CALL; }   \
    case 8: { constexpr int L = 8; return CALL; }   \
    case 9: { constexpr int L = 9; return CALL; }   \
    case 10: { constexpr int L = 10; return CALL; } \
    case 11: { constexpr int L = 11; return CALL; } \
    case 12: { constexpr int L = 12; return CALL; } \
    case 13: { constexpr int L = 13; return CALL; } \
  }

// axis 0: x-advection (n = nx), axis 1: v-advection (n = nv); arguments as vdfdx_f64 / edfdv_exp_f64
int bluestein_push_f64(int axis, const double* fin, double* fout, int batch, int nx, int nv, const double* v, double dt,
                       const double* k1_batch, double k1, const double* e, const double* dex, const double* pond,
                       double q, double m, const double* filt, cudaStream_t stream) {
  const int n = axis == 0 ? nx : nv;
  BsPushArgs p = {};
  p.fin = fin, p.fout = fout, p.batch = batch, p.nx = nx, p.nv = nv, p.n = n, p.axis = axis;
  p.npairs = axis == 0 ? (long long)batch * (nv / 2) : (long long)batch * (nx / 2);
  p.v = v, p.k1_batch = k1_batch, p.k1 = k1, p.dt = dt, p.e = e, p.dex = dex, p.pond = pond, p.q = q, p.m = m;
  p.filt = filt, p.zero = 0;
  int logm = 0;
  int rc = get_bluestein(n, &logm, &p.chirp, &p.bhat);
  if (rc != ADEPT_OK) return rc;
  p.tw = get_twiddles(logm);
  if (!p.tw) return ADEPT_ERR_CUDA;
  ADEPT_BS_SWITCH(logm, (launch_bs<L>(bluestein_push_kernel<L>, p, p.npairs, axis == 0 ? "vdfdx_bluestein" : "edfdv_bluestein", stream)))
  set_last_error("bluestein push: unsupported convolution length 2^%d", logm);
  return ADEPT_ERR_UNSUPPORTED;
}

int bluestein_poisson_f64(const double* rho, const double* kmul, long long kmul_stride, double* e, int batch, int nx,
                          int mode, double Te, double lambda_De, cudaStream_t stream) {
  BsPoissonArgs p = {};
  p.rho = rho, p.kmul = kmul, p.kmul_stride = kmul_stride, p.e = e, p.n = nx, p.mode = mode, p.Te = Te;
  p.lambda_De = lambda_De, p.zero = 0;
  int logm = 0;
  int rc = get_bluestein(nx, &logm, &p.chirp, &p.bhat);
  if (rc != ADEPT_OK) return rc;
  p.tw = get_twiddles(logm);
  if (!p.tw) return ADEPT_ERR_CUDA;
  ADEPT_BS_SWITCH(logm, (launch_bs<L>(bluestein_poisson_kernel<L>, p, batch, "poisson_bluestein", stream)))
  set_last_error("bluestein poisson: unsupported convolution length 2^%d", logm);
  return ADEPT_ERR_UNSUPPORTED;
}

}  // namespace adept
