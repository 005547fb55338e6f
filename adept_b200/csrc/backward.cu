// Adjoint (reverse-mode) kernels of the vlasov-1d operators: what a jax.custom_vjp backward rule calls.
//
//   x-advection, v-advection w.r.t. f   linear with a unit-modulus symbol: the adjoint is the same kernel with dt -> -dt
//                                       (adept_b200_vdfdx_f64 / adept_b200_edfdv_exp_f64); nothing new is needed
//   v-advection w.r.t. the acceleration edfdv_exp_bwd_accel below
//   velocity moments                    moments_bwd below (broadcast of the row cotangents along v)
//   Poisson solve                       antisymmetric operator: rho_bar = -poisson(e_bar); nothing new is needed
//   collisions                          collide_bwd (collide.cu)
//
// Reference forward semantics: adept/_vlasov1d/solvers/pushers/vlasov.py:74-91 (VelocityExponential.push),
// adept/_vlasov1d/solvers/pushers/field.py:186-224.
#include "internal.h"
#include "push_core.cuh"

namespace adept {

// ---- d/d(accel_i) of f'_i = irfft(exp(-i kv dt a_i) rfft(f_i)) contracted with the cotangent g_i -------------------
//   a_bar_i = sum_j g_ij (d f'_ij / d a_i),   d f'/d a = irfft(-i kv dt exp(-i kv dt a) F)
// By Parseval this is a reduction over modes of conj(G_k) (-i kv_k dt P_k F_k); F and G come out of ONE complex FFT of
// z = f_i + i g_i (two-for-one), so the kernel costs half a push and no inverse transform.
struct AccelBwdArgs {
  const double* f;     // [rows, nv] input of the forward push
  const double* g;     // [rows, nv] cotangent of its output
  const double* e;     // [rows]
  const double* dex;   // nullable
  const double* pond;  // nullable
  double q, m, dt, k1;
  double* abar;  // [rows]
  const cplx* tw;
  int zero;
};

template <int LOGN>
__global__ void __launch_bounds__(FftCfg<LOGN>::T < 32 ? 32 : FftCfg<LOGN>::T)
    edfdv_exp_bwd_accel_kernel(AccelBwdArgs p) {
  using C = FftCfg<LOGN>;
  using PC = PhaseCfg<LOGN>;
  constexpr int N = C::N, E = C::E, T = C::T, H = E / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* buf = reinterpret_cast<cplx*>(smem_raw);
  cplx* ph = buf + C::BUF;
  double* red = reinterpret_cast<double*>(ph + 2 * PC::PER_SEQ);  // [32]
  const int t = threadIdx.x;
  const bool live = t < T;
  const int tt = live ? t : 0;
  const long long row = blockIdx.x;
  const int nv = N;

  double ee = p.e[row];
  if (p.dex) ee = __dadd_rn(ee, p.dex[row]);
  const double pd = p.pond ? p.pond[row] : 0.0;
  const double accel = accel_of(ee, pd, p.q, p.q * p.q / p.m, p.m);
  const double alpha = p.k1 * (p.dt * accel);
  if (live) phase_table_fill<LOGN>(ph, alpha, 0.0, tt, T);

  cplx x[E];
  const double* fr = p.f + row * nv;
  const double* gr = p.g + row * nv;
#pragma unroll
  for (int m = 0; m < E; m++) {
    const int e = tt + T * m;
    x[m] = cmake(fr[e], gr[e]);
  }
  fft_forward<LOGN>(x, buf, p.tw, tt, p.zero);
  __syncthreads();
  if (live) {
#pragma unroll
    for (int m = H; m < E; m++) buf[fft_pad(tt + T * m)] = x[m];
  }
  __syncthreads();
  double acc = 0.0;
  if (live) {
    cplx pa = cmul(ph[tt & (PC::NLO - 1)], ph[PC::NLO + (tt >> PC::LOBT)]);  // exp(-i k alpha) / (2N), k = t
    const cplx sa = ph[PC::STEP];
#pragma unroll
    for (int m = 0; m < H; m++) {
      if (m > 0) pa = cmul(pa, sa);
      const int k = tt + T * m;
      if (k == 0) continue;
      const cplx zk = x[m];
      const cplx zq = buf[fft_pad(N - k)];
      const cplx A = cmake(zk.x + zq.x, zk.y - zq.y);  // 2 F_k
      const cplx B = cmake(zk.y + zq.y, zq.x - zk.x);  // 2 G_k
      const cplx w = cmul(A, pa);
      const double kd = (p.k1 * (double)k) * p.dt;
      // d = -i kd w = kd (w.y, -w.x);  Re(conj(B) d) = B.x d.x + B.y d.y
      acc += kd * (B.x * w.y - B.y * w.x);
    }
    if (tt == 0) {  // Nyquist: irfft keeps only Re(-i k_N dt P_N F_N) = -k_N dt sin(theta_N) F_N
      const cplx z = buf[fft_pad(N / 2)];
      const double kd = (p.k1 * (double)(N / 2)) * p.dt;
      acc += -kd * sin(alpha * (double)(N / 2)) * z.x * z.y / (double)N;
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < (int)((blockDim.x + 31) >> 5); i++) s += red[i];
    p.abar[row] = s;
  }
}

template <int LOGN>
static int launch_accel_bwd(const AccelBwdArgs& p, long long rows, cudaStream_t stream) {
  using C = FftCfg<LOGN>;
  using PC = PhaseCfg<LOGN>;
  const size_t smem = (C::BUF + 2 * PC::PER_SEQ) * sizeof(cplx) + 32 * sizeof(double);
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = edfdv_exp_bwd_accel_kernel<LOGN>;
  if (dev < 64 && !configured[dev]) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(edfdv_exp_bwd_accel, smem=%zu): %s", smem, cudaGetErrorString(err));
      return ADEPT_ERR_CUDA;
    }
    configured[dev] = true;
  }
  const int threads = C::T < 32 ? 32 : C::T;
  ProfileScope prof("edfdv_exp_bwd_accel", stream);
  kern<<<(unsigned)rows, threads, smem, stream>>>(p);
  return check_launch("edfdv_exp_bwd_accel_kernel");
}

int edfdv_exp_bwd_accel_f64(const double* f, const double* g, int batch, int nx, int nv, const double* e,
                            const double* dex, const double* pond, double q, double m, double dt, double k1,
                            double* abar, cudaStream_t stream) {
  if (batch < 1 || nx < 1 || nv < 2 || (nv & (nv - 1)) || nv > 8192) {
    set_last_error("edfdv_exp_bwd_accel: nv=%d must be a power of two in [2, 8192] (batch=%d nx=%d)", nv, batch, nx);
    return ADEPT_ERR_UNSUPPORTED;
  }
  int logn = 0;
  while ((1 << logn) < nv) logn++;
  AccelBwdArgs p = {f, g, e, dex, pond, q, m, dt, k1, abar, get_twiddles(logn), 0};
  if (!p.tw) return ADEPT_ERR_CUDA;
  const long long rows = (long long)batch * nx;
  switch (logn) {
#define ADEPT_CASE(L) \
  case L:             \
    return launch_accel_bwd<L>(p, rows, stream);
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
  }
  return ADEPT_ERR_UNSUPPORTED;
}

// ---- adjoint of the velocity moments: fbar[row, j] (+)= sum_k coef[k] obar_k[row] v_j^k ------------------------------
struct MomentBwdArgs {
  const double* obar[3];  // nullable each
  double coef[3];         // scale_a * scale_b[k] of the forward call
  const double* v;
  long long rows;
  int nv;
  int accumulate;
  double* fbar;
};

__global__ void __launch_bounds__(256) moments_bwd_kernel(MomentBwdArgs p) {
  const long long row = blockIdx.x;
  double c[3];
#pragma unroll
  for (int k = 0; k < 3; k++) c[k] = p.obar[k] ? p.coef[k] * p.obar[k][row] : 0.0;
  double* out = p.fbar + row * p.nv;
  for (int j = threadIdx.x; j < p.nv; j += blockDim.x) {
    const double vv = p.v ? __ldg(p.v + j) : 0.0;
    const double val = c[0] + vv * (c[1] + vv * c[2]);
    out[j] = p.accumulate ? out[j] + val : val;
  }
}

int moments_bwd_f64(const double* const* obar, const double* coef, int batch, int nx, int nv, const double* v,
                    int accumulate, double* fbar, cudaStream_t stream) {
  if (batch < 1 || nx < 1 || nv < 1) {
    set_last_error("moments_bwd: bad shape batch=%d nx=%d nv=%d", batch, nx, nv);
    return ADEPT_ERR_BAD_SHAPE;
  }
  MomentBwdArgs p = {};
  for (int k = 0; k < 3; k++) p.obar[k] = obar[k], p.coef[k] = coef[k];
  if ((p.obar[1] || p.obar[2]) && !v) {
    set_last_error("moments_bwd: v grid required for first/second moments");
    return ADEPT_ERR_BAD_ARG;
  }
  p.v = v, p.rows = (long long)batch * nx, p.nv = nv, p.accumulate = accumulate, p.fbar = fbar;
  ProfileScope prof("moments_bwd", stream);
  moments_bwd_kernel<<<(unsigned)p.rows, 256, 0, stream>>>(p);
  return check_launch("moments_bwd_kernel");
}

}  // namespace adept
