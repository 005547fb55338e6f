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

// ---- adjoint of the cubic-spline (semi-Lagrangian) v-advection: rowops.cu::spline_push_kernel ---------------------
// Forward (vlasov.py:106-172): out_j = h00(t) f_l + h10(t) m0 + h01(t) f_{l+1} + h11(t) m1 with l = clamp(j + off),
// t = clamp(j - s - l), s = accel dt / dv, m0 / m1 the central (one-sided at the ends) slopes; 1e-30 outside the grid.
// f_bar: the four taps of every output scatter back (shared-memory atomics, one row per CTA); accel_bar: the chain
// through t (d t / d s = -1 wherever t is not clamped; the integer offset is piecewise constant).
struct SplineBwdArgs {
  const double* f;     // forward input [rows, nv]
  const double* g;     // cotangent of the forward output
  const double* e;
  const double* dex;   // nullable
  const double* pond;  // nullable
  double q, m, dt, dv;
  int nv;
  double* fbar;   // nullable [rows, nv]
  double* abar;   // nullable [rows]: cotangent of the acceleration
};

__global__ void __launch_bounds__(256) spline_push_bwd_kernel(SplineBwdArgs p) {
  extern __shared__ double sacc[];  // [nv] f_bar of this row, then 8 doubles of reduction scratch
  const long long row = blockIdx.x;
  const int nv = p.nv;
  const double* fr = p.f + row * nv;
  const double* gr = p.g + row * nv;
  double ee = p.e[row];
  if (p.dex) ee = __dadd_rn(ee, p.dex[row]);
  const double pd = p.pond ? p.pond[row] : 0.0;
  const double shift = __dmul_rn(accel_of(ee, pd, p.q, p.q * p.q / p.m, p.m), p.dt);
  const double scaled = __ddiv_rn(shift, p.dv);
  const int row_offset = (int)floor(-scaled);
  for (int j = threadIdx.x; j < nv; j += blockDim.x) sacc[j] = 0.0;
  __syncthreads();
  double ds = 0.0;  // sum_j g_j d out_j / d t
  for (int j = threadIdx.x; j < nv; j += blockDim.x) {
    const double query = __dsub_rn((double)j, scaled);
    if (query < 0.0 || query > (double)(nv - 1)) continue;  // constant output: no dependence on f or the shift
    int left = j + row_offset;
    left = left < 0 ? 0 : (left > nv - 2 ? nv - 2 : left);
    double t = __dsub_rn(query, (double)left);
    const bool clamped = t < 0.0 || t > 1.0;
    t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
    const int im1 = left - 1 < 0 ? 0 : left - 1;
    const int ip2 = left + 2 > nv - 1 ? nv - 1 : left + 2;
    const double t2 = t * t, t3 = t2 * t;
    const double h00 = 2.0 * t3 - 3.0 * t2 + 1.0, h10 = t3 - 2.0 * t2 + t, h01 = -2.0 * t3 + 3.0 * t2, h11 = t3 - t2;
    const double gj = gr[j];
    // out = h00 f0 + h10 m0 + h01 f1 + h11 m1;  m0 = (left == 0) ? f1 - f0 : (f1 - fm1)/2;  m1 = (left == nv-2) ? f1 - f0 : (f2 - f0)/2
    double c_m1 = 0.0, c_0 = h00, c_1 = h01, c_2 = 0.0;
    if (left == 0) c_0 -= h10, c_1 += h10; else c_m1 -= 0.5 * h10, c_1 += 0.5 * h10;
    if (left == nv - 2) c_0 -= h11, c_1 += h11; else c_2 += 0.5 * h11, c_0 -= 0.5 * h11;
    if (p.fbar) {
      atomicAdd(&sacc[left], gj * c_0);
      atomicAdd(&sacc[left + 1], gj * c_1);
      if (c_m1 != 0.0) atomicAdd(&sacc[im1], gj * c_m1);
      if (c_2 != 0.0) atomicAdd(&sacc[ip2], gj * c_2);
    }
    if (p.abar && !clamped) {
      const double fm1 = fr[im1], f0 = fr[left], f1 = fr[left + 1], f2 = fr[ip2];
      const double m0 = (left == 0) ? (f1 - f0) : 0.5 * (f1 - fm1);
      const double m1 = (left == nv - 2) ? (f1 - f0) : 0.5 * (f2 - f0);
      const double dh00 = 6.0 * t2 - 6.0 * t, dh10 = 3.0 * t2 - 4.0 * t + 1.0, dh01 = -dh00, dh11 = 3.0 * t2 - 2.0 * t;
      ds += gj * (dh00 * f0 + dh10 * m0 + dh01 * f1 + dh11 * m1);
    }
  }
  __syncthreads();
  if (p.fbar) {
    double* out = p.fbar + row * nv;
    for (int j = threadIdx.x; j < nv; j += blockDim.x) out[j] = sacc[j];
  }
  if (p.abar) {
    double* red = sacc + nv;
    ds = warp_sum(ds);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ds;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += red[w];
      p.abar[row] = tot * (-p.dt / p.dv);  // t = j - accel dt / dv - left
    }
  }
}

int edfdv_spline_bwd_f64(const double* f, const double* g, int batch, int nx, int nv, const double* e, const double* dex,
                         const double* pond, double q, double m, double dt, double dv, double* fbar, double* abar,
                         cudaStream_t stream) {
  if (batch < 1 || nx < 1 || nv < 2 || (size_t)(nv + 8) * sizeof(double) > 227 * 1024) {
    set_last_error("edfdv_spline_bwd: bad shape batch=%d nx=%d nv=%d", batch, nx, nv);
    return ADEPT_ERR_BAD_SHAPE;
  }
  SplineBwdArgs p = {f, g, e, dex, pond, q, m, dt, dv, nv, fbar, abar};
  const size_t smem = (size_t)(nv + 8) * sizeof(double);
  static size_t configured[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && configured[dev] < smem && smem > 48 * 1024) {
    cudaError_t err = cudaFuncSetAttribute(spline_push_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(spline_push_bwd, smem=%zu): %s", smem, cudaGetErrorString(err));
      return ADEPT_ERR_CUDA;
    }
    configured[dev] = smem;
  }
  ProfileScope prof("edfdv_spline_bwd", stream);
  spline_push_bwd_kernel<<<(unsigned)((long long)batch * nx), 256, smem, stream>>>(p);
  return check_launch("spline_push_bwd_kernel");
}

// ---- adjoint of the Krook step  f' = f e^{-nu dt} + n f_mx (1 - e^{-nu dt}),  n = dv sum_j f_j  (fokker_planck.py:463-484)
//   f_bar_j = g_j e^{-nu dt} + dv (1 - e^{-nu dt}) sum_k g_k f_mx_k;   nu_bar = dt e^{-nu dt} sum_j g_j (n f_mx_j - f_j)
__global__ void __launch_bounds__(256) krook_bwd_kernel(const double* __restrict__ f, const double* __restrict__ g,
                                                        const double* __restrict__ nu_K, const double* __restrict__ f_mx,
                                                        int nv, double dv, double dt, double* __restrict__ fbar,
                                                        double* __restrict__ nubar) {
  __shared__ double red[3][8];
  const long long row = blockIdx.x;
  const double* fr = f + row * nv;
  const double* gr = g + row * nv;
  double s[3] = {0.0, 0.0, 0.0};  // sum f, sum g f_mx, sum g f
  for (int j = threadIdx.x; j < nv; j += blockDim.x) {
    const double fj = fr[j], gj = gr[j];
    s[0] += fj;
    s[1] = fma(gj, __ldg(f_mx + j), s[1]);
    s[2] = fma(gj, fj, s[2]);
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    s[k] = warp_sum(s[k]);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = s[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 3; k++) {
    s[k] = 0.0;
    for (int w = 0; w < 8; w++) s[k] += red[k][w];
  }
  const double ex = exp(-(dt * nu_K[row]));
  const double bias = dv * (1.0 - ex) * s[1];
  if (fbar) {
    double* out = fbar + row * nv;
    for (int j = threadIdx.x; j < nv; j += blockDim.x) out[j] = fma(gr[j], ex, bias);
  }
  if (nubar && threadIdx.x == 0) nubar[row] = dt * ex * ((s[0] * dv) * s[1] - s[2]);
}

int krook_bwd_f64(const double* f, const double* g, int batch, int nx, int nv, double dv, double dt, const double* nu_K,
                  const double* f_mx, double* fbar, double* nubar, cudaStream_t stream) {
  if (batch < 1 || nx < 1 || nv < 1) {
    set_last_error("krook_bwd: bad shape batch=%d nx=%d nv=%d", batch, nx, nv);
    return ADEPT_ERR_BAD_SHAPE;
  }
  ProfileScope prof("krook_bwd", stream);
  krook_bwd_kernel<<<(unsigned)((long long)batch * nx), 256, 0, stream>>>(f, g, nu_K, f_mx, nv, dv, dt, fbar, nubar);
  return check_launch("krook_bwd_kernel");
}

}  // namespace adept
