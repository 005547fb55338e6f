// O(nx) field kernels: spectral Poisson / Boltzmann-Poisson solve, ponderomotive force, wave-equation step.
// Reference semantics (file:line relative to /root/reference):
//   Poisson            adept/_vlasov1d/solvers/pushers/field.py:210-224  E = Re ifft(-i (1/kx) fft(rho))
//   Boltzmann-Poisson  adept/_vlasov1d/solvers/pushers/field.py:282-298
//   ponderomotive      adept/_vlasov1d/solvers/pushers/field.py:495      -0.5 * gradient(a^2, dx)[1:-1]
//   wave equation      adept/_vlasov1d/solvers/pushers/field.py:109-157  (WaveSolver + 2nd-order ABC)
#include "fft_core.cuh"
#include "internal.h"

namespace adept {

struct PoissonArgs {
  const double* rho;     // [batch, nx]
  const double* kmul;    // POISSON: one_over_kx[nx]; BOLTZMANN: kx[nx]   (per batch member if kmul_stride != 0)
  long long kmul_stride;
  double* e;             // [batch, nx]
  int mode;              // 0 poisson, 1 boltzmann
  double Te, lambda_De;  // boltzmann: lambda_De < 0 -> sqrt(Te / rho_0)
  const cplx* tw;
  int zero;  // always 0; opaque to ptxas (see FftPass)
};

template <int LOGN>
__global__ void __launch_bounds__(FftCfg<LOGN>::T) poisson_kernel(PoissonArgs p) {
  using C = FftCfg<LOGN>;
  constexpr int N = C::N, E = C::E, T = C::T;
  __shared__ __align__(16) cplx buf[C::BUF];
  __shared__ double rho0_s;
  const int t = threadIdx.x, tt = threadIdx.x;
  const double* rho = p.rho + (long long)blockIdx.x * N;
  const double* kmul = p.kmul + (long long)blockIdx.x * p.kmul_stride;
  double* eo = p.e + (long long)blockIdx.x * N;

  fft_prefetch_twiddles<LOGN>(p.tw, tt);
  cplx x[E];
#pragma unroll
  for (int m = 0; m < E; m++) x[m] = cmake(rho[tt + T * m], 0.0);
  fft_forward<LOGN>(x, buf, p.tw, tt, p.zero);
  if (t == 0) rho0_s = x[0].x / (double)N;  // mean(rho) = DC / N
  __syncthreads();
  const double rho0 = rho0_s;
#pragma unroll
  for (int m = 0; m < E; m++) {
    const int k = tt + T * m;
    double mult;
    if (p.mode == 0) {
      mult = kmul[k];
    } else {
      const double kx = kmul[k];
      const double lam_sq = p.lambda_De < 0.0 ? p.Te / rho0 : p.lambda_De * p.lambda_De;
      mult = kx * (p.Te / rho0) / (1.0 + lam_sq * kx * kx);
    }
    // Y = -i * mult * X = (mult*Xi, -mult*Xr); store swapped for inverse-by-forward
    const cplx y = cmake(mult * x[m].y, -(mult * x[m].x));
    x[m] = cmake(y.y, y.x);
  }
  if (C::NPASS > 1) __syncthreads();
  fft_forward<LOGN>(x, buf, p.tw, tt, p.zero);
#pragma unroll
  for (int m = 0; m < E; m++) eo[t + T * m] = x[m].y / (double)N;  // Re(ifft) = Im of swapped result
}

static int ilog2_exact(int n) {
  if (n < 2 || (n & (n - 1))) return -1;
  int l = 0;
  while ((1 << l) < n) l++;
  return l;
}

int poisson_f64(const double* rho, const double* kmul, long long kmul_stride, double* e, int batch, int nx, int mode,
                double Te, double lambda_De, cudaStream_t stream) {
  const int logn = ilog2_exact(nx);
  if (batch < 1 || logn < 1 || logn > 13) {
    set_last_error("poisson: nx=%d must be a power of two in [2, 8192] (batch=%d)", nx, batch);
    return logn < 1 || logn > 13 ? ADEPT_ERR_UNSUPPORTED : ADEPT_ERR_BAD_SHAPE;
  }
  PoissonArgs p = {rho, kmul, kmul_stride, e, mode, Te, lambda_De, get_twiddles(logn), 0};
  if (!p.tw) return ADEPT_ERR_CUDA;
  ProfileScope prof("poisson", stream);
  switch (logn) {
#define ADEPT_CASE(L)                                                                              \
  case L:                                                                                          \
    poisson_kernel<L><<<batch, FftCfg<L>::T, 0, stream>>>(p); \
    break;
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
#undef ADEPT_CASE
    default:
      // static shared memory is limited to 48 KiB: nx >= 4096 uses the dynamic-smem variant below
      return ADEPT_ERR_UNSUPPORTED;
  }
  return check_launch("poisson_kernel");
}

// Poisson / Boltzmann-Poisson solve of one member by the T threads of a CTA; `buf` is the FFT exchange buffer
// (FftCfg<LOGN>::BUF complex), `rho0_s` one shared double.  LDCG: rho may have been written by other CTAs of the same
// launch (fused field kernel), so the loads bypass L1.
template <int LOGN>
__device__ __forceinline__ void poisson_body(const PoissonArgs& p, const double* rho, const double* kmul, double* eo,
                                             cplx* buf, double* rho0_s) {
  using C = FftCfg<LOGN>;
  constexpr int N = C::N, E = C::E, T = C::T;
  const int t = threadIdx.x;
  fft_prefetch_twiddles<LOGN>(p.tw, t);
  cplx x[E];
#pragma unroll
  for (int m = 0; m < E; m++) x[m] = cmake(__ldcg(rho + t + T * m), 0.0);
  fft_forward<LOGN>(x, buf, p.tw, t, p.zero);
  if (t == 0) *rho0_s = x[0].x / (double)N;
  __syncthreads();
  const double rho0 = *rho0_s;
#pragma unroll
  for (int m = 0; m < E; m++) {
    const int k = t + T * m;
    double mult;
    if (p.mode == 0) {
      mult = kmul[k];
    } else {
      const double kx = kmul[k];
      const double lam_sq = p.lambda_De < 0.0 ? p.Te / rho0 : p.lambda_De * p.lambda_De;
      mult = kx * (p.Te / rho0) / (1.0 + lam_sq * kx * kx);
    }
    const cplx y = cmake(mult * x[m].y, -(mult * x[m].x));
    x[m] = cmake(y.y, y.x);
  }
  __syncthreads();
  fft_forward<LOGN>(x, buf, p.tw, t, p.zero);
#pragma unroll
  for (int m = 0; m < E; m++) eo[t + T * m] = x[m].y / (double)N;
}

// nx = 4096 / 8192: same solve with dynamic shared memory (> 48 KiB)
template <int LOGN>
__global__ void __launch_bounds__(FftCfg<LOGN>::T) poisson_kernel_big(PoissonArgs p) {
  constexpr int N = FftCfg<LOGN>::N;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double rho0_s;
  poisson_body<LOGN>(p, p.rho + (long long)blockIdx.x * N, p.kmul + (long long)blockIdx.x * p.kmul_stride,
                     p.e + (long long)blockIdx.x * N, reinterpret_cast<cplx*>(smem_raw), &rho0_s);
}

template <int LOGN>
static int launch_poisson_big(const PoissonArgs& p, int batch, cudaStream_t stream) {
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  const size_t smem = FftCfg<LOGN>::BUF * sizeof(cplx);
  if (dev < 64 && !configured[dev]) {
    cudaError_t err =
        cudaFuncSetAttribute(poisson_kernel_big<LOGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(poisson): %s", cudaGetErrorString(err));
      return ADEPT_ERR_CUDA;
    }
    configured[dev] = true;
  }
  ProfileScope prof("poisson", stream);
  poisson_kernel_big<LOGN><<<batch, FftCfg<LOGN>::T, smem, stream>>>(p);
  return check_launch("poisson_kernel_big");
}

int poisson_dispatch_f64(const double* rho, const double* kmul, long long kmul_stride, double* e, int batch, int nx,
                         int mode, double Te, double lambda_De, cudaStream_t stream) {
  const int logn = ilog2_exact(nx);
  if (logn < 0 && bluestein_supported(nx))
    return bluestein_poisson_f64(rho, kmul, kmul_stride, e, batch, nx, mode, Te, lambda_De, stream);
  if (logn < 0 && bigx_supported(nx, 64))  // long mixed-length grids (nx = 17280 = 128 x 135, ...)
    return bigx_poisson_f64(rho, kmul, kmul_stride, e, batch, nx, mode, Te, lambda_De, stream);
  if (logn == 12 || logn == 13) {
    PoissonArgs p = {rho, kmul, kmul_stride, e, mode, Te, lambda_De, get_twiddles(logn), 0};
    if (!p.tw) return ADEPT_ERR_CUDA;
    return logn == 12 ? launch_poisson_big<12>(p, batch, stream) : launch_poisson_big<13>(p, batch, stream);
  }
  return poisson_f64(rho, kmul, kmul_stride, e, batch, nx, mode, Te, lambda_De, stream);
}

// ---- ponderomotive force: pond_i = -0.5 * (a_{i+2}^2 - a_i^2) / (2 dx), a has nx+2 cells -----------------------
__global__ void pond_kernel(const double* __restrict__ a, double* __restrict__ pond, int nx, double dx,
                            long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long b = idx / nx;
  const int i = (int)(idx % nx);
  const double* ab = a + b * (nx + 2);
  const double lo = __dmul_rn(ab[i], ab[i]), hi = __dmul_rn(ab[i + 2], ab[i + 2]);
  pond[idx] = __dmul_rn(-0.5, __ddiv_rn(__dsub_rn(hi, lo), __dmul_rn(2.0, dx)));
}

int ponderomotive_f64(const double* a, double* pond, int batch, int nx, double dx, cudaStream_t stream) {
  if (batch < 1 || nx < 1) {
    set_last_error("ponderomotive: bad shape batch=%d nx=%d", batch, nx);
    return ADEPT_ERR_BAD_SHAPE;
  }
  const long long total = (long long)batch * nx;
  ProfileScope prof("ponderomotive", stream);
  pond_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(a, pond, nx, dx, total);
  return check_launch("pond_kernel");
}

// ---- wave equation step ----------------------------------------------------------------------------------
struct WaveArgs {
  const double* a;      // [batch, nx+2]
  const double* aold;   // [batch, nx+2]
  const double* djy;    // [batch, nx+2]
  const double* ne_n;   // [batch, nx] electron charge density at t_n       (nullable -> 0)
  const double* ne_np1; // [batch, nx] electron charge density at t_{n+1}   (nullable -> 0)
  double* a_new;        // [batch, nx+2]
  int nx;
  double c, dx, dt;
};

__device__ __forceinline__ double wave_interior(const WaveArgs& p, const double* a, const double* aold,
                                                const double* djy, const double* n0, const double* n1, int i) {
  // i in [0, nx): anew[i] of field.py:146-153 (a index i+1)
  const double d2dx2 = (a[i] - 2.0 * a[i + 1] + a[i + 2]) / (p.dx * p.dx);
  double ed = 0.0;
  if (n0 && n1) ed = -0.5 * (n0[i] + n1[i]);  // vector_field.py:346
  return 2.0 * a[i + 1] - aold[i + 1] + (p.dt * p.dt) * ((p.c * p.c) * d2dx2 - ed * a[i + 1] + djy[i + 1]);
}

__global__ void wave_kernel(WaveArgs p) {
  const int nx = p.nx;
  const long long b = blockIdx.y;
  const double* a = p.a + b * (nx + 2);
  const double* aold = p.aold + b * (nx + 2);
  const double* djy = p.djy + b * (nx + 2);
  const double* n0 = p.ne_n ? p.ne_n + b * nx : nullptr;
  const double* n1 = p.ne_np1 ? p.ne_np1 + b * nx : nullptr;
  double* out = p.a_new + b * (nx + 2);
  const int j = blockIdx.x * blockDim.x + threadIdx.x;  // index into a_new, 0..nx+1
  if (j > nx + 1) return;
  if (j >= 1 && j <= nx) {
    out[j] = wave_interior(p, a, aold, djy, n0, n1, j - 1);
    return;
  }
  const double c_over_dx = p.c / p.dx;
  const double cst = c_over_dx * p.dt;
  const double ooc = 1.0 / p.dt / c_over_dx;
  const double coeff = -1.0 / (ooc + 2.0 + cst);
  if (j == 0) {
    const double an0 = wave_interior(p, a, aold, djy, n0, n1, 0), an1 = wave_interior(p, a, aold, djy, n0, n1, 1);
    double al = (ooc - 2.0 + cst) * (an1 + aold[0]);
    al += 2.0 * (cst - ooc) * (a[0] + a[2] - an0 - aold[1]);
    al -= 4.0 * (ooc + cst) * a[1];
    al *= coeff;
    al -= aold[2];
    out[0] = al;
  } else {
    const double anm1 = wave_interior(p, a, aold, djy, n0, n1, nx - 1),
                 anm2 = wave_interior(p, a, aold, djy, n0, n1, nx - 2);
    double ar = (ooc - 2.0 + cst) * (anm2 + aold[nx + 1]);
    ar += 2.0 * (cst - ooc) * (a[nx + 1] + a[nx - 1] - anm1 - aold[nx]);
    ar -= 4.0 * (ooc + cst) * a[nx];
    ar *= coeff;
    ar -= aold[nx - 1];
    out[nx + 1] = ar;
  }
}

int wave_step_f64(const double* a, const double* aold, const double* djy, const double* ne_n, const double* ne_np1,
                  double* a_new, int batch, int nx, double c, double dx, double dt, cudaStream_t stream) {
  if (batch < 1 || nx < 2) {
    set_last_error("wave_step: bad shape batch=%d nx=%d", batch, nx);
    return ADEPT_ERR_BAD_SHAPE;
  }
  if (a_new == a || a_new == aold) {
    set_last_error("wave_step: a_new must not alias a or aold");
    return ADEPT_ERR_BAD_ARG;
  }
  WaveArgs p = {a, aold, djy, ne_n, ne_np1, a_new, nx, c, dx, dt};
  dim3 grid((nx + 2 + 255) / 256, batch);
  ProfileScope prof("wave_step", stream);
  wave_kernel<<<grid, 256, 0, stream>>>(p);
  return check_launch("wave_kernel");
}


// ---- fused field solve for one large grid (batch == 1) ------------------------------------------------------------
// One launch replaces ex_driver + ponderomotive + reduce_parts (per species) + Poisson of the leapfrog step
// (vector_field.py:87-95 calling field.py:479-497, 197-224, 21-33): every CTA finishes the charge density of 64 grid
// points from the per-CTA partial sums the x-advection left behind (fixed summation order), evaluates the
// ponderomotive force and the driver field there, then all CTAs solve Poisson's equation together (four-step transform
// with device-wide barriers on a ticket counter; the nx/64 <= 64 CTAs are co-resident on any B200).  The counter is
// re-armed by the CTA that exits last, so consecutive launches on one stream need no host work; it must be zero before
// the first launch.
struct FieldFusedArgs {
  int nsp;
  const double* parts[4];
  int nparts[4];
  double dv[4], charge[4];
  const double* base;  // static background (nullable)
  double* rho;
  int nx;
  const double* a;  // [nx + 2]
  double* pond;
  double dx;
  int n_ex;  // 0: the driver field is not evaluated here
  const double* ex_space;
  const double* ex_kx;
  double* dex;
  double ex_w[8], ex_a0[8], ex_tenv[8], ex_wt[8];
  const double* trow;  // nullable device-resident time row (common.cuh)
  PoissonArgs po;
  unsigned int* counter;
  double* scratch;  // 4 nx doubles: two nx-long complex transposition buffers of the distributed solve
};

// sum over the 4 lanes q = 0..3 that share one output
__device__ __forceinline__ cplx quad_sum(cplx v) {
  v.x += __shfl_xor_sync(0xffffffffu, v.x, 1), v.y += __shfl_xor_sync(0xffffffffu, v.y, 1);
  v.x += __shfl_xor_sync(0xffffffffu, v.x, 2), v.y += __shfl_xor_sync(0xffffffffu, v.y, 2);
  return v;
}

// The Poisson solve is spread over the M = nx/64 CTAs as a four-step transform, N = M x 64 (n = 64 n1 + n2,
// k = k1 + M k2):  stage A (CTA j, its share of the n2): length-M DFTs over n1;  barrier;  stage B (CTA k1): twiddle
// W_N^(n2 k1), 64-point DFT over n2 -> X[k1 + M k2], the -i kmul multiply, and at once the transposed inverse over k2;
// barrier;  stage A' (CTA j): inverse length-M DFTs over k1 -> E.  Two device-wide barriers instead of two
// 4096-point FFTs in one CTA (about 15 us with 147 SMs idle).  The small DFTs are direct sums (<= 64 terms).
__global__ void __launch_bounds__(256) field_fused_kernel(FieldFusedArgs p) {
  constexpr int NG = 4;  // thread groups that split the partial-sum rows
  __shared__ double sm[4][64];
  __shared__ cplx w64[64], wM[64], vin[64], vout[64], twn[64];
  __shared__ double red_s[8];
  const int r = threadIdx.x & 63, g = threadIdx.x >> 6;
  const int i = blockIdx.x * 64 + r;
  const int n = p.nx;
  const int M = gridDim.x;  // nx / 64
  double acc = p.base ? p.base[i] : 0.0;
  for (int k = 0; k < p.nsp; k++) {
    const double* col = p.parts[k] + i;
    const int np = p.nparts[k];
    double s[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int q = g; q < np; q += 24 * NG) {  // 24 independent L2 loads in flight per thread (148 rows: 2 round trips)
      double x[24];
#pragma unroll
      for (int u = 0; u < 24; u++) x[u] = (q + u * NG < np) ? __ldcg(col + (size_t)(q + u * NG) * n) : 0.0;
#pragma unroll
      for (int u = 0; u < 24; u++) s[u & 7] += x[u];
    }
    sm[g][r] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
    __syncthreads();
    if (g == 0) {
      const double tot = (sm[0][r] + sm[1][r]) + (sm[2][r] + sm[3][r]);
      const double term = __dmul_rn(p.charge[k], __dmul_rn(tot, p.dv[k]));
      acc = (k == 0 && !p.base) ? term : __dadd_rn(acc, term);
    }
    __syncthreads();
  }
  if (g == 0) {
    p.rho[i] = acc;
    // ponderomotive force, field.py:495
    const double lo = __dmul_rn(p.a[i], p.a[i]), hi = __dmul_rn(p.a[i + 2], p.a[i + 2]);
    p.pond[i] = __dmul_rn(-0.5, __ddiv_rn(__dsub_rn(hi, lo), __dmul_rn(2.0, p.dx)));
  }
  if (p.n_ex > 0 && g == 1) {  // driver field at the first substep time, field.py:21-33
    double total = 0.0;
    for (int d = 0; d < p.n_ex; d++) {
      const double tenv = p.trow ? p.trow[TROW_TENV + d] : p.ex_tenv[d];
      const double wt = p.trow ? p.trow[TROW_WT + d] : p.ex_wt[d];
      const double factor = __dmul_rn(tenv, p.ex_space[(size_t)d * n + i]);
      const double amp = __dmul_rn(__dmul_rn(factor, p.ex_w[d]), p.ex_a0[d]);
      total = __dadd_rn(total, __dmul_rn(amp, sin(__dsub_rn(p.ex_kx[(size_t)d * n + i], wt))));
    }
    p.dex[i] = total;
  }
  // tables while the other CTAs finish their densities: W_64^j, W_M^j, and this CTA's twiddles W_N^(n2 k1), k1 = CTA
  if (g == 2) {
    double sn, cs;
    sincospi(-2.0 * (double)r / 64.0, &sn, &cs);
    w64[r] = cmake(cs, sn);
    sincospi(-2.0 * (double)(r & (M - 1)) / (double)M, &sn, &cs);
    wM[r] = cmake(cs, sn);
  } else if (g == 3) {
    double sn, cs;
    sincospi(-2.0 * (double)(r * (int)blockIdx.x) / (double)n, &sn, &cs);
    twn[r] = cmake(cs, sn);
  }
  const unsigned int G = gridDim.x;
  grid_barrier(p.counter, G);  // rho complete; the partial-sum rows are dead from here on

  cplx* S1 = reinterpret_cast<cplx*>(p.scratch);  // [M][64]  stage A output, A[k1][n2]
  cplx* S2 = S1 + n;                              // [64][M]  stage B' output, P[n2][k1]
  const int o = threadIdx.x >> 2, q4 = threadIdx.x & 3;
  const int S = 64 / M;  // n2 values this CTA owns in stages A / A'
  double rho0 = 0.0;
  if (p.po.mode != 0) {  // Boltzmann electrons need mean(rho): every CTA sums the grid itself (nx <= 4096)
    double t = 0.0;
    for (int j = threadIdx.x; j < n; j += 256) t += __ldcg(p.rho + j);
    t = warp_sum(t);
    if ((threadIdx.x & 31) == 0) red_s[threadIdx.x >> 5] = t;
    __syncthreads();
    for (int w = 0; w < 8; w++) rho0 += red_s[w];
    rho0 /= (double)n;
  }
  // ---- stage A: A[k1][n2] = sum_n1 rho[64 n1 + n2] W_M^(n1 k1), output o -> (n2 = j S + o / M, k1 = o % M) ----
  // (all loads of a stage are issued before the sums: the stages are latency chains of L2 round trips otherwise)
  {
    const int n2 = (int)blockIdx.x * S + o / M, k1 = o & (M - 1);
    double xv[16];
#pragma unroll
    for (int u = 0; u < 16; u++) xv[u] = (q4 + 4 * u < M) ? __ldcg(p.rho + 64 * (q4 + 4 * u) + n2) : 0.0;
    cplx a = cmake(0.0, 0.0);
#pragma unroll
    for (int u = 0; u < 16; u++) {
      const cplx w = wM[((q4 + 4 * u) * k1) & (M - 1)];
      a.x = fma(xv[u], w.x, a.x), a.y = fma(xv[u], w.y, a.y);
    }
    a = quad_sum(a);
    if (q4 == 0) S1[(size_t)k1 * 64 + n2] = a;
  }
  // the multiplier of this thread's mode k = k1 + M k2 (k1 = CTA, k2 = o) does not depend on the transform
  const double kmul_k = __ldg(p.po.kmul + (int)blockIdx.x + M * o);
  grid_barrier(p.counter, 2 * G);
  // ---- stage B (k1 = CTA): X[k1 + M k2] = sum_n2 (A[k1][n2] W_N^(n2 k1)) W_64^(n2 k2);  Y = -i mult X;
  //      stage B': P[n2] = conj(W_N^(n2 k1)) sum_k2 Y[k2] conj(W_64^(n2 k2)) ----
  {
    const int k1 = blockIdx.x;
    if (threadIdx.x < 64) {
      const cplx v = __ldcg(reinterpret_cast<const double2*>(S1 + (size_t)k1 * 64 + threadIdx.x));
      vin[threadIdx.x] = cmul(v, twn[threadIdx.x]);
    }
    __syncthreads();
    cplx x0 = cmake(0.0, 0.0), x1 = cmake(0.0, 0.0);
#pragma unroll
    for (int u = 0; u < 16; u += 2) {
      const int na = q4 + 4 * u, nb = na + 4;
      const cplx wa = w64[(na * o) & 63], va = vin[na], wb = w64[(nb * o) & 63], vb = vin[nb];
      x0.x += va.x * wa.x - va.y * wa.y, x0.y += va.x * wa.y + va.y * wa.x;
      x1.x += vb.x * wb.x - vb.y * wb.y, x1.y += vb.x * wb.y + vb.y * wb.x;
    }
    const cplx x = quad_sum(cadd(x0, x1));
    if (q4 == 0) {
      double mult = kmul_k;
      if (p.po.mode != 0) {
        const double kx = kmul_k;
        const double lam_sq = p.po.lambda_De < 0.0 ? p.po.Te / rho0 : p.po.lambda_De * p.po.lambda_De;
        mult = kx * (p.po.Te / rho0) / (1.0 + lam_sq * kx * kx);
      }
      vout[o] = cmake(mult * x.y, -(mult * x.x));  // -i mult X
    }
    __syncthreads();
    cplx y0 = cmake(0.0, 0.0), y1 = cmake(0.0, 0.0);
#pragma unroll
    for (int u = 0; u < 16; u += 2) {  // o = n2 here; multiplies by conj(W_64)
      const int ka = q4 + 4 * u, kb = ka + 4;
      const cplx wa = w64[(ka * o) & 63], va = vout[ka], wb = w64[(kb * o) & 63], vb = vout[kb];
      y0.x += va.x * wa.x + va.y * wa.y, y0.y += va.y * wa.x - va.x * wa.y;
      y1.x += vb.x * wb.x + vb.y * wb.y, y1.y += vb.y * wb.x - vb.x * wb.y;
    }
    const cplx y = quad_sum(cadd(y0, y1));
    if (q4 == 0) {
      const cplx w = twn[o];
      S2[(size_t)o * M + k1] = cmake(y.x * w.x + y.y * w.y, y.y * w.x - y.x * w.y);  // times conj(W_N^(n2 k1))
    }
  }
  grid_barrier(p.counter, 3 * G);
  // ---- stage A': E[64 n1 + n2] = Re sum_k1 P[n2][k1] conj(W_M^(n1 k1)) / N, output o -> (n2, n1 = o % M) ----
  {
    const int n2 = (int)blockIdx.x * S + o / M, n1 = o & (M - 1);
    cplx pv[16];
#pragma unroll
    for (int u = 0; u < 16; u++)
      pv[u] = (q4 + 4 * u < M) ? __ldcg(reinterpret_cast<const double2*>(S2 + (size_t)n2 * M + q4 + 4 * u))
                               : cmake(0.0, 0.0);
    double e = 0.0;
#pragma unroll
    for (int u = 0; u < 16; u++) {
      const cplx w = wM[(n1 * (q4 + 4 * u)) & (M - 1)];
      e += pv[u].x * w.x + pv[u].y * w.y;  // Re(v conj(w))
    }
    e += __shfl_xor_sync(0xffffffffu, e, 1);
    e += __shfl_xor_sync(0xffffffffu, e, 2);
    if (q4 == 0) p.po.e[64 * n1 + n2] = e / (double)n;
  }
  // exit tickets: the CTA that leaves last re-arms the counter for the next launch
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(p.counter, 1u) == 4 * G - 1) *p.counter = 0u;
}

// ---- Poisson solve as a circular convolution with the Green's function, spread over nx/32 CTAs ------------------------
// E = Re ifft(-i (1/kx) fft(rho)) is linear in rho (field.py:221-224): E_i = sum_j green[(i - j) mod nx] rho_j with
// green = Re ifft(-i / kx).  One 4096-point FFT solve keeps a single SM busy for ~15-19 us with 147 SMs idle; the direct
// sum is nx^2 fused multiply-adds spread over the whole GPU.  One warp per output, lanes stride j, four running sums
// (the same arithmetic as the field tail of the x-advection, vdfdx_tma.cu).
__global__ void __launch_bounds__(1024) poisson_green_kernel(const double* __restrict__ rho,
                                                             const double* __restrict__ green, long long green_stride,
                                                             double* __restrict__ e, int nx) {
  // Shared memory: G2[2 nx] (the Green's function twice, so that a window never wraps), rho[nx], partial[8][32].
  // The direct sum reads 16 bytes of shared memory per multiply-add when every product fetches its own operands; that
  // is 2.1 M wavefronts at nx = 4096, ~8 us over 148 SMs.  Register tile instead: a lane owns two consecutive j and
  // eight consecutive outputs, the 9 Green's-function values it needs come from 5 aligned 16-byte loads:
  // 6 loads per 16 multiply-adds.
  extern __shared__ __align__(16) double sm_pg[];
  double* g2 = sm_pg;
  double* rho_s = sm_pg + 2 * nx;
  double* part = rho_s + nx;
  const int b = blockIdx.y;
  const double* rb = rho + (long long)b * nx;
  const double* gb = green + (long long)b * green_stride;
  {  // all loads of the CTA in flight at once (nx <= 8192: at most 8 per thread and array), then the stores
    double rv[8], gv[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int j = threadIdx.x + u * 1024;
      rv[u] = j < nx ? rb[j] : 0.0;
      gv[u] = j < nx ? __ldg(gb + j) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int j = threadIdx.x + u * 1024;
      if (j < nx) rho_s[j] = rv[u], g2[j] = gv[u], g2[j + nx] = gv[u];
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int og = warp & 3, jr = warp >> 2;       // output group (8 outputs), j range (nx / 8 values)
  const int i0 = blockIdx.x * 32 + 8 * og;
  const int jn = nx >> 3;
  double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int j = jr * jn + 2 * lane; j < (jr + 1) * jn; j += 64) {
    const double2 r2 = *reinterpret_cast<const double2*>(rho_s + j);
    // G2 index of (output i0 + r, column j + d): i0 + r - j - d + nx = m0 + r + 2 - d with m0 = i0 - j - 2 + nx (even)
    const double2* wp = reinterpret_cast<const double2*>(g2 + (i0 - j - 2 + nx));
    double w[10];
#pragma unroll
    for (int q = 0; q < 5; q++) {
      const double2 t2 = wp[q];
      w[2 * q] = t2.x, w[2 * q + 1] = t2.y;
    }
#pragma unroll
    for (int r = 0; r < 8; r++) acc[r] = fma(w[r + 1], r2.y, fma(w[r + 2], r2.x, acc[r]));
  }
#pragma unroll
  for (int r = 0; r < 8; r++) acc[r] = warp_sum(acc[r]);
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < 8; r++) part[jr * 32 + og * 8 + r] = acc[r];
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 8; q++) s += part[q * 32 + threadIdx.x];
    e[(long long)b * nx + blockIdx.x * 32 + threadIdx.x] = s;
  }
}

int poisson_green_f64(const double* rho, const double* green, long long green_stride, double* e, int batch, int nx,
                      cudaStream_t stream) {
  if (batch < 1 || nx < 512 || nx > 8192 || (nx & (nx - 1))) {
    set_last_error("poisson_green: nx=%d must be a power of two in [512, 8192] (batch=%d)", nx, batch);
    return batch < 1 ? ADEPT_ERR_BAD_SHAPE : ADEPT_ERR_UNSUPPORTED;
  }
  const size_t smem = ((size_t)3 * nx + 8 * 32) * sizeof(double);
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (smem > 48 * 1024 && dev < 64 && !configured[dev]) {
    cudaError_t err = cudaFuncSetAttribute(poisson_green_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (err != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(poisson_green): %s", cudaGetErrorString(err));
      return ADEPT_ERR_CUDA;
    }
    configured[dev] = true;
  }
  ProfileScope prof("poisson_green", stream);
  poisson_green_kernel<<<dim3(nx / 32, batch), 1024, smem, stream>>>(rho, green, green_stride, e, nx);
  return check_launch("poisson_green_kernel");
}

// ---- small grids / ensembles: the whole field solve of one member in one CTA ----------------------------------------
// Replaces, for nx <= 256 (one launch instead of 3 + n_species): pond_kernel, moments_kernel per species, poisson_kernel
// and, for the leapfrog step, ex_driver_kernel.  A 64 x 512 member is 256 KB: an ensemble of them lives in L2, the step
// is bound by launch latency and dependent round trips, not by bandwidth (BASELINE.json configs[1] and [3]).
// Velocity sums: one warp per row with the lane striding and rounding of moments_kernel (bit-identical densities).
// Poisson: direct length-nx DFT sums against a W_N table in shared memory (<= 256 terms), then the reference's
// Re(ifft(-i mult fft(rho))) (field.py:221-224, 293-298) with every mode kept.
struct FieldMemberArgs {
  int nsp;
  const double* f[4];
  int nv[4];
  double dv[4], charge[4];
  const double* base;  // [batch, nx] static background (nullable)
  double* rho;         // [batch, nx]
  int nx;
  const double* a;     // [batch, nx + 2]
  double* pond;
  double dx;
  int n_ex;            // 0: the driver field is not evaluated here
  long long n_rows;    // batch * nx (row stride of the driver tables)
  const double* ex_space;
  const double* ex_kx;
  double* dex;
  double ex_w[8], ex_a0[8], ex_tenv[8], ex_wt[8];
  const double* ex_w_row;
  const double* ex_a0_row;
  double ex_t0;
  const double* trow;  // nullable device-resident time row (common.cuh)
  const double* kmul;
  long long kmul_stride;
  double* e;
  int mode;
  double Te, lambda_De;
};

__global__ void __launch_bounds__(1024) field_member_kernel(FieldMemberArgs p) {
  __shared__ double rho_s[256];
  __shared__ cplx w_s[256], y_s[256];
  const int N = p.nx;
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int per_round = blockDim.x >> 2;  // outputs per round of the DFT sums (4 lanes each)
  const long long row0 = (long long)b * N;
  if (threadIdx.x < N) {
    double sn, cs;
    sincospi(-2.0 * (double)threadIdx.x / (double)N, &sn, &cs);
    w_s[threadIdx.x] = cmake(cs, sn);
  }
  // ---- charge density, field.py:197-208 ----
  for (int i = warp; i < N; i += nwarp) {
    double acc = p.base ? p.base[row0 + i] : 0.0;
    for (int k = 0; k < p.nsp; k++) {
      const int nv = p.nv[k];
      const double* fr = p.f[k] + (row0 + i) * nv;
      double s0 = 0.0;
      if ((nv & 1) == 0 && ((reinterpret_cast<uintptr_t>(fr) & 15) == 0)) {
        const double2* f2 = reinterpret_cast<const double2*>(fr);
        for (int j = lane; j < (nv >> 1); j += 32) {
          const double2 x = f2[j];
          s0 += x.x + x.y;
        }
      } else {
        for (int j = lane; j < nv; j += 32) s0 += fr[j];
      }
      s0 = warp_sum(s0);
      const double term = __dmul_rn(p.charge[k], __dmul_rn(s0, p.dv[k]));
      acc = (k == 0 && !p.base) ? term : __dadd_rn(acc, term);
    }
    if (lane == 0) {
      rho_s[i] = acc;
      p.rho[row0 + i] = acc;
    }
  }
  // ---- ponderomotive force (field.py:495) and the driver field at the first substep time (field.py:21-33) ----
  if (threadIdx.x < N) {
    const int i = threadIdx.x;
    const double* a = p.a + (long long)b * (N + 2);
    const double lo = __dmul_rn(a[i], a[i]), hi = __dmul_rn(a[i + 2], a[i + 2]);
    p.pond[row0 + i] = __dmul_rn(-0.5, __ddiv_rn(__dsub_rn(hi, lo), __dmul_rn(2.0, p.dx)));
    if (p.n_ex > 0) {
      double total = 0.0;
      for (int d = 0; d < p.n_ex; d++) {
        const long long o = d * p.n_rows + row0 + i;
        const double factor = __dmul_rn(p.trow ? p.trow[TROW_TENV + d] : p.ex_tenv[d], p.ex_space[o]);
        const double w = p.ex_w_row ? p.ex_w_row[o] : p.ex_w[d];
        const double a0 = p.ex_a0_row ? p.ex_a0_row[o] : p.ex_a0[d];
        const double wt = p.ex_w_row ? __dmul_rn(w, p.trow ? p.trow[TROW_EX_T] : p.ex_t0)
                                     : (p.trow ? p.trow[TROW_WT + d] : p.ex_wt[d]);
        const double amp = __dmul_rn(__dmul_rn(factor, w), a0);
        total = __dadd_rn(total, __dmul_rn(amp, sin(__dsub_rn(p.ex_kx[o], wt))));
      }
      p.dex[row0 + i] = total;
    }
  }
  __syncthreads();
  // ---- X_k = sum_j rho_j W^(jk); 4 lanes per output, blockDim / 4 outputs per round ----
  const int o = threadIdx.x >> 2, q4 = threadIdx.x & 3;
  const int mask = N - 1;
  double rho0 = 0.0;
  if (p.mode != 0) {  // Boltzmann electrons: mean(rho)
    for (int j = 0; j < N; j++) rho0 += rho_s[j];
    rho0 /= (double)N;
  }
  const double* kmul = p.kmul + (long long)b * p.kmul_stride;
  for (int k0 = 0; k0 < N; k0 += per_round) {
    const int k = k0 + o;
    cplx x = cmake(0.0, 0.0);
    if (k < N) {
      for (int j = q4; j < N; j += 4) {
        const cplx w = w_s[(j * k) & mask];
        const double r = rho_s[j];
        x.x = fma(r, w.x, x.x), x.y = fma(r, w.y, x.y);
      }
    }
    x = quad_sum(x);
    if (k < N && q4 == 0) {
      double mult = kmul[k];
      if (p.mode != 0) {
        const double kx = mult;
        const double lam_sq = p.lambda_De < 0.0 ? p.Te / rho0 : p.lambda_De * p.lambda_De;
        mult = kx * (p.Te / rho0) / (1.0 + lam_sq * kx * kx);
      }
      y_s[k] = cmake(mult * x.y, -(mult * x.x));  // -i mult X
    }
  }
  __syncthreads();
  // ---- E_i = Re sum_k Y_k conj(W^(ik)) / N ----
  for (int i0 = 0; i0 < N; i0 += per_round) {
    const int i = i0 + o;
    double e = 0.0;
    if (i < N) {
      for (int k = q4; k < N; k += 4) {
        const cplx w = w_s[(i * k) & mask], y = y_s[k];
        e += y.x * w.x + y.y * w.y;
      }
    }
    e += __shfl_xor_sync(0xffffffffu, e, 1);
    e += __shfl_xor_sync(0xffffffffu, e, 2);
    if (i < N && q4 == 0) p.e[row0 + i] = e / (double)N;
  }
}

bool field_member_supported(int nx) { return nx >= 4 && nx <= 256 && (nx & (nx - 1)) == 0; }

int field_member_f64(int nsp, const double* const* f, const int* nv, const double* dv, const double* charge,
                     const double* base, double* rho, int batch, int nx, const double* a, double* pond, double dx,
                     int n_ex, const double* ex_space, const double* ex_kx, double* dex, const double* ex_w,
                     const double* ex_a0, const double* ex_tenv, const double* ex_wt, const double* ex_w_row,
                     const double* ex_a0_row, double ex_t0, const double* kmul, long long kmul_stride, double* e,
                     int mode, double Te, double lambda_De, cudaStream_t stream) {
  if (!field_member_supported(nx) || batch < 1 || nsp < 1 || nsp > 4 || n_ex < 0 || n_ex > 8) {
    set_last_error("field_member: unsupported batch=%d nx=%d n_species=%d n_ex=%d", batch, nx, nsp, n_ex);
    return ADEPT_ERR_UNSUPPORTED;
  }
  FieldMemberArgs p = {};
  p.nsp = nsp;
  for (int k = 0; k < nsp; k++) p.f[k] = f[k], p.nv[k] = nv[k], p.dv[k] = dv[k], p.charge[k] = charge[k];
  p.base = base, p.rho = rho, p.nx = nx, p.a = a, p.pond = pond, p.dx = dx;
  p.n_ex = n_ex, p.n_rows = (long long)batch * nx, p.ex_space = ex_space, p.ex_kx = ex_kx, p.dex = dex;
  for (int d = 0; d < n_ex; d++) p.ex_w[d] = ex_w[d], p.ex_a0[d] = ex_a0[d], p.ex_tenv[d] = ex_tenv[d], p.ex_wt[d] = ex_wt[d];
  p.trow = current_time_row();
  p.ex_w_row = ex_w_row, p.ex_a0_row = ex_a0_row, p.ex_t0 = ex_t0;
  p.kmul = kmul, p.kmul_stride = kmul_stride, p.e = e, p.mode = mode, p.Te = Te, p.lambda_De = lambda_De;
  ProfileScope prof("field_member", stream);
  // few members: latency-bound, 32 warps per member shorten the chain of L2 round trips (121 members: 18.3 -> 14.0 us);
  // many members: bandwidth-bound, several small CTAs per SM stream better (1024 members: 73 us against 89 us)
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int threads = (nx >= 32 && batch < 2 * sms) ? 1024 : 256;
  field_member_kernel<<<batch, threads, 0, stream>>>(p);
  return check_launch("field_member_kernel");
}

bool field_fused_supported(int batch, int nx) { return batch == 1 && (nx == 1024 || nx == 2048 || nx == 4096); }

int field_fused_f64(int nsp, const double* const* parts, const int* nparts, const double* dv, const double* charge,
                    const double* base, double* rho, int nx, const double* a, double* pond, double dx, int n_ex,
                    const double* ex_space, const double* ex_kx, double* dex, const double* ex_w, const double* ex_a0,
                    const double* ex_tenv, const double* ex_wt, const double* kmul, double* e, int mode, double Te,
                    double lambda_De, unsigned int* counter, cudaStream_t stream) {
  if (!field_fused_supported(1, nx) || nsp < 1 || nsp > 4 || n_ex < 0 || n_ex > 8 || !counter) {
    set_last_error("field_fused: unsupported nx=%d / n_species=%d / n_ex=%d", nx, nsp, n_ex);
    return ADEPT_ERR_UNSUPPORTED;
  }
  FieldFusedArgs p = {};
  p.nsp = nsp;
  for (int k = 0; k < nsp; k++) p.parts[k] = parts[k], p.nparts[k] = nparts[k], p.dv[k] = dv[k], p.charge[k] = charge[k];
  if (nparts[0] < 4) {
    set_last_error("field_fused: needs at least 4 partial-sum rows as scratch (got %d)", nparts[0]);
    return ADEPT_ERR_UNSUPPORTED;
  }
  p.scratch = const_cast<double*>(parts[0]);  // rows 0..3 of the first partial-sum array, dead once rho is complete
  p.base = base, p.rho = rho, p.nx = nx, p.a = a, p.pond = pond, p.dx = dx;
  p.n_ex = n_ex, p.ex_space = ex_space, p.ex_kx = ex_kx, p.dex = dex;
  for (int d = 0; d < n_ex; d++) p.ex_w[d] = ex_w[d], p.ex_a0[d] = ex_a0[d], p.ex_tenv[d] = ex_tenv[d], p.ex_wt[d] = ex_wt[d];
  p.trow = current_time_row();
  p.po = PoissonArgs{rho, kmul, 0, e, mode, Te, lambda_De, nullptr, 0};
  p.counter = counter;
  ProfileScope prof("field_fused", stream);
  // cooperative launch: the ticket barriers need all nx/64 CTAs resident whatever other streams are running
  void* args[1] = {&p};
  cudaError_t err = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(field_fused_kernel), dim3(nx / 64), dim3(256),
                                                args, 0, stream);
  if (err != cudaSuccess) {
    set_last_error("cudaLaunchCooperativeKernel(field_fused): %s", cudaGetErrorString(err));
    (void)cudaGetLastError();
    return ADEPT_ERR_CUDA;
  }
  return check_launch("field_fused_kernel");
}

}  // namespace adept
