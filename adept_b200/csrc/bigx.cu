// x-direction spectral operators for LONG pencils whose length is not a power of two: nx = N1 * N2 with N1 = 64 / 128 /
// 256 and N2 <= 150 -- e.g. nx = 17280 = 128 * 135 of configs/vlasov-1d/iaw-turbulence-big*.yaml, whose x-pencil
// (276 KB per column pair) does not fit one SM's shared memory.
//
// Reference semantics: out = irfft(M rfft(in, axis=x), axis=x) per column, with M the advection phase
// exp(-i kx v dt) (SpaceExponential, adept/_vlasov1d/solvers/pushers/vlasov.py:234-251) or a field-solve symbol
// (SpectralPoissonSolver / BoltzmannPoissonSolver, adept/_vlasov1d/solvers/pushers/field.py:210-224, 282-298).
//
// Cooley-Tukey split n = N2 n1 + n2, k = k1 + N1 k2, three launches through a scratch array Y of the size of f:
//   K1  N1-point FFTs over n1 (the power-of-two core of fft_core.cuh, 32 column pairs interleaved per CTA so that
//       every global and shared access of a warp is contiguous), twiddle W_N^(n2 k1)           f  -> Y[n2, k1, col]
//   K2  per pair of groups (k1, N1 - k1): N2-point DFTs over n2 (direct sums against a shared-memory table of W_N2),
//       two-for-one separation of the two real columns, multiplier, recombination, inverse DFTs over k2,
//       conjugate twiddle                                                                       Y  -> Y (in place)
//   K3  inverse N1-point FFTs over k1 (swap . forward . swap)                                   Y  -> f_out
// Every thread block reads and writes whole 512-byte row segments.  The N2-point transforms are one Cooley-Tukey level
// N2 = P Q of direct sums (135 = 9 x 15: 24 multiply-adds per output instead of 135).
#include <math.h>

#include <mutex>
#include <vector>

#include "fft_core.cuh"
#include "internal.h"

namespace adept {

namespace {

constexpr int BX_COLS = 32;   // column pairs per CTA
constexpr int BX_MAXK = 10;   // N2 <= 16 * BX_MAXK
constexpr int BX_MAXN2 = 150; // three [N2][32] complex arrays + the table must fit 227 KB of shared memory

struct BigXArgs {
  const double* in;
  double* out;
  int N1, N2, nv;
  int P2;          // N2 = P2 * Q2, the factor pair closest to sqrt(N2) (1 for a prime N2)
  const cplx* tw;  // Stockham tables of the N1-point transform
  const cplx* wN;  // [N1 * N2]  exp(-2 pi i j / N)
  const cplx* w2;  // [N2]       exp(-2 pi i j / N2)
  // multiplier: advection phase exp(-i k alpha_col), alpha = k1 (v_col dt), or a table mtab[k], k = 0 .. N/2
  const double* v;
  double dt, k1;
  const double* k1_batch;
  const cplx* mtab;
  long long mtab_stride;
};

// ---- K1: forward N1-point transforms + twiddle -------------------------------------------------------------------
template <int LOGN1>
__global__ void __launch_bounds__(BX_COLS * FftCfg<LOGN1>::T) bigx_k1_kernel(BigXArgs p) {
  using C = FftCfg<LOGN1>;
  constexpr int E = C::E, T = C::T, N1 = C::N;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* buf = reinterpret_cast<cplx*>(smem_raw);
  const int g = threadIdx.x & (BX_COLS - 1), t = threadIdx.x / BX_COLS;
  const int n2 = blockIdx.y, N2 = p.N2;
  const long long col = 2 * ((long long)blockIdx.x * BX_COLS + g);
  const long long base = (long long)blockIdx.z * N1 * N2 * p.nv + col;
  cplx x[E];
#pragma unroll
  for (int m = 0; m < E; m++)
    x[m] = *reinterpret_cast<const double2*>(p.in + base + (long long)(n2 + N2 * (t + T * m)) * p.nv);
  fft_forward<LOGN1, BX_COLS>(x, buf + g, p.tw, t, 0);
#pragma unroll
  for (int m = 0; m < E; m++) {
    const int kk = t + T * m;
    const cplx w = __ldg(p.wN + n2 * kk);  // n2 k1 < N
    *reinterpret_cast<double2*>(p.out + base + (long long)(n2 * N1 + kk) * p.nv) = cmul(x[m], w);
  }
}

// ---- K3: inverse N1-point transforms ---------------------------------------------------------------------------------
template <int LOGN1>
__global__ void __launch_bounds__(BX_COLS * FftCfg<LOGN1>::T) bigx_k3_kernel(BigXArgs p) {
  using C = FftCfg<LOGN1>;
  constexpr int E = C::E, T = C::T, N1 = C::N;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* buf = reinterpret_cast<cplx*>(smem_raw);
  const int g = threadIdx.x & (BX_COLS - 1), t = threadIdx.x / BX_COLS;
  const int n2 = blockIdx.y, N2 = p.N2;
  const long long col = 2 * ((long long)blockIdx.x * BX_COLS + g);
  const long long base = (long long)blockIdx.z * N1 * N2 * p.nv + col;
  cplx x[E];
#pragma unroll
  for (int m = 0; m < E; m++) {
    const double2 y = *reinterpret_cast<const double2*>(p.in + base + (long long)(n2 * N1 + t + T * m) * p.nv);
    x[m] = cmake(y.y, y.x);  // inverse = swap . forward . swap
  }
  fft_forward<LOGN1, BX_COLS>(x, buf + g, p.tw, t, 0);
#pragma unroll
  for (int m = 0; m < E; m++)
    *reinterpret_cast<double2*>(p.out + base + (long long)(n2 + N2 * (t + T * m)) * p.nv) = make_double2(x[m].y, x[m].x);
}

// ---- K2: N2-point DFTs, spectrum update, inverse N2-point DFTs ---------------------------------------------------------
// acc[j] = sum_n S[n][lane] W^(sign n k), k = s + 16 j, by one Cooley-Tukey level N2 = P Q (n = P a + b, k = c + Q d):
//   T[b Q + c] = sum_a S[P a + b] W_Q^(a c),   X[k] = sum_b T[b Q + (k mod Q)] W_N2^(b k)
// (P + Q multiply-adds per output instead of N2; P = 1 for a prime N2 is the plain sum).  T: a third [N2][32] array.
// Every thread of the CTA must call it (two barriers).
template <bool INVERSE>
__device__ __forceinline__ void dft_n2(const cplx* __restrict__ S, cplx* __restrict__ Tm, const cplx* __restrict__ W,
                                       int N2, int P, int Q, int lane, int s, cplx (&acc)[BX_MAXK]) {
  __syncthreads();  // T may still be read by the previous call
  int oidx[BX_MAXK], ostep[BX_MAXK], obase[BX_MAXK];
#pragma unroll
  for (int j = 0; j < BX_MAXK; j++) {
    const int o = s + 16 * j;          // stage-1 output (b, c) = (o / Q, o % Q)
    const int b = o / Q, c = o - b * Q;
    acc[j] = cmake(0.0, 0.0);
    oidx[j] = 0;
    ostep[j] = (P * c) % N2;           // W_Q^(a c) = W_N2^(P a c)
    obase[j] = b;
  }
  for (int a = 0; a < Q; a++) {
#pragma unroll
    for (int j = 0; j < BX_MAXK; j++) {
      if (s + 16 * j < N2) {
        const cplx x = S[(P * a + obase[j]) * BX_COLS + lane];
        cplx w = W[oidx[j]];
        if (INVERSE) w.y = -w.y;
        acc[j].x = fma(x.x, w.x, fma(-x.y, w.y, acc[j].x));
        acc[j].y = fma(x.x, w.y, fma(x.y, w.x, acc[j].y));
        oidx[j] += ostep[j];
        if (oidx[j] >= N2) oidx[j] -= N2;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < BX_MAXK; j++)
    if (s + 16 * j < N2) Tm[(s + 16 * j) * BX_COLS + lane] = acc[j];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < BX_MAXK; j++) {
    const int k = s + 16 * j;
    acc[j] = cmake(0.0, 0.0);
    if (k < N2) {
      const int c = k % Q;
      int idx = 0;
      for (int b = 0; b < P; b++) {
        const cplx x = Tm[(b * Q + c) * BX_COLS + lane];
        cplx w = W[idx];
        if (INVERSE) w.y = -w.y;
        acc[j].x = fma(x.x, w.x, fma(-x.y, w.y, acc[j].x));
        acc[j].y = fma(x.x, w.y, fma(x.y, w.x, acc[j].y));
        idx += k;
        if (idx >= N2) idx -= N2;
      }
    }
  }
}

__global__ void __launch_bounds__(512, 1) bigx_k2_kernel(BigXArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N1 = p.N1, N2 = p.N2, N = N1 * N2;
  cplx* SA = reinterpret_cast<cplx*>(smem_raw);
  cplx* SB = SA + (size_t)N2 * BX_COLS;
  cplx* ST = SB + (size_t)N2 * BX_COLS;
  cplx* W = ST + (size_t)N2 * BX_COLS;
  const int P = p.P2, Q = N2 / P;
  const int lane = threadIdx.x & 31, s = threadIdx.x >> 5;
  const int ka = blockIdx.y;              // group A: k = ka + N1 k2
  const int kb = (N1 - ka) % N1;          // group B holds the partners N - k
  const bool self = (kb == ka);           // ka = 0 or N1 / 2
  const long long col = 2 * ((long long)blockIdx.x * BX_COLS + lane);
  double* Y = p.out + (long long)blockIdx.z * N * p.nv + col;  // in place on the scratch array

  for (int i = threadIdx.x; i < N2; i += blockDim.x) W[i] = __ldg(p.w2 + i);
  for (int n = s; n < N2; n += 16) {
    SA[n * BX_COLS + lane] = *reinterpret_cast<const double2*>(Y + (long long)(n * N1 + ka) * p.nv);
    if (!self) SB[n * BX_COLS + lane] = *reinterpret_cast<const double2*>(Y + (long long)(n * N1 + kb) * p.nv);
  }
  __syncthreads();

  cplx acc[BX_MAXK];
  dft_n2<false>(SA, ST, W, N2, P, Q, lane, s, acc);
  __syncthreads();
#pragma unroll
  for (int j = 0; j < BX_MAXK; j++)
    if (s + 16 * j < N2) SA[(s + 16 * j) * BX_COLS + lane] = acc[j];
  if (!self) {
    dft_n2<false>(SB, ST, W, N2, P, Q, lane, s, acc);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < BX_MAXK; j++)
      if (s + 16 * j < N2) SB[(s + 16 * j) * BX_COLS + lane] = acc[j];
  }
  __syncthreads();

  // ---- spectrum update on the pairs (k, N - k); lo = min(k, N - k) is the non-negative frequency index -------------
  {
    const double k1x = p.k1_batch ? p.k1_batch[blockIdx.z] : p.k1;
    double alpha_a = 0.0, alpha_b = 0.0;
    if (!p.mtab) {
      alpha_a = k1x * (p.v[col] * p.dt);
      alpha_b = k1x * (p.v[col + 1] * p.dt);
    }
    const cplx* mt = p.mtab ? p.mtab + (long long)blockIdx.z * p.mtab_stride : nullptr;
    const double inv_n = 1.0 / (double)N;
    cplx* SQ = self ? SA : SB;
    for (int k2 = s; k2 < N2; k2 += 16) {
      const int k = ka + N1 * k2;
      const int q = (N - k) % N;
      const int q2 = q / N1;  // q % N1 == kb
      if (self && k > q) continue;  // each pair once
      const bool swap = k > q;
      const int lo = swap ? q : k;
      const cplx zk = SA[k2 * BX_COLS + lane];
      cplx ma, mb;
      if (mt) {
        ma = mb = mt[lo];
      } else {
        double sn, cs;
        sincos((double)lo * alpha_a, &sn, &cs);
        ma = cmake(cs * inv_n, -sn * inv_n);
        sincos((double)lo * alpha_b, &sn, &cs);
        mb = cmake(cs * inv_n, -sn * inv_n);
      }
      if (k == q) {  // DC or Nyquist: real coefficients, irfft keeps the real part of the product
        SA[k2 * BX_COLS + lane] = cmake(zk.x * ma.x, zk.y * mb.x);
        continue;
      }
      const cplx zq = SQ[q2 * BX_COLS + lane];
      const cplx zl = swap ? zq : zk, zh = swap ? zk : zq;
      // A = (zl + conj zh) / 2, B = (zl - conj zh) / (2 i): spectra of the two real columns at frequency lo
      const cplx A = cmake(0.5 * (zl.x + zh.x), 0.5 * (zl.y - zh.y));
      const cplx B = cmake(0.5 * (zl.y + zh.y), 0.5 * (zh.x - zl.x));
      const cplx Ap = cmul(A, ma), Bp = cmul(B, mb);
      const cplx nl = cmake(Ap.x - Bp.y, Ap.y + Bp.x);   // A' + i B'
      const cplx nh = cmake(Ap.x + Bp.y, Bp.x - Ap.y);   // conj(A') + i conj(B')
      SA[k2 * BX_COLS + lane] = swap ? nh : nl;
      SQ[q2 * BX_COLS + lane] = swap ? nl : nh;
    }
  }
  __syncthreads();

  // ---- inverse N2-point DFTs, conjugate twiddle, write back -----------------------------------------------------------
  dft_n2<true>(SA, ST, W, N2, P, Q, lane, s, acc);
#pragma unroll
  for (int j = 0; j < BX_MAXK; j++) {
    const int n = s + 16 * j;
    if (n < N2) {
      cplx w = __ldg(p.wN + n * ka);
      w.y = -w.y;
      *reinterpret_cast<double2*>(Y + (long long)(n * N1 + ka) * p.nv) = cmul(acc[j], w);
    }
  }
  if (!self) {
    dft_n2<true>(SB, ST, W, N2, P, Q, lane, s, acc);
#pragma unroll
    for (int j = 0; j < BX_MAXK; j++) {
      const int n = s + 16 * j;
      if (n < N2) {
        cplx w = __ldg(p.wN + n * kb);
        w.y = -w.y;
        *reinterpret_cast<double2*>(Y + (long long)(n * N1 + kb) * p.nv) = cmul(acc[j], w);
      }
    }
  }
}

struct BigXTables {
  int n1, n2;
  cplx* wN;
  cplx* w2;
};
std::mutex g_bx_mutex;
std::vector<BigXTables> g_bx[64];

int get_tables(int n1, int n2, const cplx** wN, const cplx** w2) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    set_last_error("bigx: no CUDA device");
    (void)cudaGetLastError();
    return ADEPT_ERR_CUDA;
  }
  std::lock_guard<std::mutex> lock(g_bx_mutex);
  for (const BigXTables& t : g_bx[dev])
    if (t.n1 == n1 && t.n2 == n2) {
      *wN = t.wN, *w2 = t.w2;
      return ADEPT_OK;
    }
  const long long N = (long long)n1 * n2;
  const long double two_pi = 6.283185307179586476925286766559005768L;
  std::vector<cplx> hN((size_t)N), h2((size_t)n2);
  for (long long j = 0; j < N; j++) {
    const long double a = two_pi * (long double)j / (long double)N;
    hN[(size_t)j] = make_double2((double)cosl(a), (double)(-sinl(a)));
  }
  for (int j = 0; j < n2; j++) {
    const long double a = two_pi * (long double)j / (long double)n2;
    h2[(size_t)j] = make_double2((double)cosl(a), (double)(-sinl(a)));
  }
  BigXTables t = {n1, n2, nullptr, nullptr};
  cudaError_t err = cudaMalloc(&t.wN, hN.size() * sizeof(cplx));
  if (err == cudaSuccess) err = cudaMalloc(&t.w2, h2.size() * sizeof(cplx));
  if (err == cudaSuccess) err = cudaMemcpy(t.wN, hN.data(), hN.size() * sizeof(cplx), cudaMemcpyHostToDevice);
  if (err == cudaSuccess) err = cudaMemcpy(t.w2, h2.data(), h2.size() * sizeof(cplx), cudaMemcpyHostToDevice);
  if (err != cudaSuccess) {
    set_last_error("bigx tables (%d x %d): %s", n1, n2, cudaGetErrorString(err));
    (void)cudaGetLastError();
    return ADEPT_ERR_CUDA;
  }
  g_bx[dev].push_back(t);
  *wN = t.wN, *w2 = t.w2;
  return ADEPT_OK;
}

bool factor(int nx, int* logn1, int* n2) {
  if (nx < 2) return false;
  int a = 0, m = nx;
  while ((m & 1) == 0) m >>= 1, a++;
  if (m == 1 || a < 6) return false;  // powers of two and short pencils have their own kernels
  while (a > 8) m <<= 1, a--;         // fold surplus factors of two into N2
  if (m > BX_MAXN2) return false;
  *logn1 = a, *n2 = m;
  return true;
}

template <int LOGN1, bool K3>
int launch_k13(const BigXArgs& p, int batch, cudaStream_t stream) {
  using C = FftCfg<LOGN1>;
  const size_t smem = (size_t)C::BUF * BX_COLS * sizeof(cplx);
  auto kern = K3 ? bigx_k3_kernel<LOGN1> : bigx_k1_kernel<LOGN1>;
  static bool configured[64][2] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !configured[dev][K3]) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(bigx k1/k3, smem=%zu): %s", smem, cudaGetErrorString(err));
      return ADEPT_ERR_CUDA;
    }
    configured[dev][K3] = true;
  }
  ProfileScope prof(K3 ? "bigx_k3" : "bigx_k1", stream);
  kern<<<dim3(p.nv / (2 * BX_COLS), p.N2, batch), BX_COLS * C::T, smem, stream>>>(p);
  return check_launch(K3 ? "bigx_k3_kernel" : "bigx_k1_kernel");
}

}  // namespace

bool bigx_supported(int nx, int nv) {
  int l, m;
  return factor(nx, &l, &m) && nv > 0 && nv % (2 * BX_COLS) == 0;
}

// out = irfft(M rfft(in, axis = x), axis = x) for in / out [batch, nx, nv]; scratch: one more array of that size (must not
// alias in or out; in == out is allowed).  mtab == nullptr: advection phase from (v, dt, k1 / k1_batch).
int bigx_apply_f64(const double* in, double* out, double* scratch, int batch, int nx, int nv, const double* v, double dt,
                   const double* k1_batch, double k1, const cplx* mtab, long long mtab_stride, cudaStream_t stream) {
  int logn1 = 0, n2 = 0;
  if (batch < 1 || !factor(nx, &logn1, &n2) || nv % (2 * BX_COLS)) {
    set_last_error("bigx: unsupported batch=%d nx=%d nv=%d (nx = 2^a m, 6 <= a, m odd, m 2^max(a-8,0) <= %d; nv %% %d == 0)",
                   batch, nx, nv, BX_MAXN2, 2 * BX_COLS);
    return ADEPT_ERR_UNSUPPORTED;
  }
  if (!scratch || scratch == in || scratch == out || (!mtab && !v)) {
    set_last_error("bigx: needs a scratch array distinct from the input and output (and v for the advection phase)");
    return ADEPT_ERR_BAD_ARG;
  }
  if (batch > 65535) {
    set_last_error("bigx: batch=%d exceeds the grid limit", batch);
    return ADEPT_ERR_UNSUPPORTED;
  }
  BigXArgs p = {};
  p.N1 = 1 << logn1, p.N2 = n2, p.nv = nv;
  p.P2 = 1;
  for (int d = 2; d * d <= n2; d++)
    if (n2 % d == 0) p.P2 = d;
  p.tw = get_twiddles(logn1);
  if (!p.tw) return ADEPT_ERR_CUDA;
  const int rc = get_tables(p.N1, p.N2, &p.wN, &p.w2);
  if (rc != ADEPT_OK) return rc;
  p.v = v, p.dt = dt, p.k1 = k1, p.k1_batch = k1_batch, p.mtab = mtab, p.mtab_stride = mtab_stride;
  // K1: in -> scratch
  p.in = in, p.out = scratch;
  int r = logn1 == 6 ? launch_k13<6, false>(p, batch, stream)
                     : (logn1 == 7 ? launch_k13<7, false>(p, batch, stream) : launch_k13<8, false>(p, batch, stream));
  if (r != ADEPT_OK) return r;
  // K2: scratch in place
  {
    const size_t smem = ((size_t)3 * n2 * BX_COLS + n2) * sizeof(cplx);
    if (smem > 227 * 1024) {
      set_last_error("bigx: N2=%d does not fit shared memory", n2);
      return ADEPT_ERR_UNSUPPORTED;
    }
    static size_t configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && configured[dev] < smem) {
      cudaError_t err = cudaFuncSetAttribute(bigx_k2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (err != cudaSuccess) {
        set_last_error("cudaFuncSetAttribute(bigx k2, smem=%zu): %s", smem, cudaGetErrorString(err));
        return ADEPT_ERR_CUDA;
      }
      configured[dev] = smem;
    }
    p.in = scratch, p.out = scratch;
    ProfileScope prof("bigx_k2", stream);
    bigx_k2_kernel<<<dim3(nv / (2 * BX_COLS), p.N1 / 2 + 1, batch), 512, smem, stream>>>(p);
    r = check_launch("bigx_k2_kernel");
    if (r != ADEPT_OK) return r;
  }
  // K3: scratch -> out
  p.in = scratch, p.out = out;
  return logn1 == 6 ? launch_k13<6, true>(p, batch, stream)
                    : (logn1 == 7 ? launch_k13<7, true>(p, batch, stream) : launch_k13<8, true>(p, batch, stream));
}

// ---- field solve on a long mixed-length grid: the same three launches on a [nx, 64] array whose first column is rho ---
namespace {
constexpr int BF_COLS = 2 * BX_COLS;

__global__ void __launch_bounds__(1024) bigx_mean_kernel(const double* __restrict__ rho, int n, double* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += rho[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = warp_sum(red[threadIdx.x]);
    if (threadIdx.x == 0) *out = s / (double)n;  // np.mean
  }
}

// in2d[i, 0] = rho[i], other columns 0; mtab[k] = -i K(k) / n, K = 1/kx (mode 0) or kx (Te/rho0) / (1 + lambda^2 kx^2)
__global__ void __launch_bounds__(256) bigx_field_prep_kernel(const double* __restrict__ rho,
                                                              const double* __restrict__ kmul, int n, int mode, double Te,
                                                              double lambda_De, const double* __restrict__ mean,
                                                              double* __restrict__ in2d, cplx* __restrict__ mtab) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (long long)n * BF_COLS) in2d[i] = (i % BF_COLS == 0) ? rho[i / BF_COLS] : 0.0;
  if (i <= n / 2) {
    double K = kmul[i];
    if (mode == 1) {
      const double rho0 = *mean;
      const double lam2 = lambda_De < 0.0 ? Te / rho0 : lambda_De * lambda_De;
      K = K * (Te / rho0) / (1.0 + lam2 * K * K);
    }
    mtab[i] = cmake(0.0, -K / (double)n);
  }
}

__global__ void __launch_bounds__(256) bigx_field_take_kernel(const double* __restrict__ out2d, int n, double* __restrict__ e) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) e[i] = out2d[(long long)i * BF_COLS];
}

struct FieldScratch {
  int n;
  double* buf;  // in2d | out2d | scratch (3 n BF_COLS doubles) | mean (2 doubles) | mtab (n / 2 + 1 cplx)
};
std::mutex g_bf_mutex;
std::vector<FieldScratch> g_bf[64];

double* field_scratch(int n) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(g_bf_mutex);
  for (const FieldScratch& f : g_bf[dev])
    if (f.n == n) return f.buf;
  FieldScratch f = {n, nullptr};
  const size_t bytes = ((size_t)3 * n * BF_COLS + 2) * sizeof(double) + ((size_t)n / 2 + 1) * sizeof(cplx);
  if (cudaMalloc(&f.buf, bytes) != cudaSuccess) {
    (void)cudaGetLastError();
    return nullptr;
  }
  g_bf[dev].push_back(f);
  return f.buf;
}
}  // namespace

int bigx_poisson_f64(const double* rho, const double* kmul, long long kmul_stride, double* e, int batch, int nx,
                     int mode, double Te, double lambda_De, cudaStream_t stream) {
  if (!bigx_supported(nx, BF_COLS) || batch < 1) {
    set_last_error("poisson(bigx): unsupported nx=%d", nx);
    return ADEPT_ERR_UNSUPPORTED;
  }
  double* buf = field_scratch(nx);  // per-(device, nx) work arrays, allocated on first use like the twiddle tables
  if (!buf) {
    set_last_error("poisson(bigx): cannot allocate the work arrays for nx=%d", nx);
    return ADEPT_ERR_CUDA;
  }
  double* in2d = buf;
  double* out2d = buf + (size_t)nx * BF_COLS;
  double* scr = out2d + (size_t)nx * BF_COLS;
  double* mean = scr + (size_t)nx * BF_COLS;
  cplx* mtab = reinterpret_cast<cplx*>(mean + 2);
  for (int b = 0; b < batch; b++) {
    const double* rb = rho + (size_t)b * nx;
    const double* kb = kmul + (size_t)b * kmul_stride;
    {
      ProfileScope prof("bigx_field_prep", stream);
      if (mode == 1) bigx_mean_kernel<<<1, 1024, 0, stream>>>(rb, nx, mean);
      const long long total = (long long)nx * BF_COLS;
      bigx_field_prep_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(rb, kb, nx, mode, Te, lambda_De, mean,
                                                                               in2d, mtab);
      const int rc = check_launch("bigx_field_prep_kernel");
      if (rc != ADEPT_OK) return rc;
    }
    const int rc = bigx_apply_f64(in2d, out2d, scr, 1, nx, BF_COLS, nullptr, 0.0, nullptr, 0.0, mtab, 0, stream);
    if (rc != ADEPT_OK) return rc;
    ProfileScope prof("bigx_field_take", stream);
    bigx_field_take_kernel<<<(nx + 255) / 256, 256, 0, stream>>>(out2d, nx, e + (size_t)b * nx);
    const int rc2 = check_launch("bigx_field_take_kernel");
    if (rc2 != ADEPT_OK) return rc2;
  }
  return ADEPT_OK;
}

}  // namespace adept
