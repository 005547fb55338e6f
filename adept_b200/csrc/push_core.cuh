// Shared pieces of the spectral advection kernels (push.cu, vdfdx_tma.cu): phase tables and the half-spectrum update
// that sits between the forward and the inverse transform.
#pragma once
#include "fft_core.cuh"

namespace adept {

// Per-sequence phase table (shared memory): the thread's base phase exp(-i t alpha) / (2N) is the product of a "lo"
// and a "hi" entry (t = (hi << LOBT) + lo); the phases of its other modes t + T m follow by repeated multiplication
// with step = exp(-i T alpha).  nyq = cos(alpha N/2) / (2N) (irfft drops the imaginary part of the Nyquist mode).
template <int LOGN>
struct PhaseCfg {
  static constexpr int N = 1 << LOGN;
  static constexpr int T = FftCfg<LOGN>::T;
  static constexpr int LOGT = LOGN - (LOGN < 4 ? LOGN : 4);
  static constexpr int LOBT = LOGT > 4 ? 4 : LOGT;
  static constexpr int NLO = 1 << LOBT;
  static constexpr int NHI = T >> LOBT;
  static constexpr int STEP = NLO + NHI;  // index of exp(-i T alpha)
  static constexpr int NYQ = NLO + NHI + 1;
  static constexpr int PER_SEQ = NLO + NHI + 2;
};

// Fill the two phase tables of one sequence pair; called by the `stride` threads t = 0..stride-1 of the FFT group.
// F64-BEGIN (the generated fp32 build keeps the phase arithmetic in fp64 and rounds the table entries once)
template <int LOGN>
__device__ __forceinline__ void phase_table_fill(cplx* ph, double alpha_a, double alpha_b, int t, int stride) {
  using PC = PhaseCfg<LOGN>;
  constexpr int N = PC::N, T = PC::T;
  for (int i = t; i < 2 * PC::PER_SEQ; i += stride) {
    const int s = i / PC::PER_SEQ, j = i % PC::PER_SEQ;
    const double al = s ? alpha_b : alpha_a;
    const double sc = 0.5 / (double)N;
    // one sincos per entry on a common path (divergent per-kind branches cost the filling warps ~3000 cycles)
    const bool lo = j < PC::NLO, nyq = j == PC::NYQ;
    const int mult = lo ? j : (j < PC::STEP ? ((j - PC::NLO) << PC::LOBT) : (j == PC::STEP ? T : N / 2));
    double sn, cs;
    sincos((double)mult * al, &sn, &cs);
    const double amp = (lo || nyq) ? sc : 1.0;
    ph[i] = cmake(cs * amp, nyq ? 0.0 : -sn * amp);
  }
}
// F64-END

// Half-spectrum update.  On entry thread t holds Z[t + T m] in x[m], Z = FFT(a + i b).  Its lower register half
// (m < E/2) are the modes k = t + T m < N/2; the partner N - k of each lives in the upper register half of thread
// (T - t) mod T.  Upper halves are published in natural order, every thread updates its E/2 (k, N-k) pairs (one phase
// evaluation and one spectrum separation per pair: A' = pa A, B' = pb B, Z' = A' + i B'), writes the partner value
// back, and upper halves are read back.  Results are left swapped (im, re): the inverse transform is
// swap . forward FFT . swap.  `filt` (nullable, [N/2+1]) is an extra real multiplier per mode.  All threads of the CTA
// must call it (it uses __syncthreads()).
template <int LOGN, int BS>
__device__ __forceinline__ void half_spectrum_update(cplx (&x)[FftCfg<LOGN>::E], cplx* buf, const cplx* ph, int t,
                                                     const double* __restrict__ filt = nullptr) {
  using C = FftCfg<LOGN>;
  using PC = PhaseCfg<LOGN>;
  constexpr int N = C::N, E = C::E, T = C::T, H = E / 2;
  __syncthreads();  // last forward pass has finished reading buf; phase tables are complete
#pragma unroll
  for (int m = H; m < E; m++) buf[fft_pad(t + T * m) * BS] = x[m];
  __syncthreads();
  {
    const cplx* pha = ph;
    const cplx* phb = ph + PC::PER_SEQ;
    cplx pa = cmul(pha[t & (PC::NLO - 1)], pha[PC::NLO + (t >> PC::LOBT)]);
    cplx pb = cmul(phb[t & (PC::NLO - 1)], phb[PC::NLO + (t >> PC::LOBT)]);
    const cplx sa = pha[PC::STEP], sb = phb[PC::STEP];
#pragma unroll
    for (int m = 0; m < H; m++) {
      if (m > 0) {
        pa = cmul(pa, sa);
        pb = cmul(pb, sb);
      }
      const int k = t + T * m;
      const bool self = (k == 0);  // DC pairs with itself (its phase is real: alpha * 0)
      const int q = fft_pad((N - k) & (N - 1)) * BS;
      const cplx zk = x[m];
      const cplx zq = self ? zk : buf[q];
      const cplx A = cmake(zk.x + zq.x, zk.y - zq.y);  // 2 * spectrum of a at k
      const cplx B = cmake(zk.y + zq.y, zq.x - zk.x);  // 2 * spectrum of b at k
      cplx Ap = cmul(A, pa), Bp = cmul(B, pb);
      if (filt) {  // real per-mode multiplier (Hou-Li filter, vlasov.py:211-219)
        const double s = __ldg(filt + k);
        Ap.x *= s, Ap.y *= s, Bp.x *= s, Bp.y *= s;
      }
      // Z'[k] = A' + i B' ;  Z'[N-k] = conj(A') + i conj(B')
      x[m] = cmake(Ap.y + Bp.x, Ap.x - Bp.y);
      if (!self) buf[q] = cmake(Bp.x - Ap.y, Ap.x + Bp.y);
    }
    if (t == 0) {  // Nyquist mode: real phase cos(alpha N/2), pairs with itself
      const int q = fft_pad(N / 2) * BS;
      const cplx z = buf[q];
      const double s = filt ? __ldg(filt + N / 2) : 1.0;
      buf[q] = cmake(2.0 * z.y * phb[PC::NYQ].x * s, 2.0 * z.x * pha[PC::NYQ].x * s);
    }
  }
  __syncthreads();
#pragma unroll
  for (int m = H; m < E; m++) x[m] = buf[fft_pad(t + T * m) * BS];
}

}  // namespace adept
