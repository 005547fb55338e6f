// In-shared-memory Stockham autosort complex FFT, length N = 2^LOGN (2 <= N <= 8192), fp64.
//
// One FFT is computed by T = N/E threads (E = min(16, N) points per thread per pass).  Passes are
// radix-16 (as many as fit) followed by one radix-2/4/8 pass.  In every pass thread t holds the
// points e = t + T*m (m = 0..E-1) in registers x[m]; a radix-R butterfly q (q = 0..E/R-1) combines
// x[q + r*E/R].  Between passes points are exchanged through a padded shared-memory buffer
// (pad(e) = e + e/16 keeps 16-byte accesses conflict-free for the stride-16 scatter).  The first
// pass takes its inputs from registers (the caller loads them from global memory), the last pass
// leaves its outputs in registers in the same e = t + T*m order, so callers read and write global
// memory coalesced with no extra staging pass.
//
// Twiddles: per-pass tables tw[off_p + (r-1)*Ns + k] = exp(-2*pi*i*k*r/(Ns*R)), k < Ns, so that the
// T threads of a pass read consecutive entries (built on the host in long double; api.cu).
#pragma once
#include "common.cuh"

// Phase tracing hook (tools/micro/vpush_trace.cu defines it to record clock64() per warp); a no-op in the library.
#ifndef ADEPT_TRACE
#define ADEPT_TRACE(id)
#endif

namespace adept {

template <int LOGN>
struct FftCfg {
  static constexpr int N = 1 << LOGN;
  static constexpr int E = N < 16 ? N : 16;
  static constexpr int T = N / E;
  static constexpr int NP16 = N < 16 ? 0 : LOGN / 4;
  static constexpr int REM = N < 16 ? LOGN : LOGN % 4;
  static constexpr int NPASS = NP16 + (REM ? 1 : 0);
  static constexpr int BUF = N + N / 16;  // padded buffer length in cplx
  __host__ __device__ static constexpr int radix(int p) { return p < NP16 ? 16 : (1 << REM); }
  __host__ __device__ static constexpr int ns(int p) {  // product of radices before pass p
    int v = 1;
    for (int i = 0; i < p; i++) v *= radix(i);
    return v;
  }
  __host__ __device__ static constexpr int tw_off(int p) {  // table offset of pass p
    int v = 0;
    for (int i = 1; i < p; i++) v += (radix(i) - 1) * ns(i);
    return v;
  }
  static constexpr int TW_TOTAL = tw_off(NPASS);
};

__device__ __forceinline__ int fft_pad(int e) { return e + (e >> 4); }

// ---- small DFTs, forward sign (W = exp(-2 pi i / R)), natural-order in and out -----------------

__device__ __forceinline__ void dft2(cplx& a, cplx& b) {
  cplx t = a;
  a = cadd(t, b);
  b = csub(t, b);
}

__device__ __forceinline__ void dft4(cplx& x0, cplx& x1, cplx& x2, cplx& x3) {
  cplx t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = cmul_mi(csub(x1, x3));
  x0 = cadd(t0, t2);
  x1 = cadd(t1, t3);
  x2 = csub(t0, t2);
  x3 = csub(t1, t3);
}

#define ADEPT_SQRT1_2 0.70710678118654752440
#define ADEPT_COS_PI_8 0.92387953251128675613
#define ADEPT_SIN_PI_8 0.38268343236508977173

__device__ __forceinline__ void dft8(cplx* v) {  // v[0..7]
  // even / odd DIT
  cplx e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
  cplx o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
  dft4(e0, e1, e2, e3);
  dft4(o0, o1, o2, o3);
  // W8^1 = (1 - i)/sqrt2, W8^2 = -i, W8^3 = (-1 - i)/sqrt2
  o1 = cmake((o1.x + o1.y) * ADEPT_SQRT1_2, (o1.y - o1.x) * ADEPT_SQRT1_2);
  o2 = cmul_mi(o2);
  o3 = cmake((o3.y - o3.x) * ADEPT_SQRT1_2, -(o3.x + o3.y) * ADEPT_SQRT1_2);
  v[0] = cadd(e0, o0);
  v[4] = csub(e0, o0);
  v[1] = cadd(e1, o1);
  v[5] = csub(e1, o1);
  v[2] = cadd(e2, o2);
  v[6] = csub(e2, o2);
  v[3] = cadd(e3, o3);
  v[7] = csub(e3, o3);
}

__device__ __forceinline__ void dft16(cplx* v) {  // v[0..15]
  // n = c + 4d, k = k1 + 4 k2:  y[k1+4k2] = sum_c W4^{c k2} W16^{c k1} sum_d x[c+4d] W4^{d k1}
#pragma unroll
  for (int c = 0; c < 4; c++) dft4(v[c], v[c + 4], v[c + 8], v[c + 12]);  // -> v[c + 4 k1]
  const cplx w1 = cmake(ADEPT_COS_PI_8, -ADEPT_SIN_PI_8);
  const cplx w3 = cmake(ADEPT_SIN_PI_8, -ADEPT_COS_PI_8);
  // k1 = 1: c = 1,2,3 -> W16^1, W16^2, W16^3
  v[5] = cmul(v[5], w1);
  v[6] = cmake((v[6].x + v[6].y) * ADEPT_SQRT1_2, (v[6].y - v[6].x) * ADEPT_SQRT1_2);
  v[7] = cmul(v[7], w3);
  // k1 = 2: W16^2, W16^4, W16^6
  v[9] = cmake((v[9].x + v[9].y) * ADEPT_SQRT1_2, (v[9].y - v[9].x) * ADEPT_SQRT1_2);
  v[10] = cmul_mi(v[10]);
  v[11] = cmake((v[11].y - v[11].x) * ADEPT_SQRT1_2, -(v[11].x + v[11].y) * ADEPT_SQRT1_2);
  // k1 = 3: W16^3, W16^6, W16^9 = -W16^1
  v[13] = cmul(v[13], w3);
  v[14] = cmake((v[14].y - v[14].x) * ADEPT_SQRT1_2, -(v[14].x + v[14].y) * ADEPT_SQRT1_2);
  v[15] = cmul(v[15], cmake(-ADEPT_COS_PI_8, ADEPT_SIN_PI_8));
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);  // -> v[4k1 + k2]
  // transpose register names so that v[k] = y[k]
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++)
#pragma unroll
    for (int k2 = k1 + 1; k2 < 4; k2++) {
      cplx t = v[4 * k1 + k2];
      v[4 * k1 + k2] = v[4 * k2 + k1];
      v[4 * k2 + k1] = t;
    }
}

template <int R>
__device__ __forceinline__ void dft_r(cplx* v) {
  if constexpr (R == 2) dft2(v[0], v[1]);
  if constexpr (R == 4) dft4(v[0], v[1], v[2], v[3]);
  if constexpr (R == 8) dft8(v);
  if constexpr (R == 16) dft16(v);
}

// ---- Stockham passes -------------------------------------------------------------------------

// compiler-level scheduling fence: keeps the loads of one butterfly column from being hoisted above the previous
// column's arithmetic (a radix-16 pass would otherwise hold 16 points + 15 twiddles = 124 registers in flight)
__device__ __forceinline__ void sched_fence() { asm volatile("" ::: "memory"); }

// second half of the radix-16 butterfly: inter-stage twiddles W16^{c k1}, row DFT4s, register transposition.
// On entry v[c + 4 k1] holds the column DFT4 outputs; on exit v[k] = y[k].
__device__ __forceinline__ void dft16_finish(cplx* v) {
  const cplx w1 = cmake(ADEPT_COS_PI_8, -ADEPT_SIN_PI_8);
  const cplx w3 = cmake(ADEPT_SIN_PI_8, -ADEPT_COS_PI_8);
  v[5] = cmul(v[5], w1);
  v[6] = cmake((v[6].x + v[6].y) * ADEPT_SQRT1_2, (v[6].y - v[6].x) * ADEPT_SQRT1_2);
  v[7] = cmul(v[7], w3);
  v[9] = cmake((v[9].x + v[9].y) * ADEPT_SQRT1_2, (v[9].y - v[9].x) * ADEPT_SQRT1_2);
  v[10] = cmul_mi(v[10]);
  v[11] = cmake((v[11].y - v[11].x) * ADEPT_SQRT1_2, -(v[11].x + v[11].y) * ADEPT_SQRT1_2);
  v[13] = cmul(v[13], w3);
  v[14] = cmake((v[14].y - v[14].x) * ADEPT_SQRT1_2, -(v[14].x + v[14].y) * ADEPT_SQRT1_2);
  v[15] = cmul(v[15], cmake(-ADEPT_COS_PI_8, ADEPT_SIN_PI_8));
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++)
#pragma unroll
    for (int k2 = k1 + 1; k2 < 4; k2++) {
      cplx t = v[4 * k1 + k2];
      v[4 * k1 + k2] = v[4 * k2 + k1];
      v[4 * k2 + k1] = t;
    }
}

// Hook called by the last pass as soon as this thread has finished reading the exchange buffer (before the second half
// of its butterfly): callers that hand the buffer to an asynchronous producer (the next tile's TMA load) start it there.
struct NoHook {
  __device__ __forceinline__ void operator()() const {}
};
// A hook type that declares `static constexpr bool HAS_AT = true` is also called as hook.at<P, W>() at the start of
// the two arithmetic stretches of every radix-16 pass P: W = 0 after the pass's exchange loads have been issued (before
// the column butterflies), W = 1 after its mid-pass barrier (before the second half).  Callers queue asynchronous work
// for the load/store pipe there (deferred global stores), which then drains under the fp64 work that follows.
template <class H, class = void>
struct hook_has_at {
  static constexpr bool value = false;
};
template <class H>
struct hook_has_at<H, decltype((void)H::HAS_AT)> {
  static constexpr bool value = true;
};
template <int P, int W, class H>
__device__ __forceinline__ void hook_at(const H& h) {
  if constexpr (hook_has_at<H>::value) h.template at<P, W>();
}

__device__ __forceinline__ cplx csqr(const cplx a) { return cmake(fma(a.x, a.x, -(a.y * a.y)), (a.x + a.x) * a.y); }

// BS: stride (in cplx) between consecutive buffer slots -- 2 when two transforms are interleaved slot by slot
// TW: 0 = six twiddle loads per radix-16 pass (w^1..3, w^4, w^8, w^12); 1 = two loads (w^1, w^4), the other four by
// squaring / multiplication (14 more fp64 instructions, 4 fewer 16-byte loads per thread and pass)
template <int LOGN, int P, int BS = 1, int TW = 0>
struct FftPass {
  using C = FftCfg<LOGN>;
  template <class HOOK>
  static __device__ __forceinline__ void run(cplx (&x)[C::E], cplx* __restrict__ buf, const cplx* __restrict__ tw,
                                             int t, int opaque_zero, const HOOK& buffer_free) {
    constexpr int R = C::radix(P);
    constexpr int NS = C::ns(P);
    constexpr int Q = C::E / R;
    constexpr int T = C::T;
    if constexpr (R == 16) {
      // one radix-16 butterfly per thread (Q == 1), processed column by column: load 4 points (+ their twiddles),
      // column DFT4, next column; then the second half of the butterfly
      static_assert(Q == 1, "radix-16 passes use E == 16");
      const int k = t & (NS - 1);
      const cplx* twp = tw + C::tw_off(P) + k;
      if constexpr (P > 0) {
#pragma unroll
        for (int m = 0; m < 16; m++) x[m] = buf[fft_pad(t + T * m) * BS];
      }
      hook_at<P, 0>(buffer_free);
      ADEPT_TRACE(10 * P + 0);
      // Twiddles: only w^1, w^2, w^3 and w^4, w^8, w^12 are loaded (6 of the 15 powers).  u[c + 4j] = x[c + 4j] w^(c + 4j)
      // = w^c (x[c + 4j] (w^4)^j): the inputs are scaled by (w^4)^j, the column DFT4 runs, and its four outputs are
      // scaled by w^c.  24 complex multiplications instead of 15, but 9 fewer 16-byte loads per butterfly: the L1/LSU
      // data pipe is the busiest unit of the spectral pushes (about 30 % of its time went to twiddle loads), the fp64
      // pipe has the headroom.  (With 15 loads ptxas also hoisted all of them and spilled.)
      if constexpr (NS > 1) {
        cplx w4, w8, w12, wc[4];
        if constexpr (TW == 1) {
          wc[1] = __ldg(twp), w4 = __ldg(twp + 3 * NS);
          wc[2] = csqr(wc[1]), w8 = csqr(w4);
          wc[3] = cmul(wc[1], wc[2]), w12 = cmul(w4, w8);
        } else {
          w4 = __ldg(twp + 3 * NS), w8 = __ldg(twp + 7 * NS), w12 = __ldg(twp + 11 * NS);
#pragma unroll
          for (int c = 1; c < 4; c++) wc[c] = __ldg(twp + (c - 1) * NS);
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {
          x[c + 4] = cmul(x[c + 4], w4);
          x[c + 8] = cmul(x[c + 8], w8);
          x[c + 12] = cmul(x[c + 12], w12);
          dft4(x[c], x[c + 4], x[c + 8], x[c + 12]);
          if (c > 0) {
#pragma unroll
            for (int j = 0; j < 4; j++) x[c + 4 * j] = cmul(x[c + 4 * j], wc[c]);
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < 4; c++) dft4(x[c], x[c + 4], x[c + 8], x[c + 12]);
      }
      (void)opaque_zero;
      // every read of buf by this thread is complete: the barrier that protects the exchange buffer goes here, so the
      // second half of the butterfly overlaps (fp64 pipe) with the stores of the exchange (LSU pipe)
      ADEPT_TRACE(10 * P + 1);
      if constexpr (P < C::NPASS - 1) __syncthreads();
      if constexpr (P == C::NPASS - 1) buffer_free();
      hook_at<P, 1>(buffer_free);
      ADEPT_TRACE(10 * P + 2);
      dft16_finish(x);
      ADEPT_TRACE(10 * P + 3);
    } else {
      if constexpr (P > 0) {
#pragma unroll
        for (int m = 0; m < C::E; m++) x[m] = buf[fft_pad(t + T * m) * BS];
      }
#pragma unroll
      for (int q = 0; q < Q; q++) {
        const int k = (t + T * q) & (NS - 1);
        cplx v[R];
#pragma unroll
        for (int r = 0; r < R; r++) v[r] = x[q + r * Q];
        if constexpr (NS > 1) {
          const cplx* twp = tw + C::tw_off(P) + k;
#pragma unroll
          for (int r = 1; r < R; r++) v[r] = cmul(v[r], __ldg(twp + (r - 1) * NS));
        }
        dft_r<R>(v);
#pragma unroll
        for (int r = 0; r < R; r++) x[q + r * Q] = v[r];
      }
      if constexpr (P == C::NPASS - 1) buffer_free();
    }
    if constexpr (P < C::NPASS - 1) {
      if constexpr (R != 16) __syncthreads();  // all reads of buf for this pass (and any earlier use) are done
#pragma unroll
      for (int q = 0; q < Q; q++) {
        const int b = t + T * q;
        const int k = b & (NS - 1);
        const int j0 = (b - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; r++) buf[fft_pad(j0 + r * NS) * BS] = x[q + r * Q];
      }
      ADEPT_TRACE(10 * P + 4);
      __syncthreads();
      ADEPT_TRACE(10 * P + 5);
      FftPass<LOGN, P + 1, BS, TW>::run(x, buf, tw, t, opaque_zero, buffer_free);
    }
  }
};

// Pull this thread's twiddles of every pass towards L1 (single-CTA kernels start with a cold table).
template <int LOGN>
__device__ __forceinline__ void fft_prefetch_twiddles(const cplx* tw, int t) {
  using C = FftCfg<LOGN>;
  if constexpr (C::NP16 >= 2) {
#pragma unroll
    for (int p = 1; p < C::NP16; p++) {
      const int ns = C::ns(p);
      const cplx* twp = tw + C::tw_off(p) + (t & (ns - 1));
#pragma unroll
      for (int r = 1; r < 16; r++) asm volatile("prefetch.global.L1 [%0];" ::"l"(twp + (r - 1) * ns));
    }
  }
}

// Forward complex FFT of the N points held as x[m] = z[t + T*m]; result X[t + T*m] in x[m].
// All threads of the CTA must call this together (it uses __syncthreads()).
template <int LOGN, int BS = 1, int TW = 0, class HOOK = NoHook>
__device__ __forceinline__ void fft_forward(cplx (&x)[FftCfg<LOGN>::E], cplx* buf, const cplx* tw, int t,
                                            int opaque_zero, const HOOK& buffer_free = HOOK()) {
  FftPass<LOGN, 0, BS, TW>::run(x, buf, tw, t, opaque_zero, buffer_free);
}

}  // namespace adept
