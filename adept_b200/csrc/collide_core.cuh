// Shared device code of the collision kernels (collide.cu, vrow.cu): fast reciprocal, row reductions and the
// fast-path Fokker-Planck row solve (central differencing, Lenard-Bernstein / Dougherty) on a row held in shared memory.
// Reference semantics: adept/_vlasov1d/solvers/pushers/fokker_planck.py:368-433, adept/driftdiffusion.py:106-137,
// 359-378, 563-600 (see collide.cu for the formulation).
#pragma once
#include "common.cuh"

namespace adept {

enum { FP_LB = 0, FP_DOUGHERTY = 1, FP_SUPERGAUSSIAN = 2 };
enum { FP_CENTRAL = 0, FP_CHANG_COOPER = 1 };

#ifndef ADEPT_F32_BUILD  // the generated fp32 build takes both from common32.cuh
// couplings of the unit-diagonal reduced system below 2^-56 (an eighth of the fp64 machine epsilon) cannot change
// the solution at rounding level: parallel cyclic reduction stops there
#define PCR_TOL 1.3877787807814457e-17

// 1/x for |x| in the normal range: MUFU.RCP64H seed (about 20 bits) + two Newton steps -> rounding-level accuracy.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}
#endif  // ADEPT_F32_BUILD

// Chang-Cooper delta on the fast path: for |w| < 1/4 (w = C dv / D = dv (v_edge - vbar) / T, a few 1e-2 on production
// grids) the Bernoulli series of 1/w - 1/(e^w - 1) through w^11 (next term < 2e-19) replaces expm1 and two reciprocals
// -- and is free of the cancellation the closed form has at small w.
__device__ __forceinline__ double cc_delta_fast(double w) {
  if (fabs(w) < 0.25) {
    const double w2 = w * w;
    double p = 691.0 / 1307674368000.0;
    p = fma(p, w2, -1.0 / 47900160.0);
    p = fma(p, w2, 1.0 / 1209600.0);
    p = fma(p, w2, -1.0 / 30240.0);
    p = fma(p, w2, 1.0 / 720.0);
    p = fma(p, w2, -1.0 / 12.0);
    return fma(p, w, 0.5);
  }
  return fast_rcp(w) - fast_rcp(expm1(w));
}

// d delta / d w of the Chang-Cooper weight delta(w) = 1/w - 1/(e^w - 1) (adjoint kernels): series through w^10 for
// |w| < 0.3 (next term < 1e-16), closed form -1/w^2 + e^w / (e^w - 1)^2 beyond.
__device__ __forceinline__ double cc_delta_prime_fast(double w) {
  if (fabs(w) < 0.3) {
    const double w2 = w * w;
    double p = 691.0 / 118879488000.0;      // 11 * 691 / 1307674368000
    p = fma(p, w2, -1.0 / 5322240.0);       // -9 / 47900160
    p = fma(p, w2, 1.0 / 172800.0);         // 7 / 1209600
    p = fma(p, w2, -1.0 / 6048.0);          // -5 / 30240
    p = fma(p, w2, 1.0 / 240.0);            // 3 / 720
    return fma(p, w2, -1.0 / 12.0);
  }
  const double em = expm1(w);
  return (em + 1.0) / (em * em) - 1.0 / (w * w);
}

// Sum NVAL values over the T threads of row r; every thread of the CTA must call it.  warp_mode: T % 32 == 0 (warps do
// not straddle rows).  `red` is a scratch area of 2 * S * NVAL doubles used with alternating halves, S = 32 in warp
// mode and max(R*T, 32) otherwise.
template <int NVAL>
__device__ __forceinline__ void row_reduce(double (&val)[NVAL], double* red, int& parity, int r, int t, int T, int RT,
                                           bool warp_mode, bool live) {
  double* base = red + (size_t)parity * (warp_mode ? (RT >> 5) : (RT > 32 ? RT : 32)) * NVAL;  // warp mode: RT/32 slots
  parity ^= 1;
  if (warp_mode) {
#pragma unroll
    for (int i = 0; i < NVAL; i++) val[i] = warp_sum(val[i]);
    const int nw = T >> 5, w = t >> 5;
    double* slot = base + (size_t)r * nw * NVAL;
    if ((t & 31) == 0) {
#pragma unroll
      for (int i = 0; i < NVAL; i++) slot[w * NVAL + i] = val[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NVAL; i++) val[i] = 0.0;
    for (int j = 0; j < nw; j++) {
#pragma unroll
      for (int i = 0; i < NVAL; i++) val[i] += slot[j * NVAL + i];
    }
  } else {
    double* slot = base + (size_t)r * T * NVAL;
    if (live) {
#pragma unroll
      for (int i = 0; i < NVAL; i++) slot[t * NVAL + i] = val[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NVAL; i++) val[i] = 0.0;
    for (int j = 0; j < T; j++) {
#pragma unroll
      for (int i = 0; i < NVAL; i++) val[i] += slot[j * NVAL + i];
    }
  }
}

// Fast-path Fokker-Planck solve of ONE row by the T = nv/E threads of a CTA (all threads of the CTA must call it; it
// uses __syncthreads()).  rowbuf: the row in chunk-padded layout (cell i at i + i/E), overwritten with f + delta.
// red: 2*32*3 doubles when T % 32 == 0 (else 2*T*3), pcr: 6 T doubles.
// DENSE_OUT: the result is written as a dense row whose 128-byte chunks are XOR-swizzled in 16-byte units (unit j of
// chunk c at j ^ (c & 7)): the layout a TMA tensor store with CU_TENSOR_MAP_SWIZZLE_128B reads, conflict-free for the
// 16-byte stores of a warp; rowbuf must then be 1024-byte aligned.  Each thread fences its stores for the async proxy.
// CC: Chang-Cooper weighting of the drag term instead of central differencing.
// Velocity moments of the calling thread's chunk of a chunk-padded row: {sum f, sum f v, sum f v^2} with v = vc + l dv
// (index-space sums, shifted once).  Callers reduce them over the row with row_reduce.
template <int E>
__device__ __forceinline__ void fp_chunk_moments(const double* rowbuf, int tt, double vc, double dv, double* mom) {
  const double* chunk = rowbuf + E * tt + tt;
  double m0 = 0.0, m1 = 0.0, m2 = 0.0;
#pragma unroll
  for (int l = 0; l < E; l++) {
    const double fl = chunk[l];
    m0 += fl;
    m1 = fma(fl, (double)l, m1);
    m2 = fma(fl, (double)(l * l), m2);
  }
  m1 *= dv, m2 *= dv * dv;
  mom[0] = m0;
  mom[1] = fma(vc, m0, m1);
  mom[2] = fma(vc * vc, m0, fma(2.0 * vc, m1, m2));
}

// `sums` (nullable): the row's three moments, already reduced over the row by the caller (the fused v-row kernel
// reduces both rows of its pair behind one barrier).
template <int E, bool DENSE_OUT = false, bool CC = false>
__device__ __forceinline__ void fp_row_fast(double* rowbuf, double* red, double* pcr, int& parity, int tt,
                                            int T, int nv, double vc, double dv, double dt, double nu, int model,
                                            const double* sums = nullptr) {
  const bool warp_mode = (T & 31) == 0;
  const int i0 = E * tt;
  const double* chunk = rowbuf + i0 + tt;
  double mom[3];
  if (sums) {
    mom[0] = sums[0], mom[1] = sums[1], mom[2] = sums[2];
  } else {
    fp_chunk_moments<E>(rowbuf, tt, vc, dv, mom);
    row_reduce<3>(mom, red, parity, 0, tt, T, T, warp_mode, true);
  }
  const double s0 = mom[0], s1 = mom[1], s2 = mom[2];
  const double vbar = (model == FP_LB) ? 0.0 : s1 / s0;
  const double Temp = (s2 - 2.0 * vbar * s1 + vbar * vbar * s0) / s0;
  const double beta = 1.0 / (2.0 * Temp);
  const double D = 1.0 / (2.0 * beta);
  const double dtnu = dt * nu;
  const double pD = dtnu * D / (dv * dv);
  const double q = dtnu * (2.0 * beta * D) / (2.0 * dv);
  const double w0 = q * (vc + 0.5 * dv - vbar), dq = q * dv;
  const double ww0 = (2.0 * beta * dv) * (vc + 0.5 * dv - vbar), dww = (2.0 * beta * dv) * dv;  // Chang-Cooper w
  auto edge = [&](int l, double& U, double& L) {
    const int e = i0 + l;
    if (e < 0 || e > nv - 2) {
      U = 0.0;
      L = 0.0;
      return;
    }
    const double wq = fma((double)l, dq, w0);  // dt nu C / (2 dv)
    if (!CC) {
      U = pD + wq;
      L = pD - wq;
    } else {  // U = dt nu (C (1 - delta) + D/dv) / dv, L = dt nu (-C delta + D/dv) / dv, delta(w = C dv / D)
      const double dl = cc_delta_fast(fma((double)l, dww, ww0));
      const double cq = 2.0 * wq;
      U = fma(cq, 1.0 - dl, pD);
      L = fma(-cq, dl, pD);
    }
  };
  double cpn[E], ypn[E], apn[E];  // Thomas ratios, y-form right-hand sides and the left spike, all in registers
  double rpn_last, apn_last;
  {
    double Um, Lm;
    edge(-1, Um, Lm);
    double f_m = (tt > 0) ? rowbuf[i0 - 1 + (tt - 1)] : 0.0;
    double f_c = chunk[0];
    double G_m = Um * f_c - Lm * f_m;
    double cp_prev = 0.0, ap_prev = 0.0, rp_prev = 0.0;
#pragma unroll
    for (int l = 0; l < E; l++) {
      double Uc, Lc;
      edge(l, Uc, Lc);
      const double f_p = (l < E - 1) ? chunk[l + 1] : ((tt < T - 1) ? rowbuf[i0 + E + (tt + 1)] : 0.0);
      const double G_c = Uc * f_p - Lc * f_c;
      const double rhs = G_c - G_m;
      const double a = -Lm;
      const double b = (1.0 + Lc) + Um;
      double bp, apv, rpv;
      if (l == 0) {
        bp = b, apv = a, rpv = rhs;
      } else {
        bp = fma(-a, cp_prev, b);
        apv = -a * ap_prev;
        rpv = fma(-a, rp_prev, rhs);
      }
      const double inv = fast_rcp(bp);
      const double cp = -Uc * inv, ap = apv * inv, rp = rpv * inv;
      cpn[l] = cp;
      apn[l] = ap;
      ypn[l] = (l < E - 1) ? fma(cp, f_p, f_c + rp) : f_c;
      if (l == E - 1) rpn_last = rp, apn_last = ap;
      cp_prev = cp, ap_prev = ap, rp_prev = rp;
      Um = Uc, Lm = Lc, G_m = G_c, f_m = f_c, f_c = f_p;
    }
  }
  double A0, C0, R0;
  {
    double RY = ypn[E - 2], A = apn[E - 2], Cc = cpn[E - 2];
#pragma unroll
    for (int l = E - 3; l >= 0; l--) {
      const double cp = cpn[l];
      RY = fma(-cp, RY, ypn[l]);
      A = fma(-cp, A, apn[l]);
      Cc = -cp * Cc;
    }
    A0 = A, C0 = Cc;
    R0 = RY - chunk[0] - Cc * chunk[E - 1];
  }
  double* xb = pcr + 3 * T;
  xb[tt] = A0, xb[T + tt] = C0, xb[2 * T + tt] = R0;
  __syncthreads();
  double al = apn_last, ga = 0.0, rh = rpn_last;
  {
    double be = 1.0;
    if (tt < T - 1) {
      const double k = cpn[E - 1];
      be = fma(-k, xb[tt + 1], 1.0);
      ga = -k * xb[T + tt + 1];
      rh = fma(-k, xb[2 * T + tt + 1], rpn_last);
    }
    const double ib = fast_rcp(be);
    al *= ib, ga *= ib, rh *= ib;
  }
  double* cur = pcr;
  double* nxt = pcr + 3 * T;
  __syncthreads();
  cur[tt] = al, cur[T + tt] = ga, cur[2 * T + tt] = rh;
  // The reduced system has unit diagonal and couplings (al, ga) that are products of E Thomas ratios of a
  // diagonally dominant row, so they are usually far below rounding (about 1e-17 at dt nu D / dv^2 = 0.1) and every
  // PCR step squares them: the loop stops as soon as all couplings of the row are below PCR_TOL (a vote on the
  // barrier the exchange needs anyway).  Strongly collisional rows still take all log2(T) steps.
  int more = __syncthreads_or(fabs(al) > PCR_TOL || fabs(ga) > PCR_TOL);
  for (int s = 1; s < T && more; s <<= 1) {
    double alj = 0.0, gaj = 0.0, rhj = 0.0, alk = 0.0, gak = 0.0, rhk = 0.0;
    if (tt - s >= 0) alj = cur[tt - s], gaj = cur[T + tt - s], rhj = cur[2 * T + tt - s];
    if (tt + s < T) alk = cur[tt + s], gak = cur[T + tt + s], rhk = cur[2 * T + tt + s];
    const double be = fma(-al, gaj, fma(-ga, alk, 1.0));
    const double ib = fast_rcp(be);
    rh = fma(-al, rhj, fma(-ga, rhk, rh)) * ib;
    al = -al * alj * ib;
    ga = -ga * gak * ib;
    nxt[tt] = al, nxt[T + tt] = ga, nxt[2 * T + tt] = rh;
    more = __syncthreads_or(fabs(al) > PCR_TOL || fabs(ga) > PCR_TOL);
    double* tmp = cur;
    cur = nxt;
    nxt = tmp;
  }
  const double s_me = rh;
  const double s_left = (tt > 0) ? cur[2 * T + tt - 1] : 0.0;
  {
    double y = ypn[E - 1] + s_me;
    if constexpr (DENSE_OUT) {
      static_assert(E == 16, "dense output layout: 16 cells = one 128-byte chunk per thread");
      double2* units = reinterpret_cast<double2*>(rowbuf) + 8 * tt;
      const int sw = tt & 7;
      double y_up = y;
#pragma unroll
      for (int l = E - 2; l >= 0; l--) {
        y = fma(-cpn[l], y, fma(-apn[l], s_left, ypn[l]));
        if ((l & 1) == 0) units[(l >> 1) ^ sw] = make_double2(y, y_up);
        y_up = y;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    } else {
      double* outc = rowbuf + i0 + tt;
      outc[E - 1] = y;
#pragma unroll
      for (int l = E - 2; l >= 0; l--) {
        y = fma(-cpn[l], y, fma(-apn[l], s_left, ypn[l]));
        outc[l] = y;
      }
    }
  }
  __syncthreads();  // pcr / red may be reused by the caller for the next row
}

}  // namespace adept
