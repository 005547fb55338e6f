// Field solve in the tail of a persistent x-advection launch (vdfdx_tma.cu, vdfdx_dual.cu): see FieldTail in internal.h.
//
// Reference semantics: compute_charge_density + SpectralPoissonSolver.__call__ (adept/_vlasov1d/solvers/pushers/
// field.py:197-224), the ponderomotive force (field.py:495) and the longitudinal driver (field.py:21-33).
//
// After its last tile every persistent CTA has written its partial row sums to partial[blockIdx.x][0..N).  All CTAs
// then (1) join a device-wide ticket barrier, (2) sum their slice of x over the partial rows in a fixed order and write
// rho, pond and the driver field there, (3) join a second barrier, (4) load rho into shared memory and evaluate
// E_i = sum_j green[(i - j) mod N] rho_j for their slice (field.py:221-224 is linear in rho, so the circular
// convolution with green = Re ifft(-i / kx) is the same operator).  The launch must be cooperative (all CTAs resident).
#pragma once
#include "common.cuh"
#include "internal.h"

namespace adept {

// scratch: at least 3 N + 8 NGRP doubles of dead shared memory (rho[N] | green[2 N] | partial outputs); red: NGRP * 32 doubles.
// Every thread of the CTA must call it.  THREADS must be a multiple of 128 (NGRP a multiple of 4).
template <int N, int THREADS>
__device__ __forceinline__ void field_tail_solve(const FieldTail& ft, const double* __restrict__ partial,
                                                 double* scratch, double* red) {
  static_assert(THREADS % 128 == 0 && N % THREADS == 0, "field tail: thread count");
  const int tid = threadIdx.x;
  const unsigned int G = gridDim.x;
  double* rho_s = scratch;  // rho[N] | green[N] twice in a row
  double* g_s = rho_s + N;
  constexpr int NGRP = THREADS / 32;
  constexpr int PER_T = N / THREADS;
  const int colr = tid & 31, grp = tid >> 5;  // warp `grp` sums the partial rows grp, grp + NGRP, ...
  grid_arrive(ft.counter);  // this CTA's partial row is complete
  {  // the Green's function does not depend on the other CTAs: fetched while the barrier fills
#pragma unroll
    for (int u0 = 0; u0 < PER_T; u0 += 8) {
      double gv[8];
#pragma unroll
      for (int u = 0; u < 8; u++) gv[u] = (u0 + u < PER_T) ? __ldg(ft.green + tid + (u0 + u) * THREADS) : 0.0;
#pragma unroll
      for (int u = 0; u < 8; u++)
        if (u0 + u < PER_T) {  // twice in a row: a window of the convolution never wraps
          g_s[tid + (u0 + u) * THREADS] = gv[u];
          g_s[tid + (u0 + u) * THREADS + N] = gv[u];
        }
    }
  }
  grid_wait(ft.counter, G);  // every CTA's partial row is complete
  const int per = (N + (int)G - 1) / (int)G;
  const int i_lo = (int)blockIdx.x * per, i_hi = min(N, i_lo + per);
  for (int c0 = i_lo; c0 < i_hi; c0 += 32) {
    const int i = c0 + colr;
    double s0 = 0.0, s1 = 0.0;
    if (i < i_hi) {
      for (unsigned int q0 = grp; q0 < G; q0 += 10 * NGRP) {
        double x[10];
#pragma unroll
        for (int u = 0; u < 10; u++) {
          const unsigned int q = q0 + u * NGRP;
          x[u] = q < G ? __ldcg(partial + (size_t)q * N + i) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 10; u += 2) s0 += x[u], s1 += x[u + 1];
      }
    }
    __syncthreads();
    red[grp * 32 + colr] = s0 + s1;
    __syncthreads();
    if (grp == 0 && i < i_hi) {
      double tot = 0.0;
#pragma unroll
      for (int g2 = 0; g2 < NGRP; g2++) tot += red[g2 * 32 + colr];
      const double term = __dmul_rn(ft.charge, __dmul_rn(tot, ft.dv));  // field.py:197-208
      const double mine = ft.base ? __dadd_rn(ft.base[i], term) : term;
      if (ft.n_peers > 1) {  // my share of rho at i goes to slot my_rank of every rank's inbox
        const size_t slot = ((size_t)(ft.epoch & 1) * ft.n_peers + ft.my_rank) * N + i;
        for (int r = 0; r < ft.n_peers; r++) ft.share_in[r][slot] = mine;
        __threadfence_system();
      } else {
        ft.rho[i] = mine;
      }
    }
    if (grp == 1 && i < i_hi) {
      const double lo = __dmul_rn(ft.a[i], ft.a[i]), hi = __dmul_rn(ft.a[i + 2], ft.a[i + 2]);
      ft.pond[i] = __dmul_rn(-0.5, __ddiv_rn(__dsub_rn(hi, lo), __dmul_rn(2.0, ft.dx)));  // field.py:495
    }
    if (grp == 2 % NGRP && i < i_hi && ft.n_ex > 0) {  // field.py:21-33
      double total = 0.0;
      for (int d = 0; d < ft.n_ex; d++) {
        const double tenv = ft.trow ? ft.trow[TROW_TENV + d] : ft.ex_tenv[d];
        const double wt = ft.trow ? ft.trow[TROW_WT + d] : ft.ex_wt[d];
        const double factor = __dmul_rn(tenv, ft.ex_space[(size_t)d * N + i]);
        const double amp = __dmul_rn(__dmul_rn(factor, ft.ex_w[d]), ft.ex_a0[d]);
        total = __dadd_rn(total, __dmul_rn(amp, sin(__dsub_rn(ft.ex_kx[(size_t)d * N + i], wt))));
      }
      ft.dex[i] = total;
    }
  }
  grid_barrier(ft.counter, 2 * G);  // rho (or this rank's share of it, pushed to every inbox) complete on every CTA
  unsigned int nbar = 3;     // arrivals per CTA on the ticket counter by the end of the tail
  if (ft.n_peers > 1) {
    nbar = 4;
    if (blockIdx.x == 0 && tid < ft.n_peers) {  // all CTAs of this rank have fenced their pushes (barrier above)
      __threadfence_system();
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(ft.flag_in[tid] + ft.my_rank), "l"(ft.epoch) : "memory");
    }
    if (tid < ft.n_peers) {  // every CTA watches the rank's own inbox flags (local memory)
      const unsigned long long* fl = ft.flag_in[ft.my_rank] + tid;
      unsigned long long seen;
      do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(fl) : "memory");
      } while (seen < ft.epoch);
    }
    __syncthreads();
    // rho at my slice = the P shares in rank order (identical on every rank); L2 loads: the inbox was written by peers
    const double* inbox = ft.share_in[ft.my_rank] + (size_t)(ft.epoch & 1) * ft.n_peers * N;
    for (int i = i_lo + tid; i < i_hi; i += THREADS) {
      double x[8];
#pragma unroll
      for (int r = 0; r < 8; r++) x[r] = r < ft.n_peers ? __ldcg(inbox + (size_t)r * N + i) : 0.0;
      double tot = x[0];
#pragma unroll
      for (int r = 1; r < 8; r++)
        if (r < ft.n_peers) tot = __dadd_rn(tot, x[r]);
      ft.rho[i] = tot;
    }
    grid_barrier(ft.counter, 3 * G);  // rho complete on every CTA
  }
  {
#pragma unroll
    for (int u0 = 0; u0 < PER_T; u0 += 8) {
      double rv[8];
#pragma unroll
      for (int u = 0; u < 8; u++) rv[u] = (u0 + u < PER_T) ? __ldcg(ft.rho + tid + (u0 + u) * THREADS) : 0.0;
#pragma unroll
      for (int u = 0; u < 8; u++)
        if (u0 + u < PER_T) rho_s[tid + (u0 + u) * THREADS] = rv[u];
    }
  }
  __syncthreads();
  // E_i = sum_j green[(i - j) mod N] rho_j.  Fetching both operands of every product costs 16 bytes of shared memory
  // per multiply-add (the largest item of the tail), so where the slice allows it the sum is register-tiled like
  // poisson_green_kernel (field.cu): a lane owns two consecutive j and eight consecutive outputs, 6 aligned 16-byte
  // loads per 16 multiply-adds; the warps split into 4 output groups x NGRP/4 ranges of j.
  if ((per & 1) == 0 && per <= 32) {
    constexpr int NJR = NGRP / 4;
    double* part = rho_s + 3 * N;  // [NJR j ranges][32 outputs]
    const int og = grp & 3, jr = grp >> 2;
    const int i0 = i_lo + 8 * og;
    double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (i0 < i_hi) {
      for (int j = jr * (N / NJR) + 2 * colr; j < (jr + 1) * (N / NJR); j += 64) {
        const double2 r2 = *reinterpret_cast<const double2*>(rho_s + j);
        // g2 index of (output i0 + r, column j + d): m0 + r + 2 - d with m0 = i0 - j - 2 + N (even)
        const double2* wp = reinterpret_cast<const double2*>(g_s + (i0 - j - 2 + N));
        double w[10];
#pragma unroll
        for (int q = 0; q < 5; q++) {
          const double2 t2 = wp[q];
          w[2 * q] = t2.x, w[2 * q + 1] = t2.y;
        }
#pragma unroll
        for (int r = 0; r < 8; r++) acc[r] = fma(w[r + 1], r2.y, fma(w[r + 2], r2.x, acc[r]));
      }
    }
#pragma unroll
    for (int r = 0; r < 8; r++) acc[r] = warp_sum(acc[r]);
    if (colr == 0) {
#pragma unroll
      for (int r = 0; r < 8; r++) part[jr * 32 + og * 8 + r] = acc[r];
    }
    __syncthreads();
    if (tid < 32 && i_lo + tid < i_hi) {
      double e = 0.0;
      if constexpr (NJR == 4) {
        e = (part[tid] + part[32 + tid]) + (part[64 + tid] + part[96 + tid]);
      } else {
#pragma unroll
        for (int q = 0; q < NJR; q++) e += part[q * 32 + tid];
      }
      ft.e[i_lo + tid] = e;
    }
  } else {
    for (int i = i_lo + grp; i < i_hi; i += NGRP) {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 4
      for (int j = colr; j < N; j += 128) {  // N is a multiple of 128
        a0 = fma(g_s[(i - j) & (N - 1)], rho_s[j], a0);
        a1 = fma(g_s[(i - j - 32) & (N - 1)], rho_s[j + 32], a1);
        a2 = fma(g_s[(i - j - 64) & (N - 1)], rho_s[j + 64], a2);
        a3 = fma(g_s[(i - j - 96) & (N - 1)], rho_s[j + 96], a3);
      }
      const double e = warp_sum((a0 + a1) + (a2 + a3));
      if (colr == 0) ft.e[i] = e;
    }
  }
  __syncthreads();
  if (tid == 0 && atomicAdd(ft.counter, 1u) == nbar * G - 1) *ft.counter = 0u;  // last one out re-arms the counter
}

}  // namespace adept
