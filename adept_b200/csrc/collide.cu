// Fused Fokker-Planck + Krook collision step: per-x moments (n, u, T), tridiagonal assembly in v, delta-form
// implicit solve, Krook relaxation and the post-collision density moment, one HBM read + one HBM write of f.
//
// Reference semantics (file:line relative to /root/reference):
//   Collisions._collide / _solve_one_x   adept/_vlasov1d/solvers/pushers/fokker_planck.py:368-433
//   LenardBernstein / Dougherty / SuperGaussianDougherty   fokker_planck.py:31-255
//   AbstractBetaBasedModel.compute_C_and_D                 adept/driftdiffusion.py:359-378
//   discrete_temperature                                   adept/driftdiffusion.py:106-137
//   CentralDifferencing / ChangCooper get_operator         adept/driftdiffusion.py:563-600 / 614-658
//   chang_cooper_delta                                     adept/driftdiffusion.py:77-103
//   Krook                                                  fokker_planck.py:446-484
//
// Formulation.  With U_e = dt nu bare_upper_e and L_e = dt nu bare_lower_e on edge e (between cells e and e+1) the
// operator A = I - dt nu L of both differencing schemes has  a_i = -L_{i-1},  c_i = -U_i,  b_i = 1 + L_i + U_{i-1}
// (zero column sums = zero-flux boundaries), and the delta-form right-hand side  f - A f  is the flux difference
// G_i - G_{i-1},  G_e = U_e f_{e+1} - L_e f_e.  For central differencing U_e, L_e are LINEAR in the edge index
// (C_e = 2 beta D (v_e - vbar)), so the fast path needs no division and no table per cell.
//
// Parallel solve: T = nv/E threads per x-row, each owning E contiguous velocity cells.  A thread eliminates its chunk
// downwards carrying the "spike" towards the previous chunk's last unknown (rows normalised to unit diagonal with a
// MUFU.RCP64H + 2 Newton steps reciprocal), an upward sweep expresses the chunk's first unknown through the two
// neighbouring chunk-last unknowns, the T chunk-last unknowns form a reduced tridiagonal system solved with parallel
// cyclic reduction in shared memory (unit-diagonal form: one reciprocal per step), then chunks back-substitute
// directly for f + delta.  The matrix is strictly diagonally dominant, so no pivoting is needed; the reference's
// LAPACK gtsv agrees to rounding.
#include "collide_core.cuh"
#include <math_constants.h>
#include "internal.h"

namespace adept {


struct CollideArgs {
  const double* fin;
  double* fout;
  long long rows;  // batch * nx
  int nv;
  const double* v;  // f64  [nv]
  double dv, dt;
  const double* nu_fp;  // f64  [rows] or null (Fokker-Planck off)
  const double* nu_K;   // f64  [rows] or null (Krook off)
  const double* f_mx;   // f64  [nv] Krook Maxwellian (unit density)
  int model, scheme, nodrag;
  double sg_m, sg_ratio;  // super-Gaussian exponent m and Gamma(3/m)/Gamma(1/m)
  double* n_out;          // [rows] or null: sum_j f_out dv
  int rows_per_cta;       // R: x-rows handled by one CTA (set by the launcher)
  double nu_fp_scale, nu_K_scale;  // nu = scale * nu[row] (time envelope applied in the kernel)
  const double* trow;              // f64  nullable device-resident time row (common.cuh) that replaces the two scales
  int sc_steps;                    // self-consistent beta: Newton iterations (0 = off), fokker_planck.py:296-301
  double sc_rtol, sc_atol;
  // vlasov-1d2v (adept/_vlasov1d2v/solvers/pushers/fokker_planck.py:104-140): the operator coefficients of a group of
  // coef_div consecutive rows (the v_perp slices of one x) come from the group's marginal, not from the row itself.
  // coef_out [rows, 2] (nullable): (vbar, beta) of every row as computed here; coef_in [rows / coef_div, 2] (nullable):
  // (vbar, beta) to use instead, and nu_fp is then indexed per group.  General (non-fast) path only.
  const double* coef_in;  // f64
  double* coef_out;       // f64
  int coef_div;
};

// d delta / d w of the Chang-Cooper weight, branch by branch as autodiff differentiates driftdiffusion.py:96-103
__device__ __forceinline__ double cc_delta_prime(double w) {
  if (fabs(w) < 1.0e-8) return -1.0 / 12.0 + w * w * (1.0 / 240.0);
  const double em = expm1(w);
  const double r = exp(w) / (em * em);
  return -1.0 / (w * w) + (isfinite(r) ? r : 0.0);
}

// One Newton update of optimistix 0.1.0's root finder for a scalar (oracle/vlasov1d.py::newton_root_find): the Cauchy
// test runs before the step and a finished row keeps its value.
struct NewtonState {
  double diff, fprev;
  bool done;
};
__device__ __forceinline__ void newton_update(double& y, NewtonState& st, double fx, double slope, double rtol,
                                              double atol) {
  st.done = st.done || (fabs(st.diff) < atol + rtol * fabs(y) && fabs(st.fprev) < atol);
  if (!st.done) {
    const double d = fx / slope;
    y -= d;
    st.diff = d;
    st.fprev = fx;
  }
}

// Reciprocals here are fast_rcp (MUFU seed + two Newton steps, <= 1 ulp) and the per-row quotients D/dv, 1/dv, dv/D are
// formed once per row by the caller: an IEEE division costs ~25 fp64 instructions and this runs once per cell edge.
__device__ __forceinline__ double cc_delta(double w) {  // driftdiffusion.py:96-103
  if (fabs(w) < 1.0e-8) return 0.5 - w * (1.0 / 12.0) + w * w * w * (1.0 / 720.0);
  return fast_rcp(w) - fast_rcp(expm1(w));
}

// bare upper / lower entries of one edge (generic path); D_dv = max(D, 1e-30) / dv, inv_dv = 1 / dv,
// dv_D = dv / max(D, 1e-30)
__device__ __forceinline__ void bare_edge(double C, double D_dv, double inv_dv, double dv_D, int scheme, double& bu,
                                          double& bl) {
  if (scheme == FP_CENTRAL) {  // driftdiffusion.py:585-590
    bu = (0.5 * C + D_dv) * inv_dv;
    bl = (-0.5 * C + D_dv) * inv_dv;
  } else {  // driftdiffusion.py:637-648
    const double w = C * dv_D;
    const double dl = cc_delta(w);
    const double alpha = -C * dl + D_dv;
    const double beta = -C * (1.0 - dl) - D_dv;
    bu = -beta * inv_dv;
    bl = alpha * inv_dv;
  }
}

template <int E>
struct CollideSmem {
  // doubles per CTA for R rows of nv cells handled by RT = R*T threads
  static __host__ __device__ size_t doubles(int R, int nv, int RT) {
    const int red = 2 * (RT > 32 ? RT : 32) * 3;
    return (size_t)R * (nv + nv / E) + red + 6 * (size_t)RT;  // the spike of the downward sweep lives in registers
  }
};

// FAST: LB / Dougherty on the uniform grid (central differencing or Chang-Cooper), no `nodrag` -- the production path.
// CC (fast path only): Chang-Cooper instead of central differencing, a separate instantiation so that the central
// kernel carries no branch or extra registers.
template <int E, int MAXT, int MINB, bool FAST, bool CC = false>
__global__ void __launch_bounds__(MAXT, MINB) collide_kernel(CollideArgs p) {
  extern __shared__ __align__(16) double sm[];
  const int nv = p.nv;
  const int T = nv / E;
  const int R = p.rows_per_cta;
  const int RT = R * T;
  const int nvp = nv + T;  // one pad double per chunk: chunk stride E+1 (odd) -> conflict-free 8-byte accesses
  const int r = threadIdx.x / T, t = threadIdx.x - r * T;
  const bool live = r < R;  // blockDim.x may exceed R*T when T does not divide it; spare threads only hit barriers
  double* rowbuf = sm + (size_t)(live ? r : 0) * nvp;
  double* red = sm + (size_t)R * nvp;
  double* pcr = red + 2 * (RT > 32 ? RT : 32) * 3;  // 2 buffers x 3 arrays x RT
  const bool warp_mode = (T & 31) == 0;
  int parity = 0;

  const long long row_raw = (long long)blockIdx.x * R + r;
  const bool active = live && row_raw < p.rows;
  const long long row = active ? row_raw : p.rows - 1;
  const int tt = live ? t : 0;
  const int me = (live ? r : 0) * T + tt;
  const double dv = p.dv, dt = p.dt;

  // ---- 1. row -> shared (coalesced 16-byte loads when aligned) ---------------------------------------------------
  {
    const double* fin = p.fin + row * nv;
    if (live) {
      if ((nv & 1) == 0 && ((reinterpret_cast<uintptr_t>(fin) & 15) == 0)) {
        const double2* f2 = reinterpret_cast<const double2*>(fin);
        for (int i = t; i < (nv >> 1); i += T) {
          const double2 x = f2[i];
          const int j = 2 * i;
          rowbuf[j + j / E] = x.x;
          rowbuf[j + 1 + (j + 1) / E] = x.y;
        }
      } else {
        for (int i = t; i < nv; i += T) rowbuf[i + i / E] = fin[i];
      }
    }
  }
  __syncthreads();
  const int i0 = E * tt;
  const double* chunk = rowbuf + i0 + tt;  // chunk[l] = f[i0 + l]
  const double vc = __ldg(p.v + i0);        // v of the chunk's first cell; v[i0 + l] = vc + l dv (uniform grid)

  if (p.nu_fp) {
    const long long grp = p.coef_in ? row / p.coef_div : row;
    const double nu = __dmul_rn(p.trow ? p.trow[TROW_NU_FP] : p.nu_fp_scale, p.nu_fp[grp]);
    // ---- 2. moments in chunk-local index space: sum f, sum f l, sum f l^2 ------------------------------------------
    double mom[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int l = 0; l < E; l++) {
      const double fl = chunk[l];
      mom[0] += fl;
      mom[1] = fma(fl, (double)l, mom[1]);
      mom[2] = fma(fl, (double)(l * l), mom[2]);
    }
    if (FAST) {
      // sum f v = vc m0 + dv m1 ; sum f v^2 = vc^2 m0 + 2 vc dv m1 + dv^2 m2
      const double m0 = mom[0], m1 = mom[1] * dv, m2 = mom[2] * (dv * dv);
      mom[1] = fma(vc, m0, m1);
      mom[2] = fma(vc * vc, m0, fma(2.0 * vc, m1, m2));
    } else {  // generic path: the reference's own summation (fokker_planck.py:88, driftdiffusion.py:125-137)
      mom[1] = 0.0;
#pragma unroll
      for (int l = 0; l < E; l++) mom[1] += chunk[l] * __ldg(p.v + i0 + l);
    }
    if (!live) mom[0] = mom[1] = mom[2] = 0.0;
    row_reduce<3>(mom, red, parity, live ? r : 0, tt, T, RT, warp_mode, live);
    const double s0 = mom[0], s1 = mom[1], s2 = mom[2];
    const double vbar = (!FAST && p.coef_in) ? p.coef_in[2 * grp] : ((p.model == FP_LB) ? 0.0 : s1 / s0);
    double beta, D;
    if (!FAST && p.coef_in) {  // coefficients of the group's marginal (computed by an earlier launch with coef_out)
      beta = p.coef_in[2 * grp + 1];
      D = 1.0 / (2.0 * beta);
    } else if (!FAST && p.model == FP_SUPERGAUSSIAN) {
      double sp[1] = {0.0};
#pragma unroll
      for (int l = 0; l < E; l++) sp[0] += chunk[l] * pow(fabs(__ldg(p.v + i0 + l) - vbar), p.sg_m);
      if (!live) sp[0] = 0.0;
      row_reduce<1>(sp, red, parity, live ? r : 0, tt, T, RT, warp_mode, live);
      beta = s0 / (p.sg_m * sp[0]);
      if (p.sc_steps > 0) {
        // SuperGaussianDougherty.compute_beta, fokker_planck.py:195-207: Newton on the discrete energy-flux condition
        // h(beta) = sum_e v_e (w ftilde(w) + df), w = beta dpsi, over the edges e = i0 + l owned by this chunk
        NewtonState st = {CUDART_INF, CUDART_INF, false};
        for (int it = 0; it < p.sc_steps; it++) {
          double hs[2] = {0.0, 0.0};
#pragma unroll
          for (int l = 0; l < E; l++) {
            const int e = i0 + l;
            if (e <= nv - 2) {
              const double va = __ldg(p.v + e), vb = __ldg(p.v + e + 1);
              const double fa = chunk[l], fb = (l < E - 1) ? chunk[l + 1] : rowbuf[i0 + E + (tt + 1)];
              const double dpsi = pow(fabs(vb - vbar), p.sg_m) - pow(fabs(va - vbar), p.sg_m);
              const double w = beta * dpsi;
              double dl;
              if (fabs(w) < 1.0e-8) dl = 0.5 - w * (1.0 / 12.0) + w * w * w * (1.0 / 720.0);
              else dl = 1.0 / w - 1.0 / expm1(w);
              const double ft = dl * fa + (1.0 - dl) * fb;
              const double ve = 0.5 * (vb + va);
              hs[0] += ve * (w * ft + (fb - fa));
              hs[1] += ve * dpsi * (ft + w * cc_delta_prime(w) * (fa - fb));
            }
          }
          if (!live) hs[0] = hs[1] = 0.0;
          row_reduce<2>(hs, red, parity, live ? r : 0, tt, T, RT, warp_mode, live);
          newton_update(beta, st, hs[0], hs[1], p.sc_rtol, p.sc_atol);
        }
      }
      D = pow(beta, -2.0 / p.sg_m) * p.sg_ratio;
    } else {
      double Temp;
      if (FAST) {
        // T = sum f (v - vbar)^2 / sum f  (driftdiffusion.py:125-137; dv cancels)
        Temp = (s2 - 2.0 * vbar * s1 + vbar * vbar * s0) / s0;
      } else {
        double tm[2] = {0.0, 0.0};
#pragma unroll
        for (int l = 0; l < E; l++) {
          const double vs = __ldg(p.v + i0 + l) - vbar;
          tm[0] += chunk[l] * (vs * vs) * dv;
          tm[1] += chunk[l] * dv;
        }
        if (!live) tm[0] = tm[1] = 0.0;
        row_reduce<2>(tm, red, parity, live ? r : 0, tt, T, RT, warp_mode, live);
        Temp = tm[0] / tm[1];
      }
      beta = 1.0 / (2.0 * Temp);
      if (!FAST && p.sc_steps > 0) {
        // find_self_consistent_beta, driftdiffusion.py:161-233: beta* whose sampled Maxwellian exp(-beta (v - vbar)^2)
        // has the discrete temperature of f; slope = d(v2 / norm) / d beta
        NewtonState st = {CUDART_INF, CUDART_INF, false};
        for (int it = 0; it < p.sc_steps; it++) {
          double ms[3] = {0.0, 0.0, 0.0};
#pragma unroll
          for (int l = 0; l < E; l++) {
            const double vs = __ldg(p.v + i0 + l) - vbar;
            const double sq = vs * vs;
            const double fm = exp(-beta * sq);
            ms[0] += fm * dv;
            ms[1] += fm * sq * dv;
            ms[2] += fm * sq * sq * dv;
          }
          if (!live) ms[0] = ms[1] = ms[2] = 0.0;
          row_reduce<3>(ms, red, parity, live ? r : 0, tt, T, RT, warp_mode, live);
          const double norm = ms[0], v2 = ms[1];
          newton_update(beta, st, v2 / norm - Temp, (-ms[2] * norm + v2 * ms[1]) / (norm * norm), p.sc_rtol, p.sc_atol);
        }
      }
      D = 1.0 / (2.0 * beta);
    }
    if (!FAST && p.coef_out && tt == 0 && active) p.coef_out[2 * row] = vbar, p.coef_out[2 * row + 1] = beta;
    const double dtnu = dt * nu;

    // ---- 3. edge coefficients U_l, L_l of edge (i0 + l) for l = -1 .. E-1 ------------------------------------------
    // FAST: U = pD + q (v_e - vbar), L = pD - q (v_e - vbar), v_e = vc + (l + 1/2) dv
    const double pD = dtnu * D / (dv * dv);
    const double q = dtnu * (2.0 * beta * D) / (2.0 * dv);
    const double w0 = q * (vc + 0.5 * dv - vbar), dq = q * dv;
    // Chang-Cooper cell Peclet number w = C dv / D = 2 beta dv (v_edge - vbar), linear in the edge index
    const double ww0 = (2.0 * beta * dv) * (vc + 0.5 * dv - vbar), dww = (2.0 * beta * dv) * dv;
    // per-row quotients of the generic path (one IEEE division each per row instead of several per cell edge)
    const double sD = fmax(D, 1.0e-30), inv_dv = 1.0 / dv;
    const double D_dv = ((p.scheme == FP_CENTRAL) ? D : sD) / dv, dv_D = dv / sD;
    auto edge = [&](int l, double& U, double& L) {
      const int e = i0 + l;  // global edge index, valid for 0 <= e <= nv-2
      if (e < 0 || e > nv - 2) {
        U = 0.0;
        L = 0.0;
        return;
      }
      if (FAST) {
        const double wq = fma((double)l, dq, w0);  // dt nu C / (2 dv)
        if (!CC) {
          U = pD + wq;
          L = pD - wq;
        } else {  // Chang-Cooper: U = dt nu (C (1 - delta) + D/dv) / dv, L = dt nu (-C delta + D/dv) / dv
          const double dl = cc_delta_fast(fma((double)l, dww, ww0));
          const double cq = 2.0 * wq;
          U = fma(cq, 1.0 - dl, pD);
          L = fma(-cq, dl, pD);
        }
      } else {
        double C;
        const double va = __ldg(p.v + e), vb = __ldg(p.v + e + 1);
        if (p.nodrag) {
          C = 0.0;
        } else if (p.model == FP_SUPERGAUSSIAN) {
          const double ph0 = beta * pow(fabs(va - vbar), p.sg_m), ph1 = beta * pow(fabs(vb - vbar), p.sg_m);
          C = D * (ph1 - ph0) * inv_dv;
        } else {
          C = (2.0 * beta * D) * (0.5 * (vb + va) - vbar);
        }
        double bu, bl;
        bare_edge(C, D_dv, inv_dv, dv_D, p.scheme, bu, bl);
        U = dtnu * bu;
        L = dtnu * bl;
      }
    };

    // ---- 4. downward elimination, rows normalised to unit diagonal ---------------------------------------------------
    //   apn_l s_left + x_l + cpn_l x_{l+1} = rpn_l ;  y-form for f + x:  ypn_l = f_l + rpn_l + cpn_l f_{l+1}
    double cpn[E], ypn[E], apn[E];  // apn: spike of the downward sweep (coupling to the left neighbour chunk)
    double rpn_last, apn_last;
    {
      double Um, Lm;  // edge l-1
      edge(-1, Um, Lm);
      double f_m = (tt > 0) ? rowbuf[i0 - 1 + (tt - 1)] : 0.0;
      double f_c = chunk[0];
      double G_m = Um * f_c - Lm * f_m;  // flux through edge l-1
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
          bp = b;
          apv = a;
          rpv = rhs;
        } else {
          bp = fma(-a, cp_prev, b);
          apv = -a * ap_prev;
          rpv = fma(-a, rp_prev, rhs);
        }
        const double inv = fast_rcp(bp);
        const double cp = -Uc * inv, ap = apv * inv, rp = rpv * inv;
        cpn[l] = cp;
        apn[l] = ap;
        ypn[l] = (l < E - 1) ? fma(cp, f_p, f_c + rp) : f_c;  // the last row is solved by the reduced system
        if (l == E - 1) {
          rpn_last = rp;
          apn_last = ap;
        }
        cp_prev = cp, ap_prev = ap, rp_prev = rp;
        Um = Uc, Lm = Lc, G_m = G_c, f_m = f_c, f_c = f_p;
      }
    }

    // ---- 5. upward sweep: x_first = R0 - A0 s_left - C0 s_me --------------------------------------------------------
    double A0 = 0.0, C0 = 0.0, R0 = 0.0;
    if (E >= 2) {
      double RY = ypn[E - 2], A = apn[E - 2], Cc = cpn[E - 2];
#pragma unroll
      for (int l = E - 3; l >= 0; l--) {
        const double cp = cpn[l];
        RY = fma(-cp, RY, ypn[l]);
        A = fma(-cp, A, apn[l]);
        Cc = -cp * Cc;
      }
      A0 = A, C0 = Cc;
      R0 = RY - chunk[0] - Cc * chunk[E - 1];  // back from y-form: y = f + x
    }
    double* xb = pcr + 3 * RT;  // second PCR buffer doubles as the exchange area
    if (live) {
      xb[me] = A0;
      xb[RT + me] = C0;
      xb[2 * RT + me] = R0;
    }
    __syncthreads();
    // reduced row of this chunk's last unknown: al s_{t-1} + s_t + ga s_{t+1} = rh
    double al, ga, rh;
    {
      double be = 1.0;
      al = apn_last, ga = 0.0, rh = rpn_last;
      if (E >= 2 && tt < T - 1) {
        const double k = cpn[E - 1];
        be = fma(-k, xb[me + 1], 1.0);
        ga = -k * xb[RT + me + 1];
        rh = fma(-k, xb[2 * RT + me + 1], rpn_last);
      } else if (E == 1 && tt < T - 1) {
        ga = cpn[0];
      }
      const double ib = fast_rcp(be);
      al *= ib, ga *= ib, rh *= ib;
    }
    // ---- 6. parallel cyclic reduction (unit diagonal) on the T chunk-last unknowns ----------------------------------
    double* cur = pcr;
    double* nxt = pcr + 3 * RT;
    __syncthreads();  // exchange area (= nxt) fully consumed
    if (live) {
      cur[me] = al;
      cur[RT + me] = ga;
      cur[2 * RT + me] = rh;
    }
    // early termination: see fp_row_fast (collide_core.cuh); the vote covers every row of the CTA
    int more = __syncthreads_or(live && (fabs(al) > PCR_TOL || fabs(ga) > PCR_TOL));
    for (int s = 1; s < T && more; s <<= 1) {
      double alj = 0.0, gaj = 0.0, rhj = 0.0, alk = 0.0, gak = 0.0, rhk = 0.0;
      if (tt - s >= 0) {
        alj = cur[me - s];
        gaj = cur[RT + me - s];
        rhj = cur[2 * RT + me - s];
      }
      if (tt + s < T) {
        alk = cur[me + s];
        gak = cur[RT + me + s];
        rhk = cur[2 * RT + me + s];
      }
      const double be = fma(-al, gaj, fma(-ga, alk, 1.0));
      const double ib = fast_rcp(be);
      rh = fma(-al, rhj, fma(-ga, rhk, rh)) * ib;
      al = -al * alj * ib;
      ga = -ga * gak * ib;
      if (live) {
        nxt[me] = al;
        nxt[RT + me] = ga;
        nxt[2 * RT + me] = rh;
      }
      more = __syncthreads_or(live && (fabs(al) > PCR_TOL || fabs(ga) > PCR_TOL));
      double* tmp = cur;
      cur = nxt;
      nxt = tmp;
    }
    const double s_me = rh;  // = x at the chunk's last cell
    const double s_left = (tt > 0) ? cur[2 * RT + me - 1] : 0.0;

    // ---- 7. back substitution straight into f + delta, written over the row buffer -----------------------------------
    {
      double y = ypn[E - 1] + s_me;
      double* outc = rowbuf + i0 + tt;
      if (live) outc[E - 1] = y;
#pragma unroll
      for (int l = E - 2; l >= 0; l--) {
        y = fma(-cpn[l], y, fma(-apn[l], s_left, ypn[l]));
        if (live) outc[l] = y;
      }
    }

    if (!FAST && p.nodrag) {
      // fokker_planck.py:414-427: subtract dt nu lap(D f_M) with the same zero-flux stencil
      double sm0[1] = {0.0};
      double fm[E];
#pragma unroll
      for (int l = 0; l < E; l++) {
        const double d = __ldg(p.v + i0 + l) - vbar;
        fm[l] = exp(-beta * (d * d));
        sm0[0] += fm[l];
      }
      if (!live) sm0[0] = 0.0;
      row_reduce<1>(sm0, red, parity, live ? r : 0, tt, T, RT, warp_mode, live);
      const double nprof = s0 * dv;
      const double sc = nprof / (sm0[0] * dv);
      const double dl = (tt > 0 ? __ldg(p.v + i0 - 1) : 0.0) - vbar, dr = (tt < T - 1 ? __ldg(p.v + i0 + E) : 0.0) - vbar;
      const double fm_left = tt > 0 ? D * (exp(-beta * (dl * dl)) * sc) : 0.0;
      const double fm_right = tt < T - 1 ? D * (exp(-beta * (dr * dr)) * sc) : 0.0;
#pragma unroll
      for (int l = 0; l < E; l++) fm[l] = D * (fm[l] * sc);
      double* outc = rowbuf + i0 + tt;
#pragma unroll
      for (int l = 0; l < E; l++) {
        const int i = i0 + l;
        const double m_ = (l == 0) ? fm_left : fm[l > 0 ? l - 1 : 0];
        const double p_ = (l == E - 1) ? fm_right : fm[l < E - 1 ? l + 1 : l];
        double lap;
        if (i == 0)
          lap = (p_ - fm[l]) / (dv * dv);
        else if (i == nv - 1)
          lap = (m_ - fm[l]) / (dv * dv);
        else
          lap = (p_ - 2.0 * fm[l] + m_) / (dv * dv);
        if (live) outc[l] = outc[l] - dt * nu * lap;
      }
    }
  }

  // ---- 8. Krook: f e^{-nu_K dt} + n f_mx (1 - e^{-nu_K dt}); density of the result ------------------------------------
  if (p.nu_K || p.n_out) {
    double* outc = rowbuf + i0 + tt;  // own chunk only: no barrier needed after step 7
    double sn[1] = {0.0};
#pragma unroll
    for (int l = 0; l < E; l++) sn[0] += outc[l];
    if (!live) sn[0] = 0.0;
    row_reduce<1>(sn, red, parity, live ? r : 0, tt, T, RT, warp_mode, live);
    if (p.nu_K) {
      const double nprof = sn[0] * dv;
      const double ex = exp(-(dt * __dmul_rn(p.trow ? p.trow[TROW_NU_K] : p.nu_K_scale, p.nu_K[row])));
      double s2[1] = {0.0};
#pragma unroll
      for (int l = 0; l < E; l++) {
        const double val = outc[l] * ex + nprof * __ldg(p.f_mx + i0 + l) * (1.0 - ex);
        if (live) outc[l] = val;
        s2[0] += val;
      }
      if (p.n_out) {
        if (!live) s2[0] = 0.0;
        row_reduce<1>(s2, red, parity, live ? r : 0, tt, T, RT, warp_mode, live);
        sn[0] = s2[0];
      }
    }
    if (p.n_out && tt == 0 && active) p.n_out[row] = sn[0] * dv;
  }

  // ---- 9. shared -> global (coalesced) --------------------------------------------------------------------------------
  __syncthreads();
  if (active) {
    double* out = p.fout + row * nv;
    if ((nv & 1) == 0 && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
      double2* o2 = reinterpret_cast<double2*>(out);
      for (int i = t; i < (nv >> 1); i += T) {
        const int j = 2 * i;
        o2[i] = make_double2(rowbuf[j + j / E], rowbuf[j + 1 + (j + 1) / E]);
      }
    } else {
      for (int i = t; i < nv; i += T) out[i] = rowbuf[i + i / E];
    }
  }
}

#ifndef ADEPT_F32_BUILD  // the adjoint exists in fp64 only
// =====================================================================================================================
// Adjoint of the Fokker-Planck step (central differencing, Lenard-Bernstein / Dougherty):  f_new = A(m(f))^{-1} f with
// m = (vbar, T) the row moments of the INPUT f.  For a cotangent g of f_new:
//   lambda = A^{-T} g                                  (same chunked solver on the transposed matrix, delta form)
//   f_bar  = lambda + dt nu [ s_D dT/df + s_u dvbar/df ],   s_k = lambda^T (dL/dm_k) f_new
//   nu_bar = dt lambda^T L f_new
// With the flux form  dt nu (L f)_i = G_i - G_{i-1}:  lambda^T (dt nu L f) = sum_e (lambda_e - lambda_{e+1}) G_e.
struct CollideBwdArgs {
  const double* fin;   // forward input
  const double* fnew;  // forward output
  const double* g;     // cotangent of the forward output
  double* fbar;        // cotangent of the forward input
  double* nubar;       // [rows] cotangent of nu_fp (nullable)
  long long rows;
  int nv;
  const double* v;
  double dv, dt;
  const double* nu_fp;
  double nu_fp_scale;
  int model;
  int scheme;
};

template <int E, int MAXT, bool CC>
__global__ void __launch_bounds__(MAXT, 1) collide_bwd_kernel(CollideBwdArgs p) {
  extern __shared__ __align__(16) double sm[];
  const int nv = p.nv;
  const int T = nv / E;
  const int nvp = nv + T;
  const int t = threadIdx.x;
  const bool live = t < T;
  const int tt = live ? t : 0;
  double* buf_f = sm;               // f_in, later f_new
  double* buf_g = sm + nvp;         // g, later lambda
  double* apbuf = sm + 2 * nvp;     // [E][T]
  double* red = apbuf + nv;         // 2 * max(T,32) * 3
  double* pcr = red + 2 * (T > 32 ? T : 32) * 3;  // 6 T
  const bool warp_mode = (T & 31) == 0;
  int parity = 0;
  const long long row = blockIdx.x;
  const double dv = p.dv, dt = p.dt;
  const int i0 = E * tt;

  if (live) {
    const double* fin = p.fin + row * nv;
    const double* gin = p.g + row * nv;
    for (int i = t; i < nv; i += T) {
      buf_f[i + i / E] = fin[i];
      buf_g[i + i / E] = gin[i];
    }
  }
  __syncthreads();
  const double* chunk = buf_f + i0 + tt;
  double* gch = buf_g + i0 + tt;
  const double vc = __ldg(p.v + i0);
  const double nu = __dmul_rn(p.nu_fp_scale, p.nu_fp[row]);

  // moments of the forward input (same arithmetic as the forward kernel)
  double mom[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int l = 0; l < E; l++) {
    const double fl = chunk[l];
    mom[0] += fl;
    mom[1] = fma(fl, (double)l, mom[1]);
    mom[2] = fma(fl, (double)(l * l), mom[2]);
  }
  {
    const double m0 = mom[0], m1 = mom[1] * dv, m2 = mom[2] * (dv * dv);
    mom[1] = fma(vc, m0, m1);
    mom[2] = fma(vc * vc, m0, fma(2.0 * vc, m1, m2));
  }
  if (!live) mom[0] = mom[1] = mom[2] = 0.0;
  row_reduce<3>(mom, red, parity, 0, tt, T, T, warp_mode, live);
  const double s0 = mom[0], s1 = mom[1], s2 = mom[2];
  const double vbar = (p.model == FP_LB) ? 0.0 : s1 / s0;
  const double Temp = (s2 - 2.0 * vbar * s1 + vbar * vbar * s0) / s0;
  const double beta = 1.0 / (2.0 * Temp);
  const double D = 1.0 / (2.0 * beta);
  const double dtnu = dt * nu;
  const double pD = dtnu * D / (dv * dv);
  const double c2 = 2.0 * beta * D;
  const double q = dtnu * c2 / (2.0 * dv);
  const double w0 = q * (vc + 0.5 * dv - vbar), dq = q * dv;
  const double ww0 = (2.0 * beta * dv) * (vc + 0.5 * dv - vbar), dww = (2.0 * beta * dv) * dv;  // Chang-Cooper w
  auto edge = [&](int l, double& U, double& L) {
    const int e = i0 + l;
    if (e < 0 || e > nv - 2) {
      U = 0.0;
      L = 0.0;
      return;
    }
    const double wq = fma((double)l, dq, w0);
    if (!CC) {
      U = pD + wq;
      L = pD - wq;
    } else {  // same coefficients as the forward kernel (collide_core.cuh)
      const double dl = cc_delta_fast(fma((double)l, dww, ww0));
      const double cq = 2.0 * wq;
      U = fma(cq, 1.0 - dl, pD);
      L = fma(-cq, dl, pD);
    }
  };

  // transposed system: a_i = -U_{i-1}, c_i = -L_i, b_i = 1 + L_i + U_{i-1};
  // rhs_i = g_i - (A^T g)_i = L_i (g_{i+1} - g_i) - U_{i-1} (g_i - g_{i-1})
  double cpn[E], ypn[E];
  double rpn_last, apn_last;
  {
    double Um, Lm;
    edge(-1, Um, Lm);
    double g_m = (tt > 0) ? buf_g[i0 - 1 + (tt - 1)] : 0.0;
    double g_c = gch[0];
    double cp_prev = 0.0, ap_prev = 0.0, rp_prev = 0.0;
#pragma unroll
    for (int l = 0; l < E; l++) {
      double Uc, Lc;
      edge(l, Uc, Lc);
      const double g_p = (l < E - 1) ? gch[l + 1] : ((tt < T - 1) ? buf_g[i0 + E + (tt + 1)] : 0.0);
      const double rhs = Lc * (g_p - g_c) - Um * (g_c - g_m);
      const double a = -Um;
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
      const double cp = -Lc * inv, ap = apv * inv, rp = rpv * inv;
      cpn[l] = cp;
      if (live) apbuf[l * T + tt] = ap;
      ypn[l] = (l < E - 1) ? fma(cp, g_p, g_c + rp) : g_c;
      if (l == E - 1) rpn_last = rp, apn_last = ap;
      cp_prev = cp, ap_prev = ap, rp_prev = rp;
      Um = Uc, Lm = Lc, g_m = g_c, g_c = g_p;
    }
  }
  double A0 = 0.0, C0 = 0.0, R0 = 0.0;
  {
    double RY = ypn[E - 2], A = apbuf[(E - 2) * T + tt], Cc = cpn[E - 2];
#pragma unroll
    for (int l = E - 3; l >= 0; l--) {
      const double cp = cpn[l];
      RY = fma(-cp, RY, ypn[l]);
      A = fma(-cp, A, apbuf[l * T + tt]);
      Cc = -cp * Cc;
    }
    A0 = A, C0 = Cc;
    R0 = RY - gch[0] - Cc * gch[E - 1];
  }
  double* xb = pcr + 3 * T;
  if (live) xb[tt] = A0, xb[T + tt] = C0, xb[2 * T + tt] = R0;
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
  if (live) cur[tt] = al, cur[T + tt] = ga, cur[2 * T + tt] = rh;
  __syncthreads();
  for (int s = 1; s < T; s <<= 1) {
    double alj = 0.0, gaj = 0.0, rhj = 0.0, alk = 0.0, gak = 0.0, rhk = 0.0;
    if (tt - s >= 0) alj = cur[tt - s], gaj = cur[T + tt - s], rhj = cur[2 * T + tt - s];
    if (tt + s < T) alk = cur[tt + s], gak = cur[T + tt + s], rhk = cur[2 * T + tt + s];
    const double be = fma(-al, gaj, fma(-ga, alk, 1.0));
    const double ib = fast_rcp(be);
    rh = fma(-al, rhj, fma(-ga, rhk, rh)) * ib;
    al = -al * alj * ib;
    ga = -ga * gak * ib;
    if (live) nxt[tt] = al, nxt[T + tt] = ga, nxt[2 * T + tt] = rh;
    __syncthreads();
    double* tmp = cur;
    cur = nxt;
    nxt = tmp;
  }
  const double s_me = rh;
  const double s_left = (tt > 0) ? cur[2 * T + tt - 1] : 0.0;
  __syncthreads();  // every thread has read its neighbours' g values and the reduced solution
  // lambda over g; f_new over f_in
  {
    double y = ypn[E - 1] + s_me;
    if (live) gch[E - 1] = y;
#pragma unroll
    for (int l = E - 2; l >= 0; l--) {
      y = fma(-cpn[l], y, fma(-apbuf[l * T + tt], s_left, ypn[l]));
      if (live) gch[l] = y;
    }
  }
  if (live) {
    const double* fn = p.fnew + row * nv;
    for (int i = t; i < nv; i += T) buf_f[i + i / E] = fn[i];  // rows of buf_f are only read by their owners below
  }
  __syncthreads();
  // bilinear forms over the edges e = i0 + l, l = 0..E-1 (edge nv-1 does not exist)
  double bl[3] = {0.0, 0.0, 0.0};
  {
    const double* fch = buf_f + i0 + tt;
    const double lam_next_chunk = (tt < T - 1) ? buf_g[i0 + E + (tt + 1)] : 0.0;
    const double f_next_chunk = (tt < T - 1) ? buf_f[i0 + E + (tt + 1)] : 0.0;
#pragma unroll
    for (int l = 0; l < E; l++) {
      if (i0 + l > nv - 2) continue;
      const double lp = (l < E - 1) ? gch[l + 1] : lam_next_chunk;
      const double fp = (l < E - 1) ? fch[l + 1] : f_next_chunk;
      const double dl = gch[l] - lp;
      const double ve = vc + ((double)l + 0.5) * dv - vbar;
      if (!CC) {
        bl[0] += dl * (fp - fch[l]);
        bl[1] += dl * (fp + fch[l]);
        bl[2] += dl * ve * (fp + fch[l]);
      } else {
        // edge flux (dt nu / dv) [D/dv (f+ - f) + C ((1 - delta) f+ + delta f)], C = c2 (v_e - vbar), delta(w = C dv / D):
        // its derivatives w.r.t. D and vbar carry delta'(w) through dw/dD = -w/D and dw/dvbar = -c2 dv/D
        const double w = fma((double)l, dww, ww0);
        const double de = cc_delta_fast(w), dp = cc_delta_prime_fast(w);
        const double drag = fma(1.0 - de, fp, de * fch[l]);
        bl[0] += dl * (fp - fch[l]) * fma(dp, w * w, 1.0);
        bl[1] += dl * (dp * w * (fp - fch[l]) - drag);
        bl[2] += dl * ((D / dv) * (fp - fch[l]) + c2 * ve * drag);
      }
    }
  }
  if (!live) bl[0] = bl[1] = bl[2] = 0.0;
  row_reduce<3>(bl, red, parity, 0, tt, T, T, warp_mode, live);
  const double cD = (dtnu / (dv * dv)) * bl[0];                  // dt nu s_D
  const double cu = (p.model == FP_LB) ? 0.0 : (CC ? (dtnu * c2 / dv) * bl[1] : -q * bl[1]);  // dt nu s_u
  if (p.nubar && t == 0)
    p.nubar[row] = CC ? p.nu_fp_scale * (dt / dv) * bl[2]
                      : p.nu_fp_scale * dt * ((D / (dv * dv)) * bl[0] + (c2 / (2.0 * dv)) * bl[2]);
  if (live) {
    double* out = p.fbar + row * nv;
    const double is0 = 1.0 / s0;
    for (int i = t; i < nv; i += T) {
      const double vv = (__ldg(p.v + i) - vbar);
      out[i] = buf_g[i + i / E] + (cD * (vv * vv - Temp) + cu * vv) * is0;
    }
  }
}

template <int E, int MAXT, bool CC>
static int launch_collide_bwd_t(const CollideBwdArgs& p, cudaStream_t stream) {
  const int T = p.nv / E;
  const int threads = ((T + 31) / 32) * 32;
  const size_t smem = ((size_t)2 * (p.nv + T) + p.nv + 2 * (T > 32 ? T : 32) * 3 + 6 * (size_t)T) * sizeof(double);
  if (threads > MAXT || smem > 227 * 1024) {
    set_last_error("collide_bwd: nv=%d does not fit (threads=%d, smem=%zu)", p.nv, threads, smem);
    return ADEPT_ERR_UNSUPPORTED;
  }
  static size_t configured[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = collide_bwd_kernel<E, MAXT, CC>;
  if (dev < 64 && configured[dev] < smem) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(collide_bwd, smem=%zu): %s", smem, cudaGetErrorString(err));
      return ADEPT_ERR_CUDA;
    }
    configured[dev] = smem;
  }
  ProfileScope prof("collide_bwd", stream);
  kern<<<(unsigned)p.rows, threads, smem, stream>>>(p);
  return check_launch("collide_bwd_kernel");
}

template <int E, int MAXT>
static int launch_collide_bwd(const CollideBwdArgs& p, cudaStream_t stream) {
  return p.scheme == FP_CHANG_COOPER ? launch_collide_bwd_t<E, MAXT, true>(p, stream)
                                     : launch_collide_bwd_t<E, MAXT, false>(p, stream);
}

int collide_bwd_f64(const double* fin, const double* fnew, const double* g, double* fbar, double* nubar, int batch,
                    int nx, int nv, const double* v, double dv, double dt, const double* nu_fp, double nu_fp_scale,
                    int model, int scheme, cudaStream_t stream) {
  if (batch < 1 || nx < 1 || nv < 4) {
    set_last_error("collide_bwd: bad shape batch=%d nx=%d nv=%d", batch, nx, nv);
    return ADEPT_ERR_BAD_SHAPE;
  }
  if ((scheme != FP_CENTRAL && scheme != FP_CHANG_COOPER) || (model != FP_LB && model != FP_DOUGHERTY)) {
    set_last_error("collide_bwd: only the Lenard-Bernstein / Dougherty models (central or Chang-Cooper) are implemented");
    return ADEPT_ERR_UNSUPPORTED;
  }
  CollideBwdArgs p = {fin, fnew, g, fbar, nubar, (long long)batch * nx, nv, v, dv, dt, nu_fp, nu_fp_scale, model, scheme};
  if (nv % 16 == 0 && nv / 16 <= 256) return launch_collide_bwd<16, 256>(p, stream);
  if (nv % 16 == 0 && nv / 16 <= 512) return launch_collide_bwd<16, 512>(p, stream);
  if (nv % 8 == 0 && nv / 8 <= 256) return launch_collide_bwd<8, 256>(p, stream);
  if (nv % 4 == 0 && nv / 4 <= 256) return launch_collide_bwd<4, 256>(p, stream);
  if (nv % 2 == 0 && nv / 2 <= 256) return launch_collide_bwd<2, 256>(p, stream);
  set_last_error("collide_bwd: unsupported nv=%d", nv);
  return ADEPT_ERR_UNSUPPORTED;
}

#endif  // ADEPT_F32_BUILD

template <int E, int MAXT, int MINB, bool FAST, bool CC = false>
static int launch_collide_t(CollideArgs p, cudaStream_t stream) {
  const int T = p.nv / E;
  int R = MAXT / T;
  if (R < 1) R = 1;
  if ((long long)R > p.rows) R = (int)p.rows;
  p.rows_per_cta = R;
  const int RT = R * T;
  const int threads = ((RT + 31) / 32) * 32;  // whole warps; spare threads only take part in barriers
  const size_t smem = CollideSmem<E>::doubles(R, p.nv, RT) * sizeof(double);
  if (threads > MAXT || smem > 227 * 1024) {
    set_last_error("collide: nv=%d does not fit (threads=%d, smem=%zu)", p.nv, threads, smem);
    return ADEPT_ERR_UNSUPPORTED;
  }
  static size_t configured[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = collide_kernel<E, MAXT, MINB, FAST, CC>;
  if (dev < 64 && configured[dev] < smem) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(collide, smem=%zu): %s", smem, cudaGetErrorString(err));
      return ADEPT_ERR_CUDA;
    }
    configured[dev] = smem;
  }
  const long long blocks = (p.rows + R - 1) / R;
  ProfileScope prof("collide", stream);
  kern<<<(unsigned)blocks, threads, smem, stream>>>(p);
  return check_launch("collide_kernel");
}

template <int E, int MAXT, int MINB>
static int launch_collide(const CollideArgs& p, bool fast, cudaStream_t stream) {
  if (fast && p.scheme == FP_CHANG_COOPER) return launch_collide_t<E, MAXT, MINB, true, true>(p, stream);
  return fast ? launch_collide_t<E, MAXT, MINB, true>(p, stream) : launch_collide_t<E, MAXT, MINB, false>(p, stream);
}

int collide_f64(const double* fin, double* fout,
                int batch, int nx, int nv, const double* v, double dv, double dt,                   // f64
                const double* nu_fp, const double* nu_K, const double* f_mx,                        // f64
                int model, int scheme, int nodrag, double sg_m, double sg_ratio,                    // f64
                double* n_out,
                double nu_fp_scale, double nu_K_scale,                                              // f64
                cudaStream_t stream, int sc_steps, double sc_rtol, double sc_atol,                  // f64
                const double* coef_in, double* coef_out, int coef_div) {                            // f64
  if (batch < 1 || nx < 1 || nv < 4) {
    set_last_error("collide: bad shape batch=%d nx=%d nv=%d", batch, nx, nv);
    return ADEPT_ERR_BAD_SHAPE;
  }
  if (model < 0 || model > 2 || scheme < 0 || scheme > 1) {
    set_last_error("collide: unknown model=%d / scheme=%d", model, scheme);
    return ADEPT_ERR_BAD_ARG;
  }
  if (nu_K && !f_mx) {
    set_last_error("collide: Krook needs the Maxwellian table f_mx");
    return ADEPT_ERR_BAD_ARG;
  }
  CollideArgs p = {fin, fout, (long long)batch * nx, nv, v, dv, dt, nu_fp, nu_K, f_mx,
                   model, scheme, nodrag, sg_m, sg_ratio, n_out, 1, nu_fp_scale, nu_K_scale,
                   current_time_row(), sc_steps, sc_rtol, sc_atol, coef_in, coef_out, coef_div > 0 ? coef_div : 1};
  if ((coef_in || coef_out) && model == FP_SUPERGAUSSIAN) {
    set_last_error("collide: marginal coefficients (coef_in / coef_out) are defined for Lenard-Bernstein / Dougherty");
    return ADEPT_ERR_UNSUPPORTED;
  }
  if (coef_in && (coef_div < 1 || ((long long)batch * nx) % coef_div)) {
    set_last_error("collide: coef_div=%d must divide the number of rows", coef_div);
    return ADEPT_ERR_BAD_ARG;
  }
  if (sc_steps < 0 || sc_steps > 64) {
    set_last_error("collide: self-consistent beta max_steps=%d out of range [0, 64]", sc_steps);
    return ADEPT_ERR_BAD_ARG;
  }
  // uniform-grid arithmetic, central or Chang-Cooper; the Newton refinement of beta lives in the general kernel
  const bool fast = model != FP_SUPERGAUSSIAN && !nodrag && sc_steps == 0 && !coef_in && !coef_out;
#ifdef ADEPT_F32_BUILD
  if (!fast) {  // the closed-form Chang-Cooper weights and the Newton iteration of the general path need fp64
    set_last_error("collide(f32): only Lenard-Bernstein / Dougherty (central or Chang-Cooper) + Krook are offered in fp32");
    return ADEPT_ERR_UNSUPPORTED;
  }
#endif
  if (nv % 16 == 0 && nv / 16 <= 256) return launch_collide<16, 256, 2>(p, fast, stream);
  if (nv % 16 == 0 && nv / 16 <= 512) return launch_collide<16, 512, 1>(p, fast, stream);
  if (nv % 16 == 0 && nv / 16 <= 1024) return launch_collide<16, 1024, 1>(p, fast, stream);
  if (nv % 8 == 0 && nv / 8 <= 256) return launch_collide<8, 256, 2>(p, fast, stream);
  if (nv % 4 == 0 && nv / 4 <= 256) return launch_collide<4, 256, 2>(p, fast, stream);
  if (nv % 2 == 0 && nv / 2 <= 256) return launch_collide<2, 256, 2>(p, fast, stream);
  set_last_error("collide: unsupported nv=%d (need nv %% 16 == 0 and nv <= 16384, or a small even nv)", nv);
  return ADEPT_ERR_UNSUPPORTED;
}

}  // namespace adept
