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
// Parallel solve: T = nv/E threads per x-row, each owning E contiguous velocity cells in registers.  Every
// thread eliminates its chunk (modified Thomas with a left "spike"), the T chunk-last unknowns form a reduced
// tridiagonal system solved with parallel cyclic reduction in shared memory, then chunks back-substitute.
// The matrix is strictly diagonally dominant (I - dt nu L, zero-flux L), so no pivoting is needed; the reference's
// LAPACK gtsv agrees to rounding.
#include "common.cuh"

namespace adept {

enum { FP_LB = 0, FP_DOUGHERTY = 1, FP_SUPERGAUSSIAN = 2 };
enum { FP_CENTRAL = 0, FP_CHANG_COOPER = 1 };

struct CollideArgs {
  const double* fin;
  double* fout;
  long long rows;  // batch * nx
  int nv;
  const double* v;  // [nv]
  double dv, dt;
  const double* nu_fp;  // [rows] or null (Fokker-Planck off)
  const double* nu_K;   // [rows] or null (Krook off)
  const double* f_mx;   // [nv] Krook Maxwellian (unit density)
  int model, scheme, nodrag;
  double sg_m, sg_ratio;  // super-Gaussian exponent m and Gamma(3/m)/Gamma(1/m)
  double* n_out;          // [rows] or null: sum_j f_out dv
};

__device__ __forceinline__ double cc_delta(double w) {  // driftdiffusion.py:96-103
  if (fabs(w) < 1.0e-8) return 0.5 - w / 12.0 + w * w * w / 720.0;
  return 1.0 / w - 1.0 / expm1(w);
}

struct Edge {
  double bu, bl, X, Y;  // bare upper / lower entries of this edge; its contributions to bd_i and bd_{i+1}
};

__device__ __forceinline__ Edge make_edge(double C, double D, double dv, int scheme) {
  Edge g;
  if (scheme == FP_CENTRAL) {  // driftdiffusion.py:585-590
    g.X = (C / 2.0 - D / dv) / dv;
    g.Y = -(C / 2.0 + D / dv) / dv;
    g.bu = (C / 2.0 + D / dv) / dv;
    g.bl = (-C / 2.0 + D / dv) / dv;
  } else {  // driftdiffusion.py:637-648
    const double sD = fmax(D, 1.0e-30);
    const double w = C * dv / sD;
    const double dl = cc_delta(w);
    const double alpha = -C * dl + sD / dv;
    const double beta = -C * (1.0 - dl) - sD / dv;
    g.X = -alpha / dv;
    g.Y = beta / dv;
    g.bu = -beta / dv;
    g.bl = alpha / dv;
  }
  return g;
}

// Sum two values over the T threads of a row; every thread of the CTA must call it.
//   mode 0: T % 32 == 0 (warps do not straddle rows); mode 1: T < 32, power of two; mode 2: generic tree.
__device__ __forceinline__ void row_sum2(double& a, double& b, double* red, double* tree, int& parity, int r, int t,
                                         int T, int mode, unsigned amask) {
  if (mode == 1) {
    for (int o = T >> 1; o > 0; o >>= 1) {
      a += __shfl_xor_sync(amask, a, o);
      b += __shfl_xor_sync(amask, b, o);
    }
    return;
  }
  if (mode == 0) {
    a = warp_sum(a);
    b = warp_sum(b);
    const int nw = T >> 5, w = t >> 5;
    double* slot = red + parity * 64 + r * nw * 2;  // R*T <= 1024 -> at most 32 warps per CTA
    if ((t & 31) == 0) {
      slot[2 * w] = a;
      slot[2 * w + 1] = b;
    }
    __syncthreads();
    a = 0.0;
    b = 0.0;
    for (int i = 0; i < nw; i++) {
      a += slot[2 * i];
      b += slot[2 * i + 1];
    }
    parity ^= 1;
    return;
  }
  double* ta = tree + (size_t)r * T * 2;
  __syncthreads();
  ta[2 * t] = a;
  ta[2 * t + 1] = b;
  __syncthreads();
  int s = 1;
  while (s < T) s <<= 1;
  for (s >>= 1; s > 0; s >>= 1) {
    if (t < s && t + s < T) {
      ta[2 * t] += ta[2 * (t + s)];
      ta[2 * t + 1] += ta[2 * (t + s) + 1];
    }
    __syncthreads();
  }
  a = ta[0];
  b = ta[1];
}

template <int E, int MAXT>
__global__ void __launch_bounds__(MAXT) collide_kernel(CollideArgs p) {
  extern __shared__ __align__(16) double sm[];
  const int nv = p.nv;
  const int T = nv / E;
  const int R = blockDim.x / T;
  const int r = threadIdx.x / T, t = threadIdx.x % T;
  const int nvp = nv + nv / E;
  double* rowbuf = sm + (size_t)r * nvp;
  double* red = sm + (size_t)R * nvp;  // 2 parities * 32 warps * 2 values
  double* pcr = red + 128;              // 2 buffers * 4 arrays * R*T doubles
  const unsigned amask = __activemask();
  const int RT = R * T;
  const int mode = (T % 32 == 0) ? 0 : ((T < 32 && (T & (T - 1)) == 0) ? 1 : 2);
  int parity = 0;

  const long long row_raw = (long long)blockIdx.x * R + r;
  const bool active = row_raw < p.rows;
  const long long row = active ? row_raw : p.rows - 1;
  const double* fin = p.fin + row * nv;

  // ---- 1. row -> shared (coalesced), chunk -> registers ---------------------------------------------------
  for (int i = t; i < nv; i += T) rowbuf[i + i / E] = fin[i];
  __syncthreads();
  const int i0 = E * t;
  double f[E], vv[E];
#pragma unroll
  for (int l = 0; l < E; l++) {
    f[l] = rowbuf[i0 + l + t];
    vv[l] = __ldg(p.v + i0 + l);
  }
  const double f_left = t > 0 ? rowbuf[i0 - 1 + (t - 1)] : 0.0;
  const double f_right = t < T - 1 ? rowbuf[i0 + E + (t + 1)] : 0.0;
  const double v_left = t > 0 ? __ldg(p.v + i0 - 1) : 0.0;
  const double v_right = t < T - 1 ? __ldg(p.v + i0 + E) : 0.0;
  const double dv = p.dv, dt = p.dt;

  double fo[E];  // result of the Fokker-Planck stage
  if (p.nu_fp) {
    const double nu = p.nu_fp[row];
    // ---- 2. moments: vbar, T (or the super-Gaussian beta closure) -----------------------------------------
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int l = 0; l < E; l++) {
      s0 += f[l];
      s1 += f[l] * vv[l];
    }
    row_sum2(s0, s1, red, pcr, parity, r, t, T, mode, amask);
    const double vbar = (p.model == FP_LB) ? 0.0 : s1 / s0;
    double beta, D;
    if (p.model == FP_SUPERGAUSSIAN) {
      double sp = 0.0, dummy = 0.0;
#pragma unroll
      for (int l = 0; l < E; l++) sp += f[l] * pow(fabs(vv[l] - vbar), p.sg_m);
      row_sum2(sp, dummy, red, pcr, parity, r, t, T, mode, amask);
      beta = s0 / (p.sg_m * sp);
      D = pow(beta, -2.0 / p.sg_m) * p.sg_ratio;
    } else {
      double v2 = 0.0, nrm = 0.0;
#pragma unroll
      for (int l = 0; l < E; l++) {
        const double vs = vv[l] - vbar;
        v2 += f[l] * (vs * vs) * dv;
        nrm += f[l] * dv;
      }
      row_sum2(v2, nrm, red, pcr, parity, r, t, T, mode, amask);
      const double Temp = v2 / nrm;
      beta = 1.0 / (2.0 * Temp);
      D = 1.0 / (2.0 * beta);
    }

    // ---- 3. edges i0-1 .. i0+E-1 and the tridiagonal rows of this chunk ------------------------------------
    const double c2 = 2.0 * beta * D;
    const double mdtnu = -dt * nu, dtnu = dt * nu;
    double a[E], bdiag[E], c[E], rhs[E];
    {
      Edge prev;  // edge (i0 - 1), between cells i0-1 and i0
      prev.bu = prev.bl = prev.X = prev.Y = 0.0;
      if (t > 0) {
        double C;
        if (p.nodrag) {
          C = 0.0;
        } else if (p.model == FP_SUPERGAUSSIAN) {
          const double ph0 = beta * pow(fabs(v_left - vbar), p.sg_m), ph1 = beta * pow(fabs(vv[0] - vbar), p.sg_m);
          C = D * (ph1 - ph0) / dv;
        } else {
          C = c2 * (0.5 * (vv[0] + v_left) - vbar);
        }
        prev = make_edge(C, D, dv, p.scheme);
      }
#pragma unroll
      for (int l = 0; l < E; l++) {
        const int i = i0 + l;
        const bool has_lo = i >= 1, has_up = i <= nv - 2;
        Edge cur;
        cur.bu = cur.bl = cur.X = cur.Y = 0.0;
        if (has_up) {
          const double vn = (l < E - 1) ? vv[l < E - 1 ? l + 1 : l] : v_right;
          double C;
          if (p.nodrag) {
            C = 0.0;
          } else if (p.model == FP_SUPERGAUSSIAN) {
            const double ph0 = beta * pow(fabs(vv[l] - vbar), p.sg_m), ph1 = beta * pow(fabs(vn - vbar), p.sg_m);
            C = D * (ph1 - ph0) / dv;
          } else {
            C = c2 * (0.5 * (vn + vv[l]) - vbar);
          }
          cur = make_edge(C, D, dv, p.scheme);
        }
        a[l] = has_lo ? mdtnu * prev.bl : 0.0;
        c[l] = has_up ? mdtnu * cur.bu : 0.0;
        const double bd = (has_up ? cur.X : 0.0) + (has_lo ? prev.Y : 0.0);
        bdiag[l] = 1.0 - dtnu * bd;
        const double fl = (l == 0) ? f_left : f[l > 0 ? l - 1 : 0];
        const double fr = (l == E - 1) ? f_right : f[l < E - 1 ? l + 1 : l];
        rhs[l] = f[l] - ((bdiag[l] * f[l] + c[l] * fr) + a[l] * fl);  // delta form: fokker_planck.py:374
        prev = cur;
      }
    }

    // ---- 4. chunk elimination (spike toward the previous chunk's last unknown) -------------------------------
    double inv[E], ap[E], rp[E];
    inv[0] = 1.0 / bdiag[0];
    ap[0] = a[0];
    rp[0] = rhs[0];
    double bp_last = bdiag[0];
#pragma unroll
    for (int l = 1; l < E; l++) {
      const double m = a[l] * inv[l - 1];
      bp_last = bdiag[l] - m * c[l - 1];
      ap[l] = -m * ap[l - 1];
      rp[l] = rhs[l] - m * rp[l - 1];
      inv[l] = 1.0 / bp_last;
    }
    double A0 = ap[E - 2], C0 = c[E - 2], R0 = rp[E - 2];
#pragma unroll
    for (int l = E - 3; l >= 0; l--) {
      const double m = c[l] * inv[l + 1];
      A0 = ap[l] - m * A0;
      C0 = -m * C0;
      R0 = rp[l] - m * R0;
    }
    // publish (A0, C0, R0, inv0) for the chunk on the left
    double* xb = pcr + 4 * RT;  // buffer 1 doubles as the exchange area
    const int me = r * T + t;
    __syncthreads();  // (mode 2 reductions used pcr as tree scratch)
    xb[me] = A0;
    xb[RT + me] = C0;
    xb[2 * RT + me] = R0;
    xb[3 * RT + me] = inv[0];
    __syncthreads();
    double al = ap[E - 1], be = bp_last, ga = 0.0, rh = rp[E - 1];
    if (t < T - 1) {
      const double k = c[E - 1] * xb[3 * RT + me + 1];
      be = bp_last - k * xb[me + 1];
      ga = -k * xb[RT + me + 1];
      rh = rp[E - 1] - k * xb[2 * RT + me + 1];
    }
    // ---- 5. parallel cyclic reduction on the T chunk-last unknowns ------------------------------------------
    double* cur = pcr;
    double* nxt = pcr + 4 * RT;
    cur[me] = al;
    cur[RT + me] = be;
    cur[2 * RT + me] = ga;
    cur[3 * RT + me] = rh;
    __syncthreads();
    for (int s = 1; s < T; s <<= 1) {
      double al2 = 0.0, ga2 = 0.0, be2 = be, rh2 = rh;
      if (t - s >= 0) {
        const int j = me - s;
        const double k1 = al / cur[RT + j];
        al2 = -k1 * cur[j];
        be2 -= k1 * cur[2 * RT + j];
        rh2 -= k1 * cur[3 * RT + j];
      }
      if (t + s < T) {
        const int j = me + s;
        const double k2 = ga / cur[RT + j];
        ga2 = -k2 * cur[2 * RT + j];
        be2 -= k2 * cur[j];
        rh2 -= k2 * cur[3 * RT + j];
      }
      al = al2, be = be2, ga = ga2, rh = rh2;
      nxt[me] = al;
      nxt[RT + me] = be;
      nxt[2 * RT + me] = ga;
      nxt[3 * RT + me] = rh;
      __syncthreads();
      double* tmp = cur;
      cur = nxt;
      nxt = tmp;
    }
    const double s_me = rh / be;
    nxt[me] = s_me;
    __syncthreads();
    const double s_left = t > 0 ? nxt[me - 1] : 0.0;

    // ---- 6. back substitution; f + delta ----------------------------------------------------------------------
    double xi = s_me;
    fo[E - 1] = f[E - 1] + xi;
#pragma unroll
    for (int l = E - 2; l >= 0; l--) {
      xi = (rp[l] - ap[l] * s_left - c[l] * xi) * inv[l];
      fo[l] = f[l] + xi;
    }

    if (p.nodrag) {
      // fokker_planck.py:414-427: subtract dt nu lap(D f_M) with the same zero-flux stencil
      double sm0 = 0.0, dummy = 0.0;
      double fm[E];
#pragma unroll
      for (int l = 0; l < E; l++) {
        const double d = vv[l] - vbar;
        fm[l] = exp(-beta * (d * d));
        sm0 += fm[l];
      }
      row_sum2(sm0, dummy, red, pcr, parity, r, t, T, mode, amask);
      const double nprof = s0 * dv;
      const double sc = nprof / (sm0 * dv);
      const double dl = v_left - vbar, dr = v_right - vbar;
      const double fm_left = t > 0 ? D * (exp(-beta * (dl * dl)) * sc) : 0.0;
      const double fm_right = t < T - 1 ? D * (exp(-beta * (dr * dr)) * sc) : 0.0;
#pragma unroll
      for (int l = 0; l < E; l++) fm[l] = D * (fm[l] * sc);
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
        fo[l] = fo[l] - dt * nu * lap;
      }
    }
  } else {
#pragma unroll
    for (int l = 0; l < E; l++) fo[l] = f[l];
  }

  // ---- 7. Krook: f e^{-nu_K dt} + n f_mx (1 - e^{-nu_K dt}) ---------------------------------------------------
  if (p.nu_K) {
    double sn = 0.0, dummy = 0.0;
#pragma unroll
    for (int l = 0; l < E; l++) sn += fo[l];
    row_sum2(sn, dummy, red, pcr, parity, r, t, T, mode, amask);
    const double nprof = sn * dv;
    const double ex = exp(-(dt * p.nu_K[row]));
#pragma unroll
    for (int l = 0; l < E; l++) fo[l] = fo[l] * ex + nprof * __ldg(p.f_mx + i0 + l) * (1.0 - ex);
  }

  // ---- 8. density of the result -------------------------------------------------------------------------------
  if (p.n_out) {
    double sn = 0.0, dummy = 0.0;
#pragma unroll
    for (int l = 0; l < E; l++) sn += fo[l];
    row_sum2(sn, dummy, red, pcr, parity, r, t, T, mode, amask);
    if (t == 0 && active) p.n_out[row] = sn * dv;
  }

  // ---- 9. registers -> shared -> global (coalesced) -----------------------------------------------------------
  __syncthreads();
#pragma unroll
  for (int l = 0; l < E; l++) rowbuf[i0 + l + t] = fo[l];
  __syncthreads();
  if (active) {
    double* out = p.fout + row * nv;
    for (int i = t; i < nv; i += T) out[i] = rowbuf[i + i / E];
  }
}

template <int E, int MAXT>
static int launch_collide(const CollideArgs& p, cudaStream_t stream) {
  const int T = p.nv / E;
  int R = MAXT / T;
  if (R < 1) R = 1;
  if ((long long)R > p.rows) R = (int)p.rows;
  const int threads = R * T;
  const int nvp = p.nv + p.nv / E;
  const size_t smem = ((size_t)R * nvp + 128 + 8 * (size_t)R * T) * sizeof(double);
  static size_t configured[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = collide_kernel<E, MAXT>;
  if (dev < 64 && configured[dev] < smem) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(collide, smem=%zu): %s", smem, cudaGetErrorString(err));
      return ADEPT_ERR_CUDA;
    }
    configured[dev] = smem;
  }
  const long long blocks = (p.rows + R - 1) / R;
  kern<<<(unsigned)blocks, threads, smem, stream>>>(p);
  return check_launch("collide_kernel");
}

int collide_f64(const double* fin, double* fout, int batch, int nx, int nv, const double* v, double dv, double dt,
                const double* nu_fp, const double* nu_K, const double* f_mx, int model, int scheme, int nodrag,
                double sg_m, double sg_ratio, double* n_out, cudaStream_t stream) {
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
                   model, scheme, nodrag, sg_m, sg_ratio, n_out};
  if (nv % 8 == 0 && nv / 8 <= 256) return launch_collide<8, 256>(p, stream);
  if (nv % 8 == 0 && nv / 8 <= 512) return launch_collide<8, 512>(p, stream);
  if (nv % 16 == 0 && nv / 16 <= 512) return launch_collide<16, 512>(p, stream);
  if (nv % 16 == 0 && nv / 16 <= 1024) return launch_collide<16, 1024>(p, stream);
  if (nv % 4 == 0 && nv / 4 <= 256) return launch_collide<4, 256>(p, stream);
  if (nv % 2 == 0 && nv / 2 <= 256) return launch_collide<2, 256>(p, stream);
  set_last_error("collide: unsupported nv=%d (need nv %% 8 == 0 and nv <= 16384, or a small even nv)", nv);
  return ADEPT_ERR_UNSUPPORTED;
}

}  // namespace adept
