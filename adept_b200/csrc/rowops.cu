// Row-wise (v-contiguous) kernels that are not FFTs: cubic-spline v-advection, velocity moments, Krook-free
// helpers.  Reference semantics (file:line relative to /root/reference):
//   cubic v-advection   adept/_vlasov1d/solvers/pushers/vlasov.py:106-148 (_uniform_cubic_interp), :162-172
//   charge density      adept/_vlasov1d/solvers/pushers/field.py:186-208
//   current density     adept/_vlasov1d/solvers/pushers/field.py:319-340
#include "common.cuh"

namespace adept {

// ---- cubic-spline (semi-Lagrangian) velocity push ------------------------------------------------------------
// one CTA per x-row; out-of-place (taps of neighbouring threads are read through L1).
__global__ void __launch_bounds__(256) spline_push_kernel(const double* __restrict__ fin, double* __restrict__ fout,
                                                          int nv, const double* __restrict__ e,
                                                          const double* __restrict__ dex,
                                                          const double* __restrict__ pond, double q, double m,
                                                          double dt, double dv) {
  const long long row = blockIdx.x;
  const double* fr = fin + row * nv;
  double* out = fout + row * nv;
  double ee = e[row];
  if (dex) ee = __dadd_rn(ee, dex[row]);
  const double pd = pond ? pond[row] : 0.0;
  const double shift = __dmul_rn(accel_of(ee, pd, q, q * q / m, m), dt);
  const double scaled = __ddiv_rn(shift, dv);
  const int row_offset = (int)floor(-scaled);
  for (int j = threadIdx.x; j < nv; j += blockDim.x) {
    int left = j + row_offset;
    left = left < 0 ? 0 : (left > nv - 2 ? nv - 2 : left);
    const double query = __dsub_rn((double)j, scaled);
    double t = __dsub_rn(query, (double)left);
    t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
    const int im1 = left - 1 < 0 ? 0 : left - 1;
    const int ip2 = left + 2 > nv - 1 ? nv - 1 : left + 2;
    const double fm1 = __ldg(fr + im1), f0 = __ldg(fr + left), f1 = __ldg(fr + left + 1), f2 = __ldg(fr + ip2);
    const double m0 = (left == 0) ? (f1 - f0) : 0.5 * (f1 - fm1);
    const double m1 = (left == nv - 2) ? (f1 - f0) : 0.5 * (f2 - f0);
    const double t2 = t * t, t3 = t2 * t;
    double val = (2.0 * t3 - 3.0 * t2 + 1.0) * f0 + (t3 - 2.0 * t2 + t) * m0 + (-2.0 * t3 + 3.0 * t2) * f1 +
                 (t3 - t2) * m1;
    if (query < 0.0 || query > (double)(nv - 1)) val = 1.0e-30;
    out[j] = val;
  }
}

int edfdv_spline_f64(const double* fin, double* fout, int batch, int nx, int nv, const double* e, const double* dex,
                     const double* pond, double q, double m, double dt, double dv, cudaStream_t stream) {
  if (batch < 1 || nx < 1 || nv < 2) {
    set_last_error("edfdv_spline: bad shape batch=%d nx=%d nv=%d (cubic interpolation needs nv >= 2)", batch, nx, nv);
    return ADEPT_ERR_BAD_SHAPE;
  }
  if (fin == fout) {
    set_last_error("edfdv_spline: in-place operation is not supported (f_in == f_out)");
    return ADEPT_ERR_BAD_ARG;
  }
  ProfileScope prof("edfdv_spline", stream);
  spline_push_kernel<<<(unsigned)((long long)batch * nx), 256, 0, stream>>>(fin, fout, nv, e, dex, pond, q, m, dt, dv);
  return check_launch("spline_push_kernel");
}

// ---- velocity moments ----------------------------------------------------------------------------------------
// one warp per row: s_k = sum_j f[row, j] * v[j]^k, k = 0..2; out_k[row] = base_k[row] + scale_b_k * (s_k * scale_a)
struct MomentArgs {
  const double* f;
  const double* v;
  long long rows;
  int nv;
  double scale_a;  // dv
  const double* base[3];
  double* out[3];
  double scale_b[3];
};

__global__ void __launch_bounds__(256) moments_kernel(MomentArgs p) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const double* fr = p.f + row * p.nv;
  const bool need1 = p.out[1] != nullptr, need2 = p.out[2] != nullptr;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  if ((p.nv & 1) == 0 && ((reinterpret_cast<uintptr_t>(fr) & 15) == 0)) {
    const double2* f2 = reinterpret_cast<const double2*>(fr);
    const int n2 = p.nv >> 1;
    for (int j = lane; j < n2; j += 32) {
      const double2 x = f2[j];
      s0 += x.x + x.y;
      if (need1 || need2) {
        const double v0 = __ldg(p.v + 2 * j), v1 = __ldg(p.v + 2 * j + 1);
        s1 += x.x * v0 + x.y * v1;
        if (need2) s2 += x.x * v0 * v0 + x.y * v1 * v1;
      }
    }
  } else {
    for (int j = lane; j < p.nv; j += 32) {
      const double x = fr[j];
      s0 += x;
      if (need1 || need2) {
        const double vv = __ldg(p.v + j);
        s1 += x * vv;
        if (need2) s2 += x * vv * vv;
      }
    }
  }
  s0 = warp_sum(s0);
  if (need1) s1 = warp_sum(s1);
  if (need2) s2 = warp_sum(s2);
  if (lane == 0) {
    const double s[3] = {s0, s1, s2};
#pragma unroll
    for (int k = 0; k < 3; k++)
      if (p.out[k]) {
        const double term = __dmul_rn(p.scale_b[k], __dmul_rn(s[k], p.scale_a));
        p.out[k][row] = p.base[k] ? __dadd_rn(p.base[k][row], term) : term;
      }
  }
}

int moments_f64(const double* f, int batch, int nx, int nv, const double* v, double scale_a, const double* const* base,
                double* const* out, const double* scale_b, cudaStream_t stream) {
  if (batch < 1 || nx < 1 || nv < 1) {
    set_last_error("moments: bad shape batch=%d nx=%d nv=%d", batch, nx, nv);
    return ADEPT_ERR_BAD_SHAPE;
  }
  MomentArgs p = {};
  p.f = f, p.v = v, p.rows = (long long)batch * nx, p.nv = nv, p.scale_a = scale_a;
  for (int k = 0; k < 3; k++) {
    p.base[k] = base ? base[k] : nullptr;
    p.out[k] = out[k];
    p.scale_b[k] = scale_b ? scale_b[k] : 1.0;
  }
  if ((p.out[1] || p.out[2]) && !v) {
    set_last_error("moments: v grid required for first/second moments");
    return ADEPT_ERR_BAD_ARG;
  }
  const int wpb = 8;
  const long long blocks = (p.rows + wpb - 1) / wpb;
  ProfileScope prof("moments", stream);
  moments_kernel<<<(unsigned)blocks, wpb * 32, 0, stream>>>(p);
  return check_launch("moments_kernel");
}

// ---- tiny elementwise: out = a + s * b  (Ampere: E = E_prev - dt * j, field.py:354) ----------------------------
__global__ void axpy_kernel(const double* __restrict__ a, const double* __restrict__ b, double s,
                            double* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __dadd_rn(a[i], __dmul_rn(s, b[i]));
}

int axpy_f64(const double* a, const double* b, double s, double* out, long long n, cudaStream_t stream) {
  if (n < 1) return ADEPT_OK;
  ProfileScope prof("axpy", stream);
  axpy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(a, b, s, out, n);
  return check_launch("axpy_kernel");
}

// ---- dfdt diagnostics: out = (a - b) / dt with the reference's rounding (vector_field.py:245-250); out may alias b ----
__global__ void diff_over_dt_kernel(const double* __restrict__ a, const double* b, double dt, double* out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long j = i; j < n; j += stride) out[j] = __ddiv_rn(__dsub_rn(a[j], b[j]), dt);
}

int diff_over_dt_f64(const double* a, const double* b, double dt, double* out, long long n, cudaStream_t stream) {
  if (n < 1) return ADEPT_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  ProfileScope prof("diff_over_dt", stream);
  diff_over_dt_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a, b, dt, out, n);
  return check_launch("diff_over_dt_kernel");
}

// ---- field-energy scalars of the default save: out[b] = {mean(e_b^2), mean(de_b^2)}  (storage.py:316-317) ----------
// (e0, de0) alone, or the state interpolated linearly towards (e1, de1) with weight w (diffrax's dense output)
__global__ void __launch_bounds__(1024) field_energy_kernel(const double* __restrict__ e0, const double* __restrict__ de0,
                                                           const double* __restrict__ e1, const double* __restrict__ de1,
                                                           double w, int nx, double* __restrict__ out) {
  __shared__ double red[2][32];
  const long long off = (long long)blockIdx.x * nx;
  double s0 = 0.0, s1 = 0.0;
  // a latency-bound kernel: four grid points per thread and pass, every load issued before the first use
  for (int i0 = threadIdx.x; i0 < nx; i0 += 4 * 1024) {
    double a[4], b[4], a1[4], b1[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = i0 + u * 1024;
      const bool in = i < nx;
      a[u] = in ? e0[off + i] : 0.0, b[u] = in ? de0[off + i] : 0.0;
      a1[u] = (in && e1) ? e1[off + i] : a[u], b1[u] = (in && e1) ? de1[off + i] : b[u];
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      double x = a[u], y = b[u];
      if (e1) {
        x = __dadd_rn(x, __dmul_rn(w, __dsub_rn(a1[u], x)));
        y = __dadd_rn(y, __dmul_rn(w, __dsub_rn(b1[u], y)));
      }
      s0 = fma(x, x, s0);
      s1 = fma(y, y, s1);
    }
  }
  s0 = warp_sum(s0), s1 = warp_sum(s1);
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = s0, red[1][threadIdx.x >> 5] = s1;
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int k = 0; k < 32; k++) t += red[threadIdx.x][k];
    out[2 * blockIdx.x + threadIdx.x] = t / (double)nx;
  }
}

int field_energy_f64(const double* e0, const double* de0, const double* e1, const double* de1, double w, int batch,
                     int nx, double* out, cudaStream_t stream) {
  if (batch < 1 || nx < 1) {
    set_last_error("field_energy: bad shape batch=%d nx=%d", batch, nx);
    return ADEPT_ERR_BAD_SHAPE;
  }
  if ((e1 == nullptr) != (de1 == nullptr)) {
    set_last_error("field_energy: give both e1 and de1 or neither");
    return ADEPT_ERR_BAD_ARG;
  }
  ProfileScope prof("field_energy", stream);
  field_energy_kernel<<<batch, 1024, 0, stream>>>(e0, de0, e1, de1, w, nx, out);
  return check_launch("field_energy_kernel");
}

// ---- space averages of saved moments: out[r] = mean_i a[r, i]  (the jnp.mean over x of get_default_save_func,
// storage.py:306-323).  One CTA per row, fixed summation order; `out` may be pinned host memory (mapped): the scalars
// of a save point then reach the host without a reduction kernel of the host framework and a separate copy.
__global__ void __launch_bounds__(1024) row_means_kernel(const double* __restrict__ a, long long n,
                                                        double* __restrict__ out) {
  __shared__ double red[32];
  const double* row = a + (long long)blockIdx.x * n;
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  for (long long i0 = threadIdx.x; i0 < n; i0 += 4 * 1024) {
    double x[4];
#pragma unroll
    for (int u = 0; u < 4; u++) x[u] = (i0 + u * 1024 < n) ? row[i0 + u * 1024] : 0.0;
#pragma unroll
    for (int u = 0; u < 4; u++) s[u] += x[u];
  }
  const double t = warp_sum((s[0] + s[1]) + (s[2] + s[3]));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int k = 0; k < 32; k++) tot += red[k];
    out[blockIdx.x] = tot / (double)n;
  }
}

int row_means_f64(const double* a, int rows, long long n, double* out, cudaStream_t stream) {
  if (rows < 1 || n < 1) {
    set_last_error("row_means: bad shape rows=%d n=%lld", rows, n);
    return ADEPT_ERR_BAD_SHAPE;
  }
  ProfileScope prof("row_means", stream);
  row_means_kernel<<<rows, 1024, 0, stream>>>(a, n, out);
  return check_launch("row_means_kernel");
}

// ---- second stage of the fused x-push charge density: out[i] = base[i] + scale_b * ((sum_p parts[p, i]) * scale_a) ----
// 64 rows per CTA; the parts are dealt to 4 thread groups (p = g, g+4, ...), each with two running sums, and combined
// in a fixed order: deterministic, and short dependent-load chains (nparts/8 per thread).
// 16 groups of 64 columns per CTA; a thread keeps up to 10 loads in flight (rows g, g + 16, ...: 148 rows of a persistent
// x-advection in one batch), the sums are formed in a fixed order
__global__ void __launch_bounds__(1024) reduce_parts_kernel(const double* __restrict__ parts, int nparts, long long n,
                                                            double scale_a, double scale_b,
                                                            const double* __restrict__ base, double* __restrict__ out) {
  __shared__ double sm[16][64];
  const int r = threadIdx.x & 63, g = threadIdx.x >> 6;
  const long long i = (long long)blockIdx.x * 64 + r;
  double s = 0.0;
  if (i < n) {
    for (int p0 = g; p0 < nparts; p0 += 160) {
      double x[10];
#pragma unroll
      for (int u = 0; u < 10; u++) {
        const int pidx = p0 + 16 * u;
        x[u] = pidx < nparts ? parts[(long long)pidx * n + i] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 10; u++) s += x[u];
    }
  }
  sm[g][r] = s;
  __syncthreads();
  if (g == 0 && i < n) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 16; q++) t += sm[q][r];
    const double term = __dmul_rn(scale_b, __dmul_rn(t, scale_a));
    out[i] = base ? __dadd_rn(base[i], term) : term;
  }
}

int reduce_parts_f64(const double* parts, int nparts, long long n, double scale_a, double scale_b, const double* base,
                     double* out, cudaStream_t stream) {
  if (nparts < 1 || n < 1) {
    set_last_error("reduce_parts: bad shape nparts=%d n=%lld", nparts, n);
    return ADEPT_ERR_BAD_SHAPE;
  }
  ProfileScope prof("reduce_parts", stream);
  reduce_parts_kernel<<<(unsigned)((n + 63) / 64), 1024, 0, stream>>>(parts, nparts, n, scale_a, scale_b, base, out);
  return check_launch("reduce_parts_kernel");
}

// ---- in-loop save moments (storage.py:119-162, 286-327) -----------------------------------------------------------
// one warp per row of the (optionally time-interpolated) distribution f = f0 + w (f1 - f0):
//   out[k, row] = dv * sum_j g_k(f_j, v_j),  g = { f, f v, f v^2, f v^3, -|f| log|f|, f^2 }
// SPLIT warps share a row (each takes a contiguous part of it): the logarithm makes the pass instruction-bound, and
// 4096 rows alone fill less than half of the warp slots of 148 SMs
// Natural logarithm for the entropy moment, absolute error ~2e-16 (the library log costs ~45 fp64 instructions per cell and
// made the pass fp64-bound at 80 us for 4096^2): x = 2^e m, m in [1, 2); the top 7 mantissa bits pick c = 1 + i/128 from
// a shared-memory table of (1/c rounded, -log(1/c rounded)), r = m (1/c) - 1 in [0, 2^-7] exactly (one fma), and
// log1p(r) is its Taylor polynomial through r^8 (next term < 1.2e-20).  Zero, subnormals, inf and NaN take the library
// path (log(0) = -inf must survive: -|f| log|f| is NaN at f = 0 in the reference too).
__device__ __noinline__ double slow_log(double ax) { return log(ax); }  // out of line: keeps the streaming loop small
// true for zero, subnormals, inf and NaN (ax >= 0): everything the table path does not cover
__device__ __forceinline__ bool log_needs_library(double ax) {
  return (unsigned)(__double2hiint(ax) - 0x00100000) >= (unsigned)(0x7ff00000 - 0x00100000);
}
__device__ __forceinline__ double fast_log_normal(double ax, const double2* __restrict__ tab) {  // normal range only
  const int hi = __double2hiint(ax);
  const double2 tb = tab[(hi >> 13) & 127];
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(ax));
  const double r = fma(m, tb.x, -1.0);
  double q = fma(r, -1.0 / 8.0, 1.0 / 7.0);
  q = fma(r, q, -1.0 / 6.0);
  q = fma(r, q, 1.0 / 5.0);
  q = fma(r, q, -1.0 / 4.0);
  q = fma(r, q, 1.0 / 3.0);
  q = fma(r, q, -0.5);
  const double p1 = fma(r * r, q, r);
  return fma((double)((hi >> 20) - 1023), 0.693147180559945309417232, tb.y + p1);
}
__device__ __forceinline__ double fast_log_pos(double ax, const double2* __restrict__ tab) {
  return log_needs_library(ax) ? slow_log(ax) : fast_log_normal(ax, tab);
}
// Variant with half the table traffic (LOGV = 1): the multiplier is not looked up but taken from the hardware reciprocal
// seed of m (MUFU.RCP64H, about 20 bits) truncated to 7 mantissa bits, 1/2 <= q <= 1, so only -log(q) comes from a table
// (129 eight-byte entries indexed by q's own bits: whatever the seed returns, multiplier and table entry belong together).
// r = m q - 1 lies in (-2^-7 - 2^-20, 2^-20]; same polynomial.
__device__ __forceinline__ double fast_log_normal_rcp(double ax, const double* __restrict__ ltab) {
  const int hi = __double2hiint(ax);
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(ax));
  double r0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(m));
  int rh = __double2hiint(r0) & 0xffffe000;
  rh = min(max(rh, 0x3fe00000), 0x3ff00000);  // q in [1/2, 1] whatever the seed's last bits are
  const double qm = __hiloint2double(rh, 0);
  const int j = (rh - 0x3fe00000) >> 13;      // 0 .. 128
  const double r = fma(m, qm, -1.0);
  double q = fma(r, -1.0 / 8.0, 1.0 / 7.0);
  q = fma(r, q, -1.0 / 6.0);
  q = fma(r, q, 1.0 / 5.0);
  q = fma(r, q, -1.0 / 4.0);
  q = fma(r, q, 1.0 / 3.0);
  q = fma(r, q, -0.5);
  const double p1 = fma(r * r, q, r);
  return fma((double)((hi >> 20) - 1023), 0.693147180559945309417232, ltab[j] + p1);
}
template <int SPLIT, int LOGV = 0>
__global__ void __launch_bounds__(256, 3) save_moments_kernel(const double* __restrict__ f0, const double* __restrict__ f1,
                                                           double w, const double* __restrict__ v, long long rows,
                                                           int nv, double dv, double* __restrict__ out) {
  __shared__ double part[8][6];
  __shared__ double2 logtab[128];
  __shared__ double ltab[LOGV == 1 ? 129 : 1];
  if (threadIdx.x < 128) {
    const double inv = 1.0 / (1.0 + (double)threadIdx.x * (1.0 / 128.0));
    logtab[threadIdx.x] = make_double2(inv, -log(inv));
  }
  if (LOGV == 1 && threadIdx.x < 129)  // q = (1 + j/128) / 2 for j < 128, q = 1 for j = 128
    ltab[threadIdx.x] = threadIdx.x == 128 ? 0.0 : -log(0.5 * (1.0 + (double)threadIdx.x * (1.0 / 128.0)));
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long row = (long long)blockIdx.x * (8 / SPLIT) + wid / SPLIT;
  const int q = wid % SPLIT;
  const bool live = row < rows;
  const int len = nv / SPLIT;  // the launcher picks SPLIT so that this is a multiple of 4
  const double* a = f0 + (live ? row : 0) * nv + (size_t)q * len;
  const double* b = f1 ? f1 + (live ? row : 0) * nv + (size_t)q * len : nullptr;
  const double* vq = v + (size_t)q * len;
  double s[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  auto add = [&](double x, double y, double vv, double (&acc)[6]) {
    if (b) x = x + w * (y - x);  // diffrax's linear dense output between y0 and y1 (adept/_base_.py:40)
    const double ax = fabs(x);
    acc[0] += x;
    acc[1] += x * vv;
    acc[2] += x * (vv * vv);
    acc[3] += x * (vv * vv * vv);
    acc[4] = fma(-fast_log_pos(ax, logtab), ax, acc[4]);
    acc[5] += x * x;
  };
  if (live) {
    int j0 = 0;
    if ((len & 1) == 0 && ((reinterpret_cast<uintptr_t>(a) | (b ? reinterpret_cast<uintptr_t>(b) : 0) |
                            reinterpret_cast<uintptr_t>(vq)) & 15) == 0) {
      double s1[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      const double2* a2 = reinterpret_cast<const double2*>(a);
      const double2* b2 = reinterpret_cast<const double2*>(b);
      const double2* v2 = reinterpret_cast<const double2*>(vq);
      const int n2 = len >> 1;
      int i = lane;
      // four independent 16-byte loads of f per lane in flight (one per iteration left the pass latency-bound on DRAM:
      // 16 KB in flight per SM, 70 us for 4096^2)
      for (; i + 96 < n2; i += 128) {
        double2 xa[4], ya[4], va[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          xa[u] = __ldcs(a2 + i + 32 * u);
          ya[u] = b ? __ldcs(b2 + i + 32 * u) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) va[u] = __ldg(v2 + i + 32 * u);
        // the eight cells of a batch share ONE range test for the logarithm (a branch per cell cost more instructions
        // than the arithmetic: 70 per cell)
        double xs[8], lg[8];
        bool special = false;
#pragma unroll
        for (int u = 0; u < 4; u++) {
          xs[2 * u] = b ? xa[u].x + w * (ya[u].x - xa[u].x) : xa[u].x;  // diffrax's linear dense output (adept/_base_.py:40)
          xs[2 * u + 1] = b ? xa[u].y + w * (ya[u].y - xa[u].y) : xa[u].y;
          special = special || log_needs_library(fabs(xs[2 * u])) || log_needs_library(fabs(xs[2 * u + 1]));
        }
        // (a 32-entry table served by warp shuffles instead of the shared-memory lookups was measured: 52 -> 68 us)
        if (special) {
#pragma unroll
          for (int u = 0; u < 8; u++) lg[u] = fast_log_pos(fabs(xs[u]), logtab);
        } else if (LOGV == 1) {
#pragma unroll
          for (int u = 0; u < 8; u++) lg[u] = fast_log_normal_rcp(fabs(xs[u]), ltab);
        } else {
#pragma unroll
          for (int u = 0; u < 8; u++) lg[u] = fast_log_normal(fabs(xs[u]), logtab);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const double x = xs[u], vv = (u & 1) ? va[u >> 1].y : va[u >> 1].x;
          double(&acc)[6] = (u & 1) ? s1 : s;
          acc[0] += x;
          acc[1] += x * vv;
          acc[2] += x * (vv * vv);
          acc[3] += x * (vv * vv * vv);
          acc[4] = fma(-lg[u], fabs(x), acc[4]);
          acc[5] += x * x;
        }
      }
      for (; i < n2; i += 32) {
        const double2 xa = a2[i], va = __ldg(v2 + i);
        const double2 ya = b ? b2[i] : make_double2(0.0, 0.0);
        add(xa.x, ya.x, va.x, s);
        add(xa.y, ya.y, va.y, s1);
      }
#pragma unroll
      for (int k = 0; k < 6; k++) s[k] += s1[k];
      j0 = len;
    }
    for (int j = j0 + lane; j < len; j += 32) add(a[j], b ? b[j] : 0.0, __ldg(vq + j), s);
  }
#pragma unroll
  for (int k = 0; k < 6; k++) {
    const double t = warp_sum(s[k]);
    if (lane == 0) part[wid][k] = t;
  }
  __syncthreads();
  if (live && q == 0 && lane < 6) {
    double t = 0.0;
#pragma unroll
    for (int u = 0; u < SPLIT; u++) t += part[wid + u][lane];
    out[(long long)lane * rows + row] = t * dv;
  }
}

int save_moments_f64(const double* f0, const double* f1, double w, int batch, int nx, int nv, const double* v,
                     double dv, double* out, cudaStream_t stream) {
  if (batch < 1 || nx < 1 || nv < 1) {
    set_last_error("save_moments: bad shape batch=%d nx=%d nv=%d", batch, nx, nv);
    return ADEPT_ERR_BAD_SHAPE;
  }
  const long long rows = (long long)batch * nx;
  ProfileScope prof("save_moments", stream);
  static int logv = -1;  // ADEPT_B200_LOGVAR=1: the reciprocal-seed logarithm (fast_log_normal_rcp), A/B timing
  if (logv < 0) {
    const char* e = getenv("ADEPT_B200_LOGVAR");
    logv = (e && atoi(e) == 1) ? 1 : 0;
  }
  if (nv % 16 == 0 && nv >= 2048 && logv == 1)
    save_moments_kernel<4, 1><<<(unsigned)((rows + 1) / 2), 256, 0, stream>>>(f0, f1, w, v, rows, nv, dv, out);
  else if (nv % 16 == 0 && nv >= 2048)
    save_moments_kernel<4><<<(unsigned)((rows + 1) / 2), 256, 0, stream>>>(f0, f1, w, v, rows, nv, dv, out);
  else
    save_moments_kernel<1><<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(f0, f1, w, v, rows, nv, dv, out);
  return check_launch("save_moments_kernel");
}


// ---- distribution save on a coarser (x, v) mesh (storage.py:173-181: interpax.interp2d, method="linear") -----------
// out[a, b] = bilinear interpolation of f (or of the time-interpolated f0 + w (f1 - f0)) at (xq[a], vq[b]); NaN outside
// the grid, like interpax with extrap=False.  Node search = searchsorted(side="right") on the actual axes.
__device__ __forceinline__ int upper_bound(const double* __restrict__ a, int n, double q) {  // #{k : a[k] <= q}
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] <= q) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) interp2d_kernel(const double* __restrict__ f0, const double* __restrict__ f1,
                                                       double w, int nx, int nv, const double* __restrict__ x,
                                                       const double* __restrict__ v, const double* __restrict__ xq,
                                                       const double* __restrict__ vq, int nxq, int nvq,
                                                       double* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)nxq * nvq) return;
  const int a = (int)(idx / nvq), b = (int)(idx % nvq);
  const double qx = xq[a], qv = vq[b];
  int i = upper_bound(x, nx, qx), j = upper_bound(v, nv, qv);
  i = min(max(i, 1), nx - 1), j = min(max(j, 1), nv - 1);
  auto at = [&](int r, int c) {
    const size_t o = (size_t)r * nv + c;
    double val = f0[o];
    if (f1) val = val + w * (f1[o] - val);
    return val;
  };
  const double f00 = at(i - 1, j - 1), f01 = at(i - 1, j), f10 = at(i, j - 1), f11 = at(i, j);
  const double x0 = x[i - 1], x1 = x[i], y0 = v[j - 1], y1 = v[j];
  const double dx0 = qx - x0, dx1 = x1 - qx, dy0 = qv - y0, dy1 = y1 - qv;
  double r = (dx1 * (f00 * dy1 + f01 * dy0) + dx0 * (f10 * dy1 + f11 * dy0)) / ((x1 - x0) * (y1 - y0));
  if (qx < x[0] || qx > x[nx - 1] || qv < v[0] || qv > v[nv - 1]) r = __longlong_as_double(0x7ff8000000000000ll);
  out[idx] = r;
}

int interp2d_f64(const double* f0, const double* f1, double w, int nx, int nv, const double* x, const double* v,
                 const double* xq, const double* vq, int nxq, int nvq, double* out, cudaStream_t stream) {
  if (nx < 2 || nv < 2 || nxq < 1 || nvq < 1) {
    set_last_error("interp2d: bad shape nx=%d nv=%d nxq=%d nvq=%d", nx, nv, nxq, nvq);
    return ADEPT_ERR_BAD_SHAPE;
  }
  const long long total = (long long)nxq * nvq;
  ProfileScope prof("interp2d", stream);
  interp2d_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(f0, f1, w, nx, nv, x, v, xq, vq, nxq, nvq, out);
  return check_launch("interp2d_kernel");
}

// ---- out[i] = sum_r peers[r][i] in rank order (the same order on every rank: bit-identical results everywhere) ---------
struct SumPeersArgs {
  const double* src[8];
  int n_peers;
  long long n;
  double* out;
};
__global__ void __launch_bounds__(256) sum_peers_kernel(SumPeersArgs p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  double x[8];
#pragma unroll
  for (int r = 0; r < 8; r++) x[r] = r < p.n_peers ? __ldcv(p.src[r] + i) : 0.0;  // all peer loads in flight; no caching
  double s = 0.0;
#pragma unroll
  for (int r = 0; r < 8; r++) s += x[r];
  p.out[i] = s;
}

int sum_peers_f64(const double* const* peers, int n_peers, long long n, double* out, cudaStream_t stream) {
  if (n_peers < 1 || n_peers > 8 || n < 1) {
    set_last_error("sum_peers: bad arguments (n_peers=%d, n=%lld)", n_peers, n);
    return ADEPT_ERR_BAD_ARG;
  }
  SumPeersArgs p = {};
  for (int r = 0; r < n_peers; r++) p.src[r] = peers[r];
  p.n_peers = n_peers, p.n = n, p.out = out;
  ProfileScope prof("sum_peers", stream);
  sum_peers_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p);
  return check_launch("sum_peers_kernel");
}

// ---- vlasov-1d2v helpers (adept/_vlasov1d2v/solvers/vector_field.py:36-38, pushers/fokker_planck.py:81-83) -----------
// marginal F[row] = sum_p f[row, p] w[p]  (einsum "xvp,p->xv"), rows = nx * nv, fixed summation order per row
__global__ void __launch_bounds__(256) marginal_kernel(const double* __restrict__ f, const double* __restrict__ w,
                                                       long long rows, int np, double* __restrict__ out) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const double* fr = f + row * np;
  double s = 0.0;
  for (int j = 0; j < np; j++) s += fr[j] * __ldg(w + j);
  out[row] = s;
}

int marginal_f64(const double* f, const double* w, long long rows, int np, double* out, cudaStream_t stream) {
  if (rows < 1 || np < 1) {
    set_last_error("marginal: bad shape rows=%lld np=%d", rows, np);
    return ADEPT_ERR_BAD_SHAPE;
  }
  ProfileScope prof("marginal", stream);
  marginal_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, stream>>>(f, w, rows, np, out);
  return check_launch("marginal_kernel");
}

// batched transpose in[b, n0, n1] -> out[b, n1, n0] through a padded 32 x 32 shared-memory tile
__global__ void __launch_bounds__(256) transpose_kernel(const double* __restrict__ in, double* __restrict__ out, int n0,
                                                        int n1) {
  __shared__ double tile[32][33];
  const long long base = (long long)blockIdx.z * n0 * n1;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8)
    if (i0 + r < n0 && j0 + tx < n1) tile[r][tx] = in[base + (long long)(i0 + r) * n1 + j0 + tx];
  __syncthreads();
  for (int r = ty; r < 32; r += 8)
    if (j0 + r < n1 && i0 + tx < n0) out[base + (long long)(j0 + r) * n0 + i0 + tx] = tile[tx][r];
}

int transpose_f64(const double* in, double* out, int batch, int n0, int n1, cudaStream_t stream) {
  if (batch < 1 || n0 < 1 || n1 < 1 || batch > 65535 || in == out) {
    set_last_error("transpose: bad shape batch=%d n0=%d n1=%d (or in-place)", batch, n0, n1);
    return ADEPT_ERR_BAD_SHAPE;
  }
  ProfileScope prof("transpose", stream);
  transpose_kernel<<<dim3((n1 + 31) / 32, (n0 + 31) / 32, batch), 256, 0, stream>>>(in, out, n0, n1);
  return check_launch("transpose_kernel");
}

}  // namespace adept
