/* adept_b200 -- C ABI of libadept_b200.so: the B200-native `vlasov-1d` time-step operators of ergodicio/adept.
 *
 * Drop-in boundary.  The reference is Python/JAX; the binding a maintainer adds is a `jax.ffi` custom call (or,
 * without jax, the ctypes stub in adept_b200/_lib.py) whose target forwards to the entry points below.  Each
 * entry point replaces one reference operator (file:line relative to the reference tree):
 *
 *   adept_b200_vdfdx_f64         SpaceExponential.push/__call__      adept/_vlasov1d/solvers/pushers/vlasov.py:234-251
 *   adept_b200_edfdv_exp_f64     VelocityExponential.push            adept/_vlasov1d/solvers/pushers/vlasov.py:74-91
 *   adept_b200_edfdv_spline_f64  VelocityCubicSpline.push            adept/_vlasov1d/solvers/pushers/vlasov.py:106-172
 *   adept_b200_moments_f64       compute_charge/current_density      adept/_vlasov1d/solvers/pushers/field.py:186-208,319-340
 *   adept_b200_poisson_f64       SpectralPoisson / BoltzmannPoisson  adept/_vlasov1d/solvers/pushers/field.py:210-224,282-298
 *   adept_b200_axpy_f64          AmpereSolver.__call__               adept/_vlasov1d/solvers/pushers/field.py:342-354
 *   adept_b200_ponderomotive_f64 ElectricFieldSolver (pond)          adept/_vlasov1d/solvers/pushers/field.py:495
 *   adept_b200_wave_step_f64     WaveSolver.__call__                 adept/_vlasov1d/solvers/pushers/field.py:109-157
 *   adept_b200_collide_f64       Collisions._apply_collisions+Krook  adept/_vlasov1d/solvers/pushers/fokker_planck.py:368-484
 *   adept_b200_collide_sc_f64    same + find_self_consistent_beta    adept/driftdiffusion.py:161-283, fokker_planck.py:139-210
 *   adept_b200_step_f64          VlasovMaxwell.__call__ (whole step) adept/_vlasov1d/solvers/vector_field.py:55-361
 *
 * Conventions
 *   - All pointers are DEVICE pointers (cudaMalloc / XLA buffers) unless the name ends in `_host`.
 *   - f is [batch, nx, nv] in C order (v contiguous), fp64; `batch` independent ensemble members.
 *   - Every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*); nothing synchronises, nothing is
 *     allocated except the per-(device, n) twiddle tables on first use -- call adept_b200_prepare() before CUDA-graph
 *     capture.  Callers own every buffer; nothing is retained after return.
 *   - Return value: 0 on success, negative error code otherwise; adept_b200_last_error() gives the message
 *     (thread-local).  No CPU fallback exists: without a CUDA device every compute entry point fails with
 *     ADEPT_B200_ERR_CUDA.
 */
#ifndef ADEPT_B200_H
#define ADEPT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define ADEPT_B200_OK 0
#define ADEPT_B200_ERR_BAD_SHAPE (-1)
#define ADEPT_B200_ERR_UNSUPPORTED (-2)
#define ADEPT_B200_ERR_CUDA (-3)
#define ADEPT_B200_ERR_BAD_ARG (-4)

/* Fokker-Planck model / differencing scheme selectors (fokker_planck.py:314-342) */
#define ADEPT_B200_FP_LENARD_BERNSTEIN 0
#define ADEPT_B200_FP_DOUGHERTY 1
#define ADEPT_B200_FP_SUPER_GAUSSIAN 2
#define ADEPT_B200_FP_CENTRAL 0
#define ADEPT_B200_FP_CHANG_COOPER 1

int adept_b200_version(void);
const char* adept_b200_last_error(void);

/* Per-kernel timing with CUDA events on the launching stream.  adept_b200_profile(1) clears the record and starts
 * bracketing every kernel launch of the library; adept_b200_profile(0) stops and clears.  adept_b200_profile_report
 * synchronises on the recorded events and writes one line per kernel, "<name> <launches> <total_ms>\n", into buf. */
/* Number of kernels this library has launched in the calling process (monotonic; bench.py differences it). */
long long adept_b200_launch_count(void);
int adept_b200_profile(int enable);
int adept_b200_profile_report(char* buf, int buflen);

/* Build (once per device) the twiddle tables for a transform length n = 2^k, 2 <= n <= 8192. */
int adept_b200_prepare(int n);

/* x-advection: f_out = irfft(exp(-i kx_m v_j dt) rfft(f_in, axis=x), axis=x).
 * v[nv] velocity grid; k1x = kx_real[1] = 2 pi / (nx dx); k1x_batch[batch] (nullable) overrides it per member.
 * nx a power of two (<= 8192) or any even number <= 4096 (chirp-z path, 4-8x the work), nv even.  In-place allowed. */
int adept_b200_vdfdx_f64(const double* f_in, double* f_out, int batch, int nx, int nv, const double* v, double dt,
                         double k1x, const double* k1x_batch, void* stream);

/* x-advection of LONG pencils of mixed length, nx = 2^a m with 6 <= a and m odd (m 2^max(a - 8, 0) <= 150), nv % 64 == 0
 * -- e.g. nx = 17280 = 128 x 135 of configs/vlasov-1d/iaw-turbulence-big*.yaml, whose pencils do not fit one SM: three
 * launches (128-point FFTs, 135-point DFTs + phase, inverse FFTs) through `scratch`, an array of the size of f that
 * aliases neither f_in nor f_out (f_in == f_out is allowed).  adept_b200_step_f64 takes this path on its own and uses
 * species.f_tmp as the scratch; adept_b200_poisson_f64 solves the field on such grids with the same launches. */
int adept_b200_vdfdx_scratch_f64(const double* f_in, double* f_out, double* scratch, int batch, int nx, int nv,
                                 const double* v, double dt, double k1x, const double* k1x_batch, void* stream);

/* x-advection fused with the velocity sum of the result (first stage of the charge density that the field solve
 * needs right after the push: SpaceExponential + compute_charge_density, vlasov.py:234-251 + field.py:197-208).
 * parts[nparts, batch*nx] receives per-CTA partial sums of sum_j f_out[b, i, j] (unscaled; rows beyond those used are
 * zero); nparts >= adept_b200_vdfdx_rho_parts(batch, nx, nv).  adept_b200_reduce_parts_f64 finishes the reduction in a
 * fixed order (deterministic): out[i] = (base ? base[i] : 0) + scale_b * ((sum_p parts[p, i]) * scale_a). */
int adept_b200_vdfdx_rho_parts(int batch, int nx, int nv);
int adept_b200_vdfdx_rho_f64(const double* f_in, double* f_out, int batch, int nx, int nv, const double* v, double dt,
                             double k1x, const double* k1x_batch, double* parts, int nparts, void* stream);
int adept_b200_reduce_parts_f64(const double* parts, int nparts, long long n, double scale_a, double scale_b,
                                const double* base, double* out, void* stream);

/* Fused v-advection + Fokker-Planck step on the same rows (VelocityExponential.push followed by Collisions,
 * vector_field.py:236-238): f_out = collide(edfdv_exp(f_in)), one read and one write of f for both operators.
 * model = Lenard-Bernstein (0) or Dougherty (1), scheme = central differencing (0) or Chang-Cooper (1), no Krook; nv a
 * power of two in [512, 8192], nx even.
 * Returns ADEPT_B200_ERR_UNSUPPORTED otherwise (callers then use the two separate entry points).  In-place allowed. */
int adept_b200_vpush_collide_f64(const double* f_in, double* f_out, int batch, int nx, int nv, const double* e,
                                 const double* dex, const double* pond, double charge, double mass, double dt,
                                 double k1v, const double* v, double dv, const double* nu_fp, int model, int scheme,
                                 void* stream);

/* ---- single grid sharded over the GPUs of one node: transposes fused into the v-row kernel ----------------
 * The reference's `grid.parallel: ["x", "v"]` decomposition alternates between a v-sharded layout f[nx, nv/P]
 * (x-advection) and an x-sharded one f[nx/P, nv] (v-advection + collisions); XLA inserts all-to-all transposes.  Here
 * every buffer stays v-sharded and the fused v-advection + collision kernel of a rank reads each cell of its rows
 * straight from the rank that owns the column and writes the result straight back, over NVLink peer memory
 * (in_peers / out_peers: host arrays of n_peers device pointers to the ranks' [nx_global, nv/n_peers] buffers, each
 * mapped into this process; n_peers a power of two <= 8).  A kernel row touches one 8 nv/P-byte segment per peer,
 * so the traffic is NVLink-friendly and no transpose pass or all-to-all exists; the x-advection stays purely local.
 * Ordering between ranks is the caller's (the rho all-reduce after the x-advection, one barrier after this kernel).
 * nx = rows owned by this rank (even), row0_global = index of its first row; shapes and operators as
 * adept_b200_vpush_collide_f64.  (Scattering the x-advection's 32-byte row pieces instead was measured: 148 us
 * against 79 us local at 4096 x 2048 per rank -- small remote stores waste the link.) */
int adept_b200_vpush_collide_p2p_f64(const double* const* in_peers_host, double* const* out_peers_host, int n_peers,
                                     long long row0_global, int nx, int nv, const double* e, const double* dex,
                                     const double* pond, double charge, double mass, double dt, double k1v,
                                     const double* v, double dv, const double* nu_fp, int model, int scheme,
                                     void* stream);

/* The same with the rank's input rows gathered by dedicated mover CTAs of the same launch: the first n_movers CTAs
 * (even, dividing nx) bulk-load whole rows from the owning ranks through shared memory into stage[nx, nv] (local
 * scratch) and raise round_counters[i] (nx / n_movers unsigned ints, zeroed by the call) when round i has landed; the
 * compute CTA of a row pair waits for its round and bulk-loads the pair from the stage, so a pair's NVLink latency no
 * longer sits between its CTA's start and its arithmetic (2 CTAs per SM cannot hide it otherwise).  Results still leave
 * as one TMA tensor store per rank and row.  Needs nv >= 2048.  n_movers = 0: the caller has filled the stage before
 * the launch (adept_b200_copy2d_f64 on a copy stream, chunk by chunk, while earlier chunks compute).  nx = rows of this
 * call, nx_global = rows of the ranks' buffers (0: nx * n_peers). */
int adept_b200_vpush_collide_p2p_staged_f64(const double* const* in_peers_host, double* const* out_peers_host,
                                            int n_peers, long long row0_global, int nx, int nv, const double* e,
                                            const double* dex, const double* pond, double charge, double mass,
                                            double dt, double k1v, const double* v, double dv, const double* nu_fp,
                                            int model, int scheme, double* stage, unsigned int* round_counters,
                                            int n_movers, long long nx_global, void* stream);

/* Strided device-to-device copy of `height` rows of `width` doubles (pitches in doubles), enqueued on `stream` as one
 * cudaMemcpy2DAsync: it runs on a copy engine, over NVLink when src is a peer-mapped buffer, and takes no SM.  The
 * sharded grid gathers the v-sharded row blocks of the other ranks into its x-sharded stage with it. */
int adept_b200_copy2d_f64(double* dst, long long dst_pitch, const double* src, long long src_pitch, long long width,
                          long long height, void* stream);

/* Sharded grid, first launch of a step: x-advection of this rank's nv_local velocity columns (SpaceExponential.push,
 * pushers/vlasov.py:234-251) with the field solve of the WHOLE grid in the tail of the same launch -- what the
 * reference's shard_map'ed compute_charge_density + SpectralPoissonSolver + LongitudinalElectricFieldDriver do with an
 * implicit all-reduce (field.py:21-33, 197-224).  Every persistent CTA pushes its slice of the rank's share of the
 * charge density (ion_share + charge dv sum over the local columns) into all ranks' inboxes over NVLink peer memory,
 * the rank raises its flag in every inbox, waits for the P flags of its own inbox, adds the P shares in rank order
 * (bit-identical rho on all ranks) and solves E = green (*) rho redundantly.  No collective library call, no second
 * launch; flags carry the monotonic step count `epoch` (>= 1, the same on every rank), inboxes are double-buffered by
 * its parity, so the only other ordering a step needs is one barrier after the v-row kernel.  The launch is
 * cooperative; it needs nx in {1024, 2048, 4096}, nv_local % 4 == 0 and all CTAs resident. */
typedef struct adept_b200_field_peers {
  int n_peers, my_rank;
  unsigned long long epoch;
  double* share_in[8];             /* rank r's inbox [2][n_peers][nx], mapped into this process */
  unsigned long long* flag_in[8];  /* rank r's flags [n_peers], zero before the first step */
  unsigned int* sync_counter;      /* local zero-initialised word (device-wide barriers of the launch) */
  const double* ion_share;         /* nullable [nx]: static ion background / n_peers */
  double dv, charge, dx;
  const double* green;             /* [nx] Re ifft(-i / kx) */
  double* rho;                     /* [nx] out */
  double* e;                       /* [nx] out */
  double* dex;                     /* [nx] out, written when n_ex > 0 */
  double* pond;                    /* [nx] out (zero: a_zero is) */
  const double* a_zero;            /* [nx + 2] zeros: the sharded path carries no transverse wave */
  int n_ex;
  const double* ex_space;          /* [n_ex, nx] */
  const double* ex_kx;             /* [n_ex, nx] */
  double ex_w[8], ex_a0[8], ex_tenv[8], ex_wt[8];
} adept_b200_field_peers;
int adept_b200_vdfdx_field_peers_f64(const double* f_in, double* f_out, int nx, int nv_local, const double* v_local,
                                     double dt, double k1x, double* parts, int nparts,
                                     const adept_b200_field_peers* fp, void* stream);

/* out[i] = sum_r peers[r][i], r = 0 .. n_peers-1 in that order (the nx-long all-reduce of the charge density of a
 * sharded grid without a collective library call: every rank reads every rank's share over NVLink peer memory and adds
 * them in rank order, so all ranks hold bit-identical sums).  peers_host: host array of device pointers mapped into
 * this process; the caller orders the ranks (a barrier before: shares written; one before they are overwritten). */
int adept_b200_sum_peers_f64(const double* const* peers_host, int n_peers, long long n, double* out, void* stream);

/* In-loop save moments in one pass over f (get_default_save_func / get_field_save_func, adept/_vlasov1d/storage.py:
 * 286-327, 119-162): out[k, row] = dv sum_j g_k(f_j, v_j), g = { f, f v, f v^2, f v^3, -|f| log|f|, f^2 }, out is
 * [6, batch*nx].  With f1 != NULL the distribution is the linear interpolation f0 + w (f1 - f0) that diffrax hands to
 * the save functions between two steps; it is never materialised. */
int adept_b200_save_moments_f64(const double* f0, const double* f1, double w, int batch, int nx, int nv,
                                const double* v, double dv, double* out, void* stream);

/* Distribution save on a coarser mesh (get_dist_save_func, {t, x, v} block, adept/_vlasov1d/storage.py:173-181 ->
 * interpax.interp2d, method "linear", extrap off): out[a, b] = bilinear interpolation of f[nx, nv] at (xq[a], vq[b]) on
 * the axes x[nx], v[nv]; NaN outside the grid.  With f1 != NULL the distribution is f0 + w (f1 - f0) (diffrax's dense
 * output between two steps), never materialised. */
int adept_b200_interp2d_f64(const double* f0, const double* f1, double w, int nx, int nv, const double* x,
                            const double* v, const double* xq, const double* vq, int nxq, int nvq, double* out,
                            void* stream);

/* Spectrum save (get_dist_save_func, {t, kx, v} block, adept/_vlasov1d/storage.py:183-190): out[b, m, j] =
 * |rfft(f[b, :, j])_m|, m = 0 .. nx/2; out is [batch, nx/2 + 1, nv].  nx a power of two <= 8192, nv even.  The caller
 * interpolates on its (kx, v) sample points with adept_b200_interp2d_f64 over the one-sided axis 2 pi rfftfreq(nx, dx). */
int adept_b200_abs_rfft_x_f64(const double* f, double* out, int batch, int nx, int nv, void* stream);

/* Hou-Li spectral filter along x (HouLiFilter.__call__, vlasov.py:215-220): f_out = irfft(filt[m] rfft(f_in, axis=x)),
 * filt[nx/2+1] real.  zeros_v: a device array of nv zeros (the x-advection kernels run with zero advection speed). */
int adept_b200_filter_x_f64(const double* f_in, double* f_out, int batch, int nx, int nv, const double* filt,
                            const double* zeros_v, void* stream);

/* v-advection (spectral): accel_i = (q (e_i + dex_i) + (q^2/m) pond_i)/m;
 * f_out = irfft(exp(-i kv_n dt accel_i) rfft(f_in, axis=v), axis=v).  e, dex, pond are [batch, nx]
 * (dex, pond nullable); k1v = kv_real[1] = 2 pi / (nv dv).  nv a power of two (<= 8192) or any even number <= 4096
 * (chirp-z path), nx even.  In-place allowed. */
int adept_b200_edfdv_exp_f64(const double* f_in, double* f_out, int batch, int nx, int nv, const double* e,
                             const double* dex, const double* pond, double charge, double mass, double dt, double k1v,
                             void* stream);

/* v-advection (semi-Lagrangian, local cubic Hermite; out-of-range queries -> 1e-30).  Out-of-place only. */
int adept_b200_edfdv_spline_f64(const double* f_in, double* f_out, int batch, int nx, int nv, const double* e,
                                const double* dex, const double* pond, double charge, double mass, double dt,
                                double dv, void* stream);

/* Velocity moments per x-row: s_k = sum_j f[., j] v_j^k (k = 0, 1, 2).  For each non-null out_k:
 *   out_k[row] = (base_k ? base_k[row] : 0) + scale_b[k] * (s_k * scale_a)
 * e.g. charge density: scale_a = dv, scale_b[0] = q, base_0 = ion background.  out/base: host arrays of 3 device
 * pointers; scale_b host array of 3 doubles (nullable -> 1). */
int adept_b200_moments_f64(const double* f, int batch, int nx, int nv, const double* v, double scale_a,
                           const double* const* base_host, double* const* out_host, const double* scale_b_host,
                           void* stream);

/* Field solve, one FFT of length nx per member.  mode 0: E = Re ifft(-i kmul fft(rho)), kmul = one_over_kx[nx].
 * mode 1 (Boltzmann electrons): kmul = kx[nx], kernel kx (Te/rho0) / (1 + lambda^2 kx^2), rho0 = mean(rho),
 * lambda_De < 0 selects lambda^2 = Te/rho0.  kmul_stride = 0 shares kmul between members, nx = per-member tables.
 * nx a power of two (<= 8192) or any even number <= 4096. */
int adept_b200_poisson_f64(const double* rho, const double* kmul, long long kmul_stride, double* e, int batch, int nx,
                           int mode, double Te, double lambda_De, void* stream);

/* out[i] = a[i] + s * b[i]  (Ampere: E = E_prev - dt j). */
int adept_b200_axpy_f64(const double* a, const double* b, double s, double* out, long long n, void* stream);

/* The plain Poisson solve (field.py:221-224) as a circular convolution with the Green's function
 * green[nx] = Re ifft(-i / kx), spread over nx/32 CTAs: e[b, i] = sum_j green[b, (i - j) mod nx] rho[b, j].
 * green_stride: 0 when every member shares one table, nx for per-member tables.  nx: power of two in [512, 8192]. */
int adept_b200_poisson_green_f64(const double* rho, const double* green, long long green_stride, double* e, int batch,
                                 int nx, void* stream);

/* out[r] = mean_i a[r, i], r < rows: the average over x that get_default_save_func takes of every velocity moment
 * (adept/_vlasov1d/storage.py:306-323), one CTA per row, fixed summation order.  `out` may be a pinned (mapped) host
 * buffer like adept_b200_field_energy_f64's: the six scalars of a save point then travel to the host with the launch. */
int adept_b200_row_means_f64(const double* a, int rows, long long n, double* out, void* stream);

/* Field-energy scalars of the default save function (mean_e2, mean_de2; adept/_vlasov1d/storage.py:316-317), one
 * launch: out[b] = {mean(e_b^2), mean(de_b^2)} of (e0, de0)[batch, nx], or of the state interpolated linearly towards
 * (e1, de1) with weight w when those are given (both or neither). */
int adept_b200_field_energy_f64(const double* e0, const double* de0, const double* e1, const double* de1, double w,
                                int batch, int nx, double* out, void* stream);

/* pond[b, i] = -0.5 (a[b, i+2]^2 - a[b, i]^2) / (2 dx); a is [batch, nx+2]. */
int adept_b200_ponderomotive_f64(const double* a, double* pond, int batch, int nx, double dx, void* stream);

/* One leap-frog step of the transverse wave equation with 2nd-order absorbing boundaries.  a, aold, djy, a_new are
 * [batch, nx+2]; ne_n / ne_np1 [batch, nx] are the electron charge densities before / after the step (nullable = 0).
 * The caller sets prev_a := a afterwards. */
int adept_b200_wave_step_f64(const double* a, const double* aold, const double* djy, const double* ne_n,
                             const double* ne_np1, double* a_new, int batch, int nx, double c, double dx, double dt,
                             void* stream);

/* Collisions on every x-row: implicit Fokker-Planck (delta form) then Krook.
 * nu_fp / nu_K: [batch, nx] collision frequencies; null disables the operator.  f_mx[nv]: Krook target (unit
 * density Maxwellian, fokker_planck.py:463-464).  n_out (nullable) receives sum_j f_out dv.  In-place allowed.
 * nv % 8 == 0 and nv <= 16384 (small even nv also accepted). */
int adept_b200_collide_f64(const double* f_in, double* f_out, int batch, int nx, int nv, const double* v, double dv,
                           double dt, const double* nu_fp, const double* nu_K, const double* f_mx, int model,
                           int scheme, int nodrag, double sg_m, double sg_ratio, double* n_out, void* stream);
/* Same with the self-consistent-beta refinement (find_self_consistent_beta, adept/driftdiffusion.py:161-283;
 * SuperGaussianDougherty.compute_beta, fokker_planck.py:139-210): beta of every row is refined by up to sc_max_steps
 * Newton iterations with optimistix's Cauchy termination (rtol, atol) before the operator is assembled. */
int adept_b200_collide_sc_f64(const double* f_in, double* f_out, int batch, int nx, int nv, const double* v, double dv,
                              double dt, const double* nu_fp, const double* nu_K, const double* f_mx, int model,
                              int scheme, int nodrag, double sg_m, double sg_ratio, double* n_out, int sc_max_steps,
                              double sc_rtol, double sc_atol, void* stream);

/* ---- adjoints (what the backward rule of a jax.custom_vjp around each operator calls) -------------------------------
 * x-advection and v-advection w.r.t. f: the forward entry points with dt -> -dt (the operators are real circulants with
 * unit-modulus symbols, so transpose = inverse).  Poisson: antisymmetric, rho_bar = -poisson(e_bar).  The rest: */

/* accel_bar[b, i] = sum_j g[b, i, j] * d f_out[b, i, j] / d accel[b, i] of adept_b200_edfdv_exp_f64 (same arguments;
 * f_in is the forward INPUT).  Chain to the fields with e_bar = dex_bar = accel_bar q/m, pond_bar = accel_bar q^2/m^2. */
int adept_b200_edfdv_exp_bwd_accel_f64(const double* f_in, const double* g, int batch, int nx, int nv, const double* e,
                                       const double* dex, const double* pond, double charge, double mass, double dt,
                                       double k1v, double* accel_bar, void* stream);

/* Adjoint of adept_b200_moments_f64: f_bar[row, j] (+)= sum_k coef[k] out_bar_k[row] v_j^k with coef[k] = scale_a *
 * scale_b[k] of the forward call; out_bar_host: 3 device pointers (nullable each); accumulate != 0 adds into f_bar. */
int adept_b200_moments_bwd_f64(const double* const* out_bar_host, const double* coef_host, int batch, int nx, int nv,
                               const double* v, int accumulate, double* f_bar, void* stream);

/* Adjoint of the Fokker-Planck step of adept_b200_collide_f64 (central differencing, Lenard-Bernstein or Dougherty, no
 * Krook): given the forward input f_in, the forward output f_new and the cotangent g of f_new, writes f_bar (cotangent
 * of f_in, including the dependence of the operator on the row moments vbar, T) and nu_bar[batch*nx] (nullable). */
int adept_b200_collide_bwd_f64(const double* f_in, const double* f_new, const double* g, double* f_bar, double* nu_bar,
                               int batch, int nx, int nv, const double* v, double dv, double dt, const double* nu_fp,
                               int model, int scheme, void* stream);

/* Adjoint of adept_b200_edfdv_spline_f64 (pinned by the reference's tests/test_vlasov1d/test_velocity_cubic_spline.py:
 * 49-70, gradients w.r.t. f and the field): f_bar[b, i, .] = the four cubic-Hermite taps of every output scattered back
 * (nullable), accel_bar[b, i] = sum_j g_ij d f'_ij / d accel_i (nullable; chain to the fields as for the spectral
 * push).  Same arguments as the forward call; f_in is the forward INPUT. */
int adept_b200_edfdv_spline_bwd_f64(const double* f_in, const double* g, int batch, int nx, int nv, const double* e,
                                    const double* dex, const double* pond, double charge, double mass, double dt,
                                    double dv, double* f_bar, double* accel_bar, void* stream);

/* Adjoint of the Krook step of adept_b200_collide_f64 (f' = f e^{-nu_K dt} + n f_mx (1 - e^{-nu_K dt}), n = dv sum f;
 * fokker_planck.py:463-484): f_bar (nullable) and nu_bar[batch*nx] (nullable) from the forward input f_in and the
 * cotangent g of the output. */
int adept_b200_krook_bwd_f64(const double* f_in, const double* g, int batch, int nx, int nv, double dv, double dt,
                             const double* nu_K, const double* f_mx, double* f_bar, double* nu_bar, void* stream);

/* ---- vlasov-1d2v (f[nx, nv, nvperp], cylindrical v_perp, adept/_vlasov1d2v/) -----------------------
 * The advection kernels serve the 2V layout unchanged (v_perp is a spectator: the x-advection sees nv * nvperp columns,
 * the v_par-advection is the strided-pencil kernel with one [nv, nvperp] member per x and per-member phase increments).
 * What is new: the v_perp marginal F[x, v] = sum_p f[x, v, p] wperp[p] (vector_field.py:36-38) that feeds the 1-D field
 * solve, a batched transpose [b, n0, n1] -> [b, n1, n0] (out of place) that lays the v_par rows of every (x, v_perp)
 * slice out contiguously, and the collision step with the operator coefficients taken from the marginal
 * (pushers/fokker_planck.py:104-140): a first call on the marginal rows with coef_out[rows, 2] records (vbar, beta) of
 * every row (after the self-consistent Newton refinement, if any; its f_out is the collided marginal), a second call on
 * the nx * nvperp slice rows with coef_in[rows / coef_div, 2] and coef_div = nvperp applies the same tridiagonal
 * operator to every slice (nu_fp is then indexed per group).  Lenard-Bernstein / Dougherty / dougherty_nodrag. */
int adept_b200_marginal_f64(const double* f, const double* wperp, long long rows, int nvperp, double* out, void* stream);
int adept_b200_transpose_f64(const double* in, double* out, int batch, int n0, int n1, void* stream);
int adept_b200_collide_coef_f64(const double* f_in, double* f_out, int batch, int nx, int nv, const double* v, double dv,
                                double dt, const double* nu_fp, int model, int scheme, int nodrag, int sc_max_steps,
                                double sc_rtol, double sc_atol, const double* coef_in, double* coef_out, int coef_div,
                                void* stream);

/* ---- single precision (explicit extra of SURVEY.md 8b; the reference itself always runs in fp64, _base_.py:287-292) --
 * The same kernels with f in float (generated from the fp64 sources, adept_b200/build.py): velocity grids, fields,
 * collision profiles and all scalars stay double, so phases, accelerations and operator coefficients are formed as
 * in the fp64 path and rounded once.  Bar: relative L2 <= 1e-5 against the fp64 oracle.  Power-of-two transform
 * lengths only; collisions: Lenard-Bernstein / Dougherty (central or Chang-Cooper) and Krook. */
int adept_b200_vdfdx_f32(const float* f_in, float* f_out, int batch, int nx, int nv, const double* v, double dt,
                         double k1x, const double* k1x_batch, void* stream);
int adept_b200_edfdv_exp_f32(const float* f_in, float* f_out, int batch, int nx, int nv, const double* e,
                             const double* dex, const double* pond, double charge, double mass, double dt, double k1v,
                             void* stream);
int adept_b200_collide_f32(const float* f_in, float* f_out, int batch, int nx, int nv, const double* v, double dv,
                           double dt, const double* nu_fp, const double* nu_K, const double* f_mx, int model,
                           int scheme, float* n_out, void* stream);

/* ---- whole time step ------------------------------------------------------------------------------------------------
 * adept_b200_step_f64 enqueues every kernel of one `vlasov-1d` step y -> y' on `stream` with no host work in between:
 * it replaces VlasovMaxwell.__call__ and what it calls (adept/_vlasov1d/solvers/vector_field.py:55-95 LeapfrogIntegrator,
 * :98-186 SixthOrderHamIntegrator, :232-253 VlasovPoissonFokkerPlanck, :308-361 VlasovMaxwell).  The caller evaluates
 * the O(1) time factors of the drivers and collision profiles for this step on the host (the reference evaluates the
 * same closed forms inside its jitted loop, field.py:21-33, functions.py:72-80,112-118) and passes them by value; all
 * O(nx) work runs on the device.  All per-row arrays are [batch*nx]; every pointer is a device pointer. */
#define ADEPT_B200_MAX_SPECIES 4
#define ADEPT_B200_MAX_DRIVERS 8
#define ADEPT_B200_MAX_SUBSTEPS 6

typedef struct adept_b200_species {
  const double* f_in; /* [batch, nx, nv] distribution at t_n (not modified) */
  double* f_out;      /* [batch, nx, nv] distribution at t_{n+1}; must not alias f_in */
  double* f_tmp;      /* [batch, nx, nv] scratch, required for edfdv = cubic-spline (else may be null) */
  const double* v;    /* [nv] */
  int nv;
  double dv, k1v, charge, mass;
  double* rho_parts; /* [rho_nparts, batch*nx] scratch for the fused charge density (adept_b200_vdfdx_rho_parts) */
  int rho_nparts;
} adept_b200_species;

typedef struct adept_b200_step {
  int batch, nx, n_species;
  adept_b200_species species[ADEPT_B200_MAX_SPECIES];
  int electron_species;    /* index of the species named "electron" (wave-equation density), -1 if none */
  int collide_species;     /* index of the species the collision operators act on (reference species) */
  int time_integrator;     /* 0 leapfrog, 1 sixth-order Hamiltonian splitting */
  int edfdv;               /* 0 exponential (spectral), 1 cubic-spline */
  int field;               /* 0 poisson, 1 poisson-boltzmann, 2 ampere */
  double dt, dx, k1x;
  const double* k1x_batch; /* [batch] per-member 2 pi / (nx dx), nullable */
  const double* ion_charge; /* [batch*nx] static background added to rho (poisson only), nullable */
  const double* kmul;       /* poisson: one_over_kx; poisson-boltzmann: kx ([nx], or per member with kmul_stride = nx) */
  long long kmul_stride;
  double Te, lambda_De;
  const double* e_in; /* [batch*nx] E at t_n (ampere) */
  double* e_out;      /* [batch*nx] */
  double* dex;        /* [n_substeps, batch*nx] driver field at the substep times (output: the saved `de` is a row) */
  const double* a;    /* [batch, nx+2] */
  const double* prev_a;
  const double* djy;  /* [batch, nx+2] transverse current source at t + dt_array[1] */
  double* a_out;      /* [batch, nx+2]; only written when wave_on */
  double c_light;
  int wave_on;        /* 0: a == prev_a == 0 and no Ey driver -> the wave update is the identity and is skipped */
  double *pond, *rho, *ne_n, *ne_np1; /* [batch*nx] scratch (ne_* only when wave_on) */
  /* longitudinal drivers: dex[s, i] = sum_d ((tenv[s][d] * space[d, i]) * w[d]) * a0[d] * sin(kx[d, i] - wt[s][d]) */
  int n_ex;
  const double* ex_space; /* [n_ex, batch*nx] spatial envelope */
  const double* ex_kx;    /* [n_ex, batch*nx] k0 * x */
  double ex_w[ADEPT_B200_MAX_DRIVERS], ex_a0[ADEPT_B200_MAX_DRIVERS];
  double ex_tenv[ADEPT_B200_MAX_SUBSTEPS][ADEPT_B200_MAX_DRIVERS];
  double ex_wt[ADEPT_B200_MAX_SUBSTEPS][ADEPT_B200_MAX_DRIVERS];
  /* collisions (fokker_planck.py:272-484) */
  int fp_on, krook_on, fp_model, fp_scheme, fp_nodrag;
  double sg_m, sg_ratio;
  const double* nu_fp_space; /* [batch*nx] */
  const double* nu_K_space;  /* [batch*nx] */
  double nu_fp_time, nu_K_time; /* nu(x, t) = time * space */
  const double* f_mx;        /* [nv] Krook Maxwellian */
  /* one zero-initialised device word owned by this integration (the library resets it after every use).  When given,
   * a single large grid (batch == 1, nx in {1024, 2048, 4096}) runs driver + ponderomotive force + charge density +
   * Poisson solve as ONE launch whose last-arriving CTA solves for E; null selects the separate kernels. */
  unsigned int* sync_counter;
  /* ensembles whose members differ in driver frequency / amplitude (parameter scans): per-row values
   * [n_ex, batch*nx] that replace ex_w[d] / ex_a0[d] (both nullable), and the substep times t + dt_array[s] from which
   * the per-row phase w t is formed on the device (only read when ex_w_row is given). */
  const double* ex_w_row;
  const double* ex_a0_row;
  double ex_t[ADEPT_B200_MAX_SUBSTEPS];
  /* self-consistent beta (terms.fokker_planck.self_consistent_beta, fokker_planck.py:295-301): Newton iterations
   * (0 = off, the default), rtol, atol */
  int fp_sc_steps;
  double fp_sc_rtol, fp_sc_atol;
  /* Green's function of the Poisson solve, green[nx] = Re ifft(-i / kx) (nullable).  When given together with
   * sync_counter, a single large one-species grid (batch == 1, nx in {1024, 2048, 4096}, field = poisson) solves the
   * field in the tail of the x-advection launch: E = green (*) rho as a circular convolution (field.py:221-224 is
   * linear in rho), no separate field launch. */
  const double* poisson_green;
  /* dfdt diagnostics of the reference species (diagnostics.diag-vlasov-dfdt / diag-fp-dfdt, vector_field.py:245-250):
   * [batch, nx, nv] outputs (f_vlasov - f) / dt and (f_fp - f_vlasov) / dt, each nullable; diag_species indexes the
   * species they refer to ("electron" if present, else the first).  They need the intermediate distribution, so a
   * step that asks for them does not take the fused v-push + collision kernel. */
  double* diag_vlasov_dfdt;
  double* diag_fp_dfdt;
  int diag_species;
  /* Hou-Li spectral filter in x applied to every species after the collisions (terms.hou_li_filter, vlasov.py:187-220):
   * real multiplier per mode, [nx/2 + 1] (nullable = off) */
  const double* hou_li_filt;
  /* Device-resident time factors (nullable).  When given, the kernels read ex_tenv / ex_wt / ex_t / nu_fp_time /
   * nu_K_time from this [ADEPT_B200_TIME_ROW_LEN] device array instead of the by-value fields above, so a CUDA graph
   * captured around adept_b200_step_f64 can be replayed for later steps: the caller keeps a [n_steps, LEN] table on the
   * device and puts adept_b200_time_row_advance in front of every captured step.  Layout: tenv[s][d] at 8 s + d,
   * wt[s][d] at 48 + 8 s + d, nu_fp_time at 96, nu_K_time at 97, ex_t[s] at 98 + s. */
  const double* time_row;
} adept_b200_step;

/* Longitudinal driver field at one time (LongitudinalElectricFieldDriver.__call__, field.py:21-33), one launch:
 * dex[i] = sum_d ((tenv[d] * ex_space[d, i]) * w[d]) * a0[d] * sin(ex_kx[d, i] - wt[d]); w / a0 / tenv / wt are HOST arrays
 * of n_ex entries (the step descriptor's ex_w, ex_a0, ex_tenv[s], ex_wt[s]); ex_space / ex_kx device arrays [n_ex, n].
 * For callers that compose the step themselves (the sharded grid). */
int adept_b200_ex_driver_f64(const double* ex_space, const double* ex_kx, int n_ex, const double* w_host,
                             const double* a0_host, const double* tenv_host, const double* wt_host, long long n,
                             double* dex, void* stream);

#define ADEPT_B200_TIME_ROW_LEN 104

int adept_b200_step_f64(const adept_b200_step* step, void* stream);

/* row[0 .. LEN) = table[min(*counter, n_rows - 1)][0 .. LEN); *counter += 1.  One tiny launch; capturable.  counter is
 * a device long long owned by the caller (set it to the index of the step the next replay starts from). */
int adept_b200_time_row_advance(const double* table, long long n_rows, long long* counter, double* row, void* stream);

/* ---- whole-step backward (what the _bwd rule of a jax.custom_vjp around adept_b200_step_f64 calls) ------------------
 * Reverse mode through ONE leapfrog step (LeapfrogIntegrator + Collisions, vector_field.py:87-95, 236-238) of a single
 * species with field = poisson, edfdv = exponential or cubic-spline, Fokker-Planck (Lenard-Bernstein / Dougherty, central
 * or Chang-Cooper, no self-consistent beta) and Krook optional, transverse wave off.  `step` is the descriptor of the
 * forward call, which must have run first: its e_out, dex and pond arrays still hold the forward values, and f_in is
 * the forward input (f_out is not read).  The intermediates f* = vdfdx(f) and f** = edfdv(f*) are recomputed into the
 * scratch buffers.  Outputs: f_in_bar; dex_bar[batch*nx] (cotangent of the driver field, nullable); nu_fp_bar /
 * nu_K_bar[batch*nx] (cotangents of nu(x, t) = time * space, nullable).  Returns ADEPT_B200_ERR_UNSUPPORTED for
 * anything else (the operator-level adjoints above compose every other case). */
typedef struct adept_b200_step_bwd {
  const double* f_out_bar; /* [batch, nx, nv] cotangent of species[0].f_out */
  const double* e_out_bar; /* [batch*nx] cotangent of e_out (nullable) */
  double* f_in_bar;        /* [batch, nx, nv] */
  double* dex_bar;         /* [batch*nx] nullable */
  double* nu_fp_bar;       /* [batch*nx] nullable */
  double* nu_K_bar;        /* [batch*nx] nullable */
  double* scratch_f[3];    /* three [batch, nx, nv] buffers */
  double* scratch_row[2];  /* two [batch*nx] buffers */
} adept_b200_step_bwd;

int adept_b200_step_bwd_f64(const adept_b200_step* step, const adept_b200_step_bwd* bwd, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADEPT_B200_H */
