// adept_b200_step_f64: one whole vlasov-1d time step enqueued from native code (no host work between kernels).
// Reference composition (file:line relative to /root/reference):
//   VlasovMaxwell.__call__            adept/_vlasov1d/solvers/vector_field.py:308-361
//   VlasovPoissonFokkerPlanck         vector_field.py:232-253
//   LeapfrogIntegrator                vector_field.py:75-95
//   SixthOrderHamIntegrator           vector_field.py:113-186
//   LongitudinalElectricFieldDriver   adept/_vlasov1d/solvers/pushers/field.py:21-33
#include "../../include/adept_b200.h"
#include "common.cuh"
#include "internal.h"

namespace adept {

// ---- Ex driver field at every substep time ---------------------------------------------------------------------
struct DriverArgs {
  int n_ex, n_sub;
  long long n;  // batch * nx
  const double* space;
  const double* kx;
  double* dex;
  double w[ADEPT_B200_MAX_DRIVERS], a0[ADEPT_B200_MAX_DRIVERS];
  double tenv[ADEPT_B200_MAX_SUBSTEPS][ADEPT_B200_MAX_DRIVERS];
  double wt[ADEPT_B200_MAX_SUBSTEPS][ADEPT_B200_MAX_DRIVERS];
  const double* w_row;   // nullable [n_ex, n]: per-row frequency (ensemble members with different drivers)
  const double* a0_row;  // nullable [n_ex, n]
  double t_sub[ADEPT_B200_MAX_SUBSTEPS];
  const double* trow;    // nullable: device-resident time row (common.cuh) that replaces tenv / wt / t_sub
};

__global__ void __launch_bounds__(256) ex_driver_kernel(DriverArgs p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  for (int s = 0; s < p.n_sub; s++) {
    double total = 0.0;
    for (int d = 0; d < p.n_ex; d++) {
      // field.py:21-26: env(x, t) * (w0 + dw0) * a0 * sin(k0 x - (w0 + dw0) t), env = time_env * space_env
      const double tenv = p.trow ? p.trow[TROW_TENV + 8 * s + d] : p.tenv[s][d];
      const double factor = __dmul_rn(tenv, p.space[d * p.n + i]);
      const double w = p.w_row ? p.w_row[d * p.n + i] : p.w[d];
      const double a0 = p.a0_row ? p.a0_row[d * p.n + i] : p.a0[d];
      const double wt = p.w_row ? __dmul_rn(w, p.trow ? p.trow[TROW_EX_T + s] : p.t_sub[s])
                                : (p.trow ? p.trow[TROW_WT + 8 * s + d] : p.wt[s][d]);
      const double amp = __dmul_rn(__dmul_rn(factor, w), a0);
      total = __dadd_rn(total, __dmul_rn(amp, sin(__dsub_rn(p.kx[d * p.n + i], wt))));
    }
    p.dex[(long long)s * p.n + i] = total;
  }
}

#define ADEPT_TRY(expr)         \
  do {                          \
    const int rc_ = (expr);     \
    if (rc_ != ADEPT_OK) return rc_; \
  } while (0)

namespace {

struct StepCtx {
  const adept_b200_step& s;
  cudaStream_t st;
  long long n;  // batch * nx
  mutable bool have_parts[ADEPT_B200_MAX_SPECIES] = {false, false, false, false};  // set by the last push_x
  mutable bool field_done = false;  // set by push_x when the field solve ran in the tail of the x-advection launch

  // one species, one member, plain Poisson, a large power-of-two grid: the x-advection launch also solves the field
  bool can_tail_field(const double* fin, const double* fout) const {
    const adept_b200_species& sp = s.species[0];
    return s.n_species == 1 && s.field == 0 && s.sync_counter && s.poisson_green && sp.rho_parts &&
           vdfdx_tma_supported(fin, fout, s.nx, sp.nv) && sp.rho_nparts >= vdfdx_tma_parts(s.batch, s.nx, sp.nv) &&
           vdfdx_tma_field_supported(s.batch, s.nx, sp.nv);
  }

  // x-advection of every species, cur[k] -> dst[k]; accumulates the charge-density partial sums when `want_rho`
  int push_x(const double* const* cur, double* const* dst, double dt, bool want_rho, bool driver_here = false) const {
    field_done = false;
    if (want_rho && can_tail_field(cur[0], dst[0])) {
      const adept_b200_species& sp = s.species[0];
      FieldTail ft = {};
      ft.counter = s.sync_counter, ft.base = s.ion_charge, ft.dv = sp.dv, ft.charge = sp.charge;
      ft.rho = s.rho, ft.e = s.e_out, ft.green = s.poisson_green, ft.a = s.a, ft.pond = s.pond, ft.dx = s.dx;
      ft.n_ex = driver_here ? s.n_ex : 0, ft.ex_space = s.ex_space, ft.ex_kx = s.ex_kx, ft.dex = s.dex;
      ft.trow = s.time_row;
      for (int d = 0; d < s.n_ex; d++)
        ft.ex_w[d] = s.ex_w[d], ft.ex_a0[d] = s.ex_a0[d], ft.ex_tenv[d] = s.ex_tenv[0][d], ft.ex_wt[d] = s.ex_wt[0][d];
      ADEPT_TRY(vdfdx_tma_f64(cur[0], dst[0], s.batch, s.nx, sp.nv, sp.v, dt, s.k1x_batch, s.k1x, sp.rho_parts, st,
                              nullptr, &ft));
      have_parts[0] = true;
      field_done = true;
      return ADEPT_OK;
    }
    for (int k = 0; k < s.n_species; k++) {
      const adept_b200_species& sp = s.species[k];
      have_parts[k] = false;
      if (bigx_supported(s.nx, sp.nv) && !vdfdx_tma_supported(cur[k], dst[k], s.nx, sp.nv)) {
        // long mixed-length pencils go through a scratch array: whichever of f_tmp / f_out is dead right now
        double* scratch = (dst[k] == sp.f_out || cur[k] == sp.f_out) ? sp.f_tmp : sp.f_out;
        if (!scratch || scratch == dst[k] || scratch == cur[k]) {
          set_last_error("step: nx=%d needs species.f_tmp as the scratch array of the x-advection", s.nx);
          return ADEPT_ERR_BAD_ARG;
        }
        ADEPT_TRY(bigx_apply_f64(cur[k], dst[k], scratch, s.batch, s.nx, sp.nv, sp.v, dt, s.k1x_batch, s.k1x, nullptr, 0,
                                 st));
        continue;
      }
      if (want_rho && sp.rho_parts && vdfdx_tma_supported(cur[k], dst[k], s.nx, sp.nv) &&
          sp.rho_nparts >= vdfdx_tma_parts(s.batch, s.nx, sp.nv)) {
        if (s.batch > 1) {  // a single member is overwritten by every CTA of the x-advection; ensembles accumulate
          cudaError_t err = cudaMemsetAsync(sp.rho_parts, 0, (size_t)sp.rho_nparts * n * sizeof(double), st);
          if (err != cudaSuccess) {
            set_last_error("step: cudaMemsetAsync(rho_parts): %s", cudaGetErrorString(err));
            return ADEPT_ERR_CUDA;
          }
        }
        ADEPT_TRY(vdfdx_tma_f64(cur[k], dst[k], s.batch, s.nx, sp.nv, sp.v, dt, s.k1x_batch, s.k1x, sp.rho_parts, st));
        have_parts[k] = true;
      } else if (vdfdx_tma_supported(cur[k], dst[k], s.nx, sp.nv)) {
        ADEPT_TRY(vdfdx_tma_f64(cur[k], dst[k], s.batch, s.nx, sp.nv, sp.v, dt, s.k1x_batch, s.k1x, nullptr, st));
      } else {
        ADEPT_TRY(vdfdx_f64(cur[k], dst[k], s.batch, s.nx, sp.nv, sp.v, dt, s.k1x_batch, s.k1x, st));
      }
    }
    return ADEPT_OK;
  }

  // true when the next field_solve(from_parts = true) runs as ONE fused launch (field.cu: field_fused_kernel)
  bool can_fuse_field() const {
    if (s.field == 2 || !s.sync_counter || !field_fused_supported(s.batch, s.nx)) return false;
    for (int k = 0; k < s.n_species; k++)
      if (!have_parts[k]) return false;
    // rows 0..3 of the first partial-sum array double as the transposition scratch of the distributed solve
    return vdfdx_tma_parts(s.batch, s.nx, s.species[0].nv) >= 4;
  }

  // true when the next field_solve runs as one launch per ensemble member (field.cu: field_member_kernel): small
  // grids, Poisson or Boltzmann-Poisson, velocity sums read from f (no partial sums from a TMA x-advection)
  bool can_member_field(bool from_parts) const {
    if (s.field == 2 || !field_member_supported(s.nx)) return false;
    if (from_parts)
      for (int k = 0; k < s.n_species; k++)
        if (have_parts[k]) return false;
    return true;
  }

  // (pond, e) = field_solve(f); field.py:479-497.  from_parts: the velocity sums come from the preceding push_x.
  // driver_here: also evaluate the Ex driver field of substep 0 (leapfrog) in the fused launch.
  int field_solve(const double* const* cur, bool from_parts, double dt, bool driver_here = false) const {
    if (from_parts && field_done) {  // already solved in the tail of the x-advection launch
      field_done = false;
      return ADEPT_OK;
    }
    if (from_parts && can_fuse_field()) {
      const double* parts[ADEPT_B200_MAX_SPECIES];
      int nparts[ADEPT_B200_MAX_SPECIES];
      double dv[ADEPT_B200_MAX_SPECIES], charge[ADEPT_B200_MAX_SPECIES];
      for (int k = 0; k < s.n_species; k++) {
        const adept_b200_species& sp = s.species[k];
        parts[k] = sp.rho_parts, nparts[k] = vdfdx_tma_parts(s.batch, s.nx, sp.nv), dv[k] = sp.dv, charge[k] = sp.charge;
      }
      return field_fused_f64(s.n_species, parts, nparts, dv, charge, s.field == 0 ? s.ion_charge : nullptr, s.rho, s.nx,
                             s.a, s.pond, s.dx, driver_here ? s.n_ex : 0, s.ex_space, s.ex_kx, s.dex, s.ex_w, s.ex_a0,
                             s.ex_tenv[0], s.ex_wt[0], s.kmul, s.e_out, s.field == 1 ? 1 : 0, s.Te, s.lambda_De,
                             s.sync_counter, st);
    }
    if (can_member_field(from_parts)) {
      const double* fs[ADEPT_B200_MAX_SPECIES];
      int nvs[ADEPT_B200_MAX_SPECIES];
      double dv[ADEPT_B200_MAX_SPECIES], charge[ADEPT_B200_MAX_SPECIES];
      for (int k = 0; k < s.n_species; k++) {
        const adept_b200_species& sp = s.species[k];
        fs[k] = cur[k], nvs[k] = sp.nv, dv[k] = sp.dv, charge[k] = sp.charge;
      }
      return field_member_f64(s.n_species, fs, nvs, dv, charge, s.field == 0 ? s.ion_charge : nullptr, s.rho, s.batch,
                              s.nx, s.a, s.pond, s.dx, driver_here ? s.n_ex : 0, s.ex_space, s.ex_kx, s.dex, s.ex_w,
                              s.ex_a0, s.ex_tenv[0], s.ex_wt[0], s.ex_w_row, s.ex_a0_row, s.ex_t[0], s.kmul,
                              s.kmul_stride, s.e_out, s.field == 1 ? 1 : 0, s.Te, s.lambda_De, st);
    }
    ADEPT_TRY(ponderomotive_f64(s.a, s.pond, s.batch, s.nx, s.dx, st));
    if (s.field == 2) {  // ampere: E = E_prev - dt * sum_s q_s dv_s sum_v v f_s   (field.py:330-354)
      const double* base = nullptr;
      for (int k = 0; k < s.n_species; k++) {
        const adept_b200_species& sp = s.species[k];
        const double* bases[3] = {nullptr, base, nullptr};
        double* outs[3] = {nullptr, s.rho, nullptr};
        const double scale_b[3] = {1.0, sp.charge, 1.0};
        ADEPT_TRY(moments_f64(cur[k], s.batch, s.nx, sp.nv, sp.v, sp.dv, bases, outs, scale_b, st));
        base = s.rho;
      }
      return axpy_f64(s.e_in, s.rho, -dt, s.e_out, n, st);
    }
    // rho = sum_s q_s dv_s sum_v f_s (+ static background for plain poisson); field.py:197-208
    const double* base = (s.field == 0) ? s.ion_charge : nullptr;
    for (int k = 0; k < s.n_species; k++) {
      const adept_b200_species& sp = s.species[k];
      if (from_parts && have_parts[k]) {
        ADEPT_TRY(reduce_parts_f64(sp.rho_parts, sp.rho_nparts, n, sp.dv, sp.charge, base, s.rho, st));
      } else {
        const double* bases[3] = {base, nullptr, nullptr};
        double* outs[3] = {s.rho, nullptr, nullptr};
        const double scale_b[3] = {sp.charge, 1.0, 1.0};
        ADEPT_TRY(moments_f64(cur[k], s.batch, s.nx, sp.nv, nullptr, sp.dv, bases, outs, scale_b, st));
      }
      base = s.rho;
    }
    return poisson_dispatch_f64(s.rho, s.kmul, s.kmul_stride, s.e_out, s.batch, s.nx, s.field == 1 ? 1 : 0, s.Te,
                                s.lambda_De, st);
  }

  // v-advection of every species cur[k] -> dst[k] under e_out + dex[sub] and pond
  int push_v(const double* const* cur, double* const* dst, double dt, int sub) const {
    const double* dex = s.dex + (long long)sub * n;
    for (int k = 0; k < s.n_species; k++) {
      const adept_b200_species& sp = s.species[k];
      if (s.edfdv == 0)
        ADEPT_TRY(edfdv_exp_f64(cur[k], dst[k], s.batch, s.nx, sp.nv, s.e_out, dex, s.pond, sp.charge, sp.mass, dt,
                                sp.k1v, st));
      else
        ADEPT_TRY(edfdv_spline_f64(cur[k], dst[k], s.batch, s.nx, sp.nv, s.e_out, dex, s.pond, sp.charge, sp.mass, dt,
                                   sp.dv, st));
    }
    return ADEPT_OK;
  }

  int electron_density(const double* f, double* out) const {  // vector_field.py:297-306
    const adept_b200_species& sp = s.species[s.electron_species];
    double* outs[3] = {out, nullptr, nullptr};
    const double scale_b[3] = {sp.charge, 1.0, 1.0};
    return moments_f64(f, s.batch, s.nx, sp.nv, nullptr, sp.dv, nullptr, outs, scale_b, st);
  }
};

}  // namespace

static int validate(const adept_b200_step& s) {
  if (s.batch < 1 || s.nx < 2 || s.n_species < 1 || s.n_species > ADEPT_B200_MAX_SPECIES) {
    set_last_error("step: bad shape batch=%d nx=%d n_species=%d", s.batch, s.nx, s.n_species);
    return ADEPT_ERR_BAD_SHAPE;
  }
  if (s.time_integrator < 0 || s.time_integrator > 1 || s.edfdv < 0 || s.edfdv > 1 || s.field < 0 || s.field > 2) {
    set_last_error("step: unknown integrator=%d / edfdv=%d / field=%d", s.time_integrator, s.edfdv, s.field);
    return ADEPT_ERR_BAD_ARG;
  }
  if (s.field == 2 && s.time_integrator != 0) {
    set_last_error("step: ampere + sixth has not been implemented (vector_field.py:455-463)");
    return ADEPT_ERR_UNSUPPORTED;
  }
  if (s.n_ex < 0 || s.n_ex > ADEPT_B200_MAX_DRIVERS) {
    set_last_error("step: at most %d Ex drivers (got %d)", ADEPT_B200_MAX_DRIVERS, s.n_ex);
    return ADEPT_ERR_UNSUPPORTED;
  }
  if (!s.e_out || !s.dex || !s.pond || !s.rho || !s.a || (s.field != 2 && !s.kmul) || (s.field == 2 && !s.e_in)) {
    set_last_error("step: null field / scratch pointer");
    return ADEPT_ERR_BAD_ARG;
  }
  if (s.n_ex > 0 && (!s.ex_space || !s.ex_kx)) {
    set_last_error("step: Ex drivers need ex_space and ex_kx");
    return ADEPT_ERR_BAD_ARG;
  }
  for (int k = 0; k < s.n_species; k++) {
    const adept_b200_species& sp = s.species[k];
    if (!sp.f_in || !sp.f_out || !sp.v || sp.f_in == sp.f_out || (s.edfdv == 1 && !sp.f_tmp)) {
      set_last_error("step: species %d: null or aliased distribution buffers", k);
      return ADEPT_ERR_BAD_ARG;
    }
  }
  if ((s.fp_on || s.krook_on) && (s.collide_species < 0 || s.collide_species >= s.n_species)) {
    set_last_error("step: collide_species=%d out of range", s.collide_species);
    return ADEPT_ERR_BAD_ARG;
  }
  if ((s.fp_on && !s.nu_fp_space) || (s.krook_on && (!s.nu_K_space || !s.f_mx))) {
    set_last_error("step: collision operators need their nu profile (and f_mx for Krook)");
    return ADEPT_ERR_BAD_ARG;
  }
  if ((s.diag_vlasov_dfdt || s.diag_fp_dfdt) && (s.diag_species < 0 || s.diag_species >= s.n_species)) {
    set_last_error("step: diag_species=%d out of range", s.diag_species);
    return ADEPT_ERR_BAD_ARG;
  }
  if (s.wave_on && (!s.prev_a || !s.djy || !s.a_out || (s.electron_species >= 0 && (!s.ne_n || !s.ne_np1)))) {
    set_last_error("step: wave_on needs prev_a, djy, a_out and the ne_n / ne_np1 scratch");
    return ADEPT_ERR_BAD_ARG;
  }
  return ADEPT_OK;
}

namespace {
struct TimeRowScope {  // launchers called from this thread read the step's device-resident time row (common.cuh)
  explicit TimeRowScope(const double* row) { set_current_time_row(row); }
  ~TimeRowScope() { set_current_time_row(nullptr); }
};
}  // namespace

int step_f64(const adept_b200_step& s, cudaStream_t st) {
  ADEPT_TRY(validate(s));
  TimeRowScope time_row_scope(s.time_row);
  StepCtx c{s, st, (long long)s.batch * s.nx};
  const int n_sub = s.time_integrator == 0 ? 1 : ADEPT_B200_MAX_SUBSTEPS;

  // drivers at the substep times (vector_field.py:319); the leapfrog step defers this until it knows whether the
  // fused field kernel evaluates the driver itself
  auto launch_drivers = [&]() -> int {
    DriverArgs d = {};
    d.n_ex = s.n_ex, d.n_sub = n_sub, d.n = c.n, d.space = s.ex_space, d.kx = s.ex_kx, d.dex = s.dex;
    d.w_row = s.ex_w_row, d.a0_row = s.ex_a0_row, d.trow = s.time_row;
    for (int j = 0; j < ADEPT_B200_MAX_SUBSTEPS; j++) d.t_sub[j] = s.ex_t[j];
    for (int k = 0; k < ADEPT_B200_MAX_DRIVERS; k++) {
      d.w[k] = s.ex_w[k], d.a0[k] = s.ex_a0[k];
      for (int j = 0; j < ADEPT_B200_MAX_SUBSTEPS; j++) d.tenv[j][k] = s.ex_tenv[j][k], d.wt[j][k] = s.ex_wt[j][k];
    }
    ProfileScope prof("ex_driver", st);
    ex_driver_kernel<<<(unsigned)((c.n + 255) / 256), 256, 0, st>>>(d);
    return check_launch("ex_driver_kernel");
  };
  if (s.time_integrator != 0) ADEPT_TRY(launch_drivers());

  const bool wave = s.wave_on != 0;
  const bool wave_density = wave && s.electron_species >= 0;
  if (wave_density) ADEPT_TRY(c.electron_density(s.species[s.electron_species].f_in, s.ne_n));  // vector_field.py:336

  const double* cur[ADEPT_B200_MAX_SPECIES];
  double* out[ADEPT_B200_MAX_SPECIES];
  double* tmp[ADEPT_B200_MAX_SPECIES];
  for (int k = 0; k < s.n_species; k++) cur[k] = s.species[k].f_in, out[k] = s.species[k].f_out, tmp[k] = s.species[k].f_tmp;
  const bool spline = s.edfdv == 1;
  const bool want_diag = s.diag_vlasov_dfdt || s.diag_fp_dfdt;  // the diagnostics need the intermediate f_vlasov
  bool collided = false;  // set when the fused v-push + collision kernel already applied the operator

  if (s.time_integrator == 0) {
    // leapfrog (vector_field.py:87-95): f* = vdfdx(f); (pond, e) = field(f*); f' = edfdv(f*, e + dex[0], pond)
    const bool want_rho = s.field != 2;
    double* const* xdst = spline ? tmp : out;  // the cubic stencil cannot run in place
    // the driver field of per-row parameter scans (ex_w_row) is only evaluated by the driver kernel
    ADEPT_TRY(c.push_x(cur, xdst, s.dt, want_rho, s.n_ex > 0 && !s.ex_w_row && !s.ex_a0_row));
    const double* fstar[ADEPT_B200_MAX_SPECIES];
    for (int k = 0; k < s.n_species; k++) fstar[k] = xdst[k];
    const bool tail_driver = c.field_done && s.n_ex > 0 && !s.ex_w_row && !s.ex_a0_row;
    const bool fused_field = tail_driver || (!c.field_done && want_rho && (c.can_fuse_field() || c.can_member_field(true)));
    // with no driver the fused field kernels leave dex untouched: the driver kernel zero-fills it
    if (!fused_field || s.n_ex == 0) ADEPT_TRY(launch_drivers());
    ADEPT_TRY(c.field_solve(fstar, want_rho, s.dt, fused_field));
    // the colliding species takes the fused v-push + Fokker-Planck kernel when its shape and operator allow it
    const int kc = s.collide_species;
    if (s.fp_on && !s.krook_on && !spline && s.fp_sc_steps == 0 && !want_diag && kc >= 0 && kc < s.n_species &&
        vpush_collide_supported(s.nx, s.species[kc].nv, s.fp_model, s.fp_scheme, s.fp_nodrag)) {
      for (int k = 0; k < s.n_species; k++) {
        const adept_b200_species& sp = s.species[k];
        if (k == kc) {
          ADEPT_TRY(vpush_collide_f64(fstar[k], out[k], s.batch, s.nx, sp.nv, s.e_out, s.dex, s.pond, sp.charge,
                                      sp.mass, s.dt, sp.k1v, sp.v, sp.dv, s.nu_fp_space, s.nu_fp_time, s.fp_model,
                                      s.fp_scheme, st));
        } else {
          ADEPT_TRY(edfdv_exp_f64(fstar[k], out[k], s.batch, s.nx, sp.nv, s.e_out, s.dex, s.pond, sp.charge, sp.mass,
                                  s.dt, sp.k1v, st));
        }
      }
      collided = true;
    } else {
      ADEPT_TRY(c.push_v(fstar, out, s.dt, 0));
    }
  } else {
    // sixth-order Hamiltonian splitting (vector_field.py:118-186)
    const double dt = s.dt;
    const double a1 = 0.168735950563437422448196, a2 = 0.377851589220928303880766, a3 = -0.093175079568731452657924;
    const double b1 = 0.049086460976116245491441, b2 = 0.264177609888976700200146, b3 = 0.186735929134907054308413;
    const double c1 = -0.000069728715055305084099, c2 = -0.000625704827430047189169, c3 = -0.002213085124045325561636;
    const double d2 = -2.916600457689847816445691e-6, d3 = 3.048480261700038788680723e-5;
    const double e3 = 4.985549387875068121593988e-7;
    const double D1 = b1 + 2.0 * c1 * pow(dt, 2.0);
    const double D2 = b2 + 2.0 * c2 * pow(dt, 2.0) + 4.0 * d2 * pow(dt, 4.0);
    const double D3 = b3 + 2.0 * c3 * pow(dt, 2.0) + 4.0 * d3 * pow(dt, 4.0) - 8.0 * e3 * pow(dt, 6.0);
    const double Ds[6] = {D1, D2, D3, D3, D2, D1};
    const double As[5] = {a1, a2, a3, a2, a1};
    bool from_parts = false;
    for (int i = 0; i < 6; i++) {
      ADEPT_TRY(c.field_solve(cur, from_parts, dt));
      // destination of this v-push: exponential runs in place after the first hop; the spline alternates tmp / out
      // starting with tmp so that the sixth hop lands in f_out
      double* dst[ADEPT_B200_MAX_SPECIES];
      for (int k = 0; k < s.n_species; k++) {
        if (spline)
          dst[k] = (cur[k] == tmp[k]) ? out[k] : tmp[k];
        else
          dst[k] = out[k];
      }
      // the last v-push shares its pass over f with the collision step when shape and operator allow it
      const int kc = s.collide_species;
      if (i == 5 && s.fp_on && !s.krook_on && !spline && s.fp_sc_steps == 0 && !want_diag && kc >= 0 &&
          kc < s.n_species && vpush_collide_supported(s.nx, s.species[kc].nv, s.fp_model, s.fp_scheme, s.fp_nodrag)) {
        const double* dexi = s.dex + (long long)i * c.n;
        for (int k = 0; k < s.n_species; k++) {
          const adept_b200_species& sp = s.species[k];
          if (k == kc) {
            ADEPT_TRY(vpush_collide_f64(cur[k], dst[k], s.batch, s.nx, sp.nv, s.e_out, dexi, s.pond, sp.charge, sp.mass,
                                        Ds[i] * dt, sp.k1v, sp.v, sp.dv, s.nu_fp_space, s.nu_fp_time, s.fp_model,
                                        s.fp_scheme, st, nullptr, nullptr, 0, 0, s.dt));
          } else {
            ADEPT_TRY(edfdv_exp_f64(cur[k], dst[k], s.batch, s.nx, sp.nv, s.e_out, dexi, s.pond, sp.charge, sp.mass,
                                    Ds[i] * dt, sp.k1v, st));
          }
        }
        collided = true;
      } else {
        ADEPT_TRY(c.push_v(cur, dst, Ds[i] * dt, i));
      }
      for (int k = 0; k < s.n_species; k++) cur[k] = dst[k];
      if (i < 5) {
        double* same[ADEPT_B200_MAX_SPECIES];
        for (int k = 0; k < s.n_species; k++) same[k] = const_cast<double*>(cur[k]);
        ADEPT_TRY(c.push_x(cur, same, As[i] * dt, true));  // in place; its density feeds the next field solve
        from_parts = true;
      }
    }
  }

  // dfdt diagnostics of the reference species (vector_field.py:245-250) around the collision step
  const adept_b200_species* dsp = want_diag ? &s.species[s.diag_species] : nullptr;
  const long long n_diag = want_diag ? c.n * dsp->nv : 0;
  if (s.diag_vlasov_dfdt) ADEPT_TRY(diff_over_dt_f64(dsp->f_out, dsp->f_in, s.dt, s.diag_vlasov_dfdt, n_diag, st));
  if (s.diag_fp_dfdt) {  // keep f_vlasov in the output buffer until f_fp exists
    cudaError_t err = cudaMemcpyAsync(s.diag_fp_dfdt, dsp->f_out, (size_t)n_diag * sizeof(double),
                                      cudaMemcpyDeviceToDevice, st);
    if (err != cudaSuccess) {
      set_last_error("step: cudaMemcpyAsync(diag-fp-dfdt): %s", cudaGetErrorString(err));
      return ADEPT_ERR_CUDA;
    }
  }

  // collisions on the reference species, in place (vector_field.py:238)
  if ((s.fp_on || s.krook_on) && !collided) {
    const adept_b200_species& sp = s.species[s.collide_species];
    ADEPT_TRY(collide_f64(sp.f_out, sp.f_out, s.batch, s.nx, sp.nv, sp.v, sp.dv, s.dt, s.fp_on ? s.nu_fp_space : nullptr,
                          s.krook_on ? s.nu_K_space : nullptr, s.f_mx, s.fp_model, s.fp_scheme, s.fp_nodrag, s.sg_m,
                          s.sg_ratio, nullptr, s.nu_fp_time, s.nu_K_time, st, s.fp_sc_steps, s.fp_sc_rtol,
                          s.fp_sc_atol));
  }

  // Hou-Li filter on every species (vector_field.py:240-241): the x-advection kernels with dt = 0 (unit phases) and
  // the real per-mode multiplier, in place
  if (s.hou_li_filt) {
    for (int k = 0; k < s.n_species; k++) {
      const adept_b200_species& sp = s.species[k];
      if (vdfdx_tma_supported(sp.f_out, sp.f_out, s.nx, sp.nv))
        ADEPT_TRY(vdfdx_tma_f64(sp.f_out, sp.f_out, s.batch, s.nx, sp.nv, sp.v, 0.0, nullptr, 0.0, nullptr, st,
                                s.hou_li_filt));
      else
        ADEPT_TRY(vdfdx_f64(sp.f_out, sp.f_out, s.batch, s.nx, sp.nv, sp.v, 0.0, nullptr, 0.0, st, s.hou_li_filt));
    }
  }
  if (s.diag_fp_dfdt) ADEPT_TRY(diff_over_dt_f64(dsp->f_out, s.diag_fp_dfdt, s.dt, s.diag_fp_dfdt, n_diag, st));

  if (wave) {  // vector_field.py:340-347
    if (wave_density) ADEPT_TRY(c.electron_density(s.species[s.electron_species].f_out, s.ne_np1));
    ADEPT_TRY(wave_step_f64(s.a, s.prev_a, s.djy, wave_density ? s.ne_n : nullptr, wave_density ? s.ne_np1 : nullptr,
                            s.a_out, s.batch, s.nx, s.c_light, s.dx, s.dt, st));
  }
  return ADEPT_OK;
}

}  // namespace adept

namespace adept {

// e_bar = accel_bar * q/m (+ e_out_bar); dex_bar = accel_bar * q/m
__global__ void accel_chain_kernel(const double* __restrict__ abar, const double* __restrict__ e_out_bar, double qm,
                                   double* __restrict__ e_bar, double* __restrict__ dex_bar, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double a = abar[i] * qm;
  if (dex_bar) dex_bar[i] = a;
  e_bar[i] = e_out_bar ? a + e_out_bar[i] : a;
}
__global__ void scale_rows_kernel(double* __restrict__ x, double s, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] *= s;
}

int step_bwd_f64(const adept_b200_step& s, const adept_b200_step_bwd& b, cudaStream_t st) {
  ADEPT_TRY(validate(s));
  if (s.n_species != 1 || s.time_integrator != 0 || s.field != 0 || s.wave_on || s.fp_sc_steps != 0 ||
      s.hou_li_filt || s.ex_w_row || s.ex_a0_row || (s.fp_on && (s.fp_model == 2 || s.fp_nodrag))) {
    set_last_error("step_bwd: only one species, leapfrog, poisson, LB / Dougherty (+ Krook), wave off are implemented");
    return ADEPT_ERR_UNSUPPORTED;
  }
  if (!b.f_out_bar || !b.f_in_bar || !b.scratch_f[0] || !b.scratch_f[1] || !b.scratch_f[2] || !b.scratch_row[0] ||
      !b.scratch_row[1]) {
    set_last_error("step_bwd: null cotangent / scratch pointer");
    return ADEPT_ERR_BAD_ARG;
  }
  TimeRowScope time_row_scope(nullptr);
  const adept_b200_species& sp = s.species[0];
  const long long n = (long long)s.batch * s.nx;
  const int nv = sp.nv;
  double* fs = b.scratch_f[0];   // f* = vdfdx(f)
  double* f2 = b.scratch_f[1];   // f** = edfdv(f*)
  double* f3 = b.scratch_f[2];   // Fokker-Planck output (before Krook) / cotangent scratch
  const bool spline = s.edfdv == 1;
  // ---- recompute the intermediates of the forward step ----------------------------------------------------------
  if (vdfdx_tma_supported(sp.f_in, fs, s.nx, nv))
    ADEPT_TRY(vdfdx_tma_f64(sp.f_in, fs, s.batch, s.nx, nv, sp.v, s.dt, s.k1x_batch, s.k1x, nullptr, st));
  else
    ADEPT_TRY(vdfdx_f64(sp.f_in, fs, s.batch, s.nx, nv, sp.v, s.dt, s.k1x_batch, s.k1x, st));
  if (spline)
    ADEPT_TRY(edfdv_spline_f64(fs, f2, s.batch, s.nx, nv, s.e_out, s.dex, s.pond, sp.charge, sp.mass, s.dt, sp.dv, st));
  else
    ADEPT_TRY(edfdv_exp_f64(fs, f2, s.batch, s.nx, nv, s.e_out, s.dex, s.pond, sp.charge, sp.mass, s.dt, sp.k1v, st));
  // ---- collisions, reversed: Krook then Fokker-Planck; g lives in f_in_bar from here on --------------------------
  const double* g = b.f_out_bar;
  double* gbuf = b.f_in_bar;
  if (s.krook_on) {
    const double* f_pre_krook = f2;
    if (s.fp_on) {
      ADEPT_TRY(collide_f64(f2, f3, s.batch, s.nx, nv, sp.v, sp.dv, s.dt, s.nu_fp_space, nullptr, nullptr, s.fp_model,
                            s.fp_scheme, 0, s.sg_m, s.sg_ratio, nullptr, s.nu_fp_time, 1.0, st));
      f_pre_krook = f3;
    }
    // nu_K(x, t) = time * space: the kernel takes the product, the cotangent is w.r.t. the product
    ADEPT_TRY(axpy_f64(s.nu_K_space, s.nu_K_space, s.nu_K_time - 1.0, b.scratch_row[0], n, st));  // time * space
    ADEPT_TRY(krook_bwd_f64(f_pre_krook, g, s.batch, s.nx, nv, sp.dv, s.dt, b.scratch_row[0], s.f_mx, gbuf, b.nu_K_bar,
                            st));
    g = gbuf;
  }
  if (s.fp_on) {
    if (!s.krook_on)  // f3 = FP(f2) is the forward output itself, but f_out may have been overwritten: recompute
      ADEPT_TRY(collide_f64(f2, f3, s.batch, s.nx, nv, sp.v, sp.dv, s.dt, s.nu_fp_space, nullptr, nullptr, s.fp_model,
                            s.fp_scheme, 0, s.sg_m, s.sg_ratio, nullptr, s.nu_fp_time, 1.0, st));
    // the cotangent of f2 may land in f3's buffer: collide_bwd loads a row of f_new into shared memory before it writes
    // the same row of f_bar, and rows are independent
    double* g2 = (g == gbuf) ? f3 : gbuf;
    ADEPT_TRY(collide_bwd_f64(f2, f3, g, g2, b.nu_fp_bar, s.batch, s.nx, nv, sp.v, sp.dv, s.dt, s.nu_fp_space,
                              s.nu_fp_time, s.fp_model, s.fp_scheme, st));
    if (b.nu_fp_bar) {  // collide_bwd returns d/d(space profile) = time * d/d(nu): undo the scale
      if (s.nu_fp_time != 0.0) {
        ProfileScope prof("scale_rows", st);
        scale_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(b.nu_fp_bar, 1.0 / s.nu_fp_time, n);
        ADEPT_TRY(check_launch("scale_rows_kernel"));
      }
    }
    g = g2;
  }
  // ---- v-advection, reversed: cotangent of f* into f2's buffer, cotangent of the acceleration --------------------
  double* abar = b.scratch_row[0];
  double* ebar = b.scratch_row[1];
  double* gfs = f2;
  if (spline) {
    ADEPT_TRY(edfdv_spline_bwd_f64(fs, g, s.batch, s.nx, nv, s.e_out, s.dex, s.pond, sp.charge, sp.mass, s.dt, sp.dv,
                                   gfs, abar, st));
  } else {
    ADEPT_TRY(edfdv_exp_bwd_accel_f64(fs, g, s.batch, s.nx, nv, s.e_out, s.dex, s.pond, sp.charge, sp.mass, s.dt,
                                      sp.k1v, abar, st));
    ADEPT_TRY(edfdv_exp_f64(g, gfs, s.batch, s.nx, nv, s.e_out, s.dex, s.pond, sp.charge, sp.mass, -s.dt, sp.k1v, st));
  }
  {
    ProfileScope prof("accel_chain", st);
    accel_chain_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(abar, b.e_out_bar, sp.charge / sp.mass, ebar,
                                                                    b.dex_bar, n);
    ADEPT_TRY(check_launch("accel_chain_kernel"));
  }
  // ---- field solve, reversed: rho_bar = -poisson(e_bar) (antisymmetric), f*_bar += q dv rho_bar -------------------
  double* rhobar = b.scratch_row[0];
  ADEPT_TRY(poisson_dispatch_f64(ebar, s.kmul, s.kmul_stride, rhobar, s.batch, s.nx, 0, 0.0, 0.0, st));
  {
    const double* obar[3] = {rhobar, nullptr, nullptr};
    const double coef[3] = {-sp.dv * sp.charge, 0.0, 0.0};
    ADEPT_TRY(moments_bwd_f64(obar, coef, s.batch, s.nx, nv, nullptr, 1, gfs, st));
  }
  // ---- x-advection, reversed --------------------------------------------------------------------------------------
  if (vdfdx_tma_supported(gfs, b.f_in_bar, s.nx, nv))
    return vdfdx_tma_f64(gfs, b.f_in_bar, s.batch, s.nx, nv, sp.v, -s.dt, s.k1x_batch, s.k1x, nullptr, st);
  return vdfdx_f64(gfs, b.f_in_bar, s.batch, s.nx, nv, sp.v, -s.dt, s.k1x_batch, s.k1x, st);
}

__global__ void time_row_advance_kernel(const double* __restrict__ table, long long n_rows, long long* counter,
                                        double* __restrict__ row) {
  const long long i = *counter;
  const long long r = i < n_rows ? (i < 0 ? 0 : i) : n_rows - 1;
  row[threadIdx.x] = table[r * TROW_LEN + threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) *counter = i + 1;
}
}  // namespace adept

extern "C" int adept_b200_ex_driver_f64(const double* ex_space, const double* ex_kx, int n_ex, const double* w_host,
                                        const double* a0_host, const double* tenv_host, const double* wt_host,
                                        long long n, double* dex, void* stream) {
  using namespace adept;
  if (!dex || n < 1 || n_ex < 0 || n_ex > ADEPT_B200_MAX_DRIVERS || (n_ex > 0 && (!ex_space || !ex_kx || !w_host ||
                                                                                 !a0_host || !tenv_host || !wt_host))) {
    set_last_error("adept_b200_ex_driver_f64: bad arguments (n=%lld, n_ex=%d)", n, n_ex);
    return ADEPT_ERR_BAD_ARG;
  }
  DriverArgs d = {};
  d.n_ex = n_ex, d.n_sub = 1, d.n = n, d.space = ex_space, d.kx = ex_kx, d.dex = dex;
  for (int k = 0; k < n_ex; k++) d.w[k] = w_host[k], d.a0[k] = a0_host[k], d.tenv[0][k] = tenv_host[k], d.wt[0][k] = wt_host[k];
  ProfileScope prof("ex_driver", (cudaStream_t)stream);
  ex_driver_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d);
  return check_launch("ex_driver_kernel");
}

extern "C" int adept_b200_time_row_advance(const double* table, long long n_rows, long long* counter, double* row,
                                           void* stream) {
  if (!table || !counter || !row || n_rows < 1) {
    adept::set_last_error("adept_b200_time_row_advance: null pointer or empty table");
    return adept::ADEPT_ERR_BAD_ARG;
  }
  static_assert(adept::TROW_LEN == ADEPT_B200_TIME_ROW_LEN, "time row layout");
  adept::ProfileScope prof("time_row_advance", (cudaStream_t)stream);
  adept::time_row_advance_kernel<<<1, adept::TROW_LEN, 0, (cudaStream_t)stream>>>(table, n_rows, counter, row);
  return adept::check_launch("time_row_advance_kernel");
}

extern "C" int adept_b200_step_bwd_f64(const adept_b200_step* step, const adept_b200_step_bwd* bwd, void* stream) {
  if (!step || !bwd) {
    adept::set_last_error("adept_b200_step_bwd_f64: null descriptor");
    return adept::ADEPT_ERR_BAD_ARG;
  }
  return adept::step_bwd_f64(*step, *bwd, (cudaStream_t)stream);
}

extern "C" int adept_b200_step_f64(const adept_b200_step* step, void* stream) {
  if (!step) {
    adept::set_last_error("adept_b200_step_f64: null step descriptor");
    return adept::ADEPT_ERR_BAD_ARG;
  }
  return adept::step_f64(*step, (cudaStream_t)stream);
}
