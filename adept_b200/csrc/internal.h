// Internal C++ interface between the translation units of libadept_b200.so (kernel launchers).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace adept {

int vdfdx_f64(const double* fin, double* fout, int batch, int nx, int nv, const double* v, double dt,
              const double* k1_batch, double k1, cudaStream_t stream, const double* filt = nullptr);
int edfdv_exp_f64(const double* fin, double* fout, int batch, int nx, int nv, const double* e, const double* dex,
                  const double* pond, double q, double m, double dt, double k1, cudaStream_t stream);
int edfdv_spline_f64(const double* fin, double* fout, int batch, int nx, int nv, const double* e, const double* dex,
                     const double* pond, double q, double m, double dt, double dv, cudaStream_t stream);
int moments_f64(const double* f, int batch, int nx, int nv, const double* v, double scale_a, const double* const* base,
                double* const* out, const double* scale_b, cudaStream_t stream);
int axpy_f64(const double* a, const double* b, double s, double* out, long long n, cudaStream_t stream);
int poisson_dispatch_f64(const double* rho, const double* kmul, long long kmul_stride, double* e, int batch, int nx,
                         int mode, double Te, double lambda_De, cudaStream_t stream);
bool bluestein_supported(int n);
int bluestein_push_f64(int axis, const double* fin, double* fout, int batch, int nx, int nv, const double* v, double dt,
                       const double* k1_batch, double k1, const double* e, const double* dex, const double* pond,
                       double q, double m, const double* filt, cudaStream_t stream);
int bluestein_poisson_f64(const double* rho, const double* kmul, long long kmul_stride, double* e, int batch, int nx,
                          int mode, double Te, double lambda_De, cudaStream_t stream);
int poisson_green_f64(const double* rho, const double* green, long long green_stride, double* e, int batch, int nx,
                      cudaStream_t stream);
bool field_member_supported(int nx);
int field_member_f64(int nsp, const double* const* f, const int* nv, const double* dv, const double* charge,
                     const double* base, double* rho, int batch, int nx, const double* a, double* pond, double dx,
                     int n_ex, const double* ex_space, const double* ex_kx, double* dex, const double* ex_w,
                     const double* ex_a0, const double* ex_tenv, const double* ex_wt, const double* ex_w_row,
                     const double* ex_a0_row, double ex_t0, const double* kmul, long long kmul_stride, double* e,
                     int mode, double Te, double lambda_De, cudaStream_t stream);
bool field_fused_supported(int batch, int nx);
int field_fused_f64(int nsp, const double* const* parts, const int* nparts, const double* dv, const double* charge,
                    const double* base, double* rho, int nx, const double* a, double* pond, double dx, int n_ex,
                    const double* ex_space, const double* ex_kx, double* dex, const double* ex_w, const double* ex_a0,
                    const double* ex_tenv, const double* ex_wt, const double* kmul, double* e, int mode, double Te,
                    double lambda_De, unsigned int* counter, cudaStream_t stream);
int ponderomotive_f64(const double* a, double* pond, int batch, int nx, double dx, cudaStream_t stream);
int wave_step_f64(const double* a, const double* aold, const double* djy, const double* ne_n, const double* ne_np1,
                  double* a_new, int batch, int nx, double c, double dx, double dt, cudaStream_t stream);
int collide_f64(const double* fin, double* fout, int batch, int nx, int nv, const double* v, double dv, double dt,
                const double* nu_fp, const double* nu_K, const double* f_mx, int model, int scheme, int nodrag,
                double sg_m, double sg_ratio, double* n_out, double nu_fp_scale, double nu_K_scale,
                cudaStream_t stream, int sc_steps = 0, double sc_rtol = 1e-8, double sc_atol = 1e-12,
                const double* coef_in = nullptr, double* coef_out = nullptr, int coef_div = 1);
int diff_over_dt_f64(const double* a, const double* b, double dt, double* out, long long n, cudaStream_t stream);
int row_means_f64(const double* a, int rows, long long n, double* out, cudaStream_t stream);
int field_energy_f64(const double* e0, const double* de0, const double* e1, const double* de1, double w, int batch,
                     int nx, double* out, cudaStream_t stream);
int reduce_parts_f64(const double* parts, int nparts, long long n, double scale_a, double scale_b, const double* base,
                     double* out, cudaStream_t stream);
int save_moments_f64(const double* f0, const double* f1, double w, int batch, int nx, int nv, const double* v,
                     double dv, double* out, cudaStream_t stream);
int interp2d_f64(const double* f0, const double* f1, double w, int nx, int nv, const double* x, const double* v,
                 const double* xq, const double* vq, int nxq, int nvq, double* out, cudaStream_t stream);
int edfdv_exp_bwd_accel_f64(const double* f, const double* g, int batch, int nx, int nv, const double* e,
                            const double* dex, const double* pond, double q, double m, double dt, double k1,
                            double* abar, cudaStream_t stream);
int moments_bwd_f64(const double* const* obar, const double* coef, int batch, int nx, int nv, const double* v,
                    int accumulate, double* fbar, cudaStream_t stream);
int collide_bwd_f64(const double* fin, const double* fnew, const double* g, double* fbar, double* nubar, int batch,
                    int nx, int nv, const double* v, double dv, double dt, const double* nu_fp, double nu_fp_scale,
                    int model, int scheme, cudaStream_t stream);
int edfdv_spline_bwd_f64(const double* f, const double* g, int batch, int nx, int nv, const double* e, const double* dex,
                         const double* pond, double q, double m, double dt, double dv, double* fbar, double* abar,
                         cudaStream_t stream);
int krook_bwd_f64(const double* f, const double* g, int batch, int nx, int nv, double dv, double dt, const double* nu_K,
                  const double* f_mx, double* fbar, double* nubar, cudaStream_t stream);
// bigx.cu: x-direction spectral operators for long pencils of mixed length (nx = 2^a m, e.g. 17280 = 128 x 135)
bool bigx_supported(int nx, int nv);
int bigx_apply_f64(const double* in, double* out, double* scratch, int batch, int nx, int nv, const double* v, double dt,
                   const double* k1_batch, double k1, const double2* mtab, long long mtab_stride, cudaStream_t stream);
int bigx_poisson_f64(const double* rho, const double* kmul, long long kmul_stride, double* e, int batch, int nx,
                     int mode, double Te, double lambda_De, cudaStream_t stream);
int sum_peers_f64(const double* const* peers, int n_peers, long long n, double* out, cudaStream_t stream);
int marginal_f64(const double* f, const double* w, long long rows, int np, double* out, cudaStream_t stream);
int transpose_f64(const double* in, double* out, int batch, int n0, int n1, cudaStream_t stream);
int abs_rfft_x_f64(const double* fin, double* fout, int batch, int nx, int nv, cudaStream_t stream);
bool vpush_collide_supported(int nx, int nv, int model, int scheme, int nodrag);
int vpush_collide_f64(const double* fin, double* fout, int batch, int nx, int nv, const double* e, const double* dex,
                      const double* pond, double q, double m, double dt, double k1v, const double* v, double dv,
                      const double* nu_fp, double nu_fp_scale, int model, int scheme, cudaStream_t stream,
                      const double* const* in_peers = nullptr, double* const* out_peers = nullptr, int n_peers = 0,
                      long long row0_global = 0, double dt_fp = 0.0 /* collision time step; 0: the same as dt */,
                      double* stage = nullptr, unsigned int* round_ctr = nullptr, int n_movers = 0,
                      long long nx_global = 0);
bool tma_available();
int encode_map_2d(CUtensorMap* map, const double* base, unsigned long long dim0, unsigned long long dim1,
                  unsigned long long pitch_bytes, unsigned box0, unsigned box1, int swizzle128);
// Field solve folded into the tail of the TMA x-advection (vdfdx_tma.cu): after its last tile every persistent CTA
// joins a device-wide barrier, finishes the charge density of its slice of x from the per-CTA partial rows, evaluates
// the ponderomotive force and the driver there, and after a second barrier solves Poisson's equation for its slice as a
// circular convolution with the Green's function green = Re ifft(-i / kx) (field.py:221-224 is linear in rho).
struct FieldTail {
  unsigned int* counter;  // one zero-initialised word; re-armed by the CTA that exits last
  const double* base;     // static ion background (nullable)
  double dv, charge;
  double* rho;
  double* e;
  const double* green;    // [nx]
  const double* a;        // [nx + 2]
  double* pond;
  double dx;
  int n_ex;               // 0: the driver field is not evaluated here
  const double* ex_space;
  const double* ex_kx;
  double* dex;
  double ex_w[8], ex_a0[8], ex_tenv[8], ex_wt[8];
  const double* trow;     // nullable device-resident time row (common.cuh): replaces ex_tenv / ex_wt (substep 0)
  // Sharded grid (n_peers > 1): this launch only saw the rank's own velocity columns, so the slice sums are the rank's
  // SHARE of the charge density (base = ion / P).  Every CTA pushes its slice of the share into all ranks' inboxes over
  // peer memory, rank `my_rank` raises flag my_rank in every inbox (epoch = step count, monotonic), waits for the P flags
  // of its own inbox and sums the P shares in rank order -- the all-reduce of the reference's sharded charge density
  // (field.py:197-208 under shard_map) without a collective call or a second launch.
  int n_peers, my_rank;
  double* share_in[8];              // rank r's inbox: [2 (epoch parity)][n_peers][nx]
  unsigned long long* flag_in[8];   // rank r's flags: [n_peers]
  unsigned long long epoch;
};
bool vdfdx_tma_supported(const double* fin, const double* fout, int nx, int nv);
int vdfdx_tma_parts(int batch, int nx, int nv);
int vdfdx_tma_f64(const double* fin, double* fout, int batch, int nx, int nv, const double* v, double dt,
                  const double* k1_batch, double k1, double* partial, cudaStream_t stream,
                  const double* filt = nullptr, const FieldTail* field = nullptr);
bool vdfdx_tma_field_supported(int batch, int nx, int nv);
// vdfdx_dual.cu: two transforms per thread (nx = 4096); same contract as vdfdx_tma_f64
bool vdfdx_dual_supported(int nx);
int vdfdx_dual_parts(int batch, int nx, int nv);
int vdfdx_dual_f64(const double* fin, double* fout, int batch, int nx, int nv, const double* v, double dt,
                   const double* k1_batch, double k1, double* partial, cudaStream_t stream, const double* filt,
                   const FieldTail* field);

}  // namespace adept
