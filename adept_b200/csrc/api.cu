// extern "C" boundary of libadept_b200.so (see include/adept_b200.h), twiddle-table cache, error reporting.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/adept_b200.h"
#include "common.cuh"
#include "internal.h"

namespace adept {

static thread_local const double* tl_time_row = nullptr;
const double* current_time_row() { return tl_time_row; }
void set_current_time_row(const double* row) { tl_time_row = row; }


static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);  // every kernel launch site of the library ends here
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_last_error("%s: %s", what, cudaGetErrorString(err));
    return ADEPT_ERR_CUDA;
  }
  return ADEPT_OK;
}

// ---- per-kernel event timing --------------------------------------------------------------------------------------
struct ProfileEntry {
  const char* name;
  cudaEvent_t e0, e1;
};
static std::mutex g_prof_mutex;
static bool g_prof_on = false;
static std::vector<ProfileEntry> g_prof;

ProfileScope::ProfileScope(const char* name, cudaStream_t st) : slot(-1), stream(st) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  ProfileEntry e = {name, nullptr, nullptr};
  if (cudaEventCreate(&e.e0) != cudaSuccess || cudaEventCreate(&e.e1) != cudaSuccess) {
    (void)cudaGetLastError();
    return;
  }
  cudaEventRecord(e.e0, st);
  g_prof.push_back(e);
  slot = (int)g_prof.size() - 1;
}

ProfileScope::~ProfileScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  if (slot < (int)g_prof.size()) cudaEventRecord(g_prof[slot].e1, stream);
}

static void profile_clear() {
  for (auto& e : g_prof) {
    cudaEventDestroy(e.e0);
    cudaEventDestroy(e.e1);
  }
  g_prof.clear();
}

// ---- twiddle tables ------------------------------------------------------------------------------------------
// Layout must match FftCfg<LOGN> in fft_core.cuh: radix-16 passes then one radix 2^(logn%4) pass (n < 16: one pass);
// pass p >= 1 stores tw[off_p + (r-1)*Ns + k] = exp(-2 pi i k r / (Ns R)), k < Ns, r = 1..R-1.
static std::mutex g_tw_mutex;
static cplx* g_tw[64][16] = {};

static std::vector<int> radices_of(int logn) {
  std::vector<int> r;
  const int n = 1 << logn;
  if (n < 16) {
    r.push_back(n);
    return r;
  }
  for (int i = 0; i < logn / 4; i++) r.push_back(16);
  if (logn % 4) r.push_back(1 << (logn % 4));
  return r;
}

const cplx* get_twiddles(int logn) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || logn < 1 || logn > 13) {
    set_last_error("get_twiddles: no CUDA device or unsupported log2(n)=%d", logn);
    (void)cudaGetLastError();
    return nullptr;
  }
  std::lock_guard<std::mutex> lock(g_tw_mutex);
  if (g_tw[dev][logn]) return g_tw[dev][logn];
  const std::vector<int> rad = radices_of(logn);
  std::vector<cplx> host;
  int ns = rad[0];
  const long double two_pi = 6.283185307179586476925286766559005768L;
  for (size_t p = 1; p < rad.size(); p++) {
    const int R = rad[p];
    for (int r = 1; r < R; r++)
      for (int k = 0; k < ns; k++) {
        const long double ang = two_pi * (long double)((long long)k * r) / (long double)((long long)ns * R);
        cplx w;
        w.x = (double)cosl(ang);
        w.y = (double)(-sinl(ang));
        host.push_back(w);
      }
    ns *= R;
  }
  if (host.empty()) host.push_back(make_double2(1.0, 0.0));  // single-pass transforms never read the table
  cplx* d = nullptr;
  cudaError_t err = cudaMalloc(&d, host.size() * sizeof(cplx));
  if (err == cudaSuccess) err = cudaMemcpy(d, host.data(), host.size() * sizeof(cplx), cudaMemcpyHostToDevice);
  if (err != cudaSuccess) {
    set_last_error("get_twiddles(2^%d): %s", logn, cudaGetErrorString(err));
    (void)cudaGetLastError();
    if (d) cudaFree(d);
    return nullptr;
  }
  g_tw[dev][logn] = d;
  return d;
}

// ---- Bluestein (chirp-z) tables for transform lengths that are not powers of two ------------------------------------
// DFT_n(z)[k] = w[k] sum_j (z[j] w[j]) conj(w)[k - j], w[j] = exp(-i pi j^2 / n): a circular convolution of length
// M = 2^logm >= 2n - 1.  chirp[j] = w[j], j < n; bhat = FFT_M(b) / M with b[j] = conj(w[|j|]) wrapped (the 1/M of the
// inverse transform is folded in).  j^2 is reduced mod 2n in integers; everything is evaluated in long double.
struct BluesteinTables {
  int n, logm;
  cplx* chirp;
  cplx* bhat;
};
static std::mutex g_bs_mutex;
static std::vector<BluesteinTables> g_bs[64];

static void host_fft_ld(std::vector<long double>& re, std::vector<long double>& im) {  // radix-2, forward, in place
  const size_t m = re.size();
  for (size_t i = 1, j = 0; i < m; i++) {
    size_t bit = m >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(re[i], re[j]), std::swap(im[i], im[j]);
  }
  const long double pi = 3.141592653589793238462643383279502884L;
  for (size_t len = 2; len <= m; len <<= 1) {
    for (size_t k = 0; k < len / 2; k++) {
      const long double ang = -2.0L * pi * (long double)k / (long double)len;
      const long double wr = cosl(ang), wi = sinl(ang);
      for (size_t i = k; i < m; i += len) {
        const size_t j = i + len / 2;
        const long double tr = re[j] * wr - im[j] * wi, ti = re[j] * wi + im[j] * wr;
        re[j] = re[i] - tr, im[j] = im[i] - ti;
        re[i] += tr, im[i] += ti;
      }
    }
  }
}

int get_bluestein(int n, int* logm_out, const cplx** chirp_out, const cplx** bhat_out) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || n < 2 || n > 4096) {
    set_last_error("bluestein: no CUDA device or unsupported length n=%d (2..4096)", n);
    (void)cudaGetLastError();
    return ADEPT_ERR_UNSUPPORTED;
  }
  std::lock_guard<std::mutex> lock(g_bs_mutex);
  for (const auto& t : g_bs[dev])
    if (t.n == n) {
      *logm_out = t.logm, *chirp_out = t.chirp, *bhat_out = t.bhat;
      return ADEPT_OK;
    }
  int logm = 1;
  while ((1 << logm) < 2 * n - 1) logm++;
  const size_t m = (size_t)1 << logm;
  const long double pi = 3.141592653589793238462643383279502884L;
  std::vector<long double> wr(n), wi(n), br(m, 0.0L), bi(m, 0.0L);
  for (int j = 0; j < n; j++) {
    const long long q = ((long long)j * j) % (2LL * n);
    const long double ang = pi * (long double)q / (long double)n;
    wr[j] = cosl(ang), wi[j] = -sinl(ang);
    br[j] = wr[j], bi[j] = -wi[j];  // conj(w[j])
    if (j) br[m - j] = br[j], bi[m - j] = bi[j];
  }
  host_fft_ld(br, bi);
  std::vector<cplx> hc(n), hb(m);
  for (int j = 0; j < n; j++) hc[j] = make_double2((double)wr[j], (double)wi[j]);
  for (size_t k = 0; k < m; k++) hb[k] = make_double2((double)(br[k] / (long double)m), (double)(bi[k] / (long double)m));
  BluesteinTables t = {n, logm, nullptr, nullptr};
  cudaError_t err = cudaMalloc(&t.chirp, n * sizeof(cplx));
  if (err == cudaSuccess) err = cudaMalloc(&t.bhat, m * sizeof(cplx));
  if (err == cudaSuccess) err = cudaMemcpy(t.chirp, hc.data(), n * sizeof(cplx), cudaMemcpyHostToDevice);
  if (err == cudaSuccess) err = cudaMemcpy(t.bhat, hb.data(), m * sizeof(cplx), cudaMemcpyHostToDevice);
  if (err != cudaSuccess) {
    set_last_error("bluestein(n=%d): %s", n, cudaGetErrorString(err));
    (void)cudaGetLastError();
    return ADEPT_ERR_CUDA;
  }
  g_bs[dev].push_back(t);
  *logm_out = logm, *chirp_out = t.chirp, *bhat_out = t.bhat;
  return ADEPT_OK;
}

}  // namespace adept

using namespace adept;

#define ADEPT_REQUIRE(ptr, name)                           \
  if (!(ptr)) {                                            \
    set_last_error("%s: null pointer for %s", __func__, name); \
    return ADEPT_ERR_BAD_ARG;                              \
  }

extern "C" {

int adept_b200_version(void) { return 100; }

const char* adept_b200_last_error(void) { return g_err; }

int adept_b200_profile(int enable) {
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  profile_clear();
  g_prof_on = enable != 0;
  return ADEPT_OK;
}

int adept_b200_profile_report(char* buf, int buflen) {
  if (!buf || buflen < 1) {
    set_last_error("profile_report: null / empty buffer");
    return ADEPT_ERR_BAD_ARG;
  }
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  struct Acc {
    const char* name;
    int count;
    double ms;
  };
  std::vector<Acc> acc;
  for (auto& e : g_prof) {
    if (cudaEventSynchronize(e.e1) != cudaSuccess) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e.e0, e.e1) != cudaSuccess) continue;
    bool found = false;
    for (auto& a : acc)
      if (a.name == e.name || !strcmp(a.name, e.name)) {
        a.count++, a.ms += ms, found = true;
        break;
      }
    if (!found) acc.push_back({e.name, 1, (double)ms});
  }
  (void)cudaGetLastError();
  int off = 0;
  buf[0] = 0;
  for (auto& a : acc) {
    const int w = snprintf(buf + off, (size_t)(buflen - off), "%s %d %.6f\n", a.name, a.count, a.ms);
    if (w < 0 || w >= buflen - off) break;
    off += w;
  }
  return ADEPT_OK;
}

int adept_b200_prepare(int n) {
  int logn = 0;
  if (n < 2 || (n & (n - 1))) {
    set_last_error("prepare: n=%d is not a power of two", n);
    return ADEPT_ERR_UNSUPPORTED;
  }
  while ((1 << logn) < n) logn++;
  if (logn > 13) {
    set_last_error("prepare: n=%d exceeds 8192", n);
    return ADEPT_ERR_UNSUPPORTED;
  }
  return get_twiddles(logn) ? ADEPT_OK : ADEPT_ERR_CUDA;
}

int adept_b200_vdfdx_f64(const double* f_in, double* f_out, int batch, int nx, int nv, const double* v, double dt,
                         double k1x, const double* k1x_batch, void* stream) {
  ADEPT_REQUIRE(f_in, "f_in") ADEPT_REQUIRE(f_out, "f_out") ADEPT_REQUIRE(v, "v")
  if (batch >= 1 && vdfdx_tma_supported(f_in, f_out, nx, nv))
    return vdfdx_tma_f64(f_in, f_out, batch, nx, nv, v, dt, k1x_batch, k1x, nullptr, (cudaStream_t)stream);
  return vdfdx_f64(f_in, f_out, batch, nx, nv, v, dt, k1x_batch, k1x, (cudaStream_t)stream);
}

int adept_b200_edfdv_exp_bwd_accel_f64(const double* f_in, const double* g, int batch, int nx, int nv, const double* e,
                                       const double* dex, const double* pond, double charge, double mass, double dt,
                                       double k1v, double* accel_bar, void* stream) {
  ADEPT_REQUIRE(f_in, "f_in") ADEPT_REQUIRE(g, "g") ADEPT_REQUIRE(e, "e") ADEPT_REQUIRE(accel_bar, "accel_bar")
  return edfdv_exp_bwd_accel_f64(f_in, g, batch, nx, nv, e, dex, pond, charge, mass, dt, k1v, accel_bar,
                                 (cudaStream_t)stream);
}

int adept_b200_moments_bwd_f64(const double* const* out_bar_host, const double* coef_host, int batch, int nx, int nv,
                               const double* v, int accumulate, double* f_bar, void* stream) {
  ADEPT_REQUIRE(out_bar_host, "out_bar_host") ADEPT_REQUIRE(coef_host, "coef_host") ADEPT_REQUIRE(f_bar, "f_bar")
  return moments_bwd_f64(out_bar_host, coef_host, batch, nx, nv, v, accumulate, f_bar, (cudaStream_t)stream);
}

int adept_b200_collide_bwd_f64(const double* f_in, const double* f_new, const double* g, double* f_bar, double* nu_bar,
                               int batch, int nx, int nv, const double* v, double dv, double dt, const double* nu_fp,
                               int model, int scheme, void* stream) {
  ADEPT_REQUIRE(f_in, "f_in") ADEPT_REQUIRE(f_new, "f_new") ADEPT_REQUIRE(g, "g") ADEPT_REQUIRE(f_bar, "f_bar")
  ADEPT_REQUIRE(v, "v") ADEPT_REQUIRE(nu_fp, "nu_fp")
  return collide_bwd_f64(f_in, f_new, g, f_bar, nu_bar, batch, nx, nv, v, dv, dt, nu_fp, 1.0, model, scheme,
                         (cudaStream_t)stream);
}

int adept_b200_abs_rfft_x_f64(const double* f, double* out, int batch, int nx, int nv, void* stream) {
  ADEPT_REQUIRE(f, "f") ADEPT_REQUIRE(out, "out")
  return abs_rfft_x_f64(f, out, batch, nx, nv, (cudaStream_t)stream);
}

int adept_b200_edfdv_spline_bwd_f64(const double* f_in, const double* g, int batch, int nx, int nv, const double* e,
                                    const double* dex, const double* pond, double charge, double mass, double dt,
                                    double dv, double* f_bar, double* accel_bar, void* stream) {
  ADEPT_REQUIRE(f_in, "f_in") ADEPT_REQUIRE(g, "g") ADEPT_REQUIRE(e, "e")
  return edfdv_spline_bwd_f64(f_in, g, batch, nx, nv, e, dex, pond, charge, mass, dt, dv, f_bar, accel_bar,
                              (cudaStream_t)stream);
}

int adept_b200_krook_bwd_f64(const double* f_in, const double* g, int batch, int nx, int nv, double dv, double dt,
                             const double* nu_K, const double* f_mx, double* f_bar, double* nu_bar, void* stream) {
  ADEPT_REQUIRE(f_in, "f_in") ADEPT_REQUIRE(g, "g") ADEPT_REQUIRE(nu_K, "nu_K") ADEPT_REQUIRE(f_mx, "f_mx")
  return krook_bwd_f64(f_in, g, batch, nx, nv, dv, dt, nu_K, f_mx, f_bar, nu_bar, (cudaStream_t)stream);
}

int adept_b200_vpush_collide_p2p_f64(const double* const* in_peers_host, double* const* out_peers_host, int n_peers,
                                     long long row0_global, int nx, int nv, const double* e, const double* dex,
                                     const double* pond, double charge, double mass, double dt, double k1v,
                                     const double* v, double dv, const double* nu_fp, int model, int scheme,
                                     void* stream) {
  ADEPT_REQUIRE(in_peers_host, "in_peers") ADEPT_REQUIRE(out_peers_host, "out_peers") ADEPT_REQUIRE(e, "e")
  ADEPT_REQUIRE(v, "v") ADEPT_REQUIRE(nu_fp, "nu_fp")
  for (int j = 0; j < n_peers && j < 8; j++) {
    ADEPT_REQUIRE(in_peers_host[j], "in_peers[j]") ADEPT_REQUIRE(out_peers_host[j], "out_peers[j]")
  }
  // f_in / f_out are unused in peer mode; pass the local slots so that argument checks see non-null pointers
  return vpush_collide_f64(in_peers_host[0], out_peers_host[0], 1, nx, nv, e, dex, pond, charge, mass, dt, k1v, v, dv,
                           nu_fp, 1.0, model, scheme, (cudaStream_t)stream, in_peers_host, out_peers_host, n_peers,
                           row0_global);
}

int adept_b200_vpush_collide_p2p_staged_f64(const double* const* in_peers_host, double* const* out_peers_host,
                                            int n_peers, long long row0_global, int nx, int nv, const double* e,
                                            const double* dex, const double* pond, double charge, double mass,
                                            double dt, double k1v, const double* v, double dv, const double* nu_fp,
                                            int model, int scheme, double* stage, unsigned int* round_counters,
                                            int n_movers, long long nx_global, void* stream) {
  ADEPT_REQUIRE(in_peers_host, "in_peers") ADEPT_REQUIRE(out_peers_host, "out_peers") ADEPT_REQUIRE(e, "e")
  ADEPT_REQUIRE(v, "v") ADEPT_REQUIRE(nu_fp, "nu_fp") ADEPT_REQUIRE(stage, "stage")
  if (n_movers != 0) ADEPT_REQUIRE(round_counters, "round_counters")
  for (int j = 0; j < n_peers && j < 8; j++) {
    ADEPT_REQUIRE(in_peers_host[j], "in_peers[j]") ADEPT_REQUIRE(out_peers_host[j], "out_peers[j]")
  }
  return vpush_collide_f64(in_peers_host[0], out_peers_host[0], 1, nx, nv, e, dex, pond, charge, mass, dt, k1v, v, dv,
                           nu_fp, 1.0, model, scheme, (cudaStream_t)stream, in_peers_host, out_peers_host, n_peers,
                           row0_global, 0.0, stage, round_counters, n_movers, nx_global);
}

int adept_b200_copy2d_f64(double* dst, long long dst_pitch, const double* src, long long src_pitch, long long width,
                          long long height, void* stream) {
  ADEPT_REQUIRE(dst, "dst") ADEPT_REQUIRE(src, "src")
  if (width < 1 || height < 1 || dst_pitch < width || src_pitch < width) {
    set_last_error("copy2d: width=%lld height=%lld dst_pitch=%lld src_pitch=%lld", width, height, dst_pitch, src_pitch);
    return ADEPT_ERR_BAD_SHAPE;
  }
  cudaError_t err = cudaMemcpy2DAsync(dst, (size_t)dst_pitch * sizeof(double), src, (size_t)src_pitch * sizeof(double),
                                      (size_t)width * sizeof(double), (size_t)height, cudaMemcpyDeviceToDevice,
                                      (cudaStream_t)stream);
  if (err != cudaSuccess) {
    set_last_error("copy2d: cudaMemcpy2DAsync: %s", cudaGetErrorString(err));
    (void)cudaGetLastError();
    return ADEPT_ERR_CUDA;
  }
  return ADEPT_OK;
}

int adept_b200_interp2d_f64(const double* f0, const double* f1, double w, int nx, int nv, const double* x,
                            const double* v, const double* xq, const double* vq, int nxq, int nvq, double* out,
                            void* stream) {
  ADEPT_REQUIRE(f0, "f0") ADEPT_REQUIRE(x, "x") ADEPT_REQUIRE(v, "v") ADEPT_REQUIRE(xq, "xq") ADEPT_REQUIRE(vq, "vq")
  ADEPT_REQUIRE(out, "out")
  return interp2d_f64(f0, f1, w, nx, nv, x, v, xq, vq, nxq, nvq, out, (cudaStream_t)stream);
}

long long adept_b200_launch_count(void) { return adept::g_launches.load(std::memory_order_relaxed); }

int adept_b200_vpush_collide_f64(const double* f_in, double* f_out, int batch, int nx, int nv, const double* e,
                                 const double* dex, const double* pond, double charge, double mass, double dt,
                                 double k1v, const double* v, double dv, const double* nu_fp, int model, int scheme,
                                 void* stream) {
  ADEPT_REQUIRE(f_in, "f_in") ADEPT_REQUIRE(f_out, "f_out") ADEPT_REQUIRE(e, "e") ADEPT_REQUIRE(v, "v")
  ADEPT_REQUIRE(nu_fp, "nu_fp")
  return vpush_collide_f64(f_in, f_out, batch, nx, nv, e, dex, pond, charge, mass, dt, k1v, v, dv, nu_fp, 1.0, model,
                           scheme, (cudaStream_t)stream);
}

int adept_b200_save_moments_f64(const double* f0, const double* f1, double w, int batch, int nx, int nv,
                                const double* v, double dv, double* out, void* stream) {
  ADEPT_REQUIRE(f0, "f0") ADEPT_REQUIRE(v, "v") ADEPT_REQUIRE(out, "out")
  return save_moments_f64(f0, f1, w, batch, nx, nv, v, dv, out, (cudaStream_t)stream);
}

int adept_b200_filter_x_f64(const double* f_in, double* f_out, int batch, int nx, int nv, const double* filt,
                            const double* zeros_v, void* stream) {
  ADEPT_REQUIRE(f_in, "f_in") ADEPT_REQUIRE(f_out, "f_out") ADEPT_REQUIRE(filt, "filt") ADEPT_REQUIRE(zeros_v, "zeros_v")
  // the x-advection kernels with zero advection speed (phase 1) and the real per-mode multiplier
  if (batch >= 1 && vdfdx_tma_supported(f_in, f_out, nx, nv))
    return vdfdx_tma_f64(f_in, f_out, batch, nx, nv, zeros_v, 0.0, nullptr, 0.0, nullptr, (cudaStream_t)stream, filt);
  return vdfdx_f64(f_in, f_out, batch, nx, nv, zeros_v, 0.0, nullptr, 0.0, (cudaStream_t)stream, filt);
}

int adept_b200_vdfdx_rho_parts(int batch, int nx, int nv) {
  if (batch < 1 || nx < 2 || nv < 2) return 1;
  // alignment is checked again at call time; an unaligned buffer falls back to one part
  const int parts = vdfdx_tma_supported(nullptr, nullptr, nx, nv) ? vdfdx_tma_parts(batch, nx, nv) : 1;
  return parts < 1 ? 1 : parts;
}

int adept_b200_vdfdx_rho_f64(const double* f_in, double* f_out, int batch, int nx, int nv, const double* v, double dt,
                             double k1x, const double* k1x_batch, double* parts, int nparts, void* stream) {
  ADEPT_REQUIRE(f_in, "f_in") ADEPT_REQUIRE(f_out, "f_out") ADEPT_REQUIRE(v, "v") ADEPT_REQUIRE(parts, "parts")
  cudaStream_t st = (cudaStream_t)stream;
  if (batch < 1 || nx < 2 || nv < 2) {
    set_last_error("vdfdx_rho: bad shape batch=%d nx=%d nv=%d", batch, nx, nv);
    return ADEPT_ERR_BAD_SHAPE;
  }
  const long long n = (long long)batch * nx;
  if (nparts < adept_b200_vdfdx_rho_parts(batch, nx, nv)) {
    set_last_error("vdfdx_rho: parts buffer has %d rows, adept_b200_vdfdx_rho_parts() asks for %d", nparts,
                   adept_b200_vdfdx_rho_parts(batch, nx, nv));
    return ADEPT_ERR_BAD_ARG;
  }
  cudaError_t err = cudaMemsetAsync(parts, 0, (size_t)nparts * n * sizeof(double), st);
  if (err != cudaSuccess) {
    set_last_error("vdfdx_rho: cudaMemsetAsync: %s", cudaGetErrorString(err));
    return ADEPT_ERR_CUDA;
  }
  if (vdfdx_tma_supported(f_in, f_out, nx, nv))
    return vdfdx_tma_f64(f_in, f_out, batch, nx, nv, v, dt, k1x_batch, k1x, parts, st);
  // small or odd shapes: direct kernel, then one plain velocity sum into part 0
  int rc = vdfdx_f64(f_in, f_out, batch, nx, nv, v, dt, k1x_batch, k1x, st);
  if (rc != ADEPT_OK) return rc;
  double* outs[3] = {parts, nullptr, nullptr};
  return moments_f64(f_out, batch, nx, nv, nullptr, 1.0, nullptr, outs, nullptr, st);
}

int adept_b200_vdfdx_field_peers_f64(const double* f_in, double* f_out, int nx, int nv_local, const double* v_local,
                                     double dt, double k1x, double* parts, int nparts,
                                     const adept_b200_field_peers* fp, void* stream) {
  ADEPT_REQUIRE(f_in, "f_in") ADEPT_REQUIRE(f_out, "f_out") ADEPT_REQUIRE(v_local, "v_local")
  ADEPT_REQUIRE(parts, "parts") ADEPT_REQUIRE(fp, "fp")
  if (fp->n_peers < 2 || fp->n_peers > 8 || fp->my_rank < 0 || fp->my_rank >= fp->n_peers || fp->epoch < 1 ||
      fp->n_ex < 0 || fp->n_ex > 8) {
    set_last_error("vdfdx_field_peers: n_peers=%d my_rank=%d epoch=%llu n_ex=%d", fp->n_peers, fp->my_rank, fp->epoch,
                   fp->n_ex);
    return ADEPT_ERR_BAD_ARG;
  }
  for (int r = 0; r < fp->n_peers; r++) {
    ADEPT_REQUIRE(fp->share_in[r], "share_in[r]") ADEPT_REQUIRE(fp->flag_in[r], "flag_in[r]")
  }
  if (!vdfdx_tma_supported(f_in, f_out, nx, nv_local) || !vdfdx_tma_field_supported(1, nx, nv_local)) {
    set_last_error("vdfdx_field_peers: unsupported shape nx=%d nv_local=%d (or the grid cannot be co-resident)", nx,
                   nv_local);
    return ADEPT_ERR_UNSUPPORTED;
  }
  if (nparts < vdfdx_tma_parts(1, nx, nv_local)) {
    set_last_error("vdfdx_field_peers: parts buffer has %d rows, needs %d", nparts, vdfdx_tma_parts(1, nx, nv_local));
    return ADEPT_ERR_BAD_ARG;
  }
  FieldTail ft = {};
  ft.counter = fp->sync_counter, ft.base = fp->ion_share, ft.dv = fp->dv, ft.charge = fp->charge;
  ft.rho = fp->rho, ft.e = fp->e, ft.green = fp->green, ft.a = fp->a_zero, ft.pond = fp->pond, ft.dx = fp->dx;
  ft.n_ex = fp->n_ex, ft.ex_space = fp->ex_space, ft.ex_kx = fp->ex_kx, ft.dex = fp->dex;
  for (int d = 0; d < fp->n_ex; d++)
    ft.ex_w[d] = fp->ex_w[d], ft.ex_a0[d] = fp->ex_a0[d], ft.ex_tenv[d] = fp->ex_tenv[d], ft.ex_wt[d] = fp->ex_wt[d];
  ft.n_peers = fp->n_peers, ft.my_rank = fp->my_rank, ft.epoch = fp->epoch;
  for (int r = 0; r < fp->n_peers; r++) ft.share_in[r] = fp->share_in[r], ft.flag_in[r] = fp->flag_in[r];
  return vdfdx_tma_f64(f_in, f_out, 1, nx, nv_local, v_local, dt, nullptr, k1x, parts, (cudaStream_t)stream, nullptr, &ft);
}

int adept_b200_reduce_parts_f64(const double* parts, int nparts, long long n, double scale_a, double scale_b,
                                const double* base, double* out, void* stream) {
  ADEPT_REQUIRE(parts, "parts") ADEPT_REQUIRE(out, "out")
  return reduce_parts_f64(parts, nparts, n, scale_a, scale_b, base, out, (cudaStream_t)stream);
}

int adept_b200_edfdv_exp_f64(const double* f_in, double* f_out, int batch, int nx, int nv, const double* e,
                             const double* dex, const double* pond, double charge, double mass, double dt, double k1v,
                             void* stream) {
  ADEPT_REQUIRE(f_in, "f_in") ADEPT_REQUIRE(f_out, "f_out") ADEPT_REQUIRE(e, "e")
  return edfdv_exp_f64(f_in, f_out, batch, nx, nv, e, dex, pond, charge, mass, dt, k1v, (cudaStream_t)stream);
}

int adept_b200_edfdv_spline_f64(const double* f_in, double* f_out, int batch, int nx, int nv, const double* e,
                                const double* dex, const double* pond, double charge, double mass, double dt,
                                double dv, void* stream) {
  ADEPT_REQUIRE(f_in, "f_in") ADEPT_REQUIRE(f_out, "f_out") ADEPT_REQUIRE(e, "e")
  return edfdv_spline_f64(f_in, f_out, batch, nx, nv, e, dex, pond, charge, mass, dt, dv, (cudaStream_t)stream);
}

int adept_b200_moments_f64(const double* f, int batch, int nx, int nv, const double* v, double scale_a,
                           const double* const* base_host, double* const* out_host, const double* scale_b_host,
                           void* stream) {
  ADEPT_REQUIRE(f, "f") ADEPT_REQUIRE(out_host, "out_host")
  return moments_f64(f, batch, nx, nv, v, scale_a, base_host, out_host, scale_b_host, (cudaStream_t)stream);
}

int adept_b200_poisson_f64(const double* rho, const double* kmul, long long kmul_stride, double* e, int batch, int nx,
                           int mode, double Te, double lambda_De, void* stream) {
  ADEPT_REQUIRE(rho, "rho") ADEPT_REQUIRE(kmul, "kmul") ADEPT_REQUIRE(e, "e")
  if (mode != 0 && mode != 1) {
    set_last_error("poisson: unknown mode %d", mode);
    return ADEPT_ERR_BAD_ARG;
  }
  return poisson_dispatch_f64(rho, kmul, kmul_stride, e, batch, nx, mode, Te, lambda_De, (cudaStream_t)stream);
}

int adept_b200_poisson_green_f64(const double* rho, const double* green, long long green_stride, double* e, int batch,
                                 int nx, void* stream) {
  ADEPT_REQUIRE(rho, "rho") ADEPT_REQUIRE(green, "green") ADEPT_REQUIRE(e, "e")
  return poisson_green_f64(rho, green, green_stride, e, batch, nx, (cudaStream_t)stream);
}

int adept_b200_row_means_f64(const double* a, int rows, long long n, double* out, void* stream) {
  ADEPT_REQUIRE(a, "a") ADEPT_REQUIRE(out, "out")
  return row_means_f64(a, rows, n, out, (cudaStream_t)stream);
}

int adept_b200_field_energy_f64(const double* e0, const double* de0, const double* e1, const double* de1, double w,
                                int batch, int nx, double* out, void* stream) {
  ADEPT_REQUIRE(e0, "e0") ADEPT_REQUIRE(de0, "de0") ADEPT_REQUIRE(out, "out")
  return field_energy_f64(e0, de0, e1, de1, w, batch, nx, out, (cudaStream_t)stream);
}

int adept_b200_axpy_f64(const double* a, const double* b, double s, double* out, long long n, void* stream) {
  ADEPT_REQUIRE(a, "a") ADEPT_REQUIRE(b, "b") ADEPT_REQUIRE(out, "out")
  return axpy_f64(a, b, s, out, n, (cudaStream_t)stream);
}

int adept_b200_ponderomotive_f64(const double* a, double* pond, int batch, int nx, double dx, void* stream) {
  ADEPT_REQUIRE(a, "a") ADEPT_REQUIRE(pond, "pond")
  return ponderomotive_f64(a, pond, batch, nx, dx, (cudaStream_t)stream);
}

int adept_b200_wave_step_f64(const double* a, const double* aold, const double* djy, const double* ne_n,
                             const double* ne_np1, double* a_new, int batch, int nx, double c, double dx, double dt,
                             void* stream) {
  ADEPT_REQUIRE(a, "a") ADEPT_REQUIRE(aold, "aold") ADEPT_REQUIRE(djy, "djy") ADEPT_REQUIRE(a_new, "a_new")
  return wave_step_f64(a, aold, djy, ne_n, ne_np1, a_new, batch, nx, c, dx, dt, (cudaStream_t)stream);
}

int adept_b200_collide_f64(const double* f_in, double* f_out, int batch, int nx, int nv, const double* v, double dv,
                           double dt, const double* nu_fp, const double* nu_K, const double* f_mx, int model,
                           int scheme, int nodrag, double sg_m, double sg_ratio, double* n_out, void* stream) {
  ADEPT_REQUIRE(f_in, "f_in") ADEPT_REQUIRE(f_out, "f_out") ADEPT_REQUIRE(v, "v")
  return collide_f64(f_in, f_out, batch, nx, nv, v, dv, dt, nu_fp, nu_K, f_mx, model, scheme, nodrag, sg_m, sg_ratio,
                     n_out, 1.0, 1.0, (cudaStream_t)stream);
}

int adept_b200_collide_sc_f64(const double* f_in, double* f_out, int batch, int nx, int nv, const double* v, double dv,
                              double dt, const double* nu_fp, const double* nu_K, const double* f_mx, int model,
                              int scheme, int nodrag, double sg_m, double sg_ratio, double* n_out, int sc_max_steps,
                              double sc_rtol, double sc_atol, void* stream) {
  ADEPT_REQUIRE(f_in, "f_in") ADEPT_REQUIRE(f_out, "f_out") ADEPT_REQUIRE(v, "v")
  return collide_f64(f_in, f_out, batch, nx, nv, v, dv, dt, nu_fp, nu_K, f_mx, model, scheme, nodrag, sg_m, sg_ratio,
                     n_out, 1.0, 1.0, (cudaStream_t)stream, sc_max_steps, sc_rtol, sc_atol);
}

int adept_b200_vdfdx_scratch_f64(const double* f_in, double* f_out, double* scratch, int batch, int nx, int nv,
                                 const double* v, double dt, double k1x, const double* k1x_batch, void* stream) {
  ADEPT_REQUIRE(f_in, "f_in") ADEPT_REQUIRE(f_out, "f_out") ADEPT_REQUIRE(v, "v") ADEPT_REQUIRE(scratch, "scratch")
  return bigx_apply_f64(f_in, f_out, scratch, batch, nx, nv, v, dt, k1x_batch, k1x, nullptr, 0, (cudaStream_t)stream);
}

int adept_b200_sum_peers_f64(const double* const* peers_host, int n_peers, long long n, double* out, void* stream) {
  ADEPT_REQUIRE(peers_host, "peers") ADEPT_REQUIRE(out, "out")
  for (int j = 0; j < n_peers && j < 8; j++) ADEPT_REQUIRE(peers_host[j], "peers[j]")
  return sum_peers_f64(peers_host, n_peers, n, out, (cudaStream_t)stream);
}

/* ---- vlasov-1d2v ---- */
int adept_b200_marginal_f64(const double* f, const double* wperp, long long rows, int nvperp, double* out, void* stream) {
  ADEPT_REQUIRE(f, "f") ADEPT_REQUIRE(wperp, "wperp") ADEPT_REQUIRE(out, "out")
  return marginal_f64(f, wperp, rows, nvperp, out, (cudaStream_t)stream);
}

int adept_b200_transpose_f64(const double* in, double* out, int batch, int n0, int n1, void* stream) {
  ADEPT_REQUIRE(in, "in") ADEPT_REQUIRE(out, "out")
  return transpose_f64(in, out, batch, n0, n1, (cudaStream_t)stream);
}

int adept_b200_collide_coef_f64(const double* f_in, double* f_out, int batch, int nx, int nv, const double* v, double dv,
                                double dt, const double* nu_fp, int model, int scheme, int nodrag, int sc_max_steps,
                                double sc_rtol, double sc_atol, const double* coef_in, double* coef_out, int coef_div,
                                void* stream) {
  ADEPT_REQUIRE(f_in, "f_in") ADEPT_REQUIRE(f_out, "f_out") ADEPT_REQUIRE(v, "v") ADEPT_REQUIRE(nu_fp, "nu_fp")
  if (!coef_in && !coef_out) {
    set_last_error("adept_b200_collide_coef_f64: give coef_in or coef_out");
    return ADEPT_ERR_BAD_ARG;
  }
  return collide_f64(f_in, f_out, batch, nx, nv, v, dv, dt, nu_fp, nullptr, nullptr, model, scheme, nodrag, 2.0, 0.5,
                     nullptr, 1.0, 1.0, (cudaStream_t)stream, sc_max_steps, sc_rtol, sc_atol, coef_in, coef_out, coef_div);
}

}  // extern "C"
