// x-advection for the largest grids (nx = 4096): persistent CTAs, TMA-staged x-pencil tiles, and TWO complex FFTs per
// thread scheduled half a pass apart, so that the shared-memory exchange of one transform runs under the butterflies of
// the other.
//
// Reference semantics: SpaceExponential.push  adept/_vlasov1d/solvers/pushers/vlasov.py:234-251, followed by the
// velocity sum of compute_charge_density  adept/_vlasov1d/solvers/pushers/field.py:197-208.
//
// Why this shape.  The spectral pushes are bound by two SM pipes, not by HBM: a 4096-point two-for-one push needs ~46
// fp64 instructions and ~0.85 shared-memory wavefronts per cell, each worth about the same time, and in
// vdfdx_tma_kernel (one 512-thread CTA, both transforms of a tile in the same phase, every exchange a CTA barrier) the
// two pipes take turns: ncu shows fp64 36 %, LSU 56 %, issue-active 29 %.  The pipes do overlap when independent
// instruction streams use them (tools/micro/fp64_lsu_overlap.cu), so here ONE thread owns row t + T m of all four
// columns of the tile = element m of two independent complex transforms F and S (255 registers, one 256-thread CTA
// per SM), and the program order between two CTA barriers is always
//     loads of one transform | second butterfly half of the other | its stores | first butterfly half of the first
// -- every interval carries ~1000 cycles of LSU work and ~950 cycles of fp64 work that do not depend on each other,
// and one barrier serves the hazards of both transforms (half as many barriers per transform).
//
// Bank conflicts of the [row][4 doubles] layouts.  The TMA landing zone (and the output staging boxes) hold 32 bytes
// per row, so a warp touching "the first column pair" of 32 consecutive rows with all lanes would hit every bank twice.
// Lanes with bit 2 set therefore access the two 16-byte halves of their row in the opposite ORDER (a quarter-warp then
// covers units 0,2,4,6,9,11,13,15 -- all different mod 8) and the values are put in place with register selects
// (integer pipe, idle here).  The transforms themselves are not swapped: F is columns 0-1 for every thread, because
// all threads of the CTA cooperate on one transform and must agree on when its exchange happens.
#include <cuda.h>
#include <stdlib.h>

#include "field_tail.cuh"
#include "internal.h"
#include "push_core.cuh"
#include "tma.cuh"
#include "tmem.cuh"

namespace adept {

struct DualPushArgs {
  double* fout;
  int batch, nx, nv;
  int ntiles;           // batch * nv / 4
  const double* v;      // [nv]
  const double* k1_batch;
  double k1, dt;
  const cplx* tw;
  int zero;
  CUtensorMap out_map;  // f_out with boxes {4, BOX_ROWS} (staged output chunks)
  double* partial;      // [gridDim.x, batch*nx] per-CTA row sums of f_out, or null
  const double* filt;   // nullable [N/2+1]: real multiplier per mode (Hou-Li filter)
  FieldTail ft;         // used by the FIELD instantiation only
};

template <int LOGN, int NSTAGE_REQ = 16>
struct DualCfg {
  using C = FftCfg<LOGN>;
  using PC = PhaseCfg<LOGN>;
  static constexpr int N = C::N, T = C::T, E = C::E;
  static_assert(E == 16 && C::NPASS == 3, "dual x-advection: three passes of 16 points per thread (nx = 512 .. 4096)");
  static constexpr int THREADS = T;
  static constexpr int BOX_ROWS = T;              // one TMA box = the rows t + T m of one register index m
  static constexpr size_t BUF_BYTES = (size_t)C::BUF * sizeof(cplx);  // one padded exchange buffer
  static_assert(BUF_BYTES % 128 == 0, "exchange buffers must be bank-aligned to each other");
  static constexpr size_t LAND_BYTES = (size_t)N * 4 * sizeof(double);  // aliases the two exchange buffers
  static_assert(LAND_BYTES <= 2 * BUF_BYTES, "landing zone");
  static constexpr size_t PH_OFF = 2 * BUF_BYTES;
  static constexpr size_t PH_BYTES = (size_t)4 * PC::PER_SEQ * sizeof(cplx);
  static constexpr size_t BAR_OFF = PH_OFF + PH_BYTES;
  static constexpr size_t STAGE_OFF = (BAR_OFF + 16 + 127) / 128 * 128;
  static constexpr size_t BOX_BYTES = (size_t)BOX_ROWS * 4 * sizeof(double);
  // output chunks (register index m < NSTAGE) that leave through the staging area and TMA tensor stores; the others
  // leave as one 32-byte store per row.  As many as fit beside the exchange buffers.
  static constexpr int NSTAGE_FIT = (int)((227 * 1024 - STAGE_OFF) / BOX_BYTES);
  static constexpr int NSTAGE = NSTAGE_FIT > NSTAGE_REQ ? NSTAGE_REQ : NSTAGE_FIT;
  static constexpr size_t SMEM = STAGE_OFF + (size_t)NSTAGE * BOX_BYTES;
  static constexpr int TMEM_COLS = 64;  // row-sum accumulators: 32 columns for each of the two threads of a lane
};

// One pass of the Stockham transform of fft_core.cuh in four pieces that the kernel interleaves across two transforms.
template <int LOGN, int P>
struct DPass {
  using C = FftCfg<LOGN>;
  static constexpr int R = C::radix(P), NS = C::ns(P), T = C::T, E = C::E, Q = E / R;
  static __device__ __forceinline__ void ld(cplx (&x)[E], const cplx* __restrict__ buf, int t) {
#pragma unroll
    for (int m = 0; m < E; m++) x[m] = buf[fft_pad(t + T * m)];
  }
  // the six twiddles a radix-16 butterfly loads (w^1..3, w^4, w^8, w^12; fft_core.cuh); issued a block ahead of their
  // use: with 227 KB of shared memory the SM has no L1 left, so they come from L2
  struct Tw {
    cplx w[6];
  };
  static __device__ __forceinline__ void tw_load(Tw& q, const cplx* __restrict__ tw, int t) {
    if constexpr (R == 16 && NS > 1) {
      const cplx* twp = tw + C::tw_off(P) + (t & (NS - 1));
      q.w[0] = __ldg(twp), q.w[1] = __ldg(twp + NS), q.w[2] = __ldg(twp + 2 * NS);
      q.w[3] = __ldg(twp + 3 * NS), q.w[4] = __ldg(twp + 7 * NS), q.w[5] = __ldg(twp + 11 * NS);
    }
  }
  // first part: twiddles + column DFT4s of a radix-16 butterfly (the whole butterfly for the short last pass)
  static __device__ __forceinline__ void h1(cplx (&x)[E], const Tw& q, const cplx* __restrict__ tw, int t) {
    if constexpr (R == 16) {
      if constexpr (NS > 1) {
        const cplx w4 = q.w[3], w8 = q.w[4], w12 = q.w[5];
#pragma unroll
        for (int c = 0; c < 4; c++) {
          x[c + 4] = cmul(x[c + 4], w4);
          x[c + 8] = cmul(x[c + 8], w8);
          x[c + 12] = cmul(x[c + 12], w12);
          dft4(x[c], x[c + 4], x[c + 8], x[c + 12]);
          if (c > 0) {
#pragma unroll
            for (int j = 0; j < 4; j++) x[c + 4 * j] = cmul(x[c + 4 * j], q.w[c - 1]);
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < 4; c++) dft4(x[c], x[c + 4], x[c + 8], x[c + 12]);
      }
    } else {
#pragma unroll
      for (int q2 = 0; q2 < Q; q2++) {
        const int k = (t + T * q2) & (NS - 1);
        cplx v[R];
#pragma unroll
        for (int r = 0; r < R; r++) v[r] = x[q2 + r * Q];
        const cplx* twp = tw + C::tw_off(P) + k;
#pragma unroll
        for (int r = 1; r < R; r++) v[r] = cmul(v[r], __ldg(twp + (r - 1) * NS));
        dft_r<R>(v);
#pragma unroll
        for (int r = 0; r < R; r++) x[q2 + r * Q] = v[r];
      }
    }
  }
  static __device__ __forceinline__ void h2(cplx (&x)[E]) {
    if constexpr (R == 16) dft16_finish(x);
  }
  static __device__ __forceinline__ void st(const cplx (&x)[E], cplx* __restrict__ buf, int t) {
#pragma unroll
    for (int q = 0; q < Q; q++) {
      const int b = t + T * q;
      const int k = b & (NS - 1);
      const int j0 = (b - k) * R + k;
#pragma unroll
      for (int r = 0; r < R; r++) buf[fft_pad(j0 + r * NS)] = x[q + r * Q];
    }
  }
};

// The half-spectrum update of push_core.cuh in its three barrier-separated pieces.
template <int LOGN>
struct DHalf {
  using C = FftCfg<LOGN>;
  using PC = PhaseCfg<LOGN>;
  static constexpr int N = C::N, E = C::E, T = C::T, H = E / 2;
  // publish the upper register half (modes k >= N/2) in natural order
  static __device__ __forceinline__ void publish(const cplx (&x)[E], cplx* __restrict__ buf, int t) {
#pragma unroll
    for (int m = H; m < E; m++) buf[fft_pad(t + T * m)] = x[m];
  }
  // update the (k, N - k) pairs of the lower register half; partner values are written back
  static __device__ __forceinline__ void pairs(cplx (&x)[E], cplx* __restrict__ buf, const cplx* __restrict__ ph, int t,
                                               const double* __restrict__ filt) {
    const cplx* pha = ph;
    const cplx* phb = ph + PC::PER_SEQ;
    cplx pa = cmul(pha[t & (PC::NLO - 1)], pha[PC::NLO + (t >> PC::LOBT)]);
    cplx pb = cmul(phb[t & (PC::NLO - 1)], phb[PC::NLO + (t >> PC::LOBT)]);
    const cplx sa = pha[PC::STEP], sb = phb[PC::STEP];
#pragma unroll
    for (int m = 0; m < H; m++) {
      if (m > 0) {
        pa = cmul(pa, sa);
        pb = cmul(pb, sb);
      }
      const int k = t + T * m;
      const bool self = (k == 0);
      const int q = fft_pad((N - k) & (N - 1));
      const cplx zk = x[m];
      const cplx zq = self ? zk : buf[q];
      const cplx A = cmake(zk.x + zq.x, zk.y - zq.y);
      const cplx B = cmake(zk.y + zq.y, zq.x - zk.x);
      cplx Ap = cmul(A, pa), Bp = cmul(B, pb);
      if (filt) {
        const double s = __ldg(filt + k);
        Ap.x *= s, Ap.y *= s, Bp.x *= s, Bp.y *= s;
      }
      x[m] = cmake(Ap.y + Bp.x, Ap.x - Bp.y);
      if (!self) buf[q] = cmake(Bp.x - Ap.y, Ap.x + Bp.y);
    }
    if (t == 0) {  // Nyquist mode: real phase, pairs with itself
      const int q = fft_pad(N / 2);
      const cplx z = buf[q];
      const double s = filt ? __ldg(filt + N / 2) : 1.0;
      buf[q] = cmake(2.0 * z.y * phb[PC::NYQ].x * s, 2.0 * z.x * pha[PC::NYQ].x * s);
    }
  }
  static __device__ __forceinline__ void collect(cplx (&x)[E], const cplx* __restrict__ buf, int t) {
#pragma unroll
    for (int m = H; m < E; m++) x[m] = buf[fft_pad(t + T * m)];
  }
};

__device__ __forceinline__ void stg256(double* dst, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

template <int LOGN, bool FIELD, int NSTAGE_REQ>
__global__ void __launch_bounds__(DualCfg<LOGN>::THREADS, 1)
    vdfdx_dual_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ DualPushArgs p) {
  using K = DualCfg<LOGN, NSTAGE_REQ>;
  using C = FftCfg<LOGN>;
  using PC = PhaseCfg<LOGN>;
  using P0 = DPass<LOGN, 0>;
  using P1 = DPass<LOGN, 1>;
  using P2 = DPass<LOGN, 2>;
  using HS = DHalf<LOGN>;
  constexpr int N = C::N, E = C::E, T = C::T;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cplx* land = reinterpret_cast<cplx*>(smem_raw);  // tile[row][2 cplx]; aliases both exchange buffers
  cplx* ph_all = reinterpret_cast<cplx*>(smem_raw + K::PH_OFF);  // [4 columns][PER_SEQ]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + K::BAR_OFF);
  double* stage = reinterpret_cast<double*>(smem_raw + K::STAGE_OFF);  // [NSTAGE][BOX_ROWS][4]

  const int t = threadIdx.x;
  const int g = (t >> 2) & 1;  // access-order swap of the 32-byte rows (see above)
  cplx* bf = reinterpret_cast<cplx*>(smem_raw);                  // exchange buffer of F (columns 0-1)
  cplx* bs = reinterpret_cast<cplx*>(smem_raw + K::BUF_BYTES);   // exchange buffer of S (columns 2-3)
  const cplx* phf = ph_all;
  const cplx* phs = ph_all + (size_t)2 * PC::PER_SEQ;
  const int tiles_per_member = p.nv >> 2;
  const cplx* tw = p.tw;
  const cplx* twi = p.tw + p.zero;  // same table; a separate name keeps the two transforms' loads apart
  const unsigned zbits = (unsigned)p.zero;  // always 0: BLK() below

  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  if (t == 0) {
    mbar_init(bar, 2);  // a tile arrives in two halves (below)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (t < 32) tmem_alloc(tmem_slot, K::TMEM_COLS);
  tmem_fence_before_sync();
  __syncthreads();
  tmem_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  // rows [half N/2, (half + 1) N/2) of tile tl.  The first half lands inside F's exchange buffer, which dies one
  // interval before S's, so it is issued one barrier earlier.
  static_assert(K::LAND_BYTES / 2 <= K::BUF_BYTES, "first half of the landing zone must lie inside F's buffer");
  auto issue_half = [&](int tl, int half) {
    const int b = tl / tiles_per_member, cg = tl - b * tiles_per_member;
    mbar_expect_tx(bar, (uint32_t)(K::LAND_BYTES / 2));
#pragma unroll 1
    for (int bx = 8 * half; bx < 8 * half + 8; bx++)
      tma_load_2d(reinterpret_cast<double*>(land) + (size_t)bx * K::BOX_ROWS * 4, &in_map, bar, cg * 4,
                  b * N + bx * K::BOX_ROWS);
  };
  auto fill_phases = [&](int tl) {  // 4 columns x PER_SEQ entries, one sincos each
    const int b = tl / tiles_per_member, cg = tl - b * tiles_per_member;
    const double k1 = p.k1_batch ? p.k1_batch[b] : p.k1;
    for (int i = t; i < 4 * PC::PER_SEQ; i += K::THREADS) {
      const int s = i / PC::PER_SEQ, j = i % PC::PER_SEQ;
      const double al = k1 * (p.v[4 * cg + s] * p.dt);
      const double sc = 0.5 / (double)N;
      const bool lo = j < PC::NLO, nyq = j == PC::NYQ;
      const int mult = lo ? j : (j < PC::STEP ? ((j - PC::NLO) << PC::LOBT) : (j == PC::STEP ? T : N / 2));
      double sn, cs;
      sincos((double)mult * al, &sn, &cs);
      const double amp = (lo || nyq) ? sc : 1.0;
      ph_all[i] = cmake(cs * amp, nyq ? 0.0 : -sn * amp);
    }
  };

  // Row sums of this CTA's tiles (row t + T m is owned by this thread for every tile) live in tensor memory between
  // tiles: 16 doubles = 32 columns per thread, fetched and put back at the end of a tile when the transform registers
  // are dead.  (Held in registers they cost 32 of the 255 and the twiddles could not be loaded a block early.)
  const uint32_t racc_addr = tmem_addr(tmem_base, t >> 5, 32 * (t >> 7));
  {
    const double zeros[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    tmem_st8(racc_addr, zeros);
    tmem_st8(racc_addr + 16, zeros);
    tmem_wait_st();
  }
  auto flush_rho = [&](int b) {
    double* dst = p.partial + ((size_t)blockIdx.x * p.batch + b) * N;
    const double zeros[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int h = 0; h < 2; h++) {
      double r[8];
      tmem_ld8(racc_addr + 16 * h, r);
#pragma unroll
      for (int m = 0; m < 8; m++) {
        // a single member is visited once by every CTA: plain store, no zero-initialisation needed
        double* d = dst + t + T * (8 * h + m);
        *d = (p.batch == 1) ? r[m] : *d + r[m];
      }
      tmem_st8(racc_addr + 16 * h, zeros);
    }
    tmem_wait_st();
  };

  int tl = blockIdx.x;
  if (tl < p.ntiles) {
    if (t == 0) issue_half(tl, 0), issue_half(tl, 1);
    fill_phases(tl);
  }
  uint32_t parity = 0;
  int cur_b = -1;
  for (; tl < p.ntiles; tl += gridDim.x) {
    const int b = tl / tiles_per_member, cg = tl - b * tiles_per_member;
    if (p.partial && b != cur_b) {
      if (cur_b >= 0) flush_rho(cur_b);
      cur_b = b;
    }
    mbar_wait(bar, parity);
    parity ^= 1;

    cplx xf[E], xs[E];
#pragma unroll
    for (int m = 0; m < E; m++) {
      const cplx u0 = land[(t + T * m) * 2 + g], u1 = land[(t + T * m) * 2 + 1 - g];
      xf[m] = g ? u1 : u0;
      xs[m] = g ? u0 : u1;
    }
    __syncthreads();  // landing zone read by everybody (it aliases the exchange buffers); phase tables complete
    // Every interval between two barriers is two basic blocks -- BLK(k) is a branch on a bit of a kernel argument that
    // is always zero, which neither nvcc nor ptxas can fold -- because ptxas schedules freely inside a basic block and
    // otherwise sinks the stores of an interval behind ALL its arithmetic (measured on the SASS: the second butterfly
    // half of the other transform ended up last, the stores right in front of the barrier, where nothing overlaps
    // their drain).  Block A = loads of X | second half of Y | stores of Y; block B = first half of X.
#define BLK(k) if (((zbits >> (k)) & 1) == 0)
    // ---- forward transforms, S half a pass behind F ------------------------------------------------------------
    typename P0::Tw q1;
    BLK(0) {
      P0::tw_load(q1, tw, t);
      P0::h1(xf, typename P0::Tw(), tw, t);
      P0::h2(xf);
      P0::st(xf, bf, t);
    }
    BLK(1) { P0::h1(xs, q1, tw, t); }
    __syncthreads();
    typename P1::Tw q3;
    BLK(2) {
      P1::tw_load(q3, tw, t);
      P1::ld(xf, bf, t);
      P0::h2(xs);
      P0::st(xs, bs, t);
    }
    BLK(3) { P1::h1(xf, q3, tw, t); }
    __syncthreads();
    typename P1::Tw q5;
    BLK(4) {
      P1::tw_load(q5, tw, t);
      P1::ld(xs, bs, t);
      P1::h2(xf);
      P1::st(xf, bf, t);
    }
    BLK(5) { P1::h1(xs, q5, tw, t); }
    __syncthreads();
    typename P2::Tw q7;
    BLK(6) {
      P2::tw_load(q7, tw, t);
      P2::ld(xf, bf, t);
      P1::h2(xs);
      P1::st(xs, bs, t);
    }
    BLK(7) { P2::h1(xf, q7, tw, t); }
    __syncthreads();
    typename P2::Tw q9;
    BLK(8) {
      P2::tw_load(q9, tw, t);
      P2::ld(xs, bs, t);
      P2::h2(xf);
      HS::publish(xf, bf, t);
    }
    BLK(9) { P2::h1(xs, q9, tw, t); }
    __syncthreads();
    // ---- half-spectrum updates, then the inverse transforms (swap . forward . swap) ------------------------------
    BLK(10) {
      P2::h2(xs);
      HS::publish(xs, bs, t);
    }
    BLK(11) { HS::pairs(xf, bf, phf, t, p.filt); }
    __syncthreads();
    typename P0::Tw q13;
    BLK(12) {
      P0::tw_load(q13, twi, t);
      HS::collect(xf, bf, t);
      HS::pairs(xs, bs, phs, t, p.filt);
    }
    BLK(13) { P0::h1(xf, q13, twi, t); }
    __syncthreads();
    typename P0::Tw q15;
    BLK(14) {
      P0::tw_load(q15, twi, t);
      HS::collect(xs, bs, t);
      P0::h2(xf);
      P0::st(xf, bf, t);
    }
    BLK(15) { P0::h1(xs, q15, twi, t); }
    __syncthreads();
    typename P1::Tw q17;
    BLK(16) {
      P1::tw_load(q17, twi, t);
      P1::ld(xf, bf, t);
      P0::h2(xs);
      P0::st(xs, bs, t);
    }
    BLK(17) { P1::h1(xf, q17, twi, t); }
    __syncthreads();
    typename P1::Tw q19;
    BLK(18) {
      P1::tw_load(q19, twi, t);
      P1::ld(xs, bs, t);
      P1::h2(xf);
      P1::st(xf, bf, t);
    }
    BLK(19) { P1::h1(xs, q19, twi, t); }
    __syncthreads();
    typename P2::Tw q21;
    BLK(20) {
      P2::tw_load(q21, twi, t);
      P2::ld(xf, bf, t);
      P1::h2(xs);
      P1::st(xs, bs, t);
    }
    BLK(21) { P2::h1(xf, q21, twi, t); }
    const bool more = tl + (int)gridDim.x < p.ntiles;
    // the staging area is written below: the previous tile's TMA stores (issued a whole tile ago) must have finished
    // reading it
    if (K::NSTAGE > 0 && t == 0) tma_wait_read_all();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();  // F's exchange buffer is dead: the first half of the next tile may land; staging area free
    if (t == 0 && more) issue_half(tl + gridDim.x, 0);
    // ---- output: columns 4 cg .. 4 cg + 3 of rows t + T m; column 0 / 1 = F[m].y / F[m].x, column 2 / 3 from S ------
    // (F's and S's halves of a staged row are written an interval apart, each with a two-way bank conflict; both
    // intervals are light on shared-memory traffic)
    typename P2::Tw q23;
    BLK(22) {
      P2::tw_load(q23, twi, t);
      P2::ld(xs, bs, t);
      P2::h2(xf);
#pragma unroll
      for (int m = 0; m < K::NSTAGE; m++)
        *reinterpret_cast<double2*>(stage + ((size_t)m * K::BOX_ROWS + t) * 4) = make_double2(xf[m].y, xf[m].x);
    }
    BLK(23) { P2::h1(xs, q23, twi, t); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();  // S's exchange buffer is dead too
    if (t == 0 && more) issue_half(tl + gridDim.x, 1);
    BLK(24) {
      P2::h2(xs);
#pragma unroll
      for (int m = 0; m < K::NSTAGE; m++)
        *reinterpret_cast<double2*>(stage + ((size_t)m * K::BOX_ROWS + t) * 4 + 2) = make_double2(xs[m].y, xs[m].x);
    }
#undef BLK
    if constexpr (K::NSTAGE > 0) {
      fence_async_smem();
      __syncthreads();
      if (t == 0) {
#pragma unroll 1
        for (int m = 0; m < K::NSTAGE; m++)
          tma_store_2d(&p.out_map, stage + (size_t)m * K::BOX_ROWS * 4, cg * 4, b * N + m * K::BOX_ROWS);
        tma_commit_group();
      }
    }
    double* dst = p.fout + ((size_t)b * N) * p.nv + 4 * cg;
#pragma unroll
    for (int m = K::NSTAGE; m < E; m++)
      stg256(dst + (size_t)(t + T * m) * p.nv, xf[m].y, xf[m].x, xs[m].y, xs[m].x);
    if (p.partial) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        double r[8];
        tmem_ld8(racc_addr + 16 * h, r);
#pragma unroll
        for (int m = 0; m < 8; m++) r[m] += (xf[8 * h + m].y + xf[8 * h + m].x) + (xs[8 * h + m].y + xs[8 * h + m].x);
        tmem_st8(racc_addr + 16 * h, r);
      }
      tmem_wait_st();
    }
    if (more) fill_phases(tl + gridDim.x);  // tables were last read two barriers ago
  }
  if (p.partial && cur_b >= 0) flush_rho(cur_b);

  if constexpr (FIELD) {
    __syncthreads();
    // the exchange buffers are dead; the phase tables (not the staging area, which the last TMA stores may still be
    // reading) hold the tail's small reduction scratch
    static_assert(K::PH_BYTES >= (size_t)(K::THREADS / 32) * 32 * sizeof(double), "field tail scratch");
    field_tail_solve<N, K::THREADS>(p.ft, p.partial, reinterpret_cast<double*>(smem_raw),
                                    reinterpret_cast<double*>(smem_raw + K::PH_OFF));
  }
  // the last tile's staged stores must be complete before the CTA exits (its shared memory is their source)
  if (K::NSTAGE > 0 && t == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  tmem_fence_before_sync();
  __syncthreads();
  if (t < 32) tmem_dealloc(tmem_base, K::TMEM_COLS);
}

// ---- host side ---------------------------------------------------------------------------------------------------
template <int LOGN>
static int dual_ctas(int ntiles) {
  using K = DualCfg<LOGN>;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_sm = (int)((227 * 1024) / (K::SMEM + 1024));
  const int by_regs = 65536 / (K::THREADS * 256);
  if (per_sm > by_regs) per_sm = by_regs;
  if (per_sm < 1) per_sm = 1;
  const int grid = sms * per_sm;
  return grid < ntiles ? grid : ntiles;
}

template <int LOGN, bool FIELD, int NSTAGE_REQ>
static int launch_dual(const CUtensorMap& map, const DualPushArgs& p, int grid, cudaStream_t stream) {
  using K = DualCfg<LOGN, NSTAGE_REQ>;
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = vdfdx_dual_kernel<LOGN, FIELD, NSTAGE_REQ>;
  if (dev < 64 && !configured[dev]) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM);
    if (err != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(vdfdx_dual, smem=%zu): %s", K::SMEM, cudaGetErrorString(err));
      return ADEPT_ERR_CUDA;
    }
    configured[dev] = true;
  }
  ProfileScope prof(FIELD ? "vdfdx_dual_field" : "vdfdx_dual", stream);
  if (FIELD) {
    void* args[2] = {const_cast<CUtensorMap*>(&map), const_cast<DualPushArgs*>(&p)};
    cudaError_t err = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(grid), dim3(K::THREADS), args,
                                                  K::SMEM, stream);
    if (err != cudaSuccess) {
      set_last_error("cudaLaunchCooperativeKernel(vdfdx_dual + field): %s", cudaGetErrorString(err));
      (void)cudaGetLastError();
      return ADEPT_ERR_CUDA;
    }
    return check_launch("vdfdx_dual_kernel(field)");
  }
  kern<<<grid, K::THREADS, K::SMEM, stream>>>(map, p);
  return check_launch("vdfdx_dual_kernel");
}

bool vdfdx_dual_supported(int nx) { return nx == 4096; }

int vdfdx_dual_parts(int batch, int nx, int nv) {
  (void)nx;
  return dual_ctas<12>(batch * (nv / 4));
}

int vdfdx_dual_f64(const double* fin, double* fout, int batch, int nx, int nv, const double* v, double dt,
                   const double* k1_batch, double k1, double* partial, cudaStream_t stream, const double* filt,
                   const FieldTail* field) {
  using K = DualCfg<12>;
  if (nx != 4096 || nv % 4 || batch < 1) {
    set_last_error("vdfdx(dual): unsupported batch=%d nx=%d nv=%d", batch, nx, nv);
    return ADEPT_ERR_UNSUPPORTED;
  }
  CUtensorMap map;
  int rc = encode_map_2d(&map, fin, (unsigned long long)nv, (unsigned long long)batch * nx,
                         (unsigned long long)nv * sizeof(double), 4, K::BOX_ROWS, 0);
  if (rc != ADEPT_OK) return rc;
  DualPushArgs p = {};
  rc = encode_map_2d(&p.out_map, fout, (unsigned long long)nv, (unsigned long long)batch * nx,
                     (unsigned long long)nv * sizeof(double), 4, K::BOX_ROWS, 0);
  if (rc != ADEPT_OK) return rc;
  p.fout = fout, p.batch = batch, p.nx = nx, p.nv = nv, p.ntiles = batch * (nv / 4);
  p.v = v, p.k1_batch = k1_batch, p.k1 = k1, p.dt = dt, p.zero = 0, p.partial = partial, p.filt = filt;
  p.tw = get_twiddles(12);
  if (!p.tw) return ADEPT_ERR_CUDA;
  const int grid = vdfdx_dual_parts(batch, nx, nv);
  // ADEPT_B200_DUAL_NSTAGE=<n>: staged output chunks (development knob: 6 or the default, as many as fit)
  static int nstage = -1;
  if (nstage < 0) {
    const char* e = getenv("ADEPT_B200_DUAL_NSTAGE");
    nstage = e ? atoi(e) : 16;
  }
  if (field) p.ft = *field;
  if (nstage == 6) return field ? launch_dual<12, true, 6>(map, p, grid, stream) : launch_dual<12, false, 6>(map, p, grid, stream);
  return field ? launch_dual<12, true, 16>(map, p, grid, stream) : launch_dual<12, false, 16>(map, p, grid, stream);
}

}  // namespace adept
