// Fused v-row kernel: spectral v-advection followed by the Fokker-Planck collision step on the same rows, one HBM read
// and one HBM write of f for both operators (32 B per cell instead of 2 x 16 B + the round trip through HBM).
//
// Reference semantics: VelocityExponential.push (adept/_vlasov1d/solvers/pushers/vlasov.py:74-91) followed by
// Collisions._collide / _solve_one_x (adept/_vlasov1d/solvers/pushers/fokker_planck.py:368-433), as composed by
// VlasovPoissonFokkerPlanck.__call__ (adept/_vlasov1d/solvers/vector_field.py:236-238).
//
// One row pair per CTA of T = nv/16 threads.  The pair is transformed as one complex FFT (push_core.cuh), the result
// a'[e], b'[e] (e = t + T m) is scattered from registers into two chunk-padded row buffers that alias the Stockham
// exchange buffer (same size: 17/16 nv complex = 2 x 17/16 nv doubles), each row is solved by fp_row_fast
// (collide_core.cuh) with the same T threads (16 contiguous cells each), and the rows are stored coalesced.
#include <stdlib.h>

#include "collide_core.cuh"
#include "internal.h"
#include "push_core.cuh"
#include "tma.cuh"

namespace adept {

struct VrowArgs {
  const double* fin;
  double* fout;
  long long npairs;  // batch * nx / 2
  int nx;
  const double* e;
  const double* dex;   // nullable
  const double* pond;  // nullable
  double q, m, dt, k1;
  double dt_fp;  // time step of the collision operator (the push may be a substep of a splitting scheme)
  const cplx* tw;
  int zero;
  // collisions
  const double* v;  // [nv]
  double dv;
  const double* nu_fp;  // [batch*nx]
  double nu_fp_scale;
  const double* trow;  // nullable device-resident time row (common.cuh): nu_fp_scale = trow[TROW_NU_FP]
  int model, scheme;
  // peer mode (single grid sharded over GPUs, every buffer v-sharded [nx_global, nv / P] and mapped over NVLink): cell i
  // of local row r is READ from in_peer[i >> nvp_shift] and WRITTEN to out_peer[i >> nvp_shift], both at row
  // row0_global + r, column i & mask -- the two layout transposes of the decomposition ride on the kernel's own loads
  // and stores (whole 8 * nv / P byte row segments per peer: NVLink-friendly).  nvp_shift < 0: off.
  const double* in_peer[8];
  double* out_peer[8];
  int nvp_shift;
  long long row0_global;
  long long nx_global;  // rows of every rank's v-sharded buffer
  // Staged peer input (PEER instantiation, stage != null): the first n_movers CTAs of the grid do no arithmetic; they
  // gather the rank's rows from the owning ranks into stage[nx_local][nv] (bulk loads over NVLink into shared memory,
  // one bulk store per row), row m + i n_movers by mover m in round i, and count themselves into round_ctr[i] when the
  // row has landed.  The compute CTA of a row pair waits for its round and bulk-loads the pair from the stage -- the
  // NVLink latency of a pair's input no longer sits between that CTA's launch and its arithmetic.
  double* stage;
  unsigned int* round_ctr;
  int n_movers;
};

// Peer mode with bulk transfers (PEER instantiation): the tensor maps of the P output buffers, one per rank
struct alignas(64) PeerMaps {
  CUtensorMap m[8];
};

template <int LOGN>
struct VrowCfg {
  using C = FftCfg<LOGN>;
  using PC = PhaseCfg<LOGN>;
  static constexpr int N = C::N, T = C::T;
  static_assert(C::E == 16, "fused v-row kernel needs nv >= 16");
  static constexpr int THREADS = T < 32 ? 32 : T;
  static constexpr bool WARP_MODE = (T % 32) == 0;
  static constexpr size_t BUF_BYTES = (size_t)C::BUF * sizeof(cplx);            // == 2 rows x (N + T) doubles
  static constexpr size_t PH_BYTES = ((size_t)2 * PC::PER_SEQ * sizeof(cplx) + 127) / 128 * 128;  // phase tables
  static constexpr size_t RED_BYTES = (size_t)2 * (WARP_MODE ? T / 32 : (T > 32 ? T : 32)) * 6 * sizeof(double);  // room for the three moments of both rows (row_reduce<6>; fp_row_fast uses half)
  static constexpr size_t PCR_BYTES = (size_t)6 * T * sizeof(double);
  // ~84 KB at nv = 4096: two CTAs per SM leave ~60 KB of the 228 KB array to L1, enough for the twiddle rows in use
  static constexpr size_t SMEM = BUF_BYTES + PH_BYTES + RED_BYTES + PCR_BYTES + 16;  // + one mbarrier (peer bulk loads)
  static constexpr int OUT_BOX_ROWS = (N / 16) < 256 ? (N / 16) : 256;  // 128-byte chunks per TMA store box
};

// TMA_OUT: the solved rows leave shared memory through TMA tensor stores (no LDS + STG pass for the output); the
// tensor map views f_out as [rows * nv/16][16] with boxes {16, min(256, nv/16)}, 128-byte swizzle.
// PEER (with TMA_OUT): every row segment travels as one bulk transfer -- cp.async.bulk loads of the 8 nv/P-byte
// segments of both rows from the P owning ranks into the (still unused) exchange buffer, and one TMA tensor store per
// rank and row from the solved dense row (out_maps[j] views rank j's buffer); the NVLink sees 4-16 KB requests instead
// of a warp's 256-byte loads and stores, and the stores leave the load/store pipe.
template <int LOGN, bool TMA_OUT, bool CC, int TW, bool PEER>
__device__ __forceinline__ void vrow_body(const CUtensorMap* out_maps, const VrowArgs& p, const long long pair) {
  using K = VrowCfg<LOGN>;
  using C = FftCfg<LOGN>;
  using PC = PhaseCfg<LOGN>;
  constexpr int N = C::N, E = C::E, T = C::T;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  cplx* buf = reinterpret_cast<cplx*>(smem_raw);
  cplx* ph = reinterpret_cast<cplx*>(smem_raw + K::BUF_BYTES);
  double* red = reinterpret_cast<double*>(smem_raw + K::BUF_BYTES + K::PH_BYTES);
  double* pcr = reinterpret_cast<double*>(smem_raw + K::BUF_BYTES + K::PH_BYTES + K::RED_BYTES);
  uint64_t* ld_bar = reinterpret_cast<uint64_t*>(smem_raw + K::BUF_BYTES + K::PH_BYTES + K::RED_BYTES + K::PCR_BYTES);
  if (TMA_OUT && (smem_u32(smem_raw) & 1023u)) __trap();  // the swizzled tiles need a 1024-byte aligned window

  const int t = threadIdx.x < T ? threadIdx.x : 0;  // spare threads (T < 32) shadow thread 0 and never store
  const bool live = threadIdx.x < T;
  const long long row0 = 2 * pair;
  const double* a_in = p.fin + row0 * N;
  const double* b_in = a_in + N;

  double alpha[2];
  {
    const double q2m = p.q * p.q / p.m;
#pragma unroll
    for (int s = 0; s < 2; s++) {
      double ee = p.e[row0 + s];
      if (p.dex) ee = __dadd_rn(ee, p.dex[row0 + s]);
      const double pd = p.pond ? p.pond[row0 + s] : 0.0;
      alpha[s] = p.k1 * (p.dt * accel_of(ee, pd, p.q, q2m, p.m));
    }
  }

  cplx x[E];
  if constexpr (PEER) {
    const size_t nvp = (size_t)1 << p.nvp_shift;
    double* in_rows = reinterpret_cast<double*>(buf);  // two dense rows [2][N] in the exchange buffer
    if (threadIdx.x == 0) {
      mbar_init(ld_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      mbar_expect_tx(ld_bar, (uint32_t)(2 * N * sizeof(double)));
      if (p.stage) {  // both rows are adjacent in the stage: one bulk load, once the movers have landed their round
        if (p.n_movers > 0) {  // (n_movers == 0: the caller filled the stage before the launch, e.g. by copy engines)
          const unsigned int* ctr = p.round_ctr + row0 / p.n_movers;
          unsigned int seen;
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
          } while (seen < (unsigned int)p.n_movers);
        }
        asm volatile("fence.proxy.async;" ::: "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(in_rows)),
                     "l"(p.stage + (size_t)row0 * N), "r"((uint32_t)(2 * N * sizeof(double))), "r"(smem_u32(ld_bar))
                     : "memory");
      } else {
        const int np = N >> p.nvp_shift;
#pragma unroll 1
        for (int j = 0; j < np; j++) {
          const double* src = p.in_peer[j] + (size_t)(p.row0_global + row0) * nvp;
#pragma unroll
          for (int s2 = 0; s2 < 2; s2++)
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    smem_u32(in_rows + (size_t)s2 * N + j * nvp)),
                "l"(src + s2 * nvp), "r"((uint32_t)(nvp * sizeof(double))), "r"(smem_u32(ld_bar))
                : "memory");
        }
      }
    }
    __syncthreads();  // the barrier is initialised before anybody polls it
    if (live) phase_table_fill<LOGN>(ph, alpha[0], alpha[1], t, T);
    mbar_wait(ld_bar, 0);
#pragma unroll
    for (int m = 0; m < E; m++) {
      const int e = t + T * m;
      x[m] = cmake(in_rows[e], in_rows[N + e]);
    }
  } else if (p.nvp_shift >= 0) {
    const size_t nvp = (size_t)1 << p.nvp_shift;
#pragma unroll
    for (int m = 0; m < E; m++) {
      const int e = t + T * m;
      const double* src = p.in_peer[e >> p.nvp_shift] + (size_t)(p.row0_global + row0) * nvp + (e & (nvp - 1));
      x[m] = cmake(__ldcs(src), __ldcs(src + nvp));
    }
  } else {
#pragma unroll
    for (int m = 0; m < E; m++) {
      const int e = t + T * m;
      x[m] = cmake(__ldcs(a_in + e), __ldcs(b_in + e));
    }
  }
  if (!PEER && live) phase_table_fill<LOGN>(ph, alpha[0], alpha[1], t, T);  // sincos latency hides behind the loads in flight
  fft_forward<LOGN, 1, TW>(x, buf, p.tw, t, p.zero);
  half_spectrum_update<LOGN, 1>(x, buf, ph, t);
  fft_forward<LOGN, 1, TW>(x, buf, p.tw + p.zero, t, p.zero);

  // registers (e = t + T m) -> chunk-padded rows (cell i at i + i/16); the row buffers alias the exchange buffer
  constexpr int ROW_STRIDE = N + T;  // doubles; (N + T) * 8 bytes is a multiple of 1024 for N >= 2048
  double* rowA = reinterpret_cast<double*>(buf);
  double* rowB = rowA + ROW_STRIDE;
  __syncthreads();  // every thread is done reading the exchange buffer and the phase tables
  if (live) {
#pragma unroll
    for (int m = 0; m < E; m++) {
      const int e = t + T * m;
      rowA[e + (e >> 4)] = x[m].y;
      rowB[e + (e >> 4)] = x[m].x;
    }
  }
  __syncthreads();

  int parity = 0;
  const double vc = __ldg(p.v + 16 * t);
  // NOTE: spare threads (T < 32) would corrupt shared state in fp_row_fast; the launcher only uses this kernel for T >= 32
  // (the moments of both rows behind one barrier + the collision frequencies fetched at kernel start were measured:
  // 141.8 -> 144.3 us, the longer live ranges spill; reverted)
#pragma unroll 1
  for (int s = 0; s < 2; s++) {
    double* row = rowA + s * ROW_STRIDE;
    fp_row_fast<16, TMA_OUT, CC>(row, red, pcr, parity, t, T, N, vc, p.dv, p.dt_fp,
                                 __dmul_rn(p.trow ? p.trow[TROW_NU_FP] : p.nu_fp_scale, p.nu_fp[row0 + s]), p.model);
    if (TMA_OUT && threadIdx.x == 0) {  // the barrier that ends fp_row_fast ordered every thread's fenced stores
      if constexpr (PEER) {  // one box {16, nvp/16} per owning rank
        const int nvp = 1 << p.nvp_shift, np = N >> p.nvp_shift;
#pragma unroll 1
        for (int j = 0; j < np; j++)
          tma_store_2d(&out_maps[j], row + (size_t)j * nvp, 0, (int)((p.row0_global + row0 + s) * (nvp / 16)));
      } else {
        constexpr int BOX = K::OUT_BOX_ROWS;
#pragma unroll 1
        for (int bx = 0; bx < (N / 16) / BOX; bx++)
          tma_store_2d(out_maps, row + (size_t)bx * BOX * 16, 0, (int)((row0 + s) * (N / 16)) + bx * BOX);
      }
      tma_commit_group();
    }
  }
  if constexpr (TMA_OUT) {
    if (threadIdx.x == 0) tma_wait_read_all();  // shared memory must outlive the TMA reads
  } else {
    if (p.nvp_shift >= 0) {  // the transpose back to the v-sharded layout rides on the stores (peer memory)
      const size_t nvp = (size_t)1 << p.nvp_shift;
      for (int i = t; i < N; i += T) {
        double* dst = p.out_peer[i >> p.nvp_shift] + (size_t)(p.row0_global + row0) * nvp + (i & (nvp - 1));
        dst[0] = rowA[i + (i >> 4)];
        dst[nvp] = rowB[i + (i >> 4)];
      }
    } else {
      double* a_out = p.fout + row0 * N;
      double* b_out = a_out + N;
      for (int i = t; i < N; i += T) {
        __stcs(a_out + i, rowA[i + (i >> 4)]);
        __stcs(b_out + i, rowB[i + (i >> 4)]);
      }
    }
  }
}

template <int LOGN, bool TMA_OUT, bool CC, int TW = 0>
__global__ void __launch_bounds__(VrowCfg<LOGN>::THREADS, (VrowCfg<LOGN>::THREADS <= 256 ? 2 : 1))
    vpush_collide_kernel(const __grid_constant__ CUtensorMap out_map, VrowArgs p) {
  vrow_body<LOGN, TMA_OUT, CC, TW, false>(&out_map, p, (long long)blockIdx.x);
}

// Mover CTA of the staged peer mode (VrowArgs::stage): one thread drives a two-slot pipeline of whole rows,
// L(i) = P bulk loads of row m + i n_movers from the owning ranks into slot i & 1, S(i) = one bulk store of the slot
// into the stage; round_ctr[i] is raised when S(i) has completed.
template <int LOGN>
__device__ __forceinline__ void vrow_mover(const VrowArgs& p) {
  constexpr int N = 1 << LOGN;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  if (threadIdx.x != 0) return;
  double* slots = reinterpret_cast<double*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)2 * N * sizeof(double));
  mbar_init(&bars[0], 1);
  mbar_init(&bars[1], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const size_t nvp = (size_t)1 << p.nvp_shift;
  const int np = N >> p.nvp_shift, nm = p.n_movers, m = blockIdx.x;
  const int rounds = (int)(2 * p.npairs / nm);
  for (int i = 0; i <= rounds + 1; i++) {
    if (i < rounds) {  // L(i): the slot's previous row (i - 2) must have been read out by its store
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      const int sl = i & 1;
      const long long row = m + (long long)i * nm;
      mbar_expect_tx(&bars[sl], (uint32_t)(N * sizeof(double)));
#pragma unroll 1
      for (int j = 0; j < np; j++)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(slots + (size_t)sl * N + j * nvp)),
                     "l"(p.in_peer[j] + (size_t)(p.row0_global + row) * nvp), "r"((uint32_t)(nvp * sizeof(double))),
                     "r"(smem_u32(&bars[sl]))
                     : "memory");
    }
    if (i >= 1 && i - 1 < rounds) {  // S(i - 1)
      const int sl = (i - 1) & 1;
      mbar_wait(&bars[sl], (uint32_t)(((i - 1) >> 1) & 1));
      const long long row = m + (long long)(i - 1) * nm;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p.stage + (size_t)row * N),
                   "r"(smem_u32(slots + (size_t)sl * N)), "r"((uint32_t)(N * sizeof(double)))
                   : "memory");
      tma_commit_group();
    }
    if (i >= 2) {  // S(i - 2) has completed (at most S(i - 1) may still be pending): its round may start
      if (i - 1 < rounds)
        asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
      else
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      asm volatile("fence.proxy.async;" ::: "memory");
      __threadfence();
      atomicAdd(p.round_ctr + (i - 2), 1u);
    }
  }
}

template <int LOGN, bool CC>
__global__ void __launch_bounds__(VrowCfg<LOGN>::THREADS, (VrowCfg<LOGN>::THREADS <= 256 ? 2 : 1))
    vpush_collide_peer_kernel(const __grid_constant__ PeerMaps maps, VrowArgs p) {
  if (p.stage && p.n_movers > 0 && (int)blockIdx.x < p.n_movers) {  // the lowest block indices are dispatched first: movers are resident
    vrow_mover<LOGN>(p);                          // before any compute CTA can wait for them
    return;
  }
  vrow_body<LOGN, true, CC, 0, true>(maps.m, p, (long long)blockIdx.x - ((p.stage && p.n_movers > 0) ? p.n_movers : 0));
}

template <int LOGN, bool TMA_OUT, bool CC, int TW = 0>
static int launch_vrow(const VrowArgs& p, cudaStream_t stream) {
  using K = VrowCfg<LOGN>;
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = vpush_collide_kernel<LOGN, TMA_OUT, CC, TW>;
  if (dev < 64 && !configured[dev]) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM);
    if (err != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(vpush_collide, smem=%zu): %s", K::SMEM, cudaGetErrorString(err));
      return ADEPT_ERR_CUDA;
    }
    configured[dev] = true;
  }
  CUtensorMap map = {};
  if (TMA_OUT) {
    const int rc = encode_map_2d(&map, p.fout, 16, (unsigned long long)p.npairs * 2 * (K::N / 16), 128, 16,
                                 K::OUT_BOX_ROWS, 1);
    if (rc != ADEPT_OK) return rc;
  }
  ProfileScope prof("vpush_collide", stream);
  kern<<<(unsigned)p.npairs, K::THREADS, K::SMEM, stream>>>(map, p);
  return check_launch("vpush_collide_kernel");
}

template <int LOGN, bool CC>
static int launch_vrow_peer(const VrowArgs& p, cudaStream_t stream) {
  using K = VrowCfg<LOGN>;
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = vpush_collide_peer_kernel<LOGN, CC>;
  if (dev < 64 && !configured[dev]) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM);
    if (err != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(vpush_collide_peer, smem=%zu): %s", K::SMEM, cudaGetErrorString(err));
      return ADEPT_ERR_CUDA;
    }
    configured[dev] = true;
  }
  PeerMaps maps = {};
  const int nvp = 1 << p.nvp_shift, np = K::N >> p.nvp_shift;
  for (int j = 0; j < np; j++) {
    const int rc = encode_map_2d(&maps.m[j], p.out_peer[j], 16, (unsigned long long)p.nx_global * (nvp / 16), 128, 16,
                                 nvp / 16, 1);
    if (rc != ADEPT_OK) return rc;
  }
  unsigned grid = (unsigned)p.npairs;
  if (p.stage && p.n_movers > 0) {
    const int rounds = (int)(2 * p.npairs / p.n_movers);
    cudaError_t err = cudaMemsetAsync(p.round_ctr, 0, (size_t)rounds * sizeof(unsigned int), stream);
    if (err != cudaSuccess) {
      set_last_error("vpush_collide(peer, staged): cudaMemsetAsync: %s", cudaGetErrorString(err));
      return ADEPT_ERR_CUDA;
    }
    grid += (unsigned)p.n_movers;
  }
  ProfileScope prof("vpush_collide", stream);
  kern<<<grid, K::THREADS, K::SMEM, stream>>>(maps, p);
  return check_launch("vpush_collide_peer_kernel");
}

// Bulk transfers or per-thread peer loads / stores?  Measured on the 4096^2 grid (r02j, r02q): 2 ranks 237 us per step
// with bulk transfers against 244 us, 4 ranks 178 us against 164 us (three 8 KB segments per row and peer do not make
// up for the single issuing thread).  Default: bulk transfers for two ranks only; ADEPT_B200_PEER_TMA=0 / 1 forces one.
static bool peer_tma_enabled(int n_peers = 2) {
  static int mode = -2;
  if (mode == -2) {
    const char* e = getenv("ADEPT_B200_PEER_TMA");
    mode = e ? (atoi(e) != 0 ? 1 : 0) : -1;
  }
  return mode == -1 ? n_peers == 2 : mode == 1;
}

// TMA output needs 1024-byte aligned row buffers ((nv + nv/16) * 8 bytes apart: nv >= 2048), 16-byte aligned f_out
// and row coordinates that fit the tensor map's int32 coordinates
template <int LOGN>
static int launch_vrow_auto(const VrowArgs& p, cudaStream_t stream) {
  using K = VrowCfg<LOGN>;
  const bool tma_ok = p.nvp_shift < 0 && LOGN >= 11 && tma_available() && (reinterpret_cast<uintptr_t>(p.fout) & 15) == 0 &&
                      (unsigned long long)p.npairs * 2 * (K::N / 16) < (1ull << 31);
  const bool cc = p.scheme == FP_CHANG_COOPER;
  if constexpr (LOGN >= 11) {  // peer mode with bulk transfers: whole 1 KB-multiple row segments of at most 256 chunks
    if (p.nvp_shift >= 7 && p.nvp_shift <= 12 && tma_available() && (p.stage || peer_tma_enabled(K::N >> p.nvp_shift)) &&
        (unsigned long long)p.nx_global * ((1ull << p.nvp_shift) / 16) < (1ull << 31)) {
      bool aligned = true;
      for (int j = 0; j < (K::N >> p.nvp_shift); j++)
        aligned = aligned && ((reinterpret_cast<uintptr_t>(p.in_peer[j]) | reinterpret_cast<uintptr_t>(p.out_peer[j])) & 15) == 0;
      if (aligned) return cc ? launch_vrow_peer<LOGN, true>(p, stream) : launch_vrow_peer<LOGN, false>(p, stream);
    }
  }
  if constexpr (LOGN == 12) {  // two twiddle loads per pass (fft_core.cuh): 144.3 -> 142.1 us at 4096^2 (r02i);
    static int vtw = -1;        // ADEPT_B200_VTW=0 selects the six-load passes for A/B timing
    if (vtw < 0) {
      const char* e = getenv("ADEPT_B200_VTW");
      vtw = (e && atoi(e) == 0) ? 0 : 1;
    }
    if (tma_ok && vtw == 1)
      return cc ? launch_vrow<LOGN, true, true, 1>(p, stream) : launch_vrow<LOGN, true, false, 1>(p, stream);
  }
  if constexpr (LOGN >= 11) {
    if (tma_ok) return cc ? launch_vrow<LOGN, true, true>(p, stream) : launch_vrow<LOGN, true, false>(p, stream);
  }
  return cc ? launch_vrow<LOGN, false, true>(p, stream) : launch_vrow<LOGN, false, false>(p, stream);
}

bool vpush_collide_supported(int nx, int nv, int model, int scheme, int nodrag) {
  if (nx < 2 || (nx & 1)) return false;
  if (nv < 512 || nv > 8192 || (nv & (nv - 1))) return false;  // T = nv/16 >= 32 threads, power-of-two FFT
  return (scheme == FP_CENTRAL || scheme == FP_CHANG_COOPER) && (model == FP_LB || model == FP_DOUGHERTY) && !nodrag;
}

int vpush_collide_f64(const double* fin, double* fout, int batch, int nx, int nv, const double* e, const double* dex,
                      const double* pond, double q, double m, double dt, double k1v, const double* v, double dv,
                      const double* nu_fp, double nu_fp_scale, int model, int scheme, cudaStream_t stream,
                      const double* const* in_peers, double* const* out_peers, int n_peers, long long row0_global,
                      double dt_fp, double* stage, unsigned int* round_ctr, int n_movers, long long nx_global) {
  if (batch < 1 || !vpush_collide_supported(nx, nv, model, scheme, 0)) {
    set_last_error("vpush_collide: unsupported shape batch=%d nx=%d nv=%d / model=%d scheme=%d", batch, nx, nv, model,
                   scheme);
    return ADEPT_ERR_UNSUPPORTED;
  }
  int logn = 0;
  while ((1 << logn) < nv) logn++;
  VrowArgs p = {};
  p.fin = fin, p.fout = fout, p.npairs = (long long)batch * nx / 2, p.nx = nx;
  p.e = e, p.dex = dex, p.pond = pond, p.q = q, p.m = m, p.dt = dt, p.k1 = k1v;
  p.dt_fp = dt_fp > 0.0 ? dt_fp : dt;
  p.tw = get_twiddles(logn), p.zero = 0;
  p.v = v, p.dv = dv, p.nu_fp = nu_fp, p.nu_fp_scale = nu_fp_scale, p.model = model, p.scheme = scheme;
  p.trow = current_time_row();
  if (!p.tw) return ADEPT_ERR_CUDA;
  p.nvp_shift = -1;
  if (in_peers || out_peers) {
    if (!in_peers || !out_peers || batch != 1 || n_peers < 1 || n_peers > 8 || (n_peers & (n_peers - 1)) ||
        nv % n_peers) {
      set_last_error("vpush_collide(peer mode): needs batch == 1, both pointer tables and a power-of-two number of "
                     "peers <= 8 (got %d)", n_peers);
      return ADEPT_ERR_BAD_ARG;
    }
    int sh = 0;
    while ((nv / n_peers) >> (sh + 1)) sh++;
    p.nvp_shift = sh, p.row0_global = row0_global, p.nx_global = nx_global > 0 ? nx_global : (long long)nx * n_peers;
    if (row0_global < 0 || row0_global + nx > p.nx_global) {
      set_last_error("vpush_collide(peer mode): rows [%lld, %lld) outside the %lld rows of the grid", row0_global,
                     row0_global + nx, p.nx_global);
      return ADEPT_ERR_BAD_ARG;
    }
    if (stage) {
      const bool movers_ok = n_movers == 0 || (round_ctr && n_movers >= 2 && !(n_movers & 1) && nx % n_movers == 0);
      if (!movers_ok || nv < 2048 || !tma_available() || sh < 7) {
        set_last_error("vpush_collide(peer, staged): needs round counters, an even number of movers dividing the %d "
                       "local rows (got %d), nv >= 2048 and the bulk-transfer peer path", nx, n_movers);
        return ADEPT_ERR_UNSUPPORTED;
      }
      p.stage = stage, p.round_ctr = round_ctr, p.n_movers = n_movers;
    }
    for (int j = 0; j < n_peers; j++) p.in_peer[j] = in_peers[j], p.out_peer[j] = out_peers[j];
  }
  switch (logn) {
    case 9: return launch_vrow_auto<9>(p, stream);
    case 10: return launch_vrow_auto<10>(p, stream);
    case 11: return launch_vrow_auto<11>(p, stream);
    case 12: return launch_vrow_auto<12>(p, stream);
    case 13: return launch_vrow_auto<13>(p, stream);
  }
  return ADEPT_ERR_UNSUPPORTED;
}

}  // namespace adept
