// x-advection for large grids: persistent CTAs, TMA-staged x-pencil tiles, two interleaved in-smem FFTs per tile,
// charge-density partial sums fused into the epilogue.
//
// Reference semantics: SpaceExponential.push  adept/_vlasov1d/solvers/pushers/vlasov.py:234-251, followed by the
// velocity sum of compute_charge_density  adept/_vlasov1d/solvers/pushers/field.py:197-208.
//
// f is [batch*nx, nv] row-major, so an x-pencil is strided by nv*8 bytes.  One tile = all nx rows of 4 neighbouring
// v-columns (32 bytes per row = one DRAM sector), fetched by nx/256 TMA boxes {4, 256} of a 2-D tensor map into shared
// memory as tile[row][4] = two interleaved complex sequences z_g[row] = (col 2g, col 2g+1), g = 0, 1.  (A measured
// property of the TMA unit decides the tile shape: narrow boxes are request-bound at ~0.5 rows/clk/SM, so 16-byte
// rows reach < 3 TB/s while 32-byte rows reach 4.5 TB/s; tools/micro/tma_pencil.cu.)  The CTA's 2*T threads are
// interleaved the same way (g = tid & 1, t = tid >> 1): every shared-memory access of a warp covers 16 consecutive
// slots x 2 sequences = 512 contiguous bytes, twiddle loads are shared by lane pairs, and the direct 16-byte stores of
// a warp fill whole 32-byte sectors.  The landing zone is reused in place as the (padded) Stockham exchange buffer.
// The next tile's TMA load is issued as soon as the last inverse pass has read the buffer, so it overlaps the stores.
#include <cuda.h>

#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "field_tail.cuh"
#include "internal.h"
#include "push_core.cuh"
#include "tma.cuh"
#include "tmem.cuh"

namespace adept {

struct TmaPushArgs {
  double* fout;
  int batch, nx, nv;
  int ntiles;           // batch * nv / 4
  const double* v;      // [nv]
  const double* k1_batch;
  double k1, dt;
  const cplx* tw;
  int zero;
  CUtensorMap out_map;  // f_out with the same boxes as the input map (staged output chunks)
  double* partial;      // [gridDim.x, batch*nx] per-CTA row sums of f_out, or null
  const double* filt;   // nullable [N/2+1]: real multiplier per mode (Hou-Li filter)
  FieldTail ft;         // used by the FIELD instantiation only
};

// VAR (nx = 4096 only) divides the 80 KB beside the exchange buffer between NEARLY boxes of the NEXT tile that land
// early (their TMA loads are issued at the start of the current tile instead of inside its last pass, where the load of
// 4096 box rows at ~0.5 rows/clk is exposed) and NSTAGE staged output chunks: 0 = (0, 10), 1 = (10, 0), 2 = (5, 5),
// 3 = (7, 3); 4 = (10, 0) with DEFERRED output stores: the finished tile is parked in tensor memory (64 columns per
// thread) and its direct global stores are issued a few at a time at the start of the arithmetic stretches of the NEXT
// tile's transforms (FftPass hook points), where the load/store pipe is idle and the fp64 pipe busy -- the store phase
// (8200 LSU cycles per tile with the fp64 pipe idle, 17 % of the stall samples in r02e) disappears from the timeline.
template <int LOGN, int VAR = 0>
struct TmaCfg {
  using C = FftCfg<LOGN>;
  using PC = PhaseCfg<LOGN>;
  static constexpr int N = C::N, T = C::T;
  static constexpr int THREADS = 2 * T;
  static constexpr int BOX_ROWS = N < 256 ? N : 256;
  static constexpr int NBOX = N / BOX_ROWS;
  static constexpr size_t BUF_BYTES = (size_t)2 * C::BUF * sizeof(cplx);  // two interleaved padded buffers
  static constexpr size_t PH_BYTES = (size_t)2 * 2 * PC::PER_SEQ * sizeof(cplx);
  // nx = 4096: the per-CTA row sums live in tensor memory (16 doubles = 32 columns per thread, tmem.cuh) instead of a
  // 32 KB shared-memory accumulator: no shuffle + read-modify-write per row, and the freed shared memory stages four
  // more output chunks for TMA stores (the kernel has no matrix product, so the 256 KB of TMEM are otherwise idle)
  static constexpr bool TMEM_ACC = (LOGN == 12);
  static constexpr bool DEFER = (LOGN == 12 && VAR == 4);
  static constexpr int TW = (LOGN == 12 && VAR == 5) ? 1 : 0;  // 5 = (10, 0) with two twiddle loads per pass (fft_core.cuh)
  // 4 warps share a lane quadrant: 32 columns of row sums each, + 64 columns of parked outputs with DEFER
  static constexpr int TMEM_COLS = DEFER ? 512 : 128;
  static constexpr int TMEM_WARP_COLS = TMEM_COLS / 4;
  static constexpr size_t ACC_BYTES = TMEM_ACC ? 0 : (size_t)N * sizeof(double);
  // Output chunks (one TMA box {4, 256} = 8 KB each) that leave through a shared-memory staging area and TMA tensor
  // stores instead of direct 16-byte stores.  A warp's direct store touches 16 lines (~2 L1 cycles each), so the store
  // phase of a 4096-row tile costs ~8200 LSU cycles; the TMA unit drains a box in ~512 cycles on its own.  The staging
  // area cannot hold a whole tile (227 KB per SM), so the two paths share the tile: NSTAGE chunks by TMA, the rest
  // direct, both draining concurrently.  Only where the CTA is alone on its SM anyway (nx = 4096).
  // (with the shared-memory accumulator, r01: 4 -> 141.2 us, 6 -> 137.9 us, 7 -> 139.5 us for x-push + field tail: L1
  // shrinks with the staging area; 10 chunks beside the TMEM accumulator occupy the same 223 KB as 6 did before)
  static constexpr int NEARLY = (LOGN != 12) ? 0 : ((VAR == 1 || VAR == 4 || VAR == 5) ? 10 : (VAR == 2 ? 5 : (VAR == 3 ? 7 : 0)));
  static constexpr int NSTAGE = (LOGN == 12) ? 10 - NEARLY : 0;
  static constexpr size_t BOX_BYTES = (size_t)BOX_ROWS * 4 * sizeof(double);
  static constexpr size_t BAR_OFF = BUF_BYTES + PH_BYTES + ACC_BYTES;
  static constexpr size_t EARLY_OFF = (BAR_OFF + 32 + 127) / 128 * 128;  // two mbarriers + the TMEM slot in front
  static constexpr size_t STAGE_OFF = EARLY_OFF + (size_t)NEARLY * BOX_BYTES;
  static constexpr size_t STAGE_BYTES = (size_t)NSTAGE * BOX_BYTES;
  static constexpr size_t SMEM = (NSTAGE + NEARLY) ? STAGE_OFF + STAGE_BYTES : BAR_OFF + 32;
};

// Hook handed to the FFT passes of the nx = 4096 kernel (fft_core.cuh: hook_at).  `free_fn` is the buffer-free action of
// the last inverse pass.  With `pending`, registers base .. base + 7 of the PREVIOUS tile (parked in tensor memory at
// `park`, 4 columns each, as (col, col + 1) of row t + T m) leave as direct 16-byte stores, one or two per hook point.
template <int T, class F>
struct XHook {
  static constexpr bool HAS_AT = true;
  const F& free_fn;
  uint32_t park;
  double* dst;  // previous tile: f_out at row 0 of the member, this thread's column pair
  size_t nv;
  int t, base;
  bool pending;
  __device__ __forceinline__ void operator()() const { free_fn(); }
  template <int START, int COUNT>
  __device__ __forceinline__ void drain() const {
    if constexpr (COUNT == 2) {
      double r[4];
      tmem_ld4(park + 4 * (base + START), r);
      *reinterpret_cast<double2*>(dst + (size_t)(t + T * (base + START)) * nv) = make_double2(r[0], r[1]);
      *reinterpret_cast<double2*>(dst + (size_t)(t + T * (base + START + 1)) * nv) = make_double2(r[2], r[3]);
    } else {
      double r[2];
      tmem_ld2(park + 4 * (base + START), r);
      *reinterpret_cast<double2*>(dst + (size_t)(t + T * (base + START)) * nv) = make_double2(r[0], r[1]);
    }
  }
  template <int P, int W>
  __device__ __forceinline__ void at() const {
    if (!pending) return;
    // 8 registers over the six stretches of a three-pass transform; two where the twiddled column butterflies follow
    if constexpr (P == 0 && W == 0) drain<0, 1>();
    if constexpr (P == 0 && W == 1) drain<1, 1>();
    if constexpr (P == 1 && W == 0) drain<2, 2>();
    if constexpr (P == 1 && W == 1) drain<4, 1>();
    if constexpr (P == 2 && W == 0) drain<5, 2>();
    if constexpr (P == 2 && W == 1) drain<7, 1>();
  }
};
struct NoFree {
  __device__ __forceinline__ void operator()() const {}
};

template <int LOGN, bool FIELD = false, int VAR = 0>
__global__ void __launch_bounds__(TmaCfg<LOGN>::THREADS, 1)
    vdfdx_tma_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ TmaPushArgs p) {
  using K = TmaCfg<LOGN, VAR>;
  using C = FftCfg<LOGN>;
  using PC = PhaseCfg<LOGN>;
  constexpr int N = C::N, E = C::E, T = C::T;
  static_assert(E == 16, "TMA x-advection needs nx >= 16");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cplx* tile = reinterpret_cast<cplx*>(smem_raw);                                   // landing zone / exchange buffer
  cplx* ph_all = reinterpret_cast<cplx*>(smem_raw + K::BUF_BYTES);                  // [2 groups][2 seq][PER_SEQ]
  double* rho_acc = reinterpret_cast<double*>(smem_raw + K::BUF_BYTES + K::PH_BYTES);  // [N] row sums of this CTA
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + K::BAR_OFF);
  double* stage = reinterpret_cast<double*>(smem_raw + K::STAGE_OFF);  // [NSTAGE][BOX_ROWS][4]

  const int tid = threadIdx.x;
  const int g = tid & 1, t = tid >> 1;
  cplx* buf = tile + g;
  cplx* ph = ph_all + g * 2 * PC::PER_SEQ;
  const int tiles_per_member = p.nv >> 2;

  uint64_t* bar_early = bar + 1;
  const cplx* early = reinterpret_cast<const cplx*>(smem_raw + K::EARLY_OFF);  // boxes 0 .. NEARLY-1 of the tile: [row][2 cplx]
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar_early, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t racc_addr = 0, tmem_base = 0;
  if constexpr (K::TMEM_ACC) {
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
    if (tid < 32) tmem_alloc(tmem_slot, K::TMEM_COLS);
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    tmem_base = *tmem_slot;
    racc_addr = tmem_addr(tmem_base, tid >> 5, K::TMEM_WARP_COLS * (tid >> 7));
    const double zeros[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    tmem_st8(racc_addr, zeros);
    tmem_st8(racc_addr + 16, zeros);
    tmem_wait_st();
  } else {
    if (p.partial) {
      for (int i = tid; i < N; i += K::THREADS) rho_acc[i] = 0.0;
    }
    __syncthreads();
  }

  auto issue_load = [&](int tl) {  // the boxes that land in the exchange buffer (all of them when NEARLY == 0)
    const int b = tl / tiles_per_member, cg = tl - b * tiles_per_member;
    mbar_expect_tx(bar, (uint32_t)((K::NBOX - K::NEARLY) * K::BOX_BYTES));
#pragma unroll 1
    for (int bx = K::NEARLY; bx < K::NBOX; bx++)
      tma_load_2d(reinterpret_cast<double*>(tile) + (size_t)bx * K::BOX_ROWS * 4, &in_map, bar, cg * 4,
                  b * N + bx * K::BOX_ROWS);
  };
  auto issue_early = [&](int tl) {  // boxes 0 .. NEARLY-1 into their own landing zone
    if constexpr (K::NEARLY > 0) {
      const int b = tl / tiles_per_member, cg = tl - b * tiles_per_member;
      mbar_expect_tx(bar_early, (uint32_t)(K::NEARLY * K::BOX_BYTES));
#pragma unroll 1
      for (int bx = 0; bx < K::NEARLY; bx++)
        tma_load_2d(reinterpret_cast<double*>(smem_raw + K::EARLY_OFF) + (size_t)bx * K::BOX_ROWS * 4, &in_map, bar_early,
                    cg * 4, b * N + bx * K::BOX_ROWS);
    }
  };
  auto flush_rho = [&](int b) {  // this CTA owns row blockIdx.x of `partial`: plain read-modify-write
    double* dst = p.partial + ((size_t)blockIdx.x * p.batch + b) * N;
    if constexpr (K::TMEM_ACC) {  // rows t + T m of both lanes of a pair: sum across the pair, the even lane writes
      const double zeros[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int h = 0; h < 2; h++) {
        double r[8];
        tmem_ld8(racc_addr + 16 * h, r);
#pragma unroll
        for (int m = 0; m < 8; m++) {
          const double s = r[m] + __shfl_xor_sync(0xffffffffu, r[m], 1);
          double* d = dst + t + T * (8 * h + m);
          if (g == 0) *d = (p.batch == 1) ? s : *d + s;
        }
        tmem_st8(racc_addr + 16 * h, zeros);
      }
      tmem_wait_st();
      return;
    }
    __syncthreads();
    for (int i = tid; i < N; i += K::THREADS) {
      // a single member is visited once by every CTA: plain store, no zero-initialisation needed
      dst[i] = (p.batch == 1) ? rho_acc[i] : dst[i] + rho_acc[i];
      rho_acc[i] = 0.0;
    }
    __syncthreads();
  };

  int tl = blockIdx.x;
  if (tid == 0 && tl < p.ntiles) issue_early(tl), issue_load(tl);
  uint32_t parity = 0;
  int cur_b = -1;
  const uint32_t park_addr = racc_addr + 32;  // DEFER: 64 columns behind the row sums
  double* dst_prev = nullptr;                 // DEFER: output base of the tile parked in tensor memory, if any
  for (; tl < p.ntiles; tl += gridDim.x) {
    const int b = tl / tiles_per_member, cg = tl - b * tiles_per_member;
    if (p.partial && b != cur_b) {
      if (cur_b >= 0) flush_rho(cur_b);
      cur_b = b;
    }
    const int col = 4 * cg + 2 * g;
    {
      const double k1 = p.k1_batch ? p.k1_batch[b] : p.k1;
      const double alpha_a = k1 * (p.v[col] * p.dt);
      const double alpha_b = k1 * (p.v[col + 1] * p.dt);
      phase_table_fill<LOGN>(ph, alpha_a, alpha_b, t, T);
    }
    if constexpr (K::NEARLY > 0) mbar_wait(bar_early, parity);
    mbar_wait(bar, parity);
    parity ^= 1;

    cplx x[E];
#pragma unroll
    for (int m = 0; m < E; m++)  // unpadded landing layout; chunk m = box m (T == BOX_ROWS when NEARLY > 0)
      x[m] = (m < K::NEARLY) ? early[(t + T * m) * 2 + g] : buf[(t + T * m) * 2];
    if constexpr (K::NEARLY > 0) {  // the early zone has been read by everybody: the next tile's first boxes may land
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid == 0 && tl + (int)gridDim.x < p.ntiles) issue_early(tl + gridDim.x);
    }
    if constexpr (K::DEFER) {
      const NoFree nofree;
      const XHook<T, NoFree> hk{nofree, park_addr, dst_prev, (size_t)p.nv, t, 0, dst_prev != nullptr};
      fft_forward<LOGN, 2>(x, buf, p.tw, t, p.zero, hk);
    } else {
      fft_forward<LOGN, 2, K::TW>(x, buf, p.tw, t, p.zero);
    }
    half_spectrum_update<LOGN, 2>(x, buf, ph, t, p.filt);
    // the exchange buffer is dead once every thread has read its inputs of the last inverse pass: the next tile's TMA
    // load is issued from inside that pass, so it also overlaps the second half of the butterflies (not only the stores)
    auto buffer_free = [&]() {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      // the staging area is written again below: the previous tile's TMA stores (issued a whole tile ago) must have
      // finished reading it
      if (K::NSTAGE > 0 && tid == 0) tma_wait_read_all();
      __syncthreads();
      if (tid == 0 && tl + (int)gridDim.x < p.ntiles) issue_load(tl + gridDim.x);
    };
    if constexpr (K::DEFER) {
      const XHook<T, decltype(buffer_free)> hk{buffer_free, park_addr, dst_prev, (size_t)p.nv, t, 8, dst_prev != nullptr};
      fft_forward<LOGN, 2>(x, buf, p.tw + p.zero, t, p.zero, hk);
    } else {
      fft_forward<LOGN, 2, K::TW>(x, buf, p.tw + p.zero, t, p.zero, buffer_free);
    }

    // global stores (local/global queue) interleaved with the row-sum accumulation (shuffle + shared-memory queue)
    double* dst = p.fout + ((size_t)b * N) * p.nv + col;
    const bool want_rho = p.partial != nullptr;
    if constexpr (K::NSTAGE > 0) {
      // rows t + T m of chunk m (T == BOX_ROWS): dense [row][4] boxes, 512 contiguous bytes per warp
      static_assert(K::NSTAGE == 0 || T == K::BOX_ROWS, "one output chunk per register index");
#pragma unroll
      for (int m = 0; m < K::NSTAGE; m++)
        *reinterpret_cast<double2*>(stage + ((size_t)m * K::BOX_ROWS + t) * 4 + 2 * g) = make_double2(x[m].y, x[m].x);
      fence_async_smem();
      __syncthreads();
      if (tid == 0) {
#pragma unroll
        for (int m = 0; m < K::NSTAGE; m++)
          tma_store_2d(&p.out_map, stage + (size_t)m * K::BOX_ROWS * 4, cg * 4, b * N + m * K::BOX_ROWS);
        tma_commit_group();
      }
    }
    // DEFER: every tile but the CTA's last is parked in tensor memory and stored during the next tile's transforms
    const bool park_tile = K::DEFER && tl + (int)gridDim.x < p.ntiles;
    if constexpr (K::DEFER) {
      dst_prev = nullptr;
      if (park_tile) {
#pragma unroll
        for (int h = 0; h < 4; h++) {
          const double r[8] = {x[4 * h].y,     x[4 * h].x,     x[4 * h + 1].y, x[4 * h + 1].x,
                               x[4 * h + 2].y, x[4 * h + 2].x, x[4 * h + 3].y, x[4 * h + 3].x};
          tmem_st8(park_addr + 16 * h, r);
        }
        tmem_wait_st();
        dst_prev = dst;
      }
    }
#pragma unroll
    for (int m = 0; m < E; m++) {
      const size_t e = t + T * m;
      if (m >= K::NSTAGE && !park_tile) *reinterpret_cast<double2*>(dst + e * p.nv) = make_double2(x[m].y, x[m].x);
      if constexpr (!K::TMEM_ACC) {
        if (want_rho) {
          double s = x[m].y + x[m].x;
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          if (g == 0) rho_acc[t + T * m] += s;  // row t + T m is owned by this lane pair
        }
      }
    }
    if constexpr (K::TMEM_ACC) {
      if (want_rho) {  // this thread's share of rows t + T m: accumulated in its own TMEM columns
#pragma unroll
        for (int h = 0; h < 2; h++) {
          double r[8];
          tmem_ld8(racc_addr + 16 * h, r);
#pragma unroll
          for (int m = 0; m < 8; m++) r[m] += x[8 * h + m].y + x[8 * h + m].x;
          tmem_st8(racc_addr + 16 * h, r);
        }
        tmem_wait_st();
      }
    }
  }
  if (p.partial && cur_b >= 0) flush_rho(cur_b);
  // the last tile's staged stores must be complete before the CTA exits (its shared memory is their source); the field
  // tail does not touch the staging area or f_out, so it runs first and the wait costs nothing
  if (!FIELD && K::NSTAGE > 0 && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");

  if constexpr (FIELD) {
    // ---- field solve in the tail (batch == 1, one species): field_tail.cuh; the exchange buffer and the row-sum
    // accumulator are dead after the last flush and serve as its scratch
    // reduction scratch of the tail: the row-sum accumulator where it lives in shared memory, else the part of the
    // (dead) exchange buffer behind the tail's rho | green | partial-output arrays (3 N + 256 doubles)
    double* tail_red = K::TMEM_ACC ? reinterpret_cast<double*>(tile) + 3 * N + 256 : rho_acc;
    static_assert(!K::TMEM_ACC || (3 * N + 256 + (K::THREADS / 32) * 32) * sizeof(double) <= K::BUF_BYTES, "tail scratch");
    field_tail_solve<N, K::THREADS>(p.ft, p.partial, reinterpret_cast<double*>(tile), tail_red);
    if (K::NSTAGE > 0 && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  if constexpr (K::TMEM_ACC) {
    tmem_fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem_base, K::TMEM_COLS);
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
    (void)cudaGetLastError();
  }
  return fn;
}

// 2-D fp64 tensor map: dims {dim0 (contiguous), dim1}, row pitch `pitch_bytes`, box {box0, box1}; swizzle128 selects
// CU_TENSOR_MAP_SWIZZLE_128B (box0 * 8 must then be 128 bytes and the shared-memory tile 1024-byte aligned)
int encode_map_2d(CUtensorMap* map, const double* base, unsigned long long dim0, unsigned long long dim1,
                  unsigned long long pitch_bytes, unsigned box0, unsigned box1, int swizzle128) {
  EncodeTiledFn enc = get_encoder();
  if (!enc) {
    set_last_error("tma: cuTensorMapEncodeTiled is not available from the driver");
    return ADEPT_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)dim0, (cuuint64_t)dim1};
  cuuint64_t strides[1] = {(cuuint64_t)pitch_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("tma: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return ADEPT_ERR_CUDA;
  }
  return ADEPT_OK;
}

bool tma_available() { return get_encoder() != nullptr; }

static int encode_map(CUtensorMap* map, const double* base, long long rows, int nv, int box_rows) {
  return encode_map_2d(map, base, (unsigned long long)nv, (unsigned long long)rows,
                       (unsigned long long)nv * sizeof(double), 4, (unsigned)box_rows, 0);
}

template <int LOGN>
static int tma_ctas(int ntiles) {
  using K = TmaCfg<LOGN>;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_sm = (int)((227 * 1024) / (K::SMEM + 1024));
  const int by_threads = 2048 / K::THREADS;
  if (per_sm > by_threads) per_sm = by_threads;
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  const int grid = sms * per_sm;
  return grid < ntiles ? grid : ntiles;
}

template <int LOGN, bool FIELD = false, int VAR = 0>
static int launch_tma(const CUtensorMap& map, const TmaPushArgs& p, int grid, cudaStream_t stream) {
  using K = TmaCfg<LOGN, VAR>;
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = vdfdx_tma_kernel<LOGN, FIELD, VAR>;
  if (dev < 64 && !configured[dev]) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM);
    if (err != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(vdfdx_tma, smem=%zu): %s", K::SMEM, cudaGetErrorString(err));
      return ADEPT_ERR_CUDA;
    }
    configured[dev] = true;
  }
  ProfileScope prof(FIELD ? "vdfdx_tma_field" : "vdfdx_tma", stream);
  if (FIELD) {
    // cooperative launch: the device-wide barriers of the field tail need every CTA resident at once, and only this
    // launch mode guarantees it when kernels of other streams compete for the SMs (no deadlock by construction)
    void* args[2] = {const_cast<CUtensorMap*>(&map), const_cast<TmaPushArgs*>(&p)};
    cudaError_t err = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(grid), dim3(K::THREADS), args,
                                                  K::SMEM, stream);
    if (err != cudaSuccess) {
      set_last_error("cudaLaunchCooperativeKernel(vdfdx_tma + field): %s", cudaGetErrorString(err));
      (void)cudaGetLastError();
      return ADEPT_ERR_CUDA;
    }
    return check_launch("vdfdx_tma_kernel(field)");
  }
  kern<<<grid, K::THREADS, K::SMEM, stream>>>(map, p);
  return check_launch("vdfdx_tma_kernel");
}

// ADEPT_B200_XVAR = 0 .. 3 selects the early-landing / staging split of the nx = 4096 kernel (TmaCfg) for A/B timing
static int x_variant() {
  static int var = -1;
  if (var < 0) {
    const char* e = getenv("ADEPT_B200_XVAR");
    var = e ? atoi(e) : 1;  // measured (r02, x-push + field tail): 0 -> 136.4 us, 1 -> 129.9, 2 -> 132.9, 3 -> 133.2
    if (var < 0 || var > 5) var = 1;
  }
  return var;
}
template <bool FIELD>
static int launch_tma12(const CUtensorMap& map, const TmaPushArgs& p, int grid, cudaStream_t stream) {
  switch (x_variant()) {
    case 1: return launch_tma<12, FIELD, 1>(map, p, grid, stream);
    case 2: return launch_tma<12, FIELD, 2>(map, p, grid, stream);
    case 3: return launch_tma<12, FIELD, 3>(map, p, grid, stream);
    case 4: return launch_tma<12, FIELD, 4>(map, p, grid, stream);
    case 5: return launch_tma<12, FIELD, 5>(map, p, grid, stream);
    default: return launch_tma<12, FIELD, 0>(map, p, grid, stream);
  }
}

static int ilog2_exact_(int n) {
  if (n < 2 || (n & (n - 1))) return -1;
  int l = 0;
  while ((1 << l) < n) l++;
  return l;
}

// true when the TMA kernel handles this problem (otherwise callers use the direct kernel of push.cu)
bool vdfdx_tma_supported(const double* fin, const double* fout, int nx, int nv) {
  const int logn = ilog2_exact_(nx);
  if (logn < 8 || logn > 12) return false;
  if (nv % 4) return false;
  if ((reinterpret_cast<uintptr_t>(fin) | reinterpret_cast<uintptr_t>(fout)) & 15) return false;
  return get_encoder() != nullptr;
}

// ADEPT_B200_XPUSH=dual routes nx = 4096 to the two-transforms-per-thread kernel of vdfdx_dual.cu for A/B timing (it
// is parity-green but slower: 150 us against 128 us, profiles/r02a_*); both use one persistent CTA per SM, so the
// partial-sum contract is the same.
static bool use_dual(int nx) {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("ADEPT_B200_XPUSH");
    mode = (e && strcmp(e, "dual") == 0) ? 1 : 0;
  }
  return mode == 1 && vdfdx_dual_supported(nx);
}

// number of partial-sum rows (persistent CTAs) the TMA kernel uses for this shape
int vdfdx_tma_parts(int batch, int nx, int nv) {
  const int ntiles = batch * (nv / 4);
  if (use_dual(nx)) return vdfdx_dual_parts(batch, nx, nv);
  switch (ilog2_exact_(nx)) {
    case 8: return tma_ctas<8>(ntiles);
    case 9: return tma_ctas<9>(ntiles);
    case 10: return tma_ctas<10>(ntiles);
    case 11: return tma_ctas<11>(ntiles);
    case 12: return tma_ctas<12>(ntiles);
    default: return 0;
  }
}

// partial: [vdfdx_tma_parts(), batch*nx] zero-initialised by the caller (or null)
// the field tail needs every CTA of the launch co-resident (device-wide barriers), a single member and enough rows of
// the exchange buffer for rho + green (2 nx doubles <= the tile)
template <int LOGN>
static bool field_grid_fits(int grid) {  // cached per (device, LOGN): resident CTAs of the FIELD instantiation
  using K = TmaCfg<LOGN>;
  static int resident[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 64) return false;
  if (resident[dev] == 0) {
    auto kern = vdfdx_tma_kernel<LOGN, true>;
    int per_sm = 0, sms = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, K::THREADS, K::SMEM) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
      (void)cudaGetLastError();
      resident[dev] = -1;
    } else {
      resident[dev] = per_sm * sms > 0 ? per_sm * sms : -1;
    }
  }
  return resident[dev] > 0 && grid <= resident[dev];
}

bool vdfdx_tma_field_supported(int batch, int nx, int nv) {
  if (batch != 1 || nv % 4 != 0 || get_encoder() == nullptr) return false;
  const int grid = vdfdx_tma_parts(batch, nx, nv);
  switch (nx) {  // the device-wide barriers need the whole grid resident at once
    case 1024: return field_grid_fits<10>(grid);
    case 2048: return field_grid_fits<11>(grid);
    case 4096: return field_grid_fits<12>(grid);
    default: return false;
  }
}

int vdfdx_tma_f64(const double* fin, double* fout, int batch, int nx, int nv, const double* v, double dt,
                  const double* k1_batch, double k1, double* partial, cudaStream_t stream, const double* filt,
                  const FieldTail* field) {
  if (use_dual(nx)) {
    if (field && (!vdfdx_tma_field_supported(batch, nx, nv) || !partial || !field->counter || !field->green ||
                  !field->rho || !field->e || !field->a || !field->pond || field->n_ex < 0 || field->n_ex > 8)) {
      set_last_error("vdfdx(dual + field): unsupported batch=%d nx=%d nv=%d or missing buffers", batch, nx, nv);
      return ADEPT_ERR_UNSUPPORTED;
    }
    return vdfdx_dual_f64(fin, fout, batch, nx, nv, v, dt, k1_batch, k1, partial, stream, filt, field);
  }
  const int logn = ilog2_exact_(nx);
  CUtensorMap map;
  const int box_rows = nx < 256 ? nx : 256;
  int rc = encode_map(&map, fin, (long long)batch * nx, nv, box_rows);
  if (rc != ADEPT_OK) return rc;
  TmaPushArgs p = {};
  rc = encode_map(&p.out_map, fout, (long long)batch * nx, nv, box_rows);
  if (rc != ADEPT_OK) return rc;
  p.fout = fout, p.batch = batch, p.nx = nx, p.nv = nv, p.ntiles = batch * (nv / 4);
  p.v = v, p.k1_batch = k1_batch, p.k1 = k1, p.dt = dt, p.zero = 0, p.partial = partial, p.filt = filt;
  p.tw = get_twiddles(logn);
  if (!p.tw) return ADEPT_ERR_CUDA;
  const int grid = vdfdx_tma_parts(batch, nx, nv);
  if (field) {
    if (!vdfdx_tma_field_supported(batch, nx, nv) || !partial || !field->counter || !field->green || !field->rho ||
        !field->e || !field->a || !field->pond || field->n_ex < 0 || field->n_ex > 8) {
      set_last_error("vdfdx(tma + field): unsupported batch=%d nx=%d nv=%d or missing buffers", batch, nx, nv);
      return ADEPT_ERR_UNSUPPORTED;
    }
    p.ft = *field;
    switch (logn) {
      case 10: return launch_tma<10, true>(map, p, grid, stream);
      case 11: return launch_tma<11, true>(map, p, grid, stream);
      default: return launch_tma12<true>(map, p, grid, stream);
    }
  }
  switch (logn) {
    case 8: return launch_tma<8>(map, p, grid, stream);
    case 9: return launch_tma<9>(map, p, grid, stream);
    case 10: return launch_tma<10>(map, p, grid, stream);
    case 11: return launch_tma<11>(map, p, grid, stream);
    case 12: return launch_tma12<false>(map, p, grid, stream);
    default:
      set_last_error("vdfdx(tma): unsupported nx=%d", nx);
      return ADEPT_ERR_UNSUPPORTED;
  }
}

}  // namespace adept
