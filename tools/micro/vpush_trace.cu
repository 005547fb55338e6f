// Phase trace of the spectral v-push (same device code as csrc/push.cu, AXIS_V, N = 4096): every warp of a few CTAs
// records clock64() at the phase boundaries of each FFT pass, so the time between barriers can be attributed to
// exchange loads, butterflies, barrier waits and exchange stores.  Development tool (links libadept_b200.so for the
// twiddle tables).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -I adept_b200/csrc \
//        -o tools/micro/vpush_trace tools/micro/vpush_trace.cu -L adept_b200 -ladept_b200 -Xlinker -rpath='$ORIGIN/../../adept_b200'
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define TRACE_MAX 64
struct TraceRec { int id; long long clk; };
__device__ TraceRec g_trace[4][16][TRACE_MAX];
__device__ int g_trace_n[4][16];
// records go to shared memory (a global store before a barrier would add its own drain latency to the barrier)
__shared__ TraceRec trace_sh[16][TRACE_MAX];
__shared__ int trace_n_sh[16];
__shared__ int trace_slot_sh;
#define ADEPT_TRACE(ID)                                                                     \
  do {                                                                                      \
    if (trace_slot_sh >= 0 && (threadIdx.x & 31) == 0) {                                    \
      int w_ = threadIdx.x >> 5;                                                            \
      int n_ = trace_n_sh[w_];                                                              \
      if (n_ < TRACE_MAX) {                                                                 \
        trace_sh[w_][n_].id = (ID);                                                         \
        trace_sh[w_][n_].clk = clock64();                                                   \
        trace_n_sh[w_] = n_ + 1;                                                            \
      }                                                                                     \
    }                                                                                       \
  } while (0)

#include "push_core.cuh"

using namespace adept;

template <int LOGN, int PP>
__global__ void __launch_bounds__(PP ? 512 : 256, PP ? 1 : 2)
    vpush_trace_kernel(const double* fin, double* fout, const cplx* tw, int zero, double alpha_a, double alpha_b,
                       int trace_first) {
  using C = FftCfg<LOGN>;
  using PC = PhaseCfg<LOGN>;
  constexpr int N = C::N, E = C::E, T = C::T;
  constexpr int NW = PP ? 16 : 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int team = threadIdx.x / T;
  const int t = threadIdx.x % T;
  cplx* buf = reinterpret_cast<cplx*>(smem_raw) + (size_t)team * (C::BUF + 2 * PC::PER_SEQ);
  cplx* ph = buf + C::BUF;
  if (threadIdx.x == 0) {
    const int rel = (int)blockIdx.x - trace_first;
    trace_slot_sh = (rel >= 0 && rel < 4) ? rel : -1;
  }
  if (threadIdx.x < NW) trace_n_sh[threadIdx.x] = 0;
  __syncthreads();
  const long long pair = (long long)blockIdx.x * (PP ? 2 : 1) + team;
  const double* a = fin + 2 * pair * N;
  const double* b = a + N;
  cplx x[E];
#pragma unroll
  for (int m = 0; m < E; m++) x[m] = cmake(__ldcs(a + t + T * m), __ldcs(b + t + T * m));
  phase_table_fill<LOGN>(ph, alpha_a, alpha_b, t, T);
  ADEPT_TRACE(900);
  {
    fft_forward<LOGN>(x, buf, tw, t, zero);
    half_spectrum_update<LOGN, 1>(x, buf, ph, t);
    fft_forward<LOGN>(x, buf, tw + zero, t, zero);
  }
  ADEPT_TRACE(901);
  double* ao = fout + 2 * pair * N;
  double* bo = ao + N;
#pragma unroll
  for (int m = 0; m < E; m++) {
    __stcs(ao + t + T * m, x[m].y);
    __stcs(bo + t + T * m, x[m].x);
  }
  ADEPT_TRACE(902);
  __syncthreads();
  if (trace_slot_sh >= 0) {
    for (int i = threadIdx.x; i < NW * TRACE_MAX; i += blockDim.x)
      g_trace[trace_slot_sh][i / TRACE_MAX][i % TRACE_MAX] = trace_sh[i / TRACE_MAX][i % TRACE_MAX];
    if (threadIdx.x < NW) g_trace_n[trace_slot_sh][threadIdx.x] = trace_n_sh[threadIdx.x];
  }
}

template <int PP>
static void run(const double* fin, double* fout, const cplx* tw, int nx, int trace_first, int pad) {
  constexpr int LOGN = 12;
  constexpr int NW = PP ? 16 : 8;
  const size_t smem = (PP ? 2 : 1) * (FftCfg<LOGN>::BUF + 2 * PhaseCfg<LOGN>::PER_SEQ) * sizeof(cplx) + pad;
  auto kern = vpush_trace_kernel<LOGN, PP>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 3; rep++) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<<<nx / (PP ? 4 : 2), PP ? 512 : 256, smem>>>(fin, fout, tw, 0, 1e-3, 2e-3, trace_first);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(err)); exit(1); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("# rep %d: %.1f us (traced, pad=%d, pingpong=%d)\n", rep, ms * 1e3, pad, (int)PP);
  }
  static TraceRec tr[4][16][TRACE_MAX];
  static int tn[4][16];
  cudaMemcpyFromSymbol(tr, g_trace, sizeof(tr));
  cudaMemcpyFromSymbol(tn, g_trace_n, sizeof(tn));
  for (int s = 0; s < 1; s++) {
    long long t0 = tr[s][0][0].clk;
    for (int w = 0; w < NW; w++) if (tn[s][w] && tr[s][w][0].clk < t0) t0 = tr[s][w][0].clk;
    printf("## CTA %d: cycles since first record; rows = trace point id, columns = warps 0..%d (issue times)\n", trace_first + s, NW - 1);
    for (int i = 0; i < tn[s][0]; i++) {
      printf("%4d :", tr[s][0][i].id);
      for (int w = 0; w < NW; w += (PP ? 4 : 1)) printf(" %7lld", tr[s][w][i].clk - t0);
      printf("\n");
    }
  }
}

int main(int argc, char** argv) {
  constexpr int LOGN = 12, N = 1 << LOGN;
  const int nx = 4096;
  const int trace_first = argc > 1 ? atoi(argv[1]) : 600;  // a CTA in the middle of the run (steady state)
  const int pad = argc > 2 ? atoi(argv[2]) : 0;            // extra smem: 60000 forces one CTA per SM
  const int pp = argc > 3 ? atoi(argv[3]) : 0;
  double *fin, *fout;
  cudaMalloc(&fin, sizeof(double) * nx * N);
  cudaMalloc(&fout, sizeof(double) * nx * N);
  std::vector<double> h((size_t)nx * N);
  for (size_t i = 0; i < h.size(); i++) h[i] = 1.0 + 1e-3 * (double)(i % 977);
  cudaMemcpy(fin, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice);
  const cplx* tw = get_twiddles(LOGN);
  if (!tw) { printf("no twiddles\n"); return 1; }
  (void)pp;
  run<0>(fin, fout, tw, nx, trace_first, pad);
  return 0;
}
