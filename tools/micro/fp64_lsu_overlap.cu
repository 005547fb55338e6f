// Micro-benchmark: do the fp64 pipe and the shared-memory (LSU data) pipe of an sm_100a SM overlap?
// The spectral pushes issue ~44 fp64 instructions and ~1 shared-memory wavefront per cell; ncu shows fp64 ~45 % and
// LSU data pipe ~65 % busy with the sum near 100 %.  This test runs (1) fp64 only, (2) LDS/STS only, (3) both mixed in
// every warp, (4) both, warp-specialised (even warps fp64, odd warps LSU, each doing twice the share), with the same
// instruction counts, and prints cycles per iteration so that max() vs sum() behaviour is visible.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o fp64_lsu_overlap fp64_lsu_overlap.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int NF = 16;   // independent DFMA chains per thread
constexpr int NL = 4;    // LDS.128 + STS.128 pairs per iteration

// mode bit 0: fp64 work, bit 1: LSU work, bit 2: warp specialised
template <int FP_PER_ITER>
__global__ void __launch_bounds__(512, 1) kern(int mode, int iters, double* out, double seed) {
  extern __shared__ __align__(16) double2 sm[];
  const int tid = threadIdx.x, warp = tid >> 5;
  double acc[NF];
#pragma unroll
  for (int i = 0; i < NF; i++) acc[i] = seed + i + tid;
  double2 v[NL];
#pragma unroll
  for (int i = 0; i < NL; i++) v[i] = make_double2(tid, i);
  for (int i = tid; i < 512 * NL; i += 512) sm[i] = make_double2(i, -i);
  __syncthreads();
  bool do_fp = mode & 1, do_ls = mode & 2;
  int fp_rep = 1, ls_rep = 1;
  if (mode & 4) {
    do_fp = (warp & 1) == 0, do_ls = (warp & 1) == 1;
    fp_rep = 2, ls_rep = 2;
  }
  const double a = 1.0000001, b = 1e-9;
  for (int it = 0; it < iters; it++) {
    if (do_fp) {
      for (int r = 0; r < fp_rep; r++) {
#pragma unroll
        for (int k = 0; k < FP_PER_ITER / NF; k++) {
#pragma unroll
          for (int i = 0; i < NF; i++) acc[i] = fma(acc[i], a, b);
        }
      }
    }
    if (do_ls && (mode & 8)) {  // 32 SHFL.32 per iteration instead of shared memory (1 wavefront each if they use the LSU pipe)
      for (int r = 0; r < ls_rep; r++) {
#pragma unroll
        for (int i = 0; i < NL; i++) {
          v[i].x = __shfl_xor_sync(0xffffffffu, v[i].x, 1 + i);
          v[i].y = __shfl_xor_sync(0xffffffffu, v[i].y, 5 + i);
          v[i].x = __shfl_xor_sync(0xffffffffu, v[i].x, 9 + i);
          v[i].y = __shfl_xor_sync(0xffffffffu, v[i].y, 13 + i);
        }
      }
    } else if (do_ls) {
      for (int r = 0; r < ls_rep; r++) {
#pragma unroll
        for (int i = 0; i < NL; i++) sm[i * 512 + tid] = v[i];
        asm volatile("" ::: "memory");
#pragma unroll
        for (int i = 0; i < NL; i++) v[i] = sm[i * 512 + (tid ^ 32)];  // another warp's slot: no store forwarding tricks
        asm volatile("" ::: "memory");
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NF; i++) s += acc[i];
#pragma unroll
  for (int i = 0; i < NL; i++) s += v[i].x + v[i].y;
  out[blockIdx.x * 512 + tid] = s;
}

template <int FP>
static void run(const char* label, int mode, int threads, double* out) {
  const int iters = 2000;
  auto k = kern<FP>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 512 * NL * 16));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(e0));
    k<<<148, threads, 512 * NL * 16>>>(mode, iters, out, 1.0);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double cyc = best * 1e-3 * clk_khz * 1e3 / iters;
  printf("%-44s threads=%4d  fp64/iter=%3d  lds+sts.128/iter=%d : %8.1f cycles/iter (at max clock %d MHz)\n", label, threads,
         FP, NL, cyc, clk_khz / 1000);
}

int main() {
  double* out;
  CK(cudaMalloc(&out, 148 * 512 * sizeof(double)));
  for (int threads : {512, 256}) {
    run<64>("fp64 only", 1, threads, out);
    run<64>("lsu only", 2, threads, out);
    run<64>("fp64 + lsu mixed in every warp", 3, threads, out);
    run<64>("fp64 + lsu warp-specialised (2x each)", 7, threads, out);
    run<64>("shfl only (32 SHFL.32/iter)", 2 | 8, threads, out);
    run<64>("fp64 + shfl mixed", 3 | 8, threads, out);
    run<32>("fp64 only", 1, threads, out);
    run<32>("fp64 + lsu mixed in every warp", 3, threads, out);
    run<128>("fp64 only", 1, threads, out);
    run<128>("fp64 + lsu mixed in every warp", 3, threads, out);
  }
  return 0;
}
