// Micro-benchmark: x-pencil tiles [nx rows x W doubles] moved global -> shared -> global with 2-D TMA boxes {W, 256}.
// Answers: can narrow (16/32/64-byte inner) TMA boxes sustain HBM bandwidth for the strided x-advection access?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_pencil tma_pencil.cu   (no -lcuda needed)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}

template <int W, int TOUCH>
__global__ void __launch_bounds__(256) pencil_copy(const __grid_constant__ CUtensorMap in_map,
                                                   const __grid_constant__ CUtensorMap out_map, int nx, int ngroups) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  double* tile = reinterpret_cast<double*>(smem);
  const int nbox = nx / 256;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t phase = 0;
  for (int g = blockIdx.x; g < ngroups; g += gridDim.x) {
    if (threadIdx.x == 0) {
      mbar_expect_tx(&bar, (uint32_t)(nx * W * sizeof(double)));
      for (int b = 0; b < nbox; b++) tma_load_2d(tile + (size_t)b * 256 * W, &in_map, &bar, g * W, b * 256);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    if (TOUCH) {  // every thread touches its share (like an FFT pass would) so the data really moved through smem
      for (int i = threadIdx.x; i < nx * W; i += 256) tile[i] = tile[i] * 1.0000001;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int b = 0; b < nbox; b++) tma_store_2d(&out_map, tile + (size_t)b * 256 * W, g * W, b * 256);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

static EncodeFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) { printf("no cuTensorMapEncodeTiled\n"); exit(1); }
  return (EncodeFn)fn;
}

template <int W, int TOUCH>
void run(EncodeFn enc, double* din, double* dout, int nx, int nv, int ctas_per_sm, bool persistent) {
  CUtensorMap in_map, out_map;
  cuuint64_t dims[2] = {(cuuint64_t)nv, (cuuint64_t)nx};
  cuuint64_t strides[1] = {(cuuint64_t)nv * sizeof(double)};
  cuuint32_t box[2] = {(cuuint32_t)W, 256};
  cuuint32_t estr[2] = {1, 1};
  for (int i = 0; i < 2; i++) {
    CUresult r = enc(i ? &out_map : &in_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, i ? (void*)dout : (void*)din, dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d (W=%d)\n", (int)r, W); return; }
  }
  const int ngroups = nv / W;
  const size_t smem = (size_t)nx * W * sizeof(double);
  auto kern = pencil_copy<W, TOUCH>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = persistent ? 148 * ctas_per_sm : ngroups;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e9;
  for (int it = 0; it < 6; it++) {
    CK(cudaMemsetAsync(dout, 0, (size_t)nx * nv * 8));  // also evicts part of L2
    CK(cudaEventRecord(e0));
    kern<<<grid, 256, smem>>>(in_map, out_map, nx, ngroups);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (it > 0 && ms < best) best = ms;
  }
  printf("W=%d (%3zu-byte rows) touch=%d %s grid=%5d smem=%3zu KB: %7.1f us  %6.0f GB/s\n", W, W * sizeof(double), TOUCH,
         persistent ? "persistent" : "one-shot  ", grid, smem / 1024, best * 1e3, 2.0 * nx * nv * 8 / (best * 1e-3) / 1e9);
}

int main() {
  const int nx = 4096, nv = 4096;
  double *din, *dout;
  CK(cudaMalloc(&din, (size_t)nx * nv * 8));
  CK(cudaMalloc(&dout, (size_t)nx * nv * 8));
  CK(cudaMemset(din, 0, (size_t)nx * nv * 8));
  EncodeFn enc = get_encode();
  run<2, 0>(enc, din, dout, nx, nv, 3, false);
  run<2, 0>(enc, din, dout, nx, nv, 3, true);
  run<2, 1>(enc, din, dout, nx, nv, 3, false);
  run<4, 0>(enc, din, dout, nx, nv, 1, false);
  run<4, 0>(enc, din, dout, nx, nv, 1, true);
  run<4, 1>(enc, din, dout, nx, nv, 1, false);
  run<6, 0>(enc, din, dout, nx, nv, 1, false);
  // verify one element path: copy correctness
  double h = 3.25; CK(cudaMemcpy(din + 12345, &h, 8, cudaMemcpyHostToDevice));
  run<2, 0>(enc, din, dout, nx, nv, 3, false);
  double r; CK(cudaMemcpy(&r, dout + 12345, 8, cudaMemcpyDeviceToHost));
  printf("copy check: %s\n", r == h ? "ok" : "MISMATCH");
  return 0;
}
