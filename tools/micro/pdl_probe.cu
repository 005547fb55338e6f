// Probe: does programmatic dependent launch let a secondary kernel start while the primary still runs on this
// driver / device?  Primary: 64 CTAs, one of them spins ~20 us after griddepcontrol.launch_dependents.  Secondary:
// records %globaltimer at entry and after griddepcontrol.wait.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o pdl_probe pdl_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
__device__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__global__ void primary(unsigned long long* out, int spin_us) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = gt();
  if (blockIdx.x == gridDim.x - 1) {
    const unsigned long long t0 = gt();
    while (gt() - t0 < (unsigned long long)spin_us * 1000ull) {}
    if (threadIdx.x == 0) out[1] = gt();
  }
}
__global__ void secondary(unsigned long long* out) {
  if (blockIdx.x == 0 && threadIdx.x == 0) out[2] = gt();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (blockIdx.x == 0 && threadIdx.x == 0) out[3] = gt();
}
int main() {
  unsigned long long* d; cudaMalloc(&d, 4 * sizeof(unsigned long long));
  for (int pdl = 0; pdl < 2; pdl++) {
    for (int rep = 0; rep < 3; rep++) {
      cudaMemset(d, 0, 32);
      primary<<<64, 256>>>(d, 20);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(296), cfg.blockDim = dim3(256);
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr, cfg.numAttrs = pdl ? 1 : 0;
      cudaError_t err = cudaLaunchKernelEx(&cfg, secondary, d);
      if (err != cudaSuccess) printf("launch error %s\n", cudaGetErrorString(err));
      cudaDeviceSynchronize();
      unsigned long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
      printf("pdl=%d: primary start 0, primary end %+.1f us, secondary entry %+.1f us, secondary after wait %+.1f us\n", pdl,
             (double)(long long)(h[1] - h[0]) * 1e-3, (double)(long long)(h[2] - h[0]) * 1e-3, (double)(long long)(h[3] - h[0]) * 1e-3);
    }
  }
  return 0;
}
