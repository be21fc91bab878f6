// Micro-benchmark: cost of a kernel boundary inside a CUDA graph on B200, with and without programmatic dependent launch (PDL),
// for persistent-style kernels (148 CTAs) that either fill the SM (200 KB of shared memory: the next kernel's CTAs cannot become
// resident before this kernel's CTAs exit) or leave room for a second CTA (100 KB).  Each CTA "works" for `work_cycles` after its
// dependency wait.  Per-launch time - work time = what a boundary costs in the decode graph.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o launchbench launchbench.cu && ./launchbench
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void __launch_bounds__(320, 1) k(int work_cycles, int use_pdl, int* sink) {
  extern __shared__ int sm[];
  if (use_pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) sm[0] = 1;
  __syncthreads();
  if (use_pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  const long long t0 = clock64();
  while (clock64() - t0 < work_cycles) { }
  if (sm[0] == 12345) sink[0] = 1;
}

int main() {
  int* sink; CK(cudaMalloc(&sink, 4));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaStream_t s; CK(cudaStreamCreate(&s));
  const int N = 400;
  for (int smem_kb : {200, 100, 8}) {
    for (int work : {0, 4000, 16000}) {
      for (int pdl : {0, 1}) {
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < N; ++i) {
          cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(148); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem_kb * 1024; cfg.stream = s;
          cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
          cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
          CK(cudaLaunchKernelEx(&cfg, k, work, pdl, sink));
        }
        CK(cudaStreamEndCapture(s, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
          CK(cudaEventRecord(a, s)); CK(cudaGraphLaunch(ge, s)); CK(cudaEventRecord(b, s)); CK(cudaEventSynchronize(b));
          float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
        }
        int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
        const double per = best * 1e3 / N, work_us = work / (clk_khz * 1e-3);
        printf("smem %3d KB  work %5d cycles (%.2f us)  %s: %.2f us per launch  -> boundary %.2f us\n", smem_kb, work, work_us, pdl ? "PDL   " : "no PDL",
               per, per - work_us);
        cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
      }
    }
  }
  return 0;
}
