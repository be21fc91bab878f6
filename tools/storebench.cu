// Micro-benchmark: how fast a short kernel can push its results into L2 on B200.  `grid` CTAs of 256 threads each write `bytes / grid`
// with coalesced 16-byte stores into a buffer that stays L2-resident, 200 launches in a CUDA graph with PDL; per-launch time minus
// the 0.4 us boundary (tools/launchbench.cu) = store time.  The decoder's row kernels write 12 MB (x fp32 + LayerNorm bf16) at the
// very end of a ~7 us kernel, all CTAs at once; QKV writes 12 MB (q + K/V pages).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o storebench storebench.cu && ./storebench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void __launch_bounds__(256) k(float4* dst, size_t n16, int mode) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const size_t per = (n16 + gridDim.x - 1) / gridDim.x;
  const size_t b = blockIdx.x * per, e = min(n16, b + per);
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (size_t i = b + threadIdx.x; i < e; i += blockDim.x) {
    if (mode == 0) dst[i] = v;
    else if (mode == 1) __stcg(dst + i, v);
    else __stcs(dst + i, v);
  }
}

// mode 3: the same bytes leave through the TMA engine - the CTA's share is written from a shared-memory buffer with
// cp.async.bulk.global.shared::cta in `chunk`-byte pieces (one thread issues), completion awaited once at the end.
__global__ void __launch_bounds__(256) kb(uint8_t* dst, size_t bytes, int chunk) {
  extern __shared__ __align__(128) uint8_t sm[];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  for (int i = threadIdx.x; i < chunk / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3f800000u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0) {
    const size_t per = ((bytes + gridDim.x - 1) / gridDim.x + chunk - 1) / chunk * chunk;
    const size_t b = blockIdx.x * per, e = min(bytes, b + per);
    for (size_t o = b; o + chunk <= e; o += chunk)
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + o), "r"((uint32_t)__cvta_generic_to_shared(sm)), "r"(chunk) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

int main() {
  float4* buf; CK(cudaMalloc(&buf, 64 << 20));
  cudaStream_t s; CK(cudaStreamCreate(&s));
  const int N = 200;
  for (int grid : {128, 148, 296, 592}) {
    for (int mb : {4, 12, 24}) {
      for (int mode : {0, 1, 2}) {
        const size_t n16 = (size_t)mb << 16;
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < N; ++i) {
          cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = s;
          cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
          cfg.attrs = at; cfg.numAttrs = 1;
          CK(cudaLaunchKernelEx(&cfg, k, buf, n16, mode));
        }
        CK(cudaStreamEndCapture(s, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
          CK(cudaEventRecord(a, s)); CK(cudaGraphLaunch(ge, s)); CK(cudaEventRecord(b, s)); CK(cudaEventSynchronize(b));
          float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
        }
        const double per = best * 1e3 / N;
        printf("grid %3d  %2d MB  %s: %6.2f us per launch  %5.2f TB/s\n", grid, mb, mode == 0 ? "st      " : mode == 1 ? "st.cg   " : "st.cs   ", per,
               (double)mb * 1048576 / (per * 1e-6) / 1e12);
        cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
      }
    }
  }
  CK(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  for (int grid : {128, 148}) {
    for (int mb : {4, 12, 24}) {
      for (int chunk : {4096, 16384, 65536}) {
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < N; ++i) {
          cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = s; cfg.dynamicSmemBytes = chunk;
          cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
          cfg.attrs = at; cfg.numAttrs = 1;
          CK(cudaLaunchKernelEx(&cfg, kb, reinterpret_cast<uint8_t*>(buf), (size_t)mb << 20, chunk));
        }
        CK(cudaStreamEndCapture(s, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
          CK(cudaEventRecord(a, s)); CK(cudaGraphLaunch(ge, s)); CK(cudaEventRecord(b, s)); CK(cudaEventSynchronize(b));
          float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
        }
        const double per = best * 1e3 / N;
        printf("grid %3d  %2d MB  bulk store %5d B chunks: %6.2f us per launch  %5.2f TB/s\n", grid, mb, chunk, per, (double)mb * 1048576 / (per * 1e-6) / 1e12);
        cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
      }
    }
  }
  return 0;
}
