// Peak rate of the legacy tensor path (mma.sync.m16n8k16 bf16 -> fp32) on this GPU: register-resident operands, 8 independent accumulator
// chains per warp, W warps per CTA, one CTA wave.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mmabench tools/mmabench.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void mma_loop(float* out, int iters) {
  unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  float c[8][4] = {};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0.f;
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  if (s == 123.456f) out[0] = s;
}
int main() {
  float* d; cudaMalloc(&d, 4);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int warps : {4, 8, 16, 32}) {
    const int iters = 20000;
    mma_loop<<<sms, warps * 32>>>(d, 100);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    mma_loop<<<sms, warps * 32>>>(d, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double flop = 2.0 * 16 * 8 * 16 * 8.0 * iters * warps * sms;
    printf("mma.sync m16n8k16 bf16: %2d warps/SM x %d SMs: %.3f ms  %.1f TFLOP/s\n", warps, sms, ms, flop / ms / 1e9);
  }
  return 0;
}
