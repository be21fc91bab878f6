// tcgen05.mma issue rate by instruction shape, operands resident in shared memory (no loads): one CTA per SM issues ITERS x 4 MMAs (K = 16
// each, K-major SW128 descriptors over a zeroed 64 KB + 64 KB region), commits and waits.  Answers: what does a small-N MMA cost when
// the WEIGHTS are the M = 128 operand (blockrows.cuh), against M = 64 / N = 256 with the weights as the N operand?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I novic_b200/csrc -o tools/ummabench tools/ummabench.cu
#include <cstdio>
#include "ptx.cuh"
using namespace novic;

template <int NACC, int SAMEA>
__global__ void __launch_bounds__(128, 1) k(uint32_t M, uint32_t N, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16_f32(M, N);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 64 * 1024);
    const long long t0 = clock64();
    if (SAMEA == 2) {          // descriptors precomputed: 16 (A, B) pairs over distinct tiles, the loop body is 16 bare MMAs
      uint64_t da[16], db[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) { da[q] = umma_desc_sw128_kmajor(sa + (q >> 2) * 16384 + (q & 3) * 32); db[q] = umma_desc_sw128_kmajor(sb + (q >> 3) * 32768 + (q & 3) * 32); }
      for (int it = 0; it < iters / 4; ++it) {
#pragma unroll
        for (int q = 0; q < 16; ++q) umma_bf16_ss(slot + (q & (NACC - 1)) * N, da[q], db[q], idesc, 1u);
      }
    } else
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_bf16_ss(slot + ((it * 4 + kk) & (NACC - 1)) * N, umma_desc_sw128_kmajor(sa + (SAMEA ? 0 : (it & 3) * 16384 + kk * 32)), umma_desc_sw128_kmajor(sb + (it & 1) * 32768 + kk * 32), idesc, 1u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0, 1);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(slot);
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 8);
  cudaFuncSetAttribute(k<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 130 * 1024);
  cudaFuncSetAttribute(k<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 130 * 1024);
  cudaFuncSetAttribute(k<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 130 * 1024);
  cudaFuncSetAttribute(k<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 130 * 1024);
  const int shapes[][2] = {{128, 256}, {128, 128}, {128, 64}, {128, 32}, {128, 16}, {64, 64}};
  for (int grid : {148})
    for (auto& s : shapes) {
      const int iters = 256;
      for (int mode = 0; mode < 4; ++mode) {   // 0: one accumulator; 1: two independent accumulators, alternating; 2: two accumulators, the same A tile every time
        const int nacc = mode == 0 ? 1 : 2, same_a = mode >= 2 ? mode - 1 : 0;
        auto fn = mode == 0 ? k<1, 0> : mode == 1 ? k<2, 0> : mode == 2 ? k<2, 1> : k<2, 2>;
        fn<<<grid, 128, 130 * 1024>>>(s[0], s[1], 8, out);   // warm
        cudaDeviceSynchronize();
        fn<<<grid, 128, 130 * 1024>>>(s[0], s[1], iters, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        const double cyc = double(out[0]) / (iters * 4);
        printf("grid %3d  M=%3d N=%3d K=16 acc=%d sameA=%d: %6.1f cycles per MMA  = %6.0f FLOP/clk/SM, A operand %5.1f B/clk, B operand %5.1f B/clk\n", grid, s[0], s[1], nacc, same_a, cyc,
               2.0 * s[0] * s[1] * 16 / cyc, s[0] * 32 / cyc, s[1] * 32 / cyc);
      }
    }
  return 0;
}
