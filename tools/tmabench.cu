// Micro-benchmark: per-SM L2 -> shared-memory TMA throughput as a function of the REQUEST size (tools/l2bench.cu found 44 B/clk/SM
// with 16 KB boxes and 88 B/clk/SM with 32 KB boxes).  Boxes here are 3-D over a row-major bf16 [rows, 512] matrix:
// (64 columns, R rows, KB k-blocks) -> KB consecutive [R x 128 B] swizzled tiles in shared memory, i.e. exactly what a UMMA
// K-major operand stage of KB k-blocks looks like.  Two matrices alternate like a GEMM's A and B operands.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tmabench tmabench.cu -lcuda && ./tmabench
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t done = 0;
  while (!done) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// One thread per CTA streams `nreq` requests through a ring of `depth` slots, alternating between the two maps.
// issuers = 1: one thread issues everything; issuers = 2: a second warp's thread issues the odd requests (same ring).
__global__ void __launch_bounds__(128, 1) stream_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                         int req_bytes, int depth, int nreq, int rows_a_tiles, int rows_b_tiles, int kb_per_req,
                                                         long long* cycles) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm);
  uint8_t* ring = sm + 1024;
  if (threadIdx.x == 0) {
    for (int s = 0; s < depth; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
    const long long t0 = clock64();
    const int kreq = 8 / kb_per_req;                       // requests per 512-wide row tile
    auto issue = [&](int t) {
      const int s = t % depth;
      const int u = t >> 1;
      const bool b = t & 1;
      const int tile = (blockIdx.x * 7 + u / kreq) % (b ? rows_b_tiles : rows_a_tiles);
      mbar_expect(&full[s], req_bytes);
      tma_load_3d(ring + (size_t)s * req_bytes, b ? &map_b : &map_a, &full[s], 0, tile * 128, (u % kreq) * kb_per_req);
    };
    for (int t = 0; t < depth && t < nreq; ++t) issue(t);
    for (int t = 0; t < nreq; ++t) {
      mbar_wait(&full[t % depth], (t / depth) & 1);
      if (t + depth < nreq) issue(t + depth);
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(fn);
  const int K = 512, NA = 4096, NB = 1536;            // activations [4096, 512] (4 MB) and one layer's in_proj [1536, 512]
  uint8_t *a, *w; CK(cudaMalloc(&a, (size_t)NA * K * 2)); CK(cudaMemset(a, 1, (size_t)NA * K * 2));
  CK(cudaMalloc(&w, (size_t)NB * K * 2)); CK(cudaMemset(w, 1, (size_t)NB * K * 2));
  long long* cyc; CK(cudaMalloc(&cyc, 148 * 8));
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int kbr : {1, 2, 4, 8}) {
    CUtensorMap ma, mb;
    for (int which = 0; which < 2; ++which) {
      cuuint64_t dims[3] = {64, (cuuint64_t)(which ? NB : NA), (cuuint64_t)(K / 64)};
      cuuint64_t strides[2] = {(cuuint64_t)K * 2, 128};
      cuuint32_t box[3] = {64, 128, (cuuint32_t)kbr}; cuuint32_t es[3] = {1, 1, 1};
      if (encode(which ? &mb : &ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, which ? w : a, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
    }
    const int req_bytes = kbr * 16384;
    for (int grid : {1, 148}) {
      for (int inflight_kb : {64, 128, 192}) {
        const int depth = inflight_kb * 1024 / req_bytes;
        if (depth < 1) continue;
        const int nreq = (4 << 20) / req_bytes;           // 4 MB per CTA
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
          cudaEventRecord(e0);
          stream_kernel<<<grid, 128, depth * req_bytes + 1024>>>(ma, mb, req_bytes, depth, nreq, NA / 128, NB / 128, kbr, cyc);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        long long h[148]; CK(cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost));
        long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        cudaError_t e = cudaGetLastError();
        printf("3-D box 64 x 128 x %d (%3d KB requests) grid %3d in-flight %3d KB (%2d requests): %6.1f us  %5.1f B/clk/SM  %5.2f TB/s aggregate  %6.0f clk/request %s\n",
               kbr, req_bytes / 1024, grid, inflight_kb, depth, best * 1e3, (double)nreq * req_bytes / mx, (double)grid * nreq * req_bytes / (best * 1e-3) / 1e12,
               (double)mx / nreq, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
    }
  }
  return 0;
}
