// Micro-benchmark: L2 -> shared-memory streaming rate of one CTA per SM, the resource that bounds the decoder's small-M
// GEMMs (each CTA streams all weights of its N-slice for only 128 rows).  Compares
//   (a) 2-D TMA tile loads (128 rows x 64 bf16 box, 128B swizzle) out of a row-major [N, 512] matrix, and
//   (b) 1-D bulk copies of pre-packed contiguous 16 KB tiles,
// as a function of the number of tiles in flight per CTA and of the number of CTAs pulling at once.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2bench l2bench.cu -lcuda && ./l2bench
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
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// One thread per CTA streams `ntiles` tiles through a ring of `depth` slots; tile t of CTA b is tile (b * 7 + t) % total_tiles.
__global__ void __launch_bounds__(128, 1) stream_kernel(const __grid_constant__ CUtensorMap map, const uint8_t* packed, int mode, int tile_bytes,
                                                         int depth, int ntiles, int total_tiles, int box_rows, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm);
  uint8_t* ring = sm + 1024;
  if (threadIdx.x == 0) {
    for (int s = 0; s < depth; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
    const long long t0 = clock64();
    auto issue = [&](int t) {
      const int s = t % depth;
      const int tile = (blockIdx.x * 7 + t) % total_tiles;
      mbar_expect(&full[s], tile_bytes);
      if (mode == 0) tma_load_2d(ring + (size_t)s * tile_bytes, &map, &full[s], (tile & 7) * 64, (tile >> 3) * box_rows);
      else bulk_load(ring + (size_t)s * tile_bytes, packed + (size_t)tile * tile_bytes, tile_bytes, &full[s]);
    };
    for (int t = 0; t < depth && t < ntiles; ++t) issue(t);
    for (int t = 0; t < ntiles; ++t) {
      mbar_wait(&full[t % depth], (t / depth) & 1);
      if (t + depth < ntiles) issue(t + depth);
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
  const int N = 1536, K = 512;                      // one layer's in_proj: 1.5 MB, L2-resident
  uint8_t* w; CK(cudaMalloc(&w, (size_t)N * K * 2)); CK(cudaMemset(w, 1, (size_t)N * K * 2));
  long long* cyc; CK(cudaMalloc(&cyc, 148 * 8));
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int box_rows : {128, 256}) {
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N}; cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
    if (encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
    const int tile_bytes = box_rows * 128;
    const int total_tiles = (N / box_rows) * 8;
    for (int mode : {0, 1}) {
      for (int grid : {1, 32, 64, 128, 148}) {
        for (int inflight_kb : {64, 96, 128, 192}) {
          const int depth = inflight_kb * 1024 / tile_bytes;
          const int ntiles = (4 << 20) / tile_bytes;          // 4 MB per CTA
          float best = 1e9f;
          for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(a);
            stream_kernel<<<grid, 128, depth * tile_bytes + 1024>>>(map, w, mode, tile_bytes, depth, ntiles, total_tiles, box_rows, cyc);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
          }
          long long h[148]; CK(cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost));
          long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
          cudaError_t e = cudaGetLastError();
          printf("%s box %3d rows (%2d KB tiles) grid %3d in-flight %3d KB: %6.1f us  %5.1f B/clk/SM  %5.2f TB/s aggregate %s\n", mode ? "bulk-1D packed" : "TMA-2D row-major",
                 box_rows, tile_bytes / 1024, grid, inflight_kb, best * 1e3, (double)ntiles * tile_bytes / mx, (double)grid * ntiles * tile_bytes / (best * 1e-3) / 1e12,
                 e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
      }
    }
  }
  return 0;
}
