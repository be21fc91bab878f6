// Micro-benchmark: how fast can one SM-resident design stream "pages" (two contiguous chunks per item, like one
// sequence's K rows and V rows) out of HBM?  Compares a TMA bulk-copy ring against plain LDG.128 streaming.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o membench membench.cu && ./membench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t done = 0;
  while (!done) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- A: bulk-copy ring: 1 producer warp, ncons consumer warps, nstages stages --------------------------------
__global__ void __launch_bounds__(288, 1) ring_kernel(const uint8_t* k, const uint8_t* v, size_t item_stride, int chunk, int nitems, int nstages,
                                                      int ncons, int split, unsigned* sink) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm);
  uint64_t* empty = full + 16;
  uint8_t* ring = sm + 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = (nitems + gridDim.x - 1) / gridDim.x;
  const int i0 = blockIdx.x * per, i1 = min(nitems, i0 + per);
  if (threadIdx.x == 0) { for (int s = 0; s < nstages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); } asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  const int stage_bytes = 2 * chunk;
  if (warp == 8) {
    for (int it = i0, kk = 0; it < i1; ++it, ++kk) {
      const int st = kk % nstages; const uint32_t ph = (kk / nstages) & 1;
      if (lane == 0) { mbar_wait(&empty[st], ph ^ 1); mbar_expect(&full[st], stage_bytes); }
      __syncwarp();
      uint8_t* dst = ring + (size_t)st * stage_bytes;
      if (split == 1) {
        if (lane == 0) { bulk_load(dst, k + it * item_stride, chunk, &full[st]); bulk_load(dst + chunk, v + it * item_stride, chunk, &full[st]); }
      } else {  // split the two chunks into `split` pieces each, issued by different lanes
        const int piece = chunk / split;
        if (lane < split) { bulk_load(dst + lane * piece, k + it * item_stride + lane * piece, piece, &full[st]);
                            bulk_load(dst + chunk + lane * piece, v + it * item_stride + lane * piece, piece, &full[st]); }
      }
    }
  } else if (warp < ncons) {
    unsigned acc = 0;
    for (int it = i0 + warp, kk = warp; it < i1; it += ncons, kk += ncons) {
      const int st = kk % nstages; const uint32_t ph = (kk / nstages) & 1;
      mbar_wait(&full[st], ph);
      const uint4* p = reinterpret_cast<const uint4*>(ring + (size_t)st * stage_bytes);
      for (int i = lane; i < stage_bytes / 16; i += 32) { uint4 t = p[i]; acc ^= t.x ^ t.y ^ t.z ^ t.w; }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
    }
    if (acc == 0x12345678u) sink[0] = acc;
  }
}

// ---- B: plain LDG.128 streaming: one warp per item, UNR loads in flight per lane -----------------------------
template <int UNR>
__global__ void __launch_bounds__(256) ldg_kernel(const uint8_t* k, const uint8_t* v, size_t item_stride, int chunk, int nitems, unsigned* sink) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= nitems) return;
  unsigned acc = 0;
  for (int half = 0; half < 2; ++half) {
    const uint4* p = reinterpret_cast<const uint4*>((half ? v : k) + w * item_stride);
    const int n = chunk / 16;
    for (int i = lane; i < n; i += 32 * UNR) {
      uint4 t[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) if (i + u * 32 < n) t[u] = __ldcs(p + i + u * 32);
#pragma unroll
      for (int u = 0; u < UNR; ++u) if (i + u * 32 < n) acc ^= t[u].x ^ t[u].y ^ t[u].z ^ t[u].w;
    }
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

int main() {
  const int nitems = 4096;
  const size_t item_stride = 19 * 1024, region = (size_t)nitems * item_stride;
  uint8_t* buf; unsigned* sink;
  CK(cudaMalloc(&buf, 2 * region + (64 << 20)));
  CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(buf, 1, 2 * region));
  uint8_t* flush; CK(cudaMalloc(&flush, 256 << 20));
  CK(cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto timeit = [&](auto launch, const char* name, int chunk) {
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaMemsetAsync(flush, rep, 256 << 20);
      cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    printf("%-48s chunk %5d B: %7.2f us  %6.2f TB/s %s\n", name, chunk, best * 1e3, 2.0 * chunk * nitems / (best * 1e-3) / 1e12, e == cudaSuccess ? "" : cudaGetErrorString(e));
  };
  for (int chunk : {5 * 1024, 12 * 1024, 19 * 1024}) {
    for (int grid : {148, 296}) {
      for (int split : {1, 4}) {
        const int budget = (grid == 148 ? 200 : 100) * 1024;
        int nst = budget / (2 * chunk); if (nst > 16) nst = 16; if (nst < 2) nst = 2;
        int ncons = nst < 8 ? nst : 8; nst = nst / ncons * ncons;
        char name[96]; snprintf(name, sizeof(name), "ring grid=%d stages=%d cons=%d split=%d", grid, nst, ncons, split);
        timeit([&] { ring_kernel<<<grid, 288, nst * 2 * chunk + 256>>>(buf, buf + region, item_stride, chunk, nitems, nst, ncons, split, sink); }, name, chunk);
      }
    }
    timeit([&] { ldg_kernel<4><<<nitems / 8, 256>>>(buf, buf + region, item_stride, chunk, nitems, sink); }, "ldg warp/item unroll 4", chunk);
    timeit([&] { ldg_kernel<8><<<nitems / 8, 256>>>(buf, buf + region, item_stride, chunk, nitems, sink); }, "ldg warp/item unroll 8", chunk);
    timeit([&] { ldg_kernel<16><<<nitems / 8, 256>>>(buf, buf + region, item_stride, chunk, nitems, sink); }, "ldg warp/item unroll 16", chunk);
  }
  // reference: a plain big contiguous read
  timeit([&] { ldg_kernel<8><<<nitems / 8, 256>>>(buf, buf + (size_t)nitems * 12 * 1024, 12 * 1024, 12 * 1024, nitems, sink); }, "ldg contiguous 12 KB items (no gaps)", 12 * 1024);
  return 0;
}
