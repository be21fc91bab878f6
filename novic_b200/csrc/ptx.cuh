// Thin inline-PTX wrappers for the sm_100a features the kernels use: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (TMEM alloc / mma / commit / ld) and the proxy fences between them.
// Everything here is written against the PTX ISA for sm_100a (CUDA 12.9); there is no fallback path.
#pragma once

#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace novic {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}

// Device-side watchdog word: a kernel that waits "forever" on an mbarrier (a descriptor or byte-count bug)
// records where and traps, instead of hanging the GPU.
__device__ unsigned int g_watchdog_code = 0;
__device__ unsigned int* g_watchdog_host = nullptr;  // mapped pinned host word: survives the context-killing trap

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t site) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s at 1.9 GHz
      const unsigned int code = 0x80000000u | (site << 24) | ((blockIdx.y & 0xfffu) << 12) | (blockIdx.x & 0xfffu);
      atomicExch(&g_watchdog_code, code);
      if (g_watchdog_host != nullptr) *reinterpret_cast<volatile unsigned int*>(g_watchdog_host) = code;
      __threadfence_system();
      asm volatile("trap;\n");
    }
  }
}

// Optional phase trace (bring-up / tuning): when g_trace != nullptr, CTA (0, by) records clock64() at numbered points.
__device__ long long* g_trace = nullptr;
__device__ int g_trace_counter = 0;   // GEMM launches seen since arming
__device__ int g_trace_target = -1;   // ordinal of the launch to record
// Called by thread 0 of every CTA at kernel entry; returns whether this CTA records.
__device__ __forceinline__ bool trace_begin() {
  if (g_trace == nullptr || blockIdx.x != 0 || blockIdx.y != (gridDim.y > 1 ? 1u : 0u)) return false;
  return atomicAdd(&g_trace_counter, 1) == g_trace_target;
}
__device__ __forceinline__ void trace_point(bool on, int idx) {
  if (on) g_trace[idx] = clock64();
}

// Dropout masks of the training step (nn.Dropout at embedding_decoder.py:1290,:1297 and inside nn.TransformerEncoderLayer): a
// counter-based hash of (seed, site, element index) - reproducible in the backward pass and in the test oracle
// (the tests' CPU checker replays them) without storing a mask.  site = layer * 8 + kind (0 input, 1 attention probabilities,
// 2 attention branch, 3 feed-forward activation, 4 feed-forward branch).  Returns 1 / (1 - p) for kept elements, 0 for dropped.
struct DropCfg {
  const uint32_t* seed;   // device word holding the step's seed: the value changes every step, the pointer does not, so a captured graph of the
                          // training step replays with fresh masks (the host rewrites the word before each launch)
  uint32_t thresh;        // round(p * 2^24); 0 = dropout off (seed is not read)
  float scale;            // 1 / (1 - p)
};
__host__ __device__ __forceinline__ uint32_t drop_hash(uint32_t seed, uint32_t site, uint32_t idx) {
  uint32_t x = idx * 0x9E3779B1u + seed + site * 0x85EBCA77u;
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ float drop_factor(const DropCfg& d, uint32_t site, uint32_t idx) {
  if (d.thresh == 0u) return 1.0f;
  return (drop_hash(__ldg(d.seed), site, idx) >> 8) >= d.thresh ? d.scale : 0.f;
}
constexpr uint32_t kDropInput = 0, kDropAttn = 1, kDropBranch1 = 2, kDropFfn = 3, kDropBranch2 = 4;

// Programmatic dependent launch: pdl_trigger() lets the next kernel of the stream start its prologue while this one is
// still running; pdl_wait() blocks until every kernel this one depends on has completed and its writes are visible.
// Both are no-ops when the kernel was not launched with the programmatic-serialization attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// Fences between the generic proxy, the async proxy (TMA / UMMA smem reads) and tcgen05
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// TMA
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// 2-D tiled load: coordinates are (c0 = innermost/K element index, c1 = row index).
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0,
                                            int32_t c1, uint64_t cache_hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(cache_hint)
      : "memory");
}

// 3-D tiled load over the (64 columns, rows, k-blocks) view of a row-major bf16 matrix: c1 = row index, c2 = first k-block.
// One request brings several consecutive [rows x 128 B] swizzled k-block tiles - the TMA unit serves about two requests at a
// time at ~0.5 k cycles each whatever their size (tools/tmabench.cu), so operand delivery scales with the request size.
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c1, int32_t c2,
                                            uint64_t cache_hint) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(0), "r"(c1), "r"(c2), "l"(cache_hint)
      : "memory");
}

// The same request delivered to the same shared-memory offset of EVERY CTA of `cta_mask` (bit r = cluster rank r); each destination's
// mbarrier at the offset of `bar` receives the complete_tx of the bytes that landed there.
__device__ __forceinline__ void tma_load_3d_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c1, int32_t c2, uint16_t cta_mask,
                                               uint64_t cache_hint) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6, %7;\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(0), "r"(c1), "r"(c2), "h"(cta_mask), "l"(cache_hint)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1, uint16_t cta_mask,
                                               uint64_t cache_hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5, %6;\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask), "l"(cache_hint)
      : "memory");
}

// 3-D tiled load with all three coordinates (c0 innermost)
__device__ __forceinline__ void tma_load_3d_at(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2,
                                               uint64_t cache_hint) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(cache_hint)
      : "memory");
}

// 1-D bulk copy global -> shared (size and both addresses multiples of 16 B), completion on an mbarrier.
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ void bulk_load_1d_hint(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t cache_hint) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar)), "l"(cache_hint)
      : "memory");
}

// 16-byte global store with an L2 eviction-priority hint
__device__ __forceinline__ void st_global_v4_hint(void* gdst, const uint4& v, uint64_t cache_hint) {
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;\n" ::"l"(gdst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(cache_hint)
               : "memory");
}

// createpolicy-encoded L2 hints (same encodings CUTLASS uses for sm_90+)
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

// ------------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ------------------------------------------------------------------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {
  static_assert(kCols >= 32 && kCols <= 512 && (kCols & (kCols - 1)) == 0, "TMEM columns: power of two in [32,512]");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_dst)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}

template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(kCols) : "memory");
}

// Shared-memory matrix descriptor, K-major operand stored as rows of 64 bf16 (128 B) with the 128-byte
// swizzle TMA produces: 8-row groups are 1024 B apart (SBO), LBO is unused for swizzled K-major layouts.
// Field layout as in the PTX ISA "matrix descriptor" table for tcgen05 (version field = 1 on sm_100).
__device__ __forceinline__ uint64_t umma_desc_sw128_kmajor(uint32_t smem_addr_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr_bytes & 0x3FFFFu) >> 4);  // [0,14)  start address >> 4
  d |= static_cast<uint64_t>(1) << 16;                            // [16,30) leading byte offset >> 4 (ignored)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                    // [32,46) stride byte offset >> 4
  d |= static_cast<uint64_t>(1) << 46;                            // [46,48) descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                            // [61,64) layout: SWIZZLE_128B
  return d;
}

// MN-major operand (the weight-gradient GEMMs contract over the ROWS of two row-major activation matrices, so the contraction index is
// the strided one): the tile is stored as blocks of 64 M/N-elements - each block [K rows][128 B], written by one TMA box with the
// 128-byte swizzle - LBO apart; inside a block the 8-row groups along K are 1024 B apart (SBO).  One MMA (K = 16) reads two such
// groups, so the start address advances by 2048 B per K step.
__device__ __forceinline__ uint64_t umma_desc_sw128_mnmajor(uint32_t smem_addr_bytes, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr_bytes & 0x3FFFFu) >> 4);  // [0,14)  start address >> 4
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;               // [16,30) leading byte offset >> 4: between 64-element blocks along M/N
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                    // [32,46) stride byte offset >> 4: between 8-row groups along K
  d |= static_cast<uint64_t>(1) << 46;                            // [46,48) descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                            // [61,64) layout: SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_f32_mn(uint32_t M, uint32_t N, bool a_mn = true, bool b_mn = true) {
  return (1u << 4) | (1u << 7) | (1u << 10)
         | ((a_mn ? 1u : 0u) << 15)       // [15]    A major: M
         | ((b_mn ? 1u : 0u) << 16)       // [16]    B major: N
         | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// Instruction descriptor for kind::f16 with bf16 A/B (both K-major) and fp32 accumulation.
__host__ __device__ constexpr uint32_t umma_idesc_bf16_f32(uint32_t M, uint32_t N) {
  return (1u << 4)          // [4,6)   D format: f32
         | (1u << 7)        // [7,10)  A format: bf16
         | (1u << 10)       // [10,13) B format: bf16
         | (0u << 15)       // [15]    A major: K
         | (0u << 16)       // [16]    B major: K
         | ((N >> 3) << 17) // [17,23) N / 8
         | ((M >> 4) << 24);// [24,29) M / 16
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------------------
// CTA pairs (cta_group::2): two CTAs of a cluster (ranks 2i, 2i + 1, same TPC) run ONE MMA of M = 256.  Each CTA holds its own 128 rows
// of A, HALF of the B tile (N / 2 rows) and the 128 x N accumulator rows of its half in its own TMEM; the even ("leader") CTA issues
// the instruction, the tensor cores of both SMs read both shared memories.  Per FLOP each SM takes in half the operand bytes of a
// 128 x 128 single-CTA tile, which is what the L2 -> SM path can sustain next to the MMA rate (gemm.cuh, gemm2_kernel).
// ------------------------------------------------------------------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst) {   // by one full warp of EACH CTA of the pair, same smem offset
  static_assert(kCols >= 32 && kCols <= 512 && (kCols & (kCols - 1)) == 0, "TMEM columns: power of two in [32,512]");
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_dst)), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(kCols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B^T with M = 256 rows (128 per CTA); issued by ONE thread of the leader CTA.
__device__ __forceinline__ void umma2_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Single-CTA MMAs, arrival multicast: once this CTA's previously issued MMAs have completed, arrive on the mbarrier at this shared-memory
// offset in every CTA of `cta_mask` (the peers whose multicast loads refill the operand stage those MMAs were reading).
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
// Arrive on the mbarrier at this shared-memory offset in every CTA of `cta_mask` once the pair's previously issued MMAs have completed.
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
// 3-D tiled load into THIS CTA's shared memory whose bytes are counted on an mbarrier given by its cluster address (the leader's).
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster, int32_t c1, int32_t c2,
                                                uint64_t cache_hint) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(0), "r"(c1), "r"(c2), "l"(cache_hint)
      : "memory");
}
// Arrive on an mbarrier of any CTA of the cluster (address from dsmem_addr); release at cluster scope.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(bar_cluster) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (thread t receives lane/row
// (warp_id % 4) * 32 + t, columns [col, col + 32)).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// The same load without the wait: several can be in flight; tmem_ld_wait() before the registers are read.
__device__ __forceinline__ void tmem_ld_32x32_nowait(uint32_t taddr, float (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
        "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]),
        "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]),
        "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// Same, 16 columns (fewer live registers in register-heavy epilogues).
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// registers -> TMEM: this warp's 32 lanes x 16 consecutive fp32 columns (same addressing as tmem_ld_32x16)
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// Thread-block clusters / distributed shared memory
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
// All threads of all CTAs of the cluster (release/acquire: smem writes before are visible to peers after).
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// Exit barrier: only shared-memory lifetime matters (no data is handed over), so no release/acquire fence is needed.
__device__ __forceinline__ void cluster_sync_relaxed() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.aligned;\n" ::: "memory");
}
// Remote shared-memory address of `local_smem_ptr` in CTA `peer_rank` of the cluster.
__device__ __forceinline__ uint32_t dsmem_addr(const void* local_smem_ptr, uint32_t peer_rank) {
  uint32_t raddr;
  asm("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(raddr) : "r"(smem_u32(local_smem_ptr)), "r"(peer_rank));
  return raddr;
}
// Non-volatile so that several independent remote loads can be in flight at once (each costs ~200+ cycles).
// Bulk copy from this CTA's shared memory into a peer CTA's (TMA engine, asynchronous): `bytes` (multiple of 16) from local_src to the
// peer address dst_cluster (dsmem_addr), completion signalled as complete_tx on the PEER's mbarrier bar_cluster (dsmem_addr of it).
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t dst_cluster, const void* local_src, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst_cluster),
               "r"(smem_u32(local_src)), "r"(bytes), "r"(bar_cluster)
               : "memory");
}
// 16-byte store into a peer CTA's shared memory (address from dsmem_addr)
__device__ __forceinline__ void dsmem_st_v4(uint32_t raddr, const uint4& v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// generic-proxy writes (any state space, peers' shared memory included) -> visible to async-proxy reads ordered after it
__device__ __forceinline__ void fence_proxy_async_any() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
__device__ __forceinline__ float2 dsmem_ld_f32x2_addr(uint32_t raddr) {
  float2 v;
  asm("ld.shared::cluster.v2.f32 {%0, %1}, [%2];\n" : "=f"(v.x), "=f"(v.y) : "r"(raddr));
  return v;
}
__device__ __forceinline__ float2 dsmem_ld_f32x2(const void* local_smem_ptr, uint32_t peer_rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(raddr) : "r"(smem_u32(local_smem_ptr)), "r"(peer_rank));
  float2 v;
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];\n" : "=f"(v.x), "=f"(v.y) : "r"(raddr) : "memory");
  return v;
}

// ------------------------------------------------------------------------------------------------
// Small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(h);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

}  // namespace novic
