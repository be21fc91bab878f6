// The decoder layer stack as ONE cluster kernel (embedding_decoder.py:309-327, :714: nn.TransformerEncoder, pre-LN).
//
// Every operation of a transformer layer except attention's key/value look-up is local to a residual row, so a
// 128-row tile never needs another tile's data.  A 4-CTA cluster owns one 128-row tile for the whole launch and walks
// the phases   QKV_l -> [attention_l] -> OUT_l (out-proj + residual + LayerNorm2) -> FFN_l (linear1 + GELU + linear2
// + residual + next LayerNorm) -> QKV_{l+1} -> ...   with cluster barriers where a kernel boundary used to be:
//   * the fp32 residual tile lives in TENSOR MEMORY for the whole launch (128 rows x 128 columns per CTA): the
//     out-projection and linear2 MMAs accumulate straight onto it (D = residual + A * W^T), so the residual add costs
//     nothing and no thread carries residual registers through the attention phase,
//   * CTA c of the cluster computes output columns [128c, 128c+128) of every 512-wide product and columns
//     [384c, 384c+384) of the QKV product; LayerNorm statistics are exchanged through distributed shared memory,
//   * a phase's A operand (128 rows x 512, bf16, 128 KB) is resident in shared memory; the weight tiles (128 x 64,
//     16 KB) of ALL phases stream through one ring, so the TMA producer prefetches the next phase's weights while
//     the current phase's epilogue / barriers run,
//   * LayerNorm outputs are handed to the next phase through global memory (L2) + fence.proxy.async + cluster
//     barrier, then re-read by TMA (which applies the 128B operand swizzle).
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + tcgen05.mma issuer, warps 2..9 = epilogue
// (thread = one residual row x 64 columns), warps 10..15 idle outside the attention phase, in which all 16 warps
// are workers (decode steps: 2 of the CTA's 32 sequences each).
// TMEM: 4 slots of 128 columns - slots 0-2: QKV accumulators (slot 0 again for the linear1 hidden tile), slot 3: residual.
#pragma once

#include "kernels.cuh"

namespace novic {

constexpr int kStkMaxLayers = 6;
constexpr int kStkThreads = 512;                     // 16 warps
constexpr int kStkRing = 4;                          // weight-tile ring slots
constexpr int kStkTile = kBlockM * kBlockK * 2;      // 16 KB: 128 rows x 64 bf16 (one swizzled operand tile)
constexpr int kStkAKb = kE / kBlockK;                // 8 k-blocks of the resident A operand
constexpr int kStkABytes = kStkAKb * kStkTile;       // 128 KB
constexpr int kStkStagePitch = 64 + 16;              // staged bf16 row of 32 columns + 16 B pad
constexpr int kStkWarpStage = 32 * kStkStagePitch;   // 2560 B per epilogue warp
constexpr int kStkBarBytes = 512;
constexpr int kStkAttnWarps = 16;                    // attention workers: all 16 warps, 2 of the CTA's 32 sequences each
constexpr int kStkAttnSlots = 2;                     // per warp: 4 KB chunk slots inside the (idle) A region (K chunk, V chunk)
constexpr int kStkResSlot = 3;                       // TMEM slot of the residual tile
static_assert(kStkAttnWarps * kStkAttnSlots * kAsSlotBytes <= kStkABytes, "attention slots must fit in the A region");
constexpr int kStkSmemBytes = 1024 /*align*/ + kStkABytes + kStkRing * kStkTile + 8 * kStkWarpStage + 2 * 128 * 8 /*stats*/ +
                              128 * 4 /*gain*/ + kStkBarBytes;

enum StackPhase : int { kPhQkv = 0, kPhAttn = 1, kPhOut = 2, kPhFfn = 3 };

struct StackLayerMaps { CUtensorMap in_proj, out_proj, linear1, linear2; };   // B operands, boxes of 128 rows x 64
struct StackMaps {
  CUtensorMap xn, ao;                                                         // A operands [M, 512], boxes of 128 rows x 64
  StackLayerMaps w[kStkMaxLayers];
};

struct StackArgs {
  int M;                        // residual rows
  int ph_begin, ph_end;         // phases [ph_begin, ph_end), phase = 4 * layer + StackPhase
  int num_layers;
  float* x;                     // blocked fp32 residual stream (read at entry if load_x, written at exit if store_x)
  __nv_bfloat16* xn;            // LayerNorm rows [M, 512]
  __nv_bfloat16* xfin;          // last layer's (remapped) final-norm rows, or nullptr = write to xn
  __nv_bfloat16* q;             // [M, 512]
  __nv_bfloat16* kv;            // [L][2][slots * smax][512]
  size_t kv_layer;              // elements per (layer, k|v)
  const float* gain_out[kStkMaxLayers];   // norm2[l]
  const float* gain_ffn[kStkMaxLayers];   // norm1[l + 1], or the final norm for the last layer
  int rows_per_seq, pos0, slot_mul, smax; // residual row -> (sequence, position) -> KV page row
  int remap_in, remap_skip, remap_out;    // xn row remap of the last layer's output (0 = identity)
  int load_x, store_x;
  float eps;
  // in-kernel attention phase (decode steps: one query per sequence, no key padding)
  __nv_bfloat16* ao;            // attention output rows [M, 512] (A operand of the out-projection)
  const unsigned char* anc;     // beam ancestry table or nullptr
  int anc_ld, beams, prefix_len;
  float scale_log2e;
};

__device__ __forceinline__ int stk_num_b_tiles(int kind) { return kind == kPhQkv ? 3 * kStkAKb : (kind == kPhOut ? kStkAKb : (kind == kPhFfn ? kStkAKb + 2 : 0)); }

__device__ __forceinline__ void stk_stage_put(uint8_t* stage, int lane, const float (&v)[32]) {
  uint4* d = reinterpret_cast<uint4*>(stage + lane * kStkStagePitch);
#pragma unroll
  for (int q = 0; q < 4; ++q)
    d[q] = make_uint4(pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]), pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]),
                      pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]), pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]));
}

// Write the warp's staged 32 x 32 bf16 block: per instruction the lanes cover 8 rows x 64 contiguous bytes.
template <class RowPtr>
__device__ __forceinline__ void stk_copy_out(const uint8_t* stage, int lane, RowPtr row_ptr) {
  __syncwarp();
  const int sub = lane >> 2, chunk = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + sub;
    const uint4 v = *reinterpret_cast<const uint4*>(stage + r * kStkStagePitch + chunk * 16);
    __nv_bfloat16* dst = row_ptr(r);
    if (dst != nullptr) *reinterpret_cast<uint4*>(dst + chunk * 8) = v;
  }
  __syncwarp();
}

__device__ int g_stk_fence_mode = 2;   // tuning: 0 = none, 1 = fence.proxy.async.global, 2 = fence.proxy.async (all state spaces)
__device__ __forceinline__ void fence_proxy_async_all() {
  const int mode = g_stk_fence_mode;
  if (mode == 2) asm volatile("fence.proxy.async;\n" ::: "memory");
  else if (mode == 1) asm volatile("fence.proxy.async.global;\n" ::: "memory");
}

__global__ void __cluster_dims__(kRowCluster, 1, 1) __launch_bounds__(kStkThreads, 1)
layer_stack_kernel(const __grid_constant__ StackMaps maps, const StackArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem;                                     // [8][128 x 128 B]
  uint8_t* ring = a_smem + kStkABytes;                        // [kStkRing][128 x 128 B]
  uint8_t* staging = ring + kStkRing * kStkTile;              // [8 warps][32 x 80 B]
  float2* s_stats = reinterpret_cast<float2*>(staging + 8 * kStkWarpStage);   // [2 halves][128 rows] (sum, sumsq)
  float* s_gain = reinterpret_cast<float*>(s_stats + 2 * 128);                // [128]
  uint64_t* a_full = reinterpret_cast<uint64_t*>(s_gain + 128);               // [8]
  uint64_t* b_full = a_full + kStkAKb;                                        // [kStkRing]
  uint64_t* b_empty = b_full + kStkRing;                                      // [kStkRing]
  uint64_t* acc_full = b_empty + kStkRing;                                    // [4]
  uint64_t* h_ready = acc_full + 4;                                           // [1]
  uint64_t* r_ready = h_ready + 1;                                            // [1] residual tile loaded into TMEM
  uint64_t* attn_full = r_ready + 1;                                          // [kStkAttnWarps][kStkAttnSlots]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(attn_full + kStkAttnWarps * kStkAttnSlots);

  const int warp = threadIdx.x >> 5;
  const int lane = static_cast<int>(lane_id());
  const int m0 = blockIdx.y * kBlockM;
  const uint32_t crank = cluster_ctarank();
  const int coff = static_cast<int>(crank) * kRowBN;          // this CTA's columns of the 512-wide row
  constexpr uint32_t kIdesc = umma_idesc_bf16_f32(kBlockM, kTileN);
  pdl_trigger();
  __shared__ int s_trace;
  if (threadIdx.x == 0) { s_trace = trace_begin() ? 1 : 0; trace_point(s_trace != 0, 0); }

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&maps.xn);
      tma_prefetch_desc(&maps.ao);
      for (int i = 0; i < kStkAKb; ++i) mbar_init(&a_full[i], 1);
      for (int i = 0; i < kStkRing; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
      for (int i = 0; i < 4; ++i) mbar_init(&acc_full[i], 1);
      mbar_init(h_ready, 8);
      mbar_init(r_ready, 8);
      for (int i = 0; i < kStkAttnWarps * kStkAttnSlots; ++i) mbar_init(&attn_full[i], 1);
      fence_mbar_init();
    }
  } else if (warp == 1) {
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const bool tr = s_trace != 0 && threadIdx.x == 64;   // phase trace: recorded by the first epilogue thread of CTA (0, 1)
  if (threadIdx.x == 64) trace_point(tr, 1);

  // B tile i of a phase -> TMA source
  auto b_source = [&](int l, int kind, int i, const CUtensorMap*& map, int& c0, int& c1) {
    if (kind == kPhQkv) { map = &maps.w[l].in_proj; c0 = (i & 7) * kBlockK; c1 = static_cast<int>(crank) * 384 + (i >> 3) * kTileN; }
    else if (kind == kPhOut) { map = &maps.w[l].out_proj; c0 = i * kBlockK; c1 = coff; }
    else if (i < kStkAKb) { map = &maps.w[l].linear1; c0 = i * kBlockK; c1 = 0; }
    else { map = &maps.w[l].linear2; c0 = (i - kStkAKb) * kBlockK; c1 = coff; }
  };


  // ---- attention phase (all warps): the cluster's 128 sequences are split 32 per CTA; a warp streams the K / V pages of
  // its sequences through its own 4 KB slots with cp.async.bulk, exactly like attention_stream_kernel (kernels.cuh).
  int attn_n = 0;                                             // chunk loads this warp has issued since kernel start (slot / parity)
  auto attn_phase = [&](int l) {
    const __nv_bfloat16* kc = a.kv + (static_cast<size_t>(l) * 2 + 0) * a.kv_layer;
    const __nv_bfloat16* vc = a.kv + (static_cast<size_t>(l) * 2 + 1) * a.kv_layer;
    const int qpos = a.pos0;
    const int nkeys = qpos + 1;
    const int nchunks = (nkeys + kAsChunk - 1) / kAsChunk;
    const int per_item = 2 * nchunks;
    const int row0 = m0 + 32 * static_cast<int>(crank);
    const int aw = warp;                                      // attention worker index
    int nitems = 0;
    if (aw >= 0) for (int li = aw; li < 32 && row0 + li < a.M; li += kStkAttnWarps) ++nitems;
    const int nloads = nitems * per_item;
    uint8_t* slots = a_smem + (aw < 0 ? 0 : aw) * (kStkAttnSlots * kAsSlotBytes);
    uint64_t* fb = attn_full + (aw < 0 ? 0 : aw) * kStkAttnSlots;
    const bool contiguous = a.beams == 1;
    const int nbase = attn_n;
    auto issue = [&](int n) {
      const int i = n / per_item, r = n - i * per_item;
      const int c = r >> 1, kv = r & 1;
      const int seq = row0 + aw + i * kStkAttnWarps;
      const int j0 = c * nkeys / nchunks, rows = (c + 1) * nkeys / nchunks - j0;
      const __nv_bfloat16* base = kv ? vc : kc;
      const int sl = (nbase + n) % kStkAttnSlots;
      uint8_t* dst = slots + sl * kAsSlotBytes;
      const int own_slot = seq * a.slot_mul;
      if (lane == 0) mbar_arrive_expect_tx(&fb[sl], static_cast<uint32_t>(rows) * 1024u);
      if (contiguous) {
        if (lane == 0) bulk_load_1d(dst, base + (static_cast<size_t>(own_slot) * a.smax + j0) * kE, static_cast<uint32_t>(rows) * 1024u, &fb[sl]);
      } else {
        __syncwarp();
        if (lane < rows) {
          const int j = j0 + lane;
          const int group0 = (own_slot / a.beams) * a.beams;
          int slot;
          if (j < a.prefix_len) slot = group0;
          else if (a.anc != nullptr && j < qpos) slot = group0 + a.anc[static_cast<size_t>(seq) * a.anc_ld + (j - a.prefix_len)];
          else slot = own_slot;
          bulk_load_1d(dst + lane * 1024, base + (static_cast<size_t>(slot) * a.smax + j) * kE, 1024u, &fb[sl]);
        }
      }
    };
    int issued = 0;
    for (; issued < min(nloads, kStkAttnSlots); ++issued) issue(issued);
    float qf[16], acc[16], pj[kAsChunk];
    float m = -INFINITY, lsum = 0.f, corr = 0.f;
    uint4 qa = make_uint4(0, 0, 0, 0), qb = qa;
    if (nitems > 0) {
      const uint4* q4 = reinterpret_cast<const uint4*>(a.q + static_cast<size_t>(row0 + aw) * kE) + lane * 2;
      qa = *q4; qb = *(q4 + 1);
    }
    int i = 0, r = 0;
    for (int n = 0; n < nloads; ++n) {
      const int seq = row0 + aw + i * kStkAttnWarps;
      const int c = r >> 1;
      const int j0 = c * nkeys / nchunks, rows = (c + 1) * nkeys / nchunks - j0;
      if (r == 0) {
        bf16x8_to_f32(qa, qf);
        bf16x8_to_f32(qb, qf + 8);
#pragma unroll
        for (int k = 0; k < 16; ++k) { qf[k] *= a.scale_log2e; acc[k] = 0.f; }
        m = -INFINITY; lsum = 0.f;
        if (i + 1 < nitems) {
          const uint4* q4 = reinterpret_cast<const uint4*>(a.q + static_cast<size_t>(seq + kStkAttnWarps) * kE) + lane * 2;
          qa = *q4; qb = *(q4 + 1);
        }
      }
      const int sl = (nbase + n) % kStkAttnSlots;
      const uint8_t* src = slots + sl * kAsSlotBytes + lane * 32;
      mbar_wait(&fb[sl], (static_cast<uint32_t>(nbase + n) / kStkAttnSlots) & 1u, 28);
      if ((r & 1) == 0) {
        float sc[kAsChunk];
#pragma unroll
        for (int u = 0; u < kAsChunk; ++u) {
          sc[u] = -INFINITY;
          if (u < rows) {
            const uint4* k4 = reinterpret_cast<const uint4*>(src + u * 1024);
            float kf[16];
            bf16x8_to_f32(k4[0], kf);
            bf16x8_to_f32(k4[1], kf + 8);
            float d0 = 0.f, d1 = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) { d0 = fmaf(qf[k], kf[k], d0); d1 = fmaf(qf[8 + k], kf[8 + k], d1); }
            sc[u] = d0 + d1;
          }
        }
#pragma unroll
        for (int u = 0; u < kAsChunk; ++u) {
          if (u < rows) {
            sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], 1);
            sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], 2);
          }
        }
        const float m_new = fmaxf(fmaxf(m, fmaxf(sc[0], sc[1])), fmaxf(sc[2], sc[3]));
        corr = exp2f(m - m_new);
        float psum = 0.f;
#pragma unroll
        for (int u = 0; u < kAsChunk; ++u) { pj[u] = exp2f(sc[u] - m_new); psum += pj[u]; }
        lsum = lsum * corr + psum;
        m = m_new;
      } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[k] *= corr;
#pragma unroll
        for (int u = 0; u < kAsChunk; ++u) {
          if (u < rows) {
            const uint4* v4 = reinterpret_cast<const uint4*>(src + u * 1024);
            float vf[16];
            bf16x8_to_f32(v4[0], vf);
            bf16x8_to_f32(v4[1], vf + 8);
#pragma unroll
            for (int k = 0; k < 16; ++k) acc[k] = fmaf(pj[u], vf[k], acc[k]);
          }
        }
      }
      __syncwarp();
      if (issued < nloads) { issue(issued); ++issued; }
      if (++r == per_item) {
        const float inv = 1.0f / lsum;
        uint32_t o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = pack_bf16x2(acc[2 * k] * inv, acc[2 * k + 1] * inv);
        uint4* d = reinterpret_cast<uint4*>(a.ao + static_cast<size_t>(seq) * kE) + lane * 2;
        d[0] = make_uint4(o[0], o[1], o[2], o[3]);
        d[1] = make_uint4(o[4], o[5], o[6], o[7]);
        r = 0;
        ++i;
      }
    }
    attn_n = nbase + nloads;
    fence_proxy_async_all();     // ao rows (generic stores) -> the out-projection's TMA loads
    __syncwarp();
  };

  if (warp == 0) {
    // =============================== TMA producer ===============================
    int bt = 0;                                              // weight tiles issued so far (ring position)
    auto issue_b = [&](int ph, int i) {
      const int slot = bt % kStkRing;
      mbar_wait(&b_empty[slot], ((static_cast<uint32_t>(bt) / kStkRing) & 1u) ^ 1u, 21);
      const CUtensorMap* map; int c0, c1;
      b_source(ph >> 2, ph & 3, i, map, c0, c1);
      mbar_arrive_expect_tx(&b_full[slot], kStkTile);
      tma_load_2d(ring + slot * kStkTile, map, &b_full[slot], c0, c1, kEvictLast);
      ++bt;
    };
    int pre = 0;
    if (lane == 0) {   // weights do not depend on the previous kernel
      const int nb = stk_num_b_tiles(a.ph_begin & 3);
      pre = nb < kStkRing ? nb : kStkRing;
      for (int i = 0; i < pre; ++i) issue_b(a.ph_begin, i);
    }
    pdl_wait();
    for (int ph = a.ph_begin; ph < a.ph_end; ++ph) {
      const int kind = ph & 3;
      if (lane == 0 && kind != kPhAttn) {
        const CUtensorMap* amap = kind == kPhOut ? &maps.ao : &maps.xn;
        for (int kb = 0; kb < kStkAKb; ++kb) {
          mbar_arrive_expect_tx(&a_full[kb], kStkTile);
          tma_load_2d(a_smem + kb * kStkTile, amap, &a_full[kb], kb * kBlockK, m0, kEvictNormal);
        }
        const int nb = stk_num_b_tiles(kind);
        for (int i = pre; i < nb; ++i) issue_b(ph, i);
        pre = 0;
        int nxt = ph + 1;
        if (nxt < a.ph_end && (nxt & 3) == kPhAttn) ++nxt;
        if (nxt < a.ph_end) {   // next phase's first weight tiles: in flight during this phase's epilogue and barriers
          const int nb2 = stk_num_b_tiles(nxt & 3);
          pre = nb2 < kStkRing ? nb2 : kStkRing;
          for (int i = 0; i < pre; ++i) issue_b(nxt, i);
        }
      }
      __syncwarp();
      if (kind == kPhOut || kind == kPhFfn) { cluster_sync_all(); cluster_sync_all(); }
      else if (kind == kPhAttn) { cluster_sync_all(); attn_phase(ph >> 2); cluster_sync_all(); }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    int bt = 0;
    uint32_t a_par = 0, h_par = 0;
    bool r_waited = false;
    auto mma_tile = [&](uint32_t acc, const uint8_t* a_tile, bool first) {
      const int slot = bt % kStkRing;
      mbar_wait(&b_full[slot], (static_cast<uint32_t>(bt) / kStkRing) & 1u, 22);
      tc_fence_after_sync();
      const uint32_t sa = smem_u32(a_tile), sb = smem_u32(ring + slot * kStkTile);
#pragma unroll
      for (int k = 0; k < kBlockK / kUmmaK; ++k)
        umma_bf16_ss(acc, umma_desc_sw128_kmajor(sa + k * (kUmmaK * 2)), umma_desc_sw128_kmajor(sb + k * (kUmmaK * 2)), kIdesc,
                     (!first || k != 0) ? 1u : 0u);
      umma_commit(&b_empty[slot]);
      ++bt;
    };
    pdl_wait();
    for (int ph = a.ph_begin; ph < a.ph_end; ++ph) {
      const int kind = ph & 3;
      if (lane == 0 && kind != kPhAttn) {
        tc_fence_after_sync();
        if (kind == kPhQkv) {
          for (int t = 0; t < 3; ++t) {
            for (int kb = 0; kb < kStkAKb; ++kb) {
              if (t == 0) mbar_wait(&a_full[kb], a_par, 23);
              mma_tile(tmem_base + t * kTileN, a_smem + kb * kStkTile, kb == 0);
            }
            umma_commit(&acc_full[t]);
          }
        } else if (kind == kPhOut) {
          if (!r_waited) { mbar_wait(r_ready, 0, 29); r_waited = true; tc_fence_after_sync(); }
          for (int kb = 0; kb < kStkAKb; ++kb) {   // accumulate onto the residual tile: x += ao * Wo^T
            mbar_wait(&a_full[kb], a_par, 23);
            mma_tile(tmem_base + kStkResSlot * kTileN, a_smem + kb * kStkTile, false);
          }
          umma_commit(&acc_full[3]);
        } else {
          for (int kb = 0; kb < kStkAKb; ++kb) {
            mbar_wait(&a_full[kb], a_par, 23);
            mma_tile(tmem_base, a_smem + kb * kStkTile, kb == 0);
          }
          umma_commit(&acc_full[0]);
          mbar_wait(h_ready, h_par, 24);     // hidden tile (A operand of linear2) written by the epilogue warps
          h_par ^= 1u;
          if (!r_waited) { mbar_wait(r_ready, 0, 29); r_waited = true; tc_fence_after_sync(); }
          for (int kb = 0; kb < kFfnDim / kBlockK; ++kb) mma_tile(tmem_base + kStkResSlot * kTileN, a_smem + kb * kStkTile, false);   // x += h * W2^T
          umma_commit(&acc_full[1]);
        }
        a_par ^= 1u;
      }
      __syncwarp();
      if (kind == kPhOut || kind == kPhFfn) {
        tc_fence_before_sync();
        cluster_sync_all();
        cluster_sync_all();
        tc_fence_after_sync();
      } else if (kind == kPhAttn) {
        cluster_sync_all();
        attn_phase(ph >> 2);
        cluster_sync_all();
      }
    }
  } else if (warp < 10) {
    // =============================== epilogue warps ===============================
    const int ew = warp - 2;
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may read
    const int half = ew >> 2;                        // which 64 of a 128-column accumulator
    const int row_in_tile = quad * 32 + lane;
    const int row = m0 + row_in_tile;
    const int warp_row0 = m0 + quad * 32;
    const int c0 = coff + half * kRowCols;           // first residual column this thread owns
    const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    uint8_t* stage = staging + ew * kStkWarpStage;
    uint32_t acc_par[4] = {0u, 0u, 0u, 0u};
    const int L = a.num_layers;
    pdl_wait();
    trace_point(tr, 2);

    {
      // residual tile -> TMEM slot 3 (zeros when the launch starts with a QKV-only phase that never reads it)
      const uint32_t rt = tmem_lane + kStkResSlot * kTileN + half * kRowCols;
#pragma unroll
      for (int c = 0; c < kRowCols / 16; ++c) {
        float v[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          if (a.load_x && row < a.M) t = *reinterpret_cast<const float4*>(a.x + xblk_off(row, (c0 >> 2) + c * 4 + q));
          v[q * 4] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
        }
        tmem_st_32x16(rt + c * 16, v);
      }
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(r_ready);
    }

    // the MMAs of `bar_slot` have accumulated onto the residual tile: read this thread's 64 values, partial LayerNorm
    // statistics -> shared memory
    float r[kRowCols];
    auto read_residual_and_stats = [&](int bar_slot) {
      mbar_wait(&acc_full[bar_slot], acc_par[bar_slot], 25);
      acc_par[bar_slot] ^= 1u;
      tc_fence_after_sync();
      float sum = 0.f, sumsq = 0.f;
#pragma unroll
      for (int c = 0; c < kRowCols / 16; ++c) {
        float v[16];
        tmem_ld_32x16(tmem_lane + kStkResSlot * kTileN + half * kRowCols + c * 16, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          r[c * 16 + j] = v[j];
          sum += v[j];
          sumsq = fmaf(v[j], v[j], sumsq);
        }
      }
      s_stats[half * 128 + row_in_tile] = make_float2(sum, sumsq);
      tc_fence_before_sync();
    };
    // combine the cluster's 8 partials (fixed order), normalise, write this thread's 64 columns of LayerNorm(x) as bf16
    auto layer_norm_out = [&](__nv_bfloat16* dst, int remap_in, int remap_skip, int remap_out) {
      float mean = 0.f, rstd = 0.f;
      if (row < a.M) {
        float2 part[kRowCluster * 2];
#pragma unroll
        for (uint32_t pr = 0; pr < kRowCluster; ++pr) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) part[pr * 2 + hh] = dsmem_ld_f32x2_addr(dsmem_addr(&s_stats[hh * 128 + row_in_tile], pr));
        }
        float sum = 0.f, sumsq = 0.f;
#pragma unroll
        for (int i = 0; i < kRowCluster * 2; ++i) { sum += part[i].x; sumsq += part[i].y; }
        mean = sum * (1.0f / kE);
        rstd = rsqrtf(fmaxf(sumsq * (1.0f / kE) - mean * mean, 0.f) + a.eps);
      }
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        float y[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) y[i] = (r[pass * 32 + i] - mean) * rstd * s_gain[half * kRowCols + pass * 32 + i];
        stk_stage_put(stage, lane, y);
        stk_copy_out(stage, lane, [&](int rr) -> __nv_bfloat16* {
          const int grow = warp_row0 + rr;
          if (grow >= a.M) return nullptr;
          int nrow = grow;
          if (remap_in > 0) {
            const int seq = grow / remap_in;
            const int k = grow - seq * remap_in;
            if (k < remap_skip) return nullptr;
            nrow = seq * remap_out + (k - remap_skip);
          }
          return dst + static_cast<size_t>(nrow) * kE + c0 + pass * 32;
        });
      }
      fence_proxy_async_all();     // generic-proxy global writes -> visible to the peers' TMA (async proxy) reads
    };

    for (int ph = a.ph_begin; ph < a.ph_end; ++ph) {
      const int l = ph >> 2, kind = ph & 3;
      if (kind == kPhQkv) {
        __nv_bfloat16* kc = a.kv + (static_cast<size_t>(l) * 2 + 0) * a.kv_layer;
        __nv_bfloat16* vc = a.kv + (static_cast<size_t>(l) * 2 + 1) * a.kv_layer;
        for (int t = 0; t < 3; ++t) {
          mbar_wait(&acc_full[t], acc_par[t], 26);
          acc_par[t] ^= 1u;
          tc_fence_after_sync();
          trace_point(tr, 12 + t);
#pragma unroll
          for (int pass = 0; pass < 2; ++pass) {
            float v[32];
            tmem_ld_32x32(tmem_lane + t * kTileN + half * kRowCols + pass * 32, v);
            stk_stage_put(stage, lane, v);
            const int n = static_cast<int>(crank) * 384 + t * kTileN + half * kRowCols + pass * 32;   // column of the [*, 1536] product
            const int region = n >> 9, nin = n & (kE - 1);
            stk_copy_out(stage, lane, [&](int rr) -> __nv_bfloat16* {
              const int grow = warp_row0 + rr;
              if (grow >= a.M) return nullptr;
              if (region == 0) return a.q + static_cast<size_t>(grow) * kE + nin;
              const int seq = grow / a.rows_per_seq;
              const int pos = a.pos0 + (grow - seq * a.rows_per_seq);
              const size_t page = (static_cast<size_t>(seq) * a.slot_mul * a.smax + pos) * kE;
              return (region == 1 ? kc : vc) + page + nin;
            });
          }
        }
        fence_proxy_async_all();   // q / k / v rows (generic stores) -> bulk-copy (async proxy) reads of the attention phase
        tc_fence_before_sync();
      } else if (kind == kPhAttn) {
        trace_point(tr, 16);
        __syncwarp();
        cluster_sync_all();
        trace_point(tr, 17);
        attn_phase(l);
        trace_point(tr, 18);
        cluster_sync_all();
        trace_point(tr, 19);
      } else if (kind == kPhOut) {
        if (ew < 4) s_gain[threadIdx.x - 64] = __ldg(a.gain_out[l] + coff + (threadIdx.x - 64));
        read_residual_and_stats(3);
        trace_point(tr, 3);
        __syncwarp();
        cluster_sync_all();
        trace_point(tr, 4);
        layer_norm_out(a.xn, 0, 0, 0);
        trace_point(tr, 5);
        __syncwarp();
        cluster_sync_all();
        trace_point(tr, 6);
      } else if (kind == kPhFfn) {
        if (ew < 4) s_gain[threadIdx.x - 64] = __ldg(a.gain_ffn[l] + coff + (threadIdx.x - 64));
        mbar_wait(&acc_full[0], acc_par[0], 27);
        acc_par[0] ^= 1u;
        tc_fence_after_sync();
        trace_point(tr, 7);
        {
          // hidden tile: gelu(acc) -> bf16 -> K-major 128B-swizzled A operand of linear2 (A region, k-block `half`):
          // the row is 128 B there, 16-byte chunk c lands at chunk (c ^ (row & 7)).
          uint8_t* hrow = a_smem + half * kStkTile + row_in_tile * 128;
#pragma unroll
          for (int c = 0; c < kRowCols / 16; ++c) {
            float v[16];
            tmem_ld_32x16(tmem_lane + half * kRowCols + c * 16, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = gelu_fast(v[j]);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int chunk = c * 2 + q;
              *reinterpret_cast<uint4*>(hrow + ((chunk ^ (row_in_tile & 7)) << 4)) =
                  make_uint4(pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]), pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]),
                             pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]), pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]));
            }
          }
          fence_proxy_async_smem();
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(h_ready);
        }
        trace_point(tr, 8);
        read_residual_and_stats(1);
        trace_point(tr, 9);
        __syncwarp();
        cluster_sync_all();
        trace_point(tr, 10);
        const bool last = (l + 1 == L);
        if (last && a.xfin != nullptr) layer_norm_out(a.xfin, a.remap_in, a.remap_skip, a.remap_out);
        else layer_norm_out(a.xn, 0, 0, 0);
        __syncwarp();
        cluster_sync_all();
        trace_point(tr, 11);
      }
    }
    if (a.store_x) {
      tc_fence_after_sync();
#pragma unroll
      for (int c = 0; c < kRowCols / 16; ++c) {
        float v[16];
        tmem_ld_32x16(tmem_lane + kStkResSlot * kTileN + half * kRowCols + c * 16, v);
        if (row < a.M) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(a.x + xblk_off(row, (c0 >> 2) + c * 4 + q)) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
        }
      }
    }
  } else {
    // =============================== attention helper warps (10..15) ===============================
    for (int ph = a.ph_begin; ph < a.ph_end; ++ph) {
      const int kind = ph & 3;
      if (kind == kPhOut || kind == kPhFfn) { cluster_sync_all(); cluster_sync_all(); }
      else if (kind == kPhAttn) { cluster_sync_all(); attn_phase(ph >> 2); cluster_sync_all(); }
    }
  }

  if (threadIdx.x == 64) trace_point(tr, 15);
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

}  // namespace novic
