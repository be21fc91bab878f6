// tcgen05 GEMM for the decoder: D[M, N] = A[M, K] * W[N, K]^T with bf16 operands (both K-major), fp32
// accumulation in TMEM, operands staged by TMA into 128B-swizzled shared memory, and the layer's
// elementwise work fused into the epilogue (SURVEY.md §2 kernel inventory rows: QKV, out-proj + residual +
// LayerNorm, FFN1 + GELU, FFN2 + residual + LayerNorm, prefix projection + positions + LayerNorm,
// logits + max / argmax / log-sum-exp / top-k).
//
// One CTA computes one 128 x BN output tile.  Warp roles (192 threads):
//   warp 0      : TMA producer (one elected lane)
//   warp 1      : TMEM allocator + tcgen05.mma issuer (one elected lane)
//   warps 2..5  : epilogue; warp w owns TMEM lanes [32*(w%4), 32*(w%4)+32), thread = one output row
// Pipelines: smem full/empty mbarriers between TMA and MMA; one tmem_full mbarrier MMA -> epilogue.
#pragma once

#include "ptx.cuh"

namespace novic {

constexpr int kE = 512;           // hidden dim (config/train.yaml:256)
constexpr int kBlockM = 128;
constexpr int kBlockK = 64;       // 64 bf16 = 128 B = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 192;
constexpr int kABytes = kBlockM * kBlockK * 2;

// fp32 residual stream `x` lives in a 32-row blocked layout so that "thread = row" epilogue accesses are
// coalesced: element (row, col) is at (((row >> 5) * (kE / 4) + (col >> 2)) * 32 + (row & 31)) * 4 + (col & 3).
__host__ __device__ __forceinline__ size_t xblk_off(int row, int col4) {
  return ((static_cast<size_t>(row >> 5) * (kE / 4) + col4) * 32 + (row & 31)) * 4;
}

template <int BN>
struct GemmSmem {
  static constexpr int kBBytes = BN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int bytes(int stages) { return stages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + 2048 /*epilogue scratch*/; }
};

// ---------------------------------------------------------------------------------------------------------
// Epilogues.  Each provides Params and a static run() executed by the 128 epilogue threads; `row_in_tile`
// is the accumulator row (= TMEM lane) this thread owns, `tmem_row` the TMEM address of that lane, column 0.
// ---------------------------------------------------------------------------------------------------------

// q / k / v split (nn.MultiheadAttention in_proj, rows [Wq;Wk;Wv]): q -> q buffer, k and v -> KV cache page
// of the row's sequence at the row's position.
struct EpiQKV {
  static constexpr int BN = 128;
  struct Params {
    __nv_bfloat16* q;       // [M, 512]
    __nv_bfloat16* kcache;  // this layer: [slots, smax, 512]
    __nv_bfloat16* vcache;
    int rows_per_seq, pos0, slot_mul, smax;
  };
  __device__ static __forceinline__ void run(const Params& p, uint32_t tmem_row, int row, int n0, int M, int,
                                             float*) {
    __nv_bfloat16* dst;
    if (n0 < kE) {
      dst = p.q + static_cast<size_t>(row) * kE + n0;
    } else {
      const int seq = row / p.rows_per_seq;
      const int pos = p.pos0 + (row - seq * p.rows_per_seq);
      const size_t page = (static_cast<size_t>(seq) * p.slot_mul * p.smax + pos) * kE;
      dst = (n0 < 2 * kE) ? p.kcache + page + (n0 - kE) : p.vcache + page + (n0 - 2 * kE);
    }
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      float v[32];
      tmem_ld_32x32(tmem_row + c * 32, v);
      if (row < M) {
        uint4* d4 = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
          o.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
          o.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
          o.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
          d4[q] = o;
        }
      }
    }
  }
};

// FFN first linear + exact GELU -> bf16 h[M, 128]
struct EpiGelu {
  static constexpr int BN = 128;
  struct Params {
    __nv_bfloat16* h;
    int ldh;
  };
  __device__ static __forceinline__ void run(const Params& p, uint32_t tmem_row, int row, int n0, int M, int,
                                             float*) {
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      float v[32];
      tmem_ld_32x32(tmem_row + c * 32, v);
      if (row < M) {
        uint4* d4 = reinterpret_cast<uint4*>(p.h + static_cast<size_t>(row) * p.ldh + n0 + c * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = pack_bf16x2(gelu_erf(v[q * 8 + 0]), gelu_erf(v[q * 8 + 1]));
          o.y = pack_bf16x2(gelu_erf(v[q * 8 + 2]), gelu_erf(v[q * 8 + 3]));
          o.z = pack_bf16x2(gelu_erf(v[q * 8 + 4]), gelu_erf(v[q * 8 + 5]));
          o.w = pack_bf16x2(gelu_erf(v[q * 8 + 6]), gelu_erf(v[q * 8 + 7]));
          d4[q] = o;
        }
      }
    }
  }
};

// Full-row epilogue (BN = 512 = hidden dim): residual add (or prefix position add), write the fp32 residual
// stream, LayerNorm (biased variance, eps 1e-5, gain only) and write the normalised row as bf16 - the A
// operand of the next GEMM.  Used for out-proj, FFN2 and the prefix projection.
struct EpiRow {
  static constexpr int BN = 512;
  struct Params {
    float* x;               // blocked fp32 residual stream (read-modify-write unless prefix mode)
    __nv_bfloat16* xn;      // [rows, 512] bf16 LayerNorm output
    const float* gain;      // LayerNorm weight of the *consumer* (norm2 / next layer's norm1 / final norm)
    const float* pos;       // prefix mode: positional table [smax, 512]; else nullptr
    int prefix_rep;         // prefix mode: sequences per embedding (multi-target M or 1)
    int prefix_rows_per_seq;// prefix mode: rows per sequence in x / xn (P for decode prefill, S for teacher forcing)
    int remap_rows_in;      // xn row remap (0 = identity): rows per sequence in,
    int remap_skip;         //   leading rows per sequence to drop,
    int remap_rows_out;     //   rows per sequence out
    float eps;
  };
  __device__ static __forceinline__ void run(const Params& p, uint32_t tmem_row, int row, int n0, int M, int,
                                             float* s_gain) {
    // stage the LayerNorm gain in shared memory (128 threads x 4 floats)
    {
      const int t = (threadIdx.x - 64);
      reinterpret_cast<float4*>(s_gain)[t] = __ldg(reinterpret_cast<const float4*>(p.gain) + t);
      asm volatile("bar.sync 1, 128;\n" ::: "memory");
    }
    const bool prefix = p.pos != nullptr;
    const int reps = prefix ? p.prefix_rep : 1;
    const int ptok = n0 / kE;  // prefix mode: which prefix position this N tile produces
    for (int j = 0; j < reps; ++j) {
      const int orow = prefix ? ((row * reps + j) * p.prefix_rows_per_seq + ptok) : row;
      float sum = 0.f, sumsq = 0.f;
      __syncwarp();  // reconverge lanes that left the previous iteration early
      // tcgen05.ld is warp-collective (.sync.aligned): every thread executes it, only the global accesses
      // are predicated on row < M.
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float v[32];
        tmem_ld_32x32(tmem_row + c * 32, v);
        if (row < M) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4* xp = reinterpret_cast<float4*>(p.x + xblk_off(orow, c * 8 + q));
            float4 r;
            if (prefix) {
              r = __ldg(reinterpret_cast<const float4*>(p.pos + static_cast<size_t>(ptok) * kE + c * 32 + q * 4));
            } else {
              r = *xp;
            }
            r.x += v[q * 4 + 0]; r.y += v[q * 4 + 1]; r.z += v[q * 4 + 2]; r.w += v[q * 4 + 3];
            *xp = r;
            sum += (r.x + r.y) + (r.z + r.w);
            sumsq += (r.x * r.x + r.y * r.y) + (r.z * r.z + r.w * r.w);
          }
        }
      }
      if (row >= M) continue;  // no collective operation below this point
      const float mean = sum * (1.0f / kE);
      const float var = fmaxf(sumsq * (1.0f / kE) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + p.eps);
      int nrow = orow;
      if (p.remap_rows_in > 0) {
        const int seq = orow / p.remap_rows_in;
        const int r = orow - seq * p.remap_rows_in;
        if (r < p.remap_skip) continue;
        nrow = seq * p.remap_rows_out + (r - p.remap_skip);
      }
      uint4* dn = reinterpret_cast<uint4*>(p.xn + static_cast<size_t>(nrow) * kE);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float y[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 r = *reinterpret_cast<const float4*>(p.x + xblk_off(orow, c * 8 + q));
          const float4 g = reinterpret_cast<const float4*>(s_gain)[c * 8 + q];
          y[q * 4 + 0] = (r.x - mean) * rstd * g.x;
          y[q * 4 + 1] = (r.y - mean) * rstd * g.y;
          y[q * 4 + 2] = (r.z - mean) * rstd * g.z;
          y[q * 4 + 3] = (r.w - mean) * rstd * g.w;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = pack_bf16x2(y[q * 8 + 0], y[q * 8 + 1]);
          o.y = pack_bf16x2(y[q * 8 + 2], y[q * 8 + 3]);
          o.z = pack_bf16x2(y[q * 8 + 4], y[q * 8 + 5]);
          o.w = pack_bf16x2(y[q * 8 + 6], y[q * 8 + 7]);
          dn[c * 4 + q] = o;
        }
      }
    }
  }
};

// Per (row, vocab tile) statistics written by the logits epilogue and merged by the selection kernels.
struct __align__(32) LogitPartial {
  float max_all;     // max over the tile's valid columns (natural units)
  float sumexp_tau;  // sum exp((x - max_all) / tau)
  float sumexp_one;  // sum exp(x - max_all)
  float sum_x;       // sum of x (label smoothing term)
  float best_val;    // best selectable logit (column 0 excluded when the end token is banned)
  int best_idx;      // its vocabulary index (lowest index wins ties)
  float tgt_logit;   // logit of this row's target id if it falls in this tile, else -inf
  int pad_;
};

// Vocabulary logits (tied weights, embedding_decoder.py:725): never materialised unless asked for.
template <int BN_, int HCAP>
struct EpiLogits {
  static constexpr int BN = BN_;
  struct Params {
    float* logits;              // optional [M, ld_logits] fp32
    long long ld_logits;
    LogitPartial* part;         // [M, ntiles]
    float* topv;                // optional [M, ntiles, HCAP]
    int* topi;
    const long long* target;    // optional [M] target ids (-1 = ignore)
    int n_valid;                // V
    int ntiles;
    float inv_tau;
    int ban_eos;                // exclude id 0 from best / top-k (first generated token)
  };
  __device__ static __forceinline__ void run(const Params& p, uint32_t tmem_row, int row, int n0, int M, int tile,
                                             float*) {
    float m = -INFINITY, s_tau = 0.f, s_one = 0.f, sum_x = 0.f, best = -INFINITY, tgt_logit = -INFINITY;
    int best_i = 0x7fffffff;
    float tv[HCAP > 0 ? HCAP : 1];
    int ti[HCAP > 0 ? HCAP : 1];
#pragma unroll
    for (int i = 0; i < (HCAP > 0 ? HCAP : 1); ++i) { tv[i] = -INFINITY; ti[i] = 0x7fffffff; }
    const long long tgt = (p.target != nullptr && row < M) ? p.target[row] : -1;
    const bool tau_is_one = (p.inv_tau == 1.0f);
    const bool vec_ok = (p.ld_logits & 3) == 0;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int col0 = n0 + c * 32;
      if (col0 >= p.n_valid) break;  // tile-uniform
      float v[32];
      tmem_ld_32x32(tmem_row + c * 32, v);
      const int nv = min(32, p.n_valid - col0);
      float cm = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < nv) {
          cm = fmaxf(cm, v[j]);
          sum_x += v[j];
          const int col = col0 + j;
          if (!(p.ban_eos && col == 0)) {
            if (v[j] > best) { best = v[j]; best_i = col; }
            if (HCAP > 0) {
              if (v[j] > tv[HCAP > 0 ? HCAP - 1 : 0]) {
                float cv = v[j]; int ci = col;
#pragma unroll
                for (int i = 0; i < HCAP; ++i) {
                  if (cv > tv[i]) { const float tf = tv[i]; const int tI = ti[i]; tv[i] = cv; ti[i] = ci; cv = tf; ci = tI; }
                }
              }
            }
          }
          if (static_cast<long long>(col) == tgt) tgt_logit = v[j];
        }
      }
      const float m_new = fmaxf(m, cm);
      const float corr = __expf(m - m_new);  // m = -inf on the first chunk -> 0
      float a_one = 0.f, a_tau = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < nv) {
          const float d = v[j] - m_new;
          a_one += __expf(d);
          if (!tau_is_one) a_tau += __expf(d * p.inv_tau);
        }
      }
      s_one = s_one * corr + a_one;
      s_tau = tau_is_one ? s_one : (s_tau * __expf((m - m_new) * p.inv_tau) + a_tau);
      m = m_new;
      if (p.logits != nullptr && row < M) {
        float* lp = p.logits + static_cast<size_t>(row) * p.ld_logits + col0;
        if (vec_ok && nv == 32) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            reinterpret_cast<float4*>(lp)[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nv) lp[j] = v[j];
        }
      }
    }
    if (row < M) {
      LogitPartial o;
      o.max_all = m; o.sumexp_tau = s_tau; o.sumexp_one = s_one; o.sum_x = sum_x;
      o.best_val = best; o.best_idx = best_i; o.tgt_logit = tgt_logit; o.pad_ = 0;
      p.part[static_cast<size_t>(row) * p.ntiles + tile] = o;
      if (HCAP > 0) {
        const size_t base = (static_cast<size_t>(row) * p.ntiles + tile) * (HCAP > 0 ? HCAP : 1);
#pragma unroll
        for (int i = 0; i < (HCAP > 0 ? HCAP : 1); ++i) { p.topv[base + i] = tv[i]; p.topi[base + i] = ti[i]; }
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------------------------------------
template <class Epi, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int M,
            int num_k_blocks, typename Epi::Params ep) {
  constexpr int BN = Epi::BN;
  constexpr int UN = BN > 256 ? 256 : BN;     // N of one tcgen05.mma
  constexpr int NMMA = BN / UN;
  constexpr int kBBytes = BN * kBlockK * 2;
  constexpr int kStage = kABytes + kBBytes;
  constexpr uint32_t kIdesc = umma_idesc_bf16_f32(kBlockM, UN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * kStage);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* epi_scratch = reinterpret_cast<float*>(smem + STAGES * kStage + 256);

  const int warp = threadIdx.x >> 5;
  const int m0 = blockIdx.y * kBlockM;
  const int n0 = blockIdx.x * BN;
  __shared__ int s_trace;
  if (threadIdx.x == 0) { s_trace = trace_begin() ? 1 : 0; trace_point(s_trace != 0, 0); }

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1) {
    if (lane_id() == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      mbar_init(tmem_full_bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<BN>(tmem_slot);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const bool tr = s_trace != 0;
  if (threadIdx.x == 0) trace_point(tr, 1);

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 1);
        uint8_t* sa = smem + stage * kStage;
        uint8_t* sb = sa + kABytes;
        mbar_arrive_expect_tx(&full_bar[stage], kStage);
        tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * kBlockK, m0, kEvictNormal);
#pragma unroll
        for (int j = 0; j < NMMA; ++j)
          tma_load_2d(sb + j * (UN * kBlockK * 2), &tmap_b, &full_bar[stage], kb * kBlockK, n0 + j * UN, kEvictLast);
        if (kb == 0) trace_point(tr, 2);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      trace_point(tr, 3);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase, 2);
        if (kb == 0) trace_point(tr, 4);
        tc_fence_after_sync();
        const uint32_t sa = smem_u32(smem + stage * kStage);
        const uint32_t sb = sa + kABytes;
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
          const uint64_t da = umma_desc_sw128_kmajor(sa + k * (kUmmaK * 2));
#pragma unroll
          for (int j = 0; j < NMMA; ++j) {
            const uint64_t db = umma_desc_sw128_kmajor(sb + j * (UN * kBlockK * 2) + k * (kUmmaK * 2));
            umma_bf16_ss(tmem_base + j * UN, da, db, kIdesc, (kb | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above have read it
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tmem_full_bar);        // accumulator complete
      trace_point(tr, 5);
    }
  } else {
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane_id();
    mbar_wait(tmem_full_bar, 0, 3);
    if (threadIdx.x == 64) trace_point(tr, 6);
    tc_fence_after_sync();
    Epi::run(ep, tmem_base + (static_cast<uint32_t>(quad * 32) << 16), m0 + row_in_tile, n0, M, blockIdx.x,
             epi_scratch);
    if (threadIdx.x == 64) trace_point(tr, 8);
  }

  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) trace_point(tr, 10);
  if (warp == 1) tmem_dealloc<BN>(tmem_base);
}


// ---------------------------------------------------------------------------------------------------------
// Full-row GEMM + residual + LayerNorm, split over a 4-CTA cluster.
//
// The 512-wide output row of a 128-row tile is split across the 4 CTAs of a cluster (128 columns each), so a
// decode step over 4096 sequences fills 128 SMs instead of 32.  Each epilogue thread keeps its 128 row values in
// registers (single pass): the residual slice is prefetched while the MMAs run, the accumulator is added from
// TMEM, the per-row (sum, sum of squares) partials are exchanged through distributed shared memory, and every
// CTA then normalises and writes its own 128 columns of x (fp32) and LayerNorm(x) (bf16).
// ---------------------------------------------------------------------------------------------------------
constexpr int kRowBN = 128;
constexpr int kRowCluster = kE / kRowBN;  // 4

struct RowLnSmem {
  static constexpr int kStageBytes = kABytes + kRowBN * kBlockK * 2;
  static constexpr int bytes(int stages) { return stages * kStageBytes + 1024 + 256 + 128 * 8 /*stats*/ + kRowBN * 4 /*gain*/; }
};

template <int STAGES>
__global__ void __cluster_dims__(kRowCluster, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_rowln_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int M,
                  int num_k_blocks, EpiRow::Params ep) {
  constexpr int BN = kRowBN;
  constexpr int kStage = RowLnSmem::kStageBytes;
  constexpr uint32_t kIdesc = umma_idesc_bf16_f32(kBlockM, BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * kStage);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float2* s_stats = reinterpret_cast<float2*>(smem + STAGES * kStage + 256);   // [128] (sum, sumsq) of this CTA's columns
  float* s_gain = reinterpret_cast<float*>(smem + STAGES * kStage + 256 + 128 * 8);

  const int warp = threadIdx.x >> 5;
  const int m0 = blockIdx.y * kBlockM;
  const int n0 = blockIdx.x * BN;               // column in the [*, P*512] / [*, 512] output
  const uint32_t crank = cluster_ctarank();     // == blockIdx.x % 4
  const int coff = static_cast<int>(crank) * BN;  // column offset inside the 512-wide row
  const int ptok = blockIdx.x / kRowCluster;    // prefix mode: which prefix position
  __shared__ int s_trace;
  if (threadIdx.x == 0) { s_trace = trace_begin() ? 1 : 0; trace_point(s_trace != 0, 0); }

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1) {
    if (lane_id() == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      mbar_init(tmem_full_bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<BN>(tmem_slot);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const bool tr = s_trace != 0;
  if (threadIdx.x == 0) trace_point(tr, 1);

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 1);
        uint8_t* sa = smem + stage * kStage;
        mbar_arrive_expect_tx(&full_bar[stage], kStage);
        tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * kBlockK, m0, kEvictNormal);
        tma_load_2d(sa + kABytes, &tmap_b, &full_bar[stage], kb * kBlockK, n0, kEvictLast);
        if (kb == 0) trace_point(tr, 2);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      trace_point(tr, 3);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase, 2);
        if (kb == 0) trace_point(tr, 4);
        tc_fence_after_sync();
        const uint32_t sa = smem_u32(smem + stage * kStage);
        const uint32_t sb = sa + kABytes;
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k)
          umma_bf16_ss(tmem_base, umma_desc_sw128_kmajor(sa + k * (kUmmaK * 2)), umma_desc_sw128_kmajor(sb + k * (kUmmaK * 2)),
                       kIdesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tmem_full_bar);
      trace_point(tr, 5);
    }
  }

  // ---- epilogue part 1 (warps 2..5): residual prefetch + accumulator + partial row statistics -------------
  const bool is_epi = warp >= 2;
  const int quad = warp & 3;
  const int row_in_tile = quad * 32 + static_cast<int>(lane_id());
  const int row = m0 + row_in_tile;
  const bool prefix = ep.pos != nullptr;
  const int reps = prefix ? ep.prefix_rep : 1;
  float r[BN];
  if (is_epi) {
    s_gain[threadIdx.x - 64] = __ldg(ep.gain + coff + (threadIdx.x - 64));
    if (row < M) {
      if (prefix) {
        const float4* p4 = reinterpret_cast<const float4*>(ep.pos + static_cast<size_t>(ptok) * kE + coff);
#pragma unroll
        for (int q = 0; q < BN / 4; ++q) { const float4 t = __ldg(p4 + q); r[q * 4] = t.x; r[q * 4 + 1] = t.y; r[q * 4 + 2] = t.z; r[q * 4 + 3] = t.w; }
      } else {
#pragma unroll
        for (int q = 0; q < BN / 4; ++q) {
          const float4 t = *reinterpret_cast<const float4*>(ep.x + xblk_off(row, (coff >> 2) + q));
          r[q * 4] = t.x; r[q * 4 + 1] = t.y; r[q * 4 + 2] = t.z; r[q * 4 + 3] = t.w;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < BN; ++i) r[i] = 0.f;
    }
    if (threadIdx.x == 64) trace_point(tr, 11);
    mbar_wait(tmem_full_bar, 0, 3);
    if (threadIdx.x == 64) trace_point(tr, 6);
    tc_fence_after_sync();
    const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    float sum = 0.f, sumsq = 0.f;
#pragma unroll
    for (int c = 0; c < BN / 16; ++c) {
      float v[16];
      tmem_ld_32x16(tmem_row + c * 16, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float t = r[c * 16 + j] + v[j];
        r[c * 16 + j] = t;
        sum += t;
        sumsq = fmaf(t, t, sumsq);
      }
    }
    s_stats[row_in_tile] = make_float2(sum, sumsq);
    if (threadIdx.x == 64) trace_point(tr, 7);
  }
  __syncwarp();
  cluster_sync_all();
  if (threadIdx.x == 64) trace_point(tr, 8);   // every thread of the 4 CTAs: partial statistics are now visible cluster-wide

  // ---- epilogue part 2: combine the 4 partials (fixed order -> identical mean/rstd in all CTAs), normalise, store
  if (is_epi && row < M) {
    float sum = 0.f, sumsq = 0.f;
#pragma unroll
    for (uint32_t pr = 0; pr < kRowCluster; ++pr) {
      const float2 t = dsmem_ld_f32x2(&s_stats[row_in_tile], pr);
      sum += t.x;
      sumsq += t.y;
    }
    const float mean = sum * (1.0f / kE);
    const float rstd = rsqrtf(fmaxf(sumsq * (1.0f / kE) - mean * mean, 0.f) + ep.eps);
    for (int j = 0; j < reps; ++j) {
      const int orow = prefix ? ((row * reps + j) * ep.prefix_rows_per_seq + ptok) : row;
#pragma unroll
      for (int q = 0; q < BN / 4; ++q)
        *reinterpret_cast<float4*>(ep.x + xblk_off(orow, (coff >> 2) + q)) = make_float4(r[q * 4], r[q * 4 + 1], r[q * 4 + 2], r[q * 4 + 3]);
      int nrow = orow;
      if (ep.remap_rows_in > 0) {
        const int seq = orow / ep.remap_rows_in;
        const int rr = orow - seq * ep.remap_rows_in;
        if (rr < ep.remap_skip) continue;
        nrow = seq * ep.remap_rows_out + (rr - ep.remap_skip);
      }
      uint4* dn = reinterpret_cast<uint4*>(ep.xn + static_cast<size_t>(nrow) * kE + coff);
#pragma unroll
      for (int q = 0; q < BN / 8; ++q) {
        float y[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = (r[q * 8 + i] - mean) * rstd * s_gain[q * 8 + i];
        dn[q] = make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
      }
    }
  }

  if (threadIdx.x == 64) trace_point(tr, 9);
  tc_fence_before_sync();
  __syncwarp();
  cluster_sync_all();   // peers may still be reading this CTA's s_stats
  if (threadIdx.x == 0) trace_point(tr, 10);
  if (warp == 1) tmem_dealloc<BN>(tmem_base);
}

}  // namespace novic
