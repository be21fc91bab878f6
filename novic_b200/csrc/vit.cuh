// CLIP ViT image encoder in front of the decoder (call site embedders.py:759-764: encode_image(normalize=False)
// followed by an fp32 normalise).  BASELINE config #5 names DFN5B-CLIP-ViT-H-14-378; its definition lives in the un-vendored
// open_clip_torch (requirements.txt:8), so the architecture is an ASSUMPTION stated in DESIGN.md: open_clip's VisionTransformer,
// patch 14 on 378 x 378 (27 x 27 patches + class token = 730 tokens), width 1280, 32 pre-LN blocks of 16 heads x 80 with a 5120-wide
// QuickGELU MLP, biases on every linear, LayerNorm (weight + bias, eps 1e-5) before the blocks and on the class token after them,
// bias-free 1280 -> 1024 projection.  Parity is pinned only on an independent torch restatement (oracle/vit_oracle.py).
//
// The dense contractions (patch embedding, QKV, out-proj, both MLP layers: 99 % of the FLOPs outside attention) run on the persistent
// tcgen05 / TMEM / TMA GEMM of gemm.cuh through the epilogues below; attention (730 keys x 80 channels per head) runs on
// mma.sync.m16n8k16 tiles with an online softmax (flash-style: scores never leave the registers).
#pragma once

#include "gemm.cuh"
#include "train.cuh"   // ldsm_x4 / ldsm_x4_t / mma_bf16_16816

namespace novic {

constexpr int kVitHeadDim = 80;

// ---------------------------------------------------------------------------------------------------------
// GEMM epilogues (persistent gemm_kernel: 8 epilogue warps, thread = output row, 64 columns per thread)
// ---------------------------------------------------------------------------------------------------------

// out[row, n] = act(acc + bias[n]) as bf16 (QKV: no activation; MLP c_fc: QuickGELU x * sigmoid(1.702 x))
struct EpiVitBias {
  struct Params {
    __nv_bfloat16* out;
    int ld;                 // elements per output row
    const float* bias;      // [N] or nullptr
    int n_valid;            // N (columns >= N of the last tile are not stored)
    int quick_gelu;
  };
  template <class Release>
  __device__ static __forceinline__ void run(const Params& p, const EpiCtx& c, Release release) {
    const int lane = lane_id();
    const int nhalf = c.ncols / kEpiCols;              // 1 (128-column tiles) or 2 (256-column tiles): 64 columns are staged at a time
#pragma unroll 1
    for (int hf = 0; hf < nhalf; ++hf) {
      const int n0 = c.n0 + hf * kEpiCols;
#pragma unroll
      for (int ch = 0; ch < kEpiCols / 32; ++ch) {
        float v[32];
        tmem_ld_32x32(c.tmem_row + hf * kEpiCols + ch * 32, v);
        if (hf == nhalf - 1 && ch == kEpiCols / 32 - 1) release();
        const int col0 = n0 + ch * 32;
        if (p.bias != nullptr && col0 < p.n_valid) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + q);   // N is a multiple of 32 for every ViT GEMM
            v[q * 4] += b.x; v[q * 4 + 1] += b.y; v[q * 4 + 2] += b.z; v[q * 4 + 3] += b.w;
          }
        }
        if (p.quick_gelu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = v[j] * __fdividef(1.0f, 1.0f + __expf(-1.702f * v[j]));
        }
        stage_put32(c.stage, lane, ch * 32, v);
      }
      if (n0 < p.n_valid)
        stage_copy_out(c.stage, lane, [&](int r) -> __nv_bfloat16* {
          const int row = c.warp_row0 + r;
          return row < c.M ? p.out + static_cast<size_t>(row) * p.ld + n0 : nullptr;
        });
      else __syncwarp();
    }
  }
};

// Residual / patch-embedding epilogue on the fp32 token stream x [images * T, W] (row-major):
//   residual mode (table == nullptr):  x[row, n] += acc + bias[n]                                  (attention out-proj, MLP c_proj)
//   patch mode    (table != nullptr):  x[img * T + 1 + p, n] = acc + table[(1 + p) * W + n]        (conv1 as a GEMM over patches + positions;
//                                      GEMM row = img * (T - 1) + p)
// The thread = row accumulator chunk goes through a warp-private XOR-swizzled fp32 tile so that every global access covers 4 rows x 128
// contiguous bytes.
struct EpiVitResid {
  struct Params {
    float* x;
    int ld;                 // W
    const float* bias;      // [W] or nullptr
    const float* table;     // patch mode: positional embedding [T, W]
    int patches;            // patch mode: T - 1
    int n_valid;            // W (columns >= W of a ragged last tile are not touched)
  };
  template <class Release>
  __device__ static __forceinline__ void run(const Params& p, const EpiCtx& c, Release release) {
    const int lane = lane_id();
    float* st = reinterpret_cast<float*>(c.stage);            // 32 rows x 32 fp32 = 4096 B (kEpiStageBytes)
    const int sub = lane >> 3, chunk = lane & 7;
    const int nch = c.ncols / 32;
#pragma unroll 1
    for (int ch = 0; ch < nch; ++ch) {
      float v[32];
      tmem_ld_32x32(c.tmem_row + ch * 32, v);
      if (ch == nch - 1) release();
      if (c.n0 + ch * 32 >= p.n_valid) continue;        // warp-uniform
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(st + lane * 32 + ((q ^ (lane & 7)) << 2)) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
      __syncwarp();
      const int col = c.n0 + ch * 32 + chunk * 4;
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(p.bias + col));
#pragma unroll 4
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + sub;
        const int row = c.warp_row0 + r;
        if (row >= c.M) continue;
        const float4 a = *reinterpret_cast<const float4*>(st + r * 32 + ((chunk ^ (r & 7)) << 2));
        float4 base;
        float* dst;
        if (p.table != nullptr) {
          const int img = row / p.patches, pp = row - img * p.patches;
          dst = p.x + (static_cast<size_t>(img) * (p.patches + 1) + 1 + pp) * p.ld + col;
          base = __ldg(reinterpret_cast<const float4*>(p.table + static_cast<size_t>(1 + pp) * p.ld + col));
        } else {
          dst = p.x + static_cast<size_t>(row) * p.ld + col;
          base = *reinterpret_cast<const float4*>(dst);
        }
        *reinterpret_cast<float4*>(dst) = make_float4(base.x + a.x + b.x, base.y + a.y + b.y, base.z + a.z + b.z, base.w + a.w + b.w);
      }
    }
  }
};

// out[row, n] = acc (fp32, row-major): the final 1280 -> 1024 projection of the class-token rows (B rows only)
struct EpiVitStoreF32 {
  struct Params {
    float* out;
    int ld;
  };
  template <class Release>
  __device__ static __forceinline__ void run(const Params& p, const EpiCtx& c, Release release) {
    const int nch = c.ncols / 32;
#pragma unroll 1
    for (int ch = 0; ch < nch; ++ch) {
      float v[32];
      tmem_ld_32x32(c.tmem_row + ch * 32, v);
      if (ch == nch - 1) release();
      if (c.row < c.M && c.n0 + ch * 32 < p.ld) {
        float4* d = reinterpret_cast<float4*>(p.out + static_cast<size_t>(c.row) * p.ld + c.n0 + ch * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) d[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------------------
// Patches: images [B, 3, S, S] fp32 -> bf16 rows [B * np * np, Kp] with k = (c * 14 + py) * 14 + px (the order of conv1.weight flattened),
// zero-padded from 588 to Kp = 640 columns (a whole number of 64-wide k-blocks).  One warp per patch row.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vit_patchify_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int S, int P, int np, int Kp) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int total = B * np * np;
  if (warp >= total) return;
  const int b = warp / (np * np), pidx = warp - b * np * np, py0 = (pidx / np) * P, px0 = (pidx % np) * P;
  const int K = 3 * P * P;
  __nv_bfloat16* dst = out + static_cast<size_t>(warp) * Kp;
  for (int k = lane; k < Kp; k += 32) {
    float v = 0.f;
    if (k < K) {
      const int ch = k / (P * P), rem = k - ch * P * P, py = rem / P, px = rem - py * P;
      v = __ldg(img + ((static_cast<size_t>(b) * 3 + ch) * S + py0 + py) * S + px0 + px);
    }
    dst[k] = __float2bfloat16_rn(v);
  }
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm over rows of the fp32 token stream (weight + bias, biased variance).  One warp per row, the row in registers.
//   y = LN(x; w1, b1);  optional: x <- y (fp32, in place);  out_bf <- y (bf16), or out_bf <- LN(y; w2, b2) when w2 != nullptr
// (ln_pre followed by the first block's ln_1 is one pass), or out_f32 <- y for the class-token rows (ln_post: row stride `in_stride`).
// NV = W / 128 float4 per lane (W = 1280: 10).
// ---------------------------------------------------------------------------------------------------------
struct VitLnParams {
  float* x;                     // [rows * in_stride] fp32
  long long in_stride;          // elements between consecutive rows (W, or T * W to pick the class tokens)
  const float *w1, *b1, *w2, *b2;
  int store_x;                  // write y back to x
  __nv_bfloat16* out_bf;        // [rows, W] or nullptr
  int rows, W;
  float eps;
  const float* cls;             // optional: row r of x is first SET to cls + pos0 when (r % cls_period) == 0 (class token rows)
  const float* pos0;
  int cls_period;
};

template <int NV>
__global__ void __launch_bounds__(256) vit_layernorm_kernel(const VitLnParams p) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= p.rows) return;
  float4 v[NV];
  float* xr = p.x + static_cast<size_t>(row) * p.in_stride;
  const bool is_cls = p.cls != nullptr && (row % p.cls_period) == 0;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c4 = lane + 32 * j;
    if (is_cls) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p.cls) + c4), b = __ldg(reinterpret_cast<const float4*>(p.pos0) + c4);
      v[j] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    } else {
      v[j] = *reinterpret_cast<const float4*>(xr + c4 * 4);
    }
  }
  auto normalise = [&](const float* w, const float* b) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) s += v[j].x + v[j].y + v[j].z + v[j].w;
    s = warp_sum(s);
    const float mean = s / static_cast<float>(p.W);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float dx = v[j].x - mean, dy = v[j].y - mean, dz = v[j].z - mean, dw = v[j].w - mean;
      q += dx * dx + dy * dy + dz * dz + dw * dw;
    }
    q = warp_sum(q);
    const float rstd = rsqrtf(q / static_cast<float>(p.W) + p.eps);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c4 = lane + 32 * j;
      const float4 g = __ldg(reinterpret_cast<const float4*>(w) + c4), bb = __ldg(reinterpret_cast<const float4*>(b) + c4);
      v[j] = make_float4((v[j].x - mean) * rstd * g.x + bb.x, (v[j].y - mean) * rstd * g.y + bb.y, (v[j].z - mean) * rstd * g.z + bb.z,
                         (v[j].w - mean) * rstd * g.w + bb.w);
    }
  };
  normalise(p.w1, p.b1);
  if (p.store_x) {
#pragma unroll
    for (int j = 0; j < NV; ++j) *reinterpret_cast<float4*>(xr + (lane + 32 * j) * 4) = v[j];
  }
  if (p.out_bf != nullptr) {
    if (p.w2 != nullptr) normalise(p.w2, p.b2);
    __nv_bfloat16* o = p.out_bf + static_cast<size_t>(row) * p.W;
#pragma unroll
    for (int j = 0; j < NV; ++j)
      *reinterpret_cast<uint2*>(o + (lane + 32 * j) * 4) = make_uint2(pack_bf16x2(v[j].x, v[j].y), pack_bf16x2(v[j].z, v[j].w));
  }
}

// ---------------------------------------------------------------------------------------------------------
// Attention of one ViT block: softmax(Q K^T / sqrt(80)) V over all T tokens of an image, no mask (flash-style).
// qkv: bf16 [images * T, 3 * W] row-major (q | k | v, head h at columns h * 80); out: bf16 [images * T, W].
// One CTA = 128 queries (8 warps x 16) of one (image, head); keys in tiles of 64 through a double-buffered cp.async ring;
// Q K^T and P V on mma.sync m16n8k16 from ldmatrix fragments, probabilities go from the score accumulators straight into the A fragments.
// ---------------------------------------------------------------------------------------------------------
constexpr int kVaQ = 128, kVaK = 64, kVaWarps = 8;
constexpr int kVaPitch = kVitHeadDim + 8;                         // 88 bf16 = 176 B rows: ldmatrix rows fall into distinct bank groups
constexpr int kVaStages = 3;                                      // K/V ring: one CTA barrier per key tile
constexpr int kVaSmemBytes = (kVaQ + 2 * kVaStages * kVaK) * kVaPitch * 2;    // Q + 3 x (K, V) = 90 112 B (two CTAs per SM)

struct VitAttnParams {
  const __nv_bfloat16* qkv;
  __nv_bfloat16* out;
  int T, W, heads;
  float scale_log2e;           // log2(e) / sqrt(80)
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = smem_u32(smem_dst);
  const int n = valid ? 16 : 0;       // src-size 0: the 16 bytes are zero-filled (rows past the end of the sequence)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// 2^x on the SFU (MUFU.EX2; exp2f without fast-math expands to a dozen instructions of range handling, which made the softmax - not the
// tensor pipe - the busiest part of the kernel); relative error 2^-22, inputs here are <= 0
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kVaWarps * 32, 2) vit_attention_kernel(const VitAttnParams p) {
  extern __shared__ __align__(16) uint8_t sm_va[];
  __nv_bfloat16* sq = reinterpret_cast<__nv_bfloat16*>(sm_va);
  __nv_bfloat16* skv = sq + kVaQ * kVaPitch;                       // [kVaStages][K | V][64][pitch]
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int lm = lane >> 3, lr = lane & 7, g = lane >> 2, t = lane & 3;
  const int qt = blockIdx.x, head = blockIdx.y, img = blockIdx.z;
  const int T = p.T, ld = 3 * p.W;
  const __nv_bfloat16* base = p.qkv + static_cast<size_t>(img) * T * ld + head * kVitHeadDim;
  const int q0 = qt * kVaQ;
  constexpr int kChunks = kVitHeadDim / 8;                         // 16-byte chunks per row: 10

  // Q tile (rows past T zero-filled), then K/V tile 0
  for (int i = threadIdx.x; i < kVaQ * kChunks; i += blockDim.x) {
    const int r = i / kChunks, c = i - r * kChunks;
    const bool ok = q0 + r < T;
    cp_async16(sq + r * kVaPitch + c * 8, base + static_cast<size_t>(ok ? q0 + r : 0) * ld + c * 8, ok);
  }
  auto load_kv = [&](int tile, int stage) {
    __nv_bfloat16* sk = skv + stage * 2 * kVaK * kVaPitch;
    __nv_bfloat16* sv = sk + kVaK * kVaPitch;
    const int k0 = tile * kVaK;
    for (int i = threadIdx.x; i < 2 * kVaK * kChunks; i += blockDim.x) {
      const int which = i / (kVaK * kChunks), j = i - which * kVaK * kChunks, r = j / kChunks, c = j - r * kChunks;
      const bool ok = k0 + r < T;
      cp_async16((which ? sv : sk) + r * kVaPitch + c * 8, base + static_cast<size_t>(ok ? k0 + r : 0) * ld + (1 + which) * p.W + c * 8, ok);
    }
  };
  const int ntiles = (T + kVaK - 1) / kVaK;
  load_kv(0, 0);
  cp_async_commit();
  if (ntiles > 1) load_kv(1, 1);
  cp_async_commit();

  float o[kVitHeadDim / 8][4];
#pragma unroll
  for (int ni = 0; ni < kVitHeadDim / 8; ++ni)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[ni][e] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  uint32_t aq[kVitHeadDim / 16][4];

  int stage = 0;
  for (int tile = 0; tile < ntiles; ++tile) {
    cp_async_wait<1>();                                            // everything but the newest group: this tile's (and Q's) bytes have landed
    __syncthreads();                                               // ... for every thread; and every warp is done with tile - 1, whose stage is refilled now
    {
      const int nxt = stage == 0 ? kVaStages - 1 : stage - 1;      // stage of tile - 1 = stage of tile + 2
      if (tile + 2 < ntiles) load_kv(tile + 2, nxt);
      cp_async_commit();
    }
    if (tile == 0) {
#pragma unroll
      for (int ki = 0; ki < kVitHeadDim / 16; ++ki) ldsm_x4(aq[ki], sq + (warp * 16 + (lm & 1) * 8 + lr) * kVaPitch + ki * 16 + (lm >> 1) * 8);
    }
    const __nv_bfloat16* sk = skv + stage * 2 * kVaK * kVaPitch;
    const __nv_bfloat16* sv = sk + kVaK * kVaPitch;
    float s[kVaK / 8][4];
#pragma unroll
    for (int ni = 0; ni < kVaK / 8; ++ni)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[ni][e] = 0.f;
#pragma unroll
    for (int ki = 0; ki < kVitHeadDim / 16; ++ki) {
#pragma unroll
      for (int np = 0; np < kVaK / 16; ++np) {
        uint32_t bk[4];
        ldsm_x4(bk, sk + (np * 16 + (lm >> 1) * 8 + lr) * kVaPitch + ki * 16 + (lm & 1) * 8);
        mma_bf16_16816(s[np * 2], aq[ki], bk[0], bk[1]);
        mma_bf16_16816(s[np * 2 + 1], aq[ki], bk[2], bk[3]);
      }
    }
    // online softmax on this warp's 16 rows (thread: rows g and g + 8, columns ni * 8 + 2 t + {0, 1}); the 1 / sqrt(80) * log2(e) factor is
    // folded into the exponent: p = 2^(s * c - m * c) is one FFMA + one MUFU per score
    const int kbase = tile * kVaK;
    const bool ragged = kbase + kVaK > T;
    if (ragged) {
#pragma unroll
      for (int ni = 0; ni < kVaK / 8; ++ni)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (kbase + ni * 8 + 2 * t + (e & 1) >= T) s[ni][e] = -INFINITY;
    }
    // both row halves at once, reductions as trees of independent partials (a 16-deep FMNMX / FADD chain per row costs more issue
    // latency than the tensor work of the tile hides at four warps per scheduler)
    float mx[2][4], sm[2][4];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int q = 0; q < 4; ++q) mx[h][q] = fmaxf(fmaxf(s[2 * q][h * 2], s[2 * q][h * 2 + 1]), fmaxf(s[2 * q + 1][h * 2], s[2 * q + 1][h * 2 + 1]));
    float m_new[2], corr[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float m = fmaxf(fmaxf(mx[h][0], mx[h][1]), fmaxf(mx[h][2], mx[h][3]));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      m_new[h] = fmaxf(m_run[h], m * p.scale_log2e);               // finite: every tile holds at least one valid key (scale > 0)
      corr[h] = ex2_approx(m_run[h] - m_new[h]);
      m_run[h] = m_new[h];
    }
#pragma unroll
    for (int ni = 0; ni < kVaK / 8; ++ni)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[ni][e] = ex2_approx(fmaf(s[ni][e], p.scale_log2e, -m_new[e >> 1]));
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int q = 0; q < 4; ++q) sm[h][q] = (s[2 * q][h * 2] + s[2 * q][h * 2 + 1]) + (s[2 * q + 1][h * 2] + s[2 * q + 1][h * 2 + 1]);
      float sum = (sm[h][0] + sm[h][1]) + (sm[h][2] + sm[h][3]);
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      l_run[h] = l_run[h] * corr[h] + sum;
    }
#pragma unroll
    for (int ni = 0; ni < kVitHeadDim / 8; ++ni) { o[ni][0] *= corr[0]; o[ni][1] *= corr[0]; o[ni][2] *= corr[1]; o[ni][3] *= corr[1]; }
    // O += P V
#pragma unroll
    for (int ki = 0; ki < kVaK / 16; ++ki) {
      uint32_t ap[4];
      ap[0] = pack_bf16x2(s[2 * ki][0], s[2 * ki][1]);
      ap[1] = pack_bf16x2(s[2 * ki][2], s[2 * ki][3]);
      ap[2] = pack_bf16x2(s[2 * ki + 1][0], s[2 * ki + 1][1]);
      ap[3] = pack_bf16x2(s[2 * ki + 1][2], s[2 * ki + 1][3]);
#pragma unroll
      for (int np = 0; np < kVitHeadDim / 16; ++np) {
        uint32_t bv[4];
        ldsm_x4_t(bv, sv + (ki * 16 + (lm & 1) * 8 + lr) * kVaPitch + np * 16 + (lm >> 1) * 8);
        mma_bf16_16816(o[np * 2], ap, bv[0], bv[1]);
        mma_bf16_16816(o[np * 2 + 1], ap, bv[2], bv[3]);
      }
    }
    stage = stage + 1 == kVaStages ? 0 : stage + 1;
  }
  // normalise, stage through the Q tile (each warp reads and writes only its own 16 rows of it), store 16-byte chunks
  __nv_bfloat16* so = sq + warp * 16 * kVaPitch;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float inv = 1.0f / l_run[h];
#pragma unroll
    for (int ni = 0; ni < kVitHeadDim / 8; ++ni)
      *reinterpret_cast<uint32_t*>(so + (h * 8 + g) * kVaPitch + ni * 8 + 2 * t) = pack_bf16x2(o[ni][h * 2] * inv, o[ni][h * 2 + 1] * inv);
  }
  __syncwarp();
  __nv_bfloat16* dst = p.out + static_cast<size_t>(img) * T * p.W + head * kVitHeadDim;
  for (int i = lane; i < 16 * kChunks; i += 32) {
    const int r = i / kChunks, c = i - r * kChunks;
    const int row = q0 + warp * 16 + r;
    if (row < T) *reinterpret_cast<uint4*>(dst + static_cast<size_t>(row) * p.W + c * 8) = *reinterpret_cast<const uint4*>(so + r * kVaPitch + c * 8);
  }
}

// embed[b, :] /= ||embed[b, :]|| (fp32; embedders.py:764).  One warp per row.
__global__ void __launch_bounds__(256) vit_normalize_kernel(float* __restrict__ e, int rows, int F) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* r = e + static_cast<size_t>(row) * F;
  float s = 0.f;
  for (int i = lane; i < F; i += 32) s = fmaf(r[i], r[i], s);
  s = warp_sum(s);
  const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);
  for (int i = lane; i < F; i += 32) r[i] *= inv;
}

}  // namespace novic
