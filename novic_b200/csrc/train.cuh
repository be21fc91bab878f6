// Backward-pass kernels of the teacher-forced training step (train.py:1263-1286 -> embedding_decoder.py:659-761):
// everything that is not a GEMM.  The GEMMs (dgrad: dX = dY * W, wgrad: dW = dY^T * X) run on the same persistent
// tcgen05 kernel as the forward pass with the epilogues at the bottom of this file.
//
// Gradient tensors that flow along the residual stream are kept in three formats, each produced where it is cheap:
//   fp32 "blocked" [R, 512]  (same layout as x: coalesced for thread-per-row kernels; the accumulation format)
//   bf16 row-major [R, 512]  (A operand of the next dgrad GEMM)
//   bf16 transposed [512, R] (A operand of the next wgrad GEMM, whose contraction runs over the R rows)
#pragma once

#include "kernels.cuh"

namespace novic {

// ---------------------------------------------------------------------------------------------------------
// bf16 transpose: dst[c, r] = src[r, c]   (src row-major [R, C] with leading dimension ld_src; dst [C, ld_dst])
// ---------------------------------------------------------------------------------------------------------
// 64 x 64 tile per CTA.  Two source rows x 8 columns per thread (two 16-byte loads), stored to shared memory as (row, row + 1) pairs
// in transposed position; read back as 16 bytes = 8 consecutive source rows of one column.  Pitch 33 words: both sides conflict-free
// or 2-way.  Unaligned or ragged edges take the element-wise path.  (The first version moved single bf16 values: 15.6 us per
// [19 456, 512] matrix against 5.4 us of HBM time.)
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16* __restrict__ src, int R, int C, int ld_src,
                                                             __nv_bfloat16* __restrict__ dst, int ld_dst) {
  __shared__ uint32_t tile[64 * 33];                       // [column][row pair]
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int t = threadIdx.x;
  {
    const int chunk = t & 7, rp = t >> 3;                  // columns [8 chunk, 8 chunk + 8), rows 2 rp and 2 rp + 1
    const int c = c0 + chunk * 8;
    uint4 v[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = r0 + 2 * rp + h;
      const __nv_bfloat16* row = src + static_cast<size_t>(r) * ld_src + c;
      v[h] = make_uint4(0u, 0u, 0u, 0u);
      if (r < R) {
        if (c + 8 <= C && (reinterpret_cast<uintptr_t>(row) & 15) == 0) {
          v[h] = *reinterpret_cast<const uint4*>(row);
        } else {
          uint16_t e[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) e[j] = (c + j < C) ? reinterpret_cast<const uint16_t*>(row)[j] : uint16_t(0);
          v[h] = make_uint4(e[0] | (uint32_t(e[1]) << 16), e[2] | (uint32_t(e[3]) << 16), e[4] | (uint32_t(e[5]) << 16), e[6] | (uint32_t(e[7]) << 16));
        }
      }
    }
    const uint32_t lo[4] = {v[0].x, v[0].y, v[0].z, v[0].w}, hi[4] = {v[1].x, v[1].y, v[1].z, v[1].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      tile[(chunk * 8 + 2 * j) * 33 + rp] = __byte_perm(lo[j], hi[j], 0x5410);       // column 2j: (row, row + 1)
      tile[(chunk * 8 + 2 * j + 1) * 33 + rp] = __byte_perm(lo[j], hi[j], 0x7632);   // column 2j + 1
    }
  }
  __syncthreads();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int cl = (t >> 3) + 32 * h, rch = t & 7;         // output row (source column) cl, source rows [8 rch, 8 rch + 8)
    const int c = c0 + cl, r = r0 + rch * 8;
    if (c >= C || r >= R) continue;
    const uint32_t* w = &tile[cl * 33 + rch * 4];
    const uint4 v = make_uint4(w[0], w[1], w[2], w[3]);
    __nv_bfloat16* out = dst + static_cast<size_t>(c) * ld_dst + r;
    if (r + 8 <= R && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
      *reinterpret_cast<uint4*>(out) = v;
    } else {
      const uint32_t e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (r + j < R) reinterpret_cast<uint16_t*>(out)[j] = static_cast<uint16_t>(e[j >> 1] >> ((j & 1) * 16));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm backward + residual merge.  Thread = row (blocked fp32 tensors are coalesced that way).
//   y = (x - mean) * rstd * g          dxhat = dy * g
//   dx = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat))           dg[c] += sum_rows dy * xhat
//   out = dx + (resid ? resid : 0)     written as fp32 blocked, bf16 row-major (via smem staging) and bf16 transposed
// ---------------------------------------------------------------------------------------------------------
struct LnBwdParams {
  const float* x;        // blocked: the LayerNorm input saved by the forward pass
  const float* dy;       // blocked: gradient w.r.t. the LayerNorm output
  const float* resid;    // blocked or nullptr: gradient arriving through the residual connection
  const float* gain;     // [512]
  float* out;            // blocked
  __nv_bfloat16* out_bf; // [R, 512] row-major or nullptr
  __nv_bfloat16* out_t;  // [512, ld_t] transposed or nullptr
  float* dgain;          // [512] accumulated with atomics
  int R, ld_t;
  float eps;
  DropCfg drop;          // thresh != 0: the bf16 copies are multiplied by the dropout mask of the branch they feed (site drop_site, element
  uint32_t drop_site;    // index row * 512 + column); the fp32 residual gradient `out` is never masked
};

// One CTA = 32 rows x 16 warps; warp w owns columns [32 w, 32 w + 32) of them and loads its x, dy and residual values into registers
// up front, so every tensor is read exactly once and all of a CTA's loads are in flight together; the row reductions go through
// shared memory.  (The first version gave a whole row to one thread - 225 us per launch; the second re-read x three times from four
// warps - 157 us; 8 warps x 64 columns with the residual loaded late - 73 us.)
// Column sums of the warp's 32 rows by recursive halving: 31 shuffles for 32 columns; lane l ends up holding column l in v[0].
__device__ __forceinline__ void warp_column_sums32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float keep = upper ? v[i + off] : v[i];
      const float send = upper ? v[i] : v[i + off];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}

constexpr int kLnBwdWarps = 16;
constexpr int kLnBwdCols = kE / kLnBwdWarps;            // 32

__global__ void __launch_bounds__(kLnBwdWarps * 32, 1) ln_bwd_kernel(const LnBwdParams p) {
  __shared__ float s_red[4][kLnBwdWarps][32];           // [quantity][warp][row]
  __shared__ __align__(16) uint8_t s_stage[kLnBwdWarps][32 * kLnBwdCols * 2];   // bf16 rows of 64 B, 16-byte chunks swizzled by the row pair
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int row0 = blockIdx.x * 32;
  const int row = row0 + lane;
  const bool ok = row < p.R;
  const int c0 = warp * kLnBwdCols;
  constexpr int NC = kLnBwdCols, NQ = NC / 4;
  float x[NC], d[NC], r[NC];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), w = v, u = v;
    if (ok) {
      v = *reinterpret_cast<const float4*>(p.x + xblk_off(row, (c0 >> 2) + q));
      w = *reinterpret_cast<const float4*>(p.dy + xblk_off(row, (c0 >> 2) + q));
      if (p.resid != nullptr) u = *reinterpret_cast<const float4*>(p.resid + xblk_off(row, (c0 >> 2) + q));
    }
    x[q * 4] = v.x, x[q * 4 + 1] = v.y, x[q * 4 + 2] = v.z, x[q * 4 + 3] = v.w;
    d[q * 4] = w.x, d[q * 4 + 1] = w.y, d[q * 4 + 2] = w.z, d[q * 4 + 3] = w.w;
    r[q * 4] = u.x, r[q * 4 + 1] = u.y, r[q * 4 + 2] = u.z, r[q * 4 + 3] = u.w;
  }
  float sum = 0.f, sumsq = 0.f;
#pragma unroll
  for (int i = 0; i < NC; ++i) sum += x[i], sumsq += x[i] * x[i];
  s_red[0][warp][lane] = sum;
  s_red[1][warp][lane] = sumsq;
  __syncthreads();
  sum = 0.f, sumsq = 0.f;
#pragma unroll
  for (int w = 0; w < kLnBwdWarps; ++w) sum += s_red[0][w][lane], sumsq += s_red[1][w][lane];
  const float mean = sum * (1.0f / kE);
  const float rstd = rsqrtf(fmaxf(sumsq * (1.0f / kE) - mean * mean, 0.f) + p.eps);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const float4 g = *reinterpret_cast<const float4*>(p.gain + c0 + q * 4);
    const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float xhat = (x[q * 4 + i] - mean) * rstd;
      const float a = d[q * 4 + i] * gv[i];
      x[q * 4 + i] = xhat;                              // x now holds xhat
      s1 += a;
      s2 += a * xhat;
    }
  }
  s_red[2][warp][lane] = s1;
  s_red[3][warp][lane] = s2;
  __syncthreads();
  s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int w = 0; w < kLnBwdWarps; ++w) s1 += s_red[2][w][lane], s2 += s_red[3][w][lane];
  const float m1 = s1 * (1.0f / kE), m2 = s2 * (1.0f / kE);
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const float4 g = *reinterpret_cast<const float4*>(p.gain + c0 + q * 4);
    const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float xhat = x[q * 4 + i], dv = d[q * 4 + i];
      d[q * 4 + i] = (ok ? rstd * (dv * gv[i] - m1 - xhat * m2) : 0.f) + r[q * 4 + i];   // d now holds the output row
      x[q * 4 + i] = ok ? dv * xhat : 0.f;                                              // x now holds this row's dgain terms
    }
    if (ok) *reinterpret_cast<float4*>(p.out + xblk_off(row, (c0 >> 2) + q)) = make_float4(d[q * 4], d[q * 4 + 1], d[q * 4 + 2], d[q * 4 + 3]);
  }
  if (p.drop.thresh != 0u) {
#pragma unroll
    for (int i = 0; i < NC; ++i) d[i] *= drop_factor(p.drop, p.drop_site, static_cast<uint32_t>(row) * kE + c0 + i);
  }
  if (p.out_t != nullptr && ok) {
#pragma unroll
    for (int i = 0; i < NC; ++i) p.out_t[static_cast<size_t>(c0 + i) * p.ld_t + row] = __float2bfloat16_rn(d[i]);
  }
  if (p.out_bf != nullptr) {
    uint8_t* stage = s_stage[warp];
#pragma unroll
    for (int q = 0; q < NC / 8; ++q)
      *reinterpret_cast<uint4*>(stage + lane * (NC * 2) + ((q ^ ((lane >> 1) & 3)) << 4)) =
          make_uint4(pack_bf16x2(d[q * 8], d[q * 8 + 1]), pack_bf16x2(d[q * 8 + 2], d[q * 8 + 3]), pack_bf16x2(d[q * 8 + 4], d[q * 8 + 5]), pack_bf16x2(d[q * 8 + 6], d[q * 8 + 7]));
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rr = i * 8 + (lane >> 2), ch = lane & 3;
      const uint4 v = *reinterpret_cast<const uint4*>(stage + rr * (NC * 2) + ((ch ^ ((rr >> 1) & 3)) << 4));
      if (row0 + rr < p.R) *reinterpret_cast<uint4*>(p.out_bf + static_cast<size_t>(row0 + rr) * kE + c0 + ch * 8) = v;
    }
  }
  warp_column_sums32(x, lane);
  atomicAdd(p.dgain + c0 + lane, x[0]);
}

// ---------------------------------------------------------------------------------------------------------
// Dropout helpers of the training step (elementwise, in place).  The masks of the residual branches and of the activated feed-forward
// rows are applied where the masked tensor is produced (ln_bwd_kernel, gelu_bwd_kernel, EpiGeluTrain, the row kernel's DROP instantiation);
// separate masking passes over the bf16 GEMM operands cost 24 launches and 0.3 ms per step.
//   drop_mask_blocked_kernel : x[r, c] *= factor(site, r * 512 + c) on the blocked fp32 residual layout (gradient w.r.t. the input embedding rows).
//   input_dropout_ln_kernel : x = dropout(x) in place, xn = LayerNorm(x) * gain (forward, input dropout embedding_decoder.py:1297).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) drop_mask_blocked_kernel(float* __restrict__ x, int rows, DropCfg d, uint32_t site) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(rows) * kE) return;
  const int r = static_cast<int>(i / kE), c = static_cast<int>(i - static_cast<size_t>(r) * kE);
  float* e = x + xblk_off(r, c >> 2) + (c & 3);
  *e *= drop_factor(d, site, static_cast<uint32_t>(i));
}

__global__ void __launch_bounds__(128) input_dropout_ln_kernel(float* __restrict__ x, __nv_bfloat16* __restrict__ xn, const float* __restrict__ gain, int rows,
                                                               float eps, DropCfg d, uint32_t site) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = lane_id();
  float v[16];
  float sum = 0.f, sumsq = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float4 t = *reinterpret_cast<const float4*>(x + xblk_off(row, lane * 4 + q));
    const uint32_t c = static_cast<uint32_t>(lane * 16 + q * 4), base = static_cast<uint32_t>(row) * kE + c;
    t.x *= drop_factor(d, site, base); t.y *= drop_factor(d, site, base + 1); t.z *= drop_factor(d, site, base + 2); t.w *= drop_factor(d, site, base + 3);
    *reinterpret_cast<float4*>(x + xblk_off(row, lane * 4 + q)) = t;
    v[q * 4] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
    sum += t.x + t.y + t.z + t.w;
    sumsq += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
  }
  sum = warp_sum(sum);
  sumsq = warp_sum(sumsq);
  const float mean = sum * (1.0f / kE);
  const float rstd = rsqrtf(fmaxf(sumsq * (1.0f / kE) - mean * mean, 0.f) + eps);
  uint32_t o[8];
#pragma unroll
  for (int k = 0; k < 8; ++k)
    o[k] = pack_bf16x2((v[2 * k] - mean) * rstd * __ldg(gain + lane * 16 + 2 * k), (v[2 * k + 1] - mean) * rstd * __ldg(gain + lane * 16 + 2 * k + 1));
  uint4* dst = reinterpret_cast<uint4*>(xn + static_cast<size_t>(row) * kE) + lane * 2;
  dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
  dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

// ---------------------------------------------------------------------------------------------------------
// GELU backward, elementwise on row-major bf16: dpre = dh * (Phi(pre) + pre * phi(pre))
// ---------------------------------------------------------------------------------------------------------
// drop.thresh != 0: the activated rows were dropped in the forward pass (site drop_site, element index = flat index): dh is masked first.
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const __nv_bfloat16* __restrict__ dh, const __nv_bfloat16* __restrict__ pre,
                                                       __nv_bfloat16* __restrict__ dpre, size_t n, DropCfg drop, uint32_t drop_site) {
  const size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i >= n) return;
  const uint4 a = *reinterpret_cast<const uint4*>(dh + i), b = *reinterpret_cast<const uint4*>(pre + i);
  float d[8], x[8];
  bf16x8_to_f32(a, d);
  bf16x8_to_f32(b, x);
  if (drop.thresh != 0u) {
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] *= drop_factor(drop, drop_site, static_cast<uint32_t>(i) + k);
  }
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 8; k += 2) {
    float r[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float v = x[k + j];
      const float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752440f));
      const float pdf = 0.3989422804014327f * __expf(-0.5f * v * v);
      r[j] = d[k + j] * (cdf + v * pdf);
    }
    o[k / 2] = pack_bf16x2(r[0], r[1]);
  }
  *reinterpret_cast<uint4*>(dpre + i) = make_uint4(o[0], o[1], o[2], o[3]);
}

// ---------------------------------------------------------------------------------------------------------
// Attention backward for the teacher-forced pass (S <= 32 positions per sequence, 8 heads x 64).
//   P = softmax(Q K^T / 8 + mask)   dV = P^T dO   dP = dO V^T   dS = P o (dP - rowsum(dP o P))   dQ = dS K / 8   dK = dS^T Q / 8
// One warp per (sequence, head).  The five products are 32 x 32 x 64 or 32 x 64 x 32: they run on mma.sync m16n8k16 (bf16 in, fp32
// accumulate) from ldmatrix fragments - they are far too small for a tcgen05 tile, and far too many instructions as scalar FMAs:
// the scalar version, lanes owning keys for the scores and channels for the accumulation, issued ~15 k instructions per item and took
// 273 us per launch (702 us with fp32 shared-memory operands) against ~20 us of HBM time.  P (with the dropout factor) and dS pass
// through shared memory as bf16 to become the A operands of the second group of products.
// ---------------------------------------------------------------------------------------------------------
constexpr int kAttnBwdMaxS = 32;
constexpr int kAttnBwdWarps = 2;                  // warps (items) per CTA
constexpr int kAbPitch = kHeadDim + 8;            // q / k / v / dout rows: 64 bf16 + 16 B pad (ldmatrix rows land in distinct bank groups)
constexpr int kAbPPitch = kAttnBwdMaxS + 8;       // P / dS rows: 32 bf16 + 16 B pad
constexpr int kAbWarpBytes = (4 * kAttnBwdMaxS * kAbPitch + 2 * kAttnBwdMaxS * kAbPPitch) * 2;   // 23 552 B
__host__ __device__ constexpr int attn_bwd_smem_bytes(int /*S*/) { return kAttnBwdWarps * kAbWarpBytes; }

struct AttnBwdParams {
  const __nv_bfloat16* q;       // [nseq * S, 512]
  const __nv_bfloat16* kcache;  // [nseq, smax, 512]
  const __nv_bfloat16* vcache;
  const __nv_bfloat16* dout;    // [nseq * S, 512] gradient w.r.t. the attention output
  __nv_bfloat16* dqkv;          // [nseq * S, 1536] row-major: dq | dk | dv
  const unsigned char* keypad;  // optional [nseq, S]
  int nseq, S, smax, P, prefix_bidir;
  float scale;                  // 1 / sqrt(64)
  DropCfg drop;                 // dropout on the attention probabilities, same masks as the forward pass
  uint32_t drop_site;
};

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// out[32, 64] = A' B with the contraction over 32: A' = A (A_TRANS = false, sa is [m][k], pitch kAbPPitch) or A^T (sa is [k][m]);
// B is [k][64] row-major (pitch kAbPitch).  The result goes to `stage` (bf16, pitch kAbPitch) and from there, rows < S, to
// dst + row * 3 * kE as 16-byte stores.
template <bool A_TRANS>
__device__ __forceinline__ void attn_bwd_product(const __nv_bfloat16* sa, const __nv_bfloat16* sb, __nv_bfloat16* stage, __nv_bfloat16* dst, int S, int lane) {
  const int lm = lane >> 3, lr = lane & 7, g = lane >> 2, t = lane & 3;
  float acc[2][8][4];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 8; ++ni)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mi][ni][e] = 0.f;
#pragma unroll
  for (int ki = 0; ki < 2; ++ki) {
    uint32_t a[2][4];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      if (A_TRANS) ldsm_x4_t(a[mi], sa + (ki * 16 + (lm >> 1) * 8 + lr) * kAbPPitch + mi * 16 + (lm & 1) * 8);
      else ldsm_x4(a[mi], sa + (mi * 16 + (lm & 1) * 8 + lr) * kAbPPitch + ki * 16 + (lm >> 1) * 8);
    }
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4_t(b, sb + (ki * 16 + (lm & 1) * 8 + lr) * kAbPitch + np * 16 + (lm >> 1) * 8);
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        mma_bf16_16816(acc[mi][np * 2], a[mi], b[0], b[1]);
        mma_bf16_16816(acc[mi][np * 2 + 1], a[mi], b[2], b[3]);
      }
    }
  }
  __syncwarp();
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 8; ++ni)
#pragma unroll
      for (int h = 0; h < 2; ++h)
        *reinterpret_cast<uint32_t*>(stage + (mi * 16 + h * 8 + g) * kAbPitch + ni * 8 + 2 * t) = pack_bf16x2(acc[mi][ni][h * 2], acc[mi][ni][h * 2 + 1]);
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + lm, ch = lr;
    if (r < S) *reinterpret_cast<uint4*>(dst + static_cast<size_t>(r) * (3 * kE) + ch * 8) = *reinterpret_cast<const uint4*>(stage + r * kAbPitch + ch * 8);
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kAttnBwdWarps * 32) attn_bwd_kernel(const AttnBwdParams p) {
  extern __shared__ __align__(16) uint8_t sm_attn_b[];
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int item = blockIdx.x * kAttnBwdWarps + warp;
  if (item >= p.nseq * kHeads) return;
  const int a = item / kHeads, head = item - a * kHeads;
  const int S = p.S;
  constexpr int PT = kAbPitch, PP = kAbPPitch;
  __nv_bfloat16* sq = reinterpret_cast<__nv_bfloat16*>(sm_attn_b + static_cast<size_t>(warp) * kAbWarpBytes);
  __nv_bfloat16* sk = sq + kAttnBwdMaxS * PT;
  __nv_bfloat16* sv = sk + kAttnBwdMaxS * PT;     // after the score products: staging for the outputs
  __nv_bfloat16* sdo = sv + kAttnBwdMaxS * PT;
  __nv_bfloat16* sp = sdo + kAttnBwdMaxS * PT;    // P o dropout, [i][j]
  __nv_bfloat16* sds = sp + kAttnBwdMaxS * PP;    // dS, [i][j]
  const int lm = lane >> 3, lr = lane & 7, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int s = it * 4 + lm;
    uint4 vq = make_uint4(0u, 0u, 0u, 0u), vk = vq, vv = vq, vd = vq;
    if (s < S) {
      const size_t rq = (static_cast<size_t>(a) * S + s) * kE + head * kHeadDim + lr * 8;
      const size_t rk = (static_cast<size_t>(a) * p.smax + s) * kE + head * kHeadDim + lr * 8;
      vq = *reinterpret_cast<const uint4*>(p.q + rq);
      vk = *reinterpret_cast<const uint4*>(p.kcache + rk);
      vv = *reinterpret_cast<const uint4*>(p.vcache + rk);
      vd = *reinterpret_cast<const uint4*>(p.dout + rq);
    }
    *reinterpret_cast<uint4*>(sq + s * PT + lr * 8) = vq;
    *reinterpret_cast<uint4*>(sk + s * PT + lr * 8) = vk;
    *reinterpret_cast<uint4*>(sv + s * PT + lr * 8) = vv;
    *reinterpret_cast<uint4*>(sdo + s * PT + lr * 8) = vd;
  }
  const bool key_ok = lane < S && !(p.keypad != nullptr && lane > 0 && p.keypad[static_cast<size_t>(a) * S + lane]);
  const uint32_t kmask = __ballot_sync(0xffffffffu, key_ok);
  __syncwarp();
  // scores Q K^T and dP = dO V^T: [32 queries] x [32 keys], contraction over the 64 channels
  float sc[2][4][4], dp[2][4][4];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
      for (int e = 0; e < 4; ++e) sc[mi][ni][e] = 0.f, dp[mi][ni][e] = 0.f;
#pragma unroll
  for (int ki = 0; ki < 4; ++ki) {
    uint32_t aq[2][4], ad[2][4];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      ldsm_x4(aq[mi], sq + (mi * 16 + (lm & 1) * 8 + lr) * PT + ki * 16 + (lm >> 1) * 8);
      ldsm_x4(ad[mi], sdo + (mi * 16 + (lm & 1) * 8 + lr) * PT + ki * 16 + (lm >> 1) * 8);
    }
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t bk[4], bv[4];
      ldsm_x4(bk, sk + (np * 16 + (lm >> 1) * 8 + lr) * PT + ki * 16 + (lm & 1) * 8);
      ldsm_x4(bv, sv + (np * 16 + (lm >> 1) * 8 + lr) * PT + ki * 16 + (lm & 1) * 8);
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        mma_bf16_16816(sc[mi][np * 2], aq[mi], bk[0], bk[1]);
        mma_bf16_16816(sc[mi][np * 2 + 1], aq[mi], bk[2], bk[3]);
        mma_bf16_16816(dp[mi][np * 2], ad[mi], bv[0], bv[1]);
        mma_bf16_16816(dp[mi][np * 2 + 1], ad[mi], bv[2], bv[3]);
      }
    }
  }
  // softmax, dropout and dS per query row: thread (g, t) holds keys ni * 8 + 2 t + {0, 1} of rows mi * 16 + h * 8 + g
#pragma unroll
  for (int mi = 0; mi < 2; ++mi) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = mi * 16 + h * 8 + g;
      int nkeys = (p.prefix_bidir && i < p.P) ? p.P : i + 1;
      if (i >= S) nkeys = 0;
      const uint32_t rowmask = kmask & (nkeys >= 32 ? 0xffffffffu : ((1u << nkeys) - 1u));
      float mx = -INFINITY;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = ni * 8 + 2 * t + e;
          const float v = ((rowmask >> j) & 1u) ? sc[mi][ni][h * 2 + e] * p.scale : -INFINITY;
          sc[mi][ni][h * 2 + e] = v;
          mx = fmaxf(mx, v);
        }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float sum = 0.f;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = ni * 8 + 2 * t + e;
          const float ex = ((rowmask >> j) & 1u) ? __expf(sc[mi][ni][h * 2 + e] - mx) : 0.f;
          sc[mi][ni][h * 2 + e] = ex;
          sum += ex;
        }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float inv = sum > 0.f ? 1.0f / sum : 0.f;
      // out_i = sum_j p_ij m_ij v_j with m the dropout factor: d p_ij = m_ij (do_i . v_j), and dv_j receives p_ij m_ij do_i
      float dsum = 0.f;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        float pd[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = ni * 8 + 2 * t + e;
          const float pj = sc[mi][ni][h * 2 + e] * inv;
          float dm = 1.f;
          if (p.drop.thresh != 0u && ((rowmask >> j) & 1u)) dm = drop_factor(p.drop, p.drop_site, ((static_cast<uint32_t>(a) * kHeads + head) * S + i) * S + j);
          const float dpv = dp[mi][ni][h * 2 + e] * dm;
          sc[mi][ni][h * 2 + e] = pj;
          dp[mi][ni][h * 2 + e] = dpv;
          dsum += pj * dpv;
          pd[e] = pj * dm;
        }
        *reinterpret_cast<uint32_t*>(sp + i * PP + ni * 8 + 2 * t) = pack_bf16x2(pd[0], pd[1]);
      }
      dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
      dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const float d0 = sc[mi][ni][h * 2] * (dp[mi][ni][h * 2] - dsum) * p.scale;       // gradient w.r.t. q_i . k_j
        const float d1 = sc[mi][ni][h * 2 + 1] * (dp[mi][ni][h * 2 + 1] - dsum) * p.scale;
        *reinterpret_cast<uint32_t*>(sds + i * PP + ni * 8 + 2 * t) = pack_bf16x2(d0, d1);
      }
    }
  }
  __syncwarp();
  __nv_bfloat16* out = p.dqkv + static_cast<size_t>(a) * S * (3 * kE) + head * kHeadDim;
  attn_bwd_product<true>(sp, sdo, sv, out + 2 * kE, S, lane);    // dv_j = sum_i (p m)_ij do_i
  attn_bwd_product<true>(sds, sq, sv, out + kE, S, lane);        // dk_j = sum_i ds_ij q_i
  attn_bwd_product<false>(sds, sk, sv, out, S, lane);            // dq_i = sum_j ds_ij k_j
}

// ---------------------------------------------------------------------------------------------------------
// Attention forward of the teacher-forced pass, same tiling as the backward kernel above: one warp per (sequence, head), all S <= 32
// queries at once, Q K^T and P V on mma.sync tiles; the probabilities go from the score accumulators straight into the A fragments of
// the second product.  Replaces the key-by-key online-softmax kernel (attention_bulk_kernel, 138 us per launch; it carried the dropout too until then) in the training step.
// Requires q0 = 0, one beam, no ancestor mask.  Dropout masks the numerator only (torch applies it to the normalised probabilities).
// ---------------------------------------------------------------------------------------------------------
constexpr int kAttnTfWarps = 2;
constexpr int kAttnTfWarpBytes = 3 * kAttnBwdMaxS * kAbPitch * 2;   // 13 824 B
constexpr int kAttnTfSmemBytes = kAttnTfWarps * kAttnTfWarpBytes;

template <bool DROP>
__global__ void __launch_bounds__(kAttnTfWarps * 32) attention_tf_kernel(const AttnParams p) {
  extern __shared__ __align__(16) uint8_t sm_attn_f[];
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int item = blockIdx.x * kAttnTfWarps + warp;
  if (item >= p.nseq * kHeads) return;
  const int a = item / kHeads, head = item - a * kHeads;
  const int S = p.nq;
  const bool small = S <= 16;                      // one 16-row query tile and one 16-key tile are enough (warp-uniform)
  constexpr int PT = kAbPitch;
  __nv_bfloat16* sq = reinterpret_cast<__nv_bfloat16*>(sm_attn_f + static_cast<size_t>(warp) * kAttnTfWarpBytes);   // later: output staging
  __nv_bfloat16* sk = sq + kAttnBwdMaxS * PT;
  __nv_bfloat16* sv = sk + kAttnBwdMaxS * PT;
  const int lm = lane >> 3, lr = lane & 7, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    if (it >= 4 && small) continue;
    const int s = it * 4 + lm;
    uint4 vq = make_uint4(0u, 0u, 0u, 0u), vk = vq, vv = vq;
    if (s < S) {
      const size_t rq = (static_cast<size_t>(a) * S + s) * kE + head * kHeadDim + lr * 8;
      const size_t rk = (static_cast<size_t>(a) * p.smax + s) * kE + head * kHeadDim + lr * 8;
      vq = *reinterpret_cast<const uint4*>(p.q + rq);
      vk = *reinterpret_cast<const uint4*>(p.kcache + rk);
      vv = *reinterpret_cast<const uint4*>(p.vcache + rk);
    }
    *reinterpret_cast<uint4*>(sq + s * PT + lr * 8) = vq;
    *reinterpret_cast<uint4*>(sk + s * PT + lr * 8) = vk;
    *reinterpret_cast<uint4*>(sv + s * PT + lr * 8) = vv;
  }
  const bool key_ok = lane < S && !(p.keypad != nullptr && lane > 0 && p.keypad[static_cast<size_t>(a) * p.keypad_ld + lane]);
  const uint32_t kmask = __ballot_sync(0xffffffffu, key_ok);
  __syncwarp();
  float sc[2][4][4];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
      for (int e = 0; e < 4; ++e) sc[mi][ni][e] = 0.f;
#pragma unroll
  for (int ki = 0; ki < 4; ++ki) {
    uint32_t aq[2][4];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
      if (mi == 0 || !small) ldsm_x4(aq[mi], sq + (mi * 16 + (lm & 1) * 8 + lr) * PT + ki * 16 + (lm >> 1) * 8);
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      if (np == 1 && small) continue;
      uint32_t bk[4];
      ldsm_x4(bk, sk + (np * 16 + (lm >> 1) * 8 + lr) * PT + ki * 16 + (lm & 1) * 8);
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        if (mi == 1 && small) continue;
        mma_bf16_16816(sc[mi][np * 2], aq[mi], bk[0], bk[1]);
        mma_bf16_16816(sc[mi][np * 2 + 1], aq[mi], bk[2], bk[3]);
      }
    }
  }
#pragma unroll
  for (int mi = 0; mi < 2; ++mi) {
    if (mi == 1 && small) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = mi * 16 + h * 8 + g;
      int nkeys = (p.prefix_bidir && i < p.P) ? p.P : i + 1;
      if (i >= S) nkeys = 0;
      const uint32_t rowmask = kmask & (nkeys >= 32 ? 0xffffffffu : ((1u << nkeys) - 1u));
      float mx = -INFINITY;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = ni * 8 + 2 * t + e;
          const float v = ((rowmask >> j) & 1u) ? sc[mi][ni][h * 2 + e] * p.scale_log2e : -INFINITY;
          sc[mi][ni][h * 2 + e] = v;
          mx = fmaxf(mx, v);
        }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float sum = 0.f;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = ni * 8 + 2 * t + e;
          const float ex = ((rowmask >> j) & 1u) ? exp2f(sc[mi][ni][h * 2 + e] - mx) : 0.f;
          sc[mi][ni][h * 2 + e] = ex;
          sum += ex;
        }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float inv = sum > 0.f ? 1.0f / sum : 0.f;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = ni * 8 + 2 * t + e;
          float pj = sc[mi][ni][h * 2 + e] * inv;
          if (DROP && ((rowmask >> j) & 1u)) pj *= drop_factor(p.drop, p.drop_site, ((static_cast<uint32_t>(a) * kHeads + head) * S + i) * S + j);
          sc[mi][ni][h * 2 + e] = pj;
        }
    }
  }
  float acc[2][8][4];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 8; ++ni)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mi][ni][e] = 0.f;
#pragma unroll
  for (int ki = 0; ki < 2; ++ki) {
    if (ki == 1 && small) continue;
    uint32_t ap[2][4];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      ap[mi][0] = pack_bf16x2(sc[mi][2 * ki][0], sc[mi][2 * ki][1]);
      ap[mi][1] = pack_bf16x2(sc[mi][2 * ki][2], sc[mi][2 * ki][3]);
      ap[mi][2] = pack_bf16x2(sc[mi][2 * ki + 1][0], sc[mi][2 * ki + 1][1]);
      ap[mi][3] = pack_bf16x2(sc[mi][2 * ki + 1][2], sc[mi][2 * ki + 1][3]);
    }
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4_t(b, sv + (ki * 16 + (lm & 1) * 8 + lr) * PT + np * 16 + (lm >> 1) * 8);
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        if (mi == 1 && small) continue;
        mma_bf16_16816(acc[mi][np * 2], ap[mi], b[0], b[1]);
        mma_bf16_16816(acc[mi][np * 2 + 1], ap[mi], b[2], b[3]);
      }
    }
  }
  __syncwarp();
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 8; ++ni)
#pragma unroll
      for (int h = 0; h < 2; ++h)
        if (mi == 0 || !small) *reinterpret_cast<uint32_t*>(sq + (mi * 16 + h * 8 + g) * PT + ni * 8 + 2 * t) = pack_bf16x2(acc[mi][ni][h * 2], acc[mi][ni][h * 2 + 1]);
  __syncwarp();
  __nv_bfloat16* out = p.out + static_cast<size_t>(a) * S * kE + head * kHeadDim;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + lm;
    if (r < S) *reinterpret_cast<uint4*>(out + static_cast<size_t>(r) * kE + lr * 8) = *reinterpret_cast<const uint4*>(sq + r * PT + lr * 8);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Input-embedding backward: dx0 (blocked fp32, rows a * S + s) ->
//   dpos[s] += sum_a dx0[a, s]                      (learned positions, embedding_decoder.py:1297)
//   dtok[target[a, i]] += dx0[a, P + i]             (tied token embedding, :692)
//   dprefix_t[(p, e), b] = sum_j dx0[(b * M + j), p][e]   bf16, the A operand of the prefix-projection wgrad
// Two kernels.  (One warp-per-row kernel with scalar atomics for everything took 600 us: 18 M atomics, a thousand per address.)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// dpos: row block b (32 rows, lane = row) holds position (32 b + lane) mod S in lane `lane`, so the row blocks b0, b0 + S, b0 + 2 S, ...
// give every lane the same position each time: blockIdx.y = b0 accumulates them in registers and adds once at the end.
// blockIdx.x * 4 + warp selects four of the 128 float4 column groups.
constexpr int kPosGradGroups = 4;
__global__ void __launch_bounds__(128) pos_grad_kernel(const float* __restrict__ dx0, int R, int S, float* __restrict__ dpos) {
  const int lane = lane_id();
  const int g0 = (blockIdx.x * 4 + (threadIdx.x >> 5)) * kPosGradGroups;
  const int b0 = blockIdx.y;
  const int nblk = (R + 31) >> 5;
  float4 acc[kPosGradGroups];
#pragma unroll
  for (int q = 0; q < kPosGradGroups; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = b0; b < nblk; b += S) {
    const int row = b * 32 + lane;
    if (row < R) {
#pragma unroll
      for (int q = 0; q < kPosGradGroups; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(dx0 + xblk_off(row, g0 + q));
        acc[q].x += t.x; acc[q].y += t.y; acc[q].z += t.z; acc[q].w += t.w;
      }
    }
  }
  const int s = (b0 * 32 + lane) % S;
#pragma unroll
  for (int q = 0; q < kPosGradGroups; ++q) red_add_v4(dpos + static_cast<size_t>(s) * kE + (g0 + q) * 4, acc[q].x, acc[q].y, acc[q].z, acc[q].w);
}

// dtok / dprefix_t: warp per row; lane owns the float4 column groups lane, lane + 32, lane + 64, lane + 96 (512-byte vector reductions).
// Rows whose gradient is exactly zero (every position after a sequence's EOS) are skipped.
__global__ void __launch_bounds__(128) embed_bwd_kernel(const float* __restrict__ dx0, const long long* __restrict__ target, int ld_target,
                                                        int nseq, int S, int P, int V, int Mrep, int B,
                                                        float* __restrict__ dtok, __nv_bfloat16* __restrict__ dprefix_t, int ld_pt) {
  const int w = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (w >= nseq * S) return;
  const int lane = lane_id();
  const int a = w / S, s = w - a * S;
  if (s < P && Mrep != 1) return;
  float4 v[4];
  bool nz = false;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    v[q] = *reinterpret_cast<const float4*>(dx0 + xblk_off(w, q * 32 + lane));
    nz = nz || v[q].x != 0.f || v[q].y != 0.f || v[q].z != 0.f || v[q].w != 0.f;
  }
  if (s >= P) {
    if (!__any_sync(0xffffffffu, nz)) return;
    long long tok = target[static_cast<size_t>(a) * ld_target + (s - P)];
    tok = tok < 0 ? 0 : (tok >= V ? V - 1 : tok);
#pragma unroll
    for (int q = 0; q < 4; ++q) red_add_v4(dtok + static_cast<size_t>(tok) * kE + (q * 32 + lane) * 4, v[q].x, v[q].y, v[q].z, v[q].w);
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float e[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) dprefix_t[static_cast<size_t>(s * kE + (q * 32 + lane) * 4 + i) * ld_pt + a] = __float2bfloat16_rn(e[i]);
    }
  }
}

// multi-target case: sum the M sequences that share an embedding before the transposed bf16 store
__global__ void __launch_bounds__(128) prefix_grad_reduce_kernel(const float* __restrict__ dx0, int B, int Mrep, int S, int P,
                                                                 __nv_bfloat16* __restrict__ dprefix_t, int ld_pt) {
  const int w = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (w >= B * P) return;
  const int lane = lane_id();
  const int b = w / P, pp = w - b * P;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int j = 0; j < Mrep; ++j) {
    const int row = (b * Mrep + j) * S + pp;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 t = *reinterpret_cast<const float4*>(dx0 + xblk_off(row, lane * 4 + q));
      acc[q * 4] += t.x; acc[q * 4 + 1] += t.y; acc[q * 4 + 2] += t.z; acc[q * 4 + 3] += t.w;
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) dprefix_t[static_cast<size_t>(pp * kE + lane * 16 + i) * ld_pt + b] = __float2bfloat16_rn(acc[i]);
}

// log-sum-exp per logits row from the forward partials (needed by the dlogits epilogue)
__global__ void __launch_bounds__(128) row_lse_kernel(const LogitPartial* __restrict__ part, int nparts, int nrows, float* __restrict__ lse) {
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= nrows) return;
  const RowStats s = merge_partials(part + static_cast<size_t>(r) * nparts, nparts, 1.0f);
  if (lane_id() == 0) lse[r] = s.lse_one;
}

// ---------------------------------------------------------------------------------------------------------
// GEMM epilogues of the backward pass
// ---------------------------------------------------------------------------------------------------------

// plain bf16 row-major store: C[M, ldc]
struct EpiStoreBF16 {
  struct Params {
    __nv_bfloat16* c;
    int ldc;
  };
  template <class Release>
  __device__ static __forceinline__ void run(const Params& p, const EpiCtx& c, Release release) {
    const int lane = lane_id();
#pragma unroll
    for (int ch = 0; ch < kEpiCols / 32; ++ch) {
      float v[32];
      tmem_ld_32x32(c.tmem_row + ch * 32, v);
      if (ch == kEpiCols / 32 - 1) release();
      stage_put32(c.stage, lane, ch * 32, v);
    }
    stage_copy_out(c.stage, lane, [&](int r) -> __nv_bfloat16* {
      const int row = c.warp_row0 + r;
      return row < c.M ? p.c + static_cast<size_t>(row) * p.ldc + c.n0 : nullptr;
    });
  }
};

// FFN1 of the training forward: keeps the pre-activation (for GELU backward) and the activation
struct EpiGeluTrain {
  struct Params {
    __nv_bfloat16* pre;
    __nv_bfloat16* h;
    int ldh;
    DropCfg drop;        // thresh != 0: dropout on the activated rows h (element index row * ldh + column); pre stays unmasked
    uint32_t drop_site;
  };
  template <class Release>
  __device__ static __forceinline__ void run(const Params& p, const EpiCtx& c, Release release) {
    const int lane = lane_id();
    float v[kEpiCols];
#pragma unroll
    for (int ch = 0; ch < kEpiCols / 32; ++ch) {
      float t[32];
      tmem_ld_32x32(c.tmem_row + ch * 32, t);
      if (ch == kEpiCols / 32 - 1) release();
#pragma unroll
      for (int j = 0; j < 32; ++j) v[ch * 32 + j] = t[j];
      stage_put32(c.stage, lane, ch * 32, t);
    }
    stage_copy_out(c.stage, lane, [&](int r) -> __nv_bfloat16* {
      const int row = c.warp_row0 + r;
      return row < c.M ? p.pre + static_cast<size_t>(row) * p.ldh + c.n0 : nullptr;
    });
#pragma unroll
    for (int ch = 0; ch < kEpiCols / 32; ++ch) {
      float t[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) t[j] = gelu_fast(__bfloat162float(__float2bfloat16_rn(v[ch * 32 + j])));  // from the stored pre-activation
      if (p.drop.thresh != 0u) {
#pragma unroll
        for (int j = 0; j < 32; ++j) t[j] *= drop_factor(p.drop, p.drop_site, static_cast<uint32_t>(c.row) * p.ldh + c.n0 + ch * 32 + j);
      }
      stage_put32(c.stage, lane, ch * 32, t);
    }
    stage_copy_out(c.stage, lane, [&](int r) -> __nv_bfloat16* {
      const int row = c.warp_row0 + r;
      return row < c.M ? p.h + static_cast<size_t>(row) * p.ldh + c.n0 : nullptr;
    });
  }
};

// fp32 store into a blocked [rows, 512] gradient tensor (N = 512 GEMMs), optional output-row remap (logits rows -> sequence rows)
struct EpiGradBlocked {
  struct Params {
    float* g;
    int remap_rows_in, remap_rows_out, remap_offset;  // out_row = (row / in) * out + offset + row % in   (in = 0: identity)
  };
  template <class Release>
  __device__ static __forceinline__ void run(const Params& p, const EpiCtx& c, Release release) {
    int orow = c.row;
    if (p.remap_rows_in > 0) {
      const int seq = c.row / p.remap_rows_in;
      orow = seq * p.remap_rows_out + p.remap_offset + (c.row - seq * p.remap_rows_in);
    }
#pragma unroll
    for (int ch = 0; ch < kEpiCols / 32; ++ch) {
      float v[32];
      tmem_ld_32x32(c.tmem_row + ch * 32, v);
      if (ch == kEpiCols / 32 - 1) release();
      if (c.row < c.M) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(p.g + xblk_off(orow, ((c.n0 + ch * 32) >> 2) + q)) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
      }
    }
  }
};

// weight gradient: fp32 atomic accumulation into dW[M = out features, N = in features] (split-K over the row dimension)
struct EpiAtomicF32 {
  struct Params {
    float* dw;
    int ldw;   // in features
    int ncols; // valid columns (in features)
  };
  // The accumulator rows go through the warp's staging tile so that one instruction adds 4 rows x 128 contiguous bytes with 16-byte
  // vector reductions.  (Thread = row with scalar atomics touched 32 sectors per instruction: ~50 us even for a 128 x 512 gradient.)
  template <class Release>
  __device__ static __forceinline__ void run(const Params& p, const EpiCtx& c, Release release) {
    const int lane = lane_id();
    const int sub = lane >> 3, chunk = lane & 7;
#pragma unroll
    for (int ch = 0; ch < kEpiCols / 32; ++ch) {
      float v[32];
      tmem_ld_32x32(c.tmem_row + ch * 32, v);
      if (ch == kEpiCols / 32 - 1) release();
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *stage_chunk(c.stage, lane, q) = make_uint4(__float_as_uint(v[q * 4]), __float_as_uint(v[q * 4 + 1]), __float_as_uint(v[q * 4 + 2]), __float_as_uint(v[q * 4 + 3]));
      __syncwarp();
      const int col = c.n0 + ch * 32 + chunk * 4;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + sub, row = c.warp_row0 + r;
        const uint4 u = *stage_chunk(c.stage, r, chunk);
        if (row >= c.M || col >= p.ncols) continue;
        float* d = p.dw + static_cast<size_t>(row) * p.ldw + col;
        if (col + 4 <= p.ncols && (reinterpret_cast<uintptr_t>(d) & 15) == 0) {
          red_add_v4(d, __uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
        } else {
          const float e[4] = {__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w)};
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (col + j < p.ncols) atomicAdd(d + j, e[j]);
        }
      }
      __syncwarp();
    }
  }
};

// dlogits = (softmax(logits) - smoothed one-hot(target)) * row_weight, bf16 row-major [rows, V]
struct EpiDLogits {
  struct Params {
    __nv_bfloat16* dl;          // [M, ld]
    long long ld;
    const float* lse;           // [M]
    const long long* target;    // [M], -1 = ignored row (zero gradient)
    const float* row_weight;    // optional [M / rows_per_seq]
    int rows_per_seq;
    int n_valid;                // V
    float label_smoothing;
  };
  template <class Release>
  __device__ static __forceinline__ void run(const Params& p, const EpiCtx& c, Release release) {
    const int lane = lane_id();
    const bool in_range = c.row < c.M;
    const long long tgt = in_range ? p.target[c.row] : -1;
    const float lse = in_range ? p.lse[c.row] : 0.f;
    float w = tgt >= 0 ? 1.0f : 0.0f;
    if (p.row_weight != nullptr && in_range) w *= p.row_weight[c.row / p.rows_per_seq];
    const float on = 1.0f - p.label_smoothing, off = p.label_smoothing / static_cast<float>(p.n_valid);
#pragma unroll
    for (int ch = 0; ch < kEpiCols / 32; ++ch) {
      const int col0 = c.n0 + ch * 32;
      float v[32];
      tmem_ld_32x32(c.tmem_row + ch * 32, v);
      if (ch == kEpiCols / 32 - 1) release();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = col0 + j;
        float g = __expf(v[j] - lse) - off - (static_cast<long long>(col) == tgt ? on : 0.f);
        v[j] = col < p.n_valid ? g * w : 0.f;
      }
      stage_put32(c.stage, lane, ch * 32, v);
    }
    const long long ld = p.ld;
    const int nv = p.n_valid;
    const int n0 = c.n0;
    // rows of dl are padded to a multiple of 64 columns by the caller (ld >= round_up(V, 64)), so the 64-column block is in range
    stage_copy_out(c.stage, lane, [&](int r) -> __nv_bfloat16* {
      const int row = c.warp_row0 + r;
      return (row < c.M && n0 < nv) ? p.dl + static_cast<size_t>(row) * ld + n0 : nullptr;
    });
  }
};

}  // namespace novic
