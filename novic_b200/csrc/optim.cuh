// Optimizer tail of the training step (train.py:1281-1286, :1108-1119): global gradient norm -> clip coefficient -> AdamW, over the ONE
// flat fp32 gradient bucket the backward pass writes (novic_b200/training.py) and one flat parameter buffer (novic_b200/optim.py).
// Three launches, no host synchronisation: the 1 / loss_basis normalisation (read from the all-reduced statistics on the device), the
// clip coefficient and the non-finite check all stay on the GPU.  Pure HBM work: 4 streams of 50.9 MB read (p, g, m, v) + 3 written.
#pragma once

#include "ptx.cuh"

namespace novic {

constexpr int kOptBlocks = 592;        // 4 x 148 SMs: every SM holds four 256-thread CTAs of these register-light kernels
constexpr int kOptThreads = 256;
constexpr int kOptChunk = 512;         // granularity of the weight-decay flags: every parameter tensor of the decoder is a multiple of 512 elements

struct OptHyper {
  float lr, beta1, beta2, eps, weight_decay, max_norm;
  float bias_correction1;        // 1 - beta1^step
  float bias_correction2_sqrt;   // sqrt(1 - beta2^step)
};

// partial[b] = sum of g[i]^2 over block b's grid-stride share (fp32 per thread over ~170 elements, fp64 across threads; fixed order)
__global__ void __launch_bounds__(kOptThreads) grad_sqnorm_kernel(const float* __restrict__ g, long long n, double* __restrict__ partial) {
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(g4 + i);
    acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float v = g[(n4 << 2) + threadIdx.x]; acc = fmaf(v, v, acc); }
  double d = static_cast<double>(acc);
  __shared__ double s[kOptThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kOptThreads / 32; ++w) t += s[w];
    partial[blockIdx.x] = t;
  }
}

// out[0] = gradient norm after the 1 / basis normalisation, out[1] = clip coefficient (torch.nn.utils.clip_grad_norm_: max_norm / (norm + 1e-6),
// at most 1), out[2] = factor applied to the raw gradients (clip coefficient / basis), out[3] = 1 when the norm is not finite (the update is
// then skipped: clip_grad_norm_(error_if_nonfinite=True) would raise; the host reads the flag whenever it chooses to).
// stats: optional device pointer to [loss_sum, loss_basis, ...] (after the all-reduce): the gradients are those of loss_sum, the step uses
// d(loss_sum / loss_basis) (train.py:1272).  nullptr: the gradients are taken as they are.
__global__ void clip_coef_kernel(const double* __restrict__ partial, int nblocks, const float* __restrict__ stats, float max_norm, float* __restrict__ out) {
  __shared__ double s[32];
  double t = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 32) t += partial[i];      // one warp, fixed assignment -> deterministic
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (threadIdx.x == 0) {
    s[0] = t;
    const float inv_basis = stats != nullptr ? 1.0f / fmaxf(stats[1], 1.0f) : 1.0f;
    const float norm = static_cast<float>(sqrt(t)) * inv_basis;
    const bool finite = isfinite(norm);
    float coef = 1.0f;
    if (max_norm > 0.f) coef = fminf(1.0f, max_norm / (norm + 1e-6f));
    out[0] = norm;
    out[1] = coef;
    out[2] = finite ? coef * inv_basis : 0.f;
    out[3] = finite ? 0.f : 1.f;
  }
}

// AdamW in the operation order of torch's fused kernel (decoupled weight decay first, bias corrections from the step count):
//   p -= lr * wd * p;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
// decay[c] != 0: chunk c (kOptChunk elements) belongs to a tensor that receives weight decay (>= 2-D parameters, train.py:1108-1115).
__global__ void __launch_bounds__(kOptThreads) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                           long long n, const unsigned char* __restrict__ decay, OptHyper h, const float* __restrict__ coef) {
  const float scale = coef != nullptr ? __ldg(coef + 2) : 1.0f;
  if (coef != nullptr && __ldg(coef + 3) != 0.f) return;          // non-finite gradient norm: leave the parameters alone
  const long long n4 = n >> 2;
  const float step_size = h.lr / h.bias_correction1;
  const float w1 = 1.0f - h.beta1, w2 = 1.0f - h.beta2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float wd = decay[(i << 2) / kOptChunk] != 0 ? h.weight_decay : 0.f;
    float4 P = reinterpret_cast<float4*>(p)[i];
    const float4 G = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 M = reinterpret_cast<float4*>(m)[i];
    float4 V = reinterpret_cast<float4*>(v)[i];
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
      gg *= scale;
      pp -= h.lr * wd * pp;
      mm = h.beta1 * mm + w1 * gg;
      vv = h.beta2 * vv + w2 * gg * gg;
      const float denom = sqrtf(vv) / h.bias_correction2_sqrt + h.eps;
      pp -= step_size * mm / denom;
    };
    upd(P.x, G.x, M.x, V.x); upd(P.y, G.y, M.y, V.y); upd(P.z, G.z, M.z, V.z); upd(P.w, G.w, M.w, V.w);
    reinterpret_cast<float4*>(p)[i] = P;
    reinterpret_cast<float4*>(m)[i] = M;
    reinterpret_cast<float4*>(v)[i] = V;
  }
}

}  // namespace novic
