// Train-step tail (run.py:199-200): torch.nn.utils.clip_grad_norm_(params, max_norm) followed by AdamW.step()
// as three launches over ALL parameter tensors instead of ~750 small ones.
//   1. per 4096-element chunk: sum of squares of the gradient            -> partial[chunk]
//   2. one CTA: total = sqrt(sum partial) (fixed order), coef = min(1, max_norm / (total + 1e-6))
//   3. per chunk: g *= coef;  p *= 1 - lr*wd;  m, v moments;  p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
// Tensors are addressed through device-side pointer tables (one entry per parameter tensor); a chunk table
// maps CTA -> (tensor, offset).  All tensors fp32.  HBM-bound: 4 reads + 3 writes of 4 bytes per element.
#include "kernels.cuh"

namespace idf {

constexpr int kOptChunk = 4096;
constexpr int kOptThreads = 256;

__global__ void __launch_bounds__(kOptThreads) grad_sqsum_kernel(const ClipAdamWParams p) {
  __shared__ float red[kOptThreads / 32];
  const int c = blockIdx.x;
  const int ti = p.chunk_tensor[c];
  const long long off = static_cast<long long>(p.chunk_offset[c]) * kOptChunk;
  const long long n = p.numel[ti];
  const float* g = static_cast<const float*>(p.grads[ti]) + off;
  const int cnt = static_cast<int>(min(static_cast<long long>(kOptChunk), n - off));
  float s = 0.f;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const int n4 = cnt >> 2;
    for (int i = threadIdx.x; i < n4; i += kOptThreads) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
      s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
    for (int i = (n4 << 2) + threadIdx.x; i < cnt; i += kOptThreads) s = fmaf(g[i], g[i], s);
  } else {
    for (int i = threadIdx.x; i < cnt; i += kOptThreads) s = fmaf(g[i], g[i], s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < kOptThreads / 32; ++w) t += red[w];
    p.partial[c] = t;
  }
}

__global__ void __launch_bounds__(1024) grad_norm_finish_kernel(const float* __restrict__ partial, int n, float max_norm,
                                                                float* __restrict__ norm_out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) s += static_cast<double>(partial[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 32; ++w) t += red[w];
    const float total = static_cast<float>(sqrt(t));
    norm_out[0] = total;
    norm_out[1] = max_norm > 0.f ? fminf(1.0f, max_norm / (total + 1e-6f)) : 1.0f;     // clip_grad_norm_'s clamp
  }
}

__global__ void __launch_bounds__(kOptThreads) adamw_kernel(const ClipAdamWParams p) {
  const int c = blockIdx.x;
  const int ti = p.chunk_tensor[c];
  const long long off = static_cast<long long>(p.chunk_offset[c]) * kOptChunk;
  const long long n = p.numel[ti];
  const int cnt = static_cast<int>(min(static_cast<long long>(kOptChunk), n - off));
  const float* g = static_cast<const float*>(p.grads[ti]) + off;
  float* w = static_cast<float*>(p.params[ti]) + off;
  float* m = static_cast<float*>(p.exp_avg[ti]) + off;
  float* v = static_cast<float*>(p.exp_avg_sq[ti]) + off;
  const float coef = p.norm_out[1];
  const float decay = 1.0f - p.lr * p.weight_decay;
  const float step_size = p.lr / p.bias_corrections[2 * ti];          // per tensor: torch keeps one step count per parameter
  const float inv_sqrt_bc2 = rsqrtf(p.bias_corrections[2 * ti + 1]);
  auto upd = [&](float gi, float& mi, float& vi, float& wi) {
    gi *= coef;
    mi = fmaf(p.beta1, mi, (1.0f - p.beta1) * gi);
    vi = fmaf(p.beta2, vi, (1.0f - p.beta2) * gi * gi);
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + p.eps;
    wi = wi * decay - step_size * (mi / denom);
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  int done = 0;
  if (vec) {
    const int n4 = cnt >> 2;
    for (int i = threadIdx.x; i < n4; i += kOptThreads) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(g) + i);
      float4 m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i], w4 = reinterpret_cast<float4*>(w)[i];
      upd(g4.x, m4.x, v4.x, w4.x); upd(g4.y, m4.y, v4.y, w4.y); upd(g4.z, m4.z, v4.z, w4.z); upd(g4.w, m4.w, v4.w, w4.w);
      reinterpret_cast<float4*>(m)[i] = m4; reinterpret_cast<float4*>(v)[i] = v4; reinterpret_cast<float4*>(w)[i] = w4;
    }
    done = n4 << 2;
  }
  for (int i = done + threadIdx.x; i < cnt; i += kOptThreads) upd(g[i], m[i], v[i], w[i]);
}

cudaError_t launch_clip_adamw(const ClipAdamWParams& p, cudaStream_t stream) {
  if (p.n_chunks <= 0) return cudaSuccess;
  grad_sqsum_kernel<<<p.n_chunks, kOptThreads, 0, stream>>>(p);
  grad_norm_finish_kernel<<<1, 1024, 0, stream>>>(p.partial, p.n_chunks, p.max_norm, p.norm_out);
  adamw_kernel<<<p.n_chunks, kOptThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace idf
