// Stand-alone throughput probe of the fused-AdaGN transform loop (development tool, not part of libidf_b200.so):
// one CTA per SM rewrites a [rows x 128 B] shared-memory tile in place as bf16(silu(A*x+B)), W warps, the thread ->
// (granule column, row) mapping of conv_halo_kernel's transform warps.  Prints ns per tile for several W and variants.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xf_bench xf_bench.cu && ./xf_bench
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) { return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }

template <int MODE>   // 0 full, 1 no tanh, 2 copy only, 3 full but two granules interleaved by hand
__device__ __forceinline__ uint4 apply(const uint4& u, const float (&A)[8], const float (&B)[8]) {
  if (MODE == 2) return u;
  const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
  float f[8] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float h = fmaf(f[j], A[j], B[j]);
    if (MODE == 1) { f[j] = h; continue; }
    float th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
    f[j] = fmaf(h, th, h);
  }
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]); o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
  return o;
}

template <int MODE>
__global__ void xf_kernel(int rows, int iters, const float* coef, long long* clk_out) {
  extern __shared__ uint8_t smem[];
  const uint32_t base0 = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const int tt = threadIdx.x, gi = tt & 7, rs = tt >> 3, R = blockDim.x >> 3;
  for (int i = tt; i < rows * 8; i += blockDim.x) sts128(base0 + i * 16, make_uint4(0x3f803f80u, 0x3f003f00u, 0x40004000u, 0xbf80bf80u));
  float A[8], B[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { A[j] = coef[gi * 8 + j]; B[j] = coef[64 + gi * 8 + j]; }
  __syncthreads();
  const long long t0 = clock64();
  const uint32_t base = base0 + gi * 16;
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int i0 = rs; i0 < rows; i0 += 4 * R) {
      uint4 u[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { const int i = i0 + R * q; u[q] = lds128(base + (i < rows ? i : i0) * 128); }
#pragma unroll
      for (int q = 0; q < 4; ++q) { const int i = i0 + R * q; if (i < rows) sts128(base + i * 128, apply<MODE>(u[q], A, B)); }
    }
    __syncthreads();
  }
  if (tt == 0) clk_out[blockIdx.x] = clock64() - t0;
}


__device__ __forceinline__ int lds_s16(uint32_t a) { int v; asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts16(uint32_t a, uint16_t v) { asm volatile("st.shared.b16 [%0], %1;" ::"r"(a), "h"(v) : "memory"); }

// the transform branch of conv_halo_kernel, minus the TMA / mbarrier hand-shakes: FLAGS bit0 row table, bit1 reload from global
template <int FLAGS>
__global__ void xf_real(int rows, int iters, const float2* ctab, int ctot, long long* clk_out, int Wp, int Hp, int W, int H) {
  extern __shared__ uint8_t smem[];
  const uint32_t base0 = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const uint32_t tab = base0 + rows * 128;
  const int tt = threadIdx.x, gi = tt & 7, rs = tt >> 3, R = blockDim.x >> 3, NXT = blockDim.x;
  const int gl = gi ^ (rs & 7);
  for (int i = tt; i < rows * 8; i += blockDim.x) sts128(base0 + i * 16, make_uint4(0x3f803f80u, 0x3f003f00u, 0x40004000u, 0xbf80bf80u));
  for (int i = tt; i < rows; i += blockDim.x) sts16(tab + 2 * i, 0);
  float A[8], B[8];
  int cur = -1;
  const float inv_wp = 1.0f / Wp, inv_R = 1.0f / (Hp * Wp);
  const int Rr = Hp * Wp;
  __syncthreads();
  const long long t0 = clock64();
  const uint32_t base = base0 + gi * 16;
  const float2* ct = ctab + gl * 8;
  for (int it = 0; it < iters; ++it) {
    const int rbase = (blockIdx.x + it * 148) * 512 - 66;
    const int rfirst = rbase < 0 ? 0 : rbase;
    const int img0 = __float2int_rd((rfirst + 0.5f) * inv_R);
    if (FLAGS & 1) {
      for (int i = tt; i < rows; i += NXT) {
        const int r = rbase + i;
        int inf = -1;
        if (r >= 0) {
          const int img = __float2int_rd((r + 0.5f) * inv_R);
          const int rr = r - img * Rr;
          const int y = __float2int_rd((rr + 0.5f) * inv_wp);
          const int x = rr - y * Wp;
          if (x < W && y < H) inf = img - img0;
        }
        sts16(tab + 2 * i, static_cast<uint16_t>(inf));
      }
      __syncthreads();
    }
    auto reload = [&](int img) {
      cur = img;
      if (FLAGS & 2) {
        const float4* c4 = reinterpret_cast<const float4*>(ct + static_cast<long long>(img & 255) * ctot);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 c = __ldg(c4 + j);
          A[2 * j] = c.x * 0.5f; B[2 * j] = c.y * 0.5f; A[2 * j + 1] = c.z * 0.5f; B[2 * j + 1] = c.w * 0.5f;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { A[j] = 0.3f + 0.01f * j; B[j] = 0.1f; }
      }
    };
#pragma unroll 1
    for (int i0 = rs; i0 < rows; i0 += 4 * R) {
      int info[4];
      uint4 u[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = i0 + R * q;
        const int ic = i < rows ? i : i0;
        info[q] = i < rows ? lds_s16(tab + 2u * ic) : -1;
        u[q] = lds128(base + ic * 128u);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (info[q] < 0) continue;
        if (info[q] + img0 != cur) reload(info[q] + img0);
        sts128(base + (i0 + R * q) * 128u, apply<0>(u[q], A, B));
      }
    }
    __syncthreads();
  }
  if (tt == 0) clk_out[blockIdx.x] = clock64() - t0;
}

template <int FLAGS>
void run_real(const char* name, int warps, int rows, int iters, const float2* ctab, long long* clk) {
  const int smem = rows * 128 + rows * 2 + 64;
  cudaFuncSetAttribute(xf_real<FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  xf_real<FLAGS><<<148, warps * 32, smem>>>(rows, 2, ctab, 64, clk, 65, 65, 64, 64);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  xf_real<FLAGS><<<148, warps * 32, smem>>>(rows, iters, ctab, 64, clk, 65, 65, 64, 64);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("%-28s warps %2d: %7.2f us per %d-row halo  %s\n", name, warps, ms * 1e3 / iters, rows, cudaGetErrorString(cudaGetLastError()));
}


// NG granules per step in three phases (all affines, all tanh, all second FMAs + packs): more independent chains in flight
template <int NG>
__global__ void xf_phased(int rows, int iters, const float* coef, long long* clk_out) {
  extern __shared__ uint8_t smem[];
  const uint32_t base0 = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const int tt = threadIdx.x, gi = tt & 7, rs = tt >> 3, R = blockDim.x >> 3;
  for (int i = tt; i < rows * 8; i += blockDim.x) sts128(base0 + i * 16, make_uint4(0x3f803f80u, 0x3f003f00u, 0x40004000u, 0xbf80bf80u));
  float A[8], B[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { A[j] = coef[gi * 8 + j]; B[j] = coef[64 + gi * 8 + j]; }
  __syncthreads();
  const long long t0 = clock64();
  const uint32_t base = base0 + gi * 16;
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int i0 = rs; i0 < rows; i0 += NG * R) {
      uint4 u[NG];
#pragma unroll
      for (int q = 0; q < NG; ++q) { const int i = i0 + R * q; u[q] = lds128(base + (i < rows ? i : i0) * 128); }
      float h[NG][8], th[NG][8];
#pragma unroll
      for (int q = 0; q < NG; ++q) {
        const float2 a0 = unpack_bf16x2(u[q].x), a1 = unpack_bf16x2(u[q].y), a2 = unpack_bf16x2(u[q].z), a3 = unpack_bf16x2(u[q].w);
        const float f[8] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y};
#pragma unroll
        for (int j = 0; j < 8; ++j) h[q][j] = fmaf(f[j], A[j], B[j]);
      }
#pragma unroll
      for (int q = 0; q < NG; ++q)
#pragma unroll
        for (int j = 0; j < 8; ++j) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(th[q][j]) : "f"(h[q][j]));
#pragma unroll
      for (int q = 0; q < NG; ++q) {
        uint4 o;
        o.x = pack_bf16x2(fmaf(h[q][0], th[q][0], h[q][0]), fmaf(h[q][1], th[q][1], h[q][1]));
        o.y = pack_bf16x2(fmaf(h[q][2], th[q][2], h[q][2]), fmaf(h[q][3], th[q][3], h[q][3]));
        o.z = pack_bf16x2(fmaf(h[q][4], th[q][4], h[q][4]), fmaf(h[q][5], th[q][5], h[q][5]));
        o.w = pack_bf16x2(fmaf(h[q][6], th[q][6], h[q][6]), fmaf(h[q][7], th[q][7], h[q][7]));
        const int i = i0 + R * q;
        if (i < rows) sts128(base + i * 128, o);
      }
    }
    __syncthreads();
  }
  if (tt == 0) clk_out[blockIdx.x] = clock64() - t0;
}
template <int NG>
void run_phased(int warps, int rows, int iters, const float* coef, long long* clk) {
  const int smem = rows * 128;
  cudaFuncSetAttribute(xf_phased<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  xf_phased<NG><<<148, warps * 32, smem>>>(rows, 2, coef, clk);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  xf_phased<NG><<<148, warps * 32, smem>>>(rows, iters, coef, clk);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("phased NG=%d                  warps %2d: %7.2f us per %d-row halo  %s\n", NG, warps, ms * 1e3 / iters, rows, cudaGetErrorString(cudaGetLastError()));
}

template <int MODE>
void run(const char* name, int warps, int rows, int iters, const float* coef, long long* clk) {
  const int smem = rows * 128;
  cudaFuncSetAttribute(xf_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  xf_kernel<MODE><<<148, warps * 32, smem>>>(rows, 2, coef, clk);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  xf_kernel<MODE><<<148, warps * 32, smem>>>(rows, iters, coef, clk);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long c;
  cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
  printf("%-10s warps %2d: %7.2f us per %d-row tile (%.0f clk, %.2f clk/granule/SM)  %s\n", name, warps, ms * 1e3 / iters, rows,
         double(c) / iters, double(c) / iters / (rows * 8), cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float h[128];
  for (int i = 0; i < 128; ++i) h[i] = 0.3f + 0.01f * i;
  float* coef; long long* clk;
  cudaMalloc(&coef, sizeof(h)); cudaMalloc(&clk, 148 * 8);
  cudaMemcpy(coef, h, sizeof(h), cudaMemcpyHostToDevice);
  const int rows = 648, iters = 200;
  for (int w : {4, 8, 12, 16, 24, 32}) run<0>("full", w, rows, iters, coef, clk);
  for (int w : {8, 16}) run<1>("no tanh", w, rows, iters, coef, clk);
  for (int w : {8, 16}) run<2>("copy", w, rows, iters, coef, clk);
  for (int w : {8, 12, 16}) { run_phased<1>(w, rows, iters, coef, clk); run_phased<2>(w, rows, iters, coef, clk); run_phased<4>(w, rows, iters, coef, clk); }
  float2* ctab;
  cudaMalloc(&ctab, 256 * 64 * sizeof(float2));
  cudaMemset(ctab, 0, 256 * 64 * sizeof(float2));
  for (int w : {8, 12, 16}) {
    run_real<0>("real loop, const coef", w, rows, iters, ctab, clk);
    run_real<1>("+ row table + barrier", w, rows, iters, ctab, clk);
    run_real<2>("+ reload from global", w, rows, iters, ctab, clk);
    run_real<3>("+ both", w, rows, iters, ctab, clk);
  }
  return 0;
}
