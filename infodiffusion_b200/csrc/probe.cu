// Hardware probe (debug entry point, not on the product path): does a K-major SW128 UMMA descriptor
// whose start address is shifted by an arbitrary number of 128-byte rows read the rows TMA wrote?
// Decides how the halo-reuse convolution addresses its 3x3 taps (base_offset semantics).
#include "kernels.cuh"

namespace idf {

struct alignas(64) ProbeParams {
  CUtensorMap tmA;   // [256, 64] bf16, box {64, 128}
  CUtensorMap tmB;   // [64, 64]  bf16, box {64, 64}
  float* out;        // [128, 64] fp32
  int shift;         // rows
  int mode;          // 0: base_offset field = 0, 1: base_offset = (start_addr >> 7) & 7
};

__global__ void __launch_bounds__(128, 1) shift_probe_kernel(const __grid_constant__ ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;                 // 256 rows x 128 B
  uint8_t* smB = smem + 256 * 128;     // 64 rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + 64 * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bars + 0, 1);
    mbar_init(bars + 1, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bars + 0, 256 * 128 + 64 * 128);
    tma_load_2d(smA, &p.tmA, bars + 0, 0, 0);
    tma_load_2d(smA + 128 * 128, &p.tmA, bars + 0, 0, 128);
    tma_load_2d(smB, &p.tmB, bars + 0, 0, 0);
    mbar_wait(bars + 0, 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(smA) + static_cast<uint32_t>(p.shift) * 128u;
    uint64_t da = umma_desc_k_sw128(a_addr);
    if (p.mode == 1) da |= static_cast<uint64_t>((a_addr >> 7) & 7u) << 49;
    const uint64_t db = umma_desc_k_sw128(smem_u32(smB));
    constexpr uint32_t idesc = umma_idesc_f16(128, 64, kFmtBF16);
    for (int k = 0; k < 4; ++k) umma_f16(tmem, da + 2 * k, db + 2 * k, idesc, k != 0 ? 1u : 0u);
    umma_commit(bars + 1);
  }
  mbar_wait(bars + 1, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) p.out[row * 64 + c * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 64);
  }
}

cudaError_t launch_shift_probe(const CUtensorMap& tmA, const CUtensorMap& tmB, float* out, int shift, int mode,
                               cudaStream_t stream) {
  ProbeParams p;
  p.tmA = tmA; p.tmB = tmB; p.out = out; p.shift = shift; p.mode = mode;
  const int smem = 256 * 128 + 64 * 128 + 64 + 1024;
  cudaError_t e = cudaFuncSetAttribute(shift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  shift_probe_kernel<<<1, 128, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace idf
