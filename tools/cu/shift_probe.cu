// Hardware probe (STANDALONE program, not part of libidf_b200.so): does a K-major SW128 UMMA descriptor
// whose start address is shifted by an arbitrary number of 128-byte rows read the rows TMA wrote?
// Decides how the halo-reuse convolution addresses its 3x3 taps (base_offset semantics).
#include <cudaTypedefs.h>
#include <cstdio>
#include <vector>

#include "../../infodiffusion_b200/csrc/ptx.cuh"

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

// ---- host driver: A = 256 x 64 ramp, B = identity => D = A rows [shift, shift + 128)
static int encode(PFN_cuTensorMapEncodeTiled_v12000 enc, CUtensorMap* tm, const void* base, int rows, int box_rows) {
  cuuint64_t gdim[2] = {64, static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {128};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t es[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return 2;
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  std::vector<__nv_bfloat16> ha(256 * 64), hb(64 * 64);
  for (int i = 0; i < 256 * 64; ++i) ha[i] = __float2bfloat16(((i % 509) - 254) / 4.0f);
  for (int i = 0; i < 64 * 64; ++i) hb[i] = __float2bfloat16((i / 64 == i % 64) ? 1.0f : 0.0f);
  __nv_bfloat16 *da, *db;
  float* dout;
  cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  if (encode(enc, &tmA, da, 256, 128) || encode(enc, &tmB, db, 64, 64)) return 3;
  const int shifts[] = {0, 1, 2, 7, 8, 9, 16, 65, 66, 73, 127, 128};
  std::vector<float> out(128 * 64);
  for (int shift : shifts)
    for (int mode = 0; mode < 2; ++mode) {
      cudaMemset(dout, 0xff, 128 * 64 * 4);
      if (idf::launch_shift_probe(tmA, tmB, dout, shift, mode, 0) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
        printf("shift %3d mode %d: CUDA error %s\n", shift, mode, cudaGetErrorString(cudaGetLastError()));
        return 4;
      }
      cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int r = 0; r < 128; ++r)
        for (int c = 0; c < 64; ++c) bad += out[r * 64 + c] != __bfloat162float(ha[(r + shift) * 64 + c]);
      printf("shift %3d mode %d: mismatches %d\n", shift, mode, bad);
    }
  return 0;
}
