// Implicit-GEMM convolution for sm_100a: TMA-fed, tcgen05.mma with TMEM accumulators,
// warp-specialised persistent kernel.  See include/idf_b200.h for the contract.
//
// Roles (256 threads):  warp 0 = TMA producer (1 thread)   warp 1 = UMMA issuer (1 thread)
//                       warp 2 = TMEM allocator            warps 4..7 = epilogue (TMEM -> regs -> HBM)
// Pipelines: smem ring full/empty (TMA <-> UMMA), TMEM accumulator full/empty x2 (UMMA <-> epilogue).
//
// One tile = 128 pad-flat output rows x BN output channels.  For k-block kb the A operand is the
// [128 x 64] slice  src[kb_src][row0 + kb_rowoff : +128, kb_c0 : +64]  (TMA zero-fills rows outside
// the tensor) and the B operand is Wp[n0 : n0+BN, 64*kb : 64*kb+64]; both land in shared memory in
// the K-major 128B-swizzled layout that the UMMA descriptors of ptx.cuh describe.
#include "kernels.cuh"

namespace idf {

template <int BN>
struct ConvCfg {
  static constexpr int STAGES = (BN == 128) ? 6 : 8;
  static constexpr uint32_t A_BYTES = kBM * kBK * 2;
  static constexpr uint32_t B_BYTES = BN * kBK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 256 /*barriers*/ + 1024 /*align slack*/;
};

template <int BN>
__global__ void __launch_bounds__(256, 1) conv_igemm_kernel(const __grid_constant__ ConvKernelParams p) {
  using Cfg = ConvCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.n_src; ++i) tma_prefetch_desc(&p.tmA[i]);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar + a, 1);
      mbar_init(tempty_bar + a, 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = tile / p.n_tiles;
        const int nt = tile - mt * p.n_tiles;
        const int row0 = mt * kBM;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(empty_bar + stage, phase ^ 1u);
          mbar_arrive_expect_tx(full_bar + stage, Cfg::STAGE_BYTES);
          tma_load_2d(smA + stage * Cfg::A_BYTES, &p.tmA[p.kb_src[kb]], full_bar + stage, p.kb_c0[kb],
                      row0 + p.kb_rowoff[kb]);
          tma_load_2d(smB + stage * Cfg::B_BYTES, &p.tmB, full_bar + stage, kb * kBK, nt * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ UMMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(kBM, BN, kFmtBF16);
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
        const int as = iter & 1;
        const uint32_t aphase = (iter >> 1) & 1u;
        mbar_wait(tempty_bar + as, aphase ^ 1u);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint64_t da = umma_desc_k_sw128(smem_u32(smA + stage * Cfg::A_BYTES));
          const uint64_t db = umma_desc_k_sw128(smem_u32(smB + stage * Cfg::B_BYTES));
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // +32 bytes (16 bf16) along K inside the 128B swizzle atom == +2 in the (addr>>4) field
            umma_f16(d_tmem, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                     (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar + stage);                          // smem slot free once these MMAs retire
          if (kb == p.num_kb - 1) umma_commit(tfull_bar + as);     // accumulator complete
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    float cx = 0.f, ce = 0.f, cn = 0.f;
    if (p.epilogue == IDF_EPI_SAMPLER) {
      const int step = p.step_ptr ? *p.step_ptr : 0;
      cx = p.coef[3 * step + 0];
      ce = p.coef[3 * step + 1];
      cn = p.coef[3 * step + 2];
    }
    int iter = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
      const int mt = tile / p.n_tiles;
      const int nt = tile - mt * p.n_tiles;
      const int as = iter & 1;
      const uint32_t aphase = (iter >> 1) & 1u;
      mbar_wait(tfull_bar + as, aphase);
      tc_fence_after();

      const int64_t r = static_cast<int64_t>(mt) * kBM + q * 32 + lane;
      bool valid = r < p.rows;
      int img = 0, y = 0, x = 0;
      if (valid) {
        const int rq = static_cast<int>(r / p.Wp);
        x = static_cast<int>(r - static_cast<int64_t>(rq) * p.Wp);
        img = rq / p.Hp;
        y = rq - img * p.Hp;
        valid = (x < p.W) && (y < p.H);
      }
      const uint32_t taddr = tmem_base + static_cast<uint32_t>(as * BN) + (static_cast<uint32_t>(q * 32) << 16);

      if constexpr (BN >= 32) {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + static_cast<uint32_t>(c * 32), v);
          tmem_ld_wait();
          if (valid) {
            const int col0 = nt * BN + c * 32;
            const float4* bp = reinterpret_cast<const float4*>(p.bias + col0);
            float f[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(bp + j);
              f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + b.x;
              f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b.y;
              f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b.z;
              f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b.w;
            }
            if (p.residual != nullptr) {
              const uint4* rp = reinterpret_cast<const uint4*>(p.residual + r * p.res_ld + col0);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 u = __ldg(rp + j);
                const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z),
                             a3 = unpack_bf16x2(u.w);
                f[8 * j + 0] += a0.x; f[8 * j + 1] += a0.y; f[8 * j + 2] += a1.x; f[8 * j + 3] += a1.y;
                f[8 * j + 4] += a2.x; f[8 * j + 5] += a2.y; f[8 * j + 6] += a3.x; f[8 * j + 7] += a3.y;
              }
            }
            uint4* op = reinterpret_cast<uint4*>(p.out + r * p.out_ld + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 u;
              u.x = pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
              u.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
              u.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
              u.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
              op[j] = u;
            }
          }
        }
      } else {
        // BN == 16: narrow outputs (eps / encoder map), fp32 NCHW store or fused sampler update
        uint32_t v[16];
        tmem_ld_32x16(taddr, v);
        tmem_ld_wait();
        if (valid) {
          const int64_t plane = static_cast<int64_t>(p.H) * p.W;
          const int64_t base = static_cast<int64_t>(img) * p.cout * plane + static_cast<int64_t>(y) * p.W + x;
#pragma unroll
          for (int ch = 0; ch < 16; ++ch) {
            if (ch < p.cout) {
              const float e = __uint_as_float(v[ch]) + __ldg(p.bias + ch);
              const int64_t o = base + ch * plane;
              if (p.epilogue == IDF_EPI_SAMPLER) {
                const float xv = p.x_io[o];
                const float nz = (p.noise != nullptr) ? __ldg(p.noise + o) : 0.f;
                p.x_io[o] = cx * xv + ce * e + cn * nz;
                if (p.out_f32 != nullptr) p.out_f32[o] = e;
              } else {
                p.out_f32[o] = e;
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar + as);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN>
static cudaError_t launch_bn(const ConvKernelParams& p, int grid, cudaStream_t stream) {
  using Cfg = ConvCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(Cfg::SMEM_BYTES));
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  conv_igemm_kernel<BN><<<grid, 256, Cfg::SMEM_BYTES, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_conv_igemm(const ConvKernelParams& p, int block_n, int grid, cudaStream_t stream) {
  switch (block_n) {
    case 128: return launch_bn<128>(p, grid, stream);
    case 64: return launch_bn<64>(p, grid, stream);
    case 16: return launch_bn<16>(p, grid, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace idf
