// wgrad_gemm: weight-gradient contraction of the PPO update's backward pass.
//
//   D[m, n] += alpha * sum_k A[k, m] * B[k, n]        fp16 x fp16 -> fp32 (TMEM) -> fp32 red.global.add
//
// A [Kdim, M1] and B [Kdim, N1] are the row-major activation / gradient matrices exactly as the forward and
// dgrad GEMMs left them in HBM (rows = frames x pixels): the contraction runs over ROWS, so both operands are
// "MN-major" for the tensor core.  Nothing is transposed: TMA brings [64 k][64 mn] boxes (128-B rows, 128B
// swizzle) and the UMMA shared-memory descriptors describe them as MN-major canonical tiles
// (cute make_umma_desc<Major::MN>: ((8,n),(8,k)) : ((1,LBO),(8,SBO)) in 16-B units, i.e. LBO = byte stride
// between 64-element MN groups = one box, SBO = byte stride between 8-row K groups = 1024 B); the instruction
// descriptor sets a_major = b_major = MN.
//
// K is huge (rows = 376,320 for the 2048->128 compressor) and the output tiny, so the grid is
// (m tiles x n tiles x K splits) and every CTA adds its partial tile into the zero-initialised fp32 gradient
// with red.global.add.  `alpha` (device scalar, may be null) divides out the loss scale carried by the fp16
// gradient operand.  Output addressing is general (m stride, n stride, per-n-tile offset) so a transposed
// weight or the (c,h,w)-flatten permutation of weight_ih_l0 is written in place.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue.
#pragma once
#include "ptx.cuh"

namespace embclip {

struct WgradParams {
  int M1, N1;                 // logical output extents
  int kb_total;               // ceil(Kdim / 64)
  int kb_per_split;
  int num_m_tiles, num_n_tiles, splits;
  float* out;
  long long ldo_m, ldo_n;     // element strides of D
  long long n_tile_off;       // added per n-tile index (normally BN * ldo_n)
  const float* alpha;         // device scalar or nullptr (= 1)
  int vec;                    // 1: rows of D are contiguous and 16-B aligned -> red.global.add.v4.f32
};

template <int BN>
struct WgradCfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr int kBoxN = BN < 64 ? BN : 64;              // B box width (elements)
  static constexpr int kABox = 64 * BK * 2;                    // 8 KB: [64 k][64 m]
  static constexpr int kABytes = 2 * kABox;
  static constexpr int kBBox = kBoxN * BK * 2;
  static constexpr int kBBoxes = BN / kBoxN;
  static constexpr int kBBytes = kBBoxes * kBBox;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagesRaw = (200 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
  static constexpr int kThreads = 192;
  static constexpr size_t kSmemBytes = 1024 + size_t(kStages) * kStageBytes + 256;
  static_assert(BN == 32 || BN == 64 || BN == 128 || BN == 256, "tile N");
  static_assert(kBBox % 1024 == 0, "boxes keep 1024-B alignment");
};

template <int BN>
__global__ void __launch_bounds__(192, 1)
wgrad_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradParams p) {
  using Cfg = WgradCfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + S * Cfg::kABytes;
  const uint32_t sBar = sB + S * Cfg::kBBytes;
  const uint32_t bar_full = sBar, bar_empty = sBar + 8 * S, bar_tfull = sBar + 16 * S;
  const uint32_t tmem_slot = bar_tfull + 8;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int t = blockIdx.x;
  const int split = t % p.splits;  t /= p.splits;
  const int n_tile = t % p.num_n_tiles;
  const int m_tile = t / p.num_n_tiles;
  const int kb0 = split * p.kb_per_split;
  const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
  const int nkb = kb1 - kb0;                    // >= 1 by construction of the launcher

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < S; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_tfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
        const uint32_t full = bar_full + 8 * stage;
        mbar_arrive_expect_tx(full, Cfg::kStageBytes);
        const uint32_t a = sA + stage * Cfg::kABytes, b = sB + stage * Cfg::kBBytes;
        tma_load_2d(&tmA, full, a, m_tile * 128, kb * 64);
        tma_load_2d(&tmA, full, a + Cfg::kABox, m_tile * 128 + 64, kb * 64);
#pragma unroll
        for (int j = 0; j < Cfg::kBBoxes; ++j)
          tma_load_2d(&tmB, full, b + j * Cfg::kBBox, n_tile * BN + j * Cfg::kBoxN, kb * 64);
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // fp16 x fp16 -> fp32, A and B both MN-major (bits 15, 16)
    constexpr uint32_t idesc = make_idesc_f16_f32(128, BN) | (1u << 15) | (1u << 16);
    constexpr int kBRow = Cfg::kBoxN * 2;
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < nkb; ++i) {
      mbar_wait(bar_full + 8 * stage, phase);
      tcgen05_fence_after();
      const uint32_t a = sA + stage * Cfg::kABytes, b = sB + stage * Cfg::kBBytes;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {           // 16 k-rows per MMA = two 8-row groups
          const uint64_t da = make_mnmajor_desc<128>(a + k * 2 * 1024, Cfg::kABox);
          const uint64_t db = make_mnmajor_desc<kBRow>(b + k * 2 * (8 * kBRow), Cfg::kBBox);
          umma_f16_ss(tmem_base, da, db, idesc, (i | k) != 0 ? 1u : 0u);
        }
        umma_commit(bar_empty + 8 * stage);
      }
      __syncwarp();
      if (++stage == S) { stage = 0; phase ^= 1u; }
    }
    if (elect_one()) umma_commit(bar_tfull);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int m = m_tile * 128 + q * 32 + lane;
    const float alpha = p.alpha ? __ldg(p.alpha) : 1.f;
    mbar_wait(bar_tfull, 0);
    tcgen05_fence_after();
    constexpr int CH = BN < 32 ? BN : 32;
    float* const orow = p.out + (long long)m * p.ldo_m + (long long)n_tile * p.n_tile_off;
#pragma unroll 1
    for (int c = 0; c < BN / CH; ++c) {
      uint32_t v[CH];
      tmem_ld_32x32b<CH>(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(c * CH), v);
      tmem_ld_wait();
      if (m < p.M1 && p.vec && n_tile * BN + (c + 1) * CH <= p.N1) {
        // contiguous row piece: 16-B vector reductions (ldo_m and the tile offset are multiples of 4 elements)
        float* o = orow + c * CH;
#pragma unroll
        for (int i = 0; i < CH; i += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + i), "f"(__uint_as_float(v[i]) * alpha),
                       "f"(__uint_as_float(v[i + 1]) * alpha), "f"(__uint_as_float(v[i + 2]) * alpha),
                       "f"(__uint_as_float(v[i + 3]) * alpha) : "memory");
      } else if (m < p.M1) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          const int n = c * CH + i;
          if (n_tile * BN + n < p.N1) atomicAdd(orow + (long long)n * p.ldo_n, __uint_as_float(v[i]) * alpha);
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace embclip
