// bneck_tail: the tail of one bottleneck and the head of the next one in a single pass over the wide tensor.
//
//   x'  = relu( [y2 | x0] . [W3 | Wd]^T + b3 (+ x) )          conv3 (+ K-concatenated downsample conv, or identity residual)
//   y1' = relu( x' . W1'^T + b1' )                             the NEXT block's conv1, fed from shared memory
//
// Why: in layer 1 (56x56, 256 channels) every 1x1 conv runs at the HBM roofline (profiles/r1e_launches.txt: 411 MB
// tensors, written once and read twice).  The conv3 epilogue already holds the x' tile in shared memory as fp16 in the
// 128-B-swizzled K-major layout (it is staged there for the TMA store) -- which is exactly the A operand the next
// block's conv1 needs.  So conv1' runs as a second UMMA on the staged tile and x' is never re-read from HBM.
// (clip/model.py Bottleneck.forward [UPSTREAM]: out = relu(bn3(conv3(..)) + identity); next block: relu(bn1(conv1(out))).)
//
// Pipeline unit = one 64-column QUARTER of a 128 x 256 output tile (staging buffer q, TMEM accumulator q):
//   R-thread   residual quarter  --TMA-->  staging[q]                                   (res_full[q])
//   MMA warp   acc3[q] = A . W3[64q..64q+63]^T      (N = 64, K = 64 or 128)             (acc3_full[q])
//   epilogue   staging[q] = fp16(relu(acc3[q] + b3 + staging[q]))   in place            (c_ready[q]; 4 warp arrivals)
//   R-thread   staging[q] --TMA store--> x'                    MMA warp: acc1 += staging[q] . W1'[:, 64q..]^T  (c_mma_done[q])
//   R-thread   (store read done, MMA retired) -> next tile's residual quarter into staging[q]
// so three residual quarters are always in flight while one is being consumed.  After the 4th quarter the conv1'
// accumulator (double-buffered by tile parity) is complete; its epilogue (bias, ReLU, fp16, st.global) runs in the
// middle of the NEXT tile's quarters, when the MMAs have long retired.  W3 and W1' stay resident in shared memory.
//
// Warp roles (640 threads): warp 0 = A producer (+ weights once), warp 1 = TMEM owner + MMA issuer, warp 2 = R-thread
// (residual loads, x' stores), warp 3 idle, warps 4..19 = epilogue (four per TMEM lane quarter, 16 columns each).
//
// kPool (the LAST block of a stage, whose x' is read only by the next stage's conv1 and by its downsample branch's
// AvgPool2d(2) / stride-2 subsample): a tile is two image rows (2 W <= 128 pixels; the UMMA still runs 128 rows, the spare ones
// are ignored) brought in WINDOW-MAJOR order by 5-D tensor maps (dims: channel, dx, dy, window, row pair), so the four pixels
// of a 2x2 window are four consecutive tile rows = TMEM lanes of ONE epilogue warp.  After staging its rows the warp reduces
// its eight windows straight from shared memory and writes the POOLED tensor; x' itself never goes to HBM: at 56x56x256 that
// is one 411 MB write, one 411 MB read and the pool launch saved (clip/model.py Bottleneck.downsample [UPSTREAM]:
// AvgPool2d(stride) -> conv1x1 -> bn on the block input).
#pragma once
#include "ptx.cuh"

namespace embclip {

struct TailParams {
  int num_tiles;             // ceil(M / 128)
  int M;
  int reverse;
  int tile_rows;             // 128, or 2 W with kPool
  int pool_w;                // kPool: W / 2 pooled pixels per tile
  int pool_mode;             // kPool: 1 = average of the 2x2 window, 2 = its top-left pixel
  __half* pool_out;          // kPool: [M / 4, 256]
  const float* bias3;        // [256]
  const float* bias1;        // [N1]
  __half* y1;                // [M, N1]
};

template <int K3C, int N1>
struct TailCfg {
  static constexpr int kN3 = 256;
  static constexpr int kW3Bytes = K3C * kN3 * 128;
  static constexpr int kW1Bytes = 4 * N1 * 128;
  static constexpr int kAStage = 128 * 128;                    // one 64-channel k-chunk of 128 rows
  static constexpr int kAStages = 4;
  static constexpr int kCQuarter = 128 * 128;
  static constexpr int kBiasBytes = (kN3 + N1) * 4;
  static constexpr int kBarBytes = 320;
  static constexpr int kParts = 4;                             // epilogue warps per TMEM lane quarter: 64 / kParts columns each
  static constexpr int kEpiWarps = 4 * kParts;
  static constexpr int kThreads = 128 + 32 * kEpiWarps;
  static constexpr int kCP = 64 / kParts;                      // conv3 columns per epilogue warp and quarter
  static constexpr int kC1 = N1 / kParts;                      // conv1' columns per epilogue warp
  static constexpr size_t kSmemBytes = 1024 + kW3Bytes + kW1Bytes + kAStages * kAStage + 4 * kCQuarter + kBiasBytes + kBarBytes;
  static_assert(kSmemBytes <= 232448, "shared memory budget");
  static_assert(kAStages % K3C == 0, "a tile's k-chunks must not straddle the ring wrap");
  static_assert(N1 == 64 || N1 == 128, "next conv1 width");
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1) : "memory");
}

template <int K3C, int N1, bool kRes, bool kPool = false>
__global__ void __launch_bounds__(TailCfg<K3C, N1>::kThreads, 1)
bneck_tail_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                  const __grid_constant__ CUtensorMap tmW3, const __grid_constant__ CUtensorMap tmW1,
                  const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmC, const TailParams p) {
  using Cfg = TailCfg<K3C, N1>;
  constexpr int SA = Cfg::kAStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sW3 = smem_base;
  const uint32_t sW1 = sW3 + Cfg::kW3Bytes;
  const uint32_t sA = sW1 + Cfg::kW1Bytes;
  const uint32_t sC = sA + SA * Cfg::kAStage;
  const uint32_t sBias = sC + 4 * Cfg::kCQuarter;
  const uint32_t sBar = sBias + Cfg::kBiasBytes;
  const uint32_t bar_w = sBar;                       // weights landed
  const uint32_t bar_afull = sBar + 8;               // SA
  const uint32_t bar_aempty = bar_afull + 8 * SA;    // SA
  const uint32_t bar_acc3 = bar_aempty + 8 * SA;     // 4: conv3 quarter accumulated
  const uint32_t bar_res = bar_acc3 + 32;            // 4: staging quarter free (+ residual landed)
  const uint32_t bar_cready = bar_res + 32;          // 4: x' quarter written to staging (4 warp arrivals: one per lane quarter)
  const uint32_t bar_cdone = bar_cready + 32;        // 4: conv1' MMAs on the quarter retired
  const uint32_t bar_acc1f = bar_cdone + 32;         // 2
  const uint32_t bar_acc1e = bar_acc1f + 16;         // 2 (8 warp arrivals)
  const uint32_t tmem_slot = bar_acc1e + 16;
  float* const sBias3 = reinterpret_cast<float*>(gen_base + (sBias - smem_base));
  float* const sBias1 = sBias3 + Cfg::kN3;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_tiles;
  const int my_tiles = (int(blockIdx.x) < num_tiles) ? (num_tiles - 1 - int(blockIdx.x)) / int(gridDim.x) + 1 : 0;
  auto tile_idx = [&](int it) {
    const int t = int(blockIdx.x) + it * int(gridDim.x);
    return p.reverse ? num_tiles - 1 - t : t;
  };
  auto tile_m0 = [&](int it) { return tile_idx(it) * (kPool ? p.tile_rows : 128); };
  const uint32_t a_bytes = kPool ? uint32_t(p.tile_rows) * 128u : uint32_t(Cfg::kAStage);    // one TMA box: tile rows x 64 channels

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmW3);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmC);
    if (K3C == 2) tma_prefetch_desc(&tmA1);
    if (kRes) tma_prefetch_desc(&tmR);
    mbar_init(bar_w, 1);
    for (int s = 0; s < SA; ++s) { mbar_init(bar_afull + 8 * s, 1); mbar_init(bar_aempty + 8 * s, 1); }
    for (int q = 0; q < 4; ++q) {
      mbar_init(bar_acc3 + 8 * q, 1);
      mbar_init(bar_res + 8 * q, 1);
      mbar_init(bar_cready + 8 * q, 4);
      mbar_init(bar_cdone + 8 * q, 1);
    }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_acc1f + 8 * a, 1); mbar_init(bar_acc1e + 8 * a, Cfg::kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  for (int i = threadIdx.x; i < Cfg::kN3 + N1; i += blockDim.x)
    sBias3[i] = i < Cfg::kN3 ? __ldg(p.bias3 + i) : __ldg(p.bias1 + i - Cfg::kN3);       // parameters: not produced upstream
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (warp == 0 && lane == 0 && my_tiles > 0) {
    // resident weights (frozen parameters: safe to fetch while the previous kernel is still draining)
    mbar_arrive_expect_tx(bar_w, Cfg::kW3Bytes + Cfg::kW1Bytes);
    for (int kc = 0; kc < K3C; ++kc) tma_load_2d(&tmW3, bar_w, sW3 + kc * (Cfg::kN3 * 128), kc * 64, 0);
    for (int q = 0; q < 4; ++q) tma_load_2d(&tmW1, bar_w, sW1 + q * (N1 * 128), q * 64, 0);
  }
  griddep_launch_dependents();
  griddep_wait();

  if (warp == 0) {
    // ============================ A producer ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int m0 = tile_m0(it);
        for (int kc = 0; kc < K3C; ++kc) {
          mbar_wait(bar_aempty + 8 * stage, phase ^ 1u);
          const uint32_t full = bar_afull + 8 * stage;
          mbar_arrive_expect_tx(full, a_bytes);
          if (kPool) tma_load_5d(kc == 0 ? &tmA0 : &tmA1, full, sA + stage * Cfg::kAStage, 0, 0, 0, 0, tile_idx(it));
          else tma_load_2d(kc == 0 ? &tmA0 : &tmA1, full, sA + stage * Cfg::kAStage, 0, m0);
          if (++stage == SA) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ============================ R-thread: residual quarters in, x' quarters out ============================
    if (lane == 0 && my_tiles > 0) {
      {
        const int m0 = tile_m0(0);
        for (int q = 0; q < 4; ++q) {
          if (kRes) {
            mbar_arrive_expect_tx(bar_res + 8 * q, a_bytes);
            if (kPool) tma_load_5d(&tmR, bar_res + 8 * q, sC + q * Cfg::kCQuarter, q * 64, 0, 0, 0, tile_idx(0));
            else tma_load_2d(&tmR, bar_res + 8 * q, sC + q * Cfg::kCQuarter, q * 64, m0);
          } else {
            mbar_arrive(bar_res + 8 * q);
          }
        }
      }
      for (int it = 0; it < my_tiles; ++it) {
        const int m0 = tile_m0(it);
        const uint32_t par = uint32_t(it & 1);
        const bool more = it + 1 < my_tiles;
        const int m0n = more ? tile_m0(it + 1) : 0;
        for (int q = 0; q < 4; ++q) {
          if (!kPool) {
            mbar_wait(bar_cready + 8 * q, par);
            tma_store_2d(&tmC, sC + q * Cfg::kCQuarter, q * 64, m0);
            tma_store_commit();
          }
          if (more) {
            if (!kPool) tma_store_wait_read0();                // the store has finished reading staging[q]
            mbar_wait(bar_cdone + 8 * q, par);                 // ... and so has conv1' (issued after c_ready: the pooling reads too)
            if (kRes) {
              mbar_arrive_expect_tx(bar_res + 8 * q, a_bytes);
              if (kPool) tma_load_5d(&tmR, bar_res + 8 * q, sC + q * Cfg::kCQuarter, q * 64, 0, 0, 0, tile_idx(it + 1));
              else tma_load_2d(&tmR, bar_res + 8 * q, sC + q * Cfg::kCQuarter, q * 64, m0n);
            } else {
              mbar_arrive(bar_res + 8 * q);
            }
          }
        }
      }
      tma_store_wait_all0();
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    constexpr uint32_t idesc3 = make_idesc_f16_f32(128, 64);
    constexpr uint32_t idesc1 = make_idesc_f16_f32(128, N1);
    constexpr uint32_t dhi = kmajor_desc_hi<128>();
    const uint32_t sA_lo = kmajor_desc_lo(sA), sW3_lo = kmajor_desc_lo(sW3), sW1_lo = kmajor_desc_lo(sW1), sC_lo = kmajor_desc_lo(sC);
    const uint32_t acc3 = tmem_base, acc1 = tmem_base + 256;
    if (my_tiles > 0) mbar_wait(bar_w, 0);
    int stage = 0;
    uint32_t aphase = 0;
    // conv1' on quarter q of tile `it`
    auto mma1 = [&](int it, int q) {
      mbar_wait(bar_cready + 8 * q, uint32_t(it & 1));
      if (q == 0) mbar_wait(bar_acc1e + 8 * (it & 1), (uint32_t(it >> 1) & 1u) ^ 1u);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t d = acc1 + uint32_t((it & 1) * N1);
        const uint32_t a_lo = sC_lo + uint32_t(q) * (Cfg::kCQuarter / 16);
        const uint32_t b_lo = sW1_lo + uint32_t(q) * (N1 * 128 / 16);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(d, desc64(a_lo + 2 * k, dhi), desc64(b_lo + 2 * k, dhi), idesc1, uint32_t((q | k) != 0));
        umma_commit(bar_cdone + 8 * q);
        if (q == 3) umma_commit(bar_acc1f + 8 * (it & 1));
      }
      __syncwarp();
    };
    for (int it = 0; it < my_tiles; ++it) {
      for (int kc = 0; kc < K3C; ++kc) {
        int s = stage + kc;
        mbar_wait(bar_afull + 8 * s, aphase);
      }
      tcgen05_fence_after();
      for (int q = 0; q < 4; ++q) {
        // accumulator q was drained by the previous tile's epilogue before it signalled c_ready[q], which mma1() of
        // that tile has waited on; quarter 3 of the previous tile is issued here, between this tile's quarters 2 and 3
        if (q == 3 && it > 0) mma1(it - 1, 3);
        if (elect_one()) {
#pragma unroll
          for (int kc = 0; kc < K3C; ++kc) {
            const uint32_t a_lo = sA_lo + uint32_t(stage + kc) * (Cfg::kAStage / 16);
            const uint32_t b_lo = sW3_lo + uint32_t(kc) * (Cfg::kN3 * 128 / 16) + uint32_t(q) * (64 * 128 / 16);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ss(acc3 + uint32_t(q * 64), desc64(a_lo + 2 * k, dhi), desc64(b_lo + 2 * k, dhi), idesc3, uint32_t((kc | k) != 0));
          }
          umma_commit(bar_acc3 + 8 * q);
          if (q == 3) {
#pragma unroll
            for (int kc = 0; kc < K3C; ++kc) umma_commit(bar_aempty + 8 * (stage + kc));
          }
        }
        __syncwarp();
      }
      stage += K3C;
      if (stage == SA) { stage = 0; aphase ^= 1u; }
      for (int q = 0; q < 3; ++q) mma1(it, q);
    }
    if (my_tiles > 0) mma1(my_tiles - 1, 3);
  } else if (warp >= 4) {
    // ============================ epilogue (warps 4 .. 4 + kEpiWarps - 1) ============================
    // kParts warps share a TMEM lane quarter (32 tile rows); warp (lq, part) owns quarter `part` of every tile and a 1/kParts
    // column slice of the conv1' output.
    constexpr int CP = Cfg::kCP, C1 = Cfg::kC1;
    const int lq = warp & 3;                                   // TMEM lane quarter
    const int row = lq * 32 + lane;
    const int part = (warp - 4) >> 2;
    const uint32_t lane_addr = tmem_base + (uint32_t(lq * 32) << 16);
    auto epi1 = [&](int it) {                                  // conv1' of tile `it`: bias, ReLU, fp16, straight to global
      const int m0 = tile_m0(it);
      mbar_wait(bar_acc1f + 8 * (it & 1), uint32_t(it >> 1) & 1u);
      tcgen05_fence_after();
      // kPool: tile row r = 4 window + 2 dy + dx  ->  pixel m0 + dy W + 2 window + dx
      const int prow = kPool ? ((row >> 1) & 1) * (2 * p.pool_w) + 2 * (row >> 2) + (row & 1) : row;
      const bool ok = m0 + prow < p.M && (!kPool || row < p.tile_rows);
      __half* const dst = p.y1 + size_t(m0 + prow) * N1;
      const int col = part * C1;
      uint32_t v[C1];
      tmem_ld_32x32b<C1>(lane_addr + uint32_t(256 + (it & 1) * N1 + col), v);
      tmem_ld_wait();
      if (ok) {
#pragma unroll
        for (int i = 0; i < C1 / 8; ++i) {
          uint32_t h[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float f0 = __uint_as_float(v[8 * i + 2 * j]), f1 = __uint_as_float(v[8 * i + 2 * j + 1]);
            add_f32x2(f0, f1, sBias1[col + 8 * i + 2 * j], sBias1[col + 8 * i + 2 * j + 1]);
            h[j] = relu_pack_half2(f0, f1);
          }
          *reinterpret_cast<uint4*>(dst + col + 8 * i) = make_uint4(h[0], h[1], h[2], h[3]);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc1e + 8 * (it & 1));
    };
    // Quarter q of EVERY tile belongs to the four warps with part == q (one per TMEM lane quarter): the four quarter chains
    // (accumulator wait -> TMEM load -> residual -> fp16 -> staging -> c_ready) of a tile run side by side instead of one after
    // the other in each warp, which is what bounded the tile period (measured ~1 us per quarter whatever the byte count).
    const int q = part;
    const uint32_t qbase = sC + uint32_t(q) * Cfg::kCQuarter;
    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t par = uint32_t(it & 1);
      mbar_wait(bar_acc3 + 8 * q, par);
      tcgen05_fence_after();
      uint32_t v[2][CP];
      tmem_ld_32x32b<CP>(lane_addr + uint32_t(q * 64), v[0]);
      mbar_wait(bar_res + 8 * q, par);
#pragma unroll
      for (int sub = 0; sub < 64 / CP; ++sub) {
        const int col = q * 64 + sub * CP;
        uint4 r[CP / 8];
        if (kRes) {
#pragma unroll
          for (int i = 0; i < CP / 8; ++i) {
            const uint32_t a = qbase + swizzle_off<128>(uint32_t(row), uint32_t(sub * (CP / 8) + i));
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[i].x), "=r"(r[i].y), "=r"(r[i].z), "=r"(r[i].w) : "r"(a));
          }
        }
        tmem_ld_wait();
        if (sub + 1 < 64 / CP) tmem_ld_32x32b<CP>(lane_addr + uint32_t(col + CP), v[(sub + 1) & 1]);   // next chunk in flight
        const uint32_t (&vv)[CP] = v[sub & 1];
#pragma unroll
        for (int i = 0; i < CP / 8; ++i) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(vv[8 * i + j]);
#pragma unroll
          for (int j = 0; j < 4; ++j) add_f32x2(f[2 * j], f[2 * j + 1], sBias3[col + 8 * i + 2 * j], sBias3[col + 8 * i + 2 * j + 1]);
          if (kRes) {
            const __half2* h = reinterpret_cast<const __half2*>(&r[i]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 r2 = __half22float2(h[j]);
              add_f32x2(f[2 * j], f[2 * j + 1], r2.x, r2.y);
            }
          }
          const uint32_t a = qbase + swizzle_off<128>(uint32_t(row), uint32_t(sub * (CP / 8) + i));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a),
                       "r"(relu_pack_half2(f[0], f[1])), "r"(relu_pack_half2(f[2], f[3])),
                       "r"(relu_pack_half2(f[4], f[5])), "r"(relu_pack_half2(f[6], f[7]))
                       : "memory");
        }
      }
      if (kPool) {
        // this warp's 32 rows are 8 whole windows: item = (window, 16-B chunk of the quarter's 64 channels), two per lane.
        // Same arithmetic as avgpool2_kernel (fp32 sum in window order, * 0.25, one fp16 rounding): bit-identical to
        // pooling a stored x'.
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int wl = (lane >> 3) + 4 * k, j = lane & 7;
          const int r0 = lq * 32 + 4 * wl;
          if (r0 < p.tile_rows) {
            uint4 x4[4];
#pragma unroll
            for (int wi = 0; wi < 4; ++wi)
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x4[wi].x), "=r"(x4[wi].y), "=r"(x4[wi].z), "=r"(x4[wi].w)
                           : "r"(qbase + swizzle_off<128>(uint32_t(r0 + wi), uint32_t(j))));
            uint4 o = x4[0];                                   // pool_mode 2: the window's top-left pixel
            if (p.pool_mode == 1) {
              float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
              for (int wi = 0; wi < 4; ++wi) {
                const __half2* h = reinterpret_cast<const __half2*>(&x4[wi]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = __half22float2(h[e]);
                  add_f32x2(a[2 * e], a[2 * e + 1], f.x, f.y);
                }
              }
              o = make_uint4(pack_half2(a[0] * .25f, a[1] * .25f), pack_half2(a[2] * .25f, a[3] * .25f),
                             pack_half2(a[4] * .25f, a[5] * .25f), pack_half2(a[6] * .25f, a[7] * .25f));
            }
            const size_t prow = size_t(tile_idx(it)) * size_t(p.pool_w) + size_t(r0 >> 2);
            *reinterpret_cast<uint4*>(p.pool_out + prow * Cfg::kN3 + q * 64 + 8 * j) = o;
          }
        }
      }
      tcgen05_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_cready + 8 * q);
      if (it > 0) epi1(it - 1);
    }
    if (my_tiles > 0) epi1(my_tiles - 1);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace embclip
