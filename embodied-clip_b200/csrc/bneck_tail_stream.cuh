// bneck_tail_stream: bneck_tail (conv3 + identity residual + ReLU, then the NEXT block's conv1 on the staged tile) for the stages
// whose weights do not fit shared memory -- layer 2: conv3 128 -> 512, next conv1 512 -> 128: W3 and W1' are 128 KB each.
//
//   x'  = relu( y2 . W3^T + b3 + x )             [M, N3]   N3 = 64 * nq columns, produced quarter by quarter (64 columns)
//   y1' = relu( x' . W1'^T + b1' )               [M, N1]   accumulated over the nq quarters as they are staged
//
// Why: the layer-2 conv3 launches run at 5.4 TB/s of HBM with the tensor pipe idle (profiles/r2_launches.txt: 462 MB in 86 us),
// and the conv1 that follows re-reads the 205 MB tensor they just wrote (56 us more).  Fused, the x' tile never comes back from HBM.
//
// Same pipeline as bneck_tail.cuh with two changes: (1) the pipeline unit is a GLOBAL quarter counter u = tile * nq + q -- every
// ring (4 staging buffers, 4 conv3 accumulators, the 2-deep weight rings) is indexed u % depth with phase (u / depth) & 1, so a
// tile may have any number of quarters; (2) the weights STREAM: two producer threads bring W3[64q .. 64q+63, :] and
// W1'[:, 64q .. 64q+63] of every quarter through two 2-deep rings (32 KB per quarter at layer 2; per 128-row tile the weights are
// 256 KB of L2 -> smem traffic next to 160 KB of activations -- they come from L2, not HBM).  The conv1' UMMA of quarter u is issued
// kLag = 3 quarters behind the conv3 UMMA, so the epilogue always has two quarters of slack.
//
// Warp roles (640 threads): warp 0 = A producer, warp 1 = TMEM owner + MMA issuer, warp 2 = R-thread (residual loads, x' stores),
// warp 3 = W1' producer (warp 0 also streams W3), warps 4..19 = epilogue: ring slot s (units u with u % 4 == s) belongs to the
// four warps with part == s, one per TMEM lane quarter, so four units are in flight in the epilogue at once (as in bneck_tail).
#pragma once
#include "bneck_tail.cuh"

namespace embclip {

struct TailStreamParams {
  int num_tiles;             // ceil(M / 128)
  int M;
  int nq;                    // N3 / 64
  int reverse;
  const float* bias3;        // [N3]
  const float* bias1;        // [N1]
  __half* y1;                // [M, N1]
};

template <int K3C, int N1>
struct TailStreamCfg {
  static constexpr int kMaxNQ = 16;
  static constexpr int kAStage = 128 * 128;                    // one 64-channel k-chunk of 128 rows
  static constexpr int kAStages = 2 * K3C;                     // two tiles of A
  static constexpr int kW3Q = K3C * 64 * 128;                  // W3 rows of one quarter: K3C chunks of [64 rows][128 B]
  static constexpr int kW1Q = N1 * 128;                        // W1' columns of one quarter: [N1 rows][128 B]
  static constexpr int kWDepth = 2;                            // W1' ring
  static constexpr int kW3Depth = 3;                           // W3 ring (runs ahead of the MMAs; W1' is consumed kLag quarters later)
  static constexpr int kCQuarter = 128 * 128;
  static constexpr int kBiasBytes = (kMaxNQ * 64 + N1) * 4;
  static constexpr int kBarBytes = 384;
  static constexpr int kEpiWarps = 16;
  static constexpr int kThreads = 128 + 32 * kEpiWarps;
  static constexpr int kCP = 16;                               // conv3 columns per TMEM load
  static constexpr int kC1 = N1 / 4;                           // conv1' columns per epilogue warp
  static constexpr int kLag = 3;
  static constexpr size_t kSmemBytes = 1024 + kAStages * kAStage + (kW3Depth * kW3Q + kWDepth * kW1Q) + 4 * kCQuarter + kBiasBytes + kBarBytes;
  static_assert(kSmemBytes <= 232448, "shared memory budget");
  static_assert(2 * N1 <= 256, "conv1' accumulators (double-buffered by tile parity) share TMEM with the 4 x 64 conv3 columns");
  static_assert(N1 % 16 == 0 && N1 >= 32, "UMMA N");
};

template <int K3C, int N1>
__global__ void __launch_bounds__(TailStreamCfg<K3C, N1>::kThreads, 1)
bneck_tail_stream_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW3,
                         const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmR,
                         const __grid_constant__ CUtensorMap tmC, const TailStreamParams p) {
  using Cfg = TailStreamCfg<K3C, N1>;
  constexpr int SA = Cfg::kAStages, WD = Cfg::kWDepth, W3D = Cfg::kW3Depth;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sA = smem_base;
  const uint32_t sW3 = sA + SA * Cfg::kAStage;
  const uint32_t sW1 = sW3 + W3D * Cfg::kW3Q;
  const uint32_t sC = sW1 + WD * Cfg::kW1Q;
  const uint32_t sBias = sC + 4 * Cfg::kCQuarter;
  const uint32_t sBar = sBias + Cfg::kBiasBytes;
  const uint32_t bar_afull = sBar;                   // SA
  const uint32_t bar_aempty = bar_afull + 8 * SA;    // SA
  const uint32_t bar_w3f = bar_aempty + 8 * SA;      // W3D
  const uint32_t bar_w3e = bar_w3f + 8 * W3D;        // W3D
  const uint32_t bar_w1f = bar_w3e + 8 * W3D;        // WD
  const uint32_t bar_w1e = bar_w1f + 8 * WD;         // WD
  const uint32_t bar_acc3 = bar_w1e + 8 * WD;        // 4: conv3 quarter accumulated
  const uint32_t bar_res = bar_acc3 + 32;            // 4: staging quarter free + residual landed
  const uint32_t bar_cready = bar_res + 32;          // 4: x' quarter written to staging (8 warp arrivals)
  const uint32_t bar_cdone = bar_cready + 32;        // 4: conv1' MMAs on the quarter retired
  const uint32_t bar_acc1f = bar_cdone + 32;         // 2
  const uint32_t bar_acc1e = bar_acc1f + 16;         // 2 (8 warp arrivals)
  const uint32_t tmem_slot = bar_acc1e + 16;
  static_assert(16 * SA + 16 * W3D + 16 * WD + 4 * 32 + 32 + 4 <= Cfg::kBarBytes, "barrier block");
  float* const sBias3 = reinterpret_cast<float*>(gen_base + (sBias - smem_base));
  float* const sBias1 = sBias3 + Cfg::kMaxNQ * 64;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_tiles, nq = p.nq;
  const int my_tiles = (int(blockIdx.x) < num_tiles) ? (num_tiles - 1 - int(blockIdx.x)) / int(gridDim.x) + 1 : 0;
  const int units = my_tiles * nq;
  auto tile_m0 = [&](int it) {
    const int t = int(blockIdx.x) + it * int(gridDim.x);
    return (p.reverse ? num_tiles - 1 - t : t) * 128;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW3); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmC); tma_prefetch_desc(&tmR);
    for (int s = 0; s < SA; ++s) { mbar_init(bar_afull + 8 * s, 1); mbar_init(bar_aempty + 8 * s, 1); }
    for (int s = 0; s < W3D; ++s) { mbar_init(bar_w3f + 8 * s, 1); mbar_init(bar_w3e + 8 * s, 1); }
    for (int s = 0; s < WD; ++s) { mbar_init(bar_w1f + 8 * s, 1); mbar_init(bar_w1e + 8 * s, 1); }
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
  for (int i = threadIdx.x; i < nq * 64 + N1; i += blockDim.x)
    sBias3[i < nq * 64 ? i : Cfg::kMaxNQ * 64 + (i - nq * 64)] = i < nq * 64 ? __ldg(p.bias3 + i) : __ldg(p.bias1 + i - nq * 64);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  griddep_launch_dependents();
  griddep_wait();

  if (warp == 0) {
    // ============================ A + W3 producer ============================
    // One thread, in need order: A of tile it + 1 is requested before the W3 quarters of tile it, so the next tile's operand is
    // in flight while this tile's quarters stream (the A ring holds two tiles; W3 runs at most two quarters ahead of the MMAs).
    if (lane == 0 && my_tiles > 0) {
      auto load_a = [&](int it) {
        const int m0 = tile_m0(it), sbase = (it & 1) * K3C;
        const uint32_t ph = (uint32_t(it >> 1) & 1u) ^ 1u;
        for (int kc = 0; kc < K3C; ++kc) {
          mbar_wait(bar_aempty + 8 * (sbase + kc), ph);
          mbar_arrive_expect_tx(bar_afull + 8 * (sbase + kc), Cfg::kAStage);
          tma_load_2d(&tmA, bar_afull + 8 * (sbase + kc), sA + (sbase + kc) * Cfg::kAStage, kc * 64, m0);
        }
      };
      load_a(0);
      int u = 0;
      for (int it = 0; it < my_tiles; ++it) {
        for (int q = 0; q < nq; ++q, ++u) {
          // (after this tile's first two weight quarters: by then the previous tile's conv3 MMAs have retired and its A stages are free)
          if (q == 2 && it + 1 < my_tiles) load_a(it + 1);
          const int ws = u % W3D;
          mbar_wait(bar_w3e + 8 * ws, (uint32_t(u / W3D) & 1u) ^ 1u);
          mbar_arrive_expect_tx(bar_w3f + 8 * ws, Cfg::kW3Q);
          for (int kc = 0; kc < K3C; ++kc) tma_load_2d(&tmW3, bar_w3f + 8 * ws, sW3 + ws * Cfg::kW3Q + kc * (64 * 128), kc * 64, q * 64);
        }
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ============================ W1' producer (its consumer runs kLag quarters behind conv3: its own thread) ============================
    if (lane == 0) {
      for (int u = 0; u < units; ++u) {
        const int q = u % nq, ws = u % WD;
        mbar_wait(bar_w1e + 8 * ws, (uint32_t(u / WD) & 1u) ^ 1u);
        mbar_arrive_expect_tx(bar_w1f + 8 * ws, Cfg::kW1Q);
        tma_load_2d(&tmW1, bar_w1f + 8 * ws, sW1 + ws * Cfg::kW1Q, q * 64, 0);
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ============================ R-thread: residual quarters in, x' quarters out ============================
    if (lane == 0 && units > 0) {
      auto load_res = [&](int u) {
        const int it = u / nq, q = u - it * nq;
        mbar_arrive_expect_tx(bar_res + 8 * (u & 3), Cfg::kCQuarter);
        tma_load_2d(&tmR, bar_res + 8 * (u & 3), sC + (u & 3) * Cfg::kCQuarter, q * 64, tile_m0(it));
      };
      for (int u = 0; u < 4 && u < units; ++u) load_res(u);
      for (int u = 0; u < units; ++u) {
        const int it = u / nq, q = u - it * nq;
        const uint32_t par = uint32_t(u >> 2) & 1u;
        mbar_wait(bar_cready + 8 * (u & 3), par);
        tma_store_2d(&tmC, sC + (u & 3) * Cfg::kCQuarter, q * 64, tile_m0(it));
        tma_store_commit();
        if (u + 4 < units) {
          tma_store_wait_read0();                              // the store has finished reading staging[u & 3]
          mbar_wait(bar_cdone + 8 * (u & 3), par);             // ... and so has conv1'
          load_res(u + 4);
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
    // conv1' on unit v (tile v / nq, quarter v % nq)
    auto mma1 = [&](int v) {
      const int it = v / nq, q = v - it * nq, ws = v % WD;
      mbar_wait(bar_cready + 8 * (v & 3), uint32_t(v >> 2) & 1u);
      if (q == 0) mbar_wait(bar_acc1e + 8 * (it & 1), (uint32_t(it >> 1) & 1u) ^ 1u);
      mbar_wait(bar_w1f + 8 * ws, uint32_t(v / WD) & 1u);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t d = acc1 + uint32_t((it & 1) * N1);
        const uint32_t a_lo = sC_lo + uint32_t(v & 3) * (Cfg::kCQuarter / 16);
        const uint32_t b_lo = sW1_lo + uint32_t(ws) * (Cfg::kW1Q / 16);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(d, desc64(a_lo + 2 * k, dhi), desc64(b_lo + 2 * k, dhi), idesc1, uint32_t((q | k) != 0));
        umma_commit(bar_cdone + 8 * (v & 3));
        umma_commit(bar_w1e + 8 * ws);
        if (q == nq - 1) umma_commit(bar_acc1f + 8 * (it & 1));
      }
      __syncwarp();
    };
    int u = 0;
    for (int it = 0; it < my_tiles; ++it) {
      const int stage = (it & 1) * K3C;
      const uint32_t aphase = uint32_t(it >> 1) & 1u;
      for (int kc = 0; kc < K3C; ++kc) mbar_wait(bar_afull + 8 * (stage + kc), aphase);
      tcgen05_fence_after();
      for (int q = 0; q < nq; ++q, ++u) {
        // accumulator slot u & 3 was drained by the epilogue of unit u - 4 before it signalled c_ready, which mma1(u - 4) has
        // waited on: kLag <= 4 keeps that order
        if (u >= Cfg::kLag) mma1(u - Cfg::kLag);
        const int ws = u % W3D;
        mbar_wait(bar_w3f + 8 * ws, uint32_t(u / W3D) & 1u);
        tcgen05_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int kc = 0; kc < K3C; ++kc) {
            const uint32_t a_lo = sA_lo + uint32_t(stage + kc) * (Cfg::kAStage / 16);
            const uint32_t b_lo = sW3_lo + uint32_t(ws) * (Cfg::kW3Q / 16) + uint32_t(kc) * (64 * 128 / 16);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ss(acc3 + uint32_t((u & 3) * 64), desc64(a_lo + 2 * k, dhi), desc64(b_lo + 2 * k, dhi), idesc3, uint32_t((kc | k) != 0));
          }
          umma_commit(bar_acc3 + 8 * (u & 3));
          umma_commit(bar_w3e + 8 * ws);
          if (q == nq - 1) {
#pragma unroll
            for (int kc = 0; kc < K3C; ++kc) umma_commit(bar_aempty + 8 * (stage + kc));
          }
        }
        __syncwarp();
      }
    }
    for (int v = units > Cfg::kLag ? units - Cfg::kLag : 0; v < units; ++v) mma1(v);
  } else if (warp >= 4) {
    // ============================ epilogue (warps 4..19) ============================
    constexpr int CP = Cfg::kCP, C1 = Cfg::kC1;
    const int lq = warp & 3;                                   // TMEM lane quarter
    const int row = lq * 32 + lane;
    const int part = (warp - 4) >> 2;                          // ring slot owned by this warp; conv1' column slice
    const uint32_t lane_addr = tmem_base + (uint32_t(lq * 32) << 16);
    auto epi1 = [&](int it) {                                  // conv1' of tile `it`: bias, ReLU, fp16, straight to global
      const int m0 = tile_m0(it);
      mbar_wait(bar_acc1f + 8 * (it & 1), uint32_t(it >> 1) & 1u);
      tcgen05_fence_after();
      const bool ok = m0 + row < p.M;
      __half* const dst = p.y1 + size_t(m0 + row) * N1;
      const int col = part * C1;
      uint32_t v[C1];
      tmem_ld_32x32b<C1>(lane_addr + uint32_t(256 + (it & 1) * N1 + col), v);
      tmem_ld_wait();
      if (ok) {
#pragma unroll
        for (int i = 0; i < C1 / 8; ++i) {
          uint32_t h[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            h[j] = pack_half2(fmaxf(__uint_as_float(v[8 * i + 2 * j]) + sBias1[col + 8 * i + 2 * j], 0.f),
                              fmaxf(__uint_as_float(v[8 * i + 2 * j + 1]) + sBias1[col + 8 * i + 2 * j + 1], 0.f));
          *reinterpret_cast<uint4*>(dst + col + 8 * i) = make_uint4(h[0], h[1], h[2], h[3]);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc1e + 8 * (it & 1));
    };
    const uint32_t qbase = sC + uint32_t(part) * Cfg::kCQuarter;
    const int total_units = my_tiles * nq;
    int done_tile = -1;                                        // conv1' epilogues this warp has run: tiles 0 .. done_tile
    for (int u = part; u < total_units; u += 4) {
      const int it = u / nq, q = u - it * nq;
      const uint32_t par = uint32_t(u >> 2) & 1u;
      mbar_wait(bar_acc3 + 8 * part, par);
      tcgen05_fence_after();
      uint32_t v[2][CP];
      tmem_ld_32x32b<CP>(lane_addr + uint32_t(part * 64), v[0]);
      mbar_wait(bar_res + 8 * part, par);
#pragma unroll
      for (int sub = 0; sub < 64 / CP; ++sub) {
        const int col = q * 64 + sub * CP;
        uint4 r[CP / 8];
#pragma unroll
        for (int i = 0; i < CP / 8; ++i) {
          const uint32_t a = qbase + swizzle_off<128>(uint32_t(row), uint32_t(sub * (CP / 8) + i));
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[i].x), "=r"(r[i].y), "=r"(r[i].z), "=r"(r[i].w) : "r"(a));
        }
        tmem_ld_wait();
        if (sub + 1 < 64 / CP) tmem_ld_32x32b<CP>(lane_addr + uint32_t(part * 64 + (sub + 1) * CP), v[(sub + 1) & 1]);   // next chunk in flight
        const uint32_t (&vv)[CP] = v[sub & 1];
#pragma unroll
        for (int i = 0; i < CP / 8; ++i) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(vv[8 * i + j]) + sBias3[col + 8 * i + j];
          const __half2* h = reinterpret_cast<const __half2*>(&r[i]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 r2 = __half22float2(h[j]);
            f[2 * j] += r2.x;
            f[2 * j + 1] += r2.y;
          }
          const uint32_t a = qbase + swizzle_off<128>(uint32_t(row), uint32_t(sub * (CP / 8) + i));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a),
                       "r"(pack_half2(fmaxf(f[0], 0.f), fmaxf(f[1], 0.f))), "r"(pack_half2(fmaxf(f[2], 0.f), fmaxf(f[3], 0.f))),
                       "r"(pack_half2(fmaxf(f[4], 0.f), fmaxf(f[5], 0.f))), "r"(pack_half2(fmaxf(f[6], 0.f), fmaxf(f[7], 0.f)))
                       : "memory");
        }
      }
      tcgen05_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_cready + 8 * part);
      // conv1' of the previous tile: its last UMMA is issued kLag quarters into this tile, after units this warp has finished
      if (it - 1 > done_tile) { epi1(it - 1); done_tile = it - 1; }
    }
    for (int it = done_tile + 1; it < my_tiles; ++it) epi1(it);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace embclip
