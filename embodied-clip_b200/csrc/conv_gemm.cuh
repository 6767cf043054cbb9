// conv_gemm: the one tensor-core kernel of the encoder path.
//
//   C[m, n] = epilogue( sum_k A[m, k] * W[n, k] )      fp16 x fp16 -> fp32 (TMEM) -> fp16 / fp32
//
// * A is an NHWC fp16 activation seen through a 4-D TMA tensor map {C, W, H, N}; a 1x1 conv / linear
//   layer is the degenerate case {K, M, 1, 1}.  A 3x3 (pad 1, stride 1) conv is an implicit GEMM: the
//   M-tile is a box of box_w x box_h x box_n pixels, and each of the 9 taps is the same box loaded at a
//   shifted (w, h) coordinate -- TMA's out-of-bounds zero fill IS the conv padding, nothing is materialised.
// * An optional second A source (2-D) is concatenated along K: this fuses a bottleneck's downsample 1x1
//   conv into its conv3 ( [W3 | Wd] . [t ; x] ), so the identity tensor never touches HBM.
// * "Grouped" mode offsets the K window per N-group (per attention head), which is how AttentionPool2d's
//   per-head contractions run on the same kernel.
// * Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + single-thread tcgen05.mma
//   issuer, warps 2..9 = epilogue (TMEM -> regs -> +bias (+residual) -> ReLU -> fp16 -> swizzled smem -> TMA
//   store).  smem stages are handed over with full/empty mbarriers; the accumulator is double-buffered
//   in TMEM (2 x BN columns) so tile i's epilogue overlaps tile i+1's MMAs.  Persistent grid.
#pragma once
#include "ptx.cuh"

namespace embclip {

struct ConvGemmParams {
  int num_m_blks, num_n_blks;
  int tiles_w, tiles_h;         // M-tile grid inside one image group (2-D mode: tiles_w = num_m_blks, tiles_h = 1)
  int box_w, box_h, box_n;      // pixels per M-tile = box_w * box_h * box_n <= 128 (2-D mode: 128, 1, 1)
  int taps;                     // 1 or 9
  int kb_per_tap;               // Cin / BK of source 0
  int kb_src0;                  // k-blocks read from A0 (= taps * kb_per_tap); the rest come from A1
  int kb_total;
  uint32_t a0_box_bytes;        // bytes one A0 box load delivers (box may hold fewer than 128 rows)
  int relu;                     // epilogue activation: 0 none, 1 ReLU, 2 QuickGELU x * sigmoid(1.702 x)
  int res_mode;                 // kRes only. 0: out += residual;  1: out = residual > 0 ? out : 0  (ReLU backward mask)
  int out_f32;                  // 1: epilogue stores fp32 rows straight to `out_f32_ptr` (2-D mode only)
  int M, N;                     // logical GEMM extents (rows valid for residual / fp32 stores)
  const float* bias;            // [N] or nullptr
  float* out_f32_ptr;           // [M, ldo]
  const float* res_f32_ptr;     // out_f32 mode: optional fp32 residual [M, ldo] added before the store (may alias out_f32_ptr)
  int ldo;
  // grouped mode (0 = off): N-group g = n0 / grp_n reads A at k + g*grp_a_koff, W at k + g*grp_b_koff,
  // and W rows n0 % grp_b_nmod (if grp_b_nmod != 0)
  int grp_n, grp_a_koff, grp_b_koff, grp_b_nmod;
  int reverse;                  // 1: walk the tiles last-to-first (snake order: start where the previous layer's output is still in L2)
};

template <int BN, int BK, bool kRes = false>
struct ConvGemmCfg {
  static constexpr int BM = 128;
  static constexpr int kSwz = BK * 2;                          // operand swizzle span (bytes)
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kCS = BN < 64 ? BN : 64;                // channels per C staging chunk / TMA store box
  static constexpr int kCSwz = kCS * 2;
  static constexpr int kCChunkBytes = BM * kCS * 2;
  static constexpr int kCBytes = BM * BN * 2;
  // two staging buffers: tile i's TMA store drains while tile i+1's epilogue fills the other one, and
  // (kRes) residual tiles are prefetched one tile ahead.  BN = 256 only has room for one.
  static constexpr int kCBufs = (kRes || BN <= 128) ? 2 : 1;
  static constexpr int kEpiWarps = 8;                          // two warps per TMEM lane quarter, half the columns each
  static constexpr int kEpiThreads = kEpiWarps * 32;
  // kRes: one more warp, whose lane 0 moves the residual chunks in and the output chunks out (see the kernel)
  static constexpr int kThreads = 64 + kEpiThreads + (kRes ? 32 : 0);
  static constexpr int kColsPerWarp = BN / 2;
  static constexpr int kChunk = kColsPerWarp < 32 ? kColsPerWarp : 32;   // columns per tcgen05.ld
  // kRes: a ring of 4 staging CHUNKS (128 rows x kCS columns) instead of two whole-tile buffers
  static constexpr int kResBufs = 4;
  static constexpr int kChunksPerTile = BN / kCS;
  // warps writing into one staging chunk: the 4 TMEM lane quarters x the column halves that fall inside the chunk
  static constexpr int kWarpsPerChunk = 4 * (kColsPerWarp >= kCS ? 1 : kCS / kColsPerWarp);
  static constexpr int kCTotal = kRes ? kResBufs * kCChunkBytes : kCBufs * kCBytes;
  // (kRes, BN = 256: 48-KB stages -- three of them need the larger budget; K = 256 layers are bound by the L2 -> smem fill, and
  //  a 128 x 256 tile moves 20 % fewer operand bytes per FLOP than two 128 x 128 tiles)
  static constexpr int kBudget = (kRes && BN == 256) ? 214 * 1024 : 200 * 1024;
  static constexpr int kStagesRaw = (kBudget - kCTotal) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int kBarBytes = 256;                        // mbarriers + tmem ptr
  static constexpr int kBiasBytes = kRes ? kEpiWarps * kColsPerWarp * 4 : BN * 4;   // kRes: a private slice per epilogue warp
  static constexpr size_t kSmemBytes = 1024 /*align slack*/ + size_t(kStages) * kStageBytes + kCTotal + kBiasBytes + kBarBytes;
  static_assert(kStages >= 2, "pipeline needs at least two stages");
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M=128");
  static_assert(kABytes % 1024 == 0 && kBBytes % 1024 == 0, "stage tiles must keep 1024-B alignment");
};

template <int BN, int BK, bool kRes>
__global__ void __launch_bounds__(64 + 9 * 32, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                 const __grid_constant__ CUtensorMap tmR, const ConvGemmParams p) {
  using Cfg = ConvGemmCfg<BN, BK, kRes>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + S * Cfg::kABytes;
  const uint32_t sC = sB + S * Cfg::kBBytes;
  const uint32_t sBias = sC + Cfg::kCTotal;
  const uint32_t sBar = sBias + Cfg::kBiasBytes;
  const uint32_t bar_full = sBar;                 // S x 8 B
  const uint32_t bar_empty = sBar + 8 * S;        // S x 8 B
  const uint32_t bar_tfull = sBar + 16 * S;       // 2 x 8 B
  const uint32_t bar_tempty = bar_tfull + 16;     // 2 x 8 B
  const uint32_t bar_res = bar_tempty + 16;       // 4 x 8 B (residual landed in staging buffer / chunk i)
  const uint32_t bar_cready = bar_res + 32;       // 4 x 8 B (kRes: output chunk i written by its epilogue warps)
  const uint32_t tmem_slot = bar_cready + 32;     // 4 B
  static_assert(16 * S + 32 + 64 + 4 <= Cfg::kBarBytes, "barrier block");
  uint8_t* const gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  float* const sBias_ptr = reinterpret_cast<float*>(gen_base + (sBias - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_blks * p.num_n_blks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    if (p.kb_src0 < p.kb_total) tma_prefetch_desc(&tmA1);
    if (kRes) tma_prefetch_desc(&tmR);
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, Cfg::kEpiWarps);   // one arrive per epilogue warp
    }
    for (int a = 0; a < 4; ++a) {
      mbar_init(bar_res + 8 * a, 1);
      mbar_init(bar_cready + 8 * a, Cfg::kWarpsPerChunk);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  griddep_launch_dependents();
  griddep_wait();                                  // the previous kernel's outputs (our A / residual) are complete below here

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int tt = p.reverse ? num_tiles - 1 - t : t;
        const int n_blk = tt % p.num_n_blks;
        const int m_blk = tt / p.num_n_blks;
        const int tw = m_blk % p.tiles_w;
        const int th = (m_blk / p.tiles_w) % p.tiles_h;
        const int ng = m_blk / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.box_w, h0 = th * p.box_h, i0 = ng * p.box_n;
        const int n0 = n_blk * BN;
        int a_koff = 0, b_koff = 0, b_row = n0;
        if (p.grp_n) {
          const int g = n0 / p.grp_n;
          a_koff = g * p.grp_a_koff;
          b_koff = g * p.grp_b_koff;
          if (p.grp_b_nmod) b_row = n0 % p.grp_b_nmod;
        }
        for (int kb = 0; kb < p.kb_total; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
          const uint32_t full = bar_full + 8 * stage;
          if (kb < p.kb_src0) {
            mbar_arrive_expect_tx(full, p.a0_box_bytes + Cfg::kBBytes);
            const int tap = kb / p.kb_per_tap;
            const int c0 = (kb - tap * p.kb_per_tap) * BK;
            int dw = 0, dh = 0;
            if (p.taps == 9) { dh = tap / 3 - 1; dw = tap % 3 - 1; }
            tma_load_4d(&tmA0, full, sA + stage * Cfg::kABytes, c0 + a_koff, w0 + dw, h0 + dh, i0);
          } else {
            mbar_arrive_expect_tx(full, Cfg::kABytes + Cfg::kBBytes);
            tma_load_4d(&tmA1, full, sA + stage * Cfg::kABytes, (kb - p.kb_src0) * BK, w0, 0, 0);
          }
          tma_load_2d(&tmB, full, sB + stage * Cfg::kBBytes, kb * BK + b_koff, b_row);
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    // warp-uniform loop, one elected lane issues; descriptors are stepped with 32-bit adds (see ptx.cuh)
    constexpr uint32_t idesc = make_idesc_f16_f32(128, BN);
    constexpr uint32_t dhi = kmajor_desc_hi<Cfg::kSwz>();
    const uint32_t sA_lo = kmajor_desc_lo(sA), sB_lo = kmajor_desc_lo(sB);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1u);       // epilogue has drained this accumulator
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + uint32_t(acc * BN);
      // (one k-block per issue round: these GEMMs are bound by the L2 -> smem fill, so a stage is consumed -- and
      // released -- as soon as it lands; polling two stages per round measured 10 % slower)
      for (int kb = 0; kb < p.kb_total; ++kb) {
        mbar_wait(bar_full + 8 * stage, phase);
        tcgen05_fence_after();
        const uint32_t a_lo = sA_lo + uint32_t(stage) * (Cfg::kABytes / 16);
        const uint32_t b_lo = sB_lo + uint32_t(stage) * (Cfg::kBBytes / 16);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)                  // 16 elements (32 B) along K inside the swizzle span: +2
            umma_f16_ss(d_tmem, desc64(a_lo + 2 * k, dhi), desc64(b_lo + 2 * k, dhi), idesc, k == 0 ? uint32_t(kb != 0) : 1u);
          umma_commit(bar_empty + 8 * stage);                // smem stage free once these MMAs retire
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
      if (elect_one()) umma_commit(bar_tfull + 8 * acc);     // accumulator complete
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else if (kRes && warp == 10) {
    // ============================ kRes: chunk mover ============================
    // The epilogue works on CHUNKS of 128 rows x kCS columns that live in a ring of 4 staging buffers: the residual chunk
    // lands there by TMA, the epilogue warps rewrite it in place with the output, this thread stores it and -- as soon as
    // the store has read the buffer -- fetches the residual of the chunk four steps ahead.  (The first version fetched one
    // whole tile ahead, after draining the previous store: ncu showed DRAM 41 %, L2 31 %, tensor 25 % -- a latency chain.)
    if (lane == 0) {
      constexpr int NC = Cfg::kChunksPerTile;
      const int my_tiles = int(blockIdx.x) < num_tiles ? (num_tiles - 1 - int(blockIdx.x)) / int(gridDim.x) + 1 : 0;
      const int steps = my_tiles * NC;
      auto coords = [&](int s_, int& n_col, int& m0) {
        const int t_ = int(blockIdx.x) + (s_ / NC) * int(gridDim.x);
        const int tt = p.reverse ? num_tiles - 1 - t_ : t_;
        n_col = (tt % p.num_n_blks) * BN + (s_ % NC) * Cfg::kCS;
        m0 = (tt / p.num_n_blks) * 128;
      };
      auto load_res = [&](int s_) {
        int n_col, m0;
        coords(s_, n_col, m0);
        const uint32_t bar = bar_res + 8 * (s_ & 3);
        mbar_arrive_expect_tx(bar, Cfg::kCChunkBytes);
        tma_load_4d(&tmR, bar, sC + uint32_t(s_ & 3) * Cfg::kCChunkBytes, n_col, m0, 0, 0);
      };
      for (int s_ = 0; s_ < steps && s_ < 4; ++s_) load_res(s_);
      for (int s_ = 0; s_ < steps; ++s_) {
        mbar_wait(bar_cready + 8 * (s_ & 3), uint32_t(s_ >> 2) & 1u);
        if (!p.out_f32) {
          int n_col, m0;
          coords(s_, n_col, m0);
          tma_store_4d(&tmC, sC + uint32_t(s_ & 3) * Cfg::kCChunkBytes, n_col, m0, 0, 0);
          tma_store_commit();
        }
        if (s_ + 4 < steps) {
          if (!p.out_f32) tma_store_wait_read0();
          load_res(s_ + 4);
        }
      }
      tma_store_wait_all0();
    }
    __syncwarp();
  } else if (kRes) {
    // ============================ kRes epilogue (warps 2..9) ============================
    if (warp < 10) {
      constexpr int NC = Cfg::kChunksPerTile;
      constexpr int CH = Cfg::kChunk, NP = CH / 8;
      const int q = warp & 3;
      const int row = q * 32 + lane;
      const int wi = warp - 2;                                   // 0..7
      const int col_base = (wi >> 2) * Cfg::kColsPerWarp;        // this warp's half of the tile's columns
      // the warp's columns as segments that each lie inside ONE staging chunk (BN = 256: two 64-column chunks per warp)
      constexpr int SEG = Cfg::kColsPerWarp < Cfg::kCS ? Cfg::kColsPerWarp : Cfg::kCS;
      constexpr int NSEG = Cfg::kColsPerWarp / SEG;
      float* const bias_w = sBias_ptr + wi * Cfg::kColsPerWarp;  // private bias slice
      int acc = 0;
      uint32_t acc_phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int tt = p.reverse ? num_tiles - 1 - t : t;
        const int n0 = (tt % p.num_n_blks) * BN;
        const int grow = (tt / p.num_n_blks) * 128 + row;
        const bool row_ok = grow < p.M;
        for (int i = lane; i < Cfg::kColsPerWarp; i += 32) bias_w[i] = p.bias ? __ldg(p.bias + n0 + col_base + i) : 0.f;
        __syncwarp();
        mbar_wait(bar_tfull + 8 * acc, acc_phase);
        tcgen05_fence_after();
#pragma unroll 1
        for (int sg = 0; sg < NSEG; ++sg) {
          const int seg_col = col_base + sg * SEG;
          const int step = it * NC + seg_col / Cfg::kCS;           // position of this chunk in the ring of 4
          const uint32_t sCt = sC + uint32_t(step & 3) * Cfg::kCChunkBytes;
          mbar_wait(bar_res + 8 * (step & 3), uint32_t(step >> 2) & 1u);
#pragma unroll 1
          for (int c = 0; c < SEG / CH; ++c) {
            const int col = seg_col + c * CH;
            uint32_t v[CH];
            tmem_ld_32x32b<CH>(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * BN + col), v);
            const int piece0 = (col % Cfg::kCS) / 8;
            uint4 rres[NP];
#pragma unroll
            for (int i = 0; i < NP; ++i) {
              const uint32_t a = sCt + swizzle_off<Cfg::kCSwz>(uint32_t(row), uint32_t(piece0 + i));
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rres[i].x), "=r"(rres[i].y), "=r"(rres[i].z), "=r"(rres[i].w) : "r"(a));
            }
            tmem_ld_wait();
            float f[CH];
#pragma unroll
            for (int i = 0; i < CH; ++i) f[i] = __uint_as_float(v[i]) + bias_w[col - col_base + i];
#pragma unroll
            for (int i = 0; i < NP; ++i) {
              const __half2* h = reinterpret_cast<const __half2*>(&rres[i]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 r2 = __half22float2(h[j]);
                if (p.res_mode == 0) {
                  f[i * 8 + j * 2] += r2.x;
                  f[i * 8 + j * 2 + 1] += r2.y;
                } else {
                  f[i * 8 + j * 2] = r2.x > 0.f ? f[i * 8 + j * 2] : 0.f;
                  f[i * 8 + j * 2 + 1] = r2.y > 0.f ? f[i * 8 + j * 2 + 1] : 0.f;
                }
              }
            }
            if (p.relu == 1) {
#pragma unroll
              for (int i = 0; i < CH; ++i) f[i] = fmaxf(f[i], 0.f);
            } else if (p.relu == 2) {
#pragma unroll
              for (int i = 0; i < CH; ++i) f[i] = __fdividef(f[i], 1.f + __expf(-1.702f * f[i]));
            }
            if (p.out_f32) {
              if (row_ok) {
                float4* op = reinterpret_cast<float4*>(p.out_f32_ptr + size_t(grow) * p.ldo + n0 + col);
#pragma unroll
                for (int i = 0; i < CH / 4; ++i) op[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < NP; ++i) {
                const uint32_t a = sCt + swizzle_off<Cfg::kCSwz>(uint32_t(row), uint32_t(piece0 + i));
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a),
                             "r"(pack_half2(f[8 * i], f[8 * i + 1])), "r"(pack_half2(f[8 * i + 2], f[8 * i + 3])),
                             "r"(pack_half2(f[8 * i + 4], f[8 * i + 5])), "r"(pack_half2(f[8 * i + 6], f[8 * i + 7]))
                             : "memory");
              }
            }
          }
          if (sg == NSEG - 1) tcgen05_fence_before();              // (last TMEM read of this tile is behind us)
          if (!p.out_f32) fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (sg == NSEG - 1) mbar_arrive(bar_tempty + 8 * acc);   // accumulator columns of this warp drained
            mbar_arrive(bar_cready + 8 * (step & 3));              // its share of the chunk is in the staging buffer
          }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ============================ epilogue (warps 2..9) ============================
    const int q = warp & 3;                                    // TMEM lane quarter this warp may touch
    const int row = q * 32 + lane;                             // row of the 128-row tile
    const int col_base = ((warp - 2) >> 2) * Cfg::kColsPerWarp;   // this warp's half of the tile's columns
    constexpr int CH = Cfg::kChunk;                            // 32 or 16 columns per TMEM load
    constexpr int NP = CH / 8;                                 // 16-B pieces per chunk
    const int epi_tid = threadIdx.x - 64;
    const bool store_leader = (threadIdx.x == 64);
    int acc = 0;
    uint32_t acc_phase = 0;
    int cbuf = 0;                                              // staging buffer of this tile (kRes: alternates)
    uint32_t res_phase = 0;
    // residual tile of tile `tt` -> staging buffer `buf` (TMA, one tile ahead of its use)
    auto prefetch_residual = [&](int t_, int buf) {
      const int tt = p.reverse ? num_tiles - 1 - t_ : t_;
      const int nb = tt % p.num_n_blks, mb = tt / p.num_n_blks;
      const uint32_t bar = bar_res + 8 * buf;
      mbar_arrive_expect_tx(bar, Cfg::kCBytes);
#pragma unroll
      for (int cc = 0; cc < BN / Cfg::kCS; ++cc)
        tma_load_4d(&tmR, bar, sC + buf * Cfg::kCBytes + cc * Cfg::kCChunkBytes, nb * BN + cc * Cfg::kCS, mb * 128, 0, 0);
    };
    if (kRes && store_leader && int(blockIdx.x) < num_tiles) prefetch_residual(blockIdx.x, 0);
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int tt = p.reverse ? num_tiles - 1 - t : t;
      const int n_blk = tt % p.num_n_blks;
      const int m_blk = tt / p.num_n_blks;
      const int tw = m_blk % p.tiles_w;
      const int th = (m_blk / p.tiles_w) % p.tiles_h;
      const int ng = m_blk / (p.tiles_w * p.tiles_h);
      const int w0 = tw * p.box_w, h0 = th * p.box_h, i0 = ng * p.box_n;
      const int n0 = n_blk * BN;
      const int grow = w0 + row;                               // global row (2-D mode)
      const bool row_ok = grow < p.M;
      const uint32_t sCt = sC + uint32_t(cbuf) * Cfg::kCBytes;

      // staging smem + bias tile are reused: the TMA store that last read them must be done reading
      if (store_leader) {
        if (kRes) {
          // buffer cbuf^1 (tile t-1's output) must be drained before tile t+1's residual lands in it
          tma_store_wait_read0();
          if (t + int(gridDim.x) < num_tiles) prefetch_residual(t + gridDim.x, cbuf ^ 1);
        } else if (Cfg::kCBufs == 2) {
          tma_store_wait_read1();                              // only tile t-2's store (same buffer) must be done
        } else {
          tma_store_wait_read0();
        }
      }
      for (int i = epi_tid; i < BN; i += Cfg::kEpiThreads) sBias_ptr[i] = p.bias ? __ldg(p.bias + n0 + i) : 0.f;
      named_bar_sync(1, Cfg::kEpiThreads);

      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tcgen05_fence_after();
      if (kRes) mbar_wait(bar_res + 8 * cbuf, res_phase);
#pragma unroll 1
      for (int c = 0; c < Cfg::kColsPerWarp / CH; ++c) {
        const int col = col_base + c * CH;
        uint32_t v[CH];
        tmem_ld_32x32b<CH>(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * BN + col), v);
        const uint32_t chunk_base = sCt + uint32_t(col / Cfg::kCS) * Cfg::kCChunkBytes;
        const int piece0 = (col % Cfg::kCS) / 8;
        uint4 rres[NP];
        if (kRes) {
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            const uint32_t a = chunk_base + swizzle_off<Cfg::kCSwz>(uint32_t(row), uint32_t(piece0 + i));
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rres[i].x), "=r"(rres[i].y), "=r"(rres[i].z), "=r"(rres[i].w) : "r"(a));
          }
        }
        tmem_ld_wait();
        float f[CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) f[i] = __uint_as_float(v[i]) + sBias_ptr[col + i];
        if (kRes) {
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            const __half2* h = reinterpret_cast<const __half2*>(&rres[i]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 r2 = __half22float2(h[j]);
              if (p.res_mode == 0) {
                f[i * 8 + j * 2] += r2.x;
                f[i * 8 + j * 2 + 1] += r2.y;
              } else {
                f[i * 8 + j * 2] = r2.x > 0.f ? f[i * 8 + j * 2] : 0.f;
                f[i * 8 + j * 2 + 1] = r2.y > 0.f ? f[i * 8 + j * 2 + 1] : 0.f;
              }
            }
          }
        }
        if (p.relu == 1) {
#pragma unroll
          for (int i = 0; i < CH; ++i) f[i] = fmaxf(f[i], 0.f);
        } else if (p.relu == 2) {
#pragma unroll
          for (int i = 0; i < CH; ++i) f[i] = __fdividef(f[i], 1.f + __expf(-1.702f * f[i]));
        }
        if (p.out_f32) {
          if (row_ok) {
            float4* op = reinterpret_cast<float4*>(p.out_f32_ptr + size_t(grow) * p.ldo + n0 + col);
            if (p.res_f32_ptr) {
              const float4* rp = reinterpret_cast<const float4*>(p.res_f32_ptr + size_t(grow) * p.ldo + n0 + col);
#pragma unroll
              for (int i = 0; i < CH / 4; ++i) {
                const float4 r = rp[i];
                f[4 * i] += r.x; f[4 * i + 1] += r.y; f[4 * i + 2] += r.z; f[4 * i + 3] += r.w;
              }
            }
#pragma unroll
            for (int i = 0; i < CH / 4; ++i) op[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
          }
        } else {
          // chunk of kCS channels = one TMA store box; 16-B pieces land at their swizzled slot
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            const uint32_t a = chunk_base + swizzle_off<Cfg::kCSwz>(uint32_t(row), uint32_t(piece0 + i));
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a),
                         "r"(pack_half2(f[8 * i], f[8 * i + 1])), "r"(pack_half2(f[8 * i + 2], f[8 * i + 3])),
                         "r"(pack_half2(f[8 * i + 4], f[8 * i + 5])), "r"(pack_half2(f[8 * i + 6], f[8 * i + 7]))
                         : "memory");
          }
        }
      }
      // accumulator drained -> MMA warp may overwrite it
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }

      // all epilogue threads are done with sBias / have written their staging rows
      if (!p.out_f32) fence_proxy_async_smem();
      named_bar_sync(1, Cfg::kEpiThreads);
      if (!p.out_f32) {
        if (store_leader) {
#pragma unroll
          for (int cc = 0; cc < BN / Cfg::kCS; ++cc)
            tma_store_4d(&tmC, sCt + cc * Cfg::kCChunkBytes, n0 + cc * Cfg::kCS, w0, h0, i0);
          tma_store_commit();
        }
      }
      if (Cfg::kCBufs == 2) { cbuf ^= 1; if (cbuf == 0) res_phase ^= 1u; }
    }
    if (store_leader) tma_store_wait_all0();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace embclip
