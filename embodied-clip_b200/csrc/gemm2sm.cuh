// gemm2sm: the CTA-pair (tcgen05 cta_group::2) variant of conv_gemm for plain 2-D GEMMs with large K or N
//
//   C[m, n] = epilogue( sum_k [A0 | A1][m, k] * W[n, k] )      fp16 x fp16 -> fp32 (TMEM) -> fp16 / fp32
//
// Why: with one CTA per 128 x 128 tile every k-block moves 32 KB from L2 for 2 MFLOP -- at B200's ~12 TB/s L2 ceiling
// that caps the tensor pipe near 50 % (profiles/r1b: layer3/4 1x1 convs sit at 8-9 TB/s of L2->SM traffic and
// 450-580 TFLOP/s).  A CTA pair computes a 256 x 256 tile as ONE UMMA (M = 256): each CTA stages only its own 128 rows
// of A and HALF of the B tile (128 of the 256 weight rows); the tensor cores of both SMs read both halves.  Per CTA that
// is the same 32 KB per k-block for twice the FLOPs.
//
// Structure (cluster of 2, same warp roles as conv_gemm, 320 threads per CTA):
//   * warp 0 (both CTAs): TMA producer.  Loads use the .cta_group::2 form whose mbarrier operand, with the peer bit
//     cleared, names the LEADER's full barrier, so the leader's barrier counts the bytes of both CTAs.
//   * warp 1: TMEM owner (tcgen05.alloc.cta_group::2 in both CTAs); in the leader it is the single MMA issuer.  Stage
//     release and accumulator-ready are tcgen05.commit ... multicast to the barriers of both CTAs.
//   * warps 2..9 (both CTAs): epilogue of the CTA's own 128 accumulator rows; the peer's warps arrive remotely
//     (mapa + mbarrier.arrive.shared::cluster) on the leader's accumulator-empty barrier.
//   Accumulators are double buffered (2 x 256 TMEM columns), so tile i's epilogue overlaps tile i+1's MMAs.
//   The epilogue walks a tile in 128-column units through two 32 KB staging buffers: unit u's TMA store drains while
//   unit u+1 is computed, and (kRes) unit u+1's residual tile is TMA-prefetched into the other buffer.
#pragma once
#include "ptx.cuh"

namespace embclip {

struct Gemm2Params {
  int num_m_pairs, num_n_blks;  // tiles of 256 rows x BN columns
  int kb_src0, kb_total;        // k-blocks (64) read from A0; the rest from A1
  int relu;                     // 0 none, 1 ReLU, 2 QuickGELU
  int out_f32;
  int M, N;
  const float* bias;
  int res_mode;                 // kRes: 0 add the fp16 residual tile, 1 mask (out = residual > 0 ? out : 0)
  float* out_f32_ptr;           // [M, N]
  const float* res_f32_ptr;     // out_f32: fp32 residual [M, N] (may alias out_f32_ptr)
  int reverse;
};

template <int BN>
struct Gemm2Cfg {
  static constexpr int BK = 64;
  static constexpr int kBHalf = BN / 2;
  static constexpr int kABytes = 128 * BK * 2;
  static constexpr int kBBytes = kBHalf * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kCS = 64;
  static constexpr int kCChunkBytes = 128 * kCS * 2;
  static constexpr int kUnitN = 128;                         // the epilogue walks the tile in 128-column units
  static constexpr int kUnits = BN / kUnitN;
  static constexpr int kCBytes = 128 * kUnitN * 2;           // one staging buffer = one unit (32 KB)
  static constexpr int kCBufs = 2;
  static constexpr int kEpiWarps = 8;
  static constexpr int kEpiThreads = kEpiWarps * 32;
  static constexpr int kThreads = 64 + kEpiThreads;
  static constexpr int kBarBytes = 256;
  static constexpr int kBiasBytes = 128 * 4;
  static constexpr int kBudget = 227 * 1024 - 1024 - kCBufs * kCBytes - kBiasBytes - kBarBytes;
  static constexpr int kStagesRaw = kBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr size_t kSmemBytes = 1024 + size_t(kStages) * kStageBytes + kCBufs * kCBytes + kBiasBytes + kBarBytes;
  static_assert(BN == 128 || BN == 256, "pair tile N");
  static_assert(kStages >= 3, "pipeline depth");
};

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;     // shared::cluster address of the same offset in the pair's even CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_ss_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs of the pair once all MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(uint16_t(3)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(kCols) : "memory");
}

template <int BN, bool kRes>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 8 * 32, 1)
gemm2sm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
               const __grid_constant__ CUtensorMap tmR, const Gemm2Params p) {
  using Cfg = Gemm2Cfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + S * Cfg::kABytes;
  const uint32_t sC = sB + S * Cfg::kBBytes;
  const uint32_t sBias = sC + Cfg::kCBufs * Cfg::kCBytes;
  const uint32_t sBar = sBias + Cfg::kBiasBytes;
  const uint32_t bar_full = sBar;                 // S x 8 B   (used in the leader)
  const uint32_t bar_empty = sBar + 8 * S;        // S x 8 B   (per CTA)
  const uint32_t bar_tfull = sBar + 16 * S;       // 2 x 8 B   (per CTA)
  const uint32_t bar_tempty = bar_tfull + 16;     // 2 x 8 B   (used in the leader)
  const uint32_t bar_res = bar_tempty + 16;       // 2 x 8 B   (per CTA: residual unit landed in staging buffer i)
  const uint32_t tmem_slot = bar_res + 16;
  uint8_t* const gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  float* const sBias_ptr = reinterpret_cast<float*>(gen_base + (sBias - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();        // 0 = leader
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_tiles = p.num_m_pairs * p.num_n_blks;

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
      mbar_init(bar_tempty + 8 * a, 2 * Cfg::kEpiWarps);   // epilogue warps of both CTAs
      mbar_init(bar_res + 8 * a, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm<Cfg::kTmemCols>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                              // the peer's barriers exist before anything signals them
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  griddep_launch_dependents();
  griddep_wait();

  if (warp == 0) {
    // ============================ TMA producer (both CTAs) ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        const int tt = p.reverse ? num_tiles - 1 - t : t;
        const int n_blk = tt % p.num_n_blks;
        const int m_pair = tt / p.num_n_blks;
        const int m0 = m_pair * 256 + int(rank) * 128;
        const int nrow0 = n_blk * BN + int(rank) * Cfg::kBHalf;
        for (int kb = 0; kb < p.kb_total; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
          const uint32_t full = bar_full + 8 * stage;
          if (rank == 0) mbar_arrive_expect_tx(full, 2u * Cfg::kStageBytes);   // bytes of both CTAs land on the leader's barrier
          if (kb < p.kb_src0) tma_load_4d_2sm(&tmA0, full, sA + stage * Cfg::kABytes, kb * 64, m0, 0, 0);
          else tma_load_4d_2sm(&tmA1, full, sA + stage * Cfg::kABytes, (kb - p.kb_src0) * 64, m0, 0, 0);
          tma_load_2d_2sm(&tmB, full, sB + stage * Cfg::kBBytes, kb * 64, nrow0);
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================ MMA issuer (leader CTA only) ============================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16_f32(256, BN);
      constexpr uint32_t dhi = kmajor_desc_hi<128>();
      const uint32_t sA_lo = kmajor_desc_lo(sA), sB_lo = kmajor_desc_lo(sB);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1u);       // both CTAs' epilogues have drained this accumulator
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * BN);
        for (int kb = 0; kb < p.kb_total; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tcgen05_fence_after();
          const uint32_t a_lo = sA_lo + uint32_t(stage) * (Cfg::kABytes / 16);
          const uint32_t b_lo = sB_lo + uint32_t(stage) * (Cfg::kBBytes / 16);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ss_2sm(d_tmem, desc64(a_lo + 2 * k, dhi), desc64(b_lo + 2 * k, dhi), idesc, k == 0 ? uint32_t(kb != 0) : 1u);
            umma_commit_2sm(bar_empty + 8 * stage);            // frees the stage in both CTAs
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) umma_commit_2sm(bar_tfull + 8 * acc);
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ============================ epilogue (warps 2..9, both CTAs: own 128 rows) ============================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int col_base = ((warp - 2) >> 2) * 64;               // this warp's half of a 128-column unit
    constexpr int CH = 32, NP = 4, U = Cfg::kUnits;
    const int epi_tid = threadIdx.x - 64;
    const bool store_leader = (threadIdx.x == 64);
    int acc = 0, cbuf = 0;
    uint32_t acc_phase = 0, res_phase = 0;
    // residual unit (tile t_, unit hh) -> staging buffer `buf` (TMA, one unit ahead of its use)
    auto prefetch_residual = [&](int t_, int hh, int buf) {
      const int tt = p.reverse ? num_tiles - 1 - t_ : t_;
      const int nb = tt % p.num_n_blks, mp = tt / p.num_n_blks;
      const uint32_t bar = bar_res + 8 * buf;
      mbar_arrive_expect_tx(bar, Cfg::kCBytes);
#pragma unroll
      for (int cc = 0; cc < 2; ++cc)
        tma_load_4d(&tmR, bar, sC + buf * Cfg::kCBytes + cc * Cfg::kCChunkBytes, nb * BN + hh * 128 + cc * 64, mp * 256 + int(rank) * 128, 0, 0);
    };
    if (kRes && store_leader && pair < num_tiles) prefetch_residual(pair, 0, 0);
    for (int t = pair; t < num_tiles; t += num_pairs) {
      const int tt = p.reverse ? num_tiles - 1 - t : t;
      const int n_blk = tt % p.num_n_blks;
      const int m_pair = tt / p.num_n_blks;
      const int m0 = m_pair * 256 + int(rank) * 128;
      const int grow = m0 + row;
      const bool row_ok = grow < p.M;
#pragma unroll 1
      for (int hh = 0; hh < U; ++hh) {
        const int n0 = n_blk * BN + hh * 128;
        const uint32_t sCt = sC + uint32_t(cbuf) * Cfg::kCBytes;
        if (store_leader) {
          if (kRes) {
            // buffer cbuf^1 (the previous unit's output) must be drained before the next unit's residual lands in it
            tma_store_wait_read0();
            if (hh + 1 < U) prefetch_residual(t, hh + 1, cbuf ^ 1);
            else if (t + num_pairs < num_tiles) prefetch_residual(t + num_pairs, 0, cbuf ^ 1);
          } else {
            tma_store_wait_read1();                            // only the store from two units ago used this buffer
          }
        }
        for (int i = epi_tid; i < 128; i += Cfg::kEpiThreads) sBias_ptr[i] = p.bias ? __ldg(p.bias + n0 + i) : 0.f;
        named_bar_sync(1, Cfg::kEpiThreads);
        if (hh == 0) {
          mbar_wait(bar_tfull + 8 * acc, acc_phase);
          tcgen05_fence_after();
        }
        if (kRes) mbar_wait(bar_res + 8 * cbuf, res_phase);
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          const int col = col_base + c * CH;                   // column inside the unit
          uint32_t v[CH];
          tmem_ld_32x32b<CH>(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * BN + hh * 128 + col), v);
          const uint32_t chunk_base = sCt + uint32_t(col / 64) * Cfg::kCChunkBytes;
          const int piece0 = (col % 64) / 8;
          uint4 rres[NP];
          if (kRes) {
#pragma unroll
            for (int i = 0; i < NP; ++i) {
              const uint32_t a = chunk_base + swizzle_off<128>(uint32_t(row), uint32_t(piece0 + i));
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
            // fp32 rows (+ fp32 residual): a thread owns one ROW of the 32 x 32 block, so direct global access would touch 32
            // cache lines per instruction (the first version: the ViT out / c_proj GEMMs ran at 300 / 950 TFLOP/s, LSU-bound).
            // Go through a per-warp scratch block in the (otherwise unused) staging buffers instead: global <-> scratch is
            // coalesced (lane = (row % 4, 16-B piece): 4 full lines per instruction), scratch <-> registers is row-wise.
            static_assert(CH == 32, "fp32 epilogue block is 32 x 32");
            constexpr uint32_t kPitch = 144;                           // 128 B + 16 B pad: conflict-free for 8-lane phases
            const uint32_t scr = sC + uint32_t(warp - 2) * (32u * kPitch);
            const int sub_r = lane >> 3, piece = lane & 7;
            const size_t gbase = size_t(m0 + q * 32) * p.N + n0 + col;
            if (p.res_f32_ptr) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const int r = 4 * k + sub_r;
                float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m0 + q * 32 + r < p.M) v4 = *(reinterpret_cast<const float4*>(p.res_f32_ptr + gbase + size_t(r) * p.N) + piece);
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(scr + uint32_t(r) * kPitch + uint32_t(piece) * 16u),
                             "f"(v4.x), "f"(v4.y), "f"(v4.z), "f"(v4.w) : "memory");
              }
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float4 r4;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r4.x), "=f"(r4.y), "=f"(r4.z), "=f"(r4.w)
                             : "r"(scr + uint32_t(lane) * kPitch + uint32_t(i) * 16u));
                f[4 * i] += r4.x; f[4 * i + 1] += r4.y; f[4 * i + 2] += r4.z; f[4 * i + 3] += r4.w;
              }
              __syncwarp();
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(scr + uint32_t(lane) * kPitch + uint32_t(i) * 16u),
                           "f"(f[4 * i]), "f"(f[4 * i + 1]), "f"(f[4 * i + 2]), "f"(f[4 * i + 3]) : "memory");
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int r = 4 * k + sub_r;
              float4 o4;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o4.x), "=f"(o4.y), "=f"(o4.z), "=f"(o4.w)
                           : "r"(scr + uint32_t(r) * kPitch + uint32_t(piece) * 16u));
              if (m0 + q * 32 + r < p.M) *(reinterpret_cast<float4*>(p.out_f32_ptr + gbase + size_t(r) * p.N) + piece) = o4;
            }
            __syncwarp();
          } else {
#pragma unroll
            for (int i = 0; i < NP; ++i) {
              const uint32_t a = chunk_base + swizzle_off<128>(uint32_t(row), uint32_t(piece0 + i));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a),
                           "r"(pack_half2(f[8 * i], f[8 * i + 1])), "r"(pack_half2(f[8 * i + 2], f[8 * i + 3])),
                           "r"(pack_half2(f[8 * i + 4], f[8 * i + 5])), "r"(pack_half2(f[8 * i + 6], f[8 * i + 7]))
                           : "memory");
            }
          }
        }
        if (hh == U - 1) {
          // accumulator drained -> the leader's MMA warp may overwrite it
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (rank == 0) mbar_arrive(bar_tempty + 8 * acc);
            else mbar_arrive_cluster(bar_tempty + 8 * acc, 0);
          }
          if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (!p.out_f32) fence_proxy_async_smem();
        named_bar_sync(1, Cfg::kEpiThreads);
        if (!p.out_f32 && store_leader) {
#pragma unroll
          for (int cc = 0; cc < 2; ++cc)
            tma_store_4d(&tmC, sCt + cc * Cfg::kCChunkBytes, n0 + cc * 64, m0, 0, 0);
          tma_store_commit();
        }
        cbuf ^= 1;
        if (cbuf == 0) res_phase ^= 1u;
      }
    }
    if (store_leader) tma_store_wait_all0();
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                              // neither CTA frees TMEM / exits while the peer may still signal it
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_2sm<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace embclip
