// Non-GEMM kernels of the encoder path: HBM-bound vectorised / warp-shuffle work.
//   stem_conv1_kernel     3->C 3x3 stride-2 conv + folded BN + ReLU straight from the fp32 NHWC frames
//   avgpool2_kernel       nn.AvgPool2d(2), NHWC fp16
//   attnpool_tokens       mean token + positional embedding  -> fp16 tokens [B, HW+1, C]
//   attnpool_core         single-query softmax attention over HW+1 keys, per head, warp-shuffle
//   avg_head / nhwc_to_nchw   the two trivial output heads
#pragma once
#include "ptx.cuh"

namespace embclip {

// ------------------------------------------------------------------------------------------------
// Stem conv1 (clip/model.py ModifiedResNet.conv1+bn1+relu): K = 27 is too thin for the tensor pipe and the
// layer is bound by the fp32 frame read, so it runs on CUDA cores: one thread = one output pixel x COUT
// channels, weights broadcast from shared memory.
// ------------------------------------------------------------------------------------------------
// One thread = TWO horizontally adjacent output pixels x COUT channels: every weight vector fetched from shared
// memory (a broadcast LDS.128) feeds 24 FMAs instead of 12 -- the first version (one pixel per thread) spent as many
// issue slots on weight loads as on FMAs.  The two pixels share the middle input column (5 x 3 input pixels, not 6 x 3).
// TIn = float: frames already mean/std normalised (what the AllenAct sensor hands over).  TIn = uint8_t: raw RGB bytes;
// (v / 255 - mean) / std is applied on load as one FMA (norm = {scale_r, scale_g, scale_b, offset_r, offset_g, offset_b}),
// so the host never touches the pixels and the H2D copy is 4x smaller.
struct StemNorm { float scale[3], offset[3]; };
template <int COUT, typename TIn>
__global__ void __launch_bounds__(128)
stem_conv1_kernel(const TIn* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                  __half* __restrict__ y, int B, int R, const StemNorm norm) {
  __shared__ float sw[27 * COUT];
  __shared__ float sb[COUT];
  for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) sb[i] = bias[i];
  __syncthreads();
  griddep_wait();                                  // (weights above are constants; the output buffer may still be read by the previous forward)
  const int Ro = R / 2, Rp = Ro / 2;               // output resolution, pixel pairs per output row
  const long long total = (long long)B * Ro * Rp;
  const long long pair = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= total) return;
  const int op = int(pair % Rp);
  const int oh = int((pair / Rp) % Ro);
  const int b = int(pair / ((long long)Rp * Ro));
  const int ow = 2 * op;
  float acc0[COUT], acc1[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) { acc0[c] = sb[c]; acc1[c] = sb[c]; }
  const TIn* xb = x + (size_t)b * R * R * 3;
  constexpr bool kRaw = sizeof(TIn) == 1;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int ih = 2 * oh - 1 + kh;
    if (ih < 0 || ih >= R) continue;
    // input columns 2*ow-1 .. 2*ow+3 (5 pixels x 3 channels); column -1 is the zero pad (zero AFTER normalisation)
    float in[5][3];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int iw = 2 * ow - 1 + j;
      const bool ok = iw >= 0 && iw < R;
      const TIn* px = xb + ((size_t)ih * R + (ok ? iw : 0)) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float v = float(__ldg(px + c));
        if (kRaw) v = v * norm.scale[c] + norm.offset[c];
        in[j][c] = ok ? v : 0.f;
      }
    }
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const float4* w0 = reinterpret_cast<const float4*>(sw + ((kh * 3 + kw) * 3 + 0) * COUT);
      const float4* w1 = reinterpret_cast<const float4*>(sw + ((kh * 3 + kw) * 3 + 1) * COUT);
      const float4* w2 = reinterpret_cast<const float4*>(sw + ((kh * 3 + kw) * 3 + 2) * COUT);
      const float a0 = in[kw][0], a1 = in[kw][1], a2 = in[kw][2];              // pixel ow
      const float b0 = in[kw + 2][0], b1 = in[kw + 2][1], b2 = in[kw + 2][2];  // pixel ow + 1
#pragma unroll
      for (int c4 = 0; c4 < COUT / 4; ++c4) {
        const float4 p = w0[c4], q = w1[c4], r = w2[c4];
        acc0[4 * c4 + 0] += a0 * p.x + a1 * q.x + a2 * r.x;
        acc0[4 * c4 + 1] += a0 * p.y + a1 * q.y + a2 * r.y;
        acc0[4 * c4 + 2] += a0 * p.z + a1 * q.z + a2 * r.z;
        acc0[4 * c4 + 3] += a0 * p.w + a1 * q.w + a2 * r.w;
        acc1[4 * c4 + 0] += b0 * p.x + b1 * q.x + b2 * r.x;
        acc1[4 * c4 + 1] += b0 * p.y + b1 * q.y + b2 * r.y;
        acc1[4 * c4 + 2] += b0 * p.z + b1 * q.z + b2 * r.z;
        acc1[4 * c4 + 3] += b0 * p.w + b1 * q.w + b2 * r.w;
      }
    }
  }
  uint4* out = reinterpret_cast<uint4*>(y + (((size_t)b * Ro + oh) * Ro + ow) * COUT);
#pragma unroll
  for (int i = 0; i < COUT / 8; ++i) {
    uint4 o;
    o.x = pack_half2(fmaxf(acc0[8 * i + 0], 0.f), fmaxf(acc0[8 * i + 1], 0.f));
    o.y = pack_half2(fmaxf(acc0[8 * i + 2], 0.f), fmaxf(acc0[8 * i + 3], 0.f));
    o.z = pack_half2(fmaxf(acc0[8 * i + 4], 0.f), fmaxf(acc0[8 * i + 5], 0.f));
    o.w = pack_half2(fmaxf(acc0[8 * i + 6], 0.f), fmaxf(acc0[8 * i + 7], 0.f));
    out[i] = o;
  }
#pragma unroll
  for (int i = 0; i < COUT / 8; ++i) {
    uint4 o;
    o.x = pack_half2(fmaxf(acc1[8 * i + 0], 0.f), fmaxf(acc1[8 * i + 1], 0.f));
    o.y = pack_half2(fmaxf(acc1[8 * i + 2], 0.f), fmaxf(acc1[8 * i + 3], 0.f));
    o.z = pack_half2(fmaxf(acc1[8 * i + 4], 0.f), fmaxf(acc1[8 * i + 5], 0.f));
    o.w = pack_half2(fmaxf(acc1[8 * i + 6], 0.f), fmaxf(acc1[8 * i + 7], 0.f));
    out[COUT / 8 + i] = o;
  }
}

// ------------------------------------------------------------------------------------------------
// Stem conv1 on the tensor cores.  K = 27 is thin, and fp16 operands would cost first-layer precision, so every fp32 input
// value v and weight w is split into fp16 halves (v = v_hi + v_lo, w = w_hi + w_lo, each exact to 2^-22) and the im2col row
// of an output pixel is laid out as  [v_hi(27) 0(5) | v_lo(27) 0(5) | v_hi(27) 0(5) | 0(32)]  against weight rows
// [w_hi | w_hi | w_lo | 0]:  sum = v_hi w_hi + v_lo w_hi + v_hi w_lo, every product exact in the fp32 accumulator.
// One tile = 128 output pixels: thread <-> pixel gathers its 27 inputs (3 runs of 9 contiguous values; raw uint8 frames are
// normalised here), writes its 256-B row into the two swizzled K-major k-blocks, one thread issues 8 UMMAs (128 x 32 x 16),
// and the same threads read their accumulator row back: bias, ReLU, fp16, one 64-B store per pixel.  ~150 instructions
// per pixel instead of ~1100 on the CUDA cores.  Persistent CTAs (TMEM allocated once), several per SM; the next tile's
// gather is issued before the current tile's MMA is awaited.
// ------------------------------------------------------------------------------------------------
constexpr int kStemTcSmem = 1024 + 2 * 16384 + 2 * 4096 + 64;
template <typename TIn>
__global__ void __launch_bounds__(128)
stem_conv1_tc_kernel(const TIn* __restrict__ x, const __half* __restrict__ wtc, const float* __restrict__ bias,
                     __half* __restrict__ y, int B, int R, const StemNorm norm) {
  constexpr int COUT = 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;                        // 2 k-blocks x [128 rows][128 B]
  const uint32_t sW = base + 32768;                // 2 k-blocks x [32 rows][128 B]
  const uint32_t bar = sW + 8192, tmem_slot = bar + 8;
  uint8_t* const gen = smem_raw + (base - smem_u32(smem_raw));
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr bool kRaw = sizeof(TIn) == 1;
  const bool vec4 = !kRaw && (R % 4 == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0);

  // weights [32][128] fp16 (row n: k-block 0 = cols 0..63, k-block 1 = cols 64..127) -> swizzled K-major tiles
  for (int i = tid; i < 32 * 16; i += 128) {
    const int n = i >> 4, piece = i & 15;          // 16-B pieces of the 256-B row
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(wtc + n * 128) + piece);
    *reinterpret_cast<uint4*>(gen + 32768 + (piece >> 3) * 4096 + swizzle_off<128>(uint32_t(n), uint32_t(piece & 7))) = v;
  }
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<32>(tmem_slot);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  griddep_wait();

  const int Ro = R / 2;
  const long long total = (long long)B * Ro * Ro;
  const long long num_tiles = (total + 127) / 128;
  float bv[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) bv[c] = __ldg(bias + c);

  auto gather = [&](long long pix, float (&v)[27]) {
    if (pix >= total) {
#pragma unroll
      for (int i = 0; i < 27; ++i) v[i] = 0.f;
      return;
    }
    const int ow = int(pix % Ro);
    const int oh = int((pix / Ro) % Ro);
    const int b = int(pix / ((long long)Ro * Ro));
    const TIn* xb = x + (size_t)b * R * R * 3;
    if constexpr (!kRaw) {
      if (vec4) {
        // fp32 frames: the 9 values of a kernel row are 36 contiguous bytes starting 12 B past a 24-B pixel-pair boundary;
        // three aligned 16-B loads cover them (window starts 1 float early for even ow, 3 floats early for odd ow)
        const int odd = ow & 1;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const int ih = 2 * oh - 1 + kh;
          const bool rok = ih >= 0 && ih < R;
          float e[12];
          if (rok) {
            const float4* rp = reinterpret_cast<const float4*>(xb + (size_t)ih * R * 3 + 6 * ow - (odd ? 6 : 4));
            const float4 a = ow > 0 ? __ldg(rp) : make_float4(0.f, 0.f, 0.f, 0.f);      // column -1 is the zero pad
            const float4 b4 = __ldg(rp + 1), c4 = __ldg(rp + 2);
            e[0] = a.x; e[1] = a.y; e[2] = a.z; e[3] = a.w; e[4] = b4.x; e[5] = b4.y; e[6] = b4.z; e[7] = b4.w;
            e[8] = c4.x; e[9] = c4.y; e[10] = c4.z; e[11] = c4.w;
          } else {
#pragma unroll
            for (int i = 0; i < 12; ++i) e[i] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < 9; ++j) v[kh * 9 + j] = odd ? e[3 + j] : e[1 + j];
        }
        return;
      }
    }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = 2 * oh - 1 + kh;
      const bool rok = ih >= 0 && ih < R;
      const TIn* row = xb + ((size_t)(rok ? ih : 0) * R + 2 * ow) * 3 - 3;     // -> input column 2*ow - 1
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const bool ok = rok && (ow > 0 || j >= 3);                             // column -1 is the zero pad
        float t = 0.f;
        if (ok) {
          t = float(__ldg(row + j));
          if (kRaw) t = t * norm.scale[j % 3] + norm.offset[j % 3];
        }
        v[kh * 9 + j] = t;
      }
    }
  };

  float cur[27];
  long long tile = blockIdx.x;
  if (tile < num_tiles) gather(tile * 128 + tid, cur);
  uint32_t phase = 0;
  for (; tile < num_tiles; tile += gridDim.x) {
    // ---- im2col row -> smem (hi / lo split)
    uint32_t hi[16], lo[16];                       // 32 halves each, pairs packed
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float a = 2 * i < 27 ? cur[2 * i] : 0.f, b2 = 2 * i + 1 < 27 ? cur[2 * i + 1] : 0.f;
      const __half ha = __float2half_rn(a), hb = __float2half_rn(b2);
      const __half la = __float2half_rn(a - __half2float(ha)), lb = __float2half_rn(b2 - __half2float(hb));
      hi[i] = uint32_t(__half_as_ushort(ha)) | (uint32_t(__half_as_ushort(hb)) << 16);
      lo[i] = uint32_t(__half_as_ushort(la)) | (uint32_t(__half_as_ushort(lb)) << 16);
    }
#pragma unroll
    for (int piece = 0; piece < 8; ++piece) {      // k-block 0: [hi | lo]
      const uint32_t* src = piece < 4 ? hi + 4 * piece : lo + 4 * (piece - 4);
      const uint32_t a = sA + swizzle_off<128>(uint32_t(tid), uint32_t(piece));
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(src[0]), "r"(src[1]), "r"(src[2]), "r"(src[3]) : "memory");
    }
#pragma unroll
    for (int piece = 0; piece < 8; ++piece) {      // k-block 1: [hi | 0]
      const uint32_t a = sA + 16384 + swizzle_off<128>(uint32_t(tid), uint32_t(piece));
      if (piece < 4)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(hi[4 * piece]), "r"(hi[4 * piece + 1]), "r"(hi[4 * piece + 2]), "r"(hi[4 * piece + 3]) : "memory");
      else
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(0u) : "memory");
    }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();                               // all rows written; all accumulator reads of the previous tile done
    if (tid == 0) {
      tcgen05_fence_after();
      constexpr uint32_t idesc = make_idesc_f16_f32(128, COUT);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_f16_ss(tmem_base, make_kmajor_desc<128>(sA + uint32_t(k >> 2) * 16384u + 32u * uint32_t(k & 3)),
                    make_kmajor_desc<128>(sW + uint32_t(k >> 2) * 4096u + 32u * uint32_t(k & 3)), idesc, k != 0);
      umma_commit(bar);
    }
    __syncwarp();
    // ---- next tile's inputs are requested while the MMA runs
    const long long pix = tile * 128 + tid;
    if (tile + gridDim.x < num_tiles) gather((tile + gridDim.x) * 128 + tid, cur);
    mbar_wait(bar, phase);
    phase ^= 1u;
    tcgen05_fence_after();
    uint32_t v[32];
    tmem_ld_32x32b<32>(tmem_base + (uint32_t(warp * 32) << 16), v);
    tmem_ld_wait();
    if (pix < total) {
      uint4* out = reinterpret_cast<uint4*>(y + (size_t)pix * COUT);
#pragma unroll
      for (int i = 0; i < COUT / 8; ++i) {
        uint4 o;
        o.x = pack_half2(fmaxf(__uint_as_float(v[8 * i + 0]) + bv[8 * i + 0], 0.f), fmaxf(__uint_as_float(v[8 * i + 1]) + bv[8 * i + 1], 0.f));
        o.y = pack_half2(fmaxf(__uint_as_float(v[8 * i + 2]) + bv[8 * i + 2], 0.f), fmaxf(__uint_as_float(v[8 * i + 3]) + bv[8 * i + 3], 0.f));
        o.z = pack_half2(fmaxf(__uint_as_float(v[8 * i + 4]) + bv[8 * i + 4], 0.f), fmaxf(__uint_as_float(v[8 * i + 5]) + bv[8 * i + 5], 0.f));
        o.w = pack_half2(fmaxf(__uint_as_float(v[8 * i + 6]) + bv[8 * i + 6], 0.f), fmaxf(__uint_as_float(v[8 * i + 7]) + bv[8 * i + 7], 0.f));
        out[i] = o;
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<32>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// Stem conv1, row-tiled: the same hi/lo-split UMMA as stem_conv1_tc_kernel, but one tile = ONE OUTPUT ROW (Ro <= 128 pixels)
// and the three input rows it needs are brought into a shared-memory ring by 1-D bulk copies (cp.async.bulk) two tiles
// ahead.  The ncu capture of the gather version (profiles/r1f: 1.9 TB/s, 24 % warp occupancy, long-scoreboard stalls) shows its
// per-tile chain  global gather -> convert -> UMMA -> store  exposes the DRAM latency every tile; here DRAM requests are always
// in flight without holding registers, and the gather reads shared memory.  Needs 16-B aligned rows (R*3*sizeof(TIn) % 16 == 0).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
constexpr int kStemRowsStages = 3;
template <typename TIn>
constexpr int stem_rows_stage_bytes(int R) { return ((R * 3 * int(sizeof(TIn)) + 127) / 128 * 128) * 3; }
// n / d for n < 2^31 without the integer-division sequence (which goes through the 1/8-rate conversion unit)
struct FastDiv { uint32_t d, m, l; };
inline FastDiv make_fastdiv(uint32_t d) {
  uint32_t l = 0;
  while ((1u << l) < d) ++l;
  const uint64_t m = ((uint64_t(1) << 32) * ((uint64_t(1) << l) - d)) / d + 1;
  return FastDiv{d, uint32_t(m), l};
}
__device__ __forceinline__ uint32_t fast_div(uint32_t n, const FastDiv& f) { return (__umulhi(f.m, n) + n) >> f.l; }
// COUT = 32 (width 64) or 64 (width 96: 48 real channels + 16 zero-weight pad channels).  An output row longer than 128
// pixels (R = 384) is cut into equal segments; every segment's tile loads the three full input rows.
// Software-pipelined over tiles: two im2col buffers [128][hi(32) | lo(32)] and two TMEM accumulators, so the gather of tile
// n+1 runs while the UMMAs of tile n are in flight and the MMA -> commit -> mbarrier latency is off the per-tile chain.
// The six UMMAs per tile are hi x w_hi, lo x w_hi (k-block 0 of the weights) and hi x w_lo (the first half of k-block 1,
// reading the SAME hi columns again: no second copy of hi in shared memory).
template <typename TIn, int COUT, bool kFastU8 = false>
__global__ void __launch_bounds__(128)
stem_conv1_rows_kernel(const TIn* __restrict__ x, const __half* __restrict__ wtc, const float* __restrict__ bias,
                       __half* __restrict__ y, int B, int R, const FastDiv dsegs, const FastDiv dro, const StemNorm norm) {
  constexpr int S = kStemRowsStages;
  constexpr int kWBytes = 2 * COUT * 128;
  constexpr bool kRaw = sizeof(TIn) == 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t row_bytes = uint32_t(R) * 3u * uint32_t(sizeof(TIn));
  const uint32_t row_pitch = (row_bytes + 127u) & ~127u;
  const uint32_t stage_bytes = row_pitch * 3u;
  const uint32_t sA = base;                        // 2 buffers x [128 rows][128 B]
  const uint32_t sW = base + 32768;                // 2 k-blocks x [COUT rows][128 B]
  const uint32_t sIn = sW + kWBytes;               // S stages x 3 rows
  const uint32_t sBar = sIn + S * stage_bytes;     // S full barriers, 2 MMA barriers, TMEM slot
  const uint32_t bar_mma = sBar + 8 * S, tmem_slot = bar_mma + 16;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t Ro = dro.d, segs = dsegs.d;
  const int seg_len = int(Ro / segs);              // <= 128 output pixels per tile

  for (int i = tid; i < COUT * 16; i += 128) {     // weights -> swizzled K-major tiles (as in stem_conv1_tc_kernel)
    const int n = i >> 4, piece = i & 15;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(wtc + n * 128) + piece);
    *reinterpret_cast<uint4*>(gen + 32768 + (piece >> 3) * (COUT * 128) + swizzle_off<128>(uint32_t(n), uint32_t(piece & 7))) = v;
  }
  for (int i = tid; i < 2 * 16384 / 16; i += 128)  // A rows >= seg_len are never written again: they must read as zero
    reinterpret_cast<uint4*>(gen)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int s2 = 0; s2 < S; ++s2) mbar_init(sBar + 8 * s2, 1);
    mbar_init(bar_mma, 1);
    mbar_init(bar_mma + 8, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<2 * COUT>(tmem_slot);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  griddep_wait();

  const uint32_t num_tiles = uint32_t(B) * Ro * segs;              // < 2^31 (checked by the launcher)
  float bv[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) bv[c] = __ldg(bias + c);
  __half2 fk[3], fsh[3], fsl[3], fcl[3];           // kFastU8: per pair phase (channels (0,1), (2,0), (1,2)) constants
  if constexpr (kFastU8) {
    float k1[3], s_h[3], s_l[3], c_l[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float s = norm.scale[c], m = -norm.offset[c] / s, mr = rintf(m);
      k1[c] = 1024.f + mr;                         // exact in fp16 for |mr| <= 1023 (checked by the launcher)
      s_h[c] = __half2float(__float2half_rn(s));
      s_l[c] = s - s_h[c];
      c_l[c] = -s * (m - mr);
    }
#pragma unroll
    for (int ph = 0; ph < 3; ++ph) {
      const int ca = (2 * ph) % 3, cb = (2 * ph + 1) % 3;
      fk[ph] = __floats2half2_rn(k1[ca], k1[cb]);
      fsh[ph] = __floats2half2_rn(s_h[ca], s_h[cb]);
      fsl[ph] = __floats2half2_rn(s_l[ca], s_l[cb]);
      fcl[ph] = __floats2half2_rn(c_l[ca], c_l[cb]);
    }
  }

  // tile -> (image, output row); input rows 2*oh-1 .. 2*oh+1 (row -1 is the zero pad: not loaded, not read)
  auto issue = [&](uint32_t tile, int stage) {
    const uint32_t rowi = fast_div(tile, dsegs);                       // = b * Ro + oh
    const uint32_t oh = rowi - fast_div(rowi, dro) * Ro;
    const int first = oh == 0 ? 1 : 0;
    const uint32_t bar = sBar + 8 * stage;
    mbar_arrive_expect_tx(bar, row_bytes * uint32_t(3 - first));
    for (int kh = first; kh < 3; ++kh)
      bulk_load_1d(sIn + stage * stage_bytes + kh * row_pitch, x + ((size_t)rowi * 2 + size_t(kh) - 1) * R * 3, row_bytes, bar);   // R = 2 Ro: input row b R + 2 oh - 1 + kh
  };
  // im2col rows of one tile (thread <-> output pixel) -> A buffer `buf`
  auto gather = [&](uint32_t tile, int stage, int buf) {
    const uint32_t rowi = fast_div(tile, dsegs);
    const int oh = int(rowi - fast_div(rowi, dro) * Ro);
    const int ow = int(tile - rowi * segs) * seg_len + tid;           // this thread's output column
    if (tid >= seg_len) return;
    uint32_t hi[16], lo[16];
    if constexpr (kFastU8) {
      // Raw uint8 frames without the conversion unit (I2F / F2F issue at 1/8 rate and bounded this kernel): bytes become
      // fp16 through the exponent trick 0x6400 | b = 1024 + b, and  v = s (b - m)  is split into fp16 halves in half2
      // arithmetic:  d = b - round(m) exact,  hi = fl(s_h d),  lo = fma(s_h, d, -hi) + fma(s_l, d, -s (m - round(m))).
      const uint32_t q = uint32_t(6 * ow - 3);                         // first byte of the 9-byte run (row-relative)
      const uint32_t sh = (q & 3u) * 8u;
      uint32_t xr[3][3];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const uint32_t a = sIn + stage * stage_bytes + kh * row_pitch + (q & ~3u);   // ow == 0 reads 4 B before the row: masked
        uint32_t w0, w1, w2;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w0) : "r"(a));
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w1) : "r"(a + 4u));
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w2) : "r"(a + 8u));
        xr[kh][0] = __funnelshift_r(w0, w1, sh);
        xr[kh][1] = __funnelshift_r(w1, w2, sh);
        xr[kh][2] = w2 >> sh;
      }
      uint32_t sb[7];                                                  // the 27 bytes as one stream: pair i = bytes 2i, 2i+1
      sb[0] = xr[0][0];
      sb[1] = xr[0][1];
      sb[2] = __byte_perm(xr[0][2], xr[1][0], 0x6540);
      sb[3] = __byte_perm(xr[1][0], xr[1][1], 0x6543);
      sb[4] = __byte_perm(__byte_perm(xr[1][1], xr[1][2], 0x0043), xr[2][0], 0x5410);
      sb[5] = __byte_perm(xr[2][0], xr[2][1], 0x5432);
      sb[6] = __byte_perm(xr[2][1], xr[2][2], 0x0432);
      const uint32_t mrow = oh > 0 ? 0xffffffffu : 0u, mcol = ow > 0 ? 0xffffffffu : 0u;
#pragma unroll
      for (int i = 0; i < 14; ++i) {
        const uint32_t raw = __byte_perm(sb[i >> 1], 0x64646464u, (i & 1) ? 0x4342 : 0x4140);
        const __half2 d = __hsub2(*reinterpret_cast<const __half2*>(&raw), fk[i % 3]);
        const __half2 h = __hmul2(fsh[i % 3], d);
        const __half2 l = __hadd2(__hfma2(fsh[i % 3], d, __hneg2(h)), __hfma2(fsl[i % 3], d, fcl[i % 3]));
        uint32_t mk = 0u;                                              // zero padding taps: row -1, column -1, element 27
#pragma unroll
        for (int e2 = 0; e2 < 2; ++e2) {
          const int e = 2 * i + e2, kh = e / 9, j = e % 9;
          uint32_t m1 = e < 27 ? 0xffffu : 0u;
          if (kh == 0) m1 &= mrow;
          if (j < 3) m1 &= mcol;
          mk |= m1 << (16 * e2);
        }
        hi[i] = *reinterpret_cast<const uint32_t*>(&h) & mk;
        lo[i] = *reinterpret_cast<const uint32_t*>(&l) & mk;
      }
      hi[14] = hi[15] = lo[14] = lo[15] = 0u;
    } else {
      const uint8_t* st = gen + (sIn - base) + stage * stage_bytes;
      float v[28];
      v[27] = 0.f;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const bool rok = kh > 0 || oh > 0;
        const TIn* row = reinterpret_cast<const TIn*>(st + kh * row_pitch) + (2 * ow - 1) * 3;
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          float t = 0.f;
          if (rok && (ow > 0 || j >= 3)) {
            t = float(row[j]);
            if (kRaw) t = t * norm.scale[j % 3] + norm.offset[j % 3];
          }
          v[kh * 9 + j] = t;
        }
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {                                   // pairwise F2FP packs (scalar F2F issues at 1/8 rate)
        const float a = i < 14 ? v[2 * i] : 0.f, b2 = i < 14 ? v[2 * i + 1] : 0.f;
        const __half2 h = __floats2half2_rn(a, b2);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(a - hf.x, b2 - hf.y);
        hi[i] = *reinterpret_cast<const uint32_t*>(&h);
        lo[i] = *reinterpret_cast<const uint32_t*>(&l);
      }
    }
#pragma unroll
    for (int piece = 0; piece < 8; ++piece) {      // [hi | lo]
      const uint32_t* src = piece < 4 ? hi + 4 * piece : lo + 4 * (piece - 4);
      const uint32_t a = sA + uint32_t(buf) * 16384u + swizzle_off<128>(uint32_t(tid), uint32_t(piece));
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(src[0]), "r"(src[1]), "r"(src[2]), "r"(src[3]) : "memory");
    }
  };
  auto mma = [&](int buf) {                        // one thread
    tcgen05_fence_after();
    constexpr uint32_t idesc = make_idesc_f16_f32(128, COUT);
    const uint32_t a0 = sA + uint32_t(buf) * 16384u, acc = tmem_base + uint32_t(buf * COUT);
#pragma unroll
    for (int k = 0; k < 4; ++k)                    // hi x w_hi, lo x w_hi
      umma_f16_ss(acc, make_kmajor_desc<128>(a0 + 32u * uint32_t(k)), make_kmajor_desc<128>(sW + 32u * uint32_t(k)), idesc, k != 0);
#pragma unroll
    for (int k = 0; k < 2; ++k)                    // hi x w_lo
      umma_f16_ss(acc, make_kmajor_desc<128>(a0 + 32u * uint32_t(k)), make_kmajor_desc<128>(sW + uint32_t(COUT * 128) + 32u * uint32_t(k)), idesc, 1);
    umma_commit(bar_mma + 8u * uint32_t(buf));
  };
  auto epilogue = [&](uint32_t tile, int buf) {
    const uint32_t rowi = fast_div(tile, dsegs);
    const int ow = int(tile - rowi * segs) * seg_len + tid;
#pragma unroll
    for (int c0 = 0; c0 < COUT; c0 += 32) {
      uint32_t acc[32];
      tmem_ld_32x32b<32>(tmem_base + (uint32_t(warp * 32) << 16) + uint32_t(buf * COUT + c0), acc);
      tmem_ld_wait();
      if (tid < seg_len) {
        uint4* out = reinterpret_cast<uint4*>(y + ((size_t)rowi * Ro + ow) * COUT + c0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 o;
          o.x = pack_half2(fmaxf(__uint_as_float(acc[8 * i + 0]) + bv[c0 + 8 * i + 0], 0.f), fmaxf(__uint_as_float(acc[8 * i + 1]) + bv[c0 + 8 * i + 1], 0.f));
          o.y = pack_half2(fmaxf(__uint_as_float(acc[8 * i + 2]) + bv[c0 + 8 * i + 2], 0.f), fmaxf(__uint_as_float(acc[8 * i + 3]) + bv[c0 + 8 * i + 3], 0.f));
          o.z = pack_half2(fmaxf(__uint_as_float(acc[8 * i + 4]) + bv[c0 + 8 * i + 4], 0.f), fmaxf(__uint_as_float(acc[8 * i + 5]) + bv[c0 + 8 * i + 5], 0.f));
          o.w = pack_half2(fmaxf(__uint_as_float(acc[8 * i + 6]) + bv[c0 + 8 * i + 6], 0.f), fmaxf(__uint_as_float(acc[8 * i + 7]) + bv[c0 + 8 * i + 7], 0.f));
          out[i] = o;
        }
      }
    }
  };

  if (tid == 0) {
    uint32_t t = blockIdx.x;
    for (int s2 = 0; s2 < S && t < num_tiles; ++s2, t += gridDim.x) issue(t, s2);
  }
  int stage = 0;
  uint32_t in_phase = 0, mma_phase = 0;            // mma_phase: one bit per accumulator
  uint32_t tile = blockIdx.x;
  if (tile < num_tiles) {                          // prologue: tile 0's rows and UMMAs
    mbar_wait(sBar, 0);
    gather(tile, 0, 0);
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      mma(0);
      const uint64_t nxt = uint64_t(tile) + uint64_t(S) * gridDim.x;
      if (nxt < num_tiles) issue(uint32_t(nxt), 0);
    }
    __syncwarp();
    stage = 1;
  }
  for (int n = 0; tile < num_tiles; tile += gridDim.x, ++n) {
    const int buf = n & 1;
    const uint64_t nt = uint64_t(tile) + gridDim.x;
    const bool more = nt < num_tiles;
    if (more) {
      mbar_wait(sBar + 8 * stage, in_phase);
      gather(uint32_t(nt), stage, buf ^ 1);
    }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();                               // next A written, its input stage fully read, accumulator buf^1 drained
    if (more) {
      if (tid == 0) {
        mma(buf ^ 1);
        const uint64_t nxt = nt + uint64_t(S) * gridDim.x;            // refill the stage just consumed
        if (nxt < num_tiles) issue(uint32_t(nxt), stage);
      }
      __syncwarp();
      if (++stage == S) { stage = 0; in_phase ^= 1u; }
    }
    mbar_wait(bar_mma + 8u * uint32_t(buf), (mma_phase >> buf) & 1u);
    mma_phase ^= 1u << buf;
    tcgen05_fence_after();
    epilogue(tile, buf);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<2 * COUT>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// nn.AvgPool2d(2) on NHWC fp16; one thread = 8 channels of one output pixel (16-B vectors).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void acc8(float (&a)[8], const uint4 v) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    a[2 * i] += f.x;
    a[2 * i + 1] += f.y;
  }
}
__global__ void __launch_bounds__(256)
avgpool2_kernel(const __half* __restrict__ x, __half* __restrict__ y, int B, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2, C8 = C / 8;
  const long long total = (long long)B * Ho * Wo * C8;
  griddep_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = int(i % C8);
    long long r = i / C8;
    const int ow = int(r % Wo);
    r /= Wo;
    const int oh = int(r % Ho);
    const int b = int(r / Ho);
    const uint4* p = reinterpret_cast<const uint4*>(x + (((size_t)b * H + 2 * oh) * W + 2 * ow) * C) + c8;
    const size_t rowstride = (size_t)W * C8;
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    acc8(a, __ldg(p));
    acc8(a, __ldg(p + C8));
    acc8(a, __ldg(p + rowstride));
    acc8(a, __ldg(p + rowstride + C8));
    uint4 o;
    o.x = pack_half2(a[0] * .25f, a[1] * .25f);
    o.y = pack_half2(a[2] * .25f, a[3] * .25f);
    o.z = pack_half2(a[4] * .25f, a[5] * .25f);
    o.w = pack_half2(a[6] * .25f, a[7] * .25f);
    reinterpret_cast<uint4*>(y)[i] = o;
  }
}

// ------------------------------------------------------------------------------------------------
// AttentionPool2d front end: tokens[b,0,:] = mean_p x[b,p,:] + pos[0];  tokens[b,1+p,:] = x[b,p,:] + pos[1+p].
// x is the fp32 NHWC trunk output [B, P, C]; one thread = 4 channels.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
attnpool_tokens_kernel(const float* __restrict__ x, const float* __restrict__ pos, __half* __restrict__ tok,
                       int P, int C) {
  const int b = blockIdx.y;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  griddep_wait();
  if (c >= C) return;
  const float* xb = x + (size_t)b * P * C + c;
  __half* tb = tok + (size_t)b * (P + 1) * C + c;
  float4 s = make_float4(0, 0, 0, 0);
  for (int p = 0; p < P; ++p) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(xb + (size_t)p * C));
    const float4 e = __ldg(reinterpret_cast<const float4*>(pos + (size_t)(p + 1) * C + c));
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    uint2 o;
    o.x = pack_half2(v.x + e.x, v.y + e.y);
    o.y = pack_half2(v.z + e.z, v.w + e.w);
    *reinterpret_cast<uint2*>(tb + (size_t)(p + 1) * C) = o;
  }
  const float inv = 1.f / float(P);
  const float4 e0 = __ldg(reinterpret_cast<const float4*>(pos + c));
  uint2 o;
  o.x = pack_half2(s.x * inv + e0.x, s.y * inv + e0.y);
  o.y = pack_half2(s.z * inv + e0.z, s.w * inv + e0.w);
  *reinterpret_cast<uint2*>(tb) = o;
}

// ------------------------------------------------------------------------------------------------
// AttentionPool2d core.  Only query row 0 is consumed downstream, and
//   q_h . (Wk_h t_j + bk_h) = (Wk_h^T q_h) . t_j + const_h        (const_h drops out of the softmax)
//   sum_j p_j (Wv_h t_j + bv_h) = Wv_h (sum_j p_j t_j) + bv_h     (sum_j p_j = 1)
// so per (image, head) this kernel takes the folded query qt = Wk_h^T q_h (length C) and the L = HW+1 tokens,
// computes s_j = qt . t_j, p = softmax(s) and the head's token average  xbar_h = sum_j p_j t_j  (length C).
// The two remaining contractions (with Wk^T before, Wv after) are grouped GEMMs on the tensor pipe.
// Grid (B, heads/HG); 256 threads.  Requires C == 2048 * (C/2048) multiple of 2048? no: C % 256*8 == 0.
// ------------------------------------------------------------------------------------------------
template <int HG>
__global__ void __launch_bounds__(256)
attnpool_core_kernel(const __half* __restrict__ qt,   // [B, heads, C]
                     const __half* __restrict__ tok,  // [B, L, C]
                     __half* __restrict__ xbar,       // [B, heads, C]
                     int heads, int L, int C) {
  extern __shared__ uint8_t smem[];
  __half* sq = reinterpret_cast<__half*>(smem);                        // [HG][C]
  float* sp = reinterpret_cast<float*>(smem + (size_t)HG * C * 2);     // [HG][64]
  const int b = blockIdx.x;
  const int h0 = blockIdx.y * HG;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C8 = C / 8;
  griddep_wait();

  const uint4* gq = reinterpret_cast<const uint4*>(qt + ((size_t)b * heads + h0) * C);
  for (int i = tid; i < HG * C8; i += 256) reinterpret_cast<uint4*>(sq)[i] = __ldg(gq + i);
  __syncthreads();

  // scores: warp <-> token, lanes stride the channel dimension in 16-B pieces
  const __half* tb = tok + (size_t)b * L * C;
  for (int j = warp; j < L; j += 8) {
    float part[HG];
#pragma unroll
    for (int h = 0; h < HG; ++h) part[h] = 0.f;
    const uint4* tj = reinterpret_cast<const uint4*>(tb + (size_t)j * C);
    for (int i = lane; i < C8; i += 32) {
      const uint4 tv = __ldg(tj + i);
      const __half2* t2 = reinterpret_cast<const __half2*>(&tv);
      float tf[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(t2[k]);
        tf[2 * k] = f.x;
        tf[2 * k + 1] = f.y;
      }
#pragma unroll
      for (int h = 0; h < HG; ++h) {
        const uint4 qv = reinterpret_cast<const uint4*>(sq + (size_t)h * C)[i];
        const __half2* q2 = reinterpret_cast<const __half2*>(&qv);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __half22float2(q2[k]);
          part[h] += f.x * tf[2 * k] + f.y * tf[2 * k + 1];
        }
      }
    }
#pragma unroll
    for (int h = 0; h < HG; ++h) {
      float v = part[h];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) sp[h * 64 + j] = v;
    }
  }
  __syncthreads();

  // softmax over the L keys: warp <-> head
  if (warp < HG) {
    const float v0 = lane < L ? sp[warp * 64 + lane] : -INFINITY;
    const float v1 = lane + 32 < L ? sp[warp * 64 + lane + 32] : -INFINITY;
    float m = fmaxf(v0, v1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float e0 = lane < L ? __expf(v0 - m) : 0.f;
    const float e1 = lane + 32 < L ? __expf(v1 - m) : 0.f;
    float s = e0 + e1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float inv = 1.f / s;
    if (lane < L) sp[warp * 64 + lane] = e0 * inv;
    if (lane + 32 < L) sp[warp * 64 + lane + 32] = e1 * inv;
  }
  __syncthreads();

  // weighted token average: thread <-> 8 channels, all HG heads
  for (int i = tid; i < C8; i += 256) {
    float acc[HG][8];
#pragma unroll
    for (int h = 0; h < HG; ++h)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[h][k] = 0.f;
    for (int j = 0; j < L; ++j) {
      const uint4 tv = __ldg(reinterpret_cast<const uint4*>(tb + (size_t)j * C) + i);
      const __half2* t2 = reinterpret_cast<const __half2*>(&tv);
      float tf[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(t2[k]);
        tf[2 * k] = f.x;
        tf[2 * k + 1] = f.y;
      }
#pragma unroll
      for (int h = 0; h < HG; ++h) {
        const float pj = sp[h * 64 + j];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[h][k] += pj * tf[k];
      }
    }
#pragma unroll
    for (int h = 0; h < HG; ++h) {
      uint4 o;
      o.x = pack_half2(acc[h][0], acc[h][1]);
      o.y = pack_half2(acc[h][2], acc[h][3]);
      o.z = pack_half2(acc[h][4], acc[h][5]);
      o.w = pack_half2(acc[h][6], acc[h][7]);
      reinterpret_cast<uint4*>(xbar + ((size_t)b * heads + h0 + h) * C)[i] = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The same two contractions on the tensor cores (one CTA = 4 images = 128 (image, head) rows):
//   S[128 x 256]   = QT[128 rows x C] . T[256 token rows x C]^T      K-loop over C in 64-channel TMA stages (both K-major)
//   P              = softmax over the row's OWN image's L tokens (other images' columns are exactly 0), fp16, K-major smem
//   XBAR[128 x C]  = P[128 x 256] . T[256 x C]                        per 128-channel chunk; T is the MN-major operand
// The 4x4 cross-image blocks of S are computed and discarded (free on the tensor pipe).  Tokens are read twice from L2
// (once per phase), qt once; 192 KB smem (phase-1 stages overlaid with P + the phase-3 token tiles), all 512 TMEM columns.
// ------------------------------------------------------------------------------------------------
struct AttnPoolTcParams {
  int B, heads, L, C;      // heads * 4 == 128, 4 * L <= 256, C % 128 == 0
  __half* xbar;            // [B * heads][C]
};
constexpr int kApStages = 3;
constexpr int kApStageBytes = 16384 + 32768;
constexpr int kApSmem = 1024 + 196608 + 128;
__global__ void __launch_bounds__(128, 1)
attnpool_tc_kernel(const __grid_constant__ CUtensorMap tmQT, const __grid_constant__ CUtensorMap tmTokK,
                   const __grid_constant__ CUtensorMap tmTokMN, const AttnPoolTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // phase 1: stages [A 16 KB | B 32 KB] x 3 ; phases 2-3: P 64 KB at +0, token tiles 2 x 64 KB at +64 KB
  const uint32_t sP = base, sT = base + 65536;
  const uint32_t bars = base + 196608;
  const uint32_t bar_full = bars, bar_empty = bars + 24, bar_s = bars + 48, bar_tfull = bars + 56, bar_o = bars + 72, tmem_slot = bars + 88;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int tile = blockIdx.x;
  const int rq0 = tile * 128;                      // first (image, head) row
  const int rt0 = tile * 4 * p.L;                  // first token row
  const int nkb = p.C / 64;

  if (tid == 0) {
    tma_prefetch_desc(&tmQT); tma_prefetch_desc(&tmTokK); tma_prefetch_desc(&tmTokMN);
    for (int s = 0; s < kApStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_s, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_o + 8 * i, 1); }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  griddep_wait();
  const uint32_t tS = tmem_base, tO = tmem_base + 256;

  // ---------------- phase 1: scores
  if (tid == 0) {
    constexpr uint32_t idesc = make_idesc_f16_f32(128, 256);
    auto load = [&](int kb) {
      const int s = kb % kApStages;
      const uint32_t full = bar_full + 8 * s, a = base + s * kApStageBytes;
      mbar_arrive_expect_tx(full, kApStageBytes);
      tma_load_2d(&tmQT, full, a, kb * 64, rq0);
      tma_load_2d(&tmTokK, full, a + 16384, kb * 64, rt0);
    };
    for (int kb = 0; kb < kApStages && kb < nkb; ++kb) load(kb);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % kApStages;
      const uint32_t ph = uint32_t(kb / kApStages) & 1u;
      mbar_wait(bar_full + 8 * s, ph);
      tcgen05_fence_after();
      const uint32_t a = base + s * kApStageBytes;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16_ss(tS, make_kmajor_desc<128>(a + 32 * k), make_kmajor_desc<128>(a + 16384 + 32 * k), idesc, (kb | k) != 0);
      umma_commit(bar_empty + 8 * s);
      if (kb + kApStages < nkb) {
        mbar_wait(bar_empty + 8 * s, ph);          // the MMAs reading this stage have retired
        load(kb + kApStages);
      }
    }
    umma_commit(bar_s);
  }
  __syncwarp();

  // ---------------- phase 2: softmax, thread <-> (image, head) row
  const int r = tid;
  const int img = r / p.heads;                     // image of this row inside the tile
  const int j0 = img * p.L, j1 = j0 + p.L;
  const bool live = rq0 + r < p.B * p.heads;
  mbar_wait(bar_s, 0);
  tcgen05_fence_after();
  const uint32_t trow = tS + (uint32_t(warp * 32) << 16);
  float mx = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < 8; ++c) {
    if (32 * c + 31 < j0 || 32 * c >= j1) continue;
    uint32_t v[32];
    tmem_ld_32x32b<32>(trow + uint32_t(32 * c), v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int j = 32 * c + i;
      if (j >= j0 && j < j1) mx = fmaxf(mx, __uint_as_float(v[i]));
    }
  }
  float sum = 0.f;
#pragma unroll 1
  for (int c = 0; c < 8; ++c) {
    uint32_t h2[16];
    if (32 * c + 31 < j0 || 32 * c >= j1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) h2[i] = 0u;
    } else {
      uint32_t v[32];
      tmem_ld_32x32b<32>(trow + uint32_t(32 * c), v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const int j = 32 * c + i;
        const float a = (j >= j0 && j < j1) ? __expf(__uint_as_float(v[i]) - mx) : 0.f;
        const float b = (j + 1 >= j0 && j + 1 < j1) ? __expf(__uint_as_float(v[i + 1]) - mx) : 0.f;
        sum += a + b;
        h2[i >> 1] = pack_half2(a, b);
      }
    }
    const uint32_t chunk = sP + uint32_t((32 * c) / 64) * 16384u;
    const int piece0 = ((32 * c) % 64) / 8;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t a = chunk + swizzle_off<128>(uint32_t(r), uint32_t(piece0 + q));
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(h2[4 * q]), "r"(h2[4 * q + 1]), "r"(h2[4 * q + 2]), "r"(h2[4 * q + 3]) : "memory");
    }
  }
  const float inv = sum > 0.f ? 1.f / sum : 0.f;
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();                                 // P complete; phase-1 stages are dead (their MMAs retired before bar_s)

  // ---------------- phase 3: XBAR chunks of 128 channels.  gridDim.y CTAs share a tile: each repeats phases 1-2 (11 of the
  // kernel's 60 us) and takes a contiguous slice of the channel chunks, so a small batch still fills the SMs (`c` below is the
  // LOCAL chunk index: buffers and phases follow it, channel offsets follow c0 + c)
  const int c0 = int(blockIdx.y) * (p.C / 128) / int(gridDim.y);
  const int nchunks = (int(blockIdx.y) + 1) * (p.C / 128) / int(gridDim.y) - c0;
  auto load_t = [&](int c) {
    const uint32_t bar = bar_tfull + 8 * (c & 1), dst = sT + uint32_t(c & 1) * 65536u;
    mbar_arrive_expect_tx(bar, 65536);
    tma_load_2d(&tmTokMN, bar, dst, (c0 + c) * 128, rt0);
    tma_load_2d(&tmTokMN, bar, dst + 32768, (c0 + c) * 128 + 64, rt0);
  };
  if (tid == 0) { load_t(0); if (nchunks > 1) load_t(1); }
  for (int c = 0; c <= nchunks; ++c) {
    if (tid == 0 && c < nchunks) {
      mbar_wait(bar_tfull + 8 * (c & 1), uint32_t(c >> 1) & 1u);
      tcgen05_fence_after();
      constexpr uint32_t idesc = make_idesc_f16_f32(128, 128) | (1u << 16);    // B (tokens) is MN-major
      const uint32_t t = sT + uint32_t(c & 1) * 65536u;
#pragma unroll
      for (int ks = 0; ks < 16; ++ks)
        umma_f16_ss(tO + uint32_t(c & 1) * 128u, make_kmajor_desc<128>(sP + uint32_t(ks >> 2) * 16384u + 32u * uint32_t(ks & 3)),
                    make_mnmajor_desc<128>(t + uint32_t(ks) * 2048u, 32768u), idesc, ks != 0);
      umma_commit(bar_o + 8 * (c & 1));
    }
    __syncwarp();
    if (c >= 1) {                                  // epilogue of chunk c-1 while chunk c's MMAs run
      const int e = c - 1;
      mbar_wait(bar_o + 8 * (e & 1), uint32_t(e >> 1) & 1u);
      tcgen05_fence_after();
      if (tid == 0 && e + 2 < nchunks) load_t(e + 2);          // its token tile is free again
      __half* orow = p.xbar + (size_t)(rq0 + r) * p.C + (size_t)(c0 + e) * 128;
#pragma unroll 1
      for (int q4 = 0; q4 < 4; ++q4) {
        uint32_t v[32];
        tmem_ld_32x32b<32>(tO + uint32_t(e & 1) * 128u + (uint32_t(warp * 32) << 16) + uint32_t(32 * q4), v);
        tmem_ld_wait();
        if (live) {
          uint4* o4 = reinterpret_cast<uint4*>(orow + 32 * q4);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            o.x = pack_half2(__uint_as_float(v[8 * q]) * inv, __uint_as_float(v[8 * q + 1]) * inv);
            o.y = pack_half2(__uint_as_float(v[8 * q + 2]) * inv, __uint_as_float(v[8 * q + 3]) * inv);
            o.z = pack_half2(__uint_as_float(v[8 * q + 4]) * inv, __uint_as_float(v[8 * q + 5]) * inv);
            o.w = pack_half2(__uint_as_float(v[8 * q + 6]) * inv, __uint_as_float(v[8 * q + 7]) * inv);
            o4[q] = o;
          }
        }
      }
      tcgen05_fence_before();
    }
    __syncthreads();                               // accumulator e drained by every row before chunk e+2's MMAs overwrite it
  }
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// Output heads from the fp32 NHWC trunk result [B, P, C].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
avg_head_kernel(const float* __restrict__ x, float* __restrict__ y, int P, int C) {
  const int b = blockIdx.y;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  griddep_wait();
  if (c >= C) return;
  const float* xb = x + (size_t)b * P * C + c;
  float4 s = make_float4(0, 0, 0, 0);
  for (int p = 0; p < P; ++p) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(xb + (size_t)p * C));
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  const float inv = 1.f / float(P);
  *reinterpret_cast<float4*>(y + (size_t)b * C + c) = make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv);
}

// [B, P, C] -> [B, C, P]; one block = 32 channels of one image, out slab [32][P] is contiguous.
__global__ void __launch_bounds__(256)
nhwc_to_nchw_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int P, int C) {
  extern __shared__ float tile[];   // [P][33]
  const int b = blockIdx.y, c0 = blockIdx.x * 32;
  griddep_wait();
  const float* xb = x + (size_t)b * P * C + c0;
  for (int i = threadIdx.x; i < P * 32; i += blockDim.x) {
    const int p = i >> 5, c = i & 31;
    tile[p * 33 + c] = __ldg(xb + (size_t)p * C + c);
  }
  __syncthreads();
  float* yb = y + ((size_t)b * C + c0) * P;
  for (int i = threadIdx.x; i < P * 32; i += blockDim.x) {
    const int c = i / P, p = i - c * P;
    yb[i] = tile[p * 33 + c];
  }
}

// The same with 128 channels of one image per block: the out slab [128][P] (P * 512 B, contiguous and 16-B aligned) is
// assembled in shared memory in its final order and leaves as float4 rows; loads are 128-B segments (8 lanes x float4) of four
// pixels per warp, scattered conflict-free (bank = 4 (lane & 7) + k P + pixel: P odd or not, the 8 x 4 lanes hit 32 banks when
// P is odd; P even costs a 2-way conflict).  Needs C % 128 == 0.
__global__ void __launch_bounds__(256)
nhwc_to_nchw_f32_wide_kernel(const float* __restrict__ x, float* __restrict__ y, int P, int C) {
  extern __shared__ float tile[];   // [128][P]
  const int b = blockIdx.y, c0 = blockIdx.x * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c4 = (warp & 3) * 8 + (lane & 7);       // float4 index within the 128 channels
  griddep_wait();
  const float4* xb = reinterpret_cast<const float4*>(x + (size_t)b * P * C + c0) + c4;
  const int cq = C / 4;
  for (int p = 4 * (warp >> 2) + (lane >> 3); p < P; p += 8) {
    const float4 v = __ldg(xb + (size_t)p * cq);
    float* t = tile + (4 * c4) * P + p;
    t[0] = v.x; t[P] = v.y; t[2 * P] = v.z; t[3 * P] = v.w;
  }
  __syncthreads();
  float4* yb = reinterpret_cast<float4*>(y + ((size_t)b * C + c0) * P);
  const float4* t4 = reinterpret_cast<const float4*>(tile);
  for (int i = threadIdx.x; i < 32 * P; i += 256) yb[i] = t4[i];
}

// ------------------------------------------------------------------------------------------------
// fp32 -> fp16 copy (8 elements per thread): the trunk output as the actor-critic path's pixel rows.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cast_f32_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, long long n8) {
  griddep_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(in) + 2 * i), b = __ldg(reinterpret_cast<const float4*>(in) + 2 * i + 1);
    uint4 o;
    o.x = pack_half2(a.x, a.y); o.y = pack_half2(a.z, a.w); o.z = pack_half2(b.x, b.y); o.w = pack_half2(b.z, b.w);
    reinterpret_cast<uint4*>(out)[i] = o;
  }
}

}  // namespace embclip
