// Non-GEMM kernels of the CLIP transformer towers (clip/model.py VisionTransformer / ResidualAttentionBlock /
// CLIP.encode_text / CLIP.forward).  The residual stream is fp32; LayerNorm is computed
// in fp32 (upstream's LayerNorm subclass does the same) and emits the fp16 operand of the next GEMM.
//   vit_patchify        fp32 NHWC frames -> fp16 [B*g*g, p*p*3] patch rows (K order kh, kw, c)
//   vit_embed_ln_pre    class token + positional embedding + ln_pre -> fp32 stream [B*L, D]
//   text_embed          token_embedding[ids] + positional_embedding -> fp32 stream [K*L, D]
//   layernorm_rows      fp32 rows (optionally gathered) -> fp16 rows
//   attention_kernel    softmax(q k^T [+ causal mask]) v for one (sequence, head): L <= 96, head dim 64
//   clip_logits         exp(logit_scale) * normalize(img) . normalize(txt)^T
#pragma once
#include "ptx.cuh"

namespace embclip {

// one block = one patch: y[(b, py, px)][kh*PS*3 + kw*3 + c] = x[b][py*PS + kh][px*PS + kw][c]
__global__ void __launch_bounds__(256)
vit_patchify_kernel(const float* __restrict__ x, __half* __restrict__ y, int B, int R, int PS) {
  const int g = R / PS;
  const int row_elems = PS * 3, patch_elems = PS * row_elems;
  const long long total = (long long)B * g * g;
  const bool vec8 = row_elems % 8 == 0 && (R * 3) % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(y) % 16 == 0;
  for (long long i = blockIdx.x; i < total; i += gridDim.x) {
    const int px = int(i % g);
    const int py = int((i / g) % g);
    const int b = int(i / ((long long)g * g));
    const float* src = x + (((size_t)b * R + (size_t)py * PS) * R + (size_t)px * PS) * 3;
    __half* dst = y + (size_t)i * patch_elems;
    if (vec8) {
      // a patch row is PS*3 contiguous floats: 8 per thread (two 16-B loads, one 16-B store)
      const int vpr = row_elems / 8;
      for (int j = threadIdx.x; j < PS * vpr; j += blockDim.x) {
        const int kh = j / vpr, v = j - kh * vpr;
        const float4* s4 = reinterpret_cast<const float4*>(src + (size_t)kh * R * 3 + v * 8);
        const float4 a = __ldg(s4), c = __ldg(s4 + 1);
        uint4 o;
        o.x = pack_half2(a.x, a.y); o.y = pack_half2(a.z, a.w); o.z = pack_half2(c.x, c.y); o.w = pack_half2(c.z, c.w);
        *reinterpret_cast<uint4*>(dst + kh * row_elems + v * 8) = o;
      }
      continue;
    }
    for (int j = threadIdx.x; j < patch_elems; j += blockDim.x) {
      const int kh = j / row_elems, r = j - kh * row_elems;
      dst[j] = __float2half_rn(__ldg(src + (size_t)kh * R * 3 + r));
    }
  }
}

// warp-level LayerNorm of one row held as NV float4 per lane (D = 128 * NV); two-pass in registers
template <int NV>
__device__ __forceinline__ void warp_layernorm(float4 (&v)[NV], int D, float eps) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += v[i].x + v[i].y + v[i].z + v[i].w;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / float(D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float inv = rsqrtf(q / float(D) + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) { v[i].x *= inv; v[i].y *= inv; v[i].z *= inv; v[i].w *= inv; }
}

// x[b*L + l] = ln_pre( (l == 0 ? cls : patch[b*(L-1) + l-1]) + pos[l] );  one warp per row, D = 128 * NV
template <int NV>
__global__ void __launch_bounds__(256)
vit_embed_ln_pre_kernel(const float* __restrict__ patch, const float* __restrict__ cls, const float* __restrict__ pos,
                        const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ x, int rows, int L, int D) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int b = row / L, l = row - b * L;
  const float4* src = reinterpret_cast<const float4*>(l == 0 ? cls : patch + ((size_t)b * (L - 1) + l - 1) * D);
  const float4* pp = reinterpret_cast<const float4*>(pos + (size_t)l * D);
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 a = __ldg(src + lane + 32 * i), p4 = __ldg(pp + lane + 32 * i);
    v[i] = make_float4(a.x + p4.x, a.y + p4.y, a.z + p4.z, a.w + p4.w);
  }
  warp_layernorm<NV>(v, D, 1e-5f);
  float4* out = reinterpret_cast<float4*>(x + (size_t)row * D);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i), b4 = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
    out[lane + 32 * i] = make_float4(v[i].x * g4.x + b4.x, v[i].y * g4.y + b4.y, v[i].z * g4.z + b4.z, v[i].w * g4.w + b4.w);
  }
}

// x[k*L + l] = tok_emb[ids[k*L + l]] + pos[l]
__global__ void __launch_bounds__(256)
text_embed_kernel(const long long* __restrict__ ids, const float* __restrict__ emb, const float* __restrict__ pos,
                  float* __restrict__ x, int rows, int L, int D, int vocab) {
  const long long total = (long long)rows * (D / 4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = int(i % (D / 4));
    const int row = int(i / (D / 4));
    long long id = ids[row];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const float4 a = __ldg(reinterpret_cast<const float4*>(emb + (size_t)id * D) + c4);
    const float4 p4 = __ldg(reinterpret_cast<const float4*>(pos + (size_t)(row % L) * D) + c4);
    reinterpret_cast<float4*>(x)[i] = make_float4(a.x + p4.x, a.y + p4.y, a.z + p4.z, a.w + p4.w);
  }
}

// y[r] = fp16( LayerNorm(x[src_row(r)]) * gamma + beta ).  src_row(r) = gather ? gather[r] : r * row_stride.
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_rows_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                      __half* __restrict__ y, int rows, int D, long long row_stride, const long long* __restrict__ gather) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const long long srow = gather ? gather[row] : (long long)row * row_stride;
  const float4* src = reinterpret_cast<const float4*>(x + (size_t)srow * D);
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = src[lane + 32 * i];
  warp_layernorm<NV>(v, D, 1e-5f);
  uint2* out = reinterpret_cast<uint2*>(y + (size_t)row * D);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i), b4 = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
    uint2 o;
    o.x = pack_half2(v[i].x * g4.x + b4.x, v[i].y * g4.y + b4.y);
    o.y = pack_half2(v[i].z * g4.z + b4.z, v[i].w * g4.w + b4.w);
    out[lane + 32 * i] = o;
  }
}

// eot[k] = k*L + argmax_l ids[k][l]  (first maximum, like torch.argmax): the row ln_final / text_projection read
__global__ void text_eot_rows_kernel(const long long* __restrict__ ids, long long* __restrict__ eot, int K, int L) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  long long best = ids[(size_t)k * L];
  int arg = 0;
  for (int l = 1; l < L; ++l) {
    const long long v = ids[(size_t)k * L + l];
    if (v > best) { best = v; arg = l; }
  }
  eot[k] = (long long)k * L + arg;
}

// ------------------------------------------------------------------------------------------------
// Multi-head self-attention core for short sequences.  qkv fp16 [S*L, 3*D] rows (q pre-scaled by head_dim^-0.5 at
// weight-pack time), heads of 64 channels.  grid (S, heads), 128 threads: the (sequence, head)'s K and V live in shared
// memory (rows padded to 66 halves: conflict-free), each warp owns query rows i = warp, warp+4, ...; lanes split the
// keys for the scores and the 64 output channels for the weighted sum.  fp32 math, fp16 in / out.
// ------------------------------------------------------------------------------------------------
constexpr int kAttnMaxL = 96;
__global__ void __launch_bounds__(128)
attention_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int L, int D, int causal) {
  __shared__ __half sK[kAttnMaxL][66];
  __shared__ __half sV[kAttnMaxL][66];
  __shared__ __half sQ[4][64];
  __shared__ float sP[4][kAttnMaxL];
  const int s = blockIdx.x, h = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t ld = (size_t)3 * D;
  const __half* base = qkv + (size_t)s * L * ld + (size_t)h * 64;
  for (int i = tid; i < L * 32; i += 128) {
    const int l = i >> 5, c2 = i & 31;
    *reinterpret_cast<__half2*>(&sK[l][2 * c2]) = *reinterpret_cast<const __half2*>(base + l * ld + D + 2 * c2);
    *reinterpret_cast<__half2*>(&sV[l][2 * c2]) = *reinterpret_cast<const __half2*>(base + l * ld + 2 * D + 2 * c2);
  }
  __syncthreads();
  for (int i = warp; i < L; i += 4) {
    *reinterpret_cast<__half2*>(&sQ[warp][2 * lane]) = *reinterpret_cast<const __half2*>(base + i * ld + 2 * lane);
    __syncwarp();
    float sc[3];
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < 3; ++jj) {
      const int j = lane + 32 * jj;
      float a = -INFINITY;
      if (j < L && !(causal && j > i)) {
        a = 0.f;
#pragma unroll
        for (int d2 = 0; d2 < 32; ++d2) {
          const float2 q2 = __half22float2(*reinterpret_cast<const __half2*>(&sQ[warp][2 * d2]));
          const float2 k2 = __half22float2(*reinterpret_cast<const __half2*>(&sK[j][2 * d2]));
          a += q2.x * k2.x + q2.y * k2.y;
        }
      }
      sc[jj] = a;
      mx = fmaxf(mx, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < 3; ++jj) {
      sc[jj] = sc[jj] == -INFINITY ? 0.f : expf(sc[jj] - mx);
      sum += sc[jj];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
#pragma unroll
    for (int jj = 0; jj < 3; ++jj) {
      const int j = lane + 32 * jj;
      if (j < L) sP[warp][j] = sc[jj] * inv;
    }
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    const int jmax = causal ? i + 1 : L;
    for (int j = 0; j < jmax; ++j) {
      const float pj = sP[warp][j];
      const float2 v2 = __half22float2(*reinterpret_cast<const __half2*>(&sV[j][2 * lane]));
      o0 += pj * v2.x;
      o1 += pj * v2.y;
    }
    *reinterpret_cast<__half2*>(out + ((size_t)s * L + i) * D + (size_t)h * 64 + 2 * lane) = __floats2half2_rn(o0, o1);
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// The same attention on the tensor cores.  One CTA = one head of a 128-row tile of the qkv matrix, i.e. floor(128 / L)
// whole sequences (2 images of 50 tokens, or 1 prompt of 77):
//   TMA: Q, K, V head slices [128 rows][64] (128-B rows, 128B swizzle) -> smem
//   S = Q K^T            tcgen05.mma 128 x 128 x 64 (both operands K-major)            -> TMEM, 128 fp32 columns
//   P = exp(S - rowmax)  thread <-> row; keys outside the row's own sequence (and, causal, after it) are exactly 0;
//                        written fp16 to smem in the K-major swizzled layout the next MMA reads
//   O = P V              tcgen05.mma 128 x 64 x 128; V is the MN-major operand (keys are its K dimension)  -> TMEM
//   out = O / rowsum     fp16, 128 B per row
// The cross-sequence quarter of S is computed and discarded: the tensor pipe is ~30x faster than the CUDA-core loop above,
// so the waste is free.  49 KB smem and 128 TMEM columns per CTA: four CTAs per SM hide the load -> MMA -> softmax -> MMA chain.
// ------------------------------------------------------------------------------------------------
struct AttnTcParams {
  int rows;        // S * L rows of qkv
  int L, D;
  int nseq;        // sequences per tile = 128 / L
  int causal;
  __half* out;     // [rows][D]
};
constexpr int kAttnTcSmem = 1024 + 3 * 16384 + 64;

__global__ void __launch_bounds__(128, 4)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // P (128 x 128 fp16, 32 KB) overwrites Q and K, which are dead once S = Q K^T has been accumulated; O likewise reuses the
  // first 64 TMEM columns of S after every thread has read its S row: 49 KB smem + 128 TMEM columns -> four CTAs per SM
  const uint32_t sQ = base, sK = base + 16384, sV = base + 32768, sP = base;
  const uint32_t bar_load = base + 49152, bar_s = bar_load + 8, bar_o = bar_load + 16, tmem_slot = bar_load + 24;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, h = blockIdx.y;
  const int r0 = tile * p.nseq * p.L;
  const int rows_valid = min(p.nseq * p.L, p.rows - r0);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(bar_load, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<128>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t tS = tmem_base, tO = tmem_base;

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_load, 3 * 16384);
    tma_load_2d(&tmQKV, bar_load, sQ, h * 64, r0);
    tma_load_2d(&tmQKV, bar_load, sK, p.D + h * 64, r0);
    tma_load_2d(&tmQKV, bar_load, sV, 2 * p.D + h * 64, r0);
    mbar_wait(bar_load, 0);
    tcgen05_fence_after();
    constexpr uint32_t idesc = make_idesc_f16_f32(128, 128);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_f16_ss(tS, make_kmajor_desc<128>(sQ + 32 * k), make_kmajor_desc<128>(sK + 32 * k), idesc, k != 0);
    umma_commit(bar_s);
  }
  __syncwarp();

  // ---- softmax: thread <-> row
  const int r = warp * 32 + lane;
  const int seq = r / p.L;
  const int j0 = seq * p.L;
  int j1 = j0 + p.L;
  if (p.causal) j1 = min(j1, r + 1);
  const bool live = r < rows_valid;
  mbar_wait(bar_s, 0);
  tcgen05_fence_after();
  const uint32_t trow = tS + (uint32_t(warp * 32) << 16);
  float mx = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
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
  for (int c = 0; c < 4; ++c) {
    uint32_t v[32];
    tmem_ld_32x32b<32>(trow + uint32_t(32 * c), v);
    tmem_ld_wait();
    uint32_t h2[16];
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const int j = 32 * c + i;
      const float a = (live && j >= j0 && j < j1) ? __expf(__uint_as_float(v[i]) - mx) : 0.f;
      const float b = (live && j + 1 >= j0 && j + 1 < j1) ? __expf(__uint_as_float(v[i + 1]) - mx) : 0.f;
      sum += a + b;
      h2[i >> 1] = pack_half2(a, b);
    }
    // 32 keys = 4 pieces of 16 B in chunk (32c / 64), pieces ((32c % 64) / 8) .. +3
    const uint32_t chunk = sP + uint32_t((32 * c) / 64) * 16384u;
    const int piece0 = ((32 * c) % 64) / 8;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t a = chunk + swizzle_off<128>(uint32_t(r), uint32_t(piece0 + q));
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(h2[4 * q]), "r"(h2[4 * q + 1]), "r"(h2[4 * q + 2]), "r"(h2[4 * q + 3]) : "memory");
    }
  }
  fence_proxy_async_smem();                        // P (generic-proxy stores) -> visible to the tensor core's smem reads
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    tcgen05_fence_after();
    constexpr uint32_t idesc = make_idesc_f16_f32(128, 64) | (1u << 16);      // B (= V) is MN-major
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {               // 16 keys per MMA
      const uint64_t da = make_kmajor_desc<128>(sP + uint32_t(ks / 4) * 16384u + 32u * uint32_t(ks % 4));
      const uint64_t db = make_mnmajor_desc<128>(sV + uint32_t(ks) * 2048u, 16384u);
      umma_f16_ss(tO, da, db, idesc, ks != 0);
    }
    umma_commit(bar_o);
  }
  __syncwarp();
  mbar_wait(bar_o, 0);
  tcgen05_fence_after();
  {
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    __half* orow = p.out + (size_t)(r0 + r) * p.D + (size_t)h * 64;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b<32>(tO + (uint32_t(warp * 32) << 16) + uint32_t(32 * c), v);
      tmem_ld_wait();
      if (live) {
        uint4* o4 = reinterpret_cast<uint4*>(orow + 32 * c);
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
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}

// logits[b][k] = exp(logit_scale) * <img_b, txt_k> / (|img_b| |txt_k|);  one warp per (b, k)
__global__ void __launch_bounds__(256)
clip_logits_kernel(const float* __restrict__ img, const float* __restrict__ txt, float* __restrict__ logits, int B, int K, int E,
                   float logit_scale_exp) {
  const int idx = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (idx >= B * K) return;
  const int b = idx / K, k = idx - b * K;
  float dot = 0.f, ni = 0.f, nt = 0.f;
  for (int e = lane; e < E; e += 32) {
    const float a = img[(size_t)b * E + e], t = txt[(size_t)k * E + e];
    dot += a * t; ni += a * a; nt += t * t;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dot += __shfl_xor_sync(0xffffffffu, dot, o);
    ni += __shfl_xor_sync(0xffffffffu, ni, o);
    nt += __shfl_xor_sync(0xffffffffu, nt, o);
  }
  if (lane == 0) logits[idx] = logit_scale_exp * dot / (sqrtf(ni) * sqrtf(nt));
}

}  // namespace embclip
