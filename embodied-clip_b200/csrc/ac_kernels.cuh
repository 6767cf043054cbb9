// Non-GEMM kernels of the actor-critic / PPO-update path (HBM- or latency-bound CUDA-core work, all fp32 math):
//   ac_pack_features      fp32 NCHW rollout features -> fp16 [frames*P, C] rows (once per rollout)
//   ac_goal_rows          goal-embedding rows broadcast over the P pixels of a frame (fp16)
//   pack_w                fp32 master weights -> fp16 GEMM layouts (cast / transpose / flatten permutation)
//   colsum_f16, goal_grad bias and embedding gradients
//   amax / scale_cast     device-side loss scale for the fp16 gradient GEMMs + bias column sums
//   ac_heads_fwd / bwd    LinearActorHead + LinearCriticHead, CategoricalDistr, clipped PPO loss and its gradient
//   gae, adv_norm         RolloutStorage.compute_returns / advantage normalisation
//   sumsq, adam_clip      clip_grad_norm_ + Adam on the flat parameter buffer
// Reference semantics: allenai/allenact v0.5.0, restated in oracle/allenact_models.py.
#pragma once
#include "ptx.cuh"

namespace embclip {

// ------------------------------------------------------------------------------------------------
// x [F][C][P] fp32  ->  y [F*P][C] fp16.  One block = 128 channels of one frame: the [128][P] slab is contiguous.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ac_pack_features_kernel(const float* __restrict__ x, __half* __restrict__ y, int C, int P) {
  extern __shared__ float tile[];               // [128][P + 1]
  const int f = blockIdx.y, c0 = blockIdx.x * 128;
  const int cn = min(128, C - c0);
  const float* src = x + ((size_t)f * C + c0) * P;
  const int pitch = P + 1;
  for (int i = threadIdx.x; i < cn * P; i += blockDim.x) {
    const int c = i / P, p = i - c * P;
    tile[c * pitch + p] = __ldcs(src + i);
  }
  __syncthreads();
  __half* dst = y + (size_t)f * P * C + c0;
  const int pairs = cn >> 1;
  for (int i = threadIdx.x; i < P * pairs; i += blockDim.x) {
    const int p = i / pairs, c2 = (i - p * pairs) * 2;
    const __half2 h = __floats2half2_rn(tile[c2 * pitch + p], tile[(c2 + 1) * pitch + p]);
    *reinterpret_cast<__half2*>(dst + (size_t)p * C + c2) = h;
  }
}

// rows[(f*P + p)][c] = fp16(embed[goal[f]][c])
__global__ void __launch_bounds__(256)
ac_goal_rows_kernel(const float* __restrict__ embed, const long long* __restrict__ goals, __half* __restrict__ rows,
                    int F, int P, int D, int num_goals) {
  const long long total = (long long)F * P * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = int(i % D);
    const long long fp = i / D;
    const int f = int(fp / P);
    long long g = goals[f];
    g = g < 0 ? 0 : (g >= num_goals ? num_goals - 1 : g);
    rows[i] = __float2half_rn(embed[g * D + c]);
  }
}

// dEmbed[goal[f]][c] += alpha * sum_p dG[(f*P + p)][c]
__global__ void __launch_bounds__(256)
ac_goal_grad_kernel(const __half* __restrict__ dg, const long long* __restrict__ goals, float* __restrict__ dembed,
                    int F, int P, int D, int num_goals, const float* __restrict__ alpha) {
  const float a = alpha ? __ldg(alpha) : 1.f;
  const long long total = (long long)F * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = int(i % D);
    const int f = int(i / D);
    float s = 0.f;
    for (int p = 0; p < P; ++p) s += __half2float(dg[((size_t)f * P + p) * D + c]);
    long long g = goals[f];
    g = g < 0 ? 0 : (g >= num_goals ? num_goals - 1 : g);
    atomicAdd(dembed + g * D + c, s * a);
  }
}

// fp32 [R][Cc] -> fp16.  mode 0: dst[r][c];  1: dst[c][r];  2: dst[r][p*CG+g] = src[r][g*P+p] (Cc = CG*P);
// 3: mode 2 transposed, dst[(p*CG+g)][r].
__global__ void __launch_bounds__(256)
pack_w_kernel(const float* __restrict__ src, __half* __restrict__ dst, int R, int Cc, int mode, int P, int CG) {
  const long long total = (long long)R * Cc;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int r, c;                                   // destination-major enumeration keeps the writes coalesced
    if (mode == 0 || mode == 2) { r = int(i / Cc); c = int(i % Cc); }
    else { c = int(i / R); r = int(i % R); }
    int cs = c;
    if (mode >= 2) { const int p = c / CG, g = c - p * CG; cs = g * P + p; }
    dst[i] = __float2half_rn(src[(size_t)r * Cc + cs]);
  }
}

// out[c] += alpha * sum_r x[r][c];  x fp16 [R][C] (C <= 2048, even).  One block = 256 rows.
__global__ void __launch_bounds__(256)
colsum_f16_kernel(const __half* __restrict__ x, float* __restrict__ out, long long R, int C, const float* __restrict__ alpha) {
  const float a = alpha ? __ldg(alpha) : 1.f;
  const long long r0 = (long long)blockIdx.x * 256;
  const long long r1 = r0 + 256 < R ? r0 + 256 : R;
  for (int c2 = threadIdx.x; c2 < C / 2; c2 += blockDim.x) {
    float s0 = 0.f, s1 = 0.f;
    for (long long r = r0; r < r1; ++r) {
      const float2 v = __half22float2(*reinterpret_cast<const __half2*>(x + r * C + 2 * c2));
      s0 += v.x; s1 += v.y;
    }
    atomicAdd(out + 2 * c2, s0 * a);
    atomicAdd(out + 2 * c2 + 1, s1 * a);
  }
}

// ------------------------------------------------------------------------------------------------
// Loss scale: amax (float bits, non-negative) -> scale[0] = S = 2^k with S * amax in [512, 1024], scale[1] = 1/S.
// ------------------------------------------------------------------------------------------------
__global__ void compute_scale_kernel(const unsigned int* __restrict__ amax_bits, float* __restrict__ scale) {
  const float amax = __uint_as_float(*amax_bits);
  float s = 1.f;
  if (amax > 0.f && isfinite(amax)) {
    int e;
    frexpf(amax, &e);                           // amax = m * 2^e, m in [0.5, 1)
    int k = 10 - e;                             // S * amax = m * 2^10
    k = k < -24 ? -24 : (k > 40 ? 40 : k);
    s = ldexpf(1.f, k);
  }
  scale[0] = s;
  scale[1] = 1.f / s;
}

// dst = fp16(src * scale[0]);  colsum[c] += sum_r src[r][c] (unscaled, fp32).  One block = 64 rows.
__global__ void __launch_bounds__(256)
scale_cast_kernel(const float* __restrict__ src, __half* __restrict__ dst, long long R, int C,
                  const float* __restrict__ scale, float* __restrict__ colsum) {
  const float s = scale ? __ldg(scale) : 1.f;
  const long long r0 = (long long)blockIdx.x * 64;
  const long long r1 = r0 + 64 < R ? r0 + 64 : R;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (long long r = r0; r < r1; ++r) {
      const float v = src[r * C + c];
      acc += v;
      dst[r * C + c] = __float2half_rn(v * s);
    }
    if (colsum) atomicAdd(colsum + c, acc);
  }
}

// ------------------------------------------------------------------------------------------------
// Heads.  One warp per row of h [F][H] (H multiple of 128).  W [A+1][H]: actor rows then the critic row.
//   logits[f][a], values[f].  loss != 0 additionally evaluates CategoricalDistr + PPO.loss_per_step and writes
//   dlogits [F][A], dvalues [F] (gradient of the total loss, already multiplied by grad_scale / count).
// loss_out: [0] sum action_loss, [1] sum value_loss, [2] sum entropy  (means are taken on the host side).
// ------------------------------------------------------------------------------------------------
constexpr int kMaxActions = 16;

struct HeadsParams {
  const float* h;            // [F][H]
  const float* w_actor;      // [A][H]
  const float* b_actor;      // [A]
  const float* w_critic;     // [H]
  const float* b_critic;     // [1]
  float* logits;             // [F][A]
  float* values;             // [F]
  int F, H, A;
  int loss;
  const long long* actions;  // [F]
  const float* old_logp;     // [F]
  const float* adv;          // [F] normalised advantages
  const float* old_values;   // [F]
  const float* returns;      // [F]
  float clip, vcoef, ecoef;
  float grad_scale;          // 1 / (global row count): d(mean)/d(row)
  float* dlogits;            // [F][A]
  float* dvalues;            // [F]
  float* loss_out;           // [3]
  // act(): CategoricalDistr(logits).sample() + log_prob(sample) (allenact distributions.py [UPSTREAM]) for every row in the same
  // launch -- inverse-CDF over exp(logit - max) with the caller's
  // uniforms, log-prob of the sampled action (one launch for the softmax -> multinomial -> log_softmax -> gather chain)
  const float* uniforms;     // [F] or nullptr
  long long* sampled;        // [F]
  float* sampled_logp;       // [F]
};

__global__ void __launch_bounds__(256)
ac_heads_fwd_kernel(const HeadsParams p) {
  extern __shared__ float sw[];                  // [(A+1)][H]
  const int rows_w = p.A + 1;
  for (int i = threadIdx.x; i < p.A * p.H; i += blockDim.x) sw[i] = p.w_actor[i];
  for (int i = threadIdx.x; i < p.H; i += blockDim.x) sw[p.A * p.H + i] = p.w_critic[i];
  __shared__ float sloss[3];
  if (threadIdx.x < 3) sloss[threadIdx.x] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float l_act = 0.f, l_val = 0.f, l_ent = 0.f;
  for (int f = blockIdx.x * 8 + warp; f < p.F; f += gridDim.x * 8) {
    float acc[kMaxActions + 1];
#pragma unroll
    for (int j = 0; j <= kMaxActions; ++j) acc[j] = 0.f;
    const float4* hr = reinterpret_cast<const float4*>(p.h + (size_t)f * p.H);
    for (int i = lane; i < p.H / 4; i += 32) {
      const float4 hv = hr[i];
#pragma unroll
      for (int j = 0; j <= kMaxActions; ++j) {
        if (j < rows_w) {
          const float4 wv = reinterpret_cast<const float4*>(sw + j * p.H)[i];
          acc[j] += hv.x * wv.x + hv.y * wv.y + hv.z * wv.z + hv.w * wv.w;
        }
      }
    }
#pragma unroll
    for (int j = 0; j <= kMaxActions; ++j) {
      if (j < rows_w) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
      }
    }
    if (lane == 0) {
      float lg[kMaxActions];
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < kMaxActions; ++j) {
        if (j < p.A) {
          lg[j] = acc[j] + p.b_actor[j];
          p.logits[(size_t)f * p.A + j] = lg[j];
          mx = fmaxf(mx, lg[j]);
        }
      }
      float v = p.b_critic[0];
#pragma unroll
      for (int j = 0; j <= kMaxActions; ++j) if (j == p.A) v += acc[j];
      p.values[f] = v;
      if (p.uniforms) {
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxActions; ++j) if (j < p.A) sum += __expf(lg[j] - mx);
        const float target = p.uniforms[f] * sum;
        float cdf = 0.f, lg_a = 0.f;
        int a = -1, last = 0;
#pragma unroll
        for (int j = 0; j < kMaxActions; ++j) {
          if (j < p.A) {
            const float e = __expf(lg[j] - mx);
            if (e > 0.f) last = j;
            cdf += e;
            if (a < 0 && cdf > target) a = j;
          }
        }
        if (a < 0) a = last;                              // u * sum rounded up to the total: the last action with mass
#pragma unroll
        for (int j = 0; j < kMaxActions; ++j) if (j == a) lg_a = lg[j];
        p.sampled[f] = a;
        p.sampled_logp[f] = lg_a - mx - logf(sum);
      }
      if (p.loss) {
        float se = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxActions; ++j) if (j < p.A) se += expf(lg[j] - mx);
        const float lse = mx + logf(se);
        const int a = int(p.actions[f]);
        float ent = 0.f, lp_a = 0.f;
        float pr[kMaxActions], lp[kMaxActions];
#pragma unroll
        for (int j = 0; j < kMaxActions; ++j) {
          if (j < p.A) {
            lp[j] = lg[j] - lse;
            pr[j] = expf(lp[j]);
            ent -= pr[j] * lp[j];
            if (j == a) lp_a = lp[j];
          }
        }
        const float A_ = p.adv[f];
        const float ratio = expf(lp_a - p.old_logp[f]);
        const float clamped = fminf(fmaxf(ratio, 1.f - p.clip), 1.f + p.clip);
        const float s1 = ratio * A_, s2 = clamped * A_;
        const bool pick2 = s2 < s1;
        l_act += -(pick2 ? s2 : s1);
        // d(-min)/d lp_a : through ratio when surr1 is picked, or surr2 is picked inside the clamp range
        const bool inside = ratio >= 1.f - p.clip && ratio <= 1.f + p.clip;
        const float dlpa = (!pick2 || inside) ? -A_ * ratio : 0.f;
        const float vo = p.old_values[f], R = p.returns[f];
        const float dvv = v - vo;
        const float vclip = vo + fminf(fmaxf(dvv, -p.clip), p.clip);
        const float e1 = (v - R) * (v - R), e2 = (vclip - R) * (vclip - R);
        l_val += 0.5f * fmaxf(e1, e2);
        const bool vin = dvv >= -p.clip && dvv <= p.clip;
        float dv;                                 // d(0.5 * max(e1, e2)) / dv; torch.max splits ties evenly
        if (e1 > e2) dv = (v - R);
        else if (e1 < e2) dv = vin ? (vclip - R) : 0.f;
        else dv = 0.5f * (v - R) + (vin ? 0.5f * (vclip - R) : 0.f);
        l_ent += ent;
        const float gs = p.grad_scale;
        p.dvalues[f] = p.vcoef * dv * gs;
#pragma unroll
        for (int j = 0; j < kMaxActions; ++j) {
          if (j < p.A) {
            // total = action + vcoef * value + ecoef * (-H);  d(-H)/dz_j = p_j (log p_j + H)
            const float g = dlpa * ((j == a ? 1.f : 0.f) - pr[j]) + p.ecoef * pr[j] * (lp[j] + ent);
            p.dlogits[(size_t)f * p.A + j] = g * gs;
          }
        }
      }
    }
  }
  if (p.loss) {
    if (lane == 0) { atomicAdd(&sloss[0], l_act); atomicAdd(&sloss[1], l_val); atomicAdd(&sloss[2], l_ent); }
    __syncthreads();
    if (threadIdx.x < 3) atomicAdd(p.loss_out + threadIdx.x, sloss[threadIdx.x]);
  }
}

// dh[f][c] = sum_j dlogits[f][j] Wa[j][c] + dvalues[f] Wc[c];  dWa, dWc, dba, dbc accumulated (atomics).
// One block = 64 rows; thread = columns {tid, tid + 256, ...}.
struct HeadsBwdParams {
  const float* h; const float* w_actor; const float* w_critic;
  const float* dlogits; const float* dvalues;
  float* dh;
  float* dw_actor; float* db_actor; float* dw_critic; float* db_critic;
  int F, H, A;
};
__global__ void __launch_bounds__(256)
ac_heads_bwd_kernel(const HeadsBwdParams p) {
  __shared__ float sd[64][kMaxActions + 1];
  const int f0 = blockIdx.x * 64;
  const int nf = min(64, p.F - f0);
  for (int i = threadIdx.x; i < nf * (p.A + 1); i += blockDim.x) {
    const int r = i / (p.A + 1), j = i - r * (p.A + 1);
    sd[r][j] = j < p.A ? p.dlogits[(size_t)(f0 + r) * p.A + j] : p.dvalues[f0 + r];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < p.H; c += blockDim.x) {
    float w[kMaxActions + 1], g[kMaxActions + 1];
#pragma unroll
    for (int j = 0; j <= kMaxActions; ++j) {
      g[j] = 0.f;
      w[j] = j < p.A ? p.w_actor[(size_t)j * p.H + c] : (j == p.A ? p.w_critic[c] : 0.f);
    }
    for (int r = 0; r < nf; ++r) {
      const float hv = p.h[(size_t)(f0 + r) * p.H + c];
      float d = 0.f;
#pragma unroll
      for (int j = 0; j <= kMaxActions; ++j) {
        if (j <= p.A) {
          const float s = sd[r][j];
          d += s * w[j];
          g[j] += s * hv;
        }
      }
      p.dh[(size_t)(f0 + r) * p.H + c] = d;
    }
#pragma unroll
    for (int j = 0; j <= kMaxActions; ++j) {
      if (j < p.A) atomicAdd(p.dw_actor + (size_t)j * p.H + c, g[j]);
      else if (j == p.A) atomicAdd(p.dw_critic + c, g[j]);
    }
  }
  if (threadIdx.x <= p.A) {
    float s = 0.f;
    for (int r = 0; r < nf; ++r) s += sd[r][threadIdx.x];
    if (threadIdx.x < p.A) atomicAdd(p.db_actor + threadIdx.x, s);
    else atomicAdd(p.db_critic, s);
  }
}

// ------------------------------------------------------------------------------------------------
// RolloutStorage.compute_returns (use_gae): one thread per sampler, reverse scan over T.
//   rewards [T][N], values [T+1][N] (row T = next_value), masks [T+1][N]  ->  returns [T][N], adv [T][N] = returns - values
// ------------------------------------------------------------------------------------------------
__global__ void gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values, const float* __restrict__ masks,
                           float* __restrict__ returns, float* __restrict__ adv, int T, int N, float gamma, float tau) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float gae = 0.f;
  for (int t = T - 1; t >= 0; --t) {
    const float v = values[(size_t)t * N + n], vn = values[(size_t)(t + 1) * N + n], m = masks[(size_t)(t + 1) * N + n];
    const float delta = rewards[(size_t)t * N + n] + gamma * vn * m - v;
    gae = delta + gamma * tau * m * gae;
    returns[(size_t)t * N + n] = gae + v;
    adv[(size_t)t * N + n] = gae;                // (gae + v) - v, without the rounding of the round trip
  }
}

// adv <- (adv - mean) / (std + eps), unbiased std (torch.Tensor.std default).  Single block.
__global__ void __launch_bounds__(1024)
adv_norm_kernel(const float* __restrict__ ret, const float* __restrict__ val, float* __restrict__ out, int n, float eps) {
  __shared__ double sh[32];
  __shared__ double s_mean, s_inv;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double s = 0.0;
  for (int i = tid; i < n; i += blockDim.x) s += double(ret[i] - val[i]);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) sh[warp] = s;
  __syncthreads();
  if (tid == 0) { double t = 0; for (int i = 0; i < (blockDim.x >> 5); ++i) t += sh[i]; s_mean = t / n; }
  __syncthreads();
  const double mean = s_mean;
  double q = 0.0;
  for (int i = tid; i < n; i += blockDim.x) { const double d = double(ret[i] - val[i]) - mean; q += d * d; }
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  __syncthreads();
  if (lane == 0) sh[warp] = q;
  __syncthreads();
  if (tid == 0) { double t = 0; for (int i = 0; i < (blockDim.x >> 5); ++i) t += sh[i]; s_inv = 1.0 / (sqrt(t / (n > 1 ? n - 1 : 1)) + double(eps)); }
  __syncthreads();
  const float m = float(mean), inv = float(s_inv);
  for (int i = tid; i < n; i += blockDim.x) out[i] = ((ret[i] - val[i]) - m) * inv;
}

// ------------------------------------------------------------------------------------------------
// clip_grad_norm_ + Adam on flat fp32 buffers.
// ------------------------------------------------------------------------------------------------
// DETERMINISTIC: every rank of a data-parallel job must turn the (bit-identical, all-reduced) gradient into the bit-identical
// clip coefficient, or the replicas' parameters drift apart ulp by ulp (measured on 2 GPUs with the first version, which
// combined the block partials with atomicAdd in arrival order).  Each block writes its partial to scratch[blockIdx]; the block
// that finishes last (a counter) adds the partials in index order.  out: [0] result, [1] counter (zeroed by the launcher),
// [2 ..] one partial per block.
constexpr int kSumsqMaxBlocks = 1022;
__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += g[i] * g[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float sh[8];
  __shared__ bool last;
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sh[i];
    out[2 + blockIdx.x] = t;
    __threadfence();
    last = atomicAdd(reinterpret_cast<unsigned int*>(out + 1), 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    __threadfence();
    float t = 0.f;
    for (int i = threadIdx.x; i < int(gridDim.x); i += 256) t += __ldcg(out + 2 + i);      // fixed assignment of partials to threads
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      float r = 0.f;
      for (int i = 0; i < 8; ++i) r += sh[i];
      out[0] = r;
    }
  }
}

// torch.nn.utils.clip_grad_norm_(max_norm): coef = min(1, max_norm / (norm + 1e-6));
// torch.optim.Adam (no amsgrad / weight decay): step_size = lr / bc1; denom = sqrt(v) / sqrt(bc2) + eps.
__global__ void __launch_bounds__(256)
adam_clip_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                 const float* __restrict__ sumsq, float max_norm, float lr, float beta1, float beta2, float eps,
                 float bc1, float bc2_sqrt) {
  float coef = 1.f;
  if (max_norm > 0.f) {
    coef = max_norm / (sqrtf(*sumsq) + 1e-6f);
    coef = coef < 1.f ? coef : 1.f;
  }
  const float step_size = lr / bc1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    g[i] = gi;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  }
}


}  // namespace embclip
