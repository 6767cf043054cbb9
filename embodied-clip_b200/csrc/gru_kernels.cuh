// RNNStateEncoder (allenact basic_models; 1-layer nn.GRU with episode masking) as two persistent cooperative
// kernels, forward and BPTT.  fp32 throughout (the recurrence feeds itself 128 times: no fp16 here).
//
//   hm_t = h_{t-1} * mask_t  (+ (1 - mask_t) * h_init)       (mask 0 = episode start, applied IN the kernel:
//   gh   = W_hh hm_t + b_hh                                    upstream splits the sequence on the host instead;
//                                                              h_init = RNNStateEncoder(trainable_masked_hidden_state=True))
//   r = sigmoid(gi_r + gh_r)   z = sigmoid(gi_z + gh_z)   n = tanh(gi_n + r * gh_n)
//   h_t = (1 - z) * n + z * hm_t
//
// gi = W_ih x + b_ih for all T steps is one tensor-core GEMM done beforehand; only the recurrent half is serial.
// Decomposition: CTA = (block of 8 hidden units, group of <= 32 samplers).  Its slice of W_hh (24 gate rows in
// the forward, 8 columns in the backward) stays in shared memory for all T steps; per step the CTAs exchange h_t
// (forward) / dgh_t (backward) through L2 and meet at a counter barrier shared by the CTAs of one sampler group.
// All CTAs must be co-resident: launched with cudaLaunchCooperativeKernel, grid = (H/8) * groups <= #SMs.
#pragma once
#include "ptx.cuh"

namespace embclip {

constexpr int kGruUB = 8;        // hidden units per CTA
constexpr int kGruNS = 32;       // samplers per CTA (max)
constexpr int kGruThreads = 256;

struct GruFwdParams {
  int T, N, H;
  int groups;                // sampler groups; group g holds samplers [g*ns, min(N, (g+1)*ns))
  int ns;                    // samplers per group (<= 32)
  const float* gi;           // [T][N][3H]
  const float* w_hh;         // [3H][H]
  const float* b_hh;         // [3H]
  const float* h0;           // [N][H]
  const float* masks;        // [T][N]
  const float* h_init;       // [H] learned state an episode starts from, or null (zeros)
  float* out;                // [T][N][H]
  float* r; float* z; float* n; float* hn;   // [T][N][H] each, or all null (inference)
  unsigned int* bar;         // [groups], zeroed before launch
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Counter barrier among the `count` CTAs of a group; `target` = count * (number of barriers passed so far + 1).
__device__ __forceinline__ void group_barrier(unsigned int* bar, unsigned int target) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(bar, 1u);
    const uint64_t t0 = global_timer_ns();
    uint32_t spins = 0;
    while (ld_acquire_u32(bar) < target) {
      if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > 4000000000ull) {
        printf("embclip: gru group barrier timeout (block %d target %u)\n", blockIdx.x, target);
        __trap();
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// smem: W slice [24][H+4], hm tile [32][H+4], partials [8 k-slices][32 s][24 rows]
__global__ void __launch_bounds__(kGruThreads, 1)
gru_forward_kernel(const GruFwdParams p) {
  extern __shared__ float smem_f[];
  const int H = p.H, HP = H + 4;
  float* sW = smem_f;                               // [24][HP]
  float* sH = sW + 3 * kGruUB * HP;                 // [32][HP]
  float* sP = sH + kGruNS * HP;                     // [8][32][24]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = blockIdx.x % p.groups;
  const int ub = blockIdx.x / p.groups;
  const int u0 = ub * kGruUB;
  const int s0 = grp * p.ns;
  const int ns = min(p.ns, p.N - s0);
  const unsigned int ctas_per_group = gridDim.x / p.groups;

  // resident W_hh slice: local row lr = gate * 8 + u  <->  global row gate * H + u0 + u
  for (int i = tid; i < 3 * kGruUB * (H / 4); i += kGruThreads) {
    const int lr = i / (H / 4), k4 = i - lr * (H / 4);
    const int grow = (lr / kGruUB) * H + u0 + (lr % kGruUB);
    *reinterpret_cast<float4*>(sW + lr * HP + 4 * k4) = __ldg(reinterpret_cast<const float4*>(p.w_hh + (size_t)grow * H) + k4);
  }
  // rows of the h tile beyond ns stay zero
  for (int i = tid; i < kGruNS * HP; i += kGruThreads) sH[i] = 0.f;
  __syncthreads();

  // phase-2 role: thread (s = tid / 8, u = tid % 8) owns one hidden unit of one sampler
  const int es = tid >> 3, eu = tid & 7;
  const float bh_r = p.b_hh[u0 + eu], bh_z = p.b_hh[H + u0 + eu], bh_n = p.b_hh[2 * H + u0 + eu];
  // phase-1 role: warp = k-slice of H/8, lane = (sampler group of 4, row group of 6)
  const int sg4 = lane >> 2, rg = lane & 3;
  const int kslice = H / 8;

  for (int t = 0; t < p.T; ++t) {
    // ---- masked previous hidden state of this group's samplers -> smem
    const float* hprev = t == 0 ? p.h0 : p.out + (size_t)(t - 1) * p.N * H;
    for (int i = tid; i < ns * (H / 4); i += kGruThreads) {
      const int s = i / (H / 4), k4 = i - s * (H / 4);
      const float m = p.masks[(size_t)t * p.N + s0 + s];
      float4 v = __ldcg(reinterpret_cast<const float4*>(hprev + (size_t)(s0 + s) * H) + k4);
      v.x *= m; v.y *= m; v.z *= m; v.w *= m;
      if (p.h_init) {
        const float4 hi = __ldg(reinterpret_cast<const float4*>(p.h_init) + k4);
        const float om = 1.f - m;
        v.x += om * hi.x; v.y += om * hi.y; v.z += om * hi.z; v.w += om * hi.w;
      }
      *reinterpret_cast<float4*>(sH + s * HP + 4 * k4) = v;
    }
    __syncthreads();
    // ---- partial gh over this warp's k-slice: 4 samplers x 6 rows per thread
    float acc[4][6];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 6; ++b) acc[a][b] = 0.f;
    const float* hb = sH + (sg4 * 4) * HP + warp * kslice;
    const float* wb = sW + (rg * 6) * HP + warp * kslice;
    for (int k = 0; k < kslice; k += 4) {
      float4 hv[4], wv[6];
#pragma unroll
      for (int a = 0; a < 4; ++a) hv[a] = *reinterpret_cast<const float4*>(hb + a * HP + k);
#pragma unroll
      for (int b = 0; b < 6; ++b) wv[b] = *reinterpret_cast<const float4*>(wb + b * HP + k);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 6; ++b)
          acc[a][b] += hv[a].x * wv[b].x + hv[a].y * wv[b].y + hv[a].z * wv[b].z + hv[a].w * wv[b].w;
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 6; ++b) sP[(warp * kGruNS + sg4 * 4 + a) * 24 + rg * 6 + b] = acc[a][b];
    __syncthreads();
    // ---- gates for (sampler es, unit eu)
    if (es < ns) {
      float gr = bh_r, gz = bh_z, gn = bh_n;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const float* pp = sP + (w * kGruNS + es) * 24;
        gr += pp[eu]; gz += pp[kGruUB + eu]; gn += pp[2 * kGruUB + eu];
      }
      const size_t row = (size_t)t * p.N + s0 + es;
      const float* gi = p.gi + row * 3 * H + u0 + eu;
      const float r = sigmoidf_(gi[0] + gr);
      const float z = sigmoidf_(gi[H] + gz);
      const float n = tanhf(gi[2 * H] + r * gn);
      const float hm = sH[es * HP + u0 + eu];
      const float h = (1.f - z) * n + z * hm;
      const size_t o = row * H + u0 + eu;
      p.out[o] = h;
      if (p.r) { p.r[o] = r; p.z[o] = z; p.n[o] = n; p.hn[o] = gn; }
    }
    if (t + 1 < p.T) group_barrier(p.bar + grp, ctas_per_group * (unsigned)(t + 1));
  }
}

struct GruBwdParams {
  int T, N, H;
  int groups, ns;
  const float* w_hh;         // [3H][H]
  const float* h0;           // [N][H]
  const float* masks;        // [T][N]
  const float* out;          // [T][N][H]
  const float* r; const float* z; const float* n; const float* hn;
  const float* dout;         // [T][N][H] gradient w.r.t. the GRU outputs
  const float* dhT;          // [N][H] gradient w.r.t. the final hidden state, or null
  const float* h_init;       // [H] or null (see GruFwdParams)
  float* dh_init;            // [H] accumulated (atomicAdd) gradient of h_init, or null
  float* dgi;                // [T][N][3H]
  float* dgh;                // [T][N][3H]
  __half* hm_h;              // [T][N][H] masked previous hidden state, fp16 (operand of the dW_hh contraction)
  float* dh0;                // [N][H] or null
  unsigned int* amax;        // max |dgi| as float bits
  unsigned int* bar;         // [groups]
};

// smem: W^T slice [2][3H][4], dgh chunk tile [32][3H/2 + 4]
__global__ void __launch_bounds__(kGruThreads, 1)
gru_backward_kernel(const GruBwdParams p) {
  extern __shared__ float smem_f[];
  const int H = p.H, G3 = 3 * H, CH = G3 / 2, CP = CH + 4;
  float* sWt = smem_f;                              // [2][3H][4]: units 0..3, then units 4..7 (conflict-free float4 reads)
  float* sD = sWt + (size_t)G3 * kGruUB;            // [32][CP]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = blockIdx.x % p.groups;
  const int ub = blockIdx.x / p.groups;
  const int u0 = ub * kGruUB;
  const int s0 = grp * p.ns;
  const int ns = min(p.ns, p.N - s0);
  const unsigned int ctas_per_group = gridDim.x / p.groups;

  for (int i = tid; i < G3 * kGruUB; i += kGruThreads) {
    const int rho = i / kGruUB, u = i - rho * kGruUB;
    sWt[(size_t)(u >> 2) * G3 * 4 + rho * 4 + (u & 3)] = __ldg(p.w_hh + (size_t)rho * H + u0 + u);
  }
  for (int i = tid; i < kGruNS * CP; i += kGruThreads) sD[i] = 0.f;
  __syncthreads();

  // thread (es, eu): sampler es = tid / 8 (warp w holds samplers 4w..4w+3), unit eu = tid % 8
  const int es = tid >> 3, eu = tid & 7;
  const bool live = es < ns;
  float carry = (live && p.dhT) ? p.dhT[(size_t)(s0 + es) * H + u0 + eu] : 0.f;
  const float hinit = (live && p.h_init) ? p.h_init[u0 + eu] : 0.f;
  float dinit = 0.f;
  float amax = 0.f;
  unsigned int nbar = 0;

  for (int t = p.T - 1; t >= 0; --t) {
    float direct = 0.f, m = 0.f;
    if (live) {
      const size_t row = (size_t)t * p.N + s0 + es;
      const size_t o = row * H + u0 + eu;
      const float dh = p.dout[o] + carry;
      const float r = p.r[o], z = p.z[o], n = p.n[o], hn = p.hn[o];
      m = p.masks[row];
      const float hprev = t == 0 ? p.h0[(size_t)(s0 + es) * H + u0 + eu] : __ldcg(p.out + o - (size_t)p.N * H);
      const float hm = hprev * m + (1.f - m) * hinit;
      const float dn = dh * (1.f - z);
      const float dz = dh * (hm - n);
      direct = dh * z;
      const float dn_pre = dn * (1.f - n * n);
      const float dz_pre = dz * z * (1.f - z);
      const float dr_pre = dn_pre * hn * r * (1.f - r);
      float* gi = p.dgi + row * G3 + u0 + eu;
      float* gh = p.dgh + row * G3 + u0 + eu;
      gi[0] = dr_pre; gi[H] = dz_pre; gi[2 * H] = dn_pre;
      gh[0] = dr_pre; gh[H] = dz_pre; gh[2 * H] = dn_pre * r;
      p.hm_h[o] = __float2half_rn(hm);
      amax = fmaxf(amax, fmaxf(fabsf(dr_pre), fmaxf(fabsf(dz_pre), fabsf(dn_pre))));
    }
    // every CTA of this sampler group has published its columns of dgh[t]
    group_barrier(p.bar + grp, ctas_per_group * (++nbar));

    // dhm[s][u] = sum_rho dgh[t][s][rho] * W_hh[rho][u]: warp w <-> samplers 4w..4w+3, lanes split rho
    float acc[4][kGruUB];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < kGruUB; ++b) acc[a][b] = 0.f;
    for (int half = 0; half < 2; ++half) {
      __syncthreads();                                         // previous chunk fully consumed
      for (int i = tid; i < ns * (CH / 4); i += kGruThreads) {
        const int s = i / (CH / 4), k4 = i - s * (CH / 4);
        const float4 v = __ldcg(reinterpret_cast<const float4*>(p.dgh + ((size_t)t * p.N + s0 + s) * G3 + half * CH) + k4);
        *reinterpret_cast<float4*>(sD + s * CP + 4 * k4) = v;
      }
      __syncthreads();
      const float* db = sD + (warp * 4) * CP;
      const float* wt = sWt + (size_t)half * CH * 4;
      for (int rho = lane; rho < CH; rho += 32) {
        const float4 w0 = *reinterpret_cast<const float4*>(wt + rho * 4);
        const float4 w1 = *reinterpret_cast<const float4*>(wt + (size_t)G3 * 4 + rho * 4);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const float d = db[a * CP + rho];
          acc[a][0] += d * w0.x; acc[a][1] += d * w0.y; acc[a][2] += d * w0.z; acc[a][3] += d * w0.w;
          acc[a][4] += d * w1.x; acc[a][5] += d * w1.y; acc[a][6] += d * w1.z; acc[a][7] += d * w1.w;
        }
      }
    }
    // warp all-reduce of the 32 partial sums; lane (a*8 + b) keeps acc[a][b] = its own (sampler, unit)
    float mine = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < kGruUB; ++b) {
        float v = acc[a][b];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == a * kGruUB + b) mine = v;
      }
    dinit += (direct + mine) * (1.f - m);
    carry = (direct + mine) * m;
  }
  if (live && p.dh_init) atomicAdd(p.dh_init + u0 + eu, dinit);
  if (live && p.dh0) p.dh0[(size_t)(s0 + es) * H + u0 + eu] = carry;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if (lane == 0) atomicMax(p.amax, __float_as_uint(amax));
}

}  // namespace embclip
