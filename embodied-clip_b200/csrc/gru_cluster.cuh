// RNNStateEncoder (1-layer nn.GRU with episode masking) on THREAD-BLOCK CLUSTERS: the step-to-step exchange of the hidden
// state goes through distributed shared memory and one hardware cluster barrier instead of an L2 round trip plus a
// global-atomic grid barrier (gru_kernels.cuh: 10 us per forward step / 14 us per BPTT step at 60 samplers, of which the
// arithmetic is ~2 us -- profiles/r1i_ac_full.txt).
//
// Decomposition: one CLUSTER of CS = H / 32 CTAs (16 for H = 512: the non-portable maximum) owns a group of <= 8 samplers
// for all T steps; clusters never talk to each other, so there is no cooperative launch and no co-residency requirement.
// CTA c of a cluster owns hidden units [32c, 32c + 32): the 96 rows of W_hh that produce them (r, z, n gates), 96 x H fp32,
// stay in its shared memory for the whole sequence (192 KB at H = 512).
//
//   forward, step t    every CTA holds hm_t = mask_t * h_{t-1} (+ (1 - mask_t) * h_init) of its samplers for ALL H units in
//                      sH[t & 1]; it computes its 96 x ns gate pre-activations (fp32 FFMA, W and hm from shared memory, a
//                      lane-transposing shuffle reduction over the K split), the gates, h_t for its 32 units, and stores
//                      hm_{t+1} into sH[(t + 1) & 1] of EVERY CTA of the cluster (st.shared::cluster).  barrier.cluster.
//   backward, step t   every CTA forms dgh_t for its own 96 gate rows (so the matvec input needs no exchange), multiplies
//                      by its 96 x H slice of W_hh -- a PARTIAL dhm for all H units -- and scatters each 32-unit piece to the
//                      CTA that owns those units (reduce-scatter through DSMEM); after barrier.cluster the owner adds the
//                      CS partials in rank order (deterministic) to get the carry into step t - 1.
//
// On this B200 only 7 clusters of 16 CTAs fit at once (GPCs of 16 / 18 / 20 SMs, one of them short), so 60 samplers run as 7
// clusters of 9: with 192 KB of weights per CTA there is no room for a double-buffered exchange buffer at 9+ samplers.  Both
// kernels keep ONE buffer and cover the write-after-read hazard with a split cluster barrier (arrive when done reading, wait
// just before the remote stores) next to the full barrier that publishes the step's data.  fp32 throughout, same formulas and
// saved tensors as gru_kernels.cuh.
#pragma once
#include "gru_kernels.cuh"

namespace embclip {

constexpr int kGcUB = 32;        // hidden units per CTA
constexpr int kGcSPT = 3;        // samplers per matvec thread (forward) / granularity of the backward template
constexpr int kGcMaxNS = 12;     // samplers per cluster (max): 4 thread groups x 3
constexpr int kGcThreads = 384;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_v2f32(uint32_t addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void st_cluster_v4f32(uint32_t addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// smem: sWq [H/4][96][4] (W slice, local row = 3 * unit + gate, 4 consecutive k per 16-B item: thread `row` reads one item per
//       k-quad, conflict-free), sH [nsp][H] (hm_t of all units; nsp = samplers of this cluster rounded up to a multiple of 3),
//       sGate [nsp][96] (gate pre-activations), sX [nsp][32] (this CTA's slice of hm_{t+1}, staged for 16-B remote stores)
// The exchange buffer is SINGLE (192 KB of weights leave room for no more at 9+ samplers); the write-after-read hazard is
// covered by a split cluster barrier: arrive after the matvec has read sH, wait before the remote stores, so that phase costs
// nothing; the second (full) barrier publishes hm_{t+1}.
__global__ void __launch_bounds__(kGcThreads, 1)
gru_cluster_forward_kernel(const GruFwdParams p) {
  extern __shared__ float smem_f[];
  const int H = p.H;
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();
  const int grp = blockIdx.x / csize;
  const int u0 = int(crank) * kGcUB;
  const int s0 = grp * p.ns;
  const int ns = min(p.ns, p.N - s0);
  const int G = (p.ns + kGcSPT - 1) / kGcSPT, nsp = G * kGcSPT;
  float* sWq = smem_f;                              // [H/4][96][4]
  float* sH = sWq + 3 * kGcUB * H;                  // [nsp][H]
  float* sGate = sH + nsp * H;                      // [nsp][96]
  float* sX = sGate + nsp * 96;                     // [nsp][32]

  for (int i = tid; i < 3 * kGcUB * (H / 4); i += kGcThreads) {
    const int lr = i / (H / 4), k4 = i - lr * (H / 4);
    const int grow = (lr % 3) * H + u0 + lr / 3;
    *reinterpret_cast<float4*>(sWq + ((size_t)k4 * 96 + lr) * 4) = __ldg(reinterpret_cast<const float4*>(p.w_hh + (size_t)grow * H) + k4);
  }
  // hm_0 for all H units of this cluster's samplers (every CTA reads it from global: no exchange before step 0)
  for (int i = tid; i < nsp * (H / 4); i += kGcThreads) {
    const int s = i / (H / 4), k4 = i - s * (H / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s < ns) {
      const float m = p.masks[s0 + s];
      v = __ldg(reinterpret_cast<const float4*>(p.h0 + (size_t)(s0 + s) * H) + k4);
      v.x *= m; v.y *= m; v.z *= m; v.w *= m;
      if (p.h_init) {
        const float4 hi = __ldg(reinterpret_cast<const float4*>(p.h_init) + k4);
        const float om = 1.f - m;
        v.x += om * hi.x; v.y += om * hi.y; v.z += om * hi.z; v.w += om * hi.w;
      }
    }
    *reinterpret_cast<float4*>(sH + (size_t)s * H + 4 * k4) = v;
  }
  for (int i = tid; i < nsp * kGcUB; i += kGcThreads) sX[i] = 0.f;        // slots of samplers beyond ns stay zero
  __syncthreads();
  cluster_sync_all();                               // every CTA of the cluster is running and initialised before any remote store

  // matvec role: thread (mg = tid / 96, mr = tid % 96) owns gate row mr for samplers 3 mg .. 3 mg + 2
  const int mg = tid / 96, mr = tid - mg * 96;
  const bool mv = mg < G;
  // gate role: thread (es = tid / 32, eu = tid % 32) owns unit u0 + eu of sampler es (one warp per sampler)
  const int es = tid >> 5, eu = lane;
  const int unit = u0 + eu;
  const bool live = es < ns;
  const float bh_r = p.b_hh[unit], bh_z = p.b_hh[H + unit], bh_n = p.b_hh[2 * H + unit];
  const float hinit = p.h_init ? p.h_init[unit] : 0.f;
  const uint32_t sH_u32 = smem_u32(sH);

  for (int t = 0; t < p.T; ++t) {
    const size_t row = (size_t)t * p.N + s0 + es;
    // gi of this thread's (sampler, unit): issued now, consumed after the matvec
    float gi_r = 0.f, gi_z = 0.f, gi_n = 0.f, m_next = 0.f, hm = 0.f;
    if (live) {
      const float* gi = p.gi + row * 3 * H + unit;
      gi_r = __ldg(gi); gi_z = __ldg(gi + H); gi_n = __ldg(gi + 2 * H);
      if (t + 1 < p.T) m_next = p.masks[(size_t)(t + 1) * p.N + s0 + es];
      hm = sH[(size_t)es * H + unit];                // read BEFORE this CTA arrives at the "done reading sH" barrier
    }
    if (mv) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
      const float* hb = sH + (size_t)(kGcSPT * mg) * H;
      const float* wq = sWq + mr * 4;
#pragma unroll 4
      for (int k4 = 0; k4 < H / 4; ++k4) {
        const float4 w = *reinterpret_cast<const float4*>(wq + (size_t)k4 * 384);
        const float4 h0v = *reinterpret_cast<const float4*>(hb + 4 * k4);
        const float4 h1v = *reinterpret_cast<const float4*>(hb + H + 4 * k4);
        const float4 h2v = *reinterpret_cast<const float4*>(hb + 2 * H + 4 * k4);
        a0 = fmaf(w.x, h0v.x, a0); a1 = fmaf(w.x, h1v.x, a1); a2 = fmaf(w.x, h2v.x, a2);
        a0 = fmaf(w.y, h0v.y, a0); a1 = fmaf(w.y, h1v.y, a1); a2 = fmaf(w.y, h2v.y, a2);
        a0 = fmaf(w.z, h0v.z, a0); a1 = fmaf(w.z, h1v.z, a1); a2 = fmaf(w.z, h2v.z, a2);
        a0 = fmaf(w.w, h0v.w, a0); a1 = fmaf(w.w, h1v.w, a1); a2 = fmaf(w.w, h2v.w, a2);
      }
      float* gdst = sGate + (size_t)(kGcSPT * mg) * 96 + mr;
      gdst[0] = a0; gdst[96] = a1; gdst[192] = a2;
    }
    __syncthreads();                                 // gate sums complete; nobody in this CTA reads sH (matvec) any more ...
    if (t + 1 < p.T) cluster_arrive();               // ... which is what the peers wait for before overwriting it
    if (live) {
      const float* gs = sGate + es * 96 + 3 * eu;
      const float gr = gs[0] + bh_r, gz = gs[1] + bh_z, gn = gs[2] + bh_n;
      const float r = sigmoidf_(gi_r + gr);
      const float z = sigmoidf_(gi_z + gz);
      const float n = tanhf(gi_n + r * gn);
      const float h = (1.f - z) * n + z * hm;
      const size_t o = row * H + unit;
      p.out[o] = h;
      if (p.r) { p.r[o] = r; p.z[o] = z; p.n[o] = n; p.hn[o] = gn; }
      sX[es * kGcUB + eu] = m_next * h + (1.f - m_next) * hinit;      // hm_{t+1} of (sampler, my unit)
    }
    if (t + 1 < p.T) {
      __syncthreads();                               // sX complete
      cluster_wait();                                // every CTA of the cluster has finished reading its sH
      // nsp x 8 float4 (samplers x unit quads) to each of the csize CTAs
      const uint32_t per = uint32_t(nsp) * 8u;
      for (uint32_t item = tid; item < csize * per; item += kGcThreads) {
        const uint32_t c = item / per, v = item - c * per, s_ = v >> 3, uq = v & 7u;
        const float4 val = *reinterpret_cast<const float4*>(sX + s_ * kGcUB + 4 * uq);
        const uint32_t a = sH_u32 + uint32_t(int(s_) * H + u0 + 4 * int(uq)) * 4u;
        st_cluster_v4f32(mapa_shared(a, c), val);
      }
      cluster_sync_all();                            // hm_{t+1} is in place everywhere
    }
  }
  cluster_sync_all();                               // no CTA exits while a peer may still store into its shared memory
}

// smem: sW [96][H] (local row = 3 * unit + gate), sR [16 source ranks][NS][32] (partial dhm of MY units from every CTA),
//       sG [NS][96] (dgh of my gate rows).  NS = samplers per cluster rounded up to a multiple of 3 (compile time: the partial
//       sums of a thread live in registers).  Single exchange buffer + split barrier, as in the forward kernel.
template <int NS>
__global__ void __launch_bounds__(kGcThreads, 1)
gru_cluster_backward_kernel(const GruBwdParams p) {
  extern __shared__ float smem_f[];
  const int H = p.H, G3 = 3 * H;
  float* sW = smem_f;                                        // [96][H]
  float* sR = sW + 3 * kGcUB * H;                            // [16][NS][32]
  float* sG = sR + 16 * NS * kGcUB;                          // [NS][96]
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();
  const int grp = blockIdx.x / csize;
  const int u0 = int(crank) * kGcUB;
  const int s0 = grp * p.ns;
  const int ns = min(p.ns, p.N - s0);

  for (int i = tid; i < 3 * kGcUB * (H / 4); i += kGcThreads) {
    const int lr = i / (H / 4), k4 = i - lr * (H / 4);
    const int grow = (lr % 3) * H + u0 + lr / 3;
    *reinterpret_cast<float4*>(sW + lr * H + 4 * k4) = __ldg(reinterpret_cast<const float4*>(p.w_hh + (size_t)grow * H) + k4);
  }
  for (int i = tid; i < NS * 96; i += kGcThreads) sG[i] = 0.f;           // rows of samplers beyond ns stay zero
  __syncthreads();
  cluster_sync_all();

  // gate role: thread (es = tid / 32, eu = tid % 32) owns unit u0 + eu of sampler es (one warp per sampler)
  const int es = tid >> 5, eu = lane;
  const bool live = es < ns;
  const int unit = u0 + eu;
  float carry = (live && p.dhT) ? p.dhT[(size_t)(s0 + es) * H + unit] : 0.f;
  const float hinit = (live && p.h_init) ? p.h_init[unit] : 0.f;
  float dinit = 0.f, amax = 0.f;
  const uint32_t sR_u32 = smem_u32(sR);

  // operands of one BPTT step that do not depend on the carry: fetched one step ahead, under the previous step's matvec
  struct StepIn { float dout, r, z, n, hn, m, hprev; };
  auto fetch = [&](int t) {
    StepIn x = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (live && t >= 0) {
      const size_t row = (size_t)t * p.N + s0 + es;
      const size_t o = row * H + unit;
      x.dout = p.dout[o]; x.r = p.r[o]; x.z = p.z[o]; x.n = p.n[o]; x.hn = p.hn[o];
      x.m = p.masks[row];
      x.hprev = t == 0 ? p.h0[(size_t)(s0 + es) * H + unit] : __ldg(p.out + o - (size_t)p.N * H);
    }
    return x;
  };
  StepIn nxt = fetch(p.T - 1);
  for (int t = p.T - 1; t >= 0; --t) {
    const StepIn in = nxt;
    float direct = 0.f;
    const float m = in.m;
    if (live) {
      const size_t row = (size_t)t * p.N + s0 + es;
      const size_t o = row * H + unit;
      const float dh = in.dout + carry;
      const float r = in.r, z = in.z, n = in.n, hn = in.hn;
      const float hm = in.hprev * m + (1.f - m) * hinit;
      const float dn = dh * (1.f - z);
      const float dz = dh * (hm - n);
      direct = dh * z;
      const float dn_pre = dn * (1.f - n * n);
      const float dz_pre = dz * z * (1.f - z);
      const float dr_pre = dn_pre * hn * r * (1.f - r);
      float* gi = p.dgi + row * G3 + unit;
      float* gh = p.dgh + row * G3 + unit;
      gi[0] = dr_pre; gi[H] = dz_pre; gi[2 * H] = dn_pre;
      gh[0] = dr_pre; gh[H] = dz_pre; gh[2 * H] = dn_pre * r;
      p.hm_h[o] = __float2half_rn(hm);
      amax = fmaxf(amax, fmaxf(fabsf(dr_pre), fmaxf(fabsf(dz_pre), fabsf(dn_pre))));
      float* g = sG + es * 96 + 3 * eu;
      g[0] = dr_pre; g[1] = dz_pre; g[2] = dn_pre * r;
    }
    nxt = fetch(t - 1);
    __syncthreads();
    // partial dhm[s][k] = sum over MY 96 gate rows of dgh[s][row] * W_hh[row][k]; thread owns k = 2 * tid, 2 * tid + 1
    const int k = 2 * tid;
    float a0[NS], a1[NS];
    if (k < H) {
#pragma unroll
      for (int s = 0; s < NS; ++s) { a0[s] = 0.f; a1[s] = 0.f; }
#pragma unroll 2
      for (int lr = 0; lr < 96; lr += 4) {
        float2 w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = *reinterpret_cast<const float2*>(sW + (lr + j) * H + k);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const float4 d = *reinterpret_cast<const float4*>(sG + s * 96 + lr);       // broadcast read, 4 gate rows at once
          a0[s] = fmaf(d.x, w[0].x, a0[s]); a1[s] = fmaf(d.x, w[0].y, a1[s]);
          a0[s] = fmaf(d.y, w[1].x, a0[s]); a1[s] = fmaf(d.y, w[1].y, a1[s]);
          a0[s] = fmaf(d.z, w[2].x, a0[s]); a1[s] = fmaf(d.z, w[2].y, a1[s]);
          a0[s] = fmaf(d.w, w[3].x, a0[s]); a1[s] = fmaf(d.w, w[3].y, a1[s]);
        }
      }
    }
    __syncthreads();                                 // (sG is rewritten next step: ordered by the cluster barrier below as well; this one is
                                                     //  the CTA-scope barrier compute-sanitizer's racecheck understands)
    if (t != p.T - 1) cluster_wait();                // every owner has finished summing the previous step's partials out of sR
    if (k < H) {
      // units k, k + 1 belong to CTA k / 32: its slot [my rank][s][k % 32]
      const uint32_t owner = uint32_t(k) / kGcUB;
      const uint32_t base = sR_u32 + uint32_t((int(crank) * NS) * kGcUB + (k % kGcUB)) * 4u;
      const uint32_t remote = mapa_shared(base, owner);
#pragma unroll
      for (int s = 0; s < NS; ++s) st_cluster_v2f32(remote + uint32_t(s * kGcUB) * 4u, a0[s], a1[s]);
    }
    cluster_sync_all();                              // all partials of this step are in place
    float mine = 0.f;
    if (live) {
      const float* rb = sR + es * kGcUB + eu;
      for (uint32_t c = 0; c < csize; ++c) mine += rb[(size_t)c * NS * kGcUB];
    }
    if (t > 0) cluster_arrive();                     // done reading sR: the peers may overwrite it (they wait before their stores)
    dinit += (direct + mine) * (1.f - m);
    carry = (direct + mine) * m;
  }
  if (live && p.dh_init) atomicAdd(p.dh_init + unit, dinit);
  if (live && p.dh0) p.dh0[(size_t)(s0 + es) * H + unit] = carry;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if (lane == 0) atomicMax(p.amax, __float_as_uint(amax));
  cluster_sync_all();
}

}  // namespace embclip
