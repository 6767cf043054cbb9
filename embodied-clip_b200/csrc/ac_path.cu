// libembclip_b200.so -- actor-critic half of the C ABI (include/embclip_b200.h): the ResnetTensorNavActorCritic
// forward / backward plan of the PPO update, GAE, and the fused clip + Adam step.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "host.h"
#include "ac_kernels.cuh"
#include "gru_kernels.cuh"
#include "gru_cluster.cuh"
#include "wgrad_gemm.cuh"

using namespace embclip;

// =============================================================================================
// wgrad launcher
// =============================================================================================
namespace {

struct WgradOp {
  const void* a; int lda; int M1;          // A [Kdim][lda], logical width M1
  const void* b; int ldb; int N1;          // B [Kdim][ldb], logical width N1
  long long Kdim;
  float* out; long long ldo_m, ldo_n;
  long long n_tile_off = -1;               // -1: BN * ldo_n
  const float* alpha = nullptr;
};

template <int BN>
int launch_wgrad_cfg(const WgradOp& op, cudaStream_t st) {
  using Cfg = WgradCfg<BN>;
  { const int rc_ = ensure_smem((const void*)wgrad_gemm_kernel<BN>, (size_t)(Cfg::kSmemBytes)); if (rc_) return rc_; }
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_map_2d(&tmA, op.a, (int)op.Kdim, op.M1, op.lda, 64, 64))) return rc;
  if ((rc = make_map_2d(&tmB, op.b, (int)op.Kdim, op.N1, op.ldb, Cfg::kBoxN, 64))) return rc;
  WgradParams p;
  memset(&p, 0, sizeof p);
  p.M1 = op.M1; p.N1 = op.N1;
  p.kb_total = int((op.Kdim + 63) / 64);
  p.num_m_tiles = (op.M1 + 127) / 128;
  p.num_n_tiles = (op.N1 + BN - 1) / BN;
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  int splits = num_sms() / tiles;
  if (splits < 1) splits = 1;
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.out = op.out; p.ldo_m = op.ldo_m; p.ldo_n = op.ldo_n;
  p.n_tile_off = op.n_tile_off >= 0 ? op.n_tile_off : (long long)BN * op.ldo_n;
  p.alpha = op.alpha;
  p.vec = (op.ldo_n == 1 && op.ldo_m % 4 == 0 && p.n_tile_off % 4 == 0 && reinterpret_cast<uintptr_t>(op.out) % 16 == 0) ? 1 : 0;
  const int grid = tiles * p.splits;
  wgrad_gemm_kernel<BN><<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(tmA, tmB, p);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int launch_wgrad(const WgradOp& op, cudaStream_t st, int force_bn = 0) {
  if (op.Kdim <= 0 || op.Kdim > 0x7fffffffLL) return fail(EMBCLIP_EINVAL, "wgrad: bad K extent");
  if (op.lda % 8 || op.ldb % 8) return fail(EMBCLIP_EINVAL, "wgrad: row pitches must be multiples of 8 elements");
  if (op.N1 % 32) return fail(EMBCLIP_EINVAL, "wgrad: N extent %d must be a multiple of 32", op.N1);
  int bn = op.N1 % 256 == 0 ? 256 : (op.N1 % 128 == 0 ? 128 : (op.N1 % 64 == 0 ? 64 : 32));
  if (force_bn) bn = force_bn;
  if (op.N1 % bn) return fail(EMBCLIP_EINVAL, "wgrad: N extent %d not a multiple of tile %d", op.N1, bn);
  switch (bn) {
    case 256: return launch_wgrad_cfg<256>(op, st);
    case 128: return launch_wgrad_cfg<128>(op, st);
    case 64: return launch_wgrad_cfg<64>(op, st);
    case 32: return launch_wgrad_cfg<32>(op, st);
  }
  return fail(EMBCLIP_EINVAL, "wgrad: no kernel for tile N %d", bn);
}

int blocks_for(long long total, int threads, int cap_per_sm = 16) {
  long long b = (total + threads - 1) / threads;
  const long long cap = (long long)num_sms() * cap_per_sm;
  if (b > cap) b = cap;
  return b < 1 ? 1 : (int)b;
}

}  // namespace

extern "C" int embclip_wgrad_f16(const void* a, int lda, int M1, const void* b, int ldb, int N1, long long Kdim, float* out,
                                 long long ldo_m, long long ldo_n, const float* alpha, void* stream) {
  EMBCLIP_TRACE();
  if (!a || !b || !out) return fail(EMBCLIP_EINVAL, "wgrad: null pointer");
  WgradOp op{a, lda, M1, b, ldb, N1, Kdim, out, ldo_m, ldo_n, -1, alpha};
  return launch_wgrad(op, (cudaStream_t)stream);
}

// =============================================================================================
// GRU launchers
// =============================================================================================
namespace {

// scratch32 layout (ABI: 256 B): [0, kGruMaxGroups) barrier counters (one per sampler group), [kGruAmaxSlot] max |dgi| bits
constexpr int kGruMaxGroups = 32;
constexpr int kGruAmaxSlot = 32;
constexpr size_t kGruScratchBytes = 256;
struct GruGeom { int groups, ns, grid; };
int gru_geometry(int N, int H, GruGeom* g) {
  if (H % 64 || H <= 0) return fail(EMBCLIP_EINVAL, "gru: hidden size must be a multiple of 64");
  if (N <= 0) return fail(EMBCLIP_EINVAL, "gru: no samplers");
  const int ub = H / kGruUB;
  int max_groups = num_sms() / ub;
  if (max_groups > kGruMaxGroups) max_groups = kGruMaxGroups;      // one barrier counter per group in scratch32
  if (max_groups < 1) return fail(EMBCLIP_EINVAL, "gru: hidden size %d needs more CTAs than the device has SMs", H);
  int groups = (N + kGruNS - 1) / kGruNS;
  if (groups > max_groups) return fail(EMBCLIP_EINVAL, "gru: %d samplers exceed one launch (max %d); split the batch", N, max_groups * kGruNS);
  if (groups < max_groups && N > 1) groups = max_groups < N ? max_groups : N;   // use the idle SMs: smaller sampler groups
  g->groups = groups;
  g->ns = (N + groups - 1) / groups;
  g->groups = (N + g->ns - 1) / g->ns;
  g->grid = ub * g->groups;
  if (g->groups > kGruMaxGroups) return fail(EMBCLIP_EINVAL, "gru: %d sampler groups exceed the %d barrier slots", g->groups, kGruMaxGroups);
  return 0;
}
size_t gru_fwd_smem(int H) { return sizeof(float) * (size_t(3 * kGruUB + kGruNS) * (H + 4) + 8 * kGruNS * 24); }
size_t gru_bwd_smem(int H) { return sizeof(float) * (size_t(3 * H) * kGruUB + size_t(kGruNS) * (3 * H / 2 + 4)); }

// ---- cluster path (gru_cluster.cuh): CS = H / 32 CTAs per cluster, <= 12 samplers per cluster, no cooperative launch
struct GruClusterGeom { int cs, groups, ns, nsp, max_clusters; };
size_t gru_cluster_fwd_smem(int H, int nsp) { return sizeof(float) * (size_t(3 * kGcUB) * H + size_t(nsp) * H + size_t(nsp) * 96 + size_t(nsp) * kGcUB); }
size_t gru_cluster_bwd_smem(int H, int nsp) { return sizeof(float) * (size_t(3 * kGcUB) * H + size_t(16) * nsp * kGcUB + size_t(nsp) * 96); }

template <typename Kernel>
bool gru_cluster_geometry(Kernel kernel, int N, int H, size_t smem_max, GruClusterGeom* g) {
  static const bool off = getenv("EMBCLIP_GRU_NO_CLUSTER") != nullptr;       // A/B switch: the cooperative kernels of gru_kernels.cuh
  if (off || H % 32 || H / 32 > 16 || H / 32 < 1 || smem_max > 232448) return false;
  const int cs = H / 32;
  if (ensure_smem((const void*)kernel, smem_max)) return false;
  if (cs > 8 && cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3(cs); cfg.blockDim = dim3(kGcThreads); cfg.dynamicSmemBytes = smem_max; cfg.attrs = attr; cfg.numAttrs = 1;
  int max_clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&max_clusters, (const void*)kernel, &cfg) != cudaSuccess || max_clusters < 1) {
    cudaGetLastError();
    return false;
  }
  // one wave when it fits (as many clusters as the device keeps resident, samplers spread evenly); otherwise full clusters
  int groups = max_clusters < N ? max_clusters : N;
  int ns = (N + groups - 1) / groups;
  if (ns > kGcMaxNS) { groups = (N + kGcMaxNS - 1) / kGcMaxNS; ns = (N + groups - 1) / groups; }
  g->ns = ns;
  g->groups = (N + ns - 1) / ns;
  g->nsp = (ns + kGcSPT - 1) / kGcSPT * kGcSPT;
  g->cs = cs;
  g->max_clusters = max_clusters;
  return true;
}
template <typename Kernel, typename Params>
int launch_gru_cluster(Kernel kernel, const Params& p, const GruClusterGeom& g, size_t smem, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = g.cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3(g.cs * g.groups); cfg.blockDim = dim3(kGcThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cfg.attrs = attr; cfg.numAttrs = 1;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, p));
  return 0;
}
// the backward kernel keeps a thread's partial sums in registers: samplers per cluster is a template parameter (multiples of 3)
bool gru_cluster_bwd_geometry(int N, int H, GruClusterGeom* g) {
  return gru_cluster_geometry(gru_cluster_backward_kernel<kGcMaxNS>, N, H, gru_cluster_bwd_smem(H, kGcMaxNS), g);
}
int launch_gru_cluster_bwd(const GruBwdParams& p, const GruClusterGeom& g, cudaStream_t st) {
  const size_t smem = gru_cluster_bwd_smem(p.H, g.nsp);
#define EMBCLIP_GCB(NS_)                                                                                            \
  if (g.nsp == NS_) {                                                                                                \
    { const int rc_ = ensure_smem((const void*)gru_cluster_backward_kernel<NS_>, smem); if (rc_) return rc_; }       \
    if (g.cs > 8) CUDA_TRY(cudaFuncSetAttribute((const void*)gru_cluster_backward_kernel<NS_>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)); \
    return launch_gru_cluster(gru_cluster_backward_kernel<NS_>, p, g, smem, st);                                      \
  }
  EMBCLIP_GCB(3) EMBCLIP_GCB(6) EMBCLIP_GCB(9) EMBCLIP_GCB(12)
#undef EMBCLIP_GCB
  return fail(EMBCLIP_EINVAL, "gru: no cluster kernel for %d samplers per cluster", g.nsp);
}

int launch_gru_forward(GruFwdParams p, cudaStream_t st) {
  if (p.H % 64 || p.H <= 0) return fail(EMBCLIP_EINVAL, "gru: hidden size must be a multiple of 64");
  if (p.N <= 0) return fail(EMBCLIP_EINVAL, "gru: no samplers");
  GruClusterGeom cg;
  GruGeom g;
  // a single step (act) has no recurrence to keep on chip: staging 192 KB of weights per CTA for one use costs more than the
  // cooperative kernel's plain pass (measured 3.6 us per rollout step at 60 samplers)
  int coop_groups = num_sms() / (p.H / kGruUB);
  if (coop_groups > kGruMaxGroups) coop_groups = kGruMaxGroups;
  const bool one_step = p.T == 1 && (p.N + kGruNS - 1) / kGruNS <= coop_groups;
  if (!one_step && gru_cluster_geometry(gru_cluster_forward_kernel, p.N, p.H, gru_cluster_fwd_smem(p.H, kGcMaxNS), &cg)) {
    p.groups = cg.groups; p.ns = cg.ns;
    return launch_gru_cluster(gru_cluster_forward_kernel, p, cg, gru_cluster_fwd_smem(p.H, cg.nsp), st);
  }
  int rc;
  if ((rc = gru_geometry(p.N, p.H, &g))) return rc;
  p.groups = g.groups; p.ns = g.ns;
  const size_t smem = gru_fwd_smem(p.H);
  { const int rc_ = ensure_smem((const void*)gru_forward_kernel, (size_t)(smem)); if (rc_) return rc_; }
  CUDA_TRY(cudaMemsetAsync(p.bar, 0, sizeof(unsigned int) * kGruMaxGroups, st));
  void* args[] = {&p};
  CUDA_TRY(cudaLaunchCooperativeKernel((const void*)gru_forward_kernel, dim3(g.grid), dim3(kGruThreads), args, smem, st));
  return 0;
}
int launch_gru_backward(GruBwdParams p, cudaStream_t st) {
  if (p.H % 64 || p.H <= 0) return fail(EMBCLIP_EINVAL, "gru: hidden size must be a multiple of 64");
  if (p.N <= 0) return fail(EMBCLIP_EINVAL, "gru: no samplers");
  GruClusterGeom cg;
  if (gru_cluster_bwd_geometry(p.N, p.H, &cg)) {
    p.groups = cg.groups; p.ns = cg.ns;
    return launch_gru_cluster_bwd(p, cg, st);
  }
  GruGeom g;
  int rc;
  if ((rc = gru_geometry(p.N, p.H, &g))) return rc;
  p.groups = g.groups; p.ns = g.ns;
  const size_t smem = gru_bwd_smem(p.H);
  { const int rc_ = ensure_smem((const void*)gru_backward_kernel, (size_t)(smem)); if (rc_) return rc_; }
  CUDA_TRY(cudaMemsetAsync(p.bar, 0, sizeof(unsigned int) * kGruMaxGroups, st));
  void* args[] = {&p};
  CUDA_TRY(cudaLaunchCooperativeKernel((const void*)gru_backward_kernel, dim3(g.grid), dim3(kGruThreads), args, smem, st));
  return 0;
}

}  // namespace

/* Geometry the GRU launchers would use for (N samplers, hidden H) on the current device: out[0] = CTAs per cluster (0: the
 * cooperative kernels are used instead), out[1] = clusters (sampler groups), out[2] = samplers per cluster, out[3] = clusters
 * the device can keep resident at once (forward kernel), out[4] = the same for the backward kernel. */
extern "C" int embclip_gru_geometry(int N, int H, int* out5) {
  if (!out5 || N <= 0 || H <= 0) return fail(EMBCLIP_EINVAL, "gru_geometry: bad argument");
  GruClusterGeom f, b;
  memset(out5, 0, 5 * sizeof(int));
  if (gru_cluster_geometry(gru_cluster_forward_kernel, N, H, gru_cluster_fwd_smem(H, kGcMaxNS), &f) && gru_cluster_bwd_geometry(N, H, &b)) {
    out5[0] = f.cs; out5[1] = f.groups; out5[2] = f.ns; out5[3] = f.max_clusters; out5[4] = b.max_clusters;
  }
  return 0;
}

extern "C" int embclip_gru_forward(const float* gi, const float* w_hh, const float* b_hh, const float* h0, const float* masks,
                                   const float* h_init, int T, int N, int H, float* out, float* save_r, float* save_z, float* save_n,
                                   float* save_hn, void* scratch32, void* stream) {
  EMBCLIP_TRACE();
  if (!gi || !w_hh || !b_hh || !h0 || !masks || !out || !scratch32) return fail(EMBCLIP_EINVAL, "gru_forward: null pointer");
  if (T <= 0) return fail(EMBCLIP_EINVAL, "gru_forward: T must be positive");
  GruFwdParams p;
  memset(&p, 0, sizeof p);
  p.T = T; p.N = N; p.H = H; p.gi = gi; p.w_hh = w_hh; p.b_hh = b_hh; p.h0 = h0; p.masks = masks; p.out = out;
  p.h_init = h_init;
  p.r = save_r; p.z = save_z; p.n = save_n; p.hn = save_hn;
  p.bar = reinterpret_cast<unsigned int*>(scratch32);
  return launch_gru_forward(p, (cudaStream_t)stream);
}

extern "C" int embclip_gru_backward(const float* w_hh, const float* h0, const float* masks, const float* out, const float* save_r,
                                    const float* save_z, const float* save_n, const float* save_hn, const float* dout,
                                    const float* dh_last, const float* h_init, int T, int N, int H, float* dgi, float* dgh,
                                    void* hm_f16, float* dh0, float* dh_init, void* scratch32, void* stream) {
  EMBCLIP_TRACE();
  if (!w_hh || !h0 || !masks || !out || !save_r || !save_z || !save_n || !save_hn || !dout || !dgi || !dgh || !hm_f16 || !scratch32)
    return fail(EMBCLIP_EINVAL, "gru_backward: null pointer");
  if (T <= 0) return fail(EMBCLIP_EINVAL, "gru_backward: T must be positive");
  GruBwdParams p;
  memset(&p, 0, sizeof p);
  p.T = T; p.N = N; p.H = H; p.w_hh = w_hh; p.h0 = h0; p.masks = masks; p.out = out;
  p.r = save_r; p.z = save_z; p.n = save_n; p.hn = save_hn; p.dout = dout; p.dhT = dh_last;
  p.h_init = h_init; p.dh_init = dh_init;
  if ((h_init == nullptr) != (dh_init == nullptr)) return fail(EMBCLIP_EINVAL, "gru_backward: pass both h_init and dh_init, or neither");
  p.dgi = dgi; p.dgh = dgh; p.hm_h = reinterpret_cast<__half*>(hm_f16); p.dh0 = dh0;
  p.bar = reinterpret_cast<unsigned int*>(scratch32);
  p.amax = p.bar + kGruAmaxSlot;
  CUDA_TRY(cudaMemsetAsync(p.amax, 0, sizeof(unsigned int), (cudaStream_t)stream));
  return launch_gru_backward(p, (cudaStream_t)stream);
}

// =============================================================================================
// Actor-critic plan
// =============================================================================================
namespace {

enum PId { P_EMBED, P_C1W, P_C1B, P_C2W, P_C2B, P_M1W, P_M1B, P_M2W, P_M2B, P_WIH, P_WHH, P_BIH, P_BHH, P_AW, P_AB, P_CW, P_CB, P_HINIT, P_COUNT };

}  // namespace

struct embclip_ac {
  embclip_ac_cfg cfg;
  embclip_param_info params[P_COUNT];
  uint64_t param_floats = 0;
  int num_params = 0;            // P_COUNT - 1 unless cfg.trainable_masked_hidden_state (P_HINIT is the last slot)
  // fp16 weight layouts left in a workspace by the last forward: (params version, params, workspace, T, N).  embclip_ac_act
  // skips re-packing when its key equals this one; every other entry point that touches a workspace overwrites the key.
  uint64_t packed_version = 0;
  const void* packed_params = nullptr;
  const void* packed_ws = nullptr;
  int packed_T = 0, packed_N = 0;
};

static void ac_add_param(embclip_ac* m, int id, const char* name, std::initializer_list<int64_t> shape) {
  embclip_param_info& pi = m->params[id];
  memset(&pi, 0, sizeof pi);
  snprintf(pi.name, sizeof pi.name, "%s", name);
  pi.dtype = EMBCLIP_DTYPE_F32;
  pi.ndim = (int)shape.size();
  uint64_t n = 1;
  int i = 0;
  for (int64_t s : shape) { pi.shape[i++] = s; n *= (uint64_t)s; }
  pi.nbytes = n * 4;
  pi.offset = m->param_floats * 4;
  m->param_floats += (n + 63) & ~uint64_t(63);          // 256-B aligned slots; padding stays zero
}

extern "C" int embclip_ac_create(const embclip_ac_cfg* cfg, embclip_ac_t* out) {
  if (!cfg || !out) return fail(EMBCLIP_EINVAL, "ac_create: null argument");
  const embclip_ac_cfg& c = *cfg;
  if (c.feat_channels % 64 || c.feat_channels <= 0) return fail(EMBCLIP_EINVAL, "ac_create: feat_channels must be a multiple of 64");
  if (c.feat_pixels <= 0 || c.feat_pixels > 256) return fail(EMBCLIP_EINVAL, "ac_create: feat_pixels out of range");
  if (c.compress_hidden % 32 || c.compress_out % 32 || c.goal_dims % 32 || c.combine_hidden % 32 || c.combine_out % 32)
    return fail(EMBCLIP_EINVAL, "ac_create: conv widths must be multiples of 32");
  if (c.compress_out != 32 || c.goal_dims != 32 || c.combine_out != 32)
    return fail(EMBCLIP_EINVAL, "ac_create: compress_out, goal_dims and combine_out must be 32 (tile width of the gradient kernels)");
  if (c.compress_hidden > 128 || c.combine_hidden > 128)
    return fail(EMBCLIP_EINVAL, "ac_create: hidden conv widths above 128 are not built");
  if (c.hidden % 64 || c.hidden <= 0 || (c.hidden / 8) > 148) return fail(EMBCLIP_EINVAL, "ac_create: hidden must be a multiple of 64, at most 1184");
  if (c.num_actions < 1 || c.num_actions > kMaxActions) return fail(EMBCLIP_EINVAL, "ac_create: num_actions must be in [1, %d]", kMaxActions);
  if (c.num_goals < 1) return fail(EMBCLIP_EINVAL, "ac_create: num_goals must be positive");
  embclip_ac* m = new embclip_ac();
  m->cfg = c;
  const int64_t I = (int64_t)c.combine_out * c.feat_pixels, H = c.hidden;
  ac_add_param(m, P_EMBED, "goal_visual_encoder.embed_class.weight", {c.num_goals, c.goal_dims});
  ac_add_param(m, P_C1W, "goal_visual_encoder.resnet_compressor.0.weight", {c.compress_hidden, c.feat_channels, 1, 1});
  ac_add_param(m, P_C1B, "goal_visual_encoder.resnet_compressor.0.bias", {c.compress_hidden});
  ac_add_param(m, P_C2W, "goal_visual_encoder.resnet_compressor.2.weight", {c.compress_out, c.compress_hidden, 1, 1});
  ac_add_param(m, P_C2B, "goal_visual_encoder.resnet_compressor.2.bias", {c.compress_out});
  ac_add_param(m, P_M1W, "goal_visual_encoder.target_obs_combiner.0.weight", {c.combine_hidden, c.compress_out + c.goal_dims, 1, 1});
  ac_add_param(m, P_M1B, "goal_visual_encoder.target_obs_combiner.0.bias", {c.combine_hidden});
  ac_add_param(m, P_M2W, "goal_visual_encoder.target_obs_combiner.2.weight", {c.combine_out, c.combine_hidden, 1, 1});
  ac_add_param(m, P_M2B, "goal_visual_encoder.target_obs_combiner.2.bias", {c.combine_out});
  ac_add_param(m, P_WIH, "state_encoder.rnn.weight_ih_l0", {3 * H, I});
  ac_add_param(m, P_WHH, "state_encoder.rnn.weight_hh_l0", {3 * H, H});
  ac_add_param(m, P_BIH, "state_encoder.rnn.bias_ih_l0", {3 * H});
  ac_add_param(m, P_BHH, "state_encoder.rnn.bias_hh_l0", {3 * H});
  ac_add_param(m, P_AW, "actor.linear.weight", {c.num_actions, H});
  ac_add_param(m, P_AB, "actor.linear.bias", {c.num_actions});
  ac_add_param(m, P_CW, "critic.fc.weight", {1, H});
  ac_add_param(m, P_CB, "critic.fc.bias", {1});
  m->num_params = P_COUNT - 1;
  if (c.trainable_masked_hidden_state) {       // RNNStateEncoder(trainable_masked_hidden_state=True): learned episode-start state
    ac_add_param(m, P_HINIT, "state_encoder.init_hidden_state", {1, 1, H});
    m->num_params = P_COUNT;
  }
  *out = m;
  return 0;
}
extern "C" int embclip_ac_destroy(embclip_ac_t h) { delete h; return 0; }
extern "C" int embclip_ac_num_params(embclip_ac_t h) { return h ? h->num_params : fail(EMBCLIP_EINVAL, "null handle"); }
extern "C" int embclip_ac_param_info(embclip_ac_t h, int index, embclip_param_info* out) {
  if (!h || !out || index < 0 || index >= h->num_params) return fail(EMBCLIP_EINVAL, "ac_param_info: bad argument");
  *out = h->params[index];
  return 0;
}
extern "C" uint64_t embclip_ac_param_floats(embclip_ac_t h) { return h ? h->param_floats : 0; }

namespace {

// workspace carve-up for a [T, N] block
struct AcWs {
  // fp16 activations, rows = F * P
  __half *G, *Y1, *Y2, *Y3, *X;            // X = combiner output = GRU input [F][I]
  __half *dX, *dY3, *dY2, *dG, *dY1;
  __half *dgi_h, *dgh_h, *hm_h;
  // fp16 weights
  __half *W1, *W2, *W3, *W4, *Wih, *WihT, *W2T, *W3T, *W4T;
  // fp32
  float *GI, *Hout, *R, *Z, *Nn, *HN, *dH, *dGI, *dGH, *dlogits, *dvalues;
  float* scale;                            // [2]
  unsigned int* scratch32;                 // [0,32) group barriers, [32] amax (kGruScratchBytes)
  uint64_t total;
};

uint64_t carve(uint64_t& off, uint64_t bytes) {
  const uint64_t o = off;
  off += (bytes + 1023) & ~uint64_t(1023);
  return o;
}

void ac_workspace(const embclip_ac* m, int T, int N, uint8_t* base, AcWs* w) {
  const embclip_ac_cfg& c = m->cfg;
  const uint64_t F = (uint64_t)T * N, M = F * c.feat_pixels, H = c.hidden, I = (uint64_t)c.combine_out * c.feat_pixels;
  uint64_t off = 0;
  auto h16 = [&](uint64_t n) { return reinterpret_cast<__half*>(base + carve(off, n * 2)); };
  auto f32 = [&](uint64_t n) { return reinterpret_cast<float*>(base + carve(off, n * 4)); };
  w->G = h16(M * c.goal_dims); w->Y1 = h16(M * c.compress_hidden); w->Y2 = h16(M * c.compress_out);
  w->Y3 = h16(M * c.combine_hidden); w->X = h16(F * I);
  w->dX = h16(F * I); w->dY3 = h16(M * c.combine_hidden); w->dY2 = h16(M * c.compress_out); w->dG = h16(M * c.goal_dims);
  w->dY1 = h16(M * c.compress_hidden);
  w->dgi_h = h16(F * 3 * H); w->dgh_h = h16(F * 3 * H); w->hm_h = h16(F * H);
  w->W1 = h16((uint64_t)c.compress_hidden * c.feat_channels); w->W2 = h16((uint64_t)c.compress_out * c.compress_hidden);
  w->W3 = h16((uint64_t)c.combine_hidden * (c.compress_out + c.goal_dims)); w->W4 = h16((uint64_t)c.combine_out * c.combine_hidden);
  w->Wih = h16(3 * H * I); w->WihT = h16(3 * H * I);
  w->W2T = h16((uint64_t)c.compress_out * c.compress_hidden); w->W3T = h16((uint64_t)c.combine_hidden * (c.compress_out + c.goal_dims));
  w->W4T = h16((uint64_t)c.combine_out * c.combine_hidden);
  w->GI = f32(F * 3 * H); w->Hout = f32(F * H); w->R = f32(F * H); w->Z = f32(F * H); w->Nn = f32(F * H); w->HN = f32(F * H);
  w->dH = f32(F * H); w->dGI = f32(F * 3 * H); w->dGH = f32(F * 3 * H);
  w->dlogits = f32(F * c.num_actions); w->dvalues = f32(F);
  w->scale = f32(2);
  w->scratch32 = reinterpret_cast<unsigned int*>(base + carve(off, kGruScratchBytes));
  w->total = off;
}

const float* P(const embclip_ac* m, const float* params, int id) { return params + m->params[id].offset / 4; }
float* PG(const embclip_ac* m, float* grads, int id) { return grads + m->params[id].offset / 4; }

int pack_w(const float* src, __half* dst, int R, int C, int mode, int P_, int CG, cudaStream_t st) {
  pack_w_kernel<<<blocks_for((long long)R * C, 256), 256, 0, st>>>(src, dst, R, C, mode, P_, CG);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int check_block(const embclip_ac* m, int T, int N, const void* ws, uint64_t ws_bytes) {
  if (!m) return fail(EMBCLIP_EINVAL, "ac: null handle");
  if (T <= 0 || N <= 0) return fail(EMBCLIP_EINVAL, "ac: empty [T, N] block");
  if ((long long)T * N * m->cfg.feat_pixels > 0x7fffffffLL) return fail(EMBCLIP_EINVAL, "ac: block too large");
  if (!ws || reinterpret_cast<uintptr_t>(ws) % 1024) return fail(EMBCLIP_EINVAL, "ac: workspace must be 1024-B aligned");
  AcWs w;
  ac_workspace(m, T, N, nullptr, &w);
  if (ws_bytes < w.total) return fail(EMBCLIP_ENOSPC, "ac: workspace %llu B < required %llu B", (unsigned long long)ws_bytes, (unsigned long long)w.total);
  return 0;
}

}  // namespace

extern "C" int embclip_ac_num_acts(embclip_ac_t h) { return h ? 25 : fail(EMBCLIP_EINVAL, "null handle"); }
extern "C" int embclip_ac_act_info(embclip_ac_t h, int T, int N, int index, embclip_act_info* out) {
  if (!h || !out || T <= 0 || N <= 0 || index < 0 || index >= 25) return fail(EMBCLIP_EINVAL, "ac_act_info: bad argument");
  const embclip_ac_cfg& c = h->cfg;
  AcWs w;
  ac_workspace(h, T, N, nullptr, &w);
  const int F = T * N, M = F * c.feat_pixels, H = c.hidden, I = c.combine_out * c.feat_pixels;
  struct E { const char* name; const void* p; int dtype, rows, cols; };
  const E tab[25] = {
      {"goal_rows", w.G, 0, M, c.goal_dims}, {"compress0", w.Y1, 0, M, c.compress_hidden}, {"compress2", w.Y2, 0, M, c.compress_out},
      {"combine0", w.Y3, 0, M, c.combine_hidden}, {"x", w.X, 0, F, I}, {"d_x", w.dX, 0, F, I}, {"d_combine0", w.dY3, 0, M, c.combine_hidden},
      {"d_compress2", w.dY2, 0, M, c.compress_out}, {"d_goal_rows", w.dG, 0, M, c.goal_dims}, {"d_compress0", w.dY1, 0, M, c.compress_hidden},
      {"d_gi_f16", w.dgi_h, 0, F, 3 * H}, {"d_gh_f16", w.dgh_h, 0, F, 3 * H}, {"hm_f16", w.hm_h, 0, F, H},
      {"gi", w.GI, 1, F, 3 * H}, {"h", w.Hout, 1, F, H}, {"r", w.R, 1, F, H}, {"z", w.Z, 1, F, H}, {"n", w.Nn, 1, F, H}, {"hn", w.HN, 1, F, H},
      {"d_h", w.dH, 1, F, H}, {"d_gi", w.dGI, 1, F, 3 * H}, {"d_gh", w.dGH, 1, F, 3 * H}, {"d_logits", w.dlogits, 1, F, c.num_actions},
      {"d_values", w.dvalues, 1, F, 1}, {"loss_scale", w.scale, 1, 1, 2}};
  const E& e = tab[index];
  memset(out, 0, sizeof *out);
  snprintf(out->name, sizeof out->name, "%s", e.name);
  out->dtype = e.dtype; out->n = 1; out->h = 1; out->w = e.rows; out->c = e.cols;
  out->offset = (uint64_t)reinterpret_cast<uintptr_t>(e.p);      // carved from a null base: the pointer value IS the offset
  return 0;
}

extern "C" uint64_t embclip_ac_workspace_bytes(embclip_ac_t h, int T, int N) {
  if (!h || T <= 0 || N <= 0) return 0;
  AcWs w;
  ac_workspace(h, T, N, nullptr, &w);
  return w.total;
}

extern "C" int embclip_ac_pack_features(embclip_ac_t h, const float* feats_nchw, long long frames, void* feats_f16, void* stream) {
  EMBCLIP_TRACE();
  if (!h || !feats_nchw || !feats_f16 || frames <= 0) return fail(EMBCLIP_EINVAL, "ac_pack_features: bad argument");
  const int C = h->cfg.feat_channels, Pp = h->cfg.feat_pixels;
  const size_t smem = (size_t)128 * (Pp + 1) * 4;
  if (frames > 65535LL * 1024) return fail(EMBCLIP_EINVAL, "ac_pack_features: too many frames");
  { const int rc_ = ensure_smem((const void*)ac_pack_features_kernel, (size_t)(smem)); if (rc_) return rc_; }
  for (long long f0 = 0; f0 < frames; f0 += 65535) {       // gridDim.y limit
    const int nf = (int)(frames - f0 < 65535 ? frames - f0 : 65535);
    dim3 grid((C + 127) / 128, nf);
    ac_pack_features_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(feats_nchw + (size_t)f0 * C * Pp,
                                                                      reinterpret_cast<__half*>(feats_f16) + (size_t)f0 * Pp * C, C, Pp);
    CUDA_TRY(cudaGetLastError());
  }
  return 0;
}

// params_version != 0: the caller vouches that equal versions mean equal parameter values, so the forward-only fp16 layouts
// already sitting in this workspace (same block shape) are reused instead of re-packed (rollout steps between two updates).
struct ActSample { const float* uniforms; long long* actions; float* log_probs; };
static int ac_forward_impl(embclip_ac_t h, const float* params, uint64_t params_version, const void* feats_f16, const long long* goals,
                           const float* masks, const float* h0, int T, int N, float* logits, float* values, float* h_last,
                           void* workspace, uint64_t workspace_bytes, int save_for_backward, void* stream,
                           const ActSample* sample = nullptr) {
  int rc;
  if ((rc = check_block(h, T, N, workspace, workspace_bytes))) return rc;
  if (!params || !feats_f16 || !goals || !masks || !h0 || !logits || !values) return fail(EMBCLIP_EINVAL, "ac_forward: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const embclip_ac_cfg& c = h->cfg;
  AcWs w;
  ac_workspace(h, T, N, reinterpret_cast<uint8_t*>(workspace), &w);
  const int F = T * N, Pp = c.feat_pixels, M = F * Pp, H = c.hidden, I = c.combine_out * Pp;
  const int KC = c.compress_out + c.goal_dims;

  const bool packed = params_version != 0 && !save_for_backward && h->packed_version == params_version && h->packed_params == params &&
                      h->packed_ws == workspace && h->packed_T == T && h->packed_N == N;
  h->packed_version = save_for_backward ? 0 : params_version;     // (a training forward is followed by a backward that reuses the space)
  h->packed_params = params; h->packed_ws = workspace; h->packed_T = T; h->packed_N = N;
  // fp16 GEMM layouts of the current fp32 master weights
  if (!packed) {
  if ((rc = pack_w(P(h, params, P_C1W), w.W1, c.compress_hidden, c.feat_channels, 0, 0, 0, st))) return rc;
  if ((rc = pack_w(P(h, params, P_C2W), w.W2, c.compress_out, c.compress_hidden, 0, 0, 0, st))) return rc;
  if ((rc = pack_w(P(h, params, P_M1W), w.W3, c.combine_hidden, KC, 0, 0, 0, st))) return rc;
  if ((rc = pack_w(P(h, params, P_M2W), w.W4, c.combine_out, c.combine_hidden, 0, 0, 0, st))) return rc;
  if ((rc = pack_w(P(h, params, P_WIH), w.Wih, 3 * H, I, 2, Pp, c.combine_out, st))) return rc;
  }
  if (save_for_backward) {
    if ((rc = pack_w(P(h, params, P_C2W), w.W2T, c.compress_out, c.compress_hidden, 1, 0, 0, st))) return rc;
    if ((rc = pack_w(P(h, params, P_M1W), w.W3T, c.combine_hidden, KC, 1, 0, 0, st))) return rc;
    if ((rc = pack_w(P(h, params, P_M2W), w.W4T, c.combine_out, c.combine_hidden, 1, 0, 0, st))) return rc;
    if ((rc = pack_w(P(h, params, P_WIH), w.WihT, 3 * H, I, 3, Pp, c.combine_out, st))) return rc;
  }
  ac_goal_rows_kernel<<<blocks_for((long long)M * c.goal_dims, 256), 256, 0, st>>>(P(h, params, P_EMBED), goals, w.G, F, Pp, c.goal_dims, c.num_goals);
  CUDA_TRY(cudaGetLastError());

  auto gemm = [&](const void* a0, int k0, const void* a1, int k1, const void* wgt, const float* bias, void* out, int rows, int nout,
                  int relu, int out_f32) {
    GemmOp g;
    g.a0 = a0; g.n = 1; g.h = 1; g.w = rows; g.c0 = k0; g.lda0 = k0;
    g.a1 = a1; g.c1 = k1;
    g.wgt = wgt; g.ldw = k0 + k1; g.w_rows = nout;
    g.bias = bias; g.out = out; g.cout = nout; g.relu = relu; g.out_f32 = out_f32;
    return launch_gemm(g, st);
  };
  // resnet_compressor: conv1x1 C->128, ReLU, 128->32, ReLU;  target_obs_combiner: [.. | goal] 64->128, ReLU, 128->32
  if ((rc = gemm(feats_f16, c.feat_channels, nullptr, 0, w.W1, P(h, params, P_C1B), w.Y1, M, c.compress_hidden, 1, 0))) return rc;
  if ((rc = gemm(w.Y1, c.compress_hidden, nullptr, 0, w.W2, P(h, params, P_C2B), w.Y2, M, c.compress_out, 1, 0))) return rc;
  if ((rc = gemm(w.Y2, c.compress_out, w.G, c.goal_dims, w.W3, P(h, params, P_M1B), w.Y3, M, c.combine_hidden, 1, 0))) return rc;
  if ((rc = gemm(w.Y3, c.combine_hidden, nullptr, 0, w.W4, P(h, params, P_M2B), w.X, M, c.combine_out, 0, 0))) return rc;
  // GRU input half for all T steps at once
  if ((rc = gemm(w.X, I, nullptr, 0, w.Wih, P(h, params, P_BIH), w.GI, F, 3 * H, 0, 1))) return rc;

  GruFwdParams gp;
  memset(&gp, 0, sizeof gp);
  gp.T = T; gp.N = N; gp.H = H; gp.gi = w.GI; gp.w_hh = P(h, params, P_WHH); gp.b_hh = P(h, params, P_BHH); gp.h0 = h0;
  // one rollout step (act): the step's output IS the new memory -- written in place, no copy launch
  const bool direct_out = T == 1 && !save_for_backward && h_last && h_last != h0;
  gp.masks = masks; gp.out = direct_out ? h_last : w.Hout;
  gp.h_init = c.trainable_masked_hidden_state ? P(h, params, P_HINIT) : nullptr;
  if (save_for_backward) { gp.r = w.R; gp.z = w.Z; gp.n = w.Nn; gp.hn = w.HN; }
  gp.bar = w.scratch32;
  if ((rc = launch_gru_forward(gp, st))) return rc;
  if (h_last && !direct_out)
    CUDA_TRY(cudaMemcpyAsync(h_last, w.Hout + (size_t)(T - 1) * N * H, sizeof(float) * N * H, cudaMemcpyDeviceToDevice, st));

  HeadsParams hp;
  memset(&hp, 0, sizeof hp);
  hp.h = gp.out; hp.w_actor = P(h, params, P_AW); hp.b_actor = P(h, params, P_AB); hp.w_critic = P(h, params, P_CW);
  hp.b_critic = P(h, params, P_CB); hp.logits = logits; hp.values = values; hp.F = F; hp.H = H; hp.A = c.num_actions;
  if (sample) { hp.uniforms = sample->uniforms; hp.sampled = sample->actions; hp.sampled_logp = sample->log_probs; }
  const size_t hsmem = sizeof(float) * (size_t)(c.num_actions + 1) * H;
  { const int rc_ = ensure_smem((const void*)ac_heads_fwd_kernel, (size_t)(hsmem)); if (rc_) return rc_; }
  ac_heads_fwd_kernel<<<blocks_for(F, 8, 4), 256, hsmem, st>>>(hp);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int embclip_ac_forward(embclip_ac_t h, const float* params, const void* feats_f16, const long long* goals,
                                  const float* masks, const float* h0, int T, int N, float* logits, float* values, float* h_last,
                                  void* workspace, uint64_t workspace_bytes, int save_for_backward, void* stream) {
  EMBCLIP_TRACE();
  return ac_forward_impl(h, params, 0, feats_f16, goals, masks, h0, T, N, logits, values, h_last, workspace, workspace_bytes,
                         save_for_backward, stream);
}

extern "C" int embclip_ac_act(embclip_ac_t h, const float* params, uint64_t params_version, const void* feats_f16,
                              const long long* goals, const float* masks, const float* h0, int N, const float* uniforms,
                              long long* actions, float* action_log_probs, float* values, float* h_out, float* logits,
                              void* workspace, uint64_t workspace_bytes, void* stream) {
  EMBCLIP_TRACE();
  if (!uniforms || !actions || !action_log_probs || !h_out || !logits) return fail(EMBCLIP_EINVAL, "ac_act: null pointer");
  // sampling rides in the heads launch
  const ActSample sample{uniforms, actions, action_log_probs};
  return ac_forward_impl(h, params, params_version, feats_f16, goals, masks, h0, 1, N, logits, values, h_out, workspace,
                         workspace_bytes, 0, stream, &sample);
}

extern "C" int embclip_ac_ppo_loss(embclip_ac_t h, const float* params, int T, int N, const long long* actions,
                                   const float* old_action_log_probs, const float* norm_adv, const float* old_values,
                                   const float* returns, float clip_param, float value_loss_coef, float entropy_coef,
                                   float grad_scale, float* logits, float* values, float* loss_sums, void* workspace,
                                   uint64_t workspace_bytes, void* stream) {
  EMBCLIP_TRACE();
  int rc;
  if ((rc = check_block(h, T, N, workspace, workspace_bytes))) return rc;
  if (!params || !actions || !old_action_log_probs || !norm_adv || !old_values || !returns || !logits || !values || !loss_sums)
    return fail(EMBCLIP_EINVAL, "ac_ppo_loss: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const embclip_ac_cfg& c = h->cfg;
  AcWs w;
  ac_workspace(h, T, N, reinterpret_cast<uint8_t*>(workspace), &w);
  HeadsParams hp;
  memset(&hp, 0, sizeof hp);
  hp.h = w.Hout; hp.w_actor = P(h, params, P_AW); hp.b_actor = P(h, params, P_AB); hp.w_critic = P(h, params, P_CW);
  hp.b_critic = P(h, params, P_CB); hp.logits = logits; hp.values = values; hp.F = T * N; hp.H = c.hidden; hp.A = c.num_actions;
  hp.loss = 1; hp.actions = actions; hp.old_logp = old_action_log_probs; hp.adv = norm_adv; hp.old_values = old_values;
  hp.returns = returns; hp.clip = clip_param; hp.vcoef = value_loss_coef; hp.ecoef = entropy_coef; hp.grad_scale = grad_scale;
  hp.dlogits = w.dlogits; hp.dvalues = w.dvalues; hp.loss_out = loss_sums;
  CUDA_TRY(cudaMemsetAsync(loss_sums, 0, sizeof(float) * 3, st));
  const size_t hsmem = sizeof(float) * (size_t)(c.num_actions + 1) * c.hidden;
  ac_heads_fwd_kernel<<<blocks_for(hp.F, 8, 4), 256, hsmem, st>>>(hp);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int embclip_ac_backward(embclip_ac_t h, const float* params, const void* feats_f16, const long long* goals,
                                   const float* masks, const float* h0, int T, int N, const float* dlogits, const float* dvalues,
                                   const float* dh_last, float* grads, void* workspace, uint64_t workspace_bytes, void* stream) {
  EMBCLIP_TRACE();
  int rc;
  if ((rc = check_block(h, T, N, workspace, workspace_bytes))) return rc;
  h->packed_version = 0;                                       // the backward pass reuses workspace regions
  if (!params || !feats_f16 || !goals || !masks || !h0 || !grads) return fail(EMBCLIP_EINVAL, "ac_backward: null pointer");
  if ((dlogits == nullptr) != (dvalues == nullptr)) return fail(EMBCLIP_EINVAL, "ac_backward: pass both dlogits and dvalues, or neither");
  cudaStream_t st = (cudaStream_t)stream;
  const embclip_ac_cfg& c = h->cfg;
  AcWs w;
  ac_workspace(h, T, N, reinterpret_cast<uint8_t*>(workspace), &w);
  const int F = T * N, Pp = c.feat_pixels, M = F * Pp, H = c.hidden, I = c.combine_out * Pp;
  const int KC = c.compress_out + c.goal_dims;
  const float* inv_scale = w.scale + 1;

  // heads
  HeadsBwdParams hb;
  memset(&hb, 0, sizeof hb);
  hb.h = w.Hout; hb.w_actor = P(h, params, P_AW); hb.w_critic = P(h, params, P_CW);
  hb.dlogits = dlogits ? dlogits : w.dlogits; hb.dvalues = dvalues ? dvalues : w.dvalues;
  hb.dh = w.dH; hb.dw_actor = PG(h, grads, P_AW); hb.db_actor = PG(h, grads, P_AB); hb.dw_critic = PG(h, grads, P_CW);
  hb.db_critic = PG(h, grads, P_CB); hb.F = F; hb.H = H; hb.A = c.num_actions;
  ac_heads_bwd_kernel<<<(F + 63) / 64, 256, 0, st>>>(hb);
  CUDA_TRY(cudaGetLastError());

  // BPTT
  GruBwdParams gp;
  memset(&gp, 0, sizeof gp);
  gp.T = T; gp.N = N; gp.H = H; gp.w_hh = P(h, params, P_WHH); gp.h0 = h0; gp.masks = masks; gp.out = w.Hout;
  gp.r = w.R; gp.z = w.Z; gp.n = w.Nn; gp.hn = w.HN; gp.dout = w.dH; gp.dhT = dh_last;
  gp.dgi = w.dGI; gp.dgh = w.dGH; gp.hm_h = w.hm_h; gp.dh0 = nullptr;
  if (c.trainable_masked_hidden_state) { gp.h_init = P(h, params, P_HINIT); gp.dh_init = PG(h, grads, P_HINIT); }
  gp.bar = w.scratch32; gp.amax = w.scratch32 + kGruAmaxSlot;
  CUDA_TRY(cudaMemsetAsync(gp.amax, 0, sizeof(unsigned int), st));
  if ((rc = launch_gru_backward(gp, st))) return rc;

  // loss scale for everything downstream of the fp16 casts; bias gradients from the fp32 values
  compute_scale_kernel<<<1, 1, 0, st>>>(gp.amax, w.scale);
  scale_cast_kernel<<<(F + 63) / 64, 256, 0, st>>>(w.dGI, w.dgi_h, F, 3 * H, w.scale, PG(h, grads, P_BIH));
  scale_cast_kernel<<<(F + 63) / 64, 256, 0, st>>>(w.dGH, w.dgh_h, F, 3 * H, w.scale, PG(h, grads, P_BHH));
  CUDA_TRY(cudaGetLastError());

  auto wgrad = [&](const void* a, int lda, int M1, const void* b, int ldb, int N1, long long K, float* out, long long ldm, long long ldn,
                   long long tile_off) {
    WgradOp op{a, lda, M1, b, ldb, N1, K, out, ldm, ldn, tile_off, inv_scale};
    return launch_wgrad(op, st);
  };
  auto colsum = [&](const __half* x, long long R, int C, float* out) {
    colsum_f16_kernel<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(x, out, R, C, inv_scale);
    return cudaGetLastError() == cudaSuccess ? 0 : fail(EMBCLIP_ECUDA, "colsum launch failed");
  };
  auto dgrad = [&](const void* a0, int k0, const void* wgt, const void* mask, void* out, int rows, int nout) {
    GemmOp g;
    g.a0 = a0; g.n = 1; g.h = 1; g.w = rows; g.c0 = k0; g.lda0 = k0;
    g.wgt = wgt; g.ldw = k0; g.w_rows = nout;
    g.residual = mask; g.res_mode = 1;
    g.out = out; g.cout = nout;
    return launch_gemm(g, st);
  };

  // GRU weights: dW_hh = dgh^T hm,  dW_ih = dgi^T x  (x columns are in (pixel, channel) order: un-permute on store)
  if ((rc = wgrad(w.dgh_h, 3 * H, 3 * H, w.hm_h, H, H, F, PG(h, grads, P_WHH), H, 1, -1))) return rc;
  if ((rc = wgrad(w.dgi_h, 3 * H, 3 * H, w.X, I, I, F, PG(h, grads, P_WIH), I, Pp, 1))) return rc;
  // dx = dgi W_ih  -> gradient of the combiner output, rows = pixels again
  if ((rc = dgrad(w.dgi_h, 3 * H, w.WihT, nullptr, w.dX, F, I))) return rc;
  // target_obs_combiner.2
  if ((rc = wgrad(w.Y3, c.combine_hidden, c.combine_hidden, w.dX, c.combine_out, c.combine_out, M, PG(h, grads, P_M2W), 1, c.combine_hidden, 0))) return rc;
  if ((rc = colsum(w.dX, M, c.combine_out, PG(h, grads, P_M2B)))) return rc;
  if ((rc = dgrad(w.dX, c.combine_out, w.W4T, w.Y3, w.dY3, M, c.combine_hidden))) return rc;
  // target_obs_combiner.0 on [Y2 | G]
  if ((rc = wgrad(w.dY3, c.combine_hidden, c.combine_hidden, w.Y2, c.compress_out, c.compress_out, M, PG(h, grads, P_M1W), KC, 1, -1))) return rc;
  if ((rc = wgrad(w.dY3, c.combine_hidden, c.combine_hidden, w.G, c.goal_dims, c.goal_dims, M, PG(h, grads, P_M1W) + c.compress_out, KC, 1, -1))) return rc;
  if ((rc = colsum(w.dY3, M, c.combine_hidden, PG(h, grads, P_M1B)))) return rc;
  if ((rc = dgrad(w.dY3, c.combine_hidden, w.W3T, w.Y2, w.dY2, M, c.compress_out))) return rc;
  if ((rc = dgrad(w.dY3, c.combine_hidden, w.W3T + (size_t)c.compress_out * c.combine_hidden, nullptr, w.dG, M, c.goal_dims))) return rc;
  ac_goal_grad_kernel<<<blocks_for((long long)F * c.goal_dims, 256), 256, 0, st>>>(w.dG, goals, PG(h, grads, P_EMBED), F, Pp, c.goal_dims,
                                                                                   c.num_goals, inv_scale);
  CUDA_TRY(cudaGetLastError());
  // resnet_compressor.2
  if ((rc = wgrad(w.Y1, c.compress_hidden, c.compress_hidden, w.dY2, c.compress_out, c.compress_out, M, PG(h, grads, P_C2W), 1, c.compress_hidden, 0))) return rc;
  if ((rc = colsum(w.dY2, M, c.compress_out, PG(h, grads, P_C2B)))) return rc;
  if ((rc = dgrad(w.dY2, c.compress_out, w.W2T, w.Y1, w.dY1, M, c.compress_hidden))) return rc;
  // resnet_compressor.0: the one contraction that re-reads the features
  if ((rc = wgrad(w.dY1, c.compress_hidden, c.compress_hidden, feats_f16, c.feat_channels, c.feat_channels, M, PG(h, grads, P_C1W), c.feat_channels, 1, -1))) return rc;
  if ((rc = colsum(w.dY1, M, c.compress_hidden, PG(h, grads, P_C1B)))) return rc;
  return 0;
}

// =============================================================================================
// GAE, advantage normalisation, clip + Adam
// =============================================================================================
extern "C" int embclip_gae(const float* rewards, const float* values, const float* masks, int T, int N, float gamma, float tau,
                           float* returns, float* advantages, float* norm_advantages, float eps, void* stream) {
  EMBCLIP_TRACE();
  if (!rewards || !values || !masks || !returns || !advantages || T <= 0 || N <= 0) return fail(EMBCLIP_EINVAL, "gae: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  gae_kernel<<<(N + 127) / 128, 128, 0, st>>>(rewards, values, masks, returns, advantages, T, N, gamma, tau);
  CUDA_TRY(cudaGetLastError());
  if (norm_advantages) {
    adv_norm_kernel<<<1, 1024, 0, st>>>(returns, values, norm_advantages, T * N, eps);
    CUDA_TRY(cudaGetLastError());
  }
  return 0;
}

extern "C" int embclip_sumsq_f32(const float* x, long long n, float* out, void* stream) {
  EMBCLIP_TRACE();
  if (!x || !out || n < 0) return fail(EMBCLIP_EINVAL, "sumsq: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemsetAsync(out, 0, 2 * sizeof(float), st));               // result + the arrival counter
  if (n) {
    int blocks = blocks_for(n, 256, 4);
    if (blocks > kSumsqMaxBlocks) blocks = kSumsqMaxBlocks;
    sumsq_kernel<<<blocks, 256, 0, st>>>(x, n, out);
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int embclip_adam_clip_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                                      const float* grad_sumsq, float max_grad_norm, float lr, float beta1, float beta2, float eps,
                                      int step, void* stream) {
  EMBCLIP_TRACE();
  if (!params || !grads || !exp_avg || !exp_avg_sq || n <= 0 || step < 1) return fail(EMBCLIP_EINVAL, "adam: bad argument");
  if (max_grad_norm > 0.f && !grad_sumsq) return fail(EMBCLIP_EINVAL, "adam: clipping needs the gradient sum of squares");
  const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  adam_clip_kernel<<<blocks_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, grad_sumsq, max_grad_norm,
                                                                           lr, beta1, beta2, eps, (float)bc1, (float)sqrt(bc2));
  CUDA_TRY(cudaGetLastError());
  return 0;
}
