// libembclip_b200.so -- host side of the C ABI declared in include/embclip_b200.h:
// tensor-map construction, kernel launchers, and the ModifiedResNet (CLIP-RN50) execution plan.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "host.h"
#include "aux_kernels.cuh"
#include "conv_gemm.cuh"
#include "conv3x3_halo.cuh"
#include "gemm2sm.cuh"
#include "bneck_tail.cuh"
#include "tv_kernels.cuh"
#include "bneck_tail_stream.cuh"

using namespace embclip;

// =============================================================================================
// errors
// =============================================================================================
static thread_local std::string g_err;
int embclip::fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
extern "C" const char* embclip_last_error(void) { return g_err.c_str(); }
extern "C" int embclip_abi_version(void) { return 6; }

// =============================================================================================
// TMA descriptors (driver entry point resolved at run time: the library links only against cudart)
// =============================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
static CUtensorMapSwizzle swizzle_for_bytes(int inner_bytes) {
  return inner_bytes >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                            : (inner_bytes >= 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}
// fp16 tensor, dims fastest-first, `pitch[i]` = byte stride of dim i+1.
int embclip::make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* pitch,
                      const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(EMBCLIP_ECUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = pitch[i];
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(int(box[0]) * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(EMBCLIP_ECUDA, "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu %llu box %u %u %u %u", int(r), rank,
                (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
                (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  return 0;
}
// activation / output [n, h, w, c] with pixel pitch `ldc` elements (>= c; lets a GEMM view a K window of a wider row)
static int make_map_nhwc(CUtensorMap* m, const void* base, int n, int h, int w, int c, int ldc, int box_c, int box_w,
                         int box_h, int box_n) {
  const uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)n};
  const uint64_t pitch[3] = {(uint64_t)ldc * 2, (uint64_t)ldc * 2 * w, (uint64_t)ldc * 2 * w * h};
  const uint32_t box[4] = {(uint32_t)box_c, (uint32_t)box_w, (uint32_t)box_h, (uint32_t)box_n};
  return make_map(m, base, 4, dims, pitch, box);
}
// [M = pairs x 2 x W, cols] pixel rows as (channel, dx, dy, window, row pair): a {64, 2, 2, W / 2, 1} box lands in shared memory
// window-major (tile row = 4 window + 2 dy + dx), the order bneck_tail's pooled-output variant wants
static int make_map_windows(CUtensorMap* m, const void* base, long long M, int cols, int ld, int W) {
  const uint64_t row = (uint64_t)ld * 2;
  const uint64_t dims[5] = {(uint64_t)cols, 2, 2, (uint64_t)(W / 2), (uint64_t)(M / (2 * W))};
  const uint64_t pitch[4] = {row, row * W, row * 2, row * 2 * W};
  const uint32_t box[5] = {64, 2, 2, (uint32_t)(W / 2), 1};
  return make_map(m, base, 5, dims, pitch, box);
}
int embclip::make_map_2d(CUtensorMap* m, const void* base, int rows, int cols, int ld, int box_cols, int box_rows) {
  const uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
  const uint64_t pitch[1] = {(uint64_t)ld * 2};
  const uint32_t box[2] = {(uint32_t)box_cols, (uint32_t)box_rows};
  return make_map(m, base, 2, dims, pitch, box);
}

// =============================================================================================
// conv_gemm launcher
// =============================================================================================
// Per-device caches: the dynamic-smem opt-in is a per-device function attribute and the SM count differs by device, so both
// are keyed on the current device ordinal (one process may drive several GPUs: Preprocessor.to(other_device)).
static std::mutex g_dev_mu;
static std::map<std::pair<const void*, int>, size_t> g_smem_attr;
static int g_num_sms[64] = {0};
int embclip::ensure_smem(const void* func, size_t bytes) {
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_dev_mu);
  size_t& have = g_smem_attr[std::make_pair(func, dev)];
  if (bytes > have) {
    if (bytes > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    have = bytes;
  }
  return 0;
}
bool embclip::pdl_enabled() {
  static const bool on = getenv("EMBCLIP_NO_PDL") == nullptr;
  return on;
}
int embclip::num_sms() {
  int dev = 0;
  cudaGetDevice(&dev);
  const int slot = dev >= 0 && dev < 64 ? dev : 0;
  int n = g_num_sms[slot];
  if (!n) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    g_num_sms[slot] = n;
  }
  return n;
}

static void choose_box(int H, int W, int B, int* bw, int* bh, int* bn) {
  if (W * H <= 64) { *bw = W; *bh = H; *bn = 128 / (W * H); if (*bn > B) *bn = B; if (*bn < 1) *bn = 1; return; }
  double best = -1;
  int bbw = 1, bbh = 1;
  for (int w = 1; w <= W && w <= 128; ++w) {
    int h = 128 / w;
    if (h > H) h = H;
    if (h < 1) continue;
    const double tiles = double((W + w - 1) / w) * double((H + h - 1) / h);
    const double eff = double(W) * H / (tiles * 128.0);
    if (eff > best + 1e-9 || (eff > best - 1e-9 && w > bbw)) { best = eff; bbw = w; bbh = h; }
  }
  *bw = bbw; *bh = bbh; *bn = 1;
}

template <int BN, int BK, bool kRes>
static int launch_cfg(const GemmOp& op, cudaStream_t st) {
  using Cfg = ConvGemmCfg<BN, BK, kRes>;
  { const int rc_ = ensure_smem((const void*)conv_gemm_kernel<BN, BK, kRes>, Cfg::kSmemBytes); if (rc_) return rc_; }
  ConvGemmParams p;
  memset(&p, 0, sizeof p);
  CUtensorMap tmA0, tmA1, tmB, tmC, tmR;
  const long long M = (long long)op.n * op.h * op.w;
  const bool conv = op.taps == 9;
  int rc;
  if (conv) {
    choose_box(op.h, op.w, op.n, &p.box_w, &p.box_h, &p.box_n);
    p.tiles_w = (op.w + p.box_w - 1) / p.box_w;
    p.tiles_h = (op.h + p.box_h - 1) / p.box_h;
    p.num_m_blks = p.tiles_w * p.tiles_h * ((op.n + p.box_n - 1) / p.box_n);
    if ((rc = make_map_nhwc(&tmA0, op.a0, op.n, op.h, op.w, op.c0, op.lda0, BK, p.box_w, p.box_h, p.box_n))) return rc;
    if (!op.out_f32 &&
        (rc = make_map_nhwc(&tmC, op.out, op.n, op.h, op.w, op.cout, op.cout, Cfg::kCS, p.box_w, p.box_h, p.box_n)))
      return rc;
  } else {
    if (M > 0x7fffffffLL) return fail(EMBCLIP_EINVAL, "M too large");
    p.box_w = 128; p.box_h = 1; p.box_n = 1;
    p.num_m_blks = int((M + 127) / 128);
    p.tiles_w = p.num_m_blks; p.tiles_h = 1;
    const int acols = op.a_cols ? op.a_cols : op.c0;
    if ((rc = make_map_nhwc(&tmA0, op.a0, 1, 1, (int)M, acols, op.lda0, BK, 128, 1, 1))) return rc;
    if (!op.out_f32 && (rc = make_map_nhwc(&tmC, op.out, 1, 1, (int)M, op.cout, op.cout, Cfg::kCS, 128, 1, 1))) return rc;
  }
  if (op.out_f32) tmC = tmA0;
  if (kRes) {
    if ((rc = make_map_nhwc(&tmR, op.residual, 1, 1, (int)M, op.cout, op.cout, Cfg::kCS, 128, 1, 1))) return rc;
  } else {
    tmR = tmA0;
  }
  if (op.a1) {
    if ((rc = make_map_nhwc(&tmA1, op.a1, 1, 1, (int)M, op.c1, op.c1, BK, 128, 1, 1))) return rc;
  } else {
    tmA1 = tmA0;
  }
  if ((rc = make_map_2d(&tmB, op.wgt, op.w_rows, op.ldw, op.ldw, BK, BN))) return rc;
  p.num_n_blks = op.cout / BN;
  p.taps = op.taps;
  p.kb_per_tap = (op.c0 + BK - 1) / BK;                        // (taps == 1 may end on a zero-filled partial box)
  p.kb_src0 = op.taps * p.kb_per_tap;
  p.kb_total = p.kb_src0 + op.c1 / BK;
  p.a0_box_bytes = uint32_t(p.box_w * p.box_h * p.box_n) * BK * 2;
  p.relu = op.relu;
  p.res_mode = op.res_mode;
  p.out_f32 = op.out_f32;
  p.M = (int)M;
  p.N = op.cout;
  p.bias = op.bias;
  p.out_f32_ptr = reinterpret_cast<float*>(op.out);
  p.res_f32_ptr = op.out_f32 ? op.res_f32 : nullptr;
  p.ldo = op.cout;
  p.grp_n = op.grp_n; p.grp_a_koff = op.grp_a_koff; p.grp_b_koff = op.grp_b_koff; p.grp_b_nmod = op.grp_b_nmod;
  p.reverse = op.reverse;
  const long long tiles = (long long)p.num_m_blks * p.num_n_blks;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  if (grid <= 0) return 0;
  CUDA_TRY(launch_pdl(conv_gemm_kernel<BN, BK, kRes>, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, st, tmA0, tmA1, tmB, tmC, tmR, p));
  return 0;
}

// CTA-pair (cta_group::2) variant for plain 2-D GEMMs: 256 x 256 tiles
template <int BN, bool kRes>
static int launch_gemm2sm(const GemmOp& op, cudaStream_t st) {
  using Cfg = Gemm2Cfg<BN>;
  { const int rc_ = ensure_smem((const void*)gemm2sm_kernel<BN, kRes>, Cfg::kSmemBytes); if (rc_) return rc_; }
  const long long M = (long long)op.n * op.h * op.w;
  if (M > 0x7fffffffLL) return fail(EMBCLIP_EINVAL, "M too large");
  CUtensorMap tmA0, tmA1, tmB, tmC, tmR;
  int rc;
  if ((rc = make_map_nhwc(&tmA0, op.a0, 1, 1, (int)M, op.c0, op.lda0, 64, 128, 1, 1))) return rc;
  if (op.a1) { if ((rc = make_map_nhwc(&tmA1, op.a1, 1, 1, (int)M, op.c1, op.c1, 64, 128, 1, 1))) return rc; }
  else tmA1 = tmA0;
  if (kRes) { if ((rc = make_map_nhwc(&tmR, op.residual, 1, 1, (int)M, op.cout, op.cout, 64, 128, 1, 1))) return rc; }
  else tmR = tmA0;
  if ((rc = make_map_2d(&tmB, op.wgt, op.w_rows, op.ldw, op.ldw, 64, BN / 2))) return rc;
  if (!op.out_f32) { if ((rc = make_map_nhwc(&tmC, op.out, 1, 1, (int)M, op.cout, op.cout, 64, 128, 1, 1))) return rc; }
  else tmC = tmA0;
  Gemm2Params p;
  memset(&p, 0, sizeof p);
  p.num_m_pairs = int((M + 255) / 256);
  p.num_n_blks = op.cout / BN;
  p.kb_src0 = op.c0 / 64;
  p.kb_total = p.kb_src0 + op.c1 / 64;
  p.relu = op.relu; p.out_f32 = op.out_f32; p.M = (int)M; p.N = op.cout;
  p.bias = op.bias;
  p.res_mode = op.res_mode;
  p.out_f32_ptr = reinterpret_cast<float*>(op.out);
  p.res_f32_ptr = op.out_f32 ? op.res_f32 : nullptr;
  p.reverse = op.reverse;
  const long long tiles = (long long)p.num_m_pairs * p.num_n_blks;
  const int max_pairs = num_sms() / 2;
  const int pairs = (int)(tiles < max_pairs ? tiles : max_pairs);
  if (pairs <= 0) return 0;
  CUDA_TRY(launch_pdl(gemm2sm_kernel<BN, kRes>, dim3(2 * pairs), dim3(Cfg::kThreads), Cfg::kSmemBytes, st, tmA0, tmA1, tmB, tmC, tmR, p));
  return 0;
}

static int pick_bn(int cout) {
  if (cout % 128 == 0) return 128;
  if (cout % 64 == 0) return 64;
  if (cout % 32 == 0) return 32;
  return 0;
}

int embclip::launch_gemm(const GemmOp& op, cudaStream_t st, int force_bn) {
  if (op.c0 % 32 || op.c1 % 32 || op.cout % 32) return fail(EMBCLIP_EINVAL, "channels must be multiples of 32 (c0 %d c1 %d cout %d)", op.c0, op.c1, op.cout);
  if (op.taps != 1 && op.taps != 9) return fail(EMBCLIP_EINVAL, "taps must be 1 or 9");
  if (op.taps == 9 && (op.a1 || op.residual || op.out_f32 || op.grp_n)) return fail(EMBCLIP_EINVAL, "3x3 mode supports bias+relu only");
  {
    // large plain GEMMs run on CTA pairs: half the L2 -> smem traffic per FLOP (gemm2sm.cuh)
    static const int use_2sm = getenv("EMBCLIP_2SM") ? atoi(getenv("EMBCLIP_2SM")) : 1;
    static const int min_k = getenv("EMBCLIP_2SM_MINK") ? atoi(getenv("EMBCLIP_2SM_MINK")) : 256;
    const long long M = (long long)op.n * op.h * op.w;
    if (use_2sm && !force_bn && op.taps == 1 && !op.grp_n && op.c0 % 64 == 0 && op.c1 % 64 == 0 && op.cout % 256 == 0 &&
        (op.a_cols == 0 || op.a_cols == op.c0) && op.lda0 == op.c0 && op.c0 + op.c1 >= (op.residual ? 2 * min_k : min_k) &&
        M >= 2048 && !(op.residual && op.out_f32) &&    // (short-K residual GEMMs are epilogue-bound: measured slower on pairs)
        ((M + 255) / 256) * (op.cout / 256) >= num_sms() / 2)                       // enough pair tiles for every SM pair
      return op.residual ? launch_gemm2sm<256, true>(op, st) : launch_gemm2sm<256, false>(op, st);
  }
  int bk = (op.c0 % 64 == 0 && op.c1 % 64 == 0) ? 64 : 32;
  // a long ragged K (the GRU input GEMM: K = 32 x 49 = 1568) still takes 64-wide k-blocks: the last box is half out of bounds in
  // BOTH operands and TMA fills it with zeros.  Half the k-blocks of the BK = 32 path, which is what a few-CTA GEMM's time is
  // made of (0.32 us per k-block per CTA, profiles/r2_launch_floor.txt).
  static const bool ragged64 = getenv("EMBCLIP_NO_RAGGED_K64") == nullptr;
  if (ragged64 && bk == 32 && op.taps == 1 && !op.a1 && !op.grp_n && op.c0 % 64 == 32 && op.c0 >= 256 &&
      (op.a_cols == 0 || op.a_cols == op.c0) && op.ldw == op.c0)
    bk = 64;
  int bn = force_bn ? force_bn : pick_bn(op.cout);
  {
    // few-row GEMMs (a rollout step at 8 samplers: layers 3-4 and the actor-critic's pointwise convs) take 64-wide N tiles: twice
    // the CTAs and more pipeline stages per CTA.  Measured at 8 frames: 0.570 -> 0.566 ms per rollout step; 32-wide tiles and
    // a higher row threshold are slower (0.586 ms; 1.191 -> 1.237 ms at 60 frames with the threshold at 16 K rows)
    static const int small_bn = getenv("EMBCLIP_SMALLM_BN") ? atoi(getenv("EMBCLIP_SMALLM_BN")) : 64;
    static const int small_m = getenv("EMBCLIP_SMALLM_ROWS") ? atoi(getenv("EMBCLIP_SMALLM_ROWS")) : 2048;
    if (small_bn && !force_bn && !op.grp_n && (long long)op.n * op.h * op.w <= small_m && bn > small_bn && op.cout % small_bn == 0) bn = small_bn;
  }
  if (op.grp_n && op.grp_n % bn) bn = op.grp_n % 64 == 0 ? 64 : 32;
  if (op.cout % bn) return fail(EMBCLIP_EINVAL, "cout %d not a multiple of tile N %d", op.cout, bn);
  const bool res = op.residual != nullptr;
  static const int big_bn = getenv("EMBCLIP_BN256") ? atoi(getenv("EMBCLIP_BN256")) : 0;
  if (big_bn && !force_bn && !res && !op.grp_n && op.cout % 256 == 0 && bk == 64) bn = 256;
  if (res && bn > 128) bn = 128;
  // residual GEMMs with short K (conv3 of layers 2-4: K = 128 .. 512) are bound by the L2 -> shared-memory fill (ncu: 9.3 TB/s
  // through the crossbar, tensor pipe 25 %): a 128 x 256 tile moves 20 % fewer operand bytes per FLOP than two 128 x 128 tiles
  static const int res_bn256 = getenv("EMBCLIP_RES_BN256") ? atoi(getenv("EMBCLIP_RES_BN256")) : 1;
  if (res_bn256 && res && !force_bn && !op.grp_n && !op.out_f32 && bk == 64 && op.cout % 256 == 0 && op.taps == 1 &&
      op.c0 + op.c1 >= 256 &&                            // (K = 128, layer 2: HBM-bound, measured 2 % slower on the wide tile)
      (((long long)op.n * op.h * op.w + 127) / 128) * (op.cout / 256) >= num_sms())
    bn = 256;
#define EMBCLIP_CASE(BN_, BK_) \
  if (bn == BN_ && bk == BK_) return res ? launch_cfg<BN_, BK_, true>(op, st) : launch_cfg<BN_, BK_, false>(op, st);
  if (bn == 256 && bk == 64) return res ? launch_cfg<256, 64, true>(op, st) : launch_cfg<256, 64, false>(op, st);
  EMBCLIP_CASE(128, 64)
  EMBCLIP_CASE(64, 64)
  EMBCLIP_CASE(32, 64)
  EMBCLIP_CASE(128, 32)
  EMBCLIP_CASE(64, 32)
  EMBCLIP_CASE(32, 32)
#undef EMBCLIP_CASE
  return fail(EMBCLIP_EINVAL, "no kernel for tile N %d K %d", bn, bk);
}

// =============================================================================================
// conv3x3_halo launcher: geometry (strip / image packing), ring depths, kernel variant
// =============================================================================================
struct C3Geom {
  int R, G, BH, BHo, strips, tiles, Wp;
  uint32_t rows_tma, rows_alloc;
  double eff;
};
static bool c3_geometry(int B, int H, int W, int MS, bool pool, C3Geom* g) {
  const int Wp = W + 1, cap = MS * 128;
  g->Wp = Wp;
  if (Wp > 256) return false;
  if ((H - 1) * Wp + W - 1 < cap) {               // whole images: G per tile, (H+1) rows each
    int G = (cap - ((H - 1) * Wp + W)) / ((H + 1) * Wp) + 1;
    if (G > B) G = B;
    if (H + 1 > 256 || G > 256) return false;
    g->R = H; g->G = G; g->BH = H + 1; g->BHo = H + 1; g->strips = 0;
    g->tiles = (B + G - 1) / G;
    g->rows_tma = uint32_t(G) * (H + 1) * Wp;
    g->eff = double(B) * H * W / (double(g->tiles) * cap);
  } else {                                        // row strips of one image
    int R = (cap - W) / Wp + 1;
    if (pool) R &= ~1;
    if (R < (pool ? 2 : 1)) return false;
    const int strips = (H + R - 1) / R;
    R = (H + strips - 1) / strips;
    if (pool && (R & 1)) ++R;
    if ((R - 1) * Wp + W - 1 >= cap || R + 2 > 256) return false;
    g->R = R; g->G = 1; g->BH = R + 2; g->BHo = R + 2; g->strips = strips;
    g->tiles = B * strips;
    g->rows_tma = uint32_t(R + 2) * Wp;
    g->eff = double(B) * H * W / (double(g->tiles) * cap);
  }
  uint32_t need = uint32_t(cap + 2 * Wp + 2);
  if (need < g->rows_tma) need = g->rows_tma;
  g->rows_alloc = (need + 15u) & ~15u;
  return true;
}

struct Conv3Op {
  const void* in; const void* wgt; const float* bias; void* out;
  int B, H, W, C, N, relu, pool;
  int reverse = 0;
};

template <int BN, int MS, int KC, bool kPool>
static int launch_c3_cfg(const Conv3Op& op, const C3Geom& g, cudaStream_t st) {
  constexpr int SWZ = KC * 2;
  constexpr uint32_t kBBytes = BN * SWZ;
  const uint32_t plane_bytes = (g.rows_alloc * SWZ + 1023u) & ~1023u;
  const uint32_t stage_bytes = kPool ? uint32_t(MS) * 128u * (BN * 2 + 16) : 0u;
  const uint32_t budget = 227u * 1024u - 1024u - kC3BarBytes - stage_bytes;
  const int chunks = op.C / KC;
  static const int env_na = getenv("EMBCLIP_C3_NA") ? atoi(getenv("EMBCLIP_C3_NA")) : 0;
  int n_a = 2;
  if (2u * plane_bytes + 4u * kBBytes > budget) return fail(EMBCLIP_EINVAL, "conv3x3: strip does not fit shared memory (plane %u B)", plane_bytes);
  int n_b = int((budget - 2u * plane_bytes) / kBBytes);
  if (chunks > 1 && n_b >= 6 + int(plane_bytes / kBBytes)) { n_a = 3; n_b -= int((plane_bytes + kBBytes - 1) / kBBytes); }
  if (env_na >= 2 && env_na <= kC3MaxA && uint32_t(env_na) * plane_bytes + 4u * kBBytes <= budget) {
    n_a = env_na;
    n_b = int((budget - uint32_t(n_a) * plane_bytes) / kBBytes);
  }
  if (n_b > 12) n_b = 12;
  // Small convs (stem conv2 / conv3, layer-1 conv2: one chunk, one N block, 18 .. 72 KB of weights): every tile of a CTA reads the
  // SAME nine weight tiles, so they are loaded once and stay; the MMA warp then issues a tile's 9 taps without a barrier round per
  // kernel row (at N = 32 a weight stage feeds only 128 cycles of UMMAs: the rounds, not the traffic, were the cost)
  static const bool resident_ok = getenv("EMBCLIP_C3_NO_RESIDENT") == nullptr;
  const bool resident = resident_ok && chunks == 1 && op.N == BN && uint32_t(n_a) * plane_bytes + 9u * kBBytes <= budget;
  if (resident) n_b = 9;
  const size_t smem = 1024 + size_t(n_a) * plane_bytes + size_t(n_b) * kBBytes + kC3BarBytes + stage_bytes;
  { const int rc_ = ensure_smem((const void*)conv3x3_halo_kernel<BN, MS, KC, kPool>, (size_t)(smem)); if (rc_) return rc_; }
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_map_nhwc(&tmA, op.in, op.B, op.H, op.W, op.C, op.C, KC, g.Wp, g.BH, g.G))) return rc;
  if ((rc = make_map_2d(&tmB, op.wgt, op.N, 9 * op.C, 9 * op.C, KC, BN))) return rc;
  Conv3Params p;
  memset(&p, 0, sizeof p);
  p.num_m_tiles = g.tiles; p.num_n_blks = op.N / BN;
  p.strips_per_image = g.strips; p.R = g.R; p.G = g.G; p.BHo = g.BHo;
  p.H = op.H; p.W = op.W; p.Wp = g.Wp; p.B = op.B; p.C = op.C; p.chunks = chunks;
  p.n_a = n_a; p.n_b = n_b; p.resident = resident ? 1 : 0;
  p.plane_bytes = plane_bytes; p.plane_tx_bytes = g.rows_tma * SWZ;
  p.plane_rows_tma = g.rows_tma; p.plane_rows_alloc = plane_bytes / SWZ;
  p.relu = op.relu; p.N = op.N; p.pool = op.pool; p.reverse = op.reverse;
  p.bias = op.bias; p.out = reinterpret_cast<__half*>(op.out);
  const long long tiles = (long long)p.num_m_tiles * p.num_n_blks;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  if (grid <= 0) return 0;
  CUDA_TRY(launch_pdl(conv3x3_halo_kernel<BN, MS, KC, kPool>, dim3(grid), dim3(kC3Threads), smem, st, tmA, tmB, p));
  return 0;
}

static int launch_conv3x3_halo(const Conv3Op& op, cudaStream_t st) {
  if (op.C % 32 || op.N % 32) return fail(EMBCLIP_EINVAL, "conv3x3: channels must be multiples of 32 (cin %d cout %d)", op.C, op.N);
  if (op.pool && ((op.H | op.W) & 1)) return fail(EMBCLIP_EINVAL, "conv3x3: fused 2x2 pool needs even H, W");
  int kc = op.C % 64 == 0 ? 64 : 32;
  static const int env_bn = getenv("EMBCLIP_C3_BN") ? atoi(getenv("EMBCLIP_C3_BN")) : 0;
  static const int env_ms = getenv("EMBCLIP_C3_MS") ? atoi(getenv("EMBCLIP_C3_MS")) : 0;
  int bn = op.N % 128 == 0 ? 128 : (op.N % 64 == 0 ? 64 : 32);
  if (env_bn && op.N % env_bn == 0 && env_bn >= 128 && op.N % 128 == 0) bn = env_bn;
  int ms = bn == 256 ? 1 : (bn == 128 ? 2 : 4);
  C3Geom g;
  if (env_ms && env_ms * bn * 2 <= 512 && c3_geometry(op.B, op.H, op.W, env_ms, op.pool != 0, &g)) ms = env_ms;   // (a hint: shapes it does not fit keep the default)
  const int ms0 = ms;
  for (;; ms >>= 1) {
    if (ms < 1) {
      // wide rows (e.g. 192-pixel rows with the fused pool need a 2-row strip): retry with 32-channel chunks, whose planes
      // are half the size -- only tile widths that have a 32-chunk instantiation
      if (kc == 64 && bn <= 64) { kc = 32; ms = ms0 << 1; continue; }
      return fail(EMBCLIP_EINVAL, "conv3x3: no strip geometry for H %d W %d", op.H, op.W);
    }
    if (!c3_geometry(op.B, op.H, op.W, ms, op.pool != 0, &g)) continue;
    const uint32_t plane = ((g.rows_alloc * (kc * 2)) + 1023u) & ~1023u;
    const uint32_t stage = op.pool ? uint32_t(ms) * 128u * (bn * 2 + 16) : 0u;
    if (2u * plane + 4u * uint32_t(bn * kc * 2) + stage + 1024u + kC3BarBytes <= 227u * 1024u) break;
  }
  // Under-filled launches (rollout batches: a few dozen frames): halving the sub-tile count doubles the tiles.  Cost model =
  // rounds over the SMs x work per tile; ties keep the larger tile (less weight traffic per MAC).  Large launches (> 2 rounds)
  // keep the default -- there the weight traffic matters more than the last partial round.
  static const bool small_ok = getenv("EMBCLIP_C3_NO_SMALL") == nullptr;
  while (small_ok && !env_ms && ms > 1) {
    const long long nb = op.N / bn, sms = num_sms();
    const long long tiles = (long long)g.tiles * nb;
    if (tiles > 2 * sms) break;
    C3Geom g2;
    if (!c3_geometry(op.B, op.H, op.W, ms / 2, op.pool != 0, &g2)) break;
    const uint32_t plane2 = ((g2.rows_alloc * (kc * 2)) + 1023u) & ~1023u;
    const uint32_t stage2 = op.pool ? uint32_t(ms / 2) * 128u * (bn * 2 + 16) : 0u;
    if (2u * plane2 + 4u * uint32_t(bn * kc * 2) + stage2 + 1024u + kC3BarBytes > 227u * 1024u) break;
    const long long rounds = (tiles + sms - 1) / sms, rounds2 = ((long long)g2.tiles * nb + sms - 1) / sms;
    if (rounds2 * (ms / 2) >= rounds * ms) break;
    ms /= 2;
    g = g2;
  }
#define EMBCLIP_C3(BN_, MS_, KC_) \
  if (bn == BN_ && ms == MS_ && kc == KC_) \
    return op.pool ? launch_c3_cfg<BN_, MS_, KC_, true>(op, g, st) : launch_c3_cfg<BN_, MS_, KC_, false>(op, g, st);
  EMBCLIP_C3(32, 4, 32) EMBCLIP_C3(32, 2, 32) EMBCLIP_C3(32, 1, 32)
  EMBCLIP_C3(32, 4, 64) EMBCLIP_C3(32, 2, 64) EMBCLIP_C3(32, 1, 64)
  EMBCLIP_C3(64, 4, 32) EMBCLIP_C3(64, 2, 32) EMBCLIP_C3(64, 1, 32)
  EMBCLIP_C3(64, 4, 64) EMBCLIP_C3(64, 2, 64) EMBCLIP_C3(64, 1, 64)
  EMBCLIP_C3(128, 2, 64) EMBCLIP_C3(128, 1, 64)
  EMBCLIP_C3(256, 1, 64)
#undef EMBCLIP_C3
  return fail(EMBCLIP_EINVAL, "conv3x3: no kernel for tile N %d x %d sub-tiles, chunk %d", bn, ms, kc);
}

// =============================================================================================
// bneck_tail launcher: conv3 (+ K-concat downsample | + identity residual) + ReLU, then the next block's conv1
// =============================================================================================
struct TailOp {
  const void* a0 = nullptr;        // y2 [M, 64]
  const void* a1 = nullptr;        // block input [M, 64] for the K-concatenated downsample conv (or null)
  const void* w3 = nullptr;        // [256, 64 (+64)]
  const float* b3 = nullptr;
  const void* residual = nullptr;  // identity [M, 256] (or null)
  void* out = nullptr;             // x' [M, 256]
  const void* w1 = nullptr;        // next conv1 [N1, 256]
  const float* b1 = nullptr;
  void* y1 = nullptr;              // [M, N1]
  long long M = 0;
  int n1 = 0;
  int reverse = 0;
  void* pool_out = nullptr;        // [M / 4, 256]: the 2x2-pooled x' INSTEAD of x' (image width pool_w2, M = images x H x W)
  int pool_mode = 0;               // 1 average, 2 top-left pixel
  int pool_w2 = 0;                 // image width W (even, 2 W <= 128; H even)
};
template <int K3C, int N1, bool kRes, bool kPool = false>
static int launch_tail_cfg(const TailOp& op, cudaStream_t st) {
  using Cfg = TailCfg<K3C, N1>;
  { const int rc_ = ensure_smem((const void*)bneck_tail_kernel<K3C, N1, kRes, kPool>, Cfg::kSmemBytes); if (rc_) return rc_; }
  const int M = (int)op.M;
  const int rows = kPool ? 2 * op.pool_w2 : 128;              // tile = two image rows when the epilogue pools
  CUtensorMap tmA0, tmA1, tmW3, tmW1, tmR, tmC;
  int rc;
  if (kPool) { if ((rc = make_map_windows(&tmA0, op.a0, M, 64, 64, op.pool_w2))) return rc; }
  else if ((rc = make_map_2d(&tmA0, op.a0, M, 64, 64, 64, rows))) return rc;
  if (K3C == 2) { if ((rc = make_map_2d(&tmA1, op.a1, M, 64, 64, 64, rows))) return rc; }
  else tmA1 = tmA0;
  if ((rc = make_map_2d(&tmW3, op.w3, 256, 64 * K3C, 64 * K3C, 64, 256))) return rc;
  if ((rc = make_map_2d(&tmW1, op.w1, N1, 256, 256, 64, N1))) return rc;
  if (kRes && kPool) { if ((rc = make_map_windows(&tmR, op.residual, M, 256, 256, op.pool_w2))) return rc; }
  else if (kRes) { if ((rc = make_map_2d(&tmR, op.residual, M, 256, 256, 64, rows))) return rc; }
  if (!kPool) { if ((rc = make_map_2d(&tmC, op.out, M, 256, 256, 64, 128))) return rc; }
  else tmC = tmR;
  if (!kRes) tmR = tmC;
  TailParams p;
  memset(&p, 0, sizeof p);
  p.num_tiles = (M + rows - 1) / rows;
  p.M = M;
  p.reverse = op.reverse;
  p.tile_rows = rows;
  p.pool_w = op.pool_w2 / 2; p.pool_mode = op.pool_mode; p.pool_out = reinterpret_cast<__half*>(op.pool_out);
  p.bias3 = op.b3; p.bias1 = op.b1;
  p.y1 = reinterpret_cast<__half*>(op.y1);
  const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  if (grid <= 0) return 0;
  CUDA_TRY(launch_pdl(bneck_tail_kernel<K3C, N1, kRes, kPool>, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, st, tmA0, tmA1, tmW3, tmW1, tmR, tmC, p));
  return 0;
}
static int launch_bneck_tail(const TailOp& op, cudaStream_t st) {
  if (op.M <= 0) return 0;
  if (op.M > 0x7fffffffLL) return fail(EMBCLIP_EINVAL, "bneck_tail: M too large");
  if (!op.a0 || !op.w3 || !op.b3 || (!op.out && !op.pool_out) || !op.w1 || !op.b1 || !op.y1) return fail(EMBCLIP_EINVAL, "bneck_tail: null argument");
  if (op.pool_out) {
    const int W = op.pool_w2;
    if (!op.residual || op.n1 != 128) return fail(EMBCLIP_EINVAL, "bneck_tail: the pooled-output variant is built for the identity residual and a 128-wide next conv1");
    if (W <= 0 || W % 2 || 2 * W > 128 || (2 * W) % 8 || op.M % (2 * W)) return fail(EMBCLIP_EINVAL, "bneck_tail: pooled output needs an even image width <= 64 and whole row pairs (W %d, M %lld)", W, op.M);
    if (op.pool_mode != 1 && op.pool_mode != 2) return fail(EMBCLIP_EINVAL, "bneck_tail: pool mode %d", op.pool_mode);
    return launch_tail_cfg<1, 128, true, true>(op, st);
  }
  if ((op.a1 != nullptr) == (op.residual != nullptr)) return fail(EMBCLIP_EINVAL, "bneck_tail: exactly one of downsample source / identity residual");
  if (op.n1 != 64 && op.n1 != 128) return fail(EMBCLIP_EINVAL, "bneck_tail: next conv1 width must be 64 or 128 (got %d)", op.n1);
  if (op.a1 && op.n1 != 64) return fail(EMBCLIP_EINVAL, "bneck_tail: the K-concatenated variant is built for a 64-wide next conv1 only");
  if (op.a1) return launch_tail_cfg<2, 64, false>(op, st);
  return op.n1 == 64 ? launch_tail_cfg<1, 64, true>(op, st) : launch_tail_cfg<1, 128, true>(op, st);
}

// bneck_tail_stream: identity-residual conv3 [M, K3] -> [M, N3] + the next conv1 [M, N3] -> [M, N1], weights streamed per quarter
struct TailStreamOp {
  const void* a = nullptr; const void* w3 = nullptr; const float* b3 = nullptr; const void* residual = nullptr; void* out = nullptr;
  const void* w1 = nullptr; const float* b1 = nullptr; void* y1 = nullptr;
  long long M = 0;
  int k3 = 0, n3 = 0, n1 = 0, reverse = 0;
};
template <int K3C, int N1>
static int launch_tail_stream_cfg(const TailStreamOp& op, cudaStream_t st) {
  using Cfg = TailStreamCfg<K3C, N1>;
  { const int rc_ = ensure_smem((const void*)bneck_tail_stream_kernel<K3C, N1>, Cfg::kSmemBytes); if (rc_) return rc_; }
  const int M = (int)op.M;
  CUtensorMap tmA, tmW3, tmW1, tmR, tmC;
  int rc;
  if ((rc = make_map_2d(&tmA, op.a, M, op.k3, op.k3, 64, 128))) return rc;
  if ((rc = make_map_2d(&tmW3, op.w3, op.n3, op.k3, op.k3, 64, 64))) return rc;
  if ((rc = make_map_2d(&tmW1, op.w1, N1, op.n3, op.n3, 64, N1))) return rc;
  if ((rc = make_map_2d(&tmR, op.residual, M, op.n3, op.n3, 64, 128))) return rc;
  if ((rc = make_map_2d(&tmC, op.out, M, op.n3, op.n3, 64, 128))) return rc;
  TailStreamParams p;
  memset(&p, 0, sizeof p);
  p.num_tiles = (M + 127) / 128;
  p.M = M; p.nq = op.n3 / 64; p.reverse = op.reverse;
  p.bias3 = op.b3; p.bias1 = op.b1;
  p.y1 = reinterpret_cast<__half*>(op.y1);
  const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  if (grid <= 0) return 0;
  CUDA_TRY(launch_pdl(bneck_tail_stream_kernel<K3C, N1>, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, st, tmA, tmW3, tmW1, tmR, tmC, p));
  return 0;
}
static bool tail_stream_supported(int k3, int n3, int n1) { return k3 == 128 && n1 == 128 && n3 % 64 == 0 && n3 >= 256 && n3 <= 1024; }
static int launch_bneck_tail_stream(const TailStreamOp& op, cudaStream_t st) {
  if (op.M <= 0) return 0;
  if (op.M > 0x7fffffffLL) return fail(EMBCLIP_EINVAL, "bneck_tail_stream: M too large");
  if (!op.a || !op.w3 || !op.b3 || !op.residual || !op.out || !op.w1 || !op.b1 || !op.y1) return fail(EMBCLIP_EINVAL, "bneck_tail_stream: null argument");
  if (!tail_stream_supported(op.k3, op.n3, op.n1))
    return fail(EMBCLIP_EINVAL, "bneck_tail_stream: built for K3 = 128, N1 = 128, N3 a multiple of 64 in [256, 1024] (got %d, %d, %d)", op.k3, op.n1, op.n3);
  return launch_tail_stream_cfg<2, 128>(op, st);
}

// =============================================================================================
// primitive-op entry points
// =============================================================================================
extern "C" int embclip_bneck_tail_stream_f16(const void* y2, const void* w3, const float* b3, const void* residual, void* out, const void* w1,
                                             const float* b1, void* y1, int64_t M, int K3, int N3, int n1, void* stream) {
  EMBCLIP_TRACE();
  TailStreamOp op;
  op.a = y2; op.w3 = w3; op.b3 = b3; op.residual = residual; op.out = out; op.w1 = w1; op.b1 = b1; op.y1 = y1;
  op.M = M; op.k3 = K3; op.n3 = N3; op.n1 = n1;
  return launch_bneck_tail_stream(op, (cudaStream_t)stream);
}

extern "C" int embclip_gemm_f16(const void* a0, const void* a1, const void* w, const float* bias, const void* residual,
                                void* out, int M, int N, int K0, int K1, int relu, int out_f32, void* stream) {
  EMBCLIP_TRACE();
  if (!a0 || !w || !out || M <= 0) return fail(EMBCLIP_EINVAL, "gemm: null pointer or empty M");
  GemmOp op;
  op.a0 = a0; op.n = 1; op.h = 1; op.w = M; op.c0 = K0; op.lda0 = K0;
  op.a1 = K1 ? a1 : nullptr; op.c1 = K1;
  op.wgt = w; op.ldw = K0 + K1; op.w_rows = N;
  op.bias = bias; op.residual = residual; op.out = out; op.cout = N; op.relu = relu; op.out_f32 = out_f32;
  return launch_gemm(op, (cudaStream_t)stream);
}

extern "C" int embclip_gemm_grouped_f16(const void* a, int lda, const void* w, int ldw, int w_rows, const float* bias,
                                        void* out, int M, int N, int K, int grp_n, int grp_a_koff, int grp_b_koff,
                                        int grp_b_nmod, int relu, int out_f32, void* stream) {
  if (!a || !w || !out || M <= 0) return fail(EMBCLIP_EINVAL, "gemm_grouped: null pointer or empty M");
  GemmOp op;
  op.a0 = a; op.n = 1; op.h = 1; op.w = M; op.c0 = K; op.lda0 = lda; op.a_cols = lda;
  op.wgt = w; op.ldw = ldw; op.w_rows = w_rows;
  op.bias = bias; op.out = out; op.cout = N; op.relu = relu; op.out_f32 = out_f32;
  op.grp_n = grp_n; op.grp_a_koff = grp_a_koff; op.grp_b_koff = grp_b_koff; op.grp_b_nmod = grp_b_nmod;
  return launch_gemm(op, (cudaStream_t)stream);
}

extern "C" int embclip_conv3x3_f16(const void* in, const void* w, const float* bias, void* out, int B, int H, int W,
                                   int Cin, int Cout, int relu, int pool, void* stream) {
  EMBCLIP_TRACE();
  if (!in || !w || !out || B <= 0) return fail(EMBCLIP_EINVAL, "conv3x3: null pointer or empty batch");
  static const bool legacy = getenv("EMBCLIP_CONV3_LEGACY") != nullptr;   // 9-box-loads implicit GEMM (first version), for A/B timing
  if (legacy && !pool) {
    GemmOp op;
    op.a0 = in; op.n = B; op.h = H; op.w = W; op.c0 = Cin; op.lda0 = Cin; op.taps = 9;
    op.wgt = w; op.ldw = 9 * Cin; op.w_rows = Cout;
    op.bias = bias; op.out = out; op.cout = Cout; op.relu = relu;
    return launch_gemm(op, (cudaStream_t)stream);
  }
  Conv3Op op{in, w, bias, out, B, H, W, Cin, Cout, relu, pool};
  return launch_conv3x3_halo(op, (cudaStream_t)stream);
}

extern "C" int embclip_bneck_tail_f16(const void* y2, const void* x0, const void* w3, const float* b3, const void* residual, void* out,
                                      const void* w1, const float* b1, void* y1, int64_t M, int n1, void* stream) {
  EMBCLIP_TRACE();
  TailOp op;
  op.a0 = y2; op.a1 = x0; op.w3 = w3; op.b3 = b3; op.residual = residual; op.out = out;
  op.w1 = w1; op.b1 = b1; op.y1 = y1; op.M = M; op.n1 = n1;
  return launch_bneck_tail(op, (cudaStream_t)stream);
}

extern "C" int embclip_bneck_tail_pool_f16(const void* y2, const void* w3, const float* b3, const void* residual, void* pool_out, int pool_mode,
                                           int width, const void* w1, const float* b1, void* y1, int64_t M, int n1, void* stream) {
  EMBCLIP_TRACE();
  if (!pool_out) return fail(EMBCLIP_EINVAL, "bneck_tail_pool: null pooled output");
  TailOp op;
  op.a0 = y2; op.w3 = w3; op.b3 = b3; op.residual = residual; op.pool_out = pool_out; op.pool_mode = pool_mode; op.pool_w2 = width;
  op.w1 = w1; op.b1 = b1; op.y1 = y1; op.M = M; op.n1 = n1;
  return launch_bneck_tail(op, (cudaStream_t)stream);
}

// mode 1: nn.AvgPool2d(2); 2: x[:, ::2, ::2] (input of a stride-2 1x1 conv); 3: nn.MaxPool2d(3, 2, 1)
static int launch_avgpool2(const void* in, void* out, int B, int H, int W, int C, cudaStream_t st, int mode = 1) {
  if (H % 2 || W % 2 || C % 8) return fail(EMBCLIP_EINVAL, "pool: H, W must be even and C a multiple of 8");
  const long long total = (long long)B * (H / 2) * (W / 2) * (C / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks <= 0) return 0;
  auto k = mode == 3 ? maxpool3x3s2_kernel : (mode == 2 ? subsample2_kernel : avgpool2_kernel);
  CUDA_TRY(launch_pdl(k, dim3((int)blocks), dim3(256), 0, st, reinterpret_cast<const __half*>(in), reinterpret_cast<__half*>(out), B, H, W, C));
  return 0;
}
static int launch_im2col7(const void* x, int x_u8, const float* norm6, void* y, int B, int R, cudaStream_t st) {
  if (R % 2) return fail(EMBCLIP_EINVAL, "stem: resolution must be even");
  StemNorm nm;
  for (int c = 0; c < 3; ++c) { nm.scale[c] = norm6 ? norm6[c] : 1.f; nm.offset[c] = norm6 ? norm6[3 + c] : 0.f; }
  const long long tiles = ((long long)B * (R / 2) * (R / 2) + 31) / 32;
  long long blocks = (long long)num_sms() * 8;
  if (blocks > tiles) blocks = tiles;
  if (blocks <= 0) return 0;
  if (x_u8)
    CUDA_TRY(launch_pdl(im2col7x7s2_kernel<uint8_t>, dim3((unsigned)blocks), dim3(256), 0, st, reinterpret_cast<const uint8_t*>(x), reinterpret_cast<__half*>(y), B, R, nm));
  else
    CUDA_TRY(launch_pdl(im2col7x7s2_kernel<float>, dim3((unsigned)blocks), dim3(256), 0, st, reinterpret_cast<const float*>(x), reinterpret_cast<__half*>(y), B, R, nm));
  return 0;
}
extern "C" int embclip_pool2_f16(const void* in, void* out, int B, int H, int W, int C, int mode, void* stream) {
  if (!in || !out) return fail(EMBCLIP_EINVAL, "pool2: null pointer");
  if (mode < 1 || mode > 3) return fail(EMBCLIP_EINVAL, "pool2: mode must be 1 (avg 2x2), 2 (subsample ::2) or 3 (max 3x3 stride 2 pad 1)");
  return launch_avgpool2(in, out, B, H, W, C, (cudaStream_t)stream, mode);
}
extern "C" int embclip_avgpool2_f16(const void* in, void* out, int B, int H, int W, int C, void* stream) {
  if (!in || !out) return fail(EMBCLIP_EINVAL, "avgpool2: null pointer");
  return launch_avgpool2(in, out, B, H, W, C, (cudaStream_t)stream);
}

// tensor-core version (hi/lo-split im2col rows); wtc = fp16 [32][128]
template <typename TIn, int COUT, bool kFast = false>
static int launch_stem_rows(const void* x, const void* wtc, const float* b, void* y, int B, int R, const StemNorm& nm, cudaStream_t st) {
  const int Ro = R / 2;
  int segs = (Ro + 127) / 128;
  while (Ro % segs) ++segs;                                  // equal segments of <= 128 output pixels
  const size_t smem = 1024 + 32768 + 2 * COUT * 128 + size_t(kStemRowsStages) * stem_rows_stage_bytes<TIn>(R) + 64;
  if (smem > 227u * 1024u) return fail(EMBCLIP_EINVAL, "stem: resolution %d does not fit the row ring", R);
  { const int rc_ = ensure_smem((const void*)stem_conv1_rows_kernel<TIn, COUT, kFast>, (size_t)(smem)); if (rc_) return rc_; }
  const long long tiles = (long long)B * Ro * segs;
  if (tiles >= (1ll << 31)) return fail(EMBCLIP_EINVAL, "stem: batch %d too large for one launch", B);
  int per_sm = int((227u * 1024u) / smem);
  if (per_sm > 4) per_sm = 4;
  long long g = (long long)num_sms() * per_sm;
  if (g > tiles) g = tiles;
  if (g <= 0) return 0;
  CUDA_TRY(launch_pdl(stem_conv1_rows_kernel<TIn, COUT, kFast>, dim3((unsigned)g), dim3(128), smem, st, reinterpret_cast<const TIn*>(x),
                      reinterpret_cast<const __half*>(wtc), b, reinterpret_cast<__half*>(y), B, R, make_fastdiv((uint32_t)segs),
                      make_fastdiv((uint32_t)Ro), nm));
  return 0;
}

static int launch_stem_conv1_tc(const void* x, int x_u8, const float* norm6, const void* wtc, const float* b, void* y, int B, int R, int Cout, cudaStream_t st) {
  if (R % 2) return fail(EMBCLIP_EINVAL, "stem: resolution must be even");
  StemNorm nm;
  for (int c = 0; c < 3; ++c) { nm.scale[c] = norm6 ? norm6[c] : 1.f; nm.offset[c] = norm6 ? norm6[3 + c] : 0.f; }
  // row-tiled variant (bulk-copied input rows, see aux_kernels.cuh) when rows are 16-B granular
  static const bool gather_only = getenv("EMBCLIP_STEM_GATHER") != nullptr;
  const size_t esz = x_u8 ? 1 : 4;
  const bool rows_ok = (size_t(R) * 3 * esz) % 16 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0;
  // raw uint8 frames: the half2 split (no conversion-unit instructions) needs 1024 + round(255 mean) exact in fp16
  static const bool no_fast = getenv("EMBCLIP_STEM_NO_FAST_U8") != nullptr;
  bool fast = x_u8 && !no_fast;
  for (int c = 0; c < 3 && fast; ++c) {
    const float m = -nm.offset[c] / nm.scale[c];
    fast = std::isfinite(m) && std::fabs(m) <= 1000.f && std::fabs(nm.scale[c]) >= 1e-3f && std::fabs(nm.scale[c]) <= 16.f;
  }
  if (Cout == 64) {
    if (!rows_ok) return fail(EMBCLIP_EINVAL, "stem: the 64-channel stem needs 16-B aligned frame rows (resolution %d)", R);
    if (fast) return launch_stem_rows<uint8_t, 64, true>(x, wtc, b, y, B, R, nm, st);
    return x_u8 ? launch_stem_rows<uint8_t, 64>(x, wtc, b, y, B, R, nm, st) : launch_stem_rows<float, 64>(x, wtc, b, y, B, R, nm, st);
  }
  if (Cout != 32) return fail(EMBCLIP_EINVAL, "stem conv1: Cout %d not built (32 or 64)", Cout);
  if (!gather_only && rows_ok && fast) return launch_stem_rows<uint8_t, 32, true>(x, wtc, b, y, B, R, nm, st);
  if (!gather_only && rows_ok)
    return x_u8 ? launch_stem_rows<uint8_t, 32>(x, wtc, b, y, B, R, nm, st) : launch_stem_rows<float, 32>(x, wtc, b, y, B, R, nm, st);
  { const int rc_ = ensure_smem((const void*)stem_conv1_tc_kernel<float>, (size_t)(kStemTcSmem)); if (rc_) return rc_; }
  { const int rc_ = ensure_smem((const void*)stem_conv1_tc_kernel<uint8_t>, (size_t)(kStemTcSmem)); if (rc_) return rc_; }
  const long long tiles = ((long long)B * (R / 2) * (R / 2) + 127) / 128;
  long long grid = (long long)num_sms() * 4;
  if (grid > tiles) grid = tiles;
  if (grid <= 0) return 0;
  if (x_u8)
    CUDA_TRY(launch_pdl(stem_conv1_tc_kernel<uint8_t>, dim3((unsigned)grid), dim3(128), (size_t)kStemTcSmem, st, reinterpret_cast<const uint8_t*>(x),
                        reinterpret_cast<const __half*>(wtc), b, reinterpret_cast<__half*>(y), B, R, nm));
  else
    CUDA_TRY(launch_pdl(stem_conv1_tc_kernel<float>, dim3((unsigned)grid), dim3(128), (size_t)kStemTcSmem, st, reinterpret_cast<const float*>(x),
                        reinterpret_cast<const __half*>(wtc), b, reinterpret_cast<__half*>(y), B, R, nm));
  return 0;
}

static int launch_stem_conv1(const void* x, int x_u8, const float* norm6, const float* w, const float* b, void* y, int B, int R, int Cout, cudaStream_t st) {
  if (R % 4) return fail(EMBCLIP_EINVAL, "stem: resolution must be a multiple of 4");
  const long long total = (long long)B * (R / 2) * (R / 4);      // one thread per PAIR of output pixels
  const int blocks = (int)((total + 127) / 128);
  const size_t smem = 0;
  if (blocks <= 0) return 0;
  StemNorm nm;
  for (int c = 0; c < 3; ++c) { nm.scale[c] = norm6 ? norm6[c] : 1.f; nm.offset[c] = norm6 ? norm6[3 + c] : 0.f; }
  if (Cout == 32 && x_u8)
    CUDA_TRY(launch_pdl(stem_conv1_kernel<32, uint8_t>, dim3(blocks), dim3(128), smem, st, reinterpret_cast<const uint8_t*>(x), w, b, reinterpret_cast<__half*>(y), B, R, nm));
  else if (Cout == 32)
    CUDA_TRY(launch_pdl(stem_conv1_kernel<32, float>, dim3(blocks), dim3(128), smem, st, reinterpret_cast<const float*>(x), w, b, reinterpret_cast<__half*>(y), B, R, nm));
  else
    return fail(EMBCLIP_EINVAL, "stem conv1: only Cout == 32 (width 64) is built");
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int embclip_stem_conv1(const float* frames, const float* w, const float* bias, void* out, int B, int R, int Cout,
                                  void* stream) {
  if (!frames || !w || !bias || !out) return fail(EMBCLIP_EINVAL, "stem_conv1: null pointer");
  return launch_stem_conv1(frames, 0, nullptr, w, bias, out, B, R, Cout, (cudaStream_t)stream);
}

// =============================================================================================
// ModifiedResNet plan
// =============================================================================================
namespace {

struct Act {            // workspace tensor, NHWC; batch dim scales with B
  std::string name;
  int dtype;            // EMBCLIP_DTYPE_*
  int h, w, c;          // per image
  int rows_per_image;   // h*w, or tokens etc.
  int hidden = 0;       // not materialised: the producer hands it to its fused consumers on chip (no workspace bytes)
};
struct Param {
  embclip_param_info info;
};
enum OpKind { K_STEM1, K_GEMM, K_POOL, K_TOKENS, K_ATTN_CORE, K_AVGHEAD, K_NCHW, K_IM2COL };
enum Head { H_TRUNK_ALWAYS = 0, H_NCHW = 1, H_AVG = 2, H_ATTN = 4 };
struct Op {
  OpKind kind;
  std::string name;
  int head = H_TRUNK_ALWAYS;     // run only if this head is requested (0 = always)
  int in0 = -1, in1 = -1, res = -1, out = -1;   // Act ids (-1: none; out -2/-3/-4: external outputs)
  int wp = -1, bp = -1;          // Param ids
  int taps = 1, c0 = 0, c1 = 0, cout = 0, relu = 0, out_f32 = 0;
  int pool = 0;                  // 3x3 only: 2x2 average pool fused into the epilogue
  int rows_mode = 0;             // 0: M = B*h*w of in0;  1: M = B (one row per image)
  int lda0 = 0, a_cols = 0, ldw = 0, w_rows = 0;
  int grp_n = 0, grp_a_koff = 0, grp_b_koff = 0, grp_b_nmod = 0;
  int force_bn = 0;
  int reverse = 0;               // tile walk direction (alternates layer to layer: snake order through L2)
  int fuse_next = -1;            // conv3 only: index of the next block's conv1 op, computed by the same launch (bneck_tail)
  int fuse_pool = -1;            // conv3 only: index of the K_POOL op whose output this launch writes INSTEAD of its own (bneck_tail kPool)
  int fuse_stream = 0;           // ... by bneck_tail_stream (weights streamed: layer 2) instead of bneck_tail (weights resident: layer 1)
  int side = 0;                  // independent of the ops that follow it: launched on the handle's side stream (fork / join)
  int fused_away = 0;            // conv1 only: produced by the previous block's bneck_tail launch, not launched itself
};

}  // namespace

struct embclip_rn50 {
  embclip_rn50_cfg cfg;
  std::vector<Act> acts;
  std::vector<Param> params;
  std::vector<Op> ops;
  uint64_t blob_bytes = 0;
  const uint8_t* blob = nullptr;
  int embed = 0, fres = 0, tokens = 0;
  int act_trunk_f32 = -1;
  int p_stem_wtc = -1;
  bool attn_ok = true;           // false: more than 64 tokens, the attention-pool head is not in the plan
  // fork / join plumbing for ops marked `side` (identity-branch pools, layout heads): they overlap the tensor-bound
  // kernels that do not depend on them.  Created lazily on the device of the first forward.
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  ~embclip_rn50() {
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (side_stream) cudaStreamDestroy(side_stream);
  }
};

static int add_act(embclip_rn50* m, const std::string& name, int dtype, int h, int w, int c) {
  m->acts.push_back(Act{name, dtype, h, w, c, h * w});
  return (int)m->acts.size() - 1;
}
static int add_param(embclip_rn50* m, const std::string& name, int dtype, std::initializer_list<int64_t> shape) {
  Param p;
  memset(&p.info, 0, sizeof p.info);
  snprintf(p.info.name, sizeof p.info.name, "%s", name.c_str());
  p.info.dtype = dtype;
  p.info.ndim = (int)shape.size();
  uint64_t n = 1;
  int i = 0;
  for (int64_t s : shape) { p.info.shape[i++] = s; n *= (uint64_t)s; }
  p.info.nbytes = n * (dtype == EMBCLIP_DTYPE_F16 ? 2 : 4);
  p.info.offset = m->blob_bytes;
  m->blob_bytes += (p.info.nbytes + 255) & ~uint64_t(255);
  m->params.push_back(p);
  return (int)m->params.size() - 1;
}
// conv (+folded BN) as GEMM: weights fp16 [cout, taps*c0 + c1], bias fp32 [cout]
static int add_conv(embclip_rn50* m, const std::string& name, int in0, int in1, int res, int taps, int cout, int relu,
                    int out_f32 = 0, int pool = 0) {
  const Act& a = m->acts[in0];
  Op op;
  op.kind = K_GEMM;
  op.name = name;
  op.in0 = in0; op.in1 = in1; op.res = res;
  op.taps = taps; op.c0 = a.c; op.c1 = in1 >= 0 ? m->acts[in1].c : 0; op.cout = cout; op.relu = relu; op.out_f32 = out_f32;
  op.lda0 = a.c;
  op.ldw = taps * op.c0 + op.c1; op.w_rows = cout;
  op.wp = add_param(m, name + ".w", EMBCLIP_DTYPE_F16, {cout, op.ldw});
  op.bp = add_param(m, name + ".b", EMBCLIP_DTYPE_F32, {cout});
  op.pool = pool;
  const int oh = pool ? a.h / 2 : a.h, ow = pool ? a.w / 2 : a.w;   // (a is a reference into m->acts: read before add_act)
  op.out = add_act(m, name, out_f32 ? EMBCLIP_DTYPE_F32 : EMBCLIP_DTYPE_F16, oh, ow, cout);
  m->ops.push_back(op);
  return op.out;
}
static int add_pool(embclip_rn50* m, const std::string& name, int in0, int mode = 1, int side = 1) {
  const Act a = m->acts[in0];
  Op op;
  op.kind = K_POOL;
  op.name = name;
  op.in0 = in0;
  op.pool = mode;
  op.side = side;
  op.out = add_act(m, name, EMBCLIP_DTYPE_F16, a.h / 2, a.w / 2, a.c);
  m->ops.push_back(op);
  return op.out;
}

extern "C" int embclip_rn50_create(const embclip_rn50_cfg* cfg, embclip_rn50_t* out) {
  if (!cfg || !out) return fail(EMBCLIP_EINVAL, "rn50_create: null argument");
  const int width = cfg->width, R = cfg->input_resolution;
  const bool tv = cfg->arch == 1;        // torchvision ResNet (v1.5 Bottleneck): 7x7/2 stem + max-pool, stride on the 3x3, strided 1x1 downsample
  if (cfg->arch != 0 && cfg->arch != 1) return fail(EMBCLIP_EINVAL, "rn50_create: arch must be 0 (CLIP ModifiedResNet) or 1 (torchvision ResNet)");
  if (tv && (width != 64 || cfg->output_dim != 0)) return fail(EMBCLIP_EINVAL, "rn50_create: the torchvision plan is width 64 without an attention-pool head (output_dim 0)");
  if (width != 64 && width != 96) return fail(EMBCLIP_EINVAL, "rn50_create: width 64 (RN50 / RN101) and 96 (RN50x16) are built, got %d", width);
  if (R <= 0 || R % 32) return fail(EMBCLIP_EINVAL, "rn50_create: input_resolution must be a positive multiple of 32");
  for (int i = 0; i < 4; ++i)
    if (cfg->layers[i] < 1) return fail(EMBCLIP_EINVAL, "rn50_create: layers[%d] < 1", i);
  const int embed = width * 32, fres = R / 32, L = fres * fres + 1;
  if (!tv && (cfg->heads <= 0 || embed % cfg->heads || embed / cfg->heads != 64))
    return fail(EMBCLIP_EINVAL, "rn50_create: head dim must be 64 (embed %d heads %d)", embed, cfg->heads);
  const bool attn_ok = L <= 64 && cfg->output_dim > 0;   // output_dim 0 = "no attention-pool head" (positional embedding of another resolution)      // (RN50x16 at its native 384 x 384 has 145 tokens: trunk and avg-pool heads only)
  if (cfg->output_dim < 0 || cfg->output_dim % 32) return fail(EMBCLIP_EINVAL, "rn50_create: output_dim must be a non-negative multiple of 32");

  embclip_rn50* m = new embclip_rn50();
  m->cfg = *cfg;
  m->embed = embed; m->fres = fres; m->tokens = L; m->attn_ok = attn_ok;
  // the stem's width/2 channels are carried in a multiple of 32 (width 96: 48 real + 16 zero-weight channels, exact)
  const int stem_c = (width / 2 + 31) / 32 * 32;

  // ---- stem
  int t;
  if (tv) {
    // conv 7x7 / 2 / pad 3 (3 -> 64) + BN + ReLU as im2col rows x [64, 160] GEMM, then MaxPool2d(3, 2, 1)
    Op ic;
    ic.kind = K_IM2COL;
    ic.name = "stem.im2col";
    ic.out = add_act(m, "stem.im2col", EMBCLIP_DTYPE_F16, R / 2, R / 2, kStem7K);
    m->ops.push_back(ic);
    t = add_conv(m, "stem.conv1", ic.out, -1, -1, 1, 64, 1);
    t = add_pool(m, "stem.maxpool", t, /*mode=*/3, /*side=*/0);
  } else {
    Op op;
    op.kind = K_STEM1;
    op.name = "stem.conv1";
    op.cout = stem_c;
    op.wp = add_param(m, "stem.conv1.w", EMBCLIP_DTYPE_F32, {27, stem_c});
    op.bp = add_param(m, "stem.conv1.b", EMBCLIP_DTYPE_F32, {stem_c});
    m->p_stem_wtc = add_param(m, "stem.conv1.wtc", EMBCLIP_DTYPE_F16, {stem_c, 128});   // hi/lo-split rows for the tensor-core stem
    op.out = add_act(m, "stem.conv1", EMBCLIP_DTYPE_F16, R / 2, R / 2, stem_c);
    m->ops.push_back(op);
    t = (int)m->acts.size() - 1;
    t = add_conv(m, "stem.conv2", t, -1, -1, 9, stem_c, 1);
    t = add_conv(m, "stem.conv3", t, -1, -1, 9, width, 1, 0, /*pool=*/1);   // AvgPool2d(2) fused into the epilogue
  }

  // ---- bottleneck stages
  int inplanes = width;
  (void)inplanes;
  for (int li = 0; li < 4; ++li) {
    const int planes = width << li;
    for (int bi = 0; bi < cfg->layers[li]; ++bi) {
      const int stride = (bi == 0 && li > 0) ? 2 : 1;
      const bool down = (bi == 0);   // stride > 1 or inplanes != planes*4: true exactly for the first block of a stage
      const bool last = (li == 3 && bi == cfg->layers[3] - 1);
      char pfx[48];
      snprintf(pfx, sizeof pfx, "layer%d.%d", li + 1, bi);
      const std::string P(pfx);
      const int x = t;
      int xp = x;
      // identity-branch AvgPool2d (CLIP) / stride-2 subsample (torchvision): needs only x, so it is issued first (side stream)
      if (stride == 2) xp = add_pool(m, P + ".xpool", x, tv ? 2 : 1);
      int a = add_conv(m, P + ".conv1", x, -1, -1, 1, planes, 1);
      // CLIP: avgpool(stride) after the stride-1 3x3, fused; torchvision: the 3x3 itself has the stride (epilogue keeps ::2)
      int b = add_conv(m, P + ".conv2", a, -1, -1, 9, planes, 1, 0, /*pool=*/stride == 2 ? (tv ? 2 : 1) : 0);
      // conv3 (+ downsample conv fused along K when the block has one, else identity residual)
      if (down) t = add_conv(m, P + ".conv3", b, xp, -1, 1, planes * 4, 1, last ? 1 : 0);
      else      t = add_conv(m, P + ".conv3", b, -1, x, 1, planes * 4, 1, last ? 1 : 0);
      inplanes = planes * 4;
    }
  }
  (void)inplanes;
  m->act_trunk_f32 = t;

  // ---- heads
  {
    Op op;
    op.kind = K_NCHW; op.name = "head.trunk_nchw"; op.head = H_NCHW; op.in0 = t; op.out = -2; op.side = 1;
    m->ops.push_back(op);
    Op op2;
    op2.kind = K_AVGHEAD; op2.name = "head.avgpool"; op2.head = H_AVG; op2.in0 = t; op2.out = -3; op2.side = 1;
    m->ops.push_back(op2);
  }
  if (attn_ok) {
    const int heads = cfg->heads, E = embed;
    const int p_pos = add_param(m, "attnpool.pos", EMBCLIP_DTYPE_F32, {L, E});
    Op tk;
    tk.kind = K_TOKENS; tk.name = "attnpool.tokens"; tk.head = H_ATTN; tk.in0 = t; tk.wp = p_pos;
    tk.out = add_act(m, "attnpool.tokens", EMBCLIP_DTYPE_F16, 1, L, E);
    m->ops.push_back(tk);
    // q = (Wq t0 + bq) / sqrt(64): A = tokens viewed as [B, L*E], K window = first E columns
    Op q;
    q.kind = K_GEMM; q.name = "attnpool.q"; q.head = H_ATTN; q.in0 = tk.out; q.rows_mode = 1;
    q.c0 = E; q.lda0 = L * E; q.a_cols = E; q.cout = E; q.ldw = E; q.w_rows = E;
    q.wp = add_param(m, "attnpool.q.w", EMBCLIP_DTYPE_F16, {E, E});
    q.bp = add_param(m, "attnpool.q.b", EMBCLIP_DTYPE_F32, {E});
    q.out = add_act(m, "attnpool.q", EMBCLIP_DTYPE_F16, 1, 1, E);
    m->ops.push_back(q);
    // qt[b, h, :] = Wk_h^T q_h : grouped GEMM, N = heads*E, group = head, K = 64
    Op qt;
    qt.kind = K_GEMM; qt.name = "attnpool.qk"; qt.head = H_ATTN; qt.in0 = q.out; qt.rows_mode = 1;
    qt.c0 = 64; qt.lda0 = E; qt.a_cols = E; qt.cout = heads * E; qt.ldw = E; qt.w_rows = E;
    qt.grp_n = E; qt.grp_a_koff = 64; qt.grp_b_koff = 64; qt.grp_b_nmod = E;
    qt.wp = add_param(m, "attnpool.kT.w", EMBCLIP_DTYPE_F16, {E, E});   // Wk transposed: [c, (h,d)]
    qt.out = add_act(m, "attnpool.qk", EMBCLIP_DTYPE_F16, 1, heads, E);
    m->ops.push_back(qt);
    Op core;
    core.kind = K_ATTN_CORE; core.name = "attnpool.core"; core.head = H_ATTN; core.in0 = qt.out; core.in1 = tk.out;
    core.out = add_act(m, "attnpool.xbar", EMBCLIP_DTYPE_F16, 1, heads, E);
    m->ops.push_back(core);
    // o[b, (h,d)] = Wv_h xbar_h + bv : grouped GEMM, N = E, group = head (64 columns), K = E at A offset h*E
    Op v;
    v.kind = K_GEMM; v.name = "attnpool.v"; v.head = H_ATTN; v.in0 = core.out; v.rows_mode = 1;
    v.c0 = E; v.lda0 = heads * E; v.a_cols = heads * E; v.cout = E; v.ldw = E; v.w_rows = E;
    v.grp_n = 64; v.grp_a_koff = E; v.grp_b_koff = 0; v.grp_b_nmod = 0;
    v.wp = add_param(m, "attnpool.v.w", EMBCLIP_DTYPE_F16, {E, E});
    v.bp = add_param(m, "attnpool.v.b", EMBCLIP_DTYPE_F32, {E});
    v.out = add_act(m, "attnpool.v", EMBCLIP_DTYPE_F16, 1, 1, E);
    m->ops.push_back(v);
    Op c;
    c.kind = K_GEMM; c.name = "attnpool.c"; c.head = H_ATTN; c.in0 = v.out; c.rows_mode = 1;
    c.c0 = E; c.lda0 = E; c.cout = cfg->output_dim; c.ldw = E; c.w_rows = cfg->output_dim; c.out_f32 = 1;
    c.wp = add_param(m, "attnpool.c.w", EMBCLIP_DTYPE_F16, {cfg->output_dim, E});
    c.bp = add_param(m, "attnpool.c.b", EMBCLIP_DTYPE_F32, {cfg->output_dim});
    c.out = -4;
    m->ops.push_back(c);
  }
  // snake order: consecutive tensor-core layers walk their tiles in opposite directions, so each starts on the part
  // of its input that the previous layer wrote last (still resident in the 126 MB L2)
  // bneck_tail fusion (bneck_tail.cuh): a 64 -> 256 conv3 followed by the next block's 256 -> 64/128 conv1 on the same
  // pixels is ONE launch; the 256-channel tensor is written once and not re-read by the conv1
  static const bool tail_fuse = getenv("EMBCLIP_NO_TAILFUSE") == nullptr;
  if (tail_fuse) {
    for (size_t i = 0; i + 1 < m->ops.size(); ++i) {
      Op& c3 = m->ops[i];
      size_t j = i + 1;
      if (m->ops[j].kind == K_POOL && j + 1 < m->ops.size()) ++j;     // the next block's identity pool sits between them
      Op& c1 = m->ops[j];
      if (c3.kind != K_GEMM || c1.kind != K_GEMM || c3.rows_mode || c1.rows_mode || c3.head || c1.head) continue;
      if (c3.taps != 1 || c3.c0 != 64 || c3.cout != 256 || !c3.relu || c3.out_f32 || c3.grp_n) continue;
      if (!((c3.in1 >= 0 && c3.c1 == 64 && c3.res < 0) || (c3.in1 < 0 && c3.res >= 0))) continue;
      if (c1.taps != 1 || c1.in0 != c3.out || c1.in1 >= 0 || c1.res >= 0 || !c1.relu || c1.out_f32 || c1.grp_n) continue;
      if (c1.c0 != 256 || (c1.cout != 64 && c1.cout != 128) || (c3.in1 >= 0 && c1.cout != 64)) continue;
      c3.fuse_next = (int)j;
      c1.fused_away = 1;
      // ... and when that pool is the ONLY other reader of x' (the next stage's downsample branch), the launch writes the pooled
      // tensor instead of x' (bneck_tail kPool): x' is never materialised
      static const bool tail_pool = getenv("EMBCLIP_NO_TAILPOOL") == nullptr;
      Op& pl = m->ops[i + 1];
      const Act& xa = m->acts[c3.out];
      if (tail_pool && j == i + 2 && pl.kind == K_POOL && pl.in0 == c3.out && (pl.pool == 0 || pl.pool == 1 || pl.pool == 2) &&
          c3.res >= 0 && c1.cout == 128 && xa.w % 2 == 0 && xa.h % 2 == 0 && 2 * xa.w <= 128 && (2 * xa.w) % 8 == 0) {
        bool other = false;
        for (size_t k = 0; k < m->ops.size(); ++k)
          if (k != i + 1 && k != j && (m->ops[k].in0 == c3.out || m->ops[k].in1 == c3.out || m->ops[k].res == c3.out)) other = true;
        if (!other) {
          c3.fuse_pool = (int)(i + 1);
          pl.fused_away = 1;
          m->acts[c3.out].hidden = 1;
        }
      }
    }
    // the same fusion where the weights must stream (layer 2 identity blocks: conv3 128 -> 512, next conv1 512 -> 128)
    static const bool stream_fuse = getenv("EMBCLIP_NO_TAILSTREAM") == nullptr;
    for (size_t i = 0; stream_fuse && i + 1 < m->ops.size(); ++i) {
      Op& c3 = m->ops[i];
      Op& c1 = m->ops[i + 1];
      if (c3.kind != K_GEMM || c1.kind != K_GEMM || c3.rows_mode || c1.rows_mode || c3.head || c1.head || c3.fuse_next >= 0) continue;
      if (c3.taps != 1 || c3.in1 >= 0 || c3.res < 0 || !c3.relu || c3.out_f32 || c3.grp_n) continue;
      if (c1.taps != 1 || c1.in0 != c3.out || c1.in1 >= 0 || c1.res >= 0 || !c1.relu || c1.out_f32 || c1.grp_n || c1.c0 != c3.cout) continue;
      if (!tail_stream_supported(c3.c0, c3.cout, c1.cout)) continue;
      c3.fuse_next = (int)(i + 1);
      c3.fuse_stream = 1;
      c1.fused_away = 1;
    }
  }
  static const bool snake = getenv("EMBCLIP_NO_SNAKE") == nullptr;
  int dir = 0;
  for (Op& op : m->ops)
    if (op.kind == K_GEMM && op.rows_mode == 0 && !op.fused_away) { op.reverse = snake ? dir : 0; dir ^= 1; }
  *out = m;
  return 0;
}

extern "C" int embclip_rn50_destroy(embclip_rn50_t h) {
  delete h;
  return 0;
}
extern "C" int embclip_rn50_num_params(embclip_rn50_t h) { return h ? (int)h->params.size() : fail(EMBCLIP_EINVAL, "null handle"); }
extern "C" int embclip_rn50_param_info(embclip_rn50_t h, int index, embclip_param_info* out) {
  if (!h || !out || index < 0 || index >= (int)h->params.size()) return fail(EMBCLIP_EINVAL, "param_info: bad argument");
  *out = h->params[index].info;
  return 0;
}
extern "C" uint64_t embclip_rn50_blob_bytes(embclip_rn50_t h) { return h ? h->blob_bytes : 0; }
extern "C" int embclip_rn50_bind_weights(embclip_rn50_t h, const void* device_blob, uint64_t nbytes) {
  if (!h || !device_blob) return fail(EMBCLIP_EINVAL, "bind_weights: null argument");
  if (nbytes < h->blob_bytes) return fail(EMBCLIP_EINVAL, "bind_weights: blob holds %llu bytes, plan needs %llu",
                                          (unsigned long long)nbytes, (unsigned long long)h->blob_bytes);
  if (reinterpret_cast<uintptr_t>(device_blob) % 256) return fail(EMBCLIP_EINVAL, "bind_weights: blob must be 256-B aligned");
  h->blob = reinterpret_cast<const uint8_t*>(device_blob);
  return 0;
}

static uint64_t act_bytes(const Act& a, int B) {
  if (a.hidden) return 0;
  const uint64_t n = (uint64_t)B * a.h * a.w * a.c * (a.dtype == EMBCLIP_DTYPE_F16 ? 2 : 4);
  return (n + 1023) & ~uint64_t(1023);
}
static uint64_t act_offset(const embclip_rn50* m, int B, int id) {
  uint64_t off = 0;
  for (int i = 0; i < id; ++i) off += act_bytes(m->acts[i], B);
  return off;
}
extern "C" uint64_t embclip_rn50_workspace_bytes(embclip_rn50_t h, int batch) {
  if (!h || batch <= 0) return 0;
  return act_offset(h, batch, (int)h->acts.size());
}
extern "C" int embclip_rn50_num_acts(embclip_rn50_t h) { return h ? (int)h->acts.size() : fail(EMBCLIP_EINVAL, "null handle"); }
extern "C" int embclip_rn50_act_info(embclip_rn50_t h, int batch, int index, embclip_act_info* out) {
  if (!h || !out || index < 0 || index >= (int)h->acts.size() || batch <= 0) return fail(EMBCLIP_EINVAL, "act_info: bad argument");
  const Act& a = h->acts[index];
  memset(out, 0, sizeof *out);
  snprintf(out->name, sizeof out->name, "%s", a.name.c_str());
  out->dtype = a.dtype; out->n = a.hidden ? 0 : batch; out->h = a.h; out->w = a.w; out->c = a.c;
  out->offset = act_offset(h, batch, index);
  return 0;
}

struct FramesIn { const void* ptr; int u8; float norm[6]; void* rows_f16 = nullptr; };   // rows_f16: see embclip_rn50_encode_rows_f16
static int run_op(embclip_rn50* m, const Op& op, const std::vector<uint64_t>& offs, const FramesIn& frames, int B,
                  float* o_nchw, float* o_avg, float* o_attn, uint8_t* ws, cudaStream_t st) {
  auto act_ptr = [&](int id) -> void* { return id >= 0 ? (void*)(ws + offs[id]) : nullptr; };
  auto param_ptr = [&](int id) -> const void* { return id >= 0 ? (const void*)(m->blob + m->params[id].info.offset) : nullptr; };
  const int P = m->fres * m->fres;
  switch (op.kind) {
    case K_STEM1: {
      static const bool cuda_core = getenv("EMBCLIP_STEM_CUDA_CORE") != nullptr;       // first version, kept for A/B timing
      if ((!cuda_core && op.cout == 32) || op.cout == 64)
        return launch_stem_conv1_tc(frames.ptr, frames.u8, frames.u8 ? frames.norm : nullptr, param_ptr(m->p_stem_wtc),
                                    (const float*)param_ptr(op.bp), act_ptr(op.out), B, m->cfg.input_resolution, op.cout, st);
      return launch_stem_conv1(frames.ptr, frames.u8, frames.u8 ? frames.norm : nullptr, (const float*)param_ptr(op.wp),
                               (const float*)param_ptr(op.bp), act_ptr(op.out), B, m->cfg.input_resolution, op.cout, st);
    }
    case K_POOL: {
      const Act& a = m->acts[op.in0];
      return launch_avgpool2(act_ptr(op.in0), act_ptr(op.out), B, a.h, a.w, a.c, st, op.pool ? op.pool : 1);
    }
    case K_IM2COL:
      return launch_im2col7(frames.ptr, frames.u8, frames.u8 ? frames.norm : nullptr, act_ptr(op.out), B, m->cfg.input_resolution, st);
    case K_GEMM: {
      const Act& a = m->acts[op.in0];
      if (op.taps == 9) {
        Conv3Op c{act_ptr(op.in0), param_ptr(op.wp), (const float*)param_ptr(op.bp), act_ptr(op.out), B, a.h, a.w, op.c0, op.cout, op.relu, op.pool};
        c.reverse = op.reverse;
        return launch_conv3x3_halo(c, st);
      }
      if (op.fuse_next >= 0 && op.fuse_stream) {
        const Op& c1 = m->ops[op.fuse_next];
        TailStreamOp t;
        t.a = act_ptr(op.in0); t.w3 = param_ptr(op.wp); t.b3 = (const float*)param_ptr(op.bp); t.residual = act_ptr(op.res);
        t.out = act_ptr(op.out); t.w1 = param_ptr(c1.wp); t.b1 = (const float*)param_ptr(c1.bp); t.y1 = act_ptr(c1.out);
        t.M = (long long)B * a.h * a.w; t.k3 = op.c0; t.n3 = op.cout; t.n1 = c1.cout; t.reverse = op.reverse;
        return launch_bneck_tail_stream(t, st);
      }
      if (op.fuse_next >= 0) {
        const Op& c1 = m->ops[op.fuse_next];
        TailOp t;
        t.a0 = act_ptr(op.in0); t.a1 = act_ptr(op.in1); t.w3 = param_ptr(op.wp); t.b3 = (const float*)param_ptr(op.bp);
        t.residual = act_ptr(op.res); t.out = act_ptr(op.out);
        t.w1 = param_ptr(c1.wp); t.b1 = (const float*)param_ptr(c1.bp); t.y1 = act_ptr(c1.out);
        t.M = (long long)B * a.h * a.w; t.n1 = c1.cout; t.reverse = op.reverse;
        if (op.fuse_pool >= 0) {
          const Op& pl = m->ops[op.fuse_pool];
          t.out = nullptr; t.pool_out = act_ptr(pl.out); t.pool_mode = pl.pool ? pl.pool : 1; t.pool_w2 = a.w;
        }
        return launch_bneck_tail(t, st);
      }
      GemmOp g;
      g.a0 = act_ptr(op.in0);
      if (op.rows_mode == 1) { g.n = 1; g.h = 1; g.w = B; }
      else if (op.taps == 9) { g.n = B; g.h = a.h; g.w = a.w; }
      else { g.n = 1; g.h = 1; g.w = B * a.h * a.w; }
      g.c0 = op.c0; g.lda0 = op.lda0; g.a_cols = op.a_cols; g.taps = op.taps;
      g.a1 = act_ptr(op.in1); g.c1 = op.c1;
      g.wgt = param_ptr(op.wp); g.ldw = op.ldw; g.w_rows = op.w_rows;
      g.bias = (const float*)param_ptr(op.bp);
      g.residual = act_ptr(op.res);
      g.out = op.out == -4 ? (void*)o_attn : act_ptr(op.out);
      g.cout = op.cout; g.relu = op.relu; g.out_f32 = op.out_f32;
      // rows-only forward: the last conv rounds its fp32 result to fp16 in its own epilogue and stores the policy's pixel rows
      // (the fp32 NHWC tensor is not written: no head reads it in this mode)
      if (frames.rows_f16 && op.out == m->act_trunk_f32) { g.out = frames.rows_f16; g.out_f32 = 0; }
      g.grp_n = op.grp_n; g.grp_a_koff = op.grp_a_koff; g.grp_b_koff = op.grp_b_koff; g.grp_b_nmod = op.grp_b_nmod;
      g.reverse = op.reverse;
      return launch_gemm(g, st, op.force_bn);
    }
    case K_TOKENS: {
      dim3 grid((m->embed / 4 + 255) / 256, B);
      CUDA_TRY(launch_pdl(attnpool_tokens_kernel, grid, dim3(256), 0, st, (const float*)act_ptr(op.in0), (const float*)param_ptr(op.wp),
                          (__half*)act_ptr(op.out), P, m->embed));
      return 0;
    }
    case K_ATTN_CORE: {
      static const bool cuda_core = getenv("EMBCLIP_ATTNPOOL_CUDA_CORE") != nullptr;   // first version, kept for A/B timing
      if (!cuda_core && m->cfg.heads == 32 && 4 * m->tokens <= 256 && m->embed % 128 == 0) {
        { const int rc_ = ensure_smem((const void*)attnpool_tc_kernel, (size_t)(kApSmem)); if (rc_) return rc_; }
        CUtensorMap tq, tk, tmn;
        int rc;
        if ((rc = make_map_2d(&tq, act_ptr(op.in0), B * m->cfg.heads, m->embed, m->embed, 64, 128))) return rc;
        if ((rc = make_map_2d(&tk, act_ptr(op.in1), B * m->tokens, m->embed, m->embed, 64, 256))) return rc;
        if ((rc = make_map_2d(&tmn, act_ptr(op.in1), B * m->tokens, m->embed, m->embed, 64, 256))) return rc;
        AttnPoolTcParams ap;
        ap.B = B; ap.heads = m->cfg.heads; ap.L = m->tokens; ap.C = m->embed; ap.xbar = (__half*)act_ptr(op.out);
        // channel-chunk split: the largest power of two that still gives every CTA an SM (B = 256: 64 tiles x 2; B = 8: 2 x 16)
        const int tiles = (B + 3) / 4, nch = m->embed / 128;
        int split = 1;
        while (split * 2 <= nch && nch % (split * 2) == 0 && tiles * split * 2 <= num_sms()) split *= 2;
        CUDA_TRY(launch_pdl(attnpool_tc_kernel, dim3(tiles, split), dim3(128), (size_t)kApSmem, st, tq, tk, tmn, ap));
        return 0;
      }
      constexpr int HG = 8;
      if (m->cfg.heads % HG) return fail(EMBCLIP_EINVAL, "attention pool: heads must be a multiple of %d", HG);
      const size_t smem = (size_t)HG * m->embed * 2 + HG * 64 * 4;
      { const int rc_ = ensure_smem((const void*)attnpool_core_kernel<HG>, (size_t)(smem)); if (rc_) return rc_; }
      dim3 grid(B, m->cfg.heads / HG);
      CUDA_TRY(launch_pdl(attnpool_core_kernel<HG>, grid, dim3(256), smem, st, (const __half*)act_ptr(op.in0), (const __half*)act_ptr(op.in1),
                          (__half*)act_ptr(op.out), m->cfg.heads, m->tokens, m->embed));
      return 0;
    }
    case K_AVGHEAD: {
      dim3 grid((m->embed / 4 + 255) / 256, B);
      CUDA_TRY(launch_pdl(avg_head_kernel, grid, dim3(256), 0, st, (const float*)act_ptr(op.in0), o_avg, P, m->embed));
      return 0;
    }
    case K_NCHW: {
      const size_t wide_smem = (size_t)P * 128 * 4;
      if (m->embed % 128 == 0 && wide_smem <= 200u * 1024u && reinterpret_cast<uintptr_t>(o_nchw) % 16 == 0) {
        { const int rc_ = ensure_smem((const void*)nhwc_to_nchw_f32_wide_kernel, wide_smem); if (rc_) return rc_; }
        dim3 grid(m->embed / 128, B);
        CUDA_TRY(launch_pdl(nhwc_to_nchw_f32_wide_kernel, grid, dim3(256), wide_smem, st, (const float*)act_ptr(op.in0), o_nchw, P, m->embed));
        return 0;
      }
      dim3 grid(m->embed / 32, B);
      CUDA_TRY(launch_pdl(nhwc_to_nchw_f32_kernel, grid, dim3(256), (size_t)P * 33 * 4, st, (const float*)act_ptr(op.in0), o_nchw, P, m->embed));
      return 0;
    }
  }
  return fail(EMBCLIP_EINVAL, "unknown op kind");
}

static int forward_impl(embclip_rn50* m, const FramesIn& frames, int B, float* o_nchw, float* o_avg, float* o_attn, void* ws,
                        uint64_t ws_bytes, cudaStream_t st, float* op_ms, char* names, int max_ops) {
  if (!m || !frames.ptr || !ws || B <= 0) return fail(EMBCLIP_EINVAL, "forward: null argument or empty batch");
  if (!m->blob) return fail(EMBCLIP_ESTATE, "forward: weights not bound (call embclip_rn50_bind_weights first)");
  if (o_attn && !m->attn_ok) return fail(EMBCLIP_EINVAL, "forward: the attention-pool head is built for at most 64 tokens (this plan has %d)", m->tokens);
  const uint64_t need = embclip_rn50_workspace_bytes(m, B);
  if (ws_bytes < need) return fail(EMBCLIP_ENOSPC, "forward: workspace %llu B < required %llu B", (unsigned long long)ws_bytes, (unsigned long long)need);
  if (reinterpret_cast<uintptr_t>(ws) % 1024) return fail(EMBCLIP_EINVAL, "forward: workspace must be 1024-B aligned");
  std::vector<uint64_t> offs(m->acts.size());
  uint64_t off = 0;
  for (size_t i = 0; i < m->acts.size(); ++i) { offs[i] = off; off += act_bytes(m->acts[i], B); }
  const int want = (o_nchw ? H_NCHW : 0) | (o_avg ? H_AVG : 0) | (o_attn ? H_ATTN : 0);
  std::vector<cudaEvent_t> ev;
  int nrun = 0;
  // side-stream ops (not when profiling per op: the events would time an empty main stream)
  static const bool side_ok = getenv("EMBCLIP_NO_SIDE") == nullptr;
  const bool use_side = side_ok && !op_ms;
  if (use_side && !m->side_stream) {
    CUDA_TRY(cudaStreamCreateWithFlags(&m->side_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming));
  }
  bool forked = false;                     // side stream holds work the main stream has not joined yet
  std::vector<int> side_outs;              // acts produced by that work
  auto join = [&]() -> int {
    CUDA_TRY(cudaEventRecord(m->ev_join, m->side_stream));
    CUDA_TRY(cudaStreamWaitEvent(st, m->ev_join, 0));
    forked = false;
    side_outs.clear();
    return 0;
  };
  for (const Op& op : m->ops) {
    if (op.head && !(op.head & want)) continue;
    if (op.fused_away) continue;
    if (use_side && op.side) {
      // everything enqueued on `st` so far (in particular this op's input) precedes the side work
      CUDA_TRY(cudaEventRecord(m->ev_fork, st));
      CUDA_TRY(cudaStreamWaitEvent(m->side_stream, m->ev_fork, 0));
      const int rc = run_op(m, op, offs, frames, B, o_nchw, o_avg, o_attn, reinterpret_cast<uint8_t*>(ws), m->side_stream);
      if (rc) return rc;
      forked = true;
      if (op.out >= 0) side_outs.push_back(op.out);
      ++nrun;
      continue;
    }
    if (forked) {
      bool dep = false;
      for (int id : side_outs) dep = dep || id == op.in0 || id == op.in1 || id == op.res;
      if (dep) { const int rc = join(); if (rc) return rc; }
    }
    if (op_ms) {
      cudaEvent_t e;
      CUDA_TRY(cudaEventCreate(&e));
      CUDA_TRY(cudaEventRecord(e, st));
      ev.push_back(e);
      if (names && nrun < max_ops) snprintf(names + (size_t)nrun * 64, 64, "%s", op.name.c_str());
    }
    const int rc = run_op(m, op, offs, frames, B, o_nchw, o_avg, o_attn, reinterpret_cast<uint8_t*>(ws), st);
    if (rc) return rc;
    ++nrun;
  }
  if (forked) { const int rc = join(); if (rc) return rc; }    // caller's stream order covers the side work
  if (op_ms) {
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreate(&e));
    CUDA_TRY(cudaEventRecord(e, st));
    ev.push_back(e);
    CUDA_TRY(cudaEventSynchronize(e));
    for (int i = 0; i < nrun && i < max_ops; ++i) CUDA_TRY(cudaEventElapsedTime(&op_ms[i], ev[i], ev[i + 1]));
    for (cudaEvent_t x : ev) cudaEventDestroy(x);
  }
  return nrun;
}

extern "C" int embclip_rn50_forward(embclip_rn50_t h, const float* frames_nhwc, int batch, float* out_trunk_nchw,
                                    float* out_avgpool, float* out_attnpool, void* workspace, uint64_t workspace_bytes,
                                    void* stream) {
  EMBCLIP_TRACE();
  const FramesIn in{frames_nhwc, 0, {1, 1, 1, 0, 0, 0}};
  const int rc = forward_impl(h, in, batch, out_trunk_nchw, out_avgpool, out_attnpool, workspace, workspace_bytes,
                              (cudaStream_t)stream, nullptr, nullptr, 0);
  return rc < 0 ? rc : 0;
}
extern "C" int embclip_rn50_forward_u8(embclip_rn50_t h, const uint8_t* frames_nhwc_u8, const float* mean3, const float* std3, int batch,
                                       float* out_trunk_nchw, float* out_avgpool, float* out_attnpool, void* workspace,
                                       uint64_t workspace_bytes, void* stream) {
  EMBCLIP_TRACE();
  if (!mean3 || !std3) return fail(EMBCLIP_EINVAL, "forward_u8: mean / std required");
  FramesIn in{frames_nhwc_u8, 1, {0, 0, 0, 0, 0, 0}};
  for (int c = 0; c < 3; ++c) {
    if (!(std3[c] > 0.f)) return fail(EMBCLIP_EINVAL, "forward_u8: std must be positive");
    in.norm[c] = 1.f / (255.f * std3[c]);
    in.norm[3 + c] = -mean3[c] / std3[c];
  }
  const int rc = forward_impl(h, in, batch, out_trunk_nchw, out_avgpool, out_attnpool, workspace, workspace_bytes,
                              (cudaStream_t)stream, nullptr, nullptr, 0);
  return rc < 0 ? rc : 0;
}
extern "C" int embclip_rn50_profile(embclip_rn50_t h, const float* frames_nhwc, int batch, float* out_trunk_nchw,
                                    float* out_avgpool, float* out_attnpool, void* workspace, uint64_t workspace_bytes,
                                    void* stream, float* op_ms, char* names, int max_ops) {
  EMBCLIP_TRACE();
  if (!op_ms || max_ops <= 0) return fail(EMBCLIP_EINVAL, "profile: need op_ms buffer");
  const FramesIn in{frames_nhwc, 0, {1, 1, 1, 0, 0, 0}};
  return forward_impl(h, in, batch, out_trunk_nchw, out_avgpool, out_attnpool, workspace, workspace_bytes,
                      (cudaStream_t)stream, op_ms, names, max_ops);
}
extern "C" int embclip_rn50_profile_u8(embclip_rn50_t h, const uint8_t* frames_nhwc_u8, const float* mean3, const float* std3, int batch,
                                       float* out_trunk_nchw, float* out_avgpool, float* out_attnpool, void* workspace,
                                       uint64_t workspace_bytes, void* stream, float* op_ms, char* names, int max_ops) {
  EMBCLIP_TRACE();
  if (!op_ms || max_ops <= 0) return fail(EMBCLIP_EINVAL, "profile: need op_ms buffer");
  if (!mean3 || !std3) return fail(EMBCLIP_EINVAL, "profile_u8: mean / std required");
  FramesIn in{frames_nhwc_u8, 1, {0, 0, 0, 0, 0, 0}};
  for (int c = 0; c < 3; ++c) {
    if (!(std3[c] > 0.f)) return fail(EMBCLIP_EINVAL, "profile_u8: std must be positive");
    in.norm[c] = 1.f / (255.f * std3[c]);
    in.norm[3 + c] = -mean3[c] / std3[c];
  }
  return forward_impl(h, in, batch, out_trunk_nchw, out_avgpool, out_attnpool, workspace, workspace_bytes,
                      (cudaStream_t)stream, op_ms, names, max_ops);
}
extern "C" int embclip_rn50_encode_rows_f16(embclip_rn50_t h, const void* frames_nhwc, int frames_are_u8, const float* mean3, const float* std3,
                                            int batch, void* out_rows_f16, void* workspace, uint64_t workspace_bytes, void* stream) {
  EMBCLIP_TRACE();
  if (!out_rows_f16) return fail(EMBCLIP_EINVAL, "encode_rows: null output");
  if (reinterpret_cast<uintptr_t>(out_rows_f16) % 16) return fail(EMBCLIP_EINVAL, "encode_rows: output must be 16-B aligned");
  FramesIn in{frames_nhwc, frames_are_u8 ? 1 : 0, {1, 1, 1, 0, 0, 0}};
  if (frames_are_u8) {
    if (!mean3 || !std3) return fail(EMBCLIP_EINVAL, "encode_rows: mean / std required for uint8 frames");
    for (int c = 0; c < 3; ++c) {
      if (!(std3[c] > 0.f)) return fail(EMBCLIP_EINVAL, "encode_rows: std must be positive");
      in.norm[c] = 1.f / (255.f * std3[c]);
      in.norm[3 + c] = -mean3[c] / std3[c];
    }
  }
  in.rows_f16 = out_rows_f16;
  const int rc = forward_impl(h, in, batch, nullptr, nullptr, nullptr, workspace, workspace_bytes, (cudaStream_t)stream, nullptr, nullptr, 0);
  return rc < 0 ? rc : 0;
}
extern "C" int embclip_rn50_export_rows_f16(embclip_rn50_t h, int batch, const void* workspace, uint64_t workspace_bytes,
                                            void* out_rows_f16, void* stream) {
  EMBCLIP_TRACE();
  if (!h || !workspace || !out_rows_f16 || batch <= 0) return fail(EMBCLIP_EINVAL, "export_rows: null argument or empty batch");
  if (workspace_bytes < embclip_rn50_workspace_bytes(h, batch)) return fail(EMBCLIP_ENOSPC, "export_rows: workspace too small for batch %d", batch);
  if (reinterpret_cast<uintptr_t>(out_rows_f16) % 16) return fail(EMBCLIP_EINVAL, "export_rows: output must be 16-B aligned");
  const float* src = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(workspace) + act_offset(h, batch, h->act_trunk_f32));
  const long long n8 = (long long)batch * h->fres * h->fres * h->embed / 8;
  long long blocks = (n8 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  CUDA_TRY(launch_pdl(cast_f32_f16_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, src, reinterpret_cast<__half*>(out_rows_f16), n8));
  return 0;
}
extern "C" int embclip_rn50_launches_per_forward(embclip_rn50_t h, int want_trunk, int want_avgpool, int want_attnpool) {
  if (!h) return fail(EMBCLIP_EINVAL, "null handle");
  const int want = (want_trunk ? H_NCHW : 0) | (want_avgpool ? H_AVG : 0) | (want_attnpool ? H_ATTN : 0);
  int n = 0;
  for (const Op& op : h->ops)
    if ((!op.head || (op.head & want)) && !op.fused_away) ++n;
  return n;
}
