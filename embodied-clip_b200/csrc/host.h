// Host-side helpers shared by the translation units of libembclip_b200.so (embclip.cu: encoder plan and GEMM
// launchers; ac_path.cu: actor-critic / PPO-update plan).  Internal: nothing here is part of the C ABI.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <utility>

#include <nvtx3/nvToolsExt.h>

#include "../../include/embclip_b200.h"

namespace embclip {

int fail(int code, const char* fmt, ...);
#define CUDA_TRY(expr)                                                                                  \
  do {                                                                                                  \
    cudaError_t e_ = (expr);                                                                            \
    if (e_ != cudaSuccess) return ::embclip::fail(EMBCLIP_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

// NVTX range over one C-ABI entry point (SURVEY.md section 5 "tracing"): shows up as a named span on the host timeline of
// Nsight Systems / any NVTX consumer; a no-op (one relaxed load) when no tool is attached.  Header-only NVTX v3: no link dependency.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
#define EMBCLIP_TRACE() ::embclip::NvtxRange nvtx_range_(__func__)

int num_sms();         // SM count of the CURRENT device (cached per device ordinal)
bool pdl_enabled();   // false when $EMBCLIP_NO_PDL is set
// Raises the kernel's MaxDynamicSharedMemorySize to `bytes` on the CURRENT device if it is not there yet.  The attribute is
// per (function, device): state is keyed on both, so a second GPU in the same process gets its own opt-in.  Thread-safe.
int ensure_smem(const void* func, size_t bytes);

// Launch with programmatic stream serialization: the kernel MUST execute griddepcontrol.wait before touching global
// memory written by earlier kernels in the stream (see ptx.cuh).
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// fp16 tensor maps; dims fastest-first, `pitch[i]` = byte stride of dim i+1; swizzle follows the box's inner bytes
int make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* pitch, const uint32_t* box);
int make_map_2d(CUtensorMap* m, const void* base, int rows, int cols, int ld, int box_cols, int box_rows);

struct GemmOp {
  // A0: NHWC view
  const void* a0 = nullptr;
  int n = 1, h = 1, w = 1, c0 = 0, lda0 = 0;   // c0 = channels of source 0 (K per tap); lda0 pixel pitch (elements)
  int taps = 1;
  // A1: optional 2-D source [M, c1]
  const void* a1 = nullptr;
  int c1 = 0;
  // weights [w_rows, ldw] (row n holds K values), bias
  const void* wgt = nullptr;
  int ldw = 0, w_rows = 0;
  const float* bias = nullptr;
  const void* residual = nullptr;   // fp16 [M, cout]
  int res_mode = 0;                 // 0: add residual; 1: zero the output where residual <= 0 (ReLU backward)
  void* out = nullptr;              // fp16 NHWC [n,h,w,cout] or fp32 [M, cout]
  int cout = 0;
  int relu = 0, out_f32 = 0;        // relu: activation code (0 none, 1 ReLU, 2 QuickGELU)
  const float* res_f32 = nullptr;   // out_f32 only: fp32 residual [M, cout] added in the epilogue (may alias out)
  int grp_n = 0, grp_a_koff = 0, grp_b_koff = 0, grp_b_nmod = 0;
  int reverse = 0;                  // walk the output tiles last-to-first
  int a_cols = 0;                   // logical width of an A0 row for the tensor map (>= c0; grouped mode: full row)
};
int launch_gemm(const GemmOp& op, cudaStream_t st, int force_bn = 0);

}  // namespace embclip
