// Experiment: does the UMMA shared-memory descriptor's base_offset field (bits [49,52)) let a K-major
// swizzled A operand start at a row that is NOT a multiple of 8 (i.e. not aligned to the swizzle repeat)?
// The implicit-GEMM 3x3 conv wants this: one halo tile in smem, 9 taps = 9 row-shifted windows of it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o exp_baseoffset exp_baseoffset.cu && ./exp_baseoffset
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include "../embodied-clip_b200/csrc/ptx.cuh"

using namespace embclip;

template <int SWZ>  // 128 or 64: bytes per row (K = SWZ/2 fp16)
__global__ void __launch_bounds__(128, 1) probe(int* mism /* [24][8] */, float* dump) {
  constexpr int K = SWZ / 2;
  constexpr int ROWS = 128 + 32;
  constexpr int N = 64;
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  const uint32_t sA = base, sB = base + ROWS * SWZ;          // ROWS*SWZ is a multiple of 1024
  const uint32_t bar = sB + N * SWZ, slot = bar + 8;
  auto Aval = [](int i, int k) { return float(((i * 7 + k * 3) % 13) - 6); };
  auto Bval = [](int n, int k) { return float(((n * 5 + k) % 7) - 3); };
  for (int idx = threadIdx.x; idx < ROWS * K; idx += blockDim.x) {
    const int i = idx / K, k = idx % K;
    const uint32_t off = swizzle_off<SWZ>(i, k / 8) + (k % 8) * 2;
    *reinterpret_cast<__half*>(gen + off) = __float2half(Aval(i, k));
  }
  for (int idx = threadIdx.x; idx < N * K; idx += blockDim.x) {
    const int n = idx / K, k = idx % K;
    const uint32_t off = swizzle_off<SWZ>(n, k / 8) + (k % 8) * 2;
    *reinterpret_cast<__half*>(gen + (sB - base) + off) = __float2half(Bval(n, k));
  }
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc<64>(slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = warp * 32 + lane;
  uint32_t phase = 0;
  constexpr uint32_t idesc = make_idesc_f16_f32(128, N);
  for (int s = 0; s < 24; ++s) {
    for (int b = 0; b < 8; ++b) {
      if (threadIdx.x == 0) {
        const uint64_t a_desc = make_kmajor_desc<SWZ>(sA + s * SWZ) | (uint64_t(b) << 49);
        const uint64_t b_desc = make_kmajor_desc<SWZ>(sB);
        for (int k = 0; k < K / 16; ++k) umma_f16_ss(tmem, a_desc + uint64_t(2 * k), b_desc + uint64_t(2 * k), idesc, k != 0);
        umma_commit(bar);
      }
      mbar_wait(bar, phase);
      phase ^= 1u;
      tcgen05_fence_after();
      int bad = 0;
      for (int c = 0; c < N; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b<32>(tmem + (uint32_t(warp * 32) << 16) + c, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) {
          float ref = 0.f;
          for (int k = 0; k < K; ++k) ref += Aval(row + s, k) * Bval(c + j, k);
          if (__uint_as_float(v[j]) != ref) ++bad;
          if (s == 1 && b == 1 && dump) dump[row * N + c + j] = __uint_as_float(v[j]);
        }
      }
      if (bad) atomicAdd(&mism[s * 8 + b], bad);
      tcgen05_fence_before();
      __syncthreads();
      tcgen05_fence_after();
    }
  }
  if (threadIdx.x < 32) tmem_dealloc<64>(tmem);
}

template <int SWZ>
static void run() {
  int* d;
  cudaMalloc(&d, 24 * 8 * sizeof(int));
  cudaMemset(d, 0, 24 * 8 * sizeof(int));
  const size_t smem = 1024 + (128 + 32) * SWZ + 64 * SWZ + 64;
  cudaFuncSetAttribute(probe<SWZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<SWZ><<<1, 128, smem>>>(d, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  printf("swizzle %dB: %s\n", SWZ, cudaGetErrorString(e));
  int h[24 * 8];
  cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  printf("mismatches (of 8192) by row shift s (rows) x base_offset b (cols)\n");
  for (int s = 0; s < 24; ++s) {
    printf("s=%2d:", s);
    for (int b = 0; b < 8; ++b) printf(" %5d", h[s * 8 + b]);
    printf("\n");
  }
  cudaFree(d);
}

int main() {
  run<128>();
  run<64>();
  return 0;
}
