// Experiment: tcgen05.mma (kind::f16, cta_group::1, M=128, SS operands) issue rate vs N, with operands cycling
// through several smem stages like a real pipeline (no loads: smem content is whatever it is; finite is not needed
// for timing).  Prints cycles per MMA and the implied MAC/clk/SM, for 1 CTA and for 148 CTAs.
#include <cstdio>
#include <cuda_fp16.h>
#include "../embodied-clip_b200/csrc/ptx.cuh"
using namespace embclip;

template <int N, int MS, int SWZ>
__global__ void __launch_bounds__(128, 1) rate(long long* out, int iters, int shift) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  constexpr int STAGES = 4;
  constexpr int ABYTES = 128 * SWZ + 4096, BBYTES = N * SWZ;
  const uint32_t sA = base, sB = base + STAGES * MS * ABYTES;
  const uint32_t bar = sB + STAGES * BBYTES, slot = bar + 8;
  for (uint32_t i = threadIdx.x; i < (STAGES * (MS * ABYTES + BBYTES)) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc<512>(slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_f16_f32(128, N);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int st = it % STAGES;
      const uint64_t b_desc = make_kmajor_desc<SWZ>(sB + st * BBYTES);
#pragma unroll
      for (int k = 0; k < SWZ / 32; ++k) {
#pragma unroll
        for (int s = 0; s < MS; ++s) {
          const uint64_t a_desc = make_kmajor_desc<SWZ>(sA + (st * MS + s) * ABYTES + shift * SWZ);
          umma_f16_ss(tmem + s * N, a_desc + uint64_t(2 * k), b_desc + uint64_t(2 * k), idesc, 1);
        }
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

template <int N, int MS, int SWZ>
static void run(int grid, int shift) {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  const size_t smem = 1024 + 4 * (MS * (128 * SWZ + 4096) + N * SWZ) + 64;
  cudaFuncSetAttribute(rate<N, MS, SWZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 4000;
  rate<N, MS, SWZ><<<grid, 128, smem>>>(d, iters, shift);
  rate<N, MS, SWZ><<<grid, 128, smem>>>(d, iters, shift);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  const double per = double(mx) / (iters * (SWZ / 32.0) * MS);
  printf("SWZ=%d shift=%d N=%3d MS=%d grid=%3d: %s  %.1f cyc/MMA (floor %.0f)  %.0f MAC/clk/SM  smem operand read %.0f B/clk\n", SWZ, shift, N, MS, grid,
         cudaGetErrorString(e), per, 128.0 * N / 256.0, 128.0 * N * 16 / per, (128 * 32 + N * 32) / per);
  cudaFree(d);
}

int main() {
  for (int shift : {0, 1, 3, 4, 8}) {
    run<32, 1, 128>(148, shift); run<64, 1, 128>(148, shift); run<128, 1, 128>(148, shift); run<256, 1, 128>(148, shift);
    run<32, 1, 64>(148, shift); run<64, 1, 64>(148, shift);
  }
  return 0;
}
