// Helper kernels of the torchvision ResNet-50 plan (the reference's ImageNet baseline encoder:
// models.resnet50 -> children()[:-2], primitive_probing/generate_data/thor_image_features.py:46-49,101-105):
//   im2col7x7s2_kernel    stem conv 7x7 / stride 2 / pad 3 as a GEMM: frames -> fp16 K-major rows [B*Ro*Ro, 160]
//                         (147 taps in (kh, kw, c) order + 13 zero columns), consumed by conv_gemm (N = 64, bias, ReLU)
//   maxpool3x3s2_kernel   nn.MaxPool2d(3, stride 2, padding 1), NHWC fp16
//   subsample2_kernel     x[:, ::2, ::2, :] -- the input of a stride-2 1x1 downsample conv, NHWC fp16
// All three are HBM-bound element movers: one 16-B vector per thread access, grid-stride loops.
#pragma once
#include "ptx.cuh"
#include "aux_kernels.cuh"

namespace embclip {

constexpr int kStem7K = 160;     // 7 * 7 * 3 = 147 taps, padded to a multiple of the 32-element k-block

// One thread = one output pixel x one kernel row (kh): 21 contiguous input values (7 pixels x 3 channels of NHWC) ->
// 21 fp16 values at columns [kh*21, kh*21 + 21) of the pixel's row.  Rows are assembled in shared memory (8 pixels per
// block-iteration x 160 columns) and written out as whole 16-B vectors, so global stores are coalesced 320-B rows.
template <typename TIn>
__global__ void __launch_bounds__(256)
im2col7x7s2_kernel(const TIn* __restrict__ x, __half* __restrict__ y, int B, int R, const StemNorm norm) {
  constexpr int PX = 32;                                   // output pixels per block iteration
  __shared__ __align__(16) __half tile[PX][kStem7K];
  const int Ro = R / 2;
  const long long total = (long long)B * Ro * Ro;
  griddep_wait();
  for (long long base = (long long)blockIdx.x * PX; base < total; base += (long long)gridDim.x * PX) {
    // zero the 13 padding columns once per tile (cheap, keeps the loop body uniform)
    for (int i = threadIdx.x; i < PX * (kStem7K - 147); i += blockDim.x) tile[i / 13][147 + i % 13] = __float2half_rn(0.f);
    for (int i = threadIdx.x; i < PX * 7; i += blockDim.x) {
      const int px = i / 7, kh = i - px * 7;
      const long long o = base + px;
      __half* dst = &tile[px][kh * 21];
      if (o >= total) {
#pragma unroll
        for (int j = 0; j < 21; ++j) dst[j] = __float2half_rn(0.f);
        continue;
      }
      const int ow = int(o % Ro), oh = int((o / Ro) % Ro), b = int(o / ((long long)Ro * Ro));
      const int ih = 2 * oh - 3 + kh;
      const int iw0 = 2 * ow - 3;
      if (ih < 0 || ih >= R) {
#pragma unroll
        for (int j = 0; j < 21; ++j) dst[j] = __float2half_rn(0.f);
        continue;
      }
      const TIn* src = x + ((size_t)b * R + ih) * R * 3;
#pragma unroll
      for (int kw = 0; kw < 7; ++kw) {
        const int iw = iw0 + kw;
        const bool ok = iw >= 0 && iw < R;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float v = 0.f;
          if (ok) v = float(src[(size_t)iw * 3 + c]) * norm.scale[c] + norm.offset[c];
          dst[kw * 3 + c] = __float2half_rn(v);
        }
      }
    }
    __syncthreads();
    const long long nvalid = total - base < PX ? total - base : PX;
    uint4* out = reinterpret_cast<uint4*>(y + (size_t)base * kStem7K);
    const uint4* in = reinterpret_cast<const uint4*>(&tile[0][0]);
    for (int i = threadIdx.x; i < int(nvalid) * (kStem7K / 8); i += blockDim.x) out[i] = in[i];
    __syncthreads();
  }
}

__device__ __forceinline__ void max8(__half2 (&m)[4], const uint4& v) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) m[i] = __hmax2(m[i], h[i]);
}

// y[b, oh, ow, :] = max over the 3x3 window at (2oh-1.., 2ow-1..) clipped to the image (padding never wins: -inf)
__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const __half* __restrict__ x, __half* __restrict__ y, int B, int H, int W, int C) {
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
    __half2 m[4];
    const __half2 ninf = __float2half2_rn(-65504.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) m[k] = ninf;
#pragma unroll
    for (int dh = -1; dh <= 1; ++dh) {
      const int ih = 2 * oh + dh;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int dw = -1; dw <= 1; ++dw) {
        const int iw = 2 * ow + dw;
        if (iw < 0 || iw >= W) continue;
        max8(m, __ldg(reinterpret_cast<const uint4*>(x + (((size_t)b * H + ih) * W + iw) * C) + c8));
      }
    }
    reinterpret_cast<uint4*>(y)[i] = *reinterpret_cast<const uint4*>(m);
  }
}

__global__ void __launch_bounds__(256)
subsample2_kernel(const __half* __restrict__ x, __half* __restrict__ y, int B, int H, int W, int C) {
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
    reinterpret_cast<uint4*>(y)[i] = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)b * H + 2 * oh) * W + 2 * ow) * C) + c8);
  }
}

}  // namespace embclip
