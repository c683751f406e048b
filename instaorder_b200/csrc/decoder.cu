// Element-wise pieces of the MiDaS decoder of InstaDepthNet (reference midas/blocks.py:124-195, midas_net.py:126-140,
// 189-198): the 3x3 convolutions run through io_conv_bn_act; what is left between them is an addition with an optional
// ReLU (ResidualConvUnit's in-place ReLU makes every consumer read relu(x), so the sums are stored ReLU'd) and the x2
// bilinear up-sampling (align_corners = True in the fusion blocks, False in output_conv's Interpolate).  Both are
// per-IMAGE tensors (the disparity does not depend on the masks): HBM-trivial next to the per-pair trunks.
#include "common.cuh"

namespace io {

__global__ void __launch_bounds__(256) add_relu_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b,
                                                       uint4* __restrict__ out, size_t n8, int relu) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint4 x = __ldg(a + i), y = __ldg(b + i);
    float f[8] = {bf16_lo(x.x) + bf16_lo(y.x), bf16_hi(x.x) + bf16_hi(y.x), bf16_lo(x.y) + bf16_lo(y.y),
                  bf16_hi(x.y) + bf16_hi(y.y), bf16_lo(x.z) + bf16_lo(y.z), bf16_hi(x.z) + bf16_hi(y.z),
                  bf16_lo(x.w) + bf16_lo(y.w), bf16_hi(x.w) + bf16_hi(y.w)};
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.0f);
    }
    uint4 o;
    o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]); o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
    out[i] = o;
  }
}

// NHWC bf16 [b, h, w, c] -> [b, 2h, 2w, c]; one thread = 8 channels of one output pixel.  Source coordinates as
// torch.nn.functional.interpolate(mode="bilinear") computes them in fp32:
//   align_corners: src = dst * (in - 1) / (out - 1);  else: src = max((dst + 0.5) * in / out - 0.5, 0)
__global__ void __launch_bounds__(256) upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int h, int w,
                                                         int c8, int align) {
  const int ho = 2 * h, wo = 2 * w;
  const int oy = blockIdx.x % ho, n = blockIdx.x / ho;
  const float sy = align ? (ho > 1 ? static_cast<float>(h - 1) / static_cast<float>(ho - 1) : 0.0f) : 0.5f;
  const float sx = align ? (wo > 1 ? static_cast<float>(w - 1) / static_cast<float>(wo - 1) : 0.0f) : 0.5f;
  float fy = align ? sy * oy : fmaxf(sy * (oy + 0.5f) - 0.5f, 0.0f);
  const int y0 = min(static_cast<int>(fy), h - 1), y1 = min(y0 + 1, h - 1);
  const float ly = fy - y0;
  const uint4* __restrict__ r0 = x + (static_cast<size_t>(n) * h + y0) * w * c8;
  const uint4* __restrict__ r1 = x + (static_cast<size_t>(n) * h + y1) * w * c8;
  uint4* __restrict__ out = y + (static_cast<size_t>(n) * ho + oy) * wo * c8;
  for (int i = threadIdx.x; i < wo * c8; i += blockDim.x) {
    const int ox = i / c8, cg = i - ox * c8;
    float fx = align ? sx * ox : fmaxf(sx * (ox + 0.5f) - 0.5f, 0.0f);
    const int x0 = min(static_cast<int>(fx), w - 1), x1 = min(x0 + 1, w - 1);
    const float lx = fx - x0;
    const uint4 a = __ldg(r0 + x0 * c8 + cg), b = __ldg(r0 + x1 * c8 + cg);
    const uint4 c = __ldg(r1 + x0 * c8 + cg), d = __ldg(r1 + x1 * c8 + cg);
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    auto mix = [&](uint32_t pa, uint32_t pb, uint32_t pc, uint32_t pd) {
      const float lo = w00 * bf16_lo(pa) + w01 * bf16_lo(pb) + w10 * bf16_lo(pc) + w11 * bf16_lo(pd);
      const float hi = w00 * bf16_hi(pa) + w01 * bf16_hi(pb) + w10 * bf16_hi(pc) + w11 * bf16_hi(pd);
      return pack_bf16(lo, hi);
    };
    uint4 o;
    o.x = mix(a.x, b.x, c.x, d.x); o.y = mix(a.y, b.y, c.y, d.y);
    o.z = mix(a.z, b.z, c.z, d.z); o.w = mix(a.w, b.w, c.w, d.w);
    out[i] = o;
  }
}

}  // namespace io

using namespace io;

extern "C" int io_add_relu(const void* a, const void* b, void* out, int64_t n, int relu, void* stream) {
  IO_REQUIRE(a && b && out && n >= 0 && n % 8 == 0, "io_add_relu: bad arguments (n %% 8 == 0)");
  if (n == 0) return IO_OK;
  const size_t n8 = static_cast<size_t>(n) / 8;
  const int grid = static_cast<int>(std::min<size_t>((n8 + 255) / 256, static_cast<size_t>(num_sms()) * 8));
  add_relu_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b),
                                                      reinterpret_cast<uint4*>(out), n8, relu);
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}

extern "C" int io_upsample2x_bilinear(const void* x, int b, int h, int w, int c, int align_corners, void* y, void* stream) {
  IO_REQUIRE(x && y && b >= 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0, "io_upsample2x_bilinear: bad arguments");
  if (b == 0) return IO_OK;
  upsample2x_kernel<<<static_cast<unsigned>(b) * 2 * h, 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), h, w, c / 8, align_corners ? 1 : 0);
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}
