// Small HBM-bound kernels around the tensor-core convolutions: stem max-pool, the fused pool + FC tail and the
// decision / order-matrix scatter.
#include "tail.cuh"

namespace io {

// ---- nn.MaxPool2d(3, 2, 1) on NHWC bf16 (models/backbone/resnet_cls.py:144,207) -------------------------------
// one thread = 8 channels of one output pixel; padding never wins because the input is post-ReLU (>= 0) and the
// window always contains at least one real pixel.  One CTA per output row (image, oy): the three input row pointers
// are block-uniform and (ox, channel group) come from the thread index by shifts -- the first version decoded a flat
// index with three 64-bit divisions per output and was issue-bound at 60 % of the HBM roofline.
__global__ void __launch_bounds__(256) maxpool_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int h, int w,
                                                      int c8_shift) {
  const int ho = h / 2, wo = w / 2;
  const int n = blockIdx.x / ho, oy = blockIdx.x - n * ho;
  const int c8 = 1 << c8_shift;
  const uint4* __restrict__ xin = x + static_cast<size_t>(n) * h * w * c8;
  const int iy0 = 2 * oy - 1;
  const bool top_ok = iy0 >= 0;               // rows 2oy and 2oy + 1 always exist (h even)
  const uint4* __restrict__ r0 = xin + static_cast<size_t>(top_ok ? iy0 : 2 * oy) * w * c8;   // duplicate = harmless for max
  const uint4* __restrict__ r1 = xin + static_cast<size_t>(2 * oy) * w * c8;
  const uint4* __restrict__ r2 = xin + static_cast<size_t>(2 * oy + 1) * w * c8;
  uint4* __restrict__ yout = y + static_cast<size_t>(blockIdx.x) * wo * c8;
  for (int i = threadIdx.x; i < wo * c8; i += blockDim.x) {
    const int ox = i >> c8_shift, cg = i & (c8 - 1);
    const int xl = (ox > 0) ? 2 * ox - 1 : 0;  // left tap clamped to column 0 (duplicate of the centre tap)
    const int o0 = xl * c8 + cg, o1 = (2 * ox) * c8 + cg, o2 = (2 * ox + 1) * c8 + cg;
    uint4 v[9];
    v[0] = __ldg(r0 + o0); v[1] = __ldg(r0 + o1); v[2] = __ldg(r0 + o2);
    v[3] = __ldg(r1 + o0); v[4] = __ldg(r1 + o1); v[5] = __ldg(r1 + o2);
    v[6] = __ldg(r2 + o0); v[7] = __ldg(r2 + o1); v[8] = __ldg(r2 + o2);
    __nv_bfloat162 m[4];
    const __nv_bfloat162* p0 = reinterpret_cast<const __nv_bfloat162*>(&v[0]);
    m[0] = p0[0]; m[1] = p0[1]; m[2] = p0[2]; m[3] = p0[3];
#pragma unroll
    for (int k = 1; k < 9; ++k) {
      const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&v[k]);
      m[0] = __hmax2(m[0], pv[0]); m[1] = __hmax2(m[1], pv[1]);
      m[2] = __hmax2(m[2], pv[2]); m[3] = __hmax2(m[3], pv[3]);
    }
    yout[i] = *reinterpret_cast<uint4*>(m);
  }
}

int maxpool_launch(const void* x, void* y, int b, int h, int w, int c, cudaStream_t stream) {
  IO_REQUIRE(c % 8 == 0 && h % 2 == 0 && w % 2 == 0, "maxpool: bad shape");
  const int c8 = c / 8;
  int shift = 0;
  while ((1 << shift) < c8) ++shift;
  IO_REQUIRE((1 << shift) == c8, "maxpool: channel count %d (8 x a power of two expected)", c);
  const long long blocks = static_cast<long long>(b) * (h / 2);
  if (blocks == 0) return IO_OK;
  maxpool_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(reinterpret_cast<const uint4*>(x),
                                                                     reinterpret_cast<uint4*>(y), h, w, shift);
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}

// ---- AdaptiveAvgPool2d(1) + flatten + Linear head(s) (resnet_cls.py:214-221) ----------------------------------
// feat: [2P, HW, 2048] bf16, image 2p = direction (A,B) of pair p, image 2p + 1 = direction (B,A).
// logits: [P][2][K] fp32.  One CTA per image, 256 threads x 8 channels.
constexpr int TAIL_C = 2048;
constexpr int TAIL_MAXK = 8;

__global__ void __launch_bounds__(256) tail_kernel(const uint4* __restrict__ feat, int hw, int pairs,
                                                   const float* __restrict__ fcw, const float* __restrict__ fcb,
                                                   int k_total, float* __restrict__ logits) {
  const int img = blockIdx.x;
  const int pair = img >> 1, dir = img & 1;
  const int t = threadIdx.x;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const uint4* p = feat + static_cast<size_t>(img) * hw * (TAIL_C / 8) + t;
  for (int i = 0; i < hw; ++i) {
    const uint4 v = __ldg(p + static_cast<size_t>(i) * (TAIL_C / 8));
    s[0] += bf16_lo(v.x); s[1] += bf16_hi(v.x); s[2] += bf16_lo(v.y); s[3] += bf16_hi(v.y);
    s[4] += bf16_lo(v.z); s[5] += bf16_hi(v.z); s[6] += bf16_lo(v.w); s[7] += bf16_hi(v.w);
  }
  const float inv = static_cast<float>(hw);
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = s[j] / inv;
  __shared__ float red[TAIL_MAXK][8];
  for (int k = 0; k < k_total; ++k) {
    const float4* w4 = reinterpret_cast<const float4*>(fcw + static_cast<size_t>(k) * TAIL_C + t * 8);
    const float4 a = __ldg(w4), b = __ldg(w4 + 1);
    float d = s[0] * a.x + s[1] * a.y + s[2] * a.z + s[3] * a.w + s[4] * b.x + s[5] * b.y + s[6] * b.z + s[7] * b.w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if ((t & 31) == 0) red[k][t >> 5] = d;
  }
  __syncthreads();
  if (t < k_total) {
    float d = 0.f;
#pragma unroll
    for (int wp = 0; wp < 8; ++wp) d += red[t][wp];
    logits[(static_cast<size_t>(pair) * 2 + dir) * k_total + t] = d + fcb[t];
  }
}

int tail_launch(const void* feat, int hw, int pairs, const float* fcw, const float* fcb, int k_total, float* logits,
                cudaStream_t stream) {
  IO_REQUIRE(k_total >= 1 && k_total <= TAIL_MAXK, "tail: %d logits not supported (max %d)", k_total, TAIL_MAXK);
  if (pairs == 0) return IO_OK;
  tail_kernel<<<2 * pairs, 256, 0, stream>>>(reinterpret_cast<const uint4*>(feat), hw, pairs, fcw, fcb, k_total,
                                             logits);
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}

// ---- decisions + scatter (inference.py:44-76, 172-214, 416-434) ---------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ void softmax_k(const float* z, int k, float* out) {
  float m = z[0];
  for (int i = 1; i < k; ++i) m = fmaxf(m, z[i]);
  float s = 0.f;
  for (int i = 0; i < k; ++i) { out[i] = expf(z[i] - m); s += out[i]; }
  for (int i = 0; i < k; ++i) out[i] = out[i] / s;
}

__global__ void __launch_bounds__(128) decide_kernel(const float* __restrict__ logits, int pairs, int k_total,
                                                     int head_kind, int head_off, int head_k,
                                                     const int32_t* __restrict__ pair_ij,
                                                     const int64_t* __restrict__ mat_off,
                                                     const int32_t* __restrict__ mat_n, int64_t* __restrict__ mat,
                                                     float* __restrict__ margin) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pairs) return;
  const float* l1 = logits + (static_cast<size_t>(p) * 2 + 0) * k_total + head_off;
  const float* l2 = logits + (static_cast<size_t>(p) * 2 + 1) * k_total + head_off;
  const int i = pair_ij[2 * p], j = pair_ij[2 * p + 1];
  const int n = mat_n[p];
  int64_t* m = mat + mat_off[p];
  float mg;
  if (head_kind == IO_HEAD_OCC) {
    const float p12 = (sigmoidf_(l1[1]) + sigmoidf_(l2[0])) / 2.0f;
    const float p21 = (sigmoidf_(l1[0]) + sigmoidf_(l2[1])) / 2.0f;
    if (p12 > 0.5f) m[static_cast<size_t>(i) * n + j] = 1;
    if (p21 > 0.5f) m[static_cast<size_t>(j) * n + i] = 1;
    mg = fminf(fabsf(p12 - 0.5f), fabsf(p21 - 0.5f));
  } else {
    float o1[4], o2[4], pr[4];
    softmax_k(l1, head_k, o1);
    softmax_k(l2, head_k, o2);
    int np;
    if (head_kind == IO_HEAD_DEPTH) {
      pr[0] = (o1[0] + o2[1]) / 2.0f;  // i closer than j
      pr[1] = (o1[1] + o2[0]) / 2.0f;  // i farther
      pr[2] = (o1[2] + o2[2]) / 2.0f;  // equal
      np = 3;
    } else {
      pr[0] = (o1[1] + o2[0]) / 2.0f;  // 1 over 2
      pr[1] = (o1[0] + o2[1]) / 2.0f;  // 2 over 1
      pr[2] = (o1[2] + o2[2]) / 2.0f;  // none
      pr[3] = head_k == 4 ? (o1[3] + o2[3]) / 2.0f : 0.0f;  // both (OrderNet_ext)
      np = 4;
    }
    int a = 0;
    for (int q = 1; q < np; ++q)
      if (pr[q] > pr[a]) a = q;  // np.argmax: first maximum
    float second = -1.0f;
    for (int q = 0; q < np; ++q)
      if (q != a) second = fmaxf(second, pr[q]);
    mg = pr[a] - second;
    if (head_kind == IO_HEAD_DEPTH) {
      const int64_t vij = a == 0 ? 1 : (a == 1 ? 0 : 2);
      const int64_t vji = a == 0 ? 0 : (a == 1 ? 1 : 2);
      m[static_cast<size_t>(i) * n + j] = vij;
      m[static_cast<size_t>(j) * n + i] = vji;
    } else {
      if (a == 0 || a == 3) m[static_cast<size_t>(i) * n + j] = 1;
      if (a == 1 || a == 3) m[static_cast<size_t>(j) * n + i] = 1;
    }
  }
  if (margin) margin[p] = mg;
}

}  // namespace io

extern "C" int io_order_decide(const float* logits, int p, int k_total, int head_kind, int head_off, int head_k,
                               const int32_t* pair_ij, const int64_t* mat_off, const int32_t* mat_n, int64_t* mat,
                               float* margin, void* stream) {
  IO_REQUIRE(logits && pair_ij && mat_off && mat_n && mat && p >= 0, "io_order_decide: bad arguments");
  IO_REQUIRE(head_off >= 0 && head_off + head_k <= k_total, "io_order_decide: head columns out of range");
  IO_REQUIRE((head_kind == IO_HEAD_OCC && head_k == 2) || (head_kind == IO_HEAD_DEPTH && head_k == 3) ||
                 (head_kind == IO_HEAD_ORDERNET && (head_k == 3 || head_k == 4)),
             "io_order_decide: head kind %d with %d logits", head_kind, head_k);
  if (p == 0) return IO_OK;
  io::decide_kernel<<<(p + 127) / 128, 128, 0, io::as_stream(stream)>>>(logits, p, k_total, head_kind, head_off,
                                                                       head_k, pair_ij, mat_off, mat_n, mat, margin);
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}
