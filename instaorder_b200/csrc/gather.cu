// Pair construction: crop-window geometry (host, float64), bordering test and the fused gather kernels.
//
// Replaces, per pair, utils.crop_padding x3 + cv2.resize(INTER_CUBIC) + cv2.resize(INTER_NEAREST) x2 +
// transform_rgb + 2 H2D copies + torch.cat (reference inference.py:360-375, :141-145; utils/data_utils.py:28-34,
// :104-124) by ONE kernel that reads the u8 image / masks in HBM and writes the bf16 NHWC pair tensor once.
//
// Arithmetic follows OpenCV 4.13's generic (non-IPP) resize exactly (see oracle/oracle.py): 11-bit fixed-point
// cubic taps (A = -0.75) built with un-fused fp32 operations, integer horizontal pass, fp32 vertical pass in the
// order 3,2,1,0 without fma, round-half-even, saturate; nearest = floor(dst * (1 / (D / S))) in double.
#include "common.cuh"

#include <math.h>
#include <vector>

namespace io {

constexpr int GATHER_ROWS = 8;      // output rows per CTA (32-row bands: 10 % fewer instructions, same time -- fewer CTAs)
constexpr int GATHER_THREADS = 128;
constexpr int MAX_D = 512;

struct NormLut {
  uint16_t v[3][256];  // bf16 bits of (u8 / 255 - mean) / std
};

static inline uint16_t f32_to_bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return static_cast<uint16_t>((u >> 16) | 0x40);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}

static void fill_lut_f32(const float* mean, const float* stdv, float* out) {
  for (int c = 0; c < 3; ++c)
    for (int v = 0; v < 256; ++v) {
      volatile float x = static_cast<float>(v) / 255.0f;  // image / 255.   (fp32, as torch)
      volatile float y = x - mean[c];                      // Normalize: sub_(mean)
      out[c * 256 + v] = y / stdv[c];                      //            div_(std)
    }
}

// ---- per-axis resampling tables ---------------------------------------------------------------------------
struct AxisTab {
  int near;       // nearest source index
  int first;      // first cubic tap (may be < 0 / > S-4: taps are clamped to the crop)
  short coef[4];  // 11-bit fixed-point cubic weights
};

__device__ __forceinline__ AxisTab axis_entry(int dst, int src_len, int dst_len) {
  AxisTab t;
  const double inv = __ddiv_rn(static_cast<double>(dst_len), static_cast<double>(src_len));
  const double scale = __ddiv_rn(1.0, inv);
  // nearest: cvFloor(x * ifx), clamped
  int n = static_cast<int>(floor(__dmul_rn(static_cast<double>(dst), scale)));
  t.near = min(n, src_len - 1);
  // cubic: fx = (float)((dx + 0.5) * scale - 0.5)
  float fx = static_cast<float>(__dsub_rn(__dmul_rn(static_cast<double>(dst) + 0.5, scale), 0.5));
  const float sxf = floorf(fx);
  const int sx = static_cast<int>(sxf);
  const float x = __fsub_rn(fx, sxf);
  const float A = -0.75f;
  const float xp1 = __fadd_rn(x, 1.0f);
  const float omx = __fsub_rn(1.0f, x);
  // coeffs[0] = ((A*(x + 1) - 5*A)*(x + 1) + 8*A)*(x + 1) - 4*A
  float c0 = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, xp1), __fmul_rn(5.0f, A)), xp1),
                                            __fmul_rn(8.0f, A)), xp1), __fmul_rn(4.0f, A));
  // coeffs[1] = ((A + 2)*x - (A + 3))*x*x + 1
  float c1 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(A, 2.0f), x), __fadd_rn(A, 3.0f)), x), x),
                       1.0f);
  // coeffs[2] = ((A + 2)*(1 - x) - (A + 3))*(1 - x)*(1 - x) + 1
  float c2 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(A, 2.0f), omx), __fadd_rn(A, 3.0f)), omx),
                                 omx), 1.0f);
  float c3 = __fsub_rn(__fsub_rn(__fsub_rn(1.0f, c0), c1), c2);
  const float c[4] = {c0, c1, c2, c3};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int iv = __float2int_rn(__fmul_rn(c[k], 2048.0f));
    t.coef[k] = static_cast<short>(max(-32768, min(32767, iv)));
  }
  t.first = sx - 1;
  return t;
}

__device__ __forceinline__ void store_pixel(__nv_bfloat16* dst, float ma, float mb, uint16_t r, uint16_t g,
                                            uint16_t b) {
  uint4 o;
  o.x = pack_bf16(ma, mb);
  o.y = static_cast<uint32_t>(r) | (static_cast<uint32_t>(g) << 16);
  o.z = static_cast<uint32_t>(b);
  o.w = 0u;
  *reinterpret_cast<uint4*>(dst) = o;
}

// zero the 3-pixel border (and the pitch padding) that belongs to this CTA's rows
__device__ __forceinline__ void zero_borders(__nv_bfloat16* out_pair, int d, int pitch, int band, int n_bands,
                                             int row0, int row1, int dh = 0) {
  const uint4 z = make_uint4(0, 0, 0, 0);
  const int hp = (dh > 0 ? dh : d) + 6;      // d = interior width; dh = interior height when it differs (`orig` mode)
  // left / right borders of the band's interior rows
  const int side = 3 + (pitch - d - 3);
  for (int i = threadIdx.x; i < (row1 - row0) * side; i += blockDim.x) {
    const int rr = row0 + i / side, k = i % side;
    const int col = k < 3 ? k : d + 3 + (k - 3);
    *reinterpret_cast<uint4*>(out_pair + (static_cast<size_t>(rr + 3) * pitch + col) * 8) = z;
  }
  if (band == 0)
    for (int i = threadIdx.x; i < 3 * pitch; i += blockDim.x)
      *reinterpret_cast<uint4*>(out_pair + static_cast<size_t>(i) * 8) = z;
  if (band == n_bands - 1)
    for (int i = threadIdx.x; i < 3 * pitch; i += blockDim.x)
      *reinterpret_cast<uint4*>(out_pair + (static_cast<size_t>(hp - 3) * pitch + i) * 8) = z;
}

// Separable form of the same arithmetic (OpenCV's generic resize is separable too): the integer horizontal pass is
// done ONCE per source row of the band -- its three channel sums per output column go into a ring of 8 rows in shared
// memory -- and the fp32 vertical pass combines four ring rows per output row.  A CTA owns (pair, 8 output rows) and
// walks them top to bottom, so every source row is resampled once per band instead of once per output row and tap
// (ncu: the direct form issued 353 warp instructions per pixel and was issue / L1 bound at 1.1 TB/s).
// Per-column tap offsets / validity / weights depend only on the column and are hoisted into registers (NPX columns
// per thread).  Dynamic shared memory: AxisTab[d] + int32 ring[8][3][d].
template <int NPX>
__global__ void __launch_bounds__(GATHER_THREADS) gather_patch_kernel(const uint8_t* __restrict__ images,
                                                                      const uint8_t* __restrict__ masks,
                                                                      const io_pair_desc* __restrict__ descs, int d,
                                                                      int pitch, NormLut lut,
                                                                      __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t gsm[];
  AxisTab* tab = reinterpret_cast<AxisTab*>(gsm);
  int* ring = reinterpret_cast<int*>(gsm + static_cast<size_t>(d) * sizeof(AxisTab));   // [8][3][d]
  __shared__ uint16_t slut[3][256];
  const io_pair_desc ds = descs[blockIdx.y];
  const int S = ds.s;
  for (int i = threadIdx.x; i < d; i += blockDim.x) tab[i] = axis_entry(i, S, d);
  for (int i = threadIdx.x; i < 768; i += blockDim.x) (&slut[0][0])[i] = (&lut.v[0][0])[i];
  __syncthreads();

  const uint8_t* __restrict__ img = images + ds.image_off;
  const uint8_t* __restrict__ ma = masks + ds.mask_a_off;
  const uint8_t* __restrict__ mb = masks + ds.mask_b_off;
  const int H = ds.h, W = ds.w, X = ds.x, Y = ds.y;
  const bool flip = (ds.rgb_slot & 1) != 0;   // training augmentation: horizontal flip AFTER the resize
  __nv_bfloat16* out_pair = out + static_cast<size_t>(blockIdx.y) * (d + 6) * pitch * 8;
  const int row0 = blockIdx.x * GATHER_ROWS;
  const int row1 = min(d, row0 + GATHER_ROWS);
  zero_borders(out_pair, d, pitch, blockIdx.x, gridDim.x, row0, row1);
  const float kscale = 1.0f / (2048.0f * 2048.0f);

  // ---- per-column state (registers): byte offsets of the four taps inside an image row (-1 = outside the image:
  // crop_padding's zero), their 11-bit weights, and the nearest-neighbour mask column
  int off[NPX][4];
  int cf[NPX][4];
  int mcol[NPX];
#pragma unroll
  for (int m = 0; m < NPX; ++m) {
    const int dx = threadIdx.x + m * GATHER_THREADS;
    if (dx < d) {
      const AxisTab tx = tab[flip ? d - 1 - dx : dx];
      const int mx = X + tx.near;
      mcol[m] = (mx >= 0 && mx < W) ? mx : -1;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int xx = min(max(tx.first + j, 0), S - 1);
        const int ix = X + xx;
        off[m][j] = (ix >= 0 && ix < W) ? ix * 3 : -1;
        cf[m][j] = tx.coef[j];
      }
    } else {
      mcol[m] = -1;
#pragma unroll
      for (int j = 0; j < 4; ++j) { off[m][j] = -1; cf[m][j] = 0; }
    }
  }

  int hi = -1;   // crop rows < hi have been resampled into the ring (or skipped for good)
  for (int dy = row0; dy < row1; ++dy) {
    const AxisTab ty = tab[dy];
    int cr[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) cr[k] = min(max(ty.first + k, 0), S - 1);   // crop rows of the four taps (non-decreasing)
    const int start = max(hi, cr[0]);
    if (start <= cr[3]) {              // uniform over the CTA
      __syncthreads();                 // the previous row's vertical pass has finished reading the ring
      for (int c = start; c <= cr[3]; ++c) {
        const int iy = Y + c;
        const bool oky = iy >= 0 && iy < H;
        const uint8_t* rowp = img + static_cast<size_t>(oky ? iy : 0) * W * 3;
        int* rr = ring + (c & 7) * 3 * d;
#pragma unroll
        for (int m = 0; m < NPX; ++m) {
          const int dx = threadIdx.x + m * GATHER_THREADS;
          if (dx < d) {
            int h0 = 0, h1 = 0, h2 = 0;
            if (oky) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (off[m][j] >= 0) {
                  const uint8_t* px = rowp + off[m][j];
                  h0 += static_cast<int>(px[0]) * cf[m][j];
                  h1 += static_cast<int>(px[1]) * cf[m][j];
                  h2 += static_cast<int>(px[2]) * cf[m][j];
                }
              }
            }
            rr[dx] = h0;
            rr[d + dx] = h1;
            rr[2 * d + dx] = h2;
          }
        }
      }
      hi = cr[3] + 1;
      __syncthreads();
    }
    // ---- vertical pass (fp32, taps 3,2,1,0, no fma) + modal masks (nearest) + normalisation LUT
    const int my = Y + ty.near;
    const bool my_ok = my >= 0 && my < H;
    float by[4];
    const int* rk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      by[k] = __fmul_rn(static_cast<float>(ty.coef[k]), kscale);
      rk[k] = ring + (cr[k] & 7) * 3 * d;
    }
#pragma unroll
    for (int m = 0; m < NPX; ++m) {
      const int dx = threadIdx.x + m * GATHER_THREADS;
      if (dx < d) {
        float va = 0.0f, vb = 0.0f;
        if (my_ok && mcol[m] >= 0) {
          const size_t o = static_cast<size_t>(my) * W + mcol[m];
          va = static_cast<float>(ma[o]);
          vb = static_cast<float>(mb[o]);
        }
        uint16_t rgb[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float v = __fmul_rn(static_cast<float>(rk[3][c * d + dx]), by[3]);
          v = __fadd_rn(__fmul_rn(static_cast<float>(rk[2][c * d + dx]), by[2]), v);
          v = __fadd_rn(__fmul_rn(static_cast<float>(rk[1][c * d + dx]), by[1]), v);
          v = __fadd_rn(__fmul_rn(static_cast<float>(rk[0][c * d + dx]), by[0]), v);
          const int u = min(max(__float2int_rn(v), 0), 255);
          rgb[c] = slut[c][u];
        }
        store_pixel(out_pair + (static_cast<size_t>(dy + 3) * pitch + dx + 3) * 8, va, vb, rgb[0], rgb[1], rgb[2]);
      }
    }
  }
}

// ---- resize mode --------------------------------------------------------------------------------------------
struct AxisTabF {
  int first;
  float coef[4];
};

__device__ __forceinline__ AxisTabF axis_entry_f(int dst, int src_len, int dst_len) {
  AxisTabF t;
  const double inv = __ddiv_rn(static_cast<double>(dst_len), static_cast<double>(src_len));
  const double scale = __ddiv_rn(1.0, inv);
  float fx = static_cast<float>(__dsub_rn(__dmul_rn(static_cast<double>(dst) + 0.5, scale), 0.5));
  const float sxf = floorf(fx);
  const float x = __fsub_rn(fx, sxf);
  const float A = -0.75f;
  const float xp1 = __fadd_rn(x, 1.0f);
  const float omx = __fsub_rn(1.0f, x);
  t.coef[0] = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, xp1), __fmul_rn(5.0f, A)), xp1),
                                            __fmul_rn(8.0f, A)), xp1), __fmul_rn(4.0f, A));
  t.coef[1] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(A, 2.0f), x), __fadd_rn(A, 3.0f)), x), x),
                        1.0f);
  t.coef[2] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(A, 2.0f), omx), __fadd_rn(A, 3.0f)), omx),
                                  omx), 1.0f);
  t.coef[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.0f, t.coef[0]), t.coef[1]), t.coef[2]);
  t.first = static_cast<int>(sxf) - 1;
  return t;
}

struct MeanStd {
  double mean[3], stdv[3];
};

// One CTA per output row: image/255. (double) -> cubic with fp32 taps, double accumulation -> normalise -> fp32.
__global__ void __launch_bounds__(GATHER_THREADS) resize_rgb_kernel(const uint8_t* __restrict__ img, int H, int W, int dh,
                                                                    int d, MeanStd ms, float* __restrict__ plane) {
  const int dy = blockIdx.x;                 // output [dh][d][3]: d = output width, dh = output height (`orig` mode: != d)
  const AxisTabF ty = axis_entry_f(dy, H, dh);
  for (int dx = threadIdx.x; dx < d; dx += blockDim.x) {
    const AxisTabF tx = axis_entry_f(dx, W, d);
    double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int yy = min(max(ty.first + k, 0), H - 1);
      double h[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int xx = min(max(tx.first + j, 0), W - 1);
        const uint8_t* px = img + (static_cast<size_t>(yy) * W + xx) * 3;
        const double cj = static_cast<double>(tx.coef[j]);
#pragma unroll
        for (int c = 0; c < 3; ++c) h[c] += (static_cast<double>(px[c]) / 255.0) * cj;
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[c] += h[c] * static_cast<double>(ty.coef[k]);
    }
    float* o = plane + (static_cast<size_t>(dy) * d + dx) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = static_cast<float>((acc[c] - ms.mean[c]) / ms.stdv[c]);
  }
}

// ---- image mode (reference inference.py:377-393): zero-pad to a centred max(H,W) square, cv2.INTER_LINEAR (8-bit
// generic path, bit-exact: 11-bit weights, integer horizontal pass, (((b0*(S0>>4))>>16)+((b1*(S1>>4))>>16)+2)>>2
// vertical pass), transform_rgb -> fp32 plane [d][d][3] shared by all pairs of the image.
struct LinTab {
  int idx;     // first tap (clamped); second tap = idx + 1
  int a0, a1;  // weights (a1 unused when single)
  int single;
};

__device__ __forceinline__ LinTab linear_entry(int dst, int src_len, int dst_len, bool horizontal) {
  LinTab t;
  const double inv = __ddiv_rn(static_cast<double>(dst_len), static_cast<double>(src_len));
  const double scale = __ddiv_rn(1.0, inv);
  float fx = static_cast<float>(__dsub_rn(__dmul_rn(static_cast<double>(dst) + 0.5, scale), 0.5));
  const float sxf = floorf(fx);
  int sx = static_cast<int>(sxf);
  fx = __fsub_rn(fx, sxf);
  t.single = 0;
  if (horizontal) {
    if (sx < 0) { fx = 0.f; sx = 0; }
    if (sx >= src_len - 1) { fx = 0.f; sx = src_len - 1; }
    t.single = (sx + 1 >= src_len);
  }
  t.idx = sx;
  t.a0 = max(-32768, min(32767, __float2int_rn(__fmul_rn(__fsub_rn(1.0f, fx), 2048.0f))));
  t.a1 = max(-32768, min(32767, __float2int_rn(__fmul_rn(fx, 2048.0f))));
  return t;
}

struct LutF {
  float v[3][256];
};

// square = 1: image mode (source = the centred zero-padded max(H, W) square); square = 0: plain H x W -> d x d resize
// (training `resize` mode, datasets/depth_occ_order_dataset.py:83-86)
__global__ void __launch_bounds__(GATHER_THREADS) square_linear_rgb_kernel(const uint8_t* __restrict__ img, int H, int W,
                                                                           int d, const float* __restrict__ lut,
                                                                           float* __restrict__ plane, int square) {
  const int S = max(H, W);
  const int SH = square ? S : H, SW = square ? S : W;
  const int left = square ? (S - W) / 2 : 0, top = square ? (S - H) / 2 : 0;
  const int dy = blockIdx.x;
  const LinTab ty = linear_entry(dy, SH, d, false);
  const int r0 = min(max(ty.idx, 0), SH - 1) - top, r1 = min(max(ty.idx + 1, 0), SH - 1) - top;
  const bool ok0 = r0 >= 0 && r0 < H, ok1 = r1 >= 0 && r1 < H;
  for (int dx = threadIdx.x; dx < d; dx += blockDim.x) {
    const LinTab tx = linear_entry(dx, SW, d, true);
    const int c0 = tx.idx - left, c1 = tx.idx + 1 - left;
    const bool okc0 = c0 >= 0 && c0 < W, okc1 = !tx.single && c1 >= 0 && c1 < W;
    float* o = plane + (static_cast<size_t>(dy) * d + dx) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      int h0, h1;
      {
        const int p0 = (ok0 && okc0) ? img[(static_cast<size_t>(r0) * W + c0) * 3 + c] : 0;
        const int p1 = (ok0 && okc1) ? img[(static_cast<size_t>(r0) * W + c1) * 3 + c] : 0;
        h0 = tx.single ? p0 * 2048 : p0 * tx.a0 + p1 * tx.a1;
      }
      {
        const int p0 = (ok1 && okc0) ? img[(static_cast<size_t>(r1) * W + c0) * 3 + c] : 0;
        const int p1 = (ok1 && okc1) ? img[(static_cast<size_t>(r1) * W + c1) * 3 + c] : 0;
        h1 = tx.single ? p0 * 2048 : p0 * tx.a0 + p1 * tx.a1;
      }
      int v = (((ty.a0 * (h0 >> 4)) >> 16) + ((ty.a1 * (h1 >> 4)) >> 16) + 2) >> 2;
      v = min(max(v, 0), 255);
      o[c] = lut[c * 256 + v];
    }
  }
}

__global__ void __launch_bounds__(GATHER_THREADS) gather_resize_kernel(const float* __restrict__ planes,
                                                                       const uint8_t* __restrict__ masks,
                                                                       const io_pair_desc* __restrict__ descs, int dh,
                                                                       int d, int pitch, __nv_bfloat16* __restrict__ out) {
  // network input dh x d (height x width); dh == d except in `orig` mode
  const io_pair_desc ds = descs[blockIdx.y];
  const uint8_t* __restrict__ ma = masks + ds.mask_a_off;
  const uint8_t* __restrict__ mb = masks + ds.mask_b_off;
  const float* __restrict__ plane = planes + static_cast<size_t>(ds.rgb_slot & 0x3FFFFFFF) * dh * d * 3;
  const bool flip = (ds.rgb_slot & 0x40000000) != 0;   // training augmentation: mirror the resized pair horizontally
  const int H = ds.h, W = ds.w;
  __nv_bfloat16* out_pair = out + static_cast<size_t>(blockIdx.y) * (dh + 6) * pitch * 8;
  const int row0 = blockIdx.x * GATHER_ROWS;
  const int row1 = min(dh, row0 + GATHER_ROWS);
  zero_borders(out_pair, d, pitch, blockIdx.x, gridDim.x, row0, row1, dh);
  // resize mode: nearest over the H x W image; image mode (ds.s > 0): nearest over the centred ds.s-square whose
  // image window starts at (ds.x, ds.y); pixels of the zero padding read as 0
  const int SH = ds.s > 0 ? ds.s : H, SW = ds.s > 0 ? ds.s : W;
  const int left = ds.s > 0 ? ds.x : 0, top = ds.s > 0 ? ds.y : 0;
  const double sy = __ddiv_rn(1.0, __ddiv_rn(static_cast<double>(dh), static_cast<double>(SH)));
  const double sx = __ddiv_rn(1.0, __ddiv_rn(static_cast<double>(d), static_cast<double>(SW)));
  for (int dy = row0; dy < row1; ++dy) {
    const int my = min(static_cast<int>(floor(__dmul_rn(static_cast<double>(dy), sy))), SH - 1) - top;
    for (int dx = threadIdx.x; dx < d; dx += blockDim.x) {
      const int sxp = flip ? d - 1 - dx : dx;           // output column dx shows the un-flipped column sxp
      const int mx = min(static_cast<int>(floor(__dmul_rn(static_cast<double>(sxp), sx))), SW - 1) - left;
      float va = 0.f, vb = 0.f;
      if (my >= 0 && my < H && mx >= 0 && mx < W) {
        const size_t o = static_cast<size_t>(my) * W + mx;
        va = static_cast<float>(ma[o]);
        vb = static_cast<float>(mb[o]);
      }
      const float* px = plane + (static_cast<size_t>(dy) * d + sxp) * 3;
      const uint32_t rg = pack_bf16(px[0], px[1]);
      const uint32_t b = pack_bf16(px[2], 0.0f);
      store_pixel(out_pair + (static_cast<size_t>(dy + 3) * pitch + dx + 3) * 8, va, vb,
                  static_cast<uint16_t>(rg & 0xFFFF), static_cast<uint16_t>(rg >> 16),
                  static_cast<uint16_t>(b & 0xFFFF));
    }
  }
}

// ---- pack: training / validation batches arrive as the reference's collated tensors ---------------------------
// rgb [B,3,D,D], modal1 [B,1,D,D], modal2 [B,1,D,D] fp32 NCHW (what Dataset.__getitem__ + default collate produce,
// reference datasets/*_order_dataset.py) -> pair tensor, i.e. torch.cat([modal1, modal2, rgb], 1) of
// models/supervised_order.py:52 in the stem's layout.
__global__ void __launch_bounds__(GATHER_THREADS) pack_nchw_kernel(const float* __restrict__ rgb,
                                                                   const float* __restrict__ m1,
                                                                   const float* __restrict__ m2, int d, int pitch,
                                                                   __nv_bfloat16* __restrict__ out) {
  const int b = blockIdx.y;
  __nv_bfloat16* out_pair = out + static_cast<size_t>(b) * (d + 6) * pitch * 8;
  const int row0 = blockIdx.x * GATHER_ROWS;
  const int row1 = min(d, row0 + GATHER_ROWS);
  zero_borders(out_pair, d, pitch, blockIdx.x, gridDim.x, row0, row1);
  const size_t plane = static_cast<size_t>(d) * d;
  for (int dy = row0; dy < row1; ++dy) {
    for (int dx = threadIdx.x; dx < d; dx += blockDim.x) {
      const size_t o = static_cast<size_t>(dy) * d + dx;
      const float* px = rgb + static_cast<size_t>(b) * 3 * plane + o;
      const uint32_t rg = pack_bf16(px[0], px[plane]);
      const uint32_t bb = pack_bf16(px[2 * plane], 0.0f);
      store_pixel(out_pair + (static_cast<size_t>(dy + 3) * pitch + dx + 3) * 8, m1[b * plane + o], m2[b * plane + o],
                  static_cast<uint16_t>(rg & 0xFFFF), static_cast<uint16_t>(rg >> 16),
                  static_cast<uint16_t>(bb & 0xFFFF));
    }
  }
}

// ---- bordering ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bordering_kernel(const uint8_t* __restrict__ masks, int H, int W,
                                                        const int32_t* __restrict__ pairs,
                                                        uint8_t* __restrict__ flags) {
  const uint8_t* __restrict__ a = masks + static_cast<size_t>(pairs[2 * blockIdx.x]) * H * W;
  const uint8_t* __restrict__ b = masks + static_cast<size_t>(pairs[2 * blockIdx.x + 1]) * H * W;
  int hit = 0;
  const int total = H * W;
  for (int i = threadIdx.x; i < total && !hit; i += blockDim.x) {
    if (b[i] & 1) {
      const int y = i / W, x = i - y * W;
      uint8_t m = a[i];
      if (y > 0) m = max(m, a[i - W]);
      if (y + 1 < H) m = max(m, a[i + W]);
      if (x > 0) m = max(m, a[i - 1]);
      if (x + 1 < W) m = max(m, a[i + 1]);
      hit |= (m == 1);
    }
  }
  hit = __syncthreads_or(hit);
  if (threadIdx.x == 0) flags[blockIdx.x] = hit ? 1 : 0;
}

// ---- infer_gt_order (reference inference.py:719-739): KINS ground-truth occlusion order from modal / amodal masks --
// one CTA per pair (i < j): skipped unless bordering(i, j); occ_ij = |modal_i == 1 & amodal_j == 1| and vice versa;
// both zero -> no order; occ_ij >= occ_ji -> gt[i,j] = 1, gt[j,i] = 0, else the reverse.
__global__ void __launch_bounds__(256) gt_order_kernel(const uint8_t* __restrict__ modal,
                                                       const uint8_t* __restrict__ amodal, int n, int H, int W,
                                                       const int32_t* __restrict__ pairs, int64_t* __restrict__ mat) {
  const int i = pairs[2 * blockIdx.x], j = pairs[2 * blockIdx.x + 1];
  const size_t hw = static_cast<size_t>(H) * W;
  const uint8_t* __restrict__ a = modal + i * hw;
  const uint8_t* __restrict__ b = modal + j * hw;
  const uint8_t* __restrict__ aa = amodal + i * hw;
  const uint8_t* __restrict__ ab = amodal + j * hw;
  int hit = 0, cij = 0, cji = 0;
  const int total = H * W;
  for (int k = threadIdx.x; k < total; k += blockDim.x) {
    const uint8_t av = a[k], bv = b[k];
    if (bv & 1) {
      const int y = k / W, x = k - y * W;
      uint8_t m = av;
      if (y > 0) m = max(m, a[k - W]);
      if (y + 1 < H) m = max(m, a[k + W]);
      if (x > 0) m = max(m, a[k - 1]);
      if (x + 1 < W) m = max(m, a[k + 1]);
      hit |= (m == 1);
    }
    cij += (av == 1) & (ab[k] == 1);
    cji += (bv == 1) & (aa[k] == 1);
  }
  hit = __syncthreads_or(hit);
  __shared__ int red[2][8];
  for (int o = 16; o > 0; o >>= 1) {
    cij += __shfl_xor_sync(0xffffffffu, cij, o);
    cji += __shfl_xor_sync(0xffffffffu, cji, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = cij; red[1][threadIdx.x >> 5] = cji; }
  __syncthreads();
  if (threadIdx.x == 0 && hit) {
    int sij = 0, sji = 0;
    for (int w = 0; w < 8; ++w) { sij += red[0][w]; sji += red[1][w]; }
    if (sij != 0 || sji != 0) {
      const bool i_over = sij >= sji;
      mat[static_cast<size_t>(i) * n + j] = i_over ? 1 : 0;
      mat[static_cast<size_t>(j) * n + i] = i_over ? 0 : 1;
    }
  }
}

static int check_d(int d) {
  IO_REQUIRE(d >= 32 && d <= MAX_D && d % 2 == 0, "input size %d not supported (even, 32..%d)", d, MAX_D);
  return IO_OK;
}

}  // namespace io

using namespace io;

extern "C" int io_pair_enumerate(int n, int32_t* out) {
  IO_REQUIRE(n >= 0 && (out || n < 2), "io_pair_enumerate: bad arguments");
  int c = 0;
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) {
      out[2 * c] = i;
      out[2 * c + 1] = j;
      ++c;
    }
  return c;
}

static inline double max3(double a, double b, double c) {
  double m = a;
  if (b > m) m = b;
  if (c > m) m = c;
  return m;
}

extern "C" int io_expand_bbox(const double* boxes, int n, double enlarge, int32_t* out) {
  IO_REQUIRE(boxes && out && n >= 0, "io_expand_bbox: bad arguments");
  for (int k = 0; k < n; ++k) {
    const double x = boxes[4 * k], y = boxes[4 * k + 1], w = boxes[4 * k + 2], h = boxes[4 * k + 3];
    const double cx = x + w / 2.0, cy = y + h / 2.0;
    const double size = max3(sqrt(w * h * enlarge), w * 1.1, h * 1.1);
    out[4 * k + 0] = static_cast<int32_t>(cx - size / 2.0);  // python int(): truncation toward zero
    out[4 * k + 1] = static_cast<int32_t>(cy - size / 2.0);
    out[4 * k + 2] = static_cast<int32_t>(size);
    out[4 * k + 3] = static_cast<int32_t>(size);
  }
  return IO_OK;
}

extern "C" int io_pair_crop_boxes(const double* boxes, const int32_t* pairs, int p, int32_t* out) {
  IO_REQUIRE(boxes && pairs && out && p >= 0, "io_pair_crop_boxes: bad arguments");
  int degenerate = -1;
  for (int k = 0; k < p; ++k) {
    const double* a = boxes + 4 * pairs[2 * k];
    const double* b = boxes + 4 * pairs[2 * k + 1];
    const double l = a[0] < b[0] ? a[0] : b[0];
    const double u = a[1] < b[1] ? a[1] : b[1];
    const double ra = a[0] + a[2], rb = b[0] + b[2];
    const double ba = a[1] + a[3], bb = b[1] + b[3];
    const double r = ra > rb ? ra : rb;
    const double bt = ba > bb ? ba : bb;
    const double w = r - l, h = bt - u;
    const double cx = l + w / 2.0, cy = u + h / 2.0;
    const double size = max3(sqrt(w * h * 2.0), w * 1.1, h * 1.1);
    out[4 * k + 0] = static_cast<int32_t>(cx - size / 2.0);
    out[4 * k + 1] = static_cast<int32_t>(cy - size / 2.0);
    out[4 * k + 2] = static_cast<int32_t>(size);
    out[4 * k + 3] = static_cast<int32_t>(size);
    if (out[4 * k + 2] <= 0 && degenerate < 0) degenerate = k;
  }
  if (degenerate >= 0) {
    set_error("pair %d (%d,%d) is degenerate: crop side int(size) == %d (cv2.resize asserts in the reference)",
              degenerate, pairs[2 * degenerate], pairs[2 * degenerate + 1], out[4 * degenerate + 2]);
    return IO_ERR_DEGENERATE;
  }
  return IO_OK;
}

extern "C" int64_t io_pair_tensor_row_pitch(int d) { return (static_cast<int64_t>(d) + 6 + 7) / 8 * 8; }
extern "C" int64_t io_pair_tensor_bytes(int p, int d) {
  return static_cast<int64_t>(p) * (d + 6) * io_pair_tensor_row_pitch(d) * 8 * 2;
}

extern "C" int io_normalize_lut(const float* mean, const float* stdv, float* out) {
  IO_REQUIRE(mean && stdv && out, "io_normalize_lut: null pointer");
  fill_lut_f32(mean, stdv, out);
  return IO_OK;
}

extern "C" int io_pair_bordering(const uint8_t* masks, int n, int h, int w, const int32_t* pairs, int p,
                                 uint8_t* flags, void* stream) {
  IO_REQUIRE(masks && pairs && flags && n > 0 && h > 0 && w > 0 && p >= 0, "io_pair_bordering: bad arguments");
  if (p == 0) return IO_OK;
  bordering_kernel<<<p, 256, 0, as_stream(stream)>>>(masks, h, w, pairs, flags);
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}

// ---- per-instance mask statistics for the heuristic baselines (reference inference.py:272-346) -----------------
// out[i] = (sum of mask values, number of pixels == 1, sum of the row index over pixels == 1), exact integers
__global__ void __launch_bounds__(256) mask_stats_kernel(const uint8_t* __restrict__ masks, int h, int w,
                                                         unsigned long long* __restrict__ out) {
  const int inst = blockIdx.y;
  const uint8_t* m = masks + static_cast<size_t>(inst) * h * w;
  unsigned long long sv = 0, cnt = 0, sy = 0;
  const size_t total = static_cast<size_t>(h) * w;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const unsigned v = m[i];
    sv += v;
    if (v == 1u) { cnt += 1; sy += static_cast<unsigned long long>(i / w); }
  }
  for (int o = 16; o > 0; o >>= 1) {
    sv += __shfl_xor_sync(0xffffffffu, sv, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out + inst * 3 + 0, sv);
    atomicAdd(out + inst * 3 + 1, cnt);
    atomicAdd(out + inst * 3 + 2, sy);
  }
}

extern "C" int io_mask_stats(const uint8_t* masks, int n, int h, int w, int64_t* out, void* stream) {
  IO_REQUIRE(masks && out && n > 0 && h > 0 && w > 0, "io_mask_stats: bad arguments");
  IO_CUDA(cudaMemsetAsync(out, 0, sizeof(int64_t) * 3 * n, as_stream(stream)));
  const int gx = static_cast<int>(std::min<size_t>((static_cast<size_t>(h) * w + 256 * 8 - 1) / (256 * 8), 64));
  mask_stats_kernel<<<dim3(gx, n), 256, 0, as_stream(stream)>>>(masks, h, w,
                                                                reinterpret_cast<unsigned long long*>(out));
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}

extern "C" int io_infer_gt_order(const uint8_t* modal, const uint8_t* amodal, int n, int h, int w, const int32_t* pairs,
                                 int p, int64_t* mat, void* stream) {
  IO_REQUIRE(modal && amodal && pairs && mat && n > 0 && h > 0 && w > 0 && p >= 0, "io_infer_gt_order: bad arguments");
  if (p == 0) return IO_OK;
  gt_order_kernel<<<p, 256, 0, as_stream(stream)>>>(modal, amodal, n, h, w, pairs, mat);
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}

extern "C" int io_pair_gather_patch(const uint8_t* images, const uint8_t* masks, const io_pair_desc* descs, int p,
                                    int d, const float* mean, const float* stdv, void* out, void* stream) {
  IO_REQUIRE(images && masks && descs && mean && stdv && out && p >= 0, "io_pair_gather_patch: bad arguments");
  if (int rc = check_d(d)) return rc;
  if (p == 0) return IO_OK;
  float lutf[768];
  fill_lut_f32(mean, stdv, lutf);
  NormLut lut;
  for (int i = 0; i < 768; ++i) (&lut.v[0][0])[i] = f32_to_bf16_rn(lutf[i]);
  dim3 grid((d + GATHER_ROWS - 1) / GATHER_ROWS, p);
  const size_t smem = static_cast<size_t>(d) * (sizeof(AxisTab) + 8 * 3 * sizeof(int));
  const int pitch = static_cast<int>(io_pair_tensor_row_pitch(d));
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  if (d <= 2 * GATHER_THREADS) {
    gather_patch_kernel<2><<<grid, GATHER_THREADS, smem, as_stream(stream)>>>(images, masks, descs, d, pitch, lut, o);
  } else {
    static bool attr_set = false;
    if (!attr_set) {
      IO_CUDA(cudaFuncSetAttribute(gather_patch_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   MAX_D * (sizeof(AxisTab) + 8 * 3 * sizeof(int))));
      attr_set = true;
    }
    gather_patch_kernel<4><<<grid, GATHER_THREADS, smem, as_stream(stream)>>>(images, masks, descs, d, pitch, lut, o);
  }
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}

extern "C" int io_pair_pack_nchw(const float* rgb, const float* m1, const float* m2, int b, int d, void* out,
                                 void* stream) {
  IO_REQUIRE(rgb && m1 && m2 && out && b >= 0, "io_pair_pack_nchw: bad arguments");
  if (int rc = check_d(d)) return rc;
  if (b == 0) return IO_OK;
  dim3 grid((d + GATHER_ROWS - 1) / GATHER_ROWS, b);
  pack_nchw_kernel<<<grid, GATHER_THREADS, 0, as_stream(stream)>>>(rgb, m1, m2, d,
                                                                   static_cast<int>(io_pair_tensor_row_pitch(d)),
                                                                   reinterpret_cast<__nv_bfloat16*>(out));
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}

extern "C" int io_image_resize_rgb(const uint8_t* image, int h, int w, int d, const float* mean, const float* stdv,
                                   float* plane, void* stream) {
  IO_REQUIRE(image && mean && stdv && plane && h > 0 && w > 0, "io_image_resize_rgb: bad arguments");
  if (int rc = check_d(d)) return rc;
  MeanStd ms;
  for (int c = 0; c < 3; ++c) {
    ms.mean[c] = static_cast<double>(mean[c]);
    ms.stdv[c] = static_cast<double>(stdv[c]);
  }
  resize_rgb_kernel<<<d, GATHER_THREADS, 0, as_stream(stream)>>>(image, h, w, d, d, ms, plane);
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}

// `orig` mode (reference inference.py:401-408): the image at its own size rounded to multiples of 32 -- a dh x dw network
// input.  Same arithmetic as io_image_resize_rgb (transform_resize: /255, cv2.INTER_CUBIC on float64, normalise).
extern "C" int io_image_resize_rgb_hw(const uint8_t* image, int h, int w, int dh, int dw, const float* mean,
                                      const float* stdv, float* plane, void* stream) {
  IO_REQUIRE(image && mean && stdv && plane && h > 0 && w > 0, "io_image_resize_rgb_hw: bad arguments");
  IO_REQUIRE(dh >= 32 && dw >= 32 && dh <= 1024 && dw <= 1024 && dh % 32 == 0 && dw % 32 == 0,
             "io_image_resize_rgb_hw: network input %d x %d (multiples of 32 in [32, 1024])", dh, dw);
  MeanStd ms;
  for (int c = 0; c < 3; ++c) {
    ms.mean[c] = static_cast<double>(mean[c]);
    ms.stdv[c] = static_cast<double>(stdv[c]);
  }
  resize_rgb_kernel<<<dh, GATHER_THREADS, 0, as_stream(stream)>>>(image, h, w, dh, dw, ms, plane);
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}

extern "C" int64_t io_pair_tensor_bytes_hw(int p, int dh, int dw) {
  return static_cast<int64_t>(p) * (dh + 6) * io_pair_tensor_row_pitch(dw) * 8 * 2;
}

// per pair: nearest-resized masks (cv2.INTER_NEAREST to (dw, dh)) + the image's plane -> pair tensor [p][dh + 6][pitch(dw)][8]
extern "C" int io_pair_gather_resize_hw(const float* planes, const uint8_t* masks, const io_pair_desc* descs, int p,
                                        int dh, int dw, void* out, void* stream) {
  IO_REQUIRE(planes && masks && descs && out && p >= 0, "io_pair_gather_resize_hw: bad arguments");
  IO_REQUIRE(dh >= 32 && dw >= 32 && dh <= 1024 && dw <= 1024 && dh % 32 == 0 && dw % 32 == 0,
             "io_pair_gather_resize_hw: network input %d x %d (multiples of 32 in [32, 1024])", dh, dw);
  if (p == 0) return IO_OK;
  dim3 grid((dh + GATHER_ROWS - 1) / GATHER_ROWS, p);
  gather_resize_kernel<<<grid, GATHER_THREADS, 0, as_stream(stream)>>>(
      planes, masks, descs, dh, dw, static_cast<int>(io_pair_tensor_row_pitch(dw)), reinterpret_cast<__nv_bfloat16*>(out));
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}

extern "C" int io_image_square_linear_rgb(const uint8_t* image, int h, int w, int d, const float* mean, const float* stdv,
                                          float* lut_dev_scratch, float* plane, void* stream) {
  IO_REQUIRE(image && mean && stdv && lut_dev_scratch && plane && h > 0 && w > 0,
             "io_image_square_linear_rgb: bad arguments");
  if (int rc = check_d(d)) return rc;
  float lutf[768];
  fill_lut_f32(mean, stdv, lutf);
  IO_CUDA(cudaMemcpyAsync(lut_dev_scratch, lutf, sizeof(lutf), cudaMemcpyHostToDevice, as_stream(stream)));
  IO_CUDA(cudaStreamSynchronize(as_stream(stream)));   // lutf lives on this stack frame
  square_linear_rgb_kernel<<<d, GATHER_THREADS, 0, as_stream(stream)>>>(image, h, w, d, lut_dev_scratch, plane, 1);
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}

extern "C" int io_image_resize_linear_rgb(const uint8_t* image, int h, int w, int d, const float* mean, const float* stdv,
                                          float* lut_dev_scratch, float* plane, void* stream) {
  IO_REQUIRE(image && mean && stdv && lut_dev_scratch && plane && h > 0 && w > 0,
             "io_image_resize_linear_rgb: bad arguments");
  if (int rc = check_d(d)) return rc;
  float lutf[768];
  fill_lut_f32(mean, stdv, lutf);
  IO_CUDA(cudaMemcpyAsync(lut_dev_scratch, lutf, sizeof(lutf), cudaMemcpyHostToDevice, as_stream(stream)));
  IO_CUDA(cudaStreamSynchronize(as_stream(stream)));   // lutf lives on this stack frame
  square_linear_rgb_kernel<<<d, GATHER_THREADS, 0, as_stream(stream)>>>(image, h, w, d, lut_dev_scratch, plane, 0);
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}

extern "C" int io_pair_gather_resize(const float* planes, const uint8_t* masks, const io_pair_desc* descs, int p,
                                     int d, void* out, void* stream) {
  IO_REQUIRE(planes && masks && descs && out && p >= 0, "io_pair_gather_resize: bad arguments");
  if (int rc = check_d(d)) return rc;
  if (p == 0) return IO_OK;
  dim3 grid((d + GATHER_ROWS - 1) / GATHER_ROWS, p);
  gather_resize_kernel<<<grid, GATHER_THREADS, 0, as_stream(stream)>>>(
      planes, masks, descs, d, d, static_cast<int>(io_pair_tensor_row_pitch(d)), reinterpret_cast<__nv_bfloat16*>(out));
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}
