// Implicit-GEMM convolution on tcgen05 tensor cores (declarations shared by conv_tc.cu and net.cu).
#pragma once
#include "common.cuh"

namespace io {

enum ConvMode : int {
  CONV_GEMM = 0,   // 1x1 stride 1: A is the flat [M, Cin] matrix (2-D tensor map)
  CONV_S1 = 1,     // 3x3 stride 1 pad 1: A is {C, W, H, N} (4-D), one box per filter tap, OOB = zero padding
  CONV_S2 = 2,     // k x k stride 2: A is the parity view {2C, W/2, 2, H/2, N} (5-D) of the input
  CONV_STEM = 3,   // conv1 7x7 stride 2 on the padded 8-channel pair tensor: overlapping-window 4-D map
};

struct ConvDesc {
  int b, h, w, cin, cout, kernel, stride;
};

struct ConvParams {
  CUtensorMap map_a;
  CUtensorMap map_b;
  CUtensorMap map_out;  // [rows, ldc] bf16 output, box 64 columns x 32 rows, 128B swizzle (epilogue TMA store)
  CUtensorMap map_res;  // same geometry over the residual tensor (valid only if residual != nullptr)
  const float* bias;              // [n_total] fp32 (folded BN shift)
  const __nv_bfloat16* residual;  // optional, indexed like out
  __nv_bfloat16* out;             // [rows, ldc] bf16
  int mode;
  int m_total;        // valid output rows (pixels over the whole batch)
  int n_total;        // GEMM N (= Cout, or 2*64 for the two-direction stem)
  int k_iters;        // number of 64-wide K blocks
  int kpt;            // K blocks per filter tap (Cin / 64)
  int taps_w;         // filter width (3 or 1)
  int pad;            // filter padding (1 or 0)
  int cin;            // input channels (parity offset in CONV_S2)
  int a_bytes;        // bytes one A box brings (rows_per_tile * 128)
  // tile -> output rows
  int m_tiles, n_tiles;
  int rows_per_tile;  // <= 128
  int tpg;            // tiles per image group
  int bi, bh;         // images / output rows per tile
  int w_out, hw_out;
  int tpr;            // CONV_STEM: tiles per output row
  int ldc;            // output row stride in elements
  int n_split;        // columns >= n_split go to rows + split_row_off (stem: second direction); else n_total
  int split_row_off;
  int relu;
  // weight-stationary mode: when the layer has a single N tile and all its K blocks of weights fit in shared memory
  // they are loaded once per CTA and only the activation tiles stream through the ring
  int prefetch;       // producer pulls the next tile's activation rows into L2 one tile ahead
  int b_resident;
  int stages;         // ring depth (runtime: depends on how much shared memory the resident weights take)
  int smem_bytes;     // dynamic shared memory of the launch
  // training (data-gradient) mode: B = the FORWARD weights [Cout_f][taps * Cin_f] used as an MN-major operand
  // (K = Cout_f rows, N = Cin_f contiguous), filter taps visited mirrored (tap offsets negated) -- dgrad_plan()
  int b_mn;           // 1: B boxes are {64 n, 64 k} MN-major slabs of 8 KB, UMMA b_major = MN
  int flip;           // 1: activation box of tap (r, s) is shifted by (1 - r, 1 - s) instead of (r - 1, s - 1)
  int mn_lbo, mn_sbo; // b_mn: descriptor byte offsets (8192 / 1024)
  int b_tap_cols;     // b_mn: columns per filter tap in the weight matrix (= Cin_f)
  int img_mul;        // CONV_STEM: output image of pair n, direction 0 is img_mul * n (2 = interleaved, 1 = [dir][pair])
  // dual-source K (conv_plan_dual): the first k1 K blocks of every tile come from a second, flat [m_total, 64*k1]
  // matrix (map_a2, box 64 x rows_per_tile at the tile's first output row); the remaining blocks from map_a as usual
  CUtensorMap map_a2;
  int k1;
  // CTA-pair kernel (cta_group::2, N = 256 tiles): map_b boxes are 64 x 128 (each CTA of the pair stages half the tile's
  // weight rows), 4-stage ring of 32 KB
  int pair;
  int epi_one_slot;   // pair kernel without a residual: one epilogue staging slot per warp, one more ring stage
};

// conv3 (+ residual + ReLU) of one bottleneck fused with conv1 (+ ReLU) of the next one (conv_fused.cu)
struct FusedParams {
  ConvParams c;          // the conv3 GEMM: CONV_GEMM mode, 128-column units (map_b box 64 x 128), residual required
  CUtensorMap map_b2;    // next conv1 weights [n2][n_total] bf16, box 64 x n2
  __nv_bfloat16* out2;   // next block's T1 [rows][n2] bf16
  const float* bias2;
  int n2;
  int st1, st2;          // ring depths: (x + w3) stages of 32 KB, next-conv1 weight slots of n2 * 128 B
  int smem_bytes;
  // block 0 of a layer: K blocks [0, k1a) of the first GEMM come from c.map_a (conv2 output), the rest from c.map_a2,
  // the block input seen through the downsample convolution's activation view (flat, or the stride-2 parity view)
  int k1a;
  int a2_mode, a2_tpg, a2_bi, a2_bh;
  // sub_w > 0: the block output is written sub-sampled (even rows / columns of sub_w-wide images of sub_hw pixels) as a
  // compact [img][h/2][w/2][n1] tensor at c.out -- set by conv_fused_set_subsampled()
  int sub_w, sub_hw;
  // 1: the identity is added by the tensor pipe (D1 += residual tile x 64 x 64 identity matrix) instead of by the epilogue
  // warps, which were the bottleneck of the layer1 / layer2 launches
  int res_mma;
  int st256;   // next-conv1 output rows written with 256-bit stores (whole 32-byte sectors per lane)
  int pair;    // CTA-pair kernel (cta_group::2): map_b boxes 64 x 64, map_b2 boxes 64 x n2 / 2, halved weight slots
};
void conv_fused_set_subsampled(FusedParams* fp, int h, int w);
bool conv_fused_supported(int cmid, int n1, int n2, const ConvDesc* ds);
int conv_fused_plan(FusedParams* fp, int rows, int cmid, int n1, int n2, const void* t2, const void* wb1,
                    const float* bias1, const void* residual, void* y, const void* w1n, const float* bias2, void* y2,
                    const ConvDesc* ds, const void* x);
int conv_fused_launch(const FusedParams& fp, cudaStream_t stream);

// Transposed kernel (conv_tn.cu): 128 output channels on the MMA's M side, 256 pixels on its N side
struct TnParams {
  CUtensorMap map_w;    // weights [128][Ktot] bf16, box 64 x 128
  CUtensorMap map_x;    // activation view (as in ConvParams::map_a) with 256-pixel boxes (128 for the stem)
  CUtensorMap map_out;  // [rows][ldc] bf16, box 64 x 32
  const float* bias;
  int ch;               // 128 or 64 output channels (MMA M)
  int mode, k_iters, kpt, taps_w, pad, cin;
  int tiles, tpi, bh, w_out, hw_out;
  int px;               // pixels per tile = MMA N: 256, or 192 (two 96-wide / four 48-wide rows at 384^2)
  int ldc, n_split, split_row_off, relu;
  int img_mul;          // CONV_STEM: output image of pair n, direction 0 (2 = interleaved 2n + dir, 1 = [dir][pair])
};
bool conv_tn_supported(const ConvDesc& d);
int conv_tn_plan(TnParams* p, const ConvDesc& d, const void* x, const void* wgt, const float* bias, void* y, int relu);
bool stem_tn_supported(int d);
int stem_tn_plan(TnParams* p, int pairs, int d, const void* x, const void* wgt, const float* bias, void* y);
int conv_tn_launch(const TnParams& p, cudaStream_t stream);
bool tn_enabled();

// 3x3 stride-1 64 -> 64 convolution on 64 x 64 maps with the input rows resident in shared memory (conv_halo.cu)
struct HaloParams {
  CUtensorMap map_w;    // weights [64][9 * 64] bf16, box 64 x 64 (one filter tap), 128B swizzle
  CUtensorMap map_x;    // {C, W, H, N} view of the NHWC input, box 64 x 72 x 1 x 1 (x = -1 .. 70), 128B swizzle
  CUtensorMap map_out;  // [rows][64] bf16, box 64 x 32
  const float* bias;
  int items, strips, strip_len, relu;
};
bool conv_halo_supported(const ConvDesc& d);
int conv_halo_plan(HaloParams* p, const ConvDesc& d, const void* x, const void* wgt, const float* bias, void* y, int relu);
int conv_halo_launch(const HaloParams& p, cudaStream_t stream);
// same geometry at the full tensor rate: horizontal taps stacked along N (N = 192), shift in the epilogue (conv_row3.cu)
bool conv_row3_supported(const ConvDesc& d);
int conv_row3_plan(HaloParams* p, const ConvDesc& d, const void* x, const void* wgt, const float* bias, void* y, int relu);
int conv_row3_launch(const HaloParams& p, cudaStream_t stream);

// conv1 + BN + ReLU + MaxPool2d(3, 2, 1) in one kernel for 256 x 256 inputs (stem_pool.cu)
struct StemPoolParams {
  CUtensorMap map_w;    // weights [128][448] bf16 in stem_pool_pack_k order, box 64 x 128, 128B swizzle
  CUtensorMap map_x;    // parity view {8, 2, pitch / 2, d + 6, pairs} of the pair tensor, box = one parity array of a row
  CUtensorMap map_out;  // pooled output [2 * pairs * (d/4)^2][64] bf16, box 64 x 32, 128B swizzle
  const float* bias;    // [128]
  int items, strips, strip_len;
  int img_mul;          // output image of pair n, direction 0 (2 = interleaved 2n + dir, 1 = [dir][pair])
  int split_row_off;    // output rows added for direction 1
};
bool stem_pool_supported(int d);
int stem_pool_plan(StemPoolParams* p, int pairs, int d, const void* x, const void* wgt, const float* bias, void* y);
int stem_pool_pack_k(int r, int s, int c);
int stem_pool_launch(const StemPoolParams& p, cudaStream_t stream);

// Launches the persistent kernel for one convolution. bn_tile in {64, 128, 256}.
int conv_tc_launch(const ConvParams& p, int bn_tile, cudaStream_t stream);

// Fills tile geometry + tensor maps for a conv over NHWC bf16 input [b, h, w, cin] (or the pair tensor for the stem).
int conv_plan(ConvParams* p, int* bn_tile, const ConvDesc& d, const void* x, const void* wgt, const float* bias,
              const void* residual, void* y, int relu);
// First block of a ResNet layer: out = ReLU(conv3(t2) + bn3 + downsample(x) + bn_ds) as ONE GEMM over the concatenated
// K = [t2 channels | x channels] (reference resnet_cls.py:107-116 with :112-113).  `ds` describes the 1x1 downsample
// convolution over x (stride 1 or 2); t2 is the flat [b*ho*wo, cmid] conv2 output; wcat = [cout][cmid + ds.cin] bf16
// (conv3 columns first), bias = the two folded-BN shifts added.  The identity tensor is never written to HBM.
int conv_plan_dual(ConvParams* p, int* bn_tile, const ConvDesc& ds, const void* x, const void* t2, int cmid,
                   const void* wcat, const float* bias, void* y, int relu);
// Stem plan: x is the padded pair tensor [pairs, d+6, pitch, 8]; y is [2*pairs, d/2, d/2, 64] (image 2p + dir).
int stem_plan(ConvParams* p, int* bn_tile, int pairs, int d, const void* x, const void* wgt, const float* bias, void* y);
int stem_plan_hw(ConvParams* p, int* bn_tile, int pairs, int h, int w, const void* x, const void* wgt,
                 const float* bias, void* y);
// Data gradient of a stride-1 convolution (1x1 or 3x3 pad 1) as a convolution over dy [b, h, w, cout_f] with the
// forward weights wgt_f [cout_f][k*k*cin_f] read MN-major and the taps mirrored; dx [b, h, w, cin_f] (+ residual).
int dgrad_plan(ConvParams* p, int* bn_tile, int b, int h, int w, int cin_f, int cout_f, int kernel, const void* dy,
               const void* wgt_f, const float* zero_bias, const void* residual, void* dx);

}  // namespace io
