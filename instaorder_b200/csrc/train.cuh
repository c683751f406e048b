// Training step of the order networks (reference models/supervised_order.py:83-95 `step`, train-mode
// models/backbone/resnet_cls.py:75-222): declarations shared by wgrad.cu, train_ew.cu and train.cu.
#pragma once
#include "conv_tc.cuh"

namespace io {

// ---- weight gradient on tcgen05 (wgrad.cu) ------------------------------------------------------------------
// dW[cout][tap][cin] += sum over output pixels of dy[pixel][cout] * x[pixel shifted by tap][cin]: a GEMM whose
// contraction index is the PIXEL, so both operands are consumed MN-major straight from their NHWC tensors.
struct WgradParams {
  CUtensorMap map_dy;   // [rows_out][cout] bf16, box {64 channels, kp rows}
  CUtensorMap map_x;    // forward-input view of the convolution (as ConvParams::map_a) with kp-pixel boxes
  float* dw;            // [cout][ldw] fp32, accumulated with red.global.add (zero it first)
  int mode;             // ConvMode of the forward convolution
  int cout;             // GEMM M (128 for the two-direction stem)
  int cin;              // GEMM N per filter tap (64 = 8 taps x 8 channels for the stem); also the S2 parity offset
  int taps, taps_w, pad;
  int ldw;
  int m_tiles, n_tiles, bn;
  // filter taps handled by one work item: when cin <= 128 several taps share one accumulator tile (N = ntaps * cin
  // <= 256), so the dy tile is loaded -- and the M = 128 side of the MMA paid for -- once per group instead of once
  // per tap
  int ngroups;
  int g_tap0[9], g_ntaps[9];
  int max_b_slabs;      // B slabs per stage (largest group)
  int acc_cols;         // TMEM columns of one accumulator (largest group's N)
  int tmem_cols;        // allocation: 2 accumulators, power of two
  int a_slabs;          // 64-channel slabs loaded per A stage (1 when cout == 64: upper accumulator rows unused)
  int a_split_rows;     // stem: slab j = channels [0,64) of rows + j * a_split_rows (direction j); 0 = plain
  int kblocks, ksplit, kp;
  int wseg, bh, bi, bpr, bpi;   // K-block geometry: box {wseg, bh, bi} pixels, blocks per row / per image
  int w_out, h_out;             // output size of the convolution (K blocks walk it row-major)
  int stages, smem_bytes;
  int mn_lbo, mn_sbo;
  uint32_t debug_idesc_xor;   // bring-up timing experiments only
  int debug_skip_flush;
  double flops;
};
int wgrad_plan(WgradParams* p, const ConvDesc& d, const void* x, const void* dy, float* dw);
// x = pair tensor [pairs, d+6, pitch, 8]; dy = [2][pairs][d/2][d/2][64] ([direction][pair] image order);
// dw_scratch = [128][448] fp32 in the packed stem-GEMM layout (rows 64.. = swapped-direction weights)
int stem_wgrad_plan(WgradParams* p, int pairs, int d, const void* x, const void* dy, float* dw_scratch);
int wgrad_launch(const WgradParams& p, cudaStream_t stream);

// ---- element-wise / reduction kernels (train_ew.cu) ---------------------------------------------------------
// All activations NHWC bf16 viewed as [groups][rows][c]; BatchNorm statistics are per (group, channel): the two
// directions of a training step are two separate forward passes in the reference (two BN batches).
// stats: sums[g][0][c] = sum x, sums[g][1][c] = sum x^2 (double, must be zero on entry)
int bn_stats_launch(const void* y, int groups, int rows, int c, double* sums, cudaStream_t stream);
// finalize + apply in one launch: mean / biased var from `sums` -> a = [relu](y * scale + shift [+ residual]);
// block (0,0) stores save[4][groups][c] = (scale, shift, mean, invstd) for the backward pass, updates running_mean /
// running_var (momentum, unbiased variance, once per group in group order) and zeroes `zero_me` (optional: the sums
// buffer the NEXT BatchNorm will accumulate into -- never the one being read)
int bn_apply_launch(const void* y, const void* residual, void* a, int groups, int rows, int c, const double* sums,
                    const float* gamma, const float* beta, float eps, float momentum, float* save, float* running_mean,
                    float* running_var, double* zero_me, int relu, uint8_t* mask_out, cudaStream_t stream);
// backward of a = [relu](bn(y) [+ residual]):  g = da * mask;  red[g][0][c] = sum g, red[g][1][c] = sum g * xhat
// (double, must be zero on entry).  mask_mode 0: none, 1: a > 0 (stored activation), 2: y * scale + shift > 0,
// 3: bit mask written by bn_apply_launch's mask_out (passed through `a`: one byte per 8 channels)
int bn_bwd_reduce_launch(const void* da, const void* a, const void* y, int groups, int rows, int c, const float* save,
                         int mask_mode, double* red, cudaStream_t stream);
// dy = gamma * invstd * (g - sum_g / M - xhat * sum_gx / M)  (bf16);  if g_out != nullptr also writes g (the
// gradient flowing into the residual branch);  accumulates dgamma += sum_gx, dbeta += sum_g into the flat grads;
// zeroes `zero_me` (optional, as above)
int bn_bwd_apply_launch(const void* da, const void* a, const void* y, void* dy, void* g_out, int groups, int rows,
                        int c, const float* gamma, const float* save, const double* red, int mask_mode, float* dgamma,
                        float* dbeta, double* zero_me, cudaStream_t stream);
// one-launch variants (cooperative grid, grid-wide barrier between the two passes); `barrier` = one zeroed counter
int bn_fwd_fused_launch(const void* y, const void* residual, void* a, int groups, int rows, int c, double* sums,
                        const float* gamma, const float* beta, float eps, float momentum, float* save,
                        float* running_mean, float* running_var, double* zero_me, int relu, uint8_t* mask_out,
                        unsigned int* barrier, cudaStream_t stream);
int bn_bwd_fused_launch(const void* da, const void* a, const void* y, void* dy, void* g_out, int groups, int rows,
                        int c, const float* gamma, const float* save, double* red, int mask_mode, float* dgamma,
                        float* dbeta, double* zero_me, unsigned int* barrier, cudaStream_t stream);
// max-pool 3x3 s2 p1 forward with arg-max (first maximum in window scan order) and its backward
int maxpool_fwd_idx_launch(const void* x, void* y, uint8_t* idx, int b, int h, int w, int c, cudaStream_t stream);
int maxpool_bwd_launch(const void* dy, const uint8_t* idx, void* dx, int b, int h, int w, int c, cudaStream_t stream);
// zero-insertion up-sampling of a stride-2 gradient: z[n, 2i, 2j, :] = dy[n, i, j, :], zero elsewhere
int upsample2_zero_launch(const void* dy, void* z, int b, int ho, int wo, int c, cudaStream_t stream);
// dx[n, 2i, 2j, :] += d[n, i, j, :]   (data gradient of the stride-2 1x1 downsample convolution)
int scatter_add2_launch(const void* d, void* dx, int b, int ho, int wo, int c, cudaStream_t stream);
// average pool + FC heads: pooled [imgs][2048] fp32 (saved), logits [imgs][k] fp32
int pool_fc_fwd_launch(const void* feat, int hw, int imgs, const float* fcw, const float* fcb, int k_total,
                       float* pooled, float* logits, cudaStream_t stream);
// dW_fc[k][2048] += dlogits^T pooled, db += sum dlogits, dfeat[img][hw][2048] = (dlogits W_fc) / hw  (bf16)
int pool_fc_bwd_launch(const float* dlogits, const float* pooled, const float* fcw, int hw, int imgs, int k_total,
                       float* dfcw, float* dfcb, void* dfeat, cudaStream_t stream);

}  // namespace io
