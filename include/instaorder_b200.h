/*
 * instaorder_b200 -- C ABI of the B200-native pairwise-order hot path (libinstaorder_b200.so).
 *
 * The reference (POSTECH-CVLab/InstaOrder) is pure Python and has no FFI of its own; its boundary for this path is
 * the Python API in inference.py / models/*.py.  This header is the interface a reference maintainer would bind
 * with ctypes to replace the per-pair cv2 / torch-eager work with hand-written sm_100a kernels (INTEGRATION.md
 * shows the stub).  Every entry point cites the reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / C++ types.  `_dev` pointers are device memory owned by the caller
 *     (e.g. torch tensors), `_host` pointers are host memory.  The library owns only what lives inside io_net_t.
 *   - every device call takes the cudaStream_t to launch on (as void*) and never synchronises.
 *   - return value: 0 = ok, negative = error (IO_ERR_*); io_last_error() returns the message (thread-local).
 *   - one io_net_t per device; a handle is not thread-safe.
 *   - there is no CPU fallback: a device entry point on a box without a GPU returns IO_ERR_CUDA.
 */
#ifndef INSTAORDER_B200_H_
#define INSTAORDER_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IO_ABI_VERSION 1

#if defined(IO_BUILD)
#define IO_API __attribute__((visibility("default")))
#else
#define IO_API
#endif

#define IO_OK 0
#define IO_ERR_ARG (-1)       /* bad argument (shape, null pointer, unsupported size) */
#define IO_ERR_CUDA (-2)      /* CUDA runtime / driver error */
#define IO_ERR_DEGENERATE (-3) /* a pair whose crop side int(size) == 0: cv2.resize asserts in the reference */
#define IO_ERR_STATE (-4)     /* handle used before weights were loaded, etc. */

/* head kinds: how logits are turned into order decisions */
#define IO_HEAD_OCC 1       /* 2 logits, sigmoid          -- inference.py:196-214 net_forward_occ        */
#define IO_HEAD_DEPTH 2     /* 3 logits, softmax          -- inference.py:172-193 net_forward_depth      */
#define IO_HEAD_ORDERNET 3  /* 3 or 4 logits, softmax     -- inference.py:44-76   net_forward_OrderNet   */

IO_API int io_abi_version(void);
IO_API const char* io_last_error(void);

/* ------------------------------------------------------------------------------------------------------------ */
/* G -- pair construction                                                                                        */
/* ------------------------------------------------------------------------------------------------------------ */

/* Row-major (i, j), i < j, into out_pairs_host[2 * N(N-1)/2]; returns the count.
 * Replaces the loops at inference.py:355-356, 443-444, 521-522. */
IO_API int io_pair_enumerate(int n, int32_t* out_pairs_host);

/* Tester.expand_bbox, tools/test.py:155-163: xywh (float64) -> square int boxes; out_host[4 n]. */
IO_API int io_expand_bbox(const double* boxes_host, int n, double enlarge_box, int32_t* out_host);

/* Crop window of every pair: combine_bbox (utils/data_utils.py:61-72) + inference.py:362-365.
 * boxes_host[4 N] xywh float64 (integer boxes are passed as their float64 image), pairs_host[2 P],
 * out_host[4 P] = (x, y, S, S) with python int() truncation.  Returns IO_ERR_DEGENERATE if some S <= 0
 * (out is still filled so the caller can see which). */
IO_API int io_pair_crop_boxes(const double* boxes_host, const int32_t* pairs_host, int p, int32_t* out_host);

/* bordering(), inference.py:691-696, for P candidate pairs of one image:
 * flags_dev[k] = any((dilate(mask[i], 3x3 cross) == 1) & mask[j]).  masks_dev is [N, H, W] u8. */
IO_API int io_pair_bordering(const uint8_t* masks_dev, int n, int h, int w, const int32_t* pairs_dev, int p,
                      uint8_t* flags_dev, void* stream);

/* infer_gt_order(), inference.py:719-739 (KINS has no order annotation; tools/test.py:129,417-418 derives it from the
 * modal / amodal masks): for every pair (i < j) that borders, compares |modal_i & amodal_j| with |modal_j & amodal_i|.
 * modal_dev / amodal_dev: [N, H, W] u8; mat_dev: [N, N] int64, zero-initialised by the caller. */
IO_API int io_infer_gt_order(const uint8_t* modal_dev, const uint8_t* amodal_dev, int n, int h, int w,
                             const int32_t* pairs_dev, int p, int64_t* mat_dev, void* stream);

/* Per-instance statistics behind the heuristic baselines infer_occ_order_area / _yaxis and infer_depth_order_area /
 * _yaxis (inference.py:272-346): out_dev[i] = (sum of mask values, #pixels == 1, sum of the row index over pixels == 1)
 * as exact int64; the N x N comparison (with io_pair_bordering for the occlusion variants) is host logic. */
IO_API int io_mask_stats(const uint8_t* masks_dev, int n, int h, int w, int64_t* out_dev, void* stream);

/* One record per pair for the fused gather.  Offsets are in bytes from the base pointers passed to the call. */
typedef struct io_pair_desc {
  int64_t image_off;  /* start of the pair's H x W x 3 u8 image                         */
  int64_t mask_a_off; /* start of instance i's H x W u8 modal mask                       */
  int64_t mask_b_off; /* start of instance j's H x W u8 modal mask                       */
  int32_t h, w;       /* image size                                                      */
  int32_t x, y, s;    /* crop window (io_pair_crop_boxes); ignored in resize mode        */
  int32_t rgb_slot;   /* resize mode: index of the image's pre-resized rgb plane;
                         patch mode: flags, bit 0 = horizontal flip of the resized crop
                         (training augmentation, datasets/depth_occ_order_dataset.py:166-181) */
} io_pair_desc;

/* Geometry of the gather output ("pair tensor"): [P, D + 6, row_pitch, 8] bf16, NHWC with 3 zero pixels of border
 * on every side (conv1's padding, materialised) and channels (maskA, maskB, R, G, B, 0, 0, 0).  The interior
 * [3:3+D, 3:3+D, 0:5] equals bf16(torch.cat([modal_i, modal_j, transform_rgb(rgb)], 1)) of inference.py:141-145;
 * the (B, A) direction is the same tensor with channels 0/1 exchanged and is never materialised. */
IO_API int64_t io_pair_tensor_row_pitch(int d);          /* pixels per padded row (multiple of 8, >= d + 6)          */
IO_API int64_t io_pair_tensor_bytes(int p, int d);       /* bytes of a [p] pair tensor                                */

/* Fused `patch` mode pair construction (inference.py:360-375 + utils/data_utils.py:28-34,104-124 + :141-145):
 * crop_padding + cv2.INTER_CUBIC (OpenCV generic 8-bit path, bit-exact) of the rgb crop + cv2.INTER_NEAREST of both
 * modal masks + /255, (x - mean) / std + bf16, written once into the pair tensor.
 * images_dev / masks_dev: base pointers of the packed u8 images / masks; descs_dev[P]; mean/std: 3 floats (host). */
IO_API int io_pair_gather_patch(const uint8_t* images_dev, const uint8_t* masks_dev, const io_pair_desc* descs_dev, int p,
                         int d, const float* mean_host, const float* std_host, void* out_dev, void* stream);

/* `resize` mode (inference.py:395-399, utils/data_utils.py:37-53, midas/transforms.py:48-235):
 * step 1, once per image: image/255 -> cubic (float taps, no u8 rounding) to d x d -> normalise -> fp32 plane
 * rgb_planes_dev[slot][d][d][3]; step 2, per pair: nearest-resized masks + the image's plane -> pair tensor. */
IO_API int io_image_resize_rgb(const uint8_t* image_dev, int h, int w, int d, const float* mean_host, const float* std_host,
                        float* rgb_plane_dev, void* stream);
/* `image` mode rgb (inference.py:377-393): zero-pad to the centred max(h, w) square, cv2.INTER_LINEAR (8-bit generic
 * path, bit-exact) to d x d, transform_rgb -> fp32 plane [d][d][3].  lut_scratch_dev: 768 floats of device scratch.
 * The pairs are then assembled by io_pair_gather_resize with desc.s = max(h, w), desc.x = (s - w) / 2,
 * desc.y = (s - h) / 2 (nearest over the padded square; s = 0 selects the plain `resize` mode). */
IO_API int io_image_square_linear_rgb(const uint8_t* image_dev, int h, int w, int d, const float* mean_host,
                                      const float* std_host, float* lut_scratch_dev, float* rgb_plane_dev,
                                      void* stream);
/* Training `resize` mode rgb (datasets/depth_occ_order_dataset.py:83-86): cv2.resize(image_u8, (d, d), INTER_LINEAR)
 * (8-bit generic path, bit-exact) + Normalize -> fp32 plane [d][d][3]; pairs are assembled by io_pair_gather_resize
 * with desc.s = 0.  In both resize-type modes bit 30 of desc.rgb_slot requests the horizontal flip of the training
 * augmentation (:91-94, :122-129). */
IO_API int io_image_resize_linear_rgb(const uint8_t* image_dev, int h, int w, int d, const float* mean_host,
                                      const float* std_host, float* lut_scratch_dev, float* rgb_plane_dev,
                                      void* stream);
IO_API int io_pair_gather_resize(const float* rgb_planes_dev, const uint8_t* masks_dev, const io_pair_desc* descs_dev,
                          int p, int d, void* out_dev, void* stream);

/* Collated dataset tensors (rgb [B,3,D,D], modal1 / modal2 [B,1,D,D], fp32 NCHW, device) -> pair tensor: the
 * torch.cat([modal1, modal2, rgb], 1) of models/supervised_order.py:52 for validation / arbitrary `model.model(x)`. */
IO_API int io_pair_pack_nchw(const float* rgb_dev, const float* modal1_dev, const float* modal2_dev, int b, int d,
                             void* out_dev, void* stream);

/* (u8 / 255 - mean) / std in fp32 exactly as torchvision computes it; out_host[3 * 256].  Host only. */
IO_API int io_normalize_lut(const float* mean_host, const float* std_host, float* out_host);

/* ------------------------------------------------------------------------------------------------------------ */
/* N -- network: 5-channel ResNet-50 classifier, models/backbone/resnet_cls.py:75-222                            */
/* ------------------------------------------------------------------------------------------------------------ */

typedef struct io_net io_net_t;

/* n_heads = 1 (`fc`, num_classes int) or 2 (`fc_occ`, `fc_depth`; num_classes list) -- resnet_cls.py:153-160.
 * input_size: 256 or 384 (any multiple of 32); max_pairs: largest P of a forward call (workspace is sized once). */
IO_API int io_net_create(const int32_t* num_classes, int n_heads, int input_size, int max_pairs, io_net_t** out);
/* Generalised handle for the bottleneck-ResNet family (InstaDepthNet^od, reference midas/midas_net.py:113-212): per layer the
 * 3x3 width, the block output channels and the block count (ResNet-50: {64,128,256,512} / {256,..,2048} / {3,4,6,3};
 * the ResNeXt-101 32x8d encoder: {256,512,1024,2048} / same / {3,4,23,3} with its grouped 3x3 weights passed to
 * io_net_load_state as dense block-diagonal [cout, cin, 3, 3] tensors).  n_layers = 3 stops after layer3; keep_layers = 1
 * keeps every layer's output (io_net_feature); n_heads = 0 makes a feature extractor (logits may be NULL in
 * io_net_forward_pairs).  The stem is always the 5-channel pair-tensor stem: embed 3-channel (RGB -> channels 2..4) or
 * 2-channel (masks -> channels 0..1) conv1 weights with zeros. */
IO_API int io_net_create_arch(const int32_t* widths4, const int32_t* outs4, const int32_t* blocks4, int n_layers,
                       int keep_layers, const int32_t* num_classes, int n_heads, int input_size, int max_pairs,
                       io_net_t** out);
/* Layer `layer` (0-based) output of a keep_layers handle: bf16 [images, D >> (2 + layer), same, outs[layer]]; a head-less
 * handle (RGB-only encoder) computes ONE direction: images = the `p` of io_net_forward_pairs, else 2 * p. */
IO_API int io_net_feature(io_net_t* net, int layer, void** ptr, int64_t* elems_per_image);
/* Trunks of InstaDepthNet (midas_net.py:200-210): after layers 1, 2, 3 add f_l[idx_dev[image]] (image = 2 * pair +
 * direction; idx_dev is read at run time, int32 [2 * max_pairs]).  All NULL switches the injection off. */
IO_API int io_net_set_inject(io_net_t* net, const void* f1, const void* f2, const void* f3, const int32_t* idx_dev);
IO_API int io_net_destroy(io_net_t* net);

/* Loads a reference checkpoint's state_dict (models/single_stage_model.py:54-61, utils/common_utils.py:128-149):
 * names[k] are the reference keys *without* the `module.` prefix ("conv1.weight", "layer1.0.bn1.running_var",
 * "fc_occ.bias", ...), ptrs_host[k] the fp32 tensors in the reference's own layout (conv: [Cout, Cin, kh, kw]).
 * Eval-mode BatchNorm is folded into the bf16 GEMM weights (scale) and an fp32 bias; missing keys are an error
 * (num_batches_tracked is ignored). */
IO_API int io_net_load_state(io_net_t* net, const char* const* names, const float* const* ptrs_host,
                      const int64_t* numels, int n);

/* Eval-mode forward of both directions of P pairs from a pair tensor (io_pair_gather_*):
 * logits_dev[P][2][K] fp32 with K = sum(num_classes); [p][0] = f(A, B), [p][1] = f(B, A) (inference.py:144-145). */
IO_API int io_net_forward_pairs(io_net_t* net, const void* pair_tensor_dev, int p, float* logits_dev, void* stream);

/* number of kernel launches the last io_net_forward_pairs issued (bench.py's gpu_launches) */
IO_API int io_net_last_launches(const io_net_t* net);

/* `orig` mode (reference inference.py:401-408, utils/data_utils.py:13-17): the network input is the image at its own
 * size rounded to the nearest multiples of 32 -- dh x dw, not square.  io_image_resize_rgb_hw / io_pair_gather_resize_hw /
 * io_pair_tensor_bytes_hw are the `resize`-mode entry points with separate height and width (plane [dh][dw][3], pair
 * tensor [p][dh + 6][pitch(dw)][8]); io_net_forward_pairs_hw runs the pairs of one such image through a handle created
 * with input_size >= max(dh, dw) (general tilers, one launch per convolution; plans are rebuilt when the geometry changes). */
IO_API int64_t io_pair_tensor_bytes_hw(int p, int dh, int dw);
IO_API int io_image_resize_rgb_hw(const uint8_t* image_dev, int h, int w, int dh, int dw, const float* mean_host,
                                  const float* std_host, float* rgb_plane_dev, void* stream);
IO_API int io_pair_gather_resize_hw(const float* rgb_planes_dev, const uint8_t* masks_dev, const io_pair_desc* descs_dev,
                                    int p, int dh, int dw, void* out_dev, void* stream);
IO_API int io_net_forward_pairs_hw(io_net_t* net, const void* pair_tensor_dev, int p, int h, int w, float* logits_dev,
                                   void* stream);

/* Optional per-kernel timing for bench.py's roofline: when enabled, io_net_forward_pairs brackets every launch
 * with CUDA events on the launching stream.  io_net_profile_read (after the caller synchronised the stream)
 * returns the number of launches of the last forward and fills ms_host[i] (duration), kind_host[i]
 * (0 stem conv, 1 max-pool, 2 bottleneck conv, 3 pool+FC tail), flop_host[i] (algorithmic 2*MAC of the launch:
 * real taps / channels only, no padding) and, if non-null, bytes_host[i] (algorithmic HBM bytes) and tag_host[i]
 * (layer * 100 + block * 10 + conv index; 1 = stem, 2 = max-pool) for i < max_n. */
IO_API int io_net_profile(io_net_t* net, int enable);
IO_API int io_net_profile_read(io_net_t* net, float* ms_host, int32_t* kind_host, double* flop_host,
                               double* bytes_host, int32_t* tag_host, int max_n);

/* H1-H5: probabilities, direction average, decision and scatter into the per-image order matrices.
 * logits_dev[P][2][K]; head_kind/head_off/head_k select the columns of one head.
 * pair_ij_dev[2 P]; mat_off_dev[P] = element offset of the pair's image matrix inside mat_dev (int64, row stride
 * mat_n_dev[P]).  Matrices must be zero-initialised: like the reference, only decided entries are written
 * (occ: ones; depth: both (i,j) and (j,i)) -- inference.py:416-434, 507-510, 612-623.
 * margin_dev[P] (optional): decision margin, min |p - 0.5| (occ) or top1 - top2 (softmax heads). */
IO_API int io_order_decide(const float* logits_dev, int p, int k_total, int head_kind, int head_off, int head_k,
                    const int32_t* pair_ij_dev, const int64_t* mat_off_dev, const int32_t* mat_n_dev,
                    int64_t* mat_dev, float* margin_dev, void* stream);

/* ---- annotation -> modal masks (replaces pycocotools' frPyObjects / merge / decode as called by
 * reference datasets/reader.py:20-66 read_KINS / read_LVIS / read_COCOA) ------------------------------------------
 * Host: run lengths (column-major, starting with zeros) of one RLE part.  io_rle_from_string decodes the compressed
 * ASCII `counts` of an RLE dict; io_rle_from_polygon rasterises one polygon [x0, y0, x1, y1, ...] (k points) of an
 * h x w image.  counts[max_counts]; *n_out = number of runs. */
IO_API int io_rle_from_string(const char* s, int64_t len, uint32_t* counts, int max_counts, int* n_out);
IO_API int io_rle_from_polygon(const double* xy, int k, int h, int w, uint32_t* counts, int max_counts, int* n_out);
/* Device: out_dev[n_inst][h][w] uint8 {0,1}, instance i = the union of its parts comp in [inst_off[i], inst_off[i+1]);
 * part c owns the inclusive prefix sums cum_dev[comp_off[c] .. comp_off[c+1]) of its run lengths. */
IO_API int io_masks_from_rle(const uint32_t* cum_dev, const int32_t* comp_off_dev, const int32_t* inst_off_dev, int n_inst,
                      int h, int w, uint8_t* out_dev, void* stream);

/* ---- element-wise pieces of InstaDepthNet's MiDaS decoder (reference midas/blocks.py:124-195, midas_net.py:126-140;
 * its 3x3 convolutions go through io_conv_bn_act) ------------------------------------------------------------------
 * out = a + b (ReLU'd if relu != 0) over n bf16 elements, n % 8 == 0 (ResidualConvUnit's in-place ReLU makes every
 * consumer of a sum read relu(sum)). */
IO_API int io_add_relu(const void* a_dev, const void* b_dev, void* out_dev, int64_t n, int relu, void* stream);
/* nn.functional.interpolate(scale_factor=2, mode="bilinear", align_corners=...) on NHWC bf16 [b,h,w,c] -> [b,2h,2w,c]. */
IO_API int io_upsample2x_bilinear(const void* x_dev, int b, int h, int w, int c, int align_corners, void* y_dev, void* stream);

/* conv1 + bn1 + ReLU + MaxPool2d(3, 2, 1) (reference models/backbone/resnet_cls.py:140-146, 205-208) of both
 * directions of every pair as ONE kernel, for 256 x 256 inputs -- the first launch of io_net_forward_pairs, exposed for
 * the layer-level parity test.  w_host: conv1 weights [64][5][7][7] fp32 with the BN scale folded in, bias_host: [64]
 * (host pointers); out_dev: [2 * pairs][64][64][64] bf16 NHWC, image 2 * pair + direction.  Synchronises the stream. */
IO_API int io_stem_pool(const void* pair_tensor_dev, int pairs, int d, const float* w_host, const float* bias_host,
                        void* out_dev, void* stream);

/* Single convolution + folded BN (+ residual) (+ ReLU) on NHWC bf16, the building block of io_net_forward_pairs,
 * exported for the per-layer parity tests.  w_dev: [Cout][kh*kw*Cin] bf16 (tap-major, channel-minor);
 * kernel 1 or 3, stride 1 or 2, padding = kernel / 2; Cin, Cout multiples of 64; output rows wider than 128 pixels
 * (MiDaS decoder: 192, 384) are supported for 3x3 stride 1. */
IO_API int io_conv_bn_act(const void* x_dev, int b, int h, int w, int cin, const void* w_dev, const float* bias_dev,
                   const void* residual_dev, int cout, int kernel, int stride, int relu, void* y_dev, void* stream);

/* First bottleneck of a ResNet layer (resnet_cls.py:107-116 with the downsample branch :112-113): y = ReLU(conv3(t2) +
 * downsample(x)) as ONE GEMM over the concatenated K = [cmid channels of t2 | cin channels of x]; x_dev [b][h][w][cin],
 * t2_dev [b*(h/stride)*(w/stride)][cmid], wcat_dev [cout][cmid + cin] (conv3 columns first, BN folded), bias_dev =
 * sum of the two folded-BN shifts; stride 1 or 2.  The identity tensor never exists in HBM.  Exported for the parity
 * test; io_net_forward_pairs uses it for layerN.0 (INSTAORDER_FUSE_DS=0 falls back to two launches). */
IO_API int io_conv_dual(const void* x_dev, int b, int h, int w, int cin, int stride, const void* t2_dev, int cmid,
                 const void* wcat_dev, const float* bias_dev, int cout, int relu, void* y_dev, void* stream);

/* conv3 (1x1, cmid -> 4*cmid, + residual + ReLU -> y) of one bottleneck fused with conv1 (1x1, 4*cmid -> n2, + ReLU
 * -> y2) of the next one (resnet_cls.py:107-116 then :99-101): the block output is written once and consumed from
 * shared memory by the second GEMM.  x_dev: [rows][cmid], w3_dev: [4*cmid][cmid], w1n_dev: [n2][4*cmid], all bf16;
 * cmid in {64, 128, 256}, n2 in {64, 128, 256}.  Exported for the parity test. */
IO_API int io_conv_fused_pair(const void* x_dev, int rows, int cmid, const void* w3_dev, const float* bias3_dev,
                              const void* residual_dev, void* y_dev, const void* w1n_dev, const float* bias1n_dev,
                              int n2, void* y2_dev, void* stream);

/* The same fusion for the FIRST bottleneck of a layer: y = ReLU([t2 | x(stride)] * wcat^T + bias) (io_conv_dual) and
 * y2 = ReLU(y * w1n^T + bias1n), the next block's conv1, computed from the block-output tile while it is in shared
 * memory.  cmid in {64, 128}; n2 in {64, 128, 256}; stride 2 needs M tiles of exactly 128 output pixels. */
IO_API int io_conv_fused_dual(const void* x_dev, int b, int h, int w, int cin, int stride, const void* t2_dev, int cmid,
                       const void* wcat_dev, const float* bias_dev, void* y_dev, const void* w1n_dev,
                       const float* bias1n_dev, int n2, void* y2_dev, void* stream);

/* Validation losses (forward only) -- models/supervised_order.py:60-81 (^od), :397-411 (^d), :465-479 (OrderNet),
 * :518-533 (^o), including the swapped-direction labels of set_input and the reference's softmax-then-CrossEntropy
 * / sigmoid-then-BCELoss quirks.  logits_dev[n][2][k_total]; occ_off >= 0 selects a 2-logit sigmoid head with
 * occ_target_dev[n][2] fp32; class_off >= 0 selects a class_k-way softmax head with class_target_dev[n] int64.
 * is_overlap_dev (optional, int64[n]) switches on the overlap / distinct weighting of ^od.  out_dev[3] fp32 =
 * (loss / world_size, occlusion loss, class loss). */
IO_API int io_loss_forward(const float* logits_dev, int n, int k_total, int occ_off, int class_off, int class_k,
                           const float* occ_target_dev, const int64_t* class_target_dev,
                           const int64_t* is_overlap_dev, float overlap_w, float distinct_w, int world_size,
                           float* out_dev, void* stream);

/* ------------------------------------------------------------------------------------------------------------ */
/* T -- training step: models/supervised_order.py:83-95 (^od), :413-438 (^d), :481-493 (OrderNet), :535-548 (^o)   */
/* ------------------------------------------------------------------------------------------------------------ */

typedef struct io_train io_train_t;

/* One handle = one model replica with a fixed per-rank batch of `batch_pairs` pairs at `input_size`^2 (workspace for
 * all saved activations is allocated by io_train_bind: ~120 MB per pair at 256^2).  num_classes / n_heads as
 * io_net_create. */
IO_API int io_train_create(const int32_t* num_classes, int n_heads, int input_size, int batch_pairs, io_train_t** out);
IO_API int io_train_destroy(io_train_t* t);

/* Flat buffers.  Parameters, gradients and optimiser state are flat fp32 arrays of io_train_param_count() elements
 * owned by the CALLER (torch tensors); BatchNorm running_mean / running_var live in a second flat array of
 * io_train_stat_count() elements.  The segment table maps them to the reference's state_dict names (without the
 * `module.` prefix): buffer_out = 0 (parameter buffer) or 1 (statistics buffer), offset_out in elements, and
 * dims4_out = {cout, kh, kw, cin} for a convolution weight -- stored tap-major / channel-minor, i.e. the reference
 * tensor [cout, cin, kh, kw] permuted (0, 2, 3, 1) -- or {n, 0, 0, 0} / {rows, cols, 0, 0} for tensors kept in
 * the reference's own layout (BN weight / bias / running stats, FC weight / bias).  Segments are 64-element
 * aligned; the gaps stay zero.  Because the optimisers are element-wise, momentum / Adam buffers use the same table
 * (utils/common_utils.py:128-149 resume path). */
IO_API int64_t io_train_param_count(const io_train_t* t);
IO_API int64_t io_train_stat_count(const io_train_t* t);
IO_API int io_train_num_segments(const io_train_t* t);
IO_API int io_train_segment(const io_train_t* t, int i, char* name_out, int name_cap, int32_t* buffer_out,
                            int64_t* offset_out, int32_t* dims4_out);
IO_API int io_train_bind(io_train_t* t, float* params_dev, float* grads_dev, float* stats_dev);
/* Refreshes the bf16 GEMM copies of the weights from the fp32 masters (after loading / editing params_dev). */
IO_API int io_train_sync_weights(io_train_t* t, void* stream);

/* forward (both directions, train-mode BatchNorm: batch statistics per direction, running statistics updated once
 * per direction as the reference's two forward passes do) + loss (+ backward into grads_dev, which is zeroed first).
 * pair_tensor_dev: [batch_pairs] pair tensor (io_pair_pack_nchw / io_pair_gather_*).  Heads and targets as
 * io_loss_forward (the swapped-direction labels of set_input are derived inside).  out_losses_dev[3] fp32 =
 * (loss / world_size, occlusion loss, class loss).  The gradient all-reduce (utils/distributed_utils.py:27-31) is
 * the caller's: ONE all-reduce of grads_dev. */
IO_API int io_train_forward_backward(io_train_t* t, const void* pair_tensor_dev, int occ_off, int class_off, int class_k,
                                     const float* occ_target_dev, const int64_t* class_target_dev,
                                     const int64_t* is_overlap_dev, float overlap_w, float distinct_w, int world_size,
                                     float* out_losses_dev, int run_backward, void* stream);
/* Gradient buckets for the data-parallel all-reduce (reference utils/distributed_utils.py:27-31: one blocking
 * all_reduce per parameter AFTER backward).  The flat gradient buffer becomes final in io_train_num_buckets() contiguous
 * ranges [begin, end) (elements), in this order: (layer4 + heads), layer3, layer2, (stem + layer1).
 * io_train_wait_bucket makes `stream` wait (device side) for bucket k of the last io_train_forward_backward, so the
 * caller can all-reduce that range on a communication stream while the rest of the backward pass still runs. */
IO_API int io_train_num_buckets(const io_train_t* t);
IO_API int io_train_bucket(const io_train_t* t, int k, int64_t* begin, int64_t* end);
IO_API int io_train_wait_bucket(io_train_t* t, int k, void* stream);

/* torch.optim.SGD(lr, momentum, weight_decay) / torch.optim.Adam(lr, betas, eps) over the bound flat buffers
 * (models/single_stage_model.py:34-42), fused with the refresh of the bf16 GEMM weights. */
IO_API int io_train_sgd_step(io_train_t* t, float* momentum_buf_dev, float lr, float momentum, float weight_decay,
                             int first_step, void* stream);
IO_API int io_train_adam_step(io_train_t* t, float* m_dev, float* v_dev, float lr, float beta1, float beta2, float eps,
                              int step, void* stream);
/* logits of the last forward, copied (device to device) into out_dev[2][batch_pairs][K] fp32 ([direction][pair]) */
IO_API int io_train_read_logits(const io_train_t* t, float* out_dev, void* stream);
/* Saved tensors of the last forward (layer-by-layer parity tests): conv_name = reference module name ("conv1",
 * "layer3.2.conv2", "layer2.0.downsample.0"), which = 0: raw convolution output, 1: activation after BN (+ residual)
 * (+ ReLU); bf16 [2 * batch_pairs, h, w, c] in [direction][pair] image order.  numel_out receives the element count
 * (out_dev may be NULL to query it). */
IO_API int io_train_read_activation(const io_train_t* t, const char* conv_name, int which, void* out_dev,
                                    int64_t* numel_out, void* stream);
IO_API int io_train_last_launches(const io_train_t* t);
/* per-launch CUDA-event timing as io_net_profile; kind: 0 conv forward, 1 data gradient, 2 weight gradient,
 * 3 element-wise / reduction, 4 loss */
IO_API int io_train_profile(io_train_t* t, int enable);
IO_API int io_train_profile_read(io_train_t* t, float* ms_host, int32_t* kind_host, double* flop_host,
                                 double* bytes_host, int32_t* tag_host, int max_n);

/* The same optimiser kernels on arbitrary flat buffers (w_bf16_dev optional: first n_bf16 elements re-cast). */
IO_API int io_optim_sgd(float* w_dev, const float* g_dev, float* buf_dev, int64_t n, float lr, float momentum,
                        float weight_decay, int first_step, void* w_bf16_dev, int64_t n_bf16, void* stream);
IO_API int io_optim_adam(float* w_dev, const float* g_dev, float* m_dev, float* v_dev, int64_t n, float lr, float beta1,
                         float beta2, float eps, int step, void* w_bf16_dev, int64_t n_bf16, void* stream);

/* Building blocks of the training step, exported for the per-kernel parity tests (NHWC bf16 tensors).
 * io_conv_wgrad: dw_dev[cout][k*k*cin] fp32 += weight gradient of conv(x[b,h,w,cin]; k, stride, pad k/2) given
 *   dy[b,h/stride,w/stride,cout].  io_conv_dgrad: dx[b,h,w,cin] = data gradient of the stride-1 convolution given
 *   dy[b,h,w,cout] and the forward weights w_dev[cout][k*k*cin] bf16 (+ residual).  io_stem_wgrad: packed
 *   two-direction stem gradient [128][448] fp32 from the pair tensor and dy[2][pairs][d/2][d/2][64]. */
IO_API int io_conv_wgrad(const void* x_dev, int b, int h, int w, int cin, const void* dy_dev, int cout, int kernel,
                         int stride, float* dw_dev, void* stream);
IO_API int io_conv_dgrad(const void* dy_dev, int b, int h, int w, int cin, int cout, int kernel, const void* w_dev,
                         const float* zero_bias_dev, const void* residual_dev, void* dx_dev, void* stream);
IO_API int io_stem_wgrad(const void* pair_tensor_dev, int pairs, int d, const void* dy_dev, float* dw_scratch_dev,
                         void* stream);
/* train-mode BatchNorm over [groups][rows][c] (+ residual) (+ ReLU) and its backward; save_dev = 4 x [groups][c] fp32
 * (scale, shift, mean, invstd), scratch_dev = [groups][2][c] doubles.  Backward mask_mode: 0 = no ReLU, 1 = ReLU mask
 * from the stored activation a_dev (needed when a residual was added), 2 = mask recomputed from y_dev (a_dev unused),
 * 3 = a_dev points to the bit mask written by io_bn_train_forward's mask_out_dev (one byte per 8 channels; optional). */
IO_API int io_bn_train_forward(const void* y_dev, const void* residual_dev, void* a_dev, int groups, int rows, int c,
                               const float* gamma_dev, const float* beta_dev, float eps, float momentum,
                               float* running_mean_dev, float* running_var_dev, float* save_dev, double* scratch_dev,
                               int relu, uint8_t* mask_out_dev, void* stream);
IO_API int io_bn_train_backward(const void* da_dev, const void* a_dev, const void* y_dev, void* dy_dev, void* g_out_dev,
                                int groups, int rows, int c, const float* gamma_dev, const float* save_dev,
                                double* scratch_dev, int mask_mode, float* dgamma_dev, float* dbeta_dev, void* stream);
/* nn.MaxPool2d(3, 2, 1) with arg-max (idx_dev: one byte per output element) and, if dy_dev / dx_dev are given, its
 * backward */
IO_API int io_maxpool_train(const void* x_dev, void* y_dev, uint8_t* idx_dev, const void* dy_dev, void* dx_dev, int b,
                            int h, int w, int c, void* stream);

/* ------------------------------------------------------------------------------------------------------------ */
/* M -- metrics, batched over images                                                                             */
/* ------------------------------------------------------------------------------------------------------------ */

/* eval_order_recall_precision_f1, inference.py:794-802 (sklearn binary scores over entries with gt != -1, x100,
 * zero_division = zd).  order_dev / gt_dev: packed int64 matrices, image b at mat_off_host-style offsets
 * off_dev[b] with side n_dev[b].  out_dev[b][3] = (recall, precision, f1) float64, bit-identical to sklearn.
 * An image with no entry gt != -1 yields NaN (the reference raises). */
IO_API int io_metrics_prf(const int64_t* order_dev, const int64_t* gt_dev, const int64_t* off_dev, const int32_t* n_dev,
                   int batch, int zd, double* out_dev, void* stream);

/* eval_depth_order_whdr, inference.py:757-791: 9 WHDR variants over the strict upper triangle; out_dev[b][9] in
 * the order ovlX_eq, ovlX_neq, ovlX_all, ovlO_eq, ..., ovlOX_all; -1 where the selection mask is empty.
 * Sums follow numpy's pairwise summation order so results are bit-identical to the reference. */
IO_API int io_metrics_whdr(const int64_t* order_dev, const int64_t* gt_order_dev, const int64_t* gt_overlap_dev,
                    const int64_t* gt_count_dev, const int64_t* off_dev, const int32_t* n_dev, int batch,
                    double* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* INSTAORDER_B200_H_ */
