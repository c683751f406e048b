// io_train_t: one training step (forward in train mode + loss + backward) of the 5-channel ResNet-50 order
// classifiers as a static list of launches over bf16 NHWC activations resident in HBM.
//
// Replaces InstaOrderNet_od / _d / _o / OrderNet `.step()` up to (not including) the gradient all-reduce and the
// optimiser update (reference models/supervised_order.py:83-95, 413-438, 481-493, 535-548): two forward passes
// (A,B) and (B,A) in train-mode BatchNorm -- two separate BN batches, running statistics updated twice --, the loss
// of :60-81, and `loss.backward()`.
//
// Both directions run through every launch together as image groups [direction][pair]; only the BatchNorm statistics
// are per group.  Convolutions (forward and data gradient) are the tcgen05 implicit-GEMM kernels of conv_tc.cu /
// conv_tn.cu with un-folded bf16 weights; the data gradient reads the SAME weight matrix as an MN-major operand with
// mirrored taps; weight gradients are wgrad.cu.  fp32 master parameters, gradients and optimiser state live in flat
// caller-owned buffers (GEMM layout [cout][kh][kw][cin] per convolution; the segment table maps them to the
// reference's state_dict names), so the data-parallel all-reduce (utils/distributed_utils.py:27-31) is ONE NCCL
// call on the flat gradient buffer.
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "train.cuh"
#include "tail.cuh"

namespace io {

int loss_train_launch(const float* logits, int n, int k_total, int occ_off, int cls_off, int cls_k,
                      const float* occ_target, const int64_t* class_target, const int64_t* is_overlap,
                      float overlap_w, float distinct_w, int world_size, float* out, float* dlogits,
                      cudaStream_t stream);
int cast_bf16_launch(const float* w, void* w16, int64_t n, cudaStream_t stream);
int stem_pack_launch(const float* w, void* pk, cudaStream_t stream);
int stem_unpack_grad_launch(const float* scratch, float* dw, cudaStream_t stream);

struct Segment {
  std::string name;
  int buffer;        // 0 = parameter / gradient flat buffer, 1 = statistics (running_mean / running_var) buffer
  int64_t offset;
  int dims[4];       // conv weight: {cout, kh, kw, cin} (GEMM layout; reference layout is [cout, cin, kh, kw]);
                     // everything else: {n, 0, 0, 0} or {rows, cols, 0, 0} in the reference's own layout
};

struct Unit {        // one convolution + BatchNorm
  std::string conv, bn;
  int cin, cout, k, stride;
  int h_in, w_in, h_out, w_out;
  int64_t w_off, g_off, b_off;   // parameter offsets (weight, BN gamma, BN beta)
  int64_t rm_off, rv_off;        // statistics offsets
  __nv_bfloat16* y = nullptr;    // raw convolution output [imgs, h_out, w_out, cout]
  __nv_bfloat16* a = nullptr;    // activation after BN (+ residual) (+ ReLU)
  float* save = nullptr;         // [4][2][cout]: scale, shift, mean, invstd of the last forward (per direction)
  uint8_t* mask = nullptr;       // residual units: ReLU bit mask of `a` (1 byte per 8 channels), read by the backward
};

struct TOp {
  std::function<int(cudaStream_t)> run;
  int kind;        // 0 conv fwd, 1 dgrad, 2 wgrad, 3 element-wise / reduction, 4 other
  double flops, bytes;
  int tag;
  // side-stream scheduling of the weight gradients (they feed nothing but the gradient buffer): a `side` op reads the
  // dy buffer `reads`; a main-stream op that overwrites a buffer (`writes`) first waits for its last side reader
  int side = 0;
  const void* reads = nullptr;
  const void* writes = nullptr;
};

}  // namespace io

struct io_train {
  int n_heads = 0;
  int num_classes[2] = {0, 0};
  int k_total = 0;
  int d = 0;
  int pairs = 0, imgs = 0;
  std::vector<io::Unit> units;          // [0] = stem, then execution order (conv1, conv2, [downsample], conv3)
  std::vector<io::Segment> segs;
  int64_t n_params = 0, n_stats = 0;
  int64_t fc_w_off = 0, fc_b_off = 0;
  // caller-owned
  float* params = nullptr;
  float* grads = nullptr;
  float* stats = nullptr;
  // library-owned
  __nv_bfloat16* w16 = nullptr;         // bf16 copy of the flat parameter buffer (same offsets)
  __nv_bfloat16* stem_pk = nullptr;     // [128][448] packed two-direction stem weights
  float* stem_scratch = nullptr;        // [128][448] stem wgrad accumulator
  float* zero_bias = nullptr;           // [2048] zeros
  double* red[2] = {nullptr, nullptr};  // two [2][2][2048] reduction scratch buffers used alternately: the apply
                                        // kernel of BatchNorm i clears the buffer BatchNorm i + 1 accumulates into
  int red_sel = 0;                      // build-time cursor
  uint8_t* pool_idx = nullptr;
  __nv_bfloat16* pool_out = nullptr;
  float* pooled = nullptr;              // [imgs][2048]
  float* logits = nullptr;              // [imgs][k]
  float* dlogits = nullptr;
  __nv_bfloat16* gbuf[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // GA, GG, GY, GT, GD, GZ
  __nv_bfloat16* dy_ring[3] = {nullptr, nullptr, nullptr};   // dy buffers (gbuf[2] + 2 more) used in rotation, so a
  int dy_cursor = 0;                                         // weight gradient may lag two units behind the main chain
  unsigned int* barriers = nullptr;     // [256] grid-barrier counters of the fused BatchNorm kernels (zeroed per step)
  int barrier_cursor = 0;
  bool fused_bn = true;
  cudaStream_t side_stream = nullptr;
  std::vector<cudaEvent_t> side_ev;     // fork / done event per side op
  bool use_side = true;
  std::vector<void*> allocs;
  std::vector<io::TOp> fwd, bwd;
  // Gradient buckets for the data-parallel all-reduce (reference utils/distributed_utils.py:27-31): backward finishes the
  // parameters back to front, so the flat gradient buffer is complete in four contiguous ranges -- (layer4 + heads),
  // layer3, layer2, (stem + layer1) -- at known points of the launch list.  An event pair (main / weight-gradient
  // stream) is recorded there; the caller's communication stream waits on it (io_train_wait_bucket) and all-reduces
  // that range while the rest of the backward pass is still running.
  static constexpr int kBuckets = 4;
  int64_t bucket_begin[kBuckets] = {0, 0, 0, 0}, bucket_end[kBuckets] = {0, 0, 0, 0};
  size_t bucket_after_op[kBuckets] = {0, 0, 0, 0};     // bucket k is complete after bwd op index bucket_after_op[k] - 1
  cudaEvent_t bucket_ev[kBuckets][2] = {};
  bool bucket_side[kBuckets] = {false, false, false, false};
  bool built = false;
  bool weights_synced = false;
  // per-step inputs referenced by the op closures
  const void* pair_tensor = nullptr;
  int occ_off = -1, cls_off = -1, cls_k = 0;
  const float* occ_target = nullptr;
  const int64_t* class_target = nullptr;
  const int64_t* is_overlap = nullptr;
  float overlap_w = 1.f, distinct_w = 1.f;
  int world_size = 1;
  float* out_losses = nullptr;
  int last_launches = 0;
  // profiling
  bool profile = false;
  std::vector<cudaEvent_t> ev;
  std::vector<int> prof_kind, prof_tag;
  std::vector<double> prof_flops, prof_bytes;
};

namespace io {

static int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

static void build_units(io_train* t) {
  auto add_seg = [&](const std::string& name, int buffer, int64_t numel, int d0, int d1, int d2, int d3) -> int64_t {
    int64_t& n = buffer == 0 ? t->n_params : t->n_stats;
    const int64_t off = n;
    t->segs.push_back(Segment{name, buffer, off, {d0, d1, d2, d3}});
    n = align_up(n + numel, 64);
    return off;
  };
  auto add_unit = [&](const std::string& conv, const std::string& bn, int cin, int cout, int k, int stride, int h_in,
                      int w_in) {
    Unit u;
    u.conv = conv; u.bn = bn; u.cin = cin; u.cout = cout; u.k = k; u.stride = stride;
    u.h_in = h_in; u.w_in = w_in; u.h_out = h_in / stride; u.w_out = w_in / stride;
    u.w_off = add_seg(conv + ".weight", 0, static_cast<int64_t>(cout) * k * k * cin, cout, k, k, cin);
    u.g_off = add_seg(bn + ".weight", 0, cout, cout, 0, 0, 0);
    u.b_off = add_seg(bn + ".bias", 0, cout, cout, 0, 0, 0);
    u.rm_off = add_seg(bn + ".running_mean", 1, cout, cout, 0, 0, 0);
    u.rv_off = add_seg(bn + ".running_var", 1, cout, cout, 0, 0, 0);
    t->units.push_back(u);
  };
  const int d = t->d;
  add_unit("conv1", "bn1", 5, 64, 7, 2, d, d);
  int inpl = 64, h = d / 4;
  const int planes_[4] = {64, 128, 256, 512};
  const int blocks_[4] = {3, 4, 6, 3};
  for (int li = 0; li < 4; ++li)
    for (int b = 0; b < blocks_[li]; ++b) {
      const std::string pre = "layer" + std::to_string(li + 1) + "." + std::to_string(b);
      const int planes = planes_[li];
      const int stride = (b == 0 && li > 0) ? 2 : 1;
      add_unit(pre + ".conv1", pre + ".bn1", inpl, planes, 1, 1, h, h);
      add_unit(pre + ".conv2", pre + ".bn2", planes, planes, 3, stride, h, h);
      if (b == 0) add_unit(pre + ".downsample.0", pre + ".downsample.1", inpl, planes * 4, 1, stride, h, h);
      add_unit(pre + ".conv3", pre + ".bn3", planes, planes * 4, 1, 1, h / stride, h / stride);
      inpl = planes * 4;
      h /= stride;
    }
  const char* heads1[1] = {"fc"};
  const char* heads2[2] = {"fc_occ", "fc_depth"};
  const char* const* heads = t->n_heads == 1 ? heads1 : heads2;
  // the FC heads are stored as one [k_total][2048] matrix + [k_total] bias (rows in head order): the two segments of
  // a head are consecutive slices of them
  t->fc_w_off = t->n_params;
  int row = 0;
  for (int hI = 0; hI < t->n_heads; ++hI) {
    const int k = t->num_classes[hI];
    t->segs.push_back(Segment{std::string(heads[hI]) + ".weight", 0, t->fc_w_off + static_cast<int64_t>(row) * 2048,
                              {k, 2048, 0, 0}});
    row += k;
  }
  t->n_params = align_up(t->fc_w_off + static_cast<int64_t>(t->k_total) * 2048, 64);
  t->fc_b_off = t->n_params;
  row = 0;
  for (int hI = 0; hI < t->n_heads; ++hI) {
    const int k = t->num_classes[hI];
    t->segs.push_back(Segment{std::string(heads[hI]) + ".bias", 0, t->fc_b_off + row, {k, 0, 0, 0}});
    row += k;
  }
  t->n_params = align_up(t->fc_b_off + t->k_total, 64);
}

template <typename T>
static int dev_alloc(io_train* t, T** p, size_t count) {
  void* q = nullptr;
  IO_CUDA(cudaMalloc(&q, count * sizeof(T) > 0 ? count * sizeof(T) : 16));
  t->allocs.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return IO_OK;
}

// ---- op builders ------------------------------------------------------------------------------------------------
static void push(std::vector<TOp>& v, int kind, double flops, double bytes, int tag,
                 std::function<int(cudaStream_t)> f) {
  TOp op;
  op.run = std::move(f);
  op.kind = kind; op.flops = flops; op.bytes = bytes; op.tag = tag;
  v.push_back(std::move(op));
}

// forward convolution of unit u over `x` -> u.y (raw, no bias / ReLU)
static int add_conv_fwd(io_train* t, const Unit& u, const __nv_bfloat16* x, int tag) {
  const ConvDesc d{t->imgs, u.h_in, u.w_in, u.cin, u.cout, u.k, u.stride};
  const void* w = t->w16 + u.w_off;
  const double flops = 2.0 * t->imgs * u.h_out * u.w_out * u.k * u.k * static_cast<double>(u.cin) * u.cout;
  const double bytes = 2.0 * t->imgs * (static_cast<double>(u.h_in) * u.w_in * u.cin + u.h_out * u.w_out * u.cout);
  if (tn_enabled() && conv_tn_supported(d)) {
    TnParams tp;
    if (int rc = conv_tn_plan(&tp, d, x, w, t->zero_bias, u.y, 0)) return rc;
    push(t->fwd, 0, flops, bytes, tag, [tp](cudaStream_t s) { return conv_tn_launch(tp, s); });
  } else {
    ConvParams p;
    int bn = 0;
    if (int rc = conv_plan(&p, &bn, d, x, w, t->zero_bias, nullptr, u.y, 0)) return rc;
    push(t->fwd, 0, flops, bytes, tag, [p, bn](cudaStream_t s) { return conv_tc_launch(p, bn, s); });
  }
  return IO_OK;
}

// train-mode BN of unit u: statistics of u.y per direction group, then a = [relu](bn(y) [+ residual])
static void add_bn_fwd(io_train* t, Unit& u, const __nv_bfloat16* residual, int relu, int tag) {
  const int rows = t->pairs * u.h_out * u.w_out, c = u.cout;
  const double act = 2.0 * 2 * rows * c;
  Unit* up = &u;
  double* mine = t->red[t->red_sel];
  double* next = t->red[t->red_sel ^ 1];
  t->red_sel ^= 1;
  if (t->fused_bn) {   // statistics + apply in one cooperative launch (grid barrier in between)
    unsigned int* bar = t->barriers + (t->barrier_cursor++ % 256);
    push(t->fwd, 3, 0, act * (residual ? 4.0 : 3.0), tag, [t, up, residual, rows, c, relu, mine, next, bar](cudaStream_t s) {
      return bn_fwd_fused_launch(up->y, residual, up->a, 2, rows, c, mine, t->params + up->g_off,
                                 t->params + up->b_off, 1e-5f, 0.1f, up->save, t->stats + up->rm_off,
                                 t->stats + up->rv_off, next, relu, (residual && relu) ? up->mask : nullptr, bar, s);
    });
    return;
  }
  push(t->fwd, 3, 0, act, tag, [up, rows, c, mine](cudaStream_t s) {
    return bn_stats_launch(up->y, 2, rows, c, mine, s);
  });
  push(t->fwd, 3, 0, act * (residual ? 3.0 : 2.0), tag, [t, up, residual, rows, c, relu, mine, next](cudaStream_t s) {
    return bn_apply_launch(up->y, residual, up->a, 2, rows, c, mine, t->params + up->g_off, t->params + up->b_off,
                           1e-5f, 0.1f, up->save, t->stats + up->rm_off, t->stats + up->rv_off, next, relu,
                           (residual && relu) ? up->mask : nullptr, s);
  });
}

// backward through BN (+ ReLU) of unit u: da -> dy (and the masked gradient g if g_out), BN parameter gradients.
// mask_mode: 0 none, 1 from the stored activation (residual blocks), 2 recomputed from y
static void add_bn_bwd(io_train* t, Unit& u, const __nv_bfloat16* da, __nv_bfloat16* dy, __nv_bfloat16* g_out,
                       int mask_mode, int tag) {
  const int rows = t->pairs * u.h_out * u.w_out, c = u.cout;
  const double act = 2.0 * 2 * rows * c;
  const double reads = mask_mode == 1 ? 3.0 : (mask_mode == 3 ? 2.0625 : 2.0);
  Unit* up = &u;
  const void* mask_src = mask_mode == 3 ? static_cast<const void*>(u.mask) : static_cast<const void*>(u.a);
  double* mine = t->red[t->red_sel];
  double* next = t->red[t->red_sel ^ 1];
  t->red_sel ^= 1;
  if (t->fused_bn) {
    unsigned int* bar = t->barriers + (t->barrier_cursor++ % 256);
    push(t->bwd, 3, 0, act * (2.0 * reads + 1.0 + (g_out ? 1.0 : 0.0)), tag,
         [t, up, da, dy, g_out, rows, c, mask_mode, mine, next, mask_src, bar](cudaStream_t s) {
           return bn_bwd_fused_launch(da, mask_src, up->y, dy, g_out, 2, rows, c, t->params + up->g_off, up->save,
                                      mine, mask_mode, t->grads + up->g_off, t->grads + up->b_off, next, bar, s);
         });
    t->bwd.back().writes = dy;
    return;
  }
  push(t->bwd, 3, 0, act * reads, tag, [up, da, rows, c, mask_mode, mine, mask_src](cudaStream_t s) {
    return bn_bwd_reduce_launch(da, mask_src, up->y, 2, rows, c, up->save, mask_mode, mine, s);
  });
  push(t->bwd, 3, 0, act * (reads + 1.0 + (g_out ? 1.0 : 0.0)), tag,
       [t, up, da, dy, g_out, rows, c, mask_mode, mine, next, mask_src](cudaStream_t s) {
         return bn_bwd_apply_launch(da, mask_src, up->y, dy, g_out, 2, rows, c, t->params + up->g_off, up->save, mine,
                                    mask_mode, t->grads + up->g_off, t->grads + up->b_off, next, s);
       });
  t->bwd.back().writes = dy;
}

static int add_wgrad(io_train* t, const Unit& u, const __nv_bfloat16* x, const __nv_bfloat16* dy, int tag) {
  WgradParams p;
  // the plan needs the final gradient pointer, which is bound later: keep the offset and patch at launch
  if (int rc = wgrad_plan(&p, ConvDesc{t->imgs, u.h_in, u.w_in, u.cin, u.cout, u.k, u.stride}, x, dy, nullptr))
    return rc;
  const int64_t off = u.w_off;
  const double bytes = 2.0 * t->imgs * (static_cast<double>(u.h_in) * u.w_in * u.cin + u.h_out * u.w_out * u.cout) +
                       4.0 * u.cout * u.k * u.k * u.cin;
  push(t->bwd, 2, p.flops, bytes, tag, [t, p, off](cudaStream_t s) {
    WgradParams q = p;
    q.dw = t->grads + off;
    return wgrad_launch(q, s);
  });
  t->bwd.back().side = 1;
  t->bwd.back().reads = dy;
  return IO_OK;
}

// data gradient of a stride-1 convolution (1x1 or 3x3) of unit u: dx = conv^T(dy) (+ residual)
static int add_dgrad(io_train* t, const Unit& u, int h, int w, const __nv_bfloat16* dy, const __nv_bfloat16* residual,
                     __nv_bfloat16* dx, int tag) {
  ConvParams p;
  int bn = 0;
  if (int rc = dgrad_plan(&p, &bn, t->imgs, h, w, u.cin, u.cout, u.k, dy, t->w16 + u.w_off, t->zero_bias, residual, dx))
    return rc;
  const double flops = 2.0 * t->imgs * h * w * u.k * u.k * static_cast<double>(u.cin) * u.cout;
  const double bytes = 2.0 * t->imgs * h * w * (static_cast<double>(u.cin) * (residual ? 2 : 1) + u.cout);
  push(t->bwd, 1, flops, bytes, tag, [p, bn](cudaStream_t s) { return conv_tc_launch(p, bn, s); });
  return IO_OK;
}

static int build_graph(io_train* t) {
  const int I = t->imgs, d = t->d;
  // ---- buffers ----
  size_t max_act = static_cast<size_t>(d / 2) * (d / 2) * 64;   // stem output per image
  size_t max_z = 0;
  for (Unit& u : t->units) {
    const size_t out = static_cast<size_t>(u.h_out) * u.w_out * u.cout;
    if (out > max_act) max_act = out;
    if (u.k == 3 && u.stride == 2) max_z = std::max(max_z, static_cast<size_t>(u.h_in) * u.w_in * u.cout);
    if (int rc = dev_alloc(t, &u.y, out * I)) return rc;
    if (int rc = dev_alloc(t, &u.a, out * I)) return rc;
    if (int rc = dev_alloc(t, &u.save, static_cast<size_t>(8) * u.cout)) return rc;
    if (u.conv.find(".conv3") != std::string::npos)
      if (int rc = dev_alloc(t, &u.mask, out * I / 8)) return rc;
  }
  for (int i = 0; i < 5; ++i)
    if (int rc = dev_alloc(t, &t->gbuf[i], max_act * I)) return rc;
  if (int rc = dev_alloc(t, &t->gbuf[5], max_z * I)) return rc;
  if (int rc = dev_alloc(t, &t->w16, static_cast<size_t>(t->n_params))) return rc;
  if (int rc = dev_alloc(t, &t->stem_pk, 128 * 448)) return rc;
  if (int rc = dev_alloc(t, &t->stem_scratch, 128 * 448)) return rc;
  if (int rc = dev_alloc(t, &t->zero_bias, 2048)) return rc;
  IO_CUDA(cudaMemset(t->zero_bias, 0, 2048 * sizeof(float)));
  if (int rc = dev_alloc(t, &t->barriers, 256)) return rc;
  if (const char* e = getenv("INSTAORDER_TRAIN_FUSED_BN")) t->fused_bn = atoi(e) != 0;
  for (int i = 0; i < 2; ++i) {
    if (int rc = dev_alloc(t, &t->red[i], 2 * 2 * 2048)) return rc;
    IO_CUDA(cudaMemset(t->red[i], 0, sizeof(double) * 2 * 2 * 2048));
  }
  const size_t pool_el = static_cast<size_t>(I) * (d / 4) * (d / 4) * 64;
  if (int rc = dev_alloc(t, &t->pool_idx, pool_el)) return rc;
  if (int rc = dev_alloc(t, &t->pool_out, pool_el)) return rc;
  if (int rc = dev_alloc(t, &t->pooled, static_cast<size_t>(I) * 2048)) return rc;
  if (int rc = dev_alloc(t, &t->logits, static_cast<size_t>(I) * t->k_total)) return rc;
  if (int rc = dev_alloc(t, &t->dlogits, static_cast<size_t>(I) * t->k_total)) return rc;
  t->dy_ring[0] = t->gbuf[2];
  for (int i = 1; i < 3; ++i)
    if (int rc = dev_alloc(t, &t->dy_ring[i], max_act * I)) return rc;
  __nv_bfloat16 *GA = t->gbuf[0], *GG = t->gbuf[1], *GY = t->gbuf[2], *GT = t->gbuf[3], *GD = t->gbuf[4],
                *GZ = t->gbuf[5];
  auto next_dy = [t]() { return t->dy_ring[t->dy_cursor++ % 3]; };

  // ---- forward ----
  Unit& stem = t->units[0];
  {
    const double flops = 2.0 * I * (d / 2) * (d / 2) * 49.0 * 5.0 * 64.0;
    const double bytes = static_cast<double>(io_pair_tensor_bytes(t->pairs, d)) + 2.0 * I * (d / 2) * (d / 2) * 64;
    Unit* sp = &stem;
    push(t->fwd, 0, flops, bytes, 1, [t, sp](cudaStream_t s) -> int {
      const int hw = (t->d / 2) * (t->d / 2);
      if (tn_enabled() && stem_tn_supported(t->d)) {
        TnParams tp;
        if (int rc = stem_tn_plan(&tp, t->pairs, t->d, t->pair_tensor, t->stem_pk, t->zero_bias, sp->y)) return rc;
        tp.img_mul = 1; tp.split_row_off = t->pairs * hw; tp.relu = 0;
        return conv_tn_launch(tp, s);
      }
      ConvParams p;
      int bn = 0;
      if (int rc = stem_plan(&p, &bn, t->pairs, t->d, t->pair_tensor, t->stem_pk, t->zero_bias, sp->y)) return rc;
      p.img_mul = 1; p.split_row_off = t->pairs * hw; p.relu = 0;
      return conv_tc_launch(p, bn, s);
    });
  }
  add_bn_fwd(t, stem, nullptr, 1, 1);
  Unit* stem_p = &stem;
  push(t->fwd, 3, 0, 2.0 * I * (d / 2) * (d / 2) * 64 * 1.25 + pool_el, 2, [t, stem_p](cudaStream_t s) {
    return maxpool_fwd_idx_launch(stem_p->a, t->pool_out, t->pool_idx, t->imgs, t->d / 2, t->d / 2, 64, s);
  });
  const int blocks_[4] = {3, 4, 6, 3};
  struct Blk { int c1, c2, ds, c3; const __nv_bfloat16* x; int tag; };
  std::vector<Blk> blks;
  size_t ui = 1;
  const __nv_bfloat16* X = t->pool_out;
  for (int li = 0; li < 4; ++li)
    for (int b = 0; b < blocks_[li]; ++b) {
      Blk k;
      k.c1 = static_cast<int>(ui++);
      k.c2 = static_cast<int>(ui++);
      k.ds = (b == 0) ? static_cast<int>(ui++) : -1;
      k.c3 = static_cast<int>(ui++);
      k.x = X;
      k.tag = (li + 1) * 100 + b * 10;
      Unit &u1 = t->units[k.c1], &u2 = t->units[k.c2], &u3 = t->units[k.c3];
      if (int rc = add_conv_fwd(t, u1, X, k.tag + 1)) return rc;
      add_bn_fwd(t, u1, nullptr, 1, k.tag + 1);
      if (int rc = add_conv_fwd(t, u2, u1.a, k.tag + 2)) return rc;
      add_bn_fwd(t, u2, nullptr, 1, k.tag + 2);
      if (int rc = add_conv_fwd(t, u3, u2.a, k.tag + 3)) return rc;
      const __nv_bfloat16* identity = X;
      if (k.ds >= 0) {
        Unit& ud = t->units[k.ds];
        if (int rc = add_conv_fwd(t, ud, X, k.tag + 4)) return rc;
        add_bn_fwd(t, ud, nullptr, 0, k.tag + 4);
        identity = ud.a;
      }
      add_bn_fwd(t, u3, identity, 1, k.tag + 3);
      X = u3.a;
      blks.push_back(k);
    }
  const Unit& last = t->units.back();
  const int hw_final = last.h_out * last.w_out;
  push(t->fwd, 3, 2.0 * I * 2048.0 * t->k_total, 2.0 * I * hw_final * 2048, 3, [t, X, hw_final](cudaStream_t s) {
    return pool_fc_fwd_launch(X, hw_final, t->imgs, t->params + t->fc_w_off, t->params + t->fc_b_off, t->k_total,
                              t->pooled, t->logits, s);
  });
  push(t->fwd, 4, 0, 0, 4, [t](cudaStream_t s) {
    return loss_train_launch(t->logits, t->pairs, t->k_total, t->occ_off, t->cls_off, t->cls_k, t->occ_target,
                             t->class_target, t->is_overlap, t->overlap_w, t->distinct_w, t->world_size,
                             t->out_losses, t->dlogits, s);
  });

  // ---- backward ----
  push(t->bwd, 3, 2.0 * I * 2048.0 * t->k_total * 2, 2.0 * I * hw_final * 2048, 3, [t, GA, hw_final](cudaStream_t s) {
    return pool_fc_bwd_launch(t->dlogits, t->pooled, t->params + t->fc_w_off, hw_final, t->imgs, t->k_total,
                              t->grads + t->fc_w_off, t->grads + t->fc_b_off, GA, s);
  });
  {
    // first block of layers 4, 3, 2 in `blks` (3 + 4 + 6 + 3 blocks): everything behind it in the flat buffer is final
    // once its backward ops have run
    const int first_blk[3] = {13, 7, 3};
    int64_t hi = t->n_params;
    for (int k = 0; k < 3; ++k) {
      t->bucket_begin[k] = t->units[blks[first_blk[k]].c1].w_off;
      t->bucket_end[k] = hi;
      hi = t->bucket_begin[k];
    }
    t->bucket_begin[3] = 0;
    t->bucket_end[3] = hi;
  }
  for (int bi = static_cast<int>(blks.size()) - 1; bi >= 0; --bi) {
    if (bi == 12) t->bucket_after_op[0] = t->bwd.size();
    if (bi == 6) t->bucket_after_op[1] = t->bwd.size();
    if (bi == 2) t->bucket_after_op[2] = t->bwd.size();
    const Blk& k = blks[bi];
    Unit &u1 = t->units[k.c1], &u2 = t->units[k.c2], &u3 = t->units[k.c3];
    // out = relu(bn3(conv3(a2)) + identity):  GA = d out  ->  GY = d y3, GG = masked gradient (identity branch)
    GY = next_dy();
    add_bn_bwd(t, u3, GA, GY, GG, 3, k.tag + 3);   // ReLU bit mask of the block output
    if (int rc = add_wgrad(t, u3, u2.a, GY, k.tag + 3)) return rc;
    if (int rc = add_dgrad(t, u3, u3.h_out, u3.w_out, GY, nullptr, GT, k.tag + 3)) return rc;   // GT = d a2
    GY = next_dy();
    add_bn_bwd(t, u2, GT, GY, nullptr, 2, k.tag + 2);                                          // GY = d y2
    if (int rc = add_wgrad(t, u2, u1.a, GY, k.tag + 2)) return rc;
    if (u2.stride == 1) {
      if (int rc = add_dgrad(t, u2, u2.h_in, u2.w_in, GY, nullptr, GT, k.tag + 2)) return rc;  // GT = d a1
    } else {
      const int ho = u2.h_out, wo = u2.w_out, c = u2.cout;
      push(t->bwd, 3, 0, 2.0 * I * ho * wo * c * 5.0, k.tag + 2, [t, GY, GZ, ho, wo, c](cudaStream_t s) {
        return upsample2_zero_launch(GY, GZ, t->imgs, ho, wo, c, s);
      });
      if (int rc = add_dgrad(t, u2, u2.h_in, u2.w_in, GZ, nullptr, GT, k.tag + 2)) return rc;
    }
    GY = next_dy();
    add_bn_bwd(t, u1, GT, GY, nullptr, 2, k.tag + 1);                                          // GY = d y1
    if (int rc = add_wgrad(t, u1, k.x, GY, k.tag + 1)) return rc;
    if (k.ds < 0) {
      if (int rc = add_dgrad(t, u1, u1.h_in, u1.w_in, GY, GG, GA, k.tag + 1)) return rc;       // GA = d x
    } else {
      Unit& ud = t->units[k.ds];
      add_bn_bwd(t, ud, GG, GD, nullptr, 0, k.tag + 4);                                        // GD = d y_ds
      if (int rc = add_wgrad(t, ud, k.x, GD, k.tag + 4)) return rc;
      if (int rc = add_dgrad(t, ud, ud.h_out, ud.w_out, GD, nullptr, GG, k.tag + 4)) return rc;  // GG = low-res d x
      if (ud.stride == 1) {
        if (int rc = add_dgrad(t, u1, u1.h_in, u1.w_in, GY, GG, GA, k.tag + 1)) return rc;
      } else {
        if (int rc = add_dgrad(t, u1, u1.h_in, u1.w_in, GY, nullptr, GA, k.tag + 1)) return rc;
        const int ho = ud.h_out, wo = ud.w_out, c = ud.cin;
        push(t->bwd, 3, 0, 2.0 * I * ho * wo * c * 3.0, k.tag + 4, [t, GG, GA, ho, wo, c](cudaStream_t s) {
          return scatter_add2_launch(GG, GA, t->imgs, ho, wo, c, s);
        });
      }
    }
  }
  // max-pool, stem BN + ReLU, stem weight gradient
  push(t->bwd, 3, 0, 2.0 * I * (d / 2) * (d / 2) * 64 * 2.0, 2, [t, GA, GT](cudaStream_t s) {
    return maxpool_bwd_launch(GA, t->pool_idx, GT, t->imgs, t->d / 2, t->d / 2, 64, s);
  });
  GY = next_dy();
  add_bn_bwd(t, stem, GT, GY, nullptr, 2, 1);
  {
    const double flops = 2.0 * I * (d / 2) * (d / 2) * 49.0 * 5.0 * 64.0;
    const double bytes = static_cast<double>(io_pair_tensor_bytes(t->pairs, d)) + 2.0 * I * (d / 2) * (d / 2) * 64;
    const int64_t off = stem.w_off;
    push(t->bwd, 2, flops, bytes, 1, [t, GY, off](cudaStream_t s) -> int {
      IO_CUDA(cudaMemsetAsync(t->stem_scratch, 0, sizeof(float) * 128 * 448, s));
      WgradParams p;
      if (int rc = stem_wgrad_plan(&p, t->pairs, t->d, t->pair_tensor, GY, t->stem_scratch)) return rc;
      if (int rc = wgrad_launch(p, s)) return rc;
      return stem_unpack_grad_launch(t->stem_scratch, t->grads + off, s);
    });
  }
  t->bucket_after_op[3] = t->bwd.size();
  for (int k = 0; k < io_train::kBuckets; ++k)
    for (int j = 0; j < 2; ++j) IO_CUDA(cudaEventCreateWithFlags(&t->bucket_ev[k][j], cudaEventDisableTiming));
  IO_CUDA(cudaStreamCreateWithFlags(&t->side_stream, cudaStreamNonBlocking));
  if (const char* e = getenv("INSTAORDER_TRAIN_SIDE_STREAM")) t->use_side = atoi(e) != 0;
  t->built = true;
  return IO_OK;
}

static int run_ops(io_train* t, std::vector<TOp>& ops, cudaStream_t stream, bool is_bwd = false) {
  // weight gradients go to a side stream (they only feed the gradient buffer) so that their tensor work overlaps the
  // HBM-bound BatchNorm passes of the main chain; per-op profiling keeps everything on one stream
  const bool side_on = t->use_side && !t->profile && t->side_stream != nullptr;
  std::vector<std::pair<const void*, cudaEvent_t>> pending;   // dy buffer -> "its side reader is done" event
  size_t ev_i = 0;
  auto next_event = [&](cudaEvent_t* out) -> int {
    if (ev_i >= t->side_ev.size()) {
      cudaEvent_t e;
      IO_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      t->side_ev.push_back(e);
    }
    *out = t->side_ev[ev_i++];
    return IO_OK;
  };
  bool forked = false;
  size_t op_i = 0;
  for (TOp& op : ops) {
    if (is_bwd)
      for (int k = 0; k < io_train::kBuckets - 1; ++k)
        if (op_i == t->bucket_after_op[k]) {       // everything of bucket k has been launched
          IO_CUDA(cudaEventRecord(t->bucket_ev[k][0], stream));
          t->bucket_side[k] = forked;
          if (forked) IO_CUDA(cudaEventRecord(t->bucket_ev[k][1], t->side_stream));
        }
    ++op_i;
    size_t idx = 0;
    if (t->profile) {
      idx = 2 * t->prof_kind.size();
      while (t->ev.size() <= idx + 1) {
        cudaEvent_t e;
        IO_CUDA(cudaEventCreate(&e));
        t->ev.push_back(e);
      }
      IO_CUDA(cudaEventRecord(t->ev[idx], stream));
    }
    if (side_on && op.writes != nullptr) {
      for (size_t k = 0; k < pending.size(); ++k)
        if (pending[k].first == op.writes) {
          IO_CUDA(cudaStreamWaitEvent(stream, pending[k].second, 0));
          pending.erase(pending.begin() + k);
          break;
        }
    }
    if (side_on && op.side) {
      cudaEvent_t fork, done;
      if (int rc = next_event(&fork)) return rc;
      if (int rc = next_event(&done)) return rc;
      IO_CUDA(cudaEventRecord(fork, stream));
      IO_CUDA(cudaStreamWaitEvent(t->side_stream, fork, 0));
      if (int rc = op.run(t->side_stream)) return rc;
      IO_CUDA(cudaEventRecord(done, t->side_stream));
      bool replaced = false;
      for (auto& pr : pending)
        if (pr.first == op.reads) { pr.second = done; replaced = true; }
      if (!replaced) pending.emplace_back(op.reads, done);
      forked = true;
    } else {
      if (int rc = op.run(stream)) return rc;
    }
    if (t->profile) {
      IO_CUDA(cudaEventRecord(t->ev[idx + 1], stream));
      t->prof_kind.push_back(op.kind);
      t->prof_tag.push_back(op.tag);
      t->prof_flops.push_back(op.flops);
      t->prof_bytes.push_back(op.bytes);
    }
    ++t->last_launches;
  }
  if (forked) {   // join: everything the side stream did is ordered before whatever follows on the caller's stream
    cudaEvent_t join;
    if (int rc = next_event(&join)) return rc;
    IO_CUDA(cudaEventRecord(join, t->side_stream));
    IO_CUDA(cudaStreamWaitEvent(stream, join, 0));
  }
  if (is_bwd) {   // the last bucket (stem + layer1) is complete when the whole backward pass is
    IO_CUDA(cudaEventRecord(t->bucket_ev[io_train::kBuckets - 1][0], stream));
    t->bucket_side[io_train::kBuckets - 1] = false;
  }
  return IO_OK;
}

}  // namespace io

using namespace io;

extern "C" int io_train_create(const int32_t* num_classes, int n_heads, int input_size, int batch_pairs,
                               io_train_t** out) {
  IO_REQUIRE(num_classes && out, "io_train_create: null pointer");
  IO_REQUIRE(n_heads == 1 || n_heads == 2, "io_train_create: n_heads must be 1 or 2");
  IO_REQUIRE(input_size >= 64 && input_size <= 512 && input_size % 32 == 0, "io_train_create: input_size %d",
             input_size);
  IO_REQUIRE(batch_pairs >= 1 && batch_pairs <= 1024, "io_train_create: batch_pairs %d", batch_pairs);
  int dev_count = 0;
  IO_CUDA(cudaGetDeviceCount(&dev_count));
  std::unique_ptr<io_train> t(new io_train());
  t->n_heads = n_heads;
  for (int i = 0; i < n_heads; ++i) {
    IO_REQUIRE(num_classes[i] >= 1 && num_classes[i] <= 4, "io_train_create: num_classes[%d] = %d", i, num_classes[i]);
    t->num_classes[i] = num_classes[i];
    t->k_total += num_classes[i];
  }
  t->d = input_size;
  t->pairs = batch_pairs;
  t->imgs = 2 * batch_pairs;
  build_units(t.get());
  *out = t.release();
  return IO_OK;
}

extern "C" int io_train_destroy(io_train_t* t) {
  if (!t) return IO_OK;
  for (void* p : t->allocs) cudaFree(p);
  for (cudaEvent_t e : t->ev) cudaEventDestroy(e);
  for (cudaEvent_t e : t->side_ev) cudaEventDestroy(e);
  for (int k = 0; k < io_train::kBuckets; ++k)
    for (int j = 0; j < 2; ++j)
      if (t->bucket_ev[k][j]) cudaEventDestroy(t->bucket_ev[k][j]);
  if (t->side_stream) cudaStreamDestroy(t->side_stream);
  delete t;
  return IO_OK;
}

extern "C" int64_t io_train_param_count(const io_train_t* t) { return t ? t->n_params : 0; }
extern "C" int64_t io_train_stat_count(const io_train_t* t) { return t ? t->n_stats : 0; }
extern "C" int io_train_num_segments(const io_train_t* t) { return t ? static_cast<int>(t->segs.size()) : 0; }

extern "C" int io_train_segment(const io_train_t* t, int i, char* name, int name_cap, int32_t* buffer, int64_t* offset,
                                int32_t* dims4) {
  IO_REQUIRE(t && name && buffer && offset && dims4, "io_train_segment: null pointer");
  IO_REQUIRE(i >= 0 && i < static_cast<int>(t->segs.size()), "io_train_segment: index %d", i);
  const Segment& s = t->segs[i];
  snprintf(name, static_cast<size_t>(name_cap), "%s", s.name.c_str());
  *buffer = s.buffer;
  *offset = s.offset;
  for (int k = 0; k < 4; ++k) dims4[k] = s.dims[k];
  return IO_OK;
}

extern "C" int io_train_bind(io_train_t* t, float* params_dev, float* grads_dev, float* stats_dev) {
  IO_REQUIRE(t && params_dev && grads_dev && stats_dev, "io_train_bind: null pointer");
  t->params = params_dev;
  t->grads = grads_dev;
  t->stats = stats_dev;
  t->weights_synced = false;
  if (!t->built) return build_graph(t);
  return IO_OK;
}

extern "C" int io_train_sync_weights(io_train_t* t, void* stream) {
  IO_REQUIRE(t && t->params, "io_train_sync_weights: bind the parameter buffers first");
  if (int rc = cast_bf16_launch(t->params, t->w16, t->n_params, as_stream(stream))) return rc;
  if (int rc = stem_pack_launch(t->params + t->units[0].w_off, t->stem_pk, as_stream(stream))) return rc;
  t->weights_synced = true;
  return IO_OK;
}

extern "C" int io_train_forward_backward(io_train_t* t, const void* pair_tensor_dev, int occ_off, int class_off,
                                         int class_k, const float* occ_target_dev, const int64_t* class_target_dev,
                                         const int64_t* is_overlap_dev, float overlap_w, float distinct_w,
                                         int world_size, float* out_losses_dev, int run_backward, void* stream_) {
  IO_REQUIRE(t && pair_tensor_dev && out_losses_dev, "io_train_forward_backward: null pointer");
  if (!t->built || !t->params) {
    set_error("io_train_forward_backward: call io_train_bind first");
    return IO_ERR_STATE;
  }
  if (!t->weights_synced) {
    set_error("io_train_forward_backward: call io_train_sync_weights after loading / changing the parameters");
    return IO_ERR_STATE;
  }
  IO_REQUIRE(occ_off < 0 || (occ_target_dev && occ_off + 2 <= t->k_total), "io_train_forward_backward: occlusion head");
  IO_REQUIRE(class_off < 0 || (class_target_dev && class_k >= 2 && class_k <= 4 && class_off + class_k <= t->k_total),
             "io_train_forward_backward: class head");
  IO_REQUIRE(world_size >= 1, "io_train_forward_backward: world_size");
  cudaStream_t stream = as_stream(stream_);
  t->pair_tensor = pair_tensor_dev;
  t->occ_off = occ_off; t->cls_off = class_off; t->cls_k = class_k;
  t->occ_target = occ_target_dev; t->class_target = class_target_dev; t->is_overlap = is_overlap_dev;
  t->overlap_w = overlap_w; t->distinct_w = distinct_w; t->world_size = world_size;
  t->out_losses = out_losses_dev;
  t->last_launches = 0;
  t->prof_kind.clear(); t->prof_tag.clear(); t->prof_flops.clear(); t->prof_bytes.clear();
  IO_CUDA(cudaMemsetAsync(t->barriers, 0, sizeof(unsigned int) * 256, stream));
  if (int rc = run_ops(t, t->fwd, stream)) return rc;
  if (run_backward) {
    IO_CUDA(cudaMemsetAsync(t->grads, 0, sizeof(float) * t->n_params, stream));
    if (int rc = run_ops(t, t->bwd, stream, true)) return rc;
  }
  return IO_OK;
}

extern "C" int io_train_num_buckets(const io_train_t* t) { return t ? io_train::kBuckets : 0; }

extern "C" int io_train_bucket(const io_train_t* t, int k, int64_t* begin, int64_t* end) {
  IO_REQUIRE(t && begin && end && t->built && k >= 0 && k < io_train::kBuckets, "io_train_bucket: bad arguments");
  *begin = t->bucket_begin[k];
  *end = t->bucket_end[k];
  return IO_OK;
}

// Makes `stream` wait until the gradients of bucket k of the LAST io_train_forward_backward call are complete (their
// last kernels on the main and on the weight-gradient stream).  No host synchronisation.
extern "C" int io_train_wait_bucket(io_train_t* t, int k, void* stream) {
  IO_REQUIRE(t && t->built && k >= 0 && k < io_train::kBuckets, "io_train_wait_bucket: bad arguments");
  IO_CUDA(cudaStreamWaitEvent(as_stream(stream), t->bucket_ev[k][0], 0));
  if (t->bucket_side[k]) IO_CUDA(cudaStreamWaitEvent(as_stream(stream), t->bucket_ev[k][1], 0));
  return IO_OK;
}

extern "C" int io_train_sgd_step(io_train_t* t, float* momentum_buf_dev, float lr, float momentum, float weight_decay,
                                 int first_step, void* stream) {
  IO_REQUIRE(t && t->params && momentum_buf_dev, "io_train_sgd_step: bind the buffers first");
  if (int rc = io_optim_sgd(t->params, t->grads, momentum_buf_dev, t->n_params, lr, momentum, weight_decay, first_step,
                            t->w16, t->n_params, stream))
    return rc;
  return stem_pack_launch(t->params + t->units[0].w_off, t->stem_pk, as_stream(stream));
}

extern "C" int io_train_adam_step(io_train_t* t, float* m_dev, float* v_dev, float lr, float beta1, float beta2,
                                  float eps, int step, void* stream) {
  IO_REQUIRE(t && t->params && m_dev && v_dev, "io_train_adam_step: bind the buffers first");
  if (int rc = io_optim_adam(t->params, t->grads, m_dev, v_dev, t->n_params, lr, beta1, beta2, eps, step, t->w16,
                             t->n_params, stream))
    return rc;
  return stem_pack_launch(t->params + t->units[0].w_off, t->stem_pk, as_stream(stream));
}

extern "C" int io_train_read_logits(const io_train_t* t, float* out_dev, void* stream) {
  IO_REQUIRE(t && out_dev && t->logits, "io_train_read_logits: null pointer / handle not bound");
  IO_CUDA(cudaMemcpyAsync(out_dev, t->logits, sizeof(float) * t->imgs * t->k_total, cudaMemcpyDeviceToDevice,
                          as_stream(stream)));
  return IO_OK;
}
extern "C" int io_train_read_activation(const io_train_t* t, const char* conv_name, int which, void* out_dev,
                                        int64_t* numel_out, void* stream) {
  IO_REQUIRE(t && conv_name && numel_out && t->built, "io_train_read_activation: null pointer / handle not bound");
  for (const Unit& u : t->units) {
    if (u.conv != conv_name) continue;
    const int64_t n = static_cast<int64_t>(t->imgs) * u.h_out * u.w_out * u.cout;
    *numel_out = n;
    if (out_dev)
      IO_CUDA(cudaMemcpyAsync(out_dev, which == 0 ? u.y : u.a, n * 2, cudaMemcpyDeviceToDevice, as_stream(stream)));
    return IO_OK;
  }
  set_error("io_train_read_activation: no convolution named '%s'", conv_name);
  return IO_ERR_ARG;
}

extern "C" int io_train_last_launches(const io_train_t* t) { return t ? t->last_launches : 0; }

extern "C" int io_train_profile(io_train_t* t, int enable) {
  IO_REQUIRE(t, "io_train_profile: null handle");
  t->profile = enable != 0;
  return IO_OK;
}

extern "C" int io_train_profile_read(io_train_t* t, float* ms, int32_t* kind, double* flop, double* bytes, int32_t* tag,
                                     int max_n) {
  IO_REQUIRE(t && ms && kind && flop, "io_train_profile_read: null pointer");
  const int n = static_cast<int>(t->prof_kind.size());
  for (int i = 0; i < n && i < max_n; ++i) {
    IO_CUDA(cudaEventElapsedTime(&ms[i], t->ev[2 * i], t->ev[2 * i + 1]));
    kind[i] = t->prof_kind[i];
    flop[i] = t->prof_flops[i];
    if (bytes) bytes[i] = t->prof_bytes[i];
    if (tag) tag[i] = t->prof_tag[i];
  }
  return n;
}

// ---- exported single kernels for the per-kernel parity tests ------------------------------------------------
extern "C" int io_bn_train_forward(const void* y_dev, const void* residual_dev, void* a_dev, int groups, int rows, int c,
                                   const float* gamma_dev, const float* beta_dev, float eps, float momentum,
                                   float* running_mean_dev, float* running_var_dev, float* save_dev,
                                   double* scratch_dev, int relu, uint8_t* mask_out_dev, void* stream_) {
  // save_dev: 4 x [groups][c] fp32 (scale, shift, mean, invstd); scratch_dev: [groups][2][c] doubles
  IO_REQUIRE(y_dev && a_dev && gamma_dev && beta_dev && running_mean_dev && running_var_dev && save_dev && scratch_dev,
             "io_bn_train_forward: null pointer");
  cudaStream_t s = as_stream(stream_);
  const size_t gc = static_cast<size_t>(groups) * c;
  IO_CUDA(cudaMemsetAsync(scratch_dev, 0, sizeof(double) * 2 * gc, s));
  if (int rc = bn_stats_launch(y_dev, groups, rows, c, scratch_dev, s)) return rc;
  return bn_apply_launch(y_dev, residual_dev, a_dev, groups, rows, c, scratch_dev, gamma_dev, beta_dev, eps, momentum,
                         save_dev, running_mean_dev, running_var_dev, nullptr, relu, mask_out_dev, s);
}

extern "C" int io_bn_train_backward(const void* da_dev, const void* a_dev, const void* y_dev, void* dy_dev,
                                    void* g_out_dev, int groups, int rows, int c, const float* gamma_dev,
                                    const float* save_dev, double* scratch_dev, int mask_mode, float* dgamma_dev,
                                    float* dbeta_dev, void* stream_) {
  IO_REQUIRE(da_dev && y_dev && dy_dev && gamma_dev && save_dev && scratch_dev && dgamma_dev && dbeta_dev,
             "io_bn_train_backward: null pointer");
  IO_REQUIRE(mask_mode >= 0 && mask_mode <= 3 && ((mask_mode != 1 && mask_mode != 3) || a_dev),
             "io_bn_train_backward: mask_mode %d (1 needs the activation, 3 the bit mask)", mask_mode);
  cudaStream_t s = as_stream(stream_);
  const size_t gc = static_cast<size_t>(groups) * c;
  IO_CUDA(cudaMemsetAsync(scratch_dev, 0, sizeof(double) * 2 * gc, s));
  if (int rc = bn_bwd_reduce_launch(da_dev, a_dev, y_dev, groups, rows, c, save_dev, mask_mode, scratch_dev, s))
    return rc;
  return bn_bwd_apply_launch(da_dev, a_dev, y_dev, dy_dev, g_out_dev, groups, rows, c, gamma_dev, save_dev,
                             scratch_dev, mask_mode, dgamma_dev, dbeta_dev, nullptr, s);
}

extern "C" int io_maxpool_train(const void* x_dev, void* y_dev, uint8_t* idx_dev, const void* dy_dev, void* dx_dev,
                                int b, int h, int w, int c, void* stream_) {
  IO_REQUIRE(x_dev && y_dev && idx_dev, "io_maxpool_train: null pointer");
  if (int rc = maxpool_fwd_idx_launch(x_dev, y_dev, idx_dev, b, h, w, c, as_stream(stream_))) return rc;
  if (dy_dev && dx_dev) return maxpool_bwd_launch(dy_dev, idx_dev, dx_dev, b, h, w, c, as_stream(stream_));
  return IO_OK;
}
