// io_net_t: the 5-channel ResNet-50 order classifier (reference models/backbone/resnet_cls.py:119-222) as a
// static plan of tcgen05 implicit-GEMM convolutions over bf16 NHWC activations resident in HBM.
//
// Eval-mode BatchNorm is folded into the bf16 GEMM weights + an fp32 bias at load time; ReLU and the residual add
// run in the convolution epilogues; both directions (A,B) / (B,A) of a pair are produced by ONE stem GEMM with
// N = 2 x 64 output channels (the second half uses conv1 weights with input channels 0/1 exchanged), so the pair
// tensor is read once and the swapped input of reference inference.py:145 is never materialised.
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "conv_tc.cuh"
#include "tail.cuh"

namespace io {

struct ConvW {
  std::string name;      // e.g. "layer1.0.conv1"
  std::string bn;        // e.g. "layer1.0.bn1"
  int cin, cout, k, stride;
  __nv_bfloat16* w = nullptr;  // [cout][k*k*cin]
  float* bias = nullptr;       // [cout]
  // downsample convs only: [cout][cmid + cin] = (conv3 | downsample) weights side by side and the two biases added,
  // for the single-GEMM block output of conv_plan_dual
  __nv_bfloat16* wcat = nullptr;
  float* bias_cat = nullptr;
  int cat_k = 0;               // K of wcat = conv3's input channels + this convolution's
};

struct Op {
  enum Kind { STEM, POOL, CONV, TAIL, FUSED, CONV_TN, STEM_TN, ADD, STEM_POOL, CONV_HALO, CONV_ROW3 } kind;
  ConvParams p;
  FusedParams fp;
  TnParams tp;
  StemPoolParams sp;
  HaloParams hp;
  int bn_tile = 0;
  double flops = 0.0;  // algorithmic 2*MAC of this launch
  double bytes = 0.0;  // algorithmic HBM bytes of this launch (activations in + out + residual + weights)
  int tag = 0;         // layer*100 + block*10 + conv index (profiling label)
  // POOL / ADD (ADD: dst += src[idx[image]], b images of h*w*c elements)
  const int32_t* idx = nullptr;
  const void* src = nullptr;
  void* dst = nullptr;
  int b = 0, h = 0, w = 0, c = 0;
};

struct Plan {
  std::vector<Op> ops;
  const void* feat = nullptr;
  int hw_final = 0;
};

}  // namespace io

struct io_net {
  // architecture: bottleneck ResNet family (reference resnet_cls.py Bottleneck / torchvision ResNeXt): per layer the
  // 3x3 width, the block output channels and the block count; n_layers < 4 = feature extractor without layer4
  int widths[4] = {64, 128, 256, 512};
  int outs[4] = {256, 512, 1024, 2048};
  int blocks[4] = {3, 4, 6, 3};
  int n_layers = 4;
  // feature extractor (InstaDepthNet encoder): the output of every layer is kept in its own buffer
  bool keep_layers = false;
  // RGB-only feature extractor: the stem's two directions are identical, so the second one is written behind the
  // first ([direction][image] order) and every later launch works on the first half only
  bool single_dir = false;
  __nv_bfloat16* keep[4] = {nullptr, nullptr, nullptr, nullptr};
  // feature injection (InstaDepthNet trunks, midas_net.py:201-203): x += enc_l[inject_idx[image]] after layer l
  const __nv_bfloat16* inject[3] = {nullptr, nullptr, nullptr};
  const int32_t* inject_idx = nullptr;
  int n_heads = 0;
  int num_classes[2] = {0, 0};
  int k_total = 0;
  int d = 0;
  // network input of the CURRENT plans: d x d, or h x w <= d x d in `orig` mode (io_net_forward_pairs_hw; reference
  // inference.py:401-408).  The back-to-back / dual-source fusions are planned for the default geometry only.
  int h = 0, w = 0;
  bool plain = false;   // plans of an io_net_forward_pairs_hw call: always the general one-launch-per-convolution schedule
  int max_pairs = 0;
  bool loaded = false;
  int last_launches = 0;
  std::vector<io::ConvW> convs;  // [0] = stem, then the 52 bottleneck convs in execution order
  __nv_bfloat16* stem_w = nullptr;
  __nv_bfloat16* stem_w2 = nullptr;     // the same filter in the fused stem + pool kernel's K order (stem_pool.cu)
  float* stem_bias = nullptr;
  float* fc_w = nullptr;
  float* fc_b = nullptr;
  // phase A (stem .. layer2) works on sub-chunks of chunk_a pairs so that its large activations stay in L2;
  // phase B (layer3, layer4, tail) runs over chunk_b pairs at once so that its small GEMMs fill all SMs.
  int chunk_a = 0, chunk_b = 0;
  bool fuse_ds = true;  // first block of a layer: conv3 + downsample as one GEMM over concatenated K (INSTAORDER_FUSE_DS=0 disables)
  bool fuse = true;   // conv3 -> next conv1 back-to-back GEMM fusion in layer1 / layer2 (INSTAORDER_FUSE=0 disables)
  int fuse_layers = 0x7;  // bit li: fuse inside layer li+1 (INSTAORDER_FUSE_LAYERS)
  bool cross_fuse = false;  // layer2's last conv3 also produces layer3.0's conv1 output (phase A writes phase B's T1)
  __nv_bfloat16* buf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // phase A: X, Y, T1, T2, DS
  __nv_bfloat16* bufb[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // phase B: X, Y, T1, T2, DS
  __nv_bfloat16* big = nullptr;                                            // layer2 output of a whole B chunk
  std::map<std::pair<int, int>, std::unique_ptr<io::Plan>> plans_a;        // (pairs, first pair inside the B chunk)
  std::map<int, std::unique_ptr<io::Plan>> plans_b;
  bool profile = false;
  std::vector<cudaEvent_t> ev;       // 2 per launch
  std::vector<int> prof_kind;
  std::vector<double> prof_flops;
  std::vector<double> prof_bytes;
  std::vector<int> prof_tag;
};

namespace io {

static uint16_t bf16_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return static_cast<uint16_t>((u >> 16) | 0x40);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}

static void build_conv_list(io_net* net) {
  net->convs.clear();
  ConvW stem;
  stem.name = "conv1"; stem.bn = "bn1"; stem.cin = 5; stem.cout = 64; stem.k = 7; stem.stride = 2;
  net->convs.push_back(stem);
  int inpl = 64;
  for (int li = 0; li < net->n_layers; ++li) {
    for (int b = 0; b < net->blocks[li]; ++b) {
      const std::string pre = "layer" + std::to_string(li + 1) + "." + std::to_string(b);
      const int width = net->widths[li], outc = net->outs[li];
      const int stride = (b == 0 && li > 0) ? 2 : 1;
      ConvW c1{pre + ".conv1", pre + ".bn1", inpl, width, 1, 1};
      ConvW c2{pre + ".conv2", pre + ".bn2", width, width, 3, stride};
      ConvW c3{pre + ".conv3", pre + ".bn3", width, outc, 1, 1};
      net->convs.push_back(c1);
      net->convs.push_back(c2);
      if (b == 0) {
        ConvW ds{pre + ".downsample.0", pre + ".downsample.1", inpl, outc, 1, stride};
        ds.cat_k = width + inpl;
        net->convs.push_back(ds);
      }
      net->convs.push_back(c3);
      inpl = outc;
    }
  }
}

// x[img] += enc[idx[img]] over [h*w*c] bf16 elements per image, 8 channels (16 B) per thread (midas_net.py:201-203)
__global__ void __launch_bounds__(256) add_bcast_kernel(uint4* __restrict__ x, const uint4* __restrict__ enc,
                                                        const int32_t* __restrict__ idx, int per_img8) {
  const int img = blockIdx.y;
  const uint4* __restrict__ e = enc + static_cast<size_t>(idx[img]) * per_img8;
  uint4* __restrict__ xi = x + static_cast<size_t>(img) * per_img8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_img8; i += gridDim.x * blockDim.x) {
    const uint4 a = xi[i], b = __ldg(e + i);
    uint4 o;
    o.x = pack_bf16(bf16_lo(a.x) + bf16_lo(b.x), bf16_hi(a.x) + bf16_hi(b.x));
    o.y = pack_bf16(bf16_lo(a.y) + bf16_lo(b.y), bf16_hi(a.y) + bf16_hi(b.y));
    o.z = pack_bf16(bf16_lo(a.z) + bf16_lo(b.z), bf16_hi(a.z) + bf16_hi(b.z));
    o.w = pack_bf16(bf16_lo(a.w) + bf16_lo(b.w), bf16_hi(a.w) + bf16_hi(b.w));
    xi[i] = o;
  }
}

// elements per image of layer li's output at the handle's input size
static size_t per_img_out(const io_net* net, int li) {
  return (static_cast<size_t>(net->h) >> (2 + li)) * (static_cast<size_t>(net->w) >> (2 + li)) * net->outs[li];
}
static bool default_geometry(const io_net* net) { return !net->plain && net->h == net->d && net->w == net->d; }

// Layer li's output is consumed only by (a) the next layer's conv1 -- computed on chip by the fused kernel that produces
// it -- and (b) the next layer's stride-2 1x1 downsample, which reads one pixel in four: the producing kernel then writes
// just those pixels, compactly ([img][h/2][w/2][C]), and the dual-source GEMM of the next layer reads them as a flat
// matrix.  INSTAORDER_SUBSAMPLE=0 restores the full tensor.
static bool subsample_out(const io_net* net, int li) {
  static const bool on = []() {
    const char* e = getenv("INSTAORDER_SUBSAMPLE");
    return e == nullptr || atoi(e) != 0;
  }();
  if (!on || !default_geometry(net) || net->keep_layers || net->inject_idx != nullptr || !net->fuse || !net->fuse_ds) return false;
  if (li + 1 >= net->n_layers || !((net->fuse_layers >> li) & 1)) return false;
  if (li == 1 && !net->cross_fuse) return false;     // layer2 -> layer3 crosses the phase boundary
  if (li > 1) return false;
  const int h = net->h >> (2 + li), w = net->w >> (2 + li);
  if ((h & 1) || (w % 32) != 0) return false;        // a 32-pixel epilogue slab must lie inside one image row
  return conv_fused_supported(net->widths[li], net->outs[li], net->widths[li + 1], nullptr);
}

// InstaDepthNet trunks: after the last block of layers 1..3 the encoder's feature of the pair's image is added in place
static int maybe_inject(io_net* net, Plan* plan, int li, bool layer_end, int b, int h, int w, __nv_bfloat16* dst,
                        int idx_off) {
  if (!layer_end || li > 2 || net->inject_idx == nullptr || net->inject[li] == nullptr) return IO_OK;
  Op a;
  a.kind = Op::ADD;
  a.src = net->inject[li];
  a.dst = dst;
  a.idx = net->inject_idx + idx_off;
  a.b = b; a.h = h; a.w = w; a.c = net->outs[li];
  a.bytes = 2.0 * b * h * w * net->outs[li] * 2.0;
  a.tag = (li + 1) * 100 + 98;
  plan->ops.push_back(a);
  return IO_OK;
}

// Appends the bottlenecks of layers [l0, l1) for `b` images whose input [b, h, w, C] is `src`.  Block outputs
// ping-pong between P0 and P1 (`src` may be one of them or a read-only third buffer); the very last output goes
// to `final_dst` if given.
static int build_blocks(io_net* net, Plan* plan, int l0, int l1, int b, int* h_io, int* w_io,
                        const __nv_bfloat16* src, __nv_bfloat16* P0, __nv_bfloat16* P1, __nv_bfloat16* T1,
                        __nv_bfloat16* T2, __nv_bfloat16* DS, __nv_bfloat16* final_dst,
                        const __nv_bfloat16** out_ptr, __nv_bfloat16* next_t1 = nullptr, bool t1_in = false,
                        long long keep_off = -1, int idx_off = 0, bool src_sub = false) {
  const int* blocks_ = net->blocks;
  size_t ci = 1;
  for (int li = 0; li < l0; ++li) ci += 3 * blocks_[li] + 1;
  const bool injecting = net->inject_idx != nullptr;
  const bool geom_ok = default_geometry(net);
  int h = *h_io, w = *w_io;
  // next_t1: where the conv1 output of the block FOLLOWING this plan's last one goes (phase A -> phase B fusion);
  // t1_in: this plan's first conv1 output has already been produced that way
  bool t1_ready = t1_in;   // the previous block's fused conv3 already produced this block's conv1 output
  for (int li = l0; li < l1; ++li) {
    for (int blk = 0; blk < blocks_[li]; ++blk) {
      const ConvW& c1 = net->convs[ci++];
      const ConvW& c2 = net->convs[ci++];
      const ConvW* ds = (blk == 0) ? &net->convs[ci++] : nullptr;
      const ConvW& c3 = net->convs[ci++];
      const int ho = h / c2.stride, wo = w / c2.stride;
      const bool last = (li == l1 - 1) && (blk == blocks_[li] - 1);
      const bool layer_end = blk == blocks_[li] - 1;
      __nv_bfloat16* dst = (last && final_dst) ? final_dst : (src == P0 ? P1 : P0);
      if (layer_end && !last && keep_off >= 0 && net->keep[li] != nullptr) dst = net->keep[li] + keep_off * per_img_out(net, li);
      if (!t1_ready) {
        Op o1; o1.kind = Op::CONV;
        if (int rc = conv_plan(&o1.p, &o1.bn_tile, ConvDesc{b, h, w, c1.cin, c1.cout, 1, 1}, src, c1.w, c1.bias,
                               nullptr, T1, 1)) return rc;
        o1.flops = 2.0 * b * h * w * c1.cin * c1.cout;
        o1.bytes = 2.0 * b * h * w * (c1.cin + c1.cout) + 2.0 * c1.cin * c1.cout;
        o1.tag = (li + 1) * 100 + blk * 10 + 1;
        plan->ops.push_back(o1);
      }
      t1_ready = false;
      Op o2; o2.kind = Op::CONV;
      const ConvDesc d2{b, h, w, c2.cin, c2.cout, 3, c2.stride};
      if (conv_row3_supported(d2)) {
        o2.kind = Op::CONV_ROW3;
        if (int rc = conv_row3_plan(&o2.hp, d2, T1, c2.w, c2.bias, T2, 1)) return rc;
      } else if (conv_halo_supported(d2)) {
        o2.kind = Op::CONV_HALO;
        if (int rc = conv_halo_plan(&o2.hp, d2, T1, c2.w, c2.bias, T2, 1)) return rc;
      } else if (tn_enabled() && conv_tn_supported(d2)) {
        o2.kind = Op::CONV_TN;
        if (int rc = conv_tn_plan(&o2.tp, d2, T1, c2.w, c2.bias, T2, 1)) return rc;
      } else if (int rc = conv_plan(&o2.p, &o2.bn_tile, d2, T1, c2.w, c2.bias, nullptr, T2, 1)) {
        return rc;
      }
      o2.flops = 2.0 * b * ho * wo * 9.0 * c2.cin * c2.cout;
      o2.bytes = 2.0 * b * (h * w * c2.cin + ho * wo * c2.cout) + 18.0 * c2.cin * c2.cout;
      o2.tag = (li + 1) * 100 + blk * 10 + 2;
      plan->ops.push_back(o2);
      const __nv_bfloat16* identity = src;
      // the next bottleneck's conv1 (same layer, or the first block of the next layer inside this plan: its conv1 is
      // a stride-1 1x1 over this block's output) can be computed from the block-output tile while it is on chip
      // (not across a layer boundary when encoder features are added in between)
      const bool has_next = (blk + 1 < blocks_[li]) || ((li + 1 < l1 || next_t1 != nullptr) && !injecting);
      __nv_bfloat16* t1_dst = ((blk + 1 < blocks_[li]) || (li + 1 < l1)) ? T1 : next_t1;
      const bool want_fuse = net->fuse && geom_ok && ((net->fuse_layers >> li) & 1) && has_next;
      if (ds && net->fuse_ds && geom_ok) {
        // block output = ReLU(conv3(T2) + downsample(src)) as one GEMM, K = [T2 channels | src channels]; the
        // identity tensor is neither written nor re-read
        // src_sub: `src` already holds only the pixels the stride-2 downsample reads
        const ConvDesc dsd = src_sub ? ConvDesc{b, ho, wo, ds->cin, ds->cout, 1, 1}
                                     : ConvDesc{b, h, w, ds->cin, ds->cout, 1, ds->stride};
        src_sub = false;
        if (want_fuse && conv_fused_supported(c3.cin, c3.cout, net->convs[ci].cout, &dsd)) {
          const ConvW& n1 = net->convs[ci];
          Op of; of.kind = Op::FUSED;
          if (int rc = conv_fused_plan(&of.fp, b * ho * wo, c3.cin, c3.cout, n1.cout, T2, ds->wcat, ds->bias_cat,
                                       nullptr, dst, n1.w, n1.bias, t1_dst, &dsd, src)) return rc;
          of.flops = 2.0 * b * ho * wo * ((static_cast<double>(c3.cin) + ds->cin) * c3.cout +
                                          static_cast<double>(n1.cin) * n1.cout);
          of.bytes = 2.0 * b * ho * wo * (c3.cin + ds->cin + c3.cout + n1.cout) + 2.0 * (c3.cin + ds->cin) * c3.cout +
                     2.0 * n1.cin * n1.cout;
          of.tag = (li + 1) * 100 + blk * 10 + 7;
          plan->ops.push_back(of);
          t1_ready = true;
        } else {
          Op o3; o3.kind = Op::CONV;
          if (int rc = conv_plan_dual(&o3.p, &o3.bn_tile, dsd, src, T2, c3.cin, ds->wcat, ds->bias_cat, dst, 1)) return rc;
          o3.flops = 2.0 * b * ho * wo * (static_cast<double>(c3.cin) + ds->cin) * c3.cout;
          o3.bytes = 2.0 * b * ho * wo * (c3.cin + ds->cin + c3.cout) + 2.0 * (c3.cin + ds->cin) * c3.cout;
          o3.tag = (li + 1) * 100 + blk * 10 + 6;
          plan->ops.push_back(o3);
        }
        src = dst;
        h = ho; w = wo;
        if (int rc = maybe_inject(net, plan, li, layer_end, b, h, w, dst, idx_off)) return rc;
        continue;
      }
      if (ds) {
        Op od; od.kind = Op::CONV;
        if (int rc = conv_plan(&od.p, &od.bn_tile, ConvDesc{b, h, w, ds->cin, ds->cout, 1, ds->stride}, src, ds->w,
                               ds->bias, nullptr, DS, 0)) return rc;
        od.flops = 2.0 * b * ho * wo * ds->cin * ds->cout;
        od.bytes = 2.0 * b * (h * w * ds->cin + ho * wo * ds->cout) + 2.0 * ds->cin * ds->cout;
        od.tag = (li + 1) * 100 + blk * 10 + 4;
        plan->ops.push_back(od);
        identity = DS;
      }
      const bool fuse_next = want_fuse && conv_fused_supported(c3.cin, c3.cout, net->convs[ci].cout, nullptr);
      if (fuse_next) {
        const ConvW& n1 = net->convs[ci];   // next block's conv1
        Op of; of.kind = Op::FUSED;
        if (int rc = conv_fused_plan(&of.fp, b * ho * wo, c3.cin, c3.cout, n1.cout, T2, c3.w, c3.bias, identity, dst,
                                     n1.w, n1.bias, t1_dst, nullptr, nullptr)) return rc;
        of.flops = 2.0 * b * ho * wo * (static_cast<double>(c3.cin) * c3.cout + static_cast<double>(n1.cin) * n1.cout);
        of.bytes = 2.0 * b * ho * wo * (c3.cin + 2 * c3.cout + n1.cout) + 2.0 * c3.cin * c3.cout + 2.0 * n1.cin * n1.cout;
        of.tag = (li + 1) * 100 + blk * 10 + 5;
        if (layer_end && subsample_out(net, li) && ((li + 1 < l1) || next_t1 != nullptr)) {
          conv_fused_set_subsampled(&of.fp, ho, wo);
          of.bytes -= 2.0 * b * ho * wo * c3.cout * 0.75;
          of.tag += 3;   // LB8: conv3 + identity + next conv1, sub-sampled block output
          src_sub = true;
        }
        plan->ops.push_back(of);
        t1_ready = true;
      } else {
        Op o3; o3.kind = Op::CONV;
        if (int rc = conv_plan(&o3.p, &o3.bn_tile, ConvDesc{b, ho, wo, c3.cin, c3.cout, 1, 1}, T2, c3.w, c3.bias,
                               identity, dst, 1)) return rc;
        o3.flops = 2.0 * b * ho * wo * c3.cin * c3.cout;
        o3.bytes = 2.0 * b * ho * wo * (c3.cin + 2 * c3.cout) + 2.0 * c3.cin * c3.cout;
        o3.tag = (li + 1) * 100 + blk * 10 + 3;
        plan->ops.push_back(o3);
      }
      src = dst;  // the next block reads what was just written
      h = ho; w = wo;
      if (int rc = maybe_inject(net, plan, li, layer_end, b, h, w, dst, idx_off)) return rc;
    }
  }
  *h_io = h; *w_io = w;
  *out_ptr = src;
  return IO_OK;
}

// phase A: stem + max-pool + layer1 + layer2 for `pa` pairs; layer2's output goes to `dst` ([2*pa, D/8, D/8, 512])
static int build_plan_a(io_net* net, int pa, int a0, __nv_bfloat16* dst, __nv_bfloat16* next_t1, Plan* plan) {
  const int dirs = net->single_dir ? 1 : 2;
  const int b = dirs * pa, d = net->d;
  plan->ops.clear();
  __nv_bfloat16 *X = net->buf[0], *Y = net->buf[1];
  Op op;
  if (!default_geometry(net)) {
    // `orig` mode: h x w network input, one launch per convolution (general tilers: partial tiles, wide rows)
    const int H = net->h, W = net->w;
    op.kind = Op::STEM;
    op.flops = 2.0 * b * (H / 2) * (W / 2) * 49.0 * 5.0 * 64.0;
    op.bytes = static_cast<double>(io_pair_tensor_bytes_hw(pa, H, W)) + 2.0 * b * (H / 2) * (W / 2) * 64;
    op.tag = 1;
    plan->ops.push_back(op);
    Op pool;
    pool.kind = Op::POOL;
    pool.src = X; pool.dst = Y; pool.b = b; pool.h = H / 2; pool.w = W / 2; pool.c = 64;
    pool.bytes = 2.0 * b * (H / 2) * (W / 2) * 64 * 1.25;
    pool.tag = 2;
    plan->ops.push_back(pool);
    int h = H / 4, w = W / 4;
    const __nv_bfloat16* out = nullptr;
    return build_blocks(net, plan, 0, 2, b, &h, &w, Y, X, Y, net->buf[2], net->buf[3], net->buf[4], dst, &out, nullptr,
                        false, -1, dirs * a0);
  }
  if (stem_pool_supported(d)) {
    // conv1 + BN + ReLU + max-pool in one launch: only the pooled tensor is written (maps are rebuilt per call: the
    // pair tensor belongs to the caller)
    op.kind = Op::STEM_POOL;
    op.flops = 2.0 * b * (d / 2) * (d / 2) * 49.0 * 5.0 * 64.0;
    op.bytes = static_cast<double>(io_pair_tensor_bytes(pa, d)) + 2.0 * b * (d / 4) * (d / 4) * 64;
    op.tag = 3;
    plan->ops.push_back(op);
    int h = d / 4, w = d / 4;
    const __nv_bfloat16* out = nullptr;
    return build_blocks(net, plan, 0, 2, b, &h, &w, Y, X, Y, net->buf[2], net->buf[3], net->buf[4], dst, &out, next_t1,
                        false, net->keep_layers ? static_cast<long long>(dirs) * a0 : -1, dirs * a0);
  }
  op.kind = (tn_enabled() && stem_tn_supported(d)) ? Op::STEM_TN : Op::STEM;   // maps are rebuilt per call (the
                                                                               // pair tensor belongs to the caller)
  op.flops = 2.0 * b * (d / 2) * (d / 2) * 49.0 * 5.0 * 64.0;
  op.bytes = static_cast<double>(io_pair_tensor_bytes(pa, d)) + 2.0 * b * (d / 2) * (d / 2) * 64;
  op.tag = 1;
  plan->ops.push_back(op);
  Op pool;
  pool.kind = Op::POOL;
  pool.src = X; pool.dst = Y; pool.b = b; pool.h = d / 2; pool.w = d / 2; pool.c = 64;
  pool.bytes = 2.0 * b * (d / 2) * (d / 2) * 64 * 1.25;
  pool.tag = 2;
  plan->ops.push_back(pool);
  int h = d / 4, w = d / 4;
  const __nv_bfloat16* out = nullptr;
  return build_blocks(net, plan, 0, 2, b, &h, &w, Y, X, Y, net->buf[2], net->buf[3], net->buf[4], dst, &out, next_t1,
                      false, net->keep_layers ? static_cast<long long>(dirs) * a0 : -1, dirs * a0);
}

// phase B: layer3 + layer4 for `pb` pairs reading the big layer2-output buffer
static int build_plan_b(io_net* net, int pb, Plan* plan) {
  const int b = (net->single_dir ? 1 : 2) * pb;
  plan->ops.clear();
  int h = net->h / 8, w = net->w / 8;
  const __nv_bfloat16* out = nullptr;
  // the first block reads `big` (kept intact) and writes bufb[0]; afterwards bufb[0] / bufb[1] ping-pong
  int rc = build_blocks(net, plan, 2, net->n_layers, b, &h, &w, net->big, net->bufb[0], net->bufb[1], net->bufb[2],
                        net->bufb[3], net->bufb[4], net->keep_layers ? net->keep[net->n_layers - 1] : nullptr, &out,
                        nullptr, net->cross_fuse && default_geometry(net), net->keep_layers ? 0 : -1, 0,
                        net->cross_fuse && subsample_out(net, 1));
  plan->feat = out;
  plan->hw_final = h * w;
  return rc;
}

}  // namespace io

using namespace io;

extern "C" int io_net_create_arch(const int32_t* widths, const int32_t* outs, const int32_t* blocks, int n_layers,
                                  int keep_layers, const int32_t* num_classes, int n_heads, int input_size,
                                  int max_pairs, io_net_t** out) {
  IO_REQUIRE(widths && outs && blocks && out, "io_net_create_arch: null pointer");
  IO_REQUIRE(n_layers == 3 || n_layers == 4, "io_net_create_arch: n_layers %d (3 or 4)", n_layers);
  IO_REQUIRE(n_heads >= 0 && n_heads <= 2 && (n_heads == 0 || (num_classes && n_layers == 4)),
             "io_net_create_arch: n_heads must be 0 (feature extractor), 1 (fc) or 2 (fc_occ + fc_depth)");
  IO_REQUIRE(input_size >= 64 && input_size <= 1024 && input_size % 32 == 0,
             "io_net_create: input_size %d (multiple of 32 in [64, 1024])", input_size);
  IO_REQUIRE(max_pairs >= 1, "io_net_create: max_pairs %d", max_pairs);
  int dev_count = 0;
  IO_CUDA(cudaGetDeviceCount(&dev_count));
  std::unique_ptr<io_net> net(new io_net());
  for (int i = 0; i < 4; ++i) {
    IO_REQUIRE(i >= n_layers || (widths[i] % 64 == 0 && outs[i] % 64 == 0 && widths[i] >= 64 && outs[i] >= 64 &&
                                 outs[i] <= 2048 && blocks[i] >= 1),
               "io_net_create_arch: layer %d: width %d, out %d, blocks %d", i + 1, widths[i], outs[i], blocks[i]);
    net->widths[i] = widths[i]; net->outs[i] = outs[i]; net->blocks[i] = blocks[i];
  }
  IO_REQUIRE(n_heads == 0 || outs[3] == 2048, "io_net_create_arch: the head kernel expects 2048 features");
  net->n_layers = n_layers;
  net->keep_layers = keep_layers != 0;
  net->single_dir = net->keep_layers && n_heads == 0;
  net->n_heads = n_heads;
  for (int i = 0; i < n_heads; ++i) {
    IO_REQUIRE(num_classes[i] >= 1 && num_classes[i] <= 4, "io_net_create: num_classes[%d] = %d", i, num_classes[i]);
    net->num_classes[i] = num_classes[i];
    net->k_total += num_classes[i];
  }
  net->d = input_size;
  net->h = net->w = input_size;
  net->max_pairs = max_pairs;
  int chunk_a = 256, chunk_b = 256;
  if (const char* e = getenv("INSTAORDER_CHUNK_A")) chunk_a = atoi(e) > 0 ? atoi(e) : chunk_a;
  if (const char* e = getenv("INSTAORDER_CHUNK_B")) chunk_b = atoi(e) > 0 ? atoi(e) : chunk_b;
  if (const char* e = getenv("INSTAORDER_FUSE")) net->fuse = atoi(e) != 0;
  if (const char* e = getenv("INSTAORDER_FUSE_DS")) net->fuse_ds = atoi(e) != 0;
  if (const char* e = getenv("INSTAORDER_FUSE_LAYERS")) net->fuse_layers = atoi(e);
  net->cross_fuse = net->fuse && (net->fuse_layers & 2) && !net->keep_layers &&
                    conv_fused_supported(net->widths[1], net->outs[1], net->widths[2], nullptr);
  if (const char* e = getenv("INSTAORDER_FUSE_CROSS")) net->cross_fuse = net->cross_fuse && atoi(e) != 0;
  net->chunk_b = std::min(chunk_b, max_pairs);
  net->chunk_a = std::min(chunk_a, net->chunk_b);
  build_conv_list(net.get());
  // device weights
  for (size_t i = 1; i < net->convs.size(); ++i) {
    ConvW& c = net->convs[i];
    IO_CUDA(cudaMalloc(&c.w, static_cast<size_t>(c.cout) * c.k * c.k * c.cin * 2));
    IO_CUDA(cudaMalloc(&c.bias, static_cast<size_t>(c.cout) * 4));
    if (c.cat_k > 0) {
      IO_CUDA(cudaMalloc(&c.wcat, static_cast<size_t>(c.cout) * c.cat_k * 2));
      IO_CUDA(cudaMalloc(&c.bias_cat, static_cast<size_t>(c.cout) * 4));
    }
  }
  IO_CUDA(cudaMalloc(&net->stem_w, 128 * 448 * 2));
  IO_CUDA(cudaMalloc(&net->stem_w2, 128 * 448 * 2));
  IO_CUDA(cudaMalloc(&net->stem_bias, 128 * 4));
  if (n_heads > 0) {
    IO_CUDA(cudaMalloc(&net->fc_w, static_cast<size_t>(net->k_total) * 2048 * 4));
    IO_CUDA(cudaMalloc(&net->fc_b, static_cast<size_t>(net->k_total) * 4));
  }
  // activation buffers, sized by the largest tensor of each phase (elements per image):
  //   phase A (stem .. layer2): stem output, layer1 tensors at D/4, layer2.0's conv1 output (still D/4), layer2 at D/8
  //   phase B (layer3 ..):      layer3.0's conv1 output (still D/8), layer3 at D/16, layer4.0's conv1 at D/16, layer4
  // ResNet-50: 16 D^2 and 4 D^2 elements per image (the stem / layer3.0 conv1 outputs)
  const size_t d = input_size;
  const int* W = net->widths; const int* O = net->outs;
  size_t ea_img = (d / 2) * (d / 2) * 64;
  ea_img = std::max(ea_img, (d / 4) * (d / 4) * static_cast<size_t>(std::max(std::max(O[0], W[0]), W[1])));
  ea_img = std::max(ea_img, (d / 8) * (d / 8) * static_cast<size_t>(std::max(O[1], W[1])));
  size_t eb_img = (d / 8) * (d / 8) * static_cast<size_t>(W[2]);
  eb_img = std::max(eb_img, (d / 16) * (d / 16) * static_cast<size_t>(std::max(O[2], W[2])));
  if (n_layers == 4) {
    eb_img = std::max(eb_img, (d / 16) * (d / 16) * static_cast<size_t>(W[3]));
    eb_img = std::max(eb_img, (d / 32) * (d / 32) * static_cast<size_t>(std::max(O[3], W[3])));
  }
  const size_t ea = static_cast<size_t>(2 * net->chunk_a) * ea_img;
  const size_t eb = static_cast<size_t>(2 * net->chunk_b) * eb_img;
  for (int i = 0; i < 5; ++i) IO_CUDA(cudaMalloc(&net->buf[i], ea * 2));
  for (int i = 0; i < 5; ++i) IO_CUDA(cudaMalloc(&net->bufb[i], eb * 2));
  IO_CUDA(cudaMalloc(&net->big, static_cast<size_t>(2 * net->chunk_b) * per_img_out(net.get(), 1) * 2));
  if (net->keep_layers) {
    for (int li = 0; li < n_layers; ++li)
      if (li != 1)   // layer2's output lives in `big`
        IO_CUDA(cudaMalloc(&net->keep[li], static_cast<size_t>(2 * net->chunk_b) * per_img_out(net.get(), li) * 2));
  }
  *out = net.release();
  return IO_OK;
}

extern "C" int io_net_create(const int32_t* num_classes, int n_heads, int input_size, int max_pairs, io_net_t** out) {
  IO_REQUIRE(num_classes && out, "io_net_create: null pointer");
  IO_REQUIRE(n_heads == 1 || n_heads == 2, "io_net_create: n_heads must be 1 (fc) or 2 (fc_occ + fc_depth)");
  const int32_t widths[4] = {64, 128, 256, 512}, outs[4] = {256, 512, 1024, 2048}, blocks[4] = {3, 4, 6, 3};
  return io_net_create_arch(widths, outs, blocks, 4, 0, num_classes, n_heads, input_size, max_pairs, out);
}

// Output of layer `layer` (0-based) of a keep_layers handle: [images, D >> (2 + layer), same, outs[layer]] bf16; images =
// the pairs of the call for a head-less handle (RGB-only encoder: one direction), else 2 * pairs (image 2p + direction).
extern "C" int io_net_feature(io_net_t* net, int layer, void** ptr, int64_t* elems_per_image) {
  IO_REQUIRE(net && ptr && elems_per_image && net->keep_layers && layer >= 0 && layer < net->n_layers,
             "io_net_feature: bad arguments");
  *ptr = layer == 1 ? net->big : net->keep[layer];
  *elems_per_image = static_cast<int64_t>(per_img_out(net, layer));
  return IO_OK;
}

// x += f_l[idx[image]] after layers 1..3 (InstaDepthNet trunks, midas_net.py:201-203).  f_l: [n, h_l, w_l, outs[l]] bf16 with
// the handle's own layer geometry; idx_dev: int32 [2 * max_pairs] (one entry per image = pair direction), read at run
// time.  The pointers are baked into the cached launch plans: changing them drops the plans.
extern "C" int io_net_set_inject(io_net_t* net, const void* f1, const void* f2, const void* f3, const int32_t* idx_dev) {
  IO_REQUIRE(net && ((f1 && f2 && f3 && idx_dev) || (!f1 && !f2 && !f3 && !idx_dev)), "io_net_set_inject: bad arguments");
  IO_REQUIRE(net->max_pairs <= net->chunk_b, "io_net_set_inject: max_pairs %d > chunk %d", net->max_pairs, net->chunk_b);
  net->inject[0] = reinterpret_cast<const __nv_bfloat16*>(f1);
  net->inject[1] = reinterpret_cast<const __nv_bfloat16*>(f2);
  net->inject[2] = reinterpret_cast<const __nv_bfloat16*>(f3);
  net->inject_idx = idx_dev;
  if (idx_dev) net->cross_fuse = false;
  net->plans_a.clear();
  net->plans_b.clear();
  return IO_OK;
}

extern "C" int io_net_destroy(io_net_t* net) {
  if (!net) return IO_OK;
  for (auto& c : net->convs) {
    cudaFree(c.w);
    cudaFree(c.bias);
    cudaFree(c.wcat);
    cudaFree(c.bias_cat);
  }
  cudaFree(net->stem_w);
  cudaFree(net->stem_w2);
  cudaFree(net->stem_bias);
  cudaFree(net->fc_w);
  cudaFree(net->fc_b);
  for (int i = 0; i < 5; ++i) cudaFree(net->buf[i]);
  for (int i = 0; i < 5; ++i) cudaFree(net->bufb[i]);
  for (int i = 0; i < 4; ++i) cudaFree(net->keep[i]);
  cudaFree(net->big);
  for (cudaEvent_t e : net->ev) cudaEventDestroy(e);
  delete net;
  return IO_OK;
}

extern "C" int io_net_load_state(io_net_t* net, const char* const* names, const float* const* ptrs,
                                 const int64_t* numels, int n) {
  IO_REQUIRE(net && names && ptrs && numels, "io_net_load_state: null pointer");
  std::map<std::string, std::pair<const float*, int64_t>> sd;
  for (int i = 0; i < n; ++i) sd[names[i]] = {ptrs[i], numels[i]};
  auto get = [&](const std::string& key, int64_t numel, const float** out) -> int {
    auto it = sd.find(key);
    if (it == sd.end()) {
      set_error("io_net_load_state: missing key '%s'", key.c_str());
      return IO_ERR_ARG;
    }
    if (it->second.second != numel) {
      set_error("io_net_load_state: key '%s' has %lld elements, expected %lld", key.c_str(),
                static_cast<long long>(it->second.second), static_cast<long long>(numel));
      return IO_ERR_ARG;
    }
    *out = it->second.first;
    return IO_OK;
  };
  const double eps = 1e-5;  // nn.BatchNorm2d default
  std::vector<uint16_t> ds_w;   // folded downsample weights / bias, waiting for the conv3 that follows in the list
  std::vector<float> ds_bias;
  for (size_t ci = 0; ci < net->convs.size(); ++ci) {
    ConvW& c = net->convs[ci];
    const float *w, *g, *bt, *mu, *var;
    const int64_t wn = static_cast<int64_t>(c.cout) * c.cin * c.k * c.k;
    if (int rc = get(c.name + ".weight", wn, &w)) return rc;
    if (int rc = get(c.bn + ".weight", c.cout, &g)) return rc;
    if (int rc = get(c.bn + ".bias", c.cout, &bt)) return rc;
    if (int rc = get(c.bn + ".running_mean", c.cout, &mu)) return rc;
    if (int rc = get(c.bn + ".running_var", c.cout, &var)) return rc;
    std::vector<float> scale(c.cout), bias(c.cout);
    for (int co = 0; co < c.cout; ++co) {
      const double s = static_cast<double>(g[co]) / sqrt(static_cast<double>(var[co]) + eps);
      scale[co] = static_cast<float>(s);
      bias[co] = static_cast<float>(static_cast<double>(bt[co]) - static_cast<double>(mu[co]) * s);
    }
    if (ci == 0) {
      // stem: rows [0,64) = direction (A,B); rows [64,128) = direction (B,A) (input channels 0 and 1 exchanged).
      // K index = r * 64 + s * 8 + c with s < 7 real taps (+1 zero tap) and c < 5 real channels (+3 zero).
      std::vector<uint16_t> pk(128 * 448, 0), pk2(128 * 448, 0);
      std::vector<float> b2(128);
      for (int dir = 0; dir < 2; ++dir)
        for (int co = 0; co < 64; ++co) {
          b2[dir * 64 + co] = bias[co];
          for (int cc = 0; cc < 5; ++cc) {
            const int src_c = (dir == 1 && cc < 2) ? 1 - cc : cc;
            for (int r = 0; r < 7; ++r)
              for (int s = 0; s < 7; ++s) {
                const float v = w[((static_cast<size_t>(co) * 5 + src_c) * 7 + r) * 7 + s] * scale[co];
                pk[static_cast<size_t>(dir * 64 + co) * 448 + r * 64 + s * 8 + cc] = bf16_bits(v);
                pk2[static_cast<size_t>(dir * 64 + co) * 448 + stem_pool_pack_k(r, s, cc)] = bf16_bits(v);
              }
          }
        }
      IO_CUDA(cudaMemcpy(net->stem_w, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
      IO_CUDA(cudaMemcpy(net->stem_w2, pk2.data(), pk2.size() * 2, cudaMemcpyHostToDevice));
      IO_CUDA(cudaMemcpy(net->stem_bias, b2.data(), b2.size() * 4, cudaMemcpyHostToDevice));
    } else {
      const int kk = c.k * c.k;
      std::vector<uint16_t> pk(static_cast<size_t>(wn));
      for (int co = 0; co < c.cout; ++co)
        for (int cc = 0; cc < c.cin; ++cc)
          for (int t = 0; t < kk; ++t) {
            const float v = w[(static_cast<size_t>(co) * c.cin + cc) * kk + t] * scale[co];
            pk[(static_cast<size_t>(co) * kk + t) * c.cin + cc] = bf16_bits(v);
          }
      IO_CUDA(cudaMemcpy(c.w, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
      IO_CUDA(cudaMemcpy(c.bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
      if (c.wcat != nullptr) {
        ds_w = pk;
        ds_bias = bias;
      } else if (ci >= 1 && net->convs[ci - 1].wcat != nullptr) {
        // conv3 right after a downsample conv: rows = [conv3 weights (cin = cmid) | downsample weights]
        ConvW& dsc = net->convs[ci - 1];
        const size_t kc = static_cast<size_t>(c.cin) + dsc.cin;
        std::vector<uint16_t> cat(static_cast<size_t>(c.cout) * kc);
        std::vector<float> bsum(c.cout);
        for (int co = 0; co < c.cout; ++co) {
          memcpy(&cat[co * kc], &pk[static_cast<size_t>(co) * c.cin], static_cast<size_t>(c.cin) * 2);
          memcpy(&cat[co * kc + c.cin], &ds_w[static_cast<size_t>(co) * dsc.cin], static_cast<size_t>(dsc.cin) * 2);
          bsum[co] = bias[co] + ds_bias[co];
        }
        IO_CUDA(cudaMemcpy(dsc.wcat, cat.data(), cat.size() * 2, cudaMemcpyHostToDevice));
        IO_CUDA(cudaMemcpy(dsc.bias_cat, bsum.data(), bsum.size() * 4, cudaMemcpyHostToDevice));
      }
    }
  }
  if (net->n_heads == 0) {
    net->loaded = true;
    return IO_OK;
  }
  const char* heads1[1] = {"fc"};
  const char* heads2[2] = {"fc_occ", "fc_depth"};
  const char* const* heads = net->n_heads == 1 ? heads1 : heads2;
  int row = 0;
  for (int hI = 0; hI < net->n_heads; ++hI) {
    const float *w, *b;
    const int k = net->num_classes[hI];
    if (int rc = get(std::string(heads[hI]) + ".weight", static_cast<int64_t>(k) * 2048, &w)) return rc;
    if (int rc = get(std::string(heads[hI]) + ".bias", k, &b)) return rc;
    IO_CUDA(cudaMemcpy(net->fc_w + static_cast<size_t>(row) * 2048, w, static_cast<size_t>(k) * 2048 * 4,
                       cudaMemcpyHostToDevice));
    IO_CUDA(cudaMemcpy(net->fc_b + row, b, static_cast<size_t>(k) * 4, cudaMemcpyHostToDevice));
    row += k;
  }
  net->loaded = true;
  return IO_OK;
}

static int forward_pairs(io_net_t* net, const void* pair_tensor, int p, float* logits, void* stream_);

extern "C" int io_net_forward_pairs(io_net_t* net, const void* pair_tensor, int p, float* logits, void* stream_) {
  IO_REQUIRE(net, "io_net_forward_pairs: null handle");
  if (net->plain || net->h != net->d || net->w != net->d) {   // back from an `orig`-mode call: drop its plans
    net->plain = false;
    net->h = net->w = net->d;
    net->plans_a.clear();
    net->plans_b.clear();
  }
  return forward_pairs(net, pair_tensor, p, logits, stream_);
}

// `orig` mode (reference inference.py:401-408): the pairs of ONE image whose network input is h x w (its own size rounded
// to multiples of 32), pair tensor [p][h + 6][pitch(w)][8].  h, w <= the handle's input_size; the convolutions run through
// the general tilers, one launch each (no cross-convolution fusion).  Plans are rebuilt when the geometry changes.
extern "C" int io_net_forward_pairs_hw(io_net_t* net, const void* pair_tensor, int p, int h, int w, float* logits,
                                       void* stream_) {
  IO_REQUIRE(net, "io_net_forward_pairs_hw: null handle");
  IO_REQUIRE(h >= 32 && w >= 32 && h % 32 == 0 && w % 32 == 0 && h <= net->d && w <= net->d,
             "io_net_forward_pairs_hw: network input %d x %d (multiples of 32, at most the handle's %d x %d)", h, w,
             net->d, net->d);
  if (!net->plain || h != net->h || w != net->w) {
    net->plain = true;
    net->h = h;
    net->w = w;
    net->plans_a.clear();
    net->plans_b.clear();
  }
  return forward_pairs(net, pair_tensor, p, logits, stream_);
}

static int forward_pairs(io_net_t* net, const void* pair_tensor, int p, float* logits, void* stream_) {
  IO_REQUIRE(net && pair_tensor && (logits || net->n_heads == 0), "io_net_forward_pairs: null pointer");
  if (!net->loaded) {
    set_error("io_net_forward_pairs: no weights loaded (call io_net_load_state first)");
    return IO_ERR_STATE;
  }
  IO_REQUIRE(p >= 0 && p <= net->max_pairs, "io_net_forward_pairs: %d pairs (handle was created for <= %d)", p,
             net->max_pairs);
  cudaStream_t stream = as_stream(stream_);
  net->last_launches = 0;
  net->prof_kind.clear();
  net->prof_flops.clear();
  net->prof_bytes.clear();
  net->prof_tag.clear();
  auto mark = [&](int kind, double flops, bool begin, double bytes = 0.0, int tag = 0) -> int {
    if (!net->profile) return IO_OK;
    const size_t idx = 2 * net->prof_kind.size() + (begin ? 0 : 1);
    while (net->ev.size() <= idx) {
      cudaEvent_t e;
      IO_CUDA(cudaEventCreate(&e));
      net->ev.push_back(e);
    }
    IO_CUDA(cudaEventRecord(net->ev[idx], stream));
    if (!begin) {
      net->prof_kind.push_back(kind);
      net->prof_flops.push_back(flops);
      net->prof_bytes.push_back(bytes);
      net->prof_tag.push_back(tag);
    }
    return IO_OK;
  };
  const int64_t pair_bytes = io_pair_tensor_bytes_hw(1, net->h, net->w);
  const size_t l2_elems_per_pair = (net->single_dir ? 1 : 2) * per_img_out(net, 1) / (subsample_out(net, 1) ? 4 : 1);
  const size_t t1b_elems_per_pair = static_cast<size_t>(2) * (net->h / 8) * (net->w / 8) * net->widths[2];
  const bool cross = net->cross_fuse && default_geometry(net);
  auto run_ops = [&](Plan& plan, const uint8_t* pair_ptr, int pa) -> int {
    for (Op& op : plan.ops) {
      int rc = mark(static_cast<int>(op.kind), op.flops, true);
      if (rc) return rc;
      switch (op.kind) {
        case Op::STEM:
          rc = stem_plan_hw(&op.p, &op.bn_tile, pa, net->h, net->w, pair_ptr, net->stem_w, net->stem_bias, net->buf[0]);
          if (!rc && net->single_dir) { op.p.img_mul = 1; op.p.split_row_off = pa * op.p.hw_out; }
          if (!rc) rc = conv_tc_launch(op.p, op.bn_tile, stream);
          break;
        case Op::POOL:
          rc = maxpool_launch(op.src, op.dst, op.b, op.h, op.w, op.c, stream);
          break;
        case Op::CONV:
          rc = conv_tc_launch(op.p, op.bn_tile, stream);
          break;
        case Op::FUSED:
          rc = conv_fused_launch(op.fp, stream);
          break;
        case Op::CONV_TN:
          rc = conv_tn_launch(op.tp, stream);
          break;
        case Op::CONV_HALO:
          rc = conv_halo_launch(op.hp, stream);
          break;
        case Op::CONV_ROW3:
          rc = conv_row3_launch(op.hp, stream);
          break;
        case Op::ADD: {
          const int per8 = op.h * op.w * op.c / 8;
          dim3 grid(std::min((per8 + 255) / 256, 64), op.b);
          add_bcast_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<uint4*>(op.dst),
                                                     reinterpret_cast<const uint4*>(op.src), op.idx, per8);
          rc = cudaGetLastError() == cudaSuccess ? IO_OK : IO_ERR_CUDA;
          break;
        }
        case Op::STEM_POOL:
          rc = stem_pool_plan(&op.sp, pa, net->d, pair_ptr, net->stem_w2, net->stem_bias, net->buf[1]);
          if (!rc && net->single_dir) { op.sp.img_mul = 1; op.sp.split_row_off = pa * (net->d / 4) * (net->d / 4); }
          if (!rc) rc = stem_pool_launch(op.sp, stream);
          break;
        case Op::STEM_TN:
          rc = stem_tn_plan(&op.tp, pa, net->d, pair_ptr, net->stem_w, net->stem_bias, net->buf[0]);
          if (!rc && net->single_dir) { op.tp.img_mul = 1; op.tp.split_row_off = pa * op.tp.hw_out; }
          if (!rc) rc = conv_tn_launch(op.tp, stream);
          break;
        default:
          break;
      }
      if (rc) return rc;
      if ((rc = mark(static_cast<int>(op.kind), op.flops, false, op.bytes, op.tag))) return rc;
      ++net->last_launches;
    }
    return IO_OK;
  };
  for (int b0 = 0; b0 < p; b0 += net->chunk_b) {
    const int pb = std::min(net->chunk_b, p - b0);
    for (int a0 = 0; a0 < pb; a0 += net->chunk_a) {
      const int pa = std::min(net->chunk_a, pb - a0);
      auto key = std::make_pair(pa, a0);
      auto it = net->plans_a.find(key);
      if (it == net->plans_a.end()) {
        std::unique_ptr<Plan> plan(new Plan());
        __nv_bfloat16* next_t1 = cross ? net->bufb[2] + static_cast<size_t>(a0) * t1b_elems_per_pair : nullptr;
        if (int rc = build_plan_a(net, pa, a0, net->big + static_cast<size_t>(a0) * l2_elems_per_pair, next_t1, plan.get())) return rc;
        it = net->plans_a.emplace(key, std::move(plan)).first;
      }
      const uint8_t* pair_ptr = reinterpret_cast<const uint8_t*>(pair_tensor) + static_cast<int64_t>(b0 + a0) * pair_bytes;
      if (int rc = run_ops(*it->second, pair_ptr, pa)) return rc;
    }
    auto itb = net->plans_b.find(pb);
    if (itb == net->plans_b.end()) {
      std::unique_ptr<Plan> plan(new Plan());
      if (int rc = build_plan_b(net, pb, plan.get())) return rc;
      itb = net->plans_b.emplace(pb, std::move(plan)).first;
    }
    Plan& planb = *itb->second;
    if (int rc = run_ops(planb, nullptr, pb)) return rc;
    if (net->n_heads == 0) continue;   // feature extractor: the layer outputs are the result
    if (int rc = mark(3, 2.0 * 2 * pb * 2048.0 * net->k_total, true)) return rc;
    if (int rc = tail_launch(planb.feat, planb.hw_final, pb, net->fc_w, net->fc_b, net->k_total,
                             logits + static_cast<size_t>(b0) * 2 * net->k_total, stream))
      return rc;
    if (int rc = mark(3, 2.0 * 2 * pb * 2048.0 * net->k_total, false)) return rc;
    ++net->last_launches;
  }
  return IO_OK;
}

// Building block of io_net_forward_pairs for 256 x 256 inputs, exposed for the layer-level parity test: conv1 (7x7
// stride 2, weights with the BN scale already folded in) + bias + ReLU + MaxPool2d(3, 2, 1) of BOTH directions of every
// pair, from the pair tensor.  w_host: [64][5][7][7] fp32, bias_host: [64] fp32 (host pointers; packed and uploaded
// here), out_dev: [2 * pairs][d/4][d/4][64] bf16, image 2 * pair + direction.
extern "C" int io_stem_pool(const void* pair_tensor_dev, int pairs, int d, const float* w_host, const float* bias_host,
                            void* out_dev, void* stream_) {
  IO_REQUIRE(pair_tensor_dev && w_host && bias_host && out_dev && pairs >= 1, "io_stem_pool: bad arguments");
  IO_REQUIRE(d == 256, "io_stem_pool: input size %d (the fused kernel is built for 256)", d);
  std::vector<uint16_t> pk(128 * 448, 0);
  std::vector<float> b2(128);
  for (int dir = 0; dir < 2; ++dir)
    for (int co = 0; co < 64; ++co) {
      b2[dir * 64 + co] = bias_host[co];
      for (int cc = 0; cc < 5; ++cc) {
        const int src_c = (dir == 1 && cc < 2) ? 1 - cc : cc;
        for (int r = 0; r < 7; ++r)
          for (int s = 0; s < 7; ++s)
            pk[static_cast<size_t>(dir * 64 + co) * 448 + stem_pool_pack_k(r, s, cc)] =
                bf16_bits(w_host[((static_cast<size_t>(co) * 5 + src_c) * 7 + r) * 7 + s]);
      }
    }
  __nv_bfloat16* w_dev = nullptr;
  float* b_dev = nullptr;
  IO_CUDA(cudaMalloc(&w_dev, pk.size() * 2));
  IO_CUDA(cudaMalloc(&b_dev, b2.size() * 4));
  IO_CUDA(cudaMemcpy(w_dev, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
  IO_CUDA(cudaMemcpy(b_dev, b2.data(), b2.size() * 4, cudaMemcpyHostToDevice));
  StemPoolParams sp;
  HaloParams hp;
  int rc = stem_pool_plan(&sp, pairs, d, pair_tensor_dev, w_dev, b_dev, out_dev);
  if (!rc) rc = stem_pool_launch(sp, as_stream(stream_));
  cudaStreamSynchronize(as_stream(stream_));
  cudaFree(w_dev);
  cudaFree(b_dev);
  return rc;
}

extern "C" int io_net_profile(io_net_t* net, int enable) {
  IO_REQUIRE(net, "io_net_profile: null handle");
  net->profile = enable != 0;
  return IO_OK;
}

extern "C" int io_net_profile_read(io_net_t* net, float* ms, int32_t* kind, double* flop, double* bytes, int32_t* tag,
                                   int max_n) {
  IO_REQUIRE(net && ms && kind && flop, "io_net_profile_read: null pointer");
  const int n = static_cast<int>(net->prof_kind.size());
  for (int i = 0; i < n && i < max_n; ++i) {
    IO_CUDA(cudaEventElapsedTime(&ms[i], net->ev[2 * i], net->ev[2 * i + 1]));
    kind[i] = net->prof_kind[i];
    flop[i] = net->prof_flops[i];
    if (bytes) bytes[i] = net->prof_bytes[i];
    if (tag) tag[i] = net->prof_tag[i];
  }
  return n;
}

extern "C" int io_net_last_launches(const io_net_t* net) { return net ? net->last_launches : 0; }
