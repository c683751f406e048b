// "Transposed" implicit-GEMM convolution for layers with 128 output channels (layer2's 3x3 convolutions and the
// two-direction 7x7 stem): D^T[channel][pixel] = W[channel][K] * X^T[K][pixel].
//
// Why: with both operands in shared memory a tcgen05.mma of M = 128 rows costs >= ~128 cycles for reading its
// 128 x 16 "A" slice, whatever N is -- measured on B200: N = 256 tiles run at 97 % of the bf16 peak, N = 128 at
// ~50 %, N = 64 at ~25 % (profiles/r01_layer_report_events.txt).  A layer with only 128 output channels therefore
// wastes half the tensor pipe if the channels are the MMA's N.  Here the 128 channels are the M side (the weights
// are the "A" operand, K-major as stored) and 256 PIXELS are the N side (the activation tile is the "B" operand,
// also K-major: NHWC), so every MMA is a full 128 x 256 x 16.  The accumulator comes out channel-major
// (TMEM lane = channel, column = pixel) and the epilogue transposes it through shared memory into NHWC.
//
// Same structure as conv_tc_kernel: persistent CTA per SM, warp 0 = TMA producer, warp 1 = MMA issuer, 8 epilogue
// warps, 3-stage ring of {16 KB weights, 32 KB activations}, two TMEM accumulators of 256 columns.
#include "conv_tc.cuh"

namespace io {

namespace {
constexpr int BK = 64;
constexpr int PX = 256;                       // pixels per tile (MMA N) at most; 192 for the 96- / 48-wide rows of 384^2 inputs
constexpr int X_STAGE_BYTES = PX * BK * 2;    // 32 KB
constexpr int STAGES = 3;
constexpr int REGION_BYTES = 32 * 128;        // 32 pixels x 64 channels bf16, 128B-swizzled
constexpr int EPI_BYTES = 16 * REGION_BYTES;  // 4 slots per epilogue warp group
constexpr int BAR_BYTES = 256;
constexpr int TMEM_COLS = 512;

// CH = channels on the MMA's M side: 128, or 64 (M = 64: the accumulator then occupies lanes 0-15 of each of the
// four TMEM lane quadrants -- row m lives in lane (m % 16) + 32 * (m / 16), cute/atom/mma_traits_sm100.hpp tmem_frg)
template <int CH>
struct TnCfg {
  static constexpr int W_STAGE_BYTES = CH * BK * 2;   // 16 KB / 8 KB
  static constexpr int STAGE_BYTES = W_STAGE_BYTES + X_STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 512 + BAR_BYTES + 1024;
};

struct TileTn {
  int n_img, h0, base_row;
};

__device__ __forceinline__ TileTn tile_tn(const TnParams& p, int t) {
  TileTn r;
  if (p.mode == CONV_GEMM) {
    r.n_img = 0; r.h0 = 0; r.base_row = t * p.px;
  } else {
    r.n_img = t / p.tpi;
    r.h0 = (t - r.n_img * p.tpi) * p.bh;
    const int img_out = p.mode == CONV_STEM ? p.img_mul * r.n_img : r.n_img;
    r.base_row = img_out * p.hw_out + r.h0 * p.w_out;
  }
  return r;
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
}  // namespace

template <int CH>
__global__ void __launch_bounds__(320, 1) conv_tn_kernel(const __grid_constant__ TnParams p) {
  constexpr int W_STAGE_BYTES = TnCfg<CH>::W_STAGE_BYTES;
  constexpr int STAGE_BYTES = TnCfg<CH>::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;
  uint8_t* sX = smem + STAGES * W_STAGE_BYTES;
  uint8_t* sEpi = smem + STAGES * STAGE_BYTES;
  float* sBias = reinterpret_cast<float*>(sEpi + EPI_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(sEpi + EPI_BYTES + 512);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.map_w);
    prefetch_tmap(&p.map_x);
    prefetch_tmap(&p.map_out);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < CH; i += blockDim.x) sBias[i] = p.bias[i];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        const TileTn t = tile_tn(p, tile);
        int tap = 0, kb = 0;
        for (int ki = 0; ki < p.k_iters; ++ki) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], W_STAGE_BYTES + p.px * BK * 2);
          uint8_t* dX = sX + stage * X_STAGE_BYTES;
          tma_load_2d(sW + stage * W_STAGE_BYTES, &p.map_w, &full[stage], ki * BK, 0);
          if (p.mode == CONV_GEMM) {
            tma_load_2d(dX, &p.map_x, &full[stage], kb * BK, t.base_row);
          } else if (p.mode == CONV_S1) {
            const int r = tap / 3, s = tap - r * 3;
            tma_load_4d(dX, &p.map_x, &full[stage], kb * BK, s - 1, t.h0 + r - 1, t.n_img);
          } else if (p.mode == CONV_S2) {
            const int r = tap / p.taps_w, s = tap - r * p.taps_w;
            const int dr = r - p.pad, ds = s - p.pad;
            const int ph = dr & 1, pw = ds & 1;
            tma_load_5d(dX, &p.map_x, &full[stage], pw * p.cin + kb * BK, (ds - pw) / 2, ph, t.h0 + (dr - ph) / 2,
                        t.n_img);
          } else {  // CONV_STEM: filter row `tap`; two output rows = two boxes of 128 pixels
            tma_load_4d(dX, &p.map_x, &full[stage], 0, 0, 2 * t.h0 + tap, t.n_img);
            tma_load_4d(dX + X_STAGE_BYTES / 2, &p.map_x, &full[stage], 0, 0, 2 * (t.h0 + 1) + tap, t.n_img);
          }
          if (++kb == p.kpt) { kb = 0; ++tap; }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(CH, p.px);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * PX;
        for (int ki = 0; ki < p.k_iters; ++ki) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t w_addr = smem_u32(sW + stage * W_STAGE_BYTES);
          const uint32_t x_addr = smem_u32(sX + stage * X_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(d_tmem, umma_desc_sw128(w_addr + k * 32), umma_desc_sw128(x_addr + k * 32), idesc,
                      (ki > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
      }
    }
  } else {
    // ======================= epilogue (warps 2..9): transpose D^T -> NHWC =======================
    const int e = warp - 2;
    const int q = warp & 3;             // TMEM lane quadrant
    const int hsel = e >> 2;            // pixel half of the tile: pixels 128*hsel .. 128*hsel+127
    // CH = 128: quadrant q holds channels 32q..32q+31; the two warps (q even / odd) of a pixel half fill one
    //           64-channel staging region.  CH = 64: quadrant q holds channels 16q..16q+15 in its lanes 0..15; the
    //           four warps of a pixel half fill one region.
    const int chalf = (CH == 128) ? (q >> 1) : 0;
    const int pair = (CH == 128) ? hsel * 2 + chalf : hsel;
    constexpr int GROUP_THREADS = (CH == 128) ? 64 : 128;
    const bool leader = ((CH == 128) ? (q & 1) == 0 : q == 0) && lane == 0;
    const bool lane_on = (CH == 128) || lane < 16;
    const int cl = (CH == 128) ? (q & 1) * 32 + lane : q * 16 + (lane & 15);  // channel inside the 64-channel group
    const float my_bias = sBias[(CH == 128) ? q * 32 + lane : q * 16 + (lane & 15)];
    uint8_t* my_slots = sEpi + pair * 4 * REGION_BYTES;
    const int chunk_off = ((cl >> 3) << 4), sub_off = (cl & 7) * 2;
    const int half_px = p.px >> 1;        // pixels per epilogue half: 128, or 96
    const int chunks = p.px >> 6;         // 32-pixel chunks per half: 4, or 3
    int it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      const TileTn t = tile_tn(p, tile);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < chunks; ++c) {
        const int px0 = hsel * half_px + c * 32;
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * PX + px0, v);
        tmem_ld_wait();
        if (c == chunks - 1) {  // accumulator fully read by this warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
        }
        if (leader) {                               // the store that last read this slot (previous tile) is done
          if (chunks == 4) tma_store_wait_read<3>(); else tma_store_wait_read<2>();
        }
        named_bar_sync(1 + pair, GROUP_THREADS);
        uint8_t* region = my_slots + c * REGION_BYTES;
        if (lane_on) {
#pragma unroll
          for (int px = 0; px < 32; ++px) {
            float f = __uint_as_float(v[px]) + my_bias;
            if (p.relu) f = fmaxf(f, 0.0f);
            const __nv_bfloat16 h = __float2bfloat16_rn(f);
            // row = pixel, 16-byte chunks XOR-swizzled by (row & 7); px is a compile-time constant after unrolling
            *reinterpret_cast<__nv_bfloat16*>(region + px * 128 + (chunk_off ^ ((px & 7) << 4)) + sub_off) = h;
          }
        }
        fence_proxy_async();
        named_bar_sync(1 + pair, GROUP_THREADS);
        if (leader) {
          int scol = chalf * 64, srow = t.base_row + px0;
          if (scol >= p.n_split) { scol -= p.n_split; srow += p.split_row_off; }
          tma_store_2d(&p.map_out, region, scol, srow);
          tma_store_commit();
        }
      }
    }
    if (leader) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int CH>
static int conv_tn_launch_t(const TnParams& p, cudaStream_t stream);

int conv_tn_launch(const TnParams& p, cudaStream_t stream) {
  return p.ch == 64 ? conv_tn_launch_t<64>(p, stream) : conv_tn_launch_t<128>(p, stream);
}

template <int CH>
static int conv_tn_launch_t(const TnParams& p, cudaStream_t stream) {
  constexpr int SMEM_BYTES = TnCfg<CH>::SMEM_BYTES;
  static bool attr_set = false;
  if (!attr_set) {
    IO_CUDA(cudaFuncSetAttribute(conv_tn_kernel<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  if (p.tiles <= 0) return IO_OK;
  const int grid = p.tiles < num_sms() ? p.tiles : num_sms();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(320);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  IO_CUDA(cudaLaunchKernelEx(&cfg, conv_tn_kernel<CH>, p));
  return IO_OK;
}

bool tn_enabled() {
  static const bool on = []() {
    const char* e = getenv("INSTAORDER_TN");
    return e == nullptr || atoi(e) != 0;
  }();
  return on;
}

// pixels per tile (MMA N): whole output rows, 256 if the geometry allows it, else 192 (96- and 48-wide rows)
static int tn_tile_pixels(const ConvDesc& d) {
  const int h_out = d.h / d.stride, w_out = d.w / d.stride;
  for (int px : {256, 192})
    if (w_out <= px && px % w_out == 0 && h_out % (px / w_out) == 0) return px;
  return 0;
}

// true when the transposed kernel can run this convolution: 3x3 (stride 1 or 2) with 128 or 64 output channels
// whose output is tiled by whole 256-pixel tiles of full rows
bool conv_tn_supported(const ConvDesc& d) {
  if (d.kernel != 3 || (d.cout != 128 && d.cout != 64) || d.cin % 64 != 0) return false;
  return tn_tile_pixels(d) != 0;
}

int conv_tn_plan(TnParams* p, const ConvDesc& d, const void* x, const void* wgt, const float* bias, void* y, int relu) {
  IO_REQUIRE(conv_tn_supported(d), "conv_tn: unsupported geometry");
  *p = TnParams{};
  const int h_out = d.h / d.stride, w_out = d.w / d.stride;
  const int ktot = 9 * d.cin;
  p->bias = bias;
  p->mode = d.stride == 1 ? CONV_S1 : CONV_S2;
  p->k_iters = ktot / 64;
  p->kpt = d.cin / 64;
  p->taps_w = 3;
  p->pad = 1;
  p->cin = d.cin;
  p->px = tn_tile_pixels(d);
  p->bh = p->px / w_out;
  p->tpi = h_out / p->bh;
  p->tiles = d.b * p->tpi;
  p->w_out = w_out;
  p->hw_out = h_out * w_out;
  const int CH = d.cout;
  p->ch = CH;
  p->ldc = CH;
  p->n_split = CH;
  p->split_row_off = 0;
  p->relu = relu;
  int rc;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(ktot), static_cast<uint64_t>(CH)};
    const uint64_t str[1] = {static_cast<uint64_t>(ktot) * 2};
    const uint32_t box[2] = {64, static_cast<uint32_t>(CH)};
    if ((rc = make_tmap_bf16(&p->map_w, wgt, 2, dims, str, box, true))) return rc;
  }
  const uint64_t C = d.cin, W = d.w, H = d.h, B = d.b;
  if (d.stride == 1) {
    const uint64_t dims[4] = {C, W, H, B};
    const uint64_t str[3] = {C * 2, W * C * 2, H * W * C * 2};
    const uint32_t box[4] = {64, static_cast<uint32_t>(w_out), static_cast<uint32_t>(p->bh), 1};
    rc = make_tmap_bf16(&p->map_x, x, 4, dims, str, box, true);
  } else {
    const uint64_t dims[5] = {2 * C, W / 2, 2, H / 2, B};
    const uint64_t str[4] = {2 * C * 2, W * C * 2, 2 * W * C * 2, H * W * C * 2};
    const uint32_t box[5] = {64, static_cast<uint32_t>(w_out), 1, static_cast<uint32_t>(p->bh), 1};
    rc = make_tmap_bf16(&p->map_x, x, 5, dims, str, box, true);
  }
  if (rc) return rc;
  const uint64_t odims[2] = {static_cast<uint64_t>(CH), static_cast<uint64_t>(d.b) * p->hw_out};
  const uint64_t ostr[1] = {static_cast<uint64_t>(CH) * 2};
  const uint32_t obox[2] = {64, 32};
  return make_tmap_bf16(&p->map_out, y, 2, odims, ostr, obox, true);
}

// Stem (two directions = 128 GEMM channels) for d = 256: tiles of two output rows (2 x 128 pixels)
bool stem_tn_supported(int d) { return d == 256; }

int stem_tn_plan(TnParams* p, int pairs, int d, const void* x, const void* wgt, const float* bias, void* y) {
  IO_REQUIRE(stem_tn_supported(d), "stem_tn: input size %d", d);
  *p = TnParams{};
  const int h_out = d / 2, w_out = d / 2;
  const int64_t pitch = io_pair_tensor_row_pitch(d);
  const int hp = d + 6;
  constexpr int CH = 128;
  p->ch = CH;
  p->bias = bias;
  p->mode = CONV_STEM;
  p->k_iters = 7;
  p->kpt = 1;
  p->taps_w = 7;
  p->pad = 3;
  p->cin = 8;
  p->px = PX;
  p->bh = 2;
  p->tpi = h_out / 2;
  p->tiles = pairs * p->tpi;
  p->w_out = w_out;
  p->hw_out = h_out * w_out;
  p->ldc = 64;
  p->n_split = 64;
  p->split_row_off = p->hw_out;
  p->img_mul = 2;
  p->relu = 1;
  int rc;
  {
    const uint64_t dims[2] = {448, CH};
    const uint64_t str[1] = {448 * 2};
    const uint32_t box[2] = {64, CH};
    if ((rc = make_tmap_bf16(&p->map_w, wgt, 2, dims, str, box, true))) return rc;
  }
  {
    const uint64_t dims[4] = {64, static_cast<uint64_t>(w_out), static_cast<uint64_t>(hp), static_cast<uint64_t>(pairs)};
    const uint64_t str[3] = {32, static_cast<uint64_t>(pitch) * 16, static_cast<uint64_t>(hp) * pitch * 16};
    const uint32_t box[4] = {64, 128, 1, 1};
    if ((rc = make_tmap_bf16(&p->map_x, x, 4, dims, str, box, true))) return rc;
  }
  const uint64_t odims[2] = {64, static_cast<uint64_t>(2) * pairs * p->hw_out};
  const uint64_t ostr[1] = {64 * 2};
  const uint32_t obox[2] = {64, 32};
  return make_tmap_bf16(&p->map_out, y, 2, odims, ostr, obox, true);
}

}  // namespace io
