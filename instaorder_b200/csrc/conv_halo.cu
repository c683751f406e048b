// 3x3 stride-1 convolution with 64 input / 64 output channels on 64-pixel-wide feature maps (layer1's conv2 at
// 256 x 256 inputs: three launches per forward) with the input rows RESIDENT in shared memory.
//
// conv_tn_kernel<64> loads one TMA box per filter tap: every activation byte crosses L2 -> shared memory nine times
// (288 KB per 256-pixel tile), and ncu showed those launches pinned at the L2 throughput cap (11 TB/s, tensor pipe
// 58 % busy).  Here a CTA walks down a strip of an image and keeps a ring of input rows in shared memory -- each row
// is loaded ONCE (plus a 2-row halo per 16-row strip) as a 72-pixel x 128-byte box (x = -1 .. 70: TMA's zero fill is
// the left / right padding), 128B-swizzled.  All nine taps then read the SAME bytes: an output tile of two image rows
// is N = 136 consecutive positions of the flattened, 72-pixel-pitch row ring (64 + 8 skipped + 64), and tap (ky, kx)
// is the descriptor start address moved by (ky * 72 + kx) * 128 bytes.  A 128B-swizzled K-major operand may start at
// any 128-byte row: tcgen05.mma swizzles on absolute shared-memory address bits (checked on B200 with
// tools/probe/umma_probe.cu, base_offset = 0).  The 3 x 3 x 64 x 64 filter (72 KB) is resident too.
//
// Orientation as conv_tn.cu: channels = MMA M (64: half the tensor rate either way -- an M = 128, N = 64 instruction
// takes the same 64.8 cycles as N = 128), pixels = MMA N, accumulator channel-major in TMEM, transposed to NHWC through
// 128B-swizzled staging and TMA stores.  L2 -> shared-memory traffic: 18 KB per 128-pixel tile instead of 144 KB.
#include "conv_tc.cuh"

namespace io {

namespace {
constexpr int CH_ROW_BYTES = 72 * 128;          // one ring row: 72 pixels x 64 channels bf16 = 9 x 1024 B
constexpr int CH_RING = 8;                      // ring rows (4 groups of 2) + 1 mirror of slot 0 behind the last one
constexpr int CH_X_BYTES = (CH_RING + 1) * CH_ROW_BYTES;
constexpr int CH_W_BYTES = 9 * 8192;            // 9 taps x [64 cout][64 cin] bf16
constexpr int CH_REGION = 4096;                 // 32 pixels x 64 channels
constexpr int CH_EPI_BYTES = 8 * CH_REGION;     // 2 image rows x 2 chunks x 2 slots
constexpr int CH_N = 136;                       // MMA N: 64 pixels, 8 skipped positions, 64 pixels
constexpr int CH_ACC_COLS = 256;                // TMEM columns per accumulator (2 accumulators)
constexpr int CH_SMEM = CH_W_BYTES + CH_X_BYTES + CH_EPI_BYTES + 512 + 256 + 1024;

__device__ __forceinline__ void ch_named_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

struct ChItem {
  int img, t_begin, t_end;
};
__device__ __forceinline__ ChItem ch_item(const HaloParams& p, int w) {
  ChItem it;
  it.img = w / p.strips;
  const int s = w - it.img * p.strips;
  it.t_begin = s * p.strip_len;
  it.t_end = it.t_begin + p.strip_len;
  return it;
}
}  // namespace

__global__ void __launch_bounds__(320, 1) conv_halo_kernel(const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;
  uint8_t* sX = smem + CH_W_BYTES;
  uint8_t* sEpi = sX + CH_X_BYTES;
  float* sBias = reinterpret_cast<float*>(sEpi + CH_EPI_BYTES);
  uint64_t* wfull = reinterpret_cast<uint64_t*>(sEpi + CH_EPI_BYTES + 512);
  uint64_t* xfull = wfull + 1;      // [4] row groups
  uint64_t* xempty = xfull + 4;     // [4]
  uint64_t* tfull = xempty + 4;     // [2]
  uint64_t* tempty = tfull + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.map_w);
    prefetch_tmap(&p.map_x);
    prefetch_tmap(&p.map_out);
    mbar_init(wfull, 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&xfull[i], 1);
      mbar_init(&xempty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 64; i += blockDim.x) sBias[i] = p.bias[i];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // Row groups: group k of an item = input rows (2 (t_begin + k) - 1, 2 (t_begin + k)); tile t reads groups t - t_begin and
  // t - t_begin + 1.  Groups are numbered by a running count c (same in producer and issuer); group c lives in ring rows
  // 2 (c & 3), 2 (c & 3) + 1; ring row 8 mirrors ring row 0 so that a two-row window starting in ring row 7 is contiguous.
  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      mbar_expect_tx(wfull, CH_W_BYTES);
      for (int tap = 0; tap < 9; ++tap) tma_load_2d(sW + tap * 8192, &p.map_w, wfull, tap * 64, 0);
      uint32_t cnt = 0;
      for (int w = blockIdx.x; w < p.items; w += gridDim.x) {
        const ChItem it = ch_item(p, w);
        for (int g = it.t_begin; g <= it.t_end; ++g, ++cnt) {
          const int slot = cnt & 3;
          mbar_wait(&xempty[slot], ((cnt >> 2) & 1) ^ 1);
          mbar_expect_tx(&xfull[slot], (slot == 0 ? 3 : 2) * CH_ROW_BYTES);
          tma_load_4d(sX + (2 * slot) * CH_ROW_BYTES, &p.map_x, &xfull[slot], 0, -1, 2 * g - 1, it.img);
          tma_load_4d(sX + (2 * slot + 1) * CH_ROW_BYTES, &p.map_x, &xfull[slot], 0, -1, 2 * g, it.img);
          if (slot == 0) tma_load_4d(sX + CH_RING * CH_ROW_BYTES, &p.map_x, &xfull[slot], 0, -1, 2 * g - 1, it.img);
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(64, CH_N);
      const uint32_t x_addr = smem_u32(sX);
      const uint64_t w_desc0 = umma_desc_sw128(smem_u32(sW));
      mbar_wait(wfull, 0);
      tc_fence_after();
      uint32_t cnt = 0;
      int tcount = 0;
      for (int w = blockIdx.x; w < p.items; w += gridDim.x) {
        const ChItem it = ch_item(p, w);
        const uint32_t base = cnt;
        uint32_t waited = base;
        for (int t = it.t_begin; t < it.t_end; ++t, ++tcount) {
          const int acc = tcount & 1;
          mbar_wait(&tempty[acc], ((tcount >> 1) & 1) ^ 1);
          const uint32_t c0 = base + static_cast<uint32_t>(t - it.t_begin);
          for (; waited < c0 + 2; ++waited) mbar_wait(&xfull[waited & 3], (waited >> 2) & 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * CH_ACC_COLS;
          // input row 2t - 1 + ky sits in ring row (2 (c0 & 3) + ky) & 7 (the window continues into the next ring row;
          // ring row 8 = copy of ring row 0)
          const uint32_t r0 = 2 * (c0 & 3);
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const uint64_t xrow = umma_desc_sw128(x_addr + ((r0 + ky) & 7) * CH_ROW_BYTES);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(d_tmem, w_desc0 + (((ky * 3 + kx) * 8192 + k * 32) >> 4), xrow + ((kx * 128 + k * 32) >> 4), idesc,
                          (ky > 0 || kx > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&xempty[c0 & 3]);
          umma_commit(&tfull[acc]);
        }
        const uint32_t last = base + static_cast<uint32_t>(it.t_end - it.t_begin);   // the strip's bottom halo group
        umma_commit(&xempty[last & 3]);
        cnt = last + 1;
      }
    }
  } else {
    // ======================= epilogue (warps 2..9): transpose D^T -> NHWC =======================
    const int q = warp & 3;             // TMEM lane quadrant: channels 16q .. 16q+15 in its lanes 0..15 (M = 64 layout)
    const int hsel = (warp - 2) >> 2;   // image row of the tile: accumulator columns 72 * hsel .. 72 * hsel + 63
    const bool leader = q == 0 && lane == 0;
    const bool lane_on = lane < 16;
    const int cl = q * 16 + (lane & 15);
    const float my_bias = sBias[cl];
    const int chunk_off = (cl >> 3) << 4, sub_off = (cl & 7) * 2;
    int tcount = 0;
    for (int w = blockIdx.x; w < p.items; w += gridDim.x) {
      const ChItem it = ch_item(p, w);
      for (int t = it.t_begin; t < it.t_end; ++t, ++tcount) {
        const int acc = tcount & 1;
        mbar_wait(&tfull[acc], (tcount >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * CH_ACC_COLS + 72 * hsel + 32 * c, v);
          tmem_ld_wait();
          if (c == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
          }
          if (leader) tma_store_wait_read<2>();      // 4 slots per image row in flight order: (tile parity, chunk)
          ch_named_bar(1 + hsel, 128);
          uint8_t* region = sEpi + ((hsel * 2 + (tcount & 1)) * 2 + c) * CH_REGION;
          if (lane_on) {
#pragma unroll
            for (int px = 0; px < 32; ++px) {
              float f = __uint_as_float(v[px]) + my_bias;
              if (p.relu) f = fmaxf(f, 0.0f);
              *reinterpret_cast<__nv_bfloat16*>(region + px * 128 + (chunk_off ^ ((px & 7) << 4)) + sub_off) =
                  __float2bfloat16_rn(f);
            }
          }
          fence_proxy_async();
          ch_named_bar(1 + hsel, 128);
          if (leader) {
            tma_store_2d(&p.map_out, region, 0, it.img * 4096 + (2 * t + hsel) * 64 + 32 * c);
            tma_store_commit();
          }
        }
      }
    }
    if (leader) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// 3x3 stride 1, 64 -> 64 channels, 64 x 64 feature maps (layer1 conv2 at 256 x 256 inputs)
bool conv_halo_supported(const ConvDesc& d) {
  static const bool on = []() {
    const char* e = getenv("INSTAORDER_HALO");
    return e == nullptr || atoi(e) != 0;
  }();
  return on && d.kernel == 3 && d.stride == 1 && d.cin == 64 && d.cout == 64 && d.h == 64 && d.w == 64;
}

int conv_halo_plan(HaloParams* p, const ConvDesc& d, const void* x, const void* wgt, const float* bias, void* y, int relu) {
  IO_REQUIRE(conv_halo_supported(d), "conv_halo: unsupported geometry");
  *p = HaloParams{};
  p->bias = bias;
  p->relu = relu;
  p->strip_len = 8;                       // tiles (of two rows) per work item
  p->strips = (d.h / 2) / p->strip_len;
  p->items = d.b * p->strips;
  int rc;
  {
    const uint64_t dims[2] = {9 * 64, 64};
    const uint64_t str[1] = {9 * 64 * 2};
    const uint32_t box[2] = {64, 64};
    if ((rc = make_tmap_bf16(&p->map_w, wgt, 2, dims, str, box, true))) return rc;
  }
  {
    const uint64_t C = 64, W = d.w, H = d.h, B = d.b;
    const uint64_t dims[4] = {C, W, H, B};
    const uint64_t str[3] = {C * 2, W * C * 2, H * W * C * 2};
    const uint32_t box[4] = {64, 72, 1, 1};
    if ((rc = make_tmap_bf16(&p->map_x, x, 4, dims, str, box, true))) return rc;
  }
  const uint64_t odims[2] = {64, static_cast<uint64_t>(d.b) * d.h * d.w};
  const uint64_t ostr[1] = {64 * 2};
  const uint32_t obox[2] = {64, 32};
  return make_tmap_bf16(&p->map_out, y, 2, odims, ostr, obox, true);
}

int conv_halo_launch(const HaloParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    IO_CUDA(cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM));
    attr_set = true;
  }
  if (p.items <= 0) return IO_OK;
  const int grid = p.items < num_sms() ? p.items : num_sms();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(320);
  cfg.dynamicSmemBytes = CH_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  IO_CUDA(cudaLaunchKernelEx(&cfg, conv_halo_kernel, p));
  return IO_OK;
}

}  // namespace io
