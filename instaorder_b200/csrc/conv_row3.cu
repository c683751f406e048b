// 3x3 stride-1 convolution with 64 input / 64 output channels on 64-pixel-wide feature maps (layer1's conv2 at
// 256 x 256 inputs, three launches per forward) at the FULL tensor rate.
//
// With 64 output channels every formulation whose accumulator is [pixels x 64] or [64 x pixels] runs the tensor core
// at half rate (an N = 64 or M = 64 tcgen05.mma costs the same 64.8 cycles as N = 128 / M = 128: conv_halo.cu,
// profiles/r02_umma_rate_probe.txt).  Here the three horizontal taps are stacked along N instead of along K:
//
//   D[pixel m, (kx, cout)] = sum_{ky, cin} T[row(m) + ky - 1, x(m), cin] * W[cout][ky][kx][cin]      M = 128, N = 192, K = 192
//
// i.e. the A operand is always the UNSHIFTED pair of image rows (ky selects the ring rows, no horizontal shift, so no
// padding columns and M = 128 = exactly two image rows), 12 MMAs of N = 192 (96 cycles each) replace 36 of N = 64 /
// M = 64 (64.8 cycles each): 1152 instead of 2333 tensor cycles per 128 pixels.  The horizontal shift moves to the
// epilogue: out[x] = D[x - 1, kx = 0] + D[x, kx = 1] + D[x + 1, kx = 2], neighbours fetched with warp shuffles (TMEM
// lane = pixel), the lane 31 | lane 0 seam between the two warps of an image row through 2 KB of shared memory, zero
// at the image border.  Input rows are resident in a shared-memory ring as in conv_halo.cu (each row loaded once per
// 16-row strip, +2 halo rows), the 3 x 3 x 64 x 64 filter (72 KB) too.  The epilogue -- not the MMAs -- bounds the
// kernel, so two teams of eight warps work on alternate tiles (one TMEM accumulator each): 0.207 (conv_halo) -> 0.183
// (one team) -> 0.160 ms per launch of 512 images.
#include "conv_tc.cuh"

namespace io {

namespace {
constexpr int R3_ROW_BYTES = 64 * 128;            // one ring row: 64 pixels x 64 channels bf16
constexpr int R3_RING = 8;                        // ring rows (4 groups of 2) + 1 mirror of ring row 0 behind the last
constexpr int R3_X_BYTES = (R3_RING + 1) * R3_ROW_BYTES;
constexpr int R3_W_BYTES = 9 * 8192;              // slab (ky, kx) = [64 cout][64 cin]; B of ky = 3 slabs = 192 rows
constexpr int R3_REGION = 4096;                   // 32 pixels x 64 channels
constexpr int R3_EPI_BYTES = 2 * 4 * 2 * R3_REGION;   // 2 teams x 4 lane quadrants x 2 slots
constexpr int R3_XCH_BYTES = 2 * 2 * 2 * 2 * 32 * 4;   // [team][row of the tile][channel half][direction][32 floats]
constexpr int R3_N = 192;
constexpr int R3_ACC_COLS = 256;
constexpr int R3_SMEM = R3_W_BYTES + R3_X_BYTES + R3_EPI_BYTES + R3_XCH_BYTES + 512 + 256 + 1024;

__device__ __forceinline__ void r3_named_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

struct R3Item {
  int img, t_begin, t_end;
};
__device__ __forceinline__ R3Item r3_item(const HaloParams& p, int w) {
  R3Item it;
  it.img = w / p.strips;
  const int s = w - it.img * p.strips;
  it.t_begin = s * p.strip_len;
  it.t_end = it.t_begin + p.strip_len;
  return it;
}
}  // namespace

__global__ void __launch_bounds__(576, 1) conv_row3_kernel(const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;
  uint8_t* sX = smem + R3_W_BYTES;
  uint8_t* sEpi = sX + R3_X_BYTES;
  float* sXch = reinterpret_cast<float*>(sEpi + R3_EPI_BYTES);
  float* sBias = reinterpret_cast<float*>(sEpi + R3_EPI_BYTES + R3_XCH_BYTES);
  uint64_t* wfull = reinterpret_cast<uint64_t*>(sEpi + R3_EPI_BYTES + R3_XCH_BYTES + 512);
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

  // Row groups as in conv_halo.cu: group k of an item = input rows (2 (t_begin + k) - 1, 2 (t_begin + k)); tile t reads
  // groups t - t_begin and t - t_begin + 1.  Group c (running count) lives in ring rows 2 (c & 3), 2 (c & 3) + 1; ring
  // row 8 mirrors ring row 0 so that a two-row window starting in ring row 7 is contiguous.
  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      mbar_expect_tx(wfull, R3_W_BYTES);
      for (int tap = 0; tap < 9; ++tap) tma_load_2d(sW + tap * 8192, &p.map_w, wfull, tap * 64, 0);
      uint32_t cnt = 0;
      for (int w = blockIdx.x; w < p.items; w += gridDim.x) {
        const R3Item it = r3_item(p, w);
        for (int g = it.t_begin; g <= it.t_end; ++g, ++cnt) {
          const int slot = cnt & 3;
          mbar_wait(&xempty[slot], ((cnt >> 2) & 1) ^ 1);
          mbar_expect_tx(&xfull[slot], (slot == 0 ? 3 : 2) * R3_ROW_BYTES);
          tma_load_4d(sX + (2 * slot) * R3_ROW_BYTES, &p.map_x, &xfull[slot], 0, 0, 2 * g - 1, it.img);
          tma_load_4d(sX + (2 * slot + 1) * R3_ROW_BYTES, &p.map_x, &xfull[slot], 0, 0, 2 * g, it.img);
          if (slot == 0) tma_load_4d(sX + R3_RING * R3_ROW_BYTES, &p.map_x, &xfull[slot], 0, 0, 2 * g - 1, it.img);
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, R3_N);
      const uint32_t x_addr = smem_u32(sX);
      const uint64_t w_desc0 = umma_desc_sw128(smem_u32(sW));
      mbar_wait(wfull, 0);
      tc_fence_after();
      uint32_t cnt = 0;
      int tcount = 0;
      for (int w = blockIdx.x; w < p.items; w += gridDim.x) {
        const R3Item it = r3_item(p, w);
        const uint32_t base = cnt;
        uint32_t waited = base;
        for (int t = it.t_begin; t < it.t_end; ++t, ++tcount) {
          const int acc = tcount & 1;
          mbar_wait(&tempty[acc], ((tcount >> 1) & 1) ^ 1);
          const uint32_t c0 = base + static_cast<uint32_t>(t - it.t_begin);
          for (; waited < c0 + 2; ++waited) mbar_wait(&xfull[waited & 3], (waited >> 2) & 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * R3_ACC_COLS;
          const uint32_t r0 = 2 * (c0 & 3);     // input row 2t - 1 + ky sits in ring row (r0 + ky) & 7
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const uint64_t a_desc = umma_desc_sw128(x_addr + ((r0 + ky) & 7) * R3_ROW_BYTES);
            const uint64_t b_desc = w_desc0 + ((ky * 3 * 8192) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (ky > 0 || k > 0) ? 1u : 0u);
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
    // ======================= epilogue: two teams of 8 warps (2..9, 10..17), alternate tiles =======================
    // With one team the kernel was bound by the per-tile latency CHAIN of the epilogue (TMEM load -> seam exchange ->
    // shuffles -> staging -> store; ~2700 cycles against 1152 of MMAs, unchanged by splitting a tile over more warps):
    // team k takes the tiles with (tile count & 1) == k, i.e. accumulator k, so two chains overlap.
    const int team = (warp - 2) >> 3;
    const int e = (warp - 2) & 7;
    const int q = warp & 3;             // TMEM lane quadrant: pixels 32q .. 32q+31 of the tile = image row q >> 1, x = 32 (q & 1) + lane
    const int hs = e >> 2;              // output channels 32 hs .. 32 hs + 31 (two passes of 16)
    const int rsel = q >> 1;            // image row of the tile
    const bool right_half = (q & 1) != 0;
    const bool leader = hs == 0 && lane == 0;
    const int bar_id = 1 + team * 2 + rsel;   // the four warps of (team, image row)
    const int src_l = (lane + 31) & 31, src_r = (lane + 1) & 31;
    int tcount = 0, mine = 0;
    for (int w = blockIdx.x; w < p.items; w += gridDim.x) {
      const R3Item it = r3_item(p, w);
      for (int t = it.t_begin; t < it.t_end; ++t, ++tcount) {
        if ((tcount & 1) != team) continue;
        const int acc = team;
        mbar_wait(&tfull[acc], (tcount >> 1) & 1);
        tc_fence_after();
        uint8_t* region = sEpi + ((team * 4 + q) * 2 + (mine & 1)) * R3_REGION;
        uint8_t* rowp = region + lane * 128;
        float* xch = sXch + ((team * 2 + rsel) * 2 + hs) * 64;   // [0, 32): left warp's vl, [32, 64): right warp's vr
        if (leader) tma_store_wait_read<1>();     // the store that last read this staging slot
#pragma unroll 1
        for (int c2 = 0; c2 < 2; ++c2) {          // 16 channels at a time (register budget of 18 warps)
          uint32_t vl[16], vm[16], vr[16];
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * R3_ACC_COLS + 32 * hs + 16 * c2;
          tmem_ld16(taddr, vl);            // kx = 0: contribution of this pixel to the output pixel on its right
          tmem_ld16(taddr + 64, vm);       // kx = 1
          tmem_ld16(taddr + 128, vr);      // kx = 2: ... to the output pixel on its left
          tmem_ld_wait();
          if (c2 == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
          }
          // seam between the two warps of an image row: x = 31 (left warp, lane 31) | x = 32 (right warp, lane 0)
          if (!right_half && lane == 31) {
#pragma unroll
            for (int i = 0; i < 16; ++i) xch[16 * c2 + i] = __uint_as_float(vl[i]);
          }
          if (right_half && lane == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) xch[32 + 16 * c2 + i] = __uint_as_float(vr[i]);
          }
          r3_named_bar(bar_id, 128);
          // The value a warp's edge lane would hand to a pixel outside its image row is never used (left warp: lane 31's
          // vl went to the seam buffer; right warp: lane 31 is x = 63), so that register takes what the opposite edge
          // lane must RECEIVE -- the other warp's seam value, or zero at the image border -- and a rotating shuffle then
          // serves all 32 lanes without per-element selects.
          if (lane == 31) {
#pragma unroll
            for (int i = 0; i < 16; ++i) vl[i] = right_half ? __float_as_uint(xch[16 * c2 + i]) : 0u;
          }
          if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) vr[i] = right_half ? 0u : __float_as_uint(xch[32 + 16 * c2 + i]);
          }
          __syncwarp();
          const float4* bias4 = reinterpret_cast<const float4*>(sBias + 32 * hs + 16 * c2);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            float bb[8];
            *reinterpret_cast<float4*>(&bb[0]) = bias4[2 * k];  *reinterpret_cast<float4*>(&bb[4]) = bias4[2 * k + 1];
            float o[8];
#pragma unroll
            for (int el = 0; el < 8; ++el) {
              const int i = 8 * k + el;
              const float fl = __uint_as_float(__shfl_sync(0xffffffffu, vl[i], src_l));
              const float fr = __uint_as_float(__shfl_sync(0xffffffffu, vr[i], src_r));
              o[el] = (__uint_as_float(vm[i]) + fl) + fr + bb[el];
            }
            uint4 v4;
            if (p.relu) {
              v4.x = pack_bf16_relu(o[0], o[1]); v4.y = pack_bf16_relu(o[2], o[3]);
              v4.z = pack_bf16_relu(o[4], o[5]); v4.w = pack_bf16_relu(o[6], o[7]);
            } else {
              v4.x = pack_bf16(o[0], o[1]); v4.y = pack_bf16(o[2], o[3]);
              v4.z = pack_bf16(o[4], o[5]); v4.w = pack_bf16(o[6], o[7]);
            }
            *reinterpret_cast<uint4*>(rowp + (((hs * 4 + c2 * 2 + k) ^ (lane & 7)) << 4)) = v4;
          }
        }
        fence_proxy_async();
        r3_named_bar(bar_id, 128);
        if (leader) {
          tma_store_2d(&p.map_out, region, 0, it.img * 4096 + t * 128 + q * 32);
          tma_store_commit();
        }
        ++mine;
      }
    }
    if (leader) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// 3x3 stride 1, 64 -> 64 channels, 64 x 64 feature maps (layer1 conv2 at 256 x 256 inputs)
bool conv_row3_supported(const ConvDesc& d) {
  static const bool on = []() {
    const char* e = getenv("INSTAORDER_ROW3");
    return e == nullptr || atoi(e) != 0;
  }();
  return on && d.kernel == 3 && d.stride == 1 && d.cin == 64 && d.cout == 64 && d.h == 64 && d.w == 64;
}

int conv_row3_plan(HaloParams* p, const ConvDesc& d, const void* x, const void* wgt, const float* bias, void* y, int relu) {
  IO_REQUIRE(conv_row3_supported(d), "conv_row3: unsupported geometry");
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
    const uint32_t box[4] = {64, 64, 1, 1};
    if ((rc = make_tmap_bf16(&p->map_x, x, 4, dims, str, box, true))) return rc;
  }
  const uint64_t odims[2] = {64, static_cast<uint64_t>(d.b) * d.h * d.w};
  const uint64_t ostr[1] = {64 * 2};
  const uint32_t obox[2] = {64, 32};
  return make_tmap_bf16(&p->map_out, y, 2, odims, ostr, obox, true);
}

int conv_row3_launch(const HaloParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    IO_CUDA(cudaFuncSetAttribute(conv_row3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, R3_SMEM));
    attr_set = true;
  }
  if (p.items <= 0) return IO_OK;
  const int grid = p.items < num_sms() ? p.items : num_sms();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(576);
  cfg.dynamicSmemBytes = R3_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  IO_CUDA(cudaLaunchKernelEx(&cfg, conv_row3_kernel, p));
  return IO_OK;
}

}  // namespace io
