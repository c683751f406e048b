// conv1 (7x7 stride 2, both directions of every pair) + folded BN + ReLU + MaxPool2d(3, 2, 1) as ONE kernel
// (reference models/backbone/resnet_cls.py:140-146, 205-208), for 256 x 256 inputs.
//
// Two things bound the separate stem / max-pool launches of round 1 (ncu: tensor pipe 50 % busy, 4.6 KB/cycle of
// L2 -> shared-memory traffic; then a 2.1 GB HBM round trip of the 128 x 128 x 64 stem output through the pool):
//
// 1. The implicit-im2col expansion went through L2: every filter row of every tile was a TMA box of overlapping
//    128-byte windows (8 taps x 8 channels), 224 KB per 256-pixel tile.  Here the padded input rows themselves are
//    staged in shared memory ONCE (4 new rows = 17 KB per tile, a 16-row ring), split by column parity with a 5-D
//    TMA view {8 ch, parity, x/2, y, pair}: E[j] = pixel 2j, O[j] = pixel 2j+1, 16 bytes each.  Output pixel ox of
//    a stride-2 convolution needs taps s = 0..6 at input columns 2 ox + s, i.e. E[ox + 0..3] and O[ox + 0..2]: for a
//    K-major operand WITHOUT swizzle the "row" (pixel) stride inside an 8 x 16 B core matrix is 16 bytes and the
//    stride between K chunks (LBO) is free -- with LBO = 16 B the core matrices of neighbouring taps overlap in
//    memory and one descriptor start address per MMA does the whole im2col (verified on B200 with tools/probe/
//    umma_probe.cu: overlapping no-swizzle descriptors give exact results).  The whole filter (7 x 16 KB) is
//    resident in shared memory.
// 2. The pool is done on the accumulators: the GEMM is "transposed" (M = 2 x 64 output channels = TMEM lanes,
//    N = pixels = TMEM columns; two N = 128 MMAs per K step = the stem rows 2t and 2t+1 -- N = 128 runs at the full
//    tensor rate, 64.8 cycles per 128 x 128 x 16), so an epilogue thread owns ONE channel and walks along a stem row:
//    the horizontal 3-max is in registers, the vertical one needs row 2t-1, which the same thread produced for the
//    previous tile and kept in registers (a CTA walks down a strip of 16 pooled rows; the first tile of a strip is
//    recomputed as a warm-up, +6 % MMA work).  Bias + ReLU + bf16 rounding are monotonic, so they are applied after the
//    max: results are bit-identical to conv -> bias -> ReLU -> bf16 -> max-pool.  Only the pooled 64 x 64 x 64 tensor
//    goes to HBM (TMA store through 128B-swizzled staging).
//
// Warp roles as in conv_tn.cu: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = epilogue (TMEM lane quadrant
// q = warp & 3, pooled-column half h = (warp - 2) >> 2).
#include "conv_tc.cuh"

namespace io {

namespace {
constexpr int SP_W_BYTES = 7 * 16384;          // resident weights: 7 filter rows x [128 ch][64 K] bf16, 128B-swizzled
constexpr int SP_XP = 2176;                    // bytes per parity array of one input row (132 x 16 B, padded to 17 x 128)
constexpr int SP_RING_ROWS = 16;               // 4 groups of 4 input rows
constexpr int SP_X_BYTES = SP_RING_ROWS * 2 * SP_XP;
constexpr int SP_REGION = 4096;                // 32 pooled pixels x 64 channels bf16
constexpr int SP_EPI_BYTES = 8 * SP_REGION;    // 4 (column half x direction) groups x 2 slots
constexpr int SP_SMEM = SP_W_BYTES + SP_X_BYTES + SP_EPI_BYTES + 512 + 256 + 1024;
constexpr int SP_ROW_TX = 132 * 16;            // bytes one parity box brings

// K-major operand without swizzle: 8 x 16 B core matrices, row stride 16 B, LBO between K chunks, SBO between 8-row groups
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return __uint_as_float(r);
}

__device__ __forceinline__ void sp_named_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

struct SpItem {
  int n, t_begin, t_first_out, t_end;   // pair, first tile (incl. warm-up), first tile that produces output, end
};
__device__ __forceinline__ SpItem sp_item(const StemPoolParams& p, int w) {
  SpItem it;
  it.n = w / p.strips;
  const int s = w - it.n * p.strips;
  it.t_first_out = s * p.strip_len;
  it.t_begin = it.t_first_out - (s > 0 ? 1 : 0);
  it.t_end = it.t_first_out + p.strip_len;
  return it;
}
}  // namespace

__global__ void __launch_bounds__(320, 1) stem_pool_kernel(const __grid_constant__ StemPoolParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;
  uint8_t* sX = smem + SP_W_BYTES;
  uint8_t* sEpi = sX + SP_X_BYTES;
  float* sBias = reinterpret_cast<float*>(sEpi + SP_EPI_BYTES);
  uint64_t* wfull = reinterpret_cast<uint64_t*>(sEpi + SP_EPI_BYTES + 512);
  uint64_t* xfull = wfull + 1;      // [4]
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
  for (int i = threadIdx.x; i < 128; i += blockDim.x) sBias[i] = p.bias[i];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ======================= TMA producer: resident weights, then the ring of input-row groups =======================
    if (lane == 0) {
      mbar_expect_tx(wfull, SP_W_BYTES);
      for (int r = 0; r < 7; ++r) tma_load_2d(sW + r * 16384, &p.map_w, wfull, r * 64, 0);
      uint32_t cnt = 0;
      for (int w = blockIdx.x; w < p.items; w += gridDim.x) {
        const SpItem it = sp_item(p, w);
        // tile t reads padded input rows 4t .. 4t+8 = row groups t, t+1 and the first row of group t+2
        for (int g = it.t_begin; g <= it.t_end + 1; ++g, ++cnt) {
          const int slot = cnt & 3;
          mbar_wait(&xempty[slot], ((cnt >> 2) & 1) ^ 1);
          mbar_expect_tx(&xfull[slot], 8 * SP_ROW_TX);
          for (int rr = 0; rr < 4; ++rr)
            for (int par = 0; par < 2; ++par)
              tma_load_5d(sX + ((slot * 4 + rr) * 2 + par) * SP_XP, &p.map_x, &xfull[slot], 0, par, 0, 4 * g + rr, it.n);
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, 128);
      const uint32_t x_addr = smem_u32(sX);
      const uint64_t w_desc0 = umma_desc_sw128(smem_u32(sW));
      mbar_wait(wfull, 0);
      tc_fence_after();
      uint32_t cnt = 0;        // row groups consumed so far (same numbering as the producer)
      int tcount = 0;
      for (int w = blockIdx.x; w < p.items; w += gridDim.x) {
        const SpItem it = sp_item(p, w);
        const uint32_t base = cnt;                      // count of group it.t_begin
        uint32_t waited = base;                         // groups [base, waited) are known to have landed
        for (int t = it.t_begin; t < it.t_end; ++t, ++tcount) {
          const int acc = tcount & 1;
          mbar_wait(&tempty[acc], ((tcount >> 1) & 1) ^ 1);
          const uint32_t need = base + static_cast<uint32_t>(t - it.t_begin) + 3;
          for (; waited < need; ++waited) mbar_wait(&xfull[waited & 3], (waited >> 2) & 1);
          tc_fence_after();
          const uint32_t c0 = base + static_cast<uint32_t>(t - it.t_begin);     // count of group t
          // one thread issues 56 MMAs of 64 cycles each: keep the issue loop to two 64-bit adds per instruction.  A
          // descriptor's low word holds (address >> 4) in 14 bits, so a byte offset is an integer add of (offset >> 4).
          uint64_t brow[9];                              // even-column array of input rows 4t .. 4t+8
#pragma unroll
          for (int yr = 0; yr < 9; ++yr) {
            const uint32_t slot = (c0 + (yr >> 2)) & 3;
            brow[yr] = umma_desc_nosw(x_addr + ((slot * 4 + (yr & 3)) * 2) * SP_XP, 16, 128);
          }
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint32_t d_tmem = tmem_base + acc * 256 + half * 128;
#pragma unroll
            for (int r = 0; r < 7; ++r) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                // j = 0: taps (0, 2) = E[ox], E[ox+1]; 1: taps (4, 6) = E[ox+2], E[ox+3]; 2: taps (1, 3) = O[ox], O[ox+1];
                // 3: taps (5, -) = O[ox+2], O[ox+3] (zero weights on the second chunk)
                umma_bf16(d_tmem, w_desc0 + ((r * 16384 + j * 32) >> 4),
                          brow[2 * half + r] + (((j >> 1) * SP_XP + (j & 1) * 32) >> 4), idesc, (r > 0 || j > 0) ? 1u : 0u);
              }
            }
          }
          umma_commit(&xempty[c0 & 3]);                 // group t is not needed after this tile
          umma_commit(&tfull[acc]);
        }
        // the strip's last two groups were only read by its last tile
        const uint32_t last = base + static_cast<uint32_t>(it.t_end - it.t_begin);
        umma_commit(&xempty[last & 3]);
        umma_commit(&xempty[(last + 1) & 3]);
        cnt = last + 2;
      }
    }
  } else {
    // ======================= epilogue (warps 2..9): 3x3 / 2 max-pool on the accumulators =======================
    const int q = warp & 3;                  // TMEM lane quadrant: GEMM channels 32q .. 32q+31
    const int h = (warp - 2) >> 2;           // pooled columns 32h .. 32h+31 = stem columns 64h .. 64h+63
    const int dir = q >> 1;                  // GEMM channels 64..127 = direction (B, A)
    const int grp = h * 2 + dir;             // the two warps that fill one 32-pixel x 64-channel staging region
    const bool leader = (q & 1) == 0 && lane == 0;
    const int cl = (q & 1) * 32 + lane;      // channel inside the direction
    const float my_bias = sBias[q * 32 + lane];
    const int chunk_off = (cl >> 3) << 4, sub_off = (cl & 7) * 2;
    const float NEG = __int_as_float(0xff800000);   // -inf: MaxPool2d pads with -inf
    float prev[32];                          // horizontally pooled stem row 2t-1 (raw accumulators)
    int tcount = 0, ocount = 0;
    for (int w = blockIdx.x; w < p.items; w += gridDim.x) {
      const SpItem it = sp_item(p, w);
#pragma unroll
      for (int i = 0; i < 32; ++i) prev[i] = NEG;      // top image border; overwritten by the warm-up tile otherwise
      for (int t = it.t_begin; t < it.t_end; ++t, ++tcount) {
        const int acc = tcount & 1;
        mbar_wait(&tfull[acc], (tcount >> 1) & 1);
        tc_fence_after();
        const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256 + 64 * h;
        float carry_a = NEG, carry_b = NEG;            // stem column 64h - 1 (left image border for h = 0)
        if (h == 1) {
          carry_a = tmem_ld1(tbase - 1);
          carry_b = tmem_ld1(tbase + 128 - 1);
        }
        float outv[32];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint32_t va[32], vb[32];
          tmem_ld32(tbase + 32 * j, va);
          tmem_ld32(tbase + 128 + 32 * j, vb);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float la = i == 0 ? carry_a : __uint_as_float(va[2 * i - 1]);
            const float lb = i == 0 ? carry_b : __uint_as_float(vb[2 * i - 1]);
            const float a = fmaxf(fmaxf(la, __uint_as_float(va[2 * i])), __uint_as_float(va[2 * i + 1]));
            const float b = fmaxf(fmaxf(lb, __uint_as_float(vb[2 * i])), __uint_as_float(vb[2 * i + 1]));
            outv[16 * j + i] = fmaxf(fmaxf(prev[16 * j + i], a), b);
            prev[16 * j + i] = b;
          }
          carry_a = __uint_as_float(va[31]);
          carry_b = __uint_as_float(vb[31]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        if (t < it.t_first_out) continue;              // warm-up tile: only `prev` was wanted
        const int slot = ocount & 1;
        ++ocount;
        if (leader) tma_store_wait_read<1>();          // the store that read this slot two tiles ago is done
        sp_named_bar(1 + grp, 64);
        uint8_t* region = sEpi + (grp * 2 + slot) * SP_REGION;
#pragma unroll
        for (int px = 0; px < 32; ++px) {
          const float f = fmaxf(outv[px] + my_bias, 0.0f);
          *reinterpret_cast<__nv_bfloat16*>(region + px * 128 + (chunk_off ^ ((px & 7) << 4)) + sub_off) =
              __float2bfloat16_rn(f);
        }
        fence_proxy_async();
        sp_named_bar(1 + grp, 64);
        if (leader) {
          int srow = (p.img_mul * it.n) * 4096 + t * 64 + 32 * h;
          if (dir == 1) srow += p.split_row_off;
          tma_store_2d(&p.map_out, region, 0, srow);
          tma_store_commit();
        }
      }
    }
    if (leader) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

bool stem_pool_supported(int d) {
  static const bool on = []() {
    const char* e = getenv("INSTAORDER_STEM_POOL");
    return e == nullptr || atoi(e) != 0;
  }();
  return on && d == 256;
}

// x: the padded pair tensor [pairs, d+6, pitch, 8]; wgt: [128][448] bf16 in the K order of this kernel (stem_pool_pack_k);
// y: [2*pairs, d/4, d/4, 64] pooled output, image img_mul * pair (+ split_row_off / 4096 for the second direction)
int stem_pool_plan(StemPoolParams* p, int pairs, int d, const void* x, const void* wgt, const float* bias, void* y) {
  IO_REQUIRE(d == 256, "stem_pool: input size %d (256 only)", d);
  *p = StemPoolParams{};
  const int64_t pitch = io_pair_tensor_row_pitch(d);     // 264 pixels of 16 bytes
  const int hp = d + 6;
  p->bias = bias;
  p->strip_len = 16;
  p->strips = (d / 4) / p->strip_len;
  p->items = pairs * p->strips;
  p->img_mul = 2;
  p->split_row_off = (d / 4) * (d / 4);
  int rc;
  {
    const uint64_t dims[2] = {448, 128};
    const uint64_t str[1] = {448 * 2};
    const uint32_t box[2] = {64, 128};
    if ((rc = make_tmap_bf16(&p->map_w, wgt, 2, dims, str, box, true))) return rc;
  }
  {
    // {8 channels, column parity, column / 2, row, pair}: one box = one parity array (132 x 16 B) of one input row
    const uint64_t dims[5] = {8, 2, static_cast<uint64_t>(pitch / 2), static_cast<uint64_t>(hp), static_cast<uint64_t>(pairs)};
    const uint64_t str[4] = {16, 32, static_cast<uint64_t>(pitch) * 16, static_cast<uint64_t>(hp) * pitch * 16};
    const uint32_t box[5] = {8, 1, static_cast<uint32_t>(pitch / 2), 1, 1};
    if ((rc = make_tmap_bf16(&p->map_x, x, 5, dims, str, box, false))) return rc;
  }
  const uint64_t odims[2] = {64, static_cast<uint64_t>(2) * pairs * (d / 4) * (d / 4)};
  const uint64_t ostr[1] = {64 * 2};
  const uint32_t obox[2] = {64, 32};
  return make_tmap_bf16(&p->map_out, y, 2, odims, ostr, obox, true);
}

// K index of (filter row r, tap s, channel c) in this kernel's weight matrix: per filter row the four K = 16 MMAs read
// taps (0, 2), (4, 6), (1, 3), (5, -) -- even taps from the even-column array, odd taps from the odd-column array
int stem_pool_pack_k(int r, int s, int c) {
  static const int pos[7] = {0, 4, 1, 5, 2, 6, 3};   // tap -> 8-channel chunk inside the filter row's 64 K entries
  return r * 64 + pos[s] * 8 + c;
}

int stem_pool_launch(const StemPoolParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    IO_CUDA(cudaFuncSetAttribute(stem_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_SMEM));
    attr_set = true;
  }
  if (p.items <= 0) return IO_OK;
  const int grid = p.items < num_sms() ? p.items : num_sms();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(320);
  cfg.dynamicSmemBytes = SP_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  IO_CUDA(cudaLaunchKernelEx(&cfg, stem_pool_kernel, p));
  return IO_OK;
}

}  // namespace io
