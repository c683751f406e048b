// Convolution weight gradient on tcgen05 tensor cores (sm_100a).
//
// Replaces the autograd weight-gradient of every nn.Conv2d of the reference backbone
// (models/backbone/resnet_cls.py:86-94, 140, 187-190; driven by `loss.backward()`, models/supervised_order.py:92).
//
//   dW[co][tap][ci] = sum over output pixels p of dy[p][co] * x[p shifted by tap][ci]
//
// GEMM view: M = output channels (128 per tile), N = input channels of one filter tap (<= 256 per tile), K = output
// pixels.  The contraction index is the pixel, i.e. the *row* of both NHWC tensors, so both operands are MN-major:
// a 128B-swizzled TMA box {64 channels, kp pixels} is one MN slab of the UMMA canonical MN-major layout (one pixel =
// one 128-byte row, 8 pixels = one swizzle atom).  Nothing is transposed or re-laid-out in HBM; the shifted /
// strided / padded input window of a tap is the same TMA view the forward convolution uses (OOB zero fill = the
// convolution's zero padding), only with kp-pixel boxes.
//
// Work item = (K slice, tap group, M tile, N tile), K slice slowest so that CTAs running at the same time read the
// same pixels (the taps and the M / N tiles of a K slice re-read them from L2, not HBM).  A tap group is as many
// filter taps as fit one accumulator tile (N = taps x cin <= 256: 3 taps for cin = 64, 2 for cin = 128, 4 + 3 filter
// rows for the stem), so the dy slab is loaded once per group.  The K range is cut into just enough slices to fill
// the SMs once: every item ends by adding its fp32 tile into the flat gradient buffer with vector reductions
// (red.global.add.v4.f32), and that L2-atomic traffic (items x tile) is the cost to minimise.
// Persistent CTAs, warp 0 = TMA producer, warp 1 = MMA issuer / TMEM owner, warps 2..5 = epilogue; smem ring of
// {2 A slabs, up to 4 B slabs} of 8 KB; two TMEM accumulators.
#include "train.cuh"

namespace io {

namespace {
constexpr int SLAB = 8192;          // one MN slab: up to 64 pixels x 64 channels bf16
constexpr int A_BYTES = 2 * SLAB;
constexpr int WG_MAX_SMEM = 232448;
constexpr int WG_BAR_BYTES = 256;

struct KCoord {
  int row0;          // first dy row of the block
  int n0, h0, w0;    // image / output row / output column of the block's first pixel
};

__device__ __forceinline__ KCoord kblock_coord(const WgradParams& p, int kb) {
  KCoord c;
  c.row0 = kb * p.kp;
  if (p.mode == CONV_GEMM) {
    c.n0 = 0; c.h0 = 0; c.w0 = 0;
  } else if (p.bi > 1) {
    c.n0 = kb * p.bi; c.h0 = 0; c.w0 = 0;
  } else {
    c.n0 = kb / p.bpi;
    const int l = kb - c.n0 * p.bpi;
    const int lr = l / p.bpr;
    c.h0 = lr * p.bh;
    c.w0 = (l - lr * p.bpr) * p.wseg;
  }
  return c;
}
}  // namespace

__global__ void __launch_bounds__(192, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int slabs_per_tap = p.bn / 64;
  const int stage_bytes = A_BYTES + p.max_b_slabs * SLAB;
  const int n_stages = p.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + n_stages * stage_bytes);
  uint64_t* empty = full + 8;
  uint64_t* tfull = empty + 8;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t tmem_cols = p.tmem_cols;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.map_dy);
    prefetch_tmap(&p.map_x);
    for (int i = 0; i < n_stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  // PDL: the prologue above overlaps the previous kernel's tail; no global access before this point
  pdl_sync();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int per_slice = p.ngroups * p.m_tiles * p.n_tiles;
  const int items = p.ksplit * per_slice;

  if (warp == 0) {
    // ======================= TMA producer (whole warp) =======================
    // One lane per slab: lane j < a_slabs loads dy slab j, the next ntaps * slabs_per_tap lanes one x slab each.  All
    // per-slab address arithmetic is done once per work item, and the K-block coordinates advance incrementally, so
    // a K block costs one barrier wait + one TMA issue per lane (a single thread computing six box addresses per
    // block was measured to be the kernel's bottleneck: ~3000 cycles per block).
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int ks = item / per_slice;
      int rem = item - ks * per_slice;
      const int grp = rem / (p.m_tiles * p.n_tiles);
      rem -= grp * p.m_tiles * p.n_tiles;
      const int m_tile = rem / p.n_tiles, n_tile = rem - m_tile * p.n_tiles;
      const int k0 = static_cast<int>(static_cast<int64_t>(ks) * p.kblocks / p.ksplit);
      const int k1 = static_cast<int>(static_cast<int64_t>(ks + 1) * p.kblocks / p.ksplit);
      const int tap0 = p.g_tap0[grp], ntaps = p.g_ntaps[grp];
      const int n_loads = p.a_slabs + ntaps * slabs_per_tap;
      const uint32_t tx = static_cast<uint32_t>(n_loads) * p.kp * 128;
      // this lane's slab
      const bool is_a = lane < p.a_slabs;
      int c0 = 0, row_off = 0, dw_ = 0, dh_ = 0, par = 0, tap = 0;
      uint32_t dst_off = 0;
      if (is_a) {
        dst_off = lane * SLAB;
        if (p.a_split_rows) row_off = lane * p.a_split_rows;
        else c0 = m_tile * 128 + lane * 64;
      } else if (lane < n_loads) {
        const int jj = lane - p.a_slabs;
        const int tj = jj / slabs_per_tap, j = jj - tj * slabs_per_tap;
        tap = tap0 + tj;
        const int r = tap / p.taps_w, sx = tap - r * p.taps_w;
        const int dr = r - p.pad, ds = sx - p.pad;
        c0 = n_tile * p.bn + j * 64;
        dst_off = A_BYTES + jj * SLAB;
        if (p.mode == CONV_S2) {
          const int ph = dr & 1, pw = ds & 1;
          c0 += pw * p.cin;
          dw_ = (ds - pw) / 2;
          dh_ = (dr - ph) / 2;
          par = ph;
        } else {
          dw_ = ds;
          dh_ = dr;
        }
      }
      KCoord c = kblock_coord(p, k0);
      for (int kb = k0; kb < k1; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (lane == 0) mbar_expect_tx(&full[stage], tx);
        __syncwarp();
        if (lane < n_loads) {
          uint8_t* dst = smem + stage * stage_bytes + dst_off;
          if (is_a) {
            tma_load_2d(dst, &p.map_dy, &full[stage], c0, c.row0 + row_off);
          } else if (p.mode == CONV_GEMM) {
            tma_load_2d(dst, &p.map_x, &full[stage], c0, c.row0);
          } else if (p.mode == CONV_S1) {
            tma_load_4d(dst, &p.map_x, &full[stage], c0, c.w0 + dw_, c.h0 + dh_, c.n0);
          } else if (p.mode == CONV_S2) {
            tma_load_5d(dst, &p.map_x, &full[stage], c0, c.w0 + dw_, par, c.h0 + dh_, c.n0);
          } else {  // CONV_STEM: filter row `tap` of the overlapping-window view (8 taps x 8 channels per pixel)
            tma_load_4d(dst, &p.map_x, &full[stage], 0, c.w0, 2 * c.h0 + tap, c.n0);
          }
        }
        // next K block (uniform across the warp)
        c.row0 += p.kp;
        if (p.mode != CONV_GEMM) {
          if (p.bi > 1) {
            c.n0 += p.bi;
          } else {
            c.w0 += p.wseg;
            if (c.w0 >= p.w_out) {
              c.w0 = 0;
              c.h0 += p.bh;
              if (c.h0 >= p.h_out) { c.h0 = 0; ++c.n0; }
            }
          }
        }
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      const int ksteps = p.kp / 16;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        const int ks = item / per_slice;
        const int grp = (item - ks * per_slice) / (p.m_tiles * p.n_tiles);
        const uint32_t idesc = (umma_idesc_bf16(128, p.g_ntaps[grp] * p.bn) | UMMA_A_MN | UMMA_B_MN) ^ p.debug_idesc_xor;
        const int k0 = static_cast<int>(static_cast<int64_t>(ks) * p.kblocks / p.ksplit);
        const int k1 = static_cast<int>(static_cast<int64_t>(ks + 1) * p.kblocks / p.ksplit);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * p.acc_cols;
        for (int kb = k0; kb < k1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
          const uint32_t b_addr = a_addr + A_BYTES;
          for (int k = 0; k < ksteps; ++k)   // 16 pixels = two 8-row swizzle atoms = 2048 bytes per K step
            umma_bf16(d_tmem, umma_desc_mn_sw128(a_addr + k * 2048, p.mn_lbo, p.mn_sbo),
                      umma_desc_mn_sw128(b_addr + k * 2048, p.mn_lbo, p.mn_sbo), idesc, (kb > k0 || k > 0) ? 1u : 0u);
          umma_commit(&empty[stage]);
          if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
      }
    }
  } else {
    // ======================= epilogue (warps 2..5): TMEM -> red.add into dW =======================
    const int q = warp & 3;   // TMEM lane quadrant of this warp
    int it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      const int ks = item / per_slice;
      int rem = item - ks * per_slice;
      const int grp = rem / (p.m_tiles * p.n_tiles);
      rem -= grp * p.m_tiles * p.n_tiles;
      const int m_tile = rem / p.n_tiles, n_tile = rem - m_tile * p.n_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int row = m_tile * 128 + q * 32 + lane;
      // accumulator column j * bn + c belongs to tap tap0 + j, input channel n_tile * bn + c: with one N tile per
      // tap (bn == cin) the group's columns are contiguous in dW[row][tap * cin + channel]
      float* dst = p.dw + static_cast<size_t>(row) * p.ldw + p.g_tap0[grp] * p.cin + n_tile * p.bn;
      const int ncols = p.g_ntaps[grp] * p.bn;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      for (int c = 0; c < ncols; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * p.acc_cols + c, v);
        tmem_ld_wait();
        if (row < p.cout && !p.debug_skip_flush) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c + 4 * j),
                         "f"(__uint_as_float(v[4 * j])), "f"(__uint_as_float(v[4 * j + 1])),
                         "f"(__uint_as_float(v[4 * j + 2])), "f"(__uint_as_float(v[4 * j + 3]))
                         : "memory");
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

int wgrad_launch(const WgradParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    IO_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_MAX_SMEM));
    attr_set = true;
  }
  const int items = p.ksplit * p.ngroups * p.m_tiles * p.n_tiles;
  if (items <= 0 || p.kblocks <= 0) return IO_OK;
  const int grid = items < num_sms() ? items : num_sms();
  IO_CUDA(launch_pdl(wgrad_kernel, dim3(grid), dim3(192), static_cast<size_t>(p.smem_bytes), stream, p));
  return IO_OK;
}

// K-block geometry shared by both plans: picks the pixel box {wseg, bh, bi} with wseg * bh * bi <= 64 pixels,
// a multiple of 16 (one UMMA K step)
static int plan_kblocks(WgradParams* p, int b, int h_out, int w_out) {
  int wseg = w_out, bh = 1, bi = 1;
  if (w_out > 64) {
    wseg = 0;
    for (int c = 64; c >= 16; c -= 16)
      if (w_out % c == 0) { wseg = c; break; }
    IO_REQUIRE(wseg > 0, "wgrad: output width %d has no 16-multiple divisor <= 64", w_out);
  } else {
    const int hw = h_out * w_out;
    if (hw <= 64) {
      bh = h_out;
      bi = 64 / hw;
      while (bi > 1 && (bi * hw) % 16 != 0) --bi;
    } else {
      bh = 0;
      for (int c = 64 / w_out; c >= 1; --c)
        if (h_out % c == 0 && (c * w_out) % 16 == 0) { bh = c; break; }
      IO_REQUIRE(bh > 0, "wgrad: no K-block tiling for a %d x %d output", h_out, w_out);
    }
  }
  p->wseg = wseg; p->bh = bh; p->bi = bi;
  p->kp = wseg * bh * bi;
  IO_REQUIRE(p->kp % 16 == 0 && p->kp <= 64, "wgrad: K block of %d pixels (%d x %d output)", p->kp, h_out, w_out);
  p->bpr = w_out / wseg;
  p->bpi = (h_out / bh) * p->bpr;
  p->kblocks = bi > 1 ? (b + bi - 1) / bi : b * p->bpi;
  return IO_OK;
}

static void plan_split(WgradParams* p) {
  // tap groups: as many taps per accumulator tile as fit N <= 256 (only when one N tile covers all input channels)
  int per = 1;
  if (p->n_tiles == 1 && p->bn <= 128) {
    per = 256 / p->bn;
    if (per > p->taps) per = p->taps;
    const int ng = (p->taps + per - 1) / per;
    per = (p->taps + ng - 1) / ng;          // balanced: 9 taps -> 3+3+3 (cin 64) / 2+2+2+2+1 (cin 128), 7 -> 4+3
  }
  p->ngroups = 0;
  for (int t0 = 0; t0 < p->taps; t0 += per) {
    p->g_tap0[p->ngroups] = t0;
    p->g_ntaps[p->ngroups] = (p->taps - t0 < per) ? p->taps - t0 : per;
    ++p->ngroups;
  }
  p->max_b_slabs = per * (p->bn / 64);
  p->acc_cols = per * p->bn;                 // <= 256
  int cols = 2 * p->acc_cols;
  p->tmem_cols = cols <= 128 ? 128 : (cols <= 256 ? 256 : 512);
  // K split: ONE wave of work items.  Every item ends with a flush of its fp32 tile into dW through L2 atomics
  // (measured ~2 TB/s aggregate), so the partial-sum volume (items x tile) is what to minimise; the K range of a
  // (group, M tile, N tile) is cut just far enough to occupy all SMs once.
  const int base = p->ngroups * p->m_tiles * p->n_tiles;
  int ks = num_sms() / base;
  const int max_ks = p->kblocks / 2 > 0 ? p->kblocks / 2 : 1;
  if (ks > max_ks) ks = max_ks;
  if (ks < 1) ks = 1;
  p->ksplit = ks;
  const int stage_bytes = A_BYTES + p->max_b_slabs * SLAB;
  int st = (WG_MAX_SMEM - 1024 - WG_BAR_BYTES) / stage_bytes;
  p->stages = st > 8 ? 8 : st;
  p->smem_bytes = p->stages * stage_bytes + WG_BAR_BYTES + 1024;
  p->mn_lbo = mn_lbo();
  p->mn_sbo = mn_sbo();
  // timing experiments only (results become garbage): INSTAORDER_WGRAD_DEBUG_MAJOR = 1 -> A read as K-major, 2 -> B
  static const int dbg = []() { const char* e = getenv("INSTAORDER_WGRAD_DEBUG_MAJOR"); return e ? atoi(e) : 0; }();
  p->debug_idesc_xor = ((dbg & 1) ? UMMA_A_MN : 0u) | ((dbg & 2) ? UMMA_B_MN : 0u);
  p->debug_skip_flush = (dbg & 4) ? 1 : 0;
}

int wgrad_plan(WgradParams* p, const ConvDesc& d, const void* x, const void* dy, float* dw) {
  IO_REQUIRE(d.kernel == 1 || d.kernel == 3, "wgrad: kernel %d", d.kernel);
  IO_REQUIRE(d.stride == 1 || d.stride == 2, "wgrad: stride %d", d.stride);
  IO_REQUIRE(d.cin % 64 == 0 && d.cout % 64 == 0, "wgrad: channels must be multiples of 64");
  IO_REQUIRE(d.stride == 1 || (d.h % 2 == 0 && d.w % 2 == 0), "wgrad: stride 2 needs even h, w");
  *p = WgradParams{};
  const int h_out = d.h / d.stride, w_out = d.w / d.stride;
  p->dw = dw;
  p->w_out = w_out;
  p->h_out = h_out;
  p->cout = d.cout;
  p->cin = d.cin;
  p->taps = d.kernel * d.kernel;
  p->taps_w = d.kernel;
  p->pad = d.kernel / 2;
  p->ldw = p->taps * d.cin;
  p->m_tiles = (d.cout + 127) / 128;
  p->a_slabs = d.cout >= 128 ? 2 : 1;
  p->bn = d.cin >= 256 ? 256 : d.cin;
  IO_REQUIRE(d.cin % p->bn == 0, "wgrad: cin %d not a multiple of the N tile %d", d.cin, p->bn);
  p->n_tiles = d.cin / p->bn;
  p->flops = 2.0 * d.b * h_out * w_out * p->taps * static_cast<double>(d.cin) * d.cout;
  const uint64_t rows_out = static_cast<uint64_t>(d.b) * h_out * w_out;
  int rc;
  if (d.kernel == 1 && d.stride == 1) {
    p->mode = CONV_GEMM;
    p->kp = 64; p->wseg = 64; p->bh = 1; p->bi = 1; p->bpr = 1; p->bpi = 1;
    p->kblocks = static_cast<int>((rows_out + 63) / 64);
    const uint64_t dims[2] = {static_cast<uint64_t>(d.cin), rows_out};
    const uint64_t str[1] = {static_cast<uint64_t>(d.cin) * 2};
    const uint32_t box[2] = {64, 64};
    rc = make_tmap_bf16(&p->map_x, x, 2, dims, str, box, true);
  } else {
    if ((rc = plan_kblocks(p, d.b, h_out, w_out))) return rc;
    const uint64_t C = d.cin, W = d.w, H = d.h, B = d.b;
    if (d.stride == 1) {
      p->mode = CONV_S1;
      const uint64_t dims[4] = {C, W, H, B};
      const uint64_t str[3] = {C * 2, W * C * 2, H * W * C * 2};
      const uint32_t box[4] = {64, static_cast<uint32_t>(p->wseg), static_cast<uint32_t>(p->bh),
                               static_cast<uint32_t>(p->bi)};
      rc = make_tmap_bf16(&p->map_x, x, 4, dims, str, box, true);
    } else {
      p->mode = CONV_S2;
      const uint64_t dims[5] = {2 * C, W / 2, 2, H / 2, B};
      const uint64_t str[4] = {2 * C * 2, W * C * 2, 2 * W * C * 2, H * W * C * 2};
      const uint32_t box[5] = {64, static_cast<uint32_t>(p->wseg), 1, static_cast<uint32_t>(p->bh),
                               static_cast<uint32_t>(p->bi)};
      rc = make_tmap_bf16(&p->map_x, x, 5, dims, str, box, true);
    }
  }
  if (rc) return rc;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(d.cout), rows_out};
    const uint64_t str[1] = {static_cast<uint64_t>(d.cout) * 2};
    const uint32_t box[2] = {64, static_cast<uint32_t>(p->kp)};
    if ((rc = make_tmap_bf16(&p->map_dy, dy, 2, dims, str, box, true))) return rc;
  }
  plan_split(p);
  return IO_OK;
}

int stem_wgrad_plan(WgradParams* p, int pairs, int d, const void* x, const void* dy, float* dw_scratch) {
  IO_REQUIRE(d % 2 == 0 && d >= 32, "stem wgrad: input size %d", d);
  *p = WgradParams{};
  const int h_out = d / 2, w_out = d / 2;
  const int64_t pitch = io_pair_tensor_row_pitch(d);
  const int hp = d + 6;
  p->dw = dw_scratch;
  p->w_out = w_out;
  p->h_out = h_out;
  p->mode = CONV_STEM;
  p->cout = 128;
  p->cin = 64;
  p->taps = 7;
  p->taps_w = 1;
  p->pad = 0;
  p->ldw = 448;
  p->m_tiles = 1;
  p->n_tiles = 1;
  p->bn = 64;
  p->a_slabs = 2;
  p->a_split_rows = pairs * h_out * w_out;
  p->flops = 2.0 * 2 * pairs * h_out * w_out * 49.0 * 5.0 * 64.0;
  int wseg = 0;
  for (int c = 64; c >= 16; c -= 16)
    if (w_out % c == 0) { wseg = c; break; }
  IO_REQUIRE(wseg > 0, "stem wgrad: output width %d", w_out);
  p->wseg = wseg; p->bh = 1; p->bi = 1; p->kp = wseg;
  p->bpr = w_out / wseg;
  p->bpi = h_out * p->bpr;
  p->kblocks = pairs * p->bpi;
  int rc;
  {
    const uint64_t dims[4] = {64, static_cast<uint64_t>(w_out), static_cast<uint64_t>(hp), static_cast<uint64_t>(pairs)};
    const uint64_t str[3] = {32, static_cast<uint64_t>(pitch) * 16, static_cast<uint64_t>(hp) * pitch * 16};
    const uint32_t box[4] = {64, static_cast<uint32_t>(wseg), 1, 1};
    if ((rc = make_tmap_bf16(&p->map_x, x, 4, dims, str, box, true))) return rc;
  }
  {
    const uint64_t dims[2] = {64, static_cast<uint64_t>(2) * pairs * h_out * w_out};
    const uint64_t str[1] = {64 * 2};
    const uint32_t box[2] = {64, static_cast<uint32_t>(wseg)};
    if ((rc = make_tmap_bf16(&p->map_dy, dy, 2, dims, str, box, true))) return rc;
  }
  plan_split(p);
  return IO_OK;
}

}  // namespace io

// ---- exported entry points for the per-layer parity tests ----------------------------------------------------
extern "C" int io_conv_wgrad(const void* x_dev, int b, int h, int w, int cin, const void* dy_dev, int cout, int kernel,
                             int stride, float* dw_dev, void* stream) {
  IO_REQUIRE(x_dev && dy_dev && dw_dev, "io_conv_wgrad: null pointer");
  io::WgradParams p;
  if (int rc = io::wgrad_plan(&p, io::ConvDesc{b, h, w, cin, cout, kernel, stride}, x_dev, dy_dev, dw_dev)) return rc;
  return io::wgrad_launch(p, io::as_stream(stream));
}

extern "C" int io_conv_dgrad(const void* dy_dev, int b, int h, int w, int cin, int cout, int kernel, const void* w_dev,
                             const float* zero_bias_dev, const void* residual_dev, void* dx_dev, void* stream) {
  IO_REQUIRE(dy_dev && w_dev && zero_bias_dev && dx_dev, "io_conv_dgrad: null pointer");
  io::ConvParams p;
  int bn = 0;
  if (int rc = io::dgrad_plan(&p, &bn, b, h, w, cin, cout, kernel, dy_dev, w_dev, zero_bias_dev, residual_dev, dx_dev))
    return rc;
  return io::conv_tc_launch(p, bn, io::as_stream(stream));
}

extern "C" int io_stem_wgrad(const void* pair_tensor_dev, int pairs, int d, const void* dy_dev, float* dw_scratch_dev,
                             void* stream) {
  IO_REQUIRE(pair_tensor_dev && dy_dev && dw_scratch_dev, "io_stem_wgrad: null pointer");
  io::WgradParams p;
  if (int rc = io::stem_wgrad_plan(&p, pairs, d, pair_tensor_dev, dy_dev, dw_scratch_dev)) return rc;
  return io::wgrad_launch(p, io::as_stream(stream));
}
