// Implicit-GEMM convolution for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma (fp32 accumulators in
// TMEM, double buffered) -> tcgen05.ld epilogue with folded-BN bias, residual add and ReLU fused, bf16 NHWC out.
//
// Replaces nn.Conv2d + nn.BatchNorm2d(eval) + ReLU (+ `out += identity`) of the reference's Bottleneck
// (models/backbone/resnet_cls.py:96-116) and the stem conv1/bn1/relu (:204-206).
//
// One persistent CTA per SM, 6 warps: warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread) and TMEM
// owner, warps 2..5 = epilogue (one TMEM lane quadrant each).  GEMM view: M = output pixels (128 per tile),
// N = output channels (BN per tile), K = taps * Cin in 64-wide blocks.  The A operand is never materialised:
// every filter tap is one TMA box of the NHWC activation tensor, shifted by the tap offset; out-of-bounds rows /
// columns are zero-filled by TMA, which *is* the convolution's zero padding.
#include "conv_tc.cuh"

namespace io {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB

template <int BN>
struct Cfg {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (BN == 256) ? 3 : (BN == 128 ? 5 : 6);
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulator buffers: 128 / 256 / 512 columns
  // epilogue staging: per epilogue warp, BN/64 regions of 32 rows x 64 columns bf16 (4 KB, 128B-swizzled).  A
  // region first receives the residual tile (TMA load), is overwritten in place with the result, and is then
  // TMA-stored -- every global access of the epilogue is a full 128-byte line.
  static constexpr int EPI_REGION_BYTES = 32 * 128;
  // 8 epilogue warps (two per TMEM lane quadrant, splitting the 64-column groups between them) so that every SM
  // sub-partition has two warps to hide tcgen05.ld / shared-memory / barrier latencies; N = 64 uses 4 of them.
  // Every warp owns two staging slots: for N = 256 its two column groups of the tile, otherwise one group with
  // the slot alternating per tile -- so a slot is refilled two store-commits after it was last stored from.
  static constexpr int EPI_WARPS = 8;
  static constexpr int GROUPS = BN / 64;
  static constexpr int ACTIVE_EPI_WARPS = (GROUPS == 1) ? 4 : 8;
  static constexpr int EPI_BYTES = EPI_WARPS * 2 * EPI_REGION_BYTES;
  static constexpr int BIAS_BYTES = (BN == 256) ? 8192 : 512;  // whole bias vector of the layer (<= 2048 / 128 floats)
  static constexpr int BAR_BYTES = 512;  // up to 8+8+2+2+16+1 mbarriers + the TMEM base slot
  static constexpr int FIXED_BYTES = EPI_BYTES + BIAS_BYTES + BAR_BYTES + 1024;  // + 1024 B alignment slack
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + FIXED_BYTES;
  static constexpr int MAX_SMEM = 232448;  // 227 KB
  // CTA pair (cta_group::2, N = 256 only): every CTA stages its own 128 A rows and HALF of the weight rows
  static constexpr int PAIR_B_BYTES = (BN / 2) * BK * 2;
  static constexpr int PAIR_STAGE_BYTES = A_STAGE_BYTES + PAIR_B_BYTES;
  static constexpr int PAIR_STAGES = 4;
};

struct TileCoord {
  int n_img, h0, w0;   // first image / output row / output column of the tile
  int base_row;        // flat output row of tile row 0
  int limit;           // rows of the tile that exist (before the m_total clamp)
};

__device__ __forceinline__ TileCoord tile_coord(const ConvParams& p, int m_tile) {
  TileCoord t;
  if (p.mode == CONV_GEMM) {
    t.n_img = 0; t.h0 = 0; t.w0 = 0;
    t.base_row = m_tile * BM;
    t.limit = BM;
  } else {
    const int g = m_tile / p.tpg, l = m_tile - g * p.tpg;
    t.n_img = g * p.bi;
    if (p.mode == CONV_STEM) {
      t.h0 = l / p.tpr;
      t.w0 = (l - t.h0 * p.tpr) * BM;
      t.limit = min(p.rows_per_tile, p.w_out - t.w0);
    } else if (p.tpr > 1) {   // 3x3 stride 1 on rows wider than one tile: `tpr` tiles of rows_per_tile pixels per row
      t.h0 = l / p.tpr;
      t.w0 = (l - t.h0 * p.tpr) * p.rows_per_tile;
      t.limit = min(p.rows_per_tile, p.w_out - t.w0);
    } else {
      t.h0 = l * p.bh;
      t.w0 = 0;
      t.limit = min(p.rows_per_tile, p.bi * p.hw_out - t.h0 * p.w_out);
    }
    // the stem writes image 2n (direction A,B) and image 2n+1 (direction B,A; columns >= n_split)
    const int img_out = p.mode == CONV_STEM ? p.img_mul * t.n_img : t.n_img;
    t.base_row = img_out * p.hw_out + t.h0 * p.w_out + t.w0;
  }
  return t;
}

// PAIR: two CTAs of a cluster (one TPC) work on 256 output rows x 256 output channels with tcgen05.mma.cta_group::2:
// CTA `rank` owns M tile 2 * pair + rank (its own A boxes, epilogue and TMEM lanes) and stages weight rows
// [128 rank, 128 rank + 128) of the N tile; the leader (rank 0) issues the MMAs for both, its "full" barriers count
// the bytes of both CTAs' TMA loads, "empty" / "accumulator full" arrive in both CTAs by multicast commits and the
// peer's epilogue warps arrive on the leader's "accumulator drained" barriers.  Per K block a CTA takes in 32 KB
// instead of 48 KB for the same 512 tensor cycles: the 128 x 256 single-CTA tiles were bound by the SM's operand
// intake (96 B / cycle needed), not by the tensor pipe.
template <int BN, bool PAIR>
__global__ void __launch_bounds__(320, 1) conv_tc_kernel(const __grid_constant__ ConvParams p) {
  using C = Cfg<BN>;
  constexpr int B_SLOT_BYTES = PAIR ? C::PAIR_B_BYTES : C::B_STAGE_BYTES;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int n_workers = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int worker = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int n_stages = p.stages;
  const bool b_res = p.b_resident != 0;
  uint8_t* sA = smem;
  uint8_t* sB = smem + n_stages * A_STAGE_BYTES;   // ring slots, or the resident [k_iters][BN x 64] weights
  uint8_t* sEpi = sB + (b_res ? p.k_iters : n_stages) * B_SLOT_BYTES;
  // one staging slot per epilogue warp (p.epi_one_slot: pair kernel without a residual) frees half of the staging area
  // for one more ring stage
  const int epi_bytes = p.epi_one_slot ? C::EPI_BYTES / 2 : C::EPI_BYTES;
  float* sBias = reinterpret_cast<float*>(sEpi + epi_bytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(sEpi + epi_bytes + C::BIAS_BYTES);
  uint64_t* empty = full + 8;
  uint64_t* tfull = empty + 8;
  uint64_t* tempty = tfull + 2;
  uint64_t* rbar = tempty + 2;  // residual tile landed: one barrier per epilogue warp and staging slot
  uint64_t* bres_full = rbar + 16;  // resident weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.map_a);
    prefetch_tmap(&p.map_b);
    prefetch_tmap(&p.map_out);
    if (p.residual != nullptr) prefetch_tmap(&p.map_res);
    if (p.k1 > 0) prefetch_tmap(&p.map_a2);
    for (int i = 0; i < 16; ++i) mbar_init(&rbar[i], 1);
    for (int i = 0; i < n_stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(bres_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      // one arrival per participating epilogue warp (of both CTAs in pair mode)
      mbar_init(&tempty[i], PAIR ? 2 * C::ACTIVE_EPI_WARPS : C::ACTIVE_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) {
      tmem_alloc2(tmem_slot, C::TMEM_COLS);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_slot, C::TMEM_COLS);
      tmem_relinquish();
    }
  }
  for (int i = threadIdx.x; i < p.n_total; i += blockDim.x) sBias[i] = p.bias[i];  // weights: not produced upstream
  // Programmatic dependent launch: let the next convolution's CTAs start their prologue on SMs this grid has
  // already vacated, and do not touch activations before the previous grid has completed and flushed.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // pair: the peer's barriers must exist before anything arrives
  tc_fence_after();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  // work items: (M tile, N tile), or (pair of M tiles, N tile); this CTA's M tile of item `tile` is m_of(tile)
  const int m_items = PAIR ? (p.m_tiles + 1) >> 1 : p.m_tiles;
  const int total_tiles = m_items * p.n_tiles;
  auto m_of = [&](int tile) { return PAIR ? 2 * (tile / p.n_tiles) + static_cast<int>(rank) : tile / p.n_tiles; };

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      if (!PAIR && b_res && worker < total_tiles) {   // weights of the whole layer, once
        mbar_expect_tx(bres_full, p.k_iters * C::B_STAGE_BYTES);
        for (int ki = 0; ki < p.k_iters; ++ki)
          tma_load_2d(sB + ki * C::B_STAGE_BYTES, &p.map_b, bres_full, ki * BK, 0);
      }
      for (int tile = worker; tile < total_tiles; tile += n_workers) {
        const int m_tile = m_of(tile), n_tile = tile % p.n_tiles;
        const TileCoord t = tile_coord(p, m_tile);
        if (PAIR) {
          // both CTAs signal the leader's "full" barrier; the leader arms it with the bytes of both
          int tap = 0, kb = 0;
          for (int ki = 0; ki < p.k_iters; ++ki) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (rank == 0) mbar_expect_tx(&full[stage], 2 * (p.a_bytes + C::PAIR_B_BYTES));
            const uint32_t fb = mapa_u32(smem_u32(&full[stage]), 0);
            void* dA = sA + stage * A_STAGE_BYTES;
            uint8_t* dB = sB + stage * C::PAIR_B_BYTES;
            if (ki < p.k1) {
              tma2_load_2d(dA, &p.map_a2, fb, ki * BK, t.base_row);
            } else if (p.mode == CONV_GEMM) {
              tma2_load_2d(dA, &p.map_a, fb, kb * BK, t.base_row);
            } else if (p.mode == CONV_S1) {
              const int r = tap / 3, s = tap - r * 3;
              if (p.flip) tma2_load_4d(dA, &p.map_a, fb, kb * BK, t.w0 + 1 - s, t.h0 + 1 - r, t.n_img);
              else tma2_load_4d(dA, &p.map_a, fb, kb * BK, t.w0 + s - 1, t.h0 + r - 1, t.n_img);
            } else {   // CONV_S2
              const int r = tap / p.taps_w, s = tap - r * p.taps_w;
              const int dr = r - p.pad, ds = s - p.pad;
              const int ph = dr & 1, pw = ds & 1;
              tma2_load_5d(dA, &p.map_a, fb, pw * p.cin + kb * BK, (ds - pw) / 2, ph, t.h0 + (dr - ph) / 2, t.n_img);
            }
            if (p.b_mn) {
#pragma unroll
              for (int j = 0; j < BN / 128; ++j)
                tma2_load_2d(dB + j * 8192, &p.map_b, fb,
                             tap * p.b_tap_cols + n_tile * BN + (static_cast<int>(rank) * (BN / 128) + j) * 64, kb * BK);
            } else {
              tma2_load_2d(dB, &p.map_b, fb, ki * BK, n_tile * BN + static_cast<int>(rank) * (BN / 2));
            }
            if (ki >= p.k1 && ++kb == p.kpt) { kb = 0; ++tap; }
            if (++stage == n_stages) { stage = 0; phase ^= 1; }
          }
          continue;
        }
        // The ring holds less than one tile of K blocks for the small-channel layers, so the first touch of a
        // tile's activations would expose a full DRAM latency per tile: pull the NEXT tile's rows into L2 now.
        if (p.prefetch && tile + n_workers < total_tiles) {
          const int nm = (tile + n_workers) / p.n_tiles;
          if (nm != m_tile) {
            const TileCoord nx = tile_coord(p, nm);
            if (p.mode == CONV_GEMM) {
              for (int b = 0; b < p.kpt; ++b) tma_prefetch_2d(&p.map_a, b * BK, nx.base_row);
            } else if (p.mode == CONV_S1) {
              for (int b = 0; b < p.kpt; ++b) {
                tma_prefetch_4d(&p.map_a, b * BK, nx.w0, nx.h0 - 1, nx.n_img);   // taps r = 0 and r = 2 cover every
                tma_prefetch_4d(&p.map_a, b * BK, nx.w0, nx.h0 + 1, nx.n_img);   // input row the tile needs
              }
            } else if (p.mode == CONV_STEM) {
              for (int r = 0; r < 7; ++r) tma_prefetch_4d(&p.map_a, 0, nx.w0, 2 * nx.h0 + r, nx.n_img);
            }
          }
        }
        int tap = 0, kb = 0;
        for (int ki = 0; ki < p.k_iters; ++ki) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], p.a_bytes + (b_res ? 0 : C::B_STAGE_BYTES));
          void* dA = sA + stage * A_STAGE_BYTES;
          void* dB = sB + stage * C::B_STAGE_BYTES;
          if (ki < p.k1) {   // dual-source K: leading blocks from the flat second matrix
            tma_load_2d(dA, &p.map_a2, &full[stage], ki * BK, t.base_row);
          } else if (p.mode == CONV_GEMM) {
            tma_load_2d(dA, &p.map_a, &full[stage], kb * BK, t.base_row);
          } else if (p.mode == CONV_S1) {
            const int r = tap / 3, s = tap - r * 3;
            if (p.flip) tma_load_4d(dA, &p.map_a, &full[stage], kb * BK, t.w0 + 1 - s, t.h0 + 1 - r, t.n_img);
            else tma_load_4d(dA, &p.map_a, &full[stage], kb * BK, t.w0 + s - 1, t.h0 + r - 1, t.n_img);
          } else if (p.mode == CONV_S2) {
            const int r = tap / p.taps_w, s = tap - r * p.taps_w;
            const int dr = r - p.pad, ds = s - p.pad;
            const int ph = dr & 1, pw = ds & 1;
            tma_load_5d(dA, &p.map_a, &full[stage], pw * p.cin + kb * BK, (ds - pw) / 2, ph, t.h0 + (dr - ph) / 2,
                        t.n_img);
          } else {  // CONV_STEM: filter row `tap`, 8 taps x 8 channels per K block
            tma_load_4d(dA, &p.map_a, &full[stage], 0, t.w0, 2 * t.h0 + tap, t.n_img);
          }
          if (p.b_mn) {   // forward weights as MN-major slabs: {64 input channels (N)} x {64 output channels (K)}
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(static_cast<uint8_t*>(dB) + j * 8192, &p.map_b, &full[stage],
                          tap * p.b_tap_cols + n_tile * BN + j * 64, kb * BK);
          } else if (!b_res) {
            tma_load_2d(dB, &p.map_b, &full[stage], ki * BK, n_tile * BN);
          }
          if (ki >= p.k1 && ++kb == p.kpt) { kb = 0; ++tap; }
          if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * BM : BM, BN) | (p.b_mn ? UMMA_B_MN : 0u);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      if (!PAIR && b_res && worker < total_tiles) {
        mbar_wait(bres_full, 0);
        tc_fence_after();
      }
      for (int tile = worker; tile < total_tiles; tile += n_workers, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int ki = 0; ki < p.k_iters; ++ki) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(sB + (b_res ? ki : stage) * B_SLOT_BYTES);
          if (PAIR) {
            if (p.b_mn) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                umma2_bf16(d_tmem, umma_desc_sw128(a_addr + k * 32),
                           umma_desc_mn_sw128(b_addr + k * 2048, p.mn_lbo, p.mn_sbo), idesc, (ki > 0 || k > 0) ? 1u : 0u);
            } else {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                umma2_bf16(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                           (ki > 0 || k > 0) ? 1u : 0u);
            }
            umma2_commit_mc(&empty[stage]);   // frees the slot in both CTAs
          } else {
            if (p.b_mn) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                umma_bf16(d_tmem, umma_desc_sw128(a_addr + k * 32),
                          umma_desc_mn_sw128(b_addr + k * 2048, p.mn_lbo, p.mn_sbo), idesc, (ki > 0 || k > 0) ? 1u : 0u);
            } else {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                umma_bf16(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                          (ki > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty[stage]);  // frees the smem slot when these MMAs have read it
          }
          if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
        if (PAIR) umma2_commit_mc(&tfull[acc]); else umma_commit(&tfull[acc]);  // accumulator complete
      }
    }
  } else {
    // ======================= epilogue (warps 2..9) =======================
    const int e = warp - 2;      // epilogue warp index
    const int q = warp & 3;      // TMEM lane quadrant accessible to this warp: tile rows 32q .. 32q+31
    const int hsel = e >> 2;     // which of the quadrant's two warps: owns column groups hsel, hsel + 2
    const int r = q * 32 + lane;
    const bool has_res = p.residual != nullptr;
    constexpr int MY_GROUPS = (C::GROUPS + 1) / 2;     // 1 (N = 64, 128) or 2 (N = 256)
    if (hsel < C::GROUPS) {
      const bool one_slot = p.epi_one_slot != 0;
      uint8_t* my_stage = sEpi + e * (one_slot ? 1 : 2) * C::EPI_REGION_BYTES;
      uint64_t* my_rbar = &rbar[e * 2];
      uint32_t rphase[2] = {0, 0};
      int it = 0;
      // "accumulator drained" lives in the leader CTA
      const uint32_t tempty_leader0 = PAIR ? mapa_u32(smem_u32(&tempty[0]), 0) : 0u;
      for (int tile = worker; tile < total_tiles; tile += n_workers, ++it) {
        const int m_tile = m_of(tile), n_tile = tile % p.n_tiles;
        const TileCoord t = tile_coord(p, m_tile);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        const int row0 = t.base_row + q * 32;                    // first output row of this warp's slab
        // whole slab inside the tile -> TMA path (rows past the end of the tensor are clipped by TMA);
        // slab cut by the tile's row limit (384^2 geometries) -> per-thread path with a row guard
        const bool slab_full = (q * 32 + 32 <= t.limit) && (row0 < p.m_total);
        const bool slab_part = !slab_full && (q * 32 < t.limit) && (row0 < p.m_total);
        const bool valid = slab_part && (r < t.limit) && (t.base_row + r < p.m_total);
        const int col_base = n_tile * BN;

        // residual prefetch: a slot may be refilled once the store that last read it has finished reading
        if (slab_full && has_res && lane == 0) {
#pragma unroll
          for (int i = 0; i < MY_GROUPS; ++i) {
            const int slot = (MY_GROUPS == 2) ? i : (it & 1);
            const int g = hsel + 2 * i;
            if (MY_GROUPS == 2 && i == 1) tma_store_wait_read<0>(); else tma_store_wait_read<1>();
            mbar_expect_tx(&my_rbar[slot], C::EPI_REGION_BYTES);
            tma_load_2d(my_stage + slot * C::EPI_REGION_BYTES, &p.map_res, &my_rbar[slot], col_base + g * 64, row0);
          }
        }
        __syncwarp();
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int i = 0; i < MY_GROUPS; ++i) {
          const int slot = one_slot ? 0 : ((MY_GROUPS == 2) ? i : (it & 1));
          const int c0 = (hsel + 2 * i) * 64;
          uint32_t v[2][32];
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c0;
          tmem_ld32(taddr, v[0]);
          tmem_ld32(taddr + 32, v[1]);
          tmem_ld_wait();
          if (i == MY_GROUPS - 1) {   // accumulator fully read by this warp: hand it back to the MMA warp early
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (PAIR) mbar_arrive_cluster(tempty_leader0 + acc * 8); else mbar_arrive(&tempty[acc]);
            }
          }
          if (col_base + c0 >= p.n_total || !(slab_full || valid)) {
            if (lane == 0) tma_store_commit();  // empty group: keeps one store-commit per group for wait_group.read
            continue;
          }
          uint8_t* region = my_stage + slot * C::EPI_REGION_BYTES;
          uint8_t* rowp = region + lane * 128;
          if (slab_full) {
            if (has_res) {
              mbar_wait(&my_rbar[slot], rphase[slot]);
              rphase[slot] ^= 1;
            } else {
              if (lane == 0) {   // the store that last used this slot has read it
                if (one_slot) tma_store_wait_read<0>(); else tma_store_wait_read<1>();
              }
              __syncwarp();
            }
          }
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            int col = col_base + c0 + half * 32;
            const float4* bias4 = reinterpret_cast<const float4*>(sBias + col);
            float f[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = bias4[j];
              f[4 * j + 0] = __uint_as_float(v[half][4 * j + 0]) + b.x;
              f[4 * j + 1] = __uint_as_float(v[half][4 * j + 1]) + b.y;
              f[4 * j + 2] = __uint_as_float(v[half][4 * j + 2]) + b.z;
              f[4 * j + 3] = __uint_as_float(v[half][4 * j + 3]) + b.w;
            }
            if (slab_full) {
              // row `lane` of the 64-column staging region; 16-byte chunks XOR-swizzled by (row & 7)
              const int kbase = half * 4;
              if (has_res) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint4 rr = *reinterpret_cast<const uint4*>(rowp + (((kbase + j) ^ (lane & 7)) << 4));
                  f[8 * j + 0] += bf16_lo(rr.x); f[8 * j + 1] += bf16_hi(rr.x);
                  f[8 * j + 2] += bf16_lo(rr.y); f[8 * j + 3] += bf16_hi(rr.y);
                  f[8 * j + 4] += bf16_lo(rr.z); f[8 * j + 5] += bf16_hi(rr.z);
                  f[8 * j + 6] += bf16_lo(rr.w); f[8 * j + 7] += bf16_hi(rr.w);
                }
              }
              if (p.relu) {   // ReLU folded into the conversion (cvt.rn.relu.bf16x2.f32)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  uint4 o;
                  o.x = pack_bf16_relu(f[8 * j + 0], f[8 * j + 1]);
                  o.y = pack_bf16_relu(f[8 * j + 2], f[8 * j + 3]);
                  o.z = pack_bf16_relu(f[8 * j + 4], f[8 * j + 5]);
                  o.w = pack_bf16_relu(f[8 * j + 6], f[8 * j + 7]);
                  *reinterpret_cast<uint4*>(rowp + (((kbase + j) ^ (lane & 7)) << 4)) = o;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  uint4 o;
                  o.x = pack_bf16(f[8 * j + 0], f[8 * j + 1]);
                  o.y = pack_bf16(f[8 * j + 2], f[8 * j + 3]);
                  o.z = pack_bf16(f[8 * j + 4], f[8 * j + 5]);
                  o.w = pack_bf16(f[8 * j + 6], f[8 * j + 7]);
                  *reinterpret_cast<uint4*>(rowp + (((kbase + j) ^ (lane & 7)) << 4)) = o;
                }
              }
            } else {
              int row = t.base_row + r;
              if (col >= p.n_split) { col -= p.n_split; row += p.split_row_off; }
              const size_t off = static_cast<size_t>(row) * p.ldc + col;
              if (has_res) {
                const uint4* res = reinterpret_cast<const uint4*>(p.residual + off);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint4 rr = __ldg(res + j);
                  f[8 * j + 0] += bf16_lo(rr.x); f[8 * j + 1] += bf16_hi(rr.x);
                  f[8 * j + 2] += bf16_lo(rr.y); f[8 * j + 3] += bf16_hi(rr.y);
                  f[8 * j + 4] += bf16_lo(rr.z); f[8 * j + 5] += bf16_hi(rr.z);
                  f[8 * j + 6] += bf16_lo(rr.w); f[8 * j + 7] += bf16_hi(rr.w);
                }
              }
              if (p.relu) {
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
              }
              uint4* dst = reinterpret_cast<uint4*>(p.out + off);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 o;
                o.x = pack_bf16(f[8 * j + 0], f[8 * j + 1]);
                o.y = pack_bf16(f[8 * j + 2], f[8 * j + 3]);
                o.z = pack_bf16(f[8 * j + 4], f[8 * j + 5]);
                o.w = pack_bf16(f[8 * j + 6], f[8 * j + 7]);
                dst[j] = o;
              }
            }
          }
          if (slab_full) {  // the 64-column region is complete -> one TMA store
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              int scol = col_base + c0, srow = row0;
              if (scol >= p.n_split) { scol -= p.n_split; srow += p.split_row_off; }
              tma_store_2d(&p.map_out, region, scol, srow);
              tma_store_commit();
            }
          } else if (lane == 0) {
            tma_store_commit();
          }
        }
      }
      if (lane == 0) tma_store_wait_all();
    }
  }

  tc_fence_before();
  __syncwarp();
  if (PAIR) {
    cluster_sync_all();   // the leader's MMAs read the peer's shared memory; remote arrivals need live barriers
    if (warp == 1) tmem_dealloc2(tmem_base, C::TMEM_COLS);
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

static int launch_pair(const ConvParams& p, cudaStream_t stream) {
  using C = Cfg<256>;
  static bool attr_set = false;
  if (!attr_set) {
    IO_CUDA(cudaFuncSetAttribute(conv_tc_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::MAX_SMEM));
    attr_set = true;
  }
  IO_REQUIRE(p.stages >= 2 && p.stages <= 8 && p.smem_bytes <= C::MAX_SMEM && !p.b_resident,
             "conv (pair): bad smem plan (%d stages, %d B)", p.stages, p.smem_bytes);
  const int items = ((p.m_tiles + 1) / 2) * p.n_tiles;
  const int max_pairs = num_sms() / 2;
  const int pairs = items < max_pairs ? items : max_pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(320);
  cfg.dynamicSmemBytes = p.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 2;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  IO_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<256, true>, p));
  return IO_OK;
}

template <int BN>
static int launch_bn(const ConvParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    IO_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Cfg<BN>::MAX_SMEM));
    attr_set = true;
  }
  IO_REQUIRE(p.stages >= 2 && p.stages <= 8 && p.smem_bytes <= Cfg<BN>::MAX_SMEM, "conv: bad smem plan (%d stages, %d B)",
             p.stages, p.smem_bytes);
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(320);
  cfg.dynamicSmemBytes = p.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  IO_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, false>, p));
  return IO_OK;
}

int conv_tc_launch(const ConvParams& p, int bn_tile, cudaStream_t stream) {
  if (p.m_tiles <= 0 || p.n_tiles <= 0) return IO_OK;
  switch (bn_tile) {
    case 64: return launch_bn<64>(p, stream);
    case 128: return launch_bn<128>(p, stream);
    case 256: return p.pair ? launch_pair(p, stream) : launch_bn<256>(p, stream);
  }
  set_error("conv_tc_launch: unsupported N tile %d", bn_tile);
  return IO_ERR_ARG;
}

// ---------------------------------------------------------------------------------------------------------
// host-side planning
// ---------------------------------------------------------------------------------------------------------
// ring depth / weight residency for a planned convolution (bn = N tile)
template <int BN>
static void plan_smem_t(ConvParams* p, bool allow_resident) {
  using C = Cfg<BN>;
  const int bres = p->k_iters * C::B_STAGE_BYTES;
  p->b_resident = 0;
  p->stages = C::STAGES;
  if (allow_resident && p->n_tiles == 1) {
    const int room = C::MAX_SMEM - C::FIXED_BYTES - bres;
    const int st = room / A_STAGE_BYTES;
    if (st >= 3) {
      p->b_resident = 1;
      p->stages = st > 8 ? 8 : st;
    }
  }
  p->smem_bytes = p->b_resident ? bres + p->stages * A_STAGE_BYTES + C::FIXED_BYTES
                                : p->stages * C::STAGE_BYTES + C::FIXED_BYTES;
}

static void plan_prefetch(ConvParams* p) {
  static const bool allow = []() {
    const char* e = getenv("INSTAORDER_L2_PREFETCH");
    return e == nullptr || atoi(e) != 0;
  }();
  // worthwhile when a tile has few K blocks (the ring then covers less than a tile); never for stride-2 views
  p->prefetch = (allow && p->mode != CONV_S2 && p->k_iters <= 18) ? 1 : 0;
}

// CTA-pair mode for N = 256 tiles: INSTAORDER_PAIR=0 restores the single-CTA kernel everywhere, 2 forces the pair kernel
// wherever it is legal (read at plan time, so that tests can switch it)
static int pair_mode() {
  const char* e = getenv("INSTAORDER_PAIR");
  return e == nullptr ? 1 : atoi(e);
}

// switches a planned N = 256 convolution to the pair kernel: weight boxes of 128 rows, 4-stage ring of 32 KB
static int plan_pair(ConvParams* p, const void* wgt, uint64_t ktot, uint64_t cout) {
  using C = Cfg<256>;
  p->pair = 0;
  const int pm = pair_mode();
  if (pm == 0 || p->mode == CONV_STEM || p->n_total % 256 != 0) return IO_OK;
  // Measured on B200, same box, 256 pairs (profiles/r02_pair_kernel.md): -8 ... -24 % on every 3x3 layer, on 1x1 layers
  // with K >= 768 and on the dual-source GEMMs; +20 % / 0 % on the short-K (256 / 512) 1x1 expansions, which are
  // HBM-bound and only get their two CTAs' epilogues coupled.  It also needs at least two waves of pair items.
  const bool force = pm == 2;
  const bool pays = p->mode != CONV_GEMM || p->k_iters >= 12 || p->k1 > 0;
  const int items = ((p->m_tiles + 1) / 2) * p->n_tiles;
  if (p->m_tiles < 2 || (!force && (!pays || items < num_sms()))) return IO_OK;
  if (!p->b_mn) {
    const uint64_t wdims[2] = {ktot, cout};
    const uint64_t wstr[1] = {ktot * 2};
    const uint32_t wbox[2] = {64, 128};
    if (int rc = make_tmap_bf16(&p->map_b, wgt, 2, wdims, wstr, wbox, true)) return rc;
  }
  p->pair = 1;
  p->b_resident = 0;
  p->prefetch = 0;
  // no residual tile to prefetch: the two column groups of an epilogue warp share ONE staging slot (the second waits
  // until the first one's store has read it) and the 32 KB saved are a fifth ring stage -- measured: 2 -> 3 -> 4 stages
  // = -27 % / -8 % time, still falling
  p->epi_one_slot = p->residual == nullptr ? 1 : 0;
  p->stages = C::PAIR_STAGES + p->epi_one_slot;
  p->smem_bytes = p->stages * C::PAIR_STAGE_BYTES + C::FIXED_BYTES - (p->epi_one_slot ? C::EPI_BYTES / 2 : 0);
  return IO_OK;
}

static void plan_smem(ConvParams* p, int bn) {
  plan_prefetch(p);
  static const bool allow = []() {
    const char* e = getenv("INSTAORDER_WEIGHT_STATIONARY");
    return e == nullptr || atoi(e) != 0;
  }();
  if (bn == 64) plan_smem_t<64>(p, allow);
  else if (bn == 128) plan_smem_t<128>(p, allow);
  else plan_smem_t<256>(p, allow);
}

static int make_out_maps(ConvParams* p, int rows_total) {
  const uint64_t dims[2] = {static_cast<uint64_t>(p->ldc), static_cast<uint64_t>(rows_total)};
  const uint64_t str[1] = {static_cast<uint64_t>(p->ldc) * 2};
  const uint32_t box[2] = {64, 32};
  int rc = make_tmap_bf16(&p->map_out, p->out, 2, dims, str, box, true);
  if (rc) return rc;
  if (p->residual != nullptr) rc = make_tmap_bf16(&p->map_res, p->residual, 2, dims, str, box, true);
  return rc;
}

int conv_plan(ConvParams* p, int* bn_tile, const ConvDesc& d, const void* x, const void* wgt, const float* bias,
              const void* residual, void* y, int relu) {
  IO_REQUIRE(d.kernel == 1 || d.kernel == 3, "conv: kernel %d not supported (1 or 3)", d.kernel);
  IO_REQUIRE(d.stride == 1 || d.stride == 2, "conv: stride %d not supported", d.stride);
  IO_REQUIRE(d.cin % 64 == 0 && d.cout % 64 == 0, "conv: channels must be multiples of 64 (cin %d cout %d)", d.cin,
             d.cout);
  IO_REQUIRE(d.stride == 1 || (d.h % 2 == 0 && d.w % 2 == 0), "conv: stride 2 needs even h, w");
  IO_REQUIRE(d.b > 0 && d.h > 0 && d.w > 0, "conv: empty input");
  *p = ConvParams{};
  const int h_out = d.h / d.stride, w_out = d.w / d.stride;
  const int ktot = d.kernel * d.kernel * d.cin;
  p->bias = bias;
  p->residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p->out = reinterpret_cast<__nv_bfloat16*>(y);
  p->m_total = d.b * h_out * w_out;
  p->n_total = d.cout;
  p->k_iters = ktot / 64;
  p->kpt = d.cin / 64;
  p->taps_w = d.kernel;
  p->pad = d.kernel / 2;
  p->cin = d.cin;
  p->w_out = w_out;
  p->hw_out = h_out * w_out;
  p->tpr = 1;
  p->ldc = d.cout;
  p->n_split = d.cout;
  p->split_row_off = 0;
  p->relu = relu;
  const int bn = d.cout >= 256 ? 256 : d.cout;
  IO_REQUIRE(d.cout % bn == 0, "conv: cout %d not a multiple of the N tile %d", d.cout, bn);
  *bn_tile = bn;
  p->n_tiles = d.cout / bn;

  const uint64_t wdims[2] = {static_cast<uint64_t>(ktot), static_cast<uint64_t>(d.cout)};
  const uint64_t wstr[1] = {static_cast<uint64_t>(ktot) * 2};
  const uint32_t wbox[2] = {64, static_cast<uint32_t>(bn)};
  int rc = make_tmap_bf16(&p->map_b, wgt, 2, wdims, wstr, wbox, true);
  if (rc) return rc;

  if (d.kernel == 1 && d.stride == 1) {
    p->mode = CONV_GEMM;
    p->rows_per_tile = BM;
    p->tpg = 1; p->bi = 1; p->bh = 1;
    p->m_tiles = (p->m_total + BM - 1) / BM;
    const uint64_t dims[2] = {static_cast<uint64_t>(d.cin), static_cast<uint64_t>(p->m_total)};
    const uint64_t str[1] = {static_cast<uint64_t>(d.cin) * 2};
    const uint32_t box[2] = {64, BM};
    rc = make_tmap_bf16(&p->map_a, x, 2, dims, str, box, true);
  } else {
    IO_REQUIRE(w_out <= BM || (d.stride == 1 && d.kernel == 3), "conv: output width %d > %d needs a 3x3 stride-1 convolution",
               w_out, BM);
    if (w_out > BM) {
      // wide rows (MiDaS decoder, 192 / 384 pixels): tpr equal tiles per row, one image row at a time
      p->tpr = (w_out + BM - 1) / BM;
      while (w_out % p->tpr != 0) ++p->tpr;
      p->bh = 1;
      p->bi = 1;
      p->tpg = h_out * p->tpr;
      p->m_tiles = d.b * p->tpg;
    } else if (p->hw_out <= BM) {
      p->bh = h_out;
      p->bi = BM / p->hw_out;
      p->tpg = 1;
      p->m_tiles = (d.b + p->bi - 1) / p->bi;
    } else {
      p->bh = BM / w_out;
      p->bi = 1;
      p->tpg = (h_out + p->bh - 1) / p->bh;
      p->m_tiles = d.b * p->tpg;
    }
    p->rows_per_tile = p->tpr > 1 ? w_out / p->tpr : p->bi * p->bh * w_out;
    const uint64_t C = d.cin, W = d.w, H = d.h, B = d.b;
    if (d.stride == 1) {
      p->mode = CONV_S1;
      const uint64_t dims[4] = {C, W, H, B};
      const uint64_t str[3] = {C * 2, W * C * 2, H * W * C * 2};
      const uint32_t box[4] = {64, static_cast<uint32_t>(p->tpr > 1 ? p->rows_per_tile : w_out),
                               static_cast<uint32_t>(p->bh), static_cast<uint32_t>(p->bi)};
      rc = make_tmap_bf16(&p->map_a, x, 4, dims, str, box, true);
    } else {
      p->mode = CONV_S2;
      const uint64_t dims[5] = {2 * C, W / 2, 2, H / 2, B};
      const uint64_t str[4] = {2 * C * 2, W * C * 2, 2 * W * C * 2, H * W * C * 2};
      const uint32_t box[5] = {64, static_cast<uint32_t>(w_out), 1, static_cast<uint32_t>(p->bh),
                               static_cast<uint32_t>(p->bi)};
      rc = make_tmap_bf16(&p->map_a, x, 5, dims, str, box, true);
    }
  }
  p->a_bytes = p->rows_per_tile * 128;
  if (rc) return rc;
  plan_smem(p, bn);
  if (bn == 256 && (rc = plan_pair(p, wgt, ktot, d.cout))) return rc;
  return make_out_maps(p, p->m_total);
}

int conv_plan_dual(ConvParams* p, int* bn_tile, const ConvDesc& ds, const void* x, const void* t2, int cmid,
                   const void* wcat, const float* bias, void* y, int relu) {
  IO_REQUIRE(ds.kernel == 1 && cmid % 64 == 0 && cmid > 0, "dual conv: 1x1 downsample and cmid %% 64 == 0 expected");
  // geometry, activation map of x, output map: those of the downsample convolution alone
  int rc = conv_plan(p, bn_tile, ds, x, wcat, bias, nullptr, y, relu);
  if (rc) return rc;
  const int bn = *bn_tile;
  const int ktot = cmid + ds.cin;
  p->k1 = cmid / 64;
  p->k_iters = ktot / 64;
  const uint64_t wdims[2] = {static_cast<uint64_t>(ktot), static_cast<uint64_t>(ds.cout)};
  const uint64_t wstr[1] = {static_cast<uint64_t>(ktot) * 2};
  const uint32_t wbox[2] = {64, static_cast<uint32_t>(bn)};
  if ((rc = make_tmap_bf16(&p->map_b, wcat, 2, wdims, wstr, wbox, true))) return rc;
  const uint64_t dims[2] = {static_cast<uint64_t>(cmid), static_cast<uint64_t>(p->m_total)};
  const uint64_t str[1] = {static_cast<uint64_t>(cmid) * 2};
  const uint32_t box[2] = {64, static_cast<uint32_t>(p->rows_per_tile)};
  if ((rc = make_tmap_bf16(&p->map_a2, t2, 2, dims, str, box, true))) return rc;
  plan_smem(p, bn);
  if (bn == 256 && (rc = plan_pair(p, wcat, ktot, ds.cout))) return rc;
  return IO_OK;
}

int stem_plan(ConvParams* p, int* bn_tile, int pairs, int d, const void* x, const void* wgt, const float* bias,
              void* y) {
  return stem_plan_hw(p, bn_tile, pairs, d, d, x, wgt, bias, y);
}

// network input h x w (`orig` mode: the image's own size rounded to multiples of 32, reference inference.py:401-408)
int stem_plan_hw(ConvParams* p, int* bn_tile, int pairs, int h, int w, const void* x, const void* wgt,
                 const float* bias, void* y) {
  IO_REQUIRE(h % 2 == 0 && h >= 32 && w % 2 == 0 && w >= 32, "stem: input size %d x %d", h, w);
  *p = ConvParams{};
  const int h_out = h / 2, w_out = w / 2;
  const int64_t pitch = io_pair_tensor_row_pitch(w);
  const int hp = h + 6;
  p->bias = bias;
  p->residual = nullptr;
  p->out = reinterpret_cast<__nv_bfloat16*>(y);
  p->mode = CONV_STEM;
  p->m_total = 2 * pairs * h_out * w_out;  // output rows (both directions, pair-major interleaved)
  p->n_total = 128;
  p->k_iters = 7;
  p->kpt = 1;
  p->taps_w = 7;
  p->pad = 3;
  p->cin = 8;
  p->w_out = w_out;
  p->hw_out = h_out * w_out;
  p->tpr = (w_out + BM - 1) / BM;
  p->tpg = h_out * p->tpr;
  p->bi = 1;
  p->bh = 1;
  p->rows_per_tile = w_out < BM ? w_out : BM;
  p->m_tiles = pairs * p->tpg;
  p->n_tiles = 1;
  p->ldc = 64;
  p->n_split = 64;
  p->split_row_off = p->hw_out;
  p->img_mul = 2;
  p->relu = 1;
  p->a_bytes = p->rows_per_tile * 128;
  *bn_tile = 128;
  // overlapping-window view of the padded pair tensor: dim0 = 8 pixels x 8 channels starting at padded pixel 2*wo,
  // dim1 = output column (stride 2 pixels), dim2 = padded input row, dim3 = pair
  const uint64_t dims[4] = {64, static_cast<uint64_t>(w_out), static_cast<uint64_t>(hp), static_cast<uint64_t>(pairs)};
  const uint64_t str[3] = {32, static_cast<uint64_t>(pitch) * 16, static_cast<uint64_t>(hp) * pitch * 16};
  const uint32_t box[4] = {64, static_cast<uint32_t>(p->rows_per_tile), 1, 1};
  int rc = make_tmap_bf16(&p->map_a, x, 4, dims, str, box, true);
  if (rc) return rc;
  const uint64_t wdims[2] = {448, 128};
  const uint64_t wstr[1] = {448 * 2};
  const uint32_t wbox[2] = {64, 128};
  rc = make_tmap_bf16(&p->map_b, wgt, 2, wdims, wstr, wbox, true);
  if (rc) return rc;
  plan_smem(p, 128);
  return make_out_maps(p, p->m_total);
}

int dgrad_plan(ConvParams* p, int* bn_tile, int b, int h, int w, int cin_f, int cout_f, int kernel, const void* dy,
               const void* wgt_f, const float* zero_bias, const void* residual, void* dx) {
  // the data gradient is itself a stride-1 convolution dy[cout_f] -> dx[cin_f]; plan it as such (the weight map
  // built from the dummy K-major view is replaced below)
  int rc = conv_plan(p, bn_tile, ConvDesc{b, h, w, cout_f, cin_f, kernel, 1}, dy, wgt_f, zero_bias, residual, dx, 0);
  if (rc) return rc;
  const uint64_t kk = static_cast<uint64_t>(kernel) * kernel;
  const uint64_t wdims[2] = {kk * cin_f, static_cast<uint64_t>(cout_f)};
  const uint64_t wstr[1] = {kk * cin_f * 2};
  const uint32_t wbox[2] = {64, 64};
  rc = make_tmap_bf16(&p->map_b, wgt_f, 2, wdims, wstr, wbox, true);
  if (rc) return rc;
  p->b_mn = 1;
  p->mn_lbo = mn_lbo();
  p->mn_sbo = mn_sbo();
  p->flip = kernel == 3 ? 1 : 0;
  p->b_tap_cols = cin_f;
  if (p->pair) return IO_OK;   // pair ring: the 64 x 64 MN-major slabs above are what each CTA loads (two per stage)
  if (p->b_resident) {   // weight residency assumes the K-major layout: fall back to the streaming ring
    p->b_resident = 0;
    const int bn = *bn_tile;
    p->stages = bn == 256 ? Cfg<256>::STAGES : (bn == 128 ? Cfg<128>::STAGES : Cfg<64>::STAGES);
    const int stage_bytes = A_STAGE_BYTES + bn * BK * 2;
    const int fixed = bn == 256 ? Cfg<256>::FIXED_BYTES : (bn == 128 ? Cfg<128>::FIXED_BYTES : Cfg<64>::FIXED_BYTES);
    p->smem_bytes = p->stages * stage_bytes + fixed;
  }
  return IO_OK;
}

}  // namespace io

// ---- exported single-conv entry point (per-layer parity tests) ---------------------------------------------
extern "C" int io_conv_bn_act(const void* x_dev, int b, int h, int w, int cin, const void* w_dev, const float* bias_dev,
                              const void* residual_dev, int cout, int kernel, int stride, int relu, void* y_dev,
                              void* stream) {
  IO_REQUIRE(x_dev && w_dev && bias_dev && y_dev, "io_conv_bn_act: null pointer");
  io::ConvParams p;
  int bn = 0;
  io::ConvDesc d{b, h, w, cin, cout, kernel, stride};
  if (residual_dev == nullptr && io::conv_row3_supported(d)) {
    io::HaloParams hp;
    if (int rc = io::conv_row3_plan(&hp, d, x_dev, w_dev, bias_dev, y_dev, relu)) return rc;
    return io::conv_row3_launch(hp, io::as_stream(stream));
  }
  if (residual_dev == nullptr && io::conv_halo_supported(d)) {
    io::HaloParams hp;
    if (int rc = io::conv_halo_plan(&hp, d, x_dev, w_dev, bias_dev, y_dev, relu)) return rc;
    return io::conv_halo_launch(hp, io::as_stream(stream));
  }
  if (residual_dev == nullptr && io::tn_enabled() && io::conv_tn_supported(d)) {
    io::TnParams tp;
    if (int rc = io::conv_tn_plan(&tp, d, x_dev, w_dev, bias_dev, y_dev, relu)) return rc;
    return io::conv_tn_launch(tp, io::as_stream(stream));
  }
  int rc = io::conv_plan(&p, &bn, d, x_dev, w_dev, bias_dev, residual_dev, y_dev, relu);
  if (rc) return rc;
  return io::conv_tc_launch(p, bn, io::as_stream(stream));
}

// conv3 + downsample of a layer's first bottleneck as one GEMM over concatenated K (conv_plan_dual), for the parity test
extern "C" int io_conv_dual(const void* x_dev, int b, int h, int w, int cin, int stride, const void* t2_dev, int cmid,
                            const void* wcat_dev, const float* bias_dev, int cout, int relu, void* y_dev, void* stream) {
  IO_REQUIRE(x_dev && t2_dev && wcat_dev && bias_dev && y_dev, "io_conv_dual: null pointer");
  io::ConvParams p;
  int bn = 0;
  int rc = io::conv_plan_dual(&p, &bn, io::ConvDesc{b, h, w, cin, cout, 1, stride}, x_dev, t2_dev, cmid, wcat_dev,
                              bias_dev, y_dev, relu);
  if (rc) return rc;
  return io::conv_tc_launch(p, bn, io::as_stream(stream));
}
