// Back-to-back GEMM: a bottleneck's conv3 (1x1, K1 = C -> N1 = 4C, + folded BN + residual + ReLU) fused with the
// NEXT bottleneck's conv1 (1x1, K2 = 4C -> N2 = C', + folded BN + ReLU)  (reference resnet_cls.py:107-116 followed
// by :99-101 of the next block).  The block output tile is written to HBM once (it is the next block's identity)
// and, while it still sits in shared memory in the 128B-swizzled K-major layout, it is the A operand of the second
// GEMM -- the next block never re-reads it from HBM and one launch per block disappears.
//
// One persistent CTA per SM, same warp roles as conv_tc_kernel.  A CTA owns whole M tiles (128 pixels) and walks
// all N1/256 column tiles of the first GEMM itself, accumulating the second GEMM (D2, N2 <= 256 TMEM columns)
// across them.  TMEM: columns [0,256) first-GEMM accumulator (single buffer), [256, 256+N2) D2.
#include "conv_tc.cuh"

namespace io {

namespace {
constexpr int BM = 128;
constexpr int BK = 64;
constexpr int BN = 256;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;   // 32 KB (also holds one [N2 x 64] block of the second weights)
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int STAGES = 3;
constexpr int REGION_BYTES = 32 * 128;       // 32 rows x 64 columns bf16
constexpr int EPI_BYTES = 16 * REGION_BYTES; // 4 column groups x 4 warps: the [128 x 256] bf16 tile = 64 KB
constexpr int BIAS_BYTES = 4096 + 1024;      // bias1 (<= 1024 floats) + bias2 (<= 256 floats)
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BIAS_BYTES + BAR_BYTES + 1024;
constexpr int TMEM_COLS = 512;
}  // namespace

__global__ void __launch_bounds__(192, 1) conv_fused_kernel(const __grid_constant__ FusedParams fp) {
  const ConvParams& p = fp.c;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint8_t* sEpi = smem + STAGES * STAGE_BYTES;            // region (g, q) at (g * 4 + q) * REGION_BYTES
  float* sBias1 = reinterpret_cast<float*>(sEpi + EPI_BYTES);
  float* sBias2 = sBias1 + 1024;
  uint64_t* full = reinterpret_cast<uint64_t*>(sEpi + EPI_BYTES + BIAS_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;   // first-GEMM accumulator complete
  uint64_t* tempty = tfull + 1;       // ... drained by the epilogue (4 arrivals)
  uint64_t* sready = tempty + 1;      // block-output tile complete in shared memory (4 arrivals)
  uint64_t* sdone = sready + 1;       // second GEMM has finished reading the tile
  uint64_t* d2full = sdone + 1;       // D2 complete for this M tile
  uint64_t* d2empty = d2full + 1;     // ... drained (4 arrivals)
  uint64_t* rbar = d2empty + 1;       // residual tile landed, one per epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rbar + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nt = p.n_tiles;           // column tiles of the first GEMM (N1 / 256)
  const int n2 = fp.n2;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.map_a);
    prefetch_tmap(&p.map_b);
    prefetch_tmap(&p.map_out);
    prefetch_tmap(&p.map_res);
    prefetch_tmap(&fp.map_b2);
    prefetch_tmap(&fp.map_out2);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, 4);
    mbar_init(sready, 4);
    mbar_init(sdone, 1);
    mbar_init(d2full, 1);
    mbar_init(d2empty, 4);
    for (int i = 0; i < 4; ++i) mbar_init(&rbar[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < p.n_total; i += blockDim.x) sBias1[i] = p.bias[i];
  for (int i = threadIdx.x; i < n2; i += blockDim.x) sBias2[i] = fp.bias2[i];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_d2 = tmem_base + BN;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int m_tile = blockIdx.x; m_tile < p.m_tiles; m_tile += gridDim.x) {
        const int row0 = m_tile * BM;
        for (int j = 0; j < nt; ++j) {
          for (int ki = 0; ki < p.k_iters; ++ki) {       // first GEMM: A = conv2 output, B = conv3 weights
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], A_STAGE_BYTES + B_STAGE_BYTES);
            tma_load_2d(sA + stage * A_STAGE_BYTES, &p.map_a, &full[stage], ki * BK, row0);
            tma_load_2d(sB + stage * B_STAGE_BYTES, &p.map_b, &full[stage], ki * BK, j * BN);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          for (int kb = 0; kb < BN / BK; ++kb) {         // second GEMM: B = next conv1 weights, K block j*4+kb
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], n2 * 128);
            tma_load_2d(sB + stage * B_STAGE_BYTES, &fp.map_b2, &full[stage], j * BN + kb * BK, 0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(BM, BN);
      const uint32_t idesc2 = umma_idesc_bf16(BM, n2);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t u = 0;    // (M tile, column tile) counter
      uint32_t mt = 0;   // M tile counter
      for (int m_tile = blockIdx.x; m_tile < p.m_tiles; m_tile += gridDim.x, ++mt) {
        for (int j = 0; j < nt; ++j, ++u) {
          mbar_wait(tempty, (u & 1) ^ 1);                // epilogue has drained the accumulator of (u - 1)
          tc_fence_after();
          for (int ki = 0; ki < p.k_iters; ++ki) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(sA + stage * A_STAGE_BYTES);
            const uint32_t b_addr = smem_u32(sB + stage * B_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(tmem_base, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc1,
                        (ki > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit(tfull);
          // second GEMM on the finished block-output tile (bf16, in the staging regions)
          mbar_wait(sready, u & 1);
          tc_fence_after();
          if (j == 0) {
            mbar_wait(d2empty, (mt & 1) ^ 1);            // previous M tile's D2 has been drained
            tc_fence_after();
          }
          for (int kb = 0; kb < BN / BK; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(sEpi + kb * 4 * REGION_BYTES);
            const uint32_t b_addr = smem_u32(sB + stage * B_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(tmem_d2, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc2,
                        (j > 0 || kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit(sdone);
        }
        umma_commit(d2full);
      }
    }
  } else {
    // ======================= epilogue (warps 2..5) =======================
    const int q = warp & 3;
    uint64_t* my_rbar = &rbar[q];
    uint32_t u = 0, mt = 0, rphase = 0;
    for (int m_tile = blockIdx.x; m_tile < p.m_tiles; m_tile += gridDim.x, ++mt) {
      const int row0 = m_tile * BM + q * 32;
      const bool active = row0 < p.m_total;               // slab has at least one real row (TMA clips the rest)
      for (int j = 0; j < nt; ++j, ++u) {
        if (lane == 0) {
          tma_store_wait_read<0>();                       // our earlier stores no longer read the regions
          if (j > 0) mbar_wait(sdone, (u - 1) & 1);       // ... nor does the previous second GEMM
          if (active) {
            mbar_expect_tx(my_rbar, 4 * REGION_BYTES);
#pragma unroll
            for (int g = 0; g < 4; ++g)
              tma_load_2d(sEpi + (g * 4 + q) * REGION_BYTES, &p.map_res, my_rbar, j * BN + g * 64, row0);
          }
        }
        __syncwarp();
        mbar_wait(tfull, u & 1);
        tc_fence_after();
        if (active) {
          mbar_wait(my_rbar, rphase);
          rphase ^= 1;
        }
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
          uint32_t v[2][32];
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * 64;
          tmem_ld32(taddr, v[0]);
          tmem_ld32(taddr + 32, v[1]);
          tmem_ld_wait();
          uint8_t* region = sEpi + (g * 4 + q) * REGION_BYTES;
          uint8_t* rowp = region + lane * 128;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const float4* bias4 = reinterpret_cast<const float4*>(sBias1 + j * BN + g * 64 + half * 32);
            float f[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b = bias4[i];
              f[4 * i + 0] = __uint_as_float(v[half][4 * i + 0]) + b.x;
              f[4 * i + 1] = __uint_as_float(v[half][4 * i + 1]) + b.y;
              f[4 * i + 2] = __uint_as_float(v[half][4 * i + 2]) + b.z;
              f[4 * i + 3] = __uint_as_float(v[half][4 * i + 3]) + b.w;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4* cp = reinterpret_cast<uint4*>(rowp + (((half * 4 + i) ^ (lane & 7)) << 4));
              if (active) {
                const uint4 rr = *cp;
                f[8 * i + 0] += bf16_lo(rr.x); f[8 * i + 1] += bf16_hi(rr.x);
                f[8 * i + 2] += bf16_lo(rr.y); f[8 * i + 3] += bf16_hi(rr.y);
                f[8 * i + 4] += bf16_lo(rr.z); f[8 * i + 5] += bf16_hi(rr.z);
                f[8 * i + 6] += bf16_lo(rr.w); f[8 * i + 7] += bf16_hi(rr.w);
              }
              uint4 o;
              o.x = pack_bf16(fmaxf(f[8 * i + 0], 0.f), fmaxf(f[8 * i + 1], 0.f));
              o.y = pack_bf16(fmaxf(f[8 * i + 2], 0.f), fmaxf(f[8 * i + 3], 0.f));
              o.z = pack_bf16(fmaxf(f[8 * i + 4], 0.f), fmaxf(f[8 * i + 5], 0.f));
              o.w = pack_bf16(fmaxf(f[8 * i + 6], 0.f), fmaxf(f[8 * i + 7], 0.f));
              *cp = o;   // inactive slabs still write finite values: the second GEMM reads all 128 rows
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if (active) tma_store_2d(&p.map_out, region, j * BN + g * 64, row0);
            tma_store_commit();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(tempty);   // accumulator drained
          mbar_arrive(sready);   // tile complete in shared memory (writes fenced to the async proxy above)
        }
      }
      // ---- second epilogue: D2 = conv1_next(out tile) + bias2, ReLU -> T1 of the next block
      mbar_wait(d2full, mt & 1);
      tc_fence_after();
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();
#pragma unroll 1
      for (int g = 0; g < n2 / 64; ++g) {
        uint32_t v[2][32];
        const uint32_t taddr = tmem_d2 + (static_cast<uint32_t>(q * 32) << 16) + g * 64;
        tmem_ld32(taddr, v[0]);
        tmem_ld32(taddr + 32, v[1]);
        tmem_ld_wait();
        uint8_t* region = sEpi + (g * 4 + q) * REGION_BYTES;
        uint8_t* rowp = region + lane * 128;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const float4* bias4 = reinterpret_cast<const float4*>(sBias2 + g * 64 + half * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b0 = bias4[2 * i], b1 = bias4[2 * i + 1];
            uint4 o;
            o.x = pack_bf16(fmaxf(__uint_as_float(v[half][8 * i + 0]) + b0.x, 0.f),
                            fmaxf(__uint_as_float(v[half][8 * i + 1]) + b0.y, 0.f));
            o.y = pack_bf16(fmaxf(__uint_as_float(v[half][8 * i + 2]) + b0.z, 0.f),
                            fmaxf(__uint_as_float(v[half][8 * i + 3]) + b0.w, 0.f));
            o.z = pack_bf16(fmaxf(__uint_as_float(v[half][8 * i + 4]) + b1.x, 0.f),
                            fmaxf(__uint_as_float(v[half][8 * i + 5]) + b1.y, 0.f));
            o.w = pack_bf16(fmaxf(__uint_as_float(v[half][8 * i + 6]) + b1.z, 0.f),
                            fmaxf(__uint_as_float(v[half][8 * i + 7]) + b1.w, 0.f));
            *reinterpret_cast<uint4*>(rowp + (((half * 4 + i) ^ (lane & 7)) << 4)) = o;
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (active) tma_store_2d(&fp.map_out2, region, g * 64, row0);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d2empty);
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

int conv_fused_launch(const FusedParams& fp, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    IO_CUDA(cudaFuncSetAttribute(conv_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  if (fp.c.m_tiles <= 0) return IO_OK;
  const int grid = fp.c.m_tiles < num_sms() ? fp.c.m_tiles : num_sms();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  IO_CUDA(cudaLaunchKernelEx(&cfg, conv_fused_kernel, fp));
  return IO_OK;
}

// conv3 of one bottleneck (x: [rows, cmid] = conv2 output; w3: [4*cmid][cmid]; residual + ReLU -> y: [rows, 4*cmid])
// fused with conv1 of the next (w1n: [n2][4*cmid]; ReLU -> y2: [rows, n2]).
int conv_fused_plan(FusedParams* fp, int rows, int cmid, int n2, const void* x, const void* w3, const float* bias3,
                    const void* residual, void* y, const void* w1n, const float* bias1n, void* y2) {
  const int n1 = 4 * cmid;
  IO_REQUIRE(cmid % 64 == 0 && n1 % 256 == 0 && n1 <= 1024, "fused conv: C = %d not supported", cmid);
  IO_REQUIRE(n2 % 64 == 0 && n2 >= 64 && n2 <= 256, "fused conv: N2 = %d not supported", n2);
  IO_REQUIRE(residual != nullptr, "fused conv: the block output needs its identity");
  *fp = FusedParams{};
  ConvParams* p = &fp->c;
  p->bias = bias3;
  p->residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p->out = reinterpret_cast<__nv_bfloat16*>(y);
  p->mode = CONV_GEMM;
  p->m_total = rows;
  p->n_total = n1;
  p->k_iters = cmid / 64;
  p->kpt = p->k_iters;
  p->taps_w = 1;
  p->cin = cmid;
  p->a_bytes = A_STAGE_BYTES;
  p->m_tiles = (rows + BM - 1) / BM;
  p->n_tiles = n1 / BN;
  p->rows_per_tile = BM;
  p->tpg = 1; p->bi = 1; p->bh = 1; p->tpr = 1;
  p->ldc = n1;
  p->n_split = n1;
  p->relu = 1;
  fp->bias2 = bias1n;
  fp->n2 = n2;
  int rc;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(cmid), static_cast<uint64_t>(rows)};
    const uint64_t str[1] = {static_cast<uint64_t>(cmid) * 2};
    const uint32_t box[2] = {64, BM};
    if ((rc = make_tmap_bf16(&p->map_a, x, 2, dims, str, box, true))) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(cmid), static_cast<uint64_t>(n1)};
    const uint64_t str[1] = {static_cast<uint64_t>(cmid) * 2};
    const uint32_t box[2] = {64, BN};
    if ((rc = make_tmap_bf16(&p->map_b, w3, 2, dims, str, box, true))) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(n1), static_cast<uint64_t>(rows)};
    const uint64_t str[1] = {static_cast<uint64_t>(n1) * 2};
    const uint32_t box[2] = {64, 32};
    if ((rc = make_tmap_bf16(&p->map_out, y, 2, dims, str, box, true))) return rc;
    if ((rc = make_tmap_bf16(&p->map_res, residual, 2, dims, str, box, true))) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(n1), static_cast<uint64_t>(n2)};
    const uint64_t str[1] = {static_cast<uint64_t>(n1) * 2};
    const uint32_t box[2] = {64, static_cast<uint32_t>(n2)};
    if ((rc = make_tmap_bf16(&fp->map_b2, w1n, 2, dims, str, box, true))) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(n2), static_cast<uint64_t>(rows)};
    const uint64_t str[1] = {static_cast<uint64_t>(n2) * 2};
    const uint32_t box[2] = {64, 32};
    if ((rc = make_tmap_bf16(&fp->map_out2, y2, 2, dims, str, box, true))) return rc;
  }
  return IO_OK;
}

}  // namespace io

// exported for the parity test of the fused pair of convolutions
extern "C" int io_conv_fused_pair(const void* x_dev, int rows, int cmid, const void* w3_dev, const float* bias3_dev,
                                  const void* residual_dev, void* y_dev, const void* w1n_dev, const float* bias1n_dev,
                                  int n2, void* y2_dev, void* stream) {
  IO_REQUIRE(x_dev && w3_dev && bias3_dev && residual_dev && y_dev && w1n_dev && bias1n_dev && y2_dev,
             "io_conv_fused_pair: null pointer");
  io::FusedParams fp;
  int rc = io::conv_fused_plan(&fp, rows, cmid, n2, x_dev, w3_dev, bias3_dev, residual_dev, y_dev, w1n_dev, bias1n_dev,
                               y2_dev);
  if (rc) return rc;
  return io::conv_fused_launch(fp, io::as_stream(stream));
}
