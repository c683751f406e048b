// HBM-bound kernels of the training step around the tensor-core convolutions: train-mode BatchNorm (statistics,
// apply, backward), ReLU masks, max-pool with arg-max, average-pool + FC heads (forward / backward), the loss
// gradient, and the fused optimiser update.
//
// Replaces (reference): nn.BatchNorm2d in training mode + nn.ReLU + `out += identity`
// (models/backbone/resnet_cls.py:96-116, 204-213), nn.MaxPool2d :207, avgpool + fc :214-221, the autograd backward of
// all of them (`loss.backward()`, models/supervised_order.py:92), torch.optim.SGD / Adam
// (models/single_stage_model.py:34-42).
//
// Layout: NHWC bf16 activations seen as [groups][rows][C]; one thread owns 8 consecutive channels (16 bytes) of a
// row, so every global access is a full 16-byte vector and a warp covers whole 128-byte lines.
#include "train.cuh"

namespace io {

namespace {

struct Vec8 {
  float v[8];
};
__device__ __forceinline__ Vec8 unpack8(const uint4& u) {
  Vec8 r;
  r.v[0] = bf16_lo(u.x); r.v[1] = bf16_hi(u.x); r.v[2] = bf16_lo(u.y); r.v[3] = bf16_hi(u.y);
  r.v[4] = bf16_lo(u.z); r.v[5] = bf16_hi(u.z); r.v[6] = bf16_lo(u.w); r.v[7] = bf16_hi(u.w);
  return r;
}
__device__ __forceinline__ uint4 pack8(const Vec8& f) {
  uint4 o;
  o.x = pack_bf16(f.v[0], f.v[1]); o.y = pack_bf16(f.v[2], f.v[3]);
  o.z = pack_bf16(f.v[4], f.v[5]); o.w = pack_bf16(f.v[6], f.v[7]);
  return o;
}

// Geometry of the "slab" kernels: a block owns rows [slab * slab_rows, +slab_rows) of one group; thread t owns
// channel vector t % lanes_c and rows (t / lanes_c) + k * rows_par of the slab.  Slabs are kept small (16 row steps
// per thread, 32-64 KB per tensor) so that even the small late layers give hundreds of blocks, and every thread keeps
// 2-4 independent 16-byte loads per tensor in flight.
struct Slab {
  int c8, lanes_c, rows_par, slab_rows, slabs;
};
static Slab slab_geom(int rows, int c) {
  Slab s;
  s.c8 = c / 8;
  s.lanes_c = s.c8 < 256 ? s.c8 : 256;
  s.rows_par = 256 / s.lanes_c;
  s.slab_rows = s.rows_par * 16;
  s.slabs = (rows + s.slab_rows - 1) / s.slab_rows;
  return s;
}

// block-level reduction of NV per-thread 8-vectors over the rows_par threads that share a channel vector, followed by
// one double atomicAdd per channel: out[(g * NV + i) * C + c]
template <int NV>
__device__ __forceinline__ void slab_reduce_store(float (&acc)[NV][8], int lanes_c, int rows_par, int g, int c,
                                                  double* out) {
  __shared__ float red[256][NV * 8 + 1];
  const int t = threadIdx.x;
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) red[t][i * 8 + j] = acc[i][j];
  __syncthreads();
  for (int idx = t; idx < lanes_c * 8 * NV; idx += 256) {
    const int i = idx / (lanes_c * 8);
    const int rem = idx - i * lanes_c * 8;
    const int lc = rem >> 3, j = rem & 7;
    float s = 0.f;
    for (int r = 0; r < rows_par; ++r) s += red[r * lanes_c + lc][i * 8 + j];
    atomicAdd(out + (static_cast<size_t>(g) * NV + i) * c + lc * 8 + j, static_cast<double>(s));
  }
}

// batch statistics -> per-channel BN coefficients of one (group, channel)
struct BnCoef {
  float mean, invstd, scale, shift;
  double var;
};
__device__ __forceinline__ BnCoef bn_coef(const double* sums, int g, int c, int ch, double m, float gamma, float beta,
                                          float eps) {
  BnCoef k;
  const double mean = __ldcg(sums + (static_cast<size_t>(g) * 2 + 0) * c + ch) / m;
  double var = __ldcg(sums + (static_cast<size_t>(g) * 2 + 1) * c + ch) / m - mean * mean;
  if (var < 0.0) var = 0.0;
  k.var = var;
  k.mean = static_cast<float>(mean);
  k.invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  k.scale = gamma * k.invstd;
  k.shift = beta - k.mean * k.scale;
  return k;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// BatchNorm statistics
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_stats_kernel(const uint4* __restrict__ y, int rows, int c, Slab s,
                                                       double* __restrict__ sums) {
  pdl_sync();
  const int g = blockIdx.y;
  const int t = threadIdx.x;
  const int lc = t % s.lanes_c, roff = t / s.lanes_c;
  const int r0 = blockIdx.x * s.slab_rows;
  const int r1 = min(r0 + s.slab_rows, rows);
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
  const uint4* base = y + (static_cast<size_t>(g) * rows) * s.c8 + lc;
  for (int r = r0 + roff; r < r1; r += 8 * s.rows_par) {
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int rr = r + u * s.rows_par;
      v[u] = rr < r1 ? __ldg(base + static_cast<size_t>(rr) * s.c8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const Vec8 f = unpack8(v[u]);
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[0][j] += f.v[j]; acc[1][j] = fmaf(f.v[j], f.v[j], acc[1][j]); }
    }
  }
  slab_reduce_store<2>(acc, s.lanes_c, s.rows_par, g, c, sums);
}

int bn_stats_launch(const void* y, int groups, int rows, int c, double* sums, cudaStream_t stream) {
  IO_REQUIRE(c % 8 == 0 && c <= 2048 && rows > 0 && groups > 0, "bn_stats: bad shape (rows %d c %d)", rows, c);
  const Slab s = slab_geom(rows, c);
  IO_CUDA(launch_pdl(bn_stats_kernel, dim3(s.slabs, groups), dim3(256), 0, stream, reinterpret_cast<const uint4*>(y),
                     rows, c, s, sums));
  return IO_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// BatchNorm finalize + apply (+ residual) (+ ReLU), one launch: every block derives the coefficients of its 8-channel
// vectors from the batch sums; block (0, 0) also stores them for the backward pass ([4][groups][c]: scale, shift,
// mean, invstd), updates running_mean / running_var (momentum, unbiased variance; one update per group, in group
// order, as the reference's two forward passes do) and clears `zero_me` (the sums buffer of the NEXT BatchNorm).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_apply_kernel(const uint4* __restrict__ y, const uint4* __restrict__ res,
                                                       uint4* __restrict__ a, int groups, int rows, int c, Slab s,
                                                       const double* __restrict__ sums,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float eps, float momentum, float* __restrict__ save,
                                                       float* __restrict__ running_mean,
                                                       float* __restrict__ running_var, double* __restrict__ zero_me,
                                                       int relu, uint8_t* __restrict__ mask_out) {
  pdl_sync();
  const int g = blockIdx.y;
  const int t = threadIdx.x;
  const double m = static_cast<double>(rows);
  if (blockIdx.x == 0 && g == 0) {
    const size_t gc = static_cast<size_t>(groups) * c;
    for (int ch = t; ch < c; ch += 256) {
      float rm = running_mean[ch], rv = running_var[ch];
      const float ga = gamma[ch], be = beta[ch];
      for (int q = 0; q < groups; ++q) {
        const BnCoef k = bn_coef(sums, q, c, ch, m, ga, be, eps);
        save[0 * gc + q * c + ch] = k.scale;
        save[1 * gc + q * c + ch] = k.shift;
        save[2 * gc + q * c + ch] = k.mean;
        save[3 * gc + q * c + ch] = k.invstd;
        const double unbiased = rows > 1 ? k.var * m / (m - 1.0) : k.var;
        rm = (1.0f - momentum) * rm + momentum * k.mean;
        rv = (1.0f - momentum) * rv + momentum * static_cast<float>(unbiased);
      }
      running_mean[ch] = rm;
      running_var[ch] = rv;
    }
    if (zero_me != nullptr)
      for (int i = t; i < groups * 2 * c; i += 256) zero_me[i] = 0.0;
  }
  const int lc = t % s.lanes_c, roff = t / s.lanes_c;
  const int r0 = blockIdx.x * s.slab_rows;
  const int r1 = min(r0 + s.slab_rows, rows);
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = lc * 8 + j;
    const BnCoef k = bn_coef(sums, g, c, ch, m, __ldg(gamma + ch), __ldg(beta + ch), eps);
    sc[j] = k.scale;
    sh[j] = k.shift;
  }
  const size_t base = (static_cast<size_t>(g) * rows) * s.c8 + lc;
  for (int r = r0 + roff; r < r1; r += 4 * s.rows_par) {
    uint4 v[4], w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int rr = r + u * s.rows_par;
      if (rr < r1) {
        const size_t o = base + static_cast<size_t>(rr) * s.c8;
        v[u] = __ldg(y + o);
        if (res != nullptr) w[u] = __ldg(res + o);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int rr = r + u * s.rows_par;
      if (rr >= r1) continue;
      Vec8 f = unpack8(v[u]);
#pragma unroll
      for (int j = 0; j < 8; ++j) f.v[j] = fmaf(f.v[j], sc[j], sh[j]);
      if (res != nullptr) {
        const Vec8 rv = unpack8(w[u]);
#pragma unroll
        for (int j = 0; j < 8; ++j) f.v[j] += rv.v[j];
      }
      if (relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f.v[j] = fmaxf(f.v[j], 0.f);
      }
      const uint4 packed = pack8(f);
      a[base + static_cast<size_t>(rr) * s.c8] = packed;
      if (mask_out != nullptr) {   // ReLU mask of the STORED (bf16) activation, 1 bit per element
        const Vec8 st = unpack8(packed);
        uint32_t m = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) m |= (st.v[j] > 0.f ? 1u : 0u) << j;
        mask_out[base + static_cast<size_t>(rr) * s.c8] = static_cast<uint8_t>(m);
      }
    }
  }
}

int bn_apply_launch(const void* y, const void* residual, void* a, int groups, int rows, int c, const double* sums,
                    const float* gamma, const float* beta, float eps, float momentum, float* save, float* running_mean,
                    float* running_var, double* zero_me, int relu, uint8_t* mask_out, cudaStream_t stream) {
  IO_REQUIRE(c % 8 == 0 && c <= 2048 && rows > 0, "bn_apply: bad shape");
  const Slab s = slab_geom(rows, c);
  IO_CUDA(launch_pdl(bn_apply_kernel, dim3(s.slabs, groups), dim3(256), 0, stream,
                     reinterpret_cast<const uint4*>(y), reinterpret_cast<const uint4*>(residual),
                     reinterpret_cast<uint4*>(a), groups, rows, c, s, sums, gamma, beta, eps, momentum, save,
                     running_mean, running_var, zero_me, relu, mask_out));
  return IO_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// BatchNorm backward (through the optional ReLU):  g = da * mask
//   mask_mode 0: no ReLU;  1: mask = (a > 0) read from the stored activation (residual blocks);
//             2: mask = (y * scale + shift > 0) recomputed from the raw convolution output with the forward pass's
//                own fp32 expression (no residual) -- saves reading `a`;
//             3: bit mask written by bn_apply (`a` then points to one byte per 8 channels) -- 1/16 of the bytes
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const uint4* __restrict__ da, const uint4* __restrict__ a,
                                                            const uint4* __restrict__ y, int groups, int rows, int c,
                                                            Slab s, const float* __restrict__ save, int mask_mode,
                                                            double* __restrict__ red) {
  pdl_sync();
  const int g = blockIdx.y;
  const int t = threadIdx.x;
  const int lc = t % s.lanes_c, roff = t / s.lanes_c;
  const int r0 = blockIdx.x * s.slab_rows;
  const int r1 = min(r0 + s.slab_rows, rows);
  const size_t gc = static_cast<size_t>(groups) * c;
  float sc[8], sh[8], mu[8], is[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = g * c + lc * 8 + j;
    sc[j] = __ldg(save + ch);
    sh[j] = __ldg(save + gc + ch);
    mu[j] = __ldg(save + 2 * gc + ch);
    is[j] = __ldg(save + 3 * gc + ch);
  }
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
  const size_t base = (static_cast<size_t>(g) * rows) * s.c8 + lc;
  for (int r = r0 + roff; r < r1; r += 4 * s.rows_par) {
    uint4 vd[4], vy[4], va[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int rr = r + u * s.rows_par;
      if (rr < r1) {
        const size_t o = base + static_cast<size_t>(rr) * s.c8;
        vd[u] = __ldg(da + o);
        vy[u] = __ldg(y + o);
        if (mask_mode == 1) va[u] = __ldg(a + o);
        else if (mask_mode == 3) va[u].x = __ldg(reinterpret_cast<const uint8_t*>(a) + o);
      } else {
        vd[u] = make_uint4(0, 0, 0, 0);
        vy[u] = make_uint4(0, 0, 0, 0);
        va[u] = make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      Vec8 gd = unpack8(vd[u]);
      const Vec8 yv = unpack8(vy[u]);
      if (mask_mode == 1) {
        const Vec8 av = unpack8(va[u]);
#pragma unroll
        for (int j = 0; j < 8; ++j) gd.v[j] = av.v[j] > 0.f ? gd.v[j] : 0.f;
      } else if (mask_mode == 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) gd.v[j] = fmaf(yv.v[j], sc[j], sh[j]) > 0.f ? gd.v[j] : 0.f;
      } else if (mask_mode == 3) {
#pragma unroll
        for (int j = 0; j < 8; ++j) gd.v[j] = ((va[u].x >> j) & 1u) ? gd.v[j] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[0][j] += gd.v[j];
        acc[1][j] = fmaf(gd.v[j], (yv.v[j] - mu[j]) * is[j], acc[1][j]);
      }
    }
  }
  slab_reduce_store<2>(acc, s.lanes_c, s.rows_par, g, c, red);
}

int bn_bwd_reduce_launch(const void* da, const void* a, const void* y, int groups, int rows, int c, const float* save,
                         int mask_mode, double* red, cudaStream_t stream) {
  IO_REQUIRE(c % 8 == 0 && c <= 2048 && rows > 0, "bn_bwd_reduce: bad shape");
  const Slab s = slab_geom(rows, c);
  IO_CUDA(launch_pdl(bn_bwd_reduce_kernel, dim3(s.slabs, groups), dim3(256), 0, stream,
                     reinterpret_cast<const uint4*>(da), reinterpret_cast<const uint4*>(a),
                     reinterpret_cast<const uint4*>(y), groups, rows, c, s, save, mask_mode, red));
  return IO_OK;
}

// dy = gamma * invstd * (g - sum_g / M - xhat * sum_gx / M); optionally also writes g; block (0, 0) accumulates
// dgamma / dbeta (summed over the groups: both forward passes share the parameters) and clears `zero_me`
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const uint4* __restrict__ da, const uint4* __restrict__ a,
                                                           const uint4* __restrict__ y, uint4* __restrict__ dy,
                                                           uint4* __restrict__ g_out, int groups, int rows, int c,
                                                           Slab s, const float* __restrict__ gamma,
                                                           const float* __restrict__ save,
                                                           const double* __restrict__ red, int mask_mode,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                           double* __restrict__ zero_me) {
  pdl_sync();
  const int g = blockIdx.y;
  const int t = threadIdx.x;
  if (blockIdx.x == 0 && g == 0) {
    for (int ch = t; ch < c; ch += 256) {
      double sg = 0.0, sgx = 0.0;
      for (int q = 0; q < groups; ++q) {
        sg += red[(static_cast<size_t>(q) * 2 + 0) * c + ch];
        sgx += red[(static_cast<size_t>(q) * 2 + 1) * c + ch];
      }
      dbeta[ch] += static_cast<float>(sg);
      dgamma[ch] += static_cast<float>(sgx);
    }
    if (zero_me != nullptr)
      for (int i = t; i < groups * 2 * c; i += 256) zero_me[i] = 0.0;
  }
  const int lc = t % s.lanes_c, roff = t / s.lanes_c;
  const int r0 = blockIdx.x * s.slab_rows;
  const int r1 = min(r0 + s.slab_rows, rows);
  const float inv_m = 1.0f / static_cast<float>(rows);
  const size_t gc = static_cast<size_t>(groups) * c;
  float sc[8], sh[8], mu[8], is[8], k1[8], k2[8], k3[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = lc * 8 + j;
    sc[j] = __ldg(save + g * c + ch);
    sh[j] = __ldg(save + gc + g * c + ch);
    mu[j] = __ldg(save + 2 * gc + g * c + ch);
    is[j] = __ldg(save + 3 * gc + g * c + ch);
    k1[j] = __ldg(gamma + ch) * is[j];
    k2[j] = static_cast<float>(red[(static_cast<size_t>(g) * 2 + 0) * c + ch]) * inv_m;
    k3[j] = static_cast<float>(red[(static_cast<size_t>(g) * 2 + 1) * c + ch]) * inv_m;
  }
  const size_t base = (static_cast<size_t>(g) * rows) * s.c8 + lc;
  for (int r = r0 + roff; r < r1; r += 2 * s.rows_par) {
    uint4 vd[2], vy[2], va[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int rr = r + u * s.rows_par;
      if (rr < r1) {
        const size_t o = base + static_cast<size_t>(rr) * s.c8;
        vd[u] = __ldg(da + o);
        vy[u] = __ldg(y + o);
        if (mask_mode == 1) va[u] = __ldg(a + o);
        else if (mask_mode == 3) va[u].x = __ldg(reinterpret_cast<const uint8_t*>(a) + o);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int rr = r + u * s.rows_par;
      if (rr >= r1) continue;
      const size_t o = base + static_cast<size_t>(rr) * s.c8;
      Vec8 gd = unpack8(vd[u]);
      const Vec8 yv = unpack8(vy[u]);
      if (mask_mode == 1) {
        const Vec8 av = unpack8(va[u]);
#pragma unroll
        for (int j = 0; j < 8; ++j) gd.v[j] = av.v[j] > 0.f ? gd.v[j] : 0.f;
      } else if (mask_mode == 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) gd.v[j] = fmaf(yv.v[j], sc[j], sh[j]) > 0.f ? gd.v[j] : 0.f;
      } else if (mask_mode == 3) {
#pragma unroll
        for (int j = 0; j < 8; ++j) gd.v[j] = ((va[u].x >> j) & 1u) ? gd.v[j] : 0.f;
      }
      if (g_out != nullptr) g_out[o] = pack8(gd);
      Vec8 out;
#pragma unroll
      for (int j = 0; j < 8; ++j) out.v[j] = k1[j] * (gd.v[j] - k2[j] - (yv.v[j] - mu[j]) * is[j] * k3[j]);
      dy[o] = pack8(out);
    }
  }
}

int bn_bwd_apply_launch(const void* da, const void* a, const void* y, void* dy, void* g_out, int groups, int rows,
                        int c, const float* gamma, const float* save, const double* red, int mask_mode, float* dgamma,
                        float* dbeta, double* zero_me, cudaStream_t stream) {
  IO_REQUIRE(c % 8 == 0 && c <= 2048 && rows > 0, "bn_bwd_apply: bad shape");
  const Slab s = slab_geom(rows, c);
  IO_CUDA(launch_pdl(bn_bwd_apply_kernel, dim3(s.slabs, groups), dim3(256), 0, stream,
                     reinterpret_cast<const uint4*>(da), reinterpret_cast<const uint4*>(a),
                     reinterpret_cast<const uint4*>(y), reinterpret_cast<uint4*>(dy), reinterpret_cast<uint4*>(g_out),
                     groups, rows, c, s, gamma, save, red, mask_mode, dgamma, dbeta, zero_me));
  return IO_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Two-phase cooperative BatchNorm kernels: statistics + apply (forward) and reduce + apply (backward) in ONE launch
// each, separated by a grid-wide barrier.  A persistent grid (all blocks co-resident: cooperative launch) walks the
// slabs twice; the per-thread partial sums live in registers across all of a block's slabs, so there is ONE
// shared-memory reduction + atomic flush per block instead of one per slab, and for every tensor that fits the
// 126 MB L2 the second pass never touches HBM.  Halves the number of element-wise launches of a training step.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int n_blocks) {
  __threadfence();      // every thread's atomics / stores of phase 1 are visible before its block signals
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    while (*reinterpret_cast<volatile unsigned int*>(counter) < n_blocks) __nanosleep(40);
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) bn_fwd_fused_kernel(const uint4* __restrict__ y, const uint4* __restrict__ res,
                                                           uint4* __restrict__ a, int groups, int rows, int c, Slab s,
                                                           double* __restrict__ sums, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float eps, float momentum,
                                                           float* __restrict__ save, float* __restrict__ running_mean,
                                                           float* __restrict__ running_var,
                                                           double* __restrict__ zero_me, int relu,
                                                           uint8_t* __restrict__ mask_out,
                                                           unsigned int* __restrict__ barrier) {
  const int g = blockIdx.y;
  const int t = threadIdx.x;
  const int lc = t % s.lanes_c, roff = t / s.lanes_c;
  const size_t base = (static_cast<size_t>(g) * rows) * s.c8 + lc;
  // ---- phase 1: statistics
  {
    float acc[2][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
    for (int slab = blockIdx.x; slab < s.slabs; slab += gridDim.x) {
      const int r0 = slab * s.slab_rows;
      const int r1 = min(r0 + s.slab_rows, rows);
      for (int r = r0 + roff; r < r1; r += 8 * s.rows_par) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int rr = r + u * s.rows_par;
          v[u] = rr < r1 ? __ldg(y + base + static_cast<size_t>(rr) * s.c8) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const Vec8 f = unpack8(v[u]);
#pragma unroll
          for (int j = 0; j < 8; ++j) { acc[0][j] += f.v[j]; acc[1][j] = fmaf(f.v[j], f.v[j], acc[1][j]); }
        }
      }
    }
    slab_reduce_store<2>(acc, s.lanes_c, s.rows_par, g, c, sums);
  }
  grid_barrier(barrier, gridDim.x * gridDim.y);
  // ---- phase 2: coefficients, bookkeeping (block (0,0)), apply
  const double m = static_cast<double>(rows);
  if (blockIdx.x == 0 && g == 0) {
    const size_t gc = static_cast<size_t>(groups) * c;
    for (int ch = t; ch < c; ch += 256) {
      float rm = running_mean[ch], rv = running_var[ch];
      const float ga = gamma[ch], be = beta[ch];
      for (int q = 0; q < groups; ++q) {
        const BnCoef k = bn_coef(sums, q, c, ch, m, ga, be, eps);
        save[0 * gc + q * c + ch] = k.scale;
        save[1 * gc + q * c + ch] = k.shift;
        save[2 * gc + q * c + ch] = k.mean;
        save[3 * gc + q * c + ch] = k.invstd;
        const double unbiased = rows > 1 ? k.var * m / (m - 1.0) : k.var;
        rm = (1.0f - momentum) * rm + momentum * k.mean;
        rv = (1.0f - momentum) * rv + momentum * static_cast<float>(unbiased);
      }
      running_mean[ch] = rm;
      running_var[ch] = rv;
    }
    if (zero_me != nullptr)
      for (int i = t; i < groups * 2 * c; i += 256) zero_me[i] = 0.0;
  }
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = lc * 8 + j;
    const BnCoef k = bn_coef(sums, g, c, ch, m, __ldg(gamma + ch), __ldg(beta + ch), eps);
    sc[j] = k.scale;
    sh[j] = k.shift;
  }
  for (int slab = blockIdx.x; slab < s.slabs; slab += gridDim.x) {
    const int r0 = slab * s.slab_rows;
    const int r1 = min(r0 + s.slab_rows, rows);
    for (int r = r0 + roff; r < r1; r += 4 * s.rows_par) {
      uint4 v[4], w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int rr = r + u * s.rows_par;
        if (rr < r1) {
          const size_t o = base + static_cast<size_t>(rr) * s.c8;
          v[u] = __ldcg(y + o);
          if (res != nullptr) w[u] = __ldg(res + o);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int rr = r + u * s.rows_par;
        if (rr >= r1) continue;
        Vec8 f = unpack8(v[u]);
#pragma unroll
        for (int j = 0; j < 8; ++j) f.v[j] = fmaf(f.v[j], sc[j], sh[j]);
        if (res != nullptr) {
          const Vec8 rv = unpack8(w[u]);
#pragma unroll
          for (int j = 0; j < 8; ++j) f.v[j] += rv.v[j];
        }
        if (relu) {
#pragma unroll
          for (int j = 0; j < 8; ++j) f.v[j] = fmaxf(f.v[j], 0.f);
        }
        const uint4 packed = pack8(f);
        a[base + static_cast<size_t>(rr) * s.c8] = packed;
        if (mask_out != nullptr) {
          const Vec8 st = unpack8(packed);
          uint32_t mk = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) mk |= (st.v[j] > 0.f ? 1u : 0u) << j;
          mask_out[base + static_cast<size_t>(rr) * s.c8] = static_cast<uint8_t>(mk);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) bn_bwd_fused_kernel(const uint4* __restrict__ da, const uint4* __restrict__ a,
                                                           const uint4* __restrict__ y, uint4* __restrict__ dy,
                                                           uint4* __restrict__ g_out, int groups, int rows, int c,
                                                           Slab s, const float* __restrict__ gamma,
                                                           const float* __restrict__ save, double* __restrict__ red,
                                                           int mask_mode, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, double* __restrict__ zero_me,
                                                           unsigned int* __restrict__ barrier) {
  const int g = blockIdx.y;
  const int t = threadIdx.x;
  const int lc = t % s.lanes_c, roff = t / s.lanes_c;
  const size_t gc = static_cast<size_t>(groups) * c;
  const size_t base = (static_cast<size_t>(g) * rows) * s.c8 + lc;
  float sc[8], sh[8], mu[8], is[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = g * c + lc * 8 + j;
    sc[j] = __ldg(save + ch);
    sh[j] = __ldg(save + gc + ch);
    mu[j] = __ldg(save + 2 * gc + ch);
    is[j] = __ldg(save + 3 * gc + ch);
  }
  // ---- phase 1: sum g, sum g * xhat
  {
    float acc[2][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
    for (int slab = blockIdx.x; slab < s.slabs; slab += gridDim.x) {
      const int r0 = slab * s.slab_rows;
      const int r1 = min(r0 + s.slab_rows, rows);
      for (int r = r0 + roff; r < r1; r += 4 * s.rows_par) {
        uint4 vd[4], vy[4], va[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int rr = r + u * s.rows_par;
          if (rr < r1) {
            const size_t o = base + static_cast<size_t>(rr) * s.c8;
            vd[u] = __ldg(da + o);
            vy[u] = __ldg(y + o);
            if (mask_mode == 1) va[u] = __ldg(a + o);
            else if (mask_mode == 3) va[u].x = __ldg(reinterpret_cast<const uint8_t*>(a) + o);
          } else {
            vd[u] = make_uint4(0, 0, 0, 0);
            vy[u] = make_uint4(0, 0, 0, 0);
            va[u] = make_uint4(0, 0, 0, 0);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          Vec8 gd = unpack8(vd[u]);
          const Vec8 yv = unpack8(vy[u]);
          if (mask_mode == 1) {
            const Vec8 av = unpack8(va[u]);
#pragma unroll
            for (int j = 0; j < 8; ++j) gd.v[j] = av.v[j] > 0.f ? gd.v[j] : 0.f;
          } else if (mask_mode == 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) gd.v[j] = fmaf(yv.v[j], sc[j], sh[j]) > 0.f ? gd.v[j] : 0.f;
          } else if (mask_mode == 3) {
#pragma unroll
            for (int j = 0; j < 8; ++j) gd.v[j] = ((va[u].x >> j) & 1u) ? gd.v[j] : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            acc[0][j] += gd.v[j];
            acc[1][j] = fmaf(gd.v[j], (yv.v[j] - mu[j]) * is[j], acc[1][j]);
          }
        }
      }
    }
    slab_reduce_store<2>(acc, s.lanes_c, s.rows_par, g, c, red);
  }
  grid_barrier(barrier, gridDim.x * gridDim.y);
  // ---- phase 2
  if (blockIdx.x == 0 && g == 0) {
    for (int ch = t; ch < c; ch += 256) {
      double sg = 0.0, sgx = 0.0;
      for (int q = 0; q < groups; ++q) {
        sg += __ldcg(red + (static_cast<size_t>(q) * 2 + 0) * c + ch);
        sgx += __ldcg(red + (static_cast<size_t>(q) * 2 + 1) * c + ch);
      }
      dbeta[ch] += static_cast<float>(sg);
      dgamma[ch] += static_cast<float>(sgx);
    }
    if (zero_me != nullptr)
      for (int i = t; i < groups * 2 * c; i += 256) zero_me[i] = 0.0;
  }
  const float inv_m = 1.0f / static_cast<float>(rows);
  float k1[8], k2[8], k3[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = lc * 8 + j;
    k1[j] = __ldg(gamma + ch) * is[j];
    k2[j] = static_cast<float>(__ldcg(red + (static_cast<size_t>(g) * 2 + 0) * c + ch)) * inv_m;
    k3[j] = static_cast<float>(__ldcg(red + (static_cast<size_t>(g) * 2 + 1) * c + ch)) * inv_m;
  }
  for (int slab = blockIdx.x; slab < s.slabs; slab += gridDim.x) {
    const int r0 = slab * s.slab_rows;
    const int r1 = min(r0 + s.slab_rows, rows);
    for (int r = r0 + roff; r < r1; r += 2 * s.rows_par) {
      uint4 vd[2], vy[2], va[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int rr = r + u * s.rows_par;
        if (rr < r1) {
          const size_t o = base + static_cast<size_t>(rr) * s.c8;
          vd[u] = __ldg(da + o);
          vy[u] = __ldg(y + o);
          if (mask_mode == 1) va[u] = __ldg(a + o);
          else if (mask_mode == 3) va[u].x = __ldg(reinterpret_cast<const uint8_t*>(a) + o);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int rr = r + u * s.rows_par;
        if (rr >= r1) continue;
        const size_t o = base + static_cast<size_t>(rr) * s.c8;
        Vec8 gd = unpack8(vd[u]);
        const Vec8 yv = unpack8(vy[u]);
        if (mask_mode == 1) {
          const Vec8 av = unpack8(va[u]);
#pragma unroll
          for (int j = 0; j < 8; ++j) gd.v[j] = av.v[j] > 0.f ? gd.v[j] : 0.f;
        } else if (mask_mode == 2) {
#pragma unroll
          for (int j = 0; j < 8; ++j) gd.v[j] = fmaf(yv.v[j], sc[j], sh[j]) > 0.f ? gd.v[j] : 0.f;
        } else if (mask_mode == 3) {
#pragma unroll
          for (int j = 0; j < 8; ++j) gd.v[j] = ((va[u].x >> j) & 1u) ? gd.v[j] : 0.f;
        }
        if (g_out != nullptr) g_out[o] = pack8(gd);
        Vec8 out;
#pragma unroll
        for (int j = 0; j < 8; ++j) out.v[j] = k1[j] * (gd.v[j] - k2[j] - (yv.v[j] - mu[j]) * is[j] * k3[j]);
        dy[o] = pack8(out);
      }
    }
  }
}

// co-resident capacity of a cooperative 256-thread kernel
template <typename K>
static int coop_capacity(K kernel) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  return per_sm * num_sms();
}

int bn_fwd_fused_launch(const void* y, const void* residual, void* a, int groups, int rows, int c, double* sums,
                        const float* gamma, const float* beta, float eps, float momentum, float* save,
                        float* running_mean, float* running_var, double* zero_me, int relu, uint8_t* mask_out,
                        unsigned int* barrier, cudaStream_t stream) {
  IO_REQUIRE(c % 8 == 0 && c <= 2048 && rows > 0 && groups > 0, "bn_fwd_fused: bad shape");
  Slab s = slab_geom(rows, c);
  static const int cap = coop_capacity(bn_fwd_fused_kernel);
  int gx = cap / groups;
  if (gx > s.slabs) gx = s.slabs;
  if (gx < 1) gx = 1;
  const uint4* yp = reinterpret_cast<const uint4*>(y);
  const uint4* rp = reinterpret_cast<const uint4*>(residual);
  uint4* ap = reinterpret_cast<uint4*>(a);
  void* args[] = {&yp, &rp, &ap, &groups, &rows, &c, &s, &sums, &gamma, &beta, &eps, &momentum, &save, &running_mean,
                  &running_var, &zero_me, &relu, &mask_out, &barrier};
  IO_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(bn_fwd_fused_kernel), dim3(gx, groups), dim3(256), args,
                                      0, stream));
  return IO_OK;
}

int bn_bwd_fused_launch(const void* da, const void* a, const void* y, void* dy, void* g_out, int groups, int rows,
                        int c, const float* gamma, const float* save, double* red, int mask_mode, float* dgamma,
                        float* dbeta, double* zero_me, unsigned int* barrier, cudaStream_t stream) {
  IO_REQUIRE(c % 8 == 0 && c <= 2048 && rows > 0 && groups > 0, "bn_bwd_fused: bad shape");
  Slab s = slab_geom(rows, c);
  static const int cap = coop_capacity(bn_bwd_fused_kernel);
  int gx = cap / groups;
  if (gx > s.slabs) gx = s.slabs;
  if (gx < 1) gx = 1;
  const uint4* dap = reinterpret_cast<const uint4*>(da);
  const uint4* ap = reinterpret_cast<const uint4*>(a);
  const uint4* yp = reinterpret_cast<const uint4*>(y);
  uint4* dyp = reinterpret_cast<uint4*>(dy);
  uint4* gp = reinterpret_cast<uint4*>(g_out);
  void* args[] = {&dap, &ap, &yp, &dyp, &gp, &groups, &rows, &c, &s, &gamma, &save, &red, &mask_mode, &dgamma, &dbeta,
                  &zero_me, &barrier};
  IO_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(bn_bwd_fused_kernel), dim3(gx, groups), dim3(256), args,
                                      0, stream));
  return IO_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// max-pool 3x3 / 2 / 1 with arg-max, and its backward
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool_idx_kernel(const uint4* __restrict__ x, uint4* __restrict__ y,
                                                          uint2* __restrict__ idx, int b, int h, int w, int c8) {
  pdl_sync();
  const int ho = h / 2, wo = w / 2;
  const size_t total = static_cast<size_t>(b) * ho * wo * c8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint32_t i32 = static_cast<uint32_t>(i);   // total < 2^32 (checked at launch): 32-bit divisions, not 64-bit
    const int cg = static_cast<int>(i32 % static_cast<uint32_t>(c8));
    uint32_t t = i32 / static_cast<uint32_t>(c8);
    const int ox = static_cast<int>(t % wo);
    t /= wo;
    const int oy = static_cast<int>(t % ho);
    const int n = static_cast<int>(t / ho);
    float best[8];
    uint32_t bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bi[j] = 0; }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int iy = 2 * oy + k / 3 - 1, ix = 2 * ox + k % 3 - 1;
      if (iy < 0 || iy >= h || ix < 0 || ix >= w) continue;
      const Vec8 f = unpack8(__ldg(x + ((static_cast<size_t>(n) * h + iy) * w + ix) * c8 + cg));
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (f.v[j] > best[j]) { best[j] = f.v[j]; bi[j] = k; }   // strict: the first maximum wins (as ATen)
    }
    Vec8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = best[j];
    y[i] = pack8(o);
    uint2 pk;
    pk.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
    pk.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
    idx[i] = pk;
  }
}

int maxpool_fwd_idx_launch(const void* x, void* y, uint8_t* idx, int b, int h, int w, int c, cudaStream_t stream) {
  IO_REQUIRE(c % 8 == 0 && h % 2 == 0 && w % 2 == 0, "maxpool: bad shape");
  const size_t total = static_cast<size_t>(b) * (h / 2) * (w / 2) * (c / 8);
  if (total == 0) return IO_OK;
  IO_REQUIRE(total < (1ull << 32), "element-wise kernel: %zu work items (32-bit index decoding)", total);
  const int grid = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  IO_CUDA(launch_pdl(maxpool_idx_kernel, dim3(grid), dim3(256), 0, stream, reinterpret_cast<const uint4*>(x),
                     reinterpret_cast<uint4*>(y), reinterpret_cast<uint2*>(idx), b, h, w, c / 8));
  return IO_OK;
}

__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const uint4* __restrict__ dy, const uint2* __restrict__ idx,
                                                          uint4* __restrict__ dx, int b, int h, int w, int c8) {
  pdl_sync();
  const int ho = h / 2, wo = w / 2;
  const size_t total = static_cast<size_t>(b) * h * w * c8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint32_t i32 = static_cast<uint32_t>(i);   // total < 2^32 (checked at launch): 32-bit divisions, not 64-bit
    const int cg = static_cast<int>(i32 % static_cast<uint32_t>(c8));
    uint32_t t = i32 / static_cast<uint32_t>(c8);
    const int ix = static_cast<int>(t % w);
    t /= w;
    const int iy = static_cast<int>(t % h);
    const int n = static_cast<int>(t / h);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // windows containing (iy, ix): iy = 2 oy + d with d in {-1, 0, 1}
    for (int dyy = -1; dyy <= 1; ++dyy) {
      const int ty = iy - dyy;
      if (ty < 0 || (ty & 1)) continue;
      const int oy = ty >> 1;
      if (oy >= ho) continue;
      for (int dxx = -1; dxx <= 1; ++dxx) {
        const int tx = ix - dxx;
        if (tx < 0 || (tx & 1)) continue;
        const int ox = tx >> 1;
        if (ox >= wo) continue;
        const uint32_t k = static_cast<uint32_t>((dyy + 1) * 3 + dxx + 1);
        const size_t o = ((static_cast<size_t>(n) * ho + oy) * wo + ox) * c8 + cg;
        const uint2 id = __ldg(idx + o);
        const Vec8 g = unpack8(__ldg(dy + o));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t kk = ((j < 4 ? id.x : id.y) >> (8 * (j & 3))) & 0xFFu;
          if (kk == k) acc[j] += g.v[j];
        }
      }
    }
    Vec8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = acc[j];
    dx[i] = pack8(o);
  }
}

int maxpool_bwd_launch(const void* dy, const uint8_t* idx, void* dx, int b, int h, int w, int c, cudaStream_t stream) {
  const size_t total = static_cast<size_t>(b) * h * w * (c / 8);
  if (total == 0) return IO_OK;
  IO_REQUIRE(total < (1ull << 32), "element-wise kernel: %zu work items (32-bit index decoding)", total);
  const int grid = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  IO_CUDA(launch_pdl(maxpool_bwd_kernel, dim3(grid), dim3(256), 0, stream, reinterpret_cast<const uint4*>(dy),
                     reinterpret_cast<const uint2*>(idx), reinterpret_cast<uint4*>(dx), b, h, w, c / 8));
  return IO_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// stride-2 helpers
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upsample2_zero_kernel(const uint4* __restrict__ dy, uint4* __restrict__ z, int b,
                                                             int ho, int wo, int c8) {
  pdl_sync();
  const int h = 2 * ho, w = 2 * wo;
  const size_t total = static_cast<size_t>(b) * h * w * c8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint32_t i32 = static_cast<uint32_t>(i);   // total < 2^32 (checked at launch): 32-bit divisions, not 64-bit
    const int cg = static_cast<int>(i32 % static_cast<uint32_t>(c8));
    uint32_t t = i32 / static_cast<uint32_t>(c8);
    const int ix = static_cast<int>(t % w);
    t /= w;
    const int iy = static_cast<int>(t % h);
    const int n = static_cast<int>(t / h);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (((ix | iy) & 1) == 0) v = __ldg(dy + ((static_cast<size_t>(n) * ho + (iy >> 1)) * wo + (ix >> 1)) * c8 + cg);
    z[i] = v;
  }
}

int upsample2_zero_launch(const void* dy, void* z, int b, int ho, int wo, int c, cudaStream_t stream) {
  const size_t total = static_cast<size_t>(b) * ho * wo * 4 * (c / 8);
  if (total == 0) return IO_OK;
  IO_REQUIRE(total < (1ull << 32), "element-wise kernel: %zu work items (32-bit index decoding)", total);
  const int grid = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  IO_CUDA(launch_pdl(upsample2_zero_kernel, dim3(grid), dim3(256), 0, stream, reinterpret_cast<const uint4*>(dy),
                     reinterpret_cast<uint4*>(z), b, ho, wo, c / 8));
  return IO_OK;
}

__global__ void __launch_bounds__(256) scatter_add2_kernel(const uint4* __restrict__ d, uint4* __restrict__ dx, int b,
                                                           int ho, int wo, int c8) {
  pdl_sync();
  const size_t total = static_cast<size_t>(b) * ho * wo * c8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint32_t i32 = static_cast<uint32_t>(i);   // total < 2^32 (checked at launch): 32-bit divisions, not 64-bit
    const int cg = static_cast<int>(i32 % static_cast<uint32_t>(c8));
    uint32_t t = i32 / static_cast<uint32_t>(c8);
    const int ox = static_cast<int>(t % wo);
    t /= wo;
    const int oy = static_cast<int>(t % ho);
    const int n = static_cast<int>(t / ho);
    const size_t o = ((static_cast<size_t>(n) * 2 * ho + 2 * oy) * 2 * wo + 2 * ox) * c8 + cg;
    Vec8 a = unpack8(dx[o]);
    const Vec8 v = unpack8(__ldg(d + i));
#pragma unroll
    for (int j = 0; j < 8; ++j) a.v[j] += v.v[j];
    dx[o] = pack8(a);
  }
}

int scatter_add2_launch(const void* d, void* dx, int b, int ho, int wo, int c, cudaStream_t stream) {
  const size_t total = static_cast<size_t>(b) * ho * wo * (c / 8);
  if (total == 0) return IO_OK;
  IO_REQUIRE(total < (1ull << 32), "element-wise kernel: %zu work items (32-bit index decoding)", total);
  const int grid = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  IO_CUDA(launch_pdl(scatter_add2_kernel, dim3(grid), dim3(256), 0, stream, reinterpret_cast<const uint4*>(d),
                     reinterpret_cast<uint4*>(dx), b, ho, wo, c / 8));
  return IO_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// average pool + FC heads
// ---------------------------------------------------------------------------------------------------------------
constexpr int FC_C = 2048;
constexpr int FC_MAXK = 8;

__global__ void __launch_bounds__(256) pool_fc_fwd_kernel(const uint4* __restrict__ feat, int hw,
                                                          const float* __restrict__ fcw, const float* __restrict__ fcb,
                                                          int k_total, float* __restrict__ pooled,
                                                          float* __restrict__ logits) {
  const int img = blockIdx.x;
  const int t = threadIdx.x;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const uint4* p = feat + static_cast<size_t>(img) * hw * (FC_C / 8) + t;
  for (int i = 0; i < hw; ++i) {
    const Vec8 f = unpack8(__ldg(p + static_cast<size_t>(i) * (FC_C / 8)));
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += f.v[j];
  }
  const float inv = static_cast<float>(hw);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s[j] = s[j] / inv;
    pooled[static_cast<size_t>(img) * FC_C + t * 8 + j] = s[j];
  }
  __shared__ float red[FC_MAXK][8];
  for (int k = 0; k < k_total; ++k) {
    const float4* w4 = reinterpret_cast<const float4*>(fcw + static_cast<size_t>(k) * FC_C + t * 8);
    const float4 a = __ldg(w4), b = __ldg(w4 + 1);
    float d = s[0] * a.x + s[1] * a.y + s[2] * a.z + s[3] * a.w + s[4] * b.x + s[5] * b.y + s[6] * b.z + s[7] * b.w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if ((t & 31) == 0) red[k][t >> 5] = d;
  }
  __syncthreads();
  if (t < k_total) {
    float d = 0.f;
#pragma unroll
    for (int wp = 0; wp < 8; ++wp) d += red[t][wp];
    logits[static_cast<size_t>(img) * k_total + t] = d + fcb[t];
  }
}

int pool_fc_fwd_launch(const void* feat, int hw, int imgs, const float* fcw, const float* fcb, int k_total,
                       float* pooled, float* logits, cudaStream_t stream) {
  IO_REQUIRE(k_total >= 1 && k_total <= FC_MAXK, "pool_fc: %d logits", k_total);
  if (imgs == 0) return IO_OK;
  pool_fc_fwd_kernel<<<imgs, 256, 0, stream>>>(reinterpret_cast<const uint4*>(feat), hw, fcw, fcb, k_total, pooled,
                                               logits);
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}

__global__ void __launch_bounds__(256) pool_fc_bwd_feat_kernel(const float* __restrict__ dlogits,
                                                               const float* __restrict__ fcw, int hw, int k_total,
                                                               uint4* __restrict__ dfeat) {
  const int img = blockIdx.x;
  const int t = threadIdx.x;
  Vec8 d;
#pragma unroll
  for (int j = 0; j < 8; ++j) d.v[j] = 0.f;
  for (int k = 0; k < k_total; ++k) {
    const float dl = __ldg(dlogits + static_cast<size_t>(img) * k_total + k);
    const float4* w4 = reinterpret_cast<const float4*>(fcw + static_cast<size_t>(k) * FC_C + t * 8);
    const float4 a = __ldg(w4), b = __ldg(w4 + 1);
    d.v[0] += dl * a.x; d.v[1] += dl * a.y; d.v[2] += dl * a.z; d.v[3] += dl * a.w;
    d.v[4] += dl * b.x; d.v[5] += dl * b.y; d.v[6] += dl * b.z; d.v[7] += dl * b.w;
  }
  const float inv = static_cast<float>(hw);
#pragma unroll
  for (int j = 0; j < 8; ++j) d.v[j] = d.v[j] / inv;
  const uint4 pk = pack8(d);
  uint4* p = dfeat + static_cast<size_t>(img) * hw * (FC_C / 8) + t;
  for (int i = 0; i < hw; ++i) p[static_cast<size_t>(i) * (FC_C / 8)] = pk;
}

__global__ void __launch_bounds__(256) pool_fc_bwd_w_kernel(const float* __restrict__ dlogits,
                                                            const float* __restrict__ pooled, int imgs, int k_total,
                                                            float* __restrict__ dfcw, float* __restrict__ dfcb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // (k, c)
  if (i >= k_total * FC_C) return;
  const int k = i / FC_C, c = i - k * FC_C;
  float s = 0.f, sb = 0.f;
  for (int n = 0; n < imgs; ++n) {
    const float dl = __ldg(dlogits + static_cast<size_t>(n) * k_total + k);
    s += dl * __ldg(pooled + static_cast<size_t>(n) * FC_C + c);
    sb += dl;
  }
  dfcw[i] += s;
  if (c == 0) dfcb[k] += sb;
}

int pool_fc_bwd_launch(const float* dlogits, const float* pooled, const float* fcw, int hw, int imgs, int k_total,
                       float* dfcw, float* dfcb, void* dfeat, cudaStream_t stream) {
  if (imgs == 0) return IO_OK;
  pool_fc_bwd_feat_kernel<<<imgs, 256, 0, stream>>>(dlogits, fcw, hw, k_total, reinterpret_cast<uint4*>(dfeat));
  IO_CUDA(cudaGetLastError());
  pool_fc_bwd_w_kernel<<<(k_total * FC_C + 255) / 256, 256, 0, stream>>>(dlogits, pooled, imgs, k_total, dfcw, dfcb);
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// loss forward + gradient w.r.t. the logits
// ---------------------------------------------------------------------------------------------------------------
// Reference: InstaOrderNet_od.calculate_loss (models/supervised_order.py:60-81), InstaOrderNet_d.step (:413-438, the
// overlap / distinct weighting applies here too), OrderNet.step (:481-493), InstaOrderNet_o.step (:535-548).
// The reference applies nn.CrossEntropyLoss to softmax *probabilities* (a second log-softmax) and nn.BCELoss to
// sigmoid outputs; both are differentiated exactly as autograd does.
// logits: [2][n][k_total] ([direction][pair]).  dlogits: same shape.
__global__ void __launch_bounds__(256) loss_train_kernel(const float* __restrict__ logits, int n, int k_total,
                                                         int occ_off, int cls_off, int cls_k,
                                                         const float* __restrict__ occ_target,
                                                         const int64_t* __restrict__ class_target,
                                                         const int64_t* __restrict__ is_overlap, int use_masks,
                                                         float overlap_w, float distinct_w, float inv_world,
                                                         float* __restrict__ out, float* __restrict__ dlogits) {
  __shared__ double red[6][8];
  __shared__ double tot[6];
  const int t = threadIdx.x;
  // pass 1: subset sizes
  double cnt_o = 0, cnt_d = 0;
  if (use_masks)
    for (int p = t; p < n; p += 256) {
      if (is_overlap[p] == 1) cnt_o += 1;
      else if (is_overlap[p] == 0) cnt_d += 1;
    }
  {
    double v[2] = {cnt_o, cnt_d};
    for (int i = 0; i < 2; ++i) {
      double x = v[i];
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if ((t & 31) == 0) red[i][t >> 5] = x;
    }
    __syncthreads();
    if (t < 2) {
      double x = 0;
      for (int w = 0; w < 8; ++w) x += red[t][w];
      tot[t] = x;
    }
    __syncthreads();
  }
  const double n_ovl = tot[0], n_dis = tot[1];
  __syncthreads();
  double s_occ = 0, s_ovl = 0, s_dis = 0, s_all = 0;
  for (int p = t; p < n; p += 256) {
    for (int dir = 0; dir < 2; ++dir) {
      const float* z = logits + (static_cast<size_t>(dir) * n + p) * k_total;
      float* dz = dlogits + (static_cast<size_t>(dir) * n + p) * k_total;
      for (int i = 0; i < k_total; ++i) dz[i] = 0.f;
      if (occ_off >= 0) {
        // BCELoss(sigmoid(z), target), mean over n x 2; second direction: target columns exchanged
        for (int i = 0; i < 2; ++i) {
          const float tg = occ_target[2 * p + (dir == 0 ? i : 1 - i)];
          const float pr = 1.0f / (1.0f + expf(-z[occ_off + i]));
          const float l1 = fmaxf(logf(pr), -100.0f), l0 = fmaxf(logf(1.0f - pr), -100.0f);
          s_occ += -static_cast<double>(tg * l1 + (1.0f - tg) * l0);
          // autograd: dL/dpr = (pr - t) / max((1 - pr) * pr, 1e-12) / (2n); dpr/dz = pr (1 - pr)
          const float dpr = (pr - tg) / fmaxf((1.0f - pr) * pr, 1e-12f);
          dz[occ_off + i] = dpr * pr * (1.0f - pr) * inv_world / (2.0f * n);
        }
      }
      if (cls_off >= 0) {
        const int y1 = static_cast<int>(class_target[p]);
        const int y = dir == 0 ? y1 : (y1 == 0 ? 1 : (y1 == 1 ? 0 : y1));
        float m = z[cls_off];
        for (int i = 1; i < cls_k; ++i) m = fmaxf(m, z[cls_off + i]);
        float pr[4], s = 0.f;
        for (int i = 0; i < cls_k; ++i) { pr[i] = expf(z[cls_off + i] - m); s += pr[i]; }
        for (int i = 0; i < cls_k; ++i) pr[i] = pr[i] / s;
        float pm = pr[0];
        for (int i = 1; i < cls_k; ++i) pm = fmaxf(pm, pr[i]);
        float q[4], s2 = 0.f;
        for (int i = 0; i < cls_k; ++i) { q[i] = expf(pr[i] - pm); s2 += q[i]; }
        const double ce = static_cast<double>(pm + logf(s2) - pr[y]);
        // weight of this sample in the loss
        float wgt;
        if (use_masks) {
          const int64_t ov = is_overlap[p];
          if (ov == 1) { s_ovl += ce; wgt = overlap_w / static_cast<float>(n_ovl); }
          else if (ov == 0) { s_dis += ce; wgt = distinct_w / static_cast<float>(n_dis); }
          else wgt = 0.f;
        } else {
          s_all += ce;
          wgt = 1.0f / static_cast<float>(n);
        }
        wgt *= inv_world;
        // dCE/dp_i = softmax(p)_i - [i == y];  dp_i/dz_j = p_i ([i == j] - p_j)
        float dp[4], dot = 0.f;
        for (int i = 0; i < cls_k; ++i) {
          dp[i] = q[i] / s2 - (i == y ? 1.0f : 0.0f);
          dot += dp[i] * pr[i];
        }
        for (int i = 0; i < cls_k; ++i) dz[cls_off + i] = wgt * pr[i] * (dp[i] - dot);
      }
    }
  }
  double v[4] = {s_occ, s_ovl, s_dis, s_all};
  for (int i = 0; i < 4; ++i) {
    double x = v[i];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((t & 31) == 0) red[i][t >> 5] = x;
  }
  __syncthreads();
  if (t == 0) {
    for (int i = 0; i < 4; ++i) {
      double x = 0;
      for (int w = 0; w < 8; ++w) x += red[i][w];
      v[i] = x;
    }
    const double occ_loss = occ_off >= 0 ? v[0] / (2.0 * n) : 0.0;
    double cls_loss = 0.0;
    if (cls_off >= 0) {
      if (use_masks) {
        const double lo = n_ovl > 0 ? v[1] / n_ovl : 0.0;
        const double ld = n_dis > 0 ? v[2] / n_dis : 0.0;
        cls_loss = lo * overlap_w + ld * distinct_w;
      } else {
        cls_loss = v[3] / n;
      }
    }
    out[0] = static_cast<float>((cls_loss + occ_loss) * inv_world);
    out[1] = static_cast<float>(occ_loss);
    out[2] = static_cast<float>(cls_loss);
  }
}

int loss_train_launch(const float* logits, int n, int k_total, int occ_off, int cls_off, int cls_k,
                      const float* occ_target, const int64_t* class_target, const int64_t* is_overlap,
                      float overlap_w, float distinct_w, int world_size, float* out, float* dlogits,
                      cudaStream_t stream) {
  loss_train_kernel<<<1, 256, 0, stream>>>(logits, n, k_total, occ_off, cls_off, cls_k, occ_target, class_target,
                                           is_overlap, is_overlap != nullptr ? 1 : 0, overlap_w, distinct_w,
                                           1.0f / static_cast<float>(world_size), out, dlogits);
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// fused optimiser updates over the flat parameter buffer (+ bf16 copy of the GEMM weights)
// ---------------------------------------------------------------------------------------------------------------
// torch.optim.SGD(lr, momentum=0.9, weight_decay=wd) (single_stage_model.py:34-38): g += wd * w; buf = g on the
// first step, else buf = momentum * buf + g; w -= lr * buf.
__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ w, const float* __restrict__ g,
                                                  float* __restrict__ buf, int64_t n, float lr, float momentum,
                                                  float wd, int first, __nv_bfloat16* __restrict__ w16,
                                                  int64_t n16) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float wi = w[i];
    float gi = g[i] + wd * wi;
    float b = first ? gi : momentum * buf[i] + gi;
    buf[i] = b;
    wi = wi - lr * b;
    w[i] = wi;
    if (i < n16) w16[i] = __float2bfloat16_rn(wi);
  }
}

// torch.optim.Adam(lr, betas=(beta1, 0.999)), eps 1e-8, no weight decay (single_stage_model.py:39-42)
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ w, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n, float lr,
                                                   float beta1, float beta2, float eps, float bc1, float bc2_sqrt,
                                                   __nv_bfloat16* __restrict__ w16, int64_t n16) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float gi = g[i];
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    const float wi = w[i] - (lr / bc1) * (mi / denom);
    w[i] = wi;
    if (i < n16) w16[i] = __float2bfloat16_rn(wi);
  }
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ w16,
                                                        int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    w16[i] = __float2bfloat16_rn(w[i]);
}

int cast_bf16_launch(const float* w, void* w16, int64_t n, cudaStream_t stream) {
  if (n <= 0) return IO_OK;
  const int grid = static_cast<int>(std::min<int64_t>((n + 255) / 256, static_cast<int64_t>(num_sms()) * 16));
  cast_bf16_kernel<<<grid, 256, 0, stream>>>(w, reinterpret_cast<__nv_bfloat16*>(w16), n);
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}

// stem: master [64][7][7][5] fp32 (tap-major, channel-minor) <-> packed two-direction GEMM weights [128][448] bf16
// (K index r * 64 + s * 8 + c, rows 64.. with input channels 0 / 1 exchanged; zero padding taps / channels)
__global__ void __launch_bounds__(256) stem_pack_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ pk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128 * 448) return;
  const int row = i / 448, k = i - row * 448;
  const int r = k >> 6, s = (k >> 3) & 7, c = k & 7;
  float v = 0.f;
  if (s < 7 && c < 5) {
    const int dir = row >> 6, co = row & 63;
    const int sc = (dir == 1 && c < 2) ? 1 - c : c;
    v = w[((co * 7 + r) * 7 + s) * 5 + sc];
  }
  pk[i] = __float2bfloat16_rn(v);
}
__global__ void __launch_bounds__(256) stem_unpack_grad_kernel(const float* __restrict__ scratch,
                                                               float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 245) return;
  const int co = i / 245;
  int rem = i - co * 245;
  const int r = rem / 35;
  rem -= r * 35;
  const int s = rem / 5, c = rem - s * 5;
  const int sc = c < 2 ? 1 - c : c;
  dw[i] += scratch[co * 448 + r * 64 + s * 8 + c] + scratch[(64 + co) * 448 + r * 64 + s * 8 + sc];
}

int stem_pack_launch(const float* w, void* pk, cudaStream_t stream) {
  stem_pack_kernel<<<(128 * 448 + 255) / 256, 256, 0, stream>>>(w, reinterpret_cast<__nv_bfloat16*>(pk));
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}
int stem_unpack_grad_launch(const float* scratch, float* dw, cudaStream_t stream) {
  stem_unpack_grad_kernel<<<(64 * 245 + 255) / 256, 256, 0, stream>>>(scratch, dw);
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}

}  // namespace io

using namespace io;

extern "C" int io_optim_sgd(float* w_dev, const float* g_dev, float* buf_dev, int64_t n, float lr, float momentum,
                            float weight_decay, int first_step, void* w_bf16_dev, int64_t n_bf16, void* stream) {
  IO_REQUIRE(w_dev && g_dev && buf_dev && n >= 0, "io_optim_sgd: bad arguments");
  if (n == 0) return IO_OK;
  const int grid = static_cast<int>(std::min<int64_t>((n + 255) / 256, static_cast<int64_t>(num_sms()) * 16));
  sgd_kernel<<<grid, 256, 0, as_stream(stream)>>>(w_dev, g_dev, buf_dev, n, lr, momentum, weight_decay, first_step,
                                                 reinterpret_cast<__nv_bfloat16*>(w_bf16_dev),
                                                 w_bf16_dev ? n_bf16 : 0);
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}

extern "C" int io_optim_adam(float* w_dev, const float* g_dev, float* m_dev, float* v_dev, int64_t n, float lr,
                             float beta1, float beta2, float eps, int step, void* w_bf16_dev, int64_t n_bf16,
                             void* stream) {
  IO_REQUIRE(w_dev && g_dev && m_dev && v_dev && n >= 0 && step >= 1, "io_optim_adam: bad arguments");
  if (n == 0) return IO_OK;
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.0f - powf(beta2, static_cast<float>(step));
  const int grid = static_cast<int>(std::min<int64_t>((n + 255) / 256, static_cast<int64_t>(num_sms()) * 16));
  adam_kernel<<<grid, 256, 0, as_stream(stream)>>>(w_dev, g_dev, m_dev, v_dev, n, lr, beta1, beta2, eps, bc1,
                                                  sqrtf(bc2), reinterpret_cast<__nv_bfloat16*>(w_bf16_dev),
                                                  w_bf16_dev ? n_bf16 : 0);
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}
