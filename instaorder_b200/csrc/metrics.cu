// Order metrics, batched over images: occlusion recall / precision / F1 and the nine depth WHDR variants.
// Integer counts are reduced with warp shuffles; the float64 arithmetic replays scikit-learn's / numpy's exact
// operation order (including numpy's pairwise summation), so results are bit-identical to the reference's
// eval_order_recall_precision_f1 (inference.py:794-802) and eval_depth_order_whdr (:757-791).
#include "common.cuh"

#include <math_constants.h>

namespace io {

// one warp per image
__global__ void __launch_bounds__(128) prf_kernel(const int64_t* __restrict__ order, const int64_t* __restrict__ gt,
                                                  const int64_t* __restrict__ off, const int32_t* __restrict__ nn,
                                                  int batch, int zd, double* __restrict__ out) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= batch) return;
  const int n = nn[b];
  const int64_t* o = order + off[b];
  const int64_t* g = gt + off[b];
  int tp = 0, fp = 0, fn = 0, cnt = 0;
  for (int i = lane; i < n * n; i += 32) {
    const int64_t gv = g[i], ov = o[i];
    if (gv != -1) {
      ++cnt;
      tp += (gv == 1) & (ov == 1);
      fp += (gv != 1) & (ov == 1);
      fn += (gv == 1) & (ov != 1);
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    tp += __shfl_xor_sync(0xffffffffu, tp, s);
    fp += __shfl_xor_sync(0xffffffffu, fp, s);
    fn += __shfl_xor_sync(0xffffffffu, fn, s);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
  }
  if (lane == 0) {
    double r, p, f;
    if (cnt == 0) {
      r = p = f = CUDART_NAN;  // the reference (sklearn) raises on an empty selection
    } else {
      const double z = static_cast<double>(zd);
      r = (tp + fn) > 0 ? static_cast<double>(tp) / static_cast<double>(tp + fn) : z;
      p = (tp + fp) > 0 ? static_cast<double>(tp) / static_cast<double>(tp + fp) : z;
      f = (2 * tp + fp + fn) > 0 ? static_cast<double>(2 * tp) / static_cast<double>(2 * tp + fp + fn) : z;
    }
    out[3 * b + 0] = r * 100.0;
    out[3 * b + 1] = p * 100.0;
    out[3 * b + 2] = f * 100.0;
  }
}

// ---- WHDR ---------------------------------------------------------------------------------------------------
struct TriGen {  // walks the strict upper triangle row-major and yields the entries selected by `key`
  const int64_t *pred, *gt, *ovl, *cnt;
  int n, i, j, key;
  __device__ bool selected(int64_t g, int64_t o) const {
    const int ko = key / 3, ke = key - 3 * ko;
    const bool mo = ko == 0 ? (o == 0) : (ko == 1 ? (o == 1) : (o == 0 || o == 1));
    const bool me = ke == 0 ? (g == 2) : (ke == 1 ? (g == 0 || g == 1) : (g == 0 || g == 1 || g == 2));
    return mo && me;
  }
  __device__ void advance() {
    if (++j >= n) { ++i; j = i + 1; }
  }
  __device__ int count() {
    int c = 0;
    for (i = 0, j = 1; i < n - 1; advance()) {
      const size_t k = static_cast<size_t>(i) * n + j;
      c += selected(gt[k], ovl[k]);
    }
    i = 0; j = 1;
    return c;
  }
  // next selected entry: (err * score, score) with score = 2 / count (float64)
  __device__ double2 next() {
    for (;; advance()) {
      const size_t k = static_cast<size_t>(i) * n + j;
      const int64_t g = gt[k];
      if (selected(g, ovl[k])) {
        const double score = 2.0 / static_cast<double>(cnt[k]);
        const double err = (g != pred[k]) ? score : 0.0;
        advance();
        return make_double2(err, score);
      }
    }
  }
};

__device__ __forceinline__ void add2(double2& a, const double2 b) { a.x += b.x; a.y += b.y; }

// numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src: @TYPE@_pairwise_sum), applied to both
// streams at once: n < 8 sequential; n <= 128 eight interleaved accumulators; else split at n/2 rounded down to 8.
template <typename Gen>
__device__ double2 pairwise_leaf(Gen& g, int n) {
  if (n < 8) {
    double2 r = make_double2(0.0, 0.0);
    for (int i = 0; i < n; ++i) add2(r, g.next());
    return r;
  }
  double2 r[8];
  for (int k = 0; k < 8; ++k) r[k] = g.next();
  int i;
  for (i = 8; i < n - (n % 8); i += 8)
    for (int k = 0; k < 8; ++k) add2(r[k], g.next());
  double2 res;
  res.x = ((r[0].x + r[1].x) + (r[2].x + r[3].x)) + ((r[4].x + r[5].x) + (r[6].x + r[7].x));
  res.y = ((r[0].y + r[1].y) + (r[2].y + r[3].y)) + ((r[4].y + r[5].y) + (r[6].y + r[7].y));
  for (; i < n; ++i) add2(res, g.next());
  return res;
}

// the recursion of pairwise_sum unrolled onto an explicit stack (device call stacks are tiny)
template <typename Gen>
__device__ double2 pairwise(Gen& g, int n) {
  struct Frame { int n; int stage; double2 left; };
  Frame st[32];
  int sp = 0;
  st[sp++] = Frame{n, 0, make_double2(0.0, 0.0)};
  double2 ret = make_double2(0.0, 0.0);
  while (sp > 0) {
    Frame& f = st[sp - 1];
    if (f.n <= 128) {
      ret = pairwise_leaf(g, f.n);
      --sp;
      continue;
    }
    int n2 = f.n / 2;
    n2 -= n2 % 8;
    if (f.stage == 0) {
      f.stage = 1;
      st[sp++] = Frame{n2, 0, make_double2(0.0, 0.0)};
    } else if (f.stage == 1) {
      f.left = ret;
      f.stage = 2;
      st[sp++] = Frame{f.n - n2, 0, make_double2(0.0, 0.0)};
    } else {
      ret = make_double2(f.left.x + ret.x, f.left.y + ret.y);
      --sp;
    }
  }
  return ret;
}

// The selected entries of one key out of the per-warp staging arrays (filled coalesced by the whole warp)
struct StagedGen {
  const double* score;
  const uint16_t* mask;      // bits 0..8: key selection, bit 9: prediction != ground truth (err = score, else 0)
  int k, bit;
  __device__ double2 next() {
    for (;; ++k)
      if ((mask[k] >> bit) & 1) {
        const double2 r = make_double2(((mask[k] >> 9) & 1) ? score[k] : 0.0, score[k]);
        ++k;
        return r;
      }
  }
};

constexpr int WH_NMAX = 24;                              // images up to 24 instances go through shared memory
constexpr int WH_TMAX = WH_NMAX * (WH_NMAX - 1) / 2;     // 276 upper-triangle entries
constexpr int WH_WARPS = 4;
constexpr int WH_MASK_BYTES = (WH_TMAX * 2 + 15) / 16 * 16;
constexpr int WH_IDX_BYTES = (4 * WH_TMAX * 2 + 15) / 16 * 16;   // index lists of the four keys of one round
// per warp: score (float64) + mask / error bit + 4 index lists + 9 counts = 5 KB (10 KB before: err as a second float64
// array and all nine lists at once kept only 20 warps per SM resident; ncu: 29 % occupancy, latency-bound)
constexpr int WH_WARP_BYTES = WH_TMAX * 8 + WH_MASK_BYTES + WH_IDX_BYTES + 48;

// One WARP per image.  The first version ran one thread per (image, key): nine threads each walked the four int64
// matrices of their image serially -- uncoalesced 8-byte loads, every matrix read nine times (bench: 2.7 % of the HBM
// roofline on 65,536 images).  Now
//   1. the warp reads the strict upper triangle of the four matrices ONCE, coalesced (entry k of the row-major triangle
//      -> (i, j) in closed form), and stages per entry the two summands (err * score, score; float64) and a 9-bit
//      selection mask in shared memory;
//   2. per key the selected entries are compacted into an index list with ballots (order preserved);
//   3. numpy's pairwise summation (n <= 128: eight interleaved accumulators, then a fixed tree, then the tail) is
//      replayed with EIGHT LANES PER KEY -- lane j owns accumulator r[j], the tree is two shuffle steps in numpy's own
//      association ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) -- four keys at a time.
// Same operations in the same order as numpy, so the results stay bit-identical to the reference (inference.py:757-791).
// Images with more than 24 instances, or a selection longer than 128 entries, take the serial replay.
__global__ void __launch_bounds__(WH_WARPS * 32) whdr_kernel(const int64_t* __restrict__ order,
                                                             const int64_t* __restrict__ gto,
                                                             const int64_t* __restrict__ gtv,
                                                             const int64_t* __restrict__ gtc,
                                                             const int64_t* __restrict__ off,
                                                             const int32_t* __restrict__ nn, int batch,
                                                             double* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t wh_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * WH_WARPS + warp;
  if (b >= batch) return;
  const int n = nn[b];
  const int64_t o0 = off[b];
  if (n > WH_NMAX) {      // rare large image: the serial walk over global memory, one lane per key
    if (lane < 9) {
      TriGen g;
      g.pred = order + o0; g.gt = gto + o0; g.ovl = gtv + o0; g.cnt = gtc + o0;
      g.n = n; g.key = lane;
      const int cnt = g.count();
      if (cnt == 0) {
        out[9 * b + lane] = -1.0;
      } else {
        const double2 s = pairwise(g, cnt);
        out[9 * b + lane] = s.x / s.y * 100.0;
      }
    }
    return;
  }
  double* s_score = reinterpret_cast<double*>(wh_smem + warp * WH_WARP_BYTES);
  uint16_t* s_mask = reinterpret_cast<uint16_t*>(s_score + WH_TMAX);
  uint16_t* s_idx = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(s_mask) + WH_MASK_BYTES);   // [4][WH_TMAX]
  int* s_cnt = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(s_idx) + WH_IDX_BYTES);               // [9]
  const int T = n * (n - 1) / 2;
  // ---- 1. stage the triangle
  for (int k = lane; k < T; k += 32) {
    // row i of the strict upper triangle starts at entry i * n - i * (i + 1) / 2
    int i = static_cast<int>((static_cast<float>(2 * n - 1) -
                              sqrtf(static_cast<float>((2 * n - 1) * (2 * n - 1) - 8 * k))) * 0.5f);
    i = max(0, min(i, n - 2));
    while (i > 0 && i * n - i * (i + 1) / 2 > k) --i;
    while ((i + 1) * n - (i + 1) * (i + 2) / 2 <= k) ++i;
    const int j = k - (i * n - i * (i + 1) / 2) + i + 1;
    const int64_t idx = o0 + static_cast<int64_t>(i) * n + j;
    const int64_t g = gto[idx], o = gtv[idx], c = gtc[idx], pr = order[idx];
    s_score[k] = 2.0 / static_cast<double>(c);
    const uint32_t mo = (o == 0 ? 1u : 0u) | (o == 1 ? 2u : 0u) | ((o == 0 || o == 1) ? 4u : 0u);   // ovlX, ovlO, ovlOX
    const uint32_t me = (g == 2 ? 1u : 0u) | ((g == 0 || g == 1) ? 2u : 0u) | ((g == 0 || g == 1 || g == 2) ? 4u : 0u);
    uint32_t m = (g != pr) ? (1u << 9) : 0u;
#pragma unroll
    for (int ko = 0; ko < 3; ++ko)
#pragma unroll
      for (int ke = 0; ke < 3; ++ke)
        if (((mo >> ko) & 1u) && ((me >> ke) & 1u)) m |= 1u << (ko * 3 + ke);
    s_mask[k] = static_cast<uint16_t>(m);
  }
  __syncwarp();
  // ---- 2. how many entries every key selects (numpy recurses above 128: serial replay out of shared memory then)
  const uint32_t lt = (1u << lane) - 1u;
  int max_cnt = 0;
#pragma unroll 1
  for (int key = 0; key < 9; ++key) {
    int base = 0;
    for (int k0 = 0; k0 < T; k0 += 32) {
      const int k = k0 + lane;
      base += __popc(__ballot_sync(0xffffffffu, k < T && ((s_mask[k] >> key) & 1)));
    }
    if (lane == 0) s_cnt[key] = base;
    max_cnt = max(max_cnt, base);
  }
  __syncwarp();
  if (max_cnt > 128) {
    if (lane < 9) {
      const int cnt = s_cnt[lane];
      if (cnt == 0) {
        out[9 * b + lane] = -1.0;
      } else {
        StagedGen g{s_score, s_mask, 0, lane};
        const double2 s = pairwise(g, cnt);
        out[9 * b + lane] = s.x / s.y * 100.0;
      }
    }
    return;
  }
  // ---- 3. four keys per round: ordered index lists by ballot compaction, then the pairwise sums with eight lanes
  //         per key
  const int j8 = lane & 7;
#pragma unroll 1
  for (int round = 0; round < 3; ++round) {
#pragma unroll 1
    for (int kk = 0; kk < 4; ++kk) {
      const int key = round * 4 + kk;
      if (key >= 9) break;
      int base = 0;
      for (int k0 = 0; k0 < T; k0 += 32) {
        const int k = k0 + lane;
        const bool bit = k < T && ((s_mask[k] >> key) & 1);
        const uint32_t bal = __ballot_sync(0xffffffffu, bit);
        if (bit) s_idx[kk * WH_TMAX + base + __popc(bal & lt)] = static_cast<uint16_t>(k);
        base += __popc(bal);
      }
    }
    __syncwarp();
    const int key = round * 4 + (lane >> 3);
    const int cnt = key < 9 ? s_cnt[key] : 0;
    const uint16_t* ix = s_idx + (lane >> 3) * WH_TMAX;
    const int body = cnt - (cnt & 7);          // entries covered by the eight accumulators
    auto err_of = [&](int e) { return ((s_mask[e] >> 9) & 1) ? s_score[e] : 0.0; };
    double rx = 0.0, ry = 0.0;
    if (cnt >= 8) {
      rx = err_of(ix[j8]);
      ry = s_score[ix[j8]];
      for (int i = 8; i < body; i += 8) {
        rx += err_of(ix[i + j8]);
        ry += s_score[ix[i + j8]];
      }
    }
    // ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7))
    rx += __shfl_down_sync(0xffffffffu, rx, 1, 8);
    ry += __shfl_down_sync(0xffffffffu, ry, 1, 8);
    rx += __shfl_down_sync(0xffffffffu, rx, 2, 8);
    ry += __shfl_down_sync(0xffffffffu, ry, 2, 8);
    rx += __shfl_down_sync(0xffffffffu, rx, 4, 8);
    ry += __shfl_down_sync(0xffffffffu, ry, 4, 8);
    if (j8 == 0 && key < 9) {
      double sx, sy;
      int i;
      if (cnt >= 8) { sx = rx; sy = ry; i = body; } else { sx = 0.0; sy = 0.0; i = 0; }
      for (; i < cnt; ++i) {
        sx += err_of(ix[i]);
        sy += s_score[ix[i]];
      }
      out[9 * b + key] = cnt == 0 ? -1.0 : sx / sy * 100.0;
    }
    __syncwarp();     // the index lists are rebuilt for the next round
  }
}

}  // namespace io

extern "C" int io_metrics_prf(const int64_t* order, const int64_t* gt, const int64_t* off, const int32_t* n, int batch,
                              int zd, double* out, void* stream) {
  IO_REQUIRE(order && gt && off && n && out && batch >= 0, "io_metrics_prf: bad arguments");
  if (batch == 0) return IO_OK;
  io::prf_kernel<<<(batch + 3) / 4, 128, 0, io::as_stream(stream)>>>(order, gt, off, n, batch, zd, out);
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}

extern "C" int io_metrics_whdr(const int64_t* order, const int64_t* gt_order, const int64_t* gt_overlap,
                               const int64_t* gt_count, const int64_t* off, const int32_t* n, int batch, double* out,
                               void* stream) {
  IO_REQUIRE(order && gt_order && gt_overlap && gt_count && off && n && out && batch >= 0,
             "io_metrics_whdr: bad arguments");
  if (batch == 0) return IO_OK;
  constexpr int smem = io::WH_WARPS * io::WH_WARP_BYTES;
  static bool attr_set = false;
  if (!attr_set) {
    IO_CUDA(cudaFuncSetAttribute(io::whdr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  io::whdr_kernel<<<(batch + io::WH_WARPS - 1) / io::WH_WARPS, io::WH_WARPS * 32, smem, io::as_stream(stream)>>>(
      order, gt_order, gt_overlap, gt_count, off, n, batch, out);
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}
