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
__device__ double2 pairwise_leaf(TriGen& g, int n) {
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
__device__ double2 pairwise(TriGen& g, int n) {
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

// one thread per (image, key)
__global__ void __launch_bounds__(64) whdr_kernel(const int64_t* __restrict__ order, const int64_t* __restrict__ gto,
                                                  const int64_t* __restrict__ gtv, const int64_t* __restrict__ gtc,
                                                  const int64_t* __restrict__ off, const int32_t* __restrict__ nn,
                                                  int batch, double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= batch * 9) return;
  const int b = t / 9;
  TriGen g;
  g.pred = order + off[b];
  g.gt = gto + off[b];
  g.ovl = gtv + off[b];
  g.cnt = gtc + off[b];
  g.n = nn[b];
  g.key = t - 9 * b;
  const int n = g.n >= 2 ? g.count() : 0;
  if (n == 0) {
    out[t] = -1.0;
    return;
  }
  const double2 s = pairwise(g, n);
  out[t] = s.x / s.y * 100.0;
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
  const int threads = batch * 9;
  io::whdr_kernel<<<(threads + 63) / 64, 64, 0, io::as_stream(stream)>>>(order, gt_order, gt_overlap, gt_count, off, n,
                                                                       batch, out);
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}
