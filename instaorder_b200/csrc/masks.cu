// Annotation -> modal mask producer: the COCO mask API calls of the reference's reader
// (maskUtils.frPyObjects / merge / decode, datasets/reader.py:20-66) as
//   host : compressed-string and polygon -> run lengths (tiny, sequential, float64 / integer -- a few hundred runs)
//   GPU  : run lengths -> the N x H x W uint8 {0,1} tensor the gather kernel reads, written straight into HBM:
//          only the runs cross PCIe (a few KB per instance instead of H*W bytes), nothing is decoded on the host.
// The algorithm is that of pycocotools' common/maskApi.c (rleFrString, rleFrPoly, rleMerge(intersect = 0), rleDecode);
// pycocotools is not available in this image, see oracle/coco_mask_oracle.py ("parity unpinned").
#include "common.cuh"

#include <math.h>
#include <limits.h>
#include <algorithm>
#include <vector>

namespace io {

// C's (int) cast of a double as x86-64 performs it (cvttsd2si): NaN / out of range -> INT_MIN
static inline int c_int(double x) {
  if (!(x == x) || x >= 2147483648.0 || x <= -2147483649.0) return INT_MIN;
  return static_cast<int>(x);
}

// One thread per output pixel: pixel (y, x) of the row-major output is element p = x * h + y of the column-major run
// sequence; a binary search in the inclusive prefix sums of each part gives the run index, its parity the value;
// the instance is the union (OR) of its parts.
__global__ void __launch_bounds__(256) masks_from_rle_kernel(const uint32_t* __restrict__ cum,
                                                             const int32_t* __restrict__ comp_off,
                                                             const int32_t* __restrict__ inst_off, int h, int w,
                                                             uint8_t* __restrict__ out) {
  const int inst = blockIdx.y;
  const int c0 = inst_off[inst], c1 = inst_off[inst + 1];
  const int hw = h * w;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < hw; q += gridDim.x * blockDim.x) {
    const int y = q / w, x = q - y * w;
    const uint32_t p = static_cast<uint32_t>(x) * h + y;
    int v = 0;
    for (int c = c0; c < c1 && !v; ++c) {
      const uint32_t* a = cum + comp_off[c];
      int lo = 0, hi = comp_off[c + 1] - comp_off[c];   // number of prefix sums <= p  ==  index of the run holding p
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(a + mid) <= p) lo = mid + 1; else hi = mid;
      }
      v = lo & 1;
    }
    out[static_cast<size_t>(inst) * hw + q] = static_cast<uint8_t>(v);
  }
}

}  // namespace io

using namespace io;

extern "C" int io_rle_from_string(const char* s, int64_t len, uint32_t* counts, int max_counts, int* n_out) {
  IO_REQUIRE(s && counts && n_out && len >= 0, "io_rle_from_string: bad arguments");
  int m = 0;
  int64_t p = 0;
  std::vector<long long> c;
  while (p < len) {
    long long x = 0;
    int k = 0;
    bool more = true;
    while (more) {
      IO_REQUIRE(p < len, "io_rle_from_string: truncated code");
      const long long ch = static_cast<long long>(static_cast<unsigned char>(s[p])) - 48;
      x |= (ch & 0x1f) << (5 * k);
      more = (ch & 0x20) != 0;
      ++p; ++k;
      if (!more && (ch & 0x10)) x |= -1LL << (5 * k);
    }
    if (m > 2) x += c[m - 2];
    c.push_back(x);
    ++m;
  }
  IO_REQUIRE(m <= max_counts, "io_rle_from_string: %d runs, room for %d", m, max_counts);
  for (int i = 0; i < m; ++i) counts[i] = static_cast<uint32_t>(c[i]);
  *n_out = m;
  return IO_OK;
}

extern "C" int io_rle_from_polygon(const double* xy, int k, int h, int w, uint32_t* counts, int max_counts, int* n_out) {
  IO_REQUIRE(xy && counts && n_out && k >= 1 && h > 0 && w > 0, "io_rle_from_polygon: bad arguments");
  const double scale = 5.0;
  std::vector<int> x(k + 1), y(k + 1);
  for (int j = 0; j < k; ++j) {
    x[j] = c_int(scale * xy[2 * j] + .5);
    y[j] = c_int(scale * xy[2 * j + 1] + .5);
  }
  x[k] = x[0];
  y[k] = y[0];
  // dense boundary of the up-sampled polygon: every edge walked along its longer axis
  std::vector<int> u, v;
  for (int j = 0; j < k; ++j) {
    int xs = x[j], xe = x[j + 1], ys = y[j], ye = y[j + 1];
    const int dx = abs(xe - xs), dy = abs(ys - ye);
    const bool flip = (dx >= dy && xs > xe) || (dx < dy && ys > ye);
    if (flip) { std::swap(xs, xe); std::swap(ys, ye); }
    const double s = dx >= dy ? static_cast<double>(ye - ys) / dx : static_cast<double>(xe - xs) / dy;
    if (dx >= dy) {
      for (int d = 0; d <= dx; ++d) {
        const int t = flip ? dx - d : d;
        u.push_back(t + xs);
        v.push_back(c_int(ys + s * t + .5));
      }
    } else {
      for (int d = 0; d <= dy; ++d) {
        const int t = flip ? dy - d : d;
        v.push_back(t + ys);
        u.push_back(c_int(xs + s * t + .5));
      }
    }
  }
  // crossings of the pixel-centre columns, down-sampled
  std::vector<uint32_t> a;
  for (size_t j = 1; j < u.size(); ++j) {
    if (u[j] == u[j - 1]) continue;
    double xd = static_cast<double>(u[j] < u[j - 1] ? u[j] : u[j] - 1);
    xd = (xd + .5) / scale - .5;
    if (floor(xd) != xd || xd < 0 || xd > w - 1) continue;
    double yd = static_cast<double>(v[j] < v[j - 1] ? v[j] : v[j - 1]);
    yd = (yd + .5) / scale - .5;
    if (yd < 0) yd = 0; else if (yd > h) yd = h;
    yd = ceil(yd);
    a.push_back(static_cast<uint32_t>(static_cast<int>(xd) * h + static_cast<int>(yd)));
  }
  a.push_back(static_cast<uint32_t>(h) * static_cast<uint32_t>(w));
  std::sort(a.begin(), a.end());
  uint32_t prev = 0;
  for (size_t j = 0; j < a.size(); ++j) { const uint32_t t = a[j]; a[j] -= prev; prev = t; }
  std::vector<uint32_t> b;
  size_t j = 0;
  b.push_back(a[j++]);
  while (j < a.size()) {
    if (a[j] > 0) {
      b.push_back(a[j++]);
    } else {
      ++j;
      if (j < a.size()) b.back() += a[j++];
    }
  }
  IO_REQUIRE(static_cast<int>(b.size()) <= max_counts, "io_rle_from_polygon: %d runs, room for %d",
             static_cast<int>(b.size()), max_counts);
  for (size_t i = 0; i < b.size(); ++i) counts[i] = b[i];
  *n_out = static_cast<int>(b.size());
  return IO_OK;
}

extern "C" int io_masks_from_rle(const uint32_t* cum_dev, const int32_t* comp_off_dev, const int32_t* inst_off_dev,
                                 int n_inst, int h, int w, uint8_t* out_dev, void* stream) {
  IO_REQUIRE(cum_dev && comp_off_dev && inst_off_dev && out_dev && n_inst >= 0 && h > 0 && w > 0,
             "io_masks_from_rle: bad arguments");
  IO_REQUIRE(static_cast<long long>(h) * w < (1LL << 31), "io_masks_from_rle: image too large");
  if (n_inst == 0) return IO_OK;
  const int hw = h * w;
  dim3 grid(std::min((hw + 255) / 256, 4 * num_sms()), n_inst);
  masks_from_rle_kernel<<<grid, 256, 0, as_stream(stream)>>>(cum_dev, comp_off_dev, inst_off_dev, h, w, out_dev);
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}
