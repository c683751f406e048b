// Validation-time losses of the order models (forward only), reference models/supervised_order.py:
//   InstaOrderNet_od.calculate_loss :60-81, InstaOrderNet_d.forward_only :397-411, OrderNet.forward_only :465-479,
//   InstaOrderNet_o.forward_only :518-533.
// The reference feeds *probabilities* to nn.CrossEntropyLoss (softmax, then log_softmax inside the criterion --
// "double softmax") and sigmoid outputs to nn.BCELoss (log clamped at -100); both quirks are kept.
// The swapped-direction labels of set_input (:38-48, :389-392, :458-463, :514-516) are derived in the kernel.
#include "common.cuh"

namespace io {

__device__ __forceinline__ double ce_on_probs(const float* logit, int k, int label) {
  // p = softmax(logit) (fp32, as torch); CE(p, y) = logsumexp(p) - p[y]
  float m = logit[0];
  for (int i = 1; i < k; ++i) m = fmaxf(m, logit[i]);
  float e[4], s = 0.f;
  for (int i = 0; i < k; ++i) { e[i] = expf(logit[i] - m); s += e[i]; }
  float p[4];
  for (int i = 0; i < k; ++i) p[i] = e[i] / s;
  float pm = p[0];
  for (int i = 1; i < k; ++i) pm = fmaxf(pm, p[i]);
  float s2 = 0.f;
  for (int i = 0; i < k; ++i) s2 += expf(p[i] - pm);
  return static_cast<double>(pm + logf(s2) - p[label]);
}

__device__ __forceinline__ double bce(float logit, float target) {
  const float pr = 1.0f / (1.0f + expf(-logit));
  const float l1 = fmaxf(logf(pr), -100.0f), l0 = fmaxf(logf(1.0f - pr), -100.0f);
  return -static_cast<double>(target * l1 + (1.0f - target) * l0);
}

// out[0] = loss, out[1] = occ loss, out[2] = depth loss (before the 1/world_size factor), fp32.
// sums[0..5]: CE sums over (overlap, distinct, all) and their counts are reduced in double for determinism.
__global__ void __launch_bounds__(256) loss_kernel(const float* __restrict__ logits, int n, int k_total, int occ_off,
                                                   int depth_off, int depth_k, int softmax_occ,
                                                   const float* __restrict__ occ_target,
                                                   const int64_t* __restrict__ class_target,
                                                   const int64_t* __restrict__ is_overlap, int use_masks,
                                                   float overlap_w, float distinct_w, float inv_world,
                                                   float* __restrict__ out) {
  __shared__ double red[6][8];
  double s_occ = 0, s_ovl = 0, s_dis = 0, n_ovl = 0, n_dis = 0, s_all = 0;
  for (int p = threadIdx.x; p < n; p += blockDim.x) {
    const float* l1 = logits + (static_cast<size_t>(p) * 2 + 0) * k_total;
    const float* l2 = l1 + k_total;
    if (occ_off >= 0 && !softmax_occ) {   // BCE on sigmoid outputs; second direction: target columns exchanged
      const float t0 = occ_target[2 * p], t1 = occ_target[2 * p + 1];
      s_occ += bce(l1[occ_off], t0) + bce(l1[occ_off + 1], t1) + bce(l2[occ_off], t1) + bce(l2[occ_off + 1], t0);
    }
    if (depth_off >= 0) {                 // softmax classes; second direction: labels 0 <-> 1, others fixed
      const int y1 = static_cast<int>(class_target[p]);
      const int y2 = y1 == 0 ? 1 : (y1 == 1 ? 0 : y1);
      const double c = ce_on_probs(l1 + depth_off, depth_k, y1) + ce_on_probs(l2 + depth_off, depth_k, y2);
      s_all += c;
      if (use_masks) {
        if (is_overlap[p] == 1) { s_ovl += c; n_ovl += 1; }
        else if (is_overlap[p] == 0) { s_dis += c; n_dis += 1; }
      }
    }
  }
  double v[6] = {s_occ, s_ovl, s_dis, n_ovl, n_dis, s_all};
  for (int i = 0; i < 6; ++i) {
    double x = v[i];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) red[i][threadIdx.x >> 5] = x;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 6; ++i) {
      double x = 0;
      for (int w = 0; w < 8; ++w) x += red[i][w];
      v[i] = x;
    }
    // BCELoss: mean over the n x 2 elements of each direction, the two directions added
    const double occ_loss = (occ_off >= 0 && !softmax_occ) ? v[0] / (2.0 * n) : 0.0;
    double depth_loss = 0.0;
    if (depth_off >= 0) {
      if (use_masks) {
        // CE(p1[mask]) + CE(p2[mask]) = (sum over the subset of both directions) / |subset|; skipped if empty
        const double lo = v[3] > 0 ? v[1] / v[3] : 0.0;
        const double ld = v[4] > 0 ? v[2] / v[4] : 0.0;
        depth_loss = lo * overlap_w + ld * distinct_w;
      } else {
        depth_loss = v[5] / n;
      }
    }
    out[0] = static_cast<float>((depth_loss + occ_loss) * inv_world);
    out[1] = static_cast<float>(occ_loss);
    out[2] = static_cast<float>(depth_loss);
  }
}

}  // namespace io

extern "C" int io_loss_forward(const float* logits, int n, int k_total, int occ_off, int class_off, int class_k,
                               const float* occ_target, const int64_t* class_target, const int64_t* is_overlap,
                               float overlap_w, float distinct_w, int world_size, float* out, void* stream) {
  IO_REQUIRE(logits && out && n > 0 && world_size > 0, "io_loss_forward: bad arguments");
  IO_REQUIRE(occ_off < 0 || (occ_target && occ_off + 2 <= k_total), "io_loss_forward: occlusion head / targets");
  IO_REQUIRE(class_off < 0 || (class_target && class_k >= 2 && class_k <= 4 && class_off + class_k <= k_total),
             "io_loss_forward: class head / targets");
  io::loss_kernel<<<1, 256, 0, io::as_stream(stream)>>>(logits, n, k_total, occ_off, class_off, class_k, 0, occ_target,
                                                        class_target, is_overlap, is_overlap != nullptr ? 1 : 0,
                                                        overlap_w, distinct_w, 1.0f / static_cast<float>(world_size),
                                                        out);
  IO_CUDA(cudaGetLastError());
  return IO_OK;
}
