// Device helpers for the two refine losses, shared by the stand-alone entry
// points (loss.cu) and the fused refine engine (refine.cu).
//
// replaces: Optimizer.compute_loss_3d (pipelines/optimizer.py:166-198) and
//           Optimizer.compute_loss_2d (pipelines/optimizer.py:200-237).
#pragma once
#include "common.cuh"

namespace sdfr {

// Exact nearest neighbour of (qx,qy,qz) among `n` staged points.  sklearn's
// KDTree works in float64 on the float32 coordinates (optimizer.py:180-181), so
// the squared distance is accumulated in double: exact for float inputs.
__device__ __forceinline__ void nn_scan(const float* __restrict__ pts, int n, int index_base, float qx, float qy,
                                        float qz, double& best_d2, int& best_i) {
  for (int k = 0; k < n; ++k) {
    const double dx = (double)pts[k * 3] - (double)qx;
    const double dy = (double)pts[k * 3 + 1] - (double)qy;
    const double dz = (double)pts[k * 3 + 2] - (double)qz;
    const double d2 = dx * dx + dy * dy + dz * dz;
    if (d2 < best_d2) {
      best_d2 = d2;
      best_i = index_base + k;
    }
  }
}

// One rendered pixel of the 2D NOCS loss.  Returns delta_m and the minimising
// candidate T*D (cand).  Pixels farther than the 5 px radius all contribute the
// same candidate 0 (value ||colour||), so only the 9x9 window has to be searched.
__device__ __forceinline__ float loss2d_pixel(const float* __restrict__ target, int H, int W, int h, int w,
                                              float c0, float c1, float c2, float& k0, float& k1, float& k2) {
  const int P = H * W;
  float best = INFINITY;
  k0 = k1 = k2 = 0.f;
  for (int dh = -4; dh <= 4; ++dh) {
    const int hh = h + dh;
    if (hh < 0 || hh >= H) continue;
    for (int dw = -4; dw <= 4; ++dw) {
      const int ww = w + dw;
      if (ww < 0 || ww >= W) continue;
      const float dist = sqrtf((float)(dh * dh + dw * dw));
      const float wgt = fmaxf(kWinRadius - dist, 0.f);                  // optimizer.py:225
      const int j = hh * W + ww;
      const float t0 = target[j] * wgt, t1 = target[P + j] * wgt, t2 = target[2 * P + j] * wgt;   // :227
      const float e0 = t0 - c0, e1 = t1 - c1, e2 = t2 - c2;
      const float d = sqrtf(e0 * e0 + e1 * e1 + e2 * e2);               // :232
      if (d < best) { best = d; k0 = t0; k1 = t1; k2 = t2; }
    }
  }
  // does any pixel lie at distance >= radius?  (always true for crops larger than ~7 px)
  const float fh = (float)max(h, H - 1 - h), fw = (float)max(w, W - 1 - w);
  if (sqrtf(fh * fh + fw * fw) >= kWinRadius) {
    const float d = sqrtf(c0 * c0 + c1 * c1 + c2 * c2);
    if (d < best) { best = d; k0 = k1 = k2 = 0.f; }
  }
  return best;
}

}  // namespace sdfr
