// Fused refine engine: the whole Optimizer.optimize loop on the device.
//
// replaces: pipelines/optimizer.py:26-54 (parameter groups; Adam on yaw/trans,
//           SGD on scale/latent) and 56-164 (the loop body), with
//           utils/refinement.py:108-125 (rot_from_yaw) folded into the pose kernel.
//
// One iteration = 9 launches, no host synchronisation, all detections of the batch per launch
// (blockIdx.y / blockIdx.z = detection):
//   MLP forward over the lattice -> band select (one chained-scan kernel) -> MLP forward + input-gradient over
//   the band points only (the normals and d sdf/d latent are needed nowhere else: ~2.5 % of the lattice) ->
//   isosurface projection + camera projection -> splat forward -> 2D-loss pixels and 3D-loss pairs (one launch) ->
//   pixel-gradient records -> splat backward -> chain to (R, t, latent_unit, scale) partials, whose last block
//   per detection runs the update and the next iteration's pose.
// Every reduction is an ordered two-level sum, so a run is bit-reproducible.
#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <map>
#include <tuple>

#include "loss.cuh"

namespace sdfr {

namespace {

constexpr int kMaxLatent = 256;
constexpr float kPruneSlack = 0.005f;     // on top of the pre-selection threshold: twice the accepted fp16 lattice-pass error
constexpr float kPruneMaxThr = 0.15f;     // candidate threshold beyond which the reference is renewed (~12 % of a car lattice)

struct DetState {
  int width, height, n_lidar, target_ready;
  float yaw, trans[3], scale;
  float adam_m[4], adam_v[4];
  int adam_t;
  int iter;                 // history rows written
  float pose[16];           // render pose, row-major 4x4
  float loss2d, loss3d, n2, n3;
  int front_count, skip_reason;
};

struct EngineDev {          // passed by value to the engine kernels
  int batch, L, in0;
  long long ng;             // lattice points per detection
  long long cap;            // surfel capacity per detection (= ng)
  int max_pixels, max_lidar, max_iters;
  float w2d, w3d;
  int nb2, nb3, nbc;        // partial-sum blocks per detection
  DetState* det;            // [B]
  float* latent;            // [B,L]
  float* latent_unit;       // [B,L]
  float* sdf;               // [B,ng]   forward-only lattice pass
  float* dinput;            // [B*ng,in0] d sdf / d [latent, x] of the BAND points only (compact)
  float* surf_pts;          // [B,cap,3]
  float* surf_nrm;          // [B,cap,3]
  float* surf_glat;         // [B,cap,L]
  int* surf_idx;            // [B,cap]
  int* surf_count;          // [B]  (= band count per detection)
  unsigned long long* band_status;   // [B*nblk] chained-scan status words of the band selection
  int* band_ctrl;           // [4]  ticket / finished blocks / epoch of the band selection
  int* chain_done;          // [B]  chain blocks that have published their partials (self-resetting)
  // temporal pruning of the lattice pass (lip > 0): see begin_iteration
  float lip;                // certified upper bound of |d sdf / d latent_unit| over the lattice
  float pre_thr;            // pre-selection threshold of the lattice pass (band + fp16 margin)
  float* sdf_ref;           // [B,ng] lattice-pass sdf at the reference latent of each detection
  float* z_ref;             // [B,L]  that reference latent (unit)
  int* ref_valid;           // [B]    sdf_ref / z_ref describe a completed full evaluation
  float* cand_thr;          // [B]    candidate threshold of the coming iteration
  int* cand_all;            // [B]    the coming iteration evaluates the whole lattice and renews the reference
  int* cand_src;            // [B*ng] candidate rows (global lattice indices, ascending)
  float* cand_sdf;          // [B*ng] their lattice-pass sdf (compact)
  int* cand_start;          // [B]
  int* cand_count;          // [B]
  int* cand_total;          // [1]
  unsigned long long* cand_status;   // chained-scan state of the candidate selection
  int* cand_ctrl;           // [4]
  unsigned long long* lattice_rows;  // [2] rows the pruned lattice passes evaluated; detection-iterations they served
  int* band_det_start;      // [B]
  int* band_total;          // [1]
  int* band_src;            // [B*ng] compact global source indices of the band points
  float* band_sdf;          // [B*ng] sdf of the pre-selected points (second, gradient-carrying evaluation)
  unsigned char* surf_valid;// [B,cap] 1 where the accurate |sdf| < 0.03: the reference's band (grid.py:64)
  float* target;            // [B,3,max_pixels]
  float* lidar;             // [B,max_lidar,3]
  float* l3rec;             // [B,cap,4] unit direction, distance (-1 = unused)
  float* l3dot;             // [B,cap]   u . L_nn
  float* l2rec;             // [B,max_pixels,4]
  double* part2;            // [B,nb2,3]
  double* part3;            // [B,nb3,3]  sum, close count, front count
  float* partc;             // [B,nbc,16+L]
  float* history;           // [B,max_iters,4]
  float* grads;             // [B,5+2L]
  unsigned long long* mask_scratch;   // ReLU sign words of the tensor-core band pass (per engine: engines may overlap)
  int* presel_err;          // [1] max |coarse sdf - accurate sdf| over the pre-selected rows since the last read (float bits)
  float* extents;           // [B,8] label extents: min xyz, max xyz, surfel count, pad
  SplatView* views;         // [B]
};

// ---- iteration begin: pose + unit latent (optimizer.py:87-96) ----------------------
// One thread per detection.  Runs once at the start of sdfr_refine_run (iter_begin_kernel) and then at the end of
// every update (the parameters only change there), so an iteration does not pay a launch for it.
__device__ __forceinline__ void begin_iteration(const EngineDev& E, int b) {
  DetState& D = E.det[b];
  const float c = cosf(D.yaw), s = sinf(D.yaw);
  float* P = D.pose;
  // rot_from_yaw (refinement.py:124) with row 1 negated (optimizer.py:89), then translation (:90)
  P[0] = c;    P[1] = 0.f;   P[2] = s;   P[3] = D.trans[0];
  P[4] = -0.f; P[5] = -1.f;  P[6] = -0.f; P[7] = D.trans[1];
  P[8] = -s;   P[9] = 0.f;   P[10] = c;  P[11] = D.trans[2];
  P[12] = 0.f; P[13] = 0.f;  P[14] = 0.f; P[15] = 1.f;
  // F.normalize(latent, p=2, dim=0), eps 1e-12 (optimizer.py:96)
  float n2 = 0.f;
  for (int k = 0; k < E.L; ++k) n2 += E.latent[b * E.L + k] * E.latent[b * E.L + k];
  const float nrm = fmaxf(sqrtf(n2), 1e-12f);
  for (int k = 0; k < E.L; ++k) E.latent_unit[b * E.L + k] = E.latent[b * E.L + k] / nrm;
  // Temporal pruning of the lattice pass.  The decoder is Lipschitz in its latent input with constant <= E.lip, so
  //   |sdf(x; z)| >= |sdf(x; z_ref)| - lip |z - z_ref|
  // and a lattice point whose reference value satisfies |sdf_ref| >= pre_thr + slack + lip |z - z_ref| cannot be
  // pre-selected now: only the others ("candidates") are evaluated.  The slack covers the error of the fp16 lattice
  // pass on both sides (measured on every pre-selected row, limit pre-selection margin / 2).  Once the candidate
  // threshold has grown past kPruneMaxThr the whole lattice is evaluated again and becomes the new reference.
  if (E.lip > 0.f) {
    float d2 = 0.f;
    for (int k = 0; k < E.L; ++k) {
      const float d = E.latent_unit[b * E.L + k] - E.z_ref[b * E.L + k];
      d2 += d * d;
    }
    const float thr = E.pre_thr + kPruneSlack + E.lip * sqrtf(d2) * 1.001f;
    const bool renew = !E.ref_valid[b] || !(thr <= kPruneMaxThr);       // also a NaN latent
    if (renew) {
      for (int k = 0; k < E.L; ++k) E.z_ref[b * E.L + k] = E.latent_unit[b * E.L + k];
      E.ref_valid[b] = 0;                                                // set again when the new reference is stored
    }
    E.cand_thr[b] = thr;
    E.cand_all[b] = renew ? 1 : 0;
  }
}

__global__ void iter_begin_kernel(EngineDev E) {
  if (threadIdx.x == 0) begin_iteration(E, blockIdx.x);
}

// ---- nearest resize of the CSS NOCS prediction to the crop (optimizer.py:135-137) ----
__global__ void resize_target_kernel(const float* __restrict__ src, int th, int tw, float* __restrict__ dst, int H,
                                     int W, int plane_stride) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= H * W) return;
  const int h = j / W, w = j - h * W;
  const float sh = (float)th / (float)H, sw = (float)tw / (float)W;
  const int ih = min((int)floorf((float)h * sh), th - 1), iw = min((int)floorf((float)w * sw), tw - 1);
  for (int c = 0; c < 3; ++c) dst[c * plane_stride + j] = src[(c * th + ih) * tw + iw];
}

// ---- 2D loss: per rendered pixel (optimizer.py:200-237) --------------------------------
constexpr int LB = 256;

__device__ __forceinline__ void loss2d_block(const EngineDev& E, const int b, const int block) {
  __shared__ double s_sum[LB / 32], s_hw[LB / 32];
  __shared__ int s_cnt[LB / 32];
  const DetState& D = E.det[b];
  const SplatView& V = E.views[b];
  const int H = D.height, W = D.width, P = H * W;
  const int j = block * LB + threadIdx.x;
  double sum = 0.0, hw = 0.0;
  int cnt = 0;
  if (j < P) {
    const float c0 = V.color[j], c1 = V.color[P + j], c2 = V.color[2 * P + j];
    float4 r = make_float4(0.f, 0.f, 0.f, -1.f);
    if (c0 + c1 + c2 != 0.f) {
      const int h = j / W, w = j - h * W;
      hw = (double)(h + w);
      float k0, k1, k2;
      const float* T = E.target + (size_t)b * 3 * E.max_pixels;
      // target planes are packed with stride P for this detection
      const float d = loss2d_pixel(T, H, W, h, w, c0, c1, c2, k0, k1, k2);
      if (d < kNocsThr) {
        cnt = 1;
        sum = (double)d;
        if (d > 0.f) r = make_float4((c0 - k0) / d, (c1 - k1) / d, (c2 - k2) / d, d);
        else r.w = d;
      }
    }
    *reinterpret_cast<float4*>(E.l2rec + ((size_t)b * E.max_pixels + j) * 4) = r;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    hw += __shfl_xor_sync(0xffffffffu, hw, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = sum; s_hw[threadIdx.x >> 5] = hw; s_cnt[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, th = 0.0;
    int tc = 0;
    for (int w = 0; w < LB / 32; ++w) { ts += s_sum[w]; th += s_hw[w]; tc += s_cnt[w]; }
    if (block < E.nb2) {
      double* o = E.part2 + ((size_t)b * E.nb2 + block) * 3;
      o[0] = ts; o[1] = (double)tc; o[2] = th;
    }
  }
}

// ---- 3D loss: exact NN of every front-facing surfel in lidar/scale (optimizer.py:84,166-198) ----
constexpr int STAGE = 1024;

__device__ __forceinline__ void loss3d_block(const EngineDev& E, const int b, const int block) {
  __shared__ float s_pts[STAGE * 3];
  __shared__ double s_sum[LB / 32];
  __shared__ int s_cnt[LB / 32], s_front[LB / 32];
  const DetState& D = E.det[b];
  const SplatView& V = E.views[b];
  const int m = min(E.surf_count[b], (int)E.cap);
  if (block * LB >= m) return;   // partials of these blocks are never read
  const int i = block * LB + threadIdx.x;
  const bool live = i < m && V.front[i];
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (live) { qx = V.cam_v[i * 3]; qy = V.cam_v[i * 3 + 1]; qz = V.cam_v[i * 3 + 2]; }
  const float scale = D.scale;
  const float* lidar = E.lidar + (size_t)b * E.max_lidar * 3;
  double best = INFINITY;
  int bi = -1;
  for (int base = 0; base < D.n_lidar; base += STAGE) {
    const int n = min(STAGE, D.n_lidar - base);
    for (int k = threadIdx.x; k < n * 3; k += LB) s_pts[k] = lidar[base * 3 + k] / scale;   // optimizer.py:84
    __syncthreads();
    if (live) nn_scan(s_pts, n, base, qx, qy, qz, best, bi);
    __syncthreads();
  }
  float dist = -1.f, ux = 0.f, uy = 0.f, uz = 0.f, dot = 0.f;
  int close = 0;
  const double radius = 0.2 / (double)scale;                       // optimizer.py:188
  if (live && bi >= 0 && sqrt(best) < radius) {
    close = 1;
    const float lx = lidar[bi * 3] / scale, ly = lidar[bi * 3 + 1] / scale, lz = lidar[bi * 3 + 2] / scale;
    const float ex = lx - qx, ey = ly - qy, ez = lz - qz;
    dist = sqrtf(ex * ex + ey * ey + ez * ez);                      // optimizer.py:189
    if (dist > 0.f) { ux = ex / dist; uy = ey / dist; uz = ez / dist; }
    dot = ux * lx + uy * ly + uz * lz;
  }
  if (i < m) {
    *reinterpret_cast<float4*>(E.l3rec + ((size_t)b * E.cap + i) * 4) = make_float4(ux, uy, uz, dist);
    E.l3dot[(size_t)b * E.cap + i] = dot;
  }
  double s = close ? (double)dist : 0.0;
  int c = close, f = live ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
    f += __shfl_xor_sync(0xffffffffu, f, o);
  }
  if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = s; s_cnt[threadIdx.x >> 5] = c; s_front[threadIdx.x >> 5] = f; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0;
    int tc = 0, tf = 0;
    for (int w = 0; w < LB / 32; ++w) { ts += s_sum[w]; tc += s_cnt[w]; tf += s_front[w]; }
    double* o = E.part3 + ((size_t)b * E.nb3 + block) * 3;
    o[0] = ts; o[1] = (double)tc; o[2] = (double)tf;
  }
}

// Both losses in one launch: blocks [0, n2) take the 2D-loss pixels, blocks [n2, n2 + nb3) the 3D-loss surfels
// (they are independent, so the two also overlap on the device).
__global__ void __launch_bounds__(LB) losses_batch_kernel(EngineDev E, int n2) {
  if ((int)blockIdx.x < n2) loss2d_block(E, blockIdx.y, blockIdx.x);
  else loss3d_block(E, blockIdx.y, (int)blockIdx.x - n2);
}

// ---- per-pixel gradient records for the surfel gather -------------------------------------
__global__ void __launch_bounds__(LB) grad_prep_kernel(EngineDev E) {
  __shared__ float s_scale;
  const int b = blockIdx.y;
  DetState& D = E.det[b];
  const SplatView& V = E.views[b];
  const int P = D.width * D.height;
  if (threadIdx.x < 32) {     // warp 0: fixed-order (deterministic) reduction of the block partials
    const int nb = (P + LB - 1) / LB;
    double s = 0.0, c = 0.0, hw = 0.0;
    const double* p = E.part2 + (size_t)b * E.nb2 * 3;
    for (int k = threadIdx.x; k < nb; k += 32) { s += p[k * 3]; c += p[k * 3 + 1]; hw += p[k * 3 + 2]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
      hw += __shfl_xor_sync(0xffffffffu, hw, o);
    }
    if (threadIdx.x == 0) {
      float loss, n;
      if (hw == 0.0) { loss = 0.f; n = 0.f; }                         // optimizer.py:214 quirk
      else { n = (float)c; loss = c > 0.0 ? (float)(s / c) : __int_as_float(0x7fc00000); }
      s_scale = n > 0.f ? E.w2d / n : 0.f;
      if (blockIdx.x == 0) { D.loss2d = loss; D.n2 = n; }
    }
  }
  __syncthreads();
  const int j = blockIdx.x * LB + threadIdx.x;
  if (j >= P) return;
  const float sc = s_scale;
  const float4 r = *reinterpret_cast<const float4*>(E.l2rec + ((size_t)b * E.max_pixels + j) * 4);
  const float* raw = V.pix_raw + (size_t)j * 8;
  float g0 = 0.f, g1 = 0.f, g2 = 0.f;
  if (r.w >= 0.f) {
    g0 = raw[0] <= 1.f ? r.x * sc : 0.f;
    g1 = raw[1] <= 1.f ? r.y * sc : 0.f;
    g2 = raw[2] <= 1.f ? r.z * sc : 0.f;
  }
  const float G = raw[0] * g0 + raw[1] * g1 + raw[2] * g2;
  float* o = V.pix_grad + (size_t)j * 12;
  *reinterpret_cast<float4*>(o) = make_float4(g0, g1, g2, 0.f);
  *reinterpret_cast<float4*>(o + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  *reinterpret_cast<float4*>(o + 8) = make_float4(G, 0.f, 0.f, 0.f);
}

// ---- chain surfel gradients to (t, R, scale, latent_unit) block partials --------------------
__device__ __forceinline__ float block_sum(float v, float* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0)
    for (int w = 0; w < LB / 32; ++w) t += s_red[w];
  return t;   // valid on thread 0
}

__device__ void update_detection(const EngineDev& E, int b);

// The block that publishes a detection's last partial also runs its update (gradient assembly, Adam + SGD, the
// next iteration's pose): chain and update are one launch.
__global__ void __launch_bounds__(LB) chain_update_kernel(EngineDev E) {
  __shared__ float s_red[LB / 32];
  __shared__ float s_l3scale, s_loss3d, s_n3;
  __shared__ int s_front_count, s_last;
  const int b = blockIdx.y;
  DetState& D = E.det[b];
  const SplatView& V = E.views[b];
  const int m = min(E.surf_count[b], (int)E.cap);
  if (blockIdx.x > 0 && (int)(blockIdx.x * LB) >= m) return;   // block 0 always publishes the loss
  if (threadIdx.x < 32) {     // warp 0: fixed-order reduction of the 3D-loss block partials
    const int nb = (m + LB - 1) / LB;
    double s = 0.0, c = 0.0, f = 0.0;
    const double* p = E.part3 + (size_t)b * E.nb3 * 3;
    for (int k = threadIdx.x; k < nb; k += 32) { s += p[k * 3]; c += p[k * 3 + 1]; f += p[k * 3 + 2]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
      f += __shfl_xor_sync(0xffffffffu, f, o);
    }
    if (threadIdx.x == 0) {
      s_l3scale = c > 0.0 ? E.w3d / (float)c : 0.f;
      s_loss3d = c > 0.0 ? (float)(s / c) : 0.f;                      // optimizer.py:192-197
      s_n3 = (float)c;
      s_front_count = (int)f;
    }
  }
  __syncthreads();
  const int i = blockIdx.x * LB + threadIdx.x;
  float dt[3] = {0.f, 0.f, 0.f}, dR[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, dscale = 0.f, df = 0.f;
  if (i < m) {
    const float* P = D.pose;
    float dv[3] = {V.d_v[i * 3], V.d_v[i * 3 + 1], V.d_v[i * 3 + 2]};
    const float dm[3] = {V.d_m[i * 3], V.d_m[i * 3 + 1], V.d_m[i * 3 + 2]};
    const float dc[3] = {V.d_c[i * 3], V.d_c[i * 3 + 1], V.d_c[i * 3 + 2]};
    const float4 r3 = *reinterpret_cast<const float4*>(E.l3rec + ((size_t)b * E.cap + i) * 4);
    if (r3.w >= 0.f) {           // d loss3d / d v = -u/n ; d / d scale = -(u . L_nn)/(n scale)
      const float k = s_l3scale;
      dv[0] -= k * r3.x; dv[1] -= k * r3.y; dv[2] -= k * r3.z;
      dscale = -k * E.l3dot[(size_t)b * E.cap + i] / D.scale;
    }
    const float* p = E.surf_pts + ((size_t)b * E.cap + i) * 3;
    const float* n = E.surf_nrm + ((size_t)b * E.cap + i) * 3;
    for (int a = 0; a < 3; ++a) {
      dt[a] = dv[a];
      for (int c = 0; c < 3; ++c) dR[a * 3 + c] = dv[a] * p[c] + dm[a] * n[c];
    }
    // d p = R^T d v + d colour path: C = ((-p_x, p_y, p_z) + 1)/2 (projection.py:53-55, rasterer.py:114)
    float dp[3];
    for (int c = 0; c < 3; ++c) dp[c] = P[0 * 4 + c] * dv[0] + P[1 * 4 + c] * dv[1] + P[2 * 4 + c] * dv[2];
    dp[0] += -0.5f * dc[0]; dp[1] += 0.5f * dc[1]; dp[2] += 0.5f * dc[2];
    // p = g - f n_hat with n_hat constant (grid.py:57-61): d f = - n_hat . d p
    df = -(n[0] * dp[0] + n[1] * dp[1] + n[2] * dp[2]);
  }
  float* out = E.partc + ((size_t)b * E.nbc + blockIdx.x) * (16 + E.L);
  float t;
  for (int a = 0; a < 3; ++a) { t = block_sum(dt[a], s_red); if (threadIdx.x == 0) out[a] = t; }
  for (int a = 0; a < 9; ++a) { t = block_sum(dR[a], s_red); if (threadIdx.x == 0) out[3 + a] = t; }
  t = block_sum(dscale, s_red); if (threadIdx.x == 0) out[12] = t;
  for (int k = 0; k < E.L; ++k) {
    const float g = i < m ? df * E.surf_glat[((size_t)b * E.cap + i) * E.L + k] : 0.f;
    t = block_sum(g, s_red);
    if (threadIdx.x == 0) out[16 + k] = t;
  }
  // ---- last block of the detection: update ----
  if (threadIdx.x == 0) {
    __threadfence();                                               // this block's partials before its ticket
    const int nblk = max(1, (m + LB - 1) / LB);
    s_last = atomicAdd(&E.chain_done[b], 1) == nblk - 1;
    if (s_last) {
      E.chain_done[b] = 0;
      __threadfence();
      // every block reduced the same 3D-loss partials in the same order: this block's copy is THE value
      D.loss3d = s_loss3d; D.n3 = s_n3; D.front_count = s_front_count;
    }
  }
  __syncthreads();
  if (s_last) update_detection(E, b);
}

// ---- gradient assembly, skip logic, Adam + SGD (optimizer.py:34-38,49-52,146-157) ----------
__device__ void update_detection(const EngineDev& E, const int b) {
  __shared__ float s_tot[16 + kMaxLatent];
  DetState& D = E.det[b];
  const int m = min(E.surf_count[b], (int)E.cap);
  const int nb = max(1, (m + LB - 1) / LB);                          // block 0 always writes a partial
  const int ncomp = 16 + E.L;
  for (int c = threadIdx.x; c < ncomp; c += LB) {
    float s = 0.f;
    const float* p = E.partc + (size_t)b * E.nbc * ncomp + c;
    for (int k = 0; k < nb; ++k) s += __ldcg(p + (size_t)k * ncomp);  // other blocks' partials: not through L1
    s_tot[c] = s;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int L = E.L;
  float* g = E.grads + (size_t)b * (5 + 2 * L);
  const float cy = cosf(D.yaw), sy = sinf(D.yaw);
  const float* dR = s_tot + 3;
  // R = diag(1,-1,1) R_y(yaw): dR/dyaw = [[-s,0,c],[0,0,0],[-c,0,-s]]
  const float dyaw = dR[0] * (-sy) + dR[2] * cy + dR[6] * (-cy) + dR[8] * (-sy);
  const float dtr[3] = {s_tot[0], s_tot[1], s_tot[2]};
  const float dscale = s_tot[12];
  // normalize backward: l_hat = l/||l||  ->  dl = (dl_hat - l_hat (l_hat . dl_hat)) / ||l||
  float n2 = 0.f, dot = 0.f;
  for (int k = 0; k < L; ++k) {
    n2 += E.latent[b * L + k] * E.latent[b * L + k];
    dot += E.latent_unit[b * L + k] * s_tot[16 + k];
  }
  const float nrm = fmaxf(sqrtf(n2), 1e-12f);
  g[0] = dyaw; g[1] = dtr[0]; g[2] = dtr[1]; g[3] = dtr[2]; g[4] = dscale;
  for (int k = 0; k < L; ++k) {
    g[5 + k] = s_tot[16 + k];
    g[5 + L + k] = (s_tot[16 + k] - E.latent_unit[b * L + k] * dot) / nrm;
  }
  // skip logic
  const float loss = E.w3d * D.loss3d + E.w2d * D.loss2d;
  int skip = 0;
  if (D.front_count == 0 || D.n_lidar == 0) skip = 1;               // optimizer.py:127-129
  else if (isnan(loss) || loss == 0.f) skip = 2;                    // optimizer.py:149-151
  D.skip_reason = skip;
  if (D.iter < E.max_iters) {
    float* h = E.history + ((size_t)b * E.max_iters + D.iter) * 4;
    h[0] = D.loss2d; h[1] = D.loss3d; h[2] = loss; h[3] = (float)skip;
  }
  D.iter += 1;
  if (skip) return;                                                  // parameters, pose and unit latent stay as they are
  // Adam(lr 0.01, betas (0.9, 0.999), eps 1e-8) on yaw, trans
  D.adam_t += 1;
  const double b1 = 0.9, b2 = 0.999, lr = 0.01, eps = 1e-8;
  const double bc1 = 1.0 - pow(b1, (double)D.adam_t), bc2 = 1.0 - pow(b2, (double)D.adam_t);
  const float step = (float)(lr / bc1);
  const float bc2s = (float)sqrt(bc2);
  float grad4[4] = {dyaw, dtr[0], dtr[1], dtr[2]};
  float* par[4] = {&D.yaw, &D.trans[0], &D.trans[1], &D.trans[2]};
  for (int k = 0; k < 4; ++k) {
    D.adam_m[k] = D.adam_m[k] * (float)b1 + (1.f - (float)b1) * grad4[k];
    D.adam_v[k] = D.adam_v[k] * (float)b2 + (1.f - (float)b2) * grad4[k] * grad4[k];
    const float denom = sqrtf(D.adam_v[k]) / bc2s + (float)eps;
    *par[k] = *par[k] - step * (D.adam_m[k] / denom);
  }
  // SGD(momentum 0): scale lr 0.01, latent lr 3e-5
  D.scale = D.scale - 0.01f * dscale;
  for (int k = 0; k < L; ++k) E.latent[b * L + k] = E.latent[b * L + k] - 0.00003f * g[5 + L + k];
  begin_iteration(E, b);                                             // pose and unit latent of the next iteration
}

// ---- parameters straight from / to the caller's device tensors (optimizer.py:26-30) ---------------
__global__ void import_params_kernel(DetState* __restrict__ det, float* __restrict__ latent, int L,
                                     const float* __restrict__ yaw, const float* __restrict__ trans,
                                     const float* __restrict__ scale, const float* __restrict__ latent_in) {
  const int t = threadIdx.x;
  if (t == 0 && yaw) det->yaw = yaw[0];
  if (t < 3 && trans) det->trans[t] = trans[t];
  if (t == 0 && scale) det->scale = scale[0];
  if (latent_in) for (int i = t; i < L; i += blockDim.x) latent[i] = latent_in[i];
}

__global__ void export_params_kernel(const DetState* __restrict__ det, const float* __restrict__ latent, int L,
                                     float* __restrict__ yaw, float* __restrict__ trans, float* __restrict__ scale,
                                     float* __restrict__ latent_out) {
  const int t = threadIdx.x;
  if (t == 0 && yaw) yaw[0] = det->yaw;
  if (t < 3 && trans) trans[t] = det->trans[t];
  if (t == 0 && scale) scale[0] = det->scale;
  if (latent_out) for (int i = t; i < L; i += blockDim.x) latent_out[i] = latent[i];
}

// ---- Optimizer.optimize as one call (sdfr_refine_optimize): packed staging in, packed read-back out ----
// Prologue: block 0 unpacks the staged [DetState | SplatView | lidar] block of slot b (ONE host->device copy) and takes
// the parameters from the caller's device tensors (import_params); every block resizes its pixels of the NOCS target
// (optimizer.py:135-137).
__global__ void __launch_bounds__(256) optimize_prologue_kernel(EngineDev E, int b, const unsigned int* __restrict__ stage,
                                                                const float* __restrict__ nocs_src, int th, int tw,
                                                                const float* __restrict__ yaw, const float* __restrict__ trans,
                                                                const float* __restrict__ scale,
                                                                const float* __restrict__ latent_in, int n_lidar, int H, int W) {
  const int t = threadIdx.x;
  if (blockIdx.x == 0) {
    constexpr int DW = (int)(sizeof(DetState) / 4), VW = (int)(sizeof(SplatView) / 4);
    unsigned int* det = reinterpret_cast<unsigned int*>(E.det + b);
    unsigned int* view = reinterpret_cast<unsigned int*>(E.views + b);
    for (int i = t; i < DW; i += 256) det[i] = stage[i];
    for (int i = t; i < VW; i += 256) view[i] = stage[DW + i];
    float* lidar = E.lidar + (size_t)b * E.max_lidar * 3;
    const float* ls = reinterpret_cast<const float*>(stage + DW + VW);
    for (int i = t; i < n_lidar * 3; i += 256) lidar[i] = ls[i];
    for (int i = t; i < E.L; i += 256) E.latent[(size_t)b * E.L + i] = latent_in[i];
    __syncthreads();
    DetState& D = E.det[b];
    if (t == 0) { D.yaw = yaw[0]; D.scale = scale[0]; }
    if (t < 3) D.trans[t] = trans[t];
  }
  const int j = blockIdx.x * 256 + t;
  if (j >= H * W) return;
  const int h = j / W, w = j - h * W;
  const float sh = (float)th / (float)H, sw = (float)tw / (float)W;
  const int ih = min((int)floorf((float)h * sh), th - 1), iw = min((int)floorf((float)w * sw), tw - 1);
  float* dst = E.target + (size_t)b * 3 * E.max_pixels;
  for (int c = 0; c < 3; ++c) dst[c * (H * W) + j] = nocs_src[(c * th + ih) * tw + iw];
}

// Epilogue: the params tensors updated in place (export_params) and everything the host reads packed into one block
// [DetState | latent | presel_err, overflow | history rows] for ONE device->host copy.
__global__ void __launch_bounds__(256) optimize_epilogue_kernel(EngineDev E, int b, float* __restrict__ yaw,
                                                                float* __restrict__ trans, float* __restrict__ scale,
                                                                float* __restrict__ latent_out,
                                                                unsigned int* __restrict__ rb, const int* __restrict__ overflow,
                                                                int hist_rows) {
  const int t = threadIdx.x;
  constexpr int DW = (int)(sizeof(DetState) / 4);
  const DetState& D = E.det[b];
  if (t == 0) { yaw[0] = D.yaw; scale[0] = D.scale; }
  if (t < 3) trans[t] = D.trans[t];
  const float* lat = E.latent + (size_t)b * E.L;
  for (int i = t; i < E.L; i += 256) latent_out[i] = lat[i];
  const unsigned int* det = reinterpret_cast<const unsigned int*>(E.det + b);
  for (int i = t; i < DW; i += 256) rb[i] = det[i];
  for (int i = t; i < E.L; i += 256) rb[DW + i] = __float_as_uint(lat[i]);
  if (t == 0) {
    rb[DW + E.L] = (unsigned int)*E.presel_err;
    rb[DW + E.L + 1] = overflow ? (unsigned int)*overflow : 0u;
  }
  const unsigned int* hist = reinterpret_cast<const unsigned int*>(E.history + (size_t)b * E.max_iters * 4);
  for (int i = t; i < hist_rows * 4; i += 256) rb[DW + E.L + 2 + i] = hist[i];
}

// ---- dump-time label extents (utils/refinement.py:527-541) ------------------------------------------
// get_kitti_label evaluates the decoder with the refined latent AS IS (not normalised, refine_css.py:229)
__global__ void raw_latent_kernel(EngineDev E) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < E.batch * E.L) E.latent_unit[i] = E.latent[i];
}

// min / max of the band's isosurface points of one detection (exact: min and max commute with the
// reference's later multiplication by the positive scale)
__global__ void __launch_bounds__(LB) extent_kernel(EngineDev E) {
  __shared__ float s_lo[3][LB / 32], s_hi[3][LB / 32];
  __shared__ int s_n[LB / 32];
  const int b = blockIdx.x;
  const int m = min(E.surf_count[b], (int)E.cap);
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  int n = 0;
  for (int i = threadIdx.x; i < m; i += LB) {
    if (!E.surf_valid[(size_t)b * E.cap + i]) continue;
    const float* p = E.surf_pts + ((size_t)b * E.cap + i) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) { lo[c] = fminf(lo[c], p[c]); hi[c] = fmaxf(hi[c], p[c]); }
    ++n;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
      hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
    }
    n += __shfl_xor_sync(0xffffffffu, n, o);
  }
  if ((threadIdx.x & 31) == 0) {
    for (int c = 0; c < 3; ++c) { s_lo[c][threadIdx.x >> 5] = lo[c]; s_hi[c][threadIdx.x >> 5] = hi[c]; }
    s_n[threadIdx.x >> 5] = n;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int tn = 0;
    for (int w = 0; w < LB / 32; ++w) {
      for (int c = 0; c < 3; ++c) { lo[c] = fminf(lo[c], s_lo[c][w]); hi[c] = fmaxf(hi[c], s_hi[c][w]); }
      tn += s_n[w];
    }
    float* o = E.extents + (size_t)b * 8;
    for (int c = 0; c < 3; ++c) { o[c] = lo[c]; o[3 + c] = hi[c]; }
    o[6] = (float)tn; o[7] = (float)m;
  }
}

}  // namespace

}  // namespace sdfr

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
using namespace sdfr;

struct sdfr_refine {
  sdfr_decoder* dec;
  sdfr_refine_cfg cfg;
  EngineDev E;
  std::vector<void*> allocs;
  std::vector<SplatView> views_host;
  std::vector<float*> nocs_dev;       // per-detection staging of the un-resized NOCS prediction
  std::vector<size_t> nocs_cap;
  std::vector<int> det_w, det_h;      // crop of each slot (0 = never set)
  int active;                         // detections the next run covers: slots [0, active)
  // One refine iteration captured as a CUDA graph (all arguments are device-resident state, so the same
  // executable graph is replayed every iteration).  The launch shapes depend on the number of active
  // detections and on the largest active crop, rounded up to a power of two: one graph per such key is kept, so
  // a sequence of frames with 1-8 detections and ragged crops captures a handful of graphs once.
  std::map<std::tuple<int, int, int>, std::pair<cudaGraphExec_t, int>> graphs;
  cudaStream_t capture_stream;   // the caller's stream may be the legacy default stream, which cannot be captured
  int runs;
  std::vector<int> iters_enqueued;   // per detection: iterations launched since its last set_detection
  std::vector<DetState> host_state;  // what the last sdfr_refine_get / get_batch read back
  // Page-locked landing zone of the read-back: [DetState x batch | latent x batch*L | presel, overflow | history].
  // A device->host copy into pageable memory blocks the host until it has completed, so five small copies were
  // five round trips; into this block they are enqueued back to back and waited for once.
  char* rb;
  size_t rb_det, rb_lat, rb_flags, rb_hist, rb_bytes;
  // sdfr_refine_optimize: page-locked [DetState | SplatView | lidar] staging and its device copy; packed read-back
  // [DetState | latent | 2 flags | history] on the device and its page-locked landing zone
  char* opt_stage_host;
  char* opt_stage_dev;
  char* opt_rb_host;
  char* opt_rb_dev;
  size_t opt_stage_bytes, opt_rb_bytes;
};

namespace {

template <typename T>
int dev_alloc(sdfr_refine* r, T** p, size_t count) {
  void* q = nullptr;
  SDFR_CUDA(cudaMalloc(&q, count * sizeof(T) + 16));
  SDFR_CUDA(cudaMemset(q, 0, count * sizeof(T) + 16));
  r->allocs.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return SDFR_OK;
}

void invert3x3(const float* k, float* o) {
  const double a = k[0], b = k[1], c = k[2], d = k[3], e = k[4], f = k[5], g = k[6], h = k[7], i = k[8];
  const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
  o[0] = (float)((e * i - f * h) / det); o[1] = (float)((c * h - b * i) / det); o[2] = (float)((b * f - c * e) / det);
  o[3] = (float)((f * g - d * i) / det); o[4] = (float)((a * i - c * g) / det); o[5] = (float)((c * d - a * f) / det);
  o[6] = (float)((d * h - e * g) / det); o[7] = (float)((b * g - a * h) / det); o[8] = (float)((a * e - b * d) / det);
}

int pow2_ceil(int v, int lo) {
  int q = lo;
  while (q < v) q <<= 1;
  return q;
}

// launch shape of the per-pixel kernels: the largest crop among the active detections, rounded up to a power
// of two (>= 32) and clamped to the capacity; kernels exit early outside each detection's own crop
void active_shape(const sdfr_refine* r, int* qw, int* qh, int* any_fine) {
  int mw = 0, mh = 0;
  *any_fine = 0;
  for (int b = 0; b < r->active; ++b) {
    mw = std::max(mw, r->det_w[b]); mh = std::max(mh, r->det_h[b]);
    *any_fine |= r->det_w[b] * r->det_h[b] <= kFineCropPixelsHost;     // picks the splat tiling of that detection
  }
  *qw = std::min(pow2_ceil(mw, 32), std::max(r->cfg.max_width, mw));
  *qh = std::min(pow2_ceil(mh, 32), std::max(r->cfg.max_height, mh));
}

}  // namespace

extern "C" int sdfr_refine_create(sdfr_decoder* dec, const sdfr_refine_cfg* cfg, sdfr_refine** out) {
  SDFR_REQUIRE(dec && cfg && out, SDFR_E_INVALID, "null argument");
  SDFR_REQUIRE(cfg->batch > 0 && cfg->density > 1 && cfg->max_width > 0 && cfg->max_height > 0, SDFR_E_INVALID,
               "bad refine configuration");
  SDFR_REQUIRE(dec->dev.latent_size <= kMaxLatent, SDFR_E_UNSUPPORTED, "latent size %d > %d", dec->dev.latent_size,
               kMaxLatent);
  sdfr_refine* r = new sdfr_refine();
  r->dec = dec;
  r->cfg = *cfg;
  r->capture_stream = nullptr; r->runs = 0;
  r->rb = nullptr;
  r->opt_stage_host = r->opt_rb_host = r->opt_stage_dev = r->opt_rb_dev = nullptr;
  r->active = cfg->batch;
  r->iters_enqueued.assign((size_t)cfg->batch, 0);
  r->det_w.assign((size_t)cfg->batch, 0);
  r->det_h.assign((size_t)cfg->batch, 0);
  r->host_state.resize((size_t)cfg->batch);
  memset(r->host_state.data(), 0, sizeof(DetState) * (size_t)cfg->batch);
  EngineDev& E = r->E;
  const int B = cfg->batch, L = dec->dev.latent_size;
  E.batch = B; E.L = L; E.in0 = L + 3;
  E.ng = (long long)cfg->density * cfg->density * cfg->density;
  E.cap = E.ng;
  E.max_pixels = cfg->max_width * cfg->max_height;
  E.max_lidar = cfg->max_lidar > 0 ? cfg->max_lidar : 1;
  E.max_iters = cfg->max_iters > 0 ? cfg->max_iters : 1;
  E.w2d = cfg->weight_2d; E.w3d = cfg->weight_3d;
  E.nb2 = (E.max_pixels + LB - 1) / LB;
  E.nb3 = (int)((E.cap + LB - 1) / LB);
  E.nbc = E.nb3;
  int rc = 0;
#define A(ptr, count) if ((rc = dev_alloc(r, &(ptr), (size_t)(count)))) { sdfr_refine_destroy(r); return rc; }
  A(E.det, B); A(E.latent, B * L); A(E.latent_unit, B * L);
  A(E.sdf, B * E.ng); A(E.dinput, B * E.ng * E.in0);
  A(E.surf_pts, B * E.cap * 3); A(E.surf_nrm, B * E.cap * 3); A(E.surf_glat, B * E.cap * L);
  A(E.surf_idx, B * E.cap); A(E.surf_count, B);
  A(E.band_status, (size_t)B * (E.ng / 1024 + 2)); A(E.band_ctrl, 4); A(E.chain_done, B);
  E.lip = cfg->latent_lipschitz > 0.f ? cfg->latent_lipschitz : 0.f;
  {
    int impl = cfg->mlp_impl;
    if (impl == SDFR_MLP_AUTO) impl = dec->tc.ok ? SDFR_MLP_TCGEN05 : SDFR_MLP_FFMA;
    E.pre_thr = 0.03f + (impl == SDFR_MLP_TCGEN05 ? kPreselectMargin : 0.f);
  }
  if (E.lip > 0.f) {
    A(E.sdf_ref, B * E.ng); A(E.z_ref, B * L); A(E.ref_valid, B); A(E.cand_thr, B); A(E.cand_all, B);
    A(E.cand_src, (size_t)B * E.ng); A(E.cand_sdf, (size_t)B * E.ng); A(E.cand_start, B); A(E.cand_count, B);
    A(E.cand_total, 4); A(E.cand_status, (size_t)B * (E.ng / 1024 + 2)); A(E.cand_ctrl, 4); A(E.lattice_rows, 2);
  }
  A(E.band_det_start, B); A(E.band_total, 4); A(E.band_src, (size_t)B * E.ng); A(E.band_sdf, (size_t)B * E.ng);
  A(E.surf_valid, (size_t)B * E.cap);
  A(E.target, (size_t)B * 3 * E.max_pixels); A(E.lidar, (size_t)B * E.max_lidar * 3);
  A(E.l3rec, B * E.cap * 4); A(E.l3dot, B * E.cap); A(E.l2rec, (size_t)B * E.max_pixels * 4);
  A(E.part2, (size_t)B * E.nb2 * 3); A(E.part3, (size_t)B * E.nb3 * 3); A(E.partc, (size_t)B * E.nbc * (16 + L));
  A(E.history, (size_t)B * E.max_iters * 4); A(E.grads, (size_t)B * (5 + 2 * L));
  A(E.presel_err, 4); A(E.extents, (size_t)B * 8);
  A(E.mask_scratch, mlp_tc_mask_scratch_bytes(dec) / 8 + 1);
  A(E.views, B);
  r->rb_det = 0;
  r->rb_lat = r->rb_det + sizeof(DetState) * (size_t)B;
  r->rb_flags = r->rb_lat + sizeof(float) * (size_t)B * L;
  r->rb_hist = r->rb_flags + 4 * sizeof(int);
  r->rb_bytes = r->rb_hist + sizeof(float) * (size_t)B * E.max_iters * 4;
  if (cudaHostAlloc(reinterpret_cast<void**>(&r->rb), r->rb_bytes, cudaHostAllocDefault) != cudaSuccess) {
    r->rb = nullptr;
    sdfr_refine_destroy(r);
    SDFR_REQUIRE(false, SDFR_E_CUDA, "cudaHostAlloc of the %zu-byte read-back block failed", r->rb_bytes);
  }
  memset(r->rb, 0, r->rb_bytes);
  r->opt_stage_bytes = sizeof(DetState) + sizeof(SplatView) + sizeof(float) * 3 * (size_t)E.max_lidar;
  r->opt_rb_bytes = sizeof(DetState) + sizeof(float) * (size_t)(L + 2) + sizeof(float) * 4 * (size_t)E.max_iters;
  A(r->opt_stage_dev, r->opt_stage_bytes); A(r->opt_rb_dev, r->opt_rb_bytes);
  if (cudaHostAlloc(reinterpret_cast<void**>(&r->opt_stage_host), r->opt_stage_bytes, cudaHostAllocDefault) != cudaSuccess ||
      cudaHostAlloc(reinterpret_cast<void**>(&r->opt_rb_host), r->opt_rb_bytes, cudaHostAllocDefault) != cudaSuccess) {
    sdfr_refine_destroy(r);
    SDFR_REQUIRE(false, SDFR_E_CUDA, "cudaHostAlloc of the optimize staging blocks failed");
  }
  r->views_host.resize(B);
  r->nocs_dev.assign(B, nullptr);
  r->nocs_cap.assign(B, 0);
  for (int b = 0; b < B; ++b) {
    SplatView& V = r->views_host[b];
    memset(&V, 0, sizeof(V));
    V.rot = SDFR_ROT_DCM;
    V.output_nocs = 1;
    V.coords = E.surf_pts + (size_t)b * E.cap * 3;
    V.normals = E.surf_nrm + (size_t)b * E.cap * 3;
    V.colors = nullptr;
    V.pose = reinterpret_cast<const float*>(reinterpret_cast<const char*>(E.det + b) + offsetof(DetState, pose));
    V.count = E.surf_count + b;
    V.valid = E.surf_valid + (size_t)b * E.cap;
    V.capacity = (int)E.cap;
    A(V.cam_v, E.cap * 3); A(V.cam_m, E.cap * 3); A(V.cam_c, E.cap * 3); A(V.plane_a, E.cap);
    A(V.bbox, E.cap * 4); A(V.front, E.cap);
    V.cam_rgb = nullptr;
    A(V.color, (size_t)3 * E.max_pixels); A(V.mask, E.max_pixels); A(V.depth, E.max_pixels);
    A(V.nmap, (size_t)3 * E.max_pixels);
    A(V.pix_stat, (size_t)4 * E.max_pixels); A(V.pix_raw, (size_t)8 * E.max_pixels);
    A(V.pix_grad, (size_t)12 * E.max_pixels);
    A(V.d_v, E.cap * 3); A(V.d_m, E.cap * 3); A(V.d_c, E.cap * 3);
  }
#undef A
  *out = r;
  return SDFR_OK;
}

extern "C" void sdfr_refine_destroy(sdfr_refine* r) {
  if (!r) return;
  for (auto& g : r->graphs) if (g.second.first) cudaGraphExecDestroy(g.second.first);
  if (r->capture_stream) cudaStreamDestroy(r->capture_stream);
  for (void* p : r->allocs) cudaFree(p);
  for (float* p : r->nocs_dev) if (p) cudaFree(p);
  if (r->rb) cudaFreeHost(r->rb);
  if (r->opt_stage_host) cudaFreeHost(r->opt_stage_host);
  if (r->opt_rb_host) cudaFreeHost(r->opt_rb_host);
  delete r;
}

extern "C" int sdfr_refine_set_active(sdfr_refine* r, int count) {
  SDFR_REQUIRE(r && count >= 1 && count <= r->cfg.batch, SDFR_E_INVALID, "active count %d outside [1, %d]", count,
               r ? r->cfg.batch : 0);
  r->active = count;
  return SDFR_OK;
}

extern "C" int sdfr_refine_set_detection(sdfr_refine* r, int b, const float* k_host, const float* kinv_host,
                                         int width, int height, const float* nocs_host, int th, int tw,
                                         const float* lidar_host, int n_lidar, const float* yaw_host,
                                         const float* trans_host, const float* scale_host,
                                         const float* latent_host, void* stream) {
  SDFR_REQUIRE(r && k_host && nocs_host, SDFR_E_INVALID, "null argument");
  SDFR_REQUIRE(b >= 0 && b < r->cfg.batch, SDFR_E_INVALID, "detection index %d out of range", b);
  SDFR_REQUIRE(width > 0 && height > 0 && width <= r->cfg.max_width && height <= r->cfg.max_height, SDFR_E_CAPACITY,
               "crop %dx%d exceeds the configured capacity %dx%d", width, height, r->cfg.max_width, r->cfg.max_height);
  SDFR_REQUIRE(n_lidar >= 0 && n_lidar <= r->E.max_lidar, SDFR_E_CAPACITY, "%d lidar points exceed capacity %d",
               n_lidar, r->E.max_lidar);
  SDFR_REQUIRE(n_lidar == 0 || lidar_host, SDFR_E_INVALID, "null lidar array");
  r->iters_enqueued[b] = 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  EngineDev& E = r->E;
  DetState D;
  memset(&D, 0, sizeof(D));     // fresh optimiser state: a new Optimizer (optimizer.py:46-52)
  D.width = width; D.height = height; D.n_lidar = n_lidar; D.target_ready = 1;
  if (yaw_host) D.yaw = yaw_host[0];
  if (trans_host) { D.trans[0] = trans_host[0]; D.trans[1] = trans_host[1]; D.trans[2] = trans_host[2]; }
  if (scale_host) D.scale = scale_host[0];
  SDFR_CUDA(cudaMemcpyAsync(E.det + b, &D, sizeof(D), cudaMemcpyHostToDevice, s));
  if (latent_host)
    SDFR_CUDA(cudaMemcpyAsync(E.latent + (size_t)b * E.L, latent_host, E.L * sizeof(float), cudaMemcpyHostToDevice, s));
  if (n_lidar > 0)
    SDFR_CUDA(cudaMemcpyAsync(E.lidar + (size_t)b * E.max_lidar * 3, lidar_host, (size_t)n_lidar * 3 * sizeof(float),
                              cudaMemcpyHostToDevice, s));
  const size_t nocs_count = (size_t)3 * th * tw;
  if (r->nocs_cap[b] < nocs_count) {
    if (r->nocs_dev[b]) SDFR_CUDA(cudaFree(r->nocs_dev[b]));
    r->nocs_dev[b] = nullptr;
    r->nocs_cap[b] = 0;
    SDFR_CUDA(cudaMalloc(&r->nocs_dev[b], nocs_count * sizeof(float)));
    r->nocs_cap[b] = nocs_count;
  }
  SDFR_CUDA(cudaMemcpyAsync(r->nocs_dev[b], nocs_host, nocs_count * sizeof(float), cudaMemcpyHostToDevice, s));
  const int P = width * height;
  resize_target_kernel<<<(P + 255) / 256, 256, 0, s>>>(r->nocs_dev[b], th, tw, E.target + (size_t)b * 3 * E.max_pixels,
                                                        height, width, P);
  SDFR_LAUNCH_CHECK();
  SplatView& V = r->views_host[b];
  V.width = width; V.height = height;
  for (int i = 0; i < 9; ++i) V.k[i] = k_host[i];
  if (kinv_host) for (int i = 0; i < 9; ++i) V.kinv[i] = kinv_host[i];
  else invert3x3(k_host, V.kinv);
  SDFR_CUDA(cudaMemcpyAsync(E.views + b, &V, sizeof(V), cudaMemcpyHostToDevice, s));
  r->det_w[b] = width;
  r->det_h[b] = height;
  return SDFR_OK;
}

extern "C" int sdfr_refine_import(sdfr_refine* r, int b, const float* yaw_dev, const float* trans_dev,
                                  const float* scale_dev, const float* latent_dev, void* stream) {
  SDFR_REQUIRE(r && b >= 0 && b < r->cfg.batch, SDFR_E_INVALID, "bad argument");
  EngineDev& E = r->E;
  import_params_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(E.det + b, E.latent + (size_t)b * E.L, E.L,
                                                                             yaw_dev, trans_dev, scale_dev, latent_dev);
  SDFR_LAUNCH_CHECK();
  return SDFR_OK;
}

extern "C" int sdfr_refine_set_optimizer_state(sdfr_refine* r, int b, const float* adam_m_host,
                                               const float* adam_v_host, int adam_t, void* stream) {
  SDFR_REQUIRE(r && adam_m_host && adam_v_host && b >= 0 && b < r->cfg.batch && adam_t >= 0, SDFR_E_INVALID,
               "bad argument");
  struct { float m[4], v[4]; int t; } st;
  static_assert(offsetof(DetState, adam_v) == offsetof(DetState, adam_m) + 16 &&
                offsetof(DetState, adam_t) == offsetof(DetState, adam_m) + 32, "Adam fields must be contiguous");
  memcpy(st.m, adam_m_host, sizeof(st.m));
  memcpy(st.v, adam_v_host, sizeof(st.v));
  st.t = adam_t;
  SDFR_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(r->E.det + b) + offsetof(DetState, adam_m), &st, 36,
                            cudaMemcpyHostToDevice, reinterpret_cast<cudaStream_t>(stream)));
  return SDFR_OK;
}

extern "C" int sdfr_refine_get_optimizer_state(sdfr_refine* r, int b, float* adam_m_host, float* adam_v_host,
                                               int* adam_t) {
  SDFR_REQUIRE(r && b >= 0 && b < r->cfg.batch, SDFR_E_INVALID, "bad argument");
  const DetState& D = r->host_state[b];
  if (adam_m_host) memcpy(adam_m_host, D.adam_m, sizeof(D.adam_m));
  if (adam_v_host) memcpy(adam_v_host, D.adam_v, sizeof(D.adam_v));
  if (adam_t) *adam_t = D.adam_t;
  return SDFR_OK;
}

namespace {

struct IterPlan {
  MlpInputs in, in_band, in_cand;
  BandArgs ba;
  SelectArgs sel, sel_cand;   // band selection; candidate selection of the pruned lattice pass
  bool coarse, prune;
  int impl;
};

IterPlan make_iter_plan(sdfr_refine* r, int B) {
  EngineDev& E = r->E;
  IterPlan p;
  MlpInputs& in = p.in;
  in.inputs = nullptr;
  in.latent_unit = E.latent_unit;
  in.lattice = make_lattice(r->cfg.density);
  in.points_per_batch = E.ng;
  in.n = E.ng * B;
  in.index = nullptr;
  in.count_dev = nullptr;
  in.small_tiles = 0;
  p.in_band = in;                       // same lattice / latents, rows gathered through band_src
  p.in_band.index = E.band_src;
  p.in_band.count_dev = E.band_total;
  p.in_band.small_tiles = B <= 4;       // a few detections: ~2 000 rows each, spread them over all SMs
  p.in_band.mask_scratch = E.mask_scratch;
  int impl = r->cfg.mlp_impl;
  if (impl == SDFR_MLP_AUTO) impl = r->dec->tc.ok ? SDFR_MLP_TCGEN05 : SDFR_MLP_FFMA;
  p.impl = impl;
  // Pre-selection threshold = band (grid.py:43) + margin.  With the tensor-core decoder the lattice pass runs
  // at fp16 operand precision (error ~3e-4); the accurate second pass then decides the real band, so the result
  // is the same set the accurate kernel alone gives as long as the coarse error stays inside the margin.  The
  // error is MEASURED on every pre-selected row (band_surface_kernel) and checked by sdfr_refine_get.
  p.coarse = impl == SDFR_MLP_TCGEN05;
  BandArgs& ba = p.ba;
  ba.lattice = in.lattice; ba.sdf = E.sdf; ba.n = E.ng; ba.batch = B;
  ba.threshold = E.pre_thr;
  ba.out_valid = E.surf_valid; ba.final_threshold = 0.03f;
  ba.det_start = E.band_det_start;
  ba.det_count = E.surf_count; ba.total = E.band_total; ba.band_src = E.band_src;
  ba.band_sdf = E.band_sdf; ba.band_dinput = E.dinput; ba.in0 = E.in0; ba.latent = E.L;
  ba.out_pts = E.surf_pts; ba.out_nrm = E.surf_nrm; ba.out_idx = E.surf_idx; ba.out_glat = E.surf_glat; ba.cap = E.cap;
  ba.presel_err = p.coarse ? E.presel_err : nullptr;
  ba.views = E.views;
  // band selection over the whole lattice (no pruning) ...
  SelectArgs& sel = p.sel;
  memset(&sel, 0, sizeof(sel));
  sel.values = E.sdf; sel.n = E.ng; sel.batch = B; sel.threshold = ba.threshold;
  sel.out_src = E.band_src; sel.det_start = E.band_det_start; sel.det_count = E.surf_count; sel.total = E.band_total;
  sel.status = E.band_status; sel.ctrl = E.band_ctrl;
  p.prune = E.lip > 0.f;
  if (p.prune) {
    // ... or: candidates from the reference sdf, lattice pass on them only, band selection among the candidates
    SelectArgs& sc = p.sel_cand;
    memset(&sc, 0, sizeof(sc));
    sc.values = E.sdf_ref; sc.n = E.ng; sc.batch = B; sc.det_threshold = E.cand_thr; sc.det_all = E.cand_all;
    sc.out_src = E.cand_src; sc.det_start = E.cand_start; sc.det_count = E.cand_count; sc.total = E.cand_total;
    sc.status = E.cand_status; sc.ctrl = E.cand_ctrl; sc.total_accum = E.lattice_rows;
    p.in_cand = in;
    p.in_cand.index = E.cand_src;
    p.in_cand.count_dev = E.cand_total;
    sel.values = E.cand_sdf; sel.in_src = E.cand_src; sel.in_start = E.cand_start; sel.in_count = E.cand_count;
    sel.scatter_values = E.sdf;                      // the guard of the isosurface kernel reads the coarse sdf by lattice index
    sel.scatter_ref = E.sdf_ref; sel.scatter_flag = E.cand_all; sel.scatter_done = E.ref_valid;
  }
  return p;
}

// Optional per-stage timing of an un-captured iteration (sdfr_refine_profile): an event after every stage.
struct StageClock {
  std::vector<cudaEvent_t> ev;
  cudaStream_t s;
  void mark() {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    ev.push_back(e);
  }
};
#define STAGE_MARK(clk) do { if (clk) (clk)->mark(); } while (0)

const char* const kStageNames[] = {"lattice_pass", "band_select", "band_pass", "surface_project", "splat_forward",
                                   "losses", "grad_prep", "splat_backward", "chain_update"};
constexpr int kNumStages = sizeof(kStageNames) / sizeof(kStageNames[0]);

// lattice pass -> band select -> accurate pass on the selected rows -> isosurface projection
int enqueue_surface(sdfr_refine* r, const IterPlan& p, cudaStream_t s, StageClock* clk = nullptr) {
  EngineDev& E = r->E;
  int rc;
  if (p.prune) {
    if ((rc = launch_select(p.sel_cand, s))) return rc;
    rc = p.coarse ? launch_mlp_tc_coarse(r->dec, p.in_cand, E.cand_sdf, s) : launch_mlp_ffma(r->dec, p.in_cand, E.cand_sdf, nullptr, s);
  } else {
    rc = p.coarse ? launch_mlp_tc_coarse(r->dec, p.in, E.sdf, s) : launch_mlp_ffma(r->dec, p.in, E.sdf, nullptr, s);
  }
  if (rc) return rc;
  STAGE_MARK(clk);
  if ((rc = launch_select(p.sel, s))) return rc;
  STAGE_MARK(clk);
  rc = p.impl == SDFR_MLP_TCGEN05 ? launch_mlp_tc(r->dec, p.in_band, E.band_sdf, E.dinput, s)
                                  : launch_mlp_ffma(r->dec, p.in_band, E.band_sdf, E.dinput, s);
  if (rc) return rc;
  STAGE_MARK(clk);
  rc = launch_band_surface(p.ba, s);
  STAGE_MARK(clk);
  return rc;
}

// pose and unit latent of detections [0, B) from their current parameters: once per run, the updates keep them current
int enqueue_begin(sdfr_refine* r, int B, cudaStream_t s) {
  EngineDev E = r->E;
  E.batch = B;
  iter_begin_kernel<<<B, 32, 0, s>>>(E);
  SDFR_LAUNCH_CHECK();
  return SDFR_OK;
}

// enqueues the NINE kernels of one refine iteration of detections [0, B) on `s`
int enqueue_iteration(sdfr_refine* r, int B, int qw, int qh, int any_fine, cudaStream_t s, StageClock* clk = nullptr) {
  EngineDev E = r->E;
  E.batch = B;
  const IterPlan p = make_iter_plan(r, B);
  int rc;
  STAGE_MARK(clk);
  if ((rc = enqueue_surface(r, p, s, clk))) return rc;             // lattice, select, band pass, isosurface + projection
  if ((rc = launch_splat_forward(E.views, B, qw, qh, any_fine, s))) return rc;
  STAGE_MARK(clk);
  const int n2 = (qw * qh + LB - 1) / LB;
  losses_batch_kernel<<<dim3(n2 + E.nb3, B), LB, 0, s>>>(E, n2);
  SDFR_LAUNCH_CHECK();
  STAGE_MARK(clk);
  grad_prep_kernel<<<dim3(n2, B), LB, 0, s>>>(E);
  SDFR_LAUNCH_CHECK();
  STAGE_MARK(clk);
  if ((rc = launch_splat_backward(E.views, B, (int)E.cap, s))) return rc;
  STAGE_MARK(clk);
  chain_update_kernel<<<dim3(E.nbc, B), LB, 0, s>>>(E);
  SDFR_LAUNCH_CHECK();
  STAGE_MARK(clk);
  return SDFR_OK;
}

}  // namespace

extern "C" int sdfr_refine_run(sdfr_refine* r, int iters, void* stream) {
  SDFR_REQUIRE(r && iters >= 0, SDFR_E_INVALID, "bad argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int B = r->active;
  for (int b = 0; b < B; ++b)
    SDFR_REQUIRE(r->det_w[b] > 0, SDFR_E_INVALID, "detection %d of the %d active ones has not been set", b, B);
  int qw, qh, any_fine;
  active_shape(r, &qw, &qh, &any_fine);
  for (int b = 0; b < B; ++b) r->iters_enqueued[b] = (int)std::min<long long>((long long)r->iters_enqueued[b] + iters, 1 << 30);
  static int use_graph = -1;
  if (use_graph < 0) { const char* e = getenv("SDFR_REFINE_GRAPH"); use_graph = e ? atoi(e) : 1; }
  int rc;
  int done = 0;
  if (iters > 0 && (rc = enqueue_begin(r, B, s))) return rc;
  // The first call runs un-captured (lazy attribute setup inside the launchers must not happen during capture).
  if (!use_graph || r->runs == 0) {
    for (; done < (use_graph ? std::min(iters, 1) : iters); ++done)
      if ((rc = enqueue_iteration(r, B, qw, qh, any_fine, s))) return rc;
  }
  r->runs += 1;
  if (done == iters) return SDFR_OK;
  const auto key = std::make_tuple(B, qw * 2 + any_fine, qh);
  auto it = r->graphs.find(key);
  if (it == r->graphs.end()) {
    cudaGraph_t graph = nullptr;
    const long long before = sdfr_launch_count();
    if (!r->capture_stream) SDFR_CUDA(cudaStreamCreateWithFlags(&r->capture_stream, cudaStreamNonBlocking));
    SDFR_CUDA(cudaStreamBeginCapture(r->capture_stream, cudaStreamCaptureModeThreadLocal));
    rc = enqueue_iteration(r, B, qw, qh, any_fine, r->capture_stream);
    cudaError_t ce = cudaStreamEndCapture(r->capture_stream, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    SDFR_CUDA(ce);
    const int nodes = (int)(sdfr_launch_count() - before);
    count_launch(-nodes);                     // capturing launched nothing
    cudaGraphExec_t exec = nullptr;
    SDFR_CUDA(cudaGraphInstantiate(&exec, graph, 0));
    cudaGraphDestroy(graph);
    it = r->graphs.emplace(key, std::make_pair(exec, nodes)).first;
  }
  for (; done < iters; ++done) {
    SDFR_CUDA(cudaGraphLaunch(it->second.first, s));
    count_launch(it->second.second);
  }
  return SDFR_OK;
}

namespace {

// Checks riding on the read-back synchronisation: the fp16-range flag of the tensor-core decoder and the
// measured error of its fp16-operand pre-selection pass.
int check_decoder_flags(sdfr_refine* r, int overflow, int presel_bits, cudaStream_t s) {
  float presel = 0.f;
  memcpy(&presel, &presel_bits, sizeof(float));
  if (overflow) {
    tc_overflow_reset(r->dec, s);
    SDFR_REQUIRE(false, SDFR_E_UNSUPPORTED,
                 "an activation left the fp16 range of the split-operand tensor-core kernel; use SDFR_MLP_FFMA for this network");
  }
  if (!(presel <= kPreselectMargin * 0.5f)) {     // also catches a NaN sdf
    cudaMemsetAsync(r->E.presel_err, 0, sizeof(int), s);
    SDFR_REQUIRE(false, SDFR_E_UNSUPPORTED,
                 "the fp16-operand lattice pass is off by %.2e near the band (limit %.2e): band points may have been "
                 "missed by the pre-selection; use SDFR_MLP_FFMA for this network", presel, kPreselectMargin * 0.5f);
  }
  return SDFR_OK;
}

}  // namespace

extern "C" int sdfr_refine_get(sdfr_refine* r, int b, float* params_host, float* history_host, int* n_history,
                               void* stream) {
  SDFR_REQUIRE(r && params_host && b >= 0 && b < r->cfg.batch, SDFR_E_INVALID, "bad argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  EngineDev& E = r->E;
  // everything the caller reads back rides on ONE stream synchronisation: state, latent, the history rows
  // the host knows were enqueued, and the two decoder flags - all into the page-locked block
  DetState* rb_det = reinterpret_cast<DetState*>(r->rb + r->rb_det) + b;
  float* rb_lat = reinterpret_cast<float*>(r->rb + r->rb_lat) + (size_t)b * E.L;
  int* rb_flags = reinterpret_cast<int*>(r->rb + r->rb_flags);
  float* rb_hist = reinterpret_cast<float*>(r->rb + r->rb_hist) + (size_t)b * E.max_iters * 4;
  const int nh_host = std::min(r->iters_enqueued[b], E.max_iters);
  SDFR_CUDA(cudaMemcpyAsync(rb_det, E.det + b, sizeof(DetState), cudaMemcpyDeviceToHost, s));
  SDFR_CUDA(cudaMemcpyAsync(rb_lat, E.latent + (size_t)b * E.L, E.L * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (history_host && nh_host > 0)
    SDFR_CUDA(cudaMemcpyAsync(rb_hist, E.history + (size_t)b * E.max_iters * 4, (size_t)nh_host * 4 * sizeof(float),
                              cudaMemcpyDeviceToHost, s));
  SDFR_CUDA(cudaMemcpyAsync(rb_flags, E.presel_err, sizeof(int), cudaMemcpyDeviceToHost, s));
  int rc = tc_overflow_flag_enqueue(r->dec, rb_flags + 1, s);
  if (rc) return rc;
  SDFR_CUDA(cudaStreamSynchronize(s));
  DetState& D = r->host_state[b];
  D = *rb_det;
  params_host[0] = D.yaw; params_host[1] = D.trans[0]; params_host[2] = D.trans[1]; params_host[3] = D.trans[2];
  params_host[4] = D.scale;
  memcpy(params_host + 5, rb_lat, E.L * sizeof(float));
  const int nh = std::min(D.iter, E.max_iters);
  if (n_history) *n_history = nh;
  if (history_host && nh > nh_host) {      // iterations this handle did not count (not reachable through the ABI)
    SDFR_CUDA(cudaMemcpyAsync(rb_hist, E.history + (size_t)b * E.max_iters * 4, (size_t)nh * 4 * sizeof(float),
                              cudaMemcpyDeviceToHost, s));
    SDFR_CUDA(cudaStreamSynchronize(s));
  }
  if (history_host && nh > 0) memcpy(history_host, rb_hist, (size_t)nh * 4 * sizeof(float));
  return check_decoder_flags(r, rb_flags[1], rb_flags[0], s);
}

extern "C" int sdfr_refine_get_batch(sdfr_refine* r, float* params_host, float* history_host, int* n_history,
                                     void* stream) {
  SDFR_REQUIRE(r && params_host, SDFR_E_INVALID, "bad argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  EngineDev& E = r->E;
  const int B = r->active, L = E.L, W = 5 + L;
  DetState* rb_det = reinterpret_cast<DetState*>(r->rb + r->rb_det);
  float* rb_lat = reinterpret_cast<float*>(r->rb + r->rb_lat);
  int* rb_flags = reinterpret_cast<int*>(r->rb + r->rb_flags);
  float* rb_hist = reinterpret_cast<float*>(r->rb + r->rb_hist);
  SDFR_CUDA(cudaMemcpyAsync(rb_det, E.det, sizeof(DetState) * (size_t)B, cudaMemcpyDeviceToHost, s));
  SDFR_CUDA(cudaMemcpyAsync(rb_lat, E.latent, sizeof(float) * (size_t)B * L, cudaMemcpyDeviceToHost, s));
  if (history_host)
    SDFR_CUDA(cudaMemcpyAsync(rb_hist, E.history, sizeof(float) * (size_t)B * E.max_iters * 4, cudaMemcpyDeviceToHost, s));
  SDFR_CUDA(cudaMemcpyAsync(rb_flags, E.presel_err, sizeof(int), cudaMemcpyDeviceToHost, s));
  int rc = tc_overflow_flag_enqueue(r->dec, rb_flags + 1, s);
  if (rc) return rc;
  SDFR_CUDA(cudaStreamSynchronize(s));
  memcpy(r->host_state.data(), rb_det, sizeof(DetState) * (size_t)B);
  if (history_host) memcpy(history_host, rb_hist, sizeof(float) * (size_t)B * E.max_iters * 4);
  for (int b = 0; b < B; ++b) {
    const DetState& D = r->host_state[b];
    float* p = params_host + (size_t)b * W;
    p[0] = D.yaw; p[1] = D.trans[0]; p[2] = D.trans[1]; p[3] = D.trans[2]; p[4] = D.scale;
    memcpy(p + 5, rb_lat + (size_t)b * L, sizeof(float) * L);
    if (n_history) n_history[b] = std::min(D.iter, E.max_iters);
  }
  return check_decoder_flags(r, rb_flags[1], rb_flags[0], s);
}

// Optimizer.optimize as ONE call (optimizer.py:56-164 seen from its caller): inputs of slot b from host buffers,
// parameters from / to the caller's device tensors, `iters` iterations, one synchronisation, read-back.
// Two host->device copies (the packed [DetState | SplatView | lidar] block, the NOCS prediction), a prologue kernel, the
// iterations, an epilogue kernel, one device->host copy.
extern "C" int sdfr_refine_optimize(sdfr_refine* r, int b, const float* k_host, const float* kinv_host, int width,
                                    int height, const float* nocs_host, int th, int tw, const float* lidar_host,
                                    int n_lidar, float* yaw_dev, float* trans_dev, float* scale_dev, float* latent_dev,
                                    float* adam_m_host, float* adam_v_host, int* adam_t, int iters, float* params_host,
                                    float* history_host, int* n_history, void* stream) {
  SDFR_REQUIRE(r && k_host && nocs_host && yaw_dev && trans_dev && scale_dev && latent_dev && params_host, SDFR_E_INVALID,
               "null argument");
  SDFR_REQUIRE(b >= 0 && b < r->active, SDFR_E_INVALID, "slot %d is not among the %d active ones", b, r->active);
  SDFR_REQUIRE(width > 0 && height > 0 && width <= r->cfg.max_width && height <= r->cfg.max_height, SDFR_E_CAPACITY,
               "crop %dx%d exceeds the configured capacity %dx%d", width, height, r->cfg.max_width, r->cfg.max_height);
  SDFR_REQUIRE(n_lidar >= 0 && n_lidar <= r->E.max_lidar, SDFR_E_CAPACITY, "%d lidar points exceed capacity %d",
               n_lidar, r->E.max_lidar);
  SDFR_REQUIRE(n_lidar == 0 || lidar_host, SDFR_E_INVALID, "null lidar array");
  SDFR_REQUIRE(th > 0 && tw > 0 && iters >= 0, SDFR_E_INVALID, "bad argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  EngineDev& E = r->E;
  // ---- staging: what sdfr_refine_set_detection + sdfr_refine_set_optimizer_state would copy piecewise ----
  DetState D;
  memset(&D, 0, sizeof(D));     // fresh optimiser state unless the caller continues one (optimizer.py:46-52)
  D.width = width; D.height = height; D.n_lidar = n_lidar; D.target_ready = 1;
  if (adam_m_host && adam_v_host && adam_t && *adam_t > 0) {
    memcpy(D.adam_m, adam_m_host, sizeof(D.adam_m));
    memcpy(D.adam_v, adam_v_host, sizeof(D.adam_v));
    D.adam_t = *adam_t;
  }
  SplatView& V = r->views_host[b];
  V.width = width; V.height = height;
  for (int i = 0; i < 9; ++i) V.k[i] = k_host[i];
  if (kinv_host) for (int i = 0; i < 9; ++i) V.kinv[i] = kinv_host[i];
  else invert3x3(k_host, V.kinv);
  memcpy(r->opt_stage_host, &D, sizeof(D));
  memcpy(r->opt_stage_host + sizeof(D), &V, sizeof(V));
  if (n_lidar > 0) memcpy(r->opt_stage_host + sizeof(D) + sizeof(V), lidar_host, sizeof(float) * 3 * (size_t)n_lidar);
  SDFR_CUDA(cudaMemcpyAsync(r->opt_stage_dev, r->opt_stage_host, sizeof(D) + sizeof(V) + sizeof(float) * 3 * (size_t)n_lidar,
                            cudaMemcpyHostToDevice, s));
  const size_t nocs_count = (size_t)3 * th * tw;
  if (r->nocs_cap[b] < nocs_count) {
    if (r->nocs_dev[b]) SDFR_CUDA(cudaFree(r->nocs_dev[b]));
    r->nocs_dev[b] = nullptr;
    r->nocs_cap[b] = 0;
    SDFR_CUDA(cudaMalloc(&r->nocs_dev[b], nocs_count * sizeof(float)));
    r->nocs_cap[b] = nocs_count;
  }
  SDFR_CUDA(cudaMemcpyAsync(r->nocs_dev[b], nocs_host, nocs_count * sizeof(float), cudaMemcpyHostToDevice, s));
  const int P = width * height;
  optimize_prologue_kernel<<<(P + 255) / 256, 256, 0, s>>>(E, b, reinterpret_cast<const unsigned int*>(r->opt_stage_dev),
                                                           r->nocs_dev[b], th, tw, yaw_dev, trans_dev, scale_dev, latent_dev,
                                                           n_lidar, height, width);
  SDFR_LAUNCH_CHECK();
  r->det_w[b] = width;
  r->det_h[b] = height;
  r->iters_enqueued[b] = 0;
  int rc;
  if ((rc = sdfr_refine_run(r, iters, stream))) return rc;
  // ---- packed read-back ----
  const int nh_host = std::min(r->iters_enqueued[b], E.max_iters);
  optimize_epilogue_kernel<<<1, 256, 0, s>>>(E, b, yaw_dev, trans_dev, scale_dev, latent_dev,
                                             reinterpret_cast<unsigned int*>(r->opt_rb_dev), tc_overflow_ptr(r->dec),
                                             history_host ? nh_host : 0);
  SDFR_LAUNCH_CHECK();
  const size_t head = sizeof(DetState) + sizeof(float) * (size_t)(E.L + 2);
  SDFR_CUDA(cudaMemcpyAsync(r->opt_rb_host, r->opt_rb_dev, head + (history_host ? sizeof(float) * 4 * (size_t)nh_host : 0),
                            cudaMemcpyDeviceToHost, s));
  SDFR_CUDA(cudaStreamSynchronize(s));
  DetState& H = r->host_state[b];
  memcpy(&H, r->opt_rb_host, sizeof(DetState));
  const float* lat = reinterpret_cast<const float*>(r->opt_rb_host + sizeof(DetState));
  params_host[0] = H.yaw; params_host[1] = H.trans[0]; params_host[2] = H.trans[1]; params_host[3] = H.trans[2];
  params_host[4] = H.scale;
  memcpy(params_host + 5, lat, sizeof(float) * (size_t)E.L);
  const int nh = std::min(H.iter, nh_host);
  if (n_history) *n_history = nh;
  if (history_host && nh > 0) memcpy(history_host, r->opt_rb_host + head, sizeof(float) * 4 * (size_t)nh);
  int flags[2];
  memcpy(flags, lat + E.L, sizeof(flags));
  if ((rc = check_decoder_flags(r, flags[1], flags[0], s))) return rc;
  return sdfr_refine_get_optimizer_state(r, b, adam_m_host, adam_v_host, adam_t);
}

extern "C" int sdfr_refine_preselect_error(sdfr_refine* r, float* err_host, void* stream) {
  SDFR_REQUIRE(r && err_host, SDFR_E_INVALID, "bad argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int bits = 0;
  SDFR_CUDA(cudaMemcpyAsync(&bits, r->E.presel_err, sizeof(int), cudaMemcpyDeviceToHost, s));
  SDFR_CUDA(cudaStreamSynchronize(s));
  memcpy(err_host, &bits, sizeof(float));
  return SDFR_OK;
}

extern "C" int sdfr_refine_lattice_rows(sdfr_refine* r, int64_t* rows_host, int64_t* detection_iterations_host,
                                        int reset, void* stream) {
  SDFR_REQUIRE(r && rows_host && detection_iterations_host, SDFR_E_INVALID, "null argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  unsigned long long v[2] = {0ull, 0ull};
  if (r->E.lip > 0.f) {
    SDFR_CUDA(cudaMemcpyAsync(v, r->E.lattice_rows, sizeof(v), cudaMemcpyDeviceToHost, s));
    if (reset) SDFR_CUDA(cudaMemsetAsync(r->E.lattice_rows, 0, sizeof(v), s));
    SDFR_CUDA(cudaStreamSynchronize(s));
  }
  *rows_host = (int64_t)v[0];
  *detection_iterations_host = (int64_t)v[1];
  return SDFR_OK;
}

extern "C" int sdfr_refine_profile(sdfr_refine* r, int iters, float* stage_ms_host, int max_stages, int* n_stages,
                                   int32_t* band_rows_host, void* stream) {
  SDFR_REQUIRE(r && stage_ms_host && iters > 0 && max_stages >= kNumStages, SDFR_E_INVALID, "bad argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int B = r->active;
  for (int b = 0; b < B; ++b)
    SDFR_REQUIRE(r->det_w[b] > 0, SDFR_E_INVALID, "detection %d of the %d active ones has not been set", b, B);
  int qw, qh, any_fine;
  active_shape(r, &qw, &qh, &any_fine);
  std::vector<double> acc(kNumStages, 0.0);
  int rc = SDFR_OK;
  for (int it = 0; it < iters && rc == SDFR_OK; ++it) {
    StageClock clk;
    clk.s = s;
    if (it == 0 && (rc = enqueue_begin(r, B, s))) break;
    rc = enqueue_iteration(r, B, qw, qh, any_fine, s, &clk);
    r->iters_enqueued.assign(r->iters_enqueued.size(), 1 << 30);     // the history is read through D.iter
    cudaStreamSynchronize(s);
    if (rc == SDFR_OK && (int)clk.ev.size() == kNumStages + 1)
      for (int k = 0; k < kNumStages; ++k) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, clk.ev[k], clk.ev[k + 1]);
        acc[k] += ms;
      }
    for (cudaEvent_t e : clk.ev) cudaEventDestroy(e);
  }
  if (rc) return rc;
  r->runs += 1;
  for (int k = 0; k < kNumStages; ++k) stage_ms_host[k] = (float)(acc[k] / iters);
  if (n_stages) *n_stages = kNumStages;
  if (band_rows_host) SDFR_CUDA(cudaMemcpy(band_rows_host, r->E.band_total, sizeof(int32_t), cudaMemcpyDeviceToHost));
  return SDFR_OK;
}

extern "C" const char* sdfr_refine_stage_name(int stage) {
  return stage >= 0 && stage < kNumStages ? kStageNames[stage] : "";
}

extern "C" int sdfr_refine_export(sdfr_refine* r, int b, float* yaw_dev, float* trans_dev, float* scale_dev,
                                  float* latent_dev, void* stream) {
  SDFR_REQUIRE(r && b >= 0 && b < r->cfg.batch, SDFR_E_INVALID, "bad argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  EngineDev& E = r->E;
  export_params_kernel<<<1, 32, 0, s>>>(E.det + b, E.latent + (size_t)b * E.L, E.L, yaw_dev, trans_dev, scale_dev, latent_dev);
  SDFR_LAUNCH_CHECK();
  return SDFR_OK;
}

extern "C" int sdfr_refine_set_latent(sdfr_refine* r, int b, const float* latent_host, void* stream) {
  SDFR_REQUIRE(r && latent_host && b >= 0 && b < r->cfg.batch, SDFR_E_INVALID, "bad argument");
  SDFR_CUDA(cudaMemcpyAsync(r->E.latent + (size_t)b * r->E.L, latent_host, r->E.L * sizeof(float), cudaMemcpyHostToDevice,
                            reinterpret_cast<cudaStream_t>(stream)));
  return SDFR_OK;
}

extern "C" int sdfr_refine_label_extents(sdfr_refine* r, float* extents_host, void* stream) {
  SDFR_REQUIRE(r && extents_host, SDFR_E_INVALID, "bad argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int B = r->active;
  EngineDev E = r->E;
  E.batch = B;
  IterPlan p = make_iter_plan(r, B);
  p.ba.views = nullptr;                       // extents only: the slots need no crop / camera (they may never have been set)
  if (p.prune) {                              // the raw latent is not the trajectory the pruning follows: whole lattice
    p.prune = false;
    p.sel.values = E.sdf; p.sel.in_src = nullptr; p.sel.in_start = nullptr; p.sel.in_count = nullptr;
    p.sel.scatter_values = nullptr; p.sel.scatter_ref = nullptr; p.sel.scatter_flag = nullptr; p.sel.scatter_done = nullptr;
  }
  raw_latent_kernel<<<(B * E.L + 127) / 128, 128, 0, s>>>(E);
  SDFR_LAUNCH_CHECK();
  int rc = enqueue_surface(r, p, s);
  if (rc) return rc;
  extent_kernel<<<B, LB, 0, s>>>(E);
  SDFR_LAUNCH_CHECK();
  SDFR_CUDA(cudaMemcpyAsync(extents_host, E.extents, sizeof(float) * 8 * (size_t)B, cudaMemcpyDeviceToHost, s));
  SDFR_CUDA(cudaStreamSynchronize(s));
  return SDFR_OK;
}

extern "C" int sdfr_refine_view(sdfr_refine* r, int b, int kind, void** ptr_dev, int64_t* count) {
  SDFR_REQUIRE(r && ptr_dev && count && b >= 0 && b < r->cfg.batch, SDFR_E_INVALID, "bad argument");
  EngineDev& E = r->E;
  const SplatView& V = r->views_host[b];
  const int64_t P = (int64_t)V.width * V.height;
  switch (kind) {
    case 0: *ptr_dev = E.sdf + (size_t)b * E.ng; *count = E.ng; break;
    case 1: *ptr_dev = E.dinput + (size_t)b * E.ng * E.in0; *count = E.ng * E.in0; break;
    case 2: *ptr_dev = E.surf_pts + (size_t)b * E.cap * 3; *count = E.cap * 3; break;
    case 3: *ptr_dev = E.surf_nrm + (size_t)b * E.cap * 3; *count = E.cap * 3; break;
    case 4: *ptr_dev = V.color; *count = 3 * P; break;
    case 5: *ptr_dev = V.mask; *count = P; break;
    case 6: *ptr_dev = V.nmap; *count = 3 * P; break;
    case 7: *ptr_dev = E.grads + (size_t)b * (5 + 2 * E.L); *count = 5 + 2 * E.L; break;
    case 8: *ptr_dev = E.surf_count + b; *count = 1; break;
    case 9: *ptr_dev = V.depth; *count = P; break;
    case 10: *ptr_dev = V.cam_v; *count = E.cap * 3; break;
    case 11: *ptr_dev = V.front; *count = E.cap; break;
    case 12: *ptr_dev = E.surf_valid + (size_t)b * E.cap; *count = E.cap; break;
    case 13: *ptr_dev = V.cam_c; *count = E.cap * 3; break;
    case 14: *ptr_dev = E.surf_idx + (size_t)b * E.cap; *count = E.cap; break;
    default: SDFR_REQUIRE(false, SDFR_E_INVALID, "unknown view kind %d", kind);
  }
  return SDFR_OK;
}

extern "C" int sdfr_refine_copy_view(sdfr_refine* r, int b, int kind, void* dst_dev, int64_t max_count, void* stream) {
  void* src = nullptr;
  int64_t count = 0;
  int rc = sdfr_refine_view(r, b, kind, &src, &count);
  if (rc) return rc;
  SDFR_REQUIRE(dst_dev && max_count >= 0, SDFR_E_INVALID, "bad argument");
  const size_t elem = (kind == 11 || kind == 12) ? 1 : 4;
  const size_t n = (size_t)std::min<int64_t>(count, max_count);
  if (n) SDFR_CUDA(cudaMemcpyAsync(dst_dev, src, n * elem, cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream)));
  return SDFR_OK;
}
