// The two screen-space circle primitives of the reference and background compositing.
//
// replaces: inside_circle (sdfrenderer/renderer/primitives.py:4-68; called with diam=0.02, rasterer.py:93-96),
//           inside_circle_opt (primitives.py:71-162; diam=0.025, the 15 x 15 `grid_prim` stamp, rasterer.py:97-100),
//           the `add_bg` rows of all three primitives (primitives.py:58-62, 148-156, 232-237) and the
//           composition with a background image (rasterer.py:107-126).
//
// Only the stand-alone render API reaches these (sdfrenderer/main.py:62-121; the refine loop renders discs without
// a background).  Unlike the disc, a circle's softmax score depends on the point's depth alone
//     s_i = 100 (or 10 000) * max(-z_i / (||z|| + eps) + 1, 0)
// so it is computed once per point (circle_prep_kernel, which also reduces ||z||, the background score
// min_i s_i - 1 and its arg-min); the per-pixel work is the cover test and the softmax.  Two quirks of the
// reference are kept because they are visible in the images:
//   * inside_circle tests its sigmoid "soft clamp" only for > 0: a point covers every pixel until the sigmoid
//     underflows (~29.6 px beyond its radius at the factor 3), and the softmax runs over ALL points with the
//     uncovered ones at logit 0 (`z * mask`, not a masked fill) - they dilute the mask;
//   * inside_circle_opt stamps float-truncated, clamped pixel indices: stamps pile up on the image border.
// Forward: one thread per pixel, points staged through shared memory, two sweeps (max, then sums).  Backward: one
// warp per point gathers over the pixels it covers; the background row's score gradient goes to the arg-min
// point.  No atomics: bit-reproducible.
#include "common.cuh"

namespace sdfr {

namespace {

constexpr int CB = 256;      // threads per block, points per shared-memory chunk

struct CircleParams {
  float diam, gain, soft;
};
__device__ __forceinline__ CircleParams circle_params(int primitive) {
  // rasterer.py:93-100 (diam) with the defaults of primitives.py:10-13 / 80-83
  return primitive == SDFR_PRIM_CIRCLE ? CircleParams{0.02f, 100.f, 3.f} : CircleParams{0.025f, 10000.f, 5.f};
}

// scalars[]: 0 nu, 1 background score, 2 arg-min point (int bits), 3 d score of the background row (backward)
__global__ void __launch_bounds__(1024) circle_prep_kernel(const SplatView* __restrict__ views) {
  __shared__ float s_red[32];
  __shared__ int s_idx[32];
  __shared__ float s_nu;
  const SplatView& V = views[blockIdx.x];
  const int m = min(V.count ? *V.count : V.static_count, V.capacity);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const CircleParams cp = circle_params(V.primitive);
  // ||z||_2 over all points (primitives.py:54 / 144; detached)
  float sq = 0.f;
  for (int i = tid; i < m; i += 1024) { const float z = V.cam_v[i * 3 + 2]; sq += z * z; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if (lane == 0) s_red[warp] = sq;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < 32; ++w) t += s_red[w];
    s_nu = sqrtf(t);
  }
  __syncthreads();
  const float nu = s_nu;
  const float k00 = V.k[0], k01 = V.k[1], k02 = V.k[2], k10 = V.k[3], k11 = V.k[4], k12 = V.k[5], k20 = V.k[6],
              k21 = V.k[7], k22 = V.k[8];
  float best = INFINITY;
  int best_i = 0x7fffffff;
  for (int i = tid; i < m; i += 1024) {
    const float x = V.cam_v[i * 3], y = V.cam_v[i * 3 + 1], z = V.cam_v[i * 3 + 2];
    const float t = (-z) / (nu + kEps32) + 1.f;                          // primitives.py:55-56
    const float s = fmaxf(t, 0.f) * cp.gain;
    V.score[i] = s;
    // projection.py:88-93: K v, divide by (z' + eps), clamp to [-1, res]
    const float hx = k00 * x + k01 * y + k02 * z, hy = k10 * x + k11 * y + k12 * z, hz = k20 * x + k21 * y + k22 * z;
    V.p2[i * 2] = fminf(fmaxf(hx / (hz + kEps32), -1.f), (float)V.width);
    V.p2[i * 2 + 1] = fminf(fmaxf(hy / (hz + kEps32), -1.f), (float)V.height);
    V.radius[i] = fabsf(k00 * cp.diam / (z + kEps32));                  // primitives.py:44-45 / 112
    if (s < best) { best = s; best_i = i; }
  }
  // min score and its first index: the background row scores min - 1 (primitives.py:59-60)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ob < best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
  }
  __syncthreads();
  if (lane == 0) { s_red[warp] = best; s_idx[warp] = best_i; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 32; ++w)
      if (s_red[w] < best || (s_red[w] == best && s_idx[w] < best_i)) { best = s_red[w]; best_i = s_idx[w]; }
    V.prim_scalars[0] = nu;
    V.prim_scalars[1] = m > 0 ? best - 1.f : 0.f;
    reinterpret_cast<int*>(V.prim_scalars)[2] = m > 0 ? best_i : -1;
    V.prim_scalars[3] = 0.f;
  }
}

// the disc primitive's background score: min_i(-v_z * gain) - 1 (primitives.py:233); pixel independent
__global__ void __launch_bounds__(1024) disc_bg_prep_kernel(const SplatView* __restrict__ views) {
  __shared__ float s_red[32];
  const SplatView& V = views[blockIdx.x];
  const int m = min(V.count ? *V.count : V.static_count, V.capacity);
  float best = INFINITY;
  for (int i = threadIdx.x; i < m; i += 1024) best = fminf(best, -V.cam_v[i * 3 + 2] * kDepthGain);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 32; ++w) best = fminf(best, s_red[w]);
    V.prim_scalars[1] = m > 0 ? best - 1.f : 0.f;
  }
}

// ---- cover tests -----------------------------------------------------------------------------------------
// inside_circle: sigmoid((radius - dist) * soft) > 0 in float32 (primitives.py:41-50)
__device__ __forceinline__ bool circle_covers(float px, float py, float radius, float soft, float x, float y) {
  const float dx = px - x, dy = py - y;
  const float arg = (radius - sqrtf(dx * dx + dy * dy)) * soft;
  return 1.f / (1.f + expf(-arg)) > 0.f;
}
// inside_circle_opt: is `c` one of clamp(trunc(p + o), 0, hi) for o = -7..7 (primitives.py:118-127)?
__device__ __forceinline__ bool stamp_covers_axis(float p, int c, int hi) {
  if (fabsf((float)c - p) > 9.f && c != 0 && c != hi) return false;
#pragma unroll
  for (int o = -7; o <= 7; ++o) {
    const int v = min(max((int)truncf(p + (float)o), 0), hi);
    if (v == c) return true;
  }
  return false;
}

__global__ void __launch_bounds__(CB) circle_forward_kernel(const SplatView* __restrict__ views) {
  __shared__ float s_p2[CB][2], s_rad[CB], s_score[CB], s_c[CB][3], s_n[CB][3], s_z[CB];
  const SplatView& V = views[blockIdx.y];
  const int P = V.width * V.height;
  const int m = min(V.count ? *V.count : V.static_count, V.capacity);
  const int j = blockIdx.x * CB + threadIdx.x;
  const bool live = j < P;
  const int y = live ? j / V.width : 0, x = live ? j - y * V.width : 0;
  const float fx = (float)x, fy = (float)y;
  const bool opt = V.primitive == SDFR_PRIM_CIRCLE_OPT;
  const CircleParams cp = circle_params(V.primitive);
  // circle_opt takes the image size from the principal point (primitives.py:108-109)
  const int xhi = opt ? (int)V.k[2] * 2 - 1 : V.width - 1, yhi = opt ? (int)V.k[5] * 2 - 1 : V.height - 1;
  const float s_bg = V.prim_scalars[1];
  float mx = V.has_bg ? s_bg : -INFINITY;
  int covered = 0;
  auto stage = [&](int base) {
    const int i = base + threadIdx.x;
    __syncthreads();
    if (i < m) {
      s_p2[threadIdx.x][0] = V.p2[i * 2]; s_p2[threadIdx.x][1] = V.p2[i * 2 + 1];
      s_rad[threadIdx.x] = V.radius[i];
      s_score[threadIdx.x] = V.score[i];
      s_z[threadIdx.x] = V.cam_v[i * 3 + 2];
#pragma unroll
      for (int c = 0; c < 3; ++c) { s_c[threadIdx.x][c] = V.cam_c[i * 3 + c]; s_n[threadIdx.x][c] = (V.cam_m[i * 3 + c] + 1.f) / 2.f; }
    }
    __syncthreads();
  };
  auto covers = [&](int k) {
    return opt ? (stamp_covers_axis(s_p2[k][0], x, xhi) && stamp_covers_axis(s_p2[k][1], y, yhi))
               : circle_covers(s_p2[k][0], s_p2[k][1], s_rad[k], cp.soft, fx, fy);
  };
  // sweep 0: the largest logit (torch.softmax subtracts it)
  for (int base = 0; base < m; base += CB) {
    stage(base);
    if (live) {
      const int n = min(CB, m - base);
      for (int k = 0; k < n; ++k)
        if (covers(k)) { mx = fmaxf(mx, s_score[k]); ++covered; }
    }
  }
  if (!opt && covered < m) mx = fmaxf(mx, 0.f);          // inside_circle: uncovered points sit at logit 0
  // sweep 1: sums
  float den = 0.f, acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int base = 0; base < m; base += CB) {
    stage(base);
    if (live) {
      const int n = min(CB, m - base);
      for (int k = 0; k < n; ++k) {
        if (!covers(k)) continue;
        const float e = expf(s_score[k] - mx);
        den += e;
        acc[0] += e * s_c[k][0]; acc[1] += e * s_c[k][1]; acc[2] += e * s_c[k][2];
        acc[3] += e;
        acc[4] += e * s_z[k];
        acc[5] += e * s_n[k][0]; acc[6] += e * s_n[k][1]; acc[7] += e * s_n[k][2];
      }
    }
  }
  if (!live) return;
  if (!opt) den += (float)(m - covered) * expf(0.f - mx);   // the uncovered points' share of the denominator
  if (V.has_bg) {
    const float e = expf(s_bg - mx);
    den += e;
    acc[0] += e * V.bg[j]; acc[1] += e * V.bg[P + j]; acc[2] += e * V.bg[2 * P + j];
    acc[3] += e;
  }
  const float inv_den = den > 0.f ? 1.f / den : 0.f;        // a pixel nothing covers: uniform softmax times zero cover
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] *= inv_den;
  if (V.color) {
    V.color[j] = fminf(acc[0], 1.f); V.color[P + j] = fminf(acc[1], 1.f); V.color[2 * P + j] = fminf(acc[2], 1.f);
  }
  if (V.mask) V.mask[j] = fminf(acc[3], 1.f);
  if (V.depth) V.depth[j] = acc[4];
  if (V.nmap) {
    V.nmap[j] = fminf(acc[5], 1.f); V.nmap[P + j] = fminf(acc[6], 1.f); V.nmap[2 * P + j] = fminf(acc[7], 1.f);
  }
  *reinterpret_cast<float4*>(V.pix_stat + (size_t)j * 4) = make_float4(0.f, mx, inv_den, (float)covered);
  *reinterpret_cast<float4*>(V.pix_raw + (size_t)j * 8) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  *reinterpret_cast<float4*>(V.pix_raw + (size_t)j * 8 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
}

// disc + background: pixels no surfel hits show the background with weight 1 (everywhere else the background
// row's softmax weight exp(-v_z*150 - 1 - s_max) is exactly 0 in float32)
__global__ void __launch_bounds__(CB) disc_bg_compose_kernel(const SplatView* __restrict__ views) {
  const SplatView& V = views[blockIdx.y];
  const int P = V.width * V.height;
  const int j = blockIdx.x * CB + threadIdx.x;
  if (j >= P) return;
  const float4 st = *reinterpret_cast<const float4*>(V.pix_stat + (size_t)j * 4);
  if (st.w != 0.f) return;
  float* raw = V.pix_raw + (size_t)j * 8;
  for (int c = 0; c < 3; ++c) {
    const float b = V.bg[c * P + j];
    raw[c] = b;
    if (V.color) V.color[c * P + j] = fminf(b, 1.f);
  }
  raw[3] = 1.f;
  if (V.mask) V.mask[j] = 1.f;
}

// One warp per point: d loss / d (score, colour, depth, normals) gathered over the pixels the point covers.
__global__ void __launch_bounds__(CB) circle_backward_kernel(const SplatView* __restrict__ views) {
  const SplatView& V = views[blockIdx.y];
  const int m = min(V.count ? *V.count : V.static_count, V.capacity);
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (CB / 32) + (threadIdx.x >> 5);
  if (i >= m) return;
  const bool opt = V.primitive == SDFR_PRIM_CIRCLE_OPT;
  const CircleParams cp = circle_params(V.primitive);
  const int xhi = opt ? (int)V.k[2] * 2 - 1 : V.width - 1, yhi = opt ? (int)V.k[5] * 2 - 1 : V.height - 1;
  const float px = V.p2[i * 2], py = V.p2[i * 2 + 1], rad = V.radius[i], s = V.score[i];
  const float cx = V.cam_c[i * 3], cy = V.cam_c[i * 3 + 1], cz = V.cam_c[i * 3 + 2];
  const float vz = V.cam_v[i * 3 + 2];
  const float nxd = (V.cam_m[i * 3] + 1.f) / 2.f, nyd = (V.cam_m[i * 3 + 1] + 1.f) / 2.f, nzd = (V.cam_m[i * 3 + 2] + 1.f) / 2.f;
  // pixel range to visit: the whole image for inside_circle (its cover reaches ~30 px), the stamp's box otherwise
  int x0 = 0, y0 = 0, x1 = V.width - 1, y1 = V.height - 1;
  if (opt) {
    x0 = min(max((int)floorf(px) - 8, 0), V.width - 1); x1 = max(min((int)floorf(px) + 8, V.width - 1), 0);
    y0 = min(max((int)floorf(py) - 8, 0), V.height - 1); y1 = max(min((int)floorf(py) + 8, V.height - 1), 0);
  }
  const int bw = x1 - x0 + 1, bh = y1 - y0 + 1, npx = bw * bh;
  float ds = 0.f, dcx = 0.f, dcy = 0.f, dcz = 0.f, dvz = 0.f, dnx = 0.f, dny = 0.f, dnz = 0.f;
  for (int idx = lane; idx < npx; idx += 32) {
    const int yy = idx / bw, x = x0 + (idx - yy * bw), y = y0 + yy;
    const bool cov = opt ? (stamp_covers_axis(px, x, xhi) && stamp_covers_axis(py, y, yhi))
                         : circle_covers(px, py, rad, cp.soft, (float)x, (float)y);
    if (!cov) continue;
    const int j = y * V.width + x;
    const float4 st = *reinterpret_cast<const float4*>(V.pix_stat + (size_t)j * 4);
    const float w = expf(s - st.y) * st.z;
    const float* pg = V.pix_grad + (size_t)j * 12;
    const float4 g0 = *reinterpret_cast<const float4*>(pg);       // g_colour'(3), g_depth
    const float4 g1 = *reinterpret_cast<const float4*>(pg + 4);   // g_normals'(3), g_mask'
    const float G = pg[8];
    const float gbar = cx * g0.x + cy * g0.y + cz * g0.z + vz * g0.w + nxd * g1.x + nyd * g1.y + nzd * g1.z + g1.w;
    ds += w * (gbar - G);
    dcx += w * g0.x; dcy += w * g0.y; dcz += w * g0.z;
    dvz += w * g0.w;
    dnx += w * g1.x; dny += w * g1.y; dnz += w * g1.z;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ds += __shfl_xor_sync(0xffffffffu, ds, o);
    dcx += __shfl_xor_sync(0xffffffffu, dcx, o); dcy += __shfl_xor_sync(0xffffffffu, dcy, o);
    dcz += __shfl_xor_sync(0xffffffffu, dcz, o);
    dvz += __shfl_xor_sync(0xffffffffu, dvz, o);
    dnx += __shfl_xor_sync(0xffffffffu, dnx, o); dny += __shfl_xor_sync(0xffffffffu, dny, o);
    dnz += __shfl_xor_sync(0xffffffffu, dnz, o);
  }
  if (lane == 0) {
    V.d_score[i] = ds;
    V.d_v[i * 3] = 0.f; V.d_v[i * 3 + 1] = 0.f; V.d_v[i * 3 + 2] = dvz;      // completed by circle_finish_kernel
    V.d_m[i * 3] = 0.5f * dnx; V.d_m[i * 3 + 1] = 0.5f * dny; V.d_m[i * 3 + 2] = 0.5f * dnz;
    V.d_c[i * 3] = dcx; V.d_c[i * 3 + 1] = dcy; V.d_c[i * 3 + 2] = dcz;
  }
}

// d score of the background row = sum_j w_bg,j (gbar_bg,j - G_j), gbar_bg = bg . g_colour' + g_mask'
// (one block, ordered reduction); it belongs to the arg-min point (z.min() - 1, primitives.py:59).
__global__ void __launch_bounds__(1024) circle_bg_backward_kernel(const SplatView* __restrict__ views) {
  __shared__ float s_red[32];
  const SplatView& V = views[blockIdx.x];
  const int P = V.width * V.height;
  const float s_bg = V.prim_scalars[1];
  float acc = 0.f;
  for (int j = threadIdx.x; j < P; j += 1024) {
    const float4 st = *reinterpret_cast<const float4*>(V.pix_stat + (size_t)j * 4);
    const float w = expf(s_bg - st.y) * st.z;
    const float* pg = V.pix_grad + (size_t)j * 12;
    const float gbar = V.bg[j] * pg[0] + V.bg[P + j] * pg[1] + V.bg[2 * P + j] * pg[2] + pg[7];
    acc += w * (gbar - pg[8]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 32; ++w) t += s_red[w];
    V.prim_scalars[3] = t;
  }
}

// score -> depth: s = gain * max(-z / (nu + eps) + 1, 0) with nu detached
__global__ void __launch_bounds__(CB) circle_finish_kernel(const SplatView* __restrict__ views) {
  const SplatView& V = views[blockIdx.y];
  const int m = min(V.count ? *V.count : V.static_count, V.capacity);
  const int i = blockIdx.x * CB + threadIdx.x;
  if (i >= m) return;
  const CircleParams cp = circle_params(V.primitive);
  const float nu = V.prim_scalars[0];
  float ds = V.d_score[i];
  if (V.has_bg && i == reinterpret_cast<const int*>(V.prim_scalars)[2]) ds += V.prim_scalars[3];
  const float t = (-V.cam_v[i * 3 + 2]) / (nu + kEps32) + 1.f;
  if (t > 0.f) V.d_v[i * 3 + 2] += ds * cp.gain * (-1.f / (nu + kEps32));
}

}  // namespace

int launch_circle_forward(const SplatView* views_dev, int batch, int max_pixels, cudaStream_t s) {
  if (batch <= 0 || max_pixels <= 0) return SDFR_OK;
  circle_prep_kernel<<<batch, 1024, 0, s>>>(views_dev);
  SDFR_LAUNCH_CHECK();
  circle_forward_kernel<<<dim3((max_pixels + CB - 1) / CB, batch), CB, 0, s>>>(views_dev);
  SDFR_LAUNCH_CHECK();
  return SDFR_OK;
}

int launch_circle_backward(const SplatView* views_dev, int batch, int max_count, int has_bg, cudaStream_t s) {
  if (batch <= 0 || max_count <= 0) return SDFR_OK;
  circle_backward_kernel<<<dim3((max_count + CB / 32 - 1) / (CB / 32), batch), CB, 0, s>>>(views_dev);
  SDFR_LAUNCH_CHECK();
  if (has_bg) {
    circle_bg_backward_kernel<<<batch, 1024, 0, s>>>(views_dev);
    SDFR_LAUNCH_CHECK();
  }
  circle_finish_kernel<<<dim3((max_count + CB - 1) / CB, batch), CB, 0, s>>>(views_dev);
  SDFR_LAUNCH_CHECK();
  return SDFR_OK;
}

int launch_disc_background(const SplatView* views_dev, int batch, int max_pixels, cudaStream_t s) {
  if (batch <= 0 || max_pixels <= 0) return SDFR_OK;
  disc_bg_prep_kernel<<<batch, 1024, 0, s>>>(views_dev);
  SDFR_LAUNCH_CHECK();
  disc_bg_compose_kernel<<<dim3((max_pixels + CB - 1) / CB, batch), CB, 0, s>>>(views_dev);
  SDFR_LAUNCH_CHECK();
  return SDFR_OK;
}

}  // namespace sdfr
