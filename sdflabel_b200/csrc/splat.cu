// Surfel ("disc") splatting: projection, per-pixel depth-softmax forward and
// the per-surfel gather backward.
//
// replaces: project_in_2D / project_in_2D_quat (sdfrenderer/renderer/projection.py:7-101,104-199),
//           qrot (renderer/utils_rasterer.py:6-24), inside_surfel (renderer/primitives.py:165-243,
//           diam=0.04, softclamp=False, add_bg=False) and the composition in
//           Rasterer.forward (renderer/rasterer.py:113-144).
//
// The reference materialises >= 6 tensors of N_surf x P x {1,3}; here nothing of
// that size exists.  Every hit lies inside the ball of radius 0.04 around the
// surfel centre, so a surfel can only touch pixels inside the projected bounding
// box of that ball: the forward is a gather per 16x16 pixel tile over the surfels
// whose box meets the tile (two sweeps: sum zeta^2 / max, then the softmax), the
// backward a gather per surfel (one warp) over the pixels of its box.  Both are
// deterministic (no atomics).  Traffic: 44 B per surfel in, 8 floats per pixel out.
#include <algorithm>

#include "common.cuh"
#include "project.cuh"

namespace sdfr {

namespace {

__global__ void __launch_bounds__(256) project_kernel(const SplatView* __restrict__ views) {
  const SplatView& V = views[blockIdx.y];
  const int m = V.count ? *V.count : V.static_count;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m || i >= V.capacity) return;
  project_surfel(V, i, V.coords[i * 3], V.coords[i * 3 + 1], V.coords[i * 3 + 2], V.normals[i * 3], V.normals[i * 3 + 1],
                 V.normals[i * 3 + 2], !V.valid || V.valid[i]);
}

// ---------------------------------------------------------------------------
// ray / tangent-disc test shared by forward and backward (primitives.py:202-226)
// ---------------------------------------------------------------------------
struct Hit {
  bool hit, overwritten;
  float b, z;
};

__device__ __forceinline__ Hit disc_test(float rx, float ry, float rz, float vx, float vy, float vz, float mx,
                                         float my, float mz, float a) {
  Hit h;
  float b = rx * mx + ry * my + rz * mz;
  h.overwritten = fabsf(b) < kRayCutoff;
  if (h.overwritten) b = kEps32;
  h.b = b;
  h.z = a / b;
  const float dx = vx - rx * h.z, dy = vy - ry * h.z, dz = vz - rz * h.z;
  const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
  h.hit = (kDiscRadius - dist) > 0.f;
  return h;
}

constexpr int CHUNK = 256;     // threads per block = tile * tile * split
constexpr int LCAP = 1024;     // surfels of one tile kept resident in shared memory for both sweeps
// Crops above kFineCropPixels: 8 x 8 pixel tiles with FOUR threads per pixel, each sweeping a quarter of the tile's surfel
// list (a 256 x 256 crop is 1024 CTAs; the critical path of a busy tile is a quarter of what one thread per pixel on
// 16 x 16 tiles pays: 66 -> 30 us).  Crops up to kFineCropPixels (the reference's default regime: rendering_area 32 - 64,
// configs/config_refine.ini:12): 4 x 4 tiles with SIXTEEN threads per pixel - a 32 x 32 crop is 64 CTAs instead of 16,
// and the ~900 surfels that meet a central 8 x 8 tile there (close to the list's capacity: one tile past it ran the
// chunk-serial path and took 96 us) become ~400.  The mode depends on the detection's OWN crop only, so its bits do not
// depend on what else is in the batch.
constexpr int kFineCropPixels = 64 * 64;

__global__ void __launch_bounds__(CHUNK) splat_forward_kernel(const SplatView* __restrict__ views) {
  extern __shared__ __align__(16) float s_list[];
  const SplatView& V = views[blockIdx.z];
  constexpr int lcap = LCAP;
  const bool fine = V.width * V.height <= kFineCropPixels;
  const int tile = fine ? 4 : 8, split = fine ? 16 : 4;
  const int tx0 = blockIdx.x * tile, ty0 = blockIdx.y * tile;
  if (tx0 >= V.width || ty0 >= V.height) return;
  const int m = min(V.count ? *V.count : V.static_count, V.capacity);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pix = fine ? tid >> 4 : tid >> 2, part = fine ? tid & 15 : tid & 3;
  const int x = tx0 + (fine ? pix & 3 : pix & 7), y = ty0 + (fine ? pix >> 2 : pix >> 3);
  const bool live = x < V.width && y < V.height;
  const int tx1 = min(tx0 + tile - 1, V.width - 1), ty1 = min(ty0 + tile - 1, V.height - 1);

  float (*s_v)[3] = reinterpret_cast<float(*)[3]>(s_list);
  float (*s_m)[3] = reinterpret_cast<float(*)[3]>(s_list + 3 * lcap);
  float (*s_c)[3] = reinterpret_cast<float(*)[3]>(s_list + 6 * lcap);
  float* s_a = s_list + 9 * lcap;
  int4* s_bb = reinterpret_cast<int4*>(s_list + 10 * lcap);
  __shared__ int s_warp_cnt[CHUNK / 32];

  const float fx = (float)x, fy = (float)y;
  const float rx = V.kinv[0] * fx + V.kinv[1] * fy + V.kinv[2];
  const float ry = V.kinv[3] * fx + V.kinv[4] * fy + V.kinv[5];
  const float rz = V.kinv[6] * fx + V.kinv[7] * fy + V.kinv[8];

  float sumsq = 0.f, zeta_max = -INFINITY;
  int hits = 0;
  float nu = 0.f, smax = 0.f, den = 0.f;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // colour3, mask, depth, normals3

  // one candidate of the tile's list against this pixel; sweep 0: sum of squares / max / count of zeta,
  // sweep 1: softmax-weighted accumulation (primitives.py:227-240, rasterer.py:113-144)
  auto visit = [&](const int k, const int sweep) {
    // every hit of a surfel lies inside its pixel box (project_kernel): four integer compares reject the
    // ~85 % of the tile's surfels that cannot touch this pixel before the ray / disc test
    const int4 b = s_bb[k];
    if (x < b.x || x > b.z || y < b.y || y > b.w) return;
    const Hit h = disc_test(rx, ry, rz, s_v[k][0], s_v[k][1], s_v[k][2], s_m[k][0], s_m[k][1], s_m[k][2], s_a[k]);
    if (!h.hit) return;
    const float zeta = -h.z;                      // primitives.py:227
    if (sweep == 0) {
      sumsq += zeta * zeta;                       // primitives.py:228
      zeta_max = fmaxf(zeta_max, zeta);
      ++hits;
    } else {
      const float sc = fmaxf(zeta / (nu + kEps32) + 1.f, 0.f) * kDepthGain;   // primitives.py:229-230
      const float e = expf(sc - smax);            // softmax numerator (primitives.py:240)
      den += e;
      acc[0] += e * s_c[k][0]; acc[1] += e * s_c[k][1]; acc[2] += e * s_c[k][2];
      acc[3] += e;
      acc[4] += e * s_v[k][2];                    // rasterer.py:136 (surfel-centre z)
      acc[5] += e * ((s_m[k][0] + 1.f) / 2.f);
      acc[6] += e * ((s_m[k][1] + 1.f) / 2.f);
      acc[7] += e * ((s_m[k][2] + 1.f) / 2.f);
    }
  };
  // the threads of a pixel hold partial results over their shares of the list: combine them in a fixed order
  // (consecutive lanes of one warp)
  auto combine_sweep0 = [&]() {
    for (int o = 1; o < split; o <<= 1) {
      sumsq += __shfl_xor_sync(0xffffffffu, sumsq, o);
      zeta_max = fmaxf(zeta_max, __shfl_xor_sync(0xffffffffu, zeta_max, o));
      hits += __shfl_xor_sync(0xffffffffu, hits, o);
    }
  };
  auto combine_sweep1 = [&]() {
    for (int o = 1; o < split; o <<= 1) {
      den += __shfl_xor_sync(0xffffffffu, den, o);
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    }
  };
  // culls surfels [base, base + CHUNK) against the tile and appends the survivors (in index order) to the
  // shared list at `start`; entries past `cap` are dropped.  Returns the number of survivors of the chunk.
  auto cull_chunk = [&](const int base, const int start, const int cap) -> int {
    const int i = base + tid;
    bool take = false;
    int4 bb = make_int4(0, 0, 0, 0);
    if (i < m) {
      bb = *reinterpret_cast<const int4*>(V.bbox + i * 4);
      take = bb.x <= tx1 && bb.z >= tx0 && bb.y <= ty1 && bb.w >= ty0;
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, take);
    __syncthreads();                              // the previous chunk's readers of s_warp_cnt / the list are done
    if (lane == 0) s_warp_cnt[warp] = __popc(ballot);
    __syncthreads();
    int off = start + __popc(ballot & ((1u << lane) - 1u)), total = 0;
#pragma unroll
    for (int w = 0; w < CHUNK / 32; ++w) {
      const int c = s_warp_cnt[w];
      if (w < warp) off += c;
      total += c;
    }
    if (take && off < cap) {
      s_v[off][0] = V.cam_v[i * 3]; s_v[off][1] = V.cam_v[i * 3 + 1]; s_v[off][2] = V.cam_v[i * 3 + 2];
      s_m[off][0] = V.cam_m[i * 3]; s_m[off][1] = V.cam_m[i * 3 + 1]; s_m[off][2] = V.cam_m[i * 3 + 2];
      s_c[off][0] = V.cam_c[i * 3]; s_c[off][1] = V.cam_c[i * 3 + 1]; s_c[off][2] = V.cam_c[i * 3 + 2];
      s_a[off] = V.plane_a[i];
      s_bb[off] = bb;
    }
    return total;
  };

  // Pass 1: cull everything once.  If the tile's surfels fit the resident list (the usual case) both sweeps run
  // out of shared memory; the chunk-serial cull with its dependent global loads and block barriers is paid once
  // instead of twice.
  int total = 0;
  for (int base = 0; base < m && total <= lcap; base += CHUNK) total += cull_chunk(base, total, lcap);
  __syncthreads();
  if (total <= lcap) {
    if (live)
      for (int k = part; k < total; k += split) visit(k, 0);
    combine_sweep0();
    nu = sqrtf(sumsq);
    smax = fmaxf(zeta_max / (nu + kEps32) + 1.f, 0.f) * kDepthGain;
    if (live)
      for (int k = part; k < total; k += split) visit(k, 1);
    combine_sweep1();
  } else {
    // more survivors than the list holds: chunk by chunk, per sweep
    for (int sweep = 0; sweep < 2; ++sweep) {
      if (sweep == 1) {
        combine_sweep0();
        nu = sqrtf(sumsq);
        smax = fmaxf(zeta_max / (nu + kEps32) + 1.f, 0.f) * kDepthGain;
      }
      for (int base = 0; base < m; base += CHUNK) {
        const int cnt = cull_chunk(base, 0, CHUNK);
        __syncthreads();
        if (live)
          for (int k = part; k < cnt; k += split) visit(k, sweep);
      }
    }
    combine_sweep1();
  }
  if (!live || part != 0) return;
  const int P = V.width * V.height, j = y * V.width + x;
  const float inv_den = hits > 0 ? 1.f / den : 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] *= inv_den;
  if (V.color) {   // clamp(max=1): rasterer.py:124
    V.color[j] = fminf(acc[0], 1.f); V.color[P + j] = fminf(acc[1], 1.f); V.color[2 * P + j] = fminf(acc[2], 1.f);
  }
  if (V.mask) V.mask[j] = fminf(acc[3], 1.f);
  if (V.depth) V.depth[j] = acc[4];
  if (V.nmap) {
    V.nmap[j] = fminf(acc[5], 1.f); V.nmap[P + j] = fminf(acc[6], 1.f); V.nmap[2 * P + j] = fminf(acc[7], 1.f);
  }
  if (V.pix_stat) *reinterpret_cast<float4*>(V.pix_stat + (size_t)j * 4) = make_float4(nu, smax, inv_den, (float)hits);
  if (V.pix_raw) {
    *reinterpret_cast<float4*>(V.pix_raw + (size_t)j * 8) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(V.pix_raw + (size_t)j * 8 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// Upstream map gradients -> per-pixel record used by the surfel gather.
// g' = g * [unclamped <= 1] (clamp(max=1) backward); G = sum_k w_kj gbar_kj = <raw, g'>.
__global__ void pixel_grad_prep_kernel(const SplatView* __restrict__ views, const float* __restrict__ g_color,
                                       const float* __restrict__ g_mask, const float* __restrict__ g_depth,
                                       const float* __restrict__ g_nmap) {
  const SplatView& V = views[blockIdx.y];
  const int P = V.width * V.height;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= P) return;
  const float* raw = V.pix_raw + (size_t)j * 8;
  float g[8];
  g[0] = g_color ? g_color[j] : 0.f; g[1] = g_color ? g_color[P + j] : 0.f; g[2] = g_color ? g_color[2 * P + j] : 0.f;
  g[3] = g_mask ? g_mask[j] : 0.f;
  g[4] = g_depth ? g_depth[j] : 0.f;
  g[5] = g_nmap ? g_nmap[j] : 0.f; g[6] = g_nmap ? g_nmap[P + j] : 0.f; g[7] = g_nmap ? g_nmap[2 * P + j] : 0.f;
  float G = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    if (c != 4 && !(raw[c] <= 1.f)) g[c] = 0.f;
    G += raw[c] * g[c];
  }
  float* o = V.pix_grad + (size_t)j * 12;
  *reinterpret_cast<float4*>(o) = make_float4(g[0], g[1], g[2], g[4]);
  *reinterpret_cast<float4*>(o + 4) = make_float4(g[5], g[6], g[7], g[3]);
  *reinterpret_cast<float4*>(o + 8) = make_float4(G, 0.f, 0.f, 0.f);
}

// One warp per surfel: gathers d loss / d (v, m, composited colour) over the surfel's pixel box.
__device__ __forceinline__ void splat_backward_surfel(const SplatView& V, const int i, const int lane) {
  const float vx = V.cam_v[i * 3], vy = V.cam_v[i * 3 + 1], vz = V.cam_v[i * 3 + 2];
  const float mx = V.cam_m[i * 3], my = V.cam_m[i * 3 + 1], mz = V.cam_m[i * 3 + 2];
  const float cx = V.cam_c[i * 3], cy = V.cam_c[i * 3 + 1], cz = V.cam_c[i * 3 + 2];
  const float a = V.plane_a[i];
  const int4 bb = *reinterpret_cast<const int4*>(V.bbox + i * 4);
  const int bw = bb.z - bb.x + 1, bh = bb.w - bb.y + 1;
  const int npx = (bw > 0 && bh > 0) ? bw * bh : 0;
  float da = 0.f, dmx = 0.f, dmy = 0.f, dmz = 0.f, dcx = 0.f, dcy = 0.f, dcz = 0.f, dvz = 0.f;
  float dnx = 0.f, dny = 0.f, dnz = 0.f;
  const float nxd = (mx + 1.f) / 2.f, nyd = (my + 1.f) / 2.f, nzd = (mz + 1.f) / 2.f;
  for (int idx = lane; idx < npx; idx += 32) {
    const int yy = idx / bw, x = bb.x + (idx - yy * bw), y = bb.y + yy;
    const int j = y * V.width + x;
    const float4 st = *reinterpret_cast<const float4*>(V.pix_stat + (size_t)j * 4);
    if (st.w == 0.f) continue;
    const float fx = (float)x, fy = (float)y;
    const float rx = V.kinv[0] * fx + V.kinv[1] * fy + V.kinv[2];
    const float ry = V.kinv[3] * fx + V.kinv[4] * fy + V.kinv[5];
    const float rz = V.kinv[6] * fx + V.kinv[7] * fy + V.kinv[8];
    const Hit h = disc_test(rx, ry, rz, vx, vy, vz, mx, my, mz, a);
    if (!h.hit) continue;
    const float zeta = -h.z;
    const float t = zeta / (st.x + kEps32) + 1.f;
    const float sc = fmaxf(t, 0.f) * kDepthGain;
    const float w = expf(sc - st.y) * st.z;
    const float* pg = V.pix_grad + (size_t)j * 12;
    const float4 g0 = *reinterpret_cast<const float4*>(pg);       // g_colour'(3), g_depth
    const float4 g1 = *reinterpret_cast<const float4*>(pg + 4);   // g_normals'(3), g_mask'
    const float G = pg[8];
    const float gbar = cx * g0.x + cy * g0.y + cz * g0.z + vz * g0.w + nxd * g1.x + nyd * g1.y + nzd * g1.z + g1.w;
    const float ds = w * (gbar - G);                               // softmax backward
    const float dzeta = t > 0.f ? ds * kDepthGain / (st.x + kEps32) : 0.f;
    const float dz = -dzeta;                                       // zeta = -z * mu
    da += dz / h.b;                                                // z = a / b
    if (!h.overwritten) {                                          // in-place overwrite kills this path (primitives.py:210)
      const float db = -dz * h.z / h.b;
      dmx += db * rx; dmy += db * ry; dmz += db * rz;
    }
    dcx += w * g0.x; dcy += w * g0.y; dcz += w * g0.z;
    dvz += w * g0.w;
    dnx += w * g1.x; dny += w * g1.y; dnz += w * g1.z;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    da += __shfl_xor_sync(0xffffffffu, da, o);
    dmx += __shfl_xor_sync(0xffffffffu, dmx, o); dmy += __shfl_xor_sync(0xffffffffu, dmy, o);
    dmz += __shfl_xor_sync(0xffffffffu, dmz, o);
    dcx += __shfl_xor_sync(0xffffffffu, dcx, o); dcy += __shfl_xor_sync(0xffffffffu, dcy, o);
    dcz += __shfl_xor_sync(0xffffffffu, dcz, o);
    dvz += __shfl_xor_sync(0xffffffffu, dvz, o);
    dnx += __shfl_xor_sync(0xffffffffu, dnx, o); dny += __shfl_xor_sync(0xffffffffu, dny, o);
    dnz += __shfl_xor_sync(0xffffffffu, dnz, o);
  }
  if (lane == 0) {
    // a = m.v  ->  dv += da*m, dm += da*v ; depth uses v.z ; normals map uses (m+1)/2
    V.d_v[i * 3] = da * mx; V.d_v[i * 3 + 1] = da * my; V.d_v[i * 3 + 2] = da * mz + dvz;
    V.d_m[i * 3] = dmx + da * vx + 0.5f * dnx;
    V.d_m[i * 3 + 1] = dmy + da * vy + 0.5f * dny;
    V.d_m[i * 3 + 2] = dmz + da * vz + 0.5f * dnz;
    V.d_c[i * 3] = dcx; V.d_c[i * 3 + 1] = dcy; V.d_c[i * 3 + 2] = dcz;
  }
}

// One warp per surfel; the grid covers a few thousand surfels per sweep (a detection has ~2 000) instead of the whole
// capacity, whose empty blocks used to be most of this kernel's time in large batches.
__global__ void __launch_bounds__(256) splat_backward_kernel(const SplatView* __restrict__ views) {
  const SplatView& V = views[blockIdx.y];
  const int m = min(V.count ? *V.count : V.static_count, V.capacity);
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < m; i += gridDim.x * (blockDim.x >> 5))
    splat_backward_surfel(V, i, lane);
}

}  // namespace

int launch_project(const SplatView* views_dev, int batch, int max_count, cudaStream_t s) {
  if (max_count <= 0 || batch <= 0) return SDFR_OK;
  dim3 grid((max_count + 255) / 256, batch);
  project_kernel<<<grid, 256, 0, s>>>(views_dev);
  SDFR_LAUNCH_CHECK();
  return SDFR_OK;
}

int launch_splat_forward(const SplatView* views_dev, int batch, int max_w, int max_h, int any_fine, cudaStream_t s) {
  if (max_w <= 0 || max_h <= 0 || batch <= 0) return SDFR_OK;
  // any_fine: some detection of the launch is a small crop (4 x 4 tiles); the grid then covers the finer tiling and the
  // blocks of the other detections beyond their 8 x 8 tiling exit at once
  const int tile = any_fine ? 4 : 8;
  static bool attr_set = false;
  if (!attr_set) {
    SDFR_CUDA(cudaFuncSetAttribute(splat_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LCAP * 56));
    attr_set = true;
  }
  dim3 grid((max_w + tile - 1) / tile, (max_h + tile - 1) / tile, batch);
  splat_forward_kernel<<<grid, CHUNK, (size_t)LCAP * 56, s>>>(views_dev);
  SDFR_LAUNCH_CHECK();
  return SDFR_OK;
}

int launch_pixel_grad_prep(const SplatView* views_dev, int batch, int max_pixels, const float* g_color,
                           const float* g_mask, const float* g_depth, const float* g_nmap, cudaStream_t s) {
  if (max_pixels <= 0 || batch <= 0) return SDFR_OK;
  dim3 grid((max_pixels + 255) / 256, batch);
  pixel_grad_prep_kernel<<<grid, 256, 0, s>>>(views_dev, g_color, g_mask, g_depth, g_nmap);
  SDFR_LAUNCH_CHECK();
  return SDFR_OK;
}

int launch_splat_backward(const SplatView* views_dev, int batch, int max_count, cudaStream_t s) {
  if (max_count <= 0 || batch <= 0) return SDFR_OK;
  dim3 grid(std::min((max_count + 7) / 8, 512), batch);
  splat_backward_kernel<<<grid, 256, 0, s>>>(views_dev);
  SDFR_LAUNCH_CHECK();
  return SDFR_OK;
}

}  // namespace sdfr
