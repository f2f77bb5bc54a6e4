// Trace mode: per-ray sphere tracing against the DeepSDF decoder (the renderer BASELINE.json's
// north_star describes; the reference itself only has the surfel splat).
//
// Not a replacement of a reference function - a new renderer behind the same decoder, validated
// against its own torch restatement (oracle/trace_oracle.py), a closed-form field and, loosely, splat mode.
//
// Rays are "points" for the decoder kernels.  Two forms of the march, no host synchronisation in either:
//
//  * fused (tensor-core decoder with a CTA-pair pass table):
//      1. distance cache: the lattice-pass kernel (mlp_tc.cu, fp16 operands, 0.74 of the tensor roofline) evaluates a
//         regular 40^3 lattice over the box once (190 us); every ray then marches through its trilinear interpolant
//         minus a safety margin in registers (trace_grid_march_kernel) until it is within ~1.5 cells of the surface
//         or leaves the box.  Rays that never come near the surface cost no decoder evaluation at all.
//      2. speculative march (trace.cuh): ONE launch of the lattice-pass kernel per step; it generates its rows from
//         the active rays - 2^lk look-ahead samples per ray, as many as fit one round of the grid - and its epilogue
//         advances every ray over all samples whose unbounding spheres connect, drops it when it leaves the box, or -
//         once |sdf| < near_thr - hands it to the `near` list (warp-ballot compaction, one atomic per warp and list).
//         A launch costs one pass over the weights however few rays are left, so the long tail of grazing rays
//         (60 plain steps) collapses to ~10 launches.  Far from the surface the 3e-4 error of that pass only
//         perturbs the step length.
//      3. the near rays are finished at full precision by Newton steps along the ray, tau -= sdf / (grad sdf . d),
//         using the accurate forward + input-gradient kernel whose last evaluation also yields the normals and
//         d sdf / d latent: a hit is a root with |sdf| < eps.
//  * stepwise (any decoder, SDFR_MLP_FFMA): three launches per step at full precision - trace_points, the decoder
//    forward over `count` rows, trace_advance - and one gradient-carrying evaluation at the hits.
//
// The backward is implicit differentiation of f(l, o + tau d) = 0 at the hit: no storage
// of the march, one gradient evaluation per hit ray.
#include "trace.cuh"

#include <algorithm>


namespace sdfr {

namespace {

struct TraceWs {
  float* tau;        // [P] current ray parameter (distance along the unit ray, camera units)
  float* tau_exit;   // [P]
  int* list[2];      // [P] active ray ids (ping-pong)
  int* hits;         // [P] rows of the final evaluation: the hit rays (stepwise) / the near rays (fused)
  int* counters;     // [16] march counters (3, rotating), near rays [3], hits [4], Newton work lists [5..8]
  unsigned char* hit_flag;   // [P] per row of the final evaluation: 1 = hit
  float* inputs;     // [P, in0] decoder inputs of the rows being evaluated
  float* sdf;        // [P]
  float* dinput;     // [P, in0] for the rows of the final evaluation
  RayMarch* march;   // [6] per-step descriptors of the fused march
  float* fh;         // [P] fused march: predicted sdf at tau, slope estimate, last evaluated sample (trace.cuh)
  float* mh;
  float* ls;
  float* lf;
  int* nsteps;       // [P] fused march: connected samples consumed per ray
  float* cache;      // [GRID_D^3] distance cache: coarse sdf on the regular lattice over the box
  float* z_ref;      // [CACHE_LATENT] unit latent the cache was evaluated at
  int* cstate;       // [4] cache state: valid, lattice rows this call evaluates (0 = cache reused), extra margin (float)
  unsigned long long* mask_scratch;   // ReLU sign scratch of the full-precision decoder launches of THIS workspace (views
                                      // in flight on different streams must not share the decoder's own)
  float* esdf;       // [P] fused finish: sdf / input gradient of the rows evaluated in this Newton round (compact);
  float* edinput;    // [P, in0]  sdf / dinput above hold the LAST evaluation of every near ray, by near index
};

constexpr int GRID_D = 40;            // nodes per axis of the distance cache (spacing 0.052 over [-1, 1.025])
constexpr float GRID_MARGIN = 0.04f;  // subtracted from the interpolated distance: trilinear error of a 1-Lipschitz
                                      // field is below the distance to the nearest node (<= 0.045 at a cell centre);
                                      // measured 0.02 on the stock prior
constexpr float GRID_STOP = 0.05f;    // hand-over to the decoder march below this interpolated distance (0.08: 17 march
                                      // launches at 256^2, 0.05: 15; the last steps of the grid march are 0.01 long)
constexpr int GRID_ITERS = 192;       // step cap of the grid march
constexpr int CACHE_LATENT = 512;     // floats reserved for the cache's reference latent
constexpr float CACHE_SLACK = 0.01f;  // the cache is reused while lipschitz * |z - z_ref| stays below this (it widens the margin)
constexpr size_t CACHE_BYTES = (((size_t)GRID_D * GRID_D * GRID_D * 4 + 255) & ~(size_t)255) + CACHE_LATENT * 4 + 256;

// Distance cache across calls (caller-owned block, sdfr_trace_cache_bytes): the cache only depends on the latent, and the
// decoder is Lipschitz in it (sdfr_refine_cfg.latent_lipschitz):
//   sdf(x; z) >= sdf(x; z_ref) - lip |z - z_ref| ,
// so the lattice values at z_ref stay a conservative distance bound for z once lip |z - z_ref| is added to the margin.
// One warp decides on the device: reuse (no lattice rows this call) or renew (whole lattice, new reference).
__global__ void trace_cache_check_kernel(const float* __restrict__ z, int L, float lip, int n_rows, float* __restrict__ z_ref,
                                         int* __restrict__ cstate) {
  const int lane = threadIdx.x;
  float acc = 0.f;
  for (int c = lane; c < L; c += 32) { const float dz = z[c] - z_ref[c]; acc += dz * dz; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  const float extra = lip * sqrtf(acc);
  const bool reuse = cstate[0] == 1 && lip > 0.f && extra <= CACHE_SLACK;   // also false for a NaN latent
  __syncwarp();
  if (!reuse) for (int c = lane; c < L; c += 32) z_ref[c] = z[c];
  if (lane == 0) {
    cstate[0] = 1;
    cstate[1] = reuse ? 0 : n_rows;
    cstate[2] = __float_as_int(reuse ? extra : 0.f);
  }
}

__global__ void __launch_bounds__(256) trace_init_kernel(TraceParams p, TraceWs w) {
  const int P = p.width * p.height;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  bool active = false;
  if (j < P) {
    float o[3], d[3], rn[3];
    ray_of_pixel(p, j, o, d, rn);
    float t0 = 0.f, t1 = 1e30f;
    for (int a = 0; a < 3; ++a) {
      const float inv = 1.f / d[a];
      float ta = (p.lo - o[a]) * inv, tb = (p.hi - o[a]) * inv;
      if (ta > tb) { const float s = ta; ta = tb; tb = s; }
      if (d[a] == 0.f) {   // parallel to the slab
        if (o[a] < p.lo || o[a] > p.hi) { t0 = 1.f; t1 = 0.f; }
        continue;
      }
      t0 = fmaxf(t0, ta);
      t1 = fminf(t1, tb);
    }
    active = t0 <= t1;
    w.tau[j] = t0;
    w.tau_exit[j] = t1;
  }
  ray_append(active, j, w.list[0], w.counters + 0);
}

// Fused form of the start: box entry / exit of every pixel ray, then the march through the distance cache
// (trilinear, GRID_MARGIN subtracted) up to the hand-over distance.  Rays that get there join the active list with a
// predicted distance and slope (the interpolant and its directional derivative) for the speculative march.
__global__ void __launch_bounds__(256) trace_grid_march_kernel(TraceParams p, TraceWs w, float stop) {
  const int P = p.width * p.height;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  bool active = false;
  if (j < P) {
    float o[3], d[3], rn[3];
    ray_of_pixel(p, j, o, d, rn);
    float t0 = 0.f, t1 = 1e30f;
    for (int a = 0; a < 3; ++a) {
      const float inv = 1.f / d[a];
      float ta = (p.lo - o[a]) * inv, tb = (p.hi - o[a]) * inv;
      if (ta > tb) { const float s = ta; ta = tb; tb = s; }
      if (d[a] == 0.f) {
        if (o[a] < p.lo || o[a] > p.hi) { t0 = 1.f; t1 = 0.f; }
        continue;
      }
      t0 = fmaxf(t0, ta);
      t1 = fminf(t1, tb);
    }
    float tau = t0, fv = 0.f, mv = -1.f;
    if (t0 <= t1) {
      const float extra = __int_as_float(w.cstate[2]);     // lip |z - z_ref| of a reused cache
      const float grid_margin = GRID_MARGIN + extra, grid_stop = stop + extra;
      const float inv_h = (float)(GRID_D - 1) / (p.hi - p.lo);
      const float* __restrict__ G = w.cache;
      // steps are >= stop - margin = 0.01, the box diagonal is 3.5: a ray that has not arrived after GRID_ITERS steps
      // (it runs along a surface at the hand-over distance) continues in the decoder march from where it is
      for (int it = 0; it <= GRID_ITERS; ++it) {
        float u[3], fr[3];
        int i0[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          u[a] = fminf(fmaxf((o[a] + tau * d[a] - p.lo) * inv_h, 0.f), (float)(GRID_D - 1) - 1e-3f);
          i0[a] = (int)u[a];
          fr[a] = u[a] - (float)i0[a];
        }
        const float* c = G + ((size_t)i0[0] * GRID_D + i0[1]) * GRID_D + i0[2];
        const float c000 = c[0], c001 = c[1], c010 = c[GRID_D], c011 = c[GRID_D + 1];
        const float c100 = c[GRID_D * GRID_D], c101 = c[GRID_D * GRID_D + 1], c110 = c[GRID_D * GRID_D + GRID_D],
                    c111 = c[GRID_D * GRID_D + GRID_D + 1];
        const float x00 = c000 + fr[2] * (c001 - c000), x01 = c010 + fr[2] * (c011 - c010);
        const float x10 = c100 + fr[2] * (c101 - c100), x11 = c110 + fr[2] * (c111 - c110);
        const float y0 = x00 + fr[1] * (x01 - x00), y1 = x10 + fr[1] * (x11 - x10);
        const float f = y0 + fr[0] * (y1 - y0);
        if (f < grid_stop || it == GRID_ITERS) {
          // directional derivative of the interpolant along the ray
          const float gx = (y1 - y0) * inv_h;
          const float gy = ((x01 - x00) + fr[0] * ((x11 - x10) - (x01 - x00))) * inv_h;
          const float z00 = c001 - c000, z01 = c011 - c010, z10 = c101 - c100, z11 = c111 - c110;
          const float zy0 = z00 + fr[1] * (z01 - z00), zy1 = z10 + fr[1] * (z11 - z10);
          const float gz = (zy0 + fr[0] * (zy1 - zy0)) * inv_h;
          fv = fmaxf(f - extra, 0.25f * stop);
          mv = fminf(fmaxf(gx * d[0] + gy * d[1] + gz * d[2], -1.f), 0.f);
          active = true;
          break;
        }
        tau += f - grid_margin;
        if (tau > t1) break;
      }
    }
#ifdef SDFR_TRACE_DEBUG
    if (j == SDFR_TRACE_DEBUG) printf("grid ray %d: t0 %.5f t1 %.5f tau %.5f fv %.5f mv %.3f active %d\n", j, t0, t1, tau, fv, mv, (int)active);
#endif
    w.tau[j] = tau;
    w.tau_exit[j] = t1;
    w.fh[j] = fv;
    w.mh[j] = mv;
    w.ls[j] = -1e30f;
    w.lf[j] = 0.f;
    w.nsteps[j] = 0;
  }
  ray_append(active, j, w.list[0], w.counters + 0);
}

// decoder inputs [latent_unit, o + tau d] for the rows of `list` (also resets the next list's counter); with `via`,
// row i is the ray list[via[i]] (the work list of a Newton round indexes the near list)
__global__ void __launch_bounds__(256) trace_points_kernel(TraceParams p, TraceWs w, const float* __restrict__ latent_unit,
                                                           const int* __restrict__ list, const int* __restrict__ via,
                                                           const int* __restrict__ count, int* __restrict__ reset_counter) {
  if (blockIdx.x == 0 && threadIdx.x == 0 && reset_counter) *reset_counter = 0;
  const int n = *count;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = list[via ? via[i] : i];
  float o[3], d[3], rn[3];
  ray_of_pixel(p, j, o, d, rn);
  const float tau = w.tau[j];
  float* row = w.inputs + (size_t)i * p.in0;
  for (int c = 0; c < p.latent; ++c) row[c] = latent_unit[c];
  for (int a = 0; a < 3; ++a) row[p.latent + a] = o[a] + tau * d[a];
}

__global__ void __launch_bounds__(256) trace_advance_kernel(TraceParams p, TraceWs w, const int* __restrict__ list,
                                                            const int* __restrict__ count, int* __restrict__ next_list,
                                                            int* __restrict__ next_count, int last_step) {
  const int n = *count;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool keep = false, hit = false;
  int j = 0;
  if (i < n) {
    j = list[i];
    const float f = w.sdf[i];
    if (fabsf(f) < p.eps) {
      hit = true;                       // tau stays at the evaluated point
    } else {
      const float tau = w.tau[j] + f;
      w.tau[j] = tau;
      keep = tau <= w.tau_exit[j] && tau >= 0.f && !last_step;
    }
  }
  ray_append(keep, j, next_list, next_count);
  ray_append(hit, j, w.hits, w.counters + 3);
}

// One Newton round of the near rays along their ray, from the accurate evaluation (sdf, grad sdf) at o + tau d of the
// rows of the work list (`work` = indices into the near list, null = all of it).  A row whose |sdf| < stay_thr is a
// root: its evaluation is final (sdf / dinput by near index, hit flag) and it leaves the work list; the others step,
// tau -= sdf / (grad sdf . d), and are evaluated again.  The last round classifies what is left: a hit is a root with
// |sdf| < eps inside the box.
__global__ void __launch_bounds__(256) trace_newton_kernel(TraceParams p, TraceWs w, const int* __restrict__ work,
                                                           const int* __restrict__ work_count, int* __restrict__ next_work,
                                                           int* __restrict__ next_count, float max_step, float stay_thr,
                                                           int last) {
  const int n = *work_count;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  bool hit = false, again = false;
  int i = 0;
  if (r < n) {
    i = work ? work[r] : r;
    const int j = w.hits[i];
    const float f = w.esdf[r];
    const float tau = w.tau[j];
    const float* G = w.edinput + (size_t)r * p.in0;
#ifdef SDFR_TRACE_DEBUG
    if (j == SDFR_TRACE_DEBUG) printf("newton ray %d row %d near-index %d: tau %.6f f %.3e G %.4f %.4f %.4f last %d\n", j, r, i, tau, f, G[p.latent], G[p.latent + 1], G[p.latent + 2], last);
#endif
    if (last || fabsf(f) < stay_thr) {
      hit = fabsf(f) < p.eps && tau >= 0.f && tau <= w.tau_exit[j];
      w.hit_flag[i] = hit ? 1 : 0;
      w.sdf[i] = f;
      for (int c = 0; c < p.in0; ++c) w.dinput[(size_t)i * p.in0 + c] = G[c];
    } else {
      float o[3], d[3], rn[3];
      ray_of_pixel(p, j, o, d, rn);
      const float Gd = G[p.latent] * d[0] + G[p.latent + 1] * d[1] + G[p.latent + 2] * d[2];
      // f(tau + s) ~ f + s (G . d); a ray that runs (nearly) along the level set falls back to the sphere-tracing step
      float step = fabsf(Gd) > 1e-3f ? -f / Gd : f;
      step = fminf(fmaxf(step, -max_step), max_step);
      if (step == step) w.tau[j] = tau + step;
      again = true;
    }
  }
  ray_append(again, i, next_work, next_count);
  const unsigned ballot = __ballot_sync(0xffffffffu, hit);
  if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(w.counters + 4, __popc(ballot));
}

// maps of the hit rays from the gradient-carrying evaluation at the hit points
__global__ void __launch_bounds__(256) trace_finalize_kernel(TraceParams p, TraceWs w, float* __restrict__ depth,
                                                             float* __restrict__ nmap, float* __restrict__ nocs,
                                                             float* __restrict__ mask) {
  const int n = w.counters[3];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !w.hit_flag[i]) return;
  const int P = p.width * p.height;
  const int j = w.hits[i];
  float o[3], d[3], rn[3];
  ray_of_pixel(p, j, o, d, rn);
  const float tau = w.tau[j];
  const float* g = w.dinput + (size_t)i * p.in0 + p.latent;
  const float inv = rsqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
  const float n_obj[3] = {g[0] * inv, g[1] * inv, g[2] * inv};
  const float x[3] = {o[0] + tau * d[0], o[1] + tau * d[1], o[2] + tau * d[2]};
  if (depth) depth[j] = tau * rn[2];
  if (mask) mask[j] = 1.f;
  if (nmap)
    for (int a = 0; a < 3; ++a)
      nmap[a * P + j] = ((p.R[a * 3] * n_obj[0] + p.R[a * 3 + 1] * n_obj[1] + p.R[a * 3 + 2] * n_obj[2]) + 1.f) / 2.f;
  if (nocs) {   // same convention as the splat's dcm path: ((-x, y, z) + 1) / 2
    nocs[j] = (-x[0] + 1.f) / 2.f;
    nocs[P + j] = (x[1] + 1.f) / 2.f;
    nocs[2 * P + j] = (x[2] + 1.f) / 2.f;
  }
}

// implicit differentiation at the hit:  (g_x - lambda G) . d x|tau / d theta  -  lambda g_l . d l / d theta,
// lambda = (g_depth r_z + g_x . d) / (G . d)
__global__ void __launch_bounds__(256) trace_backward_kernel(TraceParams p, TraceWs w, const float* __restrict__ g_depth,
                                                             const float* __restrict__ g_nocs, float* __restrict__ d_pose,
                                                             float* __restrict__ d_latent) {
  __shared__ float s_red[8];
  const int n = w.counters[3];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int P = p.width * p.height;
  const bool live = i < n && w.hit_flag[i];
  float acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = 0.f;
  float lam = 0.f;
  if (live) {
    const int j = w.hits[i];
    float o[3], d[3], rn[3];
    ray_of_pixel(p, j, o, d, rn);
    const float tau = w.tau[j];
    const float* G = w.dinput + (size_t)i * p.in0 + p.latent;
    const float gx[3] = {g_nocs ? -0.5f * g_nocs[j] : 0.f, g_nocs ? 0.5f * g_nocs[P + j] : 0.f,
                         g_nocs ? 0.5f * g_nocs[2 * P + j] : 0.f};
    float Gd = G[0] * d[0] + G[1] * d[1] + G[2] * d[2];
    if (fabsf(Gd) < 1e-6f) Gd = Gd < 0.f ? -1e-6f : 1e-6f;
    lam = ((g_depth ? g_depth[j] * rn[2] : 0.f) + gx[0] * d[0] + gx[1] * d[1] + gx[2] * d[2]) / Gd;
    const float q[3] = {gx[0] - lam * G[0], gx[1] - lam * G[1], gx[2] - lam * G[2]};
    // x = R^T c with c = tau r_hat - t:  dL/dR[b][a] = c_b q_a ; dL/dt = -R q
    const float c[3] = {tau * rn[0] - p.t[0], tau * rn[1] - p.t[1], tau * rn[2] - p.t[2]};
    for (int b = 0; b < 3; ++b) {
      for (int a = 0; a < 3; ++a) acc[b * 4 + a] = c[b] * q[a];
      acc[b * 4 + 3] = -(p.R[b * 3] * q[0] + p.R[b * 3 + 1] * q[1] + p.R[b * 3 + 2] * q[2]);
    }
  }
  // block reductions (12 pose entries, then the latent entries), one atomic per block each
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = 0; k < 12 + p.latent; ++k) {
    float v;
    if (k < 12) v = acc[k];
    else v = live ? -lam * w.dinput[(size_t)i * p.in0 + (k - 12)] : 0.f;
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o2);
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int ww = 0; ww < 8; ++ww) t += s_red[ww];
      if (k < 12) { if (d_pose) atomicAdd(d_pose + k, t); }
      else if (d_latent) atomicAdd(d_latent + (k - 12), t);
    }
  }
}

int g_trace_stats = 0;                  // row counting of the fused march (sdfr_trace_set_stats)
long long g_trace_counts[4] = {0, 0, 0, 0};   // distance-cache rows, march rows, non-empty march launches, Newton rows

TraceWs carve(void* ws, int64_t P, int in0, size_t mask_bytes) {
  TraceWs w;
  char* p = reinterpret_cast<char*>(ws);
  auto take = [&](size_t bytes) { char* r = p; p += (bytes + 255) & ~(size_t)255; return r; };
  w.counters = reinterpret_cast<int*>(take(64));
  w.march = reinterpret_cast<RayMarch*>(take(6 * sizeof(RayMarch)));
  w.tau = reinterpret_cast<float*>(take((size_t)P * 4));
  w.tau_exit = reinterpret_cast<float*>(take((size_t)P * 4));
  w.list[0] = reinterpret_cast<int*>(take((size_t)P * 4));
  w.list[1] = reinterpret_cast<int*>(take((size_t)P * 4));
  w.hits = reinterpret_cast<int*>(take((size_t)P * 4));
  w.hit_flag = reinterpret_cast<unsigned char*>(take((size_t)P));
  w.inputs = reinterpret_cast<float*>(take((size_t)P * in0 * 4));
  w.sdf = reinterpret_cast<float*>(take((size_t)P * 4));
  w.dinput = reinterpret_cast<float*>(take((size_t)P * in0 * 4));
  w.fh = reinterpret_cast<float*>(take((size_t)P * 4));
  w.mh = reinterpret_cast<float*>(take((size_t)P * 4));
  w.ls = reinterpret_cast<float*>(take((size_t)P * 4));
  w.lf = reinterpret_cast<float*>(take((size_t)P * 4));
  w.nsteps = reinterpret_cast<int*>(take((size_t)P * 4));
  w.cache = reinterpret_cast<float*>(take((size_t)GRID_D * GRID_D * GRID_D * 4));
  w.z_ref = reinterpret_cast<float*>(take((size_t)CACHE_LATENT * 4));
  w.cstate = reinterpret_cast<int*>(take(256));
  w.mask_scratch = mask_bytes ? reinterpret_cast<unsigned long long*>(take(mask_bytes)) : nullptr;
  w.esdf = reinterpret_cast<float*>(take((size_t)P * 4));
  w.edinput = reinterpret_cast<float*>(take((size_t)P * in0 * 4));
  return w;
}

size_t ws_bytes(int64_t P, int in0, size_t mask_bytes) {
  auto r = [](size_t b) { return (b + 255) & ~(size_t)255; };
  return r(64) + r(6 * sizeof(RayMarch)) + 12 * r((size_t)P * 4) + r((size_t)P) + 3 * r((size_t)P * in0 * 4) +
         r((size_t)GRID_D * GRID_D * GRID_D * 4) + r((size_t)CACHE_LATENT * 4) + 256 + r(mask_bytes);
}

// reuse the lattice values of an earlier call while the latent has not moved beyond the slack (device-side decision),
// evaluate the lattice otherwise
int update_distance_cache(const sdfr_decoder* dec, const float* latent_unit_dev, float lip, float hi, float* cache,
                          float* z_ref, int* cstate, cudaStream_t s) {
  MlpInputs ig;
  ig.inputs = nullptr; ig.latent_unit = latent_unit_dev; ig.lattice = make_regular_lattice(GRID_D, (double)hi);
  ig.points_per_batch = (long long)GRID_D * GRID_D * GRID_D; ig.n = ig.points_per_batch;
  ig.index = nullptr; ig.small_tiles = 0;
  trace_cache_check_kernel<<<1, 32, 0, s>>>(latent_unit_dev, dec->dev.latent_size, lip, (int)ig.n, z_ref, cstate);
  SDFR_LAUNCH_CHECK();
  ig.count_dev = cstate + 1;
  return launch_mlp_tc_coarse(dec, ig, cache, s);
}

constexpr float BOX_LO = -1.0f, BOX_HI = 1.025f;   // Grid3D(40) extent the priors are sampled on (grid.py:38)

int fill_params(const sdfr_raster_cfg* cfg, const sdfr_decoder* dec, const float* pose_host, float eps, TraceParams* tp) {
  tp->width = cfg->width; tp->height = cfg->height;
  tp->in0 = dec->dev.in0; tp->latent = dec->dev.latent_size;
  for (int i = 0; i < 9; ++i) tp->kinv[i] = cfg->kinv[i];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) tp->R[r * 3 + c] = pose_host[r * 4 + c];
    tp->t[r] = pose_host[r * 4 + 3];
  }
  tp->eps = eps;
  tp->lo = BOX_LO; tp->hi = BOX_HI;
  return SDFR_OK;
}

}  // namespace

}  // namespace sdfr

using namespace sdfr;

extern "C" int64_t sdfr_trace_workspace_bytes(const sdfr_raster_cfg* cfg, const sdfr_decoder* dec) {
  if (!cfg || !dec) return -1;
  return (int64_t)ws_bytes((int64_t)cfg->width * cfg->height, dec->dev.in0, mlp_tc_mask_scratch_bytes(dec));
}

extern "C" int64_t sdfr_trace_cache_bytes(void) { return (int64_t)CACHE_BYTES; }

extern "C" int sdfr_trace_cache_update(sdfr_decoder* dec, const float* latent_unit_dev, void* cache_dev,
                                       float latent_lipschitz, void* stream) {
  SDFR_REQUIRE(dec && latent_unit_dev && cache_dev, SDFR_E_INVALID, "null argument");
  SDFR_REQUIRE(dec->tc.ok && mlp_tc_march_ok(dec), SDFR_E_UNSUPPORTED, "the distance cache belongs to the fused march");
  SDFR_REQUIRE(dec->dev.latent_size <= CACHE_LATENT, SDFR_E_UNSUPPORTED, "latent size %d > %d", dec->dev.latent_size,
               CACHE_LATENT);
  char* c = reinterpret_cast<char*>(cache_dev);
  float* z_ref = reinterpret_cast<float*>(c + (((size_t)GRID_D * GRID_D * GRID_D * 4 + 255) & ~(size_t)255));
  int* cstate = reinterpret_cast<int*>(z_ref + CACHE_LATENT);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int rc = update_distance_cache(dec, latent_unit_dev, latent_lipschitz > 0.f ? latent_lipschitz : 0.f, BOX_HI,
                                 reinterpret_cast<float*>(c), z_ref, cstate, s);
  if (!rc && g_trace_stats > 0) {
    int st[4];
    cudaStreamSynchronize(s);
    cudaMemcpy(st, cstate, sizeof(st), cudaMemcpyDeviceToHost);
    g_trace_counts[0] += st[1];
  }
  return rc;
}

extern "C" void sdfr_trace_set_stats(int on) {
  g_trace_stats = on;
  for (long long& c : g_trace_counts) c = 0;
}
extern "C" void sdfr_trace_get_stats(int64_t* counts4) {
  for (int i = 0; i < 4; ++i) counts4[i] = g_trace_counts[i];
}

extern "C" int sdfr_trace_forward(sdfr_decoder* dec, const sdfr_raster_cfg* cfg, const float* latent_unit_dev,
                                  const float* pose_host, int max_steps, float eps, float* depth_dev, float* nmap_dev,
                                  float* nocs_dev, float* mask_dev, int32_t* hit_count_dev, void* workspace_dev,
                                  void* cache_dev, float latent_lipschitz, int impl, void* stream) {
  SDFR_REQUIRE(dec && cfg && latent_unit_dev && pose_host && workspace_dev, SDFR_E_INVALID, "null argument");
  SDFR_REQUIRE(cfg->width > 0 && cfg->height > 0 && max_steps > 0 && eps > 0.f, SDFR_E_INVALID, "bad trace configuration");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t P = (int64_t)cfg->width * cfg->height;
  const int in0 = dec->dev.in0;
  TraceParams tp;
  fill_params(cfg, dec, pose_host, eps, &tp);
  TraceWs w = carve(workspace_dev, P, in0, mlp_tc_mask_scratch_bytes(dec));
  SDFR_REQUIRE(dec->dev.latent_size <= CACHE_LATENT, SDFR_E_UNSUPPORTED, "latent size %d > %d", dec->dev.latent_size,
               CACHE_LATENT);
  if (cache_dev) {      // the caller's persistent block replaces the per-call cache of the workspace
    char* c = reinterpret_cast<char*>(cache_dev);
    w.cache = reinterpret_cast<float*>(c);
    w.z_ref = reinterpret_cast<float*>(c + (((size_t)GRID_D * GRID_D * GRID_D * 4 + 255) & ~(size_t)255));
    w.cstate = reinterpret_cast<int*>(w.z_ref + CACHE_LATENT);
  } else {
    SDFR_CUDA(cudaMemsetAsync(w.cstate, 0, 16, s));
  }
  if (impl == SDFR_MLP_AUTO) impl = dec->tc.ok ? SDFR_MLP_TCGEN05 : SDFR_MLP_FFMA;
  SDFR_CUDA(cudaMemsetAsync(w.counters, 0, 64, s));
  if (depth_dev) SDFR_CUDA(cudaMemsetAsync(depth_dev, 0, (size_t)P * 4, s));
  if (mask_dev) SDFR_CUDA(cudaMemsetAsync(mask_dev, 0, (size_t)P * 4, s));
  if (nmap_dev) SDFR_CUDA(cudaMemsetAsync(nmap_dev, 0, (size_t)P * 12, s));
  if (nocs_dev) SDFR_CUDA(cudaMemsetAsync(nocs_dev, 0, (size_t)P * 12, s));
  const unsigned blocks = (unsigned)((P + 255) / 256);
  MlpInputs in;
  in.inputs = w.inputs; in.latent_unit = nullptr; in.lattice = make_lattice(2); in.points_per_batch = 1; in.n = P;
  in.index = nullptr; in.small_tiles = 0;
  in.mask_scratch = w.mask_scratch;
  int rc;
  const bool fused = impl == SDFR_MLP_TCGEN05 && mlp_tc_march_ok(dec);
  if (fused) {
    // ---- distance cache and the march through it ----
    // (a negative latent_lipschitz: the caller has brought the block up to date for this latent with
    //  sdfr_trace_cache_update - renders of one latent that run concurrently share one cache)
    if (!(cache_dev && latent_lipschitz < 0.f) &&
        (rc = update_distance_cache(dec, latent_unit_dev, cache_dev ? latent_lipschitz : 0.f, tp.hi, w.cache, w.z_ref,
                                    w.cstate, s))) return rc;
    trace_grid_march_kernel<<<blocks, 256, 0, s>>>(tp, w, GRID_STOP);
    SDFR_LAUNCH_CHECK();
    // ---- speculative march: one launch per step; lists ping-pong, counters rotate (read / append / clear) ----
    // The lattice pass is good to ~3e-4 (its error near the surface is what the refine engine measures and bounds by
    // 2.5e-3), so rays are handed to the full-precision finish well before that matters.
    // (a wider hand-over - 5e-3 / 0.02 - saves two march launches at 256^2 but the second Newton round then outgrows the
    //  16-point tiling of the band kernel: measured, no gain)
    const float near_thr = 2.5e-3f, near_reach = 0.01f;
    const int newton = 3;
    const int round_rows = mlp_tc_round_rows(dec);
    RayMarch desc[6];
    for (int k = 0; k < 6; ++k) {
      RayMarch& m = desc[k];
      m.p = tp; m.near_thr = near_thr; m.near_reach = near_reach; m.spacing = 1.2f; m.latent_unit = latent_unit_dev; m.tau = w.tau; m.tau_exit = w.tau_exit;
      m.fh = w.fh; m.mh = w.mh; m.ls = w.ls; m.lf = w.lf; m.nsteps = w.nsteps; m.step_budget = 4 * max_steps;
      m.list = w.list[k & 1]; m.count = w.counters + (k % 3);
      m.next_list = w.list[(k + 1) & 1]; m.next_count = w.counters + ((k + 1) % 3);
      m.near_list = w.hits; m.near_count = w.counters + 3;
      m.reset_count = w.counters + ((k + 2) % 3);
      m.round_rows = round_rows; m.max_log2k = 5;
    }
    SDFR_CUDA(cudaMemcpyAsync(w.march, desc, sizeof(desc), cudaMemcpyHostToDevice, s));
    MlpInputs im = in;
    im.inputs = nullptr;
    im.n = std::max<long long>(P, round_rows);
    // A launch advances every ray by at least one plain sphere-tracing step and a grazing ray by up to 32; a ray is
    // abandoned (a miss, the plain march's max_steps rule) once it has consumed 4 max_steps connected samples -
    // without that, the few rays that run parallel to a flat side a millimetre away keep every launch busy.
    // Launches that find their list empty return at once (3.5 us).  While more rays are active than fit half a round
    // every launch is a plain step for all of them, and the count falls by ~7 % per launch there: a 1024^2 image hands
    // 160 k rays to the march and still had 10 k active after 32 launches (they were dropped: 0.4 % of the hits
    // missing), so large images get three more launches per doubling of the pixel count, up to the max_steps of the
    // plain march.
    int extra_launches = 0;
    for (int64_t q = 65536; q < P; q *= 2) extra_launches += 3;
    const int launches = std::max(1, std::min(max_steps, std::max(8, max_steps / 2) + extra_launches));
    // sdfr_trace_set_stats(1): count the rows of every launch (synchronises after each: a measurement aid for
    // bench.py's utilisation figure, never on in a timed call; 2 also prints the per-launch counts)
    const int stats = g_trace_stats;
    if (stats && !(cache_dev && latent_lipschitz < 0.f)) {
      int c[4];
      cudaStreamSynchronize(s);
      cudaMemcpy(c, w.cstate, sizeof(c), cudaMemcpyDeviceToHost);
      g_trace_counts[0] += c[1];
    }
    for (int step = 0; step < launches; ++step) {
      im.march = w.march + (step % 6);
      im.count_dev = w.counters + (step % 3);
      if (stats) {
        int c[16];
        cudaStreamSynchronize(s);
        cudaMemcpy(c, w.counters, sizeof(c), cudaMemcpyDeviceToHost);
        if (stats > 1) fprintf(stderr, "trace march launch %d: active %d near %d\n", step, c[step % 3], c[3]);
        const int n = c[step % 3];
        if (n > 0) {
          int lk = 0;
          while (lk < 5 && ((long long)n << (lk + 1)) <= (long long)round_rows) ++lk;
          g_trace_counts[1] += (long long)n << lk;
          g_trace_counts[2] += 1;
        }
      }
      if ((rc = launch_mlp_tc_coarse(dec, im, nullptr, s))) return rc;
    }
    // ---- finish at full precision: Newton rounds along the ray over a shrinking work list; the evaluation that finds
    //      |sdf| < eps / 2 (or the last one) classifies the ray and feeds the maps ----
    MlpInputs ia = in;
    ia.inputs = w.inputs;
    ia.adaptive_tiles = 1;
    for (int it = 0; it <= newton; ++it) {
      const int* work = it == 0 ? nullptr : w.list[(it - 1) & 1];
      const int* work_count = it == 0 ? w.counters + 3 : w.counters + 4 + it;
      trace_points_kernel<<<blocks, 256, 0, s>>>(tp, w, latent_unit_dev, w.hits, work, work_count, nullptr);
      SDFR_LAUNCH_CHECK();
      ia.count_dev = work_count;
      if ((rc = launch_mlp_tc(dec, ia, w.esdf, w.edinput, s))) return rc;
      trace_newton_kernel<<<blocks, 256, 0, s>>>(tp, w, work, work_count, w.list[it & 1], w.counters + 5 + it,
                                                 2.f * near_reach, 0.5f * eps, it == newton);
      SDFR_LAUNCH_CHECK();
      if (stats) {
        int c[16];
        cudaStreamSynchronize(s);
        cudaMemcpy(c, w.counters, sizeof(c), cudaMemcpyDeviceToHost);
        if (stats > 1) fprintf(stderr, "trace newton round %d: rows %d -> again %d, hits so far %d\n", it, it == 0 ? c[3] : c[4 + it], c[5 + it], c[4]);
        g_trace_counts[3] += it == 0 ? c[3] : c[4 + it];
      }
    }
  } else {
    trace_init_kernel<<<blocks, 256, 0, s>>>(tp, w);
    SDFR_LAUNCH_CHECK();
    for (int step = 0; step < max_steps; ++step) {
      const int cur = step & 1, nxt = cur ^ 1;
      trace_points_kernel<<<blocks, 256, 0, s>>>(tp, w, latent_unit_dev, w.list[cur], nullptr, w.counters + cur, w.counters + nxt);
      SDFR_LAUNCH_CHECK();
      in.count_dev = w.counters + cur;
      rc = impl == SDFR_MLP_TCGEN05 ? launch_mlp_tc(dec, in, w.sdf, nullptr, s) : launch_mlp_ffma(dec, in, w.sdf, nullptr, s);
      if (rc) return rc;
      trace_advance_kernel<<<blocks, 256, 0, s>>>(tp, w, w.list[cur], w.counters + cur, w.list[nxt], w.counters + nxt,
                                                 step == max_steps - 1);
      SDFR_LAUNCH_CHECK();
    }
    // gradient-carrying evaluation at the hit points: every listed ray is a hit
    SDFR_CUDA(cudaMemsetAsync(w.hit_flag, 1, (size_t)P, s));
    SDFR_CUDA(cudaMemcpyAsync(w.counters + 4, w.counters + 3, 4, cudaMemcpyDeviceToDevice, s));
    trace_points_kernel<<<blocks, 256, 0, s>>>(tp, w, latent_unit_dev, w.hits, nullptr, w.counters + 3, nullptr);
    SDFR_LAUNCH_CHECK();
    in.count_dev = w.counters + 3;
    rc = impl == SDFR_MLP_TCGEN05 ? launch_mlp_tc(dec, in, w.sdf, w.dinput, s) : launch_mlp_ffma(dec, in, w.sdf, w.dinput, s);
    if (rc) return rc;
  }
  trace_finalize_kernel<<<blocks, 256, 0, s>>>(tp, w, depth_dev, nmap_dev, nocs_dev, mask_dev);
  SDFR_LAUNCH_CHECK();
  if (hit_count_dev) SDFR_CUDA(cudaMemcpyAsync(hit_count_dev, w.counters + 4, 4, cudaMemcpyDeviceToDevice, s));
  return SDFR_OK;
}

extern "C" int sdfr_trace_backward(sdfr_decoder* dec, const sdfr_raster_cfg* cfg, const float* pose_host, float eps,
                                   const float* g_depth_dev, const float* g_nocs_dev, float* d_pose_dev,
                                   float* d_latent_unit_dev, void* workspace_dev, void* stream) {
  SDFR_REQUIRE(dec && cfg && pose_host && workspace_dev, SDFR_E_INVALID, "null argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t P = (int64_t)cfg->width * cfg->height;
  TraceParams tp;
  fill_params(cfg, dec, pose_host, eps, &tp);
  TraceWs w = carve(workspace_dev, P, dec->dev.in0, mlp_tc_mask_scratch_bytes(dec));
  if (d_pose_dev) SDFR_CUDA(cudaMemsetAsync(d_pose_dev, 0, 12 * sizeof(float), s));
  if (d_latent_unit_dev) SDFR_CUDA(cudaMemsetAsync(d_latent_unit_dev, 0, dec->dev.latent_size * sizeof(float), s));
  trace_backward_kernel<<<(unsigned)((P + 255) / 256), 256, 0, s>>>(tp, w, g_depth_dev, g_nocs_dev, d_pose_dev,
                                                                    d_latent_unit_dev);
  SDFR_LAUNCH_CHECK();
  return SDFR_OK;
}
