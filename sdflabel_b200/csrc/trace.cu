// Trace mode: per-ray sphere tracing against the DeepSDF decoder (the renderer BASELINE.json's
// north_star describes; the reference itself only has the surfel splat).
//
// Not a replacement of a reference function - a new renderer behind the same decoder, validated
// against its own torch restatement (oracle/trace_oracle.py), a closed-form field and, loosely, splat mode.
//
// Rays are "points" for the decoder kernels.  Two forms of the march, no host synchronisation in either:
//
//  * fused (tensor-core decoder with a CTA-pair pass table): ONE kernel per march step - the lattice-pass kernel
//    (mlp_tc.cu, fp16 operands, 0.74 of the tensor roofline) in march mode generates its rows from the active rays
//    (o + tau d), and its epilogue advances every ray (tau += sdf), drops it when it leaves the box, or - once
//    |sdf| < near_thr - hands it to the `near` list (warp-ballot compaction, one atomic per warp and list).  Far from
//    the surface the 3e-4 error of that pass only perturbs the step length.  The near rays are then finished at full
//    precision by Newton steps along the ray, tau -= sdf / (grad sdf . d), using the accurate forward + input-gradient
//    kernel whose last evaluation also yields the normals and d sdf / d latent: a hit is a root with |sdf| < eps.
//  * stepwise (any decoder, SDFR_MLP_FFMA): three launches per step at full precision - trace_points, the decoder
//    forward over `count` rows, trace_advance - and one gradient-carrying evaluation at the hits.
//
// The backward is implicit differentiation of f(l, o + tau d) = 0 at the hit: no storage
// of the march, one gradient evaluation per hit ray.
#include "trace.cuh"


namespace sdfr {

namespace {

struct TraceWs {
  float* tau;        // [P] current ray parameter (distance along the unit ray, camera units)
  float* tau_exit;   // [P]
  int* list[2];      // [P] active ray ids (ping-pong)
  int* hits;         // [P] rows of the final evaluation: the hit rays (stepwise) / the near rays (fused)
  int* counters;     // [8] march counters (3, rotating), rows of the final evaluation [3], hits [4]
  unsigned char* hit_flag;   // [P] per row of the final evaluation: 1 = hit
  float* inputs;     // [P, in0] decoder inputs of the rows being evaluated
  float* sdf;        // [P]
  float* dinput;     // [P, in0] for the rows of the final evaluation
  RayMarch* march;   // [6] per-step descriptors of the fused march
};

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

// decoder inputs [latent_unit, o + tau d] for the rows of `list` (also resets the next list's counter)
__global__ void __launch_bounds__(256) trace_points_kernel(TraceParams p, TraceWs w, const float* __restrict__ latent_unit,
                                                           const int* __restrict__ list, const int* __restrict__ count,
                                                           int* __restrict__ reset_counter) {
  if (blockIdx.x == 0 && threadIdx.x == 0 && reset_counter) *reset_counter = 0;
  const int n = *count;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = list[i];
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

// Newton step of the near rays along their ray, from the accurate evaluation (sdf, grad sdf) at o + tau d;
// the LAST evaluation classifies: a hit is a root with |sdf| < eps inside the box.
__global__ void __launch_bounds__(256) trace_newton_kernel(TraceParams p, TraceWs w, float max_step, int last) {
  const int n = w.counters[3];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool hit = false;
  if (i < n) {
    const int j = w.hits[i];
    const float f = w.sdf[i];
    const float tau = w.tau[j];
    if (last) {
      hit = fabsf(f) < p.eps && tau >= 0.f && tau <= w.tau_exit[j];
      w.hit_flag[i] = hit ? 1 : 0;
    } else if (fabsf(f) >= 0.25f * p.eps) {     // already well inside the stopping band: stay
      float o[3], d[3], rn[3];
      ray_of_pixel(p, j, o, d, rn);
      const float* G = w.dinput + (size_t)i * p.in0 + p.latent;
      const float Gd = G[0] * d[0] + G[1] * d[1] + G[2] * d[2];
      // f(tau + s) ~ f + s (G . d); a ray that runs (nearly) along the level set falls back to the sphere-tracing step
      float step = fabsf(Gd) > 1e-3f ? -f / Gd : f;
      step = fminf(fmaxf(step, -max_step), max_step);
      if (step == step) w.tau[j] = tau + step;
    }
  }
  if (last) {
    const unsigned ballot = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(w.counters + 4, __popc(ballot));
  }
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

TraceWs carve(void* ws, int64_t P, int in0) {
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
  return w;
}

size_t ws_bytes(int64_t P, int in0) {
  auto r = [](size_t b) { return (b + 255) & ~(size_t)255; };
  return r(64) + r(6 * sizeof(RayMarch)) + 6 * r((size_t)P * 4) + r((size_t)P) + 2 * r((size_t)P * in0 * 4);
}

int fill_params(const sdfr_raster_cfg* cfg, const sdfr_decoder* dec, const float* pose_host, float eps, TraceParams* tp) {
  tp->width = cfg->width; tp->height = cfg->height;
  tp->in0 = dec->dev.in0; tp->latent = dec->dev.latent_size;
  for (int i = 0; i < 9; ++i) tp->kinv[i] = cfg->kinv[i];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) tp->R[r * 3 + c] = pose_host[r * 4 + c];
    tp->t[r] = pose_host[r * 4 + 3];
  }
  tp->eps = eps;
  tp->lo = -1.0f; tp->hi = 1.025f;   // Grid3D(40) extent the priors are sampled on (grid.py:38)
  return SDFR_OK;
}

}  // namespace

}  // namespace sdfr

using namespace sdfr;

extern "C" int64_t sdfr_trace_workspace_bytes(const sdfr_raster_cfg* cfg, const sdfr_decoder* dec) {
  if (!cfg || !dec) return -1;
  return (int64_t)ws_bytes((int64_t)cfg->width * cfg->height, dec->dev.in0);
}

extern "C" int sdfr_trace_forward(sdfr_decoder* dec, const sdfr_raster_cfg* cfg, const float* latent_unit_dev,
                                  const float* pose_host, int max_steps, float eps, float* depth_dev, float* nmap_dev,
                                  float* nocs_dev, float* mask_dev, int32_t* hit_count_dev, void* workspace_dev, int impl,
                                  void* stream) {
  SDFR_REQUIRE(dec && cfg && latent_unit_dev && pose_host && workspace_dev, SDFR_E_INVALID, "null argument");
  SDFR_REQUIRE(cfg->width > 0 && cfg->height > 0 && max_steps > 0 && eps > 0.f, SDFR_E_INVALID, "bad trace configuration");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t P = (int64_t)cfg->width * cfg->height;
  const int in0 = dec->dev.in0;
  TraceParams tp;
  fill_params(cfg, dec, pose_host, eps, &tp);
  TraceWs w = carve(workspace_dev, P, in0);
  if (impl == SDFR_MLP_AUTO) impl = dec->tc.ok ? SDFR_MLP_TCGEN05 : SDFR_MLP_FFMA;
  SDFR_CUDA(cudaMemsetAsync(w.counters, 0, 64, s));
  if (depth_dev) SDFR_CUDA(cudaMemsetAsync(depth_dev, 0, (size_t)P * 4, s));
  if (mask_dev) SDFR_CUDA(cudaMemsetAsync(mask_dev, 0, (size_t)P * 4, s));
  if (nmap_dev) SDFR_CUDA(cudaMemsetAsync(nmap_dev, 0, (size_t)P * 12, s));
  if (nocs_dev) SDFR_CUDA(cudaMemsetAsync(nocs_dev, 0, (size_t)P * 12, s));
  const unsigned blocks = (unsigned)((P + 255) / 256);
  trace_init_kernel<<<blocks, 256, 0, s>>>(tp, w);
  SDFR_LAUNCH_CHECK();
  MlpInputs in;
  in.inputs = w.inputs; in.latent_unit = nullptr; in.lattice = make_lattice(2); in.points_per_batch = 1; in.n = P;
  in.index = nullptr; in.small_tiles = 0;
  int rc;
  const bool fused = impl == SDFR_MLP_TCGEN05 && mlp_tc_march_ok(dec);
  if (fused) {
    // ---- fused march: one launch per step; lists ping-pong, counters rotate (read / append / clear) ----
    // The lattice pass is good to ~3e-4 (its error near the surface is what the refine engine measures and bounds by
    // 2.5e-3), so rays are handed to the full-precision finish well before that matters.
    const float near_thr = 5e-3f;
    RayMarch desc[6];
    for (int k = 0; k < 6; ++k) {
      RayMarch& m = desc[k];
      m.p = tp; m.near_thr = near_thr; m.latent_unit = latent_unit_dev; m.tau = w.tau; m.tau_exit = w.tau_exit;
      m.list = w.list[k & 1]; m.count = w.counters + (k % 3);
      m.next_list = w.list[(k + 1) & 1]; m.next_count = w.counters + ((k + 1) % 3);
      m.near_list = w.hits; m.near_count = w.counters + 3;
      m.reset_count = w.counters + ((k + 2) % 3);
    }
    SDFR_CUDA(cudaMemcpyAsync(w.march, desc, sizeof(desc), cudaMemcpyHostToDevice, s));
    MlpInputs im = in;
    im.inputs = nullptr;
    for (int step = 0; step < max_steps; ++step) {
      im.march = w.march + (step % 6);
      im.count_dev = w.counters + (step % 3);
      if ((rc = launch_mlp_tc_coarse(dec, im, nullptr, s))) return rc;
    }
    // ---- finish at full precision: Newton steps along the ray, the last evaluation classifies and feeds the maps ----
    const int newton = 3;
    in.count_dev = w.counters + 3;
    for (int it = 0; it <= newton; ++it) {
      trace_points_kernel<<<blocks, 256, 0, s>>>(tp, w, latent_unit_dev, w.hits, w.counters + 3, nullptr);
      SDFR_LAUNCH_CHECK();
      if ((rc = launch_mlp_tc(dec, in, w.sdf, w.dinput, s))) return rc;
      trace_newton_kernel<<<blocks, 256, 0, s>>>(tp, w, 4.f * near_thr, it == newton);
      SDFR_LAUNCH_CHECK();
    }
  } else {
    for (int step = 0; step < max_steps; ++step) {
      const int cur = step & 1, nxt = cur ^ 1;
      trace_points_kernel<<<blocks, 256, 0, s>>>(tp, w, latent_unit_dev, w.list[cur], w.counters + cur, w.counters + nxt);
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
    trace_points_kernel<<<blocks, 256, 0, s>>>(tp, w, latent_unit_dev, w.hits, w.counters + 3, nullptr);
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
  TraceWs w = carve(workspace_dev, P, dec->dev.in0);
  if (d_pose_dev) SDFR_CUDA(cudaMemsetAsync(d_pose_dev, 0, 12 * sizeof(float), s));
  if (d_latent_unit_dev) SDFR_CUDA(cudaMemsetAsync(d_latent_unit_dev, 0, dec->dev.latent_size * sizeof(float), s));
  trace_backward_kernel<<<(unsigned)((P + 255) / 256), 256, 0, s>>>(tp, w, g_depth_dev, g_nocs_dev, d_pose_dev,
                                                                    d_latent_unit_dev);
  SDFR_LAUNCH_CHECK();
  return SDFR_OK;
}
