// Ray bookkeeping of trace mode, shared by the trace kernels (trace.cu) and the lattice-pass kernel of the decoder
// (mlp_tc.cu), which in "march mode" generates its rows from the active rays and advances them in its epilogue.
#pragma once

#include "common.cuh"

namespace sdfr {

struct TraceParams {
  int width, height, in0, latent;
  float kinv[9];
  float R[9], t[3];       // camera pose: v_cam = R x_obj + t (R orthogonal)
  float eps;
  float lo, hi;           // lattice box [-1, hi]^3 the prior was trained on
};

// One step of the fused march (device-resident; two copies alternate between steps).
// The kernel reads `list[0 .. *count)`, evaluates the decoder at o + tau d of every listed ray and then, per ray:
//   |sdf| < near_thr           -> the ray goes to the `near` list (it is finished by Newton steps at full precision)
//   tau + sdf outside [0, exit] -> the ray has left the box: dropped
//   otherwise                  -> tau += sdf and the ray goes to `next_list`
struct RayMarch {
  TraceParams p;
  float near_thr;
  const float* latent_unit;   // [latent]
  float* tau;                 // [P]
  const float* tau_exit;      // [P]
  const int* list;
  const int* count;
  int* next_list;
  int* next_count;
  int* near_list;
  int* near_count;
  int* reset_count;           // counter the step after the next appends to: cleared by this step
};

__device__ __forceinline__ void ray_of_pixel(const TraceParams& p, int j, float (&o)[3], float (&d)[3], float (&rn)[3]) {
  const int y = j / p.width, x = j - y * p.width;
  const float fx = (float)x, fy = (float)y;
  float r[3] = {p.kinv[0] * fx + p.kinv[1] * fy + p.kinv[2], p.kinv[3] * fx + p.kinv[4] * fy + p.kinv[5],
                p.kinv[6] * fx + p.kinv[7] * fy + p.kinv[8]};
  const float inv = rsqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  for (int a = 0; a < 3; ++a) rn[a] = r[a] * inv;
  for (int a = 0; a < 3; ++a) {
    // x_obj = R^T (v_cam - t)
    o[a] = -(p.R[0 * 3 + a] * p.t[0] + p.R[1 * 3 + a] * p.t[1] + p.R[2 * 3 + a] * p.t[2]);
    d[a] = p.R[0 * 3 + a] * rn[0] + p.R[1 * 3 + a] * rn[1] + p.R[2 * 3 + a] * rn[2];
  }
}

// warp-aggregated append: one atomic per warp (every lane of the warp must call it)
__device__ __forceinline__ void ray_append(bool pred, int value, int* list, int* counter) {
  const unsigned ballot = __ballot_sync(0xffffffffu, pred);
  if (!ballot) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == (__ffs(ballot) - 1)) base = atomicAdd(counter, __popc(ballot));
  base = __shfl_sync(0xffffffffu, base, __ffs(ballot) - 1);
  if (pred) list[base + __popc(ballot & ((1u << lane) - 1u))] = value;
}

// decoder input column c of march row `row` (latent, then the point o + tau d of the listed ray)
__device__ __forceinline__ float march_input(const RayMarch& m, long long row, int c) {
  if (c < m.p.latent) return m.latent_unit[c];
  const int j = m.list[row];
  float o[3], d[3], rn[3];
  ray_of_pixel(m.p, j, o, d, rn);
  const int a = c - m.p.latent;
  return o[a] + m.tau[j] * d[a];
}

// the advance of one evaluated march row (all 32 lanes of the warp call it; `valid` = the lane holds a row)
__device__ __forceinline__ void march_advance(const RayMarch& m, bool valid, long long row, float f) {
  bool keep = false, near = false;
  int j = 0;
  if (valid) {
    j = m.list[row];
    if (fabsf(f) < m.near_thr) {
      near = true;                      // tau stays at the evaluated point
    } else if (f == f) {                // a NaN sdf drops the ray
      const float tau = m.tau[j] + f;
      keep = tau <= m.tau_exit[j] && tau >= 0.f;
      if (keep) m.tau[j] = tau;
    }
  }
  ray_append(keep, j, m.next_list, m.next_count);
  ray_append(near, j, m.near_list, m.near_count);
}

}  // namespace sdfr
