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

// One launch of the fused march (device-resident descriptors; the lists ping-pong, the counters rotate).
//
// Speculative sphere tracing.  A launch costs one pass over the weights however few rays are left, so the rows a
// round of the grid has to spare are spent on look-ahead: every listed ray is evaluated at K = 2^lk sample points
//   s_0 = tau,   s_k = s_(k-1) + dt0 r^(k-1)        (dt0 = c fh, r = 1 + c mh: the spacing a planar surface at the
//                                                    predicted distance fh and slope mh would ask for)
// with K the largest power of two <= 32 such that count * K fits one round.  The epilogue then walks the samples of a
// ray in order and keeps those whose unbounding spheres connect (s_k - |f_k| <= front, the over-relaxation test of
// enhanced sphere tracing): the safe front moves to max(s_k + f_k).  Sample 0 alone is the plain sphere-tracing step,
// so a launch never advances less than one; a grazing ray advances up to 32 of them.  Per ray afterwards:
//   a connected sample with |sdf| < near_thr -> the ray goes to the `near` list at that sample (finished by Newton
//                                               steps at full precision)
//   front beyond the exit / inside the surface -> dropped
//   otherwise                                 -> tau = front and the ray goes to `next_list`
struct RayMarch {
  TraceParams p;
  float near_thr;
  float near_reach;           // hand-over only when the predicted root |sdf| / |slope| is within this distance
  float spacing;              // c above
  const float* latent_unit;   // [latent]
  float* tau;                 // [P] safe front of the ray
  const float* tau_exit;      // [P]
  float* fh;                  // [P] predicted sdf at tau
  float* mh;                  // [P] slope estimate d sdf / d tau in [-1, 0]
  float* ls;                  // [P] ray parameter of the last evaluated connected sample (-1e30: none yet)
  float* lf;                  // [P] its sdf
  int* nsteps;                // [P] connected samples the ray has consumed (each is worth one plain step)
  int step_budget;            // a ray that has consumed more is abandoned (the plain march's max_steps rule)
  const int* list;
  const int* count;
  int* next_list;
  int* next_count;
  int* near_list;
  int* near_count;
  int* reset_count;           // counter the launch after the next appends to: cleared by this launch
  int round_rows;             // rows of one round of the grid
  int max_log2k;              // K <= 2^max_log2k (<= 5: the samples of a ray share a warp of the epilogue)
};

// samples per ray of a launch over `count` rays (uniform over the grid)
__device__ __forceinline__ int march_log2k(const RayMarch& m, int count) {
  int lk = 0;
  while (lk < m.max_log2k && ((long long)count << (lk + 1)) <= (long long)m.round_rows) ++lk;
  return lk;
}
__device__ __forceinline__ long long march_rows(const RayMarch& m) {
  const int c = *m.count;
  return c <= 0 ? 0 : (long long)c << march_log2k(m, c);
}
// ray parameter of sample k of a ray (explicitly rounded operations: the row generator and the epilogue agree)
__device__ __forceinline__ float march_sample(const RayMarch& m, float tau, float fh, float mh, int k) {
  const float r = fminf(fmaxf(__fmaf_rn(m.spacing, mh, 1.f), 0.3f), 1.f);
  float dt = __fmul_rn(m.spacing, fmaxf(fh, 0.4f * m.near_thr));   // floor: 1e-3, the resolution of this pass
  float s = tau;
  for (int i = 0; i < k; ++i) {
    s = __fadd_rn(s, dt);
    dt = __fmul_rn(dt, r);
  }
  return s;
}

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

// decoder input column c of march row `row` (latent, then the sample point o + s_k d of the listed ray)
__device__ __forceinline__ float march_input(const RayMarch& m, long long row, int c, int lk) {
  if (c < m.p.latent) return m.latent_unit[c];
  const int j = m.list[row >> lk];
  float o[3], d[3], rn[3];
  ray_of_pixel(m.p, j, o, d, rn);
  const int a = c - m.p.latent;
  return o[a] + march_sample(m, m.tau[j], m.fh[j], m.mh[j], (int)(row & ((1 << lk) - 1))) * d[a];
}

// the advance of the evaluated march rows of one warp (all 32 lanes call it; lane = row mod 32, so the 2^lk samples of
// a ray sit in consecutive lanes; `valid` = the lane holds a row)
__device__ __forceinline__ void march_advance(const RayMarch& m, bool valid, long long row, float f, int lk) {
  const int lane = threadIdx.x & 31, K = 1 << lk, k = lane & (K - 1), g0 = lane - k;
  int j = 0;
  float tau0 = 0.f, texit = 0.f, fh = 0.f, mh = 0.f, lasts = -1e30f, lastf = 0.f;
  if (valid) {
    j = m.list[row >> lk];
    tau0 = m.tau[j]; texit = m.tau_exit[j]; fh = m.fh[j]; mh = m.mh[j]; lasts = m.ls[j]; lastf = m.lf[j];
  }
  const float s_own = march_sample(m, tau0, fh, mh, k);
  float front = tau0, tnear = 0.f, prevs = -1e30f, prevf = 0.f;
  bool alive = valid, isnear = false, dead = false, inside = false;
  int used = 0;
  for (int kk = 0; kk < K; ++kk) {
    const float fk = __shfl_sync(0xffffffffu, f, g0 + kk), sk = __shfl_sync(0xffffffffu, s_own, g0 + kk);
    const bool reach = alive && (kk == 0 || sk - fabsf(fk) <= front) && sk <= texit && fk == fk;
    // hand-over to the Newton finish: close to the surface AND the root within Newton's reach - a ray that runs at a
    // shallow angle (|d sdf / d tau| small) keeps marching here, where a launch is cheap, until it is within 1e-3
    // (three times the error of this pass)
    const float slope_k = kk > 0 || lasts > -1e29f ? (fk - lastf) / fmaxf(sk - lasts, 1e-9f) : mh;
    const bool nr = reach && fabsf(fk) < m.near_thr && (fabsf(fk) < 1e-3f || fabsf(fk) <= m.near_reach * fabsf(slope_k));
    const bool ok = reach && !nr && fk > 0.f;
    if (nr) { isnear = true; tnear = sk; }
    if (reach && !nr && !ok) {
      // inside the surface (the field is not an exact distance: a step can overshoot).  Sample 0: step back by |f|
      // like the plain march does; later samples: the sign change is bracketed by the last connected sample
      inside = true;
      front = kk == 0 ? tau0 + fk : fminf(front, lasts + (sk - lasts) * lastf / (lastf - fk));
    } else if (kk == 0 && alive && !nr && !ok) {
      dead = true;                                         // NaN
    }
    if (ok) {
      ++used;
      prevs = lasts; prevf = lastf;
      lasts = sk; lastf = fk;
      front = fmaxf(front, sk + fk);
    }
    alive = ok;
  }
  bool keep = false;
  const bool leader = valid && k == 0;
#ifdef SDFR_TRACE_DEBUG
  if (valid && j == SDFR_TRACE_DEBUG)
    printf("march ray %d sample %d/%d: s %.6f f %.6f | tau0 %.6f fh %.5f mh %.3f front %.6f near %d tnear %.6f dead %d inside %d texit %.4f\n",
           j, k, K, s_own, f, tau0, fh, mh, front, (int)isnear, tnear, (int)dead, (int)inside, texit);
#endif
  if (leader) {
    if (isnear) {
      m.tau[j] = tnear;                  // tau stays at the evaluated point
    } else if (!dead && front <= texit && front >= 0.f && (used += m.nsteps[j]) <= m.step_budget) {
      keep = true;
      m.nsteps[j] = used;
      float slope = mh;
      if (prevs > -1e29f && lasts > prevs) slope = fminf(fmaxf((lastf - prevf) / (lasts - prevs), -1.f), 0.f);
      const float ahead = front - lasts;
      m.tau[j] = front;
      // predicted distance at the new front; next to a bracketed sign change the samples are spaced at near_thr
      m.fh[j] = inside ? 0.f : fmaxf(fmaxf(lastf + slope * ahead, 0.25f * ahead), 0.f);
      m.mh[j] = inside ? 0.f : slope;
      m.ls[j] = inside ? -1e30f : lasts;
      m.lf[j] = inside ? 0.f : lastf;
    }
  }
  ray_append(keep, j, m.next_list, m.next_count);
  ray_append(leader && isnear, j, m.near_list, m.near_count);
}

}  // namespace sdfr
