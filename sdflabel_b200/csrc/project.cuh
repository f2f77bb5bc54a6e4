// Surfel projection shared by the stand-alone project kernel (splat.cu) and the fused isosurface-projection
// kernel of the refine engine (surface.cu).
//
// replaces: project_in_2D / project_in_2D_quat (sdfrenderer/renderer/projection.py:7-101,104-199) and
//           qrot (renderer/utils_rasterer.py:6-24).
#pragma once

#include "common.cuh"

namespace sdfr {

struct Rot {
  float r[9];
  float t[3];
};

// Linear map applied to points/normals.  dcm: rows of pose[:3,:3].  quat: the matrix of
// v -> v + 2 (qw (u x v) + u x (u x v)), u = q_xyz (not normalised, utils_rasterer.py:21-24).
__device__ __forceinline__ Rot load_rot(const SplatView& V) {
  Rot R;
  const float* p = V.pose;
  if (V.rot == SDFR_ROT_DCM) {
    R.r[0] = p[0]; R.r[1] = p[1]; R.r[2] = p[2];  R.t[0] = p[3];
    R.r[3] = p[4]; R.r[4] = p[5]; R.r[5] = p[6];  R.t[1] = p[7];
    R.r[6] = p[8]; R.r[7] = p[9]; R.r[8] = p[10]; R.t[2] = p[11];
  } else {
    const float w = p[0], x = p[1], y = p[2], z = p[3];
    R.r[0] = 1.f - 2.f * (y * y + z * z); R.r[1] = 2.f * (x * y - w * z);       R.r[2] = 2.f * (x * z + w * y);
    R.r[3] = 2.f * (x * y + w * z);       R.r[4] = 1.f - 2.f * (x * x + z * z); R.r[5] = 2.f * (y * z - w * x);
    R.r[6] = 2.f * (x * z - w * y);       R.r[7] = 2.f * (y * z + w * x);       R.r[8] = 1.f - 2.f * (x * x + y * y);
    R.t[0] = p[4]; R.t[1] = p[5]; R.t[2] = p[6];
  }
  return R;
}

__device__ __forceinline__ void interval_div(float lo, float hi, float zlo, float zhi, float& qlo, float& qhi) {
  // [lo,hi] / [zlo,zhi] with zlo > 0
  qlo = lo >= 0.f ? lo / zhi : lo / zlo;
  qhi = hi >= 0.f ? hi / zlo : hi / zhi;
}

// One surfel i of view V: object-frame centre p and unit normal n -> camera-space centre / normal / colour,
// plane offset, front flag and the conservative pixel box (projection.py:34-70, rasterer.py:113-114).
// `is_surfel` = 0 marks a pre-selected row outside the band: invisible to every later stage.
__device__ __forceinline__ void project_surfel(const SplatView& V, const int i, const float px, const float py,
                                               const float pz, const float nx, const float ny, const float nz,
                                               const bool is_surfel) {
  const Rot R = load_rot(V);
  float vx, vy, vz, mx, my, mz;
  if (V.rot == SDFR_ROT_DCM) {
    vx = R.r[0] * px + R.r[1] * py + R.r[2] * pz + R.t[0];
    vy = R.r[3] * px + R.r[4] * py + R.r[5] * pz + R.t[1];
    vz = R.r[6] * px + R.r[7] * py + R.r[8] * pz + R.t[2];
    mx = R.r[0] * nx + R.r[1] * ny + R.r[2] * nz;
    my = R.r[3] * nx + R.r[4] * ny + R.r[5] * nz;
    mz = R.r[6] * nx + R.r[7] * ny + R.r[8] * nz;
  } else {
    // qrot evaluated as written in the reference (two cross products)
    const float w = V.pose[0], ux = V.pose[1], uy = V.pose[2], uz = V.pose[3];
    {
      const float ax = uy * pz - uz * py, ay = uz * px - ux * pz, az = ux * py - uy * px;
      const float bx = uy * az - uz * ay, by = uz * ax - ux * az, bz = ux * ay - uy * ax;
      vx = px + 2.f * (w * ax + bx) + R.t[0];
      vy = py + 2.f * (w * ay + by) + R.t[1];
      vz = pz + 2.f * (w * az + bz) + R.t[2];
    }
    {
      const float ax = uy * nz - uz * ny, ay = uz * nx - ux * nz, az = ux * ny - uy * nx;
      const float bx = uy * az - uz * ay, by = uz * ax - ux * az, bz = ux * ay - uy * ax;
      mx = nx + 2.f * (w * ax + bx);
      my = ny + 2.f * (w * ay + by);
      mz = nz + 2.f * (w * az + bz);
    }
  }
  float cx, cy, cz;
  if (V.output_nocs) {   // projection.py:53-55 (dcm negates x) / 147-149 (quat does not)
    cx = V.rot == SDFR_ROT_DCM ? -px : px; cy = py; cz = pz;
  } else {
    cx = V.colors[i * 3]; cy = V.colors[i * 3 + 1]; cz = V.colors[i * 3 + 2];
  }
  V.cam_v[i * 3] = vx; V.cam_v[i * 3 + 1] = vy; V.cam_v[i * 3 + 2] = vz;
  V.cam_m[i * 3] = mx; V.cam_m[i * 3 + 1] = my; V.cam_m[i * 3 + 2] = mz;
  if (V.output_nocs) {   // rasterer.py:113-114
    V.cam_c[i * 3] = (cx + 1.f) / 2.f; V.cam_c[i * 3 + 1] = (cy + 1.f) / 2.f; V.cam_c[i * 3 + 2] = (cz + 1.f) / 2.f;
  } else {
    V.cam_c[i * 3] = cx; V.cam_c[i * 3 + 1] = cy; V.cam_c[i * 3 + 2] = cz;
  }
  if (V.cam_rgb) {       // rasterer.py:150
    V.cam_rgb[i * 3] = (cx + 1.f) / 2.f; V.cam_rgb[i * 3 + 1] = (cy + 1.f) / 2.f; V.cam_rgb[i * 3 + 2] = (cz + 1.f) / 2.f;
  }
  const float a = mx * vx + my * vy + mz * vz;
  V.plane_a[i] = a;
  V.front[i] = (V.rot == SDFR_ROT_DCM) ? (a < 0.f ? 1 : 0) : 1;   // projection.py:61-66
  if (!is_surfel) {                 // pre-selected but outside the band: invisible to every later stage
    V.front[i] = 0;
    V.bbox[i * 4] = 1; V.bbox[i * 4 + 1] = 1; V.bbox[i * 4 + 2] = 0; V.bbox[i * 4 + 3] = 0;
    return;
  }

  // conservative pixel box of the ball B(v, radius)
  int x0 = 0, y0 = 0, x1 = V.width - 1, y1 = V.height - 1;
  const float rad = kDiscRadius * 1.001f;
  const bool affine_k = V.k[6] == 0.f && V.k[7] == 0.f && V.k[8] == 1.f;
  if (affine_k && vz - rad > 1e-6f && isfinite(vx) && isfinite(vy) && isfinite(vz)) {
    float ulo, uhi, wlo, whi;
    interval_div(vx - rad, vx + rad, vz - rad, vz + rad, ulo, uhi);
    interval_div(vy - rad, vy + rad, vz - rad, vz + rad, wlo, whi);
    const float k00 = V.k[0], k01 = V.k[1], k02 = V.k[2], k10 = V.k[3], k11 = V.k[4], k12 = V.k[5];
    const float xa = fminf(k00 * ulo, k00 * uhi) + fminf(k01 * wlo, k01 * whi) + k02;
    const float xb = fmaxf(k00 * ulo, k00 * uhi) + fmaxf(k01 * wlo, k01 * whi) + k02;
    const float ya = fminf(k10 * ulo, k10 * uhi) + fminf(k11 * wlo, k11 * whi) + k12;
    const float yb = fmaxf(k10 * ulo, k10 * uhi) + fmaxf(k11 * wlo, k11 * whi) + k12;
    const float big = 1e8f;
    x0 = max(0, (int)floorf(fmaxf(xa, -big)) - 1);
    y0 = max(0, (int)floorf(fmaxf(ya, -big)) - 1);
    x1 = min(V.width - 1, (int)ceilf(fminf(xb, big)) + 1);
    y1 = min(V.height - 1, (int)ceilf(fminf(yb, big)) + 1);
  }
  V.bbox[i * 4] = x0; V.bbox[i * 4 + 1] = y0; V.bbox[i * 4 + 2] = x1; V.bbox[i * 4 + 3] = y1;
}


}  // namespace sdfr
