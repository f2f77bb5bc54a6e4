// Rotated bird's-eye-view box overlap for the KITTI evaluator.
//
// replaces: the numba.cuda kernel of pipelines/rotate_iou.py (rotate_iou_kernel_eval, 257-286, and the
//           device functions it inlines, 22-254): corners of both boxes, vertices inside the other
//           quadrilateral, the 16 edge-edge intersections, angular sort about the centroid, fan
//           triangulation of the intersection polygon, then the criterion (-1 IoU, 0 / 1 over one area,
//           2 raw intersection).  Same float32 operand types, float64 where numba's typing promotes
//           (the triangle areas divide by the double constant 2.0 and accumulate in double; the
//           point-in-quadrilateral margins compare against the double constant 1e-4).
//
// One thread per (box, query) pair, 64 x 64 pairs per block with both box tiles staged in shared
// memory - the reference's tiling - but every thread owns a column of its row instead of looping over
// it (the reference serialises 64 pairs per thread).
#include "common.cuh"

namespace sdfr {

namespace {

__device__ __forceinline__ double triangle_area(const float* a, const float* b, const float* c) {
  return (double)((a[0] - c[0]) * (b[1] - c[1]) - (a[1] - c[1]) * (b[0] - c[0])) / 2.0;   // rotate_iou.py:23
}

__device__ double polygon_area(const float* pts, int n) {                                   // rotate_iou.py:27-31
  double area = 0.0;
  for (int i = 0; i < n - 2; ++i) area += fabs(triangle_area(pts, pts + 2 * i + 2, pts + 2 * i + 4));
  return area;
}

__device__ void sort_vertices(float* pts, int n) {                                          // rotate_iou.py:35-72
  if (n <= 0) return;
  float cx = 0.f, cy = 0.f;
  for (int i = 0; i < n; ++i) { cx += pts[2 * i]; cy += pts[2 * i + 1]; }
  cx = (float)((double)cx / (double)n);
  cy = (float)((double)cy / (double)n);
  float vs[16];
  for (int i = 0; i < n; ++i) {
    float vx = pts[2 * i] - cx, vy = pts[2 * i + 1] - cy;
    const float d = sqrtf(vx * vx + vy * vy);
    vx = vx / d;
    vy = vy / d;
    if (vy < 0.f) vx = -2.f - vx;
    vs[i] = vx;
  }
  for (int i = 1; i < n; ++i) {          // insertion sort, ascending pseudo-angle
    if (vs[i - 1] > vs[i]) {
      const float temp = vs[i], tx = pts[2 * i], ty = pts[2 * i + 1];
      int j = i;
      while (j > 0 && vs[j - 1] > temp) {
        vs[j] = vs[j - 1];
        pts[2 * j] = pts[2 * j - 2];
        pts[2 * j + 1] = pts[2 * j - 1];
        --j;
      }
      vs[j] = temp;
      pts[2 * j] = tx;
      pts[2 * j + 1] = ty;
    }
  }
}

__device__ bool segment_intersection(const float* p1, const float* p2, int i, int j, float* out) {   // :76-116
  const float ax = p1[2 * i], ay = p1[2 * i + 1];
  const float bx = p1[2 * ((i + 1) & 3)], by = p1[2 * ((i + 1) & 3) + 1];
  const float cx = p2[2 * j], cy = p2[2 * j + 1];
  const float dx = p2[2 * ((j + 1) & 3)], dy = p2[2 * ((j + 1) & 3) + 1];
  const float ba0 = bx - ax, ba1 = by - ay, da0 = dx - ax, ca0 = cx - ax, da1 = dy - ay, ca1 = cy - ay;
  const bool acd = da1 * ca0 > ca1 * da0;
  const bool bcd = (dy - by) * (cx - bx) > (cy - by) * (dx - bx);
  if (acd != bcd) {
    const bool abc = ca1 * ba0 > ba1 * ca0;
    const bool abd = da1 * ba0 > ba1 * da0;
    if (abc != abd) {
      const float dc0 = dx - cx, dc1 = dy - cy;
      const float abba = ax * by - bx * ay;
      const float cddc = cx * dy - dx * cy;
      const float dh = ba1 * dc0 - ba0 * dc1;
      out[0] = (abba * dc0 - ba0 * cddc) / dh;
      out[1] = (abba * dc1 - ba1 * cddc) / dh;
      return true;
    }
  }
  return false;
}

__device__ bool point_in_quad(float px, float py, const float* c) {                         // rotate_iou.py:158-175
  const float ab0 = c[2] - c[0], ab1 = c[3] - c[1];
  const float ad0 = c[6] - c[0], ad1 = c[7] - c[1];
  const float ap0 = px - c[0], ap1 = py - c[1];
  const float abab = ab0 * ab0 + ab1 * ab1, abap = ab0 * ap0 + ab1 * ap1;
  const float adad = ad0 * ad0 + ad1 * ad1, adap = ad0 * ap0 + ad1 * ap1;
  const double eps = 0.0001;
  return (double)abab >= (double)abap - eps && (double)abap >= 0.0 - eps && (double)adad >= (double)adap - eps &&
         (double)adap >= 0.0 - eps;
}

__device__ void box_corners(float* corners, const float* b) {                               // rotate_iou.py:203-226
  const float c = cosf(b[4]), s = sinf(b[4]);
  const float xs[4] = {-b[2] / 2.f, -b[2] / 2.f, b[2] / 2.f, b[2] / 2.f};
  const float ys[4] = {-b[3] / 2.f, b[3] / 2.f, b[3] / 2.f, -b[3] / 2.f};
  for (int i = 0; i < 4; ++i) {
    corners[2 * i] = c * xs[i] + s * ys[i] + b[0];
    corners[2 * i + 1] = -s * xs[i] + c * ys[i] + b[1];
  }
}

__device__ double intersection_area(const float* b1, const float* b2) {                     // rotate_iou.py:178-243
  float c1[8], c2[8], pts[16];
  box_corners(c1, b1);
  box_corners(c2, b2);
  int n = 0;
  for (int i = 0; i < 4; ++i) {
    if (point_in_quad(c1[2 * i], c1[2 * i + 1], c2)) { pts[2 * n] = c1[2 * i]; pts[2 * n + 1] = c1[2 * i + 1]; ++n; }
    if (point_in_quad(c2[2 * i], c2[2 * i + 1], c1)) { pts[2 * n] = c2[2 * i]; pts[2 * n + 1] = c2[2 * i + 1]; ++n; }
  }
  float t[2];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      if (n < 8 && segment_intersection(c1, c2, i, j, t)) { pts[2 * n] = t[0]; pts[2 * n + 1] = t[1]; ++n; }
  sort_vertices(pts, n);
  return polygon_area(pts, n);
}

constexpr int RT = 64;    // boxes per tile side (rotate_iou.py:259)

// block = 64 x 4 threads: threadIdx.x = query column of the tile, threadIdx.y strides over the rows
__global__ void __launch_bounds__(256) rotate_iou_kernel(long long N, long long K, const float* __restrict__ boxes,
                                                         const float* __restrict__ query, float* __restrict__ iou,
                                                         int criterion) {
  __shared__ float s_b[RT * 5], s_q[RT * 5];
  const long long row0 = (long long)blockIdx.x * RT, col0 = (long long)blockIdx.y * RT;
  const int rows = (int)min((long long)RT, N - row0), cols = (int)min((long long)RT, K - col0);
  const int tid = threadIdx.y * RT + threadIdx.x;
  for (int i = tid; i < rows * 5; i += 256) s_b[i] = boxes[row0 * 5 + i];
  for (int i = tid; i < cols * 5; i += 256) s_q[i] = query[col0 * 5 + i];
  __syncthreads();
  const int c = threadIdx.x;
  if (c >= cols) return;
  for (int r = threadIdx.y; r < rows; r += 4) {
    // devRotateIoUEval(query box, box): rbox1 is the QUERY box (rotate_iou.py:286), so criterion 0 divides
    // by the query box's area and criterion 1 by the box's
    const float* b1 = s_q + c * 5;
    const float* b2 = s_b + r * 5;
    const float area1 = b1[2] * b1[3], area2 = b2[2] * b2[3];
    const double inter = intersection_area(b1, b2);
    double v;
    if (criterion == -1) v = inter / ((double)area1 + (double)area2 - inter);
    else if (criterion == 0) v = inter / (double)area1;
    else if (criterion == 1) v = inter / (double)area2;
    else v = inter;
    iou[(row0 + r) * K + col0 + c] = (float)v;
  }
}

}  // namespace

}  // namespace sdfr

using namespace sdfr;

extern "C" int sdfr_rotate_iou(const float* boxes_dev, int64_t n, const float* query_dev, int64_t k, int criterion,
                               float* iou_dev, void* stream) {
  SDFR_REQUIRE(n >= 0 && k >= 0, SDFR_E_INVALID, "sdfr_rotate_iou: negative size");
  if (n == 0 || k == 0) return SDFR_OK;
  SDFR_REQUIRE(boxes_dev && query_dev && iou_dev, SDFR_E_INVALID, "sdfr_rotate_iou: null pointer");
  SDFR_REQUIRE((k + RT - 1) / RT <= 65535, SDFR_E_CAPACITY, "sdfr_rotate_iou: more than 65535 x 64 query boxes");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)((n + RT - 1) / RT), (unsigned)((k + RT - 1) / RT));
  rotate_iou_kernel<<<grid, dim3(RT, 4), 0, s>>>(n, k, boxes_dev, query_dev, iou_dev, criterion);
  SDFR_LAUNCH_CHECK();
  return SDFR_OK;
}
