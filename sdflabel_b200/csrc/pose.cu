// Initial-pose RANSAC on the device: exact nearest-neighbour queries and the scoring of all
// hypotheses in one launch.
//
// replaces: the two sklearn KD-trees of PoseEstimator.init_pose_3d (utils/pose.py:133-134) and
//           their queries - the NOCS correspondences of the sampled scene points (146), and, per
//           RANSAC hypothesis, the transform of the scene cloud (170), its 1-NN in the model cloud
//           (172-175) and the inlier test (178).  The reference runs the ~567 hypotheses one after
//           the other on the CPU (a KD-tree query of the whole scene cloud each); here a
//           hypothesis is blockIdx.y and every (hypothesis, scene point) pair is a thread.
//
// A KD-tree returns the exact nearest neighbour with the distance evaluated in float64 on the
// float32 coordinates, so the brute-force scan accumulates the squared distance in double, in the
// tree's order (x, y, z) and without fused multiply-adds: same neighbour, same distance bits.
#include "common.cuh"

namespace sdfr {

namespace {

constexpr int PB = 256;
constexpr int PSTAGE = 1024;   // reference points staged per shared-memory chunk

__device__ __forceinline__ void nn_scan_exact(const float* __restrict__ pts, int n, int index_base, float qx, float qy,
                                              float qz, double& best_d2, int& best_i) {
  for (int k = 0; k < n; ++k) {
    const double dx = (double)qx - (double)pts[k * 3];
    const double dy = (double)qy - (double)pts[k * 3 + 1];
    const double dz = (double)qz - (double)pts[k * 3 + 2];
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    if (d2 < best_d2) {      // strict: the first of equal candidates wins
      best_d2 = d2;
      best_i = index_base + k;
    }
  }
}

__global__ void __launch_bounds__(PB) nn_query_kernel(const float* __restrict__ q, long long nq,
                                                      const float* __restrict__ refs, long long m,
                                                      int* __restrict__ idx, double* __restrict__ dist) {
  __shared__ float s_pts[PSTAGE * 3];
  const long long i = (long long)blockIdx.x * PB + threadIdx.x;
  const bool live = i < nq;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (live) { qx = q[i * 3]; qy = q[i * 3 + 1]; qz = q[i * 3 + 2]; }
  double best = INFINITY;
  int bi = -1;
  for (long long base = 0; base < m; base += PSTAGE) {
    const int n = (int)min((long long)PSTAGE, m - base);
    for (int k = threadIdx.x; k < n * 3; k += PB) s_pts[k] = refs[base * 3 + k];
    __syncthreads();
    if (live) nn_scan_exact(s_pts, n, (int)base, qx, qy, qz, best, bi);
    __syncthreads();
  }
  if (live) {
    idx[i] = bi;
    dist[i] = sqrt(best);
  }
}

// thread = (scene point s, hypothesis h): inlier[h][s] and counts[h]
__global__ void __launch_bounds__(PB) ransac_score_kernel(const float* __restrict__ scene_pts,
                                                          const float* __restrict__ scene_cls, long long ns,
                                                          const float* __restrict__ model_pts,
                                                          const float* __restrict__ model_cls, long long m,
                                                          const float* __restrict__ transforms, double metric_thr,
                                                          float nocs_thr, int* __restrict__ counts,
                                                          unsigned char* __restrict__ masks) {
  __shared__ float s_pts[PSTAGE * 3];
  __shared__ float s_t[12];
  __shared__ int s_cnt[PB / 32];
  const int h = blockIdx.y;
  const long long s = (long long)blockIdx.x * PB + threadIdx.x;
  const bool live = s < ns;
  if (threadIdx.x < 12) s_t[threadIdx.x] = transforms[(long long)h * 12 + threadIdx.x];
  __syncthreads();
  float px = 0.f, py = 0.f, pz = 0.f;
  if (live) {
    // (trans[:, :3] @ scene_pts.T).T + trans[:, 3] in float32 (pose.py:166-170): a k = 0, 1, 2
    // multiply-add chain, then the translation as a separate rounded add
    const float x = scene_pts[s * 3], y = scene_pts[s * 3 + 1], z = scene_pts[s * 3 + 2];
    px = __fadd_rn(fmaf(s_t[2], z, fmaf(s_t[1], y, __fmul_rn(s_t[0], x))), s_t[3]);
    py = __fadd_rn(fmaf(s_t[6], z, fmaf(s_t[5], y, __fmul_rn(s_t[4], x))), s_t[7]);
    pz = __fadd_rn(fmaf(s_t[10], z, fmaf(s_t[9], y, __fmul_rn(s_t[8], x))), s_t[11]);
  }
  double best = INFINITY;
  int bi = -1;
  for (long long base = 0; base < m; base += PSTAGE) {
    const int n = (int)min((long long)PSTAGE, m - base);
    for (int k = threadIdx.x; k < n * 3; k += PB) s_pts[k] = model_pts[base * 3 + k];
    __syncthreads();
    if (live) nn_scan_exact(s_pts, n, (int)base, px, py, pz, best, bi);
    __syncthreads();
  }
  int in = 0;
  if (live && bi >= 0) {
    // dists_color = np.linalg.norm(scene_cls - model_cls[idxs], axis=1) in float32 (pose.py:175)
    const float e0 = __fsub_rn(scene_cls[s * 3], model_cls[(long long)bi * 3]);
    const float e1 = __fsub_rn(scene_cls[s * 3 + 1], model_cls[(long long)bi * 3 + 1]);
    const float e2 = __fsub_rn(scene_cls[s * 3 + 2], model_cls[(long long)bi * 3 + 2]);
    const float dc = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(e0, e0), __fmul_rn(e1, e1)), __fmul_rn(e2, e2)));
    in = (sqrt(best) < metric_thr && dc < nocs_thr) ? 1 : 0;    // pose.py:178
  }
  if (live) masks[(long long)h * ns + s] = (unsigned char)in;
  int c = in;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < PB / 32; ++w) t += s_cnt[w];
    if (t) atomicAdd(&counts[h], t);      // integer: the result does not depend on the order
  }
}

}  // namespace

}  // namespace sdfr

using namespace sdfr;

extern "C" int sdfr_nn_query(const float* queries_dev, int64_t q, const float* refs_dev, int64_t m, int32_t* idx_dev,
                             double* dist_dev, void* stream) {
  SDFR_REQUIRE(q >= 0 && m >= 0, SDFR_E_INVALID, "sdfr_nn_query: negative size");
  if (q == 0) return SDFR_OK;
  SDFR_REQUIRE(m > 0, SDFR_E_INVALID, "sdfr_nn_query: empty reference set");
  SDFR_REQUIRE(queries_dev && refs_dev && idx_dev && dist_dev, SDFR_E_INVALID, "sdfr_nn_query: null pointer");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  nn_query_kernel<<<(unsigned)((q + PB - 1) / PB), PB, 0, s>>>(queries_dev, q, refs_dev, m, idx_dev, dist_dev);
  SDFR_LAUNCH_CHECK();
  return SDFR_OK;
}

extern "C" int sdfr_ransac_score(const float* scene_pts_dev, const float* scene_cls_dev, int64_t ns,
                                 const float* model_pts_dev, const float* model_cls_dev, int64_t m,
                                 const float* transforms_dev, int num_hypotheses, double metric_thr, float nocs_thr,
                                 int32_t* counts_dev, uint8_t* masks_dev, void* stream) {
  SDFR_REQUIRE(ns >= 0 && m >= 0 && num_hypotheses >= 0, SDFR_E_INVALID, "sdfr_ransac_score: negative size");
  if (num_hypotheses == 0) return SDFR_OK;
  SDFR_REQUIRE(num_hypotheses <= 65535, SDFR_E_CAPACITY, "sdfr_ransac_score: more than 65535 hypotheses");
  SDFR_REQUIRE(counts_dev && masks_dev, SDFR_E_INVALID, "sdfr_ransac_score: null output");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  SDFR_CUDA(cudaMemsetAsync(counts_dev, 0, sizeof(int32_t) * (size_t)num_hypotheses, s));
  if (ns == 0) return SDFR_OK;
  SDFR_REQUIRE(m > 0, SDFR_E_INVALID, "sdfr_ransac_score: empty model cloud");
  SDFR_REQUIRE(scene_pts_dev && scene_cls_dev && model_pts_dev && model_cls_dev && transforms_dev, SDFR_E_INVALID,
               "sdfr_ransac_score: null pointer");
  dim3 grid((unsigned)((ns + PB - 1) / PB), (unsigned)num_hypotheses);
  ransac_score_kernel<<<grid, PB, 0, s>>>(scene_pts_dev, scene_cls_dev, ns, model_pts_dev, model_cls_dev, m,
                                          transforms_dev, metric_thr, nocs_thr, counts_dev, masks_dev);
  SDFR_LAUNCH_CHECK();
  return SDFR_OK;
}
