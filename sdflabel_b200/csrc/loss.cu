// Stand-alone loss entry points (value + gradient).  The fused engine has its
// own batched variants in refine.cu built on the same device helpers.
//
// replaces: Optimizer.compute_loss_3d (pipelines/optimizer.py:166-198): exact
//           brute-force 1-NN on the device instead of the D2H copy + sklearn
//           KD-tree; Optimizer.compute_loss_2d (pipelines/optimizer.py:200-237):
//           9x9 window + shared far candidate instead of O(M*H*W) tensors.
#include "loss.cuh"

namespace sdfr {

namespace {

constexpr int LB = 256;
constexpr int STAGE = 1024;   // lidar points staged per shared-memory chunk

// per query: nearest neighbour, thresholded pair distance; per block: (count, sum)
__global__ void __launch_bounds__(LB) loss3d_pairs_kernel(const float* __restrict__ q, long long nq,
                                                          const float* __restrict__ lidar, long long nl,
                                                          double radius, float* __restrict__ rec /*[nq,4]*/,
                                                          int* __restrict__ rec_idx, double* __restrict__ partial) {
  __shared__ float s_pts[STAGE * 3];
  __shared__ double s_sum[LB / 32];
  __shared__ int s_cnt[LB / 32];
  const long long i = (long long)blockIdx.x * LB + threadIdx.x;
  const bool live = i < nq;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (live) { qx = q[i * 3]; qy = q[i * 3 + 1]; qz = q[i * 3 + 2]; }
  double best = INFINITY;
  int bi = -1;
  for (long long base = 0; base < nl; base += STAGE) {
    const int n = (int)min((long long)STAGE, nl - base);
    for (int k = threadIdx.x; k < n * 3; k += LB) s_pts[k] = lidar[base * 3 + k];
    __syncthreads();
    if (live) nn_scan(s_pts, n, (int)base, qx, qy, qz, best, bi);
    __syncthreads();
  }
  float dist = 0.f, ux = 0.f, uy = 0.f, uz = 0.f;
  int close = 0;
  if (live && bi >= 0 && sqrt(best) < radius) {      // optimizer.py:188
    close = 1;
    const float ex = lidar[(long long)bi * 3] - qx, ey = lidar[(long long)bi * 3 + 1] - qy,
                ez = lidar[(long long)bi * 3 + 2] - qz;
    dist = sqrtf(ex * ex + ey * ey + ez * ez);       // optimizer.py:189
    if (dist > 0.f) { ux = ex / dist; uy = ey / dist; uz = ez / dist; }
  }
  if (live) {
    rec[i * 4] = ux; rec[i * 4 + 1] = uy; rec[i * 4 + 2] = uz; rec[i * 4 + 3] = close ? dist : -1.f;
    rec_idx[i] = bi;
  }
  double s = close ? (double)dist : 0.0;
  int c = close;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = s; s_cnt[threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0;
    int tc = 0;
    for (int w = 0; w < LB / 32; ++w) { ts += s_sum[w]; tc += s_cnt[w]; }
    partial[blockIdx.x * 2] = ts;
    partial[blockIdx.x * 2 + 1] = (double)tc;
  }
}

__global__ void loss_finalize_kernel(const double* __restrict__ partial, int nblocks, float* __restrict__ loss,
                                     int nan_when_empty) {
  // single thread: ordered (deterministic) reduction of the block partials
  double s = 0.0, c = 0.0;
  for (int b = 0; b < nblocks; ++b) { s += partial[b * 2]; c += partial[b * 2 + 1]; }
  loss[1] = (float)c;
  if (c > 0.0) loss[0] = (float)(s / c);
  else loss[0] = nan_when_empty ? __int_as_float(0x7fc00000) : 0.f;
}

__global__ void __launch_bounds__(LB) loss3d_grad_kernel(const float* __restrict__ rec, const int* __restrict__ rec_idx,
                                                         long long nq, const float* __restrict__ loss,
                                                         float* __restrict__ d_q, float* __restrict__ d_lidar) {
  const long long i = (long long)blockIdx.x * LB + threadIdx.x;
  if (i >= nq) return;
  const float n = loss[1];
  const bool close = rec[i * 4 + 3] >= 0.f && n > 0.f;
  const float sx = close ? rec[i * 4] / n : 0.f, sy = close ? rec[i * 4 + 1] / n : 0.f,
              sz = close ? rec[i * 4 + 2] / n : 0.f;
  if (d_q) { d_q[i * 3] = -sx; d_q[i * 3 + 1] = -sy; d_q[i * 3 + 2] = -sz; }
  if (d_lidar && close) {
    const long long k = rec_idx[i];
    atomicAdd(d_lidar + k * 3, sx); atomicAdd(d_lidar + k * 3 + 1, sy); atomicAdd(d_lidar + k * 3 + 2, sz);
  }
}

// ---- 2D -----------------------------------------------------------------------
__global__ void __launch_bounds__(LB) loss2d_pixels_kernel(const float* __restrict__ color,
                                                           const float* __restrict__ target, int H, int W,
                                                           float* __restrict__ rec /*[P,4] dir(3), delta or -1*/,
                                                           double* __restrict__ partial /*[blocks,3]*/) {
  __shared__ double s_sum[LB / 32], s_hw[LB / 32];
  __shared__ int s_cnt[LB / 32];
  const int P = H * W;
  const int j = blockIdx.x * LB + threadIdx.x;
  double sum = 0.0, hw = 0.0;
  int cnt = 0;
  if (j < P) {
    const float c0 = color[j], c1 = color[P + j], c2 = color[2 * P + j];
    float4 r = make_float4(0.f, 0.f, 0.f, -1.f);
    if (c0 + c1 + c2 != 0.f) {                       // optimizer.py:213
      const int h = j / W, w = j - h * W;
      hw = (double)(h + w);
      float k0, k1, k2;
      const float d = loss2d_pixel(target, H, W, h, w, c0, c1, c2, k0, k1, k2);
      if (d < kNocsThr) {                            // optimizer.py:234
        cnt = 1;
        sum = (double)d;
        if (d > 0.f) r = make_float4((c0 - k0) / d, (c1 - k1) / d, (c2 - k2) / d, d);
        else r.w = d;
      }
    }
    *reinterpret_cast<float4*>(rec + (size_t)j * 4) = r;
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
    partial[blockIdx.x * 3] = ts; partial[blockIdx.x * 3 + 1] = (double)tc; partial[blockIdx.x * 3 + 2] = th;
  }
}

__global__ void loss2d_finalize_kernel(const double* __restrict__ partial, int nblocks, float* __restrict__ loss) {
  double s = 0.0, c = 0.0, hw = 0.0;
  for (int b = 0; b < nblocks; ++b) { s += partial[b * 3]; c += partial[b * 3 + 1]; hw += partial[b * 3 + 2]; }
  if (hw == 0.0) {             // `if rendering_nonzero_idxs.sum()` (optimizer.py:214): empty, or only pixel (0,0)
    loss[0] = 0.f; loss[1] = 0.f;
    return;
  }
  loss[1] = (float)c;
  loss[0] = c > 0.0 ? (float)(s / c) : __int_as_float(0x7fc00000);   // mean of an empty selection is NaN
}

__global__ void __launch_bounds__(LB) loss2d_grad_kernel(const float* __restrict__ rec, int P,
                                                         const float* __restrict__ loss, float* __restrict__ d_color) {
  const int j = blockIdx.x * LB + threadIdx.x;
  if (j >= P) return;
  const float n = loss[1];
  const float4 r = *reinterpret_cast<const float4*>(rec + (size_t)j * 4);
  const bool sel = r.w >= 0.f && n > 0.f;
  d_color[j] = sel ? r.x / n : 0.f;
  d_color[P + j] = sel ? r.y / n : 0.f;
  d_color[2 * P + j] = sel ? r.z / n : 0.f;
}

// scratch cache (stand-alone calls only)
struct Scratch {
  void* ptr = nullptr;
  size_t bytes = 0;
};
thread_local Scratch g_scratch;

int get_scratch(size_t bytes, void** out) {
  if (g_scratch.bytes < bytes) {
    if (g_scratch.ptr) SDFR_CUDA(cudaFree(g_scratch.ptr));
    g_scratch.ptr = nullptr;
    g_scratch.bytes = 0;
    SDFR_CUDA(cudaMalloc(&g_scratch.ptr, bytes));
    g_scratch.bytes = bytes;
  }
  *out = g_scratch.ptr;
  return SDFR_OK;
}

}  // namespace

int launch_loss3d_standalone(const float* xyzf, long long q, const float* lidar, long long nl, double radius,
                             float* loss, float* d_xyzf, float* d_lidar, cudaStream_t s) {
  if (d_lidar && nl > 0) SDFR_CUDA(cudaMemsetAsync(d_lidar, 0, (size_t)nl * 3 * sizeof(float), s));
  if (q <= 0 || nl <= 0) {   // optimizer.py:177,197
    SDFR_CUDA(cudaMemsetAsync(loss, 0, 2 * sizeof(float), s));
    if (d_xyzf && q > 0) SDFR_CUDA(cudaMemsetAsync(d_xyzf, 0, (size_t)q * 3 * sizeof(float), s));
    return SDFR_OK;
  }
  const int nb = (int)((q + LB - 1) / LB);
  const size_t rec_bytes = (size_t)q * 4 * sizeof(float), idx_bytes = (size_t)q * sizeof(int);
  const size_t part_bytes = (size_t)nb * 2 * sizeof(double);
  void* base = nullptr;
  int rc = get_scratch(rec_bytes + idx_bytes + part_bytes + 64, &base);
  if (rc) return rc;
  double* partial = reinterpret_cast<double*>(base);
  float* rec = reinterpret_cast<float*>(reinterpret_cast<char*>(base) + ((part_bytes + 15) & ~(size_t)15));
  int* rec_idx = reinterpret_cast<int*>(reinterpret_cast<char*>(rec) + rec_bytes);
  loss3d_pairs_kernel<<<nb, LB, 0, s>>>(xyzf, q, lidar, nl, radius, rec, rec_idx, partial);
  SDFR_LAUNCH_CHECK();
  loss_finalize_kernel<<<1, 1, 0, s>>>(partial, nb, loss, 0);
  SDFR_LAUNCH_CHECK();
  if (d_xyzf || d_lidar) {
    loss3d_grad_kernel<<<nb, LB, 0, s>>>(rec, rec_idx, q, loss, d_xyzf, d_lidar);
    SDFR_LAUNCH_CHECK();
  }
  return SDFR_OK;
}

int launch_loss2d_standalone(const float* color, const float* target, int h, int w, float* loss, float* d_color,
                             cudaStream_t s) {
  const int P = h * w;
  if (P <= 0) {
    SDFR_CUDA(cudaMemsetAsync(loss, 0, 2 * sizeof(float), s));
    return SDFR_OK;
  }
  const int nb = (P + LB - 1) / LB;
  const size_t part_bytes = ((size_t)nb * 3 * sizeof(double) + 15) & ~(size_t)15;
  void* base = nullptr;
  int rc = get_scratch(part_bytes + (size_t)P * 4 * sizeof(float), &base);
  if (rc) return rc;
  double* partial = reinterpret_cast<double*>(base);
  float* rec = reinterpret_cast<float*>(reinterpret_cast<char*>(base) + part_bytes);
  loss2d_pixels_kernel<<<nb, LB, 0, s>>>(color, target, h, w, rec, partial);
  SDFR_LAUNCH_CHECK();
  loss2d_finalize_kernel<<<1, 1, 0, s>>>(partial, nb, loss);
  SDFR_LAUNCH_CHECK();
  if (d_color) {
    loss2d_grad_kernel<<<nb, LB, 0, s>>>(rec, P, loss, d_color);
    SDFR_LAUNCH_CHECK();
  }
  return SDFR_OK;
}

}  // namespace sdfr
