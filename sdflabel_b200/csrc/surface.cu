// Grid3D lattice and zero-isosurface extraction.
//
// replaces: Grid3D.generate_point_grid (sdfrenderer/grid.py:22-41) and
//           Grid3D.get_surface_points (sdfrenderer/grid.py:43-71).
//
// Both are HBM sweeps: the lattice is generated from the index (no reads), the
// extraction reads sdf (4 B) + gradient (12 B) per lattice point once and
// writes 28 B (+4L) per surviving point.  Order-preserving compaction (the
// reference's masked_select keeps ascending index order) is a two-launch
// count/scatter with warp-ballot ranking inside the block.
#include "common.cuh"

namespace sdfr {

namespace {

constexpr int SB = 1024;   // elements (= threads) per block

__global__ void lattice_points_kernel(LatticeParams lp, long long n, float* __restrict__ pts) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x, y, z;
  lattice_point(lp, i, x, y, z);
  pts[i * 3 + 0] = x;
  pts[i * 3 + 1] = y;
  pts[i * 3 + 2] = z;
}

__device__ __forceinline__ bool in_band(float sdf, float thr) { return fabsf(sdf) < thr; }

__global__ void __launch_bounds__(SB) band_count_kernel(SurfaceArgs a, int nblocks) {
  const int b = blockIdx.y;
  const long long i = (long long)blockIdx.x * SB + threadIdx.x;
  const bool keep = i < a.n && in_band(a.sdf[(long long)b * a.n + i], a.threshold);
  const int c = __syncthreads_count(keep);
  if (threadIdx.x == 0) a.scratch[(long long)b * (nblocks + 1) + blockIdx.x] = c;
}

__global__ void __launch_bounds__(SB) band_scatter_kernel(SurfaceArgs a, int nblocks) {
  __shared__ int warp_cnt[SB / 32];
  __shared__ int block_base;
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int* counts = a.scratch + (long long)b * (nblocks + 1);
  // exclusive prefix of the block counts (warp 0)
  if (warp == 0) {
    int s = 0;
    for (int j = lane; j < (int)blockIdx.x; j += 32) s += counts[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) block_base = s;
    if (blockIdx.x == nblocks - 1 && lane == 0) a.out_count[b] = s + counts[blockIdx.x];
  }
  const long long i = (long long)blockIdx.x * SB + tid;
  float f = 0.f;
  bool keep = false;
  if (i < a.n) {
    f = a.sdf[(long long)b * a.n + i];
    keep = in_band(f, a.threshold);
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) warp_cnt[warp] = __popc(ballot);
  __syncthreads();
  if (!keep) return;
  int off = block_base + __popc(ballot & ((1u << lane) - 1u));
  for (int w = 0; w < warp; ++w) off += warp_cnt[w];

  const float* g = a.grad + ((long long)b * a.n + i) * a.grad_stride;
  const float gx = g[a.grad_col], gy = g[a.grad_col + 1], gz = g[a.grad_col + 2];
  // normals /= ||normals||  (grid.py:57-58); points - sdf*normals (grid.py:61), unfused like torch
  const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz)));
  const float nx = gx / nrm, ny = gy / nrm, nz = gz / nrm;
  float px, py, pz;
  if (a.points) {
    px = a.points[i * 3]; py = a.points[i * 3 + 1]; pz = a.points[i * 3 + 2];
  } else {
    lattice_point(a.lattice, i, px, py, pz);
  }
  const long long o = (long long)b * a.cap + off;
  a.out_pts[o * 3 + 0] = __fsub_rn(px, __fmul_rn(f, nx));
  a.out_pts[o * 3 + 1] = __fsub_rn(py, __fmul_rn(f, ny));
  a.out_pts[o * 3 + 2] = __fsub_rn(pz, __fmul_rn(f, nz));
  a.out_nrm[o * 3 + 0] = nx;
  a.out_nrm[o * 3 + 1] = ny;
  a.out_nrm[o * 3 + 2] = nz;
  if (a.out_idx) a.out_idx[o] = (int)i;
  if (a.out_glat)
    for (int c = 0; c < a.glat_dim; ++c) a.out_glat[o * a.glat_dim + c] = g[c];
}

// ---- band-restricted flow (fused engine) -------------------------------------------------------
__global__ void __launch_bounds__(SB) band_count2_kernel(BandArgs a, int nblocks) {
  const int b = blockIdx.y;
  const long long i = (long long)blockIdx.x * SB + threadIdx.x;
  const bool keep = i < a.n && in_band(a.sdf[(long long)b * a.n + i], a.threshold);
  const int c = __syncthreads_count(keep);
  if (threadIdx.x == 0) a.block_counts[(long long)b * nblocks + blockIdx.x] = c;
}

// one block: exclusive scan of the batch*nblocks block counts (detection-major)
__global__ void __launch_bounds__(1024) band_prefix_kernel(BandArgs a, int nblocks) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int total_blocks = a.batch * nblocks;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < total_blocks; base += 1024) {
    const int i = base + tid;
    const int v = i < total_blocks ? a.block_counts[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      s_warp[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    const int excl = s_carry + (warp ? s_warp[warp - 1] : 0) + incl - v;
    if (i < total_blocks) {
      a.block_prefix[i] = excl;
      if (i % nblocks == 0) a.det_start[i / nblocks] = excl;
    }
    __syncthreads();
    if (tid == 1023) s_carry = excl + v;
    __syncthreads();
  }
  if (tid == 0) *a.total = s_carry;
  __syncthreads();
  const int tot = s_carry;
  for (int b = tid; b < a.batch; b += 1024) {
    const int end = b + 1 < a.batch ? a.block_prefix[(b + 1) * nblocks] : tot;
    a.det_count[b] = end - a.block_prefix[b * nblocks];
  }
}

__global__ void __launch_bounds__(SB) band_index_kernel(BandArgs a, int nblocks) {
  __shared__ int warp_cnt[SB / 32];
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long i = (long long)blockIdx.x * SB + tid;
  const bool keep = i < a.n && in_band(a.sdf[(long long)b * a.n + i], a.threshold);
  const unsigned ballot = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) warp_cnt[warp] = __popc(ballot);
  __syncthreads();
  if (!keep) return;
  int off = a.block_prefix[(long long)b * nblocks + blockIdx.x] + __popc(ballot & ((1u << lane) - 1u));
  for (int w = 0; w < warp; ++w) off += warp_cnt[w];
  a.band_src[off] = (int)((long long)b * a.n + i);
}

__global__ void __launch_bounds__(256) band_surface_kernel(BandArgs a) {
  const int total = *a.total;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= total) return;
  const long long src = a.band_src[j];
  const int b = (int)(src / a.n);
  const long long k = src - (long long)b * a.n;
  const int local = j - a.det_start[b];
  const float f = a.band_sdf[j];
  const float* g = a.band_dinput + (size_t)j * a.in0;
  const float gx = g[a.latent], gy = g[a.latent + 1], gz = g[a.latent + 2];
  const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz)));
  const float nx = gx / nrm, ny = gy / nrm, nz = gz / nrm;
  float px, py, pz;
  lattice_point(a.lattice, k, px, py, pz);
  const long long o = (long long)b * a.cap + local;
  a.out_pts[o * 3 + 0] = __fsub_rn(px, __fmul_rn(f, nx));
  a.out_pts[o * 3 + 1] = __fsub_rn(py, __fmul_rn(f, ny));
  a.out_pts[o * 3 + 2] = __fsub_rn(pz, __fmul_rn(f, nz));
  a.out_nrm[o * 3 + 0] = nx;
  a.out_nrm[o * 3 + 1] = ny;
  a.out_nrm[o * 3 + 2] = nz;
  if (a.out_idx) a.out_idx[o] = (int)k;
  if (a.out_glat)
    for (int c = 0; c < a.latent; ++c) a.out_glat[o * a.latent + c] = g[c];
  if (a.out_valid) a.out_valid[o] = in_band(f, a.final_threshold) ? 1 : 0;
  if (a.presel_err) {
    // measured error of the coarse lattice pass on the rows it pre-selected (all of them lie near the band,
    // the only place where that error can change the result); non-negative floats order like their bit patterns
    const unsigned live = __activemask();         // the warp's threads past `total` have returned
    const unsigned bits = __reduce_max_sync(live, __float_as_uint(fabsf(a.sdf[src] - f)));
    if ((threadIdx.x & 31) == (__ffs(live) - 1) && bits) atomicMax(a.presel_err, (int)bits);
  }
}

}  // namespace

int launch_band_select(const BandArgs& a, cudaStream_t s) {
  if (a.n <= 0 || a.batch <= 0) return SDFR_OK;
  const int nblocks = (int)((a.n + SB - 1) / SB);
  dim3 grid(nblocks, a.batch);
  band_count2_kernel<<<grid, SB, 0, s>>>(a, nblocks);
  SDFR_LAUNCH_CHECK();
  band_prefix_kernel<<<1, 1024, 0, s>>>(a, nblocks);
  SDFR_LAUNCH_CHECK();
  band_index_kernel<<<grid, SB, 0, s>>>(a, nblocks);
  SDFR_LAUNCH_CHECK();
  return SDFR_OK;
}

int launch_band_surface(const BandArgs& a, cudaStream_t s) {
  const long long capacity = a.n * a.batch;
  if (capacity <= 0) return SDFR_OK;
  band_surface_kernel<<<(unsigned)((capacity + 255) / 256), 256, 0, s>>>(a);
  SDFR_LAUNCH_CHECK();
  return SDFR_OK;
}

int launch_lattice_points(int density, float* pts, cudaStream_t s) {
  const long long n = (long long)density * density * density;
  lattice_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(make_lattice(density), n, pts);
  SDFR_LAUNCH_CHECK();
  return SDFR_OK;
}

int launch_surface_extract(const SurfaceArgs& a, cudaStream_t s) {
  if (a.n <= 0 || a.batch <= 0) return SDFR_OK;
  const int nblocks = (int)((a.n + SB - 1) / SB);
  dim3 grid(nblocks, a.batch);
  band_count_kernel<<<grid, SB, 0, s>>>(a, nblocks);
  SDFR_LAUNCH_CHECK();
  band_scatter_kernel<<<grid, SB, 0, s>>>(a, nblocks);
  SDFR_LAUNCH_CHECK();
  return SDFR_OK;
}

}  // namespace sdfr
