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
#include "project.cuh"

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
// Order-preserving selection of |sdf| < thr over all detections in ONE launch: a chained scan with decoupled
// look-back over 1024-point chunks (detection-major).  Chunks are handed out by an atomic ticket, so a block only
// ever waits for chunks that started before it.  A chunk publishes its count as soon as it is known and its
// inclusive prefix once its predecessors are summed; the status word carries the launch epoch, so nothing has to be
// cleared between launches (the last block of a launch resets the ticket and advances the epoch).
constexpr unsigned long long kStateAggregate = 1ull, kStateInclusive = 2ull;

__device__ __forceinline__ unsigned long long status_word(unsigned epoch, unsigned long long state, int value) {
  return ((unsigned long long)(epoch & 0x3fffffffu) << 34) | (state << 32) | (unsigned long long)(unsigned)value;
}
__device__ __forceinline__ unsigned long long status_load(const unsigned long long* p) {
  unsigned long long w;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return w;
}
__device__ __forceinline__ void status_store(unsigned long long* p, unsigned long long w) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
// spins until chunk j has published something for this epoch
__device__ __forceinline__ unsigned long long status_wait(const unsigned long long* status, int j, unsigned epoch) {
  unsigned long long w;
  do { w = status_load(status + j); } while ((unsigned)(w >> 34) != (epoch & 0x3fffffffu));
  return w;
}

__global__ void __launch_bounds__(SB) select_kernel(SelectArgs a, int nblocks) {
  __shared__ int warp_off[SB / 32];
  __shared__ int s_chunk, s_base, s_total;
  __shared__ unsigned s_epoch;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int total_blocks = a.batch * nblocks;
  if (tid == 0) {
    s_chunk = atomicAdd(&a.ctrl[0], 1);
    s_epoch = *reinterpret_cast<volatile unsigned*>(&a.ctrl[2]) + 1u;    // never 0: fresh status words match no launch
  }
  __syncthreads();
  const int chunk = s_chunk;
  const unsigned epoch = s_epoch;
  const int b = chunk / nblocks, blk = chunk - b * nblocks;
  const long long i = (long long)blk * SB + tid;
  // rows of detection b: the whole lattice [b n, b n + n), or its slice of a compact list
  const long long rows = a.in_count ? (long long)a.in_count[b] : a.n;
  const long long row0 = a.in_start ? (long long)a.in_start[b] : (long long)b * a.n;
  const bool all = a.det_all && a.det_all[b];
  const float thr = a.det_threshold ? a.det_threshold[b] : a.threshold;
  bool keep = false;
  int src = 0;
  if (i < rows) {
    const float v = a.values[row0 + i];
    src = a.in_src ? a.in_src[row0 + i] : (int)(row0 + i);
    keep = all || in_band(v, thr);
    if (a.scatter_values) a.scatter_values[src] = v;
    if (a.scatter_ref && a.scatter_flag[b]) a.scatter_ref[src] = v;
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) warp_off[warp] = __popc(ballot);
  __syncthreads();
  if (warp == 0) {
    const int c = warp_off[lane];
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    warp_off[lane] = incl - c;                                 // exclusive offset of each warp inside the chunk
    if (lane == 0) status_store(a.status + chunk, status_word(epoch, chunk == 0 ? kStateInclusive : kStateAggregate, total));
    int excl = 0;
    if (chunk > 0) {
      for (int pos = chunk - 1;; pos -= 32) {                  // 32 predecessors per round, nearest first
        const int j = pos - lane;
        unsigned long long w = status_word(epoch, kStateInclusive, 0);      // before the first chunk: prefix 0
        if (j >= 0) w = status_wait(a.status, j, epoch);
        const bool inclusive = ((w >> 32) & 3ull) == kStateInclusive;
        const unsigned have = __ballot_sync(0xffffffffu, inclusive);
        const int first = have ? __ffs(have) - 1 : 31;         // sum up to the nearest inclusive prefix
        int v = lane <= first ? (int)(unsigned)w : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        excl += v;
        if (have) break;
      }
      if (lane == 0) status_store(a.status + chunk, status_word(epoch, kStateInclusive, excl + total));
    }
    if (lane == 0) { s_base = excl; s_total = total; }
  }
  __syncthreads();
  const int base = s_base;
  if (keep) a.out_src[base + warp_off[warp] + __popc(ballot & ((1u << lane) - 1u))] = src;
  if (tid == 0) {
    if (blk == nblocks - 1) {                                  // the detection's last chunk: its start and count
      int start = 0;
      if (b > 0) {                                             // inclusive prefix of the previous detection's last chunk
        unsigned long long w;
        do { w = status_wait(a.status, b * nblocks - 1, epoch); } while (((w >> 32) & 3ull) != kStateInclusive);
        start = (int)(unsigned)w;
      }
      a.det_start[b] = start;
      a.det_count[b] = base + s_total - start;
      if (a.scatter_done && a.scatter_flag[b]) a.scatter_done[b] = 1;
      if (chunk == total_blocks - 1) {
        *a.total = base + s_total;
        if (a.total_accum) {
          atomicAdd(a.total_accum, (unsigned long long)(base + s_total));
          atomicAdd(a.total_accum + 1, (unsigned long long)a.batch);
        }
      }
    }
    __threadfence();
    if (atomicAdd(&a.ctrl[1], 1) == total_blocks - 1) {        // every block of the launch is past its look-back
      a.ctrl[0] = 0;
      a.ctrl[1] = 0;
      a.ctrl[2] = (int)(epoch & 0x3fffffffu) == 0x3fffffff ? 0 : (int)(epoch & 0x3fffffffu);
      __threadfence();
    }
  }
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
  const bool is_surfel = in_band(f, a.final_threshold);
  if (a.out_valid) a.out_valid[o] = is_surfel ? 1 : 0;
  // camera-space surfel of the detection's view (the engine's next stage), from the values just stored
  if (a.views)
    project_surfel(a.views[b], local, __fsub_rn(px, __fmul_rn(f, nx)), __fsub_rn(py, __fmul_rn(f, ny)),
                   __fsub_rn(pz, __fmul_rn(f, nz)), nx, ny, nz, is_surfel);
  if (a.presel_err) {
    // measured error of the coarse lattice pass on the rows it pre-selected (all of them lie near the band,
    // the only place where that error can change the result); non-negative floats order like their bit patterns
    const unsigned live = __activemask();         // the warp's threads past `total` have returned
    const unsigned bits = __reduce_max_sync(live, __float_as_uint(fabsf(a.sdf[src] - f)));
    if ((threadIdx.x & 31) == (__ffs(live) - 1) && bits) atomicMax(a.presel_err, (int)bits);
  }
}

}  // namespace

int launch_select(const SelectArgs& a, cudaStream_t s) {
  if (a.n <= 0 || a.batch <= 0) return SDFR_OK;
  const int nblocks = (int)((a.n + SB - 1) / SB);
  select_kernel<<<nblocks * a.batch, SB, 0, s>>>(a, nblocks);
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
