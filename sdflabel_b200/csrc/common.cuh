// Shared declarations for libsdfr.so (sm_100a).  Host glue lives in api.cu;
// the kernels are split by stage of the hot path.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/sdfr.h"

namespace sdfr {

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern thread_local long long g_launches_tls;
void count_launch(int n = 1);

#define SDFR_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::sdfr::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,               \
                        cudaGetErrorString(_e));                                          \
      return SDFR_E_CUDA;                                                                 \
    }                                                                                     \
  } while (0)

#define SDFR_LAUNCH_CHECK()                                                               \
  do {                                                                                    \
    ::sdfr::count_launch();                                                               \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) {                                                              \
      ::sdfr::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,          \
                        cudaGetErrorString(_e));                                          \
      return SDFR_E_CUDA;                                                                 \
    }                                                                                     \
  } while (0)

#define SDFR_REQUIRE(cond, code, ...)                                                     \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      ::sdfr::set_error(__VA_ARGS__);                                                     \
      return (code);                                                                      \
    }                                                                                     \
  } while (0)

// ---------------------------------------------------------------------------
// constants of the reference
// ---------------------------------------------------------------------------
constexpr float kDiscRadius = 0.04f;   // rasterer.py:103
constexpr float kDepthGain = 150.0f;   // primitives.py:172
constexpr float kRayCutoff = 0.01f;    // primitives.py:210
constexpr float kEps32 = 1.1920928955078125e-07f;  // torch.finfo(float32).eps
constexpr float kWinRadius = 5.0f;     // optimizer.py:200
constexpr float kNocsThr = 1.0f;       // optimizer.py:200
constexpr int kMaxWidthFFMA = 512;     // widest layer the kernels accept
// Margin added to the band threshold (grid.py:43, 0.03) when the lattice pass runs at fp16 operand precision:
// the accurate second pass sees every true band point as long as the coarse error stays below it.  The engine
// measures that error on every pre-selected row and refuses results when it exceeds half the margin.
constexpr float kPreselectMargin = 0.005f;

// ---------------------------------------------------------------------------
// Decoder (device resident)
// ---------------------------------------------------------------------------
constexpr int kMaxLayers = 16;

struct LayerDev {
  int in_dim, out_dim;     // logical fan-in (after concat) / fan-out
  int in_pad, out_pad;     // padded to multiples of 4
  int concat;              // 0 none, 1 input, 2 xyz (concatenated BEFORE this layer)
  int layer_norm;
  const float* wt;         // [in_pad][out_pad]  = W^T, zero padded (forward operand)
  const float* w;          // [out_pad][in_pad]  = W,   zero padded (backward operand)
  const float* bias;       // [out_pad]
  const float* ln_w;       // [out_pad] or null
  const float* ln_b;
};

struct DecoderDev {
  int latent_size, in0;    // in0 = L + 3
  int num_layers;
  int use_tanh;
  int max_width;           // max over in_pad / out_pad
  LayerDev layer[kMaxLayers];
};

// tcgen05 path: one entry per GEMM pass (8 forward + 1 + 7 backward for stock)
struct TcPass {
  int kind;        // 0 fwd hidden, 1 fwd last, 2 bwd hidden, 3 bwd first
  int layer;       // Linear index
  int m_blocks;    // ceil(rows/128)
  int k_chunks;    // ceil(K/32)
  int rows, kdim;  // logical rows (outputs of this pass) / reduction length
  long long tile_offset;   // first weight tile (in tiles) inside the packed array
};

struct DecoderTc {
  int ok;
  int num_passes;
  TcPass pass[2 * kMaxLayers];
  const uint4* tiles;      // packed fp16 hi/lo weight tiles
  long long num_tiles;
};

}  // namespace sdfr

struct sdfr_decoder {
  sdfr::DecoderDev dev;          // host copy of the device table (pointers are device pointers)
  sdfr::DecoderDev* dev_ptr;     // same table in device memory
  sdfr::DecoderTc tc;            // host copy
  sdfr::DecoderTc* tc_ptr;
  std::vector<void*> allocs;
  int device;
  float* scratch;                // LayerNorm x_hat spill (per resident CTA)
  size_t scratch_bytes;
  int sm_count;
};

namespace sdfr {

// ---------------------------------------------------------------------------
// Lattice (grid.py:22-41), shared by surface.cu and the MLP kernels
// ---------------------------------------------------------------------------
struct LatticeParams {
  int density;
  double step;    // 2/(D-1)
  double shift;   // (max-min)/D/2
};

inline LatticeParams make_lattice(int density) {
  LatticeParams lp;
  lp.density = density;
  lp.step = 2.0 / (double)(density - 1);
  double amax = (double)(density - 1) * lp.step + (-1.0);
  lp.shift = (amax - (-1.0)) / (double)density / 2.0;
  return lp;
}

// regular lattice of d^3 nodes over [-1, hi]^3 (no half-cell shift): the distance cache of trace mode
inline LatticeParams make_regular_lattice(int density, double hi) {
  LatticeParams lp;
  lp.density = density;
  lp.step = (hi + 1.0) / (double)(density - 1);
  lp.shift = 0.0;
  return lp;
}

__device__ __forceinline__ void lattice_point(const LatticeParams& lp, long long idx, float& x, float& y,
                                              float& z) {
  const int d = lp.density;
  int iz = (int)(idx % d);
  long long r = idx / d;
  int iy = (int)(r % d);
  int ix = (int)(r / d);
  // numpy: arange(d) * step + start, then += shift on odd rows, then float32 cast
  double dx = __dadd_rn(__dmul_rn((double)ix, lp.step), -1.0);
  double dy = __dadd_rn(__dmul_rn((double)iy, lp.step), -1.0);
  double dz = __dadd_rn(__dmul_rn((double)iz, lp.step), -1.0);
  if (idx & 1) {
    dx = __dadd_rn(dx, lp.shift);
    dy = __dadd_rn(dy, lp.shift);
  }
  x = (float)dx;
  y = (float)dy;
  z = (float)dz;
}

// ---------------------------------------------------------------------------
// Splat views: one per detection, resident in device memory
// ---------------------------------------------------------------------------
struct SplatView {
  // configuration
  int width, height;
  float kinv[9];
  float k[9];
  int rot;           // SDFR_ROT_*
  int output_nocs;
  int primitive;     // SDFR_PRIM_*
  int has_bg;
  const float* bg;   // [3,P] background image or null
  // inputs
  const float* coords;    // [m,3] object-frame surfel centres
  const float* normals;   // [m,3]
  const float* colors;    // [m,3] or null
  const float* pose;      // 16 (dcm) or 7 (quat) floats
  const unsigned char* valid;   // optional per-surfel flag: 0 = not a surfel (skipped by every stage)
  const int* count;       // surfel count on the device (null -> static_count)
  int static_count;
  int capacity;
  // per-surfel state
  float* cam_v;      // [cap,3]
  float* cam_m;      // [cap,3]
  float* cam_c;      // [cap,3] colour as composited ((c+1)/2 when output_nocs)
  float* plane_a;    // [cap]   n.v
  int* bbox;         // [cap,4] x0,y0,x1,y1 inclusive, clipped; x1<x0 -> empty
  unsigned char* front;  // [cap]
  float* cam_rgb;    // [cap,3] points['rgb'] (optional)
  // circle primitives (circle.cu): per-point score / 2D position / pixel radius, scalars, score gradient
  float* score;      // [cap]
  float* p2;         // [cap,2]
  float* radius;     // [cap]
  float* prim_scalars;   // [4] nu, background score, arg-min point, d background score
  float* d_score;    // [cap]
  // per-pixel outputs (unclamped sums are kept in ws_* for the backward)
  float* color;      // [3,P]
  float* mask;       // [1,P]
  float* depth;      // [1,P]
  float* nmap;       // [3,P]
  float* pix_stat;   // [P,4]  nu, smax, 1/den, hits
  float* pix_raw;    // [P,8]  unclamped colour(3), mask, depth, normals(3)
  float* pix_grad;   // [P,12] g_colour'(3), g_depth, g_normals'(3), g_mask', G, pad(3)
  // per-surfel gradients
  float* d_v;        // [cap,3]
  float* d_m;        // [cap,3]
  float* d_c;        // [cap,3] wrt composited colour
};

int launch_project(const SplatView* views_dev, int batch, int max_count, cudaStream_t s);
int launch_splat_forward(const SplatView* views_dev, int batch, int max_w, int max_h, int any_fine, cudaStream_t s);
constexpr int kFineCropPixelsHost = 64 * 64;   // = kFineCropPixels of splat.cu: crops up to this size use 4 x 4 pixel tiles
int launch_pixel_grad_prep(const SplatView* views_dev, int batch, int max_pixels, const float* g_color,
                           const float* g_mask, const float* g_depth, const float* g_nmap, cudaStream_t s);
int launch_splat_backward(const SplatView* views_dev, int batch, int max_count, cudaStream_t s);
// circle.cu
int launch_circle_forward(const SplatView* views_dev, int batch, int max_pixels, cudaStream_t s);
int launch_circle_backward(const SplatView* views_dev, int batch, int max_count, int has_bg, cudaStream_t s);
int launch_disc_background(const SplatView* views_dev, int batch, int max_pixels, cudaStream_t s);

// ---------------------------------------------------------------------------
// MLP launchers (mlp_ffma.cu / mlp_tc.cu)
// ---------------------------------------------------------------------------
struct RayMarch;   // trace.cuh

struct MlpInputs {
  const float* inputs;        // explicit [n, L+3] or null
  const float* latent_unit;   // [batch, L] for lattice mode
  LatticeParams lattice;
  long long points_per_batch; // D^3 in lattice mode
  long long n;                // total points (capacity when count_dev is set)
  const int* index;           // optional gather: row r evaluates source point index[r] (null = r)
  const int* count_dev;       // optional: number of rows, read on the device (no host sync)
  int small_tiles;            // hint: the row list is short, use 16-point tiles in the tensor-core kernel
  // tensor-core kernel: ReLU sign scratch of this launch (null = the decoder's own).  Launches that may run
  // concurrently on different streams (two engines on one decoder) must not share it.
  unsigned long long* mask_scratch = nullptr;
  // lattice-pass kernel only: march mode (trace.cuh).  The rows are the rays of march->list (count_dev = march->count),
  // and instead of writing sdf the epilogue advances every ray and sorts it into the next / near lists.
  const RayMarch* march = nullptr;
  // CTA-pair band kernels: the launch runs only when count_lo < *count_dev <= count_hi (adaptive_tiles: launch_mlp_tc
  // enqueues the 16- and the 64-point tiling with complementary windows)
  int count_lo = 0, count_hi = 0x7fffffff;
  int adaptive_tiles = 0;
};

__device__ __forceinline__ long long mlp_rows(const MlpInputs& in) {
  if (!in.count_dev) return in.n;
  const long long c = (long long)*in.count_dev;
  return c < in.n ? c : in.n;
}

int launch_mlp_ffma(const sdfr_decoder* dec, const MlpInputs& in, float* sdf, float* dinput, cudaStream_t s);
int launch_mlp_tc(const sdfr_decoder* dec, const MlpInputs& in, float* sdf, float* dinput, cudaStream_t s);
// forward only, fp16 operand precision (hi halves only): the band pre-selection pass of the fused engine
int launch_mlp_tc_coarse(const sdfr_decoder* dec, const MlpInputs& in, float* sdf, cudaStream_t s);
bool mlp_tc_march_ok(const sdfr_decoder* dec);   // the lattice-pass kernel can run trace mode's fused march for this decoder
int mlp_tc_round_rows(const sdfr_decoder* dec);  // rows of one round of the lattice-pass grid (0: no pair kernel)
size_t mlp_tc_mask_scratch_bytes(const sdfr_decoder* dec);
int build_tc_tables(sdfr_decoder* dec, const sdfr_decoder_spec* spec, const float* const* weights_host);
void free_tc_tables(sdfr_decoder* dec);
int tc_overflow_flag(const sdfr_decoder* dec, int* flag);
int tc_overflow_flag_enqueue(const sdfr_decoder* dec, int* flag_host, cudaStream_t s);
int tc_overflow_reset(const sdfr_decoder* dec, cudaStream_t s);
const int* tc_overflow_ptr(const sdfr_decoder* dec);   // the device flag itself (null: no tensor-core tables)

// surface.cu
int launch_lattice_points(int density, float* pts, cudaStream_t s);
struct SurfaceArgs {
  const float* points;      // explicit [n,3] or null -> lattice
  LatticeParams lattice;
  const float* sdf;         // [batch, n]
  const float* grad;        // [batch, n, grad_stride]
  int grad_stride, grad_col;
  long long n;              // points per detection
  int batch;
  float threshold;
  float* out_pts;           // [batch, cap, 3]
  float* out_nrm;
  int* out_idx;
  float* out_glat;          // optional [batch, cap, L] gather of grad[:, 0:L]
  int glat_dim;
  long long cap;            // rows per detection in the outputs
  int* out_count;           // [batch]
  int* scratch;             // [batch, blocks+1]
};
int launch_surface_extract(const SurfaceArgs& a, cudaStream_t s);

// Band-restricted flow of the fused engine: (1) select |sdf| < thr over all detections into one
// compact source-index list, (2) evaluate the decoder with its input gradient on that list only,
// (3) project the band points onto the zero isosurface.
// Order-preserving selection |value| < threshold over the rows of every detection, ONE launch (chained scan with
// decoupled look-back over 1024-row chunks, detection-major).  The rows of detection b are either the whole
// lattice slice values[b n .. b n + n) or, with in_start / in_count / in_src, its slice of a compact list whose
// entries carry their global lattice index.
struct SelectArgs {
  const float* values;
  const int* in_src;        // optional: global source index of each compact row
  const int* in_start;      // optional [batch]: first compact row of each detection
  const int* in_count;      // optional [batch]: compact rows of each detection (<= n)
  long long n;              // lattice points per detection (chunk layout: ceil(n / 1024) chunks per detection)
  int batch;
  float threshold;
  const float* det_threshold;   // optional [batch]: per-detection threshold
  const int* det_all;       // optional [batch]: != 0 selects every row of the detection
  int* out_src;             // [batch * n] global source indices of the selected rows, ascending
  int* det_start;           // [batch]
  int* det_count;           // [batch]
  int* total;               // [1]
  unsigned long long* status;   // [batch * nblocks] status words (epoch | state | value), zero-initialised
  int* ctrl;                // [4] ticket, finished blocks, epoch, pad: zero-initialised, self-resetting
  float* scatter_values;    // optional: scatter_values[src] = value of every visited row
  float* scatter_ref;       // optional: scatter_ref[src] = value for the detections with scatter_flag[b] != 0
  const int* scatter_flag;
  unsigned long long* total_accum;   // optional [2]: += selected rows, += batch (running totals over launches)
  int* scatter_done;        // optional [batch]: set to 1 for the flagged detections (their scatter_ref slice is complete after the launch)
};
int launch_select(const SelectArgs& a, cudaStream_t s);

struct BandArgs {
  LatticeParams lattice;
  const float* sdf;         // [batch, n] coarse sdf by global lattice index (valid at least at the selected rows)
  long long n;              // lattice points per detection
  int batch;
  float threshold;
  int* det_start;           // [batch]
  int* det_count;           // [batch]  (also the surfel count the splat stages read)
  int* total;               // [1]
  int* band_src;            // [batch * n] global source index b*n + k, ascending
  // after the band evaluation
  const float* band_sdf;    // [total]
  const float* band_dinput; // [total, in0]
  int in0, latent;
  float* out_pts;           // [batch, cap, 3]
  float* out_nrm;
  int* out_idx;
  float* out_glat;          // [batch, cap, latent]
  unsigned char* out_valid; // [batch, cap] 1 where |band_sdf| < final_threshold (the true band)
  float final_threshold;
  long long cap;
  int* presel_err;          // optional [1]: running max of |sdf[src] - band_sdf| (float bits, atomicMax)
  const SplatView* views;   // optional [batch]: the isosurface kernel also projects each surfel into its view
};
int launch_band_surface(const BandArgs& a, cudaStream_t s);

// loss.cu
int launch_loss3d_standalone(const float* xyzf, long long q, const float* lidar, long long nl, double radius,
                             float* loss, float* d_xyzf, float* d_lidar, cudaStream_t s);
int launch_loss2d_standalone(const float* color, const float* target, int h, int w, float* loss,
                             float* d_color, cudaStream_t s);

}  // namespace sdfr
