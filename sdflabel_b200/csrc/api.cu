// C ABI glue of libsdfr.so (see include/sdfr.h for the contract and the
// reference file:line each entry point replaces).
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstring>

#include "common.cuh"

namespace sdfr {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};
thread_local long long g_launches_tls = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

namespace {

inline int pad4(int v) { return (v + 3) & ~3; }

int upload(sdfr_decoder* d, const std::vector<float>& host, const float** out) {
  void* p = nullptr;
  SDFR_CUDA(cudaMalloc(&p, std::max<size_t>(host.size(), 4) * sizeof(float)));
  d->allocs.push_back(p);
  if (!host.empty()) SDFR_CUDA(cudaMemcpy(p, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
  *out = reinterpret_cast<const float*>(p);
  return SDFR_OK;
}

// ---- stand-alone splat workspace ------------------------------------------------------------
struct SplatWs {
  size_t off_view, off_count, off_v, off_m, off_c, off_a, off_bbox, off_front, off_dv, off_dm, off_dc, off_stat,
      off_raw, off_grad, off_dpose, off_score, off_p2, off_radius, off_scalars, off_dscore, total;
};

SplatWs splat_ws_layout(int64_t m, int64_t P) {
  SplatWs w;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 255) & ~(size_t)255; return r; };
  const size_t mm = (size_t)std::max<int64_t>(m, 1), pp = (size_t)std::max<int64_t>(P, 1);
  w.off_view = take(sizeof(SplatView));
  w.off_count = take(64);
  w.off_v = take(mm * 12); w.off_m = take(mm * 12); w.off_c = take(mm * 12); w.off_a = take(mm * 4);
  w.off_bbox = take(mm * 16); w.off_front = take(mm);
  w.off_dv = take(mm * 12); w.off_dm = take(mm * 12); w.off_dc = take(mm * 12);
  w.off_stat = take(pp * 16); w.off_raw = take(pp * 32); w.off_grad = take(pp * 48);
  w.off_dpose = take(64);
  w.off_score = take(mm * 4); w.off_p2 = take(mm * 8); w.off_radius = take(mm * 4); w.off_scalars = take(64);
  w.off_dscore = take(mm * 4);
  w.total = o;
  return w;
}

SplatView make_view(const sdfr_raster_cfg* cfg, const float* coords, const float* normals, const float* colors,
                    const float* pose, int64_t m, char* ws, const SplatWs& L) {
  SplatView V;
  memset(&V, 0, sizeof(V));
  V.width = cfg->width; V.height = cfg->height;
  memcpy(V.kinv, cfg->kinv, sizeof(V.kinv));
  memcpy(V.k, cfg->k, sizeof(V.k));
  V.rot = cfg->rot; V.output_nocs = cfg->output_nocs;
  V.primitive = cfg->primitive; V.bg = cfg->bg_dev; V.has_bg = cfg->bg_dev ? 1 : 0;
  V.coords = coords; V.normals = normals; V.colors = colors; V.pose = pose;
  V.count = nullptr; V.static_count = (int)m; V.capacity = (int)m;
  V.cam_v = reinterpret_cast<float*>(ws + L.off_v);
  V.cam_m = reinterpret_cast<float*>(ws + L.off_m);
  V.cam_c = reinterpret_cast<float*>(ws + L.off_c);
  V.plane_a = reinterpret_cast<float*>(ws + L.off_a);
  V.bbox = reinterpret_cast<int*>(ws + L.off_bbox);
  V.front = reinterpret_cast<unsigned char*>(ws + L.off_front);
  V.d_v = reinterpret_cast<float*>(ws + L.off_dv);
  V.d_m = reinterpret_cast<float*>(ws + L.off_dm);
  V.d_c = reinterpret_cast<float*>(ws + L.off_dc);
  V.pix_stat = reinterpret_cast<float*>(ws + L.off_stat);
  V.pix_raw = reinterpret_cast<float*>(ws + L.off_raw);
  V.pix_grad = reinterpret_cast<float*>(ws + L.off_grad);
  V.score = reinterpret_cast<float*>(ws + L.off_score);
  V.p2 = reinterpret_cast<float*>(ws + L.off_p2);
  V.radius = reinterpret_cast<float*>(ws + L.off_radius);
  V.prim_scalars = reinterpret_cast<float*>(ws + L.off_scalars);
  V.d_score = reinterpret_cast<float*>(ws + L.off_dscore);
  return V;
}

// ordered compaction of the front-facing surfels (projection.py:61-70 masked_select order)
__global__ void __launch_bounds__(1024) compact_front_kernel(const SplatView* __restrict__ views,
                                                             const float* __restrict__ cam_rgb,
                                                             float* __restrict__ xyzf, float* __restrict__ rgbf,
                                                             int* __restrict__ count_out) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const SplatView& V = views[0];
  const int m = V.static_count;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int base = 0; base < m; base += 1024) {
    const int i = base + tid;
    const bool keep = i < m && V.front[i];
    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    int off = s_base + __popc(ballot & ((1u << lane) - 1u));
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (keep) {
      for (int c = 0; c < 3; ++c) {
        if (xyzf) xyzf[off * 3 + c] = V.cam_v[i * 3 + c];
        if (rgbf) rgbf[off * 3 + c] = cam_rgb[i * 3 + c];
      }
    }
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < 32; ++w) t += s_warp[w];
      s_base += t;
    }
    __syncthreads();
  }
  if (tid == 0 && count_out) *count_out = s_base;
}

// d loss / d (coords, normals, colours, pose) from the per-surfel camera-space gradients
__global__ void __launch_bounds__(256) pose_chain_kernel(const SplatView* __restrict__ views,
                                                         const float* __restrict__ g_cam_pts,
                                                         const float* __restrict__ g_cam_rgb,
                                                         float* __restrict__ d_coords, float* __restrict__ d_normals,
                                                         float* __restrict__ d_colors, float* __restrict__ d_pose) {
  __shared__ float s_red[8][12];
  const SplatView& V = views[0];
  const int m = V.static_count;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = 0.f;
  if (i < m) {
    float dv[3], dm[3], dc[3], p[3], n[3];
    for (int c = 0; c < 3; ++c) {
      dv[c] = V.d_v[i * 3 + c] + (g_cam_pts ? g_cam_pts[i * 3 + c] : 0.f);
      dm[c] = V.d_m[i * 3 + c];
      dc[c] = (V.output_nocs ? 0.5f * V.d_c[i * 3 + c] : V.d_c[i * 3 + c]) + (g_cam_rgb ? 0.5f * g_cam_rgb[i * 3 + c] : 0.f);
      p[c] = V.coords[i * 3 + c];
      n[c] = V.normals[i * 3 + c];
    }
    float dp[3] = {0.f, 0.f, 0.f}, dn[3];
    if (V.output_nocs) {
      dp[0] = (V.rot == SDFR_ROT_DCM ? -dc[0] : dc[0]); dp[1] = dc[1]; dp[2] = dc[2];
    } else if (d_colors) {
      for (int c = 0; c < 3; ++c) d_colors[i * 3 + c] = dc[c];
    }
    float R[9];
    if (V.rot == SDFR_ROT_DCM) {
      for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[r * 3 + c] = V.pose[r * 4 + c];
    } else {
      const float w = V.pose[0], x = V.pose[1], y = V.pose[2], z = V.pose[3];
      R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - w * z);       R[2] = 2.f * (x * z + w * y);
      R[3] = 2.f * (x * y + w * z);       R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - w * x);
      R[6] = 2.f * (x * z - w * y);       R[7] = 2.f * (y * z + w * x);       R[8] = 1.f - 2.f * (x * x + y * y);
    }
    for (int c = 0; c < 3; ++c) {
      dp[c] += R[0 * 3 + c] * dv[0] + R[1 * 3 + c] * dv[1] + R[2 * 3 + c] * dv[2];
      dn[c] = R[0 * 3 + c] * dm[0] + R[1 * 3 + c] * dm[1] + R[2 * 3 + c] * dm[2];
    }
    for (int c = 0; c < 3; ++c) {
      if (d_coords) d_coords[i * 3 + c] = dp[c];
      if (d_normals) d_normals[i * 3 + c] = dn[c];
    }
    if (V.rot == SDFR_ROT_DCM) {
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) acc[r * 4 + c] = dv[r] * p[c] + dm[r] * n[c];
        acc[r * 4 + 3] = dv[r];
      }
    } else {
      // v = x + 2 (w (u cross x) + u cross (u cross x)) + t, applied to x = p (gradient dv) and x = n (gradient dm)
      const float w = V.pose[0], u[3] = {V.pose[1], V.pose[2], V.pose[3]};
      const float* xs[2] = {p, n};
      const float* gs[2] = {dv, dm};
      for (int s = 0; s < 2; ++s) {
        const float* x = xs[s];
        const float* g = gs[s];
        const float uxx[3] = {u[1] * x[2] - u[2] * x[1], u[2] * x[0] - u[0] * x[2], u[0] * x[1] - u[1] * x[0]};
        const float xxg[3] = {x[1] * g[2] - x[2] * g[1], x[2] * g[0] - x[0] * g[2], x[0] * g[1] - x[1] * g[0]};
        const float ux = u[0] * x[0] + u[1] * x[1] + u[2] * x[2];
        const float ug = u[0] * g[0] + u[1] * g[1] + u[2] * g[2];
        const float xg = x[0] * g[0] + x[1] * g[1] + x[2] * g[2];
        acc[0] += 2.f * (uxx[0] * g[0] + uxx[1] * g[1] + uxx[2] * g[2]);
        for (int c = 0; c < 3; ++c) acc[1 + c] += 2.f * w * xxg[c] + 2.f * (g[c] * ux + x[c] * ug - 2.f * xg * u[c]);
      }
      for (int c = 0; c < 3; ++c) acc[4 + c] = dv[c];
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    float v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[warp][k] = v;
  }
  __syncthreads();
  const int nout = V.rot == SDFR_ROT_DCM ? 12 : 7;
  if (threadIdx.x < nout && d_pose) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_red[w][threadIdx.x];
    atomicAdd(d_pose + threadIdx.x, t);
  }
}

}  // namespace
}  // namespace sdfr

using namespace sdfr;

extern "C" {

int sdfr_version(void) { return SDFR_VERSION; }
const char* sdfr_last_error(void) { return g_err; }
int64_t sdfr_launch_count(void) { return (int64_t)g_launches.load(); }

int sdfr_caps(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return 0; }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 1;
  return 1 | ((p.major == 10) ? 2 : 0);
}

int sdfr_decoder_create(const sdfr_decoder_spec* spec, const float* const* weights_host,
                        const float* const* bias_host, const float* const* ln_weight_host,
                        const float* const* ln_bias_host, sdfr_decoder** out) {
  SDFR_REQUIRE(spec && weights_host && bias_host && out, SDFR_E_INVALID, "null argument");
  SDFR_REQUIRE(spec->num_layers >= 2 && spec->num_layers <= kMaxLayers, SDFR_E_UNSUPPORTED,
               "num_layers %d outside [2,%d]", spec->num_layers, kMaxLayers);
  SDFR_REQUIRE(spec->latent_size >= 0, SDFR_E_INVALID, "negative latent size");
  const int NL = spec->num_layers, in0 = spec->latent_size + 3;
  SDFR_REQUIRE(spec->in_dims[0] == in0, SDFR_E_INVALID, "layer 0 fan-in %d != latent+3 = %d", spec->in_dims[0], in0);
  SDFR_REQUIRE(spec->out_dims[NL - 1] == 1, SDFR_E_UNSUPPORTED, "last layer must have one output");
  SDFR_REQUIRE(spec->concat[0] == 0, SDFR_E_INVALID, "layer 0 cannot concatenate");
  for (int l = 0; l < NL; ++l) {
    SDFR_REQUIRE(spec->in_dims[l] > 0 && spec->out_dims[l] > 0, SDFR_E_INVALID, "bad dims at layer %d", l);
    SDFR_REQUIRE(pad4(spec->in_dims[l]) <= kMaxWidthFFMA && pad4(spec->out_dims[l]) <= kMaxWidthFFMA,
                 SDFR_E_UNSUPPORTED, "layer %d is %dx%d; widths above %d are not supported", l, spec->out_dims[l],
                 spec->in_dims[l], kMaxWidthFFMA);
    if (l > 0) {
      const int cat = spec->concat[l] == 1 ? in0 : spec->concat[l] == 2 ? 3 : 0;
      SDFR_REQUIRE(spec->in_dims[l] == spec->out_dims[l - 1] + cat, SDFR_E_INVALID,
                   "layer %d fan-in %d != previous fan-out %d + concat %d", l, spec->in_dims[l],
                   spec->out_dims[l - 1], cat);
    }
    if (spec->layer_norm[l]) SDFR_REQUIRE(ln_weight_host && ln_bias_host && ln_weight_host[l] && ln_bias_host[l],
                                          SDFR_E_INVALID, "layer %d needs LayerNorm parameters", l);
  }
  int ndev = 0;
  SDFR_CUDA(cudaGetDeviceCount(&ndev));
  SDFR_REQUIRE(ndev > 0, SDFR_E_CUDA, "no CUDA device: libsdfr has no CPU path");
  sdfr_decoder* d = new sdfr_decoder();
  d->dev_ptr = nullptr; d->tc_ptr = nullptr; d->scratch = nullptr; d->scratch_bytes = 0;
  cudaGetDevice(&d->device);
  cudaDeviceGetAttribute(&d->sm_count, cudaDevAttrMultiProcessorCount, d->device);
  memset(&d->dev, 0, sizeof(d->dev));
  memset(&d->tc, 0, sizeof(d->tc));
  d->dev.latent_size = spec->latent_size; d->dev.in0 = in0; d->dev.num_layers = NL; d->dev.use_tanh = spec->use_tanh;
  int maxw = pad4(in0);
  int rc = SDFR_OK;
  for (int l = 0; l < NL && rc == SDFR_OK; ++l) {
    LayerDev& Ly = d->dev.layer[l];
    Ly.in_dim = spec->in_dims[l]; Ly.out_dim = spec->out_dims[l];
    Ly.in_pad = pad4(Ly.in_dim); Ly.out_pad = pad4(Ly.out_dim);
    Ly.concat = spec->concat[l]; Ly.layer_norm = spec->layer_norm[l];
    maxw = std::max(maxw, std::max(Ly.in_pad, Ly.out_pad));
    std::vector<float> wt((size_t)Ly.in_pad * Ly.out_pad, 0.f), w((size_t)Ly.out_pad * Ly.in_pad, 0.f),
        bias(Ly.out_pad, 0.f);
    const float* W = weights_host[l];
    for (int o = 0; o < Ly.out_dim; ++o)
      for (int i = 0; i < Ly.in_dim; ++i) {
        const float v = W[(size_t)o * Ly.in_dim + i];
        wt[(size_t)i * Ly.out_pad + o] = v;
        w[(size_t)o * Ly.in_pad + i] = v;
      }
    for (int o = 0; o < Ly.out_dim; ++o) bias[o] = bias_host[l][o];
    if ((rc = upload(d, wt, &Ly.wt))) break;
    if ((rc = upload(d, w, &Ly.w))) break;
    if ((rc = upload(d, bias, &Ly.bias))) break;
    if (Ly.layer_norm) {
      std::vector<float> lw(Ly.out_pad, 0.f), lb(Ly.out_pad, 0.f);
      for (int o = 0; o < Ly.out_dim; ++o) { lw[o] = ln_weight_host[l][o]; lb[o] = ln_bias_host[l][o]; }
      if ((rc = upload(d, lw, &Ly.ln_w))) break;
      if ((rc = upload(d, lb, &Ly.ln_b))) break;
    }
  }
  d->dev.max_width = maxw;
  if (rc == SDFR_OK) {
    void* p = nullptr;
    if (cudaMalloc(&p, sizeof(DecoderDev)) != cudaSuccess ||
        cudaMemcpy(p, &d->dev, sizeof(DecoderDev), cudaMemcpyHostToDevice) != cudaSuccess) {
      set_error("decoder table upload failed: %s", cudaGetErrorString(cudaGetLastError()));
      rc = SDFR_E_CUDA;
    } else {
      d->allocs.push_back(p);
      d->dev_ptr = reinterpret_cast<DecoderDev*>(p);
    }
  }
  if (rc == SDFR_OK) rc = build_tc_tables(d, spec, weights_host);
  if (rc != SDFR_OK) { sdfr_decoder_destroy(d); return rc; }
  *out = d;
  return SDFR_OK;
}

void sdfr_decoder_destroy(sdfr_decoder* dec) {
  if (!dec) return;
  free_tc_tables(dec);
  for (void* p : dec->allocs) cudaFree(p);
  if (dec->scratch) cudaFree(dec->scratch);
  delete dec;
}

int sdfr_decoder_tcgen05_ok(const sdfr_decoder* dec) { return dec ? dec->tc.ok : 0; }

int sdfr_decoder_check(sdfr_decoder* dec) {
  SDFR_REQUIRE(dec, SDFR_E_INVALID, "null decoder");
  int flag = 0;
  int rc = tc_overflow_flag(dec, &flag);
  if (rc) return rc;
  SDFR_REQUIRE(!flag, SDFR_E_UNSUPPORTED,
               "an activation left the fp16 range of the split-operand tensor-core kernel; use SDFR_MLP_FFMA for this network");
  return SDFR_OK;
}

static int pick_impl(const sdfr_decoder* dec, int impl) {
  if (impl == SDFR_MLP_AUTO) return dec->tc.ok ? SDFR_MLP_TCGEN05 : SDFR_MLP_FFMA;
  return impl;
}

int sdfr_decoder_eval(sdfr_decoder* dec, const float* inputs_dev, int64_t n, float* sdf_dev, float* dinput_dev,
                      int impl, void* stream) {
  SDFR_REQUIRE(dec && sdf_dev && (inputs_dev || n == 0) && n >= 0, SDFR_E_INVALID, "bad argument");
  MlpInputs in;
  in.inputs = inputs_dev; in.latent_unit = nullptr; in.lattice = make_lattice(2); in.points_per_batch = 1; in.n = n;
  in.index = nullptr; in.count_dev = nullptr; in.small_tiles = 0;
  impl = pick_impl(dec, impl);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (impl == SDFR_MLP_TCGEN05) return launch_mlp_tc(dec, in, sdf_dev, dinput_dev, s);
  if (impl == SDFR_MLP_TCGEN05_COARSE) {
    SDFR_REQUIRE(!dinput_dev, SDFR_E_INVALID, "the coarse decoder pass is forward only");
    return launch_mlp_tc_coarse(dec, in, sdf_dev, s);
  }
  SDFR_REQUIRE(impl == SDFR_MLP_FFMA, SDFR_E_INVALID, "unknown MLP implementation %d", impl);
  return launch_mlp_ffma(dec, in, sdf_dev, dinput_dev, s);
}

int sdfr_decoder_eval_lattice(sdfr_decoder* dec, const float* latent_unit_dev, int batch, int density,
                              float* sdf_dev, float* dinput_dev, int impl, void* stream) {
  SDFR_REQUIRE(dec && latent_unit_dev && sdf_dev && batch > 0 && density > 1, SDFR_E_INVALID, "bad argument");
  MlpInputs in;
  in.inputs = nullptr; in.latent_unit = latent_unit_dev; in.lattice = make_lattice(density);
  in.points_per_batch = (long long)density * density * density;
  in.n = in.points_per_batch * batch;
  in.index = nullptr; in.count_dev = nullptr; in.small_tiles = 0;
  impl = pick_impl(dec, impl);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (impl == SDFR_MLP_TCGEN05) return launch_mlp_tc(dec, in, sdf_dev, dinput_dev, s);
  if (impl == SDFR_MLP_TCGEN05_COARSE) {
    SDFR_REQUIRE(!dinput_dev, SDFR_E_INVALID, "the coarse decoder pass is forward only");
    return launch_mlp_tc_coarse(dec, in, sdf_dev, s);
  }
  SDFR_REQUIRE(impl == SDFR_MLP_FFMA, SDFR_E_INVALID, "unknown MLP implementation %d", impl);
  return launch_mlp_ffma(dec, in, sdf_dev, dinput_dev, s);
}

int sdfr_lattice_points(int density, float* points_dev, void* stream) {
  SDFR_REQUIRE(points_dev && density > 1, SDFR_E_INVALID, "bad argument");
  return launch_lattice_points(density, points_dev, reinterpret_cast<cudaStream_t>(stream));
}

int sdfr_surface_extract(const float* points_dev, int density, const float* sdf_dev, const float* grad_dev,
                         int grad_stride, int grad_col, int64_t n, float threshold, float* out_pts_dev,
                         float* out_nrm_dev, int32_t* out_idx_dev, int32_t* out_count_dev, int32_t* scratch_dev,
                         void* stream) {
  SDFR_REQUIRE(sdf_dev && grad_dev && out_pts_dev && out_nrm_dev && out_count_dev && scratch_dev && n >= 0,
               SDFR_E_INVALID, "null argument");
  SDFR_REQUIRE(points_dev || (density > 1 && (int64_t)density * density * density == n), SDFR_E_INVALID,
               "implicit lattice needs density^3 == n");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (n == 0) { SDFR_CUDA(cudaMemsetAsync(out_count_dev, 0, sizeof(int32_t), s)); return SDFR_OK; }
  SurfaceArgs a;
  a.points = points_dev; a.lattice = make_lattice(density > 1 ? density : 2); a.sdf = sdf_dev; a.grad = grad_dev;
  a.grad_stride = grad_stride; a.grad_col = grad_col; a.n = n; a.batch = 1; a.threshold = threshold;
  a.out_pts = out_pts_dev; a.out_nrm = out_nrm_dev; a.out_idx = out_idx_dev; a.out_glat = nullptr; a.glat_dim = 0;
  a.cap = n; a.out_count = out_count_dev; a.scratch = scratch_dev;
  return launch_surface_extract(a, s);
}

int64_t sdfr_splat_workspace_bytes(const sdfr_raster_cfg* cfg, int64_t m) {
  if (!cfg || m < 0) return -1;
  return (int64_t)splat_ws_layout(m, (int64_t)cfg->width * cfg->height).total;
}

int sdfr_splat_forward(const sdfr_raster_cfg* cfg, const float* coords_dev, const float* normals_dev,
                       const float* colors_dev, const float* pose_dev, int64_t m, float* color_dev, float* mask_dev,
                       float* depth_dev, float* nrm_map_dev, float* cam_pts_dev, float* cam_rgb_dev,
                       uint8_t* front_dev, float* xyzf_dev, float* rgbf_dev, int32_t* front_count_dev,
                       void* workspace_dev, void* stream) {
  SDFR_REQUIRE(cfg && pose_dev && workspace_dev && m >= 0, SDFR_E_INVALID, "null argument");
  SDFR_REQUIRE(cfg->width > 0 && cfg->height > 0, SDFR_E_INVALID, "bad resolution");
  SDFR_REQUIRE(m == 0 || (coords_dev && normals_dev), SDFR_E_INVALID, "null surfel arrays");
  SDFR_REQUIRE(cfg->output_nocs || colors_dev || m == 0, SDFR_E_INVALID, "colors required when output_nocs == 0");
  SDFR_REQUIRE(cfg->primitive >= SDFR_PRIM_DISC && cfg->primitive <= SDFR_PRIM_CIRCLE_OPT, SDFR_E_INVALID,
               "unknown primitive %d", cfg->primitive);
  SDFR_REQUIRE(!cfg->bg_dev || (!depth_dev && !nrm_map_dev), SDFR_E_UNSUPPORTED,
               "with a background only the color and mask maps exist (rasterer.py:107-144 cannot broadcast the others)");
  SDFR_REQUIRE(!cfg->bg_dev || cfg->output_nocs, SDFR_E_UNSUPPORTED,
               "a background is composited with colours shown as (c+1)/2 (rasterer.py:107-111): output_nocs must be set");
  SDFR_REQUIRE(cfg->primitive != SDFR_PRIM_CIRCLE_OPT ||
                   ((int)cfg->k[2] * 2 == cfg->width && (int)cfg->k[5] * 2 == cfg->height),
               SDFR_E_INVALID, "circle_opt takes the image size from the principal point (primitives.py:108-109): "
               "int(K[0,2])*2 x int(K[1,2])*2 must equal the resolution %dx%d", cfg->width, cfg->height);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t P = (int64_t)cfg->width * cfg->height;
  const SplatWs L = splat_ws_layout(m, P);
  char* ws = reinterpret_cast<char*>(workspace_dev);
  SplatView V = make_view(cfg, coords_dev, normals_dev, colors_dev, pose_dev, m, ws, L);
  V.color = color_dev; V.mask = mask_dev; V.depth = depth_dev; V.nmap = nrm_map_dev;
  V.cam_rgb = cam_rgb_dev;
  SDFR_CUDA(cudaMemcpyAsync(ws + L.off_view, &V, sizeof(V), cudaMemcpyHostToDevice, s));
  const SplatView* vd = reinterpret_cast<const SplatView*>(ws + L.off_view);
  int rc;
  if ((rc = launch_project(vd, 1, (int)m, s))) return rc;
  if (cfg->primitive == SDFR_PRIM_DISC) {
    if ((rc = launch_splat_forward(vd, 1, cfg->width, cfg->height, cfg->width * cfg->height <= kFineCropPixelsHost, s))) return rc;
    if (cfg->bg_dev && (rc = launch_disc_background(vd, 1, (int)P, s))) return rc;
  } else {
    if ((rc = launch_circle_forward(vd, 1, (int)P, s))) return rc;
  }
  if (m > 0 && cam_pts_dev) SDFR_CUDA(cudaMemcpyAsync(cam_pts_dev, V.cam_v, (size_t)m * 12, cudaMemcpyDeviceToDevice, s));
  if (m > 0 && front_dev) SDFR_CUDA(cudaMemcpyAsync(front_dev, V.front, (size_t)m, cudaMemcpyDeviceToDevice, s));
  if (front_count_dev) {
    if (m > 0) {
      SDFR_REQUIRE(!rgbf_dev || cam_rgb_dev, SDFR_E_INVALID, "rgbf needs cam_rgb");
      compact_front_kernel<<<1, 1024, 0, s>>>(vd, cam_rgb_dev, xyzf_dev, rgbf_dev, front_count_dev);
      SDFR_LAUNCH_CHECK();
    } else {
      SDFR_CUDA(cudaMemsetAsync(front_count_dev, 0, sizeof(int32_t), s));
    }
  }
  return SDFR_OK;
}

int sdfr_splat_backward(const sdfr_raster_cfg* cfg, const float* coords_dev, const float* normals_dev,
                        const float* colors_dev, const float* pose_dev, int64_t m, const float* g_color_dev,
                        const float* g_mask_dev, const float* g_depth_dev, const float* g_nrm_map_dev,
                        const float* g_cam_pts_dev, const float* g_cam_rgb_dev, float* d_coords_dev,
                        float* d_normals_dev, float* d_colors_dev, float* d_pose_dev, void* workspace_dev,
                        void* stream) {
  SDFR_REQUIRE(cfg && pose_dev && workspace_dev && m >= 0, SDFR_E_INVALID, "null argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t P = (int64_t)cfg->width * cfg->height;
  const SplatWs L = splat_ws_layout(m, P);
  char* ws = reinterpret_cast<char*>(workspace_dev);
  const int npose = cfg->rot == SDFR_ROT_DCM ? 12 : 7;
  if (d_pose_dev) SDFR_CUDA(cudaMemsetAsync(d_pose_dev, 0, npose * sizeof(float), s));
  if (m == 0) return SDFR_OK;
  // the forward call left the view (with the map pointers) in the workspace; refresh the input pointers
  SplatView V = make_view(cfg, coords_dev, normals_dev, colors_dev, pose_dev, m, ws, L);
  SDFR_CUDA(cudaMemcpyAsync(ws + L.off_view, &V, sizeof(V), cudaMemcpyHostToDevice, s));
  const SplatView* vd = reinterpret_cast<const SplatView*>(ws + L.off_view);
  int rc;
  if ((rc = launch_pixel_grad_prep(vd, 1, (int)P, g_color_dev, g_mask_dev, g_depth_dev, g_nrm_map_dev, s))) return rc;
  if (cfg->primitive == SDFR_PRIM_DISC) {
    if ((rc = launch_splat_backward(vd, 1, (int)m, s))) return rc;
  } else {
    if ((rc = launch_circle_backward(vd, 1, (int)m, cfg->bg_dev ? 1 : 0, s))) return rc;
  }
  pose_chain_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(vd, g_cam_pts_dev, g_cam_rgb_dev, d_coords_dev,
                                                                d_normals_dev, d_colors_dev, d_pose_dev);
  SDFR_LAUNCH_CHECK();
  return SDFR_OK;
}

int sdfr_loss3d(const float* xyzf_dev, int64_t q, const float* lidar_scaled_dev, int64_t nl, double radius,
                float* loss_dev, float* d_xyzf_dev, float* d_lidar_dev, void* stream) {
  SDFR_REQUIRE(loss_dev && q >= 0 && nl >= 0, SDFR_E_INVALID, "bad argument");
  return launch_loss3d_standalone(xyzf_dev, q, lidar_scaled_dev, nl, radius, loss_dev, d_xyzf_dev, d_lidar_dev,
                                  reinterpret_cast<cudaStream_t>(stream));
}

int sdfr_loss2d(const float* color_dev, const float* target_dev, int height, int width, float* loss_dev,
                float* d_color_dev, void* stream) {
  SDFR_REQUIRE(color_dev && target_dev && loss_dev && height >= 0 && width >= 0, SDFR_E_INVALID, "bad argument");
  return launch_loss2d_standalone(color_dev, target_dev, height, width, loss_dev, d_color_dev,
                                  reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
