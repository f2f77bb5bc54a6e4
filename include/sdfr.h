/*
 * sdfr.h - C ABI of libsdfr.so, the B200 (sm_100a) implementation of the
 * sdflabel differentiable SDF render / refine hot path.
 *
 * The reference (TRI-ML/sdflabel) has no FFI: its boundary is the Python object
 * surface used by pipelines/optimizer.py and pipelines/refine_css.py
 * (SURVEY.md section 8(b)).  Each entry point below replaces one of those
 * Python-level operations; the citation after "replaces:" is the reference
 * file:line.  The Python mirror classes in sdflabel_b200/ bind these symbols
 * with ctypes (see INTEGRATION.md for the stub a maintainer would add).
 *
 * Conventions
 *  - every pointer named *_dev is a device pointer borrowed from the caller
 *    (e.g. a torch tensor's data_ptr()); *_host pointers are host memory;
 *  - the caller owns all memory except the opaque handles;
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *  - every function returns 0 on success, a negative SDFR_E_* code otherwise;
 *    sdfr_last_error() returns a thread-local message for the last failure;
 *  - handles are thread-compatible, not thread-safe;
 *  - nothing here ever falls back to the CPU: without a CUDA device the calls
 *    fail with SDFR_E_CUDA.
 */
#ifndef SDFR_H_
#define SDFR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDFR_VERSION 100 /* 0.1.0 */

enum {
  SDFR_OK = 0,
  SDFR_E_INVALID = -1,     /* bad argument */
  SDFR_E_CUDA = -2,        /* CUDA runtime error (message has the detail) */
  SDFR_E_UNSUPPORTED = -3, /* network spec outside what the kernels cover */
  SDFR_E_CAPACITY = -4     /* a fixed-capacity buffer would overflow */
};

/* MLP kernel selection */
enum {
  SDFR_MLP_AUTO = 0,    /* tcgen05 kernel when the spec qualifies, else FFMA */
  SDFR_MLP_FFMA = 1,    /* fp32 CUDA-core kernel (any supported spec) */
  SDFR_MLP_TCGEN05 = 2, /* tensor-core kernel: fp16 hi/lo split operands, fp32 accumulate in TMEM */
  SDFR_MLP_TCGEN05_COARSE = 3 /* tensor-core kernel, hi halves only (fp16 operand precision, sdf to ~3e-4),
                                 forward only: the band pre-selection pass of the fused engine */
};

enum { SDFR_ROT_DCM = 0, SDFR_ROT_QUAT = 1 };

/* Rasterer primitives (rasterer.py:93-105) */
enum {
  SDFR_PRIM_DISC = 0,       /* inside_surfel, diam 0.04: tangent discs (the refine loop's primitive) */
  SDFR_PRIM_CIRCLE = 1,     /* inside_circle, diam 0.02: screen-space circles, sigmoid soft clamp */
  SDFR_PRIM_CIRCLE_OPT = 2  /* inside_circle_opt, diam 0.025: 15 x 15 pixel stamps */
};

int sdfr_version(void);
const char* sdfr_last_error(void);
/* bit 0: a CUDA device is visible; bit 1: that device is sm_100 (tcgen05 path usable) */
int sdfr_caps(void);

/* ------------------------------------------------------------------------- *
 * DeepSDF decoder.
 * replaces: sdfrenderer/deepsdf/networks/deep_sdf_decoder_scale.py:10-114 (Decoder)
 *           sdfrenderer/deepsdf/workspace.py:167-188 (setup_dsdf: the Python
 *           side parses the .json/.pt and hands the folded weights over here)
 * ------------------------------------------------------------------------- */
typedef struct sdfr_decoder sdfr_decoder;

typedef struct {
  int32_t latent_size;      /* L */
  int32_t num_layers;       /* number of Linear layers (stock: 9) */
  const int32_t* in_dims;   /* [num_layers] fan-in of each Linear (after any concat) */
  const int32_t* out_dims;  /* [num_layers] fan-out */
  const int32_t* concat;    /* [num_layers] 0 none, 1 = cat[x, input] before this layer
                               (latent_in, decoder.py:90-91), 2 = cat[x, xyz] (xyz_in_all, 92-93) */
  const int32_t* layer_norm;/* [num_layers] 1 = LayerNorm(out) after the Linear (decoder.py:99-101) */
  int32_t use_tanh;         /* extra tanh on the last Linear (decoder.py:96-97); the final
                               tanh (decoder.py:106-107) is always applied */
} sdfr_decoder_spec;

/* weights_host[l]: row-major [out][in] effective weight (weight_norm already folded:
 * W = g * v / ||v||_row, torch weight_norm dim=0); bias_host[l]: [out];
 * ln_weight_host[l] / ln_bias_host[l]: [out] or NULL. */
int sdfr_decoder_create(const sdfr_decoder_spec* spec, const float* const* weights_host,
                        const float* const* bias_host, const float* const* ln_weight_host,
                        const float* const* ln_bias_host, sdfr_decoder** out);
void sdfr_decoder_destroy(sdfr_decoder* dec);
/* 1 when the tcgen05 kernel covers this spec (all widths <= 512, no LayerNorm) */
int sdfr_decoder_tcgen05_ok(const sdfr_decoder* dec);
/* Synchronises the device and fails with SDFR_E_UNSUPPORTED if, since the last check, a scaled
 * activation of the split-fp16 tensor-core kernel left the fp16 range (its results are then invalid
 * and the caller should select SDFR_MLP_FFMA); SDFR_OK otherwise. */
int sdfr_decoder_check(sdfr_decoder* dec);

/* sdf[n] = Decoder(inputs[n, L+3]); if dinput_dev != NULL also the exact
 * gradient d sdf[n] / d inputs[n, :] (what the reference obtains with
 * pred_sdf_grid.sum().backward(), grid.py:55-56).
 * replaces: Decoder.forward (deep_sdf_decoder_scale.py:78-114). */
int sdfr_decoder_eval(sdfr_decoder* dec, const float* inputs_dev, int64_t n, float* sdf_dev,
                      float* dinput_dev, int impl, void* stream);

/* Same over the implicit Grid3D lattice: point k of detection b is lattice
 * point k (grid.py:22-41) with latent latent_unit_dev[b, :]; no HBM traffic for
 * the inputs.  sdf_dev: [batch, D^3]; dinput_dev: [batch, D^3, L+3] or NULL.
 * replaces: optimizer.py:99-104 (inputs = cat[latent.expand, grid.points]; dsdf(inputs)). */
int sdfr_decoder_eval_lattice(sdfr_decoder* dec, const float* latent_unit_dev, int batch, int density,
                              float* sdf_dev, float* dinput_dev, int impl, void* stream);

/* ------------------------------------------------------------------------- *
 * Grid3D.
 * replaces: sdfrenderer/grid.py:18-41 (lattice) and 43-71 (get_surface_points)
 * ------------------------------------------------------------------------- */
/* points_dev: [D^3, 3], bit-identical to Grid3D.generate_point_grid */
int sdfr_lattice_points(int density, float* points_dev, void* stream);

/* Zero-isosurface projection + band select, order preserving (ascending index).
 * points_dev [n,3] (NULL = implicit lattice of `density`), sdf_dev [n], grad_dev
 * [n, grad_stride] with the xyz gradient at columns grad_col..grad_col+2.
 * Outputs (capacity n rows each): out_pts [m,3] = p - sdf*n_hat, out_nrm [m,3],
 * out_idx [m] source index; *out_count_dev = m.  scratch_dev: >= (n/1024+2) int32.
 * replaces: Grid3D.get_surface_points (grid.py:43-71). */
int sdfr_surface_extract(const float* points_dev, int density, const float* sdf_dev, const float* grad_dev,
                         int grad_stride, int grad_col, int64_t n, float threshold, float* out_pts_dev,
                         float* out_nrm_dev, int32_t* out_idx_dev, int32_t* out_count_dev,
                         int32_t* scratch_dev, void* stream);

/* ------------------------------------------------------------------------- *
 * Rasterer.
 * replaces: sdfrenderer/renderer/rasterer.py:49-155 (forward: 'disc', 'circle', 'circle_opt', bg),
 *           renderer/projection.py:7-101 (dcm) and 104-199 (quat),
 *           renderer/primitives.py:165-243 (inside_surfel), 4-68 (inside_circle), 71-162 (inside_circle_opt),
 *           renderer/utils_rasterer.py:6-24 (qrot)
 * ------------------------------------------------------------------------- */
typedef struct {
  int32_t width, height;   /* resolution_px = (W, H) */
  float kinv[9];           /* row-major K^-1 (inverted in fp32, primitives.py:204) */
  float k[9];              /* row-major K (surfel bounding boxes, points_2d) */
  int32_t rot;             /* SDFR_ROT_DCM: pose = 4x4 row-major (first 3 rows used)
                              SDFR_ROT_QUAT: pose = [qw,qx,qy,qz,tx,ty,tz] */
  int32_t output_nocs;     /* 1: colours = object coords (x negated in the dcm path) shown as (c+1)/2;
                              0: colours = the given colour tensor, unscaled (rasterer.py:113-116) */
  int32_t primitive;       /* SDFR_PRIM_* (rasterer.py:93-105) */
  const float* bg_dev;     /* background image [3,H,W] or NULL (rasterer.py:107-111): composited into color
                              and mask; with a background the depth / normals maps are not available (the
                              reference fails to broadcast them) and colours are shown as (c+1)/2 */
} sdfr_raster_cfg;

/* Workspace sizes (bytes) for m surfels at the configured resolution. */
int64_t sdfr_splat_workspace_bytes(const sdfr_raster_cfg* cfg, int64_t m);

/* Forward.  coords/normals/colors: [m,3]; pose: 16 or 7 floats (device).
 * Outputs: color [3,H,W], mask [1,H,W], depth [1,H,W], normals [3,H,W] (any may
 * be NULL), cam_pts [m,3] (points['xyz']), cam_rgb [m,3] (points['rgb']),
 * front flags [m] (uint8, n.v < 0; all ones for quat), and the compacted
 * front-facing lists xyzf/rgbf [<=m,3] with *front_count_dev (dcm only; NULL ok).
 * workspace_dev keeps per-surfel / per-pixel state for the backward call. */
int sdfr_splat_forward(const sdfr_raster_cfg* cfg, const float* coords_dev, const float* normals_dev,
                       const float* colors_dev, const float* pose_dev, int64_t m, float* color_dev,
                       float* mask_dev, float* depth_dev, float* nrm_map_dev, float* cam_pts_dev,
                       float* cam_rgb_dev, uint8_t* front_dev, float* xyzf_dev, float* rgbf_dev,
                       int32_t* front_count_dev, void* workspace_dev, void* stream);

/* Backward.  Upstream gradients of the four maps (NULL = zero) and of the
 * per-surfel outputs cam_pts / cam_rgb (NULL = zero; gradients of the xyzf/rgbf
 * lists must be scattered into these by the caller).  Produces d coords [m,3],
 * d normals [m,3], d colors [m,3] (only when output_nocs == 0, else NULL) and
 * d pose (12 floats = rows of [R|t] for dcm, 7 for quat). */
int sdfr_splat_backward(const sdfr_raster_cfg* cfg, const float* coords_dev, const float* normals_dev,
                        const float* colors_dev, const float* pose_dev, int64_t m, const float* g_color_dev,
                        const float* g_mask_dev, const float* g_depth_dev, const float* g_nrm_map_dev,
                        const float* g_cam_pts_dev, const float* g_cam_rgb_dev, float* d_coords_dev,
                        float* d_normals_dev, float* d_colors_dev, float* d_pose_dev, void* workspace_dev,
                        void* stream);

/* ------------------------------------------------------------------------- *
 * Trace mode: per-ray sphere tracing against the decoder (BASELINE.json north_star).  Not a
 * reference function (the reference renderer is the surfel splat above); same decoder, same
 * camera conventions, validated against oracle/trace_oracle.py.
 * ------------------------------------------------------------------------- */
int64_t sdfr_trace_workspace_bytes(const sdfr_raster_cfg* cfg, const sdfr_decoder* dec);

/* Marches every pixel ray (r = K^-1 [x,y,1], object frame through pose_host = 4x4 row-major
 * [R|t], R orthogonal) by tau += sdf for at most max_steps decoder evaluations; a ray hits when
 * |sdf| < eps.  Outputs (device, may be NULL): depth [1,H,W] = z of the hit in the camera frame,
 * normals [3,H,W] = (R grad sdf/|grad sdf| + 1)/2, nocs [3,H,W] = ((-x,y,z)+1)/2 of the hit in
 * the object frame, mask [1,H,W]; zero where no hit.  latent_unit_dev: [L] normalised latent. */
int sdfr_trace_forward(sdfr_decoder* dec, const sdfr_raster_cfg* cfg, const float* latent_unit_dev,
                       const float* pose_host, int max_steps, float eps, float* depth_dev, float* nmap_dev,
                       float* nocs_dev, float* mask_dev, int32_t* hit_count_dev, void* workspace_dev,
                       void* cache_dev, float latent_lipschitz, int impl, void* stream);

/* Optional persistent block for the distance cache of the fused march (cache_dev above; NULL = rebuilt every
 * call): sdfr_trace_cache_bytes() bytes of device memory, zero-filled by the caller once, private to one decoder.
 * The cache (the decoder on a regular 40^3 lattice, 190 us) depends on the latent only; it is reused while
 * latent_lipschitz * |latent - latent of the cache| <= 0.01 with that product added to the march's safety margin
 * (latent_lipschitz: the certified bound of sdfr_refine_cfg; 0 = never reuse), and renewed otherwise - decided
 * on the device, no synchronisation.  Rendering one shape from many poses pays for the cache once. */
int64_t sdfr_trace_cache_bytes(void);
/* Brings the block up to date for this latent on `stream` (reuse or renew, as sdfr_trace_forward would).  Renders
 * that follow with latent_lipschitz < 0 use the block as it is: strips or views of one latent rendered
 * concurrently on several streams share one cache (order them after this call with an event). */
int sdfr_trace_cache_update(sdfr_decoder* dec, const float* latent_unit_dev, void* cache_dev, float latent_lipschitz,
                            void* stream);

/* Measurement aid: with on != 0 every fused-march forward synchronises after each launch and counts the decoder rows
 * it issued (process-wide; resets the counts).  counts4 = [distance-cache rows, march rows (fp16-operand forward),
 * non-empty march launches, Newton rows (full-precision forward + input gradient)].  Never on in a timed call. */
void sdfr_trace_set_stats(int on);
void sdfr_trace_get_stats(int64_t* counts4);

/* Gradient of a loss on the depth / NOCS maps with respect to the pose (12 floats, rows of
 * [R|t]) and the unit latent, by implicit differentiation of sdf(l, o + tau d) = 0 at the hits
 * found by the preceding forward call on the same workspace. */
int sdfr_trace_backward(sdfr_decoder* dec, const sdfr_raster_cfg* cfg, const float* pose_host, float eps,
                        const float* g_depth_dev, const float* g_nocs_dev, float* d_pose_dev,
                        float* d_latent_unit_dev, void* workspace_dev, void* stream);

/* ------------------------------------------------------------------------- *
 * Losses (value + gradient in one call).
 * replaces: pipelines/optimizer.py:166-198 (compute_loss_3d; exact 1-NN on the
 *           device instead of the sklearn KD-tree + D2H) and 200-237 (compute_loss_2d)
 * ------------------------------------------------------------------------- */
/* xyzf [q,3] query points; lidar_scaled [nl,3] (= lidar / scale); radius =
 * 0.2/scale.  loss_dev[0] = mean ||L_nn - v|| over pairs with NN distance <
 * radius (0 if none), loss_dev[1] = pair count.  d_xyzf [q,3] and d_lidar
 * [nl,3] receive d loss (either may be NULL). */
int sdfr_loss3d(const float* xyzf_dev, int64_t q, const float* lidar_scaled_dev, int64_t nl, double radius,
                float* loss_dev, float* d_xyzf_dev, float* d_lidar_dev, void* stream);

/* color, target: [3,H,W]; loss_dev[0] = loss (NaN when no pixel passes the
 * threshold, exactly like the reference), loss_dev[1] = selected pixel count;
 * d_color [3,H,W] = d loss / d color. */
int sdfr_loss2d(const float* color_dev, const float* target_dev, int height, int width, float* loss_dev,
                float* d_color_dev, void* stream);

/* ------------------------------------------------------------------------- *
 * Fused refine engine: the whole Optimizer.optimize loop on the device, all
 * detections of a batch per launch, no host synchronisation inside the loop.
 * replaces: pipelines/optimizer.py:26-54 (parameter groups, Adam + SGD) and
 *           56-164 (optimize), utils/refinement.py:108-125 (rot_from_yaw)
 * ------------------------------------------------------------------------- */
typedef struct sdfr_refine sdfr_refine;

typedef struct {
  int32_t batch;        /* detection slots (capacity); sdfr_refine_set_active picks how many a run covers */
  int32_t density;      /* Grid3D density D (config_refine.ini:11) */
  int32_t max_width;    /* crop capacity in pixels */
  int32_t max_height;
  int32_t max_lidar;    /* LIDAR points capacity per detection */
  int32_t max_iters;    /* loss-history capacity */
  float weight_2d;      /* config_refine.ini:26 */
  float weight_3d;      /* config_refine.ini:27 */
  int32_t mlp_impl;     /* SDFR_MLP_* */
  float latent_lipschitz; /* certified upper bound of |d sdf / d latent| of the decoder (product of the spectral norms
                           * along the latent's path): > 0 lets an iteration evaluate only the lattice points that the
                           * bound cannot exclude from the band, given the sdf at an earlier latent of the same
                           * detection (same surfels, same results); 0 evaluates the whole lattice every iteration
                           * as the reference does (optimizer.py:99-101) */
} sdfr_refine_cfg;

int sdfr_refine_create(sdfr_decoder* dec, const sdfr_refine_cfg* cfg, sdfr_refine** out);
void sdfr_refine_destroy(sdfr_refine* r);

/* Host-side inputs of detection b (copied host->device asynchronously on
 * `stream`; the host buffers must stay valid until the stream has run the copy).
 * k_host: 3x3 row-major intrinsics of the crop; kinv_host: its fp32 inverse as
 * the caller computed it (K.float().inverse(), primitives.py:204) or NULL to
 * invert here; nocs_host: [3,th,tw] CSS NOCS
 * prediction (nearest-resized to the crop on the device, optimizer.py:135-137);
 * lidar_host: [n_lidar,3] un-scaled LIDAR crop (optimizer.py:84); the initial
 * parameters follow get_opt_params (optimizer.py:26-40) and may be NULL when
 * sdfr_refine_import supplies them from device memory.  width / height must fit
 * max_width / max_height (SDFR_E_CAPACITY otherwise).  The optimiser state of
 * the slot (Adam moments, step count) is reset: a new Optimizer (optimizer.py:46-52). */
int sdfr_refine_set_detection(sdfr_refine* r, int b, const float* k_host, const float* kinv_host, int width,
                              int height, const float* nocs_host, int th, int tw, const float* lidar_host,
                              int n_lidar, const float* yaw_host, const float* trans_host,
                              const float* scale_host, const float* latent_host, void* stream);

/* The engine is created for cfg.batch detection SLOTS; the next runs cover slots [0, count).  Frames with
 * different numbers of detections share one engine: nothing is re-allocated, and one CUDA graph per
 * (count, crop size class) is captured on first use and replayed afterwards. */
int sdfr_refine_set_active(sdfr_refine* r, int count);

/* Reads the initial parameters of slot b from caller-owned DEVICE buffers (yaw [1], trans [3], scale [1],
 * latent [L]; any may be NULL) with one stream-ordered launch, after sdfr_refine_set_detection: the
 * `params` tensors of optimizer.py:26-30 never travel through the host. */
int sdfr_refine_import(sdfr_refine* r, int b, const float* yaw_dev, const float* trans_dev, const float* scale_dev,
                       const float* latent_dev, void* stream);

/* Adam state of slot b (moments of [yaw, tx, ty, tz] and the step count).  The reference builds its solver
 * once per Optimizer (optimizer.py:46-52), so repeated optimize() calls on one Optimizer continue the same
 * Adam state: the Python mirror reads it back after a run (get_..., valid after sdfr_refine_get /
 * sdfr_refine_get_batch, no synchronisation of its own) and restores it after the next
 * sdfr_refine_set_detection (set_..., stream-ordered). */
int sdfr_refine_set_optimizer_state(sdfr_refine* r, int b, const float* adam_m_host, const float* adam_v_host,
                                    int adam_t, void* stream);
int sdfr_refine_get_optimizer_state(sdfr_refine* r, int b, float* adam_m_host, float* adam_v_host, int* adam_t);

/* Enqueue `iters` iterations for the active detections. */
int sdfr_refine_run(sdfr_refine* r, int iters, void* stream);

/* Synchronises `stream` (once) and reads detection b back: params_host =
 * [yaw, tx, ty, tz, scale, latent(L)]; history_host (may be NULL): per executed
 * iteration [loss_2d, loss_3d, total, skipped] (4 floats), *n_history rows.
 * Fails with SDFR_E_UNSUPPORTED when the tensor-core decoder flagged an activation outside the
 * fp16 range since the last check (same condition as sdfr_decoder_check, read in the same sync). */
int sdfr_refine_get(sdfr_refine* r, int b, float* params_host, float* history_host, int* n_history,
                    void* stream);

/* The same for all active detections in ONE synchronisation: params_host [active, 5+L],
 * history_host [active, max_iters, 4] (may be NULL), n_history [active] (may be NULL). */
int sdfr_refine_get_batch(sdfr_refine* r, float* params_host, float* history_host, int* n_history, void* stream);

/* Largest |coarse sdf - accurate sdf| the engine has measured on pre-selected rows since the last failure
 * (tensor-core decoder only; 0 otherwise).  sdfr_refine_get fails with SDFR_E_UNSUPPORTED when it exceeds half
 * the pre-selection margin of 5e-3: a band point could then have been missed by the fp16-operand lattice pass. */
int sdfr_refine_preselect_error(sdfr_refine* r, float* err_host, void* stream);

/* Dump-time extents of get_kitti_label (utils/refinement.py:527-541) for all active detections: one more
 * lattice evaluation with the CURRENT latent as it is (not normalised, refine_css.py:229), band extraction,
 * min / max of the isosurface points.  extents_host [active, 8] = min xyz, max xyz (un-scaled object frame),
 * band point count, pre-selected row count (the length of views 2 / 3 / 12 / 14).  Overwrites the intermediates
 * of the last iteration (the isosurface points themselves stay readable through sdfr_refine_view: this is
 * also how the model clouds of the pose initialisation are produced, refine_css.py:143-151); synchronises. */
int sdfr_refine_label_extents(sdfr_refine* r, float* extents_host, void* stream);
/* Overwrites the latent of slot b (host -> device, stream-ordered) without touching anything else. */
int sdfr_refine_set_latent(sdfr_refine* r, int b, const float* latent_host, void* stream);

/* Temporal pruning (sdfr_refine_cfg.latent_lipschitz > 0): lattice points the pruned lattice passes have
 * evaluated since the last reset, and the detection-iterations they served (without pruning every
 * detection-iteration evaluates density^3 points; both are 0 for an engine created without the bound).  Synchronises. */
int sdfr_refine_lattice_rows(sdfr_refine* r, int64_t* rows_host, int64_t* detection_iterations_host, int reset,
                             void* stream);

/* Measurement aid (bench.py's per-kernel table): runs `iters` iterations of the active detections WITHOUT the
 * CUDA graph, with an event after every stage, and returns the mean device time of each stage in milliseconds
 * (stage_ms_host [>= 9]; *n_stages = 9; names from sdfr_refine_stage_name) and the number of rows the band
 * pass evaluated in the last iteration.  The iterations are real ones (the parameters move).  Synchronises. */
int sdfr_refine_profile(sdfr_refine* r, int iters, float* stage_ms_host, int max_stages, int* n_stages,
                        int32_t* band_rows_host, void* stream);
const char* sdfr_refine_stage_name(int stage);

/* Writes the current parameters of detection b into caller-owned DEVICE buffers (yaw [1], trans [3],
 * scale [1], latent [L]; any may be NULL) with one stream-ordered launch: the in-place update of the
 * `params` tensors (optimizer.py:26-30, read back at refine_css.py:229-231) without a host round trip. */
int sdfr_refine_export(sdfr_refine* r, int b, float* yaw_dev, float* trans_dev, float* scale_dev, float* latent_dev,
                       void* stream);

/* Optimizer.optimize of ONE detection as a single call (optimizer.py:56-164 as its caller sees it,
 * refine_css.py:216-231): sdfr_refine_set_detection(b, host inputs) -> sdfr_refine_import(params from the
 * caller's DEVICE tensors) -> Adam state restored when *adam_t > 0 (the solver of optimizer.py:46-52 lives as
 * long as the Optimizer) -> `iters` iterations -> sdfr_refine_export (the params tensors updated in place) ->
 * sdfr_refine_get (one synchronisation; params_host [5+L], history_host [iters,4] / n_history may be NULL) ->
 * adam_m_host [4] / adam_v_host [4] / *adam_t updated (may be NULL: a fresh solver, state not returned).
 * The host buffers only have to stay valid for the duration of the call. */
int sdfr_refine_optimize(sdfr_refine* r, int b, const float* k_host, const float* kinv_host, int width, int height,
                         const float* nocs_host, int th, int tw, const float* lidar_host, int n_lidar,
                         float* yaw_dev, float* trans_dev, float* scale_dev, float* latent_dev, float* adam_m_host,
                         float* adam_v_host, int* adam_t, int iters, float* params_host, float* history_host,
                         int* n_history, void* stream);

/* Device views of the last iteration's intermediates of detection b (for
 * parity tests and label dumps): kind 0 sdf [D^3], 1 d sdf/d[latent,x] of the band points of the
 * whole batch (compact, [sum m, L+3]),
 * 2 surfel points [m,3], 3 surfel normals [m,3], 4 color [3,H,W], 5 mask,
 * 6 normals map [3,H,W], 7 grads [dyaw,dt3,dscale,dlatent_unit(L),dlatent(L)],
 * 8 pre-selected point count m (int32; with the tensor-core decoder a superset of the band), 9 depth,
 * 10 camera-space surfel centres [m,3], 11 front-facing flags [m] (uint8), 12 band flags [m] (uint8: 1 =
 * inside the reference's band |sdf| < 0.03; rows with 0 are ignored by every stage), 13 composited surfel
 * colours [m,3] (points['rgb']), 14 lattice index of each pre-selected point [m] (int32).  Returns the device
 * pointer and element count. */
int sdfr_refine_view(sdfr_refine* r, int b, int kind, void** ptr_dev, int64_t* count);
/* Asynchronous device-to-device copy of such a view into dst_dev (at most max_count elements). */
int sdfr_refine_copy_view(sdfr_refine* r, int b, int kind, void* dst_dev, int64_t max_count, void* stream);

/* ------------------------------------------------------------------------- *
 * Initial-pose RANSAC
 * replaces: the sklearn KD-tree queries of PoseEstimator.init_pose_3d
 *           (utils/pose.py:133-134 build, 146 / 172 / 194 query) and, per RANSAC
 *           hypothesis, pose.py:166-178 (transform of the scene cloud, 1-NN in the
 *           model cloud, inlier test)
 * ------------------------------------------------------------------------- */
/* Exact 1-NN of queries [q,3] among refs [m,3] (float32 coordinates, distance in
 * float64 like a KD-tree; the lowest index wins a tie): idx [q] int32, dist [q] double. */
int sdfr_nn_query(const float* queries_dev, int64_t q, const float* refs_dev, int64_t m, int32_t* idx_dev,
                  double* dist_dev, void* stream);
/* HOST function (no kernel): `draws` times numpy's legacy `np.random.choice(range(n), 4, replace=False)`
 * (pose.py:139) on the caller's MT19937 state - key [624] and position as `np.random.get_state()` returns
 * them, updated in place so that `np.random.set_state()` continues the stream exactly where numpy would
 * be.  samples [draws,4] int32. */
int sdfr_np_choice4(uint32_t* mt_key, int32_t* mt_pos, int64_t n, int32_t draws, int32_t* samples_out);
/* All hypotheses of a RANSAC round in one launch.  transforms [h,12]: row-major 3x4
 * float32 [R*scale | t] (pose.py:166-168).  For hypothesis i and scene point j:
 * p = T_i * scene_pts[j] (float32), k = 1-NN of p in model_pts, inlier iff
 * |p - model_pts[k]| < metric_thr and |scene_cls[j] - model_cls[k]| < nocs_thr
 * (pose.py:170-178).  counts [h] int32 (zeroed here), masks [h,ns] uint8. */
int sdfr_ransac_score(const float* scene_pts_dev, const float* scene_cls_dev, int64_t ns,
                      const float* model_pts_dev, const float* model_cls_dev, int64_t m,
                      const float* transforms_dev, int num_hypotheses, double metric_thr, float nocs_thr,
                      int32_t* counts_dev, uint8_t* masks_dev, void* stream);

/* ------------------------------------------------------------------------- *
 * Rotated BEV box overlap of the KITTI evaluator
 * replaces: rotate_iou_gpu_eval / rotate_iou_kernel_eval (pipelines/rotate_iou.py:257-325),
 *           the reference's only hand-written GPU kernel (numba.cuda)
 * ------------------------------------------------------------------------- */
/* boxes [n,5], query [k,5] rows [cx, cy, w, h, angle]; iou [n,k] row-major.
 * criterion -1: intersection / union, 0: / area of the QUERY box, 1: / area of the box,
 * 2: raw intersection area (the argument order of rotate_iou.py:286 is kept). */
int sdfr_rotate_iou(const float* boxes_dev, int64_t n, const float* query_dev, int64_t k, int criterion,
                    float* iou_dev, void* stream);

/* Kernel launches issued by this library since load (bench.py's gpu_launches). */
int64_t sdfr_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SDFR_H_ */
