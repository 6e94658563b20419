/*
 * sdpc_b200.h - C ABI of the B200-native simultaneous multi-view Langevin sampling step.
 *
 * The reference (Ryan-Faulkner/Simultaneous-Diffusion-for-Pointclouds) has no FFI: its
 * hot path sits behind two Python call sites.  This header is the boundary a
 * maintainer binds instead (ctypes stub in INTEGRATION.md).  Each entry point cites the
 * reference code it replaces.
 *
 * Conventions
 *   - every function returns an int status (0 = SDPC_OK, <0 = error); nothing throws, nothing
 *     synchronises the device implicitly; sdpc_last_error() returns a thread-local message.
 *   - unless a parameter says "host", pointers are CUDA device pointers owned by the caller.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - tensors are dense, row-major, in the reference's own layouts (NCHW float32 images).
 *   - a handle is not re-entrant: one forward at a time per handle.
 */
#ifndef SDPC_B200_H
#define SDPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDPC_ABI_VERSION 2

enum sdpc_status {
  SDPC_OK = 0,
  SDPC_ERR_ARG = -1,        /* bad argument (null pointer, unsupported shape, ...) */
  SDPC_ERR_CUDA = -2,       /* a CUDA runtime / driver call failed */
  SDPC_ERR_STATE = -3,      /* call order violated (e.g. forward before finalize) */
  SDPC_ERR_WORKSPACE = -4,  /* workspace pointer null or too small */
  SDPC_ERR_NAME = -5,       /* unknown parameter name / wrong shape for that name */
  SDPC_ERR_UNSUPPORTED = -6 /* device is not sm_100 or configuration not supported */
};

/* Arithmetic of the score network's 3x3 convolutions (the 75 conv2d of ncsnv2.py:420-518). */
enum sdpc_precision {
  SDPC_PREC_FP32 = 0, /* CUDA-core fp32 FMA, strict-parity arm (matches CPU fp32 to ~1e-5) */
  SDPC_PREC_TF32 = 1, /* tcgen05 kind::tf32, fp32 activations rounded to tf32, fp32 accumulate */
  SDPC_PREC_BF16 = 2, /* tcgen05 kind::f16 (bf16 operands), fp32 accumulate in TMEM */
  SDPC_PREC_BF16X3 = 3, /* fp32-parity arm on the tensor cores: operands split into bf16 hi + lo planes,
                          X_hi.W_hi + X_hi.W_lo + X_lo.W_hi accumulated in fp32 (3 tcgen05 passes, ~16-bit operands) */
  SDPC_PREC_FP16 = 4   /* tcgen05 kind::f16 with IEEE half operands (11-bit significand: the precision class of the TF32
                          convolutions the reference runs on a GPU, at the bf16 arm's rate), fp32 accumulate; finite values
                          saturate at +-65504 */
};

int sdpc_abi_version(void);
/* sizeof of the ABI structs as this build sees them: 0 sdpc_step_params, 1 sdpc_step_buffers, 2 sdpc_score_config,
 * 3 sdpc_projection_params; any other index gives 0.  A binding compares them with its own layout at load time. */
size_t sdpc_abi_struct_bytes(int which);
const char* sdpc_last_error(void);
/* Name of the kernels this build contains, e.g. "sm_100a". */
const char* sdpc_build_arch(void);

/* ------------------------------------------------------------------------------------------
 * Score network: NCSN_LiDAR_small.forward(x, y)        LiDARGen/models/ncsnv2.py:484-518
 * (constructor / parameter inventory                    LiDARGen/models/ncsnv2.py:420-477)
 * ---------------------------------------------------------------------------------------- */
typedef struct sdpc_score sdpc_score_t;

typedef struct sdpc_score_config {
  int32_t channels;    /* config.data.channels (2: log-range, intensity) */
  int32_t height;      /* config.data.image_size  (64) */
  int32_t width;       /* config.data.image_width (1024) */
  int32_t ngf;         /* config.model.ngf (128) */
  int32_t num_classes; /* config.model.num_classes = length of the `sigmas` buffer */
  int32_t precision;   /* enum sdpc_precision */
  int32_t max_views;   /* largest batch a forward will see (sizes the activation arena) */
  int32_t reserved;    /* flags, 0 for production use: bit 0 keeps every intermediate alive for sdpc_score_read_tap (tests),
                          bit 1 launches the kernels eagerly instead of replaying the captured CUDA graph */
} sdpc_score_config;

int sdpc_score_create(const sdpc_score_config* cfg, sdpc_score_t** out);
int sdpc_score_destroy(sdpc_score_t* h);

/* Number of state_dict entries the handle expects (153 parameters + the `sigmas` buffer). */
int sdpc_score_param_count(const sdpc_score_t* h);
/* i-th expected entry: its reference state_dict key and shape (ndim <= 4). */
int sdpc_score_param_info(const sdpc_score_t* h, int i, const char** name, int64_t shape[4], int* ndim);

/* Load one float32 tensor by its reference state_dict key (checkpoint layout of
 * runners/ncsn_runner_kitti_simultaneous.py:472-489 after stripping "module.").
 * `data` may be a host or a device pointer (on_device = 0/1). */
int sdpc_score_load_param(sdpc_score_t* h, const char* name, const float* data,
                          const int64_t* shape, int ndim, int on_device, void* stream);
/* Repack the loaded weights into the kernels' layouts; fails if an entry is missing. */
int sdpc_score_finalize(sdpc_score_t* h, void* stream);

size_t sdpc_score_workspace_bytes(const sdpc_score_t* h, int n_views);

/* out[b] = s(x[b], sigma[labels[b]]);  x,out: float32 [n_views, channels, H, W] (NCHW);
 * labels: int64 [n_views].  Replaces `scorenet(x_mod, labels)` at KITTISampling.py:137,505 and
 * models/__init__.py:240,594,1401,1431. */
int sdpc_score_forward(sdpc_score_t* h, const float* x, const int64_t* labels, float* out,
                       int n_views, void* workspace, size_t workspace_bytes, void* stream);

/* Test hook: after a forward, copy a named intermediate (e.g. "res1.0", "refine4") as float32
 * NCHW into `out` (device, capacity in elements). Writes its [C,H,W] to chw. */
int sdpc_score_read_tap(sdpc_score_t* h, const char* tap, float* out, size_t capacity,
                        int n_views, int chw[3], void* stream);
/* Kernels launched by the most recent forward (for bench.py's gpu_launches). */
int sdpc_score_last_launch_count(const sdpc_score_t* h);
/* CUDA-event time (ms) of the convolution kernels alone is measured by bench.py itself; this
 * returns the algorithmic FLOPs (2*M*N*K over all convolutions) of one view-forward. */
double sdpc_score_flops_per_view(const sdpc_score_t* h);
/* Roofline support for bench.py: when on, every tensor-core convolution launch of a forward is
 * bracketed by CUDA events on the launching stream; collect() waits for them and returns the summed
 * device time, the summed algorithmic FLOPs (2*M*N*K) and the number of launches, then resets. */
int sdpc_score_set_profiling(sdpc_score_t* h, int on);
int sdpc_score_profile_collect(sdpc_score_t* h, double* total_ms, double* total_flops, int* n_launches);

/* ------------------------------------------------------------------------------------------
 * Langevin update + cross-view consistency step
 *   a-4  KITTISampling.py:137-490   (pose matrices)         variant = SDPC_VARIANT_POSE
 *   a-5  models/__init__.py:240-582 (translations)          variant = SDPC_VARIANT_TRANSLATION
 *   a-6  models/__init__.py:1401-1416 (update only)         share = 0
 * ---------------------------------------------------------------------------------------- */
enum sdpc_variant { SDPC_VARIANT_POSE = 0, SDPC_VARIANT_TRANSLATION = 1 };

typedef struct sdpc_step_params {
  int32_t n_views;       /* B: views in x (all groups) */
  int32_t group_size;    /* A = actualBatchSize; B % A == 0 */
  int32_t height;        /* H */
  int32_t width;         /* W */
  int32_t big_rows;      /* R = bigRowCount (KITTISampling.py:68) */
  int32_t variant;       /* enum sdpc_variant */
  int32_t share;         /* c >= minStepToShare (KITTISampling.py:160) */
  int32_t nan_to_num;    /* zero NaN / clamp inf of the score (KITTISampling.py:138); a-6: 0 */
  int32_t sky_filter;    /* a-5: drop candidates whose SOURCE pixel has sky == 0 (__init__.py:352) */
  int32_t tgt_first;     /* first target view this call resolves (multi-GPU sharding), else 0 */
  int32_t tgt_count;     /* number of target views, else n_views */
  int32_t scalar_div_recip; /* 1: tensor/python-scalar divisions as x*(1/s) like torch's CUDA kernels (matches the
                               reference on a GPU bit-for-bit); 0: IEEE division like torch's CPU kernels */
  int32_t key_shift_override; /* test hook: bits of the squared range dropped from the packed (range | source id) key, which
                                 makes packed winners wrong so that verification and the fix pass are exercised; 0 = auto */
  int32_t winner_mode;   /* how the nearest candidate of a z-buffer cell is identified.  The scatter keeps min(r^2) (the exact
                            squared range: the log-range is a monotone function of it) and min(packed r^2 | source id) with two
                            fire-and-forget 64-bit reductions; 0 (production): the packed winner is verified (its exact squared
                            range recomputed) only where its identity matters -
                            cells the controlled average declares "far", or every filled cell when cell-level debug output
                            is requested; 1: every filled cell is verified; 2: every filled cell goes through the exact
                            second traversal (smallest source id at exactly the nearest depth).  A cell whose packed winner
                            is not confirmed always goes through that traversal, so the three modes agree bit for bit. */
  float step_size;       /* eps: float32 value of step_lr*(sigma/sigmas[-1])**2 (KITTISampling.py:135) */
  float noise_scale;     /* float32 value of np.sqrt(step_size*2) (KITTISampling.py:156) */
  float grad_ref;        /* step_refer */
  float corr_coef;       /* correlation_coefficient */
  float sigma_mod;       /* sigmaMod (KITTISampling.py:117-119) */
  float min_depth_thr;   /* float32 log-range threshold (KITTISampling.py:273-275); < 0 disables */
  double allowance;      /* controlled-average allowance in metres (KITTISampling.py:381); < 0: plain average */
  double h_min, dh;      /* horizontalMin, horizontalAngles (KITTISampling.py:64,66) */
  double big_row_min, dv;/* bigRowMin, verticalAngles     (KITTISampling.py:65,71) */
} sdpc_step_params;

typedef struct sdpc_step_buffers {
  float* x;                 /* in/out [B,2,H,W]: x_mod */
  const float* grad;        /* [B,2,H,W] score (may be NULL when step_size == 0) */
  const float* noise;       /* [B,2,H,W] z ~ N(0,1), drawn by the caller (torch.randn_like) */
  const float* refer;       /* [B,2,H,W] refer_image */
  const int32_t* mask;      /* [B,2,H,W] refer_mask (int32 0/1) */
  const uint8_t* sky;       /* [B,H,W]   sky (bool) */
  const uint8_t* exist;     /* [A,H,W]   existMask[:A] (bool) */
  const double* to_world;   /* [B,16] row-major 4x4 (pose variant) */
  const double* from_world; /* [B,16] */
  const float* origins;     /* [A,3] originList[:A,:,0,0] (translation variant) */
  const double* cos_az;     /* [W] cos(azimuth)   (KITTISampling.py:101,176) */
  const double* sin_az;     /* [W] */
  const double* cos_el;     /* [H] cos(elevation) (KITTISampling.py:102,176) */
  const double* sin_el;     /* [H] */
  float* grad_likelihood;   /* optional out [B,2,H,W]: -mask*(x_before-refer) (KITTISampling.py:144) */
  float* new_images;        /* optional out [B,2,H,W]: newImages (KITTISampling.py:415) */
  int32_t* too_high;        /* optional out [1]: the tooHigh gate (KITTISampling.py:162) */
  /* optional debug outputs for parity tests */
  int32_t* dbg_row;         /* [B, A*H*W] row in the R-row grid, per (target, source point) */
  int32_t* dbg_col;         /* [B, A*H*W] */
  uint8_t* dbg_valid;       /* [B, A*H*W] */
  int32_t* dbg_cnt;         /* [B,R,W] candidates per pixel */
  int32_t* dbg_winner;      /* [B,R,W] source point id (a*H*W + r*W + c) of the nearest candidate, -1 if empty */
  double* dbg_min_d;        /* [B,R,W] its log-range */
} sdpc_step_buffers;

size_t sdpc_step_workspace_bytes(int n_views, int height, int width, int big_rows);
/* Arms the z-buffers (empty cells) and clears the header.  Call ONCE after allocating a step workspace, before the first
 * sdpc_crossview_share / sdpc_langevin_reproject_step on it: every share call re-arms the cells it used, so there is no
 * per-step memset.  A share call on a workspace that was never armed traps (the CUDA error surfaces at the next
 * synchronisation) instead of returning wrong images. */
int sdpc_step_workspace_init(void* workspace, size_t workspace_bytes, int n_views, int height, int width, int big_rows,
                             void* stream);

/* x <- x + eps*nan_to_num(grad) + rho*(-mask*(x-refer)) + noise_scale*noise  (KITTISampling.py:137-156);
 * also leaves max|x[:,0]| of the updated sample in the workspace for the tooHigh gate. */
int sdpc_langevin_update(const sdpc_step_params* p, const sdpc_step_buffers* b,
                         void* workspace, size_t workspace_bytes, void* stream);
/* Optional hook between update and share for sharded runs: fold other ranks' max|x0| in. */
int sdpc_step_merge_max(void* workspace, const float* other_max, int n, void* stream);
int sdpc_step_read_max(void* workspace, float* out_max, void* stream);
/* View sharding over ranks (the reference's group concatenation KITTISampling.py:189-206 happens on one
 * device): ONE exchange per step.  A rank's slot of the gather buffer holds its `views_per_rank` updated views
 * [views_per_rank,2,H,W] followed by its max |x0| word (sdpc_shard_slot_floats floats per slot, slots contiguous in rank
 * order).  pack fills the rank's own slot after sdpc_langevin_update; the caller all-gathers the slots (NCCL
 * all_gather_into_tensor in place); unpack copies the other ranks' views into x (rank-major view order) and folds their
 * maxima into the workspace, after which sdpc_crossview_share runs on the rank's target range. */
size_t sdpc_shard_slot_floats(int views_per_rank, int height, int width);
int sdpc_shard_pack(void* workspace, const float* x_own, float* slot, int views_per_rank, int height, int width, void* stream);
int sdpc_shard_unpack(void* workspace, float* x, const float* gathered, int world, int rank, int views_per_rank,
                      int height, int width, void* stream);
/* Cross-view block on the updated sample (KITTISampling.py:160-490): un-project, pose transform,
 * re-project, z-buffer (count, sums, nearest), fusion, crop/mirror, correction. */
int sdpc_crossview_share(const sdpc_step_params* p, const sdpc_step_buffers* b,
                         void* workspace, size_t workspace_bytes, void* stream);
/* Kernels one sdpc_langevin_reproject_step call with these parameters launches (for bench.py's gpu_launches):
 * update [+ scatter, resolve, fix (exits at once unless resolve flagged a cell), correct]. */
int sdpc_step_kernel_launches(const sdpc_step_params* p, const sdpc_step_buffers* b);
/* update followed by share (when p->share != 0) on one stream. */
int sdpc_langevin_reproject_step(const sdpc_step_params* p, const sdpc_step_buffers* b,
                                 void* workspace, size_t workspace_bytes, void* stream);

/* The whole sampling step with HOST sample buffers - the call a host-side caller without device tensors makes, i.e. the
 * body of the reference's inner loop (KITTISampling.py:137-490: scorenet(x_mod, labels), the update, the cross-view
 * block) in one entry point, everything enqueued on `stream`:
 *   x_host -> b->x (H2D);
 *   b->grad = score(b->x, labels) when `score` is non-NULL (labels: device int64 [B]; score_workspace as for
 *     sdpc_score_forward), else grad_host -> b->grad when grad_host is non-NULL (else b->grad as it is);
 *   noise_host -> b->noise when non-NULL (else b->noise as it is: drawn on the device by the caller);
 *   update (+ share when p->share);
 *   b->x -> x_host and, when non-NULL and p->share, b->new_images -> new_images_host (D2H).
 * Host buffers may be pinned or pageable; the static inputs (refer, mask, sky, exist, poses, LUTs) stay device pointers in
 * `b`.  bench.py's e2e leg and tests/test_gpu_host_step.py call it. */
int sdpc_langevin_reproject_step_host(const sdpc_step_params* p, const sdpc_step_buffers* b,
                                      sdpc_score_t* score, const int64_t* labels, void* score_workspace,
                                      size_t score_workspace_bytes, float* x_host, const float* grad_host,
                                      const float* noise_host, float* new_images_host, void* workspace,
                                      size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Row N1: point cloud -> range image, the projection that renders the sampler's inputs
 *   point_cloud_to_range_image        LiDARGen/datasets/lidar_utils.py:54-347
 * ---------------------------------------------------------------------------------------- */
typedef struct sdpc_projection_params {
  int32_t n_points;       /* N */
  int32_t point_stride;   /* doubles per point (>= 3: x, y, z first) */
  int32_t intensity_col;  /* column of the remission value, or -1 (return_remission = False) */
  int32_t height, width;  /* rowMax, colMax (width <= 1024) */
  int32_t reserved;
  double origin[3];       /* sensor origin subtracted from every point (lidar_utils.py:146-147) */
  double h_min, dh;       /* horizontalMin, horizontalAngles (lidar_utils.py:103,109) */
  double v_min, dv;       /* verticalMin, verticalAngles     (lidar_utils.py:104,118) */
} sdpc_projection_params;

size_t sdpc_projection_workspace_bytes(int height, int width);
/* points: float64 [N, point_stride] (device).  Outputs (device), all already flipped by 180 degrees like the reference:
 * depth [H,W] f64 (2057.701 where empty), intensity [H,W] f64 (nullable when intensity_col < 0), obfuscation [H,W] u8,
 * sky [H,W] u8 (the reference returns it cleared), index [H,W] f64 (point index of the nearest point, -1 where empty). */
int sdpc_pointcloud_to_range_image(const sdpc_projection_params* p, const double* points, double* depth,
                                   double* intensity, uint8_t* obfuscation, uint8_t* sky, double* index,
                                   void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Row N3: the output stage right after the sampler
 *   range image -> xyz point cloud     LiDARGen/visualization.py:12-43 (visualize_tensor, numeric part)
 *   L1 depth / intensity error sums    MeasureResults/QuantifyingNotebookSynthesis_Line.ipynb, cell 1
 * ---------------------------------------------------------------------------------------- */
size_t sdpc_points_workspace_bytes(int n_views, int height, int width);
/* image: float32 [V,2,H,W] (device; channel 0 = log-range r, channel 1 = intensity).  depth = 2^(6 r) - 1 in float32;
 * a pixel yields a point iff 0.5 < depth < 63 (visualization.py:41).  cos_yaw / sin_yaw [W] and cos_pitch / sin_pitch [H]
 * are float64 device tables of the reference's yaw / pitch grid (visualization.py:30-33).  Outputs (device): xyz float64
 * [V, H*W, 3], the first n_points[v] rows of view v valid and in row-major pixel order (= pts[mask, :]); intensity
 * float32 [V, H*W] and pixel int32 [V, H*W] (source pixel of every point) are nullable. */
int sdpc_range_image_to_points(const float* image, int n_views, int height, int width, const double* cos_yaw,
                               const double* sin_yaw, const double* cos_pitch, const double* sin_pitch, double* xyz,
                               float* intensity, int* pixel, int* n_points, void* workspace, size_t workspace_bytes,
                               void* stream);
/* pred, gt, input: float32 [V,2,H,W] (device).  out: float64 [V,8] (device) =
 *   { sum|d_pred - d_gt| over all pixels, sum|i_pred - i_gt| over all pixels, the same two over the input pixels,
 *     sum d_pred over the input pixels, #pixels, #input pixels, 0 }   with d = 2^(6 r) - 1 and
 *   input pixel = (input_r > 0.001) & (d_gt < 63)   (notebook: inputMask, distanceError, intensityError, ...Input). */
int sdpc_depth_intensity_errors(const float* pred, const float* gt, const float* input, int n_views, int height, int width,
                                double* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Row N2: multi-view dataset assembly around the projection of row N1
 *   KITTI360_im_8batch.__getitem__      LiDARGen/datasets/kitti360_im_8Batch.py:94-304
 *   (kitti360_im_AllForOne.py:94-355 and kitti360_im_simultenous_densification.py differ in the pose they pick and in
 *    the origin handed to the projection, not in these two steps)
 * ---------------------------------------------------------------------------------------- */
/* scan: float32 [N,4] (device; x, y, z, remission as read from a KITTI .bin).  to_world / from_world: float64 4x4 row
 * major (HOST pointers).  out: float64 [N,4] (device) = (from_world . (to_world . (x, y, z, 1)))[:3], remission. */
int sdpc_transform_scan(const float* scan, int n_points, const double* to_world, const double* from_world, double* out,
                        void* stream);
/* depth, intensity (nullable), obfuscation (nullable), sky: the images sdpc_pointcloud_to_range_image returns (device).
 * real: float64 [C,H,W] (C = 2 with intensity, else 1): log2(depth + 1) / 6 and remission, offset by 1e-4 and clipped
 * to [0,1], holes (depth >= max_range, remission >= 1) zeroed first.  known (nullable): uint8 [C,H,W] =
 * logical_not(mask) with mask = obfuscation | holes.  notsky (nullable): uint8 [H,W] = logical_not(sky shifted down by
 * three rows). */
int sdpc_range_image_postprocess(const double* depth, const double* intensity, const uint8_t* obfuscation,
                                 const uint8_t* sky, int height, int width, double max_range, double* real,
                                 uint8_t* known, uint8_t* notsky, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SDPC_B200_H */
