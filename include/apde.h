/*
 * apde.h -- C ABI of libapde (B200-native APDe-MVS hot path: deformable-PatchMatch depth/normal estimation
 * + depth/normal fusion).  Plain C types only: no C++, CUDA or torch types cross this boundary.
 *
 * Every entry point replaces a piece of the reference's driver interface (file:line into the reference tree):
 *
 *   apde_create / apde_destroy          <- cudaSetDevice(gpu_index)                      main.cpp:264
 *   apde_scene_begin / _set_view / _set_pairs / _commit
 *                                       <- GenerateSampleList + ReadImage + ReadCamera   main.cpp:44-102, APD.cpp:85-160
 *                                          (images/cameras become scene-resident instead of MemoryCache, APD.cpp:3-16)
 *   apde_problem_setup                  <- APD::APD + InuputInitialization + CudaSpaceInitialization
 *                                          + SetDataPassHelperInCuda                     APD.cpp:458-814
 *   apde_problem_stage / apde_problem_run
 *                                       <- APD::RunPatchMatch (one kernel / all kernels) APD.cu:2663-2737
 *   apde_problem_get / apde_problem_set <- GetPlaneHypothesis / GetPixelStates / GetConfidence   APD.cpp:816-826
 *   apde_problem_finish                 <- ProcessProblem tail: depth range check + WriteBinMat  main.cpp:168-190
 *   apde_pass_run                       <- ProcessProblem                                main.cpp:148-208
 *   apde_run_schedule                   <- main()'s round x iteration x problem loops    main.cpp:303-367
 *   apde_view_download / apde_view_upload
 *                                       <- ReadBinMat / WriteBinMat of depths/normals/weak/confidence.bin  APD.cpp:18-83
 *   apde_eval_costs                     <- ComputeBilateralNCCOld/New, ComputeGeomConsistencyCost (parity hook)
 *                                                                                        APD.cu:448-721, 865-902
 *   apde_weak_vis_filter / apde_fuse    <- WeakVisFilter, RunFusion                      APD.cpp:962-1227
 *   apde_fuse_variant                   <- RunFusion_TAT_I / RunFusion_TAT_A             APD.cpp:1229-1608
 *   apde_comm_* / apde_exchange / apde_fuse_collective
 *                                       <- (none: the reference is single-GPU per scan, run.py:127-153)
 *   apde_get_counters                   <- (none; the reference only prints wall-clock, main.cpp:157-161)
 *   apde_last_error                     <- CudaSafeCall / CudaCheckError                 APD.cpp:417-450
 *
 * Conventions: every function returns 0 on success and a negative apde_status otherwise (the reference prints and
 * exit()s; a library must not).  Host buffers are caller-owned; device state is owned by the context.  One context
 * per GPU; calls on one context must be serialised by the caller.  There is NO CPU fallback: without a usable CUDA
 * device apde_create fails with APDE_ERR_CUDA.
 */
#ifndef APDE_H_
#define APDE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APDE_MAX_IMAGES 32 /* main.h:40 */
#define APDE_ANCHOR_NUM 9  /* main.h:41 */

typedef enum {
    APDE_OK = 0,
    APDE_ERR_ARG = -1,
    APDE_ERR_CUDA = -2,
    APDE_ERR_STATE = -3,
    APDE_ERR_NOMEM = -4
} apde_status;

/* RunState, main.h:68-72 */
enum { APDE_FIRST_INIT = 0, APDE_REFINE_INIT = 1, APDE_REFINE_ITER = 2 };
/* PixelState, main.h:74-78 */
enum { APDE_WEAK = 0, APDE_STRONG = 1, APDE_UNKNOWN = 2 };

/* Camera, main.h:50-61 (identical layout, 120 bytes) */
typedef struct {
    float K[9];
    float R[9];
    float t[3];
    float c[3];
    int height;
    int width;
    float depth_min;
    float depth_max;
    float interval;
    float depth_num;
} apde_camera;

/* PatchMatchParams, main.h:80-100 (bools widened to int) */
typedef struct {
    int max_iterations;
    int num_images; /* filled by the library: N + 1 */
    int top_k;
    float depth_min; /* filled by the library: 0.6 * cam.depth_min, APD.cpp:554 */
    float depth_max; /* filled by the library: 1.2 * cam.depth_max, APD.cpp:555 */
    int geom_consistency;
    int use_impetus;
    int strong_radius;
    int strong_increment;
    int weak_radius;
    int weak_increment;
    int use_APD;
    int use_sa; /* consume the reference view's segment-label map (apde_view_set_sa_mask) in use_APD passes, APD.cpp:641-649 */
    int weak_peak_radius;
    int rotate_time;
    float ransac_threshold;
    float geom_factor;
    int state;
} apde_params;

/* the reference defaults, main.h:80-100 */
void apde_params_default(apde_params *p);

typedef struct apde_context apde_context;

int apde_create(int device, apde_context **out);
void apde_destroy(apde_context *ctx);
const char *apde_last_error(void);
/* library build info: "sm_100a;..." */
const char *apde_version(void);

/* ---------------------------------------------------------------- scene (resident images + cameras + maps) */
int apde_scene_begin(apde_context *ctx, int num_views, int width, int height);
/* gray: H*W bytes (what cv::imread(GRAYSCALE) yields, APD.cpp:145) ; bgr may be NULL (only fusion colours need it) */
int apde_scene_set_view(apde_context *ctx, int view, const uint8_t *gray, const uint8_t *bgr, const apde_camera *cam);
/* segment labels of a view: the CV_8UC1 map of <dense>/sa_masks/<id>.bin (tools/run_SAM.py:41-60; any size, 0 = no segment),
 * replaces APD.cpp:641-649 (ReadBinMat + nearest resize to the working size, done here per problem on the device).  Consumed
 * by passes with use_APD && use_sa: NCC-Old branch B (APD.cu:664-719) and the label tests of NCC-New (APD.cu:493-497, 526-530).
 * labels == NULL removes the map. */
int apde_view_set_sa_mask(apde_context *ctx, int view, const uint8_t *labels, int width, int height);
/* neighbour list of a view = Problem::src_image_ids (main.h:104), at most APDE_MAX_IMAGES-1 entries */
int apde_scene_set_pairs(apde_context *ctx, int view, int num_src, const int32_t *src_views);
int apde_scene_commit(apde_context *ctx);

/* maps of a view as the reference stores them in depths/normals/weak/confidence.bin.  normal = float[P][3].
 * Any pointer may be NULL.  width/height report (download) or give (upload) the map size. */
int apde_view_download(apde_context *ctx, int view, float *depth, float *normal, uint8_t *weak, uint8_t *conf,
                       int *width, int *height);
int apde_view_upload(apde_context *ctx, int view, const float *depth, const float *normal, const uint8_t *weak,
                     const uint8_t *conf, int width, int height);

/* ---------------------------------------------------------------- one problem = one reference view, one pass */
int apde_problem_setup(apde_context *ctx, int ref_view, const apde_params *params, int scale_size, uint32_t seed);

typedef enum {
    APDE_STAGE_NEAREST_STRONG = 0, /* FindNearestStrongPoint   APD.cu:2434 */
    APDE_STAGE_GEN_ANCHORS = 1,    /* GenAnchors + NeigbourUpdate APD.cu:1857, 2084 */
    APDE_STAGE_INIT = 2,           /* RandomInitialization     APD.cu:919 */
    APDE_STAGE_PROP_STRONG = 3,    /* Black/RedPixelUpdateStrong APD.cu:1654 */
    APDE_STAGE_RANSAC_FIT = 4,     /* RANSACToGetFitPlane      APD.cu:2486 */
    APDE_STAGE_PROP_WEAK = 5,      /* Black/RedPixelUpdateWeak APD.cu:1617 */
    APDE_STAGE_DEPTH_NORMAL = 6,   /* GetDepthandNormal        APD.cu:1694 */
    APDE_STAGE_MEDIAN = 7,         /* Black/RedPixelFilterStrong APD.cu:1823 */
    APDE_STAGE_DEPTH_TO_WEAK = 8,  /* DepthToWeak              APD.cu:2103 */
    APDE_STAGE_CONFIDENCE = 9,     /* ConfidenceCompute        APD.cu:2282 */
    APDE_STAGE_LOCAL_REFINE = 10   /* LocalRefine              APD.cu:2346 */
} apde_stage;

/* run one kernel of RunPatchMatch; color 0 = black ((x+y) even), 1 = red */
int apde_problem_stage(apde_context *ctx, int stage, int iter, int color);
/* run all of RunPatchMatch in the reference order */
int apde_problem_run(apde_context *ctx);

typedef enum {
    APDE_FIELD_PLANES = 0,         /* float4[P] */
    APDE_FIELD_COSTS = 1,          /* float[P] */
    APDE_FIELD_SELECTED_VIEWS = 2, /* uint32[P] */
    APDE_FIELD_VIEW_WEIGHT = 3,    /* uint8[P][32] (reference layout; stored packed on the device) */
    APDE_FIELD_WEAK_INFO = 4,      /* uint8[P] */
    APDE_FIELD_CONFIDENCE = 5,     /* uint8[P] */
    APDE_FIELD_FIT_PLANES = 6,     /* float4[P] */
    APDE_FIELD_WEAK_RELIABLE = 7,  /* uint8[P] */
    APDE_FIELD_NEAREST_STRONG = 8, /* short2[P] */
    APDE_FIELD_ANCHORS = 9,        /* short2[P][9] */
    APDE_FIELD_IMAGE = 10,         /* float[P], working-resolution reference image (read only) */
    APDE_FIELD_SRC_DEPTH = 11,     /* float[N+1][P], working-resolution depth maps, index 0 = ref (read only) */
    APDE_FIELD_RELIABLE_CURVE = 12,/* float[P][61], DepthToWeak's cost curve (read only; needs apde_problem_capture_curve) */
    APDE_FIELD_SA_MASK = 13        /* uint8[P], the reference view's segment labels at the working size (read only; error if none) */
} apde_field;

int apde_problem_get(apde_context *ctx, int field, void *host, size_t bytes);
int apde_problem_set(apde_context *ctx, int field, const void *host, size_t bytes);
int apde_problem_dims(apde_context *ctx, int *width, int *height, int *num_images);
/* working-resolution image of source index idx (0 = ref) and the working cameras, for parity checks */
int apde_problem_get_image(apde_context *ctx, int idx, float *host, size_t bytes);
int apde_problem_get_cameras(apde_context *ctx, apde_camera *cams /* [N+1] */, apde_params *params);
/* write depth/normal/weak/confidence back into the view store (depth range check of main.cpp:172-175) */
int apde_problem_finish(apde_context *ctx);

/* Keep DepthToWeak's 61-sample cost curve of the following problems (reliable_curve.bin, APD.cu:2714-2723, 2651-2661);
 * costs P * 61 floats of device memory while on.  Pixels DepthToWeak does not sweep hold zeros. */
int apde_problem_capture_curve(apde_context *ctx, int on);

/* setup + run + finish */
int apde_pass_run(apde_context *ctx, int ref_view, const apde_params *params, int scale_size, uint32_t seed);

/* parity hook: per-hypothesis costs of the current problem.  tuples = int32[n][3] (x, y, src_idx>=1),
 * planes = float[n][4] (camera-frame normal, plane distance).  mode 0 = NCC-Old, 1 = NCC-New, 2 = geometric. */
int apde_eval_costs(apde_context *ctx, int n, const int32_t *tuples, const float *planes, int mode, float *out);

/* parity hook: tex2D<float>(image of source index idx, x, y) exactly as the kernels sample it (linear filter, clamp,
 * unnormalised; texture set-up of APD.cpp:699-706).  xy = float[n][2]. */
int apde_debug_tex2d(apde_context *ctx, int idx, int n, const float *xy, float *out);

/* ---------------------------------------------------------------- whole schedule (main.cpp:303-367) */
typedef struct {
    int rounds;          /* <= 0: ComputeRoundNum rule (halve max side until <= 800), main.cpp:129-146 */
    int geom_iterations; /* reference: 3 */
    int jacobi;          /* 0: reference order (Gauss-Seidel through fresh maps); 1: all views read previous-pass maps */
    int use_impetus;
    float geom_factor;   /* 0.2 (0.05 for TaT), main.cpp:294-298 */
    uint32_t seed;
    int first_view, num_views_local; /* this rank's shard of reference views (multi-GPU); 0,0 = all */
    int use_sa;          /* main.cpp:324; only views that were given a label map are affected */
} apde_schedule;

void apde_schedule_default(apde_schedule *s);
/* the PatchMatchParams, scale and seed that pass `pass_index` of the schedule uses for every view (main.cpp:309-365), for
 * drivers that need to step through a pass view by view (apde_problem_setup / run / get / finish), e.g. to export the
 * anchors and the reliable curve of the last iteration as the reference does (main.cpp:353-358) */
int apde_schedule_pass_params(apde_context *ctx, const apde_schedule *s, int pass_index, apde_params *params, int *scale_size,
                              uint32_t *seed);

typedef struct {
    double patchmatch_ms; /* device time of all RunPatchMatch stages (CUDA events; == the reference's "RunPatchMatch time") */
    double device_ms;     /* device time of the whole pass incl. set-up / finish kernels (CUDA events on the stream) */
    double total_ms;      /* host wall clock of the same region, stream synchronised on both sides */
    uint64_t evals_ncc_old, evals_ncc_new, evals_geom;
    uint64_t kernel_launches;
    int passes;
    int pad_;
    double exchange_ms;       /* multi-GPU: device time the compute stream waited for the depth-map exchange (the exposed part) */
    uint64_t exchange_bytes;  /* multi-GPU: depth-map bytes this rank received */
} apde_timing;

int apde_run_schedule(apde_context *ctx, const apde_schedule *s, apde_timing *out);
/* run one pass (all views of the shard) of the schedule; pass_index counts from 0 as main.cpp's iteration_index */
int apde_run_schedule_pass(apde_context *ctx, const apde_schedule *s, int pass_index, apde_timing *out);
int apde_schedule_num_passes(apde_context *ctx, const apde_schedule *s);

/* cumulative device counters since the last reset: [0] NCC-Old evals, [1] NCC-New evals, [2] geom evals,
 * [3] kernel launches */
int apde_get_counters(apde_context *ctx, uint64_t out[4], int reset);
/* 3x3 anchor patches (9 samples each) sampled by the weak propagation since the last counter reset: with out[1] above the real
 * sample count of the deformable cost, 36 * out[1] + 9 * anchor_patches (APD.cu:514-563) */
int apde_get_anchor_evals(apde_context *ctx, uint64_t *anchor_patches);

/* per-kernel profile of the pass: when enabled, every stage launch is bracketed by CUDA events on the launching stream.
 * ms[11], launches[11], evals[11][3] are indexed by apde_stage. */
int apde_set_profiling(apde_context *ctx, int on);
/* DepthToWeak / LocalRefine keep 62 cost slots per (pixel, selected view) column in HBM; images whose worst case exceeds
 * this budget are processed in bands of rows (results are identical).  0 = default (48 GB, or APDE_SWEEP_BUDGET_MB). */
int apde_set_sweep_budget_mb(apde_context *ctx, size_t megabytes);
int apde_get_stage_stats(apde_context *ctx, double *ms, uint64_t *launches, uint64_t *evals, int reset);

/* roofline denominators measured on this GPU, in this process: FP32 FMA rate (TFLOP/s) and texture-unit bilinear gather
 * rate (10^9 filtered samples/s) on the current pyramid level.  Either pointer may be NULL. */
int apde_microbench(apde_context *ctx, double *fp32_tflops, double *tex_gsamples);
/* access-pattern study (thread-per-evaluation vs quad-per-evaluation, scattered hypotheses): 10^9 samples/s */
int apde_microbench_pattern(apde_context *ctx, int mode, float spread, double *tex_gsamples);

/* device pointer + byte size of the replicated depth-map pool ([V][P] float at the current map size) so that a
 * host-side collective (NCCL all-gather between passes) can exchange shards in place */
int apde_depth_pool(apde_context *ctx, void **dev_ptr, size_t *bytes, size_t *bytes_per_view);

/* ---------------------------------------------------------------- multi-GPU jobs (no counterpart in the reference,
 * which runs one process per GPU on disjoint scans, run.py:127-153)
 *
 * One context per GPU, every context holds the whole scene (images, cameras, pairs).  Reference views are dealt out in
 * contiguous blocks; within a pass every rank runs its own block, and each finished depth map is broadcast to the other
 * ranks over NCCL (NVLink / NVSwitch) while the next view of the block is being computed -- the maps a view reads from its
 * neighbours are those of the PREVIOUS pass (Jacobi order; main.cpp:309,336 is Gauss-Seidel through the file cache).
 * The contexts may live in one process (one host thread per context: `apd --gpus N`) or in one process per GPU (torchrun,
 * mpirun): rank 0 makes an id with apde_comm_create_id, the job hands its 128 bytes to every rank, every rank calls
 * apde_comm_init.  After that
 *   apde_run_schedule / apde_run_schedule_pass   run this rank's block and exchange the depth maps (collective calls);
 *   apde_exchange                                all-gathers the rows of one map pool (before fusion);
 *   apde_fuse_collective                         sharded WeakVisFilter + gathers + the greedy fusion on rank 0.
 * NCCL is loaded at run time (dlopen of libnccl.so.2) by apde_comm_create_id / apde_comm_init only: a single-GPU user needs none. */
#define APDE_COMM_ID_BYTES 128
int apde_comm_create_id(uint8_t id[APDE_COMM_ID_BYTES]);
int apde_comm_init(apde_context *ctx, const uint8_t id[APDE_COMM_ID_BYTES], int rank, int world);
/* how a job deals V views out: contiguous blocks, the first V % world ranks hold one view more (pure function, no GPU) */
int apde_comm_block_of(int num_views, int world, int rank, int *first_view, int *count);
/* rank, world and this rank's block [first_view, first_view + num_views) of the committed scene (any pointer may be NULL) */
int apde_comm_info(apde_context *ctx, int *rank, int *world, int *first_view, int *num_views);
/* collective: the rows of pool `which` (enum apde_pool) that belong to each rank's own views reach every rank, in place */
int apde_exchange(apde_context *ctx, int which);
/* collective: RunFusion / _TAT_I / _TAT_A over the maps of all ranks.  WeakVisFilter is sharded by reference view, the map and
 * skip pools are all-gathered, rank 0 runs the greedy fusion and receives the points; the other ranks get *num_points = 0. */
int apde_fuse_collective(apde_context *ctx, int variant, int use_weak_filter, float *xyz, float *bgr, int64_t max_points,
                         int64_t *num_points);
int apde_comm_destroy(apde_context *ctx);

/* ---------------------------------------------------------------- fusion (APD.cpp:962-1227) */
/* skip_weaks: uint8[V][P] out (host) or NULL to keep it on the device only */
int apde_weak_vis_filter(apde_context *ctx, uint8_t *skip_weaks);
/* RunFusion: returns the number of fused points through *num_points; xyz float[max][3], bgr float[max][3]
 * (either may be NULL to only count).  use_weak_filter mirrors the CLI flag (main.cpp:19). */
int apde_fuse(apde_context *ctx, int use_weak_filter, float *xyz, float *bgr, int64_t max_points,
              int64_t *num_points);
/* WeakVisFilter for reference views [first_view, first_view + num_views) only: the per-view work items of the reference's
 * thread pool (APD.cpp:1040-1047) sharded over GPUs.  skip_weaks: uint8[num_views][P] out (host) or NULL. */
int apde_weak_vis_filter_range(apde_context *ctx, int first_view, int num_views, uint8_t *skip_weaks);
/* use_weak_filter value for apde_fuse*: keep the skip maps that are already in the APDE_POOL_SKIP pool */
#define APDE_WEAK_FILTER_KEEP 2

/* Multi-GPU fusion: the maps of all views live in one contiguous device pool per field, [V][bytes_per_view]
 * (maps smaller than the full resolution are packed at the start of their slot).  A job with one process per GPU
 * all-gathers the rows of its own views in place (NCCL over NVLink), calls apde_views_mark_maps on every rank, shards
 * WeakVisFilter with apde_weak_vis_filter_range, all-gathers APDE_POOL_SKIP and lets one rank run apde_fuse_variant with
 * APDE_WEAK_FILTER_KEEP.  The pointer stays valid until the next apde_scene_begin with a different shape; for
 * APDE_POOL_DEPTH it is the pool the next pass reads (same as apde_depth_pool). */
enum apde_pool { APDE_POOL_DEPTH = 0, APDE_POOL_NORMAL = 1, APDE_POOL_WEAK = 2, APDE_POOL_CONFIDENCE = 3, APDE_POOL_SKIP = 4 };
int apde_map_pool(apde_context *ctx, int which, void **device_ptr, size_t *bytes_total, size_t *bytes_per_view);
/* declare that every view's slot in the pools now holds maps of width x height (after an all-gather of peer maps) */
int apde_views_mark_maps(apde_context *ctx, int width, int height);

/* The dataset-specific variants main.cpp:277-283 dispatches to: RunFusion (ETH3D and others), RunFusion_TAT_I
 * (APD.cpp:1229-1431) and RunFusion_TAT_A (APD.cpp:1433-1608). */
enum apde_fuse_kind { APDE_FUSE_DEFAULT = 0, APDE_FUSE_TAT_I = 1, APDE_FUSE_TAT_A = 2 };
int apde_fuse_variant(apde_context *ctx, int variant, int use_weak_filter, float *xyz, float *bgr, int64_t max_points,
                      int64_t *num_points);
/* A fusion call with xyz = bgr = NULL counts the points AND keeps the cloud in host memory of the context; this call copies it
 * out (up to max_points points; *num_points = size of the kept cloud) and releases it, so that a caller who sizes its buffers
 * from the count runs the fusion once.  APDE_ERR_STATE when no cloud is kept (after a scene change, a fusion call with
 * buffers, or a previous take).  After apde_fuse_collective only rank 0 holds a cloud. */
int apde_fuse_take_points(apde_context *ctx, float *xyz, float *bgr, int64_t max_points, int64_t *num_points);

#ifdef __cplusplus
}
#endif
#endif /* APDE_H_ */
