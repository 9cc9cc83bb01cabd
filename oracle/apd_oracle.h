/*
 * apd_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, never shipped, never on the product path).
 *
 * A plain single-file C++ restatement of the reference's per-pixel deformable-PatchMatch pass
 * (/root/reference/APD.cu) and of its CPU fusion (/root/reference/APD.cpp).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load the library built from it.
 *
 * PARITY PIN: the reference ships no tests / golden vectors, so this oracle is
 * pinned against outputs of the reference's own device functions, produced on a B200 by
 * oracle/ref_wrapper.cu (which #includes /root/reference/APD.cu unmodified) and committed as
 * tests/golden/ref_*.npz together with the generating script (tests/golden/make_ref_golden.py).  The HOST half (ReadCamera,
 * the .bin map format, WeakVisFilter, RunFusion, RunFusion_TAT_I / _TAT_A, PLY export) is pinned against the reference's own
 * APD.cpp, compiled unmodified into oracle/_ref/libapd_ref_host.so (oracle/ref_host_wrapper.cpp) and run on the CPU by
 * tests/test_ref_host.py: identical skip maps, points, order and colours.
 *
 * Deviations from the reference that are deliberate and shared with the CUDA product:
 *   - RNG: Philox4x32-10 counter RNG keyed by (seed, stream), counter = (pixel, site, block) instead
 *     of per-pixel XORWOW seeded by clock64() (APD.cu:916).  Whole-pass parity with the reference is
 *     therefore statistical; per-function parity is exact / 1e-4.
 *   - running SAM is out of scope; a label map it produced (sa_masks/<id>.bin) is consumed as the reference does:
 *     orc_problem::sa_mask != NULL switches on NCC-Old branch B (APD.cu:664-719) and the label tests of NCC-New
 *     (APD.cu:493-497, 526-530); NULL = all-zero labels (APD.cpp:613).
 *   - anchors are stored densely ([pixel][9]) instead of through anchors_map (APD.cpp:627-640).
 */
#ifndef APD_ORACLE_H_
#define APD_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_IMAGES 32
#define ORC_ANCHOR_NUM 9

enum { ORC_FIRST_INIT = 0, ORC_REFINE_INIT = 1, ORC_REFINE_ITER = 2 };  /* main.h:68-72 */
enum { ORC_WEAK = 0, ORC_STRONG = 1, ORC_UNKNOWN = 2 };                 /* main.h:74-78 */

/* same layout as the reference Camera (main.h:50-61), 120 bytes */
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
} orc_camera;

/* PatchMatchParams (main.h:80-100) with bools widened to int */
typedef struct {
    int max_iterations;
    int num_images;
    int top_k;
    float depth_min;
    float depth_max;
    int geom_consistency;
    int use_impetus;
    int strong_radius;
    int strong_increment;
    int weak_radius;
    int weak_increment;
    int use_APD;
    int use_sa;
    int weak_peak_radius;
    int rotate_time;
    float ransac_threshold;
    float geom_factor;
    int state;
} orc_params;

typedef struct {
    int width, height;
    int num_images;                       /* N+1, index 0 = reference view */
    const float *images[ORC_MAX_IMAGES];  /* H*W float, 0..255 */
    const float *depths[ORC_MAX_IMAGES];  /* H*W float or NULL */
    orc_camera cameras[ORC_MAX_IMAGES];
    orc_params params;
    float *planes;            /* float4[P] */
    float *costs;             /* [P] */
    uint32_t *selected_views; /* [P] */
    uint8_t *view_weight;     /* [P*32] */
    uint8_t *weak_info;       /* [P] */
    uint8_t *confidence;      /* [P] */
    float *fit_planes;        /* float4[P] */
    uint8_t *weak_reliable;   /* [P] */
    int16_t *nearest_strong;  /* short2[P] */
    int16_t *anchors;         /* short2[P*9] */
    uint32_t seed;
    uint32_t stream;
    int tex_mode;             /* 0 = exact fp32 bilinear, 1 = 8-bit weight quantisation (CUDA texture unit) */
    int num_threads;          /* OpenMP threads, <=0 -> 1 */
    uint64_t counters[8];     /* [0] NCC-Old evals, [1] NCC-New evals, [2] geom evals */
    const uint8_t *sa_mask;   /* [P] segment labels of the reference view (sa_masks/<id>.bin, APD.cpp:641-649) or NULL = all zero */
} orc_problem;

/* ---- per-hypothesis cost functions (APD.cu:334-403, 448-721, 831-902) ---- */
void orc_homography(const orc_camera *ref, const orc_camera *src, const float plane[4], float H[9]);
float orc_ncc_old(orc_problem *pb, int x, int y, int src_idx, const float plane[4]);
float orc_ncc_new(orc_problem *pb, int x, int y, int src_idx, const float plane[4]);
float orc_geom_cost(orc_problem *pb, int x, int y, int src_idx, const float plane[4]);
float orc_tex2d(const float *img, int w, int h, float x, float y, int tex_mode);
/* batched: tuples (x, y, src_idx) int32[n*3], planes float[n*4]; mode 0 = old, 1 = new, 2 = geom */
void orc_eval_costs(orc_problem *pb, int n, const int32_t *tuples, const float *planes, int mode, float *out);

/* ---- stages, in RunPatchMatch order (APD.cu:2663-2737) ---- */
void orc_nearest_strong(orc_problem *pb);                    /* K2  APD.cu:2434-2484 */
void orc_gen_anchors(orc_problem *pb);                       /* K3  APD.cu:1857-2082 */
void orc_neighbour_update(orc_problem *pb);                  /* K4  APD.cu:2084-2100 */
void orc_random_init(orc_problem *pb);                       /* K5  APD.cu:919-948 */
void orc_propagate_strong(orc_problem *pb, int iter, int color); /* K6, color 0 = black */
void orc_ransac_fit(orc_problem *pb, int iter);              /* K7  APD.cu:2486-2598 */
void orc_propagate_weak(orc_problem *pb, int iter, int color);   /* K8 */
void orc_depth_normal(orc_problem *pb);                      /* K9  APD.cu:1694-1709 */
void orc_median_filter(orc_problem *pb, int color);          /* K10 APD.cu:1711-1855 */
void orc_depth_to_weak(orc_problem *pb, float *curve);       /* K11 APD.cu:2103-2250 */
void orc_confidence(orc_problem *pb);                        /* K12 APD.cu:2282-2344 */
void orc_local_refine(orc_problem *pb);                      /* K13 APD.cu:2346-2432 */
void orc_run_pass(orc_problem *pb);                          /* whole RunPatchMatch */

/* candidate index set of the adaptive checkerboard (APD.cu:1127-1314): positions[8], flags[8] */
void orc_checkerboard_candidates(const float *costs, int w, int h, int x, int y, int32_t positions[8],
                                 uint8_t flags[8]);

/* RNG spec shared with the CUDA product */
void orc_philox(uint32_t seed, uint32_t stream, uint32_t pixel, uint32_t site, uint32_t block, uint32_t out[4]);

/* ---- fusion (APD.cpp:844-910, 962-1227) ---- */
typedef struct {
    int num_views;
    int width, height;
    const orc_camera *cameras;   /* [V] already rescaled to the map size */
    const float *depths;         /* [V][P] */
    const float *normals;        /* [V][P][3] */
    const uint8_t *weaks;        /* [V][P] */
    const uint8_t *confidences;  /* [V][P] */
    const uint8_t *colors;       /* [V][P][3] BGR, or NULL */
    const int32_t *src_offsets;  /* [V+1] CSR into src_ids */
    const int32_t *src_ids;      /* neighbour view indices (0..V-1) */
    int num_threads;
} orc_fusion_input;

void orc_weak_vis_filter(const orc_fusion_input *in, uint8_t *skip_weaks /* [V][P] */);
/* returns number of points; writes at most max_points entries: xyz float[3], bgr float[3] */
int64_t orc_fuse(const orc_fusion_input *in, const uint8_t *skip_weaks, float *points_xyz, float *points_bgr,
                 int64_t max_points);

/* RunFusion_TAT_I (variant 1, APD.cpp:1229-1431) / RunFusion_TAT_A (variant 2, APD.cpp:1433-1608) */
int64_t orc_fuse_tat(const orc_fusion_input *in, const uint8_t *skip_weaks, int variant, float *points_xyz, float *points_bgr,
                     int64_t max_points);

#ifdef __cplusplus
}
#endif
#endif
