/* mock_apde.c -- TEST DOUBLE of libapde.so for the CPU tests of the host side (tests/test_cli_host_logic.py).
 *
 * It implements the entry points of include/apde.h that the `apd` CLI and the APD / RunFusion host code call, without any
 * computation: uploads are check-summed and reported on stdout, "passes" only count, downloads return synthetic maps that
 * depend on (view, x, y).  It lets the host logic -- scene loading on threads, the pass loop, Show* writers, .bin / .ply /
 * skip.png output -- run end to end on a machine without a GPU.  It is NOT a fallback: nothing in the product links or loads
 * it (the product's libapde.so fails loudly without a CUDA device); only the test builds a second binary against it. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "apde.h"

struct apde_context {
    int V, W, H;
    int passes;
    int *src_count;
    int rank, world;  /* multi-GPU job: 0 of 1 until apde_comm_init */
    int id;           /* creation order: tells the contexts of a job apart in the log */
    int kept_variant, kept_filter, kept;  /* cloud kept by a count-only fusion call */
};
static int g_created = 0;

static unsigned long long checksum(const uint8_t *p, size_t n) {
    unsigned long long a = 0;
    for (size_t i = 0; i < n; ++i) a = a * 1315423911ull + p[i];
    return a;
}

const char *apde_last_error(void) { return "mock"; }
const char *apde_version(void) { return "mock"; }
void apde_params_default(apde_params *p) { memset(p, 0, sizeof(*p)); p->top_k = 4; p->max_iterations = 3; p->use_sa = 1; }
void apde_schedule_default(apde_schedule *s) { memset(s, 0, sizeof(*s)); s->geom_iterations = 3; s->use_impetus = 1; s->geom_factor = 0.2f; s->seed = 1; s->use_sa = 1; }
int apde_create(int device, apde_context **out) {
    *out = (apde_context *)calloc(1, sizeof(apde_context));
    (*out)->world = 1;
    (*out)->id = g_created++;
    printf("MOCK create %d device %d\n", (*out)->id, device);
    return 0;
}
void apde_destroy(apde_context *c) { if (c) { free(c->src_count); free(c); } }
int apde_scene_begin(apde_context *c, int num_views, int width, int height) {
    c->V = num_views; c->W = width; c->H = height; c->passes = 0;
    c->src_count = (int *)calloc((size_t)num_views, sizeof(int));
    if (c->id == 0) printf("MOCK begin %d %d %d\n", num_views, width, height);
    else printf("MOCK ctx%d begin %d %d %d\n", c->id, num_views, width, height);
    return 0;
}
int apde_scene_set_view(apde_context *c, int view, const uint8_t *gray, const uint8_t *bgr, const apde_camera *cam) {
    const size_t P = (size_t)c->W * c->H;
    printf(c->id == 0 ? "MOCK view %d gray %llu bgr %llu fx %.9g dmin %.9g\n" : "MOCK ctx1+ view %d gray %llu bgr %llu fx %.9g dmin %.9g\n", view, checksum(gray, P), bgr ? checksum(bgr, 3 * P) : 0ull, cam->K[0], cam->depth_min);
    return 0;
}
int apde_view_set_sa_mask(apde_context *c, int view, const uint8_t *labels, int width, int height) {
    if (c->id == 0) printf("MOCK sa %d %dx%d %llu\n", view, width, height, labels ? checksum(labels, (size_t)width * height) : 0ull);
    return 0;
}
int apde_scene_set_pairs(apde_context *c, int view, int num_src, const int32_t *src_views) {
    c->src_count[view] = num_src;
    if (c->id != 0) return 0;
    printf("MOCK pairs %d:", view);
    for (int i = 0; i < num_src; ++i) printf(" %d", src_views[i]);
    printf("\n");
    return 0;
}
int apde_scene_commit(apde_context *c) { printf(c->id == 0 ? "MOCK commit\n" : "MOCK ctx1+ commit\n"); return 0; }

static int rounds_of(const apde_context *c) {  /* ComputeRoundNum, main.cpp:129-146 */
    int m = c->W > c->H ? c->W : c->H, r = 1;
    while (m > 800) { m /= 2; r++; }
    return r;
}
int apde_schedule_num_passes(apde_context *c, const apde_schedule *s) { return (s->rounds > 0 ? s->rounds : rounds_of(c)) * (1 + s->geom_iterations); }
int apde_run_schedule_pass(apde_context *c, const apde_schedule *s, int pass_index, apde_timing *out) {
    c->passes = pass_index + 1;
    out->patchmatch_ms += 1.0;
    out->passes += 1;
    if (c->world == 1) printf("MOCK pass %d use_sa %d geom_factor %.3g\n", pass_index, s->use_sa, s->geom_factor);
    else printf("MOCK rank %d of %d pass %d\n", c->rank, c->world, pass_index);
    return 0;
}
/* synthetic maps: any host-side mix-up of views, rows or fields shows */
int apde_view_download(apde_context *c, int view, float *depth, float *normal, uint8_t *weak, uint8_t *conf, int *width, int *height) {
    if (c->passes == 0) { if (width) *width = 0; if (height) *height = 0; return 0; }
    if (width) *width = c->W;
    if (height) *height = c->H;
    for (int y = 0; y < c->H; ++y)
        for (int x = 0; x < c->W; ++x) {
            const size_t i = (size_t)y * c->W + x;
            if (depth) depth[i] = 2.0f + 0.01f * x + 0.02f * y + 0.1f * view + 0.001f * c->passes;
            if (normal) { normal[3 * i] = 0.1f * (view % 3); normal[3 * i + 1] = 0.0f; normal[3 * i + 2] = -1.0f; }
            if (weak) weak[i] = (uint8_t)((x + y + view) % 3);
            if (conf) conf[i] = (uint8_t)((x * 3 + y + view) % 200);
        }
    return 0;
}
int apde_view_upload(apde_context *c, int view, const float *depth, const float *normal, const uint8_t *weak, const uint8_t *conf, int width, int height) {
    (void)normal; (void)weak; (void)conf;
    c->passes = 1;
    printf("MOCK upload %d %dx%d d00 %.6g\n", view, width, height, depth[0]);
    return 0;
}
int apde_weak_vis_filter(apde_context *c, uint8_t *skip_weaks) {
    const size_t P = (size_t)c->W * c->H;
    for (int v = 0; v < c->V; ++v)
        for (size_t i = 0; i < P; ++i) skip_weaks[v * P + i] = (uint8_t)((i + v) % 7 == 0);
    return 0;
}
/* ---- multi-GPU job entry points (apd --gpus N): the rendezvous and the block partition are real, the transfers are not */
int apde_comm_create_id(uint8_t id[APDE_COMM_ID_BYTES]) { for (int i = 0; i < APDE_COMM_ID_BYTES; ++i) id[i] = (uint8_t)(i * 3 + 1); return 0; }
int apde_comm_init(apde_context *c, const uint8_t id[APDE_COMM_ID_BYTES], int rank, int world) {
    for (int i = 0; i < APDE_COMM_ID_BYTES; ++i) if (id[i] != (uint8_t)(i * 3 + 1)) return -1;
    c->rank = rank; c->world = world;
    printf("MOCK comm_init ctx %d rank %d of %d\n", c->id, rank, world);
    return 0;
}
int apde_comm_block_of(int V, int world, int rank, int *first, int *count) {
    const int base = V / world, extra = V % world;
    *count = base + (rank < extra ? 1 : 0);
    *first = rank * base + (rank < extra ? rank : extra);
    return 0;
}
int apde_comm_info(apde_context *c, int *rank, int *world, int *first_view, int *num_views) {
    int f, n;
    apde_comm_block_of(c->V, c->world, c->rank, &f, &n);
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    if (first_view) *first_view = f;
    if (num_views) *num_views = n;
    return 0;
}
int apde_exchange(apde_context *c, int which) { printf("MOCK rank %d exchange %d\n", c->rank, which); return 0; }
int apde_views_mark_maps(apde_context *c, int width, int height) { printf("MOCK rank %d mark %dx%d\n", c->rank, width, height); return 0; }
int apde_weak_vis_filter_range(apde_context *c, int first_view, int num_views, uint8_t *skip_weaks) {
    const size_t P = (size_t)c->W * c->H;
    for (int v = 0; v < num_views; ++v)
        for (size_t i = 0; i < P; ++i) skip_weaks[v * P + i] = (uint8_t)((i + first_view + v) % 7 == 0);
    return 0;
}
int apde_comm_destroy(apde_context *c) { (void)c; return 0; }

static void mock_points(int variant, int use_weak_filter, float *xyz, float *bgr, int64_t n, int64_t max_points) {
    for (int64_t i = 0; i < n && i < max_points; ++i)
        for (int k = 0; k < 3; ++k) { xyz[3 * i + k] = (float)(i + 0.25 * k); bgr[3 * i + k] = (float)(10 * i + k + use_weak_filter); }
}
int apde_fuse_variant(apde_context *c, int variant, int use_weak_filter, float *xyz, float *bgr, int64_t max_points, int64_t *num_points) {
    const int64_t n = 5 + variant;
    printf("MOCK fusion run variant %d\n", variant);  /* the host side must run the fusion ONCE per cloud */
    c->kept = 0;
    if (xyz && bgr) mock_points(variant, use_weak_filter, xyz, bgr, n, max_points);
    else { c->kept = 1; c->kept_variant = variant; c->kept_filter = use_weak_filter; }
    *num_points = n;
    return 0;
}
int apde_fuse_take_points(apde_context *c, float *xyz, float *bgr, int64_t max_points, int64_t *num_points) {
    if (!c->kept) return -1;
    mock_points(c->kept_variant, c->kept_filter, xyz, bgr, 5 + c->kept_variant, max_points);
    *num_points = 5 + c->kept_variant;
    c->kept = 0;
    return 0;
}
int apde_fuse(apde_context *c, int use_weak_filter, float *xyz, float *bgr, int64_t max_points, int64_t *num_points) {
    return apde_fuse_variant(c, 0, use_weak_filter, xyz, bgr, max_points, num_points);
}
/* the view-by-view path (--export_anchor / --export_curve, the APD class facade) is exercised on the GPU only */
int apde_schedule_pass_params(apde_context *c, const apde_schedule *s, int pass_index, apde_params *params, int *scale_size, uint32_t *seed) {
    (void)c; (void)s; (void)pass_index; (void)params; (void)scale_size; (void)seed; return -1;
}
int apde_problem_setup(apde_context *c, int ref_view, const apde_params *params, int scale_size, uint32_t seed) { (void)c; (void)ref_view; (void)params; (void)scale_size; (void)seed; return -1; }
int apde_problem_run(apde_context *c) { (void)c; return -1; }
int apde_problem_finish(apde_context *c) { (void)c; return -1; }
int apde_problem_dims(apde_context *c, int *w, int *h, int *n) { (void)c; (void)w; (void)h; (void)n; return -1; }
int apde_problem_get(apde_context *c, int field, void *host, size_t bytes) { (void)c; (void)field; (void)host; (void)bytes; return -1; }
int apde_problem_get_cameras(apde_context *c, apde_camera *cams, apde_params *params) { (void)c; (void)cams; (void)params; return -1; }
int apde_problem_capture_curve(apde_context *c, int on) { (void)c; (void)on; return -1; }
