#include "APD.h"

#include <cstring>
#include <iostream>
#include <map>
#include <stdexcept>

namespace apd {

static void check(int rc, const char *what) {
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + apde_last_error());
}

static std::map<std::string, std::shared_ptr<SceneSession>> g_sessions;

std::shared_ptr<SceneSession> SceneSession::get(const path &dense_folder, int gpu_index) {
    const std::string key = std::filesystem::weakly_canonical(dense_folder).string();
    auto it = g_sessions.find(key);
    if (it != g_sessions.end()) return it->second;
    return get_job(dense_folder, {gpu_index});
}

std::shared_ptr<SceneSession> SceneSession::get_job(const path &dense_folder, const std::vector<int> &gpus) {
    const std::string key = std::filesystem::weakly_canonical(dense_folder).string();
    if (gpus.empty()) throw std::runtime_error("no GPU given");
    std::shared_ptr<SceneSession> s(new SceneSession());
    std::string err;
    if (!GenerateSampleList(dense_folder, s->problems, &err)) throw std::runtime_error(err);
    if (s->problems.empty()) throw std::runtime_error("no problems in " + (dense_folder / "pair.txt").string());
    for (int g : gpus) {
        apde_context *c = nullptr;
        check(apde_create(g, &c), "apde_create");
        s->ctxs.push_back(c);
    }
    s->ctx = s->ctxs[0];
    int max_id = 0;
    for (auto &p : s->problems) max_id = std::max(max_id, p.ref_image_id);
    s->id_to_view.assign(max_id + 1, -1);
    for (size_t i = 0; i < s->problems.size(); ++i) s->id_to_view[s->problems[i].ref_image_id] = (int)i;
    // Decoding is the host cost of a scene (a 24-megapixel JPEG takes ~1.5 s for its grey and colour versions): views are read
    // by a batch of threads (LoadViews, apd_io.cpp) and uploaded in view order; one batch of decoded views is alive at a time.
    const size_t n_views = s->problems.size();
    const size_t batch_size = LoadViewsBatchSize(n_views);
    bool begun = false;
    for (size_t first = 0; first < n_views; first += batch_size) {
        std::vector<LoadedView> batch;
        LoadViews(dense_folder, s->problems, first, std::min(batch_size, n_views - first), batch);
        for (size_t k = 0; k < batch.size(); ++k) {
            const size_t i = first + k;
            const auto &p = s->problems[i];
            LoadedView &l = batch[k];
            if (!l.image_ok) throw std::runtime_error("Images may error, check it!");
            if (!begun) {
                s->width = l.gray.cols; s->height = l.gray.rows;
                for (apde_context *c : s->ctxs) check(apde_scene_begin(c, (int)n_views, s->width, s->height), "apde_scene_begin");
                begun = true;
            } else if (l.gray.cols != s->width || l.gray.rows != s->height) {
                throw std::runtime_error("Images may error, check it!");  // CheckImages, main.cpp:104-127
            }
            if (!l.cam_ok) throw std::runtime_error("can not read camera of image " + ToFormatIndex(p.ref_image_id));
            l.cam.width = s->width; l.cam.height = s->height;
            s->cameras.push_back(l.cam);
            for (apde_context *c : s->ctxs) {  // every GPU of a job holds the whole scene
                check(apde_scene_set_view(c, (int)i, l.gray.data(), l.bgr.empty() ? nullptr : l.bgr.data(), &l.cam), "apde_scene_set_view");
                if (!l.sa.empty()) check(apde_view_set_sa_mask(c, (int)i, l.sa.data(), l.sa.cols, l.sa.rows), "apde_view_set_sa_mask");
            }
            s->has_color = !l.bgr.empty();
            if (!l.sa.empty()) s->num_sa_masks++;
        }
    }
    for (size_t i = 0; i < s->problems.size(); ++i) {
        std::vector<int32_t> src;
        for (int id : s->problems[i].src_image_ids) {
            const int v = s->view_of(id);
            if (v < 0) throw std::runtime_error("pair.txt names image " + std::to_string(id) + " that is not a reference view");
            src.push_back(v);
        }
        for (apde_context *c : s->ctxs) check(apde_scene_set_pairs(c, (int)i, (int)src.size(), src.data()), "apde_scene_set_pairs");
    }
    for (apde_context *c : s->ctxs) check(apde_scene_commit(c), "apde_scene_commit");
    g_sessions[key] = s;
    return s;
}

void SceneSession::release_all() { g_sessions.clear(); }
SceneSession::~SceneSession() {
    for (apde_context *c : ctxs) apde_destroy(c);
}
int SceneSession::view_of(int image_id) const { return (image_id >= 0 && image_id < (int)id_to_view.size()) ? id_to_view[image_id] : -1; }

APD::APD(const Problem &problem_) : problem(problem_) {
    apde_params_default(&params_c);
    const PatchMatchParams &p = problem.params;
    params_c.max_iterations = p.max_iterations; params_c.top_k = p.top_k;
    params_c.geom_consistency = p.geom_consistency; params_c.use_impetus = p.use_impetus;
    params_c.strong_radius = p.strong_radius; params_c.strong_increment = p.strong_increment;
    params_c.weak_radius = p.weak_radius; params_c.weak_increment = p.weak_increment;
    params_c.use_APD = p.use_APD; params_c.use_sa = p.use_sa; params_c.weak_peak_radius = p.weak_peak_radius;
    params_c.rotate_time = p.rotate_time; params_c.ransac_threshold = p.ransac_threshold; params_c.geom_factor = p.geom_factor;
    params_c.state = (int)p.state;
}
APD::~APD() {}

void APD::InuputInitialization() {  // APD.cpp:501-685: everything it loads is already resident
    session = SceneSession::get(problem.dense_folder);
    if (session->view_of(problem.ref_image_id) < 0) throw std::runtime_error("unknown reference image id");
    std::cout << "Num images: " << problem.src_image_ids.size() + 1 << std::endl;
}
void APD::CudaSpaceInitialization() {  // APD.cpp:687-788
    if (!session) InuputInitialization();
    const uint32_t seed = 0x9E3779B1u + (uint32_t)problem.iteration * 0x85EBCA77u;
    check(apde_problem_setup(session->ctx, session->view_of(problem.ref_image_id), &params_c, problem.scale_size, seed), "apde_problem_setup");
    int n = 0;
    check(apde_problem_dims(session->ctx, &width, &height, &n), "apde_problem_dims");
    apde_params filled;
    check(apde_problem_get_cameras(session->ctx, nullptr, &filled), "apde_problem_get_cameras");
    params_c = filled;
    std::cout << "Depth range: " << params_c.depth_min << " " << params_c.depth_max << std::endl;
    std::cout << "Image size: " << width << " * " << height << std::endl;
    set_up = true;
}
void APD::SetDataPassHelperInCuda() {}  // APD.cpp:790-814: kernel arguments are passed by value at launch
void APD::RunPatchMatch() {             // APD.cu:2663-2737
    if (!set_up) CudaSpaceInitialization();
    check(apde_problem_run(session->ctx), "apde_problem_run");
    ran = true;
}
void APD::download() {
    if (downloaded) return;
    if (!ran) throw std::runtime_error("RunPatchMatch has not run");
    const size_t P = (size_t)width * height;
    planes.resize(P);
    weak.create(height, width, CV_8UC1);
    conf.create(height, width, CV_8UC1);
    check(apde_problem_get(session->ctx, APDE_FIELD_PLANES, planes.data(), P * 16), "planes");
    check(apde_problem_get(session->ctx, APDE_FIELD_WEAK_INFO, weak.data(), P), "weak");
    check(apde_problem_get(session->ctx, APDE_FIELD_CONFIDENCE, conf.data(), P), "confidence");
    downloaded = true;
}
float4_ APD::GetPlaneHypothesis(int r, int c) { download(); return planes[(size_t)c + (size_t)r * width]; }
Mat APD::GetPixelStates() { download(); return weak; }
Mat APD::GetConfidence() { download(); return conf; }
int APD::GetWidth() { return width; }
int APD::GetHeight() { return height; }
float APD::GetDepthMin() { return params_c.depth_min; }
float APD::GetDepthMax() { return params_c.depth_max; }
void APD::Commit() {
    if (committed || !ran) return;
    check(apde_problem_finish(session->ctx), "apde_problem_finish");
    committed = true;
}

static void run_fusion_variant(int variant, const path &dense_folder, const std::string &name, bool weak_filter, bool export_color) {
    auto s = SceneSession::get(dense_folder);
    int64_t n = 0;
    int filter_mode = 0;
    if (weak_filter) {  // WeakVisFilter once; its masks also go to <dense>/APD/<id>/skip.png as in the reference (APD.cpp:1026-1036)
        int mw = 0, mh = 0;
        check(apde_view_download(s->ctx, 0, nullptr, nullptr, nullptr, nullptr, &mw, &mh), "apde_view_download(dims)");
        const size_t P = (size_t)mw * mh;
        std::vector<uint8_t> skip(P * s->problems.size());
        check(apde_weak_vis_filter(s->ctx, skip.data()), "apde_weak_vis_filter");
        for (size_t v = 0; v < s->problems.size(); ++v) {
            Mat img(mh, mw, CV_8UC1);
            for (size_t i = 0; i < P; ++i) img.data()[i] = skip[v * P + i] == 1 ? 255 : 0;
            WritePNG(s->problems[v].result_folder / "skip.png", img);
        }
        filter_mode = APDE_WEAK_FILTER_KEEP;
    }
    // one fusion: the count-only call keeps the cloud in the context, apde_fuse_take_points hands it over
    check(apde_fuse_variant(s->ctx, variant, filter_mode, nullptr, nullptr, 0, &n), "apde_fuse");
    std::vector<float> xyz((size_t)n * 3), bgr((size_t)n * 3);
    int64_t n2 = 0;
    check(apde_fuse_take_points(s->ctx, xyz.data(), bgr.data(), n, &n2), "apde_fuse_take_points");
    std::vector<PointList> pc((size_t)std::min(n, n2));
    for (size_t i = 0; i < pc.size(); ++i) {
        pc[i].coord = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        pc[i].color = {bgr[3 * i], bgr[3 * i + 1], bgr[3 * i + 2]};
    }
    ExportPointCloud(dense_folder / "APD" / name, pc, export_color && s->has_color);
    std::cout << "Fused " << pc.size() << " points" << std::endl;
}

void RunFusionJob(SceneSession &s, int rank, int variant, const std::string &name, bool weak_filter, bool export_color) {
    apde_context *c = s.ctxs.at((size_t)rank);
    int first = 0, count = 0;
    check(apde_comm_info(c, nullptr, nullptr, &first, &count), "apde_comm_info");
    // normals, states and confidences of every view on every rank (depth maps already are: exchanged pass by pass)
    for (int which : {APDE_POOL_NORMAL, APDE_POOL_WEAK, APDE_POOL_CONFIDENCE}) check(apde_exchange(c, which), "apde_exchange");
    int mw = 0, mh = 0;
    check(apde_view_download(c, count > 0 ? first : 0, nullptr, nullptr, nullptr, nullptr, &mw, &mh), "apde_view_download(dims)");
    if (count == 0) { mw = s.width; mh = s.height; }
    check(apde_views_mark_maps(c, mw, mh), "apde_views_mark_maps");
    int filter_mode = 0;
    if (weak_filter) {  // the rank's own views: WeakVisFilter's work items (APD.cpp:1040-1047), skip.png by their owner (APD.cpp:1026-1036)
        const size_t P = (size_t)mw * mh;
        std::vector<uint8_t> skip(P * (size_t)std::max(count, 1));
        check(apde_weak_vis_filter_range(c, first, count, skip.data()), "apde_weak_vis_filter_range");
        for (int k = 0; k < count; ++k) {
            Mat img(mh, mw, CV_8UC1);
            for (size_t i = 0; i < P; ++i) img.data()[i] = skip[(size_t)k * P + i] == 1 ? 255 : 0;
            WritePNG(s.problems[(size_t)(first + k)].result_folder / "skip.png", img);
        }
        check(apde_exchange(c, APDE_POOL_SKIP), "apde_exchange(skip)");
        filter_mode = APDE_WEAK_FILTER_KEEP;
    }
    if (rank != 0) return;
    // the greedy claim order runs over all views through masks[] (APD.cpp:1149,1176,1209): one rank, over the gathered maps
    int64_t n = 0, n2 = 0;
    check(apde_fuse_variant(c, variant, filter_mode, nullptr, nullptr, 0, &n), "apde_fuse");
    std::vector<float> xyz((size_t)n * 3), bgr((size_t)n * 3);
    check(apde_fuse_take_points(c, xyz.data(), bgr.data(), n, &n2), "apde_fuse_take_points");
    std::vector<PointList> pc((size_t)std::min(n, n2));
    for (size_t i = 0; i < pc.size(); ++i) {
        pc[i].coord = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        pc[i].color = {bgr[3 * i], bgr[3 * i + 1], bgr[3 * i + 2]};
    }
    ExportPointCloud(s.problems[0].dense_folder / "APD" / name, pc, export_color && s.has_color);
    std::cout << ("Fused " + std::to_string(pc.size()) + " points\n") << std::flush;
}

void RunFusion(const path &dense_folder, const std::vector<Problem> &, const std::string &name, bool weak_filter, bool export_color) {
    run_fusion_variant(APDE_FUSE_DEFAULT, dense_folder, name, weak_filter, export_color);
}
void RunFusion_TAT_I(const path &dense_folder, const std::vector<Problem> &, const std::string &name, bool weak_filter, bool export_color) {
    run_fusion_variant(APDE_FUSE_TAT_I, dense_folder, name, weak_filter, export_color);
}
void RunFusion_TAT_A(const path &dense_folder, const std::vector<Problem> &, const std::string &name, bool weak_filter, bool export_color) {
    run_fusion_variant(APDE_FUSE_TAT_A, dense_folder, name, weak_filter, export_color);
}

}  // namespace apd
