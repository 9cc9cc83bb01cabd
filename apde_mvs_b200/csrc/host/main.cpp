// main.cpp -- the reference's CLI (main.cpp:7-41, 210-411) on the B200 library: same flag spellings (booleans take a value),
// same round x iteration x problem schedule, same output files (APD/%08d/{depths,normals,weak,confidence}.bin, APD/APD.ply).
#include <atomic>
#include <chrono>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <thread>

#include "APD.h"

using namespace apd;

struct Args {
    std::string dense_folder, dataset = "DTU";
    int gpu_index = 0;
    int gpus = 1;               // extension: one scene over this many GPUs (gpu_index, gpu_index + 1, ...), NCCL between them
    std::vector<int> gpu_list;  // extension: ... or over exactly these devices
    bool jacobi = false;        // extension: every view reads the PREVIOUS pass's maps of its neighbours (what a multi-GPU job does)
    bool only_fuse = false, no_fuse = false, memory_cache = true, use_sa = true, use_impetus = true, weak_filter = true, flush = false,
         export_anchor = false, export_curve = false, export_color = true;
};

static void usage() {
    std::cout << "Allowed options:\n  -d [ --dense_folder ] arg  path to dense folder\n  -g [ --gpu_index ] arg (=0)\n"
                 "  -D [ --dataset ] arg (=DTU)  DTU, ETH3D, TaT_a, TaT_i, General\n  -f [ --only_fuse ] arg (=0)\n  -F [ --no_fuse ] arg (=0)\n"
                 "  -m [ --memory_cache ] arg (=1)\n  -s [ --use_sa ] arg (=1)\n  -i [ --use_impetus ] arg (=1)\n  -w [ --weak_filter ] arg (=1)\n"
                 "  --flush arg (=0)\n  -n [ --export_anchor ] arg (=0)\n  -r [ --export_curve ] arg (=0)\n  -c [ --export_color ] arg (=1)\n  -h [ --help ]\n"
                 "  --gpus arg (=1)            one scene over N GPUs (gpu_index .. gpu_index + N - 1)\n  --gpu_list arg             ... or over the listed devices, e.g. 0,2,3\n"
                 "  --view_order arg (=reference)  reference: a view sees the maps its neighbours finished earlier in the same pass (main.cpp:309,336);\n"
                 "                             jacobi: only those of the previous pass (always so with --gpus > 1)\n";
}
static bool to_bool(const std::string &v) { return v == "1" || v == "true" || v == "True" || v == "yes" || v == "on"; }

static Args parse(int argc, char **argv) {
    Args a;
    std::map<std::string, std::string> alias = {{"-d", "--dense_folder"}, {"-g", "--gpu_index"}, {"-D", "--dataset"}, {"-f", "--only_fuse"},
        {"-F", "--no_fuse"}, {"-m", "--memory_cache"}, {"-s", "--use_sa"}, {"-i", "--use_impetus"}, {"-w", "--weak_filter"},
        {"-n", "--export_anchor"}, {"-r", "--export_curve"}, {"-c", "--export_color"}, {"-h", "--help"}};
    for (int i = 1; i < argc; ++i) {
        std::string k = argv[i], v;
        const size_t eq = k.find('=');
        if (eq != std::string::npos) { v = k.substr(eq + 1); k = k.substr(0, eq); }
        if (alias.count(k)) k = alias[k];
        if (k == "--help") { usage(); exit(0); }
        if (v.empty()) {
            if (i + 1 >= argc) { std::cout << "Error: the required argument for option '" << k << "' is missing\n"; usage(); exit(-1); }
            v = argv[++i];
        }
        if (k == "--dense_folder") a.dense_folder = v;
        else if (k == "--gpu_index") a.gpu_index = atoi(v.c_str());
        else if (k == "--dataset") a.dataset = v;
        else if (k == "--gpus") a.gpus = atoi(v.c_str());
        else if (k == "--view_order") a.jacobi = (v == "jacobi");
        else if (k == "--gpu_list") {
            std::stringstream ss(v);
            std::string tok;
            while (std::getline(ss, tok, ',')) if (!tok.empty()) a.gpu_list.push_back(atoi(tok.c_str()));
        }
        else if (k == "--only_fuse") a.only_fuse = to_bool(v);
        else if (k == "--no_fuse") a.no_fuse = to_bool(v);
        else if (k == "--memory_cache") a.memory_cache = to_bool(v);
        else if (k == "--use_sa") a.use_sa = to_bool(v);
        else if (k == "--use_impetus") a.use_impetus = to_bool(v);
        else if (k == "--weak_filter") a.weak_filter = to_bool(v);
        else if (k == "--flush") a.flush = to_bool(v);
        else if (k == "--export_anchor") a.export_anchor = to_bool(v);
        else if (k == "--export_curve") a.export_curve = to_bool(v);
        else if (k == "--export_color") a.export_color = to_bool(v);
        else { std::cout << "Error: unrecognised option '" << k << "'\n"; usage(); exit(-1); }
    }
    if (a.dense_folder.empty()) { std::cout << "Error: the option '--dense_folder' is required but missing\n"; usage(); exit(-1); }
    if (a.gpu_list.empty()) for (int g = 0; g < std::max(a.gpus, 1); ++g) a.gpu_list.push_back(a.gpu_index + g);
    a.gpus = (int)a.gpu_list.size();
    a.gpu_index = a.gpu_list[0];
    return a;
}

// maps of the views [first, first + count) as this context holds them (a rank of a multi-GPU job writes its own block)
static void write_maps(SceneSession &s, apde_context *ctx, size_t first, size_t count) {
    for (size_t i = first; i < first + count; ++i) {
        int w = 0, h = 0;
        apde_view_download(ctx, (int)i, nullptr, nullptr, nullptr, nullptr, &w, &h);
        if (w == 0) continue;
        Mat depth(h, w, CV_32FC1), normal(h, w, CV_32FC3), weak(h, w, CV_8UC1), conf(h, w, CV_8UC1);
        apde_view_download(ctx, (int)i, depth.ptr<float>(), normal.ptr<float>(), weak.data(), conf.data(), &w, &h);
        const path dir = s.problems[i].result_folder;
        WriteBinMat(dir / "depths.bin", depth);
        WriteBinMat(dir / "normals.bin", normal);
        WriteBinMat(dir / "weak.bin", weak);
        WriteBinMat(dir / "confidence.bin", conf);
    }
}

// problem.show_medium_result (main.cpp:191-204, 359): pictures of every view's maps after the last geometric iteration of a round.
// Every view's maps are those of this pass once the pass has ended, so writing them here equals writing them problem by problem.
static void write_show(SceneSession &s, apde_context *ctx, size_t first, size_t count, int iteration, ShowWriter &writer) {
    if (getenv("APDE_NO_SHOW")) return;
    for (size_t i = first; i < first + count; ++i) {
        int w = 0, h = 0;
        apde_view_download(ctx, (int)i, nullptr, nullptr, nullptr, nullptr, &w, &h);
        if (w == 0) continue;
        Mat depth(h, w, CV_32FC1), normal(h, w, CV_32FC3), weak(h, w, CV_8UC1), conf(h, w, CV_8UC1);
        apde_view_download(ctx, (int)i, depth.ptr<float>(), normal.ptr<float>(), weak.data(), conf.data(), &w, &h);
        const Camera &cam = s.cameras[i];
        // encoded and written by worker threads while the next round runs on the GPU; depth range as APD.cpp:554-555
        writer.submit(s.problems[i].result_folder, iteration, std::move(depth), std::move(normal), std::move(weak), std::move(conf),
                      cam.depth_min * 0.6f, cam.depth_max * 1.2f);
    }
}

// The last iteration with --export_anchor / --export_curve, view by view (main.cpp:353-358, APD.cu:2614-2627, 2651-2661, 2692-2723):
//   anchors_map.bin  CV_32SC1: running index of the pixels that were WEAK when the pass started, -1 elsewhere (APD.cpp:627-640)
//   anchors.bin      int weak_count, int 9, short2[weak_count][9] (x, y; -1 = empty slot)
//   reliable_curve.bin  int width, int height, int 61, float[height * width][61]: DepthToWeak's cost curve
static bool run_pass_with_exports(SceneSession &s, const apde_schedule &sched, int pass, bool export_anchor, bool export_curve) {
    apde_params prm;
    int scale = 1;
    uint32_t seed = 0;
    if (apde_schedule_pass_params(s.ctx, &sched, pass, &prm, &scale, &seed)) return false;
    if (export_curve && apde_problem_capture_curve(s.ctx, 1)) return false;
    for (size_t i = 0; i < s.problems.size(); ++i) {
        int mw = 0, mh = 0;
        apde_view_download(s.ctx, (int)i, nullptr, nullptr, nullptr, nullptr, &mw, &mh);
        Mat weak_in(mh, mw, CV_8UC1);
        if (mw) apde_view_download(s.ctx, (int)i, nullptr, nullptr, weak_in.data(), nullptr, &mw, &mh);
        if (apde_problem_setup(s.ctx, (int)i, &prm, scale, seed) || apde_problem_run(s.ctx)) return false;
        int w = 0, h = 0, n = 0;
        apde_problem_dims(s.ctx, &w, &h, &n);
        const path dir = s.problems[i].result_folder;
        if (export_anchor && prm.use_APD && mw == w && mh == h) {
            std::vector<int16_t> dense((size_t)w * h * 9 * 2);
            if (apde_problem_get(s.ctx, APDE_FIELD_ANCHORS, dense.data(), dense.size() * sizeof(int16_t))) return false;
            Mat amap(h, w, CV_32SC1);
            std::vector<int16_t> packed;
            int weak_count = 0;
            for (int px = 0; px < w * h; ++px) {
                if (weak_in.data()[px] == APDE_WEAK) {
                    amap.ptr<int>()[px] = weak_count++;
                    packed.insert(packed.end(), dense.begin() + (size_t)px * 18, dense.begin() + (size_t)(px + 1) * 18);
                } else {
                    amap.ptr<int>()[px] = -1;
                }
            }
            WriteBinMat(dir / "anchors_map.bin", amap);
            std::ofstream out(dir / "anchors.bin", std::ios::binary);
            const int num = 9;
            out.write((const char *)&weak_count, sizeof(int));
            out.write((const char *)&num, sizeof(int));
            out.write((const char *)packed.data(), (std::streamsize)(packed.size() * sizeof(int16_t)));
        }
        if (export_curve) {
            std::vector<float> curve((size_t)w * h * 61);
            if (apde_problem_get(s.ctx, APDE_FIELD_RELIABLE_CURVE, curve.data(), curve.size() * sizeof(float))) return false;
            std::ofstream out(dir / "reliable_curve.bin", std::ios::binary);
            const int num = 61;
            out.write((const char *)&w, sizeof(int));
            out.write((const char *)&h, sizeof(int));
            out.write((const char *)&num, sizeof(int));
            out.write((const char *)curve.data(), (std::streamsize)(curve.size() * sizeof(float)));
        }
        if (apde_problem_finish(s.ctx)) return false;
    }
    if (export_curve) apde_problem_capture_curve(s.ctx, 0);
    return true;
}

static bool read_maps(SceneSession &s) {  // --only_fuse: depth maps come from a previous run
    for (size_t i = 0; i < s.problems.size(); ++i) {
        const path dir = s.problems[i].result_folder;
        Mat depth, normal, weak, conf;
        if (!ReadBinMat(dir / "depths.bin", depth) || !ReadBinMat(dir / "normals.bin", normal) || !ReadBinMat(dir / "weak.bin", weak) ||
            !ReadBinMat(dir / "confidence.bin", conf)) return false;
        // the checks of RunFusion's loader (APD.cpp:1106-1119), plus the element types and the image bound the upload relies on;
        // the reference skips such a view and fuses the rest -- here the maps of all views sit in one pool, so the run stops
        auto bad = [&](const char *what) { std::cout << "Error: " << what << " (" << dir.string() << ")" << std::endl; return false; };
        if (depth.type() != CV_32FC1 || normal.type() != CV_32FC3 || weak.type() != CV_8UC1 || conf.type() != CV_8UC1) return bad("unexpected element type in the maps of a previous run");
        if (normal.cols != depth.cols || normal.rows != depth.rows) return bad("normal size is not equal to depth size");
        if (weak.cols != depth.cols || weak.rows != depth.rows) return bad("weak size is not equal to depth size");
        if (conf.cols != depth.cols || conf.rows != depth.rows) return bad("confidence size is not equal to depth size");
        if (depth.cols < 1 || depth.rows < 1 || depth.cols > s.width || depth.rows > s.height) return bad("depth map is empty or larger than the image");
        if (apde_view_upload(s.ctx, (int)i, depth.ptr<float>(), normal.ptr<float>(), weak.data(), conf.data(), depth.cols, depth.rows)) return false;
    }
    return true;
}

// `apd --gpus N`: ONE scene over N GPUs of this node.  Every GPU gets a context holding the whole scene and a host thread;
// the threads connect their contexts into a job (apde_comm_init: NCCL over NVLink / NVSwitch) and then run main()'s schedule
// (main.cpp:303-367) side by side -- every apde_run_schedule_pass is a collective call in which a rank runs its own block of
// reference views and the finished depth maps are broadcast to the other ranks while the next view is computed.  The maps a
// view reads from its neighbours are those of the previous pass (Jacobi order), on one GPU count as on any other.
static void say(const std::string &line) {  // whole lines from concurrent rank threads
    fputs((line + "\n").c_str(), stdout);
    fflush(stdout);
}

static int run_job(const Args &a) {
    if (a.export_anchor || a.export_curve) {
        std::cout << "Error: --export_anchor / --export_curve step through the last pass view by view on one GPU; run them without --gpus" << std::endl;
        return EXIT_FAILURE;
    }
    const int N = a.gpus;
    auto s = SceneSession::get_job(a.dense_folder, a.gpu_list);
    std::cout << "There are " << s->problems.size() << " problems needed to be processed! (" << N << " GPUs)" << std::endl;
    apde_schedule sched;
    apde_schedule_default(&sched);
    sched.use_impetus = a.use_impetus;
    sched.use_sa = a.use_sa;
    if (a.use_sa && s->num_sa_masks == 0) std::cout << "Can't find sa mask folder: " << (path(a.dense_folder) / "sa_masks") << std::endl;
    else if (a.use_sa) std::cout << "sa masks: " << s->num_sa_masks << " of " << s->problems.size() << " views" << std::endl;
    sched.geom_factor = (a.dataset == "TaT_a" || a.dataset == "TaT_i") ? 0.05f : 0.2f;  // main.cpp:294-298
    const int npass = apde_schedule_num_passes(s->ctx, &sched);
    std::cout << "Round nums: " << npass / (1 + sched.geom_iterations) << std::endl;
    uint8_t id[APDE_COMM_ID_BYTES];
    if (apde_comm_create_id(id)) { std::cout << "Error: " << apde_last_error() << std::endl; return EXIT_FAILURE; }
    const int variant = a.dataset == "TaT_a" ? APDE_FUSE_TAT_A : (a.dataset == "TaT_i" ? APDE_FUSE_TAT_I : APDE_FUSE_DEFAULT);
    std::vector<apde_timing> timing((size_t)N);
    const auto start = std::chrono::steady_clock::now();
    auto worker = [&](int r) {
        // a rank that fails would leave its peers waiting inside a collective: report and end the process
        auto die = [&](const std::string &what) {
            say("Error (GPU " + std::to_string(a.gpu_list[(size_t)r]) + "): " + what);
            std::_Exit(EXIT_FAILURE);
        };
        try {
            apde_context *ctx = s->ctxs[(size_t)r];
            if (apde_comm_init(ctx, id, r, N)) die(apde_last_error());
            int first = 0, count = 0;
            apde_comm_info(ctx, nullptr, nullptr, &first, &count);
            apde_timing &t = timing[(size_t)r];
            memset(&t, 0, sizeof(t));
            ShowWriter show_writer;
            for (int p = 0; p < npass; ++p) {
                if (r == 0 && p % (1 + sched.geom_iterations) == 0)
                    say("========================== Round " + std::to_string(p / (1 + sched.geom_iterations)) + " ==========================");
                if (r == 0) say("======== iteration " + std::to_string(p) + "========");
                const double before = t.patchmatch_ms;
                if (apde_run_schedule_pass(ctx, &sched, p, &t)) die(apde_last_error());
                if (r == 0) say("RunPatchMatch time: " + std::to_string((int)(t.patchmatch_ms - before)) + " ms (" + std::to_string(count) + " of " +
                                std::to_string(s->problems.size()) + " views on GPU " + std::to_string(a.gpu_list[0]) + ")");
                if (p % (1 + sched.geom_iterations) == sched.geom_iterations) write_show(*s, ctx, (size_t)first, (size_t)count, p, show_writer);
            }
            write_maps(*s, ctx, (size_t)first, (size_t)count);
            if (!a.no_fuse) {
                if (r == 0) say("Run fusion");
                RunFusionJob(*s, r, variant, "APD.ply", a.weak_filter, a.export_color);
            }
        } catch (const std::exception &e) {
            die(e.what());
        }
    };
    std::vector<std::thread> threads;
    for (int r = 0; r < N; ++r) threads.emplace_back(worker, r);
    for (auto &th : threads) th.join();
    const auto end = std::chrono::steady_clock::now();
    double pm = 0, exch = 0;
    uint64_t old_e = 0, new_e = 0, geo_e = 0, bytes = 0;
    for (auto &t : timing) { pm += t.patchmatch_ms; exch = std::max(exch, t.exchange_ms); old_e += t.evals_ncc_old; new_e += t.evals_ncc_new; geo_e += t.evals_geom; bytes += t.exchange_bytes; }
    std::cout << "Cost time: " << std::chrono::duration_cast<std::chrono::milliseconds>(end - start).count() << " ms" << std::endl;
    std::cout << "Average used time: " << (int)(pm / s->problems.size()) << " ms" << std::endl;
    std::cout << "Cost evaluations: " << old_e << " NCC, " << new_e << " deformable NCC, " << geo_e << " geometric" << std::endl;
    std::cout << "Depth-map exchange: " << bytes / 1000000 << " MB received over NCCL, " << (int)exch << " ms not hidden behind compute" << std::endl;
    std::cout << (a.no_fuse ? "Skip fusion, all done!\n" : "All done\n");
    SceneSession::release_all();
    return EXIT_SUCCESS;
}

int main(int argc, char **argv) {
    Args a = parse(argc, argv);
    // every return path below frees the contexts while the CUDA runtime is still up (not at static-destruction time)
    struct Release { ~Release() { SceneSession::release_all(); } } release_on_exit;
    if (a.only_fuse) a.memory_cache = false;
    if (a.no_fuse) a.flush = true;
    std::cout << "========================== Config ==========================" << std::endl;
    std::cout << "dense_folder : " << a.dense_folder << "\ngpu_index    : " << a.gpu_index << "\ngpus         : " << a.gpus << "\ndataset      : " << a.dataset
              << "\nonly_fuse    : " << a.only_fuse << "\nno_fuse      : " << a.no_fuse << "\nmemory_cache : " << a.memory_cache
              << "\nuse_sa       : " << a.use_sa << "\nuse_impetus  : " << a.use_impetus
              << "\nweak_filter  : " << a.weak_filter << "\nflush        : " << a.flush << "\nexport_anchor: " << a.export_anchor
              << "\nexport_curve : " << a.export_curve << "\nexport_color : " << a.export_color << std::endl;
    std::cout << "============================================================" << std::endl;
    try {
        std::filesystem::create_directories(path(a.dense_folder) / "APD");
        if (a.gpus > 1 && !a.only_fuse) return run_job(a);
        auto s = SceneSession::get(a.dense_folder, a.gpu_index);
        std::cout << "There are " << s->problems.size() << " problems needed to be processed!" << std::endl;
        auto fuse = [&]() {  // main.cpp:277-283 / 402-408
            if (a.dataset == "TaT_a") RunFusion_TAT_A(a.dense_folder, {}, "APD.ply", a.weak_filter, a.export_color);
            else if (a.dataset == "TaT_i") RunFusion_TAT_I(a.dense_folder, {}, "APD.ply", a.weak_filter, a.export_color);
            else RunFusion(a.dense_folder, {}, "APD.ply", a.weak_filter, a.export_color);
        };
        if (a.only_fuse) {
            if (!read_maps(*s)) { std::cout << "Error: can not read the depth maps of a previous run" << std::endl; return EXIT_FAILURE; }
            fuse();
            printf("Fusion done!\n");
            return EXIT_SUCCESS;
        }
        apde_schedule sched;
        apde_schedule_default(&sched);
        sched.use_impetus = a.use_impetus;
        sched.jacobi = a.jacobi ? 1 : 0;
        sched.use_sa = a.use_sa;  // label maps of <dense>/sa_masks/ (read by SceneSession), main.cpp:324
        if (a.use_sa && s->num_sa_masks == 0) std::cout << "Can't find sa mask folder: " << (path(a.dense_folder) / "sa_masks") << std::endl;
        else if (a.use_sa) std::cout << "sa masks: " << s->num_sa_masks << " of " << s->problems.size() << " views" << std::endl;
        sched.geom_factor = (a.dataset == "TaT_a" || a.dataset == "TaT_i") ? 0.05f : 0.2f;  // main.cpp:294-298
        const int npass = apde_schedule_num_passes(s->ctx, &sched);
        std::cout << "Round nums: " << npass / (1 + sched.geom_iterations) << std::endl;
        apde_timing t;
        memset(&t, 0, sizeof(t));
        ShowWriter show_writer;  // joins its worker threads when it goes out of scope (every return path below)
        const auto start = std::chrono::steady_clock::now();
        for (int p = 0; p < npass; ++p) {
            if (p % (1 + sched.geom_iterations) == 0)
                std::cout << "========================== Round " << p / (1 + sched.geom_iterations) << " ==========================" << std::endl;
            std::cout << "======== iteration " << p << "========" << std::endl;
            const double before = t.patchmatch_ms;
            if (p == npass - 1 && (a.export_anchor || a.export_curve)) {  // is_last_iteration, main.cpp:335
                if (!run_pass_with_exports(*s, sched, p, a.export_anchor, a.export_curve)) { std::cout << "Error: " << apde_last_error() << std::endl; return EXIT_FAILURE; }
                write_show(*s, s->ctx, 0, s->problems.size(), p, show_writer);
                continue;
            }
            if (apde_run_schedule_pass(s->ctx, &sched, p, &t)) { std::cout << "Error: " << apde_last_error() << std::endl; return EXIT_FAILURE; }
            printf("RunPatchMatch time: %d ms (all %zu views)\n", (int)(t.patchmatch_ms - before), s->problems.size());
            if (p % (1 + sched.geom_iterations) == sched.geom_iterations) write_show(*s, s->ctx, 0, s->problems.size(), p, show_writer);
        }
        const auto end = std::chrono::steady_clock::now();
        std::cout << "Cost time: " << std::chrono::duration_cast<std::chrono::milliseconds>(end - start).count() << " ms" << std::endl;
        std::cout << "Average used time: " << (int)(t.patchmatch_ms / s->problems.size()) << " ms" << std::endl;
        std::cout << "Cost evaluations: " << t.evals_ncc_old << " NCC, " << t.evals_ncc_new << " deformable NCC, " << t.evals_geom << " geometric" << std::endl;
        write_maps(*s, s->ctx, 0, s->problems.size());  // the reference writes after every pass; nothing reads them in between when maps stay resident
        if (a.no_fuse) { printf("Skip fusion, all done!\n"); return EXIT_SUCCESS; }
        std::cout << "Run fusion\n";
        fuse();
        std::cout << "All done\n";
    } catch (const std::exception &e) {
        std::cout << "Error: " << e.what() << std::endl;
        return EXIT_FAILURE;
    }
    SceneSession::release_all();
    return EXIT_SUCCESS;
}
