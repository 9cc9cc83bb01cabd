// ref_host_wrapper.cpp -- TEST INFRASTRUCTURE.  The reference's own HOST code (/root/reference/APD.cpp: ReadCamera,
// ReadBinMat / WriteBinMat, WeakVisFilter, RunFusion, RunFusion_TAT_I / _TAT_A, ExportPointCloud) compiled unmodified,
// included from where it lies, against the stub OpenCV / Boost headers of oracle/ref_stubs, with C entry points for the tests.
// Built by `make -C oracle ref_host` into oracle/_ref/libapd_ref_host.so; runs on the CPU (the CUDA runtime is linked only
// because the same file defines the APD class; none of it is called here).
#include "APD.cpp"

// main.cpp rides along for GenerateSampleList / ComputeRoundNum; its ProcessProblem refers to the one APD method that lives
// in APD.cu, which this CPU-only library does not contain and never calls
void APD::RunPatchMatch() {}
#define main reference_main
#include "main.cpp"
#undef main

extern "C" {

// the reference's GenerateSampleList (main.cpp:44-102): ref ids, CSR source ids (score <= 0 dropped), image extension of view 0
int ref_generate_sample_list(const char *dense_folder, int max_views, int max_src, int *ref_ids, int *src_offsets, int *src_ids, char *ext16) {
    std::vector<Problem> problems;
    GenerateSampleList(path(dense_folder), problems);
    if ((int)problems.size() > max_views) return -1;
    int k = 0;
    src_offsets[0] = 0;
    for (size_t i = 0; i < problems.size(); ++i) {
        ref_ids[i] = problems[i].ref_image_id;
        for (int s : problems[i].src_image_ids) { if (k >= max_src) return -2; src_ids[k++] = s; }
        src_offsets[i + 1] = k;
    }
    if (!problems.empty()) { strncpy(ext16, problems[0].img_ext.c_str(), 15); ext16[15] = 0; }
    return (int)problems.size();
}

// the reference's ComputeRoundNum (main.cpp:129-146) on the folder's first image
int ref_compute_round_num(const char *dense_folder) {
    std::vector<Problem> problems;
    GenerateSampleList(path(dense_folder), problems);
    return ComputeRoundNum(problems);
}

// the reference's ReadCamera (APD.cpp:85-135): out = K[9] R[9] t[3] c[3] height width depth_min depth_max interval depth_num
int ref_read_camera(const char *cam_path, float *out) {
    Camera cam;
    memset(&cam, 0, sizeof(cam));
    if (!ReadCamera(path(cam_path), cam)) return -1;
    for (int i = 0; i < 9; ++i) { out[i] = cam.K[i]; out[9 + i] = cam.R[i]; }
    for (int i = 0; i < 3; ++i) { out[18 + i] = cam.t[i]; out[21 + i] = cam.c[i]; }
    out[24] = (float)cam.height; out[25] = (float)cam.width;
    out[26] = cam.depth_min; out[27] = cam.depth_max; out[28] = cam.interval; out[29] = cam.depth_num;
    return 0;
}

// round trip through the reference's ReadBinMat / WriteBinMat (APD.cpp:18-83)
int ref_copy_bin_mat(const char *src, const char *dst) {
    cv::Mat m;
    if (!ReadBinMat(path(src), m)) return -1;
    if (!WriteBinMat(path(dst), m, true)) return -2;
    return m.type();
}

// RunFusion (variant 0), RunFusion_TAT_I (1), RunFusion_TAT_A (2) on <dense>: cams/, images/<id><ext>, APD/<id>/{depths,
// normals,weak,confidence}.bin; writes <dense>/APD/<name>.  src ids in CSR form.
int ref_run_fusion(const char *dense_folder, int num_views, const int *ref_ids, const int *src_offsets, const int *src_ids,
                   const char *img_ext, int variant, int weak_filter, int export_color, const char *name) {
    std::vector<Problem> problems(num_views);
    const path dense(dense_folder);
    for (int i = 0; i < num_views; ++i) {
        Problem &p = problems[i];
        p.ref_image_id = ref_ids[i];
        for (int k = src_offsets[i]; k < src_offsets[i + 1]; ++k) p.src_image_ids.push_back(src_ids[k]);
        p.dense_folder = dense;
        p.result_folder = dense / path("APD") / path(ToFormatIndex(ref_ids[i]));
        p.img_ext = img_ext;
        p.iteration = 0;
        p.used_time = 0;
    }
    if (variant == 1) RunFusion_TAT_I(dense, problems, name, weak_filter != 0, export_color != 0);
    else if (variant == 2) RunFusion_TAT_A(dense, problems, name, weak_filter != 0, export_color != 0);
    else RunFusion(dense, problems, name, weak_filter != 0, export_color != 0);
    return 0;
}

}  // extern "C"
