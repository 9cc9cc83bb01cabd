// APD.h -- the reference's problem / driver interface (APD.h:88-114, main.h:80-115) on top of libapde's C ABI.
//
// Same class name, same method names and call order as the reference, so main.cpp-style code keeps compiling:
//     APD apd(problem); apd.InuputInitialization(); apd.CudaSpaceInitialization(); apd.SetDataPassHelperInCuda();
//     apd.RunPatchMatch(); apd.GetPlaneHypothesis(r, c); apd.GetPixelStates(); apd.GetConfidence(); ...
// Differences that a caller can observe: cv::Mat is apd::Mat (OpenCV C++ is not available in this image); images, cameras
// and every view's maps stay resident on the GPU inside a per-dense-folder SceneSession instead of being re-read from
// disk / MemoryCache for every problem (APD.cpp:501-685); CUDA errors throw std::runtime_error instead of exit().
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "apd_io.h"

namespace apd {

enum RunState { FIRST_INIT, REFINE_INIT, REFINE_ITER };  // main.h:68-72
enum PixelState { WEAK, STRONG, UNKNOWN };               // main.h:74-78

struct PatchMatchParams {  // main.h:80-100
    int max_iterations = 3;
    int num_images = 5;
    int top_k = 4;
    float depth_min = 0.0f;
    float depth_max = 1.0f;
    bool geom_consistency = false;
    bool use_impetus = true;
    int strong_radius = 5;
    int strong_increment = 2;
    int weak_radius = 5;
    int weak_increment = 5;
    bool use_APD = true;
    bool use_sa = true;
    int weak_peak_radius = 2;
    int rotate_time = 4;
    float ransac_threshold = 0.005f;
    float geom_factor = 0.2f;
    RunState state = FIRST_INIT;
};

struct Problem {  // main.h:102-115
    int ref_image_id = 0;
    std::vector<int> src_image_ids;
    path dense_folder;
    path result_folder;
    int scale_size = 1;
    PatchMatchParams params;
    bool show_medium_result = false;
    bool export_anchor = false;
    bool export_reliable_curve = false;
    int iteration = 0;
    std::string img_ext;
    int used_time = 0;
};

struct float4_ { float x, y, z, w; };

// One GPU-resident scene per dense folder (replaces MemoryCache, APD.cpp:3-16): images + cameras uploaded once.
class SceneSession {
public:
    static std::shared_ptr<SceneSession> get(const path &dense_folder, int gpu_index = 0);
    // a multi-GPU job over one scene (`apd --gpus N`): one context per listed GPU, every context holds the whole scene (views
    // are decoded once and uploaded to each); the contexts are connected with apde_comm_init by the threads that drive them
    static std::shared_ptr<SceneSession> get_job(const path &dense_folder, const std::vector<int> &gpus);
    static void release_all();
    ~SceneSession();
    apde_context *ctx = nullptr;        // == ctxs[0]
    std::vector<apde_context *> ctxs;   // one per GPU of the job (size 1 without --gpus)
    std::vector<ProblemDesc> problems;
    std::vector<int> id_to_view;  // image id -> view index (-1: unknown)
    int width = 0, height = 0;
    int view_of(int image_id) const;
    std::vector<Camera> cameras;
    bool has_color = false;
    int num_sa_masks = 0;  // views with a label map in <dense>/sa_masks/
private:
    SceneSession() {}
};

class APD {
public:
    explicit APD(const Problem &problem);
    ~APD();
    void InuputInitialization();
    void CudaSpaceInitialization();
    void SetDataPassHelperInCuda();
    void RunPatchMatch();
    float4_ GetPlaneHypothesis(int r, int c);
    Mat GetPixelStates();
    Mat GetConfidence();
    int GetWidth();
    int GetHeight();
    float GetDepthMin();
    float GetDepthMax();
    // extension: hand the results to the resident view store (ProcessProblem's tail, main.cpp:168-190, without file I/O)
    void Commit();

private:
    void download();
    Problem problem;
    std::shared_ptr<SceneSession> session;
    apde_params params_c;
    int width = 0, height = 0;
    bool set_up = false, ran = false, downloaded = false, committed = false;
    std::vector<float4_> planes;
    Mat weak, conf;
};

// RunFusion (APD.cpp:1051-1227): WeakVisFilter + greedy fusion on the GPU, PLY written by ExportPointCloud
void RunFusion(const path &dense_folder, const std::vector<Problem> &problems, const std::string &name = "APD.ply",
               bool weak_filter = true, bool export_color = true);
// Tanks-and-Temples variants (APD.cpp:1229-1431 intermediate set, 1433-1608 advanced set), dispatched by main.cpp:277-283
void RunFusion_TAT_I(const path &dense_folder, const std::vector<Problem> &problems, const std::string &name = "APD.ply",
                     bool weak_filter = true, bool export_color = true);
void RunFusion_TAT_A(const path &dense_folder, const std::vector<Problem> &problems, const std::string &name = "APD.ply",
                     bool weak_filter = true, bool export_color = true);
// The fusion of a multi-GPU job, called by EVERY rank's thread with its own context: WeakVisFilter of the rank's own views
// (skip.png written by their owner), the gathers and, on rank 0, the greedy fusion and the PLY.  variant = enum apde_fuse_kind.
void RunFusionJob(SceneSession &s, int rank, int variant, const std::string &name, bool weak_filter, bool export_color);

}  // namespace apd
