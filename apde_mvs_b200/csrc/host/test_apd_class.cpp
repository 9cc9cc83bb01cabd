// test_apd_class.cpp -- drives the APD class exactly like the reference's ProcessProblem (main.cpp:148-208)
#include <cmath>
#include <iostream>

#include "APD.h"
using namespace apd;

int main(int argc, char **argv) {
    if (argc < 2) { std::cout << "usage: test_apd_class <dense_folder>\n"; return 2; }
    auto s = SceneSession::get(argv[1]);
    int bad = 0;
    for (auto &pd : s->problems) {
        Problem problem;
        problem.ref_image_id = pd.ref_image_id; problem.src_image_ids = pd.src_image_ids;
        problem.dense_folder = pd.dense_folder; problem.result_folder = pd.result_folder; problem.img_ext = pd.img_ext;
        problem.params.state = FIRST_INIT; problem.params.use_APD = false; problem.params.geom_consistency = false;
        problem.params.weak_peak_radius = 6; problem.scale_size = 1; problem.iteration = 0;
        APD APD(problem);
        APD.InuputInitialization();
        APD.CudaSpaceInitialization();
        APD.SetDataPassHelperInCuda();
        APD.RunPatchMatch();
        const int width = APD.GetWidth(), height = APD.GetHeight();
        Mat depth(height, width, CV_32FC1), normal(height, width, CV_32FC3);
        Mat pixel_states = APD.GetPixelStates();
        int valid = 0;
        for (int r = 0; r < height; ++r)
            for (int c = 0; c < width; ++c) {
                const float4_ ph = APD.GetPlaneHypothesis(r, c);
                depth.at<float>(r, c) = ph.w;
                if (ph.w < APD.GetDepthMin() || ph.w > APD.GetDepthMax()) { depth.at<float>(r, c) = 0; pixel_states.at<uint8_t>(r, c) = UNKNOWN; }
                else valid++;
                float *n = &normal.at<float>(r, 3 * c);
                n[0] = ph.x; n[1] = ph.y; n[2] = ph.z;
                const float len = std::sqrt(ph.x * ph.x + ph.y * ph.y + ph.z * ph.z);
                if (std::fabs(len - 1.0f) > 1e-2f) bad++;
            }
        WriteBinMat(problem.result_folder / "depths.bin", depth);
        WriteBinMat(problem.result_folder / "normals.bin", normal);
        WriteBinMat(problem.result_folder / "weak.bin", pixel_states);
        APD.Commit();
        std::cout << "image " << pd.ref_image_id << ": " << width << "x" << height << " valid depth " << valid << " of " << width * height << std::endl;
        if (valid < width * height / 2) bad++;
    }
    std::cout << (bad ? "FAILED" : "OK") << std::endl;
    SceneSession::release_all();
    return bad ? 1 : 0;
}
