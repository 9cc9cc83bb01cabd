// apd_io.h -- host data plane of the drop-in surface: the reference's file formats without OpenCV / Boost.
//   ReadBinMat / WriteBinMat   ".bin" maps ("dmb": int32 version=1, rows, cols, cv type + raw rows)  APD.cpp:18-83
//   ReadCamera                 MVSNet *_cam.txt                                                     APD.cpp:85-135
//   ReadImage / ReadImageColor PNG (zlib inflate) / JPEG (apd_jpeg.cpp) / PGM / PPM, pixel-identical to cv::imread   APD.cpp:137-160
//   Show*                      depth_k.jpg / normal_k.jpg / weak_k.png / confidence_k.png (apd_show.cpp)         APD.cpp:162-314
//   ExportPointCloud           binary little-endian PLY, float xyz + uchar blue green red             APD.cpp:316-356
//   GenerateSampleList         pair.txt -> Problem list                                              main.cpp:44-102
#pragma once
#include <cstdint>
#include <filesystem>
#include <string>
#include <vector>

#include "../../../include/apde.h"

namespace apd {

using path = std::filesystem::path;
typedef apde_camera Camera;  // identical layout to the reference Camera (main.h:50-61)

// OpenCV type codes used by the reference's maps
enum { CV_8UC1 = 0, CV_8UC3 = 16, CV_32SC1 = 4, CV_32FC1 = 5, CV_32FC3 = 21 };

struct Mat {  // the subset of cv::Mat the driver interface touches
    int rows = 0, cols = 0, type_ = 0;
    std::vector<uint8_t> buf;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    void create(int r, int c, int type);
    int type() const { return type_; }
    bool empty() const { return buf.empty(); }
    size_t elem_size() const;
    size_t step() const { return (size_t)cols * elem_size(); }
    uint8_t *data() { return buf.data(); }
    const uint8_t *data() const { return buf.data(); }
    template <typename T> T &at(int r, int c) { return *reinterpret_cast<T *>(buf.data() + r * step() + c * sizeof(T)); }
    template <typename T> const T &at(int r, int c) const { return *reinterpret_cast<const T *>(buf.data() + r * step() + c * sizeof(T)); }
    template <typename T> T *ptr(int r = 0) { return reinterpret_cast<T *>(buf.data() + r * step()); }
    template <typename T> const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(buf.data() + r * step()); }
};

struct float3_ { float x, y, z; };
struct PointList { float3_ coord; float3_ color; };  // main.h:63-66

bool ReadBinMat(const path &mat_path, Mat &mat);
bool WriteBinMat(const path &mat_path, const Mat &mat);
bool ReadCamera(const path &cam_path, Camera &cam);
bool ReadImage(const path &img_path, Mat &gray_u8);         // CV_8UC1 (the library converts to float on the device)
bool ReadImageColor(const path &img_path, Mat &bgr_u8);     // CV_8UC3, BGR order as cv::imread(IMREAD_COLOR)
bool WritePGM(const path &p, const Mat &gray_u8);
bool WritePNG(const path &p, const Mat &img_u8);             // CV_8UC1 or CV_8UC3 (BGR), what cv::imwrite("*.png") stores
bool WriteJPEG(const path &p, const Mat &img_u8, int quality = 95);  // baseline JPEG, CV_8UC1 or CV_8UC3 (BGR)
// debug imagery of the last geometric iteration of a round (APD.cpp:162-314, main.cpp:191-204)
bool ShowDepthMap(const path &depth_path, const Mat &depth, float depth_min, float depth_max);
bool ShowNormalMap(const path &normal_path, const Mat &normal);
bool ShowWeakImage(const path &weak_path, const Mat &weak);
bool ShowConfidenceMap(const path &confidence_path, const Mat &confidence);
void JetColorMap(uint8_t bgr[256][3]);  // cv::COLORMAP_JET
// The four pictures of one view (main.cpp:191-204) written by worker threads while the caller goes on with the next GPU pass:
// encoding a 1920x1080 view takes 0.5-1.2 s of one core, more than the PatchMatch pass that produced it.  At most
// min(cores, 16) views are in flight, fewer when their maps would hold more than ~2 GB; wait() (also the destructor) joins.
class ShowWriter {
public:
    ShowWriter();
    ~ShowWriter();
    void submit(const path &result_folder, int iteration, Mat depth, Mat normal, Mat weak, Mat confidence, float depth_min, float depth_max);
    void wait();
private:
    struct Impl;
    Impl *impl;
};
bool ExportPointCloud(const path &ply_path, const std::vector<PointList> &pc, bool export_color = true);
std::string ToFormatIndex(int index);

struct ProblemDesc {  // main.h:102-115 (the fields the schedule needs)
    int ref_image_id = 0;
    std::vector<int> src_image_ids;
    path dense_folder, result_folder;
    std::string img_ext;
};
// pair.txt parser with the reference's rules: score <= 0 entries dropped, result folder APD/<id>/ created
bool GenerateSampleList(const path &dense_folder, std::vector<ProblemDesc> &problems, std::string *err);

// Everything a view brings from disk: grey and BGR image (ReadImage / ReadImageColor), camera (ReadCamera) and, when
// <dense>/sa_masks/<id>.bin exists, its segment-label map (CV_8UC1, tools/run_SAM.py:41-60; APD.cpp:507, 641-649)
struct LoadedView {
    Mat gray, bgr, sa;
    Camera cam;
    bool image_ok = false, cam_ok = false;
};
// views [first, first + count) of `problems`, decoded by one thread each (the reference's ThreadPool idea, APD.cpp:1040,
// applied to CheckImages / ReadImage, main.cpp:104-127); results in view order
void LoadViews(const path &dense_folder, const std::vector<ProblemDesc> &problems, size_t first, size_t count, std::vector<LoadedView> &out);
size_t LoadViewsBatchSize(size_t num_views);  // min(hardware threads, views, 16)

}  // namespace apd
