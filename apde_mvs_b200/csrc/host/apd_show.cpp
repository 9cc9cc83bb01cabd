// apd_show.cpp -- the reference's debug imagery (APD.cpp:162-314), written on the last geometric iteration of every round
// (main.cpp:191-204, 359): depth_<it>.jpg (jet ramp over mean +- 2 sigma), normal_<it>.jpg, weak_<it>.png, confidence_<it>.png.
// Pixel values follow the reference's arithmetic; the .jpg files are baseline JPEGs from apd_jpeg.cpp's encoder (quality 95
// as cv::imwrite's default), so they decode to the same picture up to JPEG quantisation, not to the same bytes as libjpeg's.
#include <algorithm>
#include <array>
#include <cmath>
#include <deque>
#include <fstream>
#include <memory>
#include <thread>

#include "apd_io.h"

namespace apd {

bool encode_jpeg(int w, int h, int channels, const uint8_t *pix, int quality, std::vector<uint8_t> &out);  // apd_jpeg.cpp

bool WriteJPEG(const path &p, const Mat &img, int quality) {
    if (img.empty() || (img.type() != CV_8UC1 && img.type() != CV_8UC3)) return false;
    std::vector<uint8_t> bytes;
    if (!encode_jpeg(img.cols, img.rows, img.type() == CV_8UC3 ? 3 : 1, img.data(), quality, bytes)) return false;
    std::ofstream out(p, std::ios::binary);
    out.write((const char *)bytes.data(), (std::streamsize)bytes.size());
    return out.good();
}

// cv::COLORMAP_JET as OpenCV 4's 256-entry table comes out: three trapezoids with 64-entry plateaus and slopes of 4 per entry
// (blue starts half way up, red ends half way down); the off-by-one values at the joints (254, 2, 1) are OpenCV's.
// Checked entry for entry against cv2.applyColorMap in tests/test_host_io.py.
void JetColorMap(uint8_t bgr[256][3]) {
    for (int i = 0; i < 256; ++i) {
        bgr[i][0] = (uint8_t)(i <= 31 ? 128 + 4 * i : i <= 95 ? 255 : i <= 158 ? 254 - 4 * (i - 96) : i == 159 ? 1 : 0);
        bgr[i][1] = (uint8_t)(i <= 32 ? 0 : i <= 95 ? 4 * (i - 32) : i <= 159 ? 255 : i <= 223 ? 252 - 4 * (i - 160) : 0);
        bgr[i][2] = (uint8_t)(i <= 95 ? 0 : i <= 159 ? 2 + 4 * (i - 96) : i <= 223 ? 255 : 252 - 4 * (i - 224));
    }
}

bool ShowDepthMap(const path &depth_path, const Mat &depth, float depth_min, float depth_max) {
    if (depth.empty() || depth.type() != CV_32FC1) return false;
    auto usable = [&](float d) { return !(d < depth_min || d > depth_max || std::isnan(d)); };
    // mean of the column means, then the root of the mean of the column variances (APD.cpp:164-204): float accumulators
    float mean = 0, var = 0;
    int mean_count = 0, var_count = 0;
    for (int i = 0; i < depth.cols; ++i) {
        float sum = 0;
        int cnt = 0;
        for (int j = 0; j < depth.rows; ++j) {
            const float d = depth.at<float>(j, i);
            if (usable(d)) { sum += d; ++cnt; }
        }
        if (cnt > 0) { mean += sum / cnt; ++mean_count; }
    }
    if (mean_count > 0) mean /= mean_count;
    for (int i = 0; i < depth.cols; ++i) {
        float sum = 0;
        int cnt = 0;
        for (int j = 0; j < depth.rows; ++j) {
            const float d = depth.at<float>(j, i);
            if (usable(d)) { sum = (float)(sum + std::pow((double)(d - mean), 2)); ++cnt; }
        }
        if (cnt > 0) { var += sum / cnt; ++var_count; }
    }
    const float sigma = var_count > 0 ? std::sqrt(var / var_count) : 0.0f;
    const float lo = mean - sigma * 2, delta = (mean + sigma * 2) - lo;
    uint8_t jet[256][3];
    JetColorMap(jet);
    Mat img(depth.rows, depth.cols, CV_8UC3);
    for (int j = 0; j < depth.rows; ++j)
        for (int i = 0; i < depth.cols; ++i) {
            float v = (depth.at<float>(j, i) - lo) / delta;
            if (v > 1) v = 1;
            if (v < 0) v = 0;
            if (std::isnan(v)) v = 0;  // 0/0 when every depth is equal: static_cast<uchar>(NaN) is undefined in the reference
            const uint8_t g = (uint8_t)(v * 255);
            uint8_t *o = img.ptr<uint8_t>(j) + 3 * i;
            o[0] = jet[g][0]; o[1] = jet[g][1]; o[2] = jet[g][2];
        }
    return WriteJPEG(depth_path, img, 95);
}

bool ShowConfidenceMap(const path &confidence_path, const Mat &confidence) {
    if (confidence.empty() || confidence.type() != CV_8UC1) return false;
    uint8_t max_c = 0, min_c = 255;
    for (uint8_t v : confidence.buf) { if (v > max_c) max_c = v; if (v < min_c) min_c = v; }
    uint8_t delta = (uint8_t)(max_c - min_c);
    if (delta == 0) delta = 1;
    Mat img(confidence.rows, confidence.cols, CV_8UC1);
    for (size_t i = 0; i < confidence.buf.size(); ++i) img.buf[i] = (uint8_t)((confidence.buf[i] - min_c) * 255 / delta);
    return WritePNG(confidence_path, img);
}

bool ShowNormalMap(const path &normal_path, const Mat &normal) {
    if (normal.empty() || normal.type() != CV_32FC3) return false;
    Mat img(normal.rows, normal.cols, CV_8UC3);
    for (int j = 0; j < normal.rows; ++j) {
        const float *n = normal.ptr<float>(j);
        uint8_t *o = img.ptr<uint8_t>(j);
        for (int i = 0; i < normal.cols; ++i, n += 3, o += 3) {
            const float len = (float)std::sqrt(std::pow((double)n[0], 2) + std::pow((double)n[1], 2) + std::pow((double)n[2], 2));
            for (int c = 0; c < 3; ++c) {
                const float u = len == 0 ? 0.0f : n[c] / len;
                // Mat::convertTo(CV_8U, 127.5, 127.5): saturate_cast of the rounded (ties-to-even) value
                const long r = std::lrint(u * (255.f / 2.f) + (255.f / 2.f));
                o[c] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
            }
        }
    }
    return WriteJPEG(normal_path, img, 95);
}

bool ShowWeakImage(const path &weak_path, const Mat &weak) {
    if (weak.empty() || weak.type() != CV_8UC1) return false;
    Mat img(weak.rows, weak.cols, CV_8UC3);  // uninitialised cv::Mat in the reference for any other state value; zero here
    for (size_t i = 0; i < weak.buf.size(); ++i) {
        uint8_t *o = img.buf.data() + 3 * i;
        switch (weak.buf[i]) {
            case APDE_WEAK: o[0] = 255; o[1] = 255; o[2] = 255; break;
            case APDE_STRONG: o[0] = 0; o[1] = 255; o[2] = 0; break;
            case APDE_UNKNOWN: o[0] = 0; o[1] = 0; o[2] = 255; break;
            default: break;
        }
    }
    return WritePNG(weak_path, img);
}

// ------------------------------------------------------------------------------------------------ asynchronous writer
struct ShowWriter::Impl {
    std::deque<std::thread> running;  // oldest first
    size_t max_threads = 1;
};

ShowWriter::ShowWriter() : impl(new Impl) {
    const size_t hw = std::max<size_t>(1, std::thread::hardware_concurrency());
    impl->max_threads = std::min<size_t>(hw, 16);
}
ShowWriter::~ShowWriter() {
    wait();
    delete impl;
}
void ShowWriter::wait() {
    for (auto &t : impl->running) t.join();
    impl->running.clear();
}
void ShowWriter::submit(const path &result_folder, int iteration, Mat depth, Mat normal, Mat weak, Mat confidence, float depth_min,
                        float depth_max) {
    const size_t bytes = depth.buf.size() + normal.buf.size() + weak.buf.size() + confidence.buf.size();
    const size_t by_memory = std::max<size_t>(1, ((size_t)2 << 30) / std::max<size_t>(1, bytes));
    const size_t limit = std::min(impl->max_threads, by_memory);
    while (impl->running.size() >= limit) {  // the oldest view first: in-flight maps stay bounded
        impl->running.front().join();
        impl->running.pop_front();
    }
    auto maps = std::make_shared<std::array<Mat, 4>>();
    (*maps)[0] = std::move(depth); (*maps)[1] = std::move(normal); (*maps)[2] = std::move(weak); (*maps)[3] = std::move(confidence);
    impl->running.emplace_back([maps, result_folder, iteration, depth_min, depth_max]() {
        try {  // debug imagery must never take the run down (the reference ignores the return values, main.cpp:197-203)
            const std::string it = std::to_string(iteration);
            ShowDepthMap(result_folder / ("depth_" + it + ".jpg"), (*maps)[0], depth_min, depth_max);
            ShowNormalMap(result_folder / ("normal_" + it + ".jpg"), (*maps)[1]);
            ShowWeakImage(result_folder / ("weak_" + it + ".png"), (*maps)[2]);
            ShowConfidenceMap(result_folder / ("confidence_" + it + ".png"), (*maps)[3]);
        } catch (...) {
        }
    });
}

}  // namespace apd
