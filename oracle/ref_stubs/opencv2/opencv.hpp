// Minimal stand-in for the OpenCV C++ API surface that /root/reference/APD.h and APD.cu touch (OpenCV C++ headers are
// not installable in this image).  TEST INFRASTRUCTURE: lets nvcc compile the reference's own APD.cu unmodified.
#ifndef APDE_STUB_OPENCV_HPP_
#define APDE_STUB_OPENCV_HPP_
#include <cfloat>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

typedef unsigned char uchar;
#ifndef MIN
#define MIN(a, b) ((a) > (b) ? (b) : (a))
#endif
#ifndef MAX
#define MAX(a, b) ((a) < (b) ? (b) : (a))
#endif
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32SC1 4
#define CV_32FC1 5
#define CV_32FC3 21

namespace cv {
template <typename T, int N>
struct Vec {
    T v[N];
    Vec() { for (int i = 0; i < N; ++i) v[i] = T(); }
    Vec(T a, T b, T c) { v[0] = a; v[1] = b; v[2] = c; }
    T &operator[](int i) { return v[i]; }
    const T &operator[](int i) const { return v[i]; }
};
typedef Vec<float, 3> Vec3f;
typedef Vec<uchar, 3> Vec3b;
struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} };
struct Scalar { double v[4]; Scalar(double a = 0) { v[0] = a; v[1] = v[2] = v[3] = 0; } };

class Mat {
public:
    int rows = 0, cols = 0;
    uchar *data = nullptr;
    size_t step = 0;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, const Scalar &s) { create(r, c, type); std::memset(data, (int)s.v[0], step * rows); }
    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); std::memset(m.data, 0, m.step * m.rows); return m; }
    static Mat zeros(Size s, int type) { return zeros(s.height, s.width, type); }
    static Mat ones(int r, int c, int type) { Mat m(r, c, type); std::memset(m.data, 1, m.step * m.rows); return m; }
    int type() const { return type_; }
    bool empty() const { return data == nullptr; }
    Size size() const { return Size(cols, rows); }
    Mat clone() const { Mat m(rows, cols, type_); if (data) std::memcpy(m.data, data, step * rows); return m; }
    template <typename T> T &at(int r, int c) { return *reinterpret_cast<T *>(data + r * step + c * sizeof(T)); }
    template <typename T> const T &at(int r, int c) const { return *reinterpret_cast<const T *>(data + r * step + c * sizeof(T)); }
    template <typename T> T *ptr(int r = 0) { return reinterpret_cast<T *>(data + r * step); }
    template <typename T> const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(data + r * step); }
private:
    void create(int r, int c, int type) {
        static const int elem[] = {1, 1, 2, 2, 4, 4, 8, 2};
        rows = r; cols = c; type_ = type;
        const int cn = (type >> 3) + 1;
        step = (size_t)c * elem[type & 7] * cn;
        store_.reset(new uchar[step * (size_t)r + 16]);
        data = store_.get();
    }
    int type_ = 0;
    std::shared_ptr<uchar[]> store_;
};
template <typename T> class Mat_ : public Mat {};
inline bool imwrite(const std::string &, const Mat &) { return true; }
}  // namespace cv
#endif
