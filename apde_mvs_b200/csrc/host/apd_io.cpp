#include "apd_io.h"

#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <thread>
#include <algorithm>

namespace apd {

size_t Mat::elem_size() const {
    static const size_t depth_bytes[8] = {1, 1, 2, 2, 4, 4, 8, 2};
    return depth_bytes[type_ & 7] * (size_t)((type_ >> 3) + 1);
}
void Mat::create(int r, int c, int type) {
    rows = r; cols = c; type_ = type;
    buf.assign((size_t)r * c * elem_size(), 0);
}

bool ReadBinMat(const path &mat_path, Mat &mat) {
    std::ifstream in(mat_path, std::ios_base::binary);
    if (!in) { std::cout << "Error opening file: " << mat_path << std::endl; return false; }
    int32_t version, rows, cols, type;
    in.read((char *)&version, 4); in.read((char *)&rows, 4); in.read((char *)&cols, 4); in.read((char *)&type, 4);
    if (!in || version != 1) { std::cout << "Version error: " << mat_path << std::endl; return false; }
    // the header is untrusted input: a type the writer can produce (the four of WriteBinMat's callers, main.cpp:180-204 and
    // APD.cu:2615-2661), and a payload that is really in the file -- before any allocation is sized from it
    if (type != CV_8UC1 && type != CV_32SC1 && type != CV_32FC1 && type != CV_32FC3) { std::cout << "Type error: " << mat_path << std::endl; return false; }
    if (rows < 0 || cols < 0 || rows > (1 << 16) || cols > (1 << 16)) { std::cout << "Size error: " << mat_path << std::endl; return false; }
    const std::streampos body = in.tellg();
    in.seekg(0, std::ios::end);
    const std::streamoff have = in.tellg() - body;
    in.seekg(body);
    Mat probe(0, 0, type);
    if ((std::streamoff)((size_t)rows * cols * probe.elem_size()) > have) { std::cout << "Size error: " << mat_path << " is truncated" << std::endl; return false; }
    mat.create(rows, cols, type);
    in.read((char *)mat.data(), (std::streamsize)mat.buf.size());
    return (bool)in;
}

bool WriteBinMat(const path &mat_path, const Mat &mat) {
    std::ofstream out(mat_path, std::ios_base::binary);
    if (!out) { std::cout << "Error opening file: " << mat_path << std::endl; return false; }
    const int32_t version = 1, rows = mat.rows, cols = mat.cols, type = mat.type();
    out.write((const char *)&version, 4); out.write((const char *)&rows, 4); out.write((const char *)&cols, 4); out.write((const char *)&type, 4);
    out.write((const char *)mat.data(), (std::streamsize)mat.buf.size());
    return (bool)out;
}

bool ReadCamera(const path &cam_path, Camera &cam) {
    std::ifstream in(cam_path);
    if (!in) return false;
    std::string word;
    in >> word;  // "extrinsic"
    for (int i = 0; i < 3; ++i) in >> cam.R[3 * i + 0] >> cam.R[3 * i + 1] >> cam.R[3 * i + 2] >> cam.t[i];
    float tmp[4];
    in >> tmp[0] >> tmp[1] >> tmp[2] >> tmp[3];
    in >> word;  // "intrinsic"
    for (int i = 0; i < 3; ++i) in >> cam.K[3 * i + 0] >> cam.K[3 * i + 1] >> cam.K[3 * i + 2];
    for (int j = 0; j < 3; ++j)  // camera centre in double (APD.cpp:114-119)
        cam.c[j] = -float(double(cam.R[0 + j]) * double(cam.t[0]) + double(cam.R[3 + j]) * double(cam.t[1]) + double(cam.R[6 + j]) * double(cam.t[2]));
    in >> cam.depth_min >> cam.interval;
    if (!(in >> cam.depth_num >> cam.depth_max)) {
        cam.depth_num = 192;
        cam.depth_max = cam.interval * cam.depth_num + cam.depth_min;
    }
    cam.width = cam.height = 0;
    return true;
}

// ---------------------------------------------------------------------------------------------- image decoding
static bool read_file(const path &p, std::vector<uint8_t> &out) {
    std::ifstream in(p, std::ios::binary);
    if (!in) return false;
    in.seekg(0, std::ios::end);
    out.resize((size_t)in.tellg());
    in.seekg(0);
    in.read((char *)out.data(), (std::streamsize)out.size());
    return (bool)in;
}
static uint32_t be32(const uint8_t *p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

// non-interlaced 8-bit PNG: grey, grey+alpha, RGB, RGBA, palette -> interleaved RGB (or single channel)
static bool decode_png(const std::vector<uint8_t> &f, int &w, int &h, int &channels, std::vector<uint8_t> &pix) {
    static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    if (f.size() < 33 || memcmp(f.data(), sig, 8)) return false;
    size_t pos = 8;
    std::vector<uint8_t> idat, palette;
    int bit_depth = 0, color_type = 0, interlace = 0;
    while (pos + 12 <= f.size()) {
        const uint32_t len = be32(&f[pos]);
        const char *tag = (const char *)&f[pos + 4];
        const uint8_t *d = &f[pos + 8];
        if (pos + 12 + len > f.size()) return false;
        if (!memcmp(tag, "IHDR", 4)) { w = (int)be32(d); h = (int)be32(d + 4); bit_depth = d[8]; color_type = d[9]; interlace = d[12]; }
        else if (!memcmp(tag, "PLTE", 4)) palette.assign(d, d + len);
        else if (!memcmp(tag, "IDAT", 4)) idat.insert(idat.end(), d, d + len);
        else if (!memcmp(tag, "IEND", 4)) break;
        pos += 12 + len;
    }
    if (bit_depth != 8 || interlace != 0) return false;
    const int spp = color_type == 0 ? 1 : color_type == 2 ? 3 : color_type == 3 ? 1 : color_type == 4 ? 2 : color_type == 6 ? 4 : 0;
    if (!spp) return false;
    const size_t stride = (size_t)w * spp;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf rawlen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size()) != Z_OK || rawlen != raw.size()) return false;
    std::vector<uint8_t> img(stride * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t ft = raw[(stride + 1) * y];
        const uint8_t *s = &raw[(stride + 1) * y + 1];
        uint8_t *o = &img[stride * y];
        const uint8_t *up = y ? &img[stride * (y - 1)] : nullptr;
        for (size_t x = 0; x < stride; ++x) {
            const int a = x >= (size_t)spp ? o[x - spp] : 0, b = up ? up[x] : 0, c = (up && x >= (size_t)spp) ? up[x - spp] : 0;
            int v = s[x];
            switch (ft) {
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: { const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c); v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
                default: break;
            }
            o[x] = (uint8_t)v;
        }
    }
    channels = (color_type == 0 || color_type == 4) ? 1 : 3;
    pix.resize((size_t)w * h * channels);
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        if (color_type == 0) pix[i] = img[i];
        else if (color_type == 4) pix[i] = img[2 * i];
        else if (color_type == 2) { pix[3 * i] = img[3 * i]; pix[3 * i + 1] = img[3 * i + 1]; pix[3 * i + 2] = img[3 * i + 2]; }
        else if (color_type == 6) { pix[3 * i] = img[4 * i]; pix[3 * i + 1] = img[4 * i + 1]; pix[3 * i + 2] = img[4 * i + 2]; }
        else { const int k = img[i]; if ((size_t)3 * k + 2 >= palette.size()) return false; pix[3 * i] = palette[3 * k]; pix[3 * i + 1] = palette[3 * k + 1]; pix[3 * i + 2] = palette[3 * k + 2]; }
    }
    return true;
}

static bool decode_pnm(const std::vector<uint8_t> &f, int &w, int &h, int &channels, std::vector<uint8_t> &pix) {
    if (f.size() < 8 || f[0] != 'P' || (f[1] != '5' && f[1] != '6')) return false;
    channels = f[1] == '5' ? 1 : 3;
    size_t pos = 2;
    int vals[3], k = 0;
    while (k < 3 && pos < f.size()) {
        while (pos < f.size() && isspace(f[pos])) ++pos;
        if (pos < f.size() && f[pos] == '#') { while (pos < f.size() && f[pos] != '\n') ++pos; continue; }
        int v = 0;
        while (pos < f.size() && isdigit(f[pos])) v = v * 10 + (f[pos++] - '0');
        vals[k++] = v;
    }
    ++pos;
    w = vals[0]; h = vals[1];
    if (vals[2] != 255 || pos + (size_t)w * h * channels > f.size()) return false;
    pix.assign(f.begin() + pos, f.begin() + pos + (size_t)w * h * channels);
    return true;
}

bool decode_jpeg(const std::vector<uint8_t> &f, bool want_color, int &w, int &h, int &channels, std::vector<uint8_t> &pix);  // apd_jpeg.cpp

static bool decode_any(const path &p, bool want_color, int &w, int &h, int &ch, std::vector<uint8_t> &pix) {
    std::vector<uint8_t> f;
    if (!read_file(p, f)) { std::cout << "Error opening file: " << p << std::endl; return false; }
    if (decode_png(f, w, h, ch, pix) || decode_pnm(f, w, h, ch, pix) || decode_jpeg(f, want_color, w, h, ch, pix)) return true;
    std::cout << "Error: unsupported image format (8-bit PNG / PGM / PPM / Huffman JPEG) " << p << std::endl;
    return false;
}

bool ReadImage(const path &img_path, Mat &gray) {
    int w, h, ch;
    std::vector<uint8_t> pix;
    if (!decode_any(img_path, false, w, h, ch, pix)) return false;
    gray.create(h, w, CV_8UC1);
    if (ch == 1) memcpy(gray.data(), pix.data(), pix.size());
    else  // cv::imread(IMREAD_GRAYSCALE) of a colour PNG goes through libpng's png_set_rgb_to_gray(0.299, 0.587): 15-bit
          // coefficients 9797 / 19234 / 3737, truncated (pinned against cv2 4.13 on 65 536 random pixels: 0 mismatches)
        for (size_t i = 0; i < (size_t)w * h; ++i)
            gray.data()[i] = (uint8_t)((pix[3 * i] * 9797 + pix[3 * i + 1] * 19234 + pix[3 * i + 2] * 3737) >> 15);
    return true;
}

bool ReadImageColor(const path &img_path, Mat &bgr) {
    int w, h, ch;
    std::vector<uint8_t> pix;
    if (!decode_any(img_path, true, w, h, ch, pix)) return false;
    bgr.create(h, w, CV_8UC3);
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        if (ch == 1) { bgr.data()[3 * i] = bgr.data()[3 * i + 1] = bgr.data()[3 * i + 2] = pix[i]; }
        else { bgr.data()[3 * i] = pix[3 * i + 2]; bgr.data()[3 * i + 1] = pix[3 * i + 1]; bgr.data()[3 * i + 2] = pix[3 * i]; }
    }
    return true;
}

bool WritePGM(const path &p, const Mat &g) {
    std::ofstream out(p, std::ios::binary);
    if (!out) return false;
    out << "P5\n" << g.cols << " " << g.rows << "\n255\n";
    out.write((const char *)g.data(), (std::streamsize)g.buf.size());
    return (bool)out;
}

// 8-bit grey or BGR image as a non-interlaced PNG (filter 0, zlib), the format of the reference's skip.png (APD.cpp:1035)
bool WritePNG(const path &p, const Mat &img) {
    const int ch = img.type() == CV_8UC3 ? 3 : img.type() == CV_8UC1 ? 1 : 0;
    if (!ch || img.empty()) return false;
    const size_t stride = (size_t)img.cols * ch;
    std::vector<uint8_t> raw((stride + 1) * img.rows);
    for (int y = 0; y < img.rows; ++y) {
        uint8_t *o = &raw[(stride + 1) * y];
        *o++ = 0;
        const uint8_t *s = img.data() + stride * y;
        if (ch == 1) memcpy(o, s, stride);
        else for (int x = 0; x < img.cols; ++x) { o[3 * x] = s[3 * x + 2]; o[3 * x + 1] = s[3 * x + 1]; o[3 * x + 2] = s[3 * x]; }
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
    std::ofstream out(p, std::ios::binary);
    if (!out) return false;
    auto chunk = [&](const char *tag, const uint8_t *d, uint32_t len) {
        uint8_t hdr[8] = {(uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len, (uint8_t)tag[0], (uint8_t)tag[1], (uint8_t)tag[2], (uint8_t)tag[3]};
        out.write((const char *)hdr, 8);
        if (len) out.write((const char *)d, len);
        uLong crc = crc32(0L, hdr + 4, 4);
        if (len) crc = crc32(crc, d, len);
        const uint8_t c[4] = {(uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc};
        out.write((const char *)c, 4);
    };
    static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    out.write((const char *)sig, 8);
    const uint32_t w = (uint32_t)img.cols, h = (uint32_t)img.rows;
    const uint8_t ihdr[13] = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w, (uint8_t)(h >> 24), (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h,
                              8, (uint8_t)(ch == 3 ? 2 : 0), 0, 0, 0};
    chunk("IHDR", ihdr, 13);
    chunk("IDAT", comp.data(), (uint32_t)clen);
    chunk("IEND", nullptr, 0);
    return (bool)out;
}

bool ExportPointCloud(const path &ply_path, const std::vector<PointList> &pc, bool export_color) {
    std::ofstream out(ply_path, std::ios::binary);
    if (!out) return false;
    out << "ply\nformat binary_little_endian 1.0\nelement vertex " << int(pc.size()) << "\nproperty float x\nproperty float y\nproperty float z\n";
    if (export_color) out << "property uchar blue\nproperty uchar green\nproperty uchar red\n";
    out << "end_header\n";
    for (const auto &p : pc) {
        out.write((const char *)&p.coord.x, 4); out.write((const char *)&p.coord.y, 4); out.write((const char *)&p.coord.z, 4);
        if (export_color) {
            const uint8_t c[3] = {(uint8_t)p.color.x, (uint8_t)p.color.y, (uint8_t)p.color.z};
            out.write((const char *)c, 3);
        }
    }
    return (bool)out;
}

std::string ToFormatIndex(int index) {
    std::stringstream ss;
    ss << std::setw(8) << std::setfill('0') << index;
    return ss.str();
}

bool GenerateSampleList(const path &dense_folder, std::vector<ProblemDesc> &problems, std::string *err) {
    const path list = dense_folder / "pair.txt", image_folder = dense_folder / "images";
    problems.clear();
    std::ifstream file(list);
    if (!file) { if (err) *err = "can not open " + list.string(); return false; }
    static const char *exts[] = {".jpg", ".png", ".jpeg", ".JPG", ".PNG", ".JPEG", ".pgm", ".ppm"};
    std::string line;
    std::getline(file, line);
    int num_images = 0;
    { std::stringstream ss(line); ss >> num_images; }
    for (int i = 0; i < num_images; ++i) {
        ProblemDesc pr;
        std::getline(file, line);
        { std::stringstream ss(line); ss >> pr.ref_image_id; }
        pr.dense_folder = dense_folder;
        pr.result_folder = dense_folder / "APD" / ToFormatIndex(pr.ref_image_id);
        std::filesystem::create_directories(pr.result_folder);
        std::getline(file, line);
        std::stringstream ss(line);
        int n = 0;
        ss >> n;
        for (int j = 0; j < n; ++j) {
            int id; float score;
            ss >> id >> score;
            if (score <= 0.0f) continue;
            pr.src_image_ids.push_back(id);
        }
        for (const char *e : exts)
            if (std::filesystem::exists(image_folder / (ToFormatIndex(pr.ref_image_id) + e))) { pr.img_ext = e; break; }
        if (pr.img_ext.empty()) { if (err) *err = "Error: can not find image: " + ToFormatIndex(pr.ref_image_id); return false; }
        problems.push_back(pr);
    }
    return true;
}

size_t LoadViewsBatchSize(size_t num_views) {
    const size_t hw = std::max<size_t>(1, std::thread::hardware_concurrency());
    return std::max<size_t>(1, std::min<size_t>({hw, num_views, (size_t)16}));
}

void LoadViews(const path &dense_folder, const std::vector<ProblemDesc> &problems, size_t first, size_t count, std::vector<LoadedView> &out) {
    out.clear();
    out.resize(count);
    auto load_one = [&](size_t k) {
        const ProblemDesc &p = problems[first + k];
        LoadedView &l = out[k];
        const path img = dense_folder / "images" / (ToFormatIndex(p.ref_image_id) + p.img_ext);
        l.image_ok = ReadImage(img, l.gray);
        if (l.image_ok) ReadImageColor(img, l.bgr);
        l.cam_ok = ReadCamera(dense_folder / "cams" / (ToFormatIndex(p.ref_image_id) + "_cam.txt"), l.cam);
        const path sa_path = dense_folder / "sa_masks" / (ToFormatIndex(p.ref_image_id) + ".bin");
        if (std::filesystem::exists(sa_path) && !(ReadBinMat(sa_path, l.sa) && l.sa.type() == CV_8UC1 && !l.sa.empty())) l.sa = Mat();
    };
    if (count == 1) { load_one(0); return; }
    std::vector<std::thread> pool;
    for (size_t k = 0; k < count; ++k) pool.emplace_back(load_one, k);
    for (auto &t : pool) t.join();
}

}  // namespace apd
