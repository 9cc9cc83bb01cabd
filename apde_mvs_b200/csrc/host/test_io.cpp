// test_io.cpp -- CPU-only exerciser of the host I/O layer for tests/test_host_io.py (no GPU calls; links libapde only for symbols)
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "apd_io.h"
using namespace apd;

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    const std::string cmd = argv[1];
    if (cmd == "gray") {  // gray <image> <out.pgm>
        Mat g;
        if (!ReadImage(argv[2], g)) return 1;
        return WritePGM(argv[3], g) ? 0 : 1;
    }
    if (cmd == "color") {  // color <image> <out.bin>  (CV_8UC3 BGR)
        Mat c;
        if (!ReadImageColor(argv[2], c)) return 1;
        return WriteBinMat(argv[3], c) ? 0 : 1;
    }
    if (cmd == "png") {  // png <image> <out.png>: colour round trip through WritePNG
        Mat c;
        if (!ReadImageColor(argv[2], c)) return 1;
        return WritePNG(argv[3], c) ? 0 : 1;
    }
    if (cmd == "graypng") {  // graypng <image> <out.png>
        Mat g;
        if (!ReadImage(argv[2], g)) return 1;
        return WritePNG(argv[3], g) ? 0 : 1;
    }
    if (cmd == "jpg") {  // jpg <image> <out.jpg> [quality]: colour round trip through WriteJPEG
        Mat c;
        if (!ReadImageColor(argv[2], c)) return 1;
        return WriteJPEG(argv[3], c, argc > 4 ? atoi(argv[4]) : 95) ? 0 : 1;
    }
    if (cmd == "grayjpg") {  // grayjpg <image> <out.jpg>
        Mat g;
        if (!ReadImage(argv[2], g)) return 1;
        return WriteJPEG(argv[3], g) ? 0 : 1;
    }
    if (cmd == "show") {  // show <dir with depths/normals/weak/confidence.bin> <depth_min> <depth_max>: the four Show* pictures
        const path dir = argv[2];
        Mat depth, normal, weak, conf;
        if (!ReadBinMat(dir / "depths.bin", depth) || !ReadBinMat(dir / "normals.bin", normal) || !ReadBinMat(dir / "weak.bin", weak) ||
            !ReadBinMat(dir / "confidence.bin", conf)) return 1;
        const bool ok = ShowDepthMap(dir / "depth_3.jpg", depth, (float)atof(argv[3]), (float)atof(argv[4])) && ShowNormalMap(dir / "normal_3.jpg", normal) &&
                        ShowWeakImage(dir / "weak_3.png", weak) && ShowConfidenceMap(dir / "confidence_3.png", conf);
        uint8_t jet[256][3];
        JetColorMap(jet);
        Mat j(1, 256, CV_8UC3);
        memcpy(j.data(), jet, sizeof(jet));
        return ok && WriteBinMat(dir / "jet.bin", j) ? 0 : 1;
    }
    if (cmd == "showasync") {  // showasync <dir with the four .bin maps> <n> <dmin> <dmax>: n views through ShowWriter into <dir>/<k>/
        const path dir = argv[2];
        Mat depth, normal, weak, conf;
        if (!ReadBinMat(dir / "depths.bin", depth) || !ReadBinMat(dir / "normals.bin", normal) || !ReadBinMat(dir / "weak.bin", weak) ||
            !ReadBinMat(dir / "confidence.bin", conf)) return 1;
        const int n = atoi(argv[3]);
        {
            ShowWriter writer;
            for (int k = 0; k < n; ++k) {
                std::filesystem::create_directories(dir / std::to_string(k));
                Mat d = depth;
                for (size_t i = 0; i < d.buf.size() / 4; ++i) d.ptr<float>()[i] += 0.01f * k;  // another picture per view
                writer.submit(dir / std::to_string(k), 3, d, normal, weak, conf, (float)atof(argv[4]), (float)atof(argv[5]));
            }
        }  // the destructor joins
        for (int k = 0; k < n; ++k) {  // the same pictures, one by one
            Mat d = depth;
            for (size_t i = 0; i < d.buf.size() / 4; ++i) d.ptr<float>()[i] += 0.01f * k;
            if (!ShowDepthMap(dir / std::to_string(k) / "depth_sync.jpg", d, (float)atof(argv[4]), (float)atof(argv[5]))) return 1;
        }
        return 0;
    }
    if (cmd == "cam") {  // cam <cam.txt>
        Camera cam;
        if (!ReadCamera(argv[2], cam)) return 1;
        for (int i = 0; i < 9; ++i) printf("%.9g ", cam.K[i]);
        for (int i = 0; i < 9; ++i) printf("%.9g ", cam.R[i]);
        for (int i = 0; i < 3; ++i) printf("%.9g ", cam.t[i]);
        for (int i = 0; i < 3; ++i) printf("%.9g ", cam.c[i]);
        printf("%.9g %.9g %.9g %.9g\n", cam.depth_min, cam.depth_max, cam.interval, cam.depth_num);
        return 0;
    }
    if (cmd == "bin") {  // bin <in.bin> <out.bin>: round trip
        Mat m;
        if (!ReadBinMat(argv[2], m)) return 1;
        printf("%d %d %d\n", m.rows, m.cols, m.type());
        return WriteBinMat(argv[3], m) ? 0 : 1;
    }
    if (cmd == "pairs") {  // pairs <dense_folder>
        std::vector<ProblemDesc> pr;
        std::string err;
        if (!GenerateSampleList(argv[2], pr, &err)) { std::cout << err << std::endl; return 1; }
        for (auto &p : pr) { printf("%d %s:", p.ref_image_id, p.img_ext.c_str()); for (int s : p.src_image_ids) printf(" %d", s); printf("\n"); }
        return 0;
    }
    if (cmd == "load") {  // load <dense_folder>: every view through LoadViews (threaded batches), one line of checksums per view
        std::vector<ProblemDesc> pr;
        std::string err;
        if (!GenerateSampleList(argv[2], pr, &err)) { std::cout << err << std::endl; return 1; }
        const size_t bs = LoadViewsBatchSize(pr.size());
        for (size_t first = 0; first < pr.size(); first += bs) {
            std::vector<LoadedView> batch;
            LoadViews(argv[2], pr, first, std::min(bs, pr.size() - first), batch);
            for (size_t k = 0; k < batch.size(); ++k) {
                const LoadedView &l = batch[k];
                auto sum = [](const Mat &m) { unsigned long long a = 0; for (size_t i = 0; i < m.buf.size(); ++i) a = a * 1315423911ull + m.buf[i]; return a; };
                printf("%d %d %d %dx%d %llu %llu %llu %.9g %.9g\n", pr[first + k].ref_image_id, (int)l.image_ok, (int)l.cam_ok, l.gray.cols, l.gray.rows,
                       sum(l.gray), sum(l.bgr), sum(l.sa), l.cam_ok ? l.cam.K[0] : 0.0f, l.cam_ok ? l.cam.depth_min : 0.0f);
            }
        }
        return 0;
    }
    if (cmd == "ply") {  // ply <out.ply>
        std::vector<PointList> pc = {{{1.5f, -2.0f, 3.25f}, {10, 20, 30}}, {{0.0f, 1.0f, 2.0f}, {255, 0, 128}}};
        return ExportPointCloud(argv[2], pc, true) ? 0 : 1;
    }
    return 2;
}
