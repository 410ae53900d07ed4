// test_ply.cpp -- drives saveCloudAndCamerasToPLY from a scene file written by tests/test_host_ply_cpu.py (which also renders the
// expected text independently, oracle/ply_oracle.py).  Usage: test_ply <scene.bin> <out prefix>
// scene.bin (little endian): int32 nviews, per view {int32 h, w, nfeat; float32 pts[nfeat][2]; uint8 bgr[h][w][3]};
//                            int32 npoints, per point {float32 xyz[3]; int32 k; int32 (view, feat)[k]}; int32 ncams, float32 pose[ncams][12]
#include "sfm_ply.h"
#include <cstdio>
#include <cstring>

using namespace sfmtoylib;

template <typename T> static bool rd(FILE* f, T* v, size_t n = 1) { return fread(v, sizeof(T), n, f) == n; }

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t nviews = 0; if (!rd(f, &nviews)) return 2;
    std::vector<Features> feats(nviews); std::vector<cv::Mat> images(nviews);
    for (int v = 0; v < nviews; ++v) {
        int32_t h, w, nf; if (!rd(f, &h) || !rd(f, &w) || !rd(f, &nf)) return 2;
        std::vector<float> pts((size_t)nf * 2); if (nf && !rd(f, pts.data(), pts.size())) return 2;
        for (int i = 0; i < nf; ++i) feats[v].points.push_back(cv::Point2f(pts[2 * i], pts[2 * i + 1]));
        images[v] = cv::Mat(h, w, cv::CV_8UC3);
        if ((size_t)h * w && !rd(f, images[v].data, (size_t)h * w * 3)) return 2;
    }
    int32_t np = 0; if (!rd(f, &np)) return 2;
    PointCloud cloud(np);
    for (int i = 0; i < np; ++i) {
        float xyz[3]; int32_t k; if (!rd(f, xyz, 3) || !rd(f, &k)) return 2;
        cloud[i].p = cv::Point3f(xyz[0], xyz[1], xyz[2]);
        for (int j = 0; j < k; ++j) { int32_t vf[2]; if (!rd(f, vf, 2)) return 2; cloud[i].originatingViews[vf[0]] = vf[1]; }
    }
    int32_t nc = 0; if (!rd(f, &nc)) return 2;
    std::vector<Pose> poses(nc);
    for (int i = 0; i < nc; ++i) if (!rd(f, poses[i].val, 12)) return 2;
    fclose(f);
    return saveCloudAndCamerasToPLY(argv[2], cloud, feats, images, poses) ? 0 : 1;
}
