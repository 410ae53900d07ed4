// test_shim.cpp -- C++ host-side test of the drop-in shim (runs on the GPU box).  Reads like the reference's own
// SfMUnitTests.cpp: the triangulate_from_2_views scene (12 canned points, two mock cameras, tolerance 0.01), plus a
// matcher check against an in-test brute force and an adjustBundle run on a small synthetic scene.
#include "sfmtoylib_b200.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include "../../include/sfmb200.h"

using namespace sfmtoylib;

static int failures = 0;
#define EXPECT(cond, msg) do { if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, msg); ++failures; } } while (0)

static void eulerDegToR(double pitch, double roll, double yaw, float R[9]) {      // R = Rz(yaw) Ry(roll) Rx(pitch)
    const double d = M_PI / 180.0, c1 = std::cos(yaw * d), s1 = std::sin(yaw * d), c2 = std::cos(roll * d), s2 = std::sin(roll * d),
                 c3 = std::cos(pitch * d), s3 = std::sin(pitch * d);
    const double r[9] = {c1 * c2, -s1 * c3 + c1 * s2 * s3, s1 * s3 + c1 * s2 * c3, s1 * c2, c1 * c3 + s1 * s2 * s3, -c1 * s3 + s1 * s2 * c3, -s2, c2 * s3, c2 * c3};
    for (int i = 0; i < 9; ++i) R[i] = (float)r[i];
}
static cv::Matx34f makePose(const float R[9], float tx, float ty, float tz) {
    cv::Matx34f P;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) P(r, c) = R[3 * r + c];
    P(0, 3) = tx; P(1, 3) = ty; P(2, 3) = tz;
    return P;
}
static cv::Point2f project(const cv::Matx34f& P, double f, double cx, double cy, double X, double Y, double Z) {
    const double x = P(0, 0) * X + P(0, 1) * Y + P(0, 2) * Z + P(0, 3), y = P(1, 0) * X + P(1, 1) * Y + P(1, 2) * Z + P(1, 3),
                 z = P(2, 0) * X + P(2, 1) * Y + P(2, 2) * Z + P(2, 3);
    return cv::Point2f((float)(f * x / z + cx), (float)(f * y / z + cy));
}
static cv::Mat makeK(float f, float cx, float cy) {
    cv::Mat K(3, 3, cv::CV_32F);
    const float k[9] = {f, 0, cx, 0, f, cy, 0, 0, 1};
    for (int i = 0; i < 9; ++i) K.ptr<float>(0)[i] = k[i];
    return K;
}

static void test_triangulate_from_2_views() {
    const float canned[12][3] = {{4, 12, 50}, {12, 11, 55}, {22, 1, 45}, {13, 3, 60}, {11, 16, 61}, {21, 12, 65}, {24, 11, 67},
                                 {29, 6, 41}, {27, 4, 44}, {22, 7, 58}, {20, 9, 51}, {15, 10, 40}};
    float Rl[9], Rr[9];
    eulerDegToR(5, 5, 5, Rl); eulerDegToR(-5, 0, 5, Rr);
    const cv::Matx34f Pl = makePose(Rl, -10, 0, 30), Pr = makePose(Rr, 10, 0, 28);
    Features left, right; Matching matching;
    for (int i = 0; i < 12; ++i) {
        left.points.push_back(project(Pl, 700, 320, 240, canned[i][0], canned[i][1], canned[i][2]));
        right.points.push_back(project(Pr, 700, 320, 240, canned[i][0], canned[i][1], canned[i][2]));
        matching.push_back(cv::DMatch(i, i, 0));
    }
    Intrinsics intr; intr.K = makeK(700, 320, 240);
    PointCloud cloud;
    const bool ok = SfMStereoUtilities::triangulateViews(intr, ImagePair{0, 1}, matching, left, right, Pl, Pr, cloud);
    EXPECT(ok && cloud.size() == 12, "all 12 points triangulated");
    for (size_t i = 0; i < cloud.size(); ++i) {
        const double dx = cloud[i].p.x - canned[i][0], dy = cloud[i].p.y - canned[i][1], dz = cloud[i].p.z - canned[i][2];
        EXPECT(std::sqrt(dx * dx + dy * dy + dz * dz) < 0.01, "triangulated point within 0.01");
        EXPECT(cloud[i].originatingViews.at(0) == (int)i && cloud[i].originatingViews.at(1) == (int)i, "back references");
    }
    // appends, never clears (SfMStereoUtilities.cpp:202); a gross outlier is filtered
    right.points[3].y += 80;   // off the epipolar line (an x shift would only change the depth)
    SfMStereoUtilities::triangulateViews(intr, ImagePair{0, 1}, matching, left, right, Pl, Pr, cloud);
    EXPECT(cloud.size() == 23, "appended 11 more (one outlier dropped)");
}

static void test_match_features() {
    std::mt19937 rng(7);
    const int nq = 700, nt = 650;
    Features L, R;
    L.descriptors = cv::Mat(nq, 32, cv::CV_8U); R.descriptors = cv::Mat(nt, 32, cv::CV_8U);
    for (int i = 0; i < nt * 32; ++i) R.descriptors.data[i] = (uint8_t)rng();
    for (int i = 0; i < nq * 32; ++i) L.descriptors.data[i] = (uint8_t)rng();
    for (int i = 0; i < nq; i += 4) {            // near-duplicates so that the ratio test passes for a quarter of the rows
        const int src = rng() % nt;
        for (int b = 0; b < 32; ++b) L.descriptors.at<uint8_t>(i, b) = R.descriptors.at<uint8_t>(src, b);
        for (int f = 0; f < (int)(rng() % 12); ++f) L.descriptors.at<uint8_t>(i, rng() % 32) ^= (uint8_t)(1u << (rng() % 8));
    }
    const Matching m = SfM2DFeatureUtilities::matchFeatures(L, R);
    // in-test brute force
    Matching ref;
    for (int i = 0; i < nq; ++i) {
        int d0 = 1 << 30, d1 = 1 << 30, i0 = -1;
        for (int j = 0; j < nt; ++j) {
            int d = 0;
            for (int b = 0; b < 32; ++b) d += __builtin_popcount(L.descriptors.at<uint8_t>(i, b) ^ R.descriptors.at<uint8_t>(j, b));
            if (d < d0) { d1 = d0; d0 = d; i0 = j; } else if (d < d1) d1 = d;
        }
        if ((double)(float)d0 < (double)0.8f * (double)(float)d1) ref.push_back(cv::DMatch(i, i0, 0, (float)d0));
    }
    EXPECT(m.size() == ref.size() && m.size() > 100, "same number of ratio-test survivors");
    for (size_t i = 0; i < m.size() && i < ref.size(); ++i)
        EXPECT(m[i].queryIdx == ref[i].queryIdx && m[i].trainIdx == ref[i].trainIdx && m[i].distance == ref[i].distance && m[i].imgIdx == 0,
               "match identical to brute force");
}

static void test_adjust_bundle() {
    std::mt19937 rng(11);
    std::normal_distribution<double> N01(0, 1);
    std::uniform_real_distribution<double> U(-2, 2);
    const int nviews = 6, npts = 400;
    const double f_true = 2500, cx = 512, cy = 384;
    std::vector<Pose> truth(nviews), poses(nviews);
    for (int v = 0; v < nviews; ++v) {
        float R[9]; eulerDegToR(3.0 * v - 7, 10.0 * v - 25, 2.0 * v, R);
        truth[v] = makePose(R, (float)(0.8 * v - 2), (float)(0.1 * v), 9.0f + 0.2f * v);
        float Rn[9]; eulerDegToR(3.0 * v - 7 + 0.4 * N01(rng), 10.0 * v - 25 + 0.4 * N01(rng), 2.0 * v + 0.4 * N01(rng), Rn);
        poses[v] = makePose(Rn, truth[v](0, 3) + 0.03f * (float)N01(rng), truth[v](1, 3) + 0.03f * (float)N01(rng), truth[v](2, 3) + 0.03f * (float)N01(rng));
    }
    poses.push_back(Pose());                       // an "empty" placeholder pose nobody observes (SfMBundleAdjustmentUtils.cpp:118-122)
    {   // a NON-empty pose nobody observes: the reference still rewrites it from its float angle-axis (:192-215)
        float Ru[9]; eulerDegToR(11.0, -17.0, 23.0, Ru);
        poses.push_back(makePose(Ru, 0.25f, -0.5f, 3.0f));
    }
    std::vector<Features> feats(nviews + 2);
    PointCloud cloud;
    for (int i = 0; i < npts; ++i) {
        const double X = U(rng), Y = U(rng), Z = U(rng);
        Point3DInMap p; p.p = cv::Point3f((float)(X + 0.03 * N01(rng)), (float)(Y + 0.03 * N01(rng)), (float)(Z + 0.03 * N01(rng)));
        for (int k = 0; k < 3; ++k) {
            const int v = (i + 2 * k) % nviews;
            cv::Point2f uv = project(truth[v], f_true, cx, cy, X, Y, Z);
            uv.x += (float)(0.3 * N01(rng)); uv.y += (float)(0.3 * N01(rng));
            p.originatingViews[v] = (int)feats[v].points.size();
            feats[v].points.push_back(uv);
        }
        cloud.push_back(p);
    }
    Intrinsics intr; intr.K = makeK((float)(f_true * 1.03), (float)cx, (float)cy);
    const PointCloud cloud0 = cloud; const std::vector<Pose> poses0 = poses;
    SfMBundleAdjustmentUtils::adjustBundle(cloud, poses, intr, feats);
    const float f = intr.K.at<float>(0, 0);
    EXPECT(f == intr.K.at<float>(1, 1) && std::fabs(f - f_true) < 25, "focal refined towards the truth, fx == fy");
    double moved = 0;
    for (int i = 0; i < npts; ++i) moved += std::fabs(cloud[i].p.x - cloud0[i].p.x);
    EXPECT(moved > 1e-3, "points were written back (CONVERGENCE)");
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) EXPECT(poses[nviews](r, c) == 0.0f, "empty pose left untouched");
    {   // unobserved non-empty pose: R -> float angle-axis -> double rotation matrix -> float (a round trip, not a copy)
        const Pose& before = poses0[nviews + 1]; const Pose& after = poses[nviews + 1];
        float R[9], aa[3]; double aad[3], Rd[9];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[3 * r + c] = before(r, c);
        sfmb200_rotmat_to_angle_axis_f32(R, aa);
        for (int k = 0; k < 3; ++k) aad[k] = aa[k];
        sfmb200_angle_axis_to_rotmat(aad, Rd);
        double maxdiff = 0;
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) { EXPECT(after(r, c) == (float)Rd[3 * r + c], "unobserved pose = its own angle-axis round trip"); maxdiff = std::fmax(maxdiff, std::fabs(after(r, c) - before(r, c))); }
            EXPECT(after(r, 3) == before(r, 3), "unobserved pose keeps its translation");
        }
        EXPECT(maxdiff < 1e-5, "round trip stays within float rounding");
    }
    // reprojection RMS with the adjusted parameters ~ noise level
    double ss = 0; int n = 0;
    for (int i = 0; i < npts; ++i)
        for (const auto& kv : cloud[i].originatingViews) {
            const cv::Point2f uv = project(poses[kv.first], f, cx, cy, cloud[i].p.x, cloud[i].p.y, cloud[i].p.z), o = feats[kv.first].points[kv.second];
            ss += (uv.x - o.x) * (uv.x - o.x) + (uv.y - o.y) * (uv.y - o.y); ++n;
        }
    EXPECT(std::sqrt(ss / n) < 0.6, "reprojection RMS at the noise floor");
    std::printf("adjustBundle: focal %.2f (truth %.0f), reprojection RMS %.3f px over %d observations\n", f, f_true, std::sqrt(ss / n), n);
}

// extractFeatures (SfM2DFeatureUtilities.cpp:46-51) on a generated image: container invariants here; bit-for-bit parity with cv2 is
// checked by tests/test_gpu_host_shim.py through the file mode below.
static cv::Mat generated_image(int w, int h, int channels) {
    cv::Mat img(h, w, channels == 3 ? cv::CV_8UC3 : cv::CV_8U);
    uint32_t s = 12345u;
    std::vector<uint8_t> g((size_t)w * h, 120);
    for (int b = 0; b < 500; b++) {
        s = s * 1664525u + 1013904223u; const int x0 = (s >> 8) % w;
        s = s * 1664525u + 1013904223u; const int y0 = (s >> 8) % h;
        s = s * 1664525u + 1013904223u; const int sz = 4 + (s >> 8) % 36;
        s = s * 1664525u + 1013904223u; const uint8_t c = (uint8_t)(s >> 16);
        for (int y = y0; y < y0 + sz && y < h; y++) for (int x = x0; x < x0 + sz && x < w; x++) g[(size_t)y * w + x] = c;
    }
    for (size_t i = 0; i < g.size(); i++) for (int c = 0; c < channels; c++) img.data[i * channels + c] = (uint8_t)(g[i] + (c ? 0 : 0));
    return img;
}

static void test_extract_features() {
    SfM2DFeatureUtilities util;
    const cv::Mat grey = generated_image(640, 480, 1), bgr = generated_image(640, 480, 3);
    const Features a = util.extractFeatures(grey), b = util.extractFeatures(bgr);
    EXPECT(a.keyPoints.size() > 500 && a.keyPoints.size() <= 5064, "ORB finds key points");
    EXPECT(a.points.size() == a.keyPoints.size() && (size_t)a.descriptors.rows == a.keyPoints.size() && a.descriptors.cols == 32, "container sizes");
    bool same = a.keyPoints.size() == b.keyPoints.size(), pts = true, inside = true;
    for (size_t i = 0; i < a.keyPoints.size(); i++) {
        pts = pts && a.points[i].x == a.keyPoints[i].pt.x && a.points[i].y == a.keyPoints[i].pt.y;
        inside = inside && a.keyPoints[i].pt.x >= 31 && a.keyPoints[i].pt.x < 640 - 31 && a.keyPoints[i].octave >= 0 && a.keyPoints[i].octave < 8 &&
                 a.keyPoints[i].angle >= 0 && a.keyPoints[i].angle < 360 && a.keyPoints[i].class_id == -1;
        if (same) same = std::memcmp(&a.keyPoints[i], &b.keyPoints[i], sizeof(cv::KeyPoint)) == 0;
    }
    if (same) same = std::memcmp(a.descriptors.data, b.descriptors.data, 32 * a.keyPoints.size()) == 0;
    EXPECT(pts, "points = KeyPointsToPoints(keyPoints)");
    EXPECT(inside, "key point fields in range");
    EXPECT(same, "B = G = R image gives the grey image's features");
    std::printf("extractFeatures: %zu key points\n", a.keyPoints.size());
}

// file mode: test_shim --orb image.raw width height channels out.bin   (out: int32 n, n x 28-byte key points, n x 32 descriptor bytes)
static int orb_file_mode(char** argv) {
    const int w = std::atoi(argv[3]), h = std::atoi(argv[4]), ch = std::atoi(argv[5]);
    cv::Mat img(h, w, ch == 3 ? cv::CV_8UC3 : cv::CV_8U);
    FILE* f = std::fopen(argv[2], "rb");
    if (!f || std::fread(img.data, 1, (size_t)w * h * ch, f) != (size_t)w * h * ch) { std::fprintf(stderr, "cannot read %s\n", argv[2]); return 2; }
    std::fclose(f);
    SfM2DFeatureUtilities util;
    const Features ft = util.extractFeatures(img);
    FILE* o = std::fopen(argv[6], "wb");
    if (!o) return 2;
    const int32_t n = (int32_t)ft.keyPoints.size();
    std::fwrite(&n, 4, 1, o);
    std::fwrite(ft.keyPoints.data(), sizeof(cv::KeyPoint), n, o);
    std::fwrite(ft.descriptors.data, 32, n, o);
    std::fclose(o);
    return 0;
}

int main(int argc, char** argv) {
    if (argc == 7 && std::string(argv[1]) == "--orb") return orb_file_mode(argv);
    test_triangulate_from_2_views();
    test_extract_features();
    test_match_features();
    test_adjust_bundle();
    std::printf(failures ? "SHIM_TEST FAIL (%d)\n" : "SHIM_TEST PASS\n", failures);
    return failures ? 1 : 0;
}
