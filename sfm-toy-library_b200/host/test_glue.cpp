// test_glue.cpp -- sfm_glue.cpp against the naive restatement (oracle/host_glue_naive.hpp) on randomised scenes: identical
// 2D-3D match lists, identical merged clouds (points, originating-view maps, counters, debug match matrix); then a timing of
// both on a scene of the size where the scans start to dominate.  Usage: test_glue [seed] [--bench]
#include "sfm_glue.h"
#include "../../oracle/host_glue_naive.hpp"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

using namespace sfmtoylib;

struct Scene {
    size_t numImages;
    std::vector<Features> feats;
    MatchMatrix M;
    PointCloud recon, fresh;
    std::set<int> done;
};

static Scene makeScene(unsigned seed, int V, int F, int nRecon, int nFresh, bool uniqueSorted) {
    std::mt19937 rng(seed);
    auto ri = [&](int lo, int hi) { return (int)(rng() % (unsigned)(hi - lo + 1)) + lo; };
    auto rf = [&](float lo, float hi) { return lo + (hi - lo) * (float)(rng() / 4294967296.0); };
    Scene s; s.numImages = (size_t)V;
    s.feats.resize(V);
    for (int v = 0; v < V; ++v) for (int f = 0; f < F; ++f) s.feats[v].points.push_back(cv::Point2f(rf(0, 1024), rf(0, 768)));
    s.M.assign(V, std::vector<Matching>(V));
    for (int i = 0; i < V; ++i)
        for (int j = i + 1; j < V; ++j) {
            Matching& m = s.M[i][j];
            if (uniqueSorted) {                               // what matchFeatures emits: ascending, unique query indices
                for (int q = 0; q < F; ++q) if (rng() % 3 == 0) m.push_back(cv::DMatch(q, ri(0, F - 1), 0, (float)ri(0, 60)));
            } else {                                          // adversarial: repeated query / train indices, arbitrary order
                const int n = F / 2;
                for (int k = 0; k < n; ++k) m.push_back(cv::DMatch(ri(0, F / 4), ri(0, F / 4), 0, (float)ri(0, 40)));
            }
        }
    auto randomPoint = [&](cv::Point3f centre, float spread) {
        Point3DInMap p; p.p = cv::Point3f(centre.x + rf(-spread, spread), centre.y + rf(-spread, spread), centre.z + rf(-spread, spread));
        const int k = ri(2, 4);
        while ((int)p.originatingViews.size() < k) p.originatingViews[ri(0, V - 1)] = ri(0, uniqueSorted ? F - 1 : F / 4);
        return p;
    };
    for (int i = 0; i < nRecon; ++i) s.recon.push_back(randomPoint(cv::Point3f(rf(-1, 1), rf(-1, 1), rf(4, 6)), 0.0f));
    for (int i = 0; i < nFresh; ++i) {
        const int mode = ri(0, 3);
        Point3DInMap p;
        if (mode == 0 || s.recon.empty()) p = randomPoint(cv::Point3f(rf(-1, 1), rf(-1, 1), rf(4, 6)), 0.0f);          // far from everything
        else {
            const Point3DInMap& e = s.recon[ri(0, (int)s.recon.size() - 1)];
            p = randomPoint(e.p, mode == 1 ? 0.004f : 0.012f);                                                           // inside / around the merge radius
            if (mode != 3) {                                                                                             // make a confirming feature match likely
                auto it = e.originatingViews.begin(); std::advance(it, ri(0, (int)e.originatingViews.size() - 1));
                int nv = ri(0, V - 1); if (nv == it->first) nv = (nv + 1) % V;
                const int l = std::min(nv, it->first), r = std::max(nv, it->first);
                if (!s.M[l][r].empty()) {
                    const cv::DMatch& d = s.M[l][r][ri(0, (int)s.M[l][r].size() - 1)];
                    p.originatingViews.clear();
                    p.originatingViews[nv] = (nv == l) ? d.queryIdx : d.trainIdx;
                    const_cast<Point3DInMap&>(e).originatingViews[it->first] = (it->first == l) ? d.queryIdx : d.trainIdx;
                    p.originatingViews[ri(0, V - 1)] = ri(0, uniqueSorted ? F - 1 : F / 4);
                }
            }
        }
        s.fresh.push_back(p);
    }
    for (int v = 0; v < V; ++v) if (rng() % 2) s.done.insert(v);
    return s;
}

static bool samePoint(const cv::Point3f& a, const cv::Point3f& b) { return std::memcmp(&a, &b, sizeof a) == 0; }
static bool sameCloud(const PointCloud& a, const PointCloud& b) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); ++i) if (!samePoint(a[i].p, b[i].p) || a[i].originatingViews != b[i].originatingViews) return false;
    return true;
}
static bool sameMatches(const MatchMatrix& a, const MatchMatrix& b) {
    for (size_t i = 0; i < a.size(); ++i) for (size_t j = 0; j < a[i].size(); ++j) {
        if (a[i][j].size() != b[i][j].size()) return false;
        for (size_t k = 0; k < a[i][j].size(); ++k)
            if (a[i][j][k].queryIdx != b[i][j][k].queryIdx || a[i][j][k].trainIdx != b[i][j][k].trainIdx || a[i][j][k].distance != b[i][j][k].distance) return false;
    }
    return true;
}

static int checkScene(const Scene& s, const char* what) {
    const MatchIndex index(s.M);
    const Images2D3DMatches ref = sfm_oracle::find2D3DMatches(s.numImages, s.done, s.M, s.feats, s.recon);
    const Images2D3DMatches got = find2D3DMatches(s.numImages, s.done, s.M, index, s.feats, s.recon);
    if (ref.size() != got.size()) { printf("FAIL %s: 2D-3D view count\n", what); return 1; }
    size_t pairs = 0;
    for (const auto& kv : ref) {
        auto it = got.find(kv.first);
        if (it == got.end() || it->second.points2D.size() != kv.second.points2D.size()) { printf("FAIL %s: 2D-3D list of view %d\n", what, kv.first); return 1; }
        for (size_t i = 0; i < kv.second.points2D.size(); ++i)
            if (std::memcmp(&kv.second.points2D[i], &it->second.points2D[i], sizeof(cv::Point2f)) || !samePoint(kv.second.points3D[i], it->second.points3D[i])) {
                printf("FAIL %s: 2D-3D entry %zu of view %d\n", what, i, kv.first); return 1;
            }
        pairs += kv.second.points2D.size();
    }
    PointCloud reconRef = s.recon, reconGot = s.recon;
    MatchMatrix dbgRef(s.numImages, std::vector<Matching>(s.numImages)), dbgGot = dbgRef;
    const MergeCounts cRef = sfm_oracle::mergeNewPointCloud(s.fresh, reconRef, s.M, &dbgRef);
    const MergeCounts cGot = mergeNewPointCloud(s.fresh, reconGot, s.M, index, &dbgGot);
    if (cRef.newPoints != cGot.newPoints || cRef.mergedPoints != cGot.mergedPoints) { printf("FAIL %s: merge counters %zu/%zu vs %zu/%zu\n", what, cRef.newPoints, cRef.mergedPoints, cGot.newPoints, cGot.mergedPoints); return 1; }
    if (!sameCloud(reconRef, reconGot)) { printf("FAIL %s: merged cloud differs\n", what); return 1; }
    if (!sameMatches(dbgRef, dbgGot)) { printf("FAIL %s: merge match matrix differs\n", what); return 1; }
    printf("ok %s: %zu 2D-3D pairs, merge new=%zu merged=%zu dropped=%zu\n", what, pairs, cGot.newPoints, cGot.mergedPoints, s.fresh.size() - cGot.newPoints - cGot.mergedPoints);
    return 0;
}

int main(int argc, char** argv) {
    const unsigned seed = argc > 1 && argv[1][0] != '-' ? (unsigned)atoi(argv[1]) : 1u;
    bool bench = false;
    for (int i = 1; i < argc; ++i) if (!strcmp(argv[i], "--bench")) bench = true;
    int fails = 0;
    for (unsigned t = 0; t < 6; ++t) {
        char name[64];
        snprintf(name, sizeof name, "matcher-like scene %u", t);
        fails += checkScene(makeScene(seed * 100 + t, 5 + (int)t, 200, 400, 300, true), name);
        snprintf(name, sizeof name, "adversarial scene %u", t);
        fails += checkScene(makeScene(seed * 100 + 50 + t, 4 + (int)t, 60, 300, 300, false), name);
    }
    {   // empty inputs
        Scene e; e.numImages = 3; e.feats.resize(3); e.M.assign(3, std::vector<Matching>(3));
        fails += checkScene(e, "empty scene");
    }
    if (bench) {
        const Scene s = makeScene(7, 24, 3000, 40000, 8000, true);
        using clk = std::chrono::steady_clock;
        auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        auto t0 = clk::now(); const MatchIndex index(s.M); auto t1 = clk::now();
        const Images2D3DMatches g = find2D3DMatches(s.numImages, s.done, s.M, index, s.feats, s.recon); auto t2 = clk::now();
        const Images2D3DMatches r = sfm_oracle::find2D3DMatches(s.numImages, s.done, s.M, s.feats, s.recon); auto t3 = clk::now();
        PointCloud a = s.recon, b = s.recon;
        auto t4 = clk::now(); const MergeCounts cg = mergeNewPointCloud(s.fresh, a, s.M, index); auto t5 = clk::now();
        const MergeCounts cr = sfm_oracle::mergeNewPointCloud(s.fresh, b, s.M, nullptr); auto t6 = clk::now();
        size_t pairs = 0; for (const auto& kv : g) pairs += kv.second.points2D.size();
        printf("{\"stage\": \"host glue (SURVEY 8f-1)\", \"views\": %zu, \"features_per_view\": 3000, \"cloud_points\": %zu, \"new_points\": %zu, "
               "\"index_build_ms\": %.2f, \"find2D3DMatches_ms\": %.2f, \"find2D3DMatches_reference_scan_ms\": %.2f, \"pairs\": %zu, "
               "\"mergeNewPointCloud_ms\": %.2f, \"mergeNewPointCloud_reference_scan_ms\": %.2f, \"merged\": %zu, \"identical\": %s}\n",
               s.numImages, s.recon.size(), s.fresh.size(), ms(t0, t1), ms(t1, t2), ms(t2, t3), pairs, ms(t4, t5), ms(t5, t6), cg.mergedPoints,
               (sameCloud(a, b) && cg.newPoints == cr.newPoints && g.size() == r.size()) ? "true" : "false");
    }
    printf(fails ? "FAILED (%d)\n" : "all host-glue checks passed\n", fails);
    return fails ? 1 : 0;
}
