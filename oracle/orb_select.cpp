// oracle/orb_select.cpp -- TEST INFRASTRUCTURE ONLY (part of the ORB oracle, oracle/orb_oracle.py).
//
// cv::KeyPointsFilter::retainBest (OpenCV features2d, keypoint.cpp; un-vendored dependency of the reference, called by
// cv::ORB::detectAndCompute which the reference calls at SfM2DFeatureUtilities.cpp:48): when more than n_points key points are
// given, std::nth_element by descending response, then every key point behind position n_points whose response equals the
// n_points-th one is kept too (std::partition).  The ORDER in which key points come out is the permutation libstdc++'s
// introselect leaves behind -- it only depends on the comparison results, so running the same two algorithms on
// (response, original index) records reproduces it.
#include <algorithm>
#include <cstdint>
#include <vector>

namespace {
struct Rec { float response; int32_t idx; };
struct ResponseGreater { bool operator()(const Rec& a, const Rec& b) const { return a.response > b.response; } };
struct ResponseAtLeast { float v; bool operator()(const Rec& k) const { return k.response >= v; } };
}  // namespace

extern "C" int orb_oracle_retain_best(const float* response, int n, int n_points, int32_t* order) {
    std::vector<Rec> k(n);
    for (int i = 0; i < n; i++) { k[i].response = response[i]; k[i].idx = i; }
    if (n_points >= 0 && n > n_points) {
        if (n_points == 0) return 0;
        std::nth_element(k.begin(), k.begin() + n_points - 1, k.end(), ResponseGreater());
        const float ambiguous = k[n_points - 1].response;
        auto new_end = std::partition(k.begin() + n_points, k.end(), ResponseAtLeast{ambiguous});
        k.resize(new_end - k.begin());
    }
    for (size_t i = 0; i < k.size(); i++) order[i] = k[i].idx;
    return (int)k.size();
}
