// sfm_glue.cpp -- indexed versions of SfM::find2D3DMatches and SfM::mergeNewPointCloud (see sfm_glue.h).
#include "sfm_glue.h"
#include <algorithm>
#include <cmath>

namespace sfmtoylib {

MatchIndex::MatchIndex(const MatchMatrix& m) {
    pairs_.resize(m.size());
    for (size_t i = 0; i < m.size(); ++i) {
        pairs_[i].resize(m[i].size());
        for (size_t j = 0; j < m[i].size(); ++j) {
            const Matching& list = m[i][j];
            Pair& p = pairs_[i][j];
            int maxq = -1, maxt = -1;
            for (const cv::DMatch& d : list) { maxq = std::max(maxq, d.queryIdx); maxt = std::max(maxt, d.trainIdx); }
            p.qoff.assign((size_t)(maxq + 2), 0); p.tfirst.assign((size_t)(maxt + 1), -1);
            for (const cv::DMatch& d : list) if (d.queryIdx >= 0) p.qoff[d.queryIdx + 1]++;
            for (size_t q = 1; q < p.qoff.size(); ++q) p.qoff[q] += p.qoff[q - 1];
            p.qpos.resize(p.qoff.empty() ? 0 : (size_t)p.qoff.back());
            std::vector<int32_t> cursor(p.qoff.begin(), p.qoff.end());
            for (size_t k = 0; k < list.size(); ++k) {                         // stable: list order inside a query group
                const cv::DMatch& d = list[k];
                if (d.queryIdx >= 0) p.qpos[cursor[d.queryIdx]++] = (int32_t)k;
                if (d.trainIdx >= 0 && d.queryIdx >= 0 && p.tfirst[d.trainIdx] < 0) p.tfirst[d.trainIdx] = (int32_t)k;   // first USABLE hit (:512)
            }
        }
    }
}

Images2D3DMatches find2D3DMatches(size_t numImages, const std::set<int>& doneViews, const MatchMatrix& M, const MatchIndex& index,
                                  const std::vector<Features>& imageFeatures, const PointCloud& cloud) {
    Images2D3DMatches matches;
    for (size_t viewIdx = 0; viewIdx < numImages; ++viewIdx) {
        if (doneViews.find((int)viewIdx) != doneViews.end()) continue;          // SfM.cpp:476
        Image2D3DMatch match2D3D;
        const Features& newViewFeatures = imageFeatures[viewIdx];
        for (const Point3DInMap& cloudPoint : cloud) {
            for (const auto& origViewAndPoint : cloudPoint.originatingViews) {   // ascending view id (std::map)
                const int origView = origViewAndPoint.first, origFeat = origViewAndPoint.second;
                int matched = -1;
                if ((size_t)origView < viewIdx) {                                // originating view is 'left': first m with queryIdx == feature
                    const int32_t *qb, *qe;                                      // (and a usable trainIdx, :512)
                    index.byQuery(origView, (int)viewIdx, origFeat, qb, qe);
                    for (const int32_t* it = qb; it != qe && matched < 0; ++it) matched = M[origView][viewIdx][*it].trainIdx;
                } else {                                                         // originating view is 'right': first m with trainIdx == feature
                    const int pos = index.firstByTrain((int)viewIdx, origView, origFeat);
                    if (pos >= 0) matched = M[viewIdx][origView][pos].queryIdx;
                }
                if (matched >= 0) {
                    match2D3D.points2D.push_back(newViewFeatures.points[matched]);
                    match2D3D.points3D.push_back(cloudPoint.p);
                    break;                                                        // SfM.cpp:514, :519
                }
            }
        }
        matches[(int)viewIdx] = match2D3D;
    }
    return matches;
}

namespace {

// cv::norm(Point3f - Point3f): component differences in float, squares and root in double
inline double pointDistance(const cv::Point3f& a, const cv::Point3f& b) {
    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return std::sqrt((double)dx * dx + (double)dy * dy + (double)dz * dz);
}

// uniform grid over the cloud, cell edge slightly above the merge radius: every point closer than the radius to q lies in
// one of the 27 cells around q's cell
class CloudGrid {
public:
    explicit CloudGrid(double cell) : inv_(1.0 / cell) {}
    void insert(const cv::Point3f& p, int id) { cells_[key(cx(p.x), cx(p.y), cx(p.z))].push_back(id); }
    // ids of the points in the 27 cells around p, ascending (= cloud order)
    void candidates(const cv::Point3f& p, std::vector<int>& out) const {
        out.clear();
        const long long x = cx(p.x), y = cx(p.y), z = cx(p.z);
        for (long long a = x - 1; a <= x + 1; ++a)
            for (long long b = y - 1; b <= y + 1; ++b)
                for (long long c = z - 1; c <= z + 1; ++c) {
                    auto it = cells_.find(key(a, b, c));
                    if (it != cells_.end()) out.insert(out.end(), it->second.begin(), it->second.end());
                }
        std::sort(out.begin(), out.end());
    }
private:
    long long cx(float v) const { return (long long)std::floor((double)v * inv_); }
    static uint64_t key(long long a, long long b, long long c) {
        return ((uint64_t)(a & 0x1fffff) << 42) ^ ((uint64_t)(b & 0x1fffff) << 21) ^ (uint64_t)(c & 0x1fffff) ^ ((uint64_t)(a >> 21) * 0x9e3779b97f4a7c15ull)
               ^ ((uint64_t)(b >> 21) * 0xc2b2ae3d27d4eb4full) ^ ((uint64_t)(c >> 21) * 0x165667b19e3779f9ull);
    }
    double inv_;
    std::unordered_map<uint64_t, std::vector<int>> cells_;
};

}  // namespace

MergeCounts mergeNewPointCloud(const PointCloud& cloud, PointCloud& reconstructionCloud, const MatchMatrix& M, const MatchIndex& index,
                               MatchMatrix* mergeMatchMatrix) {
    MergeCounts counts;
    // A hash collision only adds candidates (the exact distance test below decides), it never hides one.
    CloudGrid grid(1.05 * (double)MERGE_CLOUD_POINT_MIN_MATCH_DISTANCE);
    for (size_t i = 0; i < reconstructionCloud.size(); ++i) grid.insert(reconstructionCloud[i].p, (int)i);
    std::vector<int> cand;
    for (const Point3DInMap& p : cloud) {
        const cv::Point3f newPoint = p.p;
        bool foundAnyMatchingExistingViews = false, foundMatching3DPoint = false;
        grid.candidates(newPoint, cand);
        for (int id : cand) {                                                    // cloud order, like the reference's scan (:543)
            Point3DInMap& existingPoint = reconstructionCloud[id];
            if (pointDistance(existingPoint.p, newPoint) < MERGE_CLOUD_POINT_MIN_MATCH_DISTANCE) {
                foundMatching3DPoint = true;
                for (const auto& newKv : p.originatingViews) {
                    // the map is extended inside this loop (:579); std::map iterators stay valid and later keys ARE visited
                    for (const auto& existingKv : existingPoint.originatingViews) {
                        const bool newIsLeft = newKv.first < existingKv.first;
                        const int leftViewIdx = newIsLeft ? newKv.first : existingKv.first;
                        const int leftViewFeatureIdx = newIsLeft ? newKv.second : existingKv.second;
                        const int rightViewIdx = newIsLeft ? existingKv.first : newKv.first;
                        const int rightViewFeatureIdx = newIsLeft ? existingKv.second : newKv.second;
                        // first match with queryIdx == left feature, trainIdx == right feature and distance < 20 (:566-569):
                        // only the matches of that query index can qualify, in list order
                        bool foundMatchingFeature = false;
                        const int32_t *qb, *qe;
                        index.byQuery(leftViewIdx, rightViewIdx, leftViewFeatureIdx, qb, qe);
                        for (const int32_t* it = qb; it != qe; ++it) {
                            const cv::DMatch& match = M[leftViewIdx][rightViewIdx][*it];
                            if (match.trainIdx == rightViewFeatureIdx && match.distance < MERGE_CLOUD_FEATURE_MIN_MATCH_DISTANCE) {
                                if (mergeMatchMatrix) (*mergeMatchMatrix)[leftViewIdx][rightViewIdx].push_back(match);
                                foundMatchingFeature = true;
                                break;
                            }
                        }
                        if (foundMatchingFeature) {
                            existingPoint.originatingViews[newKv.first] = newKv.second;
                            foundAnyMatchingExistingViews = true;
                        }
                    }
                }
            }
            if (foundAnyMatchingExistingViews) { counts.mergedPoints++; break; }
        }
        if (!foundAnyMatchingExistingViews && !foundMatching3DPoint) {
            reconstructionCloud.push_back(p);
            grid.insert(p.p, (int)reconstructionCloud.size() - 1);
            counts.newPoints++;
        }
    }
    return counts;
}

}  // namespace sfmtoylib
