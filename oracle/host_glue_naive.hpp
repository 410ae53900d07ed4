// oracle/host_glue_naive.hpp -- TEST INFRASTRUCTURE ONLY (the checker for sfm-toy-library_b200/host/sfm_glue.cpp; nothing in the
// product path includes it).  Straight restatement of the reference's two driver scans with their members turned into
// arguments: SfM::find2D3DMatches (reference SfMToyLib/SfM.cpp:471-528) and SfM::mergeNewPointCloud (SfM.cpp:530-600) --
// every loop a linear scan, every `break` where the reference has one.  The reference itself needs OpenCV/Ceres/Boost and
// cannot be built in this image, and its tests pin neither function, so parity is defined by this restatement.
#pragma once
#include "sfm_glue.h"
#include <cmath>

namespace sfm_oracle {
using namespace sfmtoylib;

inline Images2D3DMatches find2D3DMatches(size_t numImages, const std::set<int>& doneViews, const MatchMatrix& M,
                                         const std::vector<Features>& imageFeatures, const PointCloud& cloud) {
    Images2D3DMatches out;
    for (size_t view = 0; view < numImages; ++view) {
        if (doneViews.count((int)view)) continue;                                    // :476-478
        Image2D3DMatch acc;
        for (const Point3DInMap& cp : cloud) {                                       // :483
            bool found = false;
            for (const auto& kv : cp.originatingViews) {                             // :487, ascending view id
                const int oview = kv.first, ofeat = kv.second;
                const bool origIsLeft = (size_t)oview < view;                        // :493-495 (int vs size_t comparison)
                const size_t l = origIsLeft ? (size_t)oview : view, r = origIsLeft ? view : (size_t)oview;
                for (const cv::DMatch& m : M[l][r]) {                                // :498
                    int hit = -1;
                    if (origIsLeft) { if (m.queryIdx == ofeat) hit = m.trainIdx; }   // :500-503
                    else { if (m.trainIdx == ofeat) hit = m.queryIdx; }              // :504-508
                    if (hit >= 0) {                                                  // :509-516
                        acc.points2D.push_back(imageFeatures[view].points[hit]);
                        acc.points3D.push_back(cp.p);
                        found = true;
                        break;
                    }
                }
                if (found) break;                                                    // :518-520
            }
        }
        out[(int)view] = acc;                                                        // :524
    }
    return out;
}

inline MergeCounts mergeNewPointCloud(const PointCloud& cloud, PointCloud& recon, const MatchMatrix& M, MatchMatrix* mergeMatchMatrix) {
    MergeCounts counts;
    for (const Point3DInMap& np : cloud) {                                           // :538
        const cv::Point3f q = np.p;
        bool anyView = false, near3d = false;
        for (Point3DInMap& ep : recon) {                                             // :543
            const float dx = ep.p.x - q.x, dy = ep.p.y - q.y, dz = ep.p.z - q.z;     // Point3f difference, then cv::norm in double
            if (std::sqrt((double)dx * dx + (double)dy * dy + (double)dz * dz) < MERGE_CLOUD_POINT_MIN_MATCH_DISTANCE) {   // :544
                near3d = true;
                for (const auto& nkv : np.originatingViews) {                        // :549
                    for (const auto& ekv : ep.originatingViews) {                    // :553 (the map grows inside, :579)
                        const bool newLeft = nkv.first < ekv.first;                  // :559-563
                        const int lv = newLeft ? nkv.first : ekv.first, lf = newLeft ? nkv.second : ekv.second;
                        const int rv = newLeft ? ekv.first : nkv.first, rf = newLeft ? ekv.second : nkv.second;
                        bool hit = false;
                        for (const cv::DMatch& m : M[lv][rv]) {                      // :566
                            if (m.queryIdx == lf && m.trainIdx == rf && m.distance < MERGE_CLOUD_FEATURE_MIN_MATCH_DISTANCE) {   // :567-569
                                if (mergeMatchMatrix) (*mergeMatchMatrix)[lv][rv].push_back(m);   // :571
                                hit = true;
                                break;
                            }
                        }
                        if (hit) { ep.originatingViews[nkv.first] = nkv.second; anyView = true; }   // :577-583
                    }
                }
            }
            if (anyView) { counts.mergedPoints++; break; }                           // :586-589
        }
        if (!anyView && !near3d) { recon.push_back(np); counts.newPoints++; }         // :591-595
    }
    return counts;
}

}  // namespace sfm_oracle
