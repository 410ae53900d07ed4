// sfm_ply.cpp -- ASCII PLY writers of the reference (see sfm_ply.h).
#include "sfm_ply.h"
#include <cmath>
#include <fstream>

namespace sfmtoylib {

namespace {
// cv::Point_<int>(cv::Point2f) = saturate_cast<int> per coordinate = round half to even (what Mat::at<Vec3b>(Point2f) goes through)
inline int roundCoord(float v) { return (int)std::lrint((double)v); }
}  // namespace

bool saveCloudAndCamerasToPLY(const std::string& prefix, const PointCloud& cloud, const std::vector<Features>& imageFeatures,
                              const std::vector<cv::Mat>& images, const std::vector<Pose>& cameraPoses) {
    std::ofstream ofs(prefix + "_points.ply");
    if (!ofs) return false;
    ofs << "ply                 " << std::endl
        << "format ascii 1.0    " << std::endl
        << "element vertex " << cloud.size() << std::endl
        << "property float x    " << std::endl
        << "property float y    " << std::endl
        << "property float z    " << std::endl
        << "property uchar red  " << std::endl
        << "property uchar green" << std::endl
        << "property uchar blue " << std::endl
        << "end_header          " << std::endl;
    for (const Point3DInMap& p : cloud) {
        const auto first = p.originatingViews.begin();                       // colour from the first originating view (:648-652)
        const int view = first->first;
        const cv::Point2f p2d = imageFeatures[view].points[first->second];
        const cv::Mat& img = images[view];
        const unsigned char* px = img.data + ((size_t)roundCoord(p2d.y) * img.cols + roundCoord(p2d.x)) * 3;
        ofs << p.p.x << " " << p.p.y << " " << p.p.z << " " << (int)px[2] << " " << (int)px[1] << " " << (int)px[0] << " " << std::endl;
    }
    ofs.close();

    std::ofstream ofsc(prefix + "_cameras.ply");
    if (!ofsc) return false;
    ofsc << "ply                 " << std::endl
         << "format ascii 1.0    " << std::endl
         << "element vertex " << (cameraPoses.size() * 4) << std::endl
         << "property float x    " << std::endl
         << "property float y    " << std::endl
         << "property float z    " << std::endl
         << "element edge " << (cameraPoses.size() * 3) << std::endl
         << "property int vertex1" << std::endl
         << "property int vertex2" << std::endl
         << "property uchar red  " << std::endl
         << "property uchar green" << std::endl
         << "property uchar blue " << std::endl
         << "end_header          " << std::endl;
    for (const Pose& pose : cameraPoses) {                                     // centre and the three axis tips, in double (:689-692)
        const double c[3] = {(double)pose(0, 3), (double)pose(1, 3), (double)pose(2, 3)};
        ofsc << c[0] << " " << c[1] << " " << c[2] << std::endl;
        for (int axis = 0; axis < 3; ++axis)
            ofsc << c[0] + (double)pose(0, axis) * 0.2 << " " << c[1] + (double)pose(1, axis) * 0.2 << " " << c[2] + (double)pose(2, axis) * 0.2 << std::endl;
    }
    for (size_t i = 0; i < cameraPoses.size(); ++i) {
        ofsc << (i * 4 + 0) << " " << (i * 4 + 1) << " " << "255 0 0" << std::endl;
        ofsc << (i * 4 + 0) << " " << (i * 4 + 2) << " " << "0 255 0" << std::endl;
        ofsc << (i * 4 + 0) << " " << (i * 4 + 3) << " " << "0 0 255" << std::endl;
    }
    return true;
}

}  // namespace sfmtoylib
