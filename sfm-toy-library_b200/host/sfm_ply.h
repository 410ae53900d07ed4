// sfm_ply.h -- SURVEY.md 8(f-4): the reference's output format, SfM::saveCloudAndCamerasToPLY (reference SfMToyLib/SfM.cpp:630-711),
// with its members as arguments.  Two ASCII PLY files: <prefix>_points.ply (x y z + the colour of the pixel under the point's
// FIRST originating view's feature, BGR -> RGB) and <prefix>_cameras.ply (4 vertices and 3 coloured edges per camera, axis
// length 0.2).  Byte-for-byte the text std::ofstream produces there: default float formatting (6 significant digits), the
// padded header lines, the trailing blank before each vertex line's newline.
#pragma once
#include "sfmtoylib_b200.h"
#include <string>
#include <vector>

namespace sfmtoylib {

// images[v]: CV_8UC3, BGR, the image features of view v were detected in (mImages, SfM.cpp:652)
bool saveCloudAndCamerasToPLY(const std::string& prefix, const PointCloud& reconstructionCloud, const std::vector<Features>& imageFeatures,
                              const std::vector<cv::Mat>& images, const std::vector<Pose>& cameraPoses);

}  // namespace sfmtoylib
