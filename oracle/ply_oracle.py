"""oracle/ply_oracle.py -- TEST INFRASTRUCTURE ONLY.  Independent rendering of the text SfM::saveCloudAndCamerasToPLY writes
(reference SfMToyLib/SfM.cpp:630-711): std::ofstream's default formatting of float/double is printf's %g (6 significant digits);
the pixel is addressed through cv::Point(cv::Point2f), i.e. round-half-to-even per coordinate; colours are stored BGR and
written RGB; every vertex line of the point file ends with a blank."""
import numpy as np

_HEAD_POINTS = ["ply                 ", "format ascii 1.0    ", None, "property float x    ", "property float y    ", "property float z    ",
                "property uchar red  ", "property uchar green", "property uchar blue ", "end_header          "]


def _g(v):
    return "%g" % float(v)


def points_ply(cloud, feats, images):
    """cloud: list of (xyz float32[3], {view: feat}); feats: list of float32 [n,2]; images: list of uint8 [h,w,3] BGR."""
    lines = [("element vertex %d" % len(cloud)) if l is None else l for l in _HEAD_POINTS]       # :637-646
    for xyz, views in cloud:
        view = min(views)                                                                           # std::map::begin(), :649
        x, y = feats[view][views[view]]
        px = images[view][int(np.rint(np.float32(y))), int(np.rint(np.float32(x)))]                 # Mat::at<Vec3b>(Point2f), :652
        lines.append("%s %s %s %d %d %d " % (_g(np.float32(xyz[0])), _g(np.float32(xyz[1])), _g(np.float32(xyz[2])), px[2], px[1], px[0]))   # :655-660
    return "\n".join(lines) + "\n"


def cameras_ply(poses):
    """poses: float32 [n,3,4]."""
    n = len(poses)
    lines = ["ply                 ", "format ascii 1.0    ", "element vertex %d" % (4 * n), "property float x    ", "property float y    ",
             "property float z    ", "element edge %d" % (3 * n), "property int vertex1", "property int vertex2", "property uchar red  ",
             "property uchar green", "property uchar blue ", "end_header          "]                  # :668-681
    for P in poses:
        P = np.asarray(P, np.float32).astype(np.float64)
        c = P[:, 3]
        lines.append("%s %s %s" % (_g(c[0]), _g(c[1]), _g(c[2])))
        for a in range(3):                                                                          # :685-692
            t = c + P[:, a] * 0.2
            lines.append("%s %s %s" % (_g(t[0]), _g(t[1]), _g(t[2])))
    for i in range(n):                                                                              # :697-707
        lines += ["%d %d 255 0 0" % (4 * i, 4 * i + 1), "%d %d 0 255 0" % (4 * i, 4 * i + 2), "%d %d 0 0 255" % (4 * i, 4 * i + 3)]
    return "\n".join(lines) + "\n"
