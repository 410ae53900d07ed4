"""the reference's ASCII PLY output (host/sfm_ply.cpp) byte for byte against an independent rendering
(oracle/ply_oracle.py) of SfM::saveCloudAndCamerasToPLY (reference SfM.cpp:630-711)."""
import importlib.util
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import ply_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ply_bin():
    spec = importlib.util.spec_from_file_location("sfmb200_build", os.path.join(ROOT, "sfm-toy-library_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    return mod.build_ply()


def _scene(seed, nviews=4, npoints=300, ncams=4):
    rs = np.random.RandomState(seed)
    feats, images = [], []
    for v in range(nviews):
        h, w = int(rs.randint(40, 90)), int(rs.randint(50, 120))
        n = int(rs.randint(20, 60))
        pts = np.stack([rs.uniform(0, w - 1.001, n), rs.uniform(0, h - 1.001, n)], 1).astype(np.float32)
        pts[:4] = np.floor(pts[:4]) + 0.5                     # exact .5 coordinates: round-half-to-even cases
        pts[:4] = np.minimum(pts[:4], [w - 1.5, h - 1.5])
        feats.append(pts); images.append(rs.randint(0, 256, (h, w, 3)).astype(np.uint8))
    cloud = []
    for i in range(npoints):
        scale = [1.0, 1e-4, 1e4, 123456.789][i % 4]          # exercises %g's switch between fixed and exponent notation
        xyz = (rs.normal(0, 1, 3) * scale).astype(np.float32)
        views = {}
        for v in rs.choice(nviews, int(rs.randint(1, nviews + 1)), replace=False):
            views[int(v)] = int(rs.randint(0, len(feats[int(v)])))
        cloud.append((xyz, views))
    poses = rs.normal(0, 3, (ncams, 3, 4)).astype(np.float32)
    return feats, images, cloud, poses


def _write_scene(path, feats, images, cloud, poses):
    with open(path, "wb") as f:
        f.write(struct.pack("<i", len(feats)))
        for pts, img in zip(feats, images):
            f.write(struct.pack("<iii", img.shape[0], img.shape[1], len(pts))); f.write(pts.astype("<f4").tobytes()); f.write(img.tobytes())
        f.write(struct.pack("<i", len(cloud)))
        for xyz, views in cloud:
            f.write(np.asarray(xyz, "<f4").tobytes()); f.write(struct.pack("<i", len(views)))
            for v, ft in views.items():
                f.write(struct.pack("<ii", v, ft))
        f.write(struct.pack("<i", len(poses))); f.write(np.asarray(poses, "<f4").tobytes())


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_ply_files_match_reference_text(ply_bin, tmp_path, seed):
    feats, images, cloud, poses = _scene(seed)
    scene = tmp_path / "scene.bin"; _write_scene(scene, feats, images, cloud, poses)
    prefix = str(tmp_path / "out")
    subprocess.run([ply_bin, str(scene), prefix], check=True, timeout=60)
    assert open(prefix + "_points.ply").read() == ply_oracle.points_ply(cloud, feats, images)
    assert open(prefix + "_cameras.ply").read() == ply_oracle.cameras_ply(poses)


def test_ply_empty_cloud(ply_bin, tmp_path):
    feats, images, cloud, poses = _scene(5, npoints=0, ncams=0)
    scene = tmp_path / "scene.bin"; _write_scene(scene, feats, images, cloud, poses)
    prefix = str(tmp_path / "out")
    subprocess.run([ply_bin, str(scene), prefix], check=True, timeout=60)
    assert open(prefix + "_points.ply").read() == ply_oracle.points_ply([], feats, images)
    assert open(prefix + "_cameras.ply").read() == ply_oracle.cameras_ply(poses)
