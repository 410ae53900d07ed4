"""Shared loader for the BASELINE configs[0] fixture (tests/golden/cfg1_crazyhorse.npz, made by tests/golden/make_cfg1.py):
ORB(5000) features of the 7 crazyhorse images + the stage-by-stage trace of a cv2/oracle replay of SfM::runSfM."""
import os

import numpy as np

from sfm_toy_library_b200.stages import Features

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cfg1_crazyhorse.npz")


class Cfg1:
    def __init__(self):
        g = np.load(GOLDEN)
        self.g = g
        self.n_img = len(g["files"])
        self.size = tuple(int(x) for x in g["image_size"])
        self.features = [Features(points=g[f"pts_{i}"], descriptors=g[f"desc_{i}"]) for i in range(self.n_img)]
        self.pairs = [tuple(int(x) for x in p) for p in g["pairs"]]
        self.matches = [(g[f"match_{p}_q"], g[f"match_{p}_t"], g[f"match_{p}_d"]) for p in range(len(self.pairs))]
        self.n_tri = int(g["n_tri"]); self.n_ba = int(g["n_ba"]); self.n_pnp = int(g["n_pnp"])

    def tri(self, k):
        g = self.g; p = f"tri_{k}_"
        return {n: g[p + n] for n in ("pair", "K", "Pl", "Pr", "mq", "mt", "X", "back")}

    def ba(self, k):
        g = self.g; p = f"ba_{k}_"
        d = {n: g[p + n] for n in ("cams", "pts", "obs_xy", "obs_cam", "pt_off", "used", "K_after", "poses_after", "pts_after", "summary", "cost")}
        d["focal"] = float(g[p + "focal"])
        return d
