"""tools/bench_orb.py -- SURVEY.md section 8 row f-3: ORB extraction stage line.  `images` equally sized 1024 x 768 synthetic images
(cfg-1 geometry: 7 images; cfg-4 geometry: 50) through sfmb200_orb_detect_and_compute_batch from HOST buffers (upload, three host round
trips and the two retainBest selections inside the timed region), beside cv2.ORB_create(5000).detectAndCompute on the host cores.
    python tools/bench_orb.py [--images 7] [--reps 5] [--no-cpu]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from orb_util import textured  # noqa: E402
from sfm_toy_library_b200 import capi  # noqa: E402


def run(n_images=7, reps=5, cpu=True, nfeatures=5000, w=1024, h=768, channels=3, ctx=None):
    imgs = [textured(h, w, 100 + i) for i in range(n_images)]
    if channels == 3:
        imgs = [np.ascontiguousarray(np.stack([im, np.roll(im, 1, 0), np.roll(im, 1, 1)], 2)) for im in imgs]
    own = ctx is None
    ctx = ctx or capi.Context(0)
    ctx.orb_detect_and_compute(imgs, nfeatures)                       # warm-up: allocations
    l0 = ctx.kernel_launches
    out = ctx.orb_detect_and_compute(imgs, nfeatures)
    call, _, _, cnt = ctx.orb_prepare(imgs, nfeatures)                # the C-ABI call alone: host buffers in, caller-owned outputs
    call()
    l0 = ctx.kernel_launches
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); call(); ts.append(time.perf_counter() - t)
    assert [int(c) for c in cnt] == [len(k) for k, _ in out]
    launches = (ctx.kernel_launches - l0) // reps
    phases = ctx.orb_last_timings()
    call1 = ctx.orb_prepare(imgs[:1], nfeatures)[0]
    call1()
    t1 = []
    for _ in range(reps):
        t = time.perf_counter(); call1(); t1.append(time.perf_counter() - t)
    if own:
        ctx.close()
    best = min(ts)
    line = {"stage": "orb_extract", "metric": "images per second (ORB(%d) detectAndCompute, %dx%d, %d channel(s))" % (nfeatures, w, h, channels),
            "value": n_images / best, "unit": "images/s", "images": n_images, "keypoints": int(sum(len(k) for k, _ in out)),
            "ms_per_batch": best * 1e3, "ms_per_image_batched": best * 1e3 / n_images, "ms_single_image_call": min(t1) * 1e3,
            "gpu_launches": int(launches), "host_phases_ms": {k: round(v, 3) for k, v in phases.items()}, "dtype": "u8 (f32 for Harris / angle / blur)",
            "e2e": {"value": n_images / best, "unit": "images/s", "h2d_bytes_per_step": int(sum(im.nbytes for im in imgs)),
                    "d2h_bytes_per_step": int(sum(len(k) for k, _ in out)) * 60,
                    "note": "one sfmb200_orb_detect_and_compute_batch call: pageable host images in, key points + descriptors out; staging, upload, the three host round trips and both retainBest selections inside"}}
    if cpu:
        import cv2
        orb = cv2.ORB_create(nfeatures)
        res = {}
        for nt in (1, os.cpu_count() or 1):
            cv2.setNumThreads(nt)
            orb.detectAndCompute(imgs[0], None)
            t = time.perf_counter()
            for im in imgs[:7]:
                orb.detectAndCompute(im, None)
            res[nt] = (time.perf_counter() - t) / min(7, n_images)
        cv2.setNumThreads(-1)
        best_nt = min(res, key=res.get)
        line["cpu_baseline"] = {"value": 1.0 / res[best_nt], "unit": "images/s", "cores": best_nt, "kind": "reference",
                                "sample": "cv2 %s ORB_create(%d).detectAndCompute on 7 of the images; 1 thread: %.1f ms/image, %d threads: %.1f ms/image"
                                          % (cv2.__version__, nfeatures, res[1] * 1e3, max(res), res[max(res)] * 1e3)}
    return line


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=7); ap.add_argument("--reps", type=int, default=5); ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--gray", action="store_true")
    a = ap.parse_args()
    print(json.dumps(run(a.images, a.reps, not a.no_cpu, channels=1 if a.gray else 3)))
