#!/usr/bin/env python
"""Secondary metrics of SURVEY.md section 8(d): all-pairs matching (BASELINE.json config 4, reference-faithful Hamming)
and batched triangulation (config 5), each with resident-input device timing (CUDA events on the library stream),
the host-buffer C-ABI call, a roofline object and a CPU baseline (cv2 = the OpenCV code the reference calls).
Prints one JSON line per stage.  Usage: python tools/bench_stages.py [--images 50] [--features 5000] [--points 1000000]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def tensor_peak_tops():
    """int8 dense rate = 2x the bf16 dense rate; bf16 is the measured cuBLAS figure of MEASURED_PEAKS.json (burst)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return 2.0 * float(json.load(open(p))["bf16_tflops"]), "2 x measured bf16 (MEASURED_PEAKS.json bf16_tflops)"
    except Exception:
        return 2.0 * 1590.0, "2 x fallback bf16"


def _timed(ctx, stream, flush, fn, reps):
    import torch
    ts = []
    for _ in range(reps + 1):
        with torch.cuda.stream(stream):
            flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        ctx.synchronize(); e0.record(stream); fn(); e1.record(stream); ctx.synchronize(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.mean(ts[1:]))


def measure_match(ctx, stream, flush, images=50, features=5000, reps=5, norm="hamming"):
    """BASELINE configs[3]: all-pairs matching of `images` x `features` descriptors.  norm="hamming": the reference-faithful case
    (ORB-256, SfM2DFeatureUtilities.cpp:53-71); norm="l2": the BASELINE wording (SIFT-128, cv::BFMatcher(NORM_L2)), exact u8 GEMM."""
    import torch
    from sfm_toy_library_b200 import capi, synth
    if norm == "hamming":
        descs = synth.make_descriptor_set(images, n=features); kbits = 256
    else:
        descs = [synth.make_sift_like(0, features)]
        for i in range(1, images):
            descs.append(synth.make_sift_like(i, features, prev=descs[-1]))
        kbits = 128
    pairs = [(i, j) for i in range(images) for j in range(i + 1, images)]
    ds = ctx.descriptor_set(descs, norm=norm)
    rows = features * len(pairs)
    dq = torch.empty(rows, dtype=torch.int32, device="cuda"); dt_ = torch.empty(rows, dtype=torch.int32, device="cuda")
    dd = torch.empty(rows, dtype=torch.float32, device="cuda"); dst = torch.empty(len(pairs) + 1, dtype=torch.int32, device="cuda")
    dtot = torch.empty(1, dtype=torch.int64, device="cuda")
    l0 = ctx.kernel_launches
    ms = _timed(ctx, stream, flush, lambda: ds.match_pairs_device(pairs, dq.data_ptr(), dt_.data_ptr(), dd.data_ptr(), dst.data_ptr(), dtot.data_ptr()), reps)
    launches = (ctx.kernel_launches - l0) // (reps + 1)
    assert int(dtot.item()) >= 0, "tensor-core pipeline error flag"
    ds.match_pairs(pairs)                                   # untimed: first call allocates the pinned result staging
    t0 = time.perf_counter(); res = ds.match_pairs(pairs); resident_s = time.perf_counter() - t0
    ds.close()
    # true end-to-end through the batched C-ABI calls with HOST buffers: descriptor upload + expansion (descset_create),
    # all pairs, survivors read back, set destroyed -- what a host that holds cv::Mat descriptors pays
    e2e = []
    packed = np.ascontiguousarray(np.concatenate(descs, 0)); sizes = [len(d) for d in descs]     # the host's descriptor rows, as a C++ caller holds them
    bufs = None
    for _ in range(4):                                      # the three C calls alone: result buffers are the caller's (allocated and touched once)
        t0 = time.perf_counter()
        d2 = capi.DescriptorSet.from_packed(ctx, packed, sizes, norm=norm)
        call, bufs = d2.match_pairs_prepare(pairs, buffers=bufs)
        off2, cnt2 = call()
        d2.close()
        t3 = time.perf_counter()
        e2e.append(t3 - t0)
    e2e_s = min(e2e[1:])
    assert int(cnt2.sum()) == int(sum(len(r[0]) for r in res))
    n_matches = int(sum(len(r[0]) for r in res))
    h2d = int(sum(d.nbytes for d in descs)); d2h = 12 * n_matches + 4 * len(pairs)
    dist_evals = float(features) ** 2 * len(pairs)
    import cv2
    cpu = {}
    for thr in (1, os.cpu_count() or 1):
        cv2.setNumThreads(thr)
        m = cv2.DescriptorMatcher_create("BruteForce-Hamming" if norm == "hamming" else "BruteForce")
        npairs_cpu = 2 if thr == 1 else 8
        t0 = time.perf_counter()
        for (i, j) in pairs[:npairs_cpu]:
            m.knnMatch(descs[i], descs[j], 2)
        cpu[thr] = npairs_cpu / (time.perf_counter() - t0)
    peak, peak_src = tensor_peak_tops()
    ach = 2 * dist_evals * kbits / (ms * 1e-3) / 1e12
    what = "5000x5000 ORB-256 Hamming" if norm == "hamming" else "5000x5000 SIFT-128 L2"
    return {"stage": "match_" + norm, "metric": f"image pairs matched per second ({what}, knn2 + ratio)", "value": len(pairs) / (ms * 1e-3),
            "unit": "pairs/s", "ms_per_step": ms, "pairs": len(pairs), "features": features, "matches": n_matches, "dtype": "s8 x s8 -> s32" if norm == "hamming" else "u8 x u8 -> s32",
            "descriptor_pairs_per_s": dist_evals / (ms * 1e-3), "gpu_launches": int(launches),
            "e2e": {"value": len(pairs) / e2e_s, "unit": "pairs/s", "seconds": e2e_s, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": "sfmb200_descset_create (upload + operand expansion) + sfmb200_match_pairs (caller-owned host result buffers) + destroy; the first repetition (buffer allocation) is dropped",
                    "resident_descriptors_pairs_per_s": len(pairs) / resident_s},
            "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TOP/s", "frac": ach / peak, "peak_source": peak_src,
                         "kernel": "knn2_tc_kernel<L2=%s> (tcgen05 kind::i8)" % ("true" if norm != "hamming" else "false"),
                         "note": f"exact integer GEMM form: 2*Nq*Nt*{kbits} ops per pair; descriptor bytes are negligible (L2-resident)"},
            "cpu_baseline": {"value": cpu[max(cpu)], "unit": "pairs/s", "cores": max(cpu), "kind": "reference", "single_thread_pairs_per_s": cpu[1],
                             "sample": "cv2 BruteForce knnMatch(k=2) on the first pairs of the same set"}}


def measure_triangulate(ctx, stream, flush, points=1_000_000, reps=5):
    """BASELINE configs[4]: triangulateViews on `points` matches in one call."""
    import cv2
    import torch
    from sfm_toy_library_b200 import synth
    hbm, src = peaks()
    p = synth.make_triangulation_problem(points, seed=0)
    m = points
    dl = torch.from_numpy(p["ptsL"]).cuda(); dr = torch.from_numpy(p["ptsR"]).cuda()
    dX = torch.empty(m * 3, dtype=torch.float32, device="cuda"); dk = torch.empty(m, dtype=torch.uint8, device="cuda"); dn = torch.empty(1, dtype=torch.int32, device="cuda")
    ms = _timed(ctx, stream, flush, lambda: ctx.triangulate_device(p["K"], p["Pl"], p["Pr"], dl.data_ptr(), dr.data_ptr(), None, None, m, dX.data_ptr(), dk.data_ptr(), dn.data_ptr()), reps)
    ctx.triangulate(p["K"], p["Pl"], p["Pr"], p["ptsL"], p["ptsR"])      # untimed: first call grows the device scratch
    t0 = time.perf_counter(); X, keep, nk = ctx.triangulate(p["K"], p["Pl"], p["Pr"], p["ptsL"], p["ptsR"]); e2e_s = time.perf_counter() - t0
    from oracle import cv2_reference as ref
    cv2.setNumThreads(os.cpu_count() or 1)
    ns = min(m, 100_000)
    t0 = time.perf_counter(); ref.triangulate_views(p["K"], p["Pl"], p["Pr"], p["ptsL"][:ns], p["ptsR"][:ns]); cpu_s = time.perf_counter() - t0
    abytes = 29 * m
    return {"stage": "triangulate", "metric": "point pairs triangulated per second (DLT + reprojection filter)", "value": m / (ms * 1e-3), "unit": "points/s",
            "ms_per_step": ms, "points": m, "kept": int(nk), "dtype": "f64 inside, f32 in/out", "gpu_launches": 1,
            "e2e": {"value": m / e2e_s, "unit": "points/s", "h2d_bytes_per_step": 16 * m, "d2h_bytes_per_step": 13 * m, "note": "sfmb200_triangulate with host buffers"},
            "roofline": {"bound": "hbm", "achieved": abytes / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": abytes / (ms * 1e-3) / 1e9 / hbm,
                         "peak_source": src, "algorithmic_bytes": abytes, "note": "fp64 ALU bound in practice (closed-form smallest eigenvector, Jacobi SVD fallback)"},
            "cpu_baseline": {"value": ns / cpu_s, "unit": "points/s", "cores": 1, "kind": "reference",
                             "sample": f"cv2 replay of triangulateViews on the first {ns} points (cv::triangulatePoints is serial)"}}


def measure_all(images=50, features=5000, points=1_000_000, reps=5, ctx=None):
    import torch
    from sfm_toy_library_b200 import capi
    own = ctx is None
    if own:
        torch.cuda.set_device(0); ctx = capi.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream)
    flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
    out = [measure_match(ctx, stream, flush, images, features, reps, "hamming"), measure_match(ctx, stream, flush, images, features, reps, "l2"),
           measure_triangulate(ctx, stream, flush, points, reps)]
    if own:
        ctx.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=50)
    ap.add_argument("--features", type=int, default=5000)
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    for line in measure_all(args.images, args.features, args.points, args.reps):
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
