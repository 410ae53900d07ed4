#!/usr/bin/env python
"""Diagnostic: wall time per LM iteration of problem.run for profile / flush / chunk combinations (cfg3)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from sfm_toy_library_b200 import capi, synth
import bench
torch.cuda.set_device(0)
p = bench.make_shard("cfg3", 0)
a = (p["cams"], p["pts"], p["focal"], p["obs_xy"], p["obs_cam"], p["pt_off"])
for chunk in ("4", "1"):
    os.environ["SFMB200_BA_CHUNK"] = chunk
    ctx = capi.Context(0)
    prob = ctx.ba_problem(*a)
    prob.run(bench.fixed_iteration_options(capi, 100)); prob.reset()
    for profile in (0, 1):
        for flush in (0, 192):
            ts = []
            for rep in range(3):
                prob.reset(); ctx.synchronize()
                t0 = time.perf_counter()
                s = prob.run(bench.fixed_iteration_options(capi, 20, profile=profile, l2_flush_mb=flush))
                ctx.synchronize()
                ts.append((time.perf_counter() - t0) * 1e3 / 20)
            print(f"chunk={chunk} profile={profile} flush={flush}: ms/iter {['%.3f' % t for t in ts]} iters={s['num_iterations']} flush_ms_total={s['flush_ms_total']:.3f} solve_ms={s['solve_ms_total']:.3f}", flush=True)
    prob.close(); ctx.close()
