#!/usr/bin/env python
"""Diagnostic: the timed pass of bench.py under toggles (pre-heat iterations, torch event/stream use, second pass)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from sfm_toy_library_b200 import capi
import bench
torch.cuda.set_device(0)
p = bench.make_shard("cfg3", 0)
a = (p["cams"], p["pts"], p["focal"], p["obs_xy"], p["obs_cam"], p["pt_off"])
for preheat in (0, 150, 20):
    for use_torch in (0, 1):
        ctx = capi.Context(0)
        stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
        prob = ctx.ba_problem(*a)
        flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
        if preheat:
            prob.run(bench.fixed_iteration_options(capi, preheat)); prob.reset()
        s0 = prob.run(bench.fixed_iteration_options(capi, 3)); prob.reset()
        if use_torch:
            with torch.cuda.stream(stream):
                flush.zero_()
        torch.cuda.synchronize(); ctx.synchronize()
        out = []
        for rep in range(2):
            if use_torch:
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record(stream)
            t0 = time.perf_counter()
            s = prob.run(bench.fixed_iteration_options(capi, 20, profile=1, l2_flush_mb=192))
            if use_torch:
                e1.record(stream)
            torch.cuda.synchronize(); ctx.synchronize()
            out.append("%.3f" % ((time.perf_counter() - t0) * 1e3 / 20))
            prob.reset()
        print(f"preheat={preheat} torch={use_torch}: ms/iter {out}  term={s['message']!r} iters={s['num_iterations']} succ={s['num_successful_steps']} final={s['final_cost']:.6e}", flush=True)
        if preheat == 150 and use_torch == 0:
            sp = prob.run(bench.fixed_iteration_options(capi, 150)); print("  preheat summary:", sp["message"], sp["num_iterations"], sp["num_successful_steps"], sp["num_unsuccessful_steps"], "%.3f ms/iter" % (sp["total_time_s"] * 1e3 / max(1, sp["num_iterations"])), flush=True)
        prob.close(); ctx.close(); del flush
