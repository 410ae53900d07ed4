"""Short BA run for ncu captures: cfg 3 (or --workload cfg2), a handful of LM iterations, nothing else.
    ncu --set full --import-source on -k regex:ba_offdiag -s 2 -c 1 -o gpurun_out/prof python tools/prof_ba.py --iters 4"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sfm_toy_library_b200 import capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg3")
ap.add_argument("--iters", type=int, default=4)
a = ap.parse_args()
p = synth.make_ba_problem(seed=0, **synth.BA_CONFIGS[a.workload])
ctx = capi.Context(0)
prob = ctx.ba_problem(p["cams"], p["pts"], p["focal"], p["obs_xy"], p["obs_cam"], p["pt_off"])
s = prob.run(capi.ba_default_options(max_num_iterations=a.iters, max_solver_time_in_seconds=0.0, function_tolerance=-1.0,
                                     parameter_tolerance=-1.0, gradient_tolerance=-1.0))
print(s["num_iterations"], s["final_cost"], s["kernel_launches"])
prob.close(); ctx.close()
