#!/bin/bash
# quick compute-sanitizer pass over the default BA path only (racecheck, synccheck, memcheck); see gpu_sanitize.sh for the full set
set -u
cd "$(dirname "$0")/.."
run() { echo "=== $*"; timeout 600 "$@" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard|ok:|Traceback" | head -20; }
BA='import sys; sys.path.insert(0, "."); from sfm_toy_library_b200 import capi, synth
p = synth.make_ba_problem(n_cams=20, n_pts=3000, obs_per_pt=8, seed=0)
ctx = capi.Context(0)
c, q, f, s = ctx.ba_solve(p["cams"], p["pts"], p["focal"], p["obs_xy"], p["obs_cam"], p["pt_off"], capi.ba_default_options(max_num_iterations=4))
print("ok:", s["num_iterations"], s["final_cost"])'
for tool in racecheck synccheck memcheck; do SFMB200_BA_GRAPH=0 run compute-sanitizer --tool $tool python -c "$BA"; done
