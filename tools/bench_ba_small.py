"""Latency regime of adjustBundle inside runSfM (BASELINE configs[0]): the six bundle adjustments of the crazyhorse replay
(2..7 cameras, 123..1430 points), one-shot sfmb200_ba_solve with host buffers, per call and per LM iteration."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cfg1_util import Cfg1  # noqa: E402
from sfm_toy_library_b200 import capi  # noqa: E402

c = Cfg1(); ctx = capi.Context(0)
out = []
for k in range(c.n_ba):
    b = c.ba(k)
    a = (b["cams"], b["pts"], b["focal"], b["obs_xy"], b["obs_cam"], b["pt_off"])
    ctx.ba_solve(*a)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); s = ctx.ba_solve(*a)[3]; ts.append(time.perf_counter() - t0)
    out.append({"call": k, "cams": len(b["cams"]), "points": len(b["pts"]), "obs": len(b["obs_cam"]), "iterations": s["num_iterations"],
                "ms": 1e3 * min(ts), "us_per_iteration": 1e6 * min(ts) / max(1, s["num_iterations"]), "launches": s["kernel_launches"]})
print(json.dumps({"chunk": os.environ.get("SFMB200_BA_CHUNK", "default"), "calls": out, "total_ms": sum(o["ms"] for o in out)}))
