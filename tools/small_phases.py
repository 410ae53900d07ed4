import sys, time, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from cfg1_util import Cfg1
from sfm_toy_library_b200 import capi
c = Cfg1(); ctx = capi.Context(0)
for k in (0, 2, 5):
    b = c.ba(k); a = (b["cams"], b["pts"], b["focal"], b["obs_xy"], b["obs_cam"], b["pt_off"])
    for rep in range(3):
        t0 = time.perf_counter(); p = ctx.ba_problem(*a); t1 = time.perf_counter(); s = p.run(); t2 = time.perf_counter(); p.download(); t3 = time.perf_counter(); p.close(); t4 = time.perf_counter()
    print(k, s["num_iterations"], "create %.0f run %.0f download %.0f close %.0f us" % ((t1-t0)*1e6, (t2-t1)*1e6, (t3-t2)*1e6, (t4-t3)*1e6), s["kernel_launches"])
