import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sfm_toy_library_b200 import capi, synth
from oracle import oracle
ctx = capi.Context(0)
for nq, nt in ((128, 5000), (5000, 256), (5000, 768), (300, 2048), (5000, 5000)):
    t = synth.make_descriptors(nt % 97, nt); q = synth.make_descriptors(nq % 89 + 100, nq, prev=t)
    for mode in ("tc",):
        os.environ["SFMB200_MATCH"] = mode
        gq, gt, gd = ctx.match_knn2_ratio(q, t)
        oq, ot, od = oracle.match_hamming(q, t)
        idx, dist = oracle.knn2_hamming(q, t)
        extra = np.setdiff1d(gq, oq); missing = np.setdiff1d(oq, gq)
        print(nq, nt, mode, "got", len(gq), "want", len(oq), "extra", len(extra), "missing", len(missing))
        for r in extra[:6]:
            k = np.where(gq == r)[0][0]
            print("   row", r, "gpu best", gt[k], gd[k], " oracle top2 idx", idx[r], "dist", dist[r], " second-best tile", idx[r][1] // 256, "col", idx[r][1] % 256)
