"""Short matcher run for ncu captures: all pairs of --images images x 5000 descriptors, Hamming (tcgen05 s8) and L2 (tcgen05 u8).
    ncu --set full --import-source on -k regex:knn2_tc -c 2 -o gpurun_out/prof_match python tools/prof_match.py"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sfm_toy_library_b200 import capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=12)
a = ap.parse_args()
ctx = capi.Context(0)
pairs = [(i, j) for i in range(a.images) for j in range(i + 1, a.images)]
ds = ctx.descriptor_set(synth.make_descriptor_set(a.images, n=5000))
r = ds.match_pairs(pairs); ds.close()
s = [synth.make_sift_like(0, 5000)]
for i in range(1, a.images):
    s.append(synth.make_sift_like(i, 5000, prev=s[-1]))
d2 = ctx.descriptor_set(s, norm="l2")
r2 = d2.match_pairs(pairs); d2.close()
print(sum(len(x[0]) for x in r), sum(len(x[0]) for x in r2))
ctx.close()
