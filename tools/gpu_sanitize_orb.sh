#!/bin/bash
# compute-sanitizer racecheck / synccheck / memcheck over the ORB stage (two small images in one batch, checked against the oracle)
set -u
cd "$(dirname "$0")/.."
run() { echo "=== $*"; timeout 900 "$@" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard|ok:|Traceback" | head -20; }
ORB='import sys; sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from sfm_toy_library_b200 import capi
from oracle import orb_oracle as O
from orb_util import textured, blobs
imgs = [textured(240, 320, 5), blobs(240, 320, 6)]
ctx = capi.Context(0)
out = ctx.orb_detect_and_compute(imgs, 800)
for im, (k, d) in zip(imgs, out):
    ko, do = O.detect_and_compute(im, 800)
    assert np.array_equal(k[:, :6], ko) and np.array_equal(d, do)
print("ok:", [len(k) for k, _ in out])'
for tool in racecheck synccheck memcheck; do run compute-sanitizer --tool $tool python -c "$ORB"; done
