"""the indexed find2D3DMatches / mergeNewPointCloud (host/sfm_glue.cpp) against the naive restatement of
the reference's scans (oracle/host_glue_naive.hpp, reference SfM.cpp:471-600).  Pure C++: runs on the CPU box."""
import importlib.util
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def glue_bin():
    spec = importlib.util.spec_from_file_location("sfmb200_build", os.path.join(ROOT, "sfm-toy-library_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    return mod.build_glue()


@pytest.mark.parametrize("seed", [1, 2, 3, 11, 29])
def test_indexed_glue_equals_reference_scans(glue_bin, seed):
    """Matcher-like lists (ascending unique query indices), adversarial lists (repeated query/train indices, arbitrary order),
    points inside / around / outside the merge radius, with and without confirming feature matches, empty inputs."""
    r = subprocess.run([glue_bin, str(seed)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all host-glue checks passed" in r.stdout
    assert r.stdout.count("ok ") == 13


def test_glue_bench_line_is_consistent(glue_bin):
    r = subprocess.run([glue_bin, "5", "--bench"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["identical"] is True and d["pairs"] > 0 and d["merged"] > 0
    # the whole point: no scan over the match lists / the cloud any more
    assert d["find2D3DMatches_ms"] < d["find2D3DMatches_reference_scan_ms"]
    assert d["mergeNewPointCloud_ms"] < d["mergeNewPointCloud_reference_scan_ms"]
