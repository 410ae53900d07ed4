"""The C++ drop-in shim (sfm-toy-library_b200/host/shim.cpp: matchFeatures / triangulateViews / adjustBundle with the
reference's signatures over the C ABI) exercised by its C++ test program, which replays the reference's
triangulate_from_2_views scene (SfMUnitTests.cpp:221-251)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_cpp_shim_program():
    exe = os.path.join(ROOT, "sfm-toy-library_b200", "host", "build", "test_shim")
    if not os.path.exists(exe):
        import __graft_entry__ as ge
        ge.build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "SHIM_TEST PASS" in r.stdout
