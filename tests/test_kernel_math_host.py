"""The per-observation arithmetic of the CUDA kernels (csrc/ba_math.cuh, __host__ __device__) compiled for the CPU and
checked against the oracle's dual-number evaluation of SimpleReprojectionError (SfMBundleAdjustmentUtils.cpp:58-97)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hm():
    so = os.path.join(ROOT, "tests", "_build", "libhost_math.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", os.path.join(ROOT, "tests", "host_math.cpp"), "-o", so],
                   check=True)
    return C.CDLL(so)


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_obs_eval_matches_oracle_jets(hm, oracle):
    rs = np.random.RandomState(3)
    for trial in range(200):
        cam = np.concatenate([rs.normal(0, 0.9, 3), rs.normal(0, 1, 3) + [0, 0, 8]])
        if trial % 10 == 0:
            cam[:3] = rs.normal(0, 1e-9, 3)                 # first-order branch
        X = rs.uniform(-2, 2, 3); f = 2500.0 * rs.uniform(0.9, 1.1); ox, oy = rs.normal(0, 100, 2)
        r = np.empty(2); Jc = np.empty(12); Jp = np.empty(6); Jf = np.empty(2); r2 = np.empty(2)
        hm.host_obs_eval(_d(cam), _d(X), C.c_double(f), C.c_double(ox), C.c_double(oy), _d(r), _d(Jc), _d(Jp), _d(Jf))
        hm.host_obs_residual(_d(cam), _d(X), C.c_double(f), C.c_double(ox), C.c_double(oy), _d(r2))
        ro, Jco, Jpo, Jfo = oracle.ba_residual_jacobian(cam, X, f, ox, oy, mode=0)
        np.testing.assert_allclose(r, ro, rtol=0, atol=1e-9); np.testing.assert_allclose(r2, ro, rtol=0, atol=1e-9)
        np.testing.assert_allclose(Jc.reshape(2, 6), Jco, rtol=1e-10, atol=1e-8)
        np.testing.assert_allclose(Jp.reshape(2, 3), Jpo, rtol=1e-10, atol=1e-8)
        np.testing.assert_allclose(Jf, Jfo, rtol=1e-12, atol=1e-12)


def test_chol3_inverse(hm):
    rs = np.random.RandomState(4)
    for _ in range(50):
        B = rs.normal(0, 1, (3, 3)); U = B @ B.T + 1e-3 * np.eye(3)
        u = np.array([U[0, 0], U[0, 1], U[0, 2], U[1, 1], U[1, 2], U[2, 2]]); m = np.empty(6)
        assert hm.host_chol3_inverse(_d(u), _d(m)) == 1
        M = np.array([[m[0], 0, 0], [m[1], m[2], 0], [m[3], m[4], m[5]]])
        np.testing.assert_allclose(M.T @ M, np.linalg.inv(U), rtol=1e-9, atol=1e-12)
    u = np.array([1.0, 2.0, 0.0, 1.0, 0.0, 1.0]); m = np.empty(6)
    assert hm.host_chol3_inverse(_d(u), _d(m)) == 0          # indefinite
