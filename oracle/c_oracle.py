"""ctypes front end of oracle/icp_oracle.c (TEST INFRASTRUCTURE, see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libicp_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "icp_oracle.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        _lib = C.CDLL(_SO)
        _lib.icp_oracle.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def nn_bruteforce(src, tgt):
    src = np.ascontiguousarray(src, dtype=np.float64)
    tgt = np.ascontiguousarray(tgt, dtype=np.float64)
    idx = np.empty(len(src), dtype=np.int32)
    d2 = np.empty(len(src), dtype=np.float64)
    lib().nn_bruteforce(_p(src), C.c_int(len(src)), _p(tgt), C.c_int(len(tgt)), _p(idx), _p(d2))
    return idx, d2


def icp(A, B, max_iterations=20, tolerance=1e-5, init_pose=None, max_corr_dist=None, history=False):
    A = np.ascontiguousarray(A, dtype=np.float64)
    B = np.ascontiguousarray(B, dtype=np.float64)
    n = len(A)
    pt, pl = np.zeros(6), np.zeros(6)
    err, rmse, inl = C.c_double(), C.c_double(), C.c_int32()
    hist = np.full((max_iterations, n), -1, dtype=np.int32) if history else None
    src = np.zeros((n, 2))
    ip = None if init_pose is None else np.ascontiguousarray(init_pose, dtype=np.float64)
    it = lib().icp_oracle(_p(A), C.c_int(n), _p(B), C.c_int(len(B)), C.c_int(max_iterations),
                          C.c_double(tolerance), _p(ip),
                          C.c_double(0.0 if max_corr_dist is None else max_corr_dist),
                          _p(pt), _p(pl), C.byref(err), C.byref(rmse), C.byref(inl), _p(hist), _p(src))
    return dict(iterations=it, pose_total=pt, pose_last=pl, error=err.value, rmse=rmse.value,
                inliers=inl.value, history=hist, src=src)
