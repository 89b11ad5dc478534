"""TEST INFRASTRUCTURE — ctypes binding of the plain-C oracle (oracle/c/dynlab_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdynlab_oracle.so")
_lib = None

FLOW_IDS = {n: i for i, n in enumerate([
    "double_gyre", "autonomous_double_gyre", "bickley_jet", "glider", "hills_vortex", "sigmoidal",
    "autonomous_duffing_oscillator", "duffing_oscillator", "pendulum", "hurricane",
    "van_der_pol_oscillator", "bead_on_a_rotating_hoop", "lotka_volterra", "abc", "lorenz"])}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "c", "dynlab_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "c")])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
    return _lib


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _params(p):
    a = np.zeros(16)
    a[:len(p)] = p
    return a


def flow_eval(name, params, t, Y):
    """Scalar RHS at one point; Y is a length-n sequence."""
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    out = np.empty_like(Y)
    rc = lib().orc_flow_eval(FLOW_IDS[name], _p(_params(params)), C.c_double(t), _p(Y), _p(out))
    assert rc == 0
    return out


def rk45_endpoints(name, params, seeds, t0, tf, rtol, atol, max_step=np.inf, first_step=0.0,
                   nthreads=0):
    seeds = np.ascontiguousarray(seeds, dtype=np.float64)
    m, n = seeds.shape
    atol = np.ascontiguousarray(np.broadcast_to(np.asarray(atol, dtype=np.float64), (n,)))
    out = np.empty_like(seeds)
    nfev = np.empty(m, dtype=np.int64)
    nacc = np.empty(m, dtype=np.int64)
    status = np.empty(m, dtype=np.int32)
    rc = lib().orc_rk45_endpoints(
        FLOW_IDS[name], _p(_params(params)), n, _p(seeds), C.c_int64(m), C.c_double(t0),
        C.c_double(tf), C.c_double(rtol), _p(atol), C.c_double(max_step), C.c_double(first_step),
        _p(out), _p(nfev, C.c_int64), _p(nacc, C.c_int64), _p(status, C.c_int32), nthreads)
    if rc:
        raise ValueError(f"orc_rk45_endpoints rc={rc}")
    return out, nfev, nacc, status


def flow_map(name, params, x, y, t0, tf, rtol, atol, max_step=np.inf, nthreads=0):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.empty((len(y), len(x), 2))
    nfev = np.empty((len(y), len(x)), dtype=np.int64)
    status = np.empty((len(y), len(x)), dtype=np.int32)
    rc = lib().orc_flow_map(
        FLOW_IDS[name], _p(_params(params)), _p(x), C.c_int64(len(x)), _p(y), C.c_int64(len(y)),
        C.c_double(t0), C.c_double(tf), C.c_double(rtol), C.c_double(atol), C.c_double(max_step),
        _p(out), _p(nfev, C.c_int64), _p(status, C.c_int32), nthreads)
    if rc:
        raise ValueError(f"orc_flow_map rc={rc}")
    return out, nfev, status


def gradient(f, y, x, edge_order):
    f = np.ascontiguousarray(f, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    dfdy, dfdx = np.empty_like(f), np.empty_like(f)
    rc = lib().orc_gradient(_p(f), C.c_int64(f.shape[0]), C.c_int64(f.shape[1]), _p(y), _p(x),
                            edge_order, _p(dfdy), _p(dfdx))
    if rc:
        raise ValueError(f"orc_gradient rc={rc}")
    return dfdy, dfdx


def ftle(fm, y, x, t0, tf, edge_order):
    fm = np.ascontiguousarray(fm, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty(fm.shape[:2])
    rc = lib().orc_ftle(_p(fm), C.c_int64(fm.shape[0]), C.c_int64(fm.shape[1]), _p(y), _p(x),
                        C.c_double(t0), C.c_double(tf), edge_order, _p(out))
    if rc:
        raise ValueError(f"orc_ftle rc={rc}")
    return out
