"""ctypes binding of the C-ABI in ``include/dynlab_b200.h`` (``_lib/libdynlab_b200.so``).

The library is the product: if it is missing this module raises at import with the build
command — nothing in dynlab_b200 falls back to a CPU or PyTorch implementation.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DLB_LIB_PATH") or os.path.join(_HERE, "_lib", "libdynlab_b200.so")

DLB_OK, DLB_ERR_ARG, DLB_ERR_CUDA, DLB_ERR_TOO_SMALL, DLB_ERR_WORKSPACE = 0, -1, -2, -3, -4
EULER_S_MIN, EULER_S_MAX, EULER_RATE, EULER_RATIO = 0, 1, 2, 3
XI_NONE, XI_LAPACK, XI_FORCED = 0, 1, 2
CNT_NFEV, CNT_ACCEPTED, CNT_REJECTED, CNT_FAILED, CNT_SEEDS, CNT_WORDS = 0, 1, 2, 3, 4, 8

_vp, _i, _i64, _d, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_size_t
_dp = C.POINTER(C.c_double)

# name -> (restype, argtypes); must list every symbol include/dynlab_b200.h declares
SIGNATURES = {
    "dlb_abi_version": (_i, []),
    "dlb_last_error": (C.c_char_p, []),
    "dlb_flow_dim": (_i, [_i]),
    "dlb_flow_nparams": (_i, [_i]),
    "dlb_flow_name": (C.c_char_p, [_i]),
    "dlb_workspace_bytes": (_sz, [_i64, _i64]),
    "dlb_flow_eval_points": (_i, [_i, _dp, _i, _d, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "dlb_flow_eval_grid": (_i, [_i, _dp, _i, _d, _dp, _i64, _dp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "dlb_flowmap_rk45": (_i, [_i, _dp, _i, _dp, _i64, _dp, _i64, _i64, _d, _d, _d, _d, _d, _d,
                              _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dlb_flowmap_rk45_resident": (_i, [_i, _dp, _i, _vp, _i64, _vp, _i64, _i64, _d, _d, _d, _d, _d,
                                       _d, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _sz, _vp]),
    "dlb_rk45_endpoints": (_i, [_i, _dp, _i, _i, _vp, _i64, _d, _d, _d, _dp, _d, _d, _vp, _vp,
                                _vp, _vp, _vp, _sz, _vp]),
    "dlb_ftle_stencil": (_i, [_vp, _i64, _i64, _dp, _i64, _dp, _i64, _i, _d, _i64, _i64, _vp, _vp,
                              _i, _vp, _sz, _vp]),
    "dlb_euler_stencil": (_i, [_i, _vp, _vp, _i64, _i64, _dp, _i64, _dp, _i64, _i, _i64, _i64, _i,
                               _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "dlb_euler_stencil_flow": (_i, [_i, _i, _dp, _i, _d, _dp, _i64, _dp, _i64, _i, _i64, _i64, _i,
                                    _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "dlb_ridge_field": (_i, [_vp, _vp, _dp, _i64, _dp, _i64, _i, _i, _d, _vp, _vp, _vp, _vp, _vp,
                             _sz, _vp]),
    "dlb_contour_segments": (_i, [_vp, _vp, _dp, _i64, _dp, _i64, _d, _vp, _vp, _vp, _vp, _i64, _vp,
                                  _vp, _sz, _vp]),
    "dlb_link_segments": (_i, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "dlb_select_scratch_bytes": (_sz, []),
    "dlb_order_statistics": (_i, [_vp, _i64, C.POINTER(C.c_int64), _i, _vp, _vp, _vp, _sz, _vp]),
    "dlb_dfma_peak": (_i, [_i, _dp, _dp]),
    # multi-GPU helpers (csrc/peer.cu)
    "dlb_device_count": (_i, [C.POINTER(C.c_int)]),
    "dlb_peer_enable": (_i, [_i, _i]),
    "dlb_peer_alloc": (_i, [_sz, C.POINTER(_vp)]),
    "dlb_peer_free": (_i, [_vp]),
    "dlb_ipc_export": (_i, [_vp, C.c_char_p]),
    "dlb_ipc_open": (_i, [C.c_char_p, C.POINTER(_vp)]),
    "dlb_ipc_close": (_i, [_vp]),
    "dlb_peer_copy": (_i, [_vp, _vp, _sz, _vp]),
    "dlb_gather_rows": (_i, [_vp, _sz, C.POINTER(_vp), _i, C.POINTER(_vp), C.c_uint64, _vp]),
    "dlb_halo_exchange": (_i, [_vp, _sz, _i64, _i64, _vp, _vp, _vp, _vp, C.c_uint64, _vp]),
    "dlb_wait_flags": (_i, [_vp, _i, C.c_uint64, _d, _vp, _vp]),
    "dlb_ftle_stencil_gather": (_i, [_vp, _i64, _i64, _dp, _i64, _dp, _i64, _i, _d, _i64, _i64,
                                     _vp, C.POINTER(_vp), _i, C.POINTER(_vp), _i, C.c_uint64,
                                     _vp, _sz, _vp]),
}

MAX_PEERS = 16
IPC_HANDLE_BYTES = 64


class NativeLibraryMissing(ImportError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found. dynlab_b200 has no CPU fallback: build the CUDA extension "
            f"first (python -c 'import __graft_entry__ as g; g.build()' or "
            f"dynlab_b200/csrc/build.sh).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class DynlabCudaError(RuntimeError):
    pass


def check(rc: int) -> None:
    """Map C-ABI error codes onto the exceptions the reference / numpy raise."""
    if rc == DLB_OK:
        return
    msg = (lib.dlb_last_error() or b"").decode()
    if rc in (DLB_ERR_ARG, DLB_ERR_TOO_SMALL):
        raise ValueError(msg)  # numpy's np.gradient errors are ValueError too
    raise DynlabCudaError(f"dynlab_b200 native error {rc}: {msg}")


def darray(values):
    """Host double array for h_* arguments (kept alive by the caller)."""
    values = list(values)
    return (C.c_double * max(len(values), 1))(*values)


def dptr(np_array):
    return np_array.ctypes.data_as(_dp)
