"""ctypes binding of oracle/extract_c.c (TEST INFRASTRUCTURE): the C / OpenMP restatement of the extract
half's arithmetic - A1 density and the A5 surface-field marcher - fast enough to check full-size (128^3,
50 cameras) blocks and to serve as the timed CPU baseline.  Build with ``make -C oracle``."""
import ctypes as C
import os
import subprocess

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_c", "liboracle_extract.so")
_lib = None


def load(build=True):
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB) and build:
        subprocess.run(["make", "-s", "-C", HERE], check=True)
    lib = C.CDLL(LIB)
    fp, u8p = C.POINTER(C.c_float), C.POINTER(C.c_uint8)
    lib.orc_density.restype = None
    lib.orc_density.argtypes = [fp, fp, fp, fp, fp, C.c_int, fp]
    lib.orc_surface_mask.restype = None
    lib.orc_surface_mask.argtypes = [fp, fp, fp, fp, u8p, C.c_int, fp, fp, fp, C.c_int, fp, C.c_int, C.c_float,
                                     C.c_float, u8p, C.c_int, u8p, fp, C.POINTER(C.c_longlong)]
    _lib = lib
    return lib


def _f(t):
    a = np.ascontiguousarray(torch.as_tensor(t, dtype=torch.float32).detach().cpu().numpy(), dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def _u8(t):
    a = np.ascontiguousarray(torch.as_tensor(t).detach().cpu().to(torch.uint8).numpy(), dtype=np.uint8)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint8))


def query_density(x, aabb, table, w1, w2):
    """ngp.py:148-176 -> density [N] (fp32, sequential summation order)."""
    lib = load()
    xa, xp = _f(x.reshape(-1, 3))
    ta, tp = _f(table)
    w1a, w1p = _f(w1)
    w2a, w2p = _f(w2)
    aa, ap = _f(aabb)
    out = np.empty(xa.shape[0], dtype=np.float32)
    lib.orc_density(tp, w1p, w2p, ap, xp, xa.shape[0], out.ctypes.data_as(C.POINTER(C.c_float)))
    return torch.from_numpy(out)


def surface_mask(points, cam_origins, occ, res, roi_aabb, scene_aabb, step, cut_off, field, active=None,
                 all_rays=False):
    """sample_grid.py:245-318 -> (mask bool [N], best fp32 [N], number of density samples).
    ``field``: dict with aabb / table / w1 / w2 (oracle.make_goldens.make_field)."""
    lib = load()
    pa, pp = _f(points.reshape(-1, 3))
    ca, cp = _f(cam_origins.reshape(-1, 3))
    oa, op = _u8(occ.reshape(-1))
    ta, tp = _f(field["table"])
    w1a, w1p = _f(field["w1"])
    w2a, w2p = _f(field["w2"])
    aa, ap = _f(field["aabb"])
    ra, rp = _f(torch.as_tensor(roi_aabb, dtype=torch.float32))
    sa, sp = _f(torch.as_tensor(scene_aabb, dtype=torch.float32))
    n = pa.shape[0]
    act = (None, None) if active is None else _u8(active)
    out = np.zeros(n, dtype=np.uint8)
    best = np.zeros(n, dtype=np.float32)
    ns = C.c_longlong(0)
    lib.orc_surface_mask(tp, w1p, w2p, ap, op, int(res), rp, sp, pp, n, cp, ca.shape[0], float(step), float(cut_off),
                         act[1], int(bool(all_rays)), out.ctypes.data_as(C.POINTER(C.c_uint8)),
                         best.ctypes.data_as(C.POINTER(C.c_float)), C.byref(ns))
    return torch.from_numpy(out).bool(), torch.from_numpy(best), int(ns.value)
