"""ctypes loader for the CPU checkers.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Two interchangeable back ends with one API:
  kind="oracle"       oracle/_build/liboracle.so   (restatement, occupancy_grid_oracle.cpp)
  kind="ref"          oracle/_ref/libogref.so      (reference OccupancyGrid.hpp, unmodified, stock work-list order)
  kind="ref_ordered"  oracle/_ref/libogref_ordered.so (same, D3 pin: ascending work-list order)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {
    "oracle": (os.path.join(_HERE, "_build", "liboracle.so"), "ora_"),
    "ref": (os.path.join(_HERE, "_ref", "libogref.so"), "ref_"),
    "ref_ordered": (os.path.join(_HERE, "_ref", "libogref_ordered.so"), "ref_"),
    "ref_timing": (os.path.join(_HERE, "_ref", "libogref_timing.so"), "ref_"),
    "ref_libmtrig": (os.path.join(_HERE, "_ref", "libogref_libmtrig.so"), "ref_"),
}
_loaded = {}


def build(force=False):
    """Compile the restatement (and oracle/_ref when /root/reference is mounted)."""
    if force or not os.path.exists(_LIBS["oracle"][0]) or (
            os.path.getmtime(_LIBS["oracle"][0]) < os.path.getmtime(os.path.join(_HERE, "occupancy_grid_oracle.cpp"))):
        subprocess.run(["make", "-s", "-C", _HERE, "all"], check=True)


def available(kind):
    return os.path.exists(_LIBS[kind][0])


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


def _lib(kind):
    if kind in _loaded:
        return _loaded[kind]
    path, pre = _LIBS[kind]
    if kind == "oracle":
        build()
    lib = C.CDLL(path)
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
    sig = {
        "create": (vp, [vp, vp, dbl, dbl, i32]),
        "destroy": (None, [vp]),
        "dims": (None, [vp, vp]),
        "add_frame": (i64, [vp, vp, i32, i64, vp]),
        "add_points_world": (i64, [vp, vp, i32, i64, vp]),
        "update": (None, [vp]),
        "download": (i64, [vp]),
        "get_result": (None, [vp] * 8),
        "state_size": (i64, [vp]),
        "get_state": (None, [vp] * 7),
    }
    if kind == "oracle":
        sig.update({
            "clear": (None, [vp]),
            "write_csv": (i32, [vp, C.c_char_p]),
            "write_pcd": (i32, [vp, C.c_char_p]),
            "kat_transform": (None, [vp, vp, i32, i64, vp]),
            "kat_voxel": (None, [vp, vp, i64, vp, vp]),
            "kat_center": (None, [vp, vp, i64, vp]),
            "kat_project": (None, [vp, vp, vp, i64, vp, vp]),
            "kat_normal": (None, [vp, i64, vp, vp, vp]),
            "kat_eigen33": (None, [vp, vp, vp]),
            "kat_score": (None, [vp, i64, vp, vp, vp, vp, vp, vp, vp]),
        })
    else:
        sig["download_data"] = (i32, [vp, C.c_char_p, C.c_char_p])
    fns = {}
    for name, (res, args) in sig.items():
        f = getattr(lib, pre + name)
        f.restype, f.argtypes = res, args
        fns[name] = f
    _loaded[kind] = fns
    return fns


class Result:
    """Extraction output in x-major order (OG.hpp:463-480)."""

    def __init__(self, n):
        self.hash = np.zeros(n, np.uint64)
        self.centroid = np.zeros((n, 3), np.float32)
        self.normal = np.zeros((n, 3), np.float32)
        self.sd = np.zeros((n, 3), np.float32)
        self.mean_dist = np.zeros(n, np.float32)
        self.sd_dist = np.zeros(n, np.float32)
        self.count = np.zeros(n, np.int32)

    def __len__(self):
        return len(self.hash)


class State:
    def __init__(self, n):
        self.hash = np.zeros(n, np.uint64)
        self.buffer_len = np.zeros(n, np.int32)
        self.normal_found = np.zeros(n, np.uint8)
        self.count = np.zeros(n, np.int32)
        self.normal = np.zeros((n, 3), np.float32)
        self.viewpoint = np.zeros((n, 3), np.float32)


class OracleGrid:
    """Event-driven CPU grid: add_frame / update / download in any interleaving."""

    def __init__(self, box, res, clip_zmin=0.28, clip_zmax=0.6, reserve_hint=0, kind="oracle"):
        self.kind = kind
        self.f = _lib(kind)
        self._args = (box, res, clip_zmin, clip_zmax, reserve_hint)
        self._make()

    def _make(self):
        box, res, zmin, zmax, rsv = self._args
        b = np.ascontiguousarray(box, np.float64)
        r = np.ascontiguousarray(np.broadcast_to(np.asarray(res, np.float32), (3,)), np.float32)
        self.h = self.f["create"](_fp(b), _fp(r), zmin, zmax, rsv)
        d = np.zeros(3, np.int32)
        self.f["dims"](self.h, _fp(d))
        self.dims = tuple(int(x) for x in d)

    def close(self):
        if getattr(self, "h", None):
            self.f["destroy"](self.h)
            self.h = None

    def __del__(self):
        self.close()

    def add_frame(self, pts, pose):
        pts = np.ascontiguousarray(pts, np.float32)
        pose = np.ascontiguousarray(pose, np.float64).reshape(16)
        return self.f["add_frame"](self.h, _fp(pts), pts.shape[1], pts.shape[0], _fp(pose))

    def add_points_world(self, xyz, vp):
        xyz = np.ascontiguousarray(xyz, np.float32)
        vp = np.ascontiguousarray(vp, np.float32)
        return self.f["add_points_world"](self.h, _fp(xyz), xyz.shape[1], xyz.shape[0], _fp(vp))

    def update(self):
        self.f["update"](self.h)

    def download(self) -> Result:
        n = self.f["download"](self.h)
        r = Result(n)
        if n:
            self.f["get_result"](self.h, _fp(r.hash), _fp(r.centroid), _fp(r.normal), _fp(r.sd), _fp(r.mean_dist),
                                 _fp(r.sd_dist), _fp(r.count))
        return r

    def state(self) -> State:
        n = self.f["state_size"](self.h)
        s = State(n)
        if n:
            self.f["get_state"](self.h, _fp(s.hash), _fp(s.buffer_len), _fp(s.normal_found), _fp(s.count),
                                _fp(s.normal), _fp(s.viewpoint))
        return s

    def clear(self):
        if self.kind == "oracle":
            self.f["clear"](self.h)
        else:  # D5: the reference's clearVoxels leaves stale state; start from a fresh grid instead
            self.close()
            self._make()

    def write_files(self, cloud_path, meta_path):
        if self.kind == "oracle":
            self.f["write_pcd"](self.h, cloud_path.encode())
            self.f["write_csv"](self.h, meta_path.encode())
        else:
            self.f["download_data"](self.h, cloud_path.encode(), meta_path.encode())


# ---- known-answer helpers (oracle only) ---------------------------------------------------------
def kat_transform(pose, pts):
    f = _lib("oracle")
    pts = np.ascontiguousarray(pts, np.float32)
    out = np.zeros((pts.shape[0], 3), np.float32)
    f["kat_transform"](_fp(np.ascontiguousarray(pose, np.float64).reshape(16)), _fp(pts), pts.shape[1], pts.shape[0], _fp(out))
    return out


def kat_voxel(grid: OracleGrid, xyz):
    xyz = np.ascontiguousarray(xyz, np.float32)
    ijk = np.zeros((xyz.shape[0], 3), np.int32)
    valid = np.zeros(xyz.shape[0], np.uint8)
    grid.f["kat_voxel"](grid.h, _fp(xyz), xyz.shape[0], _fp(ijk), _fp(valid))
    return ijk, valid


def kat_center(grid: OracleGrid, ijk):
    ijk = np.ascontiguousarray(ijk, np.int32)
    out = np.zeros((ijk.shape[0], 3), np.float32)
    grid.f["kat_center"](grid.h, _fp(ijk), ijk.shape[0], _fp(out))
    return out


def kat_project(pt, axis_pt, nrm):
    f = _lib("oracle")
    pt, axis_pt, nrm = (np.ascontiguousarray(a, np.float32) for a in (pt, axis_pt, nrm))
    out = np.zeros_like(pt)
    dist = np.zeros(pt.shape[0], np.float64)
    f["kat_project"](_fp(pt), _fp(axis_pt), _fp(nrm), pt.shape[0], _fp(out), _fp(dist))
    return out, dist


def kat_normal(xyz):
    f = _lib("oracle")
    xyz = np.ascontiguousarray(xyz, np.float32)
    cov, nrm, ev = np.zeros(9, np.float32), np.zeros(3, np.float32), np.zeros(1, np.float32)
    f["kat_normal"](_fp(xyz), xyz.shape[0], _fp(cov), _fp(nrm), _fp(ev))
    return cov.reshape(3, 3), nrm, float(ev[0])


def kat_eigen33(cov):
    f = _lib("oracle")
    cov = np.ascontiguousarray(cov, np.float32).reshape(9)
    nrm, ev = np.zeros(3, np.float32), np.zeros(1, np.float32)
    f["kat_eigen33"](_fp(cov), _fp(nrm), _fp(ev))
    return nrm, float(ev[0])


def kat_score(pts, axis_pt, nrm):
    f = _lib("oracle")
    pts = np.ascontiguousarray(pts, np.float32)
    c, sd = np.zeros(3, np.float32), np.zeros(3, np.float32)
    md, sdd, cnt = np.zeros(1, np.float32), np.zeros(1, np.float32), np.zeros(1, np.int32)
    f["kat_score"](_fp(pts), pts.shape[0], _fp(np.ascontiguousarray(axis_pt, np.float32)),
                   _fp(np.ascontiguousarray(nrm, np.float32)), _fp(c), _fp(sd), _fp(md), _fp(sdd), _fp(cnt))
    return c, sd, float(md[0]), float(sdd[0]), int(cnt[0])
