"""ctypes binding of libpcfusion.so (include/pcfusion.h).  Mirrors the reference's OccupancyGrid / node surface:
start/stop/reset/process (node.cpp:351-440), addPoints -> push_frame (OG.hpp:185), updateThicknessVectors ->
update (OG.hpp:311), downloadData -> extract/process (OG.hpp:456), clearVoxels -> clear (OG.hpp:167)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    return os.path.join(_HERE, "libpcfusion.so")


class PcfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pcfusion error {code}: {msg}")
        self.code = code


class _Config(C.Structure):
    _fields_ = [("box", C.c_double * 6), ("res", C.c_float * 3), ("clip_zmin", C.c_double), ("clip_zmax", C.c_double),
                ("k_neighbourhood", C.c_int32), ("walk_k", C.c_int32), ("min_neighbours", C.c_int32),
                ("cylinder_radius", C.c_double), ("ball_radius", C.c_double), ("device", C.c_int32),
                ("max_frames", C.c_uint32), ("log_capacity_hint", C.c_uint64), ("stage_threads", C.c_int32), ("stage_raw_lanes", C.c_int32)]


class _Result(C.Structure):
    _fields_ = [("n", C.c_uint64), ("hash", C.c_void_p), ("centroid", C.c_void_p), ("normal", C.c_void_p),
                ("sd", C.c_void_p), ("mean_dist", C.c_void_p), ("sd_dist", C.c_void_p), ("count", C.c_void_p)]


class _State(C.Structure):
    _fields_ = [("n", C.c_uint64), ("hash", C.c_void_p), ("buffer_len", C.c_void_p), ("normal_found", C.c_void_p),
                ("count", C.c_void_p), ("normal", C.c_void_p), ("viewpoint", C.c_void_p)]


class _Stats(C.Structure):
    _fields_ = [("frames_pushed", C.c_uint64), ("points_offered", C.c_uint64), ("points_kept", C.c_uint64),
                ("occupied_voxels", C.c_uint64), ("normals_found", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("update_passes", C.c_uint32), ("staged_dropped", C.c_uint64)]


ABI_SYMBOLS = [
    "pcf_default_config", "pcf_create", "pcf_destroy", "pcf_last_error", "pcf_dims", "pcf_start", "pcf_stop", "pcf_reset",
    "pcf_push_frame", "pcf_push_pointcloud2", "pcf_add_points", "pcf_submit_frame", "pcf_submit_pointcloud2", "pcf_drain", "pcf_stage_frame", "pcf_staged_count", "pcf_wait_staged", "pcf_host_alloc", "pcf_host_free", "pcf_upload_ticket", "pcf_wait_upload", "pcf_push_frames_device", "pcf_sync", "pcf_count_kept", "pcf_update", "pcf_extract", "pcf_process", "pcf_write_result",
    "pcf_extract_hq", "pcf_clear", "pcf_dump_state", "pcf_reserve_process", "pcf_get_stats", "pcf_reset_stats", "pcf_last_timings", "pcf_stream",
    "pcf_grid_buffer", "pcf_viewpoint_table", "pcf_log_compact", "pcf_log_replace", "pcf_set_slab", "pcf_plane_counts", "pcf_plane_point_counts", "pcf_exchange_counts", "pcf_exchange_scatter", "pcf_exchange_hist", "pcf_exchange_plan", "pcf_exchange_scatter_async", "pcf_round_export", "pcf_round_install", "pcf_update_local", "pcf_update_commit", "pcf_recv_buffer", "pcf_ipc_export", "pcf_ipc_open",
    "pcf_ipc_close_all", "pcf_install_records", "pcf_get_viewpoints", "pcf_set_viewpoints", "pcf_enable_peer_access", "pcf_kat_transform_voxel",
    "pcf_kat_normal", "pcf_kat_score", "pcf_kat_format_float", "pcf_kat_clip_pack", "pcf_kat_div",
]

_lib = None


def load_library():
    """Load libpcfusion.so.  Raises if it has not been built: the product path has no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise PcfError(-5, f"{p} not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(p)
    vp = C.c_void_p
    lib.pcf_default_config.argtypes = [C.POINTER(_Config)]
    lib.pcf_default_config.restype = None
    lib.pcf_create.argtypes = [C.POINTER(_Config), C.POINTER(vp)]
    lib.pcf_destroy.argtypes = [vp]
    lib.pcf_destroy.restype = None
    lib.pcf_last_error.argtypes = [vp]
    lib.pcf_last_error.restype = C.c_char_p
    lib.pcf_dims.argtypes = [vp, C.POINTER(C.c_int32 * 3)]
    for name in ["pcf_start", "pcf_stop", "pcf_reset", "pcf_sync", "pcf_update", "pcf_clear", "pcf_reset_stats", "pcf_drain"]:
        getattr(lib, name).argtypes = [vp]
    lib.pcf_push_frame.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, C.c_uint32]
    lib.pcf_push_pointcloud2.argtypes = [vp, vp] + [C.c_uint32] * 7 + [vp, C.c_uint32]
    lib.pcf_add_points.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, C.c_uint32]
    lib.pcf_submit_frame.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, C.c_uint32]
    lib.pcf_submit_pointcloud2.argtypes = [vp, vp] + [C.c_uint32] * 7 + [vp, C.c_uint32]
    lib.pcf_stage_frame.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, C.POINTER(C.c_uint32)]
    lib.pcf_staged_count.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.pcf_wait_staged.argtypes = [vp, C.c_uint64]
    lib.pcf_host_alloc.argtypes = [C.c_size_t]
    lib.pcf_host_alloc.restype = vp
    lib.pcf_host_free.argtypes = [vp]
    lib.pcf_host_free.restype = None
    lib.pcf_upload_ticket.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.pcf_wait_upload.argtypes = [vp, C.c_uint64]
    lib.pcf_push_frames_device.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, C.c_uint32]
    lib.pcf_count_kept.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.pcf_extract.argtypes = [vp, C.POINTER(_Result)]
    lib.pcf_extract_hq.argtypes = [vp, C.c_double, C.POINTER(_Result)]
    lib.pcf_process.argtypes = [vp, C.c_char_p, C.c_char_p]
    lib.pcf_write_result.argtypes = [C.POINTER(_Result), C.c_char_p, C.c_char_p]
    lib.pcf_dump_state.argtypes = [vp, C.POINTER(_State)]
    lib.pcf_reserve_process.argtypes = [vp, C.c_uint64, C.c_uint64]
    lib.pcf_get_stats.argtypes = [vp, C.POINTER(_Stats)]
    lib.pcf_last_timings.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.pcf_stream.argtypes = [vp]
    lib.pcf_stream.restype = vp
    lib.pcf_grid_buffer.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint64)]
    lib.pcf_viewpoint_table.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint32)]
    lib.pcf_log_compact.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint64)]
    lib.pcf_log_replace.argtypes = [vp, vp, C.c_uint64]
    lib.pcf_set_slab.argtypes = [vp, C.c_int32, C.c_int32]
    lib.pcf_plane_counts.argtypes = [vp, vp]
    lib.pcf_plane_point_counts.argtypes = [vp, vp]
    lib.pcf_exchange_counts.argtypes = [vp, vp, C.c_int32, vp]
    lib.pcf_exchange_scatter.argtypes = [vp, vp, vp]
    lib.pcf_exchange_scatter_async.argtypes = [vp, vp, vp]
    lib.pcf_exchange_hist.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint32)]
    lib.pcf_round_export.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint64)]
    lib.pcf_round_install.argtypes = [vp, vp, C.c_uint64]
    lib.pcf_update_local.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_uint32)]
    lib.pcf_update_commit.argtypes = [vp, vp, vp, C.c_uint32]
    lib.pcf_exchange_plan.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(vp)]
    lib.pcf_recv_buffer.argtypes = [vp, C.c_uint64, C.POINTER(vp)]
    lib.pcf_ipc_export.argtypes = [vp, vp]
    lib.pcf_ipc_open.argtypes = [vp, vp, C.POINTER(vp)]
    lib.pcf_ipc_close_all.argtypes = [vp]
    lib.pcf_install_records.argtypes = [vp, vp, C.c_uint64]
    lib.pcf_get_viewpoints.argtypes = [vp, vp, C.c_uint32, C.c_uint32]
    lib.pcf_set_viewpoints.argtypes = [vp, vp, C.c_uint32, C.c_uint32]
    lib.pcf_enable_peer_access.argtypes = [vp, C.c_int32]
    lib.pcf_kat_transform_voxel.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, vp, vp, vp]
    lib.pcf_kat_format_float.argtypes = [C.c_float, C.c_int, C.c_char_p]
    lib.pcf_kat_clip_pack.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_float, C.c_float, C.c_int32, vp, C.POINTER(C.c_uint32)]
    lib.pcf_kat_normal.argtypes = [vp, vp, C.c_uint32, vp]
    lib.pcf_kat_div.argtypes = [vp, vp, vp, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.pcf_kat_score.argtypes = [vp, vp, C.c_uint32, vp, vp, vp, vp, vp, vp, vp]
    _lib = lib
    return lib


def _np_from(ptr, n, dtype, cols=None):
    if n == 0:
        return np.zeros((0,) if cols is None else (0, cols), dtype)
    count = n * (cols or 1)
    buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
    a = np.frombuffer(buf, dtype=dtype, count=count).copy()
    return a if cols is None else a.reshape(n, cols)


class Result:
    """Extraction output, x-major order (downloadData, OG.hpp:463-480)."""

    def __init__(self, r: _Result):
        n = int(r.n)
        self.hash = _np_from(r.hash, n, np.uint64)
        self.centroid = _np_from(r.centroid, n, np.float32, 3)
        self.normal = _np_from(r.normal, n, np.float32, 3)
        self.sd = _np_from(r.sd, n, np.float32, 3)
        self.mean_dist = _np_from(r.mean_dist, n, np.float32)
        self.sd_dist = _np_from(r.sd_dist, n, np.float32)
        self.count = _np_from(r.count, n, np.int32)

    def __len__(self):
        return len(self.hash)


class State:
    def __init__(self, s: _State):
        n = int(s.n)
        self.hash = _np_from(s.hash, n, np.uint64)
        self.buffer_len = _np_from(s.buffer_len, n, np.int32)
        self.normal_found = _np_from(s.normal_found, n, np.uint8)
        self.count = _np_from(s.count, n, np.int32)
        self.normal = _np_from(s.normal, n, np.float32, 3)
        self.viewpoint = _np_from(s.viewpoint, n, np.float32, 3)


def _ptr(a):
    """Address of a numpy array, a torch tensor (host or device) or a raw integer address."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()   # torch tensor


def kat_clip_pack(data, rows, cols, point_step, row_step, x_offset, lo, hi, isa=-1, misalign=False):
    """The staging pool's clip-and-pack on a raw message buffer (host only: works without a GPU).  Returns (xyz [m, 3], isa used)."""
    lib = load_library()
    data = np.ascontiguousarray(data, np.uint8)
    raw = np.empty(3 * rows * cols + 16 + 32, np.float32)
    shift = (-raw.ctypes.data // 4) % 16 + (1 if misalign else 0)    # 64-byte aligned output (the non-temporal path) or deliberately not
    out = raw[shift:]
    m = C.c_uint32()
    used = lib.pcf_kat_clip_pack(data.ctypes.data, rows, cols, point_step, row_step, x_offset, lo, hi, isa, out.ctypes.data, C.byref(m))
    if used < 0:
        raise PcfError(used, "pcf_kat_clip_pack: bad layout")
    return out[:3 * m.value].reshape(-1, 3).copy(), used


class Fusion:
    """One fusion context on one GPU (= the reference's PointcloudFusion + OccupancyGrid pair)."""

    def __init__(self, box, res, clip_zmin=0.28, clip_zmax=0.6, device=0, max_frames=1 << 16, log_capacity_hint=0,
                 walk_k=3, min_neighbours=20, started=True, stage_threads=0, stage_raw_lanes=0):
        self.lib = load_library()
        cfg = _Config()
        self.lib.pcf_default_config(C.byref(cfg))
        cfg.box[:] = [float(b) for b in box]
        r = np.broadcast_to(np.asarray(res, np.float32), (3,))
        cfg.res[:] = [float(x) for x in r]
        cfg.clip_zmin, cfg.clip_zmax = clip_zmin, clip_zmax
        cfg.device, cfg.max_frames, cfg.log_capacity_hint = device, max_frames, log_capacity_hint
        cfg.walk_k, cfg.min_neighbours = walk_k, min_neighbours
        cfg.stage_threads, cfg.stage_raw_lanes = stage_threads, stage_raw_lanes
        h = C.c_void_p()
        rc = self.lib.pcf_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise PcfError(rc, self.lib.pcf_last_error(None).decode())
        self.h = h
        self.device_index = device
        d = (C.c_int32 * 3)()
        self.lib.pcf_dims(self.h, C.byref(d))
        self.dims = tuple(d)
        if started:
            self.start()

    def close(self):
        if getattr(self, "h", None):
            self.lib.pcf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise PcfError(rc, self.lib.pcf_last_error(self.h).decode())
        return rc

    def start(self):
        self._ck(self.lib.pcf_start(self.h))

    def stop(self):
        self._ck(self.lib.pcf_stop(self.h))

    def reset(self):
        self._ck(self.lib.pcf_reset(self.h))

    def push_frame(self, pts, pose, frame_idx):
        """pts: host float32 [n, stride>=3] (numpy, or pinned torch tensor); pose: 4x4 float64 fusion<-camera."""
        pose = np.ascontiguousarray(pose, np.float64).reshape(16)
        n, stride = pts.shape
        return self._ck(self.lib.pcf_push_frame(self.h, _ptr(pts), n, stride, pose.ctypes.data, frame_idx))

    def push_pointcloud2(self, data, width, height, point_step, row_step, offsets, pose, frame_idx):
        """data: uint8 numpy array = sensor_msgs/PointCloud2.data; offsets = (x, y, z) field offsets in bytes."""
        pose = np.ascontiguousarray(pose, np.float64).reshape(16)
        return self._ck(self.lib.pcf_push_pointcloud2(self.h, _ptr(data), width, height, point_step, row_step, offsets[0], offsets[1],
                                                      offsets[2], pose.ctypes.data, frame_idx))

    def submit_frame(self, pts, pose, frame_idx):
        """Asynchronous push through the host staging pool (clip-and-pack, node.cpp:218-263): `pts` (host memory, pageable or
        pinned) and `pose` (float64[16], C-contiguous) must stay alive until drain() / sync() / count_kept() / update()."""
        n, stride = pts.shape
        return self._ck(self.lib.pcf_submit_frame(self.h, _ptr(pts), n, stride, _ptr(pose), frame_idx))

    def submit_pointcloud2(self, data, width, height, point_step, row_step, offsets, pose, frame_idx):
        pose = np.ascontiguousarray(pose, np.float64).reshape(16)
        return self._ck(self.lib.pcf_submit_pointcloud2(self.h, _ptr(data), width, height, point_step, row_step, offsets[0], offsets[1],
                                                        offsets[2], pose.ctypes.data, frame_idx))

    def drain(self):
        self._ck(self.lib.pcf_drain(self.h))

    def wait_staged(self, n):
        """Block until `n` submitted clouds have been handed to the GPU (or nothing is pending)."""
        self._ck(self.lib.pcf_wait_staged(self.h, int(n)))

    def staged_count(self) -> int:
        n = C.c_uint64()
        self._ck(self.lib.pcf_staged_count(self.h, C.byref(n)))
        return int(n.value)

    def stage_frame(self, pts, out=None):
        """pcf_stage_frame: clip-and-pack on the calling thread.  Returns (staged float32 [m, 3], m)."""
        n, stride = pts.shape
        if out is None:
            out = np.empty((n + 4, 3), np.float32)
        m = C.c_uint32()
        self._ck(self.lib.pcf_stage_frame(self.h, _ptr(pts), n, stride, _ptr(out), C.byref(m)))
        return out[:m.value], int(m.value)

    def add_points(self, pts_world, viewpoint, frame_idx):
        """OccupancyGrid::addPoints (OG.hpp:185): cloud already in the fusion frame + explicit viewpoint."""
        vp3 = np.ascontiguousarray(viewpoint, np.float32).reshape(3)
        n, stride = pts_world.shape
        return self._ck(self.lib.pcf_add_points(self.h, _ptr(pts_world), n, stride, vp3.ctypes.data, frame_idx))

    def push_frames_device(self, pts_dev, n_frames, n_per_frame, stride, poses, first_frame_idx):
        poses = np.ascontiguousarray(poses, np.float64).reshape(n_frames, 16)
        return self._ck(self.lib.pcf_push_frames_device(self.h, _ptr(pts_dev), n_frames, n_per_frame, stride,
                                                        poses.ctypes.data, first_frame_idx))

    def sync(self):
        self._ck(self.lib.pcf_sync(self.h))

    def count_kept(self) -> int:
        k = C.c_uint64()
        self._ck(self.lib.pcf_count_kept(self.h, C.byref(k)))
        return int(k.value)

    def update(self):
        self._ck(self.lib.pcf_update(self.h))

    def extract(self, hq_threshold=None) -> Result:
        r = _Result()
        if hq_threshold is None:
            self._ck(self.lib.pcf_extract(self.h, C.byref(r)))
        else:
            self._ck(self.lib.pcf_extract_hq(self.h, float(hq_threshold), C.byref(r)))
        return Result(r)

    def extract_raw(self) -> int:
        """Extraction without copying the result into numpy (for timing); returns the number of voxels."""
        r = _Result()
        self._ck(self.lib.pcf_extract(self.h, C.byref(r)))
        return int(r.n)

    def process(self, cloud_path=None, meta_path=None):
        self._ck(self.lib.pcf_process(self.h, cloud_path.encode() if cloud_path else None,
                                      meta_path.encode() if meta_path else None))

    def clear(self):
        self._ck(self.lib.pcf_clear(self.h))

    def reserve_process(self, max_points, max_voxels):
        """Pre-size the process() scratch (first-call allocations)."""
        self._ck(self.lib.pcf_reserve_process(self.h, int(max_points), int(max_voxels)))

    def state(self) -> State:
        s = _State()
        self._ck(self.lib.pcf_dump_state(self.h, C.byref(s)))
        return State(s)

    def stats(self) -> dict:
        s = _Stats()
        self._ck(self.lib.pcf_get_stats(self.h, C.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in _Stats._fields_}

    def reset_stats(self):
        self._ck(self.lib.pcf_reset_stats(self.h))

    def timings(self):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        self._ck(self.lib.pcf_last_timings(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"update_ms": a.value, "extract_device_ms": b.value, "extract_d2h_ms": c.value}

    @property
    def stream(self):
        return self.lib.pcf_stream(self.h)

    def grid_buffer(self):
        p, n = C.c_void_p(), C.c_uint64()
        self._ck(self.lib.pcf_grid_buffer(self.h, C.byref(p), C.byref(n)))
        return p.value, int(n.value)

    def viewpoint_table(self):
        p, n = C.c_void_p(), C.c_uint32()
        self._ck(self.lib.pcf_viewpoint_table(self.h, C.byref(p), C.byref(n)))
        return p.value, int(n.value)

    def log_compact(self):
        p, n = C.c_void_p(), C.c_uint64()
        self._ck(self.lib.pcf_log_compact(self.h, C.byref(p), C.byref(n)))
        return p.value, int(n.value)

    def log_replace(self, log_dev, n_points):
        self._ck(self.lib.pcf_log_replace(self.h, _ptr(log_dev), n_points))

    def set_slab(self, x_lo, x_hi):
        self._ck(self.lib.pcf_set_slab(self.h, x_lo, x_hi))

    def plane_counts(self):
        out = np.zeros(self.dims[0] + 2, np.uint32)
        self._ck(self.lib.pcf_plane_counts(self.h, out.ctypes.data))
        return out

    # ---- exchange v2 (slab-routed records) ----
    def plane_point_counts(self):
        out = np.zeros(self.dims[0] + 1, np.uint32)
        self._ck(self.lib.pcf_plane_point_counts(self.h, out.ctypes.data))
        return out

    def exchange_counts(self, bounds):
        b = np.ascontiguousarray(bounds, np.int32)
        out = np.zeros(len(b) - 1, np.uint64)
        self._ck(self.lib.pcf_exchange_counts(self.h, b.ctypes.data, len(b) - 1, out.ctypes.data))
        return out

    def exchange_scatter(self, dst_ptrs, dst_offsets):
        p = np.ascontiguousarray(dst_ptrs, np.uint64)
        o = np.ascontiguousarray(dst_offsets, np.uint64)
        self._ck(self.lib.pcf_exchange_scatter(self.h, p.ctypes.data, o.ctypes.data))

    # ---- interleaved schedules across ranks (replicated state) ----
    def round_export(self):
        p, n = C.c_void_p(), C.c_uint64()
        self._ck(self.lib.pcf_round_export(self.h, C.byref(p), C.byref(n)))
        return p.value, int(n.value)

    def round_install(self, records_dev, n):
        self._ck(self.lib.pcf_round_install(self.h, _ptr(records_dev), int(n)))

    def update_local(self):
        pc, pn, n = C.c_void_p(), C.c_void_p(), C.c_uint32()
        self._ck(self.lib.pcf_update_local(self.h, C.byref(pc), C.byref(pn), C.byref(n)))
        return pc.value, pn.value, int(n.value)

    def update_commit(self, cells_dev, normals_dev, n):
        self._ck(self.lib.pcf_update_commit(self.h, _ptr(cells_dev), _ptr(normals_dev), int(n)))

    def exchange_hist(self):
        p, n = C.c_void_p(), C.c_uint32()
        self._ck(self.lib.pcf_exchange_hist(self.h, C.byref(p), C.byref(n)))
        return p.value, int(n.value)

    def exchange_plan(self, n_ranks, self_rank) -> int:
        p = C.c_void_p()
        self._ck(self.lib.pcf_exchange_plan(self.h, n_ranks, self_rank, C.byref(p)))
        return p.value

    def exchange_scatter_async(self, dst_ptrs, dst_offsets):
        p = np.ascontiguousarray(dst_ptrs, np.uint64)
        o = np.ascontiguousarray(dst_offsets, np.uint64)
        self._ck(self.lib.pcf_exchange_scatter_async(self.h, p.ctypes.data, o.ctypes.data))

    def recv_buffer(self, n_records) -> int:
        p = C.c_void_p()
        self._ck(self.lib.pcf_recv_buffer(self.h, int(n_records), C.byref(p)))
        return p.value

    def ipc_export(self) -> bytes:
        buf = (C.c_char * 64)()
        self._ck(self.lib.pcf_ipc_export(self.h, buf))
        return bytes(buf)

    def ipc_open(self, handle: bytes) -> int:
        p = C.c_void_p()
        buf = (C.c_char * 64).from_buffer_copy(handle)
        self._ck(self.lib.pcf_ipc_open(self.h, buf, C.byref(p)))
        return p.value

    def ipc_close_all(self):
        self._ck(self.lib.pcf_ipc_close_all(self.h))

    def install_records(self, records_dev, n_records):
        self._ck(self.lib.pcf_install_records(self.h, _ptr(records_dev), int(n_records)))

    # ---- known-answer hooks ----
    def kat_transform_voxel(self, pts, pose):
        pts = np.ascontiguousarray(pts, np.float32)
        pose = np.ascontiguousarray(pose, np.float64).reshape(16)
        n, stride = pts.shape
        w, ijk, kept = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.int32), np.zeros(n, np.uint8)
        self._ck(self.lib.pcf_kat_transform_voxel(self.h, pts.ctypes.data, n, stride, pose.ctypes.data, w.ctypes.data,
                                                  ijk.ctypes.data, kept.ctypes.data))
        return w, ijk, kept

    def kat_div(self, x, c):
        """(mismatches, index of one) of the scoring kernel's shared-reciprocal division vs x / c on the device."""
        x, c = np.ascontiguousarray(x, np.float32), np.ascontiguousarray(c, np.float32)
        m, bad = C.c_uint32(), C.c_uint32()
        self._ck(self.lib.pcf_kat_div(self.h, x.ctypes.data, c.ctypes.data, len(x), C.byref(m), C.byref(bad)))
        return int(m.value), int(bad.value)

    def kat_normal(self, xyz):
        xyz = np.ascontiguousarray(xyz, np.float32)
        out = np.zeros(3, np.float32)
        self._ck(self.lib.pcf_kat_normal(self.h, xyz.ctypes.data, xyz.shape[0], out.ctypes.data))
        return out

    def kat_score(self, xyz, axis_pt, nrm):
        xyz = np.ascontiguousarray(xyz, np.float32)
        a, n = np.ascontiguousarray(axis_pt, np.float32), np.ascontiguousarray(nrm, np.float32)
        c, sd = np.zeros(3, np.float32), np.zeros(3, np.float32)
        md, sdd, cnt = np.zeros(1, np.float32), np.zeros(1, np.float32), np.zeros(1, np.int32)
        self._ck(self.lib.pcf_kat_score(self.h, xyz.ctypes.data, xyz.shape[0], a.ctypes.data, n.ctypes.data, c.ctypes.data,
                                        sd.ctypes.data, md.ctypes.data, sdd.ctypes.data, cnt.ctypes.data))
        return c, sd, float(md[0]), float(sdd[0]), int(cnt[0])
