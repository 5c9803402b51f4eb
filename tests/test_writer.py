"""CPU: the PCD / CSV writer of libpcfusion (host code) is byte-identical to the oracle's writer, which follows
OG.hpp:460-462,478 (CSV) and PCL's savePCDFileASCII; when the prebuilt reference build is
present its downloadData() CSV (the reference's own formatting code) must match too."""
import ctypes as C
import importlib

import numpy as np
import pytest

import golden_util as G


def _write_with_lib(pcf, res, cloud, meta):
    b = importlib.import_module(pcf.__name__ + ".binding")
    lib = pcf.load_library()
    keep = [np.ascontiguousarray(getattr(res, f)) for f in G.RES_F]
    r = b._Result()
    r.n = len(keep[0])
    for name, arr in zip(G.RES_F, keep):
        setattr(r, name, arr.ctypes.data)
    assert lib.pcf_write_result(C.byref(r), cloud.encode(), meta.encode()) == 0


@pytest.mark.parametrize("name", G.fixtures())
def test_pcd_and_csv_bytes(pcf, oracle, tmp_path, name):
    fx = G.load(name)
    every = fx.schedules[0]
    og = oracle.OracleGrid(fx.box, fx.res, fx.clip[0], fx.clip[1])
    G.replay(og, fx, every)
    og.download()
    og.write_files(str(tmp_path / "o.pcd"), str(tmp_path / "o.csv"))
    _write_with_lib(pcf, fx.result[every], str(tmp_path / "l.pcd"), str(tmp_path / "l.csv"))
    assert (tmp_path / "l.pcd").read_bytes() == (tmp_path / "o.pcd").read_bytes()
    assert (tmp_path / "l.csv").read_bytes() == (tmp_path / "o.csv").read_bytes()
    head = (tmp_path / "l.pcd").read_text().splitlines()[:11]
    assert head[2] == "FIELDS x y z rgb normal_x normal_y normal_z curvature" and head[10] == "DATA ascii"
    assert head[6] == f"WIDTH {len(fx.result[every].hash)}"
    if oracle.available("ref_ordered"):
        rg = oracle.OracleGrid(fx.box, fx.res, fx.clip[0], fx.clip[1], kind="ref_ordered")
        G.replay(rg, fx, every)
        rg.write_files(str(tmp_path / "r.pcd"), str(tmp_path / "r.csv"))
        assert (tmp_path / "r.csv").read_bytes() == (tmp_path / "l.csv").read_bytes()
        assert (tmp_path / "r.pcd").read_bytes() == (tmp_path / "l.pcd").read_bytes()


def test_special_values(pcf, tmp_path):
    import types
    res = types.SimpleNamespace(hash=np.array([1, 2], np.uint64), centroid=np.array([[0, 0, 0], [np.nan, 1e-9, -123456.789]], np.float32),
                                normal=np.array([[0, 0, 1], [0.57735026, -0.57735026, 0.57735026]], np.float32),
                                sd=np.array([[0, 0, 0], [1e-12, 2.5e-7, 1]], np.float32), mean_dist=np.array([0, 0.00051234567], np.float32),
                                sd_dist=np.array([0, 1e-10], np.float32), count=np.array([0, 321], np.int32))
    _write_with_lib(pcf, res, str(tmp_path / "c.pcd"), str(tmp_path / "m.csv"))
    rows = (tmp_path / "m.csv").read_text().splitlines()
    assert rows[0].startswith("Id,sdx,sdy,sdz,mean distance from normal")
    assert rows[1] == "0,0,0,0,0,0,0" and rows[2] == "1,1e-12,2.5e-07,1,0.000512346,1e-10,321"
    pts = (tmp_path / "c.pcd").read_text().splitlines()[11:]
    assert pts[0] == "0 0 0 4278190080 0 0 1 0"
    assert pts[1] == "nan 9.9999997e-10 -123456.79 4278190080 0.57735026 -0.57735026 0.57735026 0"


def test_float_formatting_equals_printf(pcf):
    """std::to_chars(general, precision) in the writer == printf %g / %.8g (what ostream << float and PCL's ASCII writer emit)
    over random bit patterns, denormals, powers of ten, rounding boundaries and the special values."""
    lib = pcf.load_library()
    rng = np.random.default_rng(11)
    vals = [rng.integers(0, 2**32, 150000, dtype=np.uint64).astype(np.uint32).view(np.float32),
            (rng.normal(size=50000) * 10.0 ** rng.integers(-12, 9, 50000)).astype(np.float32),
            np.array([0.0, -0.0, 1.0, -1.0, 1e-5, 1e-4, 9.9999995e-5, 0.0001, 123456.5, 1234567.0, 999999.5, 999999.44, 1e6, 1e7, 99999999.0,
                      1e8, 0.5, 0.15, 1.17549435e-38, 1e-45, 3.4028235e38, np.inf, -np.inf, 0.000512345678, 2.5e-7, 1e-10], np.float32)]
    buf = C.create_string_buffer(64)
    for v in np.concatenate(vals):
        for prec, fmt in ((6, "%g"), (8, "%.8g")):
            n = lib.pcf_kat_format_float(C.c_float(float(v)), prec, buf)
            got = buf.value[:n].decode()
            want = "nan" if np.isnan(v) else fmt % float(v)
            assert got == want, (float(v), prec, got, want)
