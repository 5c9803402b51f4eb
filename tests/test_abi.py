"""CPU: the C-ABI shared library loads, exports every symbol include/pcfusion.h declares, and fails loudly
(no CPU fallback) when there is no CUDA device.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pcfusion.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pcf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(pcf):
    lib = pcf.load_library()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"libpcfusion.so does not export {n}"
    from importlib import import_module
    binding = import_module(pcf.__name__ + ".binding")
    assert sorted(binding.ABI_SYMBOLS) == names


def test_default_config_matches_reference_constants(pcf):
    from importlib import import_module
    b = import_module(pcf.__name__ + ".binding")
    cfg = b._Config()
    pcf.load_library().pcf_default_config(C.byref(cfg))
    assert list(cfg.box) == [-0.8, 1.8, -1.5, 1.5, 0.0, 1.0]          # launch:8
    assert cfg.res[0] == np.float32(0.005)                           # node.cpp:91
    assert (cfg.clip_zmin, cfg.clip_zmax) == (0.28, 0.6)             # node.cpp:92-93
    assert (cfg.k_neighbourhood, cfg.walk_k, cfg.min_neighbours) == (2, 3, 20)
    assert (cfg.cylinder_radius, cfg.ball_radius) == (0.001, 0.015)  # OG.hpp:35-36


def test_no_cpu_fallback(pcf):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pcf.PcfError) as e:
        pcf.Fusion((-0.25, 0.25) * 3, 0.005)
    assert e.value.code == -5 and "no CPU path" in str(e.value)


def test_null_arguments_are_errors_not_crashes(pcf):
    lib = pcf.load_library()
    assert lib.pcf_start(None) == -1 and lib.pcf_update(None) == -1 and lib.pcf_sync(None) == -1
    assert lib.pcf_push_frame(None, None, 0, 4, None, 0) == -1
    lib.pcf_destroy(None)
