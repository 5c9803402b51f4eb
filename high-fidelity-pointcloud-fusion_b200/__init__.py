"""pcfusion-b200: B200-native frame-integration + process() path of the `pointcloud_fusion` node.

The product is `libpcfusion.so` (csrc/, C ABI in include/pcfusion.h).  This package is the thin Python
host layer used by the tests, the bench and the multi-GPU replay: a ctypes binding (`binding.Fusion`), the
synthetic sequence generator (`synth`) and the frame-sharded multi-GPU merge (`sharded`).
There is no CPU fallback anywhere in here: without the built library or a CUDA device, calls raise.
"""
from .binding import Fusion, PcfError, Result, State, kat_clip_pack, lib_path, load_library  # noqa: F401
