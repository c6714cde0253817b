"""anticipated-vins-mono_b200: B200-native sliding-window VIO solver + anticipated feature selector.

The product is csrc/libbvio.so (C-ABI in include/bvio.h, hand-written sm_100a CUDA).
This package is the thin Python host used by tests and bench.py: ctypes bindings
(`lib`), struct mirrors (`abi`) and synthetic inputs (`synth`).  The directory name
is not a valid identifier; import it through `__graft_entry__.load_package()`.
There is no CPU fallback: `lib.load()` raises if libbvio.so is missing.
"""
from . import abi, synth, lib, shard, slider, horizon, replay  # noqa: F401
