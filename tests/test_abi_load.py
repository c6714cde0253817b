"""CPU suite: the product library loads and exports every symbol include/bvio.h declares
(no compute calls -- there is no GPU here and no CPU fallback behind the ABI)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "bvio.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bvio_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_listed(pkg):
    assert sorted(pkg.lib.EXPORTS) == _declared_symbols()


def test_library_exports_every_symbol(pkg):
    if not os.path.exists(pkg.lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    L = pkg.lib.load()
    for name in _declared_symbols():
        assert hasattr(L, name), name
    assert L.bvio_abi_version() == 2
    o = pkg.abi.Opts()
    L.bvio_default_opts(C.byref(o))
    assert o.max_iters == 8 and o.focal_length == 460.0 and o.strategy == 0


def test_no_cpu_fallback_without_device(pkg):
    """bvio_create must fail (not fall back) when no CUDA device is usable."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = pkg.lib.load()
    h = C.c_void_p()
    assert L.bvio_create(0, C.byref(h)) == -2  # BVIO_ERR_CUDA
    assert not h.value


def test_product_does_not_link_oracle(pkg):
    import subprocess
    out = subprocess.run(["nm", "-D", pkg.lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle_" not in out
    ldd = subprocess.run(["ldd", pkg.lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "liboracle" not in ldd


def test_ctypes_mirror_matches_header_layout(pkg, tmp_path):
    """abi.py restates include/bvio.h by hand: compile a probe against the header and compare sizeof / offsetof of every
    struct (a silent mismatch would shift every field after it)."""
    import subprocess
    abi = pkg.abi
    pairs = [("bvio_preint", abi.Preint), ("bvio_prior", abi.Prior), ("bvio_window", abi.WindowS), ("bvio_opts", abi.Opts),
             ("bvio_summary", abi.Summary), ("bvio_prior_out", abi.PriorOut), ("bvio_imu_segment", abi.ImuSegment),
             ("bvio_camera", abi.Camera), ("bvio_select_in", abi.SelectIn), ("bvio_select_summary", abi.SelectSummary)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT}/include/bvio.h"', "int main(void) {"]
    for cname, cls in pairs:
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-o", str(exe), str(src)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in pairs:
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)
