"""The C-ABI libraries load on a CPU-only box and export every symbol the headers declare."""
import ctypes
import os
import re

import pytest

from you_can_not_recommend_b200 import build, front_end, native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(%s\w+)\s*\(" % prefix, text)))


def test_cuda_library_exports_header_symbols():
    lib = ctypes.CDLL(build.build_cuda())
    names = declared("ycnr_als.h", "ycnr_")
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(native.EXPORTS) == names


def test_host_library_exports_header_symbols():
    lib = ctypes.CDLL(build.build_host())
    names = declared("ycnr_host.h", "ycnr_")
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n


def test_cuda_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", build.build_cuda()], capture_output=True, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_(?!100a)\d+", out)


def test_create_fails_loudly_without_gpu():
    if native.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback"):
        native.Context(20, 10, 10)


def test_option_validation_messages():
    for kw, pat in ((dict(use_double_precision=True), "useDoublePrecision"), (dict(lowmem=True), "lowmem")):
        with pytest.raises(RuntimeError, match=pat):
            native.Context(20, 10, 10, **kw)
    with pytest.raises(RuntimeError, match="factorsCount"):
        native.Context(0, 10, 10)
    with pytest.raises(RuntimeError, match="factorsCount"):
        native.Context(1000, 10, 10)
