"""The C-ABI library loads on a CPU-only box and exports every symbol include/vlr_gpu.h declares;
compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "vlr_gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(vlr_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from libprosic_b200 import _ffi
    lib = _ffi.load()
    syms = _declared_symbols()
    assert len(syms) >= 12
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(_ffi.EXPORTS) == syms
    assert lib.vlr_version() >= 100


def test_build_stamp_matches_the_sources(monkeypatch):
    """The library carries the hash of the sources it was built from; a binary that does not match
    the tree next to it is refused at load time (a GPU box has no nvcc and uses the shipped .so)."""
    from libprosic_b200 import _ffi, build
    lib = _ffi.load()
    assert lib.vlr_build_stamp().decode() == build.source_stamp()
    monkeypatch.setattr(_ffi, "_lib", None)
    monkeypatch.setattr(build, "build", lambda *a, **k: build.OUT)          # "no nvcc here"
    monkeypatch.setattr(build, "source_stamp", lambda: "0" * 40)            # the tree has moved on
    with pytest.raises(ImportError, match="stale"):
        _ffi.load()


def test_no_gpu_no_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import libprosic_b200 as P
    with pytest.raises(P.VlrError):
        P.Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "libprosic_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "vlr_oracle" not in txt and "import oracle" not in txt, f


def test_pairhmm_workspace_accounts_for_long_windows():
    """vlr_pairhmm_workspace_bytes is a pure function (no context): bookkeeping grows with the pairs,
    read windows above 248 rows add the per-warp scratch of the strip-wise and the anti-diagonal
    kernels (84 bytes per read row and warp for the latter), shorter windows add nothing."""
    from libprosic_b200 import _ffi
    lib = _ffi.load()
    w = lib.vlr_pairhmm_workspace_bytes
    assert w(1000, 0, 0) < w(100000, 0, 0)
    assert w(1000, 500, 248) == w(1000, 0, 0)
    base, long_rows = w(1000, 300, 0), w(1000, 300, 1024)
    assert long_rows - base >= 1184 * 84 * 1024            # 148 SMs x 2 CTAs x 4 warps at least
    assert w(1000, 20000, 15000) - base >= 1184 * 6 * 8 * 20000   # boundary rows of the strip-wise variant
