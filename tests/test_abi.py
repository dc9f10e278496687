"""CPU-side checks of the drop-in boundary: libhegpu.so loads and exports exactly the
symbols include/hegpu.h declares; without a GPU the library refuses to create a context
(no CPU fallback)."""
import ctypes
import os
import re

import pytest

import hegpu_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "hegpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hegpu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    hg = hegpu_loader.load()
    L = hg.lib()
    names = _declared()
    assert len(names) >= 45
    for name in names:
        assert hasattr(L, name), f"libhegpu.so does not export {name}"
    assert b"sm_100a" in L.hegpu_version()


def test_python_binding_covers_the_abi():
    hg = hegpu_loader.load()
    L = hg.lib()
    for name in _declared():
        fn = getattr(L, name)
        if name in ("hegpu_last_error", "hegpu_version", "hegpu_ctx_stream", "hegpu_ctx_psi", "hegpu_launch_count", "hegpu_profile_kind_name"):
            continue
        assert fn.argtypes is not None, name


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    hg = hegpu_loader.load()
    with pytest.raises(hg.HegpuError) as e:
        hg.Context(8192, [0xFFFFFFFFFFE8001, 0xFFFFF4C001, 0xFFFFFDC001, 0xFFFFFFFFFFFC001])
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = hegpu_loader.PKG_DIR
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "ckks_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_seal_facing_sources_compile():
    """include/he_gpu_bridge.hpp (the reference-side binding) and tools/seal_golden.cpp (the real-SEAL vector dump)
    are written against SEAL 4.1, which this image lacks: type-check them against the declaration-only stand-in
    tests/mock_seal/seal/seal.h so that neither carries elided or untested lines."""
    import subprocess
    import tempfile

    inc = ["-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "mock_seal")]
    with tempfile.NamedTemporaryFile("w", suffix=".cpp") as f:
        f.write('#include "he_gpu_bridge.hpp"\nint main() { return 0; }\n')
        f.flush()
        for src in (f.name, os.path.join(ROOT, "tools", "seal_golden.cpp")):
            r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only"] + inc + [src], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
