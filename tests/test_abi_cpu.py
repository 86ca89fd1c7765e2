"""CPU: the C-ABI library loads without a GPU and exports every symbol declared in include/airv2x_b200.h."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "airv2x_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(a2x_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_all_declared_symbols(pkg):
    import a2x_import

    lib = a2x_import.pkg("_lib").load()
    syms = declared_symbols()
    assert len(syms) >= 35, syms
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.a2x_version() >= 100
    assert lib.a2x_last_error() is not None


def test_bad_arguments_fail_loudly_without_gpu(pkg):
    """argument validation happens before any CUDA call: status 1 + message, never a silent fallback."""
    import a2x_import

    libm = a2x_import.pkg("_lib")
    lib = libm.load()
    sh = libm.ConvShape(1, 8, 8, 30, 32, 3, 1)  # cin not a multiple of 32
    rc = lib.a2x_conv2d_fwd(ctypes.byref(sh), None, None, None, None, None, 0, None, None)
    assert rc == 1
    assert b"multiples of 32" in lib.a2x_last_error()
    rc = lib.a2x_split(None, ctypes.c_longlong(8), None, None)
    assert rc == 1


def test_module_refuses_cpu(pkg):
    import json

    import pytest
    import torch

    import a2x_import

    M = a2x_import.pkg("opencood.models.airv2x_where2com")
    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "w2c_small_config.json")))
    m = M.Airv2xWhere2com(cfg["model_args"])
    with pytest.raises(RuntimeError, match="no CPU path"):
        _ = m.engine


def test_every_entry_point_is_documented():
    """INTEGRATION.md names every C entry point the header declares (brace groups like a2x_conv2d_{fwd,dgrad} expanded),
    so the reference-side binding table cannot silently go stale."""
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "airv2x_b200.h")).read()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    names = sorted(set(re.findall(r"\b(a2x_[a-z0-9_]+)\s*\(", hdr)))
    known = set(re.findall(r"a2x_[a-z0-9_]+", doc))
    for m in re.finditer(r"(a2x_[a-z0-9_]*)\{([^}]+)\}", doc):
        known.update(m.group(1) + part.strip() for part in m.group(2).split(","))
    prefixes = [m.group(1) for m in re.finditer(r"(a2x_[a-z0-9_]+_)\*", doc)]
    missing = [n for n in names if n not in known and not any(n.startswith(p) for p in prefixes)]
    assert len(names) > 70 and not missing, missing


def test_header_is_plain_c_and_cpp(tmp_path):
    """the boundary is a C ABI: include/airv2x_b200.h compiles as C99 (-pedantic, no warnings) and as C++17 with no
    dependency beyond the standard headers (no torch / CUDA types in any signature)"""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        import pytest
        pytest.skip("no gcc")
    src = tmp_path / "h.c"
    src.write_text('#include "airv2x_b200.h"\nint (*probe)(void) = a2x_version;\nint main(void) { return probe == 0; }\n')
    inc = os.path.join(ROOT, "include")
    for cmd in (["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only"],
                ["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++"]):
        r = subprocess.run(cmd + ["-I", inc, str(src)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    text = open(os.path.join(inc, "airv2x_b200.h")).read()
    assert "torch" not in re.sub(r"/\*.*?\*/", "", text, flags=re.S) and "cuda_runtime" not in text


def test_missing_library_fails_loudly(pkg, monkeypatch, tmp_path):
    """no silent fallback: without the built library (and no way to build it) loading raises, and so does every op"""
    import pytest

    import a2x_import

    libm = a2x_import.pkg("_lib")
    bld = a2x_import.pkg("build")
    monkeypatch.setattr(libm, "_lib", None)
    monkeypatch.setattr(libm, "LIB_PATH", str(tmp_path / "libairv2x_b200.so"))
    monkeypatch.setattr(bld, "build", lambda *a, **k: None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        libm.load()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        libm.call("a2x_version")
