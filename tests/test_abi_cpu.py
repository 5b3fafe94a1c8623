"""The C-ABI boundary without a GPU: every entry point include/aptp_sm100.h declares is bound in _lib.SIGNATURES with the
same number of arguments, and the built library exports it. No compute entry point is called (aptp_version and
aptp_last_error need no device)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "aptp_sm100.h")


def _declared():
    """{name: number of parameters} of every `int aptp_*(...)` / `int64_t aptp_*(...)` / `const char* aptp_*(...)` prototype in the header."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)   # comments may mention entry points
    out = {}
    for m in re.finditer(r"\b(?:int|int64_t|const\s+char\s*\*)\s+(aptp_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        params = m.group(2).strip()
        out[m.group(1)] = 0 if params in ("", "void") else len([p for p in params.split(",") if p.strip()])
    return out


def test_header_and_ctypes_bindings_agree():
    from diffusion_pruning_b200 import _lib
    decl = _declared()
    assert len(decl) >= 40, sorted(decl)
    assert set(decl) == set(_lib.SIGNATURES), (sorted(set(decl) - set(_lib.SIGNATURES)),
                                               sorted(set(_lib.SIGNATURES) - set(decl)))
    for name, n in decl.items():
        assert len(_lib.SIGNATURES[name][1]) == n, f"{name}: header has {n} parameters, the binding {len(_lib.SIGNATURES[name][1])}"


def test_library_builds_loads_and_exports_every_declared_symbol():
    from diffusion_pruning_b200 import _lib, build
    path = build.build()                       # no-op when the in-tree .so is up to date (nvcc cross-compiles here)
    raw = ctypes.CDLL(os.fspath(path))
    for name in _declared():
        assert hasattr(raw, name), f"{path} does not export {name}"
    lib = _lib.load()
    assert lib.aptp_version() >= 1
    assert isinstance(lib.aptp_last_error(), bytes)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No CPU / eager fallback: without the extension the product path raises instead of computing something else."""
    from diffusion_pruning_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setenv("APTP_LIB", os.path.join(tmp_path, "nope.so"))
    try:
        _lib.load()
    except _lib.AptpError as e:
        assert "missing" in str(e)
    else:
        raise AssertionError("loading a missing library must raise")
    finally:
        monkeypatch.setattr(_lib, "_lib", None)
