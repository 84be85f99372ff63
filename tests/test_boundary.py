"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares, the ctypes table matches the header, and the product never routes through the oracle."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_decls():
    src = open(os.path.join(ROOT, "include", "dig_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"(?:int|int64_t|const char \*)\s*(dig_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        decls[m.group(1)] = n
    return decls


def test_library_exports_every_declared_symbol():
    from digdriver_b200 import build, _lib
    build.build_library()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    decls = _header_decls()
    assert len(decls) >= 15
    for name in decls:
        assert hasattr(lib, name), "libdigb200.so does not export %s" % name


def test_ctypes_table_matches_header():
    from digdriver_b200 import _lib
    decls = _header_decls()
    assert set(decls) == set(_lib.SIGNATURES), set(decls) ^ set(_lib.SIGNATURES)
    for name, n in decls.items():
        assert len(_lib.SIGNATURES[name][1]) == n, name
    lib = _lib.load()
    assert lib.dig_version() >= 100
    assert lib.dig_packed_words(33) == 4 and lib.dig_nmask_words(33) == 2


def test_product_never_imports_the_oracle():
    bad = []
    for base in ("digdriver_b200", "scripts"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"^\s*(from|import)\s+oracle\b|^\s*import\s+.*\bdig_oracle\b|libdig_oracle|"
                                 r"#\s*include\s+[<\"][^>\"]*oracle", txt, flags=re.M):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_missing_extension_fails_loudly(monkeypatch):
    from digdriver_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libdigb200.so")
    import pytest
    with pytest.raises(_lib.DigError):
        _lib.load()
