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


def test_scratch_sizing_entry_points():
    """A C caller sizes every scratch buffer from the header alone (the K5 capacity rule lived in Python)."""
    from digdriver_b200 import _lib
    lib = _lib.load()
    for n, want in ((0, 1024), (1, 1024), (504, 1024), (505, 2048), (1_000_000, 2_097_152), (-5, 1024)):
        assert lib.dig_tabulate_capacity(n) == want, n
    cap = lib.dig_tabulate_capacity(300_000)
    assert cap >= 2 * 300_000 + 16 and cap & (cap - 1) == 0
    assert lib.dig_tabulate_elements_workspace_bytes(300_000) == cap * 16      # key 8 + snv 4 + indel 4
    assert lib.dig_tabulate_genes_workspace_bytes(300_000) == cap * 28         # key 8 + 5 class counters
    assert lib.dig_scan_workspace_bytes(310_000) >= 16 + 4 * 310_000           # redo count + one int32 per region
    # the opts struct of the ctypes table has the header's layout (8-byte pointers, no padding surprises)
    import ctypes
    assert ctypes.sizeof(_lib.ScanOpts) == 8 + 8 + 4 + 4 + 8 + 4 + 4 + 8 * 8 + 8
