"""CPU: the C-ABI library loads and exports every symbol include/gmp_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gmp_b200.h")
LIB = os.path.join(ROOT, "geometric-message-passing_b200", "libgmp_b200.so")


def declared_symbols():
    txt = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(gmp_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert "gmp_schnet_cfconv_fwd" in syms and "gmp_radius_graph_fill" in syms and len(syms) >= 20


def test_library_exports_every_declared_symbol():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(LIB)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    lib.gmp_version.restype = ctypes.c_int
    assert lib.gmp_version() >= 100


def test_binding_table_matches_header():
    import gmp_b200
    assert sorted(gmp_b200._lib.exported_symbols()) == declared_symbols()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "geometric-message-passing_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "/root/reference" not in src, f


def test_no_cpu_fallback():
    import torch
    import gmp_b200
    with pytest.raises(gmp_b200._lib.GmpError):
        gmp_b200.scatter(torch.zeros(4, 4), torch.zeros(4, dtype=torch.long), dim=0, dim_size=2)


def test_binding_arity_matches_header():
    """Every ctypes signature has as many arguments as the header's declaration (a missing trailing stream argument makes
    ctypes read a garbage pointer: a host-side crash, not an error code)."""
    import gmp_b200
    txt = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    arity = {}
    for name, params in re.findall(r"\b(gmp_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", txt):
        params = params.strip()
        arity[name] = 0 if params in ("", "void") else params.count(",") + 1
    sigs = dict(gmp_b200._lib._SIGS)
    sigs.update({k: v[1] for k, v in gmp_b200._lib._PLAIN.items()})
    bad = {k: (len(v), arity.get(k)) for k, v in sigs.items() if arity.get(k) != len(v)}
    assert not bad, f"ctypes arity != header arity: {bad}"
