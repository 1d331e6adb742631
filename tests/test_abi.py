"""The C-ABI library loads and exports every symbol include/genetic_b200.h declares, and the ctypes table of the Python
host layer covers exactly those symbols (no compute calls: runs without a GPU)."""
import ctypes
import os
import re

from genetic_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "genetic_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gx_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared()
    assert len(names) > 60
    cdll = ctypes.CDLL(L.LIBPATH)
    missing = [n for n in names if not hasattr(cdll, n)]
    assert not missing, "declared in the header but not exported: %s" % missing
    unbound = [n for n in names if n not in L.PROTOTYPES]
    assert not unbound, "declared in the header but not in the ctypes table: %s" % unbound
    stale = [n for n in L.PROTOTYPES if n not in names]
    assert not stale, "bound but not declared: %s" % stale


def test_loader_reports_a_missing_library(monkeypatch):
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIBPATH", os.path.join(ROOT, "genetic_b200", "no_such_library.so"))
    try:
        L.load()
    except L.GxError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("load() must raise when the CUDA library is missing")
