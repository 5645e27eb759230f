"""CPU tests of the drop-in boundary: libsfgpu.so loads, exports every symbol include/sfgpu.h
declares, and fails loudly (never falls back to the CPU) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

from solverforge_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sfgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sfgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported_and_bound():
    names = _declared_symbols()
    assert len(names) >= 30
    lib = L.load()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sfgpu.h but not exported"
        assert n in L.SYMBOLS, f"{n} has no ctypes binding"
    assert set(L.SYMBOLS) == set(names)
    assert lib.sfgpu_abi_version() == 1


def test_struct_layouts_match_header():
    assert C.sizeof(L.Weight) == 24
    assert C.sizeof(L.ConstraintDesc) == 8 + 24 + 16 + 16 + 8
    assert C.sizeof(L.ForageParams) == 16


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_cuda(), reason="only meaningful on a box without a GPU")
def test_no_gpu_fails_loudly_without_cpu_fallback():
    lib = L.load()
    h = C.c_void_p()
    rc = lib.sfgpu_ctx_create(0, 0, None, C.byref(h))
    assert rc == L.E_CUDA
    assert not h.value
    assert b"no CPU fallback" in lib.sfgpu_last_error(None)
    from solverforge_b200 import GpuScoreDirector
    with pytest.raises(L.SfgpuError):
        GpuScoreDirector(1)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under solverforge_b200/ or include/ may reference it."""
    bad = []
    for base in ("solverforge_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    src = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"oracle_lib|liboracle|oracle/|import oracle|from oracle|sfo_", src):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_rust_sys_crate_declares_every_symbol():
    """bindings/rust/solverforge-gpu-sys cannot be compiled here (no cargo); keep its symbol list in sync."""
    src = open(os.path.join(ROOT, "bindings", "rust", "solverforge-gpu-sys", "src", "lib.rs")).read()
    declared = set(re.findall(r"pub fn (sfgpu_[a-z0-9_]+)\(", src))
    assert declared == set(_declared_symbols())


def test_rust_host_crate_calls_only_declared_symbols():
    """bindings/rust/solverforge-gpu (the `Director` + batch seam; not compilable here): every sys::sfgpu_* it calls
    is declared by the header and by the -sys crate."""
    src = open(os.path.join(ROOT, "bindings", "rust", "solverforge-gpu", "src", "lib.rs")).read()
    used = set(re.findall(r"sys::(sfgpu_[a-z0-9_]+)\(", src))
    assert len(used) >= 30
    assert used <= set(_declared_symbols())
    sys_src = open(os.path.join(ROOT, "bindings", "rust", "solverforge-gpu-sys", "src", "lib.rs")).read()
    for ty in set(re.findall(r"sys::(sfgpu_[a-z0-9_]+)\b(?!\()", src)):
        assert f"pub struct {ty}" in sys_src, ty
