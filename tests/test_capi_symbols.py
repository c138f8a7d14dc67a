"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol the
header declares, and refuses to run without a CUDA device (no compute calls are made here)."""
import ctypes as C
import os
import re

import pytest

from mars_ode_physics_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "mars_ode_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2p_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    names = _declared()
    assert len(names) > 70
    for strict in (False, True):
        lib = capi.load(strict)
        for n in names:
            assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(capi.SIGNATURES), set(names) ^ set(capi.SIGNATURES)


def test_version_and_error_string():
    lib = capi.load()
    assert lib.b2p_version() == 1
    assert isinstance(lib.b2p_last_error(), bytes)


def test_no_cpu_fallback():
    """Without a device the arena cannot be created: the product path fails loudly."""
    lib = capi.load()
    if lib.b2p_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    rc = lib.b2p_arena_create(4, 0, C.byref(h))
    assert rc == -2 and not h.value
    assert b"no CPU fallback" in lib.b2p_last_error()


def test_facade_library_exports_plugin_entry_points():
    path = os.path.join(ROOT, "mars_ode_physics_b200", "lib", "libmars_ode_physics_b200_facade.so")
    assert os.path.exists(path), "run __graft_entry__.build()"
    lib = C.CDLL(path)
    assert hasattr(lib, "create_c") and hasattr(lib, "destroy_c")  # lib_manager CREATE_LIB / DESTROY_LIB


def test_product_path_does_not_touch_the_oracle():
    """Nothing under mars_ode_physics_b200/ may import, link or execute oracle/ (bench.py's cpu
    baseline and the tests are the only users)."""
    pkg = os.path.join(ROOT, "mars_ode_physics_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for pat in (r'#include\s*[<"][^>"]*orc', r"\bimport orc\b", r"\bfrom orc\b", r"liborc",
                            r"\borc_[a-z_]+\s*\("):
                    assert not re.search(pat, text), (os.path.join(dirpath, f), pat)
