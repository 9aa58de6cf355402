"""CPU: the C-ABI library loads, exports every symbol include/rt_b200.h declares, validates its
arguments, and FAILS LOUDLY (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def header_functions():
    src = open(os.path.join(ROOT, "include", "rt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    import distributed_raytracer_b200 as rt
    lib = C.CDLL(rt._lib.LIB_PATH)
    names = header_functions()
    assert set(names) == set(rt._lib.ABI_SYMBOLS)
    for n in names:
        assert getattr(lib, n) is not None
    assert rt._lib.load().rt_abi_version() == 2


def test_argument_validation_without_compute(built, example_scene):
    import distributed_raytracer_b200 as rt
    L = rt._lib.load()
    h = C.c_void_p()
    assert L.rt_scene_create(None, C.byref(h)) == rt._lib.RT_ERR_INVALID
    assert b"NULL" in L.rt_last_error(None)
    desc, keep = rt.make_scene_desc(example_scene, precision=7)
    assert L.rt_scene_create(C.byref(desc), C.byref(h)) == rt._lib.RT_ERR_INVALID
    bad = dict(example_scene)
    bad["obj_mesh"] = np.array([3], dtype=np.uint32)
    desc, keep = rt.make_scene_desc(bad)
    assert L.rt_scene_create(C.byref(desc), C.byref(h)) == rt._lib.RT_ERR_INVALID
    assert L.rt_frame_set(None, None, None, 0, None) == rt._lib.RT_ERR_INVALID
    assert L.rt_render_region(None, 1, 1, 0, 0, 1, 1, None, 0) == rt._lib.RT_ERR_INVALID
    L.rt_scene_destroy(None)   # must be a no-op


def test_no_cpu_fallback(built, example_scene):
    import distributed_raytracer_b200 as rt
    if rt._lib.load().rt_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(rt.RtError) as ei:
        rt.GpuScene(example_scene)
    assert ei.value.code == rt._lib.RT_ERR_CUDA and "no CPU fallback" in str(ei.value)


def test_product_does_not_import_the_oracle():
    """The product path must never route through oracle/ (only tests/, smoke() and bench's CPU legs may)."""
    pkg = os.path.join(ROOT, "distributed-raytracer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "rt_oracle" not in text and "librt_oracle" not in text, f
