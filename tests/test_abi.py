"""CPU: the C-ABI shared library builds, loads and exports every symbol the header declares."""
import ctypes
import os
import re

import helpers


def test_library_exports_header_symbols():
    from waafle_b200 import build, engine
    lib_path = build.build_library()
    assert os.path.exists(lib_path)
    hdr = open(os.path.join(helpers.GOLDEN, "..", "..", "include", "waafle_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(wfl_[a-z_]+)\s*\(", hdr)))
    assert declared, "no entry points found in the header"
    lib = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(lib, name), "missing export: " + name
    assert sorted(engine.EXPORTS) == declared
    assert engine.load_library().wfl_abi_version() == engine.ABI_VERSION == 2


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the engine must fail loudly, not fall back."""
    from waafle_b200 import engine
    lib = engine.load_library()
    if lib.wfl_device_count() > 0:
        return
    try:
        engine.Engine(0)
    except engine.EngineError as exc:
        assert "no CPU fallback" in str(exc)
    else:
        raise AssertionError("Engine() succeeded without a GPU")


def test_product_never_imports_oracle():
    root = os.path.join(helpers.GOLDEN, "..", "..", "waafle_b200")
    for fn in os.listdir(root):
        if fn.endswith(".py"):
            src = open(os.path.join(root, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+.*oracle", src, re.M), fn
