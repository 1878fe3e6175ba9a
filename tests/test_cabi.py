"""The C-ABI shared library builds, loads without a GPU and exports every symbol include/*.h declares
(no compute calls here); the ctypes mirror of FlBatch has the library's layout; argument errors are
reported through return codes, never exceptions."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header="flatland_b200.h"):
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", header)).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(fl_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    import importlib.util
    spec = importlib.util.spec_from_file_location("_fl_build", os.path.join(ROOT, "flatland-marl_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    import flatland_marl_b200 as fb
    return fb._lib.lib()


def test_exports_every_declared_symbol(lib):
    import flatland_marl_b200 as fb
    syms = declared_symbols()
    assert len(syms) >= 13
    for s in syms:
        assert hasattr(lib, s), "libflatland_b200.so does not export %s" % s
    assert sorted(fb._lib.EXPORTS) == syms, "binding list and header disagree"


def test_every_header_has_a_library():
    assert sorted(f for f in os.listdir(os.path.join(ROOT, "include")) if f.endswith(".h")) == \
        ["flatland_b200.h", "flatland_policy_b200.h"]


def test_policy_library_exports_every_declared_symbol(lib):
    from flatland_marl_b200 import policy
    L = policy.lib()
    syms = declared_symbols("flatland_policy_b200.h")
    assert len(syms) >= 6
    for s in syms:
        assert hasattr(L, s), "libflatland_policy_b200.so does not export %s" % s
    assert sorted(policy.EXPORTS) == syms, "binding list and header disagree"
    assert L.fl_policy_abi_version() == 1
    n = L.fl_policy_workspace_bytes(51200)
    assert n > 51200 * 30000 and n % 256 == 0
    # argument errors come back as codes before any CUDA call
    assert L.fl_policy_choose_actions(None, None, None, 0, None) == -1


def test_struct_layout_and_version(lib):
    import flatland_marl_b200 as fb
    assert lib.fl_abi_version() == 5
    assert lib.fl_batch_sizeof() == C.sizeof(fb._lib.FlBatch)
    assert lib.fl_profile_num_kernels() >= 4
    assert lib.fl_error_string(0) == b"ok"
    assert b"bad argument" in lib.fl_error_string(-1)


def test_bad_arguments_return_codes(lib):
    import flatland_marl_b200 as fb
    b = fb._lib.FlBatch()  # all zero: rejected before any CUDA call
    assert lib.fl_step(C.byref(b), None, None, None, 0, None) == -1
    assert lib.fl_observe(C.byref(b), *([None] * 8), None) == -1
    assert lib.fl_distance_map(C.byref(b), None) == -1
    assert lib.fl_reset(C.byref(b), None, None) == -1
    b.E, b.N, b.H, b.W, b.n_slots, b.S, b.ent_cap = 1, 2000, 10, 10, 1, 1, 2000 * 501
    assert lib.fl_step(C.byref(b), None, None, None, 0, None) == -2
    with pytest.raises(fb.FlatlandB200Error):
        fb._lib.check(-2)


def test_product_never_imports_oracle():
    """The product package must not route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "flatland-marl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "flatland_oracle" not in text and "liboracle" not in text, f


def test_no_cuda_device_fails_loudly():
    import torch
    import flatland_marl_b200 as fb
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(fb.FlatlandB200Error):
        fb.BatchedRailEnv([{}])
