"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares; the product package never imports the oracle; the ops refuse to run without CUDA."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "satmvs_b200.h")).read()
    return sorted(set(re.findall(r"\b(satmvs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from satmvs_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = header_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/satmvs_b200.h but not exported"
    assert sorted(_lib.exported_symbols()) == declared, "ctypes table and header disagree"
    assert _lib.lib().satmvs_abi_version() == 1


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "satmvs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f"{f} imports the oracle"
                assert "/root/reference" not in text, f"{f} reads the reference tree"


def test_no_cpu_fallback():
    import satmvs_b200
    from satmvs_b200 import synth
    fe = synth.make_features(1, 2, 2, 8, 8)
    rp = synth.make_rpc_stack(1, 2, 8, 8)
    dv = synth.make_depth_planes(1, 2, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        satmvs_b200.rpc_warping(fe[1], rp[:, 1], rp[:, 0], dv, None)
    with pytest.raises(RuntimeError, match="CUDA"):
        satmvs_b200.build_cost_volume(fe[0], fe[1:], rp[:, 0], [rp[:, 1]], dv)


def test_invalid_arguments_are_reported():
    from satmvs_b200 import _lib
    rc = _lib.lib().satmvs_rpc_warp_fwd(None, None, None, None, 0, 1, 1, 4, 4, None, None)
    assert rc == 1 and b"invalid argument" in _lib.lib().satmvs_last_error()
    rc = _lib.lib().satmvs_cost_volume_rpc_fwd(None, None, 99, None, None, None, 0, 1, 1, 4, 4, None, None)
    assert rc == 1


def test_library_is_sm100a_sass():
    from satmvs_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout
