"""patch_reference rebinds the operators in the namespaces the reference's networks actually use."""
import types

import satmvs_b200
from satmvs_b200 import integration


def test_patch_rebinds_star_imported_names():
    fake = types.ModuleType("networks.casred")
    fake.rpc_warping = object()
    fake.homo_warping = object()
    fake.RED_Regularization = object()
    fake.unrelated = 1
    done = integration.patch_reference([fake])
    assert fake.rpc_warping is satmvs_b200.rpc_warping
    assert fake.homo_warping is satmvs_b200.homo_warping
    assert fake.RED_Regularization is satmvs_b200.RED_Regularization
    assert fake.unrelated == 1
    assert sorted(done["networks.casred"]) == ["RED_Regularization", "homo_warping", "rpc_warping"]


def test_patch_on_live_reference_when_present():
    from oracle import reference_loader
    import pytest
    if not reference_loader.available():
        pytest.skip("reference tree not present (GPU box)")
    ref = reference_loader.load()
    saved = {(m, n): getattr(m, n) for m in (ref.casred, ref.casmvs) for n in ("rpc_warping", "homo_warping")}
    try:
        done = integration.patch_reference([ref.casred, ref.casmvs])
        assert ref.casred.rpc_warping is satmvs_b200.rpc_warping
        assert "CostRegNet" in done["networks.casmvs"]
    finally:
        for (m, n), v in saved.items():
            setattr(m, n, v)
        import importlib
        importlib.reload(ref.casred)
        importlib.reload(ref.casmvs)
