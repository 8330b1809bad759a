"""The driver's entry point: `__graft_entry__.smoke()` must run on a GPU box exactly as the driver calls it."""
import pytest

pytestmark = pytest.mark.gpu


def test_smoke_entry_runs():
    import __graft_entry__ as g
    g.smoke()
