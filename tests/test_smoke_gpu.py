"""GPU: the driver's smoke entry point must pass (both kernel families)."""
import pytest

pytestmark = pytest.mark.gpu


def test_graft_entry_smoke(gpu_lib, kernel):
    import __graft_entry__ as g
    g.smoke()
