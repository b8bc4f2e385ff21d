"""Bring-up / regression of the tcgen05 building blocks (descriptors, TMEM addressing, commit/mbarrier)."""
import ctypes

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant,name", [(0, "K-major A/B, N=128"), (1, "K-major A/B, N=80")])
def test_umma_k_major(variant, name):
    from diffudf_b200 import _lib
    err = ctypes.c_float(-1)
    _lib.check(_lib.lib().dudf_selftest_umma(variant, ctypes.byref(err)), "selftest")
    print(f"umma selftest {name}: max abs err {err.value:.3e}")
    assert 0 <= err.value < 1e-3          # fp16 inputs are exact, K=256 fp32 accumulation


@pytest.mark.parametrize("variant,name", [(2, "MN-major B"), (3, "MN-major A and B")])
def test_umma_mn_major_probe(variant, name):
    """MN-major operand layouts (needed by the tensor-core weight-gradient kernel)."""
    from diffudf_b200 import _lib
    err = ctypes.c_float(-1)
    _lib.check(_lib.lib().dudf_selftest_umma(variant, ctypes.byref(err)), "selftest")
    print(f"umma selftest {name}: max abs err {err.value:.3e}")
    assert 0 <= err.value < 1e-3
