import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run with -m gpu on a B200 box)")


def rel_max(a, b):
    """max|a-b| / max|b| — the error measure of BASELINE.md's precision probe."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load


@pytest.fixture(scope="session")
def oracle():
    from oracle import dudf_oracle
    return dudf_oracle


@pytest.fixture(scope="session")
def weights(oracle):
    return {tag: oracle.load_params(os.path.join(GOLDEN, f"weights_{tag}.npz")) for tag in ("init", "trained")}


@pytest.fixture(scope="session")
def cuda_models(weights):
    """diffudf_b200.SIREN modules on cuda:0 loaded with the golden weights (GPU tests only)."""
    import torch
    from diffudf_b200 import SIREN
    out = {}
    for tag, params in weights.items():
        m = SIREN(3, 1, [256] * (len(params) - 1), w0=30, delay_init=True)
        sd = {}
        for i, (W, b) in enumerate(params):
            sd[f"net.{i}.0.weight"] = torch.from_numpy(W)
            sd[f"net.{i}.0.bias"] = torch.from_numpy(b)
        m.load_state_dict(sd)
        out[tag] = m.to("cuda:0")
    return out
