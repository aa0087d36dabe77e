"""pytest configuration: the `gpu` marker and shared helpers.

`-m "not gpu"`: oracle vs its independent numpy restatement and the golden fixtures, host logic,
ABI symbol checks (no compute calls).  `-m gpu`: the parity tests proper, through the C ABI.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_c
    oracle_c.build()
    return oracle_c


def rand_field(rng, shape, scale=1.0, offset=0.0):
    return offset + scale * rng.standard_normal(shape)


def rand_porosity(rng, shape, lo=1e-6):
    """porosity in [lo, 1] with solid, fluid and interface regions (incl. exact 1.0 and lo values)"""
    e = rng.random(shape)
    e = np.clip((e - 0.2) / 0.6, lo, 1.0)
    return e


def rel_l2(a, b):
    d = np.linalg.norm((a - b).ravel())
    n = np.linalg.norm(b.ravel())
    return d / n if n > 0 else d
