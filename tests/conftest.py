"""Test configuration.

`-m "not gpu"` : oracle vs known answers / golden vectors, host logic, C-ABI load + symbols.
`-m gpu`       : parity of the CUDA path (through the C ABI) against the oracle.
Only tests may import `oracle/` (see oracle/__init__.py).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "safe-marl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def args():
    from oracle import env_ref
    return dict(env_ref.DEFAULT_ARGS)


@pytest.fixture(scope="session")
def net():
    from oracle import ieee33
    return ieee33.create_network()


@pytest.fixture(scope="session")
def tree(net):
    from oracle import ieee33
    return ieee33.tree_arrays(net)


@pytest.fixture(scope="session")
def fonet(tree, args):
    """C mirror in the default (thread-per-env) operation order."""
    from oracle import c_mirror
    return c_mirror.make_net(tree, args, args["buildings"])


@pytest.fixture(scope="session")
def fonets(tree, args):
    """C mirrors keyed by kernel variant name."""
    from oracle import c_mirror
    return {"thread": c_mirror.make_net(tree, args, args["buildings"], variant=c_mirror.VARIANT_THREAD),
            "warp": c_mirror.make_net(tree, args, args["buildings"], variant=c_mirror.VARIANT_WARP)}


@pytest.fixture(params=["thread", "warp"])
def variant(request):
    return request.param


@pytest.fixture(scope="session")
def network():
    from flexgpu.config import DEFAULT_ENV_ARGS
    from flexgpu.network import Network, create_network
    return Network(create_network(DEFAULT_ENV_ARGS))


@pytest.fixture(scope="session")
def profiles(network):
    from flexgpu.profiles import synthetic_profiles
    return synthetic_profiles(network, 5, T=4000, seed=0)


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def rel_err(a, b, floor=1e-12):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0
