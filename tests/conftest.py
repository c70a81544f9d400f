import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib_built():
    """Builds librgbavae.so if nvcc is here and the library is stale (no-op on the GPU box)."""
    import __graft_entry__ as G

    B = G._load_builder()
    try:
        B.build()
    except RuntimeError:
        if not os.path.exists(B.LIB):
            raise
    return B.LIB


@pytest.fixture(scope="session")
def golden():
    from safetensors.torch import load_file

    def load(arch):
        return load_file(os.path.join(ROOT, "tests", "golden", f"{arch}_c1_256.safetensors"))

    return load


_ORACLES = {}


@pytest.fixture(scope="session")
def oracle_model():
    """Seed-0 random-init oracle per arch (the weights the golden vectors were made with)."""
    from oracle import vae_oracle as O

    def get(arch):
        if arch not in _ORACLES:
            _ORACLES[arch] = O.build_oracle(arch, seed=0)
        return _ORACLES[arch]

    return get
