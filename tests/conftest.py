import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", params=["euclidean", "cosine"])
def retrieval_golden(request):
    z = np.load(GOLDEN / f"retrieval_{request.param}.npz")
    d = {k: z[k] for k in z.files}
    d["loss_type"] = request.param
    return d


@pytest.fixture(scope="session")
def triplet_golden():
    z = np.load(GOLDEN / "triplet.npz")
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def sbir_lib():
    """Builds (if stale) and loads the C-ABI library; never falls back."""
    from art_sbir_b200 import _binding, _build
    _build.build()
    return _binding.load()
