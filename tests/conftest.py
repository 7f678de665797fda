"""pytest configuration: the ``gpu`` marker (tests that need a B200) and shared fixtures."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def native_lib():
    from hfnet_slam_b200 import build, lib
    if not lib.LIB_PATH.exists():
        build.build_native()
    return lib.load()


@pytest.fixture(scope="session")
def weights_blob():
    from hfnet_slam_b200 import weights
    return weights.synthetic_blob(seed=0, n_clusters=32)


@pytest.fixture(scope="session")
def weights_dict():
    from hfnet_slam_b200 import weights
    return weights.synthetic(seed=0, n_clusters=32)


@pytest.fixture()
def small_ctx(native_lib):
    """A context without network weights: matcher / database / BA / network-tail hooks only."""
    from hfnet_slam_b200.lib import Context
    ctx = Context(height=64, width=64, n_levels=1, max_keypoints=8192, max_batch=1, with_global=False)
    yield ctx
    ctx.close()
