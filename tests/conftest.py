import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def small_pack():
    import rrtmgp_b200 as R
    return R.synthetic.make_lut_pack(seed=11, dims=R.synthetic.SMALL_DIMS)


@pytest.fixture(scope="session")
def real_pack():
    import rrtmgp_b200 as R
    return R.synthetic.make_lut_pack(seed=7)
