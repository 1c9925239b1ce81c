import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import cmodel
    cmodel.lib()
    return cmodel


@pytest.fixture(scope="session")
def detector_cfg():
    from oracle import reference_glue as rg
    return rg.check_configuration(dict(rg.DEFAULT_DETECTOR_CONFIG))
