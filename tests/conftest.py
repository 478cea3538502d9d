import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# When the unmodified reference is vendored (baseline/_ref, see oracle/ref_glue.py) every test runs the plugin on the REAL
# ever.ERModule / ever.registry.MODEL (ever_b200/_ever_api.py picks them up at import time); without it the same-contract
# stand-in is used.
_REF = os.path.join(ROOT, 'baseline', '_ref')
if os.path.isdir(os.path.join(_REF, 'ever')):
    for _p in (os.path.join(ROOT, 'tests', 'golden', '_stubs'), _REF):
        if _p not in sys.path:
            sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run through gpurun)')


@pytest.fixture(scope='session')
def golden_dir():
    return os.path.join(ROOT, 'tests', 'golden')
