import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'src')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    # skip marks such as needs_ckpt are evaluated at collection time: stage the reference artefacts (cheap file copies, build
    # container only) before that, so that a fresh checkout does not skip the checkpoint / AncPhore tests on its first run
    import __graft_entry__ as ge
    ge.stage_reference_artifacts()
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def built_lib():
    import __graft_entry__ as ge
    ge.build()
    from diffphore_b200 import lib
    return lib
