import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ckpt_cache(tmp_path_factory):
    """Synthetic checkpoints are deterministic from their kwargs: write each set once per session."""
    from image2video_synthesis_using_cinns_b200 import synthetic

    cache = {}

    def get(**kw):
        key = repr(sorted(kw.items()))
        if key not in cache:
            d = tmp_path_factory.mktemp("ckpt")
            cache[key] = synthetic.write_synthetic_checkpoints(str(d), **kw)
        return cache[key]

    return get
