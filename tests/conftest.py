import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` under gpurun)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(params=[1, 0], ids=["cta-pair", "single-cta"])
def cta_pair_mode(request):
    """Select the kernel family for a GPU test: tcgen05 cta_group::2 (CTA pairs; the default) or the 1-CTA kernels."""
    from pipnet_b200 import _cabi
    prev = _cabi.lib().hcomp_set_cta_pair(request.param)
    yield request.param
    _cabi.lib().hcomp_set_cta_pair(prev)
