"""CPU-side checks of the drop-in boundary: the shared library builds, loads without a GPU and exports
every symbol declared in include/hcomp_head.h; no compute entry point is called here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib_path():
    from pipnet_b200 import build
    return build.build()


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'hcomp_head.h')).read()
    return sorted(set(re.findall(r'\b(hcomp_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f'{n} declared in hcomp_head.h but not exported'


def test_binding_covers_header(lib_path):
    from pipnet_b200 import _cabi
    assert sorted(_cabi.EXPORTS) == _declared_symbols()
    assert _cabi.lib().hcomp_abi_version() == _cabi.ABI_VERSION


def test_tile_constants_match_header():
    from pipnet_b200 import layout
    text = open(os.path.join(ROOT, 'include', 'hcomp_head.h')).read()
    assert int(re.search(r'#define HCOMP_TILE_INTS (\d+)', text).group(1)) == layout.TILE_INTS
    assert int(re.search(r'#define HCOMP_TILE_COLS (\d+)', text).group(1)) == layout.TILE_COLS
    assert int(re.search(r'#define HCOMP_MAX_SEGS (\d+)', text).group(1)) == layout.MAX_SEGS


def test_product_fails_loudly_without_cuda():
    """No CPU fallback: handing a CPU tensor to the head raises instead of computing somewhere else."""
    import torch
    from pipnet_b200 import ops, _cabi
    if torch.cuda.is_available():
        pytest.skip('CPU-only check')
    with pytest.raises(_cabi.HcompError):
        ops.feature_rows(torch.zeros(1, 8, 6, 6))
