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


def _header_prototypes():
    """{name: [parameter type strings]} for every `int hcomp_*(...)` prototype of the header"""
    text = open(os.path.join(ROOT, 'include', 'hcomp_head.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    protos = {}
    for m in re.finditer(r'\bint\s+(hcomp_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;', text, flags=re.S):
        params = [p.strip() for p in m.group(2).replace('\n', ' ').split(',')]
        protos[m.group(1)] = [] if params == ['void'] else params
    return protos


def test_ctypes_signatures_match_header_prototypes():
    """argument count and kind (pointer / float / 64-bit / int) of every bound entry point follow the header"""
    import ctypes as C
    from pipnet_b200 import _cabi
    protos = _header_prototypes()

    def kind_of_decl(p):
        if '*' in p:
            return 'ptr'
        if re.match(r'(const\s+)?float\b', p):
            return 'float'
        if re.match(r'(const\s+)?long long\b', p):
            return 'i64'
        return 'int'

    def kind_of_ctype(t):
        if t in (C.c_void_p,) or (isinstance(t, type) and issubclass(t, C._Pointer)):
            return 'ptr'
        return {C.c_float: 'float', C.c_longlong: 'i64', C.c_int: 'int'}[t]

    for name, argtypes in _cabi.SIGNATURES.items():
        assert name in protos, name
        want = [kind_of_decl(p) for p in protos[name]]
        got = [kind_of_ctype(t) for t in argtypes]
        assert got == want, f'{name}: binding {got} vs header {want}'
