"""ctypes binding of libhcomp_head.so (C ABI declared in include/hcomp_head.h).

The library is built in-tree by `__graft_entry__.build()` / `pipnet_b200.build.build()`.
There is NO fallback: if the shared object is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libhcomp_head.so')

ABI_VERSION = 6


class HcompError(RuntimeError):
    pass


class Tables(C.Structure):
    """mirror of `struct hcomp_tables`"""
    _fields_ = [
        ('n_nodes', C.c_int32), ('n_protos', C.c_int32), ('n_cols', C.c_int32), ('n_leaves', C.c_int32),
        ('n_welems', C.c_int32), ('p_max', C.c_int32),
        ('proto_off', C.c_void_p), ('cls_off', C.c_void_p), ('wc_off', C.c_void_p), ('proto_node', C.c_void_p),
        ('col_node', C.c_void_p), ('welem_col', C.c_void_p), ('welem_proto', C.c_void_p), ('child_w', C.c_void_p),
        ('path_off', C.c_void_p), ('path_col', C.c_void_p), ('anc', C.c_void_p), ('col_nleaves', C.c_void_p),
    ]


class DzBlocks(C.Structure):
    """mirror of `struct hcomp_dz_blocks`"""
    _fields_ = [('t1', C.c_void_p), ('ld1', C.c_int32), ('t2', C.c_void_p), ('ld2', C.c_int32), ('pcol', C.c_void_p),
                ('iact', C.c_void_p), ('iact_pitch', C.c_int32), ('tile_of_node', C.c_void_p),
                ('dz_only_read_through_tables', C.c_int32)]


class Spill(C.Structure):
    """mirror of `struct hcomp_spill`"""
    _fields_ = [('n_spill', C.c_int32), ('ldz', C.c_int32), ('recs_host', C.c_void_p), ('zs', C.c_void_p),
                ('stats', C.c_void_p)]


_p, _i, _f, _ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong
_T = C.POINTER(Tables)

# name -> argtypes (every function returns int except the two noted below)
SIGNATURES = {
    'hcomp_pack_weights': [_p, _p, _i, _i, _p, _p],
    'hcomp_cast_f32_to_bf16': [_p, _p, _ll, _p],
    'hcomp_nchw_to_rows_bf16': [_p, _i, _i, _i, _i, _p, _p],
    'hcomp_scale_residual_rows_bf16': [_p, _i, _p, _i, _p, _p, _i, _i, _i, _p, _p],
    'hcomp_label_tables': [_p, _T, _i, _i, _p, _p, _p, _p],
    'hcomp_split3_f32': [_p, _p, _ll, _p],
    'hcomp_pack_weights_split3': [_p, _p, _i, _i, _p, _p],
    'hcomp_proj_softmax_pool_fwd': [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _i, _p, _p, _p, _p, _p],
    'hcomp_unpack_pool': [_p, _ll, _f, _p, _p, _p],
    'hcomp_align_finalize': [_p, _p, _i, _i, _p, _p],
    'hcomp_head_bwd_dz': [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _p, _p, _p, _f, _p, _p, _p, _p, _p, _p, _p,
                          _p, _p, _p, _p],
    'hcomp_head_bwd_dx': [_p, _p, _ll, _i, _i, _p, _p, _p],
    'hcomp_head_bwd_dw': [_p, _p, _p, _ll, _i, _i, _p, _p, _p],
    'hcomp_classifier_fwd': [_p, _p, _p, _T, _i, _p, _p],
    'hcomp_classifier_bwd': [_p, _p, _p, _T, _i, _p, _i, _p, _p, _p],
    'hcomp_head_losses_fwd': [_p, _p, _p, _p, _p, _p, _p, _T, _i, _i, _i, _i, _p, _f, _f, _p, _p, _p, _p, _p, _p],
    'hcomp_head_losses_bwd': [_p, _p, _p, _p, _p, _p, _T, _i, _i, _i, _i, _p, _f, _f, _p, _p, _p, _p, _p, _p, _p],
    'hcomp_head_prologue': [_p, _p, _i, _i, _p, _p, _ll, _p, _i, _p, _T, _i, _i, _p, _p, _p, _p, _ll, _p],
    'hcomp_pool_classify_fwd': [_p, _p, _p, _p, _p, _T, _i, _i, _f, _p, _p, _p, _p, _p, _i, _f, _p, _p, _p],
    'hcomp_orth_gram': [_p, _p, _T, _i, _p, _p, _p],
    'hcomp_head_chain_fwd': [_p, _p, _p, _p, _p, _p, _p, _T, _i, _i, _i, _i, _p, _f, _f, _p, _p, _p, _p, _p, _p, _p],
    'hcomp_head_chain_bwd': [_p, _p, _p, _p, _p, _p, _p, _p, _T, _i, _i, _i, _i, _p, _f, _f, _p, _p, _p, _p, _p, _p, _p,
                             _p, _f, _p, _i, _p, _p, _p, _p],
    'hcomp_desc_losses_fwd': [_p, _p, _p, _p, _p, _p, _p, _T, _i, _i, _i, _p, _f, _f, _f, _p, _p, _p, _p],
    'hcomp_desc_losses_bwd': [_p, _p, _p, _p, _p, _p, _p, _p, _T, _i, _i, _i, _p, _f, _f, _f, _p, _p, _p, _p],
    'hcomp_joint_leaf': [_p, _T, _i, _f, _p, _p, _p, _p, _p],
    'hcomp_topk_update': [_p, _p, _p, _p, _p, _T, _i, _i, _i, _p, _p, _p, _p, _p],
    'hcomp_materialize_map': [_p, _p, _i, _i, _i, _i, _f, _p, _p],
    'hcomp_gemm_bf16': [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _ll, _p],
    'hcomp_allreduce_mean_symm': [_p, _p, _p, _p, _i, _i, _ll, _i, _i, _p],
}
EXPORTS = ['hcomp_abi_version', 'hcomp_last_error', 'hcomp_num_sms', 'hcomp_launch_count', 'hcomp_head_losses_ws_floats',
           'hcomp_desc_losses_ws_bytes', 'hcomp_head_chain_ws_floats', 'hcomp_set_cta_pair', 'hcomp_set_rider_fold', 'hcomp_set_reserved_sms', 'hcomp_init'] + list(SIGNATURES)

_lib = None


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise HcompError(f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                         f'(there is no CPU or PyTorch fallback for the prototype head)')
    L = C.CDLL(LIB_PATH)
    L.hcomp_abi_version.restype = C.c_int
    L.hcomp_last_error.restype = C.c_char_p
    L.hcomp_num_sms.restype = C.c_int
    L.hcomp_launch_count.restype = C.c_longlong
    L.hcomp_head_losses_ws_floats.restype = C.c_longlong
    L.hcomp_head_losses_ws_floats.argtypes = [_T]
    L.hcomp_head_chain_ws_floats.restype = C.c_longlong
    L.hcomp_head_chain_ws_floats.argtypes = [_T, C.c_int]
    L.hcomp_desc_losses_ws_bytes.restype = C.c_longlong
    L.hcomp_desc_losses_ws_bytes.argtypes = [_T, C.c_int]
    L.hcomp_set_cta_pair.restype = C.c_int
    L.hcomp_set_cta_pair.argtypes = [C.c_int]
    L.hcomp_set_rider_fold.restype = C.c_int
    L.hcomp_set_rider_fold.argtypes = [C.c_int]
    L.hcomp_set_reserved_sms.restype = C.c_int
    L.hcomp_set_reserved_sms.argtypes = [C.c_int]
    L.hcomp_init.restype = C.c_int
    L.hcomp_init.argtypes = []
    for name, args in SIGNATURES.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    if L.hcomp_abi_version() != ABI_VERSION:
        raise HcompError(f'ABI mismatch: library {L.hcomp_abi_version()} vs binding {ABI_VERSION}; rebuild')
    _lib = L
    return L


def call(name, *args):
    L = lib()
    rc = getattr(L, name)(*args)
    if rc != 0:
        raise HcompError(f'{name} failed ({rc}): {L.hcomp_last_error().decode(errors="replace")}')


def ptr(t):
    """device/host pointer of a torch tensor (or None -> NULL)"""
    return None if t is None else C.c_void_p(t.data_ptr())
