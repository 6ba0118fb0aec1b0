"""Host-side operators of the prototype head: thin wrappers that hand torch device memory and the
current CUDA stream to the C ABI (include/hcomp_head.h), plus the `torch.autograd.Function`s that
splice the kernels into autograd.  PyTorch is plumbing here (memory, streams, autograd graph); all
arithmetic of the path runs in libhcomp_head.so.  There is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch

from . import _cabi
from ._cabi import Tables, call, ptr
from .layout import HeadLayout


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _Profile:
    """Optional CUDA-event brackets around named C-ABI calls on the launching stream (bench.py uses this to
    time the fused kernels inside the step; off by default, zero cost)."""

    def __init__(self):
        self.enabled = False
        self.events = {}

    def start(self, name):
        if not self.enabled:
            return None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        self.events.setdefault(name, []).append((a, b))
        return b

    @staticmethod
    def stop(tok):
        if tok is not None:
            tok.record()

    def totals_ms(self):
        return {k: (sum(a.elapsed_time(b) for a, b in v), len(v)) for k, v in self.events.items()}

    def reset(self):
        self.events = {}


PROFILE = _Profile()

# Data-parallel training: when set (a torch.distributed process group), the flat prototype-kernel gradient
# is all-reduced (mean) on a side stream right after the dW GEMM so it overlaps the dX GEMM and whatever
# backbone backward follows.
GRAD_ALLREDUCE_GROUP = None
# SMs the dX GEMM leaves free while the gradient all-reduce is in flight (hcomp_set_reserved_sms).  Measured on 2xB200: the
# NCCL kernel does not make progress beside the persistent GEMM whether 0, 8 or 32 SMs are left free (its tail after the
# GEMM stays ~20 us), and capping NCCL to 8 CTAs makes the tail longer, so the default is 0; the collective is hidden
# behind the backbone backward instead (its join is at the end of the backward pass).
COLLECTIVE_SMS = int(os.environ.get('HC_COLLECTIVE_SMS', '0'))
_side_stream = None


def _side():
    global _side_stream
    if _side_stream is None:
        _side_stream = torch.cuda.Stream()
    return _side_stream


# ---- data-parallel gradient bucket ---------------------------------------------------------------------------------
# With GRAD_ALLREDUCE_GROUP set (the head's own data parallelism, not DistributedDataParallel) and ASYNC_GRAD_ALLREDUCE,
# the head's parameter gradients never travel through autograd: every producer (orth loss, classifier, presence logits,
# dW GEMM) writes into its segment of ONE zero-initialised flat buffer, ONE mean all-reduce of that buffer is issued on
# the side stream right after the dW GEMM (so it overlaps the dX GEMM and the backbone backward), and a callback at the
# END of the backward pass (the autograd-engine hook DDP's reducer uses) re-joins the stream and hands `param.grad` views
# out.  Handing the tensors to autograd instead would let it read them on the main stream while the collective is still
# in flight (input-buffer accumulation of dW + orth gradient, AccumulateGrad clones) -- tests/test_gpu_multi.py.
# ASYNC_GRAD_ALLREDUCE = False: gradients go through autograd; each all-reduce is joined before its tensor is returned
# (the dW one still overlaps the dX GEMM).
ASYNC_GRAD_ALLREDUCE = True


# How the bucket is reduced.  'symm' (default when available): the bucket lives in symmetric memory (every rank's buffer
# mapped into every peer, plus one multicast address over the NVSwitch) and the library's own kernel
# (hcomp_allreduce_mean_symm: barrier -> in-switch reduce of my shard -> broadcast -> barrier) runs on the side stream on
# the few SMs the dX GEMM is told to leave free.  'nccl': one NCCL all-reduce (the fallback when symmetric memory or the
# group's peer mapping is not available, e.g. ranks on different nodes).  Env HC_GRAD_EXCHANGE=nccl forces the fallback.
GRAD_EXCHANGE = os.environ.get('HC_GRAD_EXCHANGE', 'symm')
SYMM_CTAS = int(os.environ.get('HC_SYMM_CTAS', '8'))          # minimum CTAs of the all-reduce kernel = SMs reserved from the dX GEMM
SYMM_CTAS_MAX = 32                                            # large buckets (cub190: 11.7 MB) get one CTA per 128 KB of this rank's shard


class _SymmBuffers:
    """Two persistent symmetric-memory gradient buckets per (bucket layout, device, group): step i reduces into one while
    `param.grad` of step i-1 may still alias the other.  Created collectively (every rank at the same point of its first
    backward pass)."""

    _cache = {}

    def __init__(self, numel, device, group):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.numel = numel
        self.padded = (numel + 3) // 4 * 4
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.bufs, self.hdls = [], []
        for _ in range(2):
            t = symm.empty(self.padded, dtype=torch.float32, device=device)
            h = symm.rendezvous(t, group)
            t.zero_()
            self.bufs.append(t)
            self.hdls.append(h)
        self.turn = 0
        h = self.hdls[0]
        pad_words = int(h.signal_pad_size) // 4
        self.ctas = max(SYMM_CTAS, min(SYMM_CTAS_MAX, (numel * 4 // self.world) // (128 * 1024)))
        while self.ctas > 1 and self.ctas * self.world > pad_words:
            self.ctas //= 2
        if self.ctas * self.world > pad_words:
            raise RuntimeError('signal pad too small for the all-reduce channels')
        self.multicast = all(int(getattr(x, 'multicast_ptr', 0) or 0) != 0 for x in self.hdls)

    @classmethod
    def get(cls, numel, device, group):
        key = (numel, str(device), id(group))
        if key not in cls._cache:
            cls._cache[key] = cls(numel, device, group)
        return cls._cache[key]

    def next(self):
        self.turn ^= 1
        return self.turn


def _symm_buffers(numel, device):
    """symmetric buckets for the current group or None (-> NCCL fallback); the decision is made once per layout"""
    if GRAD_EXCHANGE != 'symm' or GRAD_ALLREDUCE_GROUP is None:
        return None
    key = (numel, str(device), id(GRAD_ALLREDUCE_GROUP))
    if key in _SymmBuffers._cache:
        return _SymmBuffers._cache[key]
    import torch.distributed as dist
    ok = 1
    try:
        sb = _SymmBuffers.get(numel, device, GRAD_ALLREDUCE_GROUP)
    except Exception as ex:             # no symmetric memory on this system / group: every rank must fall back together
        import warnings
        warnings.warn(f'symmetric-memory gradient bucket unavailable ({ex!r}); using the NCCL all-reduce')
        sb, ok = None, 0
    flag = torch.tensor([ok], device=device, dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=GRAD_ALLREDUCE_GROUP)
    if int(flag) == 0:
        sb = None
    _SymmBuffers._cache[key] = sb
    return sb


class _GradBucket:
    def __init__(self, family, device):
        self.family = family
        self.offsets, self.sizes, off = {}, {}, 0
        for grp in family:              # groups that were not used in this step (e.g. presence logits) get an idle segment
            n = sum(p_.numel() for p_ in grp.params)
            self.offsets[id(grp)], self.sizes[id(grp)] = off, n
            off += n
        self.numel = off
        self.symm = _symm_buffers(off, device)
        self.symm_turn = None
        if self.symm is not None:
            # persistent bucket: make sure no live param.grad still aliases the buffer about to be cleared (gradient
            # accumulation without zero_grad between steps), then clear it
            self.symm_turn = self.symm.next()
            buf = self.symm.bufs[self.symm_turn]
            lo, hi = buf.data_ptr(), buf.data_ptr() + buf.numel() * 4
            for grp in family:
                for p_ in grp.params:
                    if p_.grad is not None and lo <= p_.grad.data_ptr() < hi:
                        p_.grad = p_.grad.clone()
            buf.zero_()
            self.flat = buf[:off]
        elif GRAD_ALLREDUCE_GROUP is None:
            # local bucket (nothing is exchanged): no up-front clear -- a segment is cleared on first touch unless its first
            # producer overwrites every element (orth term, classifier, presence logits), so the usual step has no fill
            self.flat = torch.empty(off, device=device, dtype=torch.float32)
            self.lazy_zero = True
        else:
            self.flat = torch.zeros(off, device=device, dtype=torch.float32)
        self.produced = []
        self.reduced = False

    lazy_zero = False

    def segment(self, grp, overwrite=False):
        o = self.offsets[id(grp)]
        seg = self.flat[o:o + self.sizes[id(grp)]]
        if grp not in self.produced:
            self.produced.append(grp)
            if self.lazy_zero and not overwrite:
                seg.zero_()
        return seg


_bucket: Optional[_GradBucket] = None


# Single-process steps can use the bucket too (no exchange): every producer of a head-parameter gradient (orth term, dW
# GEMM, classifier) writes into its segment of one buffer and `param.grad` becomes a view of it -- no zero fill per
# producer, no autograd add of the orth gradient and dW (3 us of a 0.3 ms step).  Off by default: gradients must travel
# through autograd for DistributedDataParallel's hooks and for torch.autograd.grad(); the captured training step
# (train.GraphedHeadTrainStep, which owns the delivery of param.grad anyway) switches it on around its capture.
LOCAL_GRAD_BUCKET = False


class local_grad_bucket:
    """context manager: head-parameter gradients of backward passes run inside go through the flat bucket"""

    def __enter__(self):
        global LOCAL_GRAD_BUCKET
        self.prev, LOCAL_GRAD_BUCKET = LOCAL_GRAD_BUCKET, True
        return self

    def __exit__(self, *exc):
        global LOCAL_GRAD_BUCKET
        LOCAL_GRAD_BUCKET = self.prev
        return False


def _bucket_mode(grp) -> bool:
    return (((GRAD_ALLREDUCE_GROUP is not None and ASYNC_GRAD_ALLREDUCE) or LOCAL_GRAD_BUCKET) and grp is not None
            and getattr(grp, 'family', None) is not None)


def _bucket_segment(grp, device, overwrite=False) -> torch.Tensor:
    """This step's gradient segment of parameter group `grp` (zero-initialised; producers write or accumulate).
    overwrite=True: the caller writes EVERY element of the segment (lets the local bucket skip the clear)."""
    global _bucket
    if _bucket is not None and _bucket.reduced:
        # a second head forward contributed to the same loss and its backward runs after the exchange was issued: the
        # in-flight all-reduce reads the bucket on the side stream -- writing into it now would be a race
        raise _cabi.HcompError('a gradient producer ran after the head\'s gradient all-reduce of this backward pass was issued '
                               '(two head forwards in one loss?): use DistributedDataParallel or '
                               'enable_overlapped_allreduce(fresh_grads=False) for that pattern')
    if _bucket is None:
        _bucket = _GradBucket(grp.family, device)
        torch.autograd.Variable._execution_engine.queue_callback(_bucket_finalize)
    return _bucket.segment(grp, overwrite)


def _bucket_allreduce():
    import torch.distributed as dist
    b = _bucket
    if GRAD_ALLREDUCE_GROUP is None:            # local bucket: nothing to exchange (and later producers may still write)
        return
    side = _side()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        if b.symm is not None:
            sb, h = b.symm, b.symm.hdls[b.symm_turn]
            mc = int(h.multicast_ptr) if sb.multicast else 0
            call('hcomp_allreduce_mean_symm', ptr(sb.bufs[b.symm_turn]), C.c_void_p(mc) if mc else None,
                 C.c_void_p(int(h.buffer_ptrs_dev)), C.c_void_p(int(h.signal_pad_ptrs_dev)), sb.rank, sb.world,
                 C.c_longlong(sb.padded), 0, sb.ctas, C.c_void_p(side.cuda_stream))
        else:
            dist.all_reduce(b.flat, op=dist.ReduceOp.AVG, group=GRAD_ALLREDUCE_GROUP)
    b.flat.record_stream(side)
    b.reduced = True


def collective_sms() -> int:
    """SMs the dX GEMM leaves free while the gradient exchange is in flight"""
    if _bucket is not None and _bucket.symm is not None:
        return max(COLLECTIVE_SMS, _bucket.symm.ctas)
    return COLLECTIVE_SMS


def _bucket_finalize():
    """end of the backward pass: make sure the bucket was reduced, re-join the side stream, deliver param.grad"""
    global _bucket
    b = _bucket
    if b is None:
        return
    try:
        if not b.reduced:                       # the dW GEMM did not run in this pass (frozen prototypes): reduce now
            _bucket_allreduce()
        if GRAD_ALLREDUCE_GROUP is not None:    # (local bucket: nothing ran on the side stream; waiting on it would also
            torch.cuda.current_stream().wait_stream(_side())     # pull a non-captured stream into a graph capture)
        for grp in b.produced:
            seg = b.segment(grp)
            for p_, (off, numel, shape) in zip(grp.params, grp.meta):
                if not p_.requires_grad:
                    continue
                g = seg[off:off + numel].view(shape)
                if p_.grad is None:
                    p_.grad = g
                else:
                    p_.grad.add_(g)
    finally:
        _bucket = None                          # never leave a stale bucket behind (an exception above would otherwise
                                                # make the next backward pass write into it and deliver nothing)


def reset_grad_bucket():
    """Drop a half-built gradient bucket (call after an exception escaped a backward pass in bucket mode)."""
    global _bucket
    _bucket = None


def _allreduce_grad_sync_(t: torch.Tensor):
    """autograd path: mean all-reduce on the side stream; returns the stream to wait on before `t` is handed on"""
    import torch.distributed as dist
    side = _side()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        dist.all_reduce(t, op=dist.ReduceOp.AVG, group=GRAD_ALLREDUCE_GROUP)
    t.record_stream(side)
    return side


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise _cabi.HcompError(f'{what} must live on a CUDA device: the prototype head has no CPU path')


def _cabi_tile_hdr() -> int:
    """first word of the node list inside a tile record (include/hcomp_head.h: {S, nseg, umma_n, dz_col, spill_n, spill_col0,
    spill_dst, 0, node[16], len[16], poff[16]})"""
    return 8


class DeviceLayout:
    """`HeadLayout` tables resident on one GPU + the ctypes view the C ABI takes."""

    def __init__(self, layout: HeadLayout, device):
        self.layout = layout
        self.device = torch.device(device)
        d = self.device

        def up(a, dt):
            return torch.from_numpy(np.ascontiguousarray(a)).to(device=d, dtype=dt)

        L = layout
        self.proto_off, self.cls_off, self.wc_off = up(L.proto_off, torch.int32), up(L.cls_off, torch.int32), up(L.wc_off, torch.int32)
        self.proto_node, self.col_node = up(L.proto_node, torch.int32), up(L.col_node, torch.int32)
        self.welem_col, self.welem_proto = up(L.welem_col, torch.int32), up(L.welem_proto, torch.int32)
        self.child_w = up(L.child_w, torch.float32)
        self.path_off, self.path_col = up(L.path_off, torch.int32), up(L.path_col, torch.int32)
        self.anc = up(L.anc, torch.int8)
        # leaves below every child column (static; the descendant-structured loss terms count absent leaves with it)
        col_nleaves = np.zeros(L.K, dtype=np.int32)
        for n in range(L.N):
            a = L.anc[:, n]
            for c in range(int(L.cls_off[n + 1] - L.cls_off[n])):
                col_nleaves[L.cls_off[n] + c] = int((a == c).sum())
        self.col_nleaves = up(col_nleaves, torch.int32)
        self.tiles_host = torch.from_numpy(np.ascontiguousarray(L.tiles)).to(torch.int32)
        self.tiles_dev = self.tiles_host.to(d)
        self.row_map = up(L.row_map, torch.int32)
        self.row_map_c = up(L.row_map_c, torch.int32)                # compact dZ axis (backward GEMMs)
        self.row_map_all = torch.cat([self.row_map, self.row_map_c])  # one pack launch writes both weight layouts
        self.tables = Tables(L.N, L.P, L.K, L.L, L.n_welems, L.p_max,
                             self.proto_off.data_ptr(), self.cls_off.data_ptr(), self.wc_off.data_ptr(),
                             self.proto_node.data_ptr(), self.col_node.data_ptr(), self.welem_col.data_ptr(),
                             self.welem_proto.data_ptr(), self.child_w.data_ptr(), self.path_off.data_ptr(),
                             self.path_col.data_ptr(), self.anc.data_ptr(), self.col_nleaves.data_ptr())
        self.tref = C.byref(self.tables)
        self.n_tiles = int(L.tiles.shape[0])
        # spill nodes (layout.py): host records for the C ABI; the scratch matrices are per call
        self.spill_host = torch.from_numpy(np.ascontiguousarray(L.spill)).to(torch.int32)
        self.n_spill = int(L.spill.shape[0])
        self.n_wide = int((L.spill[:, 5] == 0).sum()) if self.n_spill else 0
        self.P_s = int(L.P_s)
        self.rider_pmax = int(L.spill[:, 1].max()) if self.n_spill else 0
        # compact dZ column of every flat prototype (inverse of row_map_c): block-activity marks of the sparse backward GEMMs
        pcol = np.full(L.P, -1, dtype=np.int32)
        rmc = np.asarray(L.row_map_c)
        pcol[rmc[rmc >= 0]] = np.nonzero(rmc >= 0)[0].astype(np.int32)
        self.pcol = up(pcol, torch.int32)
        # prototype tile of every node's segment (-1: spill node): work-item activity of the backward recompute kernel
        ton = np.full(L.N, -1, dtype=np.int32)
        hdr = _cabi_tile_hdr()
        for t in range(int(L.tiles.shape[0])):
            for j in range(int(L.tiles[t, 1])):
                ton[int(L.tiles[t, hdr + j])] = t
        self.tile_of_node = up(ton, torch.int32)
        # last-block-done counter of the chained loss kernel (zero between calls; one call at a time per layout)
        self.counter = torch.zeros(8, device=d, dtype=torch.int32)

    def spill_buffers(self, M, device, zs=None, stats=None):
        """(ctypes hcomp_spill or None, zs, stats): the raw-logit scratch matrix of the spill nodes [M, P_s] and the
        row statistics of the wide ones [n_wide, M, 2]; allocated by the forward, handed back for the backward"""
        if self.n_spill == 0:
            return None, None, None
        if zs is None:
            zs = torch.empty(M, self.P_s, device=device, dtype=torch.float32)
            stats = torch.empty(max(self.n_wide, 1), M, 2, device=device, dtype=torch.float32)
        sp = _cabi.Spill(self.n_spill, self.P_s, self.spill_host.data_ptr(), zs.data_ptr(), stats.data_ptr())
        return sp, zs, stats

    # convenient aliases
    @property
    def N(self): return self.layout.N
    @property
    def P(self): return self.layout.P
    @property
    def K(self): return self.layout.K
    @property
    def L(self): return self.layout.L
    @property
    def P_pad(self): return self.layout.P_pad
    @property
    def P_c(self): return self.layout.P_c


# --------------------------------------------------------------------------- operand preparation
def feature_rows(features: torch.Tensor):
    """features [V,C,H,W] (any memory format, fp32 or bf16) -> bf16 rows [V*H*W, C] for the GEMM.
    Channels-last bf16 input (the layout ConvNeXt-26 produces) is a zero-copy view."""
    _require_cuda(features, 'features')
    V, Cc, H, W = features.shape
    HW = H * W
    if features.is_contiguous(memory_format=torch.channels_last) or Cc == 1 or HW == 1:
        nhwc = features.permute(0, 2, 3, 1)
        if features.dtype == torch.bfloat16:
            return nhwc.reshape(V * HW, Cc)
        if features.dtype == torch.float32:
            src = nhwc.reshape(V * HW, Cc)
            out = torch.empty(V * HW, Cc, device=features.device, dtype=torch.bfloat16)
            call('hcomp_cast_f32_to_bf16', ptr(src), ptr(out), C.c_longlong(src.numel()), _stream())
            return out
    if features.dtype not in (torch.float32, torch.bfloat16):
        raise _cabi.HcompError(f'unsupported feature dtype {features.dtype}')
    src = features.contiguous()
    out = torch.empty(V * HW, Cc, device=features.device, dtype=torch.bfloat16)
    call('hcomp_nchw_to_rows_bf16', ptr(src), int(src.dtype == torch.bfloat16), V, Cc, HW, ptr(out), _stream())
    return out


PREC_BF16, PREC_FP32X3 = 0, 1


class ScaleResidualRows(torch.autograd.Function):
    """Tail of the last ConvNeXt block fused into the producer of the head's feature matrix:
    features = layer_scale * keep * y + residual written ONCE, as bf16 channels-last rows -- exactly what the projection
    kernel's TMA reads (torchvision `CNBlock.forward`; the reference uses the stock model, features/convnext_features.py:18-25).
    y: the block's NHWC-in-memory output [V,C,H,W]; residual: the block's input; gamma: layer_scale [C,1,1];
    keep: per-image stochastic-depth factor [V] or None.  Returns [V,C,H,W] bf16 with channels-last strides."""

    @staticmethod
    def forward(ctx, y, residual, gamma, keep):
        _require_cuda(y, 'block output')
        V, Cc, H, W = y.shape
        yr = y if y.is_contiguous(memory_format=torch.channels_last) else y.contiguous(memory_format=torch.channels_last)
        rr = residual if residual.is_contiguous(memory_format=torch.channels_last) else residual.contiguous(memory_format=torch.channels_last)
        if yr.dtype not in (torch.float32, torch.bfloat16) or rr.dtype not in (torch.float32, torch.bfloat16):
            raise _cabi.HcompError(f'unsupported backbone dtypes {yr.dtype} / {rr.dtype}')
        g = gamma.detach().reshape(-1).float().contiguous()
        k = keep.detach().float().contiguous() if keep is not None else None
        out = torch.empty(V, H, W, Cc, device=y.device, dtype=torch.bfloat16)
        call('hcomp_scale_residual_rows_bf16', ptr(yr), int(yr.dtype == torch.bfloat16), ptr(rr), int(rr.dtype == torch.bfloat16),
             ptr(g), ptr(k), V, Cc, H * W, ptr(out), _stream())
        ctx.save_for_backward(y, gamma, keep)
        ctx.res_dtype = residual.dtype
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, grad):
        y, gamma, keep = ctx.saved_tensors
        g32 = grad.float()
        kk = keep.view(-1, 1, 1, 1).float() if keep is not None else None
        gy = gres = ggamma = None
        if ctx.needs_input_grad[0]:
            t = g32 * gamma.float().view(1, -1, 1, 1)
            gy = (t * kk if kk is not None else t).to(y.dtype)
        if ctx.needs_input_grad[1]:
            gres = grad.to(ctx.res_dtype)
        if ctx.needs_input_grad[2]:
            t = g32 * y.float()
            if kk is not None:
                t = t * kk
            ggamma = t.sum(dim=(0, 2, 3)).view_as(gamma).to(gamma.dtype)
        return gy, gres, ggamma, None


def pack_weights(w_flat: torch.Tensor, dl: DeviceLayout, precision=PREC_BF16):
    """fp32 flat kernels [P,C] -> (wp, wpc): wp = tile-padded bf16 [P_pad,C] for the fused projection kernels (3 stacked
    split planes in fp32-accurate mode), wpc = bf16 [P_c,C] on the compact dZ axis for the dX GEMM."""
    _require_cuda(w_flat, 'prototype kernels')
    assert w_flat.dtype == torch.float32 and w_flat.is_contiguous() and w_flat.shape[0] == dl.P
    Cc = w_flat.shape[1]
    if precision == PREC_FP32X3:
        wp = torch.empty(3 * dl.P_pad, Cc, device=w_flat.device, dtype=torch.bfloat16)
        call('hcomp_pack_weights_split3', ptr(w_flat), ptr(dl.row_map), dl.P_pad, Cc, ptr(wp), _stream())
        wpc = torch.empty(dl.P_c, Cc, device=w_flat.device, dtype=torch.bfloat16)
        call('hcomp_pack_weights', ptr(w_flat), ptr(dl.row_map_c), dl.P_c, Cc, ptr(wpc), _stream())
        return wp, wpc
    both = torch.empty(dl.P_pad + dl.P_c, Cc, device=w_flat.device, dtype=torch.bfloat16)
    call('hcomp_pack_weights', ptr(w_flat), ptr(dl.row_map_all), dl.P_pad + dl.P_c, Cc, ptr(both), _stream())
    return both[:dl.P_pad], both[dl.P_pad:]


def feature_rows_split3(features: torch.Tensor) -> torch.Tensor:
    """fp32 features [V,C,H,W] -> 3 stacked bf16 planes [3*V*H*W, C] (hi, mid, lo) for the fp32-accurate projection"""
    _require_cuda(features, 'features')
    V, Cc, H, W = features.shape
    src = features.float().permute(0, 2, 3, 1).contiguous().view(V * H * W, Cc)
    out = torch.empty(3 * V * H * W, Cc, device=features.device, dtype=torch.bfloat16)
    call('hcomp_split3_f32', ptr(src), ptr(out), C.c_longlong(src.numel()), _stream())
    return out


class LabelTables:
    """Per-batch tables derived from the labels (pipnet/train.py:934-937): tgt[V,N] child label or -1,
    desc[V_first,N], n_desc[N].  The tables are filled lazily: the fused head forward computes them inside its prologue
    launch (`head_prologue`); any other first reader triggers the stand-alone kernel."""

    def __init__(self, ys: torch.Tensor, dl: DeviceLayout, V_first: int):
        _require_cuda(ys, 'labels')
        ys = ys.to(torch.int64).contiguous()
        V = ys.numel()
        self.V, self.V_first = V, V_first
        self.dl = dl
        self.ys = ys                      # leaf index per row (the descendant-structured loss terms group rows by leaf)
        self._tgt = torch.empty(V, dl.N, device=ys.device, dtype=torch.int8)
        self._desc = torch.empty(V_first, dl.N, device=ys.device, dtype=torch.uint8)
        self._n_desc = torch.empty(dl.N, device=ys.device, dtype=torch.int32)
        self.pending = True

    def ensure(self):
        if self.pending:
            self.pending = False
            call('hcomp_label_tables', ptr(self.ys), self.dl.tref, self.V, self.V_first, ptr(self._tgt), ptr(self._desc),
                 ptr(self._n_desc), _stream())
        return self

    @property
    def tgt(self): return self.ensure()._tgt
    @property
    def desc(self): return self.ensure()._desc
    @property
    def n_desc(self): return self.ensure()._n_desc


SPARSE_BWD = os.environ.get('HC_SPARSE_BWD', '1') != '0'      # block-sparse dX / dW GEMMs (A/B switch, same results)
DW_SPARSE_MIN_PC = int(os.environ.get('HC_DW_SPARSE_MIN_PC', '1024'))
SKIP_UNREAD_ZERO_TILES = os.environ.get('HC_SKIP_ZERO_TILES', '1') != '0'
ITEM_SKIP = os.environ.get('HC_ITEM_SKIP', '1') != '0'        # K5 skips work items without upstream gradient (needs SPARSE_BWD)


class DzBlocks:
    """Block-activity tables of dZ (include/hcomp_head.h: hcomp_dz_blocks) for one step: allocated by the forward, cleared
    by its prologue launch, marked by whichever launch builds K5's scatter table, read by the dX / dW GEMMs."""

    def __init__(self, dl: DeviceLayout, M: int, dev, V: int = 0, V_first: int = 0):
        n_r256, n_r64 = (M + 255) // 256, (M + 63) // 64
        n_c64, n_c256 = (dl.P_c + 63) // 64, (dl.P_c + 255) // 256
        ld1, ld2 = (n_c64 + 7) // 8 * 8, (n_r64 + 7) // 8 * 8
        b1 = (n_r256 * ld1 + 15) // 16 * 16
        b2 = (n_c256 * ld2 + 15) // 16 * 16
        # work items of K5: [tile][chunk of an image pair], rows padded so that an item's 8 chunk flags are one aligned word
        HW = M // max(1, V) if V else 0
        chunks = V_first * ((HW + 31) // 32)
        pitch = (chunks + 7) // 8 * 8 + 8
        b3 = (dl.n_tiles * pitch + 15) // 16 * 16 if (ITEM_SKIP and V_first > 0) else 0
        self.buf = torch.empty(b1 + b2 + b3, device=dev, dtype=torch.uint8)
        self.struct = _cabi.DzBlocks(self.buf.data_ptr(), ld1, self.buf.data_ptr() + b1, ld2, dl.pcol.data_ptr(),
                                     (self.buf.data_ptr() + b1 + b2) if b3 else None, pitch if b3 else 0,
                                     dl.tile_of_node.data_ptr() if b3 else None, 0)
        self.ref = C.byref(self.struct)
        self.dl = dl


def head_prologue(w_flat: Optional[torch.Tensor], dl: DeviceLayout, V: int, labels: Optional["LabelTables"], dev,
                  zero_extra: Optional[torch.Tensor] = None):
    """ONE launch for everything the fused projection kernel needs (hcomp_head_prologue): bf16 weight layouts (wp, wpc),
    the cleared packed max table / align accumulators, and -- when `labels` are still pending -- the label tables."""
    wp = wpc = both = None
    rows = Cc = 0
    if w_flat is not None:
        _require_cuda(w_flat, 'prototype kernels')
        Cc = w_flat.shape[1]
        rows = dl.P_pad + dl.P_c
        both = torch.empty(rows, Cc, device=dev, dtype=torch.bfloat16)
        wp, wpc = both[:dl.P_pad], both[dl.P_pad:]
    packed = torch.empty(V * dl.P, device=dev, dtype=torch.int64)
    align_sum = torch.empty(dl.N, device=dev, dtype=torch.float64) if labels is not None else None
    lab = labels if (labels is not None and labels.pending) else None
    if lab is not None:
        lab.pending = False
    call('hcomp_head_prologue', ptr(w_flat), ptr(dl.row_map_all), rows, Cc, ptr(both), ptr(packed),
         C.c_longlong(packed.numel()), ptr(align_sum), dl.N, ptr(lab.ys) if lab is not None else None, dl.tref,
         lab.V if lab is not None else 0, lab.V_first if lab is not None else 0,
         ptr(lab._tgt) if lab is not None else None, ptr(lab._desc) if lab is not None else None,
         ptr(lab._n_desc) if lab is not None else None, ptr(zero_extra),
         C.c_longlong(zero_extra.numel() * zero_extra.element_size() if zero_extra is not None else 0), _stream())
    return wp, wpc, packed, align_sum


# --------------------------------------------------------------------------- raw kernels (no autograd)
def proj_softmax_pool_raw(x_rows, wp, dl: DeviceLayout, V, V_first, HW, tau, labels: Optional[LabelTables], thresh=0.0,
                          precision=PREC_BF16, spill_out: Optional[list] = None):
    """spill_out: a list that receives (zs, stats) of the spill nodes (needed again by `head_backward_raw`)"""
    Cc = x_rows.shape[1]
    dev = x_rows.device
    sp, zs, stats = dl.spill_buffers(V * HW, dev)
    if spill_out is not None:
        spill_out[:] = [zs, stats]
    # the packed max table and the align accumulators are merged into with atomics: cleared here (outputs_zeroed = 1)
    packed = torch.zeros(V * dl.P, device=dev, dtype=torch.int64)
    align_sum = torch.zeros(dl.N, device=dev, dtype=torch.float64) if labels is not None else None
    tok = PROFILE.start('k1_proj_softmax_pool_fwd')
    call('hcomp_proj_softmax_pool_fwd', ptr(x_rows), ptr(wp), ptr(dl.tiles_host), ptr(dl.tiles_dev), dl.n_tiles, V, V_first,
         HW, Cc, dl.P, dl.P_pad, dl.N, float(tau), int(precision), 1, ptr(labels.desc) if labels is not None else None,
         ptr(packed), ptr(align_sum), C.byref(sp) if sp is not None else None, _stream())
    PROFILE.stop(tok)
    pooled = torch.empty(V, dl.P, device=dev, dtype=torch.float32)
    argmax = torch.empty(V, dl.P, device=dev, dtype=torch.int32)
    call('hcomp_unpack_pool', ptr(packed), C.c_longlong(V * dl.P), float(thresh), ptr(pooled), ptr(argmax), _stream())
    align = None
    if labels is not None:
        align = torch.empty(dl.N, device=dev, dtype=torch.float32)
        call('hcomp_align_finalize', ptr(align_sum), ptr(labels.n_desc), dl.N, HW, ptr(align), _stream())
    return pooled, argmax, align


# Riders (narrow spill nodes) can be finished by the forward-finish launch instead of the fused kernel's tail (bit-identical
# results, tests/test_gpu_chain.py).  Measured on cub27 (profiles/r2_k1_analysis.md section 7): the fused kernel loses its
# grid barrier and tail (70.4 -> 60.8 us, 0.67 of the tensor peak) but the finish launch grows from 4.4 to 19-23 us (64
# latency-bound blocks against 1776 resident epilogue warps), a net loss of 5 us per step -- off by default.
DEFER_RIDERS = os.environ.get('HC_DEFER_RIDERS', '0') != '0'


def proj_pool_classify_raw(x_rows, wp, dl: DeviceLayout, V, V_first, HW, tau, labels: Optional[LabelTables], packed, align_sum,
                           thresh=0.0, precision=PREC_BF16, spill_out: Optional[list] = None, wc=None, bias=None):
    """K1 + ONE finishing launch (hcomp_pool_classify_fwd: unpack, align finalize, classifier).  `packed` / `align_sum`
    come cleared from `head_prologue`.  Returns pooled, argmax, align (None without labels), out (None without wc)."""
    Cc = x_rows.shape[1]
    dev = x_rows.device
    sp, zs, stats = dl.spill_buffers(V * HW, dev)
    if spill_out is not None:
        spill_out[:] = [zs, stats]
    # riders (narrow spill nodes) are finished by the finishing launch below instead of K1's tail: K1 then ends with its
    # item loop (no grid barrier, no idle wait of the CTAs that finished early)
    defer = (DEFER_RIDERS and sp is not None and dl.n_wide == 0 and dl.n_spill <= 4
             and 4 * (2 * HW * (dl.rider_pmax | 1) + 4 * HW + 1184) <= 200 * 1024)      # the finish kernel's shared memory
    tok = PROFILE.start('k1_proj_softmax_pool_fwd')
    call('hcomp_proj_softmax_pool_fwd', ptr(x_rows), ptr(wp), ptr(dl.tiles_host), ptr(dl.tiles_dev), dl.n_tiles, V, V_first,
         HW, Cc, dl.P, dl.P_pad, dl.N, float(tau), int(precision), 3 if defer else 1,
         ptr(labels.desc) if labels is not None else None, ptr(packed), ptr(align_sum),
         C.byref(sp) if sp is not None else None, _stream())
    PROFILE.stop(tok)
    global ORTH_REQUEST
    if ORTH_REQUEST is not None:           # fork the orth term's Gram kernel HERE: it runs beside the finishing launch below
        req, ORTH_REQUEST = ORTH_REQUEST, None      # (forked before K1 its blocks delay K1's persistent CTAs by ~5 us)
        orth_prefetch(*req)
    pooled = torch.empty(V, dl.P, device=dev, dtype=torch.float32)
    argmax = torch.empty(V, dl.P, device=dev, dtype=torch.int32)
    align = torch.empty(dl.N, device=dev, dtype=torch.float32) if labels is not None else None
    out = torch.empty(V, dl.K, device=dev, dtype=torch.float32) if wc is not None else None
    call('hcomp_pool_classify_fwd', ptr(packed), ptr(align_sum), ptr(labels.n_desc) if labels is not None else None, ptr(wc),
         ptr(bias), dl.tref, V, HW, float(thresh), ptr(pooled), ptr(argmax), ptr(align), ptr(out),
         C.byref(sp) if defer else None, V_first, float(tau), ptr(labels.desc) if labels is not None else None,
         ptr(dl.counter), _stream())
    return pooled, argmax, align, out


def head_backward_raw(x_rows, wp, wpc, dl: DeviceLayout, V, V_first, HW, tau, argmax, g_pooled, labels, g_align, *,
                      pooled=None, thresh=0.0, need_dx=True, need_dw=True, precision=PREC_BF16, w_group=None, dz_out=None,
                      spill=None, tables=None, blocks=None):
    """spill: (zs, stats) the forward produced for the layout's spill nodes (required when it has any);
    tables: (scat, coef) already built from exactly these g_pooled / g_align by hcomp_head_chain_bwd (`_PrepSlot`);
    blocks: `DzBlocks` of this step (zeroed, or already marked together with `tables`): the dX / dW GEMMs skip the
    unmarked blocks of dZ"""
    Cc = x_rows.shape[1]
    dev = x_rows.device
    M = V * HW
    sp = None
    if dl.n_spill:
        if spill is None or spill[0] is None:
            raise _cabi.HcompError('this layout has spill nodes: pass the (zs, stats) buffers of the forward')
        sp, _zs, _st = dl.spill_buffers(M, dev, spill[0], spill[1])
    # compact column axis (layout.row_map_c); `dz_out` lets tests supply a guarded buffer
    dz = dz_out if dz_out is not None else torch.empty(M, dl.P_c, device=dev, dtype=torch.bfloat16)
    assert dz.shape == (M, dl.P_c) and dz.dtype == torch.bfloat16 and dz.is_contiguous()
    if tables is not None:
        scat, coef = tables
        argmax = None                       # "tables are ready"
    else:
        scat = torch.empty(V * dl.P * 2, device=dev, dtype=torch.int32)
        coef = torch.empty(max(1, V_first * dl.N), device=dev, dtype=torch.float32)
    use_align = labels is not None and g_align is not None
    if blocks is not None:
        # zero tiles of dZ that neither GEMM will read are not even stored -- only when BOTH GEMMs (or the one that runs)
        # go through the tables, their CTA-pair kernels are used (more than 128 GEMM rows) and nobody else gets dZ
        dw_tables = dl.P_c >= DW_SPARSE_MIN_PC
        blocks.struct.dz_only_read_through_tables = int(SKIP_UNREAD_ZERO_TILES and dz_out is None and M > 128 and dl.P_c > 128
                                                        and (dw_tables or not need_dw) and not dl.n_spill)
    tok = PROFILE.start('k5_bwd_dz')
    call('hcomp_head_bwd_dz', ptr(x_rows), ptr(wp), ptr(dl.tiles_host), ptr(dl.tiles_dev), dl.n_tiles, V, V_first, HW, Cc,
         dl.P, dl.P_pad, dl.P_c, dl.N, float(tau), int(precision), ptr(argmax), ptr(g_pooled), ptr(pooled), float(thresh),
         ptr(labels.desc) if use_align else None, ptr(labels.n_desc) if use_align else None,
         ptr(g_align) if use_align else None, ptr(scat), ptr(coef), ptr(dz), C.byref(sp) if sp is not None else None,
         blocks.ref if blocks is not None else None, ptr(dl.proto_off), ptr(dl.proto_node), _stream())
    PROFILE.stop(tok)
    dx = dw = None
    pending = None
    bucketed = need_dw and _bucket_mode(w_group)
    if need_dw:
        # bucket mode: the segment already holds the orth-loss gradient (or zeros) and K7 accumulates on top of it
        dw = (_bucket_segment(w_group, dev).view(dl.P, Cc) if bucketed
              else torch.zeros(dl.P, Cc, device=dev, dtype=torch.float32))
        tok = PROFILE.start('k7_bwd_dw')
        # (dW's blocks are 256 compact columns = ~13 nodes wide: on small trees nearly every one is marked and the flag
        # tests only cost -- measured +2.6 us on cub27 -- so the table is used from four column tiles on)
        call('hcomp_head_bwd_dw', ptr(dz), ptr(x_rows), ptr(dl.row_map_c), C.c_longlong(M), dl.P_c, Cc, ptr(dw),
             blocks.ref if (blocks is not None and dl.P_c >= DW_SPARSE_MIN_PC) else None, _stream())
        PROFILE.stop(tok)
        if bucketed:
            _bucket_allreduce()             # everything the head produces is in the bucket by now; overlaps K6
            dw = None                       # delivered to param.grad at the end of the backward pass
        elif GRAD_ALLREDUCE_GROUP is not None:
            pending = _allreduce_grad_sync_(dw)      # overlaps K6, joined before dW is returned to autograd
    if need_dx:
        dx = torch.empty(M, Cc, device=dev, dtype=torch.bfloat16)
        tok = PROFILE.start('k6_bwd_dx')
        prev = _cabi.lib().hcomp_set_reserved_sms(collective_sms()) if (pending is not None or bucketed) else None
        try:
            call('hcomp_head_bwd_dx', ptr(dz), ptr(wpc), C.c_longlong(M), dl.P_c, Cc, ptr(dx),
                 blocks.ref if blocks is not None else None, _stream())
        finally:
            if prev is not None:
                _cabi.lib().hcomp_set_reserved_sms(prev)
        PROFILE.stop(tok)
    if pending is not None:
        torch.cuda.current_stream().wait_stream(pending)      # dW is consumed by autograd on this stream
    return dx, dw, dz


def gemm_bf16(a, b, M, N, K, a_mn, b_mn, out_mode=1, splits=1):
    """self-test hook for the tcgen05 GEMM mainloop"""
    dev = a.device
    if out_mode == 0:
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    elif out_mode == 1:
        out = torch.empty(M, N, device=dev, dtype=torch.float32)
    else:
        out = torch.zeros(M, N, device=dev, dtype=torch.float32)
    call('hcomp_gemm_bf16', ptr(a), ptr(b), M, N, K, int(a_mn), int(b_mn), out_mode, splits, ptr(out), C.c_longlong(N), _stream())
    return out


# --------------------------------------------------------------------------- autograd
class _PrepSlot:
    """Hand-off between the chained loss backward and the head backward of ONE step: the loss backward can write K5's
    scatter table / align coefficients in its own launch (it produces g_pooled and g_align anyway); the head backward uses
    them iff the gradients it receives are exactly those tensors, unmodified (pointer + version), else builds its own."""

    def __init__(self, argmax, thresh, labels, V_first, HW):
        self.argmax, self.thresh, self.labels, self.V_first, self.HW = argmax, float(thresh), labels, V_first, HW
        self.ready = None            # (g_pooled ptr, version, g_align ptr or 0, version, scat, coef)
        self.blocks = None           # DzBlocks of this step (block-sparse backward GEMMs) or None
        self.blocks_used = False     # marks are conservative only for ONE backward pass per forward

    @staticmethod
    def _key(t):
        return (0, 0) if t is None else (t.data_ptr(), t._version)

    def publish(self, g_pooled, g_align, scat, coef):
        self.ready = (self._key(g_pooled), self._key(g_align), scat, coef)

    def take(self, g_pooled, g_align):
        r, self.ready = self.ready, None
        if r is not None and r[0] == self._key(g_pooled) and r[1] == self._key(g_align):
            return r[2], r[3]
        return None


class HeadProjPool(torch.autograd.Function):
    """features, flat prototype kernels -> pooled [V,P], per-node align loss [N] (zeros without labels), argmax and --
    when the classifier weights are passed -- the child logits out [V,K] (pipnet/pipnet.py:1035-1036), finished by the
    same launch that unpacks the pooled table.
    Saves only the bf16 operands + argmax: the V x P x H x W map is recomputed tile by tile in backward.
    Three launches forward (prologue, K1, finish); backward: prep, K5, dW, dX (+ the classifier's own backward kernels
    only when a gradient arrives at `out` -- the chained losses, `HeadLosses(..., chain=True)`, never send one)."""

    @staticmethod
    def forward(ctx, features, w_flat, dl: DeviceLayout, V_first, tau, labels, thresh, precision=PREC_BF16, wc_flat=None,
                bias=None):
        V, Cc, H, W = features.shape
        HW = H * W
        dev = features.device
        wf = w_flat.detach().contiguous()
        blocks = None
        if SPARSE_BWD and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]):
            blocks = DzBlocks(dl, V * HW, dev, V, V_first)           # cleared by the prologue launch below
        zx = blocks.buf if blocks is not None else None
        if precision == PREC_FP32X3:
            x_rows = feature_rows_split3(features.detach())
            wp, wpc = pack_weights(wf, dl, precision)
            _wp, _wpc, packed, align_sum = head_prologue(None, dl, V, labels, dev, zx)
        else:
            x_rows = feature_rows(features.detach())
            wp, wpc, packed, align_sum = head_prologue(wf, dl, V, labels, dev, zx)
        wc = wc_flat.detach().contiguous() if wc_flat is not None else None
        bs = bias.detach().contiguous() if bias is not None else None
        spill = []
        pooled, argmax, align, out = proj_pool_classify_raw(x_rows, wp, dl, V, V_first, HW, tau, labels, packed, align_sum,
                                                            thresh, precision, spill, wc, bs)
        ctx.spill = spill                      # raw logits / row statistics of the spill nodes (None, None without any)
        ctx.dl, ctx.geom, ctx.labels, ctx.thresh, ctx.precision = dl, (V, V_first, H, W, Cc, tau), labels, thresh, precision
        ctx.w_group = getattr(w_flat, '_hc_group', None)
        ctx.cls_groups = (getattr(wc_flat, '_hc_group', None) if wc_flat is not None else None,
                          getattr(bias, '_hc_group', None) if bias is not None else None)
        ctx.has_cls = (wc_flat is not None, bias is not None)
        ctx.set_materialize_grads(False)      # unused outputs (argmax, align without the loss) get None, not zero fills
        ctx.feat_meta = (features.dtype, features.is_contiguous(memory_format=torch.channels_last))
        ctx.save_for_backward(x_rows, wp, wpc, argmax, pooled, wc)
        ctx.mark_non_differentiable(argmax)
        ctx.prep = pooled._hc_prep = _PrepSlot(argmax, thresh, labels, V_first, HW)
        ctx.prep.blocks = blocks
        if align is None:
            align = torch.zeros(dl.N, device=dev, dtype=torch.float32)
        return pooled, align, argmax, out

    @staticmethod
    def backward(ctx, g_pooled, g_align, _g_argmax, g_out=None):
        x_rows, wp, wpc, argmax, pooled, wc = ctx.saved_tensors
        V, V_first, H, W, Cc, tau = ctx.geom
        dl = ctx.dl
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dev = x_rows.device
        g_wc = g_bias = None
        if g_out is not None and ctx.has_cls[0]:
            # somebody differentiated through `out` directly (not the chained losses): the classifier's own backward
            need_wc, need_bias = ctx.needs_input_grad[8], ctx.has_cls[1] and ctx.needs_input_grad[9]
            wc_grp, bias_grp = ctx.cls_groups
            b_wc, b_bias = need_wc and _bucket_mode(wc_grp), need_bias and _bucket_mode(bias_grp)
            # bucket mode: the chained losses of the same backward pass may already have written these segments (both the
            # loss and another consumer of `out` in one graph): then this contribution is added, not written over
            add_wc = b_wc and _bucket is not None and wc_grp in _bucket.produced
            add_bias = b_bias and _bucket is not None and bias_grp in _bucket.produced
            if need_wc:
                g_wc = _bucket_segment(wc_grp, dev, True) if (b_wc and not add_wc) else torch.empty_like(wc)
            if need_bias:
                g_bias = (_bucket_segment(bias_grp, dev, True) if (b_bias and not add_bias)
                          else torch.empty(dl.K, device=dev, dtype=torch.float32))
            accumulate = g_pooled is not None
            g_pooled = g_pooled.contiguous().float().clone() if accumulate else torch.empty(V, dl.P, device=dev, dtype=torch.float32)
            call('hcomp_classifier_bwd', ptr(g_out.contiguous().float()), ptr(pooled), ptr(wc), dl.tref, V, ptr(g_pooled),
                 int(accumulate), ptr(g_wc), ptr(g_bias), _stream())
            if add_wc:
                _bucket_segment(wc_grp, dev).add_(g_wc)
            if add_bias:
                _bucket_segment(bias_grp, dev).add_(g_bias)
            if b_wc:
                g_wc = None
            if b_bias:
                g_bias = None
            if GRAD_ALLREDUCE_GROUP is not None:
                for g in (g_wc, g_bias):
                    if g is not None:
                        torch.cuda.current_stream().wait_stream(_allreduce_grad_sync_(g))
        if g_pooled is None:
            g_pooled = torch.zeros(V, dl.P, device=dev, dtype=torch.float32)
        g_pooled = g_pooled.contiguous().float()
        if g_align is not None:
            g_align = g_align.contiguous().float()
        dx, dw, _ = head_backward_raw(x_rows, wp, wpc, dl, V, V_first, H * W, tau, argmax, g_pooled, ctx.labels, g_align,
                                      pooled=pooled, thresh=ctx.thresh, need_dx=need_dx, need_dw=need_dw,
                                      precision=ctx.precision, w_group=ctx.w_group, spill=ctx.spill,
                                      tables=ctx.prep.take(g_pooled, g_align), blocks=ctx.prep.blocks)
        d_feat = None
        if need_dx:
            dtype, _cl = ctx.feat_meta
            d_feat = dx.view(V, H, W, Cc).permute(0, 3, 1, 2)     # channels-last view of the row buffer
            if dtype != torch.bfloat16:
                d_feat = d_feat.to(dtype)
        return d_feat, dw, None, None, None, None, None, None, g_wc, g_bias


class NonNegClassifier(torch.autograd.Function):
    """out[V,K] = pooled . relu(Wc)^T per node (pipnet/pipnet.py:1035-1036) on the flat axes."""

    @staticmethod
    def forward(ctx, pooled, wc_flat, bias, dl: DeviceLayout):
        V = pooled.shape[0]
        pooled = pooled.contiguous()
        wc_flat = wc_flat.contiguous()
        out = torch.empty(V, dl.K, device=pooled.device, dtype=torch.float32)
        call('hcomp_classifier_fwd', ptr(pooled), ptr(wc_flat), ptr(bias), dl.tref, V, ptr(out), _stream())
        ctx.dl = dl
        ctx.save_for_backward(pooled, wc_flat)
        ctx.has_bias = bias is not None
        ctx.groups = (getattr(wc_flat, '_hc_group', None), getattr(bias, '_hc_group', None) if bias is not None else None)
        return out

    @staticmethod
    def backward(ctx, g_out):
        pooled, wc_flat = ctx.saved_tensors
        dl = ctx.dl
        V = pooled.shape[0]
        g_out = g_out.contiguous()
        g_pooled = torch.empty_like(pooled) if ctx.needs_input_grad[0] else None
        need_wc, need_bias = ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        wc_grp, bias_grp = ctx.groups
        b_wc, b_bias = need_wc and _bucket_mode(wc_grp), need_bias and _bucket_mode(bias_grp)
        g_wc = g_bias = None
        if need_wc:
            g_wc = _bucket_segment(wc_grp, pooled.device, True) if b_wc else torch.empty_like(wc_flat)
        if need_bias:
            g_bias = (_bucket_segment(bias_grp, pooled.device, True) if b_bias
                      else torch.empty(dl.K, device=pooled.device, dtype=torch.float32))
        call('hcomp_classifier_bwd', ptr(g_out), ptr(pooled), ptr(wc_flat), dl.tref, V, ptr(g_pooled), 0, ptr(g_wc),
             ptr(g_bias), _stream())
        if b_wc:
            g_wc = None                     # reduced with the bucket, delivered at the end of the backward pass
        if b_bias:
            g_bias = None
        if GRAD_ALLREDUCE_GROUP is not None:        # autograd path: one flat mean all-reduce for all nodes' classifiers
            for g in (g_wc, g_bias):
                if g is not None:
                    torch.cuda.current_stream().wait_stream(_allreduce_grad_sync_(g))
        return g_pooled, g_wc, g_bias, None


LOSS_TANH, LOSS_ORTH, LOSS_CLASS, LOSS_SPARSITY = 1, 2, 4, 8


LOSS_ORTH_READY = 16
ORTH_REQUEST = None       # (w_flat, wc_flat, dl, V): the model asks the next fused head forward to start `orth_prefetch` after K1
_orth_stream = None
_orth_slot = None          # (key, ws, rel, event) of the latest `orth_prefetch`


def _chain_ws(dl: DeviceLayout, V: int, dev):
    n = int(_cabi.lib().hcomp_head_chain_ws_floats(dl.tref, V))
    return torch.empty(n, device=dev, dtype=torch.float32), torch.empty(dl.P, device=dev, dtype=torch.uint8)


def _orth_key(w_flat, wc_flat, dl, V):
    return (w_flat.data_ptr(), w_flat._version, wc_flat.data_ptr(), wc_flat._version, id(dl), int(V))


def orth_prefetch(w_flat: torch.Tensor, wc_flat: torch.Tensor, dl: DeviceLayout, V: int):
    """Start the weights-only part of the kernel-orthogonality term (Gram matrices + relevance mask;
    pipnet/train.py:1136-1151) on a side stream NOW, off the critical path between the forward and the backward.
    `HeadLosses` picks the result up (and always re-joins the side stream).  The model requests it (`ORTH_REQUEST`) for
    the head forward when the previous step's loss used the term; the fused forward issues it right after K1."""
    global _orth_stream, _orth_slot
    if w_flat.shape[1] <= dl.layout.p_max:
        return
    wf, wc = w_flat.detach().contiguous(), wc_flat.detach().contiguous()
    if _orth_stream is None:
        _orth_stream = torch.cuda.Stream()
    ws, rel = _chain_ws(dl, V, wf.device)
    side = _orth_stream
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        call('hcomp_orth_gram', ptr(wf), ptr(wc), dl.tref, wf.shape[1], ptr(ws), ptr(rel), C.c_void_p(side.cuda_stream))
        ev = torch.cuda.Event()
        ev.record(side)
    ws.record_stream(side)
    rel.record_stream(side)
    _orth_slot = (_orth_key(w_flat, wc_flat, dl, V), ws, rel, ev)


def _orth_take(key):
    """join the prefetch branch (always) and return its (ws, rel) when it matches `key`"""
    global _orth_slot
    slot, _orth_slot = _orth_slot, None
    if slot is None:
        return None
    torch.cuda.current_stream().wait_event(slot[3])
    return (slot[1], slot[2]) if slot[0] == key else None


class HeadLosses(torch.autograd.Function):
    """All per-node loss terms of the shipped recipe in one forward / one backward call:
    total = sum_n (w_align*align[n] + w_tanh*tanh[n] + w_orth*orth[n] + w_class*class[n]) (weights include the 1/N
    of `pipnet/train.py:1071,1084,1149,1165`).  Returns (total, stats[4,N], n_correct[N]); only `total` is
    differentiable.  The relevance mask of the orth term comes from the classifier weights, which get no
    gradient from it (boolean indexing in the reference, `pipnet/train.py:1140`).

    chain=True (what `calculate_loss` uses when `out` is the model's own classifier output of `pooled`): the forward is ONE
    launch, and the backward -- ONE launch beside the orth term's -- goes straight THROUGH the classifier: it returns the
    gradients of `pooled` (tanh term + class term through relu(Wc)), `wc_flat` and `bias`, and nothing for `out` (so the
    classifier's own autograd node has nothing to do).  chain=False: `out` is an independent input with its own gradient."""

    @staticmethod
    def forward(ctx, pooled, out, align, w_flat, wc_flat, labels: LabelTables, dl: DeviceLayout, flags, weights, eps,
                multiplier=2.0, bias=None, chain=False):
        # multiplier: exponent of the class term's log1p(out ** m) (net._multiplier, pipnet/train.py:1158)
        pooled, out = pooled.contiguous(), out.contiguous()
        V = pooled.shape[0]
        dev = pooled.device
        use_orth = bool(flags & LOSS_ORTH)
        Cc = w_flat.shape[1] if w_flat is not None else 0
        wts = (C.c_float * 4)(*[float(x) for x in weights])
        total = torch.empty((), device=dev, dtype=torch.float32)
        stats = torch.empty(4, dl.N, device=dev, dtype=torch.float32)
        n_correct = torch.empty(dl.N, device=dev, dtype=torch.int32)
        wf = w_flat.detach().contiguous() if use_orth else None
        wc = wc_flat.detach().contiguous() if wc_flat is not None else None
        al = align.detach().contiguous() if align is not None else None
        ready = _orth_take(_orth_key(w_flat, wc_flat, dl, V) if (use_orth and chain and wc_flat is not None) else None)
        if chain:
            fl = int(flags)
            if ready is not None:
                ws, rel = ready
                fl |= LOSS_ORTH_READY
            else:
                ws, rel = _chain_ws(dl, V, dev)
            call('hcomp_head_chain_fwd', ptr(pooled), ptr(out), ptr(al), ptr(wf), ptr(wc), ptr(labels.tgt), ptr(labels.n_desc),
                 dl.tref, V, labels.V_first, Cc, fl, wts, float(eps), float(multiplier), ptr(total), ptr(stats),
                 ptr(n_correct), ptr(ws), ptr(rel), ptr(dl.counter), _stream())
        else:
            ws = torch.empty(int(_cabi.lib().hcomp_head_losses_ws_floats(dl.tref)), device=dev, dtype=torch.float32)
            rel = torch.empty(dl.P, device=dev, dtype=torch.uint8)
            call('hcomp_head_losses_fwd', ptr(pooled), ptr(out), ptr(al), ptr(wf), ptr(wc), ptr(labels.tgt), ptr(labels.n_desc),
                 dl.tref, V, labels.V_first, Cc, int(flags), wts, float(eps), float(multiplier), ptr(total), ptr(stats),
                 ptr(n_correct), ptr(ws), ptr(rel), _stream())
        ctx.dl, ctx.labels, ctx.cfg = dl, labels, (int(flags), [float(x) for x in weights], float(eps), V, Cc, float(multiplier))
        ctx.has = (align is not None, w_flat is not None and use_orth)
        ctx.chain = bool(chain)
        ctx.prep = getattr(pooled, '_hc_prep', None) if chain else None
        ctx.w_group = getattr(w_flat, '_hc_group', None) if w_flat is not None else None
        ctx.cls_groups = (getattr(wc_flat, '_hc_group', None) if wc_flat is not None else None,
                          getattr(bias, '_hc_group', None) if bias is not None else None)
        ctx.has_bias = bias is not None
        ctx.set_materialize_grads(False)      # stats / n_correct are not differentiable: no zero-filled grads for them
        ctx.save_for_backward(out, wf if use_orth else None, stats, ws, rel, pooled if chain else None, wc if chain else None)
        ctx.mark_non_differentiable(stats, n_correct)
        return total, stats, n_correct

    @staticmethod
    def backward(ctx, g_total, _gs, _gc):
        if g_total is None:
            return (None,) * 13
        out, wf, stats, ws, rel, pooled, wc = ctx.saved_tensors
        dl, labels = ctx.dl, ctx.labels
        flags, weights, eps, V, Cc, multiplier = ctx.cfg
        dev = out.device
        wts = (C.c_float * 4)(*weights)
        g_total = g_total.contiguous().float()
        need_pooled, need_out, need_align, need_w = (ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                                                     ctx.needs_input_grad[2] and ctx.has[0],
                                                     ctx.needs_input_grad[3] and ctx.has[1])
        g_pooled = torch.empty(V, dl.P, device=dev, dtype=torch.float32) if need_pooled else None
        bucketed = need_w and _bucket_mode(ctx.w_group)
        g_w = None
        if need_w:      # bucket mode: the orth gradient (identical on every rank) lands where K7 will accumulate dW
            g_w = (_bucket_segment(ctx.w_group, dev, True).view(dl.P, Cc) if bucketed
                   else torch.empty(dl.P, Cc, device=dev, dtype=torch.float32))
        g_out = g_wc = g_bias = g_align = None
        if ctx.chain:
            need_wc, need_bias = ctx.needs_input_grad[4], ctx.has_bias and ctx.needs_input_grad[11]
            wc_grp, bias_grp = ctx.cls_groups
            b_wc, b_bias = need_wc and _bucket_mode(wc_grp), need_bias and _bucket_mode(bias_grp)
            if need_wc:
                g_wc = _bucket_segment(wc_grp, dev, True) if b_wc else torch.empty_like(wc)
            if need_bias:
                g_bias = _bucket_segment(bias_grp, dev, True) if b_bias else torch.empty(dl.K, device=dev, dtype=torch.float32)
            g_align = torch.empty(dl.N, device=dev, dtype=torch.float32) if need_align else None
            # K5's scatter table / align coefficients from the same launch (used by the head backward iff these very
            # gradients reach it, `_PrepSlot`)
            slot = ctx.prep if (need_pooled and ctx.prep is not None and ctx.prep.labels is labels
                                and ctx.prep.argmax.shape == g_pooled.shape) else None
            scat = coef = None
            if slot is not None:
                scat = torch.empty(V * dl.P * 2, device=dev, dtype=torch.int32)
                coef = torch.empty(max(1, labels.V_first * dl.N), device=dev, dtype=torch.float32)
            use_coef = slot is not None and g_align is not None
            call('hcomp_head_chain_bwd', ptr(g_total), ptr(pooled), ptr(out), ptr(wf), ptr(wc), ptr(labels.tgt),
                 ptr(labels.n_desc), ptr(stats), dl.tref, V, labels.V_first, Cc, flags, wts, eps, multiplier, ptr(ws), ptr(rel),
                 ptr(g_pooled), ptr(g_wc), ptr(g_bias), ptr(g_align), ptr(g_w),
                 ptr(slot.argmax) if slot is not None else None, slot.thresh if slot is not None else 0.0,
                 ptr(labels.desc) if use_coef else None, slot.HW if slot is not None else 0, ptr(scat),
                 ptr(coef) if use_coef else None,
                 slot.blocks.ref if (slot is not None and slot.blocks is not None) else None, _stream())
            if slot is not None:
                slot.publish(g_pooled, g_align, scat, coef)
            if b_wc:
                g_wc = None                     # reduced with the bucket, delivered at the end of the backward pass
            if b_bias:
                g_bias = None
            if GRAD_ALLREDUCE_GROUP is not None:        # autograd path: one flat mean all-reduce for all nodes' classifiers
                for g in (g_wc, g_bias):
                    if g is not None:
                        torch.cuda.current_stream().wait_stream(_allreduce_grad_sync_(g))
        else:
            gvec = torch.empty(4, dl.N, device=dev, dtype=torch.float32)
            g_out = torch.empty(V, dl.K, device=dev, dtype=torch.float32) if need_out else None
            call('hcomp_head_losses_bwd', ptr(g_total), ptr(out), ptr(wf), ptr(labels.tgt), ptr(labels.n_desc), ptr(stats), dl.tref,
                 V, labels.V_first, Cc, flags, wts, eps, multiplier, ptr(ws), ptr(rel), ptr(gvec), ptr(g_pooled), ptr(g_out),
                 ptr(g_w), _stream())
            g_align = gvec[0] if need_align else None
        if bucketed:
            g_w = None
        elif g_w is not None and GRAD_ALLREDUCE_GROUP is not None:
            # the orth term is skipped for nodes without a descendant in the LOCAL batch (pipnet/train.py:941-942), so
            # its gradient differs across ranks like any other and needs the mean too
            torch.cuda.current_stream().wait_stream(_allreduce_grad_sync_(g_w))
        return g_pooled, g_out, g_align, g_w, g_wc, None, None, None, None, None, None, g_bias, None

DESC_TANH_DESC, DESC_CONTRAST, DESC_MASK_PRUNE, DESC_GEOMETRIC, DESC_SG_SCORE = 1, 2, 4, 8, 16


def gumbel_noise(dl: DeviceLayout, device) -> torch.Tensor:
    """Gumbel(0,1) noise for the mask-pruning term, one pair per classifier weight element (node, child, prototype):
    the same distribution `F.gumbel_softmax` draws from inside the reference's child loop (pipnet/train.py:978)."""
    return -torch.empty(dl.layout.n_welems, 2, device=device, dtype=torch.float32).exponential_().log()


class DescLosses(torch.autograd.Function):
    """tanh_desc + minimize_contrasting_set + mask-prune overspecificity (pipnet/train.py:946-1060, 1089-1133) for
    all nodes in one forward / one backward call.  weights = (tanh_desc_w/N, contrast_w/N, 2.0/N, 0.5/N).
    Returns (loss, stats[4,N]); gradients flow to `pooled` and the presence logits only (the classifier weights
    enter through index selections in the reference and get none)."""

    @staticmethod
    def forward(ctx, pooled, wc_flat, presence, gumbel, labels: LabelTables, dl: DeviceLayout, flags, weights, eps,
                boost, tau):
        pooled = pooled.contiguous()
        V, dev = pooled.shape[0], pooled.device
        wc = wc_flat.detach().contiguous()
        pres = presence.detach().contiguous().float() if presence is not None else None
        gum = gumbel.detach().contiguous().float() if gumbel is not None else None
        if (flags & DESC_MASK_PRUNE) and (pres is None or gum is None or tuple(pres.shape) != (dl.P, 2)
                                         or tuple(gum.shape) != (dl.layout.n_welems, 2)):
            raise _cabi.HcompError('mask pruning needs presence [P,2] and gumbel noise [n_welems,2]')
        wts = (C.c_float * 4)(*[float(x) for x in weights])
        ws = torch.empty(int(_cabi.lib().hcomp_desc_losses_ws_bytes(dl.tref, V)), device=dev, dtype=torch.uint8)
        stats = torch.empty(4, dl.N, device=dev, dtype=torch.float32)
        loss = torch.empty((), device=dev, dtype=torch.float32)
        boost = float(boost) if boost else 0.0
        call('hcomp_desc_losses_fwd', ptr(pooled), ptr(wc), ptr(pres), ptr(gum), ptr(labels.ys), ptr(labels.tgt),
             ptr(labels.n_desc), dl.tref, V, labels.V_first, int(flags), wts, float(eps), boost, float(tau), ptr(ws),
             ptr(stats), ptr(loss), _stream())
        ctx.dl, ctx.labels = dl, labels
        ctx.cfg = (int(flags), [float(x) for x in weights], float(eps), boost, float(tau), V)
        ctx.set_materialize_grads(False)
        ctx.has_presence = presence is not None
        ctx.pp_group = getattr(presence, '_hc_group', None) if presence is not None else None
        ctx.save_for_backward(pooled, wc, pres, gum, ws)
        ctx.mark_non_differentiable(stats)
        return loss, stats

    @staticmethod
    def backward(ctx, g_loss, _gs):
        if g_loss is None:
            return (None,) * 11
        pooled, wc, pres, gum, ws = ctx.saved_tensors
        dl, labels = ctx.dl, ctx.labels
        flags, weights, eps, boost, tau, V = ctx.cfg
        dev = pooled.device
        wts = (C.c_float * 4)(*weights)
        g_loss = g_loss.contiguous().float()
        g_pooled = torch.empty(V, dl.P, device=dev, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        need_pres = ctx.has_presence and ctx.needs_input_grad[2]
        bucketed = need_pres and _bucket_mode(ctx.pp_group)
        g_pres = None
        if need_pres:
            g_pres = (_bucket_segment(ctx.pp_group, dev, True).view(dl.P, 2) if bucketed
                      else torch.empty(dl.P, 2, device=dev, dtype=torch.float32))
        call('hcomp_desc_losses_bwd', ptr(g_loss), ptr(pooled), ptr(wc), ptr(pres), ptr(gum), ptr(labels.ys), ptr(labels.tgt),
             ptr(labels.n_desc), dl.tref, V, labels.V_first, flags, wts, eps, boost, tau, ptr(ws), ptr(g_pooled), ptr(g_pres),
             _stream())
        if bucketed:
            g_pres = None
        elif GRAD_ALLREDUCE_GROUP is not None and g_pres is not None:
            torch.cuda.current_stream().wait_stream(_allreduce_grad_sync_(g_pres))
        return g_pooled, None, g_pres, None, None, None, None, None, None, None, None


def joint_leaf_distribution(out_flat: torch.Tensor, dl: DeviceLayout, tau=1.0, override: Optional[torch.Tensor] = None):
    """[V,K] child logits -> ([V,L] joint leaf probabilities in sorted-leaf order, [V] argmax)
    (util/node.py:300-385 + pipnet/pipnet.py:173-185 as one pass over a flattened path table).
    override: optional float [K]; nodes whose first column is >= 0 use these child probabilities for every sample."""
    out_flat = out_flat.detach().contiguous()
    if override is not None:
        override = override.detach().to(device=out_flat.device, dtype=torch.float32).contiguous()
        assert override.numel() == dl.K
    V = out_flat.shape[0]
    dev = out_flat.device
    probs = torch.empty(V, dl.K, device=dev, dtype=torch.float32)
    joint = torch.empty(V, dl.L, device=dev, dtype=torch.float32)
    pred = torch.empty(V, device=dev, dtype=torch.int64)
    call('hcomp_joint_leaf', ptr(out_flat), dl.tref, V, float(tau), ptr(override), ptr(probs), ptr(joint), ptr(pred), _stream())
    return joint, pred


def materialize_map(features: torch.Tensor, w_node: torch.Tensor, tau=1.0) -> torch.Tensor:
    """Full softmax map [V,P_n,H,W] of ONE node (visualisation path only, util/vis_hpipnet.py:62-127)."""
    V, Cc, H, W = features.shape
    x_rows = feature_rows(features.detach())
    w = w_node.detach().reshape(w_node.shape[0], -1).contiguous().float()
    out = torch.empty(V, w.shape[0], H * W, device=features.device, dtype=torch.float32)
    call('hcomp_materialize_map', ptr(x_rows), ptr(w), V, H * W, Cc, w.shape[0], float(tau), ptr(out), _stream())
    return out.view(V, w.shape[0], H, W)
