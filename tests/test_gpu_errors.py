"""Error behaviour and edge cases at the C-ABI boundary: bad arguments come back as HCOMP_E_* with a message (never a
crash, never a silent CPU fallback); degenerate but legal batches still match the oracle."""
import ctypes as C

import pytest
import torch

from oracle import head_oracle as ho
from oracle.problems import Problem, rel_err

pytestmark = pytest.mark.gpu


def _setup(tree="cub08", C_=64, H=6, B=3, **kw):
    from pipnet_b200 import ops
    pb = Problem(tree, C_, H, B, seed=2, num_features=kw.pop('num_features', 20), **kw)
    dl = ops.DeviceLayout(pb.layout, 'cuda')
    return ops, pb, dl


def test_cpu_tensors_are_rejected():
    from pipnet_b200 import ops
    from pipnet_b200._cabi import HcompError
    ops_, pb, dl = _setup()
    with pytest.raises(HcompError):
        ops.feature_rows(pb.features('cpu'))
    with pytest.raises(HcompError):
        ops.pack_weights(pb.w_flat('cpu'), dl)
    with pytest.raises(HcompError):
        ops.LabelTables(pb.ys, dl, pb.V_first)


@pytest.mark.parametrize("bad", ["hw", "c", "vfirst", "tau", "ppad"])
def test_bad_arguments_return_error_codes(bad):
    from pipnet_b200 import _cabi
    from pipnet_b200._cabi import HcompError, call, ptr
    ops, pb, dl = _setup()
    V, HW, Cc = pb.V, pb.H * pb.H, pb.C
    xr = ops.feature_rows(pb.features('cuda'))
    wp, _ = ops.pack_weights(pb.w_flat('cuda'), dl)
    packed = torch.zeros(V * dl.P, device='cuda', dtype=torch.int64)
    kw = dict(V=V, V_first=pb.V_first, HW=HW, C=Cc, P_pad=dl.P_pad, tau=1.0)
    kw.update({'hw': dict(HW=0), 'c': dict(C=Cc + 4), 'vfirst': dict(V_first=V + 1), 'tau': dict(tau=0.0),
               'ppad': dict(P_pad=dl.P_pad + 128)}[bad])
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    with pytest.raises(HcompError) as ei:
        call('hcomp_proj_softmax_pool_fwd', ptr(xr), ptr(wp), ptr(dl.tiles_host), ptr(dl.tiles_dev), dl.n_tiles, kw['V'],
             kw['V_first'], kw['HW'], kw['C'], dl.P, kw['P_pad'], dl.N, float(kw['tau']), 0, 1, None, ptr(packed), None, None, stream)
    assert 'failed (-1)' in str(ei.value) and len(_cabi.lib().hcomp_last_error()) > 0
    torch.cuda.synchronize()                                  # the context is still healthy
    pooled, argmax, _ = ops.proj_softmax_pool_raw(xr, wp, dl, V, pb.V_first, HW, 1.0, None)
    assert float(pooled.min()) > 0


def test_unknown_precision_and_segment_class():
    from pipnet_b200._cabi import HcompError, call, ptr
    ops, pb, dl = _setup()
    V, HW = pb.V, pb.H * pb.H
    xr = ops.feature_rows(pb.features('cuda'))
    wp, _ = ops.pack_weights(pb.w_flat('cuda'), dl)
    packed = torch.zeros(V * dl.P, device='cuda', dtype=torch.int64)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    with pytest.raises(HcompError):
        call('hcomp_proj_softmax_pool_fwd', ptr(xr), ptr(wp), ptr(dl.tiles_host), ptr(dl.tiles_dev), dl.n_tiles, V, pb.V_first,
             HW, pb.C, dl.P, dl.P_pad, dl.N, 1.0, 7, 1, None, ptr(packed), None, None, stream)
    tiles = dl.tiles_host.clone()
    tiles[0, 0] = 24                                          # not an instantiated segment class
    with pytest.raises(HcompError):
        call('hcomp_proj_softmax_pool_fwd', ptr(xr), ptr(wp), ptr(tiles), ptr(dl.tiles_dev), dl.n_tiles, V, pb.V_first,
             HW, pb.C, dl.P, dl.P_pad, dl.N, 1.0, 0, 1, None, ptr(packed), None, None, stream)


@pytest.mark.parametrize("labels", ["all_same", "single_pair"])
def test_degenerate_batches_match_oracle(labels):
    """every image of the batch has the same leaf (most nodes have no descendant) / a batch of ONE image pair"""
    from pipnet_b200 import train as tr
    from oracle.problems import bf16_round, build_net, make_args
    args = make_args(num_features=20, tanh_desc='y|0.05', minimize_contrasting_set='y')
    net, root = build_net("cub18", 64, args)
    names = net.layout.node_names
    B = 1 if labels == "single_pair" else 4
    g = torch.Generator().manual_seed(9)
    x = bf16_round(torch.randn(2 * B, 64, 6, 6, generator=g))
    ys = torch.full((2 * B,), 3, dtype=torch.long)
    xs = x.cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    lab = tr.make_labels(net, ys.cuda())
    f, pf, pooled, out = net(xs, labels=lab)
    w = tr._phase_weights(False, 2, 10, args)
    res = tr.calculate_loss(2, net, {}, f, pf, pooled, out, ys.cuda(), net_normalization_multiplier=net._multiplier,
                            pretrain=False, finetune=False, criterion=None, train_iter=None, print=False, EPS=1e-8, root=root,
                            kernel_orth=True, tanh_desc=True, align=False, uni=False, align_pf=True, tanh=True, args=args,
                            device='cuda', labels=lab, **w)
    res[0].backward()
    torch.cuda.synchronize()
    aw = {n: getattr(net, '_' + n + '_add_on').weight.detach().flatten(1).double().cpu() for n in names}
    cw = {n: getattr(net, '_' + n + '_classification').weight.detach().double().cpu() for n in names}
    ref = ho.full_step(x.double(), aw, cw, root, ys, {i: n for i, n in enumerate(net.layout.leaf_names)}, pretrain=False,
                       finetune=False, epoch=2, nr_epochs=10, cl_weight=args.cl_weight, tanh_desc_weight=0.05, contrasting=0.1)
    assert abs(float(res[0].detach()) - float(ref['loss'])) <= 1e-5 * max(1.0, abs(float(ref['loss'])))
    assert set(k for k, v in res[1].items()) == set(ref['cls'])            # only nodes above the one leaf contribute
    assert rel_err(xs.grad, ref['grad_x']) <= 2e-2
