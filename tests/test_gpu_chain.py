"""Fused chains (ABI v6): the multi-role launches -- head prologue, forward finish, chained losses forward / backward,
orth prefetch -- against the separate entry points they replace, on the same inputs.  Element-wise roles must be
bit-identical; block sums may associate differently (1e-6 relative)."""
import ctypes as C

import pytest
import torch

from pipnet_b200.fixtures import make_args, build_net

pytestmark = pytest.mark.gpu


def _close(a, b, tol=2e-6):
    a, b = a.double(), b.double()
    return float((a - b).abs().max()) <= tol * max(1e-6, float(b.abs().max()))


def _problem(tree="cub27", C_=64, H=6, B=6, bias=False, **over):
    from pipnet_b200 import train as tr
    over.setdefault('num_features', 20)
    args = make_args(bias=bias, **over)
    net, root = build_net(tree, C_, args)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2 * B, C_, H, H, generator=g)
    ys = torch.randint(0, net.layout.L, (B,), generator=g)
    ys = torch.cat([ys, ys]).cuda()
    xs = x.cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    return net, root, args, xs, ys, tr


def test_prologue_matches_separate_calls():
    from pipnet_b200 import ops
    net, root, args, xs, ys, tr = _problem()
    dl = net.device_layout('cuda')
    w = net.flat_prototype_kernels().detach().contiguous()
    V = xs.shape[0]
    lab_a = ops.LabelTables(ys, dl, V // 2)
    wp, wpc, packed, align_sum = ops.head_prologue(w, dl, V, lab_a, xs.device)
    assert not lab_a.pending
    lab_b = ops.LabelTables(ys, dl, V // 2).ensure()
    wp0, wpc0 = ops.pack_weights(w, dl)
    torch.cuda.synchronize()
    assert torch.equal(wp.view(torch.int16), wp0.view(torch.int16)) and torch.equal(wpc.view(torch.int16), wpc0.view(torch.int16))
    assert int(packed.abs().max()) == 0 and float(align_sum.abs().max()) == 0.0
    assert torch.equal(lab_a._tgt, lab_b._tgt) and torch.equal(lab_a._desc, lab_b._desc) and torch.equal(lab_a._n_desc, lab_b._n_desc)
    # odd table length (scalar tail of the 16-byte clears) and labels already computed
    dirty = torch.full((V * dl.P + 1,), -1, device='cuda', dtype=torch.int64)
    extra = torch.full((48,), 7, device='cuda', dtype=torch.uint8)          # the extra clear target (16-byte units)
    ops.call('hcomp_head_prologue', None, None, 0, 0, None, ops.ptr(dirty), C.c_longlong(V * dl.P - 1), None, 0, None, dl.tref,
             0, 0, None, None, None, ops.ptr(extra), C.c_longlong(32), ops._stream())
    torch.cuda.synchronize()
    assert int(dirty[:V * dl.P - 1].abs().max()) == 0 and int(dirty[V * dl.P - 1]) == -1
    assert int(extra[:32].max()) == 0 and int(extra[32:].min()) == 7


@pytest.mark.parametrize("bias", [False, True], ids=["nobias", "bias"])
@pytest.mark.parametrize("inference", [False, True], ids=["train", "inference"])
def test_forward_finish_matches_separate_calls(bias, inference):
    from pipnet_b200 import ops
    net, root, args, xs, ys, tr = _problem(bias=bias)
    if bias:
        with torch.no_grad():
            for p in net._bias_group.params:
                p.uniform_(-0.5, 0.5)
    labels = tr.make_labels(net, ys)
    with torch.no_grad():
        _, _, pooled, out = net(xs, inference=inference, labels=labels)           # fused finish
        dl = net.device_layout('cuda')
        p0, al0, am0, _ = net.head(xs, inference=inference, labels=labels)         # same head, separate classifier kernel
        out0 = net.classify(p0, dl)
    torch.cuda.synchronize()
    assert torch.equal(pooled.flat, p0) and torch.equal(pooled.align, al0)
    assert torch.equal(out.flat, out0)


@pytest.mark.parametrize("bias", [False, True], ids=["nobias", "bias"])
@pytest.mark.parametrize("phase", [("pretrain", True, False), ("train", False, False), ("finetune", False, True)],
                         ids=lambda p: p[0])
def test_chained_losses_match_unchained(phase, bias):
    """calculate_loss with the chained backward (one launch through the classifier) against the same loss with `out`
    as an independent autograd input (separate loss kernels + the classifier's own backward + autograd's adds)"""
    from pipnet_b200 import ops
    _, pretrain, finetune = phase
    net, root, args, xs, ys, tr = _problem(bias=bias, tree="cub18", C_=128, H=7, B=5, num_features=12)
    if bias:
        with torch.no_grad():
            for p in net._bias_group.params:
                p.uniform_(-0.5, 0.5)
    w = tr._phase_weights(pretrain, 3, 10, args)
    names = net.layout.node_names
    results = []
    for chained in (True, False):
        net.zero_grad(set_to_none=True)
        x = xs.clone().requires_grad_(True)
        labels = tr.make_labels(net, ys)
        features, pf, pooled, out = net(x, labels=labels)
        if not chained:
            out.chained_from = None
        res = tr.calculate_loss(3, net, {}, features, pf, pooled, out, ys, net_normalization_multiplier=net._multiplier,
                                pretrain=pretrain, finetune=finetune, criterion=None, train_iter=None, print=False, EPS=1e-8,
                                root=root, kernel_orth=True, align=False, uni=False, align_pf=True, tanh=True, args=args,
                                device='cuda', labels=labels, **w)
        res[0].backward()
        torch.cuda.synchronize()
        grads = {'x': x.grad.float() if x.grad is not None else None}
        for pname, prm in net.named_parameters():
            if not pname.startswith('_net.'):
                grads[pname] = prm.grad.clone() if prm.grad is not None else None
        results.append((float(res[0]), res.stats.clone(), grads))
    (la, sa, ga), (lb, sb, gb) = results
    assert abs(la - lb) <= 2e-6 * max(1.0, abs(lb))
    assert _close(sa, sb)
    assert set(ga) == set(gb)
    for k in ga:
        assert (ga[k] is None) == (gb[k] is None), k
        if ga[k] is not None:
            tol = 2e-2 if (k == 'x' or k.endswith('_add_on.weight')) else 1e-5      # bf16 dZ between the two runs' g_pooled
            assert _close(ga[k], gb[k], tol), (k, float((ga[k] - gb[k]).abs().max()), float(gb[k].abs().max()))


def test_orth_prefetch_gives_the_same_loss():
    """the model starts the orth term's weights-only part beside K1 once a loss has used the term; same numbers"""
    from pipnet_b200 import ops
    net, root, args, xs, ys, tr = _problem()
    w = tr._phase_weights(False, 3, 10, args)
    vals = []
    for it in range(3):
        net.zero_grad(set_to_none=True)
        labels = tr.make_labels(net, ys)
        features, pf, pooled, out = net(xs.clone().requires_grad_(True), labels=labels)
        if it > 0:
            assert ops._orth_slot is not None, 'prefetch was not issued'
        res = tr.calculate_loss(3, net, {}, features, pf, pooled, out, ys, net_normalization_multiplier=net._multiplier,
                                pretrain=False, finetune=False, criterion=None, train_iter=None, print=False, EPS=1e-8,
                                root=root, kernel_orth=True, align=False, uni=False, align_pf=True, tanh=True, args=args,
                                device='cuda', labels=labels, **w)
        assert ops._orth_slot is None, 'prefetch branch was not re-joined'
        res[0].backward()
        gw = torch.cat([getattr(net, '_' + n + '_add_on').weight.grad.flatten() for n in net.layout.node_names])
        vals.append((float(res[0]), res.stats[2].clone(), gw.clone()))
    assert net._orth_hint
    for v in vals[1:]:
        assert abs(v[0] - vals[0][0]) <= 1e-6 * max(1.0, abs(vals[0][0])) and torch.equal(v[1], vals[0][1])
        assert _close(v[2], vals[0][2], 1e-4)          # dW: split-K red.add order varies run to run


@pytest.mark.parametrize("extra_consumer", [False, True], ids=["sole-consumer", "extra-consumer"])
def test_prep_tables_hand_off(extra_consumer):
    """the chained loss backward writes K5's scatter table / align coefficients itself; the head backward takes them iff
    the very same g_pooled / g_align reach it -- with a second consumer of `pooled` autograd sums two gradients and the
    head backward must build its own tables from the sum"""
    from pipnet_b200 import ops
    net, root, args, xs, ys, tr = _problem()
    w = tr._phase_weights(False, 3, 10, args)
    grads, taken = [], []
    real_call = ops.call
    for chained in (True, False):
        net.zero_grad(set_to_none=True)
        x = xs.clone().requires_grad_(True)
        labels = tr.make_labels(net, ys)
        features, pf, pooled, out = net(x, labels=labels)
        if not chained:
            out.chained_from = None
        res = tr.calculate_loss(3, net, {}, features, pf, pooled, out, ys, net_normalization_multiplier=net._multiplier,
                                pretrain=False, finetune=False, criterion=None, train_iter=None, print=False, EPS=1e-8,
                                root=root, kernel_orth=True, align=False, uni=False, align_pf=True, tanh=True, args=args,
                                device='cuda', labels=labels, **w)
        loss = res[0] + (0.01 * pooled.flat.square().sum() if extra_consumer else 0.0)
        try:
            ops.call = lambda name, *a: (taken.append((chained, a[15] is None)) if name == 'hcomp_head_bwd_dz' else None,
                                         real_call(name, *a))[1]
            loss.backward()
        finally:
            ops.call = real_call
        torch.cuda.synchronize()
        grads.append((x.grad.float(), torch.cat([getattr(net, '_' + n + '_add_on').weight.grad.flatten()
                                                  for n in net.layout.node_names])))
    assert taken == [(True, not extra_consumer), (False, False)], taken
    assert _close(grads[0][0], grads[1][0], 2e-2) and _close(grads[0][1], grads[1][1], 2e-2)


@pytest.mark.parametrize("inference", [False, True], ids=["train", "inference"])
def test_deferred_riders_match_folded_riders(inference):
    """cub27 at 20 prototypes per node has one rider node (25 nodes = 4 full tiles + 1): finished in the tail of the fused
    kernel behind its grid barrier (default) or by the forward-finish launch (ops.DEFER_RIDERS, the measured-and-parked
    variant) -- same pooled / argmax / logits bit for bit, same align loss"""
    from pipnet_b200 import ops
    net, root, args, xs, ys, tr = _problem(B=7)          # odd pair count, 36 locations: chunks with invalid lanes
    dl = net.device_layout('cuda')
    assert dl.n_spill > 0 and dl.n_wide == 0
    outs = []
    saved = ops.DEFER_RIDERS
    try:
        for defer in (True, False):
            ops.DEFER_RIDERS = defer
            labels = tr.make_labels(net, ys)
            with torch.no_grad():
                _, _, pooled, out = net(xs, inference=inference, labels=labels)
                argmax = net.head(xs, inference=inference, labels=tr.make_labels(net, ys), classify=True)[2]
            torch.cuda.synchronize()
            outs.append((pooled.flat.clone(), out.flat.clone(), pooled.align.clone(), argmax.clone()))
    finally:
        ops.DEFER_RIDERS = saved
    assert int(dl.counter.abs().max()) == 0
    (pa, oa, aa, ga), (pb, ob, ab, gb) = outs
    assert torch.equal(pa, pb) and torch.equal(oa, ob) and torch.equal(ga, gb)
    assert _close(aa, ab, 1e-6)


@pytest.mark.parametrize("case", [("cub27", 768, 26, 4, dict(num_features=20)), ("cub18", 128, 7, 5, dict(num_features=12)),
                                  ("cub27", 128, 6, 4, dict(num_protos_per_child=30, num_features=0)),
                                  ("synth190", 64, 8, 6, dict(num_features=20))],
                         ids=["cub27-real-geometry", "cub18-small", "cub27-wide-node", "cub190-unstored-zero-tiles"])
@pytest.mark.parametrize("phase", [("train", False, False), ("pretrain", True, False)], ids=lambda p: p[0])
def test_block_sparse_backward_matches_dense(case, phase):
    """dX / dW with the unmarked (exactly zero) blocks of dZ skipped against the dense GEMMs on the same step: dX bit for
    bit (same MMA sequence on the nonzero blocks), dW within the split-K atomics' run-to-run variation; and the marks
    must actually be sparse for hierarchical labels"""
    from pipnet_b200 import ops
    tree, C_, H, B, over = case
    _, pretrain, finetune = phase
    net, root, args, xs, ys, tr = _problem(tree=tree, C_=C_, H=H, B=B, **over)
    w = tr._phase_weights(pretrain, 3, 10, args)
    res = []
    saved = ops.SPARSE_BWD
    try:
        for sparse in (True, False):
            ops.SPARSE_BWD = sparse
            net.zero_grad(set_to_none=True)
            x = xs.clone().requires_grad_(True)
            labels = tr.make_labels(net, ys)
            features, pf, pooled, out = net(x, labels=labels)
            slot = pooled.flat._hc_prep
            assert (slot.blocks is not None) == sparse
            r = tr.calculate_loss(3, net, {}, features, pf, pooled, out, ys, net_normalization_multiplier=net._multiplier,
                                  pretrain=pretrain, finetune=finetune, criterion=None, train_iter=None, print=False, EPS=1e-8,
                                  root=root, kernel_orth=True, align=False, uni=False, align_pf=True, tanh=True, args=args,
                                  device='cuda', labels=labels, **w)
            r[0].backward()
            torch.cuda.synchronize()
            gw = torch.cat([getattr(net, '_' + n + '_add_on').weight.grad.flatten() for n in net.layout.node_names])
            frac = float(slot.blocks.buf.float().mean()) if sparse else None
            res.append((x.grad.clone(), gw.clone(), frac))
    finally:
        ops.SPARSE_BWD = saved
    (dx_s, dw_s, frac), (dx_d, dw_d, _) = res
    assert torch.equal(dx_s, dx_d)
    assert _close(dw_s, dw_d, 1e-4)
    assert 0.0 < frac < 1.0, frac          # (coarse 256 x 64 blocks: small trees mark most of them, cub190 ~25 %)


@pytest.mark.parametrize("which", ["dense", "node3", "node8", "last-node"])
def test_block_sparse_backward_with_arbitrary_upstream_gradient(which):
    """an arbitrary gradient on `pooled` -- dense (marks every block), or on ONE node only (its 20 compact dZ columns may
    straddle two 64-column blocks: the softmax couples them, all must be marked) -- sparse path == dense path"""
    from pipnet_b200 import ops
    net, root, args, xs, ys, tr = _problem(C_=128, H=7, B=5)
    L = net.layout
    g = torch.randn(xs.shape[0], L.P, device='cuda')
    if which != "dense":
        ni = {"node3": 3, "node8": 8, "last-node": L.N - 1}[which]
        keep = torch.zeros_like(g)
        keep[:, int(L.proto_off[ni]):int(L.proto_off[ni + 1])] = 1.0
        g = g * keep
    out = []
    saved = ops.SPARSE_BWD
    try:
        for sparse in (True, False):
            ops.SPARSE_BWD = sparse
            net.zero_grad(set_to_none=True)
            x = xs.clone().requires_grad_(True)
            labels = tr.make_labels(net, ys)
            _, _, pooled, _o = net(x, labels=labels)
            ((pooled.flat * g).sum() + (pooled.align.sum() if which == "dense" else 0.0 * pooled.align.sum())).backward()
            torch.cuda.synchronize()
            out.append((x.grad.clone(), torch.cat([getattr(net, '_' + n + '_add_on').weight.grad.flatten()
                                                   for n in net.layout.node_names])))
    finally:
        ops.SPARSE_BWD = saved
    assert torch.equal(out[0][0], out[1][0]) and _close(out[0][1], out[1][1], 1e-4)
