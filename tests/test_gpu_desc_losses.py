"""GPU parity of the descendant-structured loss terms the reference's shipped scripts switch on
(--tanh_desc, --minimize_contrasting_set, --mask_prune_overspecific; pipnet/train.py:946-1060, 1089-1133):
full model step through `PIPNet.forward` + `calculate_loss` + backward vs the CPU oracle (itself pinned to the
reference in tests/test_oracle_vs_reference.py, Gumbel draws included).  The Gumbel noise is injected on both sides."""
import pytest
import torch

from oracle import head_oracle as ho
from oracle.problems import bf16_round, build_net, desc_loss_kwargs, flat_gumbel, make_args, rel_err

pytestmark = pytest.mark.gpu

CASES = [
    # name, tree, C, H, B, args overrides, (pretrain, finetune), epoch
    ("shipped-B", "cub27", 64, 6, 8, dict(num_protos_per_child=4, num_features=0, tanh_desc='y|0.05',
                                          minimize_contrasting_set='y', mask_prune_overspecific='y|0|1.1'), (False, False), 3),
    ("shipped-A", "cub18", 64, 6, 6, dict(num_features=12, tanh_desc='y|0.05', minimize_contrasting_set='y|1|0.2',
                                          mask_prune_overspecific='y|0|1.1'), (False, False), 3),
    ("plain-score", "cub18", 64, 6, 6, dict(num_features=12, mask_prune_overspecific='y|0'), (False, False), 3),
    ("geometric", "cub08", 64, 6, 5, dict(num_features=20, mask_prune_overspecific='y|0',
                                          geometric_mean_overspecificity_score='y'), (False, False), 3),
    ("sg-score", "cub08", 64, 6, 5, dict(num_features=20, mask_prune_overspecific='y|0|1.1', sg_before_masking='y'),
     (False, False), 3),
    ("finetune", "cub27", 64, 6, 8, dict(num_features=20, tanh_desc='y|0.05', minimize_contrasting_set='y',
                                         mask_prune_overspecific='y|0|1.1'), (False, True), 3),
    ("not-yet", "cub08", 64, 6, 4, dict(num_features=20, mask_prune_overspecific='y|5|1.1'), (False, False), 3),
    ("wide-64", "cub27", 64, 6, 6, dict(num_protos_per_child=20, num_features=0, tanh_desc='y|0.05',
                                        minimize_contrasting_set='y', mask_prune_overspecific='y|0|1.1'), (False, False), 3),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_shipped_recipe_terms_match_oracle(case):
    from pipnet_b200 import train as tr
    _, tree, C, H, B, over, (pretrain, finetune), epoch = case
    args = make_args(**over)
    net, root = build_net(tree, C, args)
    L = net.layout
    names = L.node_names
    nodes = root.nodes_with_children()
    g = torch.Generator().manual_seed(23)
    with torch.no_grad():
        for n in names:
            pp = getattr(net, '_' + n + '_proto_presence')
            pp.copy_(torch.randn(pp.shape, generator=g).to(pp.device))
            wc = getattr(net, '_' + n + '_classification').weight
            # push some classifier weights under the two relevance thresholds (1e-3 and 1e-5) without emptying a row
            noise = torch.rand(wc.shape, generator=g)
            small = torch.where(noise < 0.15, torch.full_like(noise, 5e-4), torch.where(noise < 0.25, torch.full_like(noise, 1e-6), torch.ones_like(noise)))
            small[:, 0] = 1.0
            wc.mul_(small.to(wc.device))
    x = bf16_round(torch.randn(2 * B, C, H, H, generator=g))
    ys = torch.randint(0, L.L, (B,), generator=g)
    ys = torch.cat([ys, ys])
    label2name = {i: n for i, n in enumerate(L.leaf_names)}

    aw = {n: getattr(net, '_' + n + '_add_on').weight.detach().flatten(1).double().cpu() for n in names}
    cw = {n: getattr(net, '_' + n + '_classification').weight.detach().double().cpu() for n in names}
    pres = {n: getattr(net, '_' + n + '_proto_presence').detach().double().cpu() for n in names}
    used = {}
    torch.manual_seed(5)
    ref = ho.full_step(x.double(), aw, cw, root, ys, label2name, pretrain=pretrain, finetune=finetune, softmax_tau=1.0,
                       epoch=epoch, nr_epochs=10, cl_weight=args.cl_weight, presence=pres, gumbel_out=used,
                       **desc_loss_kwargs(args))
    # the noise the oracle drew, re-indexed like the classifier weights (node, child, prototype)
    gum = flat_gumbel(used, nodes).float()
    assert gum.shape[0] == L.n_welems

    net = net.cuda()
    xs = x.cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    labels = tr.make_labels(net, ys.cuda())
    features, proto_features, pooled, out = net(xs, labels=labels)
    w = tr._phase_weights(pretrain, epoch, 10, args)
    res = tr.calculate_loss(epoch, net, {}, features, proto_features, pooled, out, ys.cuda(),
                            net_normalization_multiplier=net._multiplier, pretrain=pretrain, finetune=finetune,
                            criterion=None, train_iter=None, print=False, EPS=1e-8, root=root, kernel_orth=True,
                            tanh_desc='y' in args.tanh_desc, align=False, uni=False, align_pf=True, tanh=True, args=args,
                            device='cuda', labels=labels, gumbel_noise=gum.cuda(), **w)
    loss = res[0]
    loss.backward()
    torch.cuda.synchronize()

    assert abs(float(loss.detach()) - float(ref['loss'])) <= 1e-5 * max(1.0, abs(float(ref['loss']))), (float(loss.detach()), float(ref['loss']))
    ds = res.desc_stats
    any_term = bool(ref['tanh_desc'] or ref['contrast'] or ref['ovsp'])
    assert (ds is not None) == any_term or (ds is not None and not any_term and float(ds.abs().sum()) == 0.0)
    if ds is not None:
        ds = ds.cpu()
        for row, key in enumerate(('tanh_desc', 'contrast', 'ovsp', 'mask_l1')):
            for i, n in enumerate(names):
                want = float(ref[key][n]) if n in ref[key] else 0.0
                assert abs(float(ds[row, i]) - want) <= 2e-5 * max(1.0, abs(want)), (key, n, float(ds[row, i]), want)
        if ref['tanh_desc']:
            mean_td = sum(float(v) for v in ref['tanh_desc'].values()) / len(ref['tanh_desc'])
            assert abs(float(res[17]) - mean_td) <= 2e-5 * max(1.0, abs(mean_td))
    # presence logits: fp32 arithmetic on [P,2]
    for n in names:
        got = getattr(net, '_' + n + '_proto_presence').grad
        want = ref['grad_presence'][n]
        if want is None:
            assert got is None or float(got.abs().max()) == 0.0
        else:
            assert got is not None
            assert (got.double().cpu() - want).abs().max() <= 2e-4 * max(1e-6, float(want.abs().max())) + 1e-9, n
    # the terms reach the prototype kernels / features through pooled: bf16 tolerance (dZ is stored in bf16)
    gw = torch.cat([getattr(net, '_' + n + '_add_on').weight.grad.flatten(1) for n in names])
    gw_ref = torch.cat([ref['grad_w'][n] if ref['grad_w'][n] is not None else torch.zeros_like(aw[n]) for n in names])
    assert rel_err(gw, gw_ref) <= 2e-2, f"dW {rel_err(gw, gw_ref)}"
    if ref['grad_x'] is not None and not finetune:
        assert rel_err(xs.grad, ref['grad_x']) <= 2e-2, f"dX {rel_err(xs.grad, ref['grad_x'])}"
    for n in names:
        gc = getattr(net, '_' + n + '_classification').weight.grad
        rc = ref['grad_cls'][n] if ref['grad_cls'][n] is not None else torch.zeros_like(cw[n])
        assert (gc.double().cpu() - rc).abs().max() <= 1e-4 * max(1e-3, float(rc.abs().max())) + 1e-7, n


def test_default_noise_is_drawn_on_device():
    """without injected noise the term still runs (fresh Gumbel draws per step) and stays finite"""
    from pipnet_b200 import train as tr
    args = make_args(num_features=20, mask_prune_overspecific='y|0|1.1', tanh_desc='y|0.05', minimize_contrasting_set='y')
    net, root = build_net("cub08", 64, args)
    net = net.cuda()
    g = torch.Generator().manual_seed(1)
    x = bf16_round(torch.randn(8, 64, 6, 6, generator=g))
    ys = torch.randint(0, net.layout.L, (4,), generator=g)
    ys = torch.cat([ys, ys]).cuda()
    xs = x.cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    labels = tr.make_labels(net, ys)
    vals = []
    for _ in range(2):
        features, pf, pooled, out = net(xs, labels=labels)
        res = tr.calculate_loss(1, net, {}, features, pf, pooled, out, ys, net_normalization_multiplier=net._multiplier,
                                pretrain=False, finetune=False, criterion=None, train_iter=None, print=False, EPS=1e-8,
                                root=root, kernel_orth=True, tanh_desc=True, align=False, uni=False, align_pf=True, tanh=True,
                                args=args, device='cuda', labels=labels, **tr._phase_weights(False, 1, 10, args))
        res[0].backward()
        vals.append(float(res[0].detach()))
    assert all(v == v and abs(v) < 1e6 for v in vals)
    assert vals[0] != vals[1]          # different Gumbel draws
