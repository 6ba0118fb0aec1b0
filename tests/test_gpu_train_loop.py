"""`train_pipnet` / `test_pipnet` end to end on a synthetic two-view loader with the reference's
optimizer layout (one AdamW group per node, `util/args.py:528-556`)."""
import pytest
import torch

from oracle.problems import build_net, make_args

pytestmark = pytest.mark.gpu


class _DS(torch.utils.data.Dataset):
    def __init__(self, n, C, H, L, two_views, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.x1 = torch.randn(n, C, H, H, generator=g).to(torch.bfloat16)
        self.x2 = (self.x1.float() + 0.1 * torch.randn(n, C, H, H, generator=g)).to(torch.bfloat16)
        self.y = torch.randint(0, L, (n,), generator=g)
        self.two = two_views

    def __len__(self): return self.y.numel()

    def __getitem__(self, i):
        return (self.x1[i], self.x2[i], self.y[i]) if self.two else (self.x1[i], self.y[i])


def _loader(net, two_views):
    ds = _DS(32, 64, 6, net.layout.L, two_views)
    ds.class_to_idx = {n: i for i, n in enumerate(net.layout.leaf_names)}
    return torch.utils.data.DataLoader(ds, batch_size=8, shuffle=False)


def test_train_and_test_epochs_run_and_learn():
    from pipnet_b200 import train as tr
    args = make_args()
    net, root = build_net("cub08", 64, args)
    names = net.layout.node_names
    opt_net = torch.optim.AdamW([{'params': [getattr(net, '_' + n + '_add_on').weight], 'lr': 5e-3} for n in names])
    opt_cls = torch.optim.AdamW([getattr(net, '_' + n + '_classification').weight for n in names], lr=5e-2)
    sch_net = torch.optim.lr_scheduler.CosineAnnealingLR(opt_net, T_max=40)
    sch_cls = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt_cls, T_0=5)
    losses = []
    for epoch in range(1, 4):
        info, log = tr.train_pipnet(net, _loader(net, True), opt_net, opt_cls, sch_net, sch_cls, None, epoch, 3, 'cuda',
                                    pretrain=False, finetune=False, kernel_orth=True, align=False, uni=False, align_pf=True,
                                    tanh=True, wandb_logging=True, args=args)
        losses.append(info['loss'])
        assert 0.0 <= info['fine_accuracy'] <= 1.0 and 'train/epoch loss' in log
        assert set(info['node_accuracy']) == set(names)
    assert losses[-1] < losses[0], losses
    info, _ = tr.test_pipnet(net, _loader(net, False), opt_net, opt_cls, sch_net, sch_cls, None, 1, 3, 'cuda',
                             kernel_orth=True, align=False, uni=False, align_pf=True, tanh=True, args=args)
    assert info['loss'] > 0
    # pretrain phase: classifier untouched
    before = torch.cat([getattr(net, '_' + n + '_classification').weight.detach().flatten() for n in names]).clone()
    tr.train_pipnet(net, _loader(net, True), opt_net, opt_cls, sch_net, None, None, 1, 3, 'cuda', pretrain=True,
                    align=False, uni=False, align_pf=True, tanh=True, args=args)
    after = torch.cat([getattr(net, '_' + n + '_classification').weight.detach().flatten() for n in names])
    assert torch.equal(before, after)


def test_convnext26_backbone_handoff_step():
    """The real hand-off: torchvision ConvNeXt-tiny with the reference's stride relaxation
    (`features/convnext_features.py:7-25`, random init -- no weights on the box) feeds 768x26x26 maps into the fused head;
    one full training-phase step with the shipped scripts' loss set runs, is finite, reaches the backbone's parameters,
    and the head's pooled scores / argmax agree with the plain map kernel on the backbone's actual output."""
    from oracle.problems import make_tree
    from pipnet_b200 import ops, pipnet as pp, train as tr
    torch.manual_seed(0)
    root = make_tree("cub08", num_features=20)
    args = make_args(net='convnext_tiny_26', num_features=20, tanh_desc='y|0.05', minimize_contrasting_set='y',
                     mask_prune_overspecific='y|0|1.1')
    feats, add_on, pool, cls_layers, k = pp.get_network(8, args, root=root)
    net = pp.PIPNet(8, k, feats, args, add_on, pool, cls_layers, len(root.nodes_with_children()), root).cuda()
    net.train()
    B = 2
    x = torch.randn(2 * B, 3, 224, 224, device='cuda')
    ys = torch.tensor([1, 5, 1, 5], device='cuda')
    labels = tr.make_labels(net, ys)
    features, pf, pooled, out = net(x, labels=labels)
    assert tuple(features.shape) == (2 * B, 768, 26, 26)
    w = tr._phase_weights(False, 1, 10, args)
    res = tr.calculate_loss(1, net, {}, features, pf, pooled, out, ys, net_normalization_multiplier=net._multiplier,
                            pretrain=False, finetune=False, criterion=None, train_iter=None, print=False, EPS=1e-8, root=root,
                            kernel_orth=True, tanh_desc=True, align=False, uni=False, align_pf=True, tanh=True, args=args,
                            device='cuda', labels=labels, **w)
    res[0].backward()
    torch.cuda.synchronize()
    assert torch.isfinite(res[0]).item()
    stem = next(net._net.parameters())
    assert stem.grad is not None and torch.isfinite(stem.grad).all() and float(stem.grad.abs().max()) > 0
    for n in net.layout.node_names:
        for suffix in ('_add_on', '_classification'):
            g = getattr(net, '_' + n + suffix).weight.grad
            assert g is not None and torch.isfinite(g).all()
    # head output vs the SIMT map kernel on the same features
    L = net.layout
    wflat = net.flat_prototype_kernels().detach()
    p0, p1 = int(L.proto_off[1]), int(L.proto_off[2])
    m = ops.materialize_map(features.detach(), wflat[p0:p1], 1.0).flatten(2)
    mv, _ = m.max(dim=2)
    torch.testing.assert_close(pooled.flat[:, p0:p1].detach(), mv, rtol=2e-2, atol=1e-4)     # fp32 features -> bf16 operands
    am = pf.argmax.flat[:, p0:p1].long()
    torch.testing.assert_close(m.gather(2, am.unsqueeze(-1)).squeeze(-1), mv, rtol=2e-2, atol=1e-4)


def test_shipped_script_flags_through_epoch_drivers():
    """`train_pipnet` / `test_pipnet` with the flag set of run_pipnet_20protos_multi_runs_seed42.sh:72-94 (tanh_desc,
    minimize_contrasting_set, mask pruning) in the training and fine-tuning phases, then the leave-out evaluation
    (`main_dist.py:727-735`) with the test-time overspecificity mask."""
    from pipnet_b200 import train as tr
    args = make_args(num_features=20, tanh_desc='y|0.05', minimize_contrasting_set='y', mask_prune_overspecific='y|0|1.1')
    net, root = build_net("cub08", 64, args)
    names = net.layout.node_names
    params = [getattr(net, '_' + n + '_add_on').weight for n in names] + [getattr(net, '_' + n + '_proto_presence') for n in names]
    opt_net = torch.optim.AdamW(params, lr=5e-3)
    opt_cls = torch.optim.AdamW([getattr(net, '_' + n + '_classification').weight for n in names], lr=5e-2)
    sch_net = torch.optim.lr_scheduler.CosineAnnealingLR(opt_net, T_max=40)
    sch_cls = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt_cls, T_0=5)
    pres_before = torch.cat([getattr(net, '_' + n + '_proto_presence').detach().flatten() for n in names]).clone()
    common = dict(kernel_orth=True, tanh_desc=True, align=False, uni=False, align_pf=True, tanh=True, wandb_logging=True, args=args)
    info, log = tr.train_pipnet(net, _loader(net, True), opt_net, opt_cls, sch_net, sch_cls, None, 1, 3, 'cuda', pretrain=False,
                                finetune=False, **common)
    assert info['loss'] == info['loss'] and 'train/epoch loss' in log
    pres_after = torch.cat([getattr(net, '_' + n + '_proto_presence').detach().flatten() for n in names])
    assert not torch.equal(pres_before, pres_after)              # the mask-pruning term trains the presence logits
    info, _ = tr.train_pipnet(net, _loader(net, True), opt_net, opt_cls, sch_net, sch_cls, None, 2, 3, 'cuda', pretrain=False,
                              finetune=True, **common)
    assert info['loss'] == info['loss']
    leaf_children = [c.name for n in root.nodes_with_children() for c in n.children if c.is_leaf()]
    info, _ = tr.test_pipnet(net, _loader(net, False), opt_net, opt_cls, sch_net, sch_cls, None, 1, 3, 'cuda',
                             kernel_orth=True, tanh_desc=True, align=False, uni=False, align_pf=True, tanh=True, args=args,
                             leave_out_classes=[leaf_children[0]], apply_overspecificity_mask=True)
    assert 0.0 <= info['fine_accuracy'] <= 1.0


@pytest.mark.parametrize("phase", ["pretrain", "train", "finetune"])
def test_graphed_head_step_matches_eager_epoch(phase, monkeypatch):
    """`train_pipnet` replays the head step (forward, losses, head backward) as ONE CUDA graph between the backbone's
    forward and backward when shapes are static (pipnet_b200.train.GraphedHeadTrainStep).  Two epochs through the graph
    path must leave the same parameters, losses and accuracies as the eager path (HC_HEAD_GRAPH=0)."""
    from pipnet_b200 import train as tr
    pretrain, finetune = phase == "pretrain", phase == "finetune"
    results = []
    for graph in ("1", "0"):
        monkeypatch.setenv("HC_HEAD_GRAPH", graph)
        tr._GRAPH_CACHE.clear()
        args = make_args()
        net, root = build_net("cub18", 64, args, seed=5)
        names = net.layout.node_names
        for n in names:
            getattr(net, '_' + n + '_add_on').weight.requires_grad = not finetune
            getattr(net, '_' + n + '_classification').weight.requires_grad = not pretrain
        opt_net = torch.optim.SGD([{'params': [getattr(net, '_' + n + '_add_on').weight], 'lr': 1e-2} for n in names])
        opt_cls = torch.optim.SGD([getattr(net, '_' + n + '_classification').weight for n in names], lr=1e-2)
        sch_net = torch.optim.lr_scheduler.CosineAnnealingLR(opt_net, T_max=40)
        sch_cls = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt_cls, T_0=5)
        infos = []
        for epoch in (1, 2):
            info, _ = tr.train_pipnet(net, _loader(net, True), opt_net, opt_cls, sch_net, sch_cls, None, epoch, 3, 'cuda',
                                      pretrain=pretrain, finetune=finetune, kernel_orth=True, align=False, uni=False,
                                      align_pf=True, tanh=True, args=args)
            infos.append(info)
        used_graph = any(v for v in tr._GRAPH_CACHE.values())
        assert used_graph == (graph == "1")
        params = torch.cat([p.detach().flatten().float() for p in net.parameters()])
        results.append((infos, params))
    (ig, pg), (ie, pe) = results
    # dW is a split-K fp32 atomic accumulation: identical up to summation order
    assert float((pg - pe).abs().max()) <= 1e-5 * float(pe.abs().max())
    for a, b in zip(ig, ie):
        assert abs(a['loss'] - b['loss']) <= 1e-5 * max(1.0, abs(b['loss']))
        assert a['fine_accuracy'] == b['fine_accuracy']
        assert a['node_accuracy'] == b['node_accuracy']


@pytest.mark.parametrize("autocast", [False, True], ids=["fp32", "bf16-autocast"])
def test_fused_convnext_tail_matches_stock_block(autocast):
    """SURVEY 8f-4, second half: the last ConvNeXt block (`features.7.2`) writes the head's bf16 channels-last feature
    matrix itself (`layer_scale * block(x)` + stochastic depth + residual in one kernel, ops.ScaleResidualRows).  Output and
    gradients must match the stock torchvision block, and the head must consume the result without any cast / layout call."""
    from torchvision.models.convnext import CNBlock
    from pipnet_b200 import ops, pipnet as pp, _cabi
    torch.manual_seed(0)
    stock = CNBlock(768, 1e-6, 0.0).cuda()
    with torch.no_grad():
        stock.layer_scale.copy_(torch.rand_like(stock.layer_scale) + 0.5)
    import copy
    fused = copy.deepcopy(stock)
    holder = torch.nn.Sequential(fused)
    assert pp.fuse_convnext_tail(holder)
    x = torch.randn(4, 768, 13, 13, device='cuda').contiguous(memory_format=torch.channels_last)
    g = torch.randn(4, 768, 13, 13, device='cuda')
    outs = []
    for blk in (stock, fused):
        xi = x.clone().requires_grad_(True)
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=autocast):
            y = blk(xi)
        (y.float() * g).sum().backward()
        outs.append((y.detach().float(), xi.grad.float(), blk.layer_scale.grad.float(),
                     blk.block[0].weight.grad.float()))
    (y0, gx0, gg0, gw0), (y1, gx1, gg1, gw1) = outs
    assert fused(x).dtype == torch.bfloat16 and fused(x).is_contiguous(memory_format=torch.channels_last)
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    assert rel(y1, y0) <= 2e-2 and rel(gx1, gx0) <= 2e-2 and rel(gg1, gg0) <= 2e-2 and rel(gw1, gw0) <= 2e-2
    # stochastic depth in training mode: dropped images are the plain residual
    drop = copy.deepcopy(stock)
    drop.stochastic_depth.p = 1.0
    hd = torch.nn.Sequential(drop)
    pp.fuse_convnext_tail(hd)
    drop.train()
    assert rel(drop(x).float(), x) <= 1e-2
    # the head reads the fused output in place: no cast / transpose call between backbone and projection kernel
    names = []
    real_call = ops.call
    try:
        ops.call = lambda name, *a: (names.append(name), real_call(name, *a))[1]
        rows = ops.feature_rows(fused(x))
    finally:
        ops.call = real_call
    assert rows.dtype == torch.bfloat16 and rows.shape == (4 * 169, 768)
    assert 'hcomp_cast_f32_to_bf16' not in names and 'hcomp_nchw_to_rows_bf16' not in names
