"""The reference's own optimizer builder (`util/args.get_optimizer_nn`, util/args.py:447-567) run UNMODIFIED on
`pipnet_b200.PIPNet`: the drop-in claim of INTEGRATION.md level 1 is that `main_dist.py:334` keeps working, i.e.
every add-on / classifier / presence parameter is discovered through `dir(net.module)` + suffix matching and lands
in the same parameter group with the same learning rate as on the reference's model -- and that re-homing the
parameters into flat buffers afterwards (first forward) does not detach the optimizer from them.

CPU test; needs the reference checkout (build container only) and is skipped on the GPU box."""
import argparse
import types

import pytest
import torch

from oracle import ref_harness as rh
from pipnet_b200.fixtures import build_net, make_args

pytestmark = pytest.mark.skipif(not rh.available(), reason='reference checkout not present')


def _opt_args(**over):
    a = make_args(**over)
    for k, v in dict(seed=1, byol='n', lr=0.05, lr_net=0.0005, lr_block=0.0005, weight_decay=0.0, optimizer='Adam').items():
        setattr(a, k, v)
    return a


def _ids(group):
    return [id(p) for p in group['params']]


@pytest.mark.parametrize('bias', [False, True])
def test_reference_get_optimizer_nn_groups_every_head_parameter(bias):
    rh.load()
    import util.args as ref_args                     # the reference's module, imported in place
    args = _opt_args(num_features=20, bias=bias)
    net, root = build_net('cub27', 64, args, device='cpu')
    wrapped = types.SimpleNamespace(module=net)      # main_dist.py hands DDP(net); the builder only touches .module
    opt_net, opt_cls, to_freeze, to_train, backbone = ref_args.get_optimizer_nn(wrapped, args)

    names = net.layout.node_names
    add_on = [getattr(net, '_' + n + '_add_on').weight for n in names]
    cls_w = [getattr(net, '_' + n + '_classification').weight for n in names]
    cls_b = [getattr(net, '_' + n + '_classification').bias for n in names]
    presence = [getattr(net, '_' + n + '_proto_presence') for n in names]

    # optimizer_net: 3 backbone groups (empty for the identity backbone) + ONE group per node's add-on at 10 x lr_block
    groups = opt_net.param_groups
    assert len(groups) == 3 + len(names)
    seen = set()
    for g in groups[3:]:
        assert g['lr'] == pytest.approx(args.lr_block * 10.0)
        assert len(g['params']) == 1
        seen.add(id(g['params'][0]))
    assert seen == {id(p) for p in add_on}

    # optimizer_classifier: weights | biases (only with --bias) | presence logits
    gw, gb, gp = opt_cls.param_groups
    assert set(_ids(gw)) == {id(p) for p in cls_w} and gw['lr'] == pytest.approx(args.lr)
    assert set(_ids(gp)) == {id(p) for p in presence}
    if bias:
        assert set(_ids(gb)) == {id(p) for p in cls_b if p is not None} and len(gb['params']) == len(names)
    else:
        assert gb['params'] == []
    # the builder freezes every classifier's normalization_multiplier (util/args.py:536-537)
    for n in names:
        assert getattr(net, '_' + n + '_classification').normalization_multiplier.requires_grad is False

    # no parameter is in two groups; nothing of the head is left out
    all_ids = [i for g in opt_net.param_groups + opt_cls.param_groups for i in _ids(g)]
    assert len(all_ids) == len(set(all_ids))
    head = {id(p) for n, p in net.named_parameters()
            if n.endswith(('_add_on.weight', '_classification.weight', '_proto_presence')) or (bias and n.endswith('_classification.bias'))}
    assert head <= set(all_ids)

    # re-homing into the flat buffers keeps the Parameter OBJECTS (the optimizer holds them by identity) ...
    before = {n: id(p) for n, p in net.named_parameters()}
    for grp in (net._w_group, net._wc_group, net._pp_group):
        grp.ensure()
    assert {n: id(p) for n, p in net.named_parameters()} == before
    # ... and an optimizer step through those objects is visible in the flat buffer the kernels read
    p0 = add_on[0]
    p0.grad = torch.ones_like(p0)
    flat_before = net._w_group.flat[: p0.numel()].clone()
    opt_net.step()
    assert not torch.equal(net._w_group.flat[: p0.numel()], flat_before)
    assert p0.data_ptr() == net._w_group.flat.data_ptr()


def test_phase_requires_grad_flips_reach_the_flat_groups():
    """main_dist.py:472-485 / :574-658 toggle requires_grad on the per-node parameters per phase; the flat gather must
    reflect them (frozen parameters get no gradient slot)."""
    args = _opt_args(num_features=20)
    net, root = build_net('cub08', 64, args, device='cpu')
    for attr in dir(net):
        if attr.endswith('_classification'):
            for p in getattr(net, attr).parameters():
                p.requires_grad = False                      # pretraining phase: classifiers frozen
    w = net._wc_group.gather()
    assert not w.requires_grad
    v = net.flat_prototype_kernels()
    assert v.requires_grad
    net._wc_group.invalidate()
    for attr in dir(net):
        if attr.endswith('_classification'):
            getattr(net, attr).weight.requires_grad = True
    assert net._wc_group.gather().requires_grad


def test_suffix_discovery_only_finds_layers_and_parameters():
    """`for attr in dir(net.module): if attr.endswith(<suffix>)` is how main_dist.py (:352-353, :413-414, :474-481, :575-658)
    and util/args.py (:528-556) find the head's layers: no method or helper attribute of the replacement may match."""
    import torch.nn as nn
    args = _opt_args(num_features=20)
    net, _ = build_net('cub08', 64, args, device='cpu')
    n_nodes = len(net.layout.node_names)
    found = {'_add_on': 0, '_classification': 0, '_proto_presence': 0}
    for attr in dir(net):
        obj = getattr(net, attr)
        if attr.endswith('_add_on'):
            assert type(obj) is nn.Conv2d, attr            # util/func.py:9 tests `type(m) == nn.Conv2d` for the xavier init
            found['_add_on'] += 1
        elif attr.endswith('_classification'):
            assert isinstance(obj, nn.Module) and hasattr(obj, 'weight') and hasattr(obj, 'normalization_multiplier'), attr
            found['_classification'] += 1
        elif attr.endswith('_proto_presence'):
            assert isinstance(obj, nn.Parameter), attr
            found['_proto_presence'] += 1
    assert found == {k: n_nodes for k in found}
