"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference (harishB97/PIPNet, mounted
read-only at /root/reference) so that the CPU oracle (`oracle/head_oracle.py`) and the golden
fixtures (`tests/golden/*.npz`) can be pinned to what the reference really computes.

Nothing in the product (`pipnet_b200/`) may import this module.  It only works in the build
container: /root/reference does not exist on the GPU box, which is why the vectors it
produces are committed under tests/golden/ (see `oracle/make_golden.py`).

Recipe: three third-party modules that the hot path never calls are
stubbed (graphviz `util/node.py:3`, torchmetrics.functional `pipnet/train.py:12`, kornia.losses
`util/custom_losses.py:3`), and the backbone factory is replaced by an identity module so that
`get_network` (`pipnet/pipnet.py:1134`) can infer C_in without downloading weights.
"""
from __future__ import annotations

import argparse
import os
import sys
import types
from collections import defaultdict

import torch
import torch.nn as nn

REF_ROOT = os.environ.get('HCOMP_REFERENCE_ROOT', '/root/reference')


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, 'pipnet', 'pipnet.py'))


_mods = None


def load():
    """Import reference modules once; returns (pipnet.pipnet, pipnet.train, util.node, util.custom_losses)."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError(f'reference not found under {REF_ROOT}')

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    if 'graphviz' not in sys.modules:
        stub('graphviz', Digraph=object)
    if 'torchmetrics' not in sys.modules:
        tm = stub('torchmetrics')
        tm.functional = stub('torchmetrics.functional', f1_score=None, recall=None, precision=None)
    if 'kornia' not in sys.modules:
        k = stub('kornia')
        k.losses = stub('kornia.losses', FocalLoss=object)
    if 'wandb' not in sys.modules:
        try:
            import wandb  # noqa: F401
        except Exception:
            stub('wandb')
    # The reference is a flat script directory whose package names (`pipnet`, `util`,
    # `features`) must win over anything else on sys.path.
    sys.path.insert(0, REF_ROOT)
    import pipnet.pipnet as ref_pipnet
    import pipnet.train as ref_train
    import util.node as ref_node
    import util.custom_losses as ref_losses
    sys.path.remove(REF_ROOT)
    _mods = (ref_pipnet, ref_train, ref_node, ref_losses)
    return _mods


class _IdentityBackbone(nn.Module):
    """Stands in for the ConvNeXt/ResNet feature net: forward is identity, and it owns one
    Conv2d so `get_network` reads `out_channels` as C_in (`pipnet/pipnet.py:1154-1155`)."""

    def __init__(self, channels):
        super().__init__()
        self.c = nn.Conv2d(3, channels, 1)

    def forward(self, x):
        return x


def make_args(**over):
    a = dict(net='convnext_tiny_26', disable_pretrained=True, basic_cnext_gaussian_multiplier='',
             stage4_reducer_net='', num_features=20, num_protos_per_descendant=0, num_protos_per_child=0,
             unitconv2d='n', projectconv2d='n', l2conv2d='n', add_on_bias=False, bias=False,
             classifier='NonNegative', protopool='n', softmax='y|1', gumbel_softmax='n', gs_tau=1.0,
             multiply_cs_softmax='n', conc_log_ip='n', sg_before_protos='n', softmax_over_channel='n',
             focal='n', mask_prune_overspecific='n', minimize_contrasting_set='n',
             tanh_during_second_phase='y', tanh_desc='n', pipnet_sparsity='y', cl_weight=2.0,
             leave_out_classes='', OOD_ent='n', image_size=224)
    a.update(over)
    return argparse.Namespace(**a)


class _Wrap:
    """`net.module` indirection the reference code expects (DDP-style)."""

    def __init__(self, m):
        self.module = m


class _Iter:
    def set_postfix_str(self, *a, **k):
        pass


def build_reference_net(tree_edges, channels, args, seed=1):
    """Reference `PIPNet` on the identity backbone, initialised like `main_dist.py:413-427`
    (xavier-uniform add-on weights via `util/func.py:8-10`, `_multiplier = 2`, frozen)."""
    ref_pipnet, _, ref_node, _ = load()
    from pipnet_b200.trees import build_tree
    root = build_tree(tree_edges, ref_node.Node)
    for node in root.nodes_with_children():
        node.set_num_protos(num_protos_per_descendant=args.num_protos_per_descendant,
                            num_protos_per_child=args.num_protos_per_child,
                            min_protos=args.num_features, split_protos=True)      # main_dist.py:188-192
        node.set_loss_weightage_using_descendants_count()                        # main_dist.py:294-297
    ref_pipnet.base_architecture_to_features['convnext_tiny_26'] = lambda pretrained=False: _IdentityBackbone(channels)
    torch.manual_seed(seed)
    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        feats, add_on, pool, cls_layers, num_protos = ref_pipnet.get_network(len(root.leaf_descendents), args, root=root)
        net = ref_pipnet.PIPNet(num_classes=len(root.leaf_descendents), num_prototypes=num_protos, feature_net=feats,
                                args=args, add_on_layers=add_on, pool_layer=pool, classification_layers=cls_layers,
                                num_parent_nodes=len(root.nodes_with_children()), root=root)
    with torch.no_grad():
        for name in add_on:
            nn.init.xavier_uniform_(getattr(net, '_' + name + '_add_on').weight, gain=1.0)
        net._multiplier.fill_(2.0)
        net._multiplier.requires_grad = False
    return net, root


def run_reference(net, root, x, ys, args, *, pretrain, finetune, epoch=1, nr_epochs=10, dtype=torch.float32,
                  kernel_orth=True, inference=False, rng_seed=None):
    """One reference step: `PIPNet.forward` (`pipnet/pipnet.py:111-171`) + `calculate_loss`
    (`pipnet/train.py:852`) with the loss weights of `train_pipnet` (`pipnet/train.py:148-177`)
    + backward.  Returns plain tensors keyed by node name."""
    _, ref_train, _, ref_losses = load()
    net = net.to(dtype)
    x = x.to(dtype).clone().requires_grad_(True)
    names = sorted(root.leaf_descendents)
    label2name = {i: n for i, n in enumerate(names)}
    node_accuracy = {}
    for node in root.nodes_with_children():
        node_accuracy[node.name] = {'n_examples': 0, 'n_correct': 0, 'accuracy': None, 'f1': None,
                                    'preds': torch.empty(0, node.num_children()), 'gts': torch.empty(0),
                                    'children': defaultdict(lambda: {'n_examples': 0, 'n_correct': 0})}
    for p in net.parameters():
        p.grad = None
    features, proto_features, pooled, out = net(x, inference=inference)
    if pretrain:
        w = dict(align_pf_weight=(epoch / nr_epochs) * 1., t_weight=5., cl_weight=0.)
    else:
        w = dict(align_pf_weight=5., t_weight=2., cl_weight=args.cl_weight)
    criterion = ref_losses.WeightedNLLLoss(device='cpu')
    if rng_seed is not None:          # the only RNG consumer on this path is the Gumbel softmax of the mask-prune term
        torch.manual_seed(rng_seed)
    res = ref_train.calculate_loss(
        epoch, _Wrap(net), {}, features, proto_features, pooled, out, ys,
        align_weight=0.5, align_pf_weight=w['align_pf_weight'], t_weight=w['t_weight'], mm_weight=0., unif_weight=3.,
        cl_weight=w['cl_weight'], OOD_loss_weight=0., orth_weight=0.5, cluster_desc_weight=0.8, sep_desc_weight=0.08,
        subspace_sep_weight=1e-2, byol_weight=0.5, net_normalization_multiplier=net._multiplier,
        pretrain=pretrain, finetune=finetune, criterion=criterion, train_iter=_Iter(), print=True, EPS=1e-8,
        root=root, label2name=label2name, node_accuracy=node_accuracy, OOD_loss_required=False,
        kernel_orth=kernel_orth, tanh_desc=('y' in args.tanh_desc), align=False, uni=False, align_pf=True, tanh=True,
        minmaximize=False, cluster_desc=False, sep_desc=False, subspace_sep=False, byol=False, args=args, device='cpu')
    loss, class_loss, _a, tanh_loss, _mm, _ood, orth_loss = res[:7]
    loss.backward()
    nodes = root.nodes_with_children()
    g = {}
    for node in nodes:
        conv = getattr(net, '_' + node.name + '_add_on')
        cls = getattr(net, '_' + node.name + '_classification')
        g[node.name] = (None if conv.weight.grad is None else conv.weight.grad.detach().flatten(1).clone(),
                        None if cls.weight.grad is None else cls.weight.grad.detach().clone())
    g_presence = {}
    for node in nodes:
        pp = getattr(net, '_' + node.name + '_proto_presence')
        g_presence[node.name] = None if pp.grad is None else pp.grad.detach().clone()
    H, W = x.shape[-2:]
    argmax = {}
    for node in nodes:
        pf = proto_features[node.name].detach()
        argmax[node.name] = pf.flatten(2).argmax(dim=2)          # first max in flat h*W+w order on CPU
    return dict(loss=loss.detach(), class_loss={k: torch.as_tensor(v).detach() for k, v in class_loss.items()},
                tanh_loss={k: torch.as_tensor(v).detach() for k, v in tanh_loss.items()},
                orth_loss={k: torch.as_tensor(v).detach() for k, v in orth_loss.items()},
                pooled={k: v.detach() for k, v in pooled.items()}, out={k: v.detach() for k, v in out.items()},
                proto_features={k: v.detach() for k, v in proto_features.items()}, argmax=argmax,
                grad_x=None if x.grad is None else x.grad.detach().clone(), grads=g, grad_presence=g_presence,
                avg_tanh_desc=float(res[17]),
                node_accuracy={k: (v['n_examples'], v['n_correct']) for k, v in node_accuracy.items()})
