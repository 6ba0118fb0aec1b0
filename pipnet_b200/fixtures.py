"""Synthetic-model builders used by bench.py, the smoke test and the parity tests: trees with prototype counts
assigned the way the reference's driver does it, the argparse namespace of the shipped recipe, and a `PIPNet` on an
identity backbone (the head's input IS the feature map) with seeded, bf16-representable prototype kernels.
Product-side helpers: nothing here touches oracle/."""
from __future__ import annotations

import argparse

import torch
import torch.nn as nn

from .trees import get_tree, build_tree


def make_tree(tree, *, num_features=0, per_child=0, per_desc=0):
    """named tree / edge list -> Node tree with per-node prototype counts (util/node.py:45-71) and class weights"""
    root = get_tree(tree) if isinstance(tree, str) else build_tree(tree)
    for node in root.nodes_with_children():
        node.set_num_protos(num_protos_per_descendant=per_desc, num_protos_per_child=per_child,
                            min_protos=num_features, split_protos=True)
        node.set_loss_weightage_using_descendants_count()
    return root


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(t.dtype)


class IdentityBackbone(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.c = nn.Conv2d(3, channels, 1)       # get_network reads out_channels of the last conv

    def forward(self, x):
        return x


def make_args(**over):
    """argparse namespace with the flags of the reference's shipped recipe (run_pipnet_20protos_multi_runs_seed42.sh:69-94)"""
    a = dict(net='identity', disable_pretrained=True, basic_cnext_gaussian_multiplier='', stage4_reducer_net='',
             num_features=20, num_protos_per_descendant=0, num_protos_per_child=0, unitconv2d='n', projectconv2d='n',
             l2conv2d='n', add_on_bias=False, bias=False, classifier='NonNegative', protopool='n', softmax='y|1',
             gumbel_softmax='n', gs_tau=1.0, multiply_cs_softmax='n', conc_log_ip='n', sg_before_protos='n',
             softmax_over_channel='n', focal='n', mask_prune_overspecific='n', minimize_contrasting_set='n',
             tanh_during_second_phase='y', tanh_desc='n', pipnet_sparsity='y', cl_weight=2.0, leave_out_classes='',
             OOD_ent='n')
    a.update(over)
    return argparse.Namespace(**a)


def build_net(tree, C, args, seed=3, device='cuda'):
    """PIPNet (this package's) on an identity backbone of `C` channels; kernels xavier-uniform rounded to bf16
    (util/func.py:8-10), `_multiplier` = 2 and frozen as main_dist.py:426-427 does."""
    from . import pipnet as pp
    root = make_tree(tree, num_features=args.num_features, per_child=args.num_protos_per_child,
                     per_desc=getattr(args, 'num_protos_per_descendant', 0))
    pp.base_architecture_to_features['identity'] = lambda pretrained=False: IdentityBackbone(C)
    torch.manual_seed(seed)
    feats, add_on, pool, cls_layers, nproto = pp.get_network(len(root.leaf_descendents), args, root=root)
    net = pp.PIPNet(len(root.leaf_descendents), nproto, feats, args, add_on, pool, cls_layers, len(root.nodes_with_children()), root)
    with torch.no_grad():
        for name in add_on:
            w = getattr(net, '_' + name + '_add_on').weight
            nn.init.xavier_uniform_(w, gain=1.0)
            w.copy_(bf16_round(w))
        net._multiplier.fill_(2.0)
        net._multiplier.requires_grad = False
    return (net.cuda() if device == 'cuda' else net), root
