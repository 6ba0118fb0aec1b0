"""Shared builders for the parity tests: one seeded problem instance expressed both ways --
per-node dicts for the CPU oracle and flat tensors for the CUDA path."""
from __future__ import annotations

import numpy as np
import torch

from pipnet_b200.layout import build_layout
from pipnet_b200.trees import get_tree, NAMED, build_tree


from pipnet_b200.fixtures import make_tree, bf16_round  # noqa: E402


class Problem:
    """Seeded weights / features / labels.  Add-on kernels are xavier-uniform (util/func.py:8-10),
    classifier weights N(1, 0.1) on each child's own prototype slice and -0.5 elsewhere
    (pipnet/pipnet.py:1026, :1235-1248), features N(0,1) rounded to bf16 (the GEMM operand type)."""

    def __init__(self, tree, C, H, B, *, seed=0, num_features=0, per_child=0, per_desc=0, paired=True, V=None,
                 feat_scale=1.0, round_bf16=True, root=None):
        # `root`: a ready tree whose internal nodes already carry num_protos (arbitrary per-node counts)
        self.root = root if root is not None else make_tree(tree, num_features=num_features, per_child=per_child,
                                                            per_desc=per_desc)
        self.layout = build_layout(self.root)
        L = self.layout
        g = torch.Generator().manual_seed(seed)
        self.C, self.H, self.W = C, H, H
        self.V = 2 * B if V is None else V
        self.V_first = B if paired else (self.V + 1) // 2
        self.w = {}
        for name, pn in zip(L.node_names, L.P_n):
            bound = float(np.sqrt(6.0 / (C + int(pn))))
            w = ((torch.rand(int(pn), C, generator=g, dtype=torch.float64) * 2 - 1) * bound)
            self.w[name] = bf16_round(w) if round_bf16 else w.float().double()
        self.wc = {}
        for node in self.root.nodes_with_children():
            pn, cn = node.num_protos, node.num_children()
            w = 1.0 + 0.1 * torch.randn(cn, pn, generator=g, dtype=torch.float64)
            if node.num_protos_per_child:
                start = 0
                for child in node.children:
                    lab = node.children_to_labels[child.name]
                    end = start + node.num_protos_per_child[child.name]
                    w[lab, :start] = -0.5
                    w[lab, end:] = -0.5
                    start = end
            self.wc[node.name] = w.float().double()
        x = torch.randn(self.V, C, H, H, generator=g, dtype=torch.float64) * feat_scale
        self.x = bf16_round(x) if round_bf16 else x.float().double()
        ys = torch.randint(0, L.L, (self.V_first,), generator=g)
        self.ys = torch.cat([ys, ys])[: self.V] if paired else torch.randint(0, L.L, (self.V,), generator=g)
        self.label2name = {i: n for i, n in enumerate(L.leaf_names)}

    # flat views for the CUDA path
    def w_flat(self, device):
        return torch.cat([self.w[n] for n in self.layout.node_names]).float().to(device)

    def wc_flat(self, device):
        return torch.cat([self.wc[n].reshape(-1) for n in self.layout.node_names]).float().to(device)

    def features(self, device, dtype=torch.bfloat16, channels_last=True):
        x = self.x.to(device=device, dtype=dtype)
        return x.contiguous(memory_format=torch.channels_last) if channels_last else x.contiguous()

    def cat_nodes(self, d, dim=1):
        return torch.cat([d[n] for n in self.layout.node_names], dim=dim)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def argmax_report(arg_gpu, arg_ref, proto_ref_flat):
    """Argmax must be bit-exact; when an entry differs report the gap between the reference's best and
    the value at the location the GPU chose (a true tie up to accumulation order has gap ~1e-7)."""
    arg_gpu, arg_ref = arg_gpu.cpu().long(), arg_ref.cpu().long()
    bad = (arg_gpu != arg_ref).nonzero()
    gaps = []
    for v, p in bad.tolist():
        row = proto_ref_flat[v, p]
        gaps.append(float(row[arg_ref[v, p]] - row[arg_gpu[v, p]]) / max(float(row[arg_ref[v, p]]), 1e-30))
    return bad.shape[0], gaps


# --------------------------------------------------------------------------- model-level builders
# product-side fixtures (synthetic net / argparse namespace builders) live in pipnet_b200/fixtures.py so that bench.py's
# repo arm imports nothing from oracle/; re-exported here for the parity tests
from pipnet_b200.fixtures import IdentityBackbone, make_args, build_net  # noqa: E402,F401


def desc_loss_kwargs(args):
    """argparse strings of the shipped scripts' optional terms (--tanh_desc "y|w", --minimize_contrasting_set 'y|k|w',
    --mask_prune_overspecific 'y|epoch|boost') -> keyword arguments of `head_oracle.head_losses`"""
    kw = {}
    if 'y' in getattr(args, 'tanh_desc', 'n'):
        kw['tanh_desc_weight'] = float(args.tanh_desc.split('|')[1])
    if 'y' in getattr(args, 'minimize_contrasting_set', 'n'):
        f = args.minimize_contrasting_set.split('|')
        assert len(f) < 2 or int(f[1]) == 1, 'TOPK 1 only'
        kw['contrasting'] = float(f[2]) if len(f) > 2 else 0.1
    if 'y' in getattr(args, 'mask_prune_overspecific', 'n'):
        f = args.mask_prune_overspecific.split('|')
        kw['mask_prune'] = dict(start_epoch=int(f[1]) if len(f) > 1 else 0, boost=float(f[2]) if len(f) > 2 else None,
                                geometric='y' in getattr(args, 'geometric_mean_overspecificity_score', 'n'),
                                sg='y' in getattr(args, 'sg_before_masking', 'n'))
    return kw


def flat_gumbel(noise_by_node, nodes):
    """{node name -> {child label -> [P_n, 2]}} -> [sum C_n*P_n, 2] indexed like the flat classifier weights"""
    total = sum(n.num_protos * n.num_children() for n in nodes)
    out = torch.zeros(total, 2, dtype=torch.float64)
    off = 0
    for n in nodes:
        for c, g in noise_by_node.get(n.name, {}).items():
            out[off + c * n.num_protos: off + (c + 1) * n.num_protos] = g.double()
        off += n.num_protos * n.num_children()
    return out


def split_gumbel(flat, nodes):
    """inverse of `flat_gumbel` (all children present; unused entries are simply never read)"""
    out, off = {}, 0
    for n in nodes:
        out[n.name] = {c: flat[off + c * n.num_protos: off + (c + 1) * n.num_protos] for c in range(n.num_children())}
        off += n.num_protos * n.num_children()
    return out
