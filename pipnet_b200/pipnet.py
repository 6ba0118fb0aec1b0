"""`PIPNet` / `get_network` with the reference's public surface (`pipnet/pipnet.py:54-185`, `:1134-1258`)
over the fused sm_100a head.

What stays identical for callers (`main_dist.py`, `pipnet/train.py`, `util/args.py`, visualisers):
  * constructor signature, `forward(xs, inference=False, apply_overspecificity_mask=False)` returning
    `(features, proto_features, pooled, out)` keyed by node name, `get_joint_distribution`;
  * attributes `_net`, `_pool`, `_softmax`, `_multiplier`, `root`, `_num_classes` and per node
    `_<node>_add_on` (a real `nn.Conv2d`, so `util/func.py:8-10` xavier init and the optimizer's
    `dir(net.module)` scan `util/args.py:528-556` keep working), `_<node>_num_protos`,
    `_<node>_classification` (`NonNegLinear`), `_<node>_proto_presence`;
  * state-dict keys and shapes.

What changes underneath: the per-node Python loop is gone.  All add-on kernels alias one flat
[P, C] buffer, all classifier weights one flat vector; `forward` makes three kernel calls for the whole
tree (projection+softmax+pool, classifier, optional threshold) and never materialises the
[V, P_n, H, W] maps -- `proto_features[node]` is produced on demand for the visualisation tools.
Unsupported research variants raise `Exception` like the reference does for invalid flag combinations
(`pipnet/pipnet.py:104-105,135`); nothing silently falls back to PyTorch.
"""
from __future__ import annotations

import argparse
import os
from collections import OrderedDict
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import ops
from .layout import build_layout
from .node import Node


# --------------------------------------------------------------------------- small modules kept for API parity
class NonNegLinear(nn.Module):
    """Parameter holder with the reference layout (`pipnet/pipnet.py:1016-1036`): weight [out, in] ~ N(1, 0.1),
    `normalization_multiplier`, optional zero bias.  `PIPNet.forward` does NOT call this module -- all nodes go
    through one classifier kernel on the flat axes; `forward` here only serves tools that call a single node's
    layer directly (`util/vis_hpipnet.py:62-127`, batch size 1)."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, device=None, dtype=None) -> None:
        kw = {'device': device, 'dtype': dtype}
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty((out_features, in_features), **kw))
        nn.init.normal_(self.weight, mean=1.0, std=0.1)
        self.normalization_multiplier = nn.Parameter(torch.ones((1,), requires_grad=True))
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_features, **kw))
        else:
            self.register_parameter('bias', None)

    def forward(self, input: Tensor) -> Tensor:
        return F.linear(input, torch.relu(self.weight), self.bias)


class _GatherParams(torch.autograd.Function):
    """Autograd bridge from the per-node `nn.Parameter`s (which alias a flat buffer) to that flat buffer:
    forward is zero-copy, backward hands each parameter its slice of the flat gradient."""

    @staticmethod
    def forward(ctx, holder, *params):
        ctx.meta = holder.meta
        ctx.n_params = len(params)
        # no zero-filled gradients when every consumer delivered its gradient out of band (ops._GradBucket)
        ctx.set_materialize_grads(False)
        return holder.flat.detach()

    @staticmethod
    def backward(ctx, g):
        if g is None:
            return (None,) * (1 + ctx.n_params)
        g = g.contiguous()
        flat = g.view(-1)
        grads = []
        for i, (off, numel, shape) in enumerate(ctx.meta):
            grads.append(flat[off:off + numel].view(shape) if ctx.needs_input_grad[i + 1] else None)
        return (None, *grads)


class _FlatGroup:
    """A set of parameters re-homed into one contiguous buffer (`param.data` becomes a view).  The
    `Parameter` objects are never replaced, so optimizers built before or after keep working."""

    def __init__(self, params: List[nn.Parameter]):
        self.params = params
        self.flat: Optional[Tensor] = None
        self.meta = []
        self._gathered: Optional[Tensor] = None
        self._gathered_grad = None
        self.family = None       # all flat groups of the model (set by PIPNet): layout of the data-parallel gradient bucket

    def ensure(self):
        p0 = self.params[0]
        ok = self.flat is not None and self.flat.device == p0.device
        if ok:
            base, es = self.flat.data_ptr(), self.flat.element_size()
            for p, (off, numel, _) in zip(self.params, self.meta):
                if p.data_ptr() != base + off * es:
                    ok = False
                    break
        if ok:
            return self.flat
        # (re)build: happens once after construction and again after `.to(device)` / load_state_dict(assign=True)
        with torch.no_grad():
            flat = torch.cat([p.detach().reshape(-1).float() for p in self.params])
            meta, off = [], 0
            for p in self.params:
                n = p.numel()
                meta.append((off, n, tuple(p.shape)))
                p.data = flat[off:off + n].view(p.shape)
                off += n
        self.flat, self.meta = flat, meta
        return flat

    def gather(self) -> Tensor:
        """Flat view with autograd edges to the parameters.  Cached until `invalidate()` (called at the start of
        every forward) so that several consumers in one step (head GEMM, orth loss) share ONE autograd node and
        their gradients are summed once on the flat buffer instead of once per node."""
        self.ensure()
        if self._gathered is None or not torch.is_grad_enabled() or self._gathered_grad != torch.is_grad_enabled():
            self._gathered = _GatherParams.apply(self, *self.params)
            self._gathered._hc_group = self      # lets the kernels' autograd functions find the parameters (ops._GradBucket)
            self._gathered_grad = torch.is_grad_enabled()
        return self._gathered

    def invalidate(self):
        self._gathered = None


class NodeDict(dict):
    """Read-only mapping node name -> column slice of a flat [V, total] tensor, shaped like the dicts the
    reference returns (`pipnet/pipnet.py:115-116`).  `.flat` is what the fused losses consume.
    It IS a `dict` (its own storage stays empty, the slices are made on demand): `DistributedDataParallel(...,
    find_unused_parameters=True)` -- how main_dist.py:330 wraps the model -- walks the forward outputs looking for
    tensors and only descends into tuples / lists / dicts; an opaque object would hide every output from its reducer
    and all parameters would be declared unused (tests/test_gpu_ddp.py)."""

    def __init__(self, flat: Tensor, names: List[str], offsets: np.ndarray):
        super().__init__()
        self.flat, self._names, self._off = flat, names, offsets
        self._idx = {n: i for i, n in enumerate(names)}
        self._cache: Dict[str, Tensor] = {}

    def __getitem__(self, name):
        t = self._cache.get(name)
        if t is None:
            i = self._idx[name]
            t = self.flat[:, int(self._off[i]):int(self._off[i + 1])]
            self._cache[name] = t
        return t

    def __contains__(self, name): return name in self._idx
    def __iter__(self): return iter(self._names)
    def __len__(self): return len(self._names)
    def __bool__(self): return len(self._names) > 0
    def keys(self): return list(self._names)
    def values(self): return [self[n] for n in self._names]
    def items(self): return [(n, self[n]) for n in self._names]
    def get(self, name, default=None): return self[name] if name in self._idx else default

    def _read_only(self, *a, **k):
        raise TypeError('NodeDict is a read-only view of the flat head outputs')
    __setitem__ = __delitem__ = pop = popitem = clear = update = setdefault = _read_only


class LazyProtoFeatures:
    """`proto_features[node]` = softmaxed map [V, P_n, H, W].  The fused head never writes these maps;
    one is rebuilt on access for the callers that really want it (`main_dist.py:444` reads a shape,
    the visualisers read batch-1 maps).  Not differentiable by design."""

    def __init__(self, net: "PIPNet", features: Tensor, tau: float, argmax: NodeDict):
        self._net, self._features, self._tau = net, features, tau
        self.argmax = argmax                      # node -> [V, P_n] flat h*W+w location of the max
        self._names = list(argmax.keys())
        self._cache: Dict[str, Tensor] = {}

    def __getitem__(self, name):
        t = self._cache.get(name)
        if t is None:
            conv = getattr(self._net, '_' + name + '_add_on')
            t = ops.materialize_map(self._features, conv.weight, self._tau)
            self._cache[name] = t
        return t

    def shape_of(self, name):
        V, _, H, W = self._features.shape
        return (V, getattr(self._net, '_' + name + '_num_protos'), H, W)

    def __contains__(self, name): return name in self._names
    def __iter__(self): return iter(self._names)
    def __len__(self): return len(self._names)
    def keys(self): return list(self._names)
    def items(self): return [(n, self[n]) for n in self._names]


# --------------------------------------------------------------------------- the model
_UNSUPPORTED = (('gumbel_softmax', 'y'), ('multiply_cs_softmax', 'y'), ('focal', 'y'), ('softmax_over_channel', 'y'))


class PIPNet(nn.Module):
    def __init__(self,
                 num_classes: int,
                 num_prototypes: int,
                 feature_net: nn.Module,
                 args: argparse.Namespace,
                 add_on_layers: dict,
                 pool_layer: nn.Module,
                 classification_layers: dict,
                 num_parent_nodes: int,
                 root: Node):
        super().__init__()
        assert num_classes > 0
        self._num_classes = num_classes
        self._net = feature_net
        for node_name, layer in add_on_layers.items():
            if type(layer) is not nn.Conv2d or layer.kernel_size != (1, 1):
                raise Exception('the B200 head supports plain 1x1 nn.Conv2d prototype layers only '
                                '(UnitConv2D / L2Conv2D / ProjectConv2D are out of scope, SURVEY.md section 2.1)')
            if layer.bias is not None:
                raise Exception('--add_on_bias is not supported by the fused head')
            setattr(self, '_' + node_name + '_add_on', layer)
            setattr(self, '_' + node_name + '_num_protos', layer.weight.shape[0])
        self._pool = pool_layer
        self._avg_pool = nn.Sequential(nn.AdaptiveAvgPool2d(output_size=(1, 1)), nn.Flatten())
        for node_name, layer in classification_layers.items():
            setattr(self, '_' + node_name + '_classification', layer)
        self._multiplier = nn.Parameter(torch.ones((1,), requires_grad=True))
        if args.softmax.split('|')[0] == 'y':
            self._softmax = nn.Softmax(dim=1)
        else:
            raise Exception('the B200 head implements the --softmax "y|tau" recipe only')
        for flag, bad in _UNSUPPORTED:
            if getattr(args, flag, 'n') == bad:
                raise Exception(f'--{flag} {bad} is not supported by the B200 head')
        self._num_parent_nodes = num_parent_nodes
        self.root = root
        for node_name in add_on_layers:
            pp = nn.Parameter(torch.zeros(getattr(self, '_' + node_name + '_num_protos'), 2), requires_grad=True)
            nn.init.xavier_normal_(pp, gain=1.0)
            setattr(self, '_' + node_name + '_proto_presence', pp)
        self.args = args
        self.conc_log_ip = ('y' in getattr(args, 'conc_log_ip', 'n'))

        parts = args.softmax.split('|')
        self.softmax_tau = float(int(parts[1])) if len(parts) > 1 else 0.2        # pipnet/pipnet.py:131-136
        # ---- flat layout (host tables) + parameter groups
        self.layout = build_layout(root)
        names = self.layout.node_names
        if list(add_on_layers.keys()) != names or list(classification_layers.keys()) != names:
            raise Exception('add_on_layers / classification_layers must follow root.nodes_with_children() order')
        self._has_cls_bias = any(getattr(self, '_' + n + '_classification').bias is not None for n in names)
        self._w_group = _FlatGroup([getattr(self, '_' + n + '_add_on').weight for n in names])
        self._wc_group = _FlatGroup([getattr(self, '_' + n + '_classification').weight for n in names])
        self._bias_group = (_FlatGroup([getattr(self, '_' + n + '_classification').bias for n in names])
                            if self._has_cls_bias else None)
        self._pp_group = _FlatGroup([getattr(self, '_' + n + '_proto_presence') for n in names])
        family = [g for g in (self._w_group, self._wc_group, self._bias_group, self._pp_group) if g is not None]
        for g in family:
            g.family = family
        self._dl: Optional[ops.DeviceLayout] = None
        self._orth_hint = False     # the last calculate_loss used the kernel-orthogonality term (ops.orth_prefetch)
        # 'bf16': bf16 GEMM operands (default, the benchmarked path); 'fp32': fp32-accurate projection (3-way bf16 split
        # operands, six cross terms through the same tcgen05 kernel) for the <= 1e-5 contract on fp32 inputs
        self.head_precision = getattr(args, 'head_precision', 'bf16')
        # Column selection of get_joint_distribution, replicating the reference exactly: np.argsort over
        # names_of_joint_distribution() (pipnet/pipnet.py:179-181).  For trees without single-child nodes this is
        # the sorted-leaf order the kernel already produces (None = identity); with a single-child node the
        # reference's name list is truncated (util/node.py:397-403) and so is its result.
        dfs = [n.name for n in self._dfs_leaves(root)]
        ref_names = root.unwrap_names_of_joint(root.names_of_joint_distribution())
        order = sorted(range(len(ref_names)), key=lambda i: ref_names[i])
        sorted_pos = {nm: i for i, nm in enumerate(self.layout.leaf_names)}
        cols = [sorted_pos[dfs[i]] for i in order]
        self._joint_cols = None if cols == list(range(self.layout.L)) else cols

    @staticmethod
    def _dfs_leaves(root):
        out, stack = [], [root]
        while stack:
            n = stack.pop()
            if n.is_leaf():
                out.append(n)
            else:
                stack.extend(reversed(n.children))
        return out

    # ------------------------------------------------------------------ plumbing
    def device_layout(self, device) -> ops.DeviceLayout:
        if self._dl is None or self._dl.device != torch.device(device):
            self._dl = ops.DeviceLayout(self.layout, device)
        return self._dl

    def flat_prototype_kernels(self) -> Tensor:
        """[P, C] view of all add-on kernels with autograd edges to the per-node parameters."""
        C = self._w_group.params[0].shape[1]
        v = self._w_group.gather().view(self.layout.P, C)
        v._hc_group = self._w_group
        return v

    def flat_classifier_weights(self) -> Tensor:
        return self._wc_group.gather()

    def flat_presence_logits(self) -> Tensor:
        """[P, 2] view of all `_<node>_proto_presence` logits with autograd edges to the per-node parameters.
        (The name must NOT end in `_proto_presence` / `_add_on` / `_classification`: the reference discovers parameters
        with `dir(net.module)` + suffix matching, util/args.py:528-556, main_dist.py:474-481 -- tests/test_reference_optimizer.py.)"""
        v = self._pp_group.gather().view(self.layout.P, 2)
        v._hc_group = self._pp_group
        return v

    # ------------------------------------------------------------------ forward
    def head(self, features: Tensor, *, inference=False, labels: Optional[ops.LabelTables] = None,
             V_first: Optional[int] = None, classify: bool = False):
        """The fused head on backbone features.  Returns flat tensors:
        pooled [V,P], align [N] (zeros unless `labels` given), argmax [V,P] int32, dl -- and with `classify=True` also
        out [V,K] (the per-node NonNegLinear, finished by the same launch that unpacks the pooled table)."""
        V = features.shape[0]
        dl = self.device_layout(features.device)
        if V_first is None:
            V_first = labels.V_first if labels is not None else (V + 1) // 2
        x = features.detach() if getattr(self.args, 'sg_before_protos', 'n') == 'y' else features
        w_flat = self.flat_prototype_kernels()
        prec = ops.PREC_FP32X3 if self.head_precision == 'fp32' else ops.PREC_BF16
        wc = bias = None
        if classify:
            wc = self.flat_classifier_weights()
            bias = self._bias_group.gather() if self._bias_group is not None else None
            if (labels is not None and self._orth_hint and torch.is_grad_enabled() and w_flat.requires_grad
                    and features.is_cuda):
                # the last loss used the kernel-orthogonality term: start its weights-only part on a side stream.  Small
                # trees: inside the fused forward, right after K1 (forked before K1 its blocks delay K1's persistent CTAs
                # by ~5 us, more than the 6 us Gram kernel is worth).  Large trees (cub190: 31 us of Gram kernel against a
                # 177 us K1): now, so that it runs beside K1 instead of in front of the loss kernel.
                if dl.N * dl.layout.p_max * dl.layout.p_max >= 40000:
                    ops.orth_prefetch(w_flat, wc, dl, V)
                else:
                    ops.ORTH_REQUEST = (w_flat, wc, dl, V)
        pooled, align, argmax, out = ops.HeadProjPool.apply(x, w_flat, dl, V_first, self.softmax_tau, labels,
                                                            0.1 if inference else 0.0, prec, wc, bias)
        if classify:
            return pooled, align, argmax, dl, out
        return pooled, align, argmax, dl

    def classify(self, pooled_flat: Tensor, dl: ops.DeviceLayout) -> Tensor:
        bias = self._bias_group.gather() if self._bias_group is not None else None
        return ops.NonNegClassifier.apply(pooled_flat, self.flat_classifier_weights(), bias, dl)

    def forward(self, xs, inference=False, apply_overspecificity_mask=False, labels: Optional[ops.LabelTables] = None):
        for grp in (self._w_group, self._wc_group, self._bias_group, self._pp_group):
            if grp is not None:
                grp.invalidate()
        features = self._net(xs)
        if apply_overspecificity_mask:
            pooled_flat, align, argmax, dl = self.head(features, inference=inference, labels=labels)
            # Gumbel hard sample on proto_presence (pipnet/pipnet.py:164-166), one draw for the whole flat axis
            pres = self.flat_presence_logits()
            mask = F.gumbel_softmax(pres, tau=0.5, hard=True, dim=-1)[:, 1].unsqueeze(0)
            pooled_flat = mask * pooled_flat
            out_flat = self.classify(pooled_flat, dl)
        else:
            pooled_flat, align, argmax, dl, out_flat = self.head(features, inference=inference, labels=labels, classify=True)
        L = self.layout
        pooled = NodeDict(pooled_flat, L.node_names, L.proto_off)
        out = NodeDict(out_flat, L.node_names, L.cls_off)
        # `out` is this model's classifier applied to exactly `pooled.flat`: lets calculate_loss chain the class term's
        # backward through the classifier in one launch (ops.HeadLosses, chain=True)
        out.chained_from = pooled_flat
        pooled.align = align                                    # per-node align_pf loss (zeros without labels)
        pooled.align_valid = labels is not None
        proto_features = LazyProtoFeatures(self, features, self.softmax_tau, NodeDict(argmax, L.node_names, L.proto_off))
        return features, proto_features, pooled, out

    def get_joint_distribution(self, out, leave_out_classes=None, apply_overspecificity_mask=False, device='cuda',
                               softmax_tau=1, presence_mask: Optional[Tensor] = None):
        """(`pipnet/pipnet.py:173-185`) -> (out['root'], [V, L] joint leaf probabilities, sorted-leaf columns).
        `leave_out_classes` / `apply_overspecificity_mask` follow `util/node.py:300-385`: they replace the child
        probabilities of whole nodes (one-hot at a left-out leaf child / leaf-count fractions when the hard Gumbel
        presence mask wipes out a class), which the kernel takes as a [K] override table.  `presence_mask` ([P] 0/1)
        injects the mask instead of drawing it (one flat Gumbel draw here, one per node in the reference)."""
        flat = out.flat if isinstance(out, NodeDict) else torch.cat([out[n] for n in self.layout.node_names], dim=1)
        dl = self.device_layout(flat.device)
        override = None
        if leave_out_classes:
            override = self._leave_out_override(tuple(leave_out_classes)).to(flat.device)
        if apply_overspecificity_mask or presence_mask is not None:
            m_ovr = self._mask_override(dl, presence_mask)
            # the leave-out rule is checked first in the reference (util/node.py:319), the mask only below it
            override = m_ovr if override is None else torch.where(override >= 0, override, m_ovr)
        joint, _ = ops.joint_leaf_distribution(flat, dl, float(softmax_tau), override)
        if self._joint_cols is not None:
            joint = joint[:, torch.as_tensor(self._joint_cols, device=joint.device)]
        return out['root'], joint

    def _leave_out_override(self, leave_out: tuple) -> Tensor:
        cache = self.__dict__.setdefault('_leave_out_cache', {})
        if leave_out not in cache:
            lo = set(leave_out)
            L = self.layout
            ovr = np.full(L.K, -1.0, dtype=np.float32)
            for i, node in enumerate(self.root.nodes_with_children()):
                if any(set(c.leaf_descendents).issubset(lo) for c in node.children):
                    left = [c for c in node.children if c.is_leaf() and c.name in lo]
                    if not left:
                        raise Exception(f'node {node.name}: a whole non-leaf child is left out; the reference indexes an '
                                        f'empty list there (util/node.py:321)')
                    k0, k1 = int(L.cls_off[i]), int(L.cls_off[i + 1])
                    ovr[k0:k1] = 0.0
                    ovr[k0 + node.children_to_labels[left[0].name]] = 1.0
            cache[leave_out] = torch.from_numpy(ovr)
        return cache[leave_out]

    def _mask_override(self, dl, presence_mask: Optional[Tensor]) -> Tensor:
        """[K] override: leaf-count fractions for nodes where the masked classifier has an all-<=1e-3 class row."""
        with torch.no_grad():
            dev = dl.device
            if presence_mask is None:
                presence_mask = F.gumbel_softmax(self.flat_presence_logits().detach(), tau=0.5, hard=True, dim=-1)[:, 1]
            m = presence_mask.to(device=dev, dtype=torch.float32)
            wc = self.flat_classifier_weights().detach()
            alive = ((m[dl.welem_proto.long()] * wc) > 1e-3).float()                       # [n_welems]
            col_alive = torch.zeros(dl.K, device=dev).scatter_reduce_(0, dl.welem_col.long(), alive, reduce='amax')
            node_dead = torch.zeros(dl.N, device=dev).scatter_reduce_(0, dl.col_node.long(), 1.0 - col_alive, reduce='amax')
            nl = dl.col_nleaves.float()
            node_leaves = torch.zeros(dl.N, device=dev).scatter_add_(0, dl.col_node.long(), nl)
            frac = nl / node_leaves[dl.col_node.long()]
            return torch.where(node_dead[dl.col_node.long()] > 0, frac, torch.full_like(frac, -1.0))

    def get_classification_layers(self):
        return [getattr(self, attr) for attr in dir(self) if attr.endswith('_classification')]


# --------------------------------------------------------------------------- backbones (hand-off only)
def _relax_strides(model: nn.Module, threshold: int) -> nn.Module:
    """Halve the stride of every stride-2 conv with more than `threshold` input channels: ConvNeXt-tiny then
    ends at 26x26 (threshold 100) or 13x13 (300) for 224px input (`features/convnext_features.py:7-16`)."""
    for m in model.modules():
        if isinstance(m, nn.Conv2d) and m.stride[0] == 2 and m.in_channels > threshold:
            m.stride = tuple(s // 2 for s in m.stride)
    return model


def _convnext_tiny(threshold, pretrained=False):
    from torchvision import models
    weights = models.ConvNeXt_Tiny_Weights.DEFAULT if pretrained else None
    model = models.convnext_tiny(weights=weights)
    model.avgpool = nn.Identity()
    model.classifier = nn.Identity()
    if threshold is not None:
        _relax_strides(model, threshold)
    if os.environ.get('HC_FUSE_BACKBONE_TAIL', '1') != '0':
        fuse_convnext_tail(model)
    return model


def fuse_convnext_tail(model) -> bool:
    """Backbone hand-off: make the LAST ConvNeXt block (`features.7.2`, the one `util/args.py:503` names) emit
    the head's feature matrix itself -- its `layer_scale * block(x)`, stochastic depth and residual add run as ONE kernel
    that writes bf16 channels-last rows (`ops.ScaleResidualRows`), so nothing (no scale / add / cast / layout pass) sits
    between the backbone and the projection kernel.  Same parameters, same state_dict; CPU tensors keep the stock path."""
    try:
        from torchvision.models.convnext import CNBlock
    except Exception:
        return False
    blocks = [m for m in model.modules() if isinstance(m, CNBlock)]
    if not blocks:
        return False
    last = blocks[-1]

    def forward(input):
        y = last.block(input)
        if not y.is_cuda:
            return last.stochastic_depth(last.layer_scale * y) + input
        keep = None
        p_drop = float(getattr(last.stochastic_depth, 'p', 0.0))
        if last.training and p_drop > 0.0:                      # torchvision.ops.stochastic_depth, mode "row"
            survival = 1.0 - p_drop
            keep = torch.empty(y.shape[0], device=y.device, dtype=torch.float32).bernoulli_(survival)
            if survival > 0.0:
                keep.div_(survival)
        return ops.ScaleResidualRows.apply(y, input, last.layer_scale, keep)

    last.forward = forward
    return True


def convnext_tiny_26_features(pretrained=False, **kw): return _convnext_tiny(100, pretrained)
def convnext_tiny_13_features(pretrained=False, **kw): return _convnext_tiny(300, pretrained)
def convnext_tiny_7_features(pretrained=False, **kw): return _convnext_tiny(None, pretrained)


class _ResNetFeatures(nn.Module):
    def __init__(self, name, pretrained):
        super().__init__()
        from torchvision import models
        m = getattr(models, name)(weights='DEFAULT' if pretrained else None)
        self.stem = nn.Sequential(m.conv1, m.bn1, m.relu, m.maxpool)
        self.layers = nn.Sequential(m.layer1, m.layer2, m.layer3, m.layer4)
        # the reference's ResNet feature nets end at 28x28 for 224px input: undo the last two strides
        for blk in (m.layer3[0], m.layer4[0]):
            for mod in blk.modules():
                if isinstance(mod, nn.Conv2d) and mod.stride == (2, 2):
                    mod.stride = (1, 1)

    def forward(self, x):
        return self.layers(self.stem(x))


base_architecture_to_features = {
    'convnext_tiny_26': convnext_tiny_26_features,
    'convnext_tiny_13': convnext_tiny_13_features,
    'convnext_tiny_7': convnext_tiny_7_features,
    'resnet18': lambda pretrained=False: _ResNetFeatures('resnet18', pretrained),
    'resnet34': lambda pretrained=False: _ResNetFeatures('resnet34', pretrained),
    'resnet50': lambda pretrained=False: _ResNetFeatures('resnet50', pretrained),
    'resnet101': lambda pretrained=False: _ResNetFeatures('resnet101', pretrained),
}


def get_network(num_classes: int, args: argparse.Namespace, root=None):
    """Same contract as `pipnet/pipnet.py:1134-1258`: returns
    (feature_net, add_on_layers, pool_layer, classification_layers, num_prototypes) with one bias-free 1x1
    `nn.Conv2d` and one `NonNegLinear` per internal node; with `--protopool n` each child's classifier row
    keeps N(1, 0.1) on its own prototype slice and -0.5 elsewhere (`:1235-1248`)."""
    for flag in ('unitconv2d', 'projectconv2d', 'l2conv2d'):
        if getattr(args, flag, 'n') == 'y':
            raise Exception(f'--{flag} y is not supported by the B200 head (plain 1x1 conv prototypes only)')
    if getattr(args, 'basic_cnext_gaussian_multiplier', '') != '' or getattr(args, 'stage4_reducer_net', '') != '':
        raise Exception('gaussian-multiplier / stage4 reducer backbones are outside the B200 head scope')
    if getattr(args, 'classifier', 'NonNegative') == 'Linear':
        raise Exception('--classifier Linear is not supported (NonNegative only)')
    if getattr(args, 'add_on_bias', False):
        raise Exception('--add_on_bias is not supported by the fused head')
    if args.net not in base_architecture_to_features:
        raise Exception('other base architecture NOT implemented')
    features = base_architecture_to_features[args.net](pretrained=not args.disable_pretrained)
    in_channels = [m for m in features.modules() if isinstance(m, nn.Conv2d)][-1].out_channels
    num_prototypes = in_channels if args.num_features == 0 else args.num_features

    parent_nodes = root.nodes_with_children()
    add_on_layers = OrderedDict()
    for node in parent_nodes:
        add_on_layers[node.name] = nn.Conv2d(in_channels, node.num_protos, kernel_size=1, stride=1, padding=0, bias=False)
    pool_layer = nn.Sequential(nn.AdaptiveMaxPool2d(output_size=(1, 1)), nn.Flatten())
    classification_layers = OrderedDict()
    for node in parent_nodes:
        layer = NonNegLinear(node.num_protos, node.num_children(), bias=bool(args.bias))
        if args.protopool == 'n':
            with torch.no_grad():
                start = 0
                for child in node.children:
                    lab = node.children_to_labels[child.name]
                    end = start + node.num_protos_per_child[child.name]
                    layer.weight[lab, :start] = -0.5
                    layer.weight[lab, end:] = -0.5
                    start = end
        classification_layers[node.name] = layer
    return features, add_on_layers, pool_layer, classification_layers, num_prototypes
