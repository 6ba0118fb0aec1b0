"""`train_pipnet` / `test_pipnet` / `calculate_loss` with the reference's signatures
(`pipnet/train.py:73-77`, `:525-530`, `:852-855`) on top of the fused head.

Differences that matter for speed, not for results:
  * the per-node Python loop of `calculate_loss` (`pipnet/train.py:933-1194`: host-built masks, H2D per node,
    `.item()` per node per loss, `.cpu()` per child) is replaced by a handful of kernel calls on flat
    [V,P] / [V,K] tensors driven by label tables built on the device (`ops.LabelTables`);
  * nothing synchronises with the host inside a step: per-node loss values are handed back as lazy
    dictionaries that copy one [N] vector on first access, epoch statistics are accumulated on the device;
  * the per-step `barrier` + broadcast of every parameter (`pipnet/train.py:54-65`) is dropped: its rank-0
    mutation tests `name.endswith('_classification')` on *parameter* names and never fires.

Loss terms implemented as kernels: align_pf, tanh, kernel_orth, class, and the three extra terms the shipped
scripts switch on -- tanh_desc, minimize_contrasting_set (TOPK 1), mask-prune overspecificity (`ops.DescLosses`,
pipnet/train.py:946-1060, 1089-1133).  OOD, BYOL, feature-level align/uni and the research-only terms raise.
"""
from __future__ import annotations

import os
from collections import defaultdict
from typing import Dict, Optional

import numpy as np
import torch

from . import ops
from .pipnet import NodeDict, PIPNet

OOD_LABEL = -1


# --------------------------------------------------------------------------- lazy per-node values
class _Scalar:
    __slots__ = ('v',)

    def __init__(self, v): self.v = float(v)
    def item(self): return self.v
    def __float__(self): return self.v
    def __repr__(self): return f'{self.v:.6f}'


class LazyNodeLosses:
    """dict-like {node name: value} over a device vector [N]; only nodes with descendants in the batch are
    present (the reference `continue`s over the others, `pipnet/train.py:941-942`).  One D2H copy on first use."""

    def __init__(self, values: Optional[torch.Tensor], n_desc: torch.Tensor, names):
        self._values, self._n_desc, self._names, self._d = values, n_desc, names, None

    def _mat(self):
        if self._d is None:
            if self._values is None:
                self._d = {}
            else:
                both = torch.stack([self._values.detach().float(), self._n_desc.float()]).cpu().numpy()
                self._d = {n: _Scalar(both[0, i]) for i, n in enumerate(self._names) if both[1, i] > 0}
        return self._d

    def items(self): return self._mat().items()
    def keys(self): return self._mat().keys()
    def values(self): return self._mat().values()
    def __getitem__(self, k): return self._mat()[k]
    def __contains__(self, k): return k in self._mat()
    def __len__(self): return len(self._mat())
    def __iter__(self): return iter(self._mat())

    def mean_tensor(self):
        """mean over present nodes as a device scalar (what `np.mean([... .item() ...])` computes)"""
        if self._values is None:
            return None
        act = (self._n_desc > 0).float()
        return (self._values.detach() * act).sum() / act.sum().clamp_min(1.0)


class LazyMean:
    """mean over the nodes present in the batch (what the reference computes with `np.mean([v.item() ...])`),
    evaluated only when somebody asks for it"""

    def __init__(self, d: LazyNodeLosses): self._d = d
    def tensor(self): return self._d.mean_tensor()
    def item(self):
        t = self._d.mean_tensor()
        return -5.0 if t is None else float(t)
    def __float__(self): return self.item()
    def __bool__(self): return True
    def __repr__(self): return f'{self.item():.6f}'


class _LossResult(tuple):
    """the reference's 21-tuple, plus the raw device statistics for sync-free epoch accounting"""

    def __new__(cls, items, stats, n_desc):
        self = super().__new__(cls, items)
        self.stats, self.n_desc = stats, n_desc
        return self


def _unwrap(net):
    return net.module if hasattr(net, 'module') else net


def _reject(flag_name, on):
    if on:
        raise Exception(f'{flag_name} is not part of the fused B200 head path yet (SURVEY.md section 8f); '
                        f'run it on the reference or disable it')


def make_labels(net, ys: torch.Tensor, V_first: Optional[int] = None) -> ops.LabelTables:
    m = _unwrap(net)
    V = ys.numel()
    return ops.LabelTables(ys, m.device_layout(ys.device), V // 2 if V_first is None else V_first)


# --------------------------------------------------------------------------- the loss
_MULT_CACHE = {}


def _multiplier_value(m) -> float:
    """`net._multiplier` as a host float (exponent of log1p(out ** m), pipnet/train.py:1158).  The reference keeps it in a
    frozen nn.Parameter; reading it costs a device sync, so the value is cached per tensor and re-read only when the
    tensor's version counter says it was written (optimizer step, load_state_dict, .fill_)."""
    if m is None:
        return 2.0
    if not torch.is_tensor(m):
        return float(m)
    key = id(m)
    ver = m._version
    hit = _MULT_CACHE.get(key)
    if hit is not None and hit[0] == ver and hit[2]() is m:
        return hit[1]
    import weakref
    val = float(m.detach().reshape(-1)[0])          # no gradient flows to it: every driver path of the reference freezes
    _MULT_CACHE[key] = (ver, val, weakref.ref(m))   # it at 2 (main_dist.py:344, :381-382, :402-403, :426-427)
    return val


def calculate_loss(epoch, net, additional_network_outputs, features, proto_features, pooled, out, ys, align_weight,
                   align_pf_weight, t_weight, mm_weight, unif_weight, cl_weight, OOD_loss_weight, orth_weight,
                   cluster_desc_weight, sep_desc_weight, subspace_sep_weight, byol_weight, net_normalization_multiplier,
                   pretrain, finetune, criterion, train_iter, print=True, EPS=1e-10, root=None, label2name=None,
                   node_accuracy=None, OOD_loss_required=False, kernel_orth=False, tanh_desc=False, align=True, uni=True,
                   align_pf=False, tanh=False, minmaximize=False, cluster_desc=False, sep_desc=False, subspace_sep=False,
                   byol=False, train=True, args=None, device=None, labels: Optional[ops.LabelTables] = None,
                   gumbel_noise: Optional[torch.Tensor] = None):
    """Same positional contract and 21-tuple as `pipnet/train.py:852-1341`.  `pooled` / `out` are the
    `NodeDict`s returned by `PIPNet.forward`; `labels` are the device label tables of this batch (built here
    from `ys` when not given).  `criterion` is accepted for signature parity: the class term is always the
    weighted NLL of `util/custom_losses.py:17-34`."""
    m: PIPNet = _unwrap(net)
    _reject('align/uni', (not finetune) and (align or uni))
    _reject('byol', (not finetune) and byol)
    _reject('minmaximize / cluster_desc / sep_desc / subspace_sep', minmaximize or cluster_desc or sep_desc or subspace_sep)
    _reject('OOD loss', OOD_loss_required)
    if args is not None:
        _reject('--OOD_ent', 'y' in getattr(args, 'OOD_ent', 'n'))
    if not isinstance(pooled, NodeDict) or not isinstance(out, NodeDict):
        raise Exception('calculate_loss expects the outputs of pipnet_b200.PIPNet.forward')
    dl = m.device_layout(pooled.flat.device)
    if labels is None:
        labels = make_labels(net, ys)
    N = dl.N
    names = m.layout.node_names
    out_device = pooled.flat.device
    losses_used = []

    flags = 0
    wts = [0.0, 0.0, 0.0, 0.0]
    align_vec = None
    use_align = (not finetune) and align_pf
    if use_align:
        align_vec = getattr(pooled, 'align', None)
        if align_vec is None or not getattr(pooled, 'align_valid', False):
            raise Exception('align_pf needs the per-node align loss of the fused forward: call net(xs, labels=...)')
        wts[0] = align_pf_weight / N
        losses_used.append('AL_PF')
    # ---- descendant-structured terms of the shipped scripts (tanh_desc, contrasting set, mask pruning)
    dflags, dw, boost = 0, [0.0, 0.0, 2.0 / N, 0.5 / N], 0.0
    mp_arg = getattr(args, 'mask_prune_overspecific', 'n') if args is not None else 'n'
    cs_arg = getattr(args, 'minimize_contrasting_set', 'n') if args is not None else 'n'
    if (not pretrain) and 'y' in mp_arg:                                           # pipnet/train.py:946-1015
        if getattr(args, 'protopool', 'n') == 'y':
            raise Exception('--mask_prune_overspecific cannot be combined with --protopool y (pipnet/train.py:947)')
        f = mp_arg.split('|')
        if not (len(f) > 1 and epoch < int(f[1])):
            dflags |= ops.DESC_MASK_PRUNE
            if len(f) > 2:
                boost = float(f[2])
            elif 'y' in getattr(args, 'geometric_mean_overspecificity_score', 'n'):
                dflags |= ops.DESC_GEOMETRIC
            if 'y' in getattr(args, 'sg_before_masking', 'n'):
                dflags |= ops.DESC_SG_SCORE
            losses_used.append('MASK_PRUNING')
    if (not pretrain) and (not finetune) and 'y' in cs_arg:                        # pipnet/train.py:1017-1060
        f = cs_arg.split('|')
        if len(f) > 1 and int(f[1]) != 1:
            raise Exception('--minimize_contrasting_set with TOPK != 1 is not supported by the B200 head')
        dflags |= ops.DESC_CONTRAST
        dw[1] = (float(f[2]) if len(f) > 2 else 0.1) / N
        losses_used.append('MIN_CONT')
        # the reference re-binds its local EPS to 1e-12 inside this block (:1025), BEFORE the tanh / tanh_desc terms of
        # the same loop iteration read it (:1080, :1108): with the term on they all see 1e-12
        EPS = 1e-12
    use_tanh = (not finetune) and tanh and not (getattr(args, 'tanh_during_second_phase', 'y') == 'n' and not pretrain)
    if use_tanh:
        flags |= ops.LOSS_TANH
        wts[1] = t_weight / N
        losses_used.append('TANH')
    use_orth = (not pretrain) and (not finetune) and kernel_orth
    if use_orth:
        flags |= ops.LOSS_ORTH
        wts[2] = orth_weight / N
        losses_used.append('KO')
    use_cls = not pretrain
    if use_cls:
        flags |= ops.LOSS_CLASS
        wts[3] = cl_weight / N
        losses_used.append('CL')
    if not (args is not None and getattr(args, 'pipnet_sparsity', 'y') == 'n'):
        flags |= ops.LOSS_SPARSITY
    mult = _multiplier_value(net_normalization_multiplier) if (use_cls and (flags & ops.LOSS_SPARSITY)) else 2.0
    m._orth_hint = bool(use_orth)             # the next head forward starts the term's weights-only part beside K1
    # `out` is the model's classifier applied to `pooled.flat` (PIPNet.forward says so): the class term's backward goes
    # straight through the classifier inside the fused loss backward, d loss / d out is never materialised
    chain = getattr(out, 'chained_from', None) is pooled.flat
    bias = m._bias_group.gather() if (chain and m._bias_group is not None) else None
    loss, stats, n_correct = ops.HeadLosses.apply(pooled.flat, out.flat, align_vec, m.flat_prototype_kernels() if use_orth else None,
                                                  m.flat_classifier_weights(), labels, dl, flags, wts, EPS, mult, bias, chain)
    if (not finetune) and (not pretrain) and tanh_desc:                             # pipnet/train.py:1089-1133
        dflags |= ops.DESC_TANH_DESC
        dw[0] = float(args.tanh_desc.split('|')[1]) / N
        losses_used.append('TANH_DESC')
    desc_stats = None
    if dflags:
        use_mp = bool(dflags & ops.DESC_MASK_PRUNE)
        if use_mp and gumbel_noise is None:
            gumbel_noise = ops.gumbel_noise(dl, pooled.flat.device)
        dloss, desc_stats = ops.DescLosses.apply(pooled.flat, m.flat_classifier_weights(),
                                                 m.flat_presence_logits() if use_mp else None,
                                                 gumbel_noise if use_mp else None, labels, dl, dflags, dw, EPS, boost, 0.5)
        loss = loss + dloss

    if node_accuracy is not None:
        acc = node_accuracy.setdefault('__device__', {'n_examples': torch.zeros(N, device=out_device, dtype=torch.int64),
                                                      'n_correct': torch.zeros(N, device=out_device, dtype=torch.int64)})
        acc['n_examples'] += labels.n_desc
        acc['n_correct'] += n_correct

    a_loss_pf = LazyNodeLosses(stats[0] if use_align else None, labels.n_desc, names)
    tanh_loss = LazyNodeLosses(stats[1] if use_tanh else None, labels.n_desc, names)
    orth_loss = LazyNodeLosses(stats[2] if use_orth else None, labels.n_desc, names)
    class_loss = LazyNodeLosses(stats[3] if use_cls else None, labels.n_desc, names)
    placeholder = -5
    avg_class_loss = LazyMean(class_loss) if use_cls else None
    avg_a_loss_pf, avg_tanh_loss, avg_orth = LazyMean(a_loss_pf), LazyMean(tanh_loss), LazyMean(orth_loss)
    if print and train_iter is not None and hasattr(train_iter, 'set_postfix_str'):
        train_iter.lazy_postfix = (loss.detach(), avg_class_loss, avg_a_loss_pf, avg_tanh_loss, avg_orth, '+'.join(losses_used))
    a_loss = torch.tensor(-5)
    uni_loss = torch.tensor(-5)
    use_td = bool(dflags & ops.DESC_TANH_DESC)
    avg_tanh_desc = LazyMean(LazyNodeLosses(desc_stats[0], labels.n_desc, names)) if use_td else -5
    res = (loss, class_loss, a_loss, tanh_loss, {}, {}, orth_loss, uni_loss, avg_class_loss, avg_a_loss_pf, avg_tanh_loss,
           placeholder, (placeholder if pretrain else -5), avg_orth, -5, -5, -5, avg_tanh_desc, -5, -5, 0.)
    out_res = _LossResult(res, stats, labels.n_desc)
    out_res.desc_stats = desc_stats          # [4,N] tanh_desc, contrast, overspecificity, mask_l1 (None when all are off)
    return out_res



# --------------------------------------------------------------------------- graphed head step (behind train_pipnet)
class _HeadReplay(torch.autograd.Function):
    """autograd node of a replayed head step: forward = the captured loss, backward = the captured d loss / d features
    (the head parameters' gradients are delivered to `.grad` directly, they never travel through autograd)"""

    @staticmethod
    def forward(ctx, features, anchor, gh):
        ctx.gh = gh
        return gh.loss.clone()

    @staticmethod
    def backward(ctx, g):
        gh = ctx.gh
        for p, gp in zip(gh.params, gh.param_grads):
            if gp is None:
                continue
            if p.grad is None:
                p.grad = gp                     # static storage, overwritten by the next replay (after optimizer.step)
            else:
                p.grad.add_(gp)
        if not ctx.needs_input_grad[0]:
            return None, None, None
        gx = gh.grad_x
        if not gh.unit_grad:
            gx = gx * g
        return (gx.to(gh.feat_dtype) if gx.dtype != gh.feat_dtype else gx), None, None


class GraphedHeadTrainStep:
    """Head forward + losses + head backward of ONE training step captured into a CUDA graph and exposed as a
    differentiable op on the backbone's feature map: `loss = step(features, ys); loss.backward()` runs the backbone's
    backward from the captured feature gradient and leaves the head parameters' gradients in `.grad`.
    Static shapes only; `train_pipnet` keys its cache on everything that changes the captured work and falls back to the
    eager path otherwise (`pipnet/train.py:229-264` is the loop this replaces: ~30 launches + autograd bookkeeping per
    step, 1.4 ms of host time for 0.3 ms of kernels)."""

    def __init__(self, net, features, ys, loss_kwargs, warmup=3, unit_grad=True):
        m = _unwrap(net)
        self.m, self.unit_grad = m, unit_grad
        backbone_ids = {id(p) for p in m._net.parameters()}
        self.params = [p for p in m.parameters() if id(p) not in backbone_ids and p.requires_grad]
        self.static_x = features.detach().clone().requires_grad_(True)
        self.static_y = ys.detach().clone()
        self.feat_dtype = features.dtype
        self.node_acc = {}
        # keeps the node in the autograd graph when the backbone is frozen (features without grad): the head parameters'
        # gradients are delivered by its backward either way
        self.anchor = torch.zeros((), device=features.device, requires_grad=True)

        def run():
            labels = make_labels(net, self.static_y)
            saved = m._net
            try:
                m._net = torch.nn.Identity()          # the backbone already ran: the head's input IS the feature map
                f, pf, pooled, out = m(self.static_x, labels=labels)
            finally:
                m._net = saved
            res = calculate_loss(loss_kwargs['epoch'], net, {}, f, pf, pooled, out, self.static_y, labels=labels,
                                 node_accuracy=self.node_acc, train_iter=None, print=False,
                                 **{k: v for k, v in loss_kwargs.items() if k != 'epoch'})
            return res, out, labels

        def zero():
            for p in self.params:
                p.grad = None
            self.static_x.grad = None

        saved_grads = [p.grad for p in self.params]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        # this object delivers param.grad itself (_HeadReplay.backward), so the head's gradients can go through the flat
        # bucket: no per-producer zero fills, no autograd add of the orth gradient and dW
        one = self._one = torch.ones((), device=features.device)         # root gradient: static, instead of a ones_like fill per step
        with ops.local_grad_bucket():
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    zero()
                    self.node_acc.clear()
                    run()[0][0].backward(gradient=one)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            zero()
            self.node_acc.clear()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.res, self.out, self.labels = run()
                self.res[0].backward(gradient=one)
        self.loss = self.res[0].detach()
        self.grad_x = self.static_x.grad
        self.param_grads = [p.grad for p in self.params]
        for p, g in zip(self.params, saved_grads):      # capture must not leave gradients behind
            p.grad = g

    def __call__(self, features, ys):
        self.static_x.detach().copy_(features, non_blocking=True)
        self.static_y.copy_(ys, non_blocking=True)
        self.graph.replay()
        return _HeadReplay.apply(features, self.anchor, self)


_GRAPH_CACHE = {}


def _graphed_head(net, features, ys, loss_kwargs, flags_key):
    """cached `GraphedHeadTrainStep` for this (net, shapes, phase, loss switches, trainable set) or None -> eager"""
    m = _unwrap(net)
    if os.environ.get('HC_HEAD_GRAPH', '1') == '0' or not features.is_cuda or net is not m:     # DDP-wrapped: eager
        return None
    head_state = tuple(bool(p.requires_grad) for p in m.parameters())
    key = (id(m), tuple(features.shape), features.dtype, tuple(features.stride()), tuple(ys.shape), flags_key, head_state,
           id(ops.GRAD_ALLREDUCE_GROUP), bool(features.requires_grad))
    gh = _GRAPH_CACHE.get(key)
    if gh is None:
        if len(_GRAPH_CACHE) >= 4:
            _GRAPH_CACHE.clear()
        try:
            gh = GraphedHeadTrainStep(net, features, ys, loss_kwargs)
        except Exception as ex:                      # capture is an optimisation of the launch path, not of the math
            import warnings
            warnings.warn(f'head step not captured ({ex!r}); running it eagerly')
            gh = False
        _GRAPH_CACHE[key] = gh
    return gh or None

# --------------------------------------------------------------------------- epoch drivers
def _class_to_idx(loader):
    ds = loader.dataset
    while not hasattr(ds, 'class_to_idx'):
        if not hasattr(ds, 'dataset'):
            raise Exception('the dataset chain must end in an object with class_to_idx (ImageFolder-like)')
        ds = ds.dataset
    return ds.class_to_idx


def _phase_weights(pretrain, epoch, nr_epochs, args):
    """`pipnet/train.py:148-177`"""
    if pretrain:
        return dict(align_pf_weight=(epoch / nr_epochs) * 1., byol_weight=0.5, align_weight=0.5, unif_weight=3., t_weight=5.,
                    mm_weight=0., cl_weight=0., OOD_loss_weight=0., orth_weight=0.5, cluster_desc_weight=0.8,
                    sep_desc_weight=0.08, subspace_sep_weight=1e-2)
    return dict(align_pf_weight=5., byol_weight=2, align_weight=0.5, unif_weight=3., t_weight=2., mm_weight=2.,
                cl_weight=args.cl_weight, OOD_loss_weight=0.2, orth_weight=0.5, cluster_desc_weight=0.8,
                sep_desc_weight=0.08, subspace_sep_weight=1e-2)


def _to_float(x):
    if x is None:
        return -5.0
    if torch.is_tensor(x):
        return float(x)
    return float(x)


class _Acc:
    """device-side running sums; one host copy at the end of the epoch"""

    def __init__(self, device):
        self.d: Dict[str, torch.Tensor] = {}
        self.device = device

    def add(self, k, v):
        if v is None:
            v = -5.0
        v = v.detach().float() if torch.is_tensor(v) else torch.tensor(float(v), device=self.device)
        self.d[k] = self.d.get(k, torch.zeros((), device=self.device)) + v.to(self.device)

    def result(self, steps):
        if not self.d:
            return {}
        keys = list(self.d)
        vals = (torch.stack([self.d[k] for k in keys]) / float(steps)).cpu().tolist()
        return dict(zip(keys, vals))


def _run_epoch(net, loader, optimizer_net, optimizer_classifier, scheduler_net, scheduler_classifier, criterion, epoch,
               nr_epochs, device, pretrain, finetune, progress_prefix, kw, train: bool):
    from tqdm import tqdm
    m: PIPNet = _unwrap(net)
    args = kw.get('args')
    _reject('OOD loader', kw.get('train_loader_OOD') is not None or kw.get('test_loader_OOD') is not None)
    name2label = _class_to_idx(loader)
    names = m.layout.node_names
    if sorted(name2label, key=name2label.get) != m.layout.leaf_names:
        raise Exception('dataset classes must be the tree leaves in sorted order (label i <-> i-th sorted leaf name)')
    node_accuracy = {}
    w = _phase_weights(pretrain, epoch, nr_epochs, args)
    iters = len(loader)
    it = tqdm(enumerate(loader), total=iters, desc=progress_prefix + '%s' % epoch, mininterval=2., ncols=0)
    acc = _Acc(device)
    n_fine_correct = torch.zeros((), device=device, dtype=torch.int64)
    n_samples = 0
    lrs_net, lrs_class = [], []
    stat_sums = torch.zeros(4, len(names), device=device)
    step_means = torch.zeros(4, device=device)
    node_cnt = torch.zeros(len(names), device=device)
    steps = 0
    ctx = torch.enable_grad() if train else torch.no_grad()
    with ctx:
        for i, batch in it:
            if train:
                xs1, xs2, ys = batch
                xs1, xs2, ys = xs1.to(device, non_blocking=True), xs2.to(device, non_blocking=True), ys.to(device, non_blocking=True)
                xs, ys = torch.cat([xs1, xs2]), torch.cat([ys, ys])                     # pipnet/train.py:213-214
                optimizer_classifier.zero_grad(set_to_none=True)
                optimizer_net.zero_grad(set_to_none=True)
            else:
                xs, ys = batch
                xs, ys = xs.to(device, non_blocking=True), ys.to(device, non_blocking=True)
                xs, ys = torch.cat([xs, xs]), torch.cat([ys, ys])                       # pipnet/train.py:652-653
            loss_kwargs = dict(epoch=epoch, net_normalization_multiplier=m._multiplier, pretrain=pretrain, finetune=finetune,
                               criterion=criterion, EPS=1e-8, root=m.root, label2name=None, OOD_loss_required=False,
                               kernel_orth=kw.get('kernel_orth', False), tanh_desc=kw.get('tanh_desc', False),
                               align=kw.get('align', True), uni=kw.get('uni', True), align_pf=kw.get('align_pf', False),
                               tanh=kw.get('tanh', False), minmaximize=kw.get('minmaximize', False),
                               cluster_desc=kw.get('cluster_desc', False), sep_desc=kw.get('sep_desc', False),
                               subspace_sep=kw.get('subspace_sep', False), byol=kw.get('byol', False), train=train, args=args,
                               device=device, **w)
            gh = None
            if train and net is m and xs.is_cuda and os.environ.get('HC_HEAD_GRAPH', '1') != '0':
                # static shapes: the whole head step (forward, losses, head backward) is ONE CUDA-graph replay between the
                # backbone's forward and backward; anything that changes the captured work is part of the cache key
                feats_bb = m._net(xs)
                flags_key = (bool(pretrain), bool(finetune), epoch if 'y' in getattr(args, 'mask_prune_overspecific', 'n') else 0,
                             tuple(sorted((k, repr(v)) for k, v in loss_kwargs.items()
                                          if k not in ('epoch', 'net_normalization_multiplier', 'root', 'args', 'criterion', 'device'))),
                             repr(sorted(vars(args).items())) if args is not None else '')
                gh = _graphed_head(net, feats_bb, ys, loss_kwargs, flags_key)
                if gh is None:                       # eager head on the features already computed
                    labels = make_labels(net, ys)
                    saved_bb = m._net
                    try:
                        m._net = torch.nn.Identity()
                        features, proto_features, pooled, out = m(feats_bb, labels=labels)
                    finally:
                        m._net = saved_bb
            if gh is not None:
                loss = gh(feats_bb, ys)
                res, out, labels = gh.res, gh.out, gh.labels
                dacc = gh.node_acc.get('__device__')
                if dacc is not None:
                    acc_ep = node_accuracy.setdefault('__device__', {'n_examples': torch.zeros_like(dacc['n_examples']),
                                                                     'n_correct': torch.zeros_like(dacc['n_correct'])})
                    acc_ep['n_examples'] += dacc['n_examples']
                    acc_ep['n_correct'] += dacc['n_correct']
                it.lazy_postfix = (gh.loss, res[8], res[9], res[10], res[13], 'graph')
            else:
                if not (train and net is m and xs.is_cuda and os.environ.get('HC_HEAD_GRAPH', '1') != '0'):
                    labels = make_labels(net, ys)
                    if train:
                        features, proto_features, pooled, out = net(xs, labels=labels)
                    else:
                        features, proto_features, pooled, out = net(xs, apply_overspecificity_mask=kw.get('apply_overspecificity_mask', False),
                                                                    labels=labels)
                res = calculate_loss(additional_network_outputs={}, features=features, proto_features=proto_features,
                                     pooled=pooled, out=out, ys=ys, net=net, train_iter=it, print=True,
                                     node_accuracy=node_accuracy, labels=labels, **loss_kwargs)
                loss = res[0]
            if train:
                loss.backward()
                if not pretrain:
                    optimizer_classifier.step()
                    scheduler_classifier.step(epoch - 1 + (i / iters))
                    lrs_class.append(scheduler_classifier.get_last_lr()[0])
                if not finetune:
                    optimizer_net.step()
                    scheduler_net.step()
                    lrs_net.append(scheduler_net.get_last_lr()[0])
                else:
                    lrs_net.append(0.)
            with torch.no_grad():
                acc.add('loss', loss)
                act = (labels.n_desc > 0).float()
                node_cnt += act
                stat_sums += res.stats                       # [4,N]: align, tanh, orth, class (0 where absent)
                step_means += res.stats.sum(dim=1) / act.sum().clamp_min(1.0)
                if train:                                     # pipnet/train.py:363 (plain) vs :713 (test-time switches)
                    _, joint = m.get_joint_distribution(out)
                else:
                    _, joint = m.get_joint_distribution(out, leave_out_classes=kw.get('leave_out_classes'),
                                                        apply_overspecificity_mask=kw.get('apply_overspecificity_mask', False),
                                                        softmax_tau=kw.get('path_prob_softmax_tau', 1))
                n_fine_correct += (joint.argmax(dim=1) == ys).sum()                      # pipnet/train.py:363-369
                n_samples += ys.numel()
            steps += 1
            if hasattr(it, 'lazy_postfix') and (i % 50 == 0):
                lz = it.lazy_postfix
                it.set_postfix_str(f'L:{float(lz[0]):.3f}, losses_used:{lz[5]}', refresh=False)

    steps = max(steps, 1)
    means = acc.result(steps)
    sm = (step_means / steps).cpu().tolist()
    means.update({'a_loss_pf': sm[0], 'tanh_loss': sm[1], 'kernel_orth_loss': sm[2], 'class_loss': sm[3]})
    node_sums = {'a_loss_pf': stat_sums[0], 'tanh_loss': stat_sums[1], 'kernel_orth_loss': stat_sums[2], 'class_loss': stat_sums[3]}
    info = dict()
    info['fine_accuracy'] = float(n_fine_correct) / max(n_samples, 1)
    info['train_accuracy' if train else 'test_accuracy'] = 0.
    info['loss'] = means.get('loss', 0.)
    info['class_loss (mean over epoch nodes)'] = means.get('class_loss', -5.)
    info['kernel_orth_loss (mean over epoch nodes)'] = means.get('kernel_orth_loss', -5.)
    info['tanh_loss (mean over epoch nodes)'] = means.get('tanh_loss', -5.)
    info['minmaximize_loss (mean over epoch nodes)'] = -5
    info['OOD_loss (mean over epoch nodes)'] = -5
    info['a_loss (mean over epoch)'] = -5
    info['a_loss_pf (mean over epoch nodes)'] = means.get('a_loss_pf', -5.)
    info['uni_loss (mean over epoch)'] = -5
    info['true_uni_loss (without averaging)'] = -5
    info['lrs_net'], info['lrs_class'] = lrs_net, lrs_class

    # node-level accuracy (pipnet/train.py:1187-1194, :484-489) from the device counters
    dev_acc = node_accuracy.pop('__device__', None)
    node_stats = {}
    if dev_acc is not None:
        ne, nc = dev_acc['n_examples'].cpu().tolist(), dev_acc['n_correct'].cpu().tolist()
        for n, e, c in zip(names, ne, nc):
            node_stats[n] = {'n_examples': e, 'n_correct': c, 'accuracy': round(100.0 * c / e, 2) if e else None, 'f1': None}
    info['node_accuracy'] = node_stats

    sub = ('pretrain' if pretrain else 'train') if train else 'test'
    log_dict = {}
    if kw.get('wandb_logging', True):
        log_dict[sub + '/epoch loss'] = info['loss']
        log_dict[sub + '/fine_accuracy'] = info['fine_accuracy']
        log_dict[sub + '/class_loss'] = info['class_loss (mean over epoch nodes)']
        log_dict[sub + '/tanh_loss'] = info['tanh_loss (mean over epoch nodes)']
        log_dict[sub + '/kernel_orth_loss'] = info['kernel_orth_loss (mean over epoch nodes)']
        log_dict[sub + '/a_loss_pf'] = info['a_loss_pf (mean over epoch nodes)']
        for n, st in node_stats.items():
            if st['accuracy'] is not None:
                log_dict[sub + f'/node_wise/acc:{n}'] = st['accuracy']
        cnt = node_cnt.clamp_min(1.0)
        per_node = {k: (v / cnt).cpu().tolist() for k, v in node_sums.items()}
        for k, vals in per_node.items():
            for n, v in zip(names, vals):
                log_dict[sub + f'/node_wise_{k}/{n}'] = v
        run = kw.get('wandb_run')
        if run is not None:
            run.log(log_dict, step=epoch if pretrain else (epoch + kw.get('pretrain_epochs', 0)))
    log = kw.get('log')
    if log is not None:
        table = f'epoch_wise_metrics_{"train" if train else "test"}'
        cols = ['fine_accuracy', 'loss', 'class_loss (mean over epoch nodes)', 'kernel_orth_loss (mean over epoch nodes)',
                'tanh_loss (mean over epoch nodes)', 'minmaximize_loss (mean over epoch nodes)',
                'OOD_loss (mean over epoch nodes)', 'a_loss (mean over epoch)', 'uni_loss (mean over epoch)',
                'true_uni_loss (without averaging)']
        try:
            log.create_log(table, 'epoch', *cols)
        except Exception:
            pass
        log.log_values(table, epoch if pretrain else (epoch + kw.get('pretrain_epochs', 0)), *[info[c] for c in cols])
    print_fn = __builtins__['print'] if isinstance(__builtins__, dict) else getattr(__builtins__, 'print')
    print_fn('\tFine accuracy:', round(info['fine_accuracy'], 2))
    return info, log_dict


def train_pipnet(net, train_loader, optimizer_net, optimizer_classifier, scheduler_net, scheduler_classifier, criterion,
                 epoch, nr_epochs, device, pretrain=False, finetune=False, progress_prefix: str = 'Train Epoch',
                 wandb_logging=True, train_loader_OOD=None, kernel_orth=False, tanh_desc=False, align=True, uni=True,
                 align_pf=False, tanh=False, minmaximize=False, cluster_desc=False, sep_desc=False, subspace_sep=False,
                 byol=False, byol_tau_base=0.9995, byol_tau_max=1., step_info=None, wandb_run=None, pretrain_epochs=0,
                 log=None, args=None, dist_training=False):
    """One training epoch (`pipnet/train.py:73-522`).  Returns (train_info, log_dict)."""
    net.train()
    m = _unwrap(net)
    # `pipnet/train.py:101-110` sets `.requires_grad` on the classification *modules* (a no-op attribute);
    # the real freezing is done by the driver on the parameters (`main_dist.py:472-485`), which we respect.
    if pretrain:
        progress_prefix = 'Pretrain Epoch'
    kw = dict(wandb_logging=wandb_logging, train_loader_OOD=train_loader_OOD, kernel_orth=kernel_orth, tanh_desc=tanh_desc,
              align=align, uni=uni, align_pf=align_pf, tanh=tanh, minmaximize=minmaximize, cluster_desc=cluster_desc,
              sep_desc=sep_desc, subspace_sep=subspace_sep, byol=byol, wandb_run=wandb_run, pretrain_epochs=pretrain_epochs,
              log=log, args=args)
    return _run_epoch(net, train_loader, optimizer_net, optimizer_classifier, scheduler_net, scheduler_classifier, criterion,
                      epoch, nr_epochs, device, pretrain, finetune, progress_prefix, kw, train=True)


def test_pipnet(net, test_loader, optimizer_net, optimizer_classifier, scheduler_net, scheduler_classifier, criterion,
                epoch, nr_epochs, device, pretrain=False, finetune=False, progress_prefix: str = 'Test Epoch',
                wandb_logging=True, test_loader_OOD=None, kernel_orth=False, tanh_desc=False, align=True, uni=True,
                align_pf=False, tanh=False, minmaximize=False, cluster_desc=False, sep_desc=False, subspace_sep=False,
                byol=False, byol_tau_base=0.9995, step_info=None, wandb_run=None, pretrain_epochs=0, log=None, args=None,
                apply_overspecificity_mask=False, leave_out_classes=None, path_prob_softmax_tau=1):
    """One evaluation epoch (`pipnet/train.py:525-849`): the test batch is duplicated like the reference does
    (`:652-653`) so the same paired kernels and losses apply.  Returns (test_info, log_dict)."""
    net.eval()
    kw = dict(wandb_logging=wandb_logging, test_loader_OOD=test_loader_OOD, kernel_orth=kernel_orth, tanh_desc=tanh_desc,
              align=align, uni=uni, align_pf=align_pf, tanh=tanh, minmaximize=minmaximize, cluster_desc=cluster_desc,
              sep_desc=sep_desc, subspace_sep=subspace_sep, byol=byol, wandb_run=wandb_run, pretrain_epochs=pretrain_epochs,
              log=log, args=args, apply_overspecificity_mask=apply_overspecificity_mask, leave_out_classes=leave_out_classes,
              path_prob_softmax_tau=path_prob_softmax_tau)
    return _run_epoch(net, test_loader, optimizer_net, optimizer_classifier, scheduler_net, scheduler_classifier, criterion,
                      epoch, nr_epochs, device, pretrain, finetune, progress_prefix, kw, train=False)


test_pipnet.__test__ = False   # not a pytest test
