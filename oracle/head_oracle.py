"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-PyTorch (CPU, fp32 or fp64) restatement of the HComP-Net per-node prototype head and
its fused losses, written against the reference's algorithm line by line.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
it, and only as the checker / reported baseline.  The product (`pipnet_b200/`) never does:
it fails loudly when its CUDA extension is missing.

Pinning: the reference has NO tests or golden vectors for this path, so
the oracle is pinned against the reference itself, imported unmodified in the build container
(`oracle/ref_harness.py`): `tests/test_oracle_vs_reference.py` compares every output below with
`PIPNet.forward` + `calculate_loss` + autograd on identical inputs, and
`oracle/make_golden.py` freezes reference outputs into `tests/golden/*.npz` for the GPU box.

Each function cites the reference lines it restates (paths relative to /root/reference).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- forward
def head_forward(x, add_on_w: Dict[str, torch.Tensor], cls_w: Dict[str, torch.Tensor], root, *,
                 softmax_tau=1.0, inference=False, cls_b: Optional[Dict[str, torch.Tensor]] = None):
    """`PIPNet.forward` after the backbone (`pipnet/pipnet.py:124-170`), canonical recipe
    (`--softmax "y|tau"`, plain Conv2d add-on without bias, NonNegLinear classifier).

    x        [V, C, H, W]  backbone features
    add_on_w name -> [P_n, C]   1x1 conv kernels (`pipnet/pipnet.py:125`, built `:1207-1208`)
    cls_w    name -> [C_n, P_n] NonNegLinear weights (`pipnet/pipnet.py:1035-1036`)
    returns proto_features (softmaxed maps), pooled, argmax (flat h*W+w, first max), out
    """
    V, C, H, W = x.shape
    proto, pooled, argmax, out = {}, {}, {}, {}
    for node in root.nodes_with_children():
        w = add_on_w[node.name]
        z = torch.einsum('vchw,pc->vphw', x, w)                    # 1x1 conv           pipnet.py:125
        z = z / softmax_tau                                        #                    pipnet.py:146
        s = torch.softmax(z, dim=1)                                # over the node's prototypes  :147
        proto[node.name] = s
        flat = s.flatten(2)                                        # [V, P_n, H*W]
        pv, pi = flat.max(dim=2)                                   # AdaptiveMaxPool2d(1) + Flatten :159
        # `Tensor.max(dim)` does not promise which index it returns on ties; the reference's
        # index convention is F.max_pool2d(return_indices=True) (pipnet.py:24-25) == first in
        # row-major order, which is what argmax documents.
        pi = flat.argmax(dim=2)
        if inference:
            pv = torch.where(pv < 0.1, torch.zeros_like(pv), pv)  #                    pipnet.py:168-169
        pooled[node.name] = pv
        argmax[node.name] = pi
        b = None if cls_b is None else cls_b.get(node.name)
        out[node.name] = F.linear(pv, torch.relu(cls_w[node.name]), b)   # NonNegLinear  pipnet.py:1036
    return proto, pooled, argmax, out


# --------------------------------------------------------------------------- per-node masks
def node_targets(root, ys, label2name):
    """For every node: boolean mask of samples whose leaf lies below the node and, for those,
    the label of the child on the path (`pipnet/train.py:934-937`)."""
    names = [label2name[int(y)] for y in ys]
    masks, targets = {}, {}
    for node in root.nodes_with_children():
        m = torch.tensor([n in node.leaf_descendents for n in names], dtype=torch.bool)
        t = [node.children_to_labels[node.closest_descendent_for(n).name] for n in names if n in node.leaf_descendents]
        masks[node.name] = m
        targets[node.name] = torch.tensor(t, dtype=torch.long)
    return masks, targets


# --------------------------------------------------------------------------- loss terms
def align_loss(inputs, targets, eps=1e-12):
    """`pipnet/train.py:1399-1405`: -mean log(<x, y> + eps) over rows."""
    return -torch.log((inputs * targets).sum(dim=1) + eps).mean()


def align_pf_term(s_node, mask):
    """`pipnet/train.py:1063-1069`: the masked batch is chunked into the two views, each
    map flattened to [n*H*W, P_n]; symmetric loss with the other side detached."""
    pf1, pf2 = s_node[mask].chunk(2)
    e1 = pf1.flatten(2).permute(0, 2, 1).flatten(end_dim=1)
    e2 = pf2.flatten(2).permute(0, 2, 1).flatten(end_dim=1)
    return (align_loss(e1, e2.detach()) + align_loss(e2, e1.detach())) / 2.


def tanh_term(pooled_node, mask, eps=1e-8):
    """`pipnet/train.py:1080-1082` (EPS passed by train_pipnet is 1e-8, `pipnet/train.py:238`)."""
    p1, p2 = pooled_node[mask].chunk(2)
    return -(torch.log(torch.tanh(p1.sum(dim=0)) + eps).mean() + torch.log(torch.tanh(p2.sum(dim=0)) + eps).mean()) / 2.


def orth_term(w_node, cls_w_node):
    """`pipnet/train.py:1137-1142` + `orth_dist :1408-1412`: prototypes with any classifier
    weight > 1e-3; Frobenius norm of (gram - I) on the smaller side."""
    rel = w_node[(cls_w_node > 0.001).any(dim=0)]
    mat = rel.reshape(rel.shape[0], -1)
    if mat.shape[0] < mat.shape[1]:
        mat = mat.permute(1, 0)
    return torch.norm(mat.t() @ mat - torch.eye(mat.shape[1], dtype=mat.dtype))


def class_term(out_node, mask, target, weights, multiplier=2.0):
    """`pipnet/train.py:1158-1163` with `WeightedNLLLoss` (`util/custom_losses.py:22-34`):
    x = log1p(out**m); per-sample NLL(log_softmax(x)) times the weight of the target class;
    mean over the node's descendants in the batch."""
    x = torch.log1p(out_node[mask] ** multiplier)
    lp = F.log_softmax(x, dim=1)
    nll = -lp.gather(1, target[:, None]).squeeze(1)
    w = weights.to(lp.dtype)[target]
    return (nll * w).mean()


def gumbel_noise_like(logits):
    """The noise `F.gumbel_softmax` draws internally (torch/nn/functional.py: -empty_like(logits).exponential_().log());
    drawing it here at the same points of the loop, under the same seed, reproduces the reference's RNG stream."""
    return -torch.empty_like(logits, memory_format=torch.legacy_contiguous_format).exponential_().log()


def mask_prune_terms(node, pooled_node, cls_w_node, presence_node, batch_names, *, boost=None, geometric=False,
                     sg_before_masking=False, noise=None, noise_out=None, tau=0.5):
    """`pipnet/train.py:946-1015` for one node: overspecificity and mask-L1 terms (before their 2.0 / 0.5 weights).
    Per child (in `node.children` order): relevant prototypes = classifier row > 1e-3 (:963); for every leaf below the
    child that occurs in the batch, the max over its rows of `pooled` (:968-973); children without any such leaf are
    skipped AFTER their relevant prototypes were counted (:965, :975-976); the presence logits go through a soft Gumbel
    softmax (tau 0.5) whose OUTPUT replaces the variable, so each further child perturbs the previous child's
    probabilities (:978); score = prod over present leaves of clamp(max * boost, max=1) (:980-985) or the plain /
    geometric-mean product (:987-995).  Both sums are divided by the total relevant count (:1003-1004).
    noise: {child label -> [P_n, 2]} to inject; None draws it like the reference does.  Returns (ovsp, l1)."""
    pres = presence_node
    ovsp, l1, total_rel = 0., 0., 0.
    for child in node.children:
        c = node.children_to_labels[child.name]
        rel = torch.nonzero(cls_w_node[c, :] > 1e-3).squeeze(-1)
        total_rel += rel.shape[0]
        rows = []
        for leaf in child.leaf_descendents:
            idx = torch.tensor([n == leaf for n in batch_names])
            if int(idx.sum()) == 0:
                continue
            rows.append(pooled_node[idx][:, rel].max(dim=0, keepdim=True)[0])
        if not rows:
            continue
        mx = torch.cat(rows, dim=0)
        g = gumbel_noise_like(pres) if noise is None else noise[c].to(pres.dtype)
        if noise_out is not None:
            noise_out[c] = g.detach().clone()
        pres = torch.softmax((pres + g) / tau, dim=-1)
        if boost is not None:
            score = torch.prod(torch.clamp(mx * boost, max=1.0), dim=0)
        elif geometric:
            score = torch.prod(mx.pow(1 / mx.shape[0]), dim=0)
        else:
            score = torch.prod(mx, dim=0)
        if sg_before_masking:
            score = score.detach()
        ovsp = ovsp + (-1) * (score * pres[rel, 1]).sum()
        l1 = l1 + pres[rel, 1].sum()
    return ovsp / total_rel, l1 / total_rel


def contrasting_set_term(node, pooled_node, cls_w_node, mask, target):
    """`pipnet/train.py:1017-1057` with TOPK = 1: for every child, over the node's descendants in the batch that do NOT
    belong to that child, the max activation of the child's prototypes (classifier row > 1e-5); mean over all
    (child, prototype) entries.  None when nothing qualifies (:1054)."""
    vals = []
    sub = pooled_node[mask]
    for child in node.children:
        c = node.children_to_labels[child.name]
        rel = torch.nonzero(cls_w_node[c, :] > 1e-5).squeeze(-1)
        if len(rel) == 0:
            continue
        rows = torch.nonzero(target != c).squeeze(-1)
        if len(rows) == 0:
            continue
        vals.append(sub[rows, :][:, rel].max(dim=0)[0])
    if not vals:
        return None
    return torch.cat(vals).mean()


def tanh_desc_term(node, pooled_node, cls_w_node, batch_names, eps=1e-8):
    """`pipnet/train.py:1089-1133`: the tanh loss of every leaf below the node, restricted to the prototypes of the
    child the leaf hangs under (classifier row > 1e-3), on that leaf's rows split by `.chunk(2)` (:1107, :1119; an
    absent leaf gives two empty halves -> log(tanh(0) + eps)); mean over leaves.  A child without relevant
    prototypes is skipped (the reference asserts / raises there, :1099-1106, :1114-1117)."""
    terms = []
    for child in node.children:
        c = node.children_to_labels[child.name]
        rel = torch.nonzero(cls_w_node[c, :] > 1e-3).squeeze(-1)
        if len(rel) == 0:
            continue
        leaves = [child.name] if child.is_leaf() else list(child.leaf_descendents)
        for leaf in leaves:
            idx = torch.tensor([n == leaf for n in batch_names])
            p1, p2 = pooled_node[idx][:, rel].chunk(2)
            terms.append(-(torch.log(torch.tanh(p1.sum(dim=0)) + eps).mean() + torch.log(torch.tanh(p2.sum(dim=0)) + eps).mean()) / 2.)
    return torch.stack(terms).mean(dim=0)


def head_losses(root, proto, pooled, out, ys, label2name, add_on_w, cls_w, *, pretrain, finetune,
                epoch=1, nr_epochs=10, cl_weight=2.0, kernel_orth=True, tanh_during_second_phase=True,
                multiplier=2.0, tanh_desc_weight=None, contrasting=None, mask_prune=None, presence=None,
                gumbel=None, gumbel_out=None):
    """The head's share of `calculate_loss` (`pipnet/train.py:852-1341`) for the canonical recipe
    (align_pf + tanh + kernel_orth + class loss), with `train_pipnet`'s weights
    (`pipnet/train.py:148-177`) and the `/len(nodes)` normaliser.  Nodes with no descendant in
    the batch are skipped (`:941-942`) but still counted in the normaliser.
    Optional terms of the shipped scripts: `tanh_desc_weight` (--tanh_desc "y|w"), `contrasting` = weight of
    --minimize_contrasting_set (TOPK 1; default 0.1), `mask_prune` = dict(start_epoch, boost, geometric, sg) for
    --mask_prune_overspecific with `presence` {node -> [P_n, 2]} logits; `gumbel` {node -> {child label -> noise}}
    injects the Gumbel noise (None draws it exactly where the reference does), `gumbel_out` collects what was used."""
    nodes = root.nodes_with_children()
    n_nodes = len(nodes)
    if pretrain:
        align_pf_weight, t_weight, cw = (epoch / nr_epochs) * 1., 5., 0.
    else:
        align_pf_weight, t_weight, cw = 5., 2., cl_weight
    orth_weight = 0.5
    masks, targets = node_targets(root, ys, label2name)
    batch_names = [label2name[int(y)] for y in ys]
    res = dict(align={}, tanh={}, orth={}, cls={}, n_desc={}, acc={}, ovsp={}, mask_l1={}, contrast={}, tanh_desc={})
    # `calculate_loss` receives EPS=1e-8 from train_pipnet (:238), but the contrasting-set block re-binds the SAME local
    # to 1e-12 (:1025) before the tanh / tanh_desc terms of the same loop iteration read it (:1080, :1108); the root
    # always has descendants, so with that term on every node sees 1e-12.
    eps = 1e-12 if ((not pretrain) and (not finetune) and contrasting is not None) else 1e-8
    loss = 0.
    for node in nodes:
        m, t = masks[node.name], targets[node.name]
        if t.numel() == 0:
            continue
        res['n_desc'][node.name] = int(t.numel())
        if (not pretrain) and mask_prune is not None and epoch >= mask_prune.get('start_epoch', 0):
            nz_out = None if gumbel_out is None else gumbel_out.setdefault(node.name, {})
            ov, l1 = mask_prune_terms(node, pooled[node.name], cls_w[node.name], presence[node.name], batch_names,
                                      boost=mask_prune.get('boost'), geometric=mask_prune.get('geometric', False),
                                      sg_before_masking=mask_prune.get('sg', False),
                                      noise=None if gumbel is None else gumbel[node.name], noise_out=nz_out)
            res['ovsp'][node.name] = 2.0 * ov / n_nodes             # the reference stores the weighted values (:1006-1010)
            res['mask_l1'][node.name] = 0.5 * l1 / n_nodes
            loss = loss + res['ovsp'][node.name] + res['mask_l1'][node.name]
        if (not pretrain) and (not finetune) and contrasting is not None:
            cs = contrasting_set_term(node, pooled[node.name], cls_w[node.name], m, t)
            if cs is not None:
                res['contrast'][node.name] = cs
                loss = loss + contrasting * cs / n_nodes
        if not finetune:
            a = align_pf_term(proto[node.name], m)
            res['align'][node.name] = a
            loss = loss + align_pf_weight * a / n_nodes
            if pretrain or tanh_during_second_phase:
                th = tanh_term(pooled[node.name], m, eps)
                res['tanh'][node.name] = th
                loss = loss + t_weight * th / n_nodes
        if (not finetune) and (not pretrain) and tanh_desc_weight is not None:
            td = tanh_desc_term(node, pooled[node.name], cls_w[node.name], batch_names, eps)
            res['tanh_desc'][node.name] = td
            loss = loss + tanh_desc_weight * td / n_nodes
        if (not pretrain) and (not finetune) and kernel_orth:
            o = orth_term(add_on_w[node.name], cls_w[node.name])
            res['orth'][node.name] = o
            loss = loss + orth_weight * o / n_nodes
        if not pretrain:
            c = class_term(out[node.name], m, t, node.weights, multiplier)
            res['cls'][node.name] = c
            loss = loss + cw * c / n_nodes
        pred = out[node.name][m].argmax(dim=1)                     # pipnet/train.py:1189-1190
        res['acc'][node.name] = (int(t.numel()), int((pred == t).sum()))
    res['loss'] = loss
    return res


# --------------------------------------------------------------------------- joint leaf distribution
def joint_distribution(root, out, softmax_tau=1.0, *, leave_out_classes=None, mask=None, cls_w=None, presence=None,
                       mask_out=None):
    """`PIPNet.get_joint_distribution` (`pipnet/pipnet.py:173-185`) -> `Node.distribution_over_furthest_descendents`
    (`util/node.py:300-385`): product along each root->leaf path of softmax(log1p(out[node]**2)/tau)[:, child];
    columns re-ordered by sorted leaf name.  Two test-time switches change the child probabilities of a node:
      leave_out_classes  a node with a child whose leaves are ALL left out puts probability 1 on its left-out LEAF child
                         for every sample (`:319-323`; the reference indexes `[0]` of that list and fails without one);
      overspecificity mask (`apply_overspecificity_mask`): a hard Gumbel sample of the node's presence logits masks the
                         classifier weights; if some class row is then entirely <= 1e-3 the node falls back to the
                         leaf-count fractions num_leaves(child) / num_leaves(node) (`:335-359`).
    mask: {node name -> [P_n] 0/1} to inject; otherwise, with `presence` given, drawn like the reference does (same
    recursion order, so the same seed gives the same draws); `mask_out` collects what was used."""
    import torch.nn.functional as Fn
    V = out[root.name].shape[0]
    dt = out[root.name].dtype
    lo = set(leave_out_classes or [])
    use_mask = mask is not None or presence is not None

    def rec(node):
        if lo and any(set(c.leaf_descendents).issubset(lo) for c in node.children):
            left = [c for c in node.children if c.is_leaf() and c.name in lo][0].name
            names = node.unwrap_names_of_joint(node.names_of_joint_distribution())
            row = torch.tensor([1.0 if n == left else 0.0 for n in names], dtype=dt)
            return row.reshape(1, -1).repeat(V, 1), list(names)
        if node.is_leaf():
            return torch.ones(V, 1, dtype=dt), [node.name]
        p = torch.softmax(torch.log1p(out[node.name] ** 2) / softmax_tau, dim=1)
        if use_mask:
            if mask is not None:
                m = mask[node.name].to(dt)
            else:
                m = Fn.gumbel_softmax(presence[node.name], tau=0.5, hard=True, dim=-1)[:, 1].to(dt)
            if mask_out is not None:
                mask_out[node.name] = m.detach().clone()
            mw = m.unsqueeze(0) * cls_w[node.name].to(dt)
            if any(bool((mw[c, :] <= 1e-3).all()) for c in range(mw.shape[0])):
                frac = [c.num_leaf_descendents() / node.num_leaf_descendents() for c in node.children]
                # the reference builds these with torch.tensor([float(...)]) = float32 (util/node.py:359)
                p = torch.tensor(frac, dtype=torch.float32).to(dt).reshape(1, -1).repeat(V, 1)
        cols, names = [], []
        for i, c in enumerate(node.children):
            sub, nm = rec(c)
            cols.append(p[:, i:i + 1] * sub)
            names += nm
        return torch.cat(cols, 1), names

    dist, _ = rec(root)
    # The reference orders (and SELECTS) columns with np.argsort over names_of_joint_distribution()
    # (pipnet/pipnet.py:179-181).  That name list stops at single-child nodes (util/node.py:397-403), so on a tree
    # with a single-child node (e.g. the CUB-08 root) it is shorter than the leaf list and the result keeps only
    # the first len(names) depth-first columns.  Restated as is.
    names = root.unwrap_names_of_joint(root.names_of_joint_distribution())
    order = sorted(range(len(names)), key=lambda i: names[i])
    return dist[:, order]


# --------------------------------------------------------------------------- convenience: one full step
def full_step(x, add_on_w, cls_w, root, ys, label2name, *, pretrain, finetune, softmax_tau=1.0, presence=None, **kw):
    """Forward + losses + autograd backward; returns outputs and gradients w.r.t. x, add-on and
    classifier weights (what `loss.backward()` at `pipnet/train.py:264` produces for the head)."""
    x = x.detach().clone().requires_grad_(True)
    aw = {k: v.detach().clone().requires_grad_(True) for k, v in add_on_w.items()}
    cw = {k: v.detach().clone().requires_grad_(True) for k, v in cls_w.items()}
    proto, pooled, argmax, out = head_forward(x, aw, cw, root, softmax_tau=softmax_tau)
    pr = None if presence is None else {k: v.detach().clone().requires_grad_(True) for k, v in presence.items()}
    res = head_losses(root, proto, pooled, out, ys, label2name, aw, cw, pretrain=pretrain, finetune=finetune, presence=pr, **kw)
    loss = res['loss']
    if torch.is_tensor(loss) and loss.requires_grad:
        loss.backward()
    res.update(proto=proto, pooled=pooled, argmax=argmax, out=out, grad_x=x.grad,
               grad_w={k: v.grad for k, v in aw.items()}, grad_cls={k: v.grad for k, v in cw.items()},
               grad_presence=None if pr is None else {k: v.grad for k, v in pr.items()})
    return res
