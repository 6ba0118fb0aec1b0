"""Visualisation feed: which image patches activate each prototype most.

The reference collects them per node with `save_images_topk` (`util/vis_hpipnet.py:184-290`): for every node it walks the
projection set with batch size 1, re-runs the whole network, and keeps one Python heap per (prototype, leaf) of
`(pooled score, ..., image path, patch box, latent map)`.  Here the fused forward already streams `pooled [V,P]` and the
first-occurrence `argmax [V,P]` for ALL nodes, so one pass with any batch size feeds device-side top-k tables
(`hcomp_topk_update`); the full softmax map of a winner is rebuilt on demand with `ops.materialize_map`.
Rendering (PIL fonts, jpg grids) stays with the caller."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from ._cabi import call, ptr


def get_patch_size(args) -> Tuple[int, int]:
    """(patchsize, skip) as `util/func.py:3-6`: 32-pixel patches, stride from the image size and the latent width."""
    patchsize = 32
    skip = round((args.image_size - patchsize) / (args.wshape - 1))
    return patchsize, skip


def get_img_coordinates(img_size, softmaxes_shape, patchsize, skip, h_idx, w_idx):
    """Pixel box (h_min, h_max, w_min, w_max) of latent location (h_idx, w_idx); same arithmetic as
    `util/vis_pipnet.py:373-409` (26x26 ConvNeXt maps: half-size receptive fields at the border)."""
    w_idx = int(w_idx)
    h_idx = int(h_idx)
    if softmaxes_shape[1] == 26 and softmaxes_shape[2] == 26:
        h_min = max(0, (h_idx - 1) * skip + 4)
        if h_idx >= softmaxes_shape[-1] - 1:
            h_min -= 4
        h_max = h_min + patchsize
        w_min = max(0, (w_idx - 1) * skip + 4)
        if w_idx >= softmaxes_shape[-1] - 1:
            w_min -= 4
        w_max = w_min + patchsize
    else:
        h_min = h_idx * skip
        h_max = min(img_size, h_idx * skip + patchsize)
        w_min = w_idx * skip
        w_max = min(img_size, w_idx * skip + patchsize)
    if h_idx == softmaxes_shape[1] - 1:
        h_max = img_size
    if w_idx == softmaxes_shape[2] - 1:
        w_max = img_size
    if h_max == img_size:
        h_min = img_size - patchsize
    if w_max == img_size:
        w_min = img_size - patchsize
    return h_min, h_max, w_min, w_max


class TopKPatches:
    """Device-side top-k tables [P, L, k] (score, image id, flat argmax location) over a pass through the data."""

    def __init__(self, net, topk: int = 10, find_non_descendants: bool = False, device=None):
        self.net = net.module if hasattr(net, 'module') else net
        self.layout = self.net.layout
        self.k = int(topk)
        self.find_non_descendants = bool(find_non_descendants)
        dev = torch.device(device) if device is not None else next(self.net.parameters()).device
        self.dl = self.net.device_layout(dev)
        P, L = self.layout.P, self.layout.L
        self.score = torch.zeros(P, L, self.k, device=dev, dtype=torch.float32)
        self.img = torch.full((P, L, self.k), -1, device=dev, dtype=torch.int64)
        self.loc = torch.zeros(P, L, self.k, device=dev, dtype=torch.int32)

    @torch.no_grad()
    def update(self, pooled_flat: torch.Tensor, argmax_flat: torch.Tensor, ys: torch.Tensor, img_ids: torch.Tensor):
        """pooled [V,P] fp32, argmax [V,P] int32 (both from `PIPNet.forward`), ys [V] leaf indices, img_ids [V] int64."""
        dev = self.score.device
        pooled_flat = pooled_flat.detach().float().contiguous()
        argmax_flat = argmax_flat.detach().to(torch.int32).contiguous()
        ys = ys.to(device=dev, dtype=torch.int64).contiguous()
        img_ids = img_ids.to(device=dev, dtype=torch.int64).contiguous()
        V = pooled_flat.shape[0]
        ws = torch.empty(2 * V, device=dev, dtype=torch.int32)
        wc = self.net.flat_classifier_weights().detach().contiguous()
        call('hcomp_topk_update', ptr(pooled_flat), ptr(argmax_flat), ptr(ys), ptr(img_ids), ptr(wc), self.dl.tref, V, self.k,
             int(self.find_non_descendants), ptr(ws), ptr(self.score), ptr(self.img), ptr(self.loc),
             C.c_void_p(torch.cuda.current_stream().cuda_stream))

    def collect(self, H: int, W: int) -> Dict[str, Dict[int, Dict[str, List[Tuple[float, int, Tuple[int, int]]]]]]:
        """{node name: {prototype index within the node: {leaf name: [(score, image id, (h, w)), ...] best first}}}"""
        score, img, loc = self.score.cpu(), self.img.cpu(), self.loc.cpu()
        L = self.layout
        out: Dict[str, Dict[int, Dict[str, list]]] = {}
        filled = (img[:, :, 0] >= 0).nonzero().tolist()
        for p, leaf in filled:
            ni = int(L.proto_node[p])
            node = out.setdefault(L.node_names[ni], {})
            per_leaf = node.setdefault(p - int(L.proto_off[ni]), {})
            items = []
            for i in range(self.k):
                if int(img[p, leaf, i]) < 0:
                    break
                l = int(loc[p, leaf, i])
                items.append((float(score[p, leaf, i]), int(img[p, leaf, i]), (l // W, l % W)))
            per_leaf[L.leaf_names[leaf]] = items
        return out


@torch.no_grad()
def collect_topk(net, loader, topk: int = 10, find_non_descendants: bool = False, device='cuda') -> TopKPatches:
    """One pass over `loader` (batches of (xs, ys) with leaf-index labels; any batch size) -> filled `TopKPatches`.
    Image ids are the running sample index, i.e. positions in the loader's (unshuffled) dataset order."""
    m = net.module if hasattr(net, 'module') else net
    was_training = m.training
    m.eval()
    tk = TopKPatches(m, topk, find_non_descendants, device)
    seen = 0
    for batch in loader:
        xs, ys = batch[0], batch[-1]
        xs = xs.to(device, non_blocking=True)
        _, pf, pooled, _ = m(xs, inference=False)         # the reference calls its forward with inference=False here
        ids = torch.arange(seen, seen + xs.shape[0], dtype=torch.int64)
        tk.update(pooled.flat, pf.argmax.flat, ys, ids)
        seen += xs.shape[0]
    m.train(was_training)
    return tk
