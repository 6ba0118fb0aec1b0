"""Flat HBM layout of the per-node prototype head.

The reference keeps one `nn.Conv2d` / `NonNegLinear` per tree node and loops over them in Python
(`pipnet/pipnet.py:124-170`).  Here all nodes share flat axes so one kernel launch covers the tree:

  flat prototype axis  P = sum_n P_n   nodes concatenated in `root.nodes_with_children()` order
  flat child axis      K = sum_n C_n   same order, children in label order
  padded prototype axis P_pad          128-column tiles for the tcgen05 GEMM.  A tile holds
                                       floor(128/S) node segments of one length class S
                                       (segment j starts at column j*S; a node with P_n < S is
                                       zero-padded and masked).  Tiles are sorted by class.

Everything here is host-side integer bookkeeping; it is pure Python/numpy and CPU-testable.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np

TILE_COLS = 128
MAX_SEGS = 16
TILE_INTS = 4 + 3 * MAX_SEGS
SEG_CLASSES = (8, 16, 20, 32, 40, 64)        # instantiated epilogues (pipnet_b200/csrc/cabi.cu)


def seg_class(p_n: int) -> int:
    for s in SEG_CLASSES:
        if p_n <= s:
            return s
    raise Exception(
        f'node with {p_n} prototypes: more than {SEG_CLASSES[-1]} prototypes per node is not supported by the '
        f'fused sm_100a head yet (SURVEY.md section 7, hard part 5)')


@dataclass
class HeadLayout:
    node_names: List[str]
    leaf_names: List[str]                 # sorted == ImageFolder.class_to_idx order
    P_n: np.ndarray                       # [N]
    C_n: np.ndarray                       # [N]
    proto_off: np.ndarray                 # [N+1]
    cls_off: np.ndarray                   # [N+1]
    wc_off: np.ndarray                    # [N+1]
    proto_node: np.ndarray                # [P]
    col_node: np.ndarray                  # [K]
    welem_col: np.ndarray                 # [sum C_n*P_n]
    welem_proto: np.ndarray
    child_w: np.ndarray                   # [K] float32
    path_off: np.ndarray                  # [L+1]
    path_col: np.ndarray
    anc: np.ndarray                       # [L,N] int8
    tiles: np.ndarray                     # [T, TILE_INTS] int32
    row_map: np.ndarray                   # [P_pad] int32 -> flat prototype or -1
    child_proto_slices: Dict[str, List[tuple]] = field(default_factory=dict)
    row_map_c: np.ndarray = None          # [P_c] int32: compact dZ column -> flat prototype or -1

    @property
    def P_c(self): return int(self.row_map_c.shape[0])     # columns of the compact dZ axis (multiple of 8)
    @property
    def N(self): return len(self.node_names)
    @property
    def P(self): return int(self.proto_off[-1])
    @property
    def K(self): return int(self.cls_off[-1])
    @property
    def L(self): return len(self.leaf_names)
    @property
    def P_pad(self): return int(self.tiles.shape[0]) * TILE_COLS
    @property
    def p_max(self): return int(self.P_n.max())
    @property
    def n_welems(self): return int(self.wc_off[-1])


def build_layout(root) -> HeadLayout:
    """`root`: a tree whose internal nodes carry `num_protos` (set by `Node.set_num_protos`)."""
    nodes = root.nodes_with_children()
    N = len(nodes)
    P_n = np.array([int(n.num_protos) for n in nodes], dtype=np.int32)
    C_n = np.array([n.num_children() for n in nodes], dtype=np.int32)
    if (P_n <= 0).any():
        raise Exception('every internal node needs at least one prototype (call set_num_protos first)')
    if (C_n > 127).any():
        raise Exception('more than 127 children per node is not supported')
    proto_off = np.concatenate([[0], np.cumsum(P_n)]).astype(np.int32)
    cls_off = np.concatenate([[0], np.cumsum(C_n)]).astype(np.int32)
    wc_off = np.concatenate([[0], np.cumsum(P_n * C_n)]).astype(np.int32)
    proto_node = np.repeat(np.arange(N, dtype=np.int32), P_n)
    col_node = np.repeat(np.arange(N, dtype=np.int32), C_n)
    welem_col, welem_proto = [], []
    for i in range(N):
        cc, pp = np.meshgrid(np.arange(C_n[i]), np.arange(P_n[i]), indexing='ij')
        welem_col.append((cls_off[i] + cc).ravel())
        welem_proto.append((proto_off[i] + pp).ravel())
    welem_col = np.concatenate(welem_col).astype(np.int32)
    welem_proto = np.concatenate(welem_proto).astype(np.int32)

    child_w = np.ones(int(cls_off[-1]), dtype=np.float32)
    for i, n in enumerate(nodes):
        if n.weights is not None:
            w = np.asarray(n.weights, dtype=np.float32).reshape(-1)
            # weights are listed in children order; the loss indexes them by label (util/custom_losses.py:30)
            child_w[cls_off[i]:cls_off[i + 1]] = w

    leaf_names = sorted(root.leaf_descendents)
    leaf_idx = {nm: i for i, nm in enumerate(leaf_names)}
    node_idx = {n.name: i for i, n in enumerate(nodes)}
    L = len(leaf_names)
    anc = np.full((L, N), -1, dtype=np.int8)
    paths: List[List[int]] = [[] for _ in range(L)]
    # walk every root->leaf path once (iterative DFS carrying the path so far)
    stack = [(root, [])]
    while stack:
        node, path = stack.pop()
        if node.is_leaf():
            l = leaf_idx[node.name]
            for (ni, lab) in path:
                anc[l, ni] = lab
                paths[l].append(int(cls_off[ni]) + lab)
            continue
        ni = node_idx[node.name]
        for child in node.children:
            stack.append((child, path + [(ni, node.children_to_labels[child.name])]))
    path_off = np.concatenate([[0], np.cumsum([len(p) for p in paths])]).astype(np.int32)
    path_col = np.array([c for p in paths for c in p], dtype=np.int32)

    # ---- tile packing
    by_class: Dict[int, List[int]] = {}
    for i in range(N):
        by_class.setdefault(seg_class(int(P_n[i])), []).append(i)
    # The padded axis (128 columns per tile) is what the MMAs of the fused kernels see.  The backward's dZ matrix and the
    # dX / dW GEMMs use a COMPACT column axis instead: a tile contributes only its used columns (segments * S rounded up
    # to 8 columns = 16 bytes), e.g. 120 instead of 128 for six 20-prototype nodes and 24 instead of 128 for a lone one.
    # rec[3] = first compact column of the tile; within a class all full tiles have the same width and are contiguous.
    recs, row_map, row_map_c = [], [], []
    col_c = 0
    for s in sorted(by_class):
        per_tile = TILE_COLS // s
        ids = by_class[s]
        for t0 in range(0, len(ids), per_tile):
            chunk = ids[t0:t0 + per_tile]
            rec = np.zeros(TILE_INTS, dtype=np.int32)
            rec[0], rec[1] = s, len(chunk)
            rec[2] = min(TILE_COLS, ((len(chunk) * s + 15) // 16) * 16)
            rec[3] = col_c
            width = ((len(chunk) * s + 7) // 8) * 8
            rows = np.full(TILE_COLS, -1, dtype=np.int32)
            for j, ni in enumerate(chunk):
                rec[4 + j] = ni
                rec[4 + MAX_SEGS + j] = P_n[ni]
                rec[4 + 2 * MAX_SEGS + j] = proto_off[ni]
                rows[j * s: j * s + P_n[ni]] = np.arange(proto_off[ni], proto_off[ni] + P_n[ni])
            recs.append(rec)
            row_map.append(rows)
            row_map_c.append(rows[:width])
            col_c += width
    tiles = np.stack(recs).astype(np.int32)
    row_map = np.concatenate(row_map).astype(np.int32)
    row_map_c = np.concatenate(row_map_c).astype(np.int32)
    # Row pitch of dZ: a multiple of 64 columns (128 bytes) keeps every 128-byte row segment of a TMA box inside one
    # cache line (pitch 1008 B measured ~15 % slower per k-block in the dX GEMM than 1024 B).  The extra columns belong to
    # the LAST tile (it stores zeros there), which works as long as that tile stays within 128 columns.
    last_width = len(row_map_c) - int(tiles[-1][3])
    pad = (-len(row_map_c)) % 64
    if pad and last_width + pad <= TILE_COLS:
        row_map_c = np.concatenate([row_map_c, np.full(pad, -1, dtype=np.int32)])

    slices = {}
    for n in nodes:
        if n.num_protos_per_child:
            start, sl = 0, []
            for c in n.children:
                k = n.num_protos_per_child[c.name]
                sl.append((n.children_to_labels[c.name], start, start + k))
                start += k
            slices[n.name] = sl
    return HeadLayout([n.name for n in nodes], leaf_names, P_n, C_n, proto_off, cls_off, wc_off, proto_node, col_node,
                      welem_col, welem_proto, child_w, path_off, path_col, anc, tiles, row_map, slices, row_map_c)
