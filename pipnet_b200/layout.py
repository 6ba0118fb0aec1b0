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
TILE_INTS = 8 + 3 * MAX_SEGS      # {S, nseg, umma_n, dz_col, spill_n, spill_col0, spill_dst, 0, node[16], len[16], poff[16]}
SEG_CLASSES = (8, 16, 20, 32, 40, 64)        # instantiated epilogues (pipnet_b200/csrc/cabi.cu)
SPILL_INTS = 8                     # spill-node record {node, P_n, poff, zoff, dz_col, S class (0 = wide), dz_width, 0}
RIDERS = True                      # move the nodes of a sparsely used last tile into spare pad columns (see build_layout)
MAX_RIDERS = 2
MAX_TILES_WITH_RIDERS = 8


def seg_class(p_n: int) -> int:
    """segment class of a node whose softmax runs inside the GEMM epilogue; 0 = wider than the widest class: the
    node is a SPILL node (raw logits -> scratch matrix -> softmax / pool / dZ in the row kernels, csrc/spill_nodes.cuh)"""
    for s in SEG_CLASSES:
        if p_n <= s:
            return s
    return 0


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
    spill: np.ndarray = None              # [n_spill, SPILL_INTS] int32 spill-node records (may be empty)
    P_s: int = 0                          # columns of the spill-logit scratch matrix Zs[M, P_s] (multiple of 4; 0 = none)

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
    # Fused nodes (P_n <= 64): 128-column tiles of equal-length segments, one launch per segment class; the softmax /
    # max-pool / align arithmetic of these nodes runs in the GEMM epilogue.
    # Spill nodes: the GEMM only produces their raw logits, which the epilogue writes to the scratch matrix Zs[M, P_s]
    # (fp32); softmax, pooling, align and dZ of these nodes run in the row kernels (csrc/spill_nodes.cuh).  Two kinds:
    #   * WIDE nodes (P_n > 64; `num_protos_per_child` x many children, flat trees with `num_protos_per_descendant`,
    #     util/node.py:45-71): their prototypes fill dedicated 128-column tiles densely (no segment structure);
    #   * RIDERS: when the last tile of a class is at most half used and its nodes fit into the spare pad columns of
    #     the other tiles (six 20-column segments leave 8 of 128 columns unused), the nodes ride there and the tile
    #     -- a full pass over the feature matrix for a handful of columns -- disappears.  cub27 with 20 prototypes per
    #     node is the motivating case: 25 nodes = 4 full tiles + ONE node; that fifth pass cost 20 % of K1 and K5.
    # Nodes are dealt to tiles in DEPTH-FIRST PREORDER, not in the flat (breadth-first, `nodes_with_children`) order of the
    # API-visible axes: an image only drives the nodes on its root-to-leaf path, and in preorder those sit next to each
    # other, so fewer (image, tile) and (image, 64-column block) pairs of the backward are active -- cub27: 56 % -> 47 % of
    # the tiles / 56 % -> 37 % of the blocks per image, cub190: 24 % -> 13.5 % / 13.5 % -> 9 % -- and the block-sparse
    # backward (csrc/small_kernels.cuh: DzBlockTables) skips more.  Purely internal: row_map / row_map_c carry the mapping.
    pre_rank = {}
    stack = [root]
    while stack:
        nd = stack.pop()
        if nd.name in node_idx:
            pre_rank[node_idx[nd.name]] = len(pre_rank)
        stack.extend(reversed(nd.children))
    by_class: Dict[int, List[int]] = {}
    for i in sorted(range(N), key=lambda j: pre_rank.get(j, N + j)):
        by_class.setdefault(seg_class(int(P_n[i])), []).append(i)
    wide_ids = by_class.pop(0, [])
    # fused tiles per class, as lists of node ids
    tile_nodes: List[tuple] = []          # (S, [node ids])
    for s_ in sorted(by_class):
        per_tile = TILE_COLS // s_
        ids = by_class[s_]
        for t0 in range(0, len(ids), per_tile):
            tile_nodes.append((s_, ids[t0:t0 + per_tile]))
    rider_ids: List[int] = []
    if RIDERS and len(tile_nodes) >= 2:
        s_last, chunk_last = tile_nodes[-1]
        others = tile_nodes[:-1]
        spare = sum(((TILE_COLS - len(ch) * s_) // 4) * 4 for s_, ch in others)
        need = sum(((int(P_n[i]) + 3) // 4) * 4 for i in chunk_last)
        # worth it only when the saved tile is a large share of the launch (a pass over the feature matrix per tile) and the
        # riders are few: each rider costs one extra row pass in the kernel tails (measured: cub27, 5 -> 4 tiles, K1 -7 %;
        # cub190 with 3 riders out of 32 tiles was slower than keeping the tile)
        if (len(chunk_last) * s_last <= TILE_COLS // 2 and need <= spare and len(chunk_last) <= MAX_RIDERS
                and len(tile_nodes) <= MAX_TILES_WITH_RIDERS):
            rider_ids = list(chunk_last)
            tile_nodes = others
    # Zs columns: riders first (their pieces are handed to the pad slots in order), then the wide nodes
    spill_ids = rider_ids + wide_ids
    zoff, zc = {}, 0
    for i in rider_ids:
        zoff[i] = zc
        zc += ((int(P_n[i]) + 3) // 4) * 4
    n_rider_cols = zc
    for i in wide_ids:
        zoff[i] = zc
        zc += int(P_n[i])
    n_wide_cols = zc - n_rider_cols
    P_s = ((zc + 3) // 4) * 4
    # compact dZ axis: spill nodes first (each rounded up to 8 columns), then the fused tiles' used columns (segments * S
    # rounded up to 8 columns = 16 bytes, e.g. 120 instead of 128 for six 20-prototype nodes).  rec[3] = first compact
    # column of a tile; within a class all full tiles have the same width and are contiguous.
    row_map_c: List[np.ndarray] = []
    col_c = 0
    spill_recs = []
    for i in spill_ids:
        width = ((int(P_n[i]) + 7) // 8) * 8
        cols = np.full(width, -1, dtype=np.int32)
        cols[:P_n[i]] = np.arange(proto_off[i], proto_off[i] + P_n[i])
        row_map_c.append(cols)
        spill_recs.append([i, int(P_n[i]), int(proto_off[i]), zoff[i], col_c, seg_class(int(P_n[i])) if i in rider_ids else 0,
                           width, 0])
        col_c += width
    recs, row_map = [], []
    # the spill columns riding in pad slots: (Zs column, flat prototype or -1) in Zs order
    rider_cols = []
    for i in rider_ids:
        w4 = ((int(P_n[i]) + 3) // 4) * 4
        rider_cols += [(zoff[i] + c, int(proto_off[i]) + c if c < P_n[i] else -1) for c in range(w4)]
    rider_pos = 0
    for s_, chunk in tile_nodes:
        rec = np.zeros(TILE_INTS, dtype=np.int32)
        rec[0], rec[1] = s_, len(chunk)
        used = len(chunk) * s_
        rows = np.full(TILE_COLS, -1, dtype=np.int32)
        for j, ni in enumerate(chunk):
            rec[8 + j] = ni
            rec[8 + MAX_SEGS + j] = P_n[ni]
            rec[8 + 2 * MAX_SEGS + j] = proto_off[ni]
            rows[j * s_: j * s_ + P_n[ni]] = np.arange(proto_off[ni], proto_off[ni] + P_n[ni])
        take = min(((TILE_COLS - used) // 4) * 4, len(rider_cols) - rider_pos)
        if take > 0:
            rec[4], rec[5], rec[6] = take, used, rider_cols[rider_pos][0]
            for c in range(take):
                rows[used + c] = rider_cols[rider_pos + c][1]
            rider_pos += take
            used += take
        rec[2] = min(TILE_COLS, ((used + 15) // 16) * 16)
        rec[3] = col_c
        width = ((len(chunk) * s_ + 7) // 8) * 8
        recs.append(rec)
        row_map.append(rows)
        row_map_c.append(rows[:width].copy())
        if rec[4] > 0:          # rider columns inside the compact width belong to the riders' own dZ columns, not to this tile
            row_map_c[-1][len(chunk) * s_:] = -1
        col_c += width
    assert rider_pos == len(rider_cols)
    # dedicated spill tiles for the wide nodes: dense columns, no segments; they join the launch of the last fused class
    # (or run as a launch of their own with the 32-column epilogue when every node is wide)
    s_attach = tile_nodes[-1][0] if tile_nodes else 32
    wide_cols = [int(proto_off[i]) + c for i in wide_ids for c in range(int(P_n[i]))]
    for t0 in range(0, len(wide_cols), TILE_COLS):
        part = wide_cols[t0:t0 + TILE_COLS]
        n4 = ((len(part) + 3) // 4) * 4
        rec = np.zeros(TILE_INTS, dtype=np.int32)
        rec[0], rec[1] = s_attach, 0
        rec[2] = min(TILE_COLS, ((n4 + 15) // 16) * 16)
        rec[3] = col_c
        rec[4], rec[5], rec[6] = n4, 0, n_rider_cols + t0
        rows = np.full(TILE_COLS, -1, dtype=np.int32)
        rows[:len(part)] = part
        recs.append(rec)
        row_map.append(rows)
    if not recs:
        raise Exception('empty prototype layout')
    tiles = np.stack(recs).astype(np.int32)
    row_map = np.concatenate(row_map).astype(np.int32)
    row_map_c = np.concatenate(row_map_c).astype(np.int32) if row_map_c else np.zeros(0, dtype=np.int32)
    # Row pitch of dZ: a multiple of 64 columns (128 bytes) keeps every 128-byte row segment of a TMA box inside one
    # cache line (pitch 1008 B measured ~15 % slower per k-block in the dX GEMM than 1024 B).  The extra columns belong to
    # the LAST fused tile (it stores zeros there), which works as long as that tile stays within 128 columns; without
    # fused tiles they belong to the last spill node (the row kernels zero-fill up to their record's dz_width).
    pad = (-len(row_map_c)) % 64
    if tile_nodes:
        last_width = len(row_map_c) - int(tiles[len(tile_nodes) - 1][3])
        if pad and last_width + pad <= TILE_COLS:
            row_map_c = np.concatenate([row_map_c, np.full(pad, -1, dtype=np.int32)])
    elif pad:
        row_map_c = np.concatenate([row_map_c, np.full(pad, -1, dtype=np.int32)])
        spill_recs[-1][6] += pad
    spill = (np.array(spill_recs, dtype=np.int32).reshape(-1, SPILL_INTS) if spill_recs
             else np.zeros((0, SPILL_INTS), dtype=np.int32))

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
                      welem_col, welem_proto, child_w, path_off, path_col, anc, tiles, row_map, slices, row_map_c, spill, P_s)
