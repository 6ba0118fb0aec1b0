"""CPU tests of the host side: tree class vs the reference's, flat layout tables, tile packing, the
data-parallel helpers on a world_size-2 gloo group."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_harness as rh
from oracle.problems import make_tree
from pipnet_b200 import layout as lay
from pipnet_b200.node import Node
from pipnet_b200.trees import CUB08, CUB18, CUB27, build_tree, get_tree, synthetic_edges


@pytest.mark.skipif(not rh.available(), reason="reference checkout not present")
@pytest.mark.parametrize("edges", [CUB08, CUB18, CUB27, synthetic_edges(40, 2)], ids=["cub08", "cub18", "cub27", "synth40"])
def test_node_matches_reference_node(edges):
    _, _, ref_node, _ = rh.load()
    a, b = build_tree(edges, Node), build_tree(edges, ref_node.Node)
    na, nb = a.nodes_with_children(), b.nodes_with_children()
    assert [n.name for n in na] == [n.name for n in nb]
    for x, y in zip(na, nb):
        assert x.children_to_labels == y.children_to_labels
        assert x.leaf_descendents == y.leaf_descendents and x.descendents == y.descendents
        assert {k: set(v) for k, v in x.leaf_descendents_of_child.items()} == {k: set(v) for k, v in y.leaf_descendents_of_child.items()}
        for leaf in x.leaf_descendents:
            assert x.closest_descendent_for(leaf).name == y.closest_descendent_for(leaf).name
        for kw in (dict(num_protos_per_descendant=0, num_protos_per_child=0, min_protos=20, split_protos=True),
                   dict(num_protos_per_descendant=0, num_protos_per_child=7, min_protos=0, split_protos=True),
                   dict(num_protos_per_descendant=2, num_protos_per_child=0, min_protos=3, split_protos=True)):
            x.set_num_protos(**kw); y.set_num_protos(**kw)
            assert x.num_protos == y.num_protos and x.num_protos_per_child == y.num_protos_per_child
        x.set_loss_weightage_using_descendants_count(); y.set_loss_weightage_using_descendants_count()
        assert torch.equal(x.weights, y.weights)
    assert a.unwrap_names_of_joint(a.names_of_joint_distribution()) == b.unwrap_names_of_joint(b.names_of_joint_distribution())


def test_tree_fixture_shapes():
    for name, nodes, leaves in (("cub08", 8, 8), ("cub18", 17, 18), ("cub27", 25, 27), ("synth190", 189, 190)):
        r = get_tree(name)
        assert len(r.nodes_with_children()) == nodes and len(r.leaf_descendents) == leaves
    assert get_tree("cub08").num_children() == 1              # single-child root
    assert max(n.num_children() for n in get_tree("cub27").nodes_with_children()) == 3


def _check_layout(L, nodes):
    """invariants of the padded / compact / spill axes that the kernels and run_pair rely on"""
    H = lay.TILE_INTS - 3 * lay.MAX_SEGS          # header words of a tile record
    assert L.P == sum(n.num_protos for n in nodes)
    assert L.P_pad == 128 * L.tiles.shape[0]
    used = L.row_map[L.row_map >= 0]
    assert sorted(used.tolist()) == list(range(L.P))          # every prototype exactly once on the padded axis
    classes = [int(t[0]) for t in L.tiles]
    assert classes == sorted(classes)                          # one launch per class
    spill_nodes = {int(r[0]): r for r in L.spill}
    n_spill_cols = 0
    # spill block of the compact axis first: one 8-column-rounded block per spill node
    col = 0
    for r in L.spill:
        ni, pn, po, zoff, dzc, cls, dzw = (int(x) for x in r[:7])
        assert pn == nodes[ni].num_protos and po == int(L.proto_off[ni]) and dzc == col and dzc % 8 == 0
        assert dzw >= (pn + 7) // 8 * 8 and dzw % 8 == 0
        assert (cls == 0) == (pn > 64) and (cls == 0 or cls == lay.seg_class(pn))
        assert L.row_map_c[col:col + pn].tolist() == list(range(po, po + pn)) and (L.row_map_c[col + pn:col + dzw] == -1).all()
        assert 0 <= zoff and zoff + pn <= L.P_s and (cls == 0 or zoff % 4 == 0)
        col += dzw
    fused = [t for t, rec in enumerate(L.tiles) if int(rec[1]) > 0]
    for t, rec in enumerate(L.tiles):
        S, nseg, umma_n, dz_col, sp_n, sp_c0, sp_dst = (int(x) for x in rec[:7])
        assert S in lay.SEG_CLASSES and 0 <= nseg <= 128 // S and umma_n % 16 == 0 and umma_n <= 128
        assert sp_n % 4 == 0 and nseg * S <= umma_n and (sp_n == 0 or (sp_c0 % 4 == 0 and sp_c0 >= nseg * S and sp_c0 + sp_n <= umma_n))
        assert nseg > 0 or sp_n > 0
        for j in range(nseg):
            ni, ln, po = int(rec[H + j]), int(rec[H + 16 + j]), int(rec[H + 32 + j])
            assert ni not in spill_nodes and lay.seg_class(nodes[ni].num_protos) == S
            assert ln == nodes[ni].num_protos <= S and po == int(L.proto_off[ni])
            assert L.row_map[t * 128 + j * S: t * 128 + j * S + ln].tolist() == list(range(po, po + ln))
        if sp_n:
            # the spill columns of the tile carry exactly the prototypes whose logits land in Zs[sp_dst, +sp_n)
            assert 0 <= sp_dst and sp_dst + sp_n <= L.P_s
            n_spill_cols += sp_n
            for c in range(sp_n):
                proto = int(L.row_map[t * 128 + sp_c0 + c])
                owner = [r for r in L.spill if int(r[3]) <= sp_dst + c < int(r[3]) + int(r[1])]
                assert (proto == int(owner[0][2]) + sp_dst + c - int(owner[0][3])) if owner else proto == -1
        if nseg > 0:
            width = (nseg * S + 7) // 8 * 8
            assert dz_col == col and dz_col % 8 == 0
            assert L.row_map_c[col:col + nseg * S].tolist() == L.row_map[t * 128: t * 128 + nseg * S].tolist()
            assert (L.row_map_c[col + nseg * S:col + width] == -1).all()
            same = [i for i in fused if int(L.tiles[i][0]) == S]
            if t != same[-1]:
                assert nseg == 128 // S      # within a class only the LAST fused tile may be partial
            col += width
        else:
            assert t > (fused[-1] if fused else -1)      # dedicated spill tiles come last
    assert n_spill_cols >= sum(int(r[1]) for r in L.spill)
    assert L.P_c % 8 == 0 and col <= L.P_c < col + 64 and (L.row_map_c[col:] == -1).all()
    if fused:
        last_w = L.P_c - int(L.tiles[fused[-1]][3])
        assert 0 < last_w <= 128
        if L.P_c % 64:
            assert L.P_c == col and (col - int(L.tiles[fused[-1]][3])) + (-col) % 64 > 128
    else:
        assert L.P_c % 64 == 0
    assert sorted(L.row_map_c[L.row_map_c >= 0].tolist()) == list(range(L.P))
    assert L.P_s % 4 == 0


@pytest.mark.parametrize("tree,kw", [("cub27", dict(num_features=20)), ("cub08", dict(per_child=20)),
                                     ("cub18", dict(num_features=12)), ("synth190", dict(num_features=20)),
                                     ("cub27", dict(per_child=40)), ("cub08", dict(per_desc=20)), ("cub27", dict(per_child=20))])
def test_layout_tables(tree, kw):
    root = make_tree(tree, **kw)
    L = lay.build_layout(root)
    nodes = root.nodes_with_children()
    assert L.P == sum(n.num_protos for n in nodes) and L.K == sum(n.num_children() for n in nodes)
    _check_layout(L, nodes)
    # anc / path tables agree with the tree
    for li, leaf in enumerate(L.leaf_names):
        for ni, n in enumerate(nodes):
            want = n.children_to_labels[n.closest_descendent_for(leaf).name] if leaf in n.leaf_descendents else -1
            assert int(L.anc[li, ni]) == want
        cols = L.path_col[L.path_off[li]:L.path_off[li + 1]].tolist()
        assert len(cols) == int((L.anc[li] >= 0).sum())


def test_layout_wide_nodes_and_riders():
    """P_n > 64 (util/node.py:45-55 with many children or 20+ prototypes per child) -> spill nodes on dedicated tiles;
    a nearly empty last tile -> its nodes ride in the pad columns of the other tiles"""
    L = lay.build_layout(make_tree("cub27", num_features=20))      # 25 nodes x 20: 4 full tiles + ONE node -> rider
    assert L.tiles.shape[0] == 4 and L.spill.shape[0] == 1 and int(L.spill[0][1]) == 20 and int(L.spill[0][5]) == 20
    assert [int(t[4]) for t in L.tiles] == [8, 8, 4, 0] and L.P_s == 20 and L.P_c == 512
    L = lay.build_layout(make_tree("cub27", per_child=30))         # 3-child node: 90 prototypes -> wide
    wide = [r for r in L.spill if int(r[5]) == 0]
    assert int(L.P_n.max()) == 90 and len(wide) == int((L.P_n > 64).sum()) >= 1
    L = lay.build_layout(make_tree("cub08", per_desc=20))          # flat-style counts: root gets 20 x leaves
    assert int(L.P_n.max()) == 160 and any(int(t[1]) == 0 for t in L.tiles)
    old = lay.RIDERS
    try:
        lay.RIDERS = False
        L0 = lay.build_layout(make_tree("cub27", num_features=20))
        assert L0.tiles.shape[0] == 5 and L0.spill.shape[0] == 0 and L0.P_s == 0
    finally:
        lay.RIDERS = old


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    from pipnet_b200 import dist as hd
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    a, b = torch.randn(7, 5, generator=g), torch.randn(3, generator=g)
    a0, b0 = a.clone(), b.clone()
    hd.flat_allreduce_mean_([a, None, b])
    gathered = [None] * world
    dist.all_gather_object(gathered, (a0, b0))
    want_a = sum(t[0] for t in gathered) / world
    want_b = sum(t[1] for t in gathered) / world
    ok = torch.allclose(a, want_a, atol=1e-6) and torch.allclose(b, want_b, atol=1e-6)
    lo, hi = hd.shard_range(11, rank, world)
    q.put((rank, bool(ok), lo, hi))
    dist.destroy_process_group()


def test_flat_allreduce_mean_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res)
    assert (res[0][2], res[0][3], res[1][2], res[1][3]) == (0, 6, 6, 11)     # shards cover the batch exactly once


@pytest.mark.parametrize("seed", list(range(12)))
def test_layout_invariants_on_random_trees(seed):
    """random trees with arbitrary per-node prototype counts (1..64): every segment class, ragged partial tiles, the
    compact-axis rules the kernels rely on (run_pair checks the same conditions on the device side of the ABI)."""
    rng = np.random.default_rng(seed)
    leaves = int(rng.integers(3, 60))
    root = build_tree(synthetic_edges(leaves, seed), Node)
    nodes = root.nodes_with_children()
    choices = ([1, 5, 8, 9, 16, 17, 20, 21, 32, 33, 40, 41, 60, 64] if seed % 2 else [20, 40, 60]) + ([65, 80, 130, 300] if seed % 3 == 0 else [])
    for n in nodes:
        n.num_protos = int(rng.choice(choices))
        n.num_protos_per_child = {}
    L = lay.build_layout(root)
    _check_layout(L, nodes)


def test_device_layout_sparse_backward_tables():
    """host tables behind the block-sparse backward: pcol is the inverse of row_map_c, tile_of_node points at the tile
    that holds the node's segment (-1 for spill nodes), and tiles are dealt in depth-first preorder"""
    import numpy as np
    from pipnet_b200 import ops
    from pipnet_b200.fixtures import make_tree
    from pipnet_b200.layout import build_layout
    for tree, kw in (("cub27", dict(num_features=20)), ("cub18", dict(num_features=12)), ("cub27", dict(per_child=30))):
        root = make_tree(tree, **kw)
        L = build_layout(root)
        dl = ops.DeviceLayout(L, 'cpu')
        pcol, rmc = dl.pcol.numpy(), np.asarray(L.row_map_c)
        assert pcol.shape == (L.P,)
        for c, pflat in enumerate(rmc):
            if pflat >= 0:
                assert pcol[pflat] == c
        assert (pcol >= 0).sum() == (rmc >= 0).sum() == L.P            # every prototype has exactly one compact column
        for n in range(L.N):                                            # a node's compact columns are contiguous
            cols = pcol[L.proto_off[n]:L.proto_off[n + 1]]
            assert (np.diff(cols) == 1).all()
        ton = dl.tile_of_node.numpy()
        spill_nodes = set(int(r[0]) for r in L.spill)
        for n in range(L.N):
            if n in spill_nodes:
                assert ton[n] == -1
            else:
                t = ton[n]
                assert t >= 0 and n in [int(x) for x in L.tiles[t, 8:8 + L.tiles[t, 1]]]
        # preorder: within a segment class the fused nodes appear in depth-first order of the tree
        nodes = root.nodes_with_children()
        idx = {nd.name: i for i, nd in enumerate(nodes)}
        pre, stack = [], [root]
        while stack:
            nd = stack.pop()
            if nd.name in idx:
                pre.append(idx[nd.name])
            stack.extend(reversed(nd.children))
        rank = {n: r for r, n in enumerate(pre)}
        by_class = {}
        for t in range(L.tiles.shape[0]):
            by_class.setdefault(int(L.tiles[t, 0]), []).extend(int(x) for x in L.tiles[t, 8:8 + L.tiles[t, 1]])
        for seq in by_class.values():
            r = [rank[n] for n in seq]
            assert r == sorted(r)
