"""Visualisation feed: device-side top-k per (prototype, leaf) vs a brute-force restatement of the
reference's heap logic (`util/vis_hpipnet.py:184-290`) on oracle outputs; patch boxes vs the reference helper."""
import heapq
from collections import defaultdict

import pytest
import torch

from oracle import head_oracle as ho
from oracle.problems import bf16_round, build_net, make_args


def test_patch_coordinates_match_reference_helper():
    from oracle import ref_harness as rh
    from pipnet_b200 import vis
    if not rh.available():
        pytest.skip("reference checkout not present")
    import importlib
    import sys
    rh.load()
    sys.path.insert(0, rh.REF_ROOT)
    try:
        try:
            ref = importlib.import_module('util.vis_pipnet')
        except Exception as ex:                      # optional third-party imports of that module are absent here
            pytest.skip(f"reference visualisation module not importable: {ex!r}")
    finally:
        sys.path.remove(rh.REF_ROOT)
    for shape, img in (((1, 26, 26), 224), ((1, 7, 7), 224), ((1, 28, 28), 224)):
        skip = round((img - 32) / (shape[-1] - 1))
        for h in (0, 1, shape[1] // 2, shape[1] - 2, shape[1] - 1):
            for w in (0, 3, shape[2] - 1):
                assert vis.get_img_coordinates(img, shape, 32, skip, h, w) == ref.get_img_coordinates(img, shape, 32, skip, h, w)


@pytest.mark.gpu
@pytest.mark.parametrize("find_non_desc", [False, True])
def test_topk_tables_match_heap_oracle(find_non_desc):
    from pipnet_b200 import vis
    args = make_args(num_features=12)
    net, root = build_net("cub18", 64, args)
    L = net.layout
    names = L.node_names
    nodes = root.nodes_with_children()
    g = torch.Generator().manual_seed(31)
    with torch.no_grad():
        for n in names:           # sparse classifier so that relevance differs per prototype
            w = getattr(net, '_' + n + '_classification').weight
            w.mul_((torch.rand(w.shape, generator=g) < 0.6).float().cuda())
    n_img, H, K = 40, 6, 3
    x = bf16_round(torch.randn(n_img, 64, H, H, generator=g))
    ys = torch.randint(0, L.L, (n_img,), generator=g)
    ds = torch.utils.data.TensorDataset(x.to(torch.bfloat16), ys)
    loader = torch.utils.data.DataLoader(ds, batch_size=7, shuffle=False)          # ragged last batch on purpose
    tk = vis.collect_topk(net, loader, topk=K, find_non_descendants=find_non_desc)
    got = tk.collect(H, H)

    aw = {n: getattr(net, '_' + n + '_add_on').weight.detach().flatten(1).double().cpu() for n in names}
    cw = {n: getattr(net, '_' + n + '_classification').weight.detach().double().cpu() for n in names}
    _, pooled, argmax, _ = ho.head_forward(x.double(), aw, cw, root)
    want = {}
    for node in nodes:
        heaps = defaultdict(lambda: defaultdict(list))
        for i in range(n_img):
            leaf = L.leaf_names[int(ys[i])]
            if leaf not in node.leaf_descendents:
                continue                                              # ModifiedLabelLoader keeps the node's images only
            c = node.children_to_labels[node.closest_descendent_for(leaf).name]
            for p in range(node.num_protos):
                rel = (cw[node.name][:, p] > 1e-3).nonzero().flatten().tolist()
                if not rel or ((c in rel) == find_non_desc):
                    continue
                item = (float(pooled[node.name][i, p]), -i, int(argmax[node.name][i, p]))
                h = heaps[p][leaf]
                if len(h) >= K:
                    heapq.heappushpop(h, item)
                else:
                    heapq.heappush(h, item)
        want[node.name] = {p: {leaf: sorted(h, reverse=True) for leaf, h in d.items()} for p, d in heaps.items()}
    want = {k: v for k, v in want.items() if v}
    assert set(got) == set(want)
    for node_name in want:
        assert set(got[node_name]) == set(want[node_name]), node_name
        for p in want[node_name]:
            assert set(got[node_name][p]) == set(want[node_name][p])
            for leaf, items in want[node_name][p].items():
                g_items = got[node_name][p][leaf]
                assert len(g_items) == len(items)
                for (gs, gi, (gh, gw)), (ws_, wi, wl) in zip(g_items, items):
                    assert gi == -wi and abs(gs - ws_) <= 1e-5 * max(1e-6, abs(ws_)) and gh * H + gw == wl
