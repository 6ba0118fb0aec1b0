"""Full-size configurations of BASELINE.json checked through size-independent properties (the CPU oracle would
take minutes there): pooled == map value at the reported argmax, argmax is the FIRST maximum, identical views give
align = -mean log sum_p S^2, determinism, gradient linearity, sparse spot checks of dX / dW against fp64 on sampled rows."""
import pytest
import torch

from oracle.problems import make_tree, bf16_round
from pipnet_b200.layout import build_layout

pytestmark = pytest.mark.gpu

CONFIGS = {
    # name: tree, protos/node, views V, C, H, channels_last
    "cub27-b64": ("cub27", 20, 128, 768, 26, True),
    "cub190-b32": ("synth190", 20, 64, 768, 26, True),
    "fish38-resnet50-b16": ("synth38", 30, 32, 2048, 28, False),
}


def _setup(name, seed=0):
    from pipnet_b200 import ops
    tree, pn, V, C, H, cl = CONFIGS[name]
    root = make_tree(tree, num_features=pn)
    L = build_layout(root)
    dl = ops.DeviceLayout(L, 'cuda')
    g = torch.Generator(device='cuda').manual_seed(seed)
    x = torch.randn(V, H, H, C, generator=g, device='cuda').to(torch.bfloat16).permute(0, 3, 1, 2)
    if not cl:
        x = x.contiguous()                                   # NCHW-contiguous like the ResNet feature nets
    bound = (6.0 / (C + pn)) ** 0.5
    w = bf16_round((torch.rand(L.P, C, generator=g, device='cuda') * 2 - 1) * bound)
    ys = torch.randint(0, L.L, (V // 2,), generator=g, device='cuda')
    ys = torch.cat([ys, ys])
    labels = ops.LabelTables(ys, dl, V // 2)
    return ops, L, dl, x, w, labels, (V, C, H)


@pytest.mark.parametrize("name", list(CONFIGS))
def test_pool_argmax_consistent_with_materialised_map(name):
    ops, L, dl, x, w, labels, (V, C, H) = _setup(name)
    pooled, align, argmax, _out = ops.HeadProjPool.apply(x, w, dl, V // 2, 1.0, labels, 0.0)
    pooled2, align2, argmax2, _out = ops.HeadProjPool.apply(x, w, dl, V // 2, 1.0, labels, 0.0)
    torch.cuda.synchronize()
    assert torch.equal(pooled, pooled2) and torch.equal(argmax, argmax2)          # deterministic values / locations
    assert float(pooled.min()) > 0 and float(pooled.max()) <= 1.0 + 1e-6
    assert int(argmax.min()) >= 0 and int(argmax.max()) < H * H
    # three nodes (first, middle, last): rebuild their full maps with the plain SIMT kernel and compare
    for ni in (0, L.N // 2, L.N - 1):
        p0, p1 = int(L.proto_off[ni]), int(L.proto_off[ni + 1])
        m = ops.materialize_map(x, w[p0:p1], 1.0).flatten(2)                      # [V, P_n, HW] fp32
        mv, mi = m.max(dim=2)
        torch.testing.assert_close(pooled[:, p0:p1], mv, rtol=2e-5, atol=1e-7)
        at = m.gather(2, argmax[:, p0:p1].long().unsqueeze(-1)).squeeze(-1)
        torch.testing.assert_close(at, mv, rtol=2e-5, atol=1e-7)                   # reported location holds the max


def test_identical_views_align_closed_form():
    """view 2 == view 1 => S1 == S2 and align_n = -mean_{b,hw} log(sum_p S^2 + 1e-12) over the node's images."""
    ops, L, dl, x, w, labels, (V, C, H) = _setup("cub27-b64", seed=3)
    x = torch.cat([x[: V // 2], x[: V // 2]]).contiguous(memory_format=torch.channels_last)
    pooled, align, argmax, _out = ops.HeadProjPool.apply(x, w, dl, V // 2, 1.0, labels, 0.0)
    torch.cuda.synchronize()
    assert torch.equal(pooled[: V // 2], pooled[V // 2:]) and torch.equal(argmax[: V // 2], argmax[V // 2:])
    for ni in (0, 7, L.N - 1):
        p0, p1 = int(L.proto_off[ni]), int(L.proto_off[ni + 1])
        mask = labels.desc[:, ni].bool()
        if mask.any():
            m = ops.materialize_map(x[: V // 2][mask], w[p0:p1], 1.0).double()
            ref = -torch.log((m * m).sum(dim=1) + 1e-12).mean()
            assert abs(float(align[ni]) - float(ref)) <= 2e-5 * max(1.0, abs(float(ref)))
        else:
            assert float(align[ni]) == 0.0


@pytest.mark.parametrize("name", ["cub27-b64", "fish38-resnet50-b16"])
def test_backward_linearity_and_spot_checks(name):
    ops, L, dl, x, w, labels, (V, C, H) = _setup(name, seed=5)
    HW = H * H
    g = torch.Generator(device='cuda').manual_seed(9)
    G1 = torch.randn(V, L.P, generator=g, device='cuda')
    G2 = torch.randn(V, L.P, generator=g, device='cuda')

    def grads(G, a):
        xr = x.detach().clone().requires_grad_(True)
        wr = w.detach().clone().requires_grad_(True)
        pooled, align, argmax, _out = ops.HeadProjPool.apply(xr, wr, dl, V // 2, 1.0, labels, 0.0)
        ((pooled * G).sum() + a * align.sum()).backward()
        return xr.grad.float(), wr.grad, argmax, pooled.detach()

    gx1, gw1, argmax, pooled = grads(G1, 0.0)
    gx2, gw2, _, _ = grads(G2, 0.0)
    gx3, gw3, _, _ = grads(G1 + G2, 0.0)
    torch.cuda.synchronize()
    # the pooled path is linear in the upstream gradient (bf16 dZ rounding => loose tolerance on the sum)
    torch.testing.assert_close(gw3, gw1 + gw2, rtol=3e-2, atol=3e-2 * float(gw3.abs().max()))
    torch.testing.assert_close(gx3, gx1 + gx2, rtol=3e-2, atol=3e-2 * float(gx3.abs().max()))
    # spot check dW for one prototype of one node against fp64 on the rows that matter: with only the pooled path,
    # dZ of node n at (v, argmax) is S*(G - sum G S); rebuild from the materialised map for a few views
    ni = L.N // 3
    p0, p1 = int(L.proto_off[ni]), int(L.proto_off[ni + 1])
    m = ops.materialize_map(x, w[p0:p1], 1.0).flatten(2).double()                 # [V, P_n, HW]
    Gn = torch.zeros_like(m)
    Gn.scatter_(2, argmax[:, p0:p1].long().unsqueeze(-1), G1[:, p0:p1].double().unsqueeze(-1))
    dz = m * (Gn - (Gn * m).sum(dim=1, keepdim=True))                             # [V, P_n, HW]
    xr = x.permute(0, 2, 3, 1).reshape(V, HW, C).double()
    dw_ref = torch.einsum('vph,vhc->pc', dz, xr)
    err = float((gw1[p0:p1].double() - dw_ref).abs().max() / dw_ref.abs().max())
    assert err <= 2e-2, err
    dx_ref = torch.einsum('vph,pc->vhc', dz, w[p0:p1].double())                   # this node's share of dX only
    # dX sums over ALL nodes; check the share of a node-only gradient by zeroing the others
    Gonly = torch.zeros_like(G1)
    Gonly[:, p0:p1] = G1[:, p0:p1]
    gxo, _, _, _ = grads(Gonly, 0.0)
    got = gxo.permute(0, 2, 3, 1).reshape(V, HW, C).double()
    err = float((got - dx_ref).abs().max() / dx_ref.abs().max())
    assert err <= 2e-2, err


def test_large_tree_inference_sweep_shape():
    """BASELINE.json configs[4] (iNat-bird-like tree, inference only): 1486 leaves => 1485 nodes, P = 29 700, single-view
    batch through the reference-facing API (`PIPNet.forward(inference=True)` + `get_joint_distribution`).  Sized down in the
    batch dimension only (the reference could not even materialise its 82 GB map at batch 1024)."""
    from oracle.problems import build_net, make_args
    from pipnet_b200 import ops
    args = make_args(num_features=20)
    net, root = build_net('synth1486', 768, args)
    L = net.layout
    assert (L.N, L.P) == (1485, 29700)
    V, H = 24, 26
    g = torch.Generator(device='cuda').manual_seed(5)
    x = torch.randn(V, H, H, 768, generator=g, device='cuda').to(torch.bfloat16).permute(0, 3, 1, 2)   # channels-last view
    with torch.no_grad():
        _, pf, pooled, out = net(x, inference=True)
        _, joint = net.get_joint_distribution(out)
        pred = joint.argmax(dim=1)
    torch.cuda.synchronize()
    assert tuple(pooled.flat.shape) == (V, L.P) and tuple(out.flat.shape) == (V, L.K) and tuple(joint.shape) == (V, L.L)
    # inference threshold (pipnet/pipnet.py:168-169): nothing in (0, 0.1)
    pv = pooled.flat
    assert bool(((pv == 0) | (pv >= 0.1)).all()) and float(pv.max()) <= 1.0 + 1e-6
    torch.testing.assert_close(joint.sum(dim=1), torch.ones(V, device='cuda'), rtol=1e-4, atol=1e-5)
    # spot-check nodes across the tile range against the plain SIMT map kernel (incl. the argmax the visualisers use)
    w = net.flat_prototype_kernels().detach()
    for ni in (0, 1, L.N // 3, L.N // 2, L.N - 2, L.N - 1):
        p0, p1 = int(L.proto_off[ni]), int(L.proto_off[ni + 1])
        m = ops.materialize_map(x, w[p0:p1], 1.0).flatten(2)
        mv, mi = m.max(dim=2)
        want = torch.where(mv < 0.1, torch.zeros_like(mv), mv)
        torch.testing.assert_close(pv[:, p0:p1], want, rtol=2e-5, atol=1e-7)
        am = pf.argmax.flat[:, p0:p1].long()
        torch.testing.assert_close(m.gather(2, am.unsqueeze(-1)).squeeze(-1), mv, rtol=2e-5, atol=1e-7)
    # the fine prediction is the leaf whose root->leaf path has the largest probability product: check two samples by hand
    names = L.node_names
    node_by_name = {n.name: n for n in root.nodes_with_children()}
    for v in (0, V - 1):
        leaf = L.leaf_names[int(pred[v])]
        prob, node = 1.0, root
        while not node.is_leaf():
            i = names.index(node.name)
            o = out.flat[v, int(L.cls_off[i]):int(L.cls_off[i + 1])].double()
            pr = torch.softmax(torch.log1p(o * o), dim=0)
            child = node.closest_descendent_for(leaf)
            prob *= float(pr[node.children_to_labels[child.name]])
            node = child
        assert abs(prob - float(joint[v, pred[v]])) <= 1e-4 * max(prob, 1e-12)
