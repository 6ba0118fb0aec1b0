"""Parity of the CUDA prototype head (through the C ABI) against the CPU oracle on identical,
bf16-representable inputs.  Tolerances: argmax bit-exact; pooled / losses <= 1e-5 relative (the GEMM
products are exact in fp32, only the accumulation order differs from the oracle's fp64); gradients are
carried through a bf16 dZ, so they get the bf16 tolerance 2e-2 of BASELINE.json."""
import pytest
import torch

from oracle import head_oracle as ho
from oracle.problems import Problem, rel_err, argmax_report

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _both_kernel_families(cta_pair_mode):
    """every head test runs on the CTA-pair (cta_group::2) kernels and on the 1-CTA ones (fixture in conftest.py)"""
    yield

FWD_CASES = [
    # tree, C, H, B, kwargs
    ("cub08", 64, 6, 4, dict(num_features=20)),
    ("cub08", 128, 7, 3, dict(per_child=20)),           # P_n = 20 / 40 (single-child root)
    ("cub27", 96, 6, 5, dict(num_features=20)),
    ("cub18", 64, 8, 4, dict(num_features=12)),          # masked tail inside the S=16 class
    ("synth12:3", 72, 6, 6, dict(per_child=16)),         # S=32 class, C not a multiple of 64
    ("cub27", 768, 26, 2, dict(num_features=20)),        # real ConvNeXt-26 geometry, image straddles tiles
    ("cub27", 128, 7, 3, dict(per_child=20)),            # recipe B: P_n = 20 / 40 / 60 -> classes 20, 40 and 64
    ("synth12:3", 64, 6, 4, dict(per_child=28)),         # P_n = 56: masked tail inside the 64 class
    # spill nodes (layout.py): wide nodes (P_n > 64) on dedicated tiles, finished by the row kernels
    ("cub27", 64, 6, 3, dict(per_child=30)),             # 3-child node: 90 prototypes, beside fused 30 / 60 nodes
    ("cub27", 96, 7, 2, dict(per_child=40)),             # every node wide (80 / 120): no fused tile at all
    ("cub08", 64, 6, 3, dict(per_desc=20)),              # num_protos_per_descendant: P_n = 20 x leaves below (up to 160)
    ("cub27", 768, 26, 1, dict(per_child=20)),           # recipe B at ConvNeXt-26 geometry: the 60-prototype node rides
]


def _oracle_forward(pb, tau=1.0):
    return ho.head_forward(pb.x, pb.w, pb.wc, pb.root, softmax_tau=tau)


# softmax temperatures the reference can run (pipnet/pipnet.py:130-136): "y|1" (shipped scripts), "y|2" (int(tau)) and the
# 0.2 default of a bare "y"
TAUS = [1.0, 2.0, 0.2]


@pytest.mark.parametrize("tau", TAUS, ids=[f"tau{t}" for t in TAUS])
@pytest.mark.parametrize("case", FWD_CASES, ids=[f"{c[0]}-C{c[1]}-H{c[2]}-B{c[3]}-{i}" for i, c in enumerate(FWD_CASES)])
def test_forward_pool_argmax_align(case, tau):
    from pipnet_b200 import ops
    tree, C, H, B, kw = case
    if tau != 1.0 and (C, H) == (768, 26) and B > 1:
        B = 1                       # the real-geometry oracle run is the slow one: one image pair for the extra temperatures
    pb = Problem(tree, C, H, B, seed=7, **kw)
    dl = ops.DeviceLayout(pb.layout, 'cuda')
    feats = pb.features('cuda')
    labels = ops.LabelTables(pb.ys.cuda(), dl, pb.V_first)
    x_rows = ops.feature_rows(feats)
    wp, _wpc = ops.pack_weights(pb.w_flat('cuda'), dl)
    pooled, argmax, align = ops.proj_softmax_pool_raw(x_rows, wp, dl, pb.V, pb.V_first, H * H, tau, labels)
    torch.cuda.synchronize()

    proto, pooled_ref, argmax_ref, _ = _oracle_forward(pb, tau)
    pr = pb.cat_nodes(pooled_ref)
    ar = pb.cat_nodes(argmax_ref)
    assert rel_err(pooled, pr) <= 1e-5, f'pooled rel err {rel_err(pooled, pr)}'
    proto_flat = torch.cat([proto[n].flatten(2) for n in pb.layout.node_names], dim=1)
    # bit-exact argmax, except where the fp64 oracle separates two locations by less than fp32 can resolve (relative gap
    # below 2^-23 = 1.2e-7: the same number in fp32, so "first occurrence" legitimately picks the earlier one)
    nbad, gaps = argmax_report(argmax, ar, proto_flat)
    assert all(abs(g) <= 1.2e-7 for g in gaps) and nbad <= 2, f'{nbad} argmax mismatches, relative gaps {gaps[:8]}'

    masks, _ = ho.node_targets(pb.root, pb.ys, pb.label2name)
    for i, name in enumerate(pb.layout.node_names):
        if masks[name].any():
            ref = ho.align_pf_term(proto[name], masks[name])
            assert abs(float(align[i]) - float(ref)) <= 1e-5 * max(1.0, abs(float(ref))), (name, float(align[i]), float(ref))
        else:
            assert float(align[i]) == 0.0


@pytest.mark.parametrize("V", [1, 3, 8])
def test_forward_unpaired_inference(V):
    """test_pipnet / visualisation call the head without view pairing and with any batch size."""
    from pipnet_b200 import ops
    pb = Problem("cub08", 64, 6, 0, seed=3, num_features=20, paired=False, V=V)
    dl = ops.DeviceLayout(pb.layout, 'cuda')
    x_rows = ops.feature_rows(pb.features('cuda'))
    wp, _wpc = ops.pack_weights(pb.w_flat('cuda'), dl)
    pooled, argmax, _ = ops.proj_softmax_pool_raw(x_rows, wp, dl, pb.V, pb.V_first, 36, 1.0, None, thresh=0.1)
    torch.cuda.synchronize()
    _, pooled_ref, argmax_ref, _ = ho.head_forward(pb.x, pb.w, pb.wc, pb.root, inference=True)
    assert rel_err(pooled, pb.cat_nodes(pooled_ref)) <= 1e-5
    assert torch.equal(argmax.cpu().long(), pb.cat_nodes(argmax_ref))


BWD_CASES = [
    ("cub08", 64, 6, 4, dict(num_features=20)),
    ("cub08", 128, 7, 3, dict(per_child=20)),
    ("cub18", 64, 8, 4, dict(num_features=12)),
    ("cub27", 768, 26, 1, dict(num_features=20)),
    ("cub27", 128, 7, 3, dict(per_child=20)),
    ("synth12:3", 64, 6, 4, dict(per_child=28)),
    ("cub27", 64, 6, 3, dict(per_child=30)),             # wide node (90 prototypes) beside fused ones
    ("cub27", 96, 7, 2, dict(per_child=40)),             # all nodes wide
    ("cub08", 64, 6, 3, dict(per_desc=20)),              # P_n up to 160
]


@pytest.mark.parametrize("tau", TAUS, ids=[f"tau{t}" for t in TAUS])
@pytest.mark.parametrize("case", BWD_CASES, ids=[f"{c[0]}-C{c[1]}-H{c[2]}-B{c[3]}-{i}" for i, c in enumerate(BWD_CASES)])
def test_backward_dx_dw(case, tau):
    """d(sum pooled*G + sum_n a_n * align_n) w.r.t. features and prototype kernels vs oracle autograd."""
    from pipnet_b200 import ops
    tree, C, H, B, kw = case
    pb = Problem(tree, C, H, B, seed=11, **kw)
    L = pb.layout
    dl = ops.DeviceLayout(L, 'cuda')
    g = torch.Generator().manual_seed(5)
    G = torch.randn(pb.V, L.P, generator=g, dtype=torch.float64)
    a = torch.rand(L.N, generator=g, dtype=torch.float64) + 0.5

    feats = pb.features('cuda').requires_grad_(True)
    w_flat = pb.w_flat('cuda').requires_grad_(True)
    labels = ops.LabelTables(pb.ys.cuda(), dl, pb.V_first)
    pooled, align, _, _out = ops.HeadProjPool.apply(feats, w_flat, dl, pb.V_first, tau, labels, 0.0)
    loss = (pooled.double() * G.cuda()).sum() + (align.double() * a.cuda()).sum()
    loss.backward()
    torch.cuda.synchronize()

    x = pb.x.clone().requires_grad_(True)
    w = {k: v.clone().requires_grad_(True) for k, v in pb.w.items()}
    proto, pooled_ref, _, _ = ho.head_forward(x, w, pb.wc, pb.root, softmax_tau=tau)
    masks, _ = ho.node_targets(pb.root, pb.ys, pb.label2name)
    ref = (pb.cat_nodes(pooled_ref) * G).sum()
    for i, name in enumerate(L.node_names):
        if masks[name].any():
            ref = ref + a[i] * ho.align_pf_term(proto[name], masks[name])
    ref.backward()
    gw_ref = torch.cat([w[n].grad for n in L.node_names])
    assert rel_err(feats.grad, x.grad) <= 2e-2, f'dX rel err {rel_err(feats.grad, x.grad)}'
    assert rel_err(w_flat.grad, gw_ref) <= 2e-2, f'dW rel err {rel_err(w_flat.grad, gw_ref)}'



@pytest.mark.parametrize("case", [("cub08", 64, 6, 4, dict(num_features=20)), ("cub27", 768, 26, 1, dict(num_features=20))],
                         ids=["cub08-small", "cub27-convnext26"])
def test_fp32_accurate_mode_on_fp32_inputs(case):
    """Genuine fp32 features / kernels (NOT bf16-representable): the fp32-accurate projection (3-way bf16 split,
    six cross terms through the same tcgen05 kernel) meets the 1e-5 contract; plain bf16 operands meet 2e-2."""
    from pipnet_b200 import ops
    tree, C, H, B, kw = case
    pb = Problem(tree, C, H, B, seed=13, round_bf16=False, **kw)
    dl = ops.DeviceLayout(pb.layout, 'cuda')
    labels = ops.LabelTables(pb.ys.cuda(), dl, pb.V_first)
    feats = pb.features('cuda', dtype=torch.float32)
    w_flat = pb.w_flat('cuda')
    proto, pooled_ref, argmax_ref, _ = _oracle_forward(pb)
    pr, ar = pb.cat_nodes(pooled_ref), pb.cat_nodes(argmax_ref)
    masks, _ = ho.node_targets(pb.root, pb.ys, pb.label2name)
    errs = {}
    for name, prec in (("fp32", ops.PREC_FP32X3), ("bf16", ops.PREC_BF16)):
        pooled, align, argmax, _out = ops.HeadProjPool.apply(feats, w_flat, dl, pb.V_first, 1.0, labels, 0.0, prec)
        torch.cuda.synchronize()
        errs[name] = rel_err(pooled, pr)
        a_err = 0.0
        for i, n in enumerate(pb.layout.node_names):
            if masks[n].any():
                ref = float(ho.align_pf_term(proto[n], masks[n]))
                a_err = max(a_err, abs(float(align[i]) - ref) / max(1.0, abs(ref)))
        errs[name + "_align"] = a_err
        if name == "fp32":
            proto_flat = torch.cat([proto[n].flatten(2) for n in pb.layout.node_names], dim=1)
            nbad, gaps = argmax_report(argmax, ar, proto_flat)
            assert nbad == 0, f'{nbad} argmax mismatches, relative gaps {gaps[:8]}'
    assert errs["fp32"] <= 1e-5 and errs["fp32_align"] <= 1e-5, errs
    assert errs["bf16"] <= 2e-2 and errs["bf16_align"] <= 2e-2, errs
    assert errs["bf16"] > errs["fp32"], errs


def test_fp32_accurate_mode_backward_runs():
    from pipnet_b200 import ops
    pb = Problem("cub08", 64, 6, 3, seed=17, num_features=20, round_bf16=False)
    dl = ops.DeviceLayout(pb.layout, 'cuda')
    labels = ops.LabelTables(pb.ys.cuda(), dl, pb.V_first)
    feats = pb.features('cuda', dtype=torch.float32).requires_grad_(True)
    w_flat = pb.w_flat('cuda').requires_grad_(True)
    g = torch.Generator().manual_seed(1)
    G = torch.randn(pb.V, pb.layout.P, generator=g, dtype=torch.float64)
    pooled, align, _, _out = ops.HeadProjPool.apply(feats, w_flat, dl, pb.V_first, 1.0, labels, 0.0, ops.PREC_FP32X3)
    ((pooled.double() * G.cuda()).sum() + align.double().sum()).backward()
    torch.cuda.synchronize()
    x = pb.x.clone().requires_grad_(True)
    w = {k: v.clone().requires_grad_(True) for k, v in pb.w.items()}
    proto, pooled_ref, _, _ = ho.head_forward(x, w, pb.wc, pb.root, softmax_tau=1.0)
    masks, _ = ho.node_targets(pb.root, pb.ys, pb.label2name)
    ref = (pb.cat_nodes(pooled_ref) * G).sum()
    for n in pb.layout.node_names:
        if masks[n].any():
            ref = ref + ho.align_pf_term(proto[n], masks[n])
    ref.backward()
    assert feats.grad.dtype == torch.float32
    assert rel_err(feats.grad, x.grad) <= 2e-2
    assert rel_err(w_flat.grad, torch.cat([w[n].grad for n in pb.layout.node_names])) <= 2e-2


@pytest.mark.parametrize("case", [("cub08", 64, 6, 3, dict(per_child=20)), ("cub18", 64, 8, 4, dict(num_features=12)),
                                  ("cub27", 64, 7, 3, dict(per_child=20)), ("cub27", 64, 26, 1, dict(num_features=20))],
                         ids=["cub08-B", "cub18-12", "cub27-B", "cub27-26x26"])
def test_dz_store_writes_exactly_its_buffer(case):
    """compute-sanitizer stand-in for the TMA stores into the compact dZ matrix: the buffer is pre-filled with NaN and
    followed by NaN guard rows; after K5 every element inside is a number (padding columns are zeros) and no guard
    element was touched -- ragged tiles, partial last tiles and view halves that end inside a 128-row tile included."""
    from pipnet_b200 import ops
    tree, C, H, B, kw = case
    pb = Problem(tree, C, H, B, seed=11, **kw)
    dl = ops.DeviceLayout(pb.layout, 'cuda')
    V, HW = pb.V, H * H
    M = V * HW
    xr = ops.feature_rows(pb.features('cuda'))
    wp, wpc = ops.pack_weights(pb.w_flat('cuda'), dl)
    labels = ops.LabelTables(pb.ys.cuda(), dl, pb.V_first)
    sp = []
    pooled, argmax, _ = ops.proj_softmax_pool_raw(xr, wp, dl, V, pb.V_first, HW, 1.0, labels, spill_out=sp)
    g = torch.Generator(device='cuda').manual_seed(1)
    gp = torch.randn(V, dl.P, device='cuda', generator=g)
    ga = torch.full((dl.N,), 0.3, device='cuda')
    guard = 256
    big = torch.full((M + guard, dl.P_c), float('nan'), device='cuda', dtype=torch.bfloat16)
    ops.head_backward_raw(xr, wp, wpc, dl, V, pb.V_first, HW, 1.0, argmax, gp, labels, ga, dz_out=big[:M], spill=sp)
    torch.cuda.synchronize()
    inside, outside = big[:M].float(), big[M:].float()
    assert not torch.isnan(inside).any(), f"{int(torch.isnan(inside).sum())} dZ elements were never written"
    assert torch.isnan(outside).all(), "K5 wrote past the end of the dZ matrix"
    pad_cols = torch.from_numpy(pb.layout.row_map_c < 0).cuda()
    assert float(inside[:, pad_cols].abs().max() if bool(pad_cols.any()) else 0.0) == 0.0       # padding columns are zeros
    assert float(inside[:, ~pad_cols].abs().max()) > 0.0


@pytest.mark.parametrize("seed", [1, 4])
def test_all_segment_classes_in_one_model(seed):
    """random tree whose nodes have arbitrary prototype counts (1..64): all six segment classes, masked tails and ragged
    partial tiles in ONE launch sequence -- forward (pooled, argmax, align) and backward (dX, dW) vs the oracle."""
    import numpy as np
    from pipnet_b200 import ops
    from pipnet_b200.node import Node
    from pipnet_b200.trees import build_tree, synthetic_edges
    rng = np.random.default_rng(seed)
    root = build_tree(synthetic_edges(14, seed), Node)
    for n in root.nodes_with_children():
        n.num_protos = int(rng.choice([3, 8, 13, 16, 20, 27, 32, 40, 49, 64]))
        n.num_protos_per_child = {}
        n.set_loss_weightage_using_descendants_count()
    pb = Problem("custom", 64, 7, 3, seed=seed, root=root)
    L = pb.layout
    assert len(set(int(t[0]) for t in L.tiles)) >= 4
    dl = ops.DeviceLayout(L, 'cuda')
    g = torch.Generator().manual_seed(5)
    G = torch.randn(pb.V, L.P, generator=g, dtype=torch.float64)
    a = torch.rand(L.N, generator=g, dtype=torch.float64) + 0.5
    feats = pb.features('cuda').requires_grad_(True)
    w_flat = pb.w_flat('cuda').requires_grad_(True)
    labels = ops.LabelTables(pb.ys.cuda(), dl, pb.V_first)
    pooled, align, argmax, _out = ops.HeadProjPool.apply(feats, w_flat, dl, pb.V_first, 1.0, labels, 0.0)
    loss = (pooled.double() * G.cuda()).sum() + (align.double() * a.cuda()).sum()
    loss.backward()
    torch.cuda.synchronize()

    x = pb.x.clone().requires_grad_(True)
    w = {k: v.clone().requires_grad_(True) for k, v in pb.w.items()}
    proto, pooled_ref, argmax_ref, _ = ho.head_forward(x, w, pb.wc, pb.root)
    assert rel_err(pooled, pb.cat_nodes(pooled_ref)) <= 1e-5
    assert torch.equal(argmax.cpu().long(), pb.cat_nodes(argmax_ref))
    masks, _ = ho.node_targets(pb.root, pb.ys, pb.label2name)
    ref = (pb.cat_nodes(pooled_ref) * G).sum()
    for i, name in enumerate(L.node_names):
        if masks[name].any():
            t = ho.align_pf_term(proto[name], masks[name])
            assert abs(float(align[i].detach()) - float(t.detach())) <= 1e-5 * max(1.0, abs(float(t.detach())))
            ref = ref + a[i] * t
    ref.backward()
    gw_ref = torch.cat([w[n].grad for n in L.node_names])
    assert rel_err(feats.grad, x.grad) <= 2e-2, f'dX rel err {rel_err(feats.grad, x.grad)}'
    assert rel_err(w_flat.grad, gw_ref) <= 2e-2, f'dW rel err {rel_err(w_flat.grad, gw_ref)}'


@pytest.mark.parametrize("geom", [(64, 6, 3), (768, 26, 2)], ids=["C64-H6", "C768-H26"])
def test_rider_tail_matches_standalone_row_kernels(geom, cta_pair_mode):
    """cub27 with 20 prototypes per node: 4 full tiles + ONE node whose columns ride in the pad columns of the other
    tiles (layout.py).  The fused kernels finish that node in their own tail (grid barrier in the forward); the
    stand-alone row kernels (hcomp_set_rider_fold(0)) must give the same pooled table, argmax, align sums and dZ."""
    from pipnet_b200 import ops, _cabi
    Cc, H, B = geom
    pb = Problem("cub27", Cc, H, B, seed=11, num_features=20)
    assert pb.layout.spill.shape[0] == 1 and int(pb.layout.spill[0][5]) == 20
    dl = ops.DeviceLayout(pb.layout, 'cuda')
    V, HW = pb.V, H * H
    xr = ops.feature_rows(pb.features('cuda'))
    wp, wpc = ops.pack_weights(pb.w_flat('cuda'), dl)
    labels = ops.LabelTables(pb.ys.cuda(), dl, pb.V_first)
    g = torch.Generator(device='cuda').manual_seed(3)
    gp = torch.randn(V, dl.P, device='cuda', generator=g)
    ga = torch.full((dl.N,), 0.3, device='cuda')
    out = []
    for fold in (1, 0):
        prev = _cabi.lib().hcomp_set_rider_fold(fold)
        try:
            for _ in range(2):                    # twice: the forward tail's grid barrier must be reusable
                sp = []
                pooled, argmax, align = ops.proj_softmax_pool_raw(xr, wp, dl, V, pb.V_first, HW, 0.7, labels, spill_out=sp)
                _, _, dz = ops.head_backward_raw(xr, wp, wpc, dl, V, pb.V_first, HW, 0.7, argmax, gp, labels, ga, pooled=pooled,
                                                 need_dx=False, need_dw=False, spill=sp)
            torch.cuda.synchronize()
            out.append((pooled.clone(), argmax.clone(), align.clone(), dz.clone()))
        finally:
            _cabi.lib().hcomp_set_rider_fold(prev)
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])
    assert torch.allclose(out[0][2], out[1][2], rtol=1e-6, atol=0)          # double atomics: order of the partial sums
    assert torch.equal(out[0][3], out[1][3])
