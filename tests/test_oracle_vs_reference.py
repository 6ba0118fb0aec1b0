"""Pins the CPU oracle to the UNMODIFIED reference (imported from /root/reference; build
container only -- skipped on the GPU box, where the committed golden fixtures take over)."""
import pytest
import torch

from oracle import head_oracle as ho
from oracle import ref_harness as rh
from oracle.problems import desc_loss_kwargs
from pipnet_b200.trees import CUB08, CUB18, CUB27, synthetic_edges

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference checkout not present")

CASES = [
    # edges, channels, H, B, args overrides, phase(pretrain, finetune)
    ("cub08-A-train", CUB08, 32, 5, 3, dict(num_features=20), (False, False)),
    ("cub08-B-train", CUB08, 24, 4, 4, dict(num_protos_per_child=6), (False, False)),
    ("cub08-A-pretrain", CUB08, 16, 4, 3, dict(num_features=8), (True, False)),
    ("cub18-A-finetune", CUB18, 16, 3, 5, dict(num_features=10), (False, True)),
    ("cub27-B-train", CUB27, 16, 3, 6, dict(num_protos_per_child=4), (False, False)),
    ("synth12-tau2", synthetic_edges(12, 3), 16, 4, 4, dict(num_features=6, softmax='y|2'), (False, False)),
    # the shipped scripts' extra terms (run_pipnet_20protos_multi_runs_seed42.sh): tanh_desc, contrasting set, mask pruning
    ("cub27-B-shipped", CUB27, 16, 3, 8, dict(num_protos_per_child=4, tanh_desc='y|0.05', minimize_contrasting_set='y',
                                              mask_prune_overspecific='y|0|1.1'), (False, False)),
    ("cub18-A-shipped", CUB18, 16, 3, 6, dict(num_features=10, tanh_desc='y|0.05', minimize_contrasting_set='y|1|0.1',
                                              mask_prune_overspecific='y|0'), (False, False)),
    ("cub18-A-shipped-finetune", CUB18, 16, 3, 6, dict(num_features=10, tanh_desc='y|0.05', minimize_contrasting_set='y',
                                                       mask_prune_overspecific='y|0|1.1'), (False, True)),
    ("cub08-B-maskprune-late", CUB08, 16, 3, 4, dict(num_protos_per_child=6, mask_prune_overspecific='y|5|1.1'), (False, False)),
]


def _run(case, dtype):
    name, edges, C, H, B, over, (pretrain, finetune) = case
    args = rh.make_args(**over)
    net, root = rh.build_reference_net(edges, C, args, seed=5)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2 * B, C, H, H, generator=g, dtype=torch.float64).to(dtype)
    L = len(root.leaf_descendents)
    ys = torch.randint(0, L, (B,), generator=g)
    ys = torch.cat([ys, ys])
    nodes = root.nodes_with_children()
    with torch.no_grad():          # break the symmetric zero init of the presence logits
        for n in nodes:
            pp = getattr(net, '_' + n.name + '_proto_presence')
            pp.copy_(torch.randn(pp.shape, generator=g, dtype=torch.float64).to(pp.dtype))
    ref = rh.run_reference(net, root, x, ys, args, pretrain=pretrain, finetune=finetune, dtype=dtype, rng_seed=77)
    aw = {n.name: getattr(net, '_' + n.name + '_add_on').weight.detach().flatten(1).to(dtype) for n in nodes}
    cw = {n.name: getattr(net, '_' + n.name + '_classification').weight.detach().to(dtype) for n in nodes}
    tau = float(args.softmax.split('|')[1])
    label2name = {i: n for i, n in enumerate(sorted(root.leaf_descendents))}
    presence = {n.name: getattr(net, '_' + n.name + '_proto_presence').detach().to(dtype) for n in nodes}
    torch.manual_seed(77)          # same seed, same draw order -> the oracle sees the reference's Gumbel noise
    ours = ho.full_step(x, aw, cw, root, ys, label2name, pretrain=pretrain, finetune=finetune, softmax_tau=tau,
                        cl_weight=args.cl_weight, presence=presence, **desc_loss_kwargs(args))
    return ref, ours, nodes


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference_fp64(case):
    ref, ours, nodes = _run(case, torch.float64)
    tol = dict(rtol=1e-11, atol=1e-12)
    for n in nodes:
        k = n.name
        torch.testing.assert_close(ours['pooled'][k], ref['pooled'][k], **tol)
        torch.testing.assert_close(ours['out'][k], ref['out'][k], **tol)
        assert torch.equal(ours['argmax'][k], ref['argmax'][k])
        for term, rkey in (('cls', 'class_loss'), ('tanh', 'tanh_loss'), ('orth', 'orth_loss')):
            assert set(ours[term]) == set(ref[rkey]), (term, set(ours[term]) ^ set(ref[rkey]))
            if k in ref[rkey]:
                torch.testing.assert_close(ours[term][k].detach(), ref[rkey][k].to(torch.float64), **tol)
        if k in ours['acc']:
            assert ours['acc'][k] == ref['node_accuracy'][k]
        gw, gc = ref['grads'][k]
        if gw is not None:
            torch.testing.assert_close(ours['grad_w'][k], gw, rtol=1e-9, atol=1e-12)
        if gc is not None:
            torch.testing.assert_close(ours['grad_cls'][k], gc, rtol=1e-9, atol=1e-12)
        gp = ref['grad_presence'][k]
        if gp is not None:
            torch.testing.assert_close(ours['grad_presence'][k], gp, rtol=1e-9, atol=1e-12)
        else:
            assert ours['grad_presence'][k] is None or float(ours['grad_presence'][k].abs().max()) == 0.0
    if ours['tanh_desc']:
        mean_td = sum(float(v) for v in ours['tanh_desc'].values()) / len(ours['tanh_desc'])
        assert abs(mean_td - ref['avg_tanh_desc']) <= 1e-9 * max(1.0, abs(mean_td))
    torch.testing.assert_close(ours['loss'].detach(), ref['loss'], **tol)
    if ref['grad_x'] is not None:
        torch.testing.assert_close(ours['grad_x'], ref['grad_x'], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("case", CASES[:2], ids=[c[0] for c in CASES[:2]])
def test_oracle_matches_reference_fp32(case):
    ref, ours, nodes = _run(case, torch.float32)
    for n in nodes:
        k = n.name
        torch.testing.assert_close(ours['pooled'][k], ref['pooled'][k], rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(ours['out'][k], ref['out'][k], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ours['loss'].detach(), ref['loss'], rtol=1e-5, atol=1e-6)


def test_joint_distribution_matches_reference():
    """`get_joint_distribution` hard-codes device='cuda' (`pipnet/pipnet.py:177-178`), so the
    reference recursion is driven directly on CPU plus the same argsort (`:179-181`)."""
    import numpy as np
    args = rh.make_args(num_features=6)
    net, root = rh.build_reference_net(CUB27, 8, args, seed=2)
    g = torch.Generator().manual_seed(3)
    out = {n.name: torch.rand(5, n.num_children(), generator=g, dtype=torch.float64) * 3 for n in root.nodes_with_children()}
    ref = root.distribution_over_furthest_descendents(net=net, batch_size=5, out=out, device='cpu', softmax_tau=1)
    names = root.unwrap_names_of_joint(root.names_of_joint_distribution())
    ref = ref[:, np.argsort(names)]
    ours = ho.joint_distribution(root, out, 1.0)
    torch.testing.assert_close(ours, ref, rtol=1e-12, atol=1e-14)
    assert torch.equal(ours.argmax(1), ref.argmax(1))


@pytest.mark.parametrize("mode", ["leave_out", "mask", "both"])
def test_joint_distribution_switches_match_reference(mode):
    """leave_out_classes and the test-time overspecificity mask of `distribution_over_furthest_descendents`
    (`util/node.py:300-385`); the hard Gumbel draws coincide under the same seed (same recursion order)."""
    import numpy as np
    args = rh.make_args(num_features=6)
    net, root = rh.build_reference_net(CUB27, 8, args, seed=2)
    g = torch.Generator().manual_seed(3)
    nodes = root.nodes_with_children()
    out = {n.name: torch.rand(5, n.num_children(), generator=g, dtype=torch.float64) * 3 for n in nodes}
    with torch.no_grad():
        for n in nodes:
            pp = getattr(net, '_' + n.name + '_proto_presence')
            pp.copy_(3 * torch.randn(pp.shape, generator=g))
            w = getattr(net, '_' + n.name + '_classification').weight
            w.mul_((torch.rand(w.shape, generator=g) < 0.5).float())          # sparse classifier: masked rows do happen
    leaves = sorted(root.leaf_descendents)
    leave_out = None
    if mode in ("leave_out", "both"):      # two leaf children of different parents
        cands = [c.name for n in nodes for c in n.children if c.is_leaf()]
        leave_out = [cands[0], cands[len(cands) // 2]]
    use_mask = mode in ("mask", "both")
    torch.manual_seed(21)
    ref = root.distribution_over_furthest_descendents(net=net, batch_size=5, out=out, leave_out_classes=leave_out,
                                                      apply_overspecificity_mask=use_mask, device='cpu', softmax_tau=1)
    names = root.unwrap_names_of_joint(root.names_of_joint_distribution())
    ref = ref[:, np.argsort(names)].to(torch.float64)
    torch.manual_seed(21)
    kw = {}
    if use_mask:
        kw = dict(presence={n.name: getattr(net, '_' + n.name + '_proto_presence').detach() for n in nodes},
                  cls_w={n.name: getattr(net, '_' + n.name + '_classification').weight.detach() for n in nodes})
    ours = ho.joint_distribution(root, out, 1.0, leave_out_classes=leave_out, **kw)
    torch.testing.assert_close(ours, ref, rtol=1e-12, atol=1e-14)
    plain = ho.joint_distribution(root, out, 1.0)
    assert not torch.allclose(ours, plain)                                      # the switches did something
