"""Model-level parity: `pipnet_b200.PIPNet.forward` + `calculate_loss` + backward (everything the
reference does between the backbone output and `optimizer.step()`) against the CPU oracle's
`full_step`, for the three training phases, plus the joint-distribution predictions."""
import argparse

import pytest
import torch
import torch.nn as nn

from oracle import head_oracle as ho
from oracle.problems import make_tree, bf16_round, rel_err

pytestmark = pytest.mark.gpu


from pipnet_b200.fixtures import IdentityBackbone, make_args, build_net


PHASES = [("pretrain", True, False), ("train", False, False), ("finetune", False, True)]
CASES = [("cub08", 64, 6, 4, dict(num_features=20)), ("cub27", 96, 6, 6, dict(num_protos_per_child=10, num_features=0)),
         ("cub18", 128, 7, 5, dict(num_features=12)),
         ("cub27", 64, 6, 5, dict(num_protos_per_child=20, num_features=0)),     # recipe B: P_n up to 60
         ("cub27", 128, 6, 4, dict(num_protos_per_child=30, num_features=0))]    # a 90-prototype (wide, spill) node


# --softmax values the reference parses (pipnet/pipnet.py:130-136): "y|1" (shipped scripts), "y|2", and a bare "y" = 0.2
SOFTMAX_ARGS = [("y|1", 1.0), ("y|2", 2.0), ("y", 0.2)]


@pytest.mark.parametrize("softmax", SOFTMAX_ARGS, ids=[f"softmax-{s[0]}" for s in SOFTMAX_ARGS])
@pytest.mark.parametrize("phase", PHASES, ids=[p[0] for p in PHASES])
@pytest.mark.parametrize("case", CASES, ids=[f'{c[0]}-{i}' for i, c in enumerate(CASES)])
def test_step_matches_oracle(case, phase, softmax):
    from pipnet_b200 import train as tr
    tree, C, H, B, over = case
    _, pretrain, finetune = phase
    softmax_arg, tau = softmax
    args = make_args(softmax=softmax_arg, **over)
    net, root = build_net(tree, C, args)
    names = net.layout.node_names
    g = torch.Generator().manual_seed(17)
    x = bf16_round(torch.randn(2 * B, C, H, H, generator=g))
    ys = torch.randint(0, net.layout.L, (B,), generator=g)
    ys = torch.cat([ys, ys])
    xs = x.cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    labels = tr.make_labels(net, ys.cuda())
    features, proto_features, pooled, out = net(xs, labels=labels)
    w = tr._phase_weights(pretrain, 3, 10, args)
    res = tr.calculate_loss(3, net, {}, features, proto_features, pooled, out, ys.cuda(),
                            net_normalization_multiplier=net._multiplier, pretrain=pretrain, finetune=finetune,
                            criterion=None, train_iter=None, print=False, EPS=1e-8, root=root, kernel_orth=True,
                            align=False, uni=False, align_pf=True, tanh=True, args=args, device='cuda', labels=labels, **w)
    loss = res[0]
    loss.backward()
    torch.cuda.synchronize()

    aw = {n: getattr(net, '_' + n + '_add_on').weight.detach().flatten(1).double().cpu() for n in names}
    cw = {n: getattr(net, '_' + n + '_classification').weight.detach().double().cpu() for n in names}
    label2name = {i: n for i, n in enumerate(net.layout.leaf_names)}
    assert net.softmax_tau == tau
    ref = ho.full_step(x.double(), aw, cw, root, ys, label2name, pretrain=pretrain, finetune=finetune, softmax_tau=tau,
                       epoch=3, nr_epochs=10, cl_weight=args.cl_weight)

    assert rel_err(pooled.flat, torch.cat([ref['pooled'][n] for n in names], 1)) <= 1e-5
    assert rel_err(out.flat, torch.cat([ref['out'][n] for n in names], 1)) <= 1e-5
    assert abs(float(loss) - float(ref['loss'])) <= 1e-5 * max(1.0, abs(float(ref['loss']))), (float(loss), float(ref['loss']))
    for key, idx in (('cls', 1), ('tanh', 3), ('orth', 6)):
        got = {k: v.item() for k, v in res[idx].items()}
        assert set(got) == set(ref[key]), (key, set(got) ^ set(ref[key]))
        for k in got:
            assert abs(got[k] - float(ref[key][k])) <= 2e-5 * max(1.0, abs(float(ref[key][k]))), (key, k, got[k], float(ref[key][k]))
    # gradients: bf16 tolerance (dZ is stored in bf16)
    if ref['grad_x'] is not None and not finetune:
        assert rel_err(xs.grad, ref['grad_x']) <= 2e-2, f"dX {rel_err(xs.grad, ref['grad_x'])}"
    gw = torch.cat([getattr(net, '_' + n + '_add_on').weight.grad.flatten(1) for n in names])
    gw_ref = torch.cat([ref['grad_w'][n] if ref['grad_w'][n] is not None else torch.zeros_like(aw[n]) for n in names])
    assert rel_err(gw, gw_ref) <= 2e-2, f"dW {rel_err(gw, gw_ref)}"
    if not pretrain:
        for n in names:
            gc = getattr(net, '_' + n + '_classification').weight.grad
            rc = ref['grad_cls'][n] if ref['grad_cls'][n] is not None else torch.zeros_like(cw[n])
            assert (gc.double().cpu() - rc).abs().max() <= 1e-4 * max(1e-3, float(rc.abs().max())) + 1e-7, n
    # per-node accuracy counters and fine predictions
    joint_ref = ho.joint_distribution(root, ref['out'], 1.0)
    _, joint = net.get_joint_distribution(out)
    assert rel_err(joint, joint_ref) <= 1e-5
    assert torch.equal(joint.argmax(1).cpu(), joint_ref.argmax(1))


def test_state_dict_keys_match_reference_layout():
    args = make_args()
    net, root = build_net("cub08", 64, args)
    sd = net.state_dict()
    assert tuple(sd['_root_add_on.weight'].shape) == (20, 64, 1, 1)
    assert tuple(sd['_016+181_classification.weight'].shape) == (2, 20)
    assert tuple(sd['_016+181_classification.normalization_multiplier'].shape) == (1,)
    assert tuple(sd['_root_proto_presence'].shape) == (20, 2)
    assert '_multiplier' in sd
    # parameters alias the flat buffers after the first forward, also after .to()/.cuda()
    xs = torch.randn(2, 64, 6, 6, device='cuda').to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    net(xs)
    flat = net._w_group.flat
    p = getattr(net, '_root_add_on').weight
    assert p.data_ptr() == flat.data_ptr()
    # optimizer updates through the per-node Parameter objects are seen by the kernels (same storage)
    with torch.no_grad():
        p.add_(1.0)
    assert float(flat[0]) == float(p.view(-1)[0])


def test_proto_features_materialised_on_demand():
    args = make_args()
    net, root = build_net("cub08", 64, args)
    g = torch.Generator().manual_seed(1)
    x = bf16_round(torch.randn(2, 64, 6, 6, generator=g))
    xs = x.cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    _, pf, pooled, _ = net(xs, inference=True)
    m = pf['root']
    assert tuple(m.shape) == (2, 20, 6, 6) and pf.shape_of('root') == (2, 20, 6, 6)
    torch.testing.assert_close(m.flatten(2).max(dim=2).values.where(pooled['root'] > 0, torch.zeros_like(pooled['root'])),
                               pooled['root'], rtol=1e-5, atol=1e-7)
    assert torch.equal(m.flatten(2).argmax(dim=2).int(), pf.argmax['root'])


def test_unsupported_variants_raise():
    from pipnet_b200 import pipnet as pp
    root = make_tree("cub08", num_features=20)
    pp.base_architecture_to_features['identity'] = lambda pretrained=False: IdentityBackbone(64)
    with pytest.raises(Exception):
        pp.get_network(8, make_args(unitconv2d='y'), root=root)
    with pytest.raises(Exception):
        pp.get_network(8, make_args(add_on_bias=True), root=root)


@pytest.mark.parametrize("mode", ["leave_out", "mask", "both"])
def test_joint_distribution_switches(mode):
    """leave_out_classes / apply_overspecificity_mask of get_joint_distribution (util/node.py:300-385) vs the oracle
    (pinned to the reference in tests/test_oracle_vs_reference.py); the hard presence mask is injected on both sides."""
    args = make_args(num_features=6)
    net, root = build_net("cub27", 64, args)
    L = net.layout
    nodes = root.nodes_with_children()
    g = torch.Generator().manual_seed(4)
    V = 7
    out_flat = (torch.rand(V, L.K, generator=g) * 3).cuda()
    with torch.no_grad():
        for n in L.node_names:
            w = getattr(net, '_' + n + '_classification').weight
            w.mul_((torch.rand(w.shape, generator=g) < 0.5).float().cuda())     # sparse rows: masking can wipe a class
    from pipnet_b200.pipnet import NodeDict
    out = NodeDict(out_flat, L.node_names, L.cls_off)
    leave_out = None
    if mode in ("leave_out", "both"):
        cands = [c.name for n in nodes for c in n.children if c.is_leaf()]
        leave_out = [cands[1], cands[-1]]
    mask_flat = None
    kw = {}
    if mode in ("mask", "both"):
        mask_flat = (torch.rand(L.P, generator=g) < 0.5).float()
        kw = dict(mask={n: mask_flat[int(L.proto_off[i]):int(L.proto_off[i + 1])].double() for i, n in enumerate(L.node_names)},
                  cls_w={n: getattr(net, '_' + n + '_classification').weight.detach().double().cpu() for n in L.node_names})
    ref = ho.joint_distribution(root, {n: out[n].double().cpu() for n in L.node_names}, 1.0, leave_out_classes=leave_out, **kw)
    _, joint = net.get_joint_distribution(out, leave_out_classes=leave_out,
                                          presence_mask=None if mask_flat is None else mask_flat.cuda())
    assert rel_err(joint, ref) <= 1e-5
    assert torch.equal(joint.argmax(1).cpu(), ref.argmax(1))
    plain = ho.joint_distribution(root, {n: out[n].double().cpu() for n in L.node_names}, 1.0)
    assert not torch.allclose(ref, plain)
    if mode != "leave_out":      # drawing the mask on the device works too and still yields a distribution
        _, j2 = net.get_joint_distribution(out, leave_out_classes=leave_out, apply_overspecificity_mask=True)
        torch.testing.assert_close(j2.sum(1), torch.ones(V, device='cuda'), rtol=1e-4, atol=1e-5)
