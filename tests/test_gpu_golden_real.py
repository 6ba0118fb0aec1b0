"""The CUDA path against REFERENCE-generated fixtures at the real ConvNeXt-tiny-26 geometry (C = 768, 26 x 26: twelve
k-blocks per tile, images straddling the 128-row tiles), cub27 in both "20 prototypes" encodings, tau = 1 and the
reference's default tau = 0.2 -- plus the reference's inference-mode forward and its
`get_joint_distribution(leave_out_classes=...)` on the same inputs.  Inputs are regenerated from the seeded recipe
(oracle/make_golden_real.py) and checked against the digest stored in the fixture."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle.make_golden_real import DW_ROW_STRIDE, DX_LOC_STRIDE, problem_digest
from oracle.problems import Problem, rel_err
from pipnet_b200.fixtures import make_args, build_net

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden_real', '*.npz')))


def _load(path):
    d = np.load(path, allow_pickle=False)
    pb = Problem(str(d['tree']), int(d['C']), int(d['H']), int(d['B']), seed=int(d['seed']),
                 num_features=int(d['num_features']), per_child=int(d['per_child']))
    assert problem_digest(pb) == str(d['digest']), 'seeded input recipe drifted: regenerate tests/golden_real'
    args = make_args(num_features=int(d['num_features']), num_protos_per_child=int(d['per_child']), softmax=str(d['softmax']))
    net, root = build_net(str(d['tree']), pb.C, args)
    names = net.layout.node_names
    assert names == [str(n) for n in d['node_names']]
    with torch.no_grad():
        for n in names:
            p = getattr(net, '_' + n + '_add_on').weight
            p.copy_(pb.w[n].float().cuda().view_as(p))
            q = getattr(net, '_' + n + '_classification').weight
            q.copy_(pb.wc[n].float().cuda().view_as(q))
    return d, pb, args, net, root, names


def test_fixtures_exist():
    assert len(GOLDEN) >= 2


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_training_step_reproduces_reference_at_convnext26_geometry(path, cta_pair_mode):
    from pipnet_b200 import train as tr
    d, pb, args, net, root, names = _load(path)
    xs = pb.features('cuda').requires_grad_(True)
    ys = pb.ys.cuda()
    labels = tr.make_labels(net, ys)
    features, pf, pooled, out = net(xs, labels=labels)
    w_ = tr._phase_weights(False, 3, 10, args)
    res = tr.calculate_loss(3, net, {}, features, pf, pooled, out, ys, net_normalization_multiplier=net._multiplier,
                            pretrain=False, finetune=False, criterion=None, train_iter=None, print=False, EPS=1e-8,
                            root=root, kernel_orth=True, align=False, uni=False, align_pf=True, tanh=True, args=args,
                            device='cuda', labels=labels, **w_)
    res[0].backward()
    torch.cuda.synchronize()
    assert rel_err(pooled.flat, torch.from_numpy(d['pooled'])) <= 1e-5
    assert torch.equal(pf.argmax.flat.cpu(), torch.from_numpy(d['argmax']))           # bit-exact prototype locations
    assert rel_err(out.flat, torch.from_numpy(d['out'])) <= 1e-5
    assert abs(float(res[0].detach()) - float(d['loss'])) <= 1e-5 * max(1.0, abs(float(d['loss'])))
    for key, idx in (('cls', 1), ('tanh', 3), ('orth', 6)):
        got = {k: v.item() for k, v in res[idx].items()}
        want = dict(zip([str(s) for s in d[key + '_nodes']], d[key + '_vals']))
        assert set(got) == set(want)
        for k in got:
            assert abs(got[k] - want[k]) <= 2e-5 * max(1.0, abs(want[k])), (key, k)
    # gradients: bf16 tolerance (dZ is bf16); fixture keeps a strided subset as max-normalised fp16
    gw = torch.cat([getattr(net, '_' + n + '_add_on').weight.grad.flatten(1) for n in names])[::DW_ROW_STRIDE]
    want_w = torch.from_numpy(d['grad_w_f16'].astype(np.float64)) * float(d['grad_w_scale'])
    assert rel_err(gw, want_w) <= 2e-2, f'dW {rel_err(gw, want_w)}'
    gx = xs.grad.float().flatten(2)[:, :, ::DX_LOC_STRIDE]
    want_x = torch.from_numpy(d['grad_x_f16'].astype(np.float64)) * float(d['grad_x_scale'])
    assert rel_err(gx, want_x) <= 2e-2, f'dX {rel_err(gx, want_x)}'
    gc = torch.cat([getattr(net, '_' + n + '_classification').weight.grad.reshape(-1) for n in names]).double().cpu()
    want_c = torch.from_numpy(d['grad_wc'])
    assert (gc - want_c).abs().max() <= 1e-4 * max(1e-3, float(want_c.abs().max())) + 1e-7
    _, joint = net.get_joint_distribution(out)
    ref_joint = torch.from_numpy(d['joint'])
    assert rel_err(joint, ref_joint) <= 1e-5
    assert torch.equal(joint.argmax(1).cpu(), ref_joint.argmax(1))                    # bit-exact predictions


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_inference_forward_and_leave_out_joint_reproduce_reference(path):
    """`net(xs, inference=True)` (pooled < 0.1 -> 0, pipnet/pipnet.py:168-169) and
    `get_joint_distribution(out, leave_out_classes=[...])` (util/node.py:319-323) against the reference's own outputs."""
    d, pb, args, net, root, names = _load(path)
    net.eval()
    with torch.no_grad():
        _, pf, pooled, out = net(pb.features('cuda'), inference=True)
        want_p = torch.from_numpy(d['pooled_inference'])
        assert rel_err(pooled.flat, want_p) <= 1e-5
        assert torch.equal(pooled.flat.cpu() == 0, want_p == 0)                       # the same prototypes are cut at 0.1
        assert rel_err(out.flat, torch.from_numpy(d['out_inference'])) <= 1e-5
        leave_out = [str(s) for s in d['leave_out']]
        _, joint = net.get_joint_distribution(out, leave_out_classes=leave_out)
        want_j = torch.from_numpy(d['joint_leave_out'])
        assert rel_err(joint, want_j) <= 1e-5
        assert torch.equal(joint.argmax(1).cpu(), want_j.argmax(1))
        _, plain = net.get_joint_distribution(out)
        assert not torch.allclose(plain.double().cpu(), want_j)                       # the switch changed the distribution
