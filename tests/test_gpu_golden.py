"""The CUDA path (through the C ABI) against the committed reference fixtures: forward outputs, total loss,
gradients, predictions -- the reference's own numbers, not the oracle's."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle.problems import make_args, build_net, rel_err

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', '*.npz')))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_head_reproduces_reference_fixture(path):
    from pipnet_b200 import train as tr
    d = np.load(path, allow_pickle=False)
    C, H, B = int(d['C']), int(d['H']), int(d['B'])
    pretrain, finetune = bool(d['pretrain']), bool(d['finetune'])
    shipped = 'gumbel' in d.files         # fixture of the shipped scripts' full recipe (extra terms + recorded Gumbel noise)
    extra_args = {k[4:]: str(d[k]) for k in d.files if k.startswith('arg_')}
    args = make_args(num_features=int(d['num_features']), num_protos_per_child=int(d['per_child']), **extra_args)
    net, root = build_net(str(d['tree']), C, args)
    names = net.layout.node_names
    assert names == [str(n) for n in d['node_names']]
    with torch.no_grad():                                   # load the fixture's weights into the per-node parameters
        w = torch.from_numpy(d['w']).cuda()
        wc = torch.from_numpy(d['wc']).float().cuda()
        o1 = o2 = 0
        for n in names:
            p = getattr(net, '_' + n + '_add_on').weight
            p.copy_(w[o1:o1 + p.shape[0]].view_as(p)); o1 += p.shape[0]
            q = getattr(net, '_' + n + '_classification').weight
            q.copy_(wc[o2:o2 + q.numel()].view_as(q)); o2 += q.numel()
        if shipped:
            pres, o3 = torch.from_numpy(d['presence']).float().cuda(), 0
            for n in names:
                pp = getattr(net, '_' + n + '_proto_presence')
                pp.copy_(pres[o3:o3 + pp.shape[0]]); o3 += pp.shape[0]
    xs = torch.from_numpy(d['x']).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    ys = torch.from_numpy(d['ys']).cuda()
    labels = tr.make_labels(net, ys)
    features, pf, pooled, out = net(xs, labels=labels)
    w_ = tr._phase_weights(pretrain, 3, 10, args)
    res = tr.calculate_loss(3, net, {}, features, pf, pooled, out, ys, net_normalization_multiplier=net._multiplier,
                            pretrain=pretrain, finetune=finetune, criterion=None, train_iter=None, print=False, EPS=1e-8,
                            root=root, kernel_orth=True, tanh_desc='y' in args.tanh_desc, align=False, uni=False,
                            align_pf=True, tanh=True, args=args, device='cuda', labels=labels,
                            gumbel_noise=torch.from_numpy(d['gumbel']).float().cuda() if shipped else None, **w_)
    res[0].backward()
    torch.cuda.synchronize()
    assert rel_err(pooled.flat, torch.from_numpy(d['pooled'])) <= 1e-5
    assert rel_err(out.flat, torch.from_numpy(d['out'])) <= 1e-5
    assert torch.equal(pf.argmax.flat.cpu(), torch.from_numpy(d['argmax']))
    assert abs(float(res[0].detach()) - float(d['loss'])) <= 1e-5 * max(1.0, abs(float(d['loss'])))
    for key, idx in (('cls', 1), ('tanh', 3), ('orth', 6)):
        got = {k: v.item() for k, v in res[idx].items()}
        want = dict(zip([str(s) for s in d[key + '_nodes']], d[key + '_vals']))
        assert set(got) == set(want)
        for k in got:
            assert abs(got[k] - want[k]) <= 2e-5 * max(1.0, abs(want[k])), (key, k)
    if d['grad_x'].size and not finetune:
        assert rel_err(xs.grad, torch.from_numpy(d['grad_x'])) <= 2e-2
    gw = torch.cat([getattr(net, '_' + n + '_add_on').weight.grad.flatten(1) for n in names])
    assert rel_err(gw, torch.from_numpy(d['grad_w'])) <= 2e-2
    if shipped:
        gp = torch.cat([getattr(net, '_' + n + '_proto_presence').grad for n in names]).double().cpu()
        want = torch.from_numpy(d['grad_presence'])
        assert (gp - want).abs().max() <= 2e-4 * float(want.abs().max()) + 1e-9
        if 'y' in args.tanh_desc and not finetune:
            assert abs(float(res[17]) - float(d['avg_tanh_desc'])) <= 2e-5 * max(1.0, abs(float(d['avg_tanh_desc'])))
    _, joint = net.get_joint_distribution(out)
    ref_joint = torch.from_numpy(d['joint'])
    assert rel_err(joint, ref_joint) <= 1e-5
    assert torch.equal(joint.argmax(1).cpu(), ref_joint.argmax(1))
