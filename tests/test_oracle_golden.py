"""The CPU oracle against the committed golden fixtures (outputs of the unmodified reference, frozen by
oracle/make_golden.py).  Runs anywhere -- this is what pins the oracle on the GPU box."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import head_oracle as ho
from oracle.problems import make_tree

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', '*.npz')))


def load(path):
    d = np.load(path, allow_pickle=False)
    root = make_tree(str(d['tree']), num_features=int(d['num_features']), per_child=int(d['per_child']))
    names = [str(n) for n in d['node_names']]
    assert names == [n.name for n in root.nodes_with_children()]
    return d, root, names


def split_nodes(root, names, flat, per_node_cols):
    out, off = {}, 0
    for n, k in zip(names, per_node_cols):
        out[n] = flat[..., off:off + k]
        off += k
    return out


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_reference_fixture(path):
    d, root, names = load(path)
    nodes = root.nodes_with_children()
    C = int(d['C'])
    pn = [n.num_protos for n in nodes]
    w = torch.from_numpy(d['w']).double()
    aw, off = {}, 0
    for n, k in zip(names, pn):
        aw[n] = w[off:off + k]
        off += k
    wc, off = {}, 0
    wcf = torch.from_numpy(d['wc']).double()
    for node in nodes:
        k = node.num_protos * node.num_children()
        wc[node.name] = wcf[off:off + k].view(node.num_children(), node.num_protos)
        off += k
    x = torch.from_numpy(d['x']).double()
    ys = torch.from_numpy(d['ys'])
    label2name = {i: n for i, n in enumerate(sorted(root.leaf_descendents))}
    extra = {}
    if 'gumbel' in d.files:               # shipped-recipe fixture: optional terms + the Gumbel noise the reference drew
        import argparse
        from oracle.problems import desc_loss_kwargs, split_gumbel
        args = argparse.Namespace(**{k[4:]: str(d[k]) for k in d.files if k.startswith('arg_')})
        pres_flat = torch.from_numpy(d['presence']).double()
        pres, off = {}, 0
        for n, k in zip(names, pn):
            pres[n] = pres_flat[off:off + k]
            off += k
        extra = dict(presence=pres, gumbel=split_gumbel(torch.from_numpy(d['gumbel']).double(), nodes), **desc_loss_kwargs(args))
    res = ho.full_step(x, aw, wc, root, ys, label2name, pretrain=bool(d['pretrain']), finetune=bool(d['finetune']),
                       epoch=3, nr_epochs=10, **extra)
    tol = dict(rtol=1e-10, atol=1e-12)
    torch.testing.assert_close(torch.cat([res['pooled'][n] for n in names], 1), torch.from_numpy(d['pooled']), **tol)
    torch.testing.assert_close(torch.cat([res['out'][n] for n in names], 1), torch.from_numpy(d['out']), **tol)
    assert torch.equal(torch.cat([res['argmax'][n] for n in names], 1).int(), torch.from_numpy(d['argmax']))
    assert abs(float(res['loss']) - float(d['loss'])) <= 1e-10 * max(1.0, abs(float(d['loss'])))
    for key in ('cls', 'tanh', 'orth'):
        got = {k: float(v) for k, v in res[key].items()}
        want = dict(zip([str(s) for s in d[key + '_nodes']], d[key + '_vals']))
        assert set(got) == set(want)
        for k in got:
            assert abs(got[k] - want[k]) <= 1e-10 * max(1.0, abs(want[k]))
    if d['grad_x'].size:
        torch.testing.assert_close(res['grad_x'], torch.from_numpy(d['grad_x']), rtol=1e-8, atol=1e-12)
    gw = torch.cat([res['grad_w'][n] if res['grad_w'][n] is not None else torch.zeros_like(aw[n]) for n in names])
    torch.testing.assert_close(gw, torch.from_numpy(d['grad_w']), rtol=1e-8, atol=1e-12)
    if 'gumbel' in d.files:
        gp = torch.cat([res['grad_presence'][n] if res['grad_presence'][n] is not None else torch.zeros_like(pres[n]) for n in names])
        torch.testing.assert_close(gp, torch.from_numpy(d['grad_presence']), rtol=1e-8, atol=1e-12)
        if res['tanh_desc']:
            mean_td = sum(float(v.detach()) for v in res['tanh_desc'].values()) / len(res['tanh_desc'])
            assert abs(mean_td - float(d['avg_tanh_desc'])) <= 1e-10 * max(1.0, abs(mean_td))
    joint = ho.joint_distribution(root, res['out'], 1.0)
    torch.testing.assert_close(joint, torch.from_numpy(d['joint']), rtol=1e-10, atol=1e-14)


def test_fixtures_exist():
    assert len(GOLDEN) >= 4
