"""The CPU oracle against the reference-generated fixtures at the real ConvNeXt-tiny-26 geometry
(tests/golden_real/*.npz, oracle/make_golden_real.py).  Runs anywhere: this pins the oracle on the GPU box for the
geometry the benchmark runs at, and checks that the seeded input recipe still produces the recorded inputs."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import head_oracle as ho
from oracle.make_golden_real import DW_ROW_STRIDE, DX_LOC_STRIDE, problem_digest
from oracle.problems import Problem

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden_real', '*.npz')))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_reference_at_convnext26_geometry(path):
    d = np.load(path, allow_pickle=False)
    pb = Problem(str(d['tree']), int(d['C']), int(d['H']), int(d['B']), seed=int(d['seed']),
                 num_features=int(d['num_features']), per_child=int(d['per_child']))
    assert problem_digest(pb) == str(d['digest'])
    names = pb.layout.node_names
    parts = str(d['softmax']).split('|')
    tau = float(int(parts[1])) if len(parts) > 1 else 0.2                      # pipnet/pipnet.py:131-136
    res = ho.full_step(pb.x, pb.w, pb.wc, pb.root, pb.ys, pb.label2name, pretrain=False, finetune=False, softmax_tau=tau,
                       epoch=3, nr_epochs=10)
    tol = dict(rtol=1e-9, atol=1e-12)
    torch.testing.assert_close(torch.cat([res['pooled'][n] for n in names], 1), torch.from_numpy(d['pooled']), **tol)
    torch.testing.assert_close(torch.cat([res['out'][n] for n in names], 1), torch.from_numpy(d['out']), **tol)
    assert torch.equal(torch.cat([res['argmax'][n] for n in names], 1).int(), torch.from_numpy(d['argmax']))
    assert abs(float(res['loss']) - float(d['loss'])) <= 1e-9 * max(1.0, abs(float(d['loss'])))
    gw = torch.cat([res['grad_w'][n] if res['grad_w'][n] is not None else torch.zeros_like(pb.w[n]) for n in names])[::DW_ROW_STRIDE]
    want_w = torch.from_numpy(d['grad_w_f16'].astype(np.float64)) * float(d['grad_w_scale'])
    assert float((gw - want_w).abs().max()) <= 1e-3 * float(want_w.abs().max())         # fp16 storage of the fixture
    gx = res['grad_x'].flatten(2)[:, :, ::DX_LOC_STRIDE]
    want_x = torch.from_numpy(d['grad_x_f16'].astype(np.float64)) * float(d['grad_x_scale'])
    assert float((gx - want_x).abs().max()) <= 1e-3 * float(want_x.abs().max())
    joint = ho.joint_distribution(pb.root, res['out'], 1.0)
    torch.testing.assert_close(joint, torch.from_numpy(d['joint']), rtol=1e-9, atol=1e-14)
    # inference-mode forward + leave_out_classes joint
    _, pooled_inf, _, out_inf = ho.head_forward(pb.x, pb.w, pb.wc, pb.root, softmax_tau=tau, inference=True)
    torch.testing.assert_close(torch.cat([pooled_inf[n] for n in names], 1), torch.from_numpy(d['pooled_inference']), **tol)
    jl = ho.joint_distribution(pb.root, out_inf, 1.0, leave_out_classes=[str(s) for s in d['leave_out']])
    torch.testing.assert_close(jl, torch.from_numpy(d['joint_leave_out']), rtol=1e-9, atol=1e-14)
