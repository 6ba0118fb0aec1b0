"""Reference-generated fixtures at the REAL geometry of the headline workload (ConvNeXt-tiny-26 feature map:
C = 768, 26 x 26 locations, 12 k-blocks per tile, images straddling the 128-row tiles), cub27 tree in both encodings of
"20 prototypes" (A: 20 per node, P = 500; B: 20 per child, P = 1020), one image pair-batch of B = 2.

    python -m oracle.make_golden_real       # (re)writes tests/golden_real/*.npz ; build container only

Storage: the inputs are NOT stored -- features / kernels / classifier weights / labels come from the seeded recipe of
`oracle.problems.Problem` (CPU torch generators, bit-reproducible across machines for one torch version; the fixture
records SHA-1 digests of the generated tensors so that a drifted generator fails loudly instead of silently).  Outputs
of the unmodified reference (fp64): pooled, argmax, child logits, loss and per-node loss terms, joint leaf
distribution in full; gradients on a strided subset (every 3rd prototype row of dW, every 13th location of dX) as
max-normalised fp16 -- they are compared at the 2e-2 bf16 tolerance, fp16 costs 5e-4 of it.  Each case adds the
inference-mode forward (`inference=True`: pooled < 0.1 -> 0, pipnet/pipnet.py:168-169) and
`get_joint_distribution(leave_out_classes=...)` (util/node.py:319-323) evaluated by the reference.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_harness as rh                 # noqa: E402
from oracle.problems import Problem                  # noqa: E402
from pipnet_b200.trees import CUB27                  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden_real')
DW_ROW_STRIDE, DX_LOC_STRIDE = 3, 13

# name, tree key, edges, C, H, B, seed, Problem kwargs, reference arg overrides (softmax arg -> tau)
CASES = [
    ('cub27_A_convnext26_tau1', 'cub27', CUB27, 768, 26, 2, 101, dict(num_features=20), dict(num_features=20, softmax='y|1')),
    ('cub27_B_convnext26_tau02', 'cub27', CUB27, 768, 26, 2, 102, dict(per_child=20),
     dict(num_protos_per_child=20, num_features=0, softmax='y')),          # bare "y": the reference's default tau = 0.2
]


def sha1(t: torch.Tensor) -> str:
    return hashlib.sha1(t.detach().contiguous().cpu().numpy().tobytes()).hexdigest()


def problem_digest(pb: Problem) -> str:
    h = hashlib.sha1()
    h.update(pb.x.float().numpy().tobytes())
    for n in pb.layout.node_names:
        h.update(pb.w[n].float().numpy().tobytes())
        h.update(pb.wc[n].float().numpy().tobytes())
    h.update(pb.ys.numpy().tobytes())
    return h.hexdigest()


def pack_f16(t: torch.Tensor):
    t = t.detach().double()
    scale = float(t.abs().max().clamp_min(1e-300))
    return (t / scale).to(torch.float16).numpy(), scale


def load_problem_into_reference(net, pb: Problem):
    with torch.no_grad():
        for n in pb.layout.node_names:
            conv = getattr(net, '_' + n + '_add_on')
            conv.weight.copy_(pb.w[n].view_as(conv.weight))
            cls = getattr(net, '_' + n + '_classification')
            cls.weight.copy_(pb.wc[n].view_as(cls.weight))


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, tree_key, edges, C, H, B, seed, pkw, over in CASES:
        pb = Problem(tree_key, C, H, B, seed=seed, **pkw)
        args = rh.make_args(**over)
        net, root = rh.build_reference_net(edges, C, args, seed=5)
        nodes = root.nodes_with_children()
        names = [n.name for n in nodes]
        assert names == pb.layout.node_names
        load_problem_into_reference(net, pb)
        ref = rh.run_reference(net, root, pb.x, pb.ys, args, pretrain=False, finetune=False, epoch=3, nr_epochs=10,
                               dtype=torch.float64)
        joint = root.distribution_over_furthest_descendents(net=net, batch_size=pb.V, out=ref['out'], device='cpu', softmax_tau=1)
        jn = root.unwrap_names_of_joint(root.names_of_joint_distribution())
        order = np.argsort(jn)
        joint = joint[:, order]
        # inference-mode forward + leave_out_classes joint (two leaf children of different parents)
        with torch.no_grad():
            _f, _pf, pooled_inf, out_inf = net(pb.x.double(), inference=True)
        # two leaf children of different NON-root parents (a left-out leaf child of the root would collapse the whole
        # distribution to that leaf's one-hot, util/node.py:319-323 -- covered by tests/test_oracle_vs_reference.py)
        cands = [c.name for n in nodes if n is not root for c in n.children if c.is_leaf()]
        leave_out = [cands[1], cands[len(cands) // 2]]
        joint_lo = root.distribution_over_furthest_descendents(net=net, batch_size=pb.V, out=out_inf, leave_out_classes=leave_out,
                                                               device='cpu', softmax_tau=1)[:, order]
        gw = torch.cat([ref['grads'][n][0] if ref['grads'][n][0] is not None else torch.zeros_like(pb.w[n]) for n in names])
        gw16, gw_scale = pack_f16(gw[::DW_ROW_STRIDE])
        gx = ref['grad_x'].flatten(2)[:, :, ::DX_LOC_STRIDE]
        gx16, gx_scale = pack_f16(gx)
        d = dict(tree=np.array(tree_key), C=C, H=H, B=B, seed=seed, softmax=np.array(over['softmax']),
                 num_features=pkw.get('num_features', 0), per_child=pkw.get('per_child', 0),
                 node_names=np.array(names), digest=np.array(problem_digest(pb)),
                 pooled=np.concatenate([ref['pooled'][n].numpy() for n in names], axis=1),
                 argmax=np.concatenate([ref['argmax'][n].numpy() for n in names], axis=1).astype(np.int32),
                 out=np.concatenate([ref['out'][n].numpy() for n in names], axis=1),
                 loss=float(ref['loss']), joint=joint.detach().numpy(),
                 grad_w_f16=gw16, grad_w_scale=gw_scale, grad_x_f16=gx16, grad_x_scale=gx_scale,
                 grad_wc=np.concatenate([(ref['grads'][n][1].numpy().reshape(-1) if ref['grads'][n][1] is not None
                                          else np.zeros(pb.wc[n].numel())) for n in names]),
                 pooled_inference=np.concatenate([pooled_inf[n].numpy() for n in names], axis=1),
                 out_inference=np.concatenate([out_inf[n].detach().numpy() for n in names], axis=1),
                 leave_out=np.array(leave_out), joint_leave_out=joint_lo.detach().numpy())
        for key, src in (('cls', 'class_loss'), ('tanh', 'tanh_loss'), ('orth', 'orth_loss')):
            d[key + '_nodes'] = np.array(sorted(ref[src].keys()))
            d[key + '_vals'] = np.array([float(ref[src][k]) for k in sorted(ref[src].keys())])
        path = os.path.join(OUT, name + '.npz')
        np.savez_compressed(path, **d)
        print(f'{name}: loss {d["loss"]:.6f}  P={pb.layout.P}  ->  {path} ({os.path.getsize(path) / 1024:.0f} KB)')


if __name__ == '__main__':
    if not rh.available():
        raise SystemExit('reference checkout not present; fixtures can only be regenerated in the build container')
    main()
