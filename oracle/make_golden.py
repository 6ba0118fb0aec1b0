"""Freezes outputs of the UNMODIFIED reference (imported from /root/reference via oracle/ref_harness.py)
into small fixtures under tests/golden/ so that the GPU box -- where /root/reference does not exist -- can
still pin both the oracle and the CUDA path to the reference.

    python -m oracle.make_golden          # (re)writes tests/golden/*.npz ; run in the build container only

Each fixture holds the inputs (bf16-representable features / kernels so that the bf16 GEMM operands are
exact), the labels and everything the reference produced: pooled, argmax, out, total loss, per-node class /
tanh / orth losses, gradients w.r.t. features, prototype kernels and classifier weights, joint predictions.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import head_oracle as ho                       # noqa: E402
from oracle import ref_harness as rh                       # noqa: E402
from oracle.problems import desc_loss_kwargs, flat_gumbel  # noqa: E402
from pipnet_b200.trees import CUB08, CUB18, CUB27          # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')

# name, edges, tree key, C, H, B, arg overrides, (pretrain, finetune)
CASES = [
    ('cub08_A_train', CUB08, 'cub08', 64, 6, 4, dict(num_features=20), (False, False)),
    ('cub08_A_pretrain', CUB08, 'cub08', 64, 6, 4, dict(num_features=20), (True, False)),
    ('cub18_B_train', CUB18, 'cub18', 64, 6, 6, dict(num_protos_per_child=8, num_features=0), (False, False)),
    ('cub27_A_finetune', CUB27, 'cub27', 64, 6, 6, dict(num_features=12), (False, True)),
    ('cub27_B_train', CUB27, 'cub27', 64, 6, 4, dict(num_protos_per_child=20, num_features=0), (False, False)),
    # the shipped scripts' full recipe (run_pipnet_20protos_multi_runs_seed42.sh:72-94): + tanh_desc, contrasting set, mask pruning
    ('cub27_B_shipped', CUB27, 'cub27', 64, 6, 8, dict(num_protos_per_child=4, num_features=0, tanh_desc='y|0.05',
                                                       minimize_contrasting_set='y', mask_prune_overspecific='y|0|1.1'), (False, False)),
    ('cub18_A_shipped_finetune', CUB18, 'cub18', 64, 6, 6, dict(num_features=12, tanh_desc='y|0.05', minimize_contrasting_set='y',
                                                                mask_prune_overspecific='y|0|1.1'), (False, True)),
]
EXTRA_ARGS = ('tanh_desc', 'minimize_contrasting_set', 'mask_prune_overspecific')


def bf16_round(t):
    return t.to(torch.bfloat16).to(t.dtype)


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, edges, tree_key, C, H, B, over, (pretrain, finetune) in CASES:
        args = rh.make_args(**over)
        net, root = rh.build_reference_net(edges, C, args, seed=5)
        nodes = root.nodes_with_children()
        with torch.no_grad():
            for n in nodes:
                w = getattr(net, '_' + n.name + '_add_on').weight
                w.copy_(bf16_round(w))
        g = torch.Generator().manual_seed(11)
        shipped = any(k in over for k in EXTRA_ARGS)
        if shipped:
            with torch.no_grad():      # non-trivial presence logits (they are zero-mean xavier draws anyway)
                for n in nodes:
                    pp = getattr(net, '_' + n.name + '_proto_presence')
                    pp.copy_(torch.randn(pp.shape, generator=g))
        x = bf16_round(torch.randn(2 * B, C, H, H, generator=g))
        L = len(root.leaf_descendents)
        ys = torch.randint(0, L, (B,), generator=g)
        ys = torch.cat([ys, ys])
        # fp64 run of the reference: the fixture values are "exact" for the given bf16-representable inputs
        ref = rh.run_reference(net, root, x, ys, args, pretrain=pretrain, finetune=finetune, epoch=3, nr_epochs=10,
                               dtype=torch.float64, rng_seed=77 if shipped else None)
        names = [n.name for n in nodes]
        import numpy as _np
        joint = root.distribution_over_furthest_descendents(net=net, batch_size=2 * B, out=ref['out'], device='cpu', softmax_tau=1)
        jn = root.unwrap_names_of_joint(root.names_of_joint_distribution())
        joint = joint[:, _np.argsort(jn)]
        d = dict(tree=np.array(tree_key), C=C, H=H, B=B, pretrain=pretrain, finetune=finetune,
                 num_features=over.get('num_features', 0), per_child=over.get('num_protos_per_child', 0),
                 node_names=np.array(names), x=x.numpy().astype(np.float32), ys=ys.numpy(),
                 w=np.concatenate([getattr(net, '_' + n + '_add_on').weight.detach().flatten(1).numpy() for n in names]).astype(np.float32),
                 wc=np.concatenate([getattr(net, '_' + n + '_classification').weight.detach().numpy().reshape(-1) for n in names]).astype(np.float64),
                 pooled=np.concatenate([ref['pooled'][n].numpy() for n in names], axis=1),
                 argmax=np.concatenate([ref['argmax'][n].numpy() for n in names], axis=1).astype(np.int32),
                 out=np.concatenate([ref['out'][n].numpy() for n in names], axis=1),
                 loss=float(ref['loss']), joint=joint.detach().numpy(),
                 grad_x=(ref['grad_x'].numpy() if ref['grad_x'] is not None else np.zeros(0)),
                 grad_w=np.concatenate([(ref['grads'][n][0].numpy() if ref['grads'][n][0] is not None
                                         else np.zeros((getattr(net, '_' + n + '_add_on').weight.shape[0], C))) for n in names]),
                 grad_wc=np.concatenate([(ref['grads'][n][1].numpy().reshape(-1) if ref['grads'][n][1] is not None
                                          else np.zeros(getattr(net, '_' + n + '_classification').weight.numel())) for n in names]))
        if shipped:
            # The Gumbel noise of the mask-pruning term is drawn inside the reference's loop; the oracle, seeded the same,
            # draws the same numbers at the same places (checked: its loss must equal the reference's) and records them.
            aw = {n: getattr(net, '_' + n + '_add_on').weight.detach().flatten(1).double() for n in names}
            cw = {n: getattr(net, '_' + n + '_classification').weight.detach().double() for n in names}
            pres = {n: getattr(net, '_' + n + '_proto_presence').detach().double() for n in names}
            used = {}
            torch.manual_seed(77)
            label2name = {i: n for i, n in enumerate(sorted(root.leaf_descendents))}
            orc = ho.full_step(x.double(), aw, cw, root, ys, label2name, pretrain=pretrain, finetune=finetune, epoch=3,
                               nr_epochs=10, presence=pres, gumbel_out=used, **desc_loss_kwargs(args))
            assert abs(float(orc['loss']) - float(ref['loss'])) <= 1e-10 * max(1.0, abs(float(ref['loss']))), 'noise replay failed'
            d['gumbel'] = flat_gumbel(used, nodes).numpy()
            d['presence'] = np.concatenate([pres[n].numpy() for n in names])
            d['grad_presence'] = np.concatenate([(ref['grad_presence'][n].numpy() if ref['grad_presence'][n] is not None
                                                  else np.zeros_like(pres[n].numpy())) for n in names])
            d['avg_tanh_desc'] = ref['avg_tanh_desc']
            for k in EXTRA_ARGS:
                d['arg_' + k] = np.array(getattr(args, k))
        for key, src in (('cls', 'class_loss'), ('tanh', 'tanh_loss'), ('orth', 'orth_loss')):
            d[key + '_nodes'] = np.array(sorted(ref[src].keys()))
            d[key + '_vals'] = np.array([float(ref[src][k]) for k in sorted(ref[src].keys())])
        path = os.path.join(OUT, name + '.npz')
        np.savez_compressed(path, **d)
        print(f'{name}: loss {d["loss"]:.6f}  ->  {path} ({os.path.getsize(path) / 1024:.0f} KB)')


if __name__ == '__main__':
    if not rh.available():
        raise SystemExit('reference checkout not present; fixtures can only be regenerated in the build container')
    main()
