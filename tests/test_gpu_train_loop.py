"""`train_pipnet` / `test_pipnet` end to end on a synthetic two-view loader with the reference's
optimizer layout (one AdamW group per node, `util/args.py:528-556`)."""
import pytest
import torch

from oracle.problems import build_net, make_args

pytestmark = pytest.mark.gpu


class _DS(torch.utils.data.Dataset):
    def __init__(self, n, C, H, L, two_views, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.x1 = torch.randn(n, C, H, H, generator=g).to(torch.bfloat16)
        self.x2 = (self.x1.float() + 0.1 * torch.randn(n, C, H, H, generator=g)).to(torch.bfloat16)
        self.y = torch.randint(0, L, (n,), generator=g)
        self.two = two_views

    def __len__(self): return self.y.numel()

    def __getitem__(self, i):
        return (self.x1[i], self.x2[i], self.y[i]) if self.two else (self.x1[i], self.y[i])


def _loader(net, two_views):
    ds = _DS(32, 64, 6, net.layout.L, two_views)
    ds.class_to_idx = {n: i for i, n in enumerate(net.layout.leaf_names)}
    return torch.utils.data.DataLoader(ds, batch_size=8, shuffle=False)


def test_train_and_test_epochs_run_and_learn():
    from pipnet_b200 import train as tr
    args = make_args()
    net, root = build_net("cub08", 64, args)
    names = net.layout.node_names
    opt_net = torch.optim.AdamW([{'params': [getattr(net, '_' + n + '_add_on').weight], 'lr': 5e-3} for n in names])
    opt_cls = torch.optim.AdamW([getattr(net, '_' + n + '_classification').weight for n in names], lr=5e-2)
    sch_net = torch.optim.lr_scheduler.CosineAnnealingLR(opt_net, T_max=40)
    sch_cls = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt_cls, T_0=5)
    losses = []
    for epoch in range(1, 4):
        info, log = tr.train_pipnet(net, _loader(net, True), opt_net, opt_cls, sch_net, sch_cls, None, epoch, 3, 'cuda',
                                    pretrain=False, finetune=False, kernel_orth=True, align=False, uni=False, align_pf=True,
                                    tanh=True, wandb_logging=True, args=args)
        losses.append(info['loss'])
        assert 0.0 <= info['fine_accuracy'] <= 1.0 and 'train/epoch loss' in log
        assert set(info['node_accuracy']) == set(names)
    assert losses[-1] < losses[0], losses
    info, _ = tr.test_pipnet(net, _loader(net, False), opt_net, opt_cls, sch_net, sch_cls, None, 1, 3, 'cuda',
                             kernel_orth=True, align=False, uni=False, align_pf=True, tanh=True, args=args)
    assert info['loss'] > 0
    # pretrain phase: classifier untouched
    before = torch.cat([getattr(net, '_' + n + '_classification').weight.detach().flatten() for n in names]).clone()
    tr.train_pipnet(net, _loader(net, True), opt_net, opt_cls, sch_net, None, None, 1, 3, 'cuda', pretrain=True,
                    align=False, uni=False, align_pf=True, tanh=True, args=args)
    after = torch.cat([getattr(net, '_' + n + '_classification').weight.detach().flatten() for n in names])
    assert torch.equal(before, after)
