"""The reference's ONLY parallel mode: `DDP(net, device_ids=[rank], find_unused_parameters=True)` (main_dist.py:330) around
`pipnet_b200.PIPNet`, driven through the three phases of main_dist.py with their per-phase `requires_grad` flips
(:472-485 pretraining, :574-593 classifier-only finetune, :637-658 full training) found by `dir(net.module)` + suffix
matching.  After `loss.backward()` every rank must hold the MEAN over ranks of the gradients each rank computes alone
(obtained under `ddp.no_sync()`), for every parameter that is trainable in the phase; frozen and unused parameters
(presence logits, classifiers during pretraining, the backbone stub) must stay without gradient.
2 GPUs; skipped on single-GPU boxes."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _set_phase(ddp, phase):
    """the flips of main_dist.py, verbatim in structure"""
    m = ddp.module
    add_on = {'pretrain': True, 'finetune': False, 'train': True}[phase]
    cls = {'pretrain': False, 'finetune': True, 'train': True}[phase]
    for attr in dir(m):
        if attr.endswith('_add_on'):
            for p in getattr(m, attr).parameters():
                p.requires_grad = add_on
    for attr in dir(m):
        if attr.endswith('_classification'):
            for name, p in getattr(m, attr).named_parameters():
                p.requires_grad = cls and 'multiplier' not in name      # util/args.py:536-537 freezes the multiplier


def _worker(rank, world, port, q):
    try:
        import torch.distributed as dist
        from torch.nn.parallel import DistributedDataParallel as DDP
        from pipnet_b200.fixtures import bf16_round, build_net, make_args
        from pipnet_b200 import train as tr
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        torch.cuda.set_device(rank)
        dev = torch.device('cuda', rank)
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
        args = make_args(num_features=20)
        net, root = build_net('cub27', 64, args, seed=3)           # same seed -> identical replicas
        net = net.to(dev)
        ddp = DDP(net, device_ids=[rank], find_unused_parameters=True, static_graph=False)      # main_dist.py:330
        B, H = 6, 6
        g = torch.Generator().manual_seed(100 + rank)              # a different shard per rank
        x = bf16_round(torch.randn(2 * B, 64, H, H, generator=g)).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        ys = torch.randint(0, net.layout.L, (B,), generator=g)
        ys = torch.cat([ys, ys]).to(dev)
        params = dict(ddp.module.named_parameters())
        bad, checked = [], {}

        def step(pretrain, finetune):
            for p in params.values():
                p.grad = None
            labels = tr.make_labels(ddp, ys)
            f, pf, pooled, out = ddp(x.detach().requires_grad_(not finetune), labels=labels)
            w = tr._phase_weights(pretrain, 1, 10, args)
            res = tr.calculate_loss(1, ddp, {}, f, pf, pooled, out, ys, net_normalization_multiplier=ddp.module._multiplier,
                                    pretrain=pretrain, finetune=finetune, criterion=None, train_iter=None, print=False,
                                    EPS=1e-8, root=root, kernel_orth=True, align=False, uni=False, align_pf=True, tanh=True,
                                    args=args, device=dev, labels=labels, **w)
            res[0].backward()
            torch.cuda.synchronize()
            return {k: p.grad.detach().clone() for k, p in params.items() if p.grad is not None}

        for phase, (pretrain, finetune) in (('pretrain', (True, False)), ('finetune', (False, True)), ('train', (False, False))):
            _set_phase(ddp, phase)
            with ddp.no_sync():
                local = step(pretrain, finetune)                    # this rank's own gradients
            want = {}
            for k in sorted(local):
                t = local[k].clone()
                dist.all_reduce(t, op=dist.ReduceOp.AVG)
                want[k] = t
            got = step(pretrain, finetune)                          # DDP's reducer
            trainable = {k for k, p in params.items() if p.requires_grad}
            expect = {k for k in trainable if k.endswith('_add_on.weight') or k.endswith('_classification.weight')}
            if not expect <= set(got):
                bad.append((phase, 'missing', sorted(expect - set(got))[:4]))
            for k in got:
                if k not in trainable:
                    bad.append((phase, k, 'gradient on a frozen parameter'))
                elif k in want:
                    err = float((got[k] - want[k]).abs().max())
                    tol = 1e-5 * float(want[k].abs().max()) + 1e-8
                    if err > tol:
                        bad.append((phase, k, err, tol))
            checked[phase] = len(expect)
        q.put((rank, bad[:6], checked))
        q.close()
        q.join_thread()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)
    except Exception as ex:                                        # surface the failure instead of hanging the parent
        import traceback
        q.put((rank, [('exception', repr(ex), traceback.format_exc()[-2000:])], {}))
        q.close()
        q.join_thread()
        os._exit(1)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_ddp_wrapped_phases_give_mean_gradients():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29950 + (os.getpid() % 40)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.kill()
    for rank, bad, checked in res:
        assert not bad, (rank, bad)
        assert checked == {'pretrain': 25, 'finetune': 25, 'train': 50}, checked       # 25 nodes: kernels / classifiers / both
