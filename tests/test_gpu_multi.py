"""Data-parallel path on >= 2 GPUs: the head's own overlapped NCCL all-reduce (side stream, joined at the
end of the backward pass) must leave on every rank the MEAN over ranks of the gradients each rank computes alone --
prototype kernels, classifier weights and presence logits -- in the eager step and in the CUDA-graph-captured step.
Skipped on single-GPU boxes (the host-side reduction helper is covered on CPU by a world-size-2 gloo test)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    try:
        import torch.distributed as dist
        from oracle.problems import bf16_round, build_net, make_args
        from pipnet_b200 import dist as hd
        from pipnet_b200 import ops, train as tr
        from pipnet_b200.graphs import GraphedHeadStep
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        torch.cuda.set_device(rank)
        dev = torch.device('cuda', rank)
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
        args = make_args(num_features=20, tanh_desc='y|0.05', minimize_contrasting_set='y', mask_prune_overspecific='y|0|1.1')
        net, root = build_net('cub27', 64, args, seed=3)          # same seed -> identical replicas
        net = net.to(dev)
        B, H = 6, 6
        g = torch.Generator().manual_seed(100 + rank)             # different shard per rank
        x = bf16_round(torch.randn(2 * B, 64, H, H, generator=g)).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        ys = torch.randint(0, net.layout.L, (B,), generator=g)
        ys = torch.cat([ys, ys]).to(dev)
        gum = -torch.empty(net.layout.n_welems, 2).exponential_(generator=g).log()
        gum = gum.to(dev)
        w = tr._phase_weights(False, 1, 10, args)
        params = dict(net.named_parameters())

        def loss_fn(xs, y):
            labels = tr.make_labels(net, y)
            f, pf, pooled, out = net(xs, labels=labels)
            return tr.calculate_loss(1, net, {}, f, pf, pooled, out, y, net_normalization_multiplier=net._multiplier,
                                     pretrain=False, finetune=False, criterion=None, train_iter=None, print=False, EPS=1e-8,
                                     root=root, kernel_orth=True, tanh_desc=True, align=False, uni=False, align_pf=True,
                                     tanh=True, args=args, device=dev, labels=labels, gumbel_noise=gum, **w)[0]

        def grads():
            torch.cuda.synchronize()
            return {k: p.grad.detach().clone() for k, p in params.items() if p.grad is not None}

        def run_eager():
            for p in params.values():
                p.grad = None
            loss_fn(x.detach().requires_grad_(True), ys).backward()
            return grads()

        local = run_eager()                                       # no collective: this rank's own gradients
        want = {}
        for k, v in local.items():
            t = v.clone()
            dist.all_reduce(t, op=dist.ReduceOp.AVG)
            want[k] = t
        results = {}
        for fresh in (True, False):
            hd.enable_overlapped_allreduce(fresh_grads=fresh)
            results[f'eager fresh={fresh}'] = run_eager()
        hd.enable_overlapped_allreduce(fresh_grads=True)
        gs = GraphedHeadStep(loss_fn, list(net.parameters()), x, ys)
        for _ in range(2):
            gs.replay()
        results['graph'] = grads()
        hd.disable_overlapped_allreduce()
        bad = []
        for name, got in results.items():
            for k in want:
                if k not in got:
                    bad.append((name, k, 'missing'))
                    continue
                err = float((got[k] - want[k]).abs().max())
                tol = 1e-5 * float(want[k].abs().max()) + 1e-8
                if err > tol:
                    bad.append((name, k, err, tol))
        n_head = sum(1 for k in want if k.endswith('_add_on.weight') or k.endswith('_classification.weight') or k.endswith('_proto_presence'))
        q.put((rank, bad[:5], n_head))
        q.close()
        q.join_thread()      # flush before the hard exit below
        torch.cuda.synchronize()
        dist.barrier()
        os._exit(0)          # graph-captured NCCL work makes process-group teardown unreliable
    except Exception as ex:                                       # surface the failure instead of hanging the parent
        import traceback
        q.put((rank, [('exception', repr(ex), traceback.format_exc()[-1500:])], 0))
        q.close()
        q.join_thread()
        os._exit(1)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_overlapped_allreduce_gives_mean_gradients():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.kill()
    for rank, bad, n_head in res:
        assert not bad, (rank, bad)
        assert n_head == 75              # 25 nodes x (prototype kernels, classifier, presence logits)


# --------------------------------------------------------------------------- DistributedDataParallel, as main_dist.py:330
PHASES = [("pretrain", True, False), ("train", False, False), ("finetune", False, True)]


def _set_phase(net, phase):
    """the per-phase requires_grad flips of main_dist.py:472-485 (pretrain), :574-588 (finetune), :630-658 (train)"""
    for name, p in net.named_parameters():
        if name.endswith('_add_on.weight'):
            p.requires_grad = phase != 'finetune'
        elif '_classification' in name:
            p.requires_grad = phase != 'pretrain'
        elif name.endswith('_proto_presence'):
            p.requires_grad = False
        elif name == '_multiplier':
            p.requires_grad = False


def _ddp_worker(rank, world, port, q):
    try:
        import torch.distributed as dist
        from torch.nn.parallel import DistributedDataParallel as DDP
        from oracle.problems import bf16_round, build_net, make_args
        from pipnet_b200 import train as tr
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        torch.cuda.set_device(rank)
        dev = torch.device('cuda', rank)
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
        args = make_args(num_features=20)
        net, root = build_net('cub27', 64, args, seed=3)          # same seed -> identical replicas
        net = net.to(dev)
        B, H = 6, 6
        g = torch.Generator().manual_seed(200 + rank)             # different shard per rank
        x = bf16_round(torch.randn(2 * B, 64, H, H, generator=g)).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        ys = torch.randint(0, net.layout.L, (B,), generator=g)
        ys = torch.cat([ys, ys]).to(dev)

        def loss_of(model, module, pretrain, finetune):
            labels = tr.make_labels(module, ys)
            f, pf, pooled, out = model(x.detach().requires_grad_(True), labels=labels)
            w = tr._phase_weights(pretrain, 1, 10, args)
            return tr.calculate_loss(1, module, {}, f, pf, pooled, out, ys, net_normalization_multiplier=module._multiplier,
                                     pretrain=pretrain, finetune=finetune, criterion=None, train_iter=None, print=False,
                                     EPS=1e-8, root=root, kernel_orth=True, tanh_desc=False, align=False, uni=False,
                                     align_pf=True, tanh=True, args=args, device=dev, labels=labels, **w)[0]

        def grads(module):
            torch.cuda.synchronize()
            return {k: p.grad.detach().clone() for k, p in module.named_parameters() if p.grad is not None}

        # per-rank gradients without any exchange -> their mean over ranks is what DDP must deliver
        want = {}
        for phase, pretrain, finetune in PHASES:
            _set_phase(net, phase)
            for p in net.parameters():
                p.grad = None
            loss_of(net, net, pretrain, finetune).backward()
            w_ = {}
            for k, v in grads(net).items():
                t = v.clone()
                dist.all_reduce(t, op=dist.ReduceOp.AVG)
                w_[k] = t
            want[phase] = w_
        for p in net.parameters():
            p.requires_grad = True
        net._multiplier.requires_grad = False
        ddp = DDP(net, device_ids=[rank], find_unused_parameters=True, static_graph=False)      # main_dist.py:330
        bad = []
        n_checked = 0
        for phase, pretrain, finetune in PHASES:
            _set_phase(ddp.module, phase)
            for p in ddp.parameters():
                p.grad = None
            loss_of(ddp, ddp.module, pretrain, finetune).backward()
            got = grads(ddp.module)
            for k, wv in want[phase].items():
                if k not in got:
                    bad.append((phase, k, 'missing'))
                    continue
                n_checked += 1
                err = float((got[k] - wv).abs().max())
                tol = 1e-5 * float(wv.abs().max()) + 1e-8
                if err > tol:
                    bad.append((phase, k, err, tol))
            extra = [k for k in got if k not in want[phase] and float(got[k].abs().max()) > 0]
            if extra:
                bad.append((phase, 'unexpected gradients', extra[:3]))
        q.put((rank, bad[:5], n_checked))
        q.close()
        q.join_thread()
        torch.cuda.synchronize()
        dist.barrier()
        os._exit(0)
    except Exception as ex:
        import traceback
        q.put((rank, [('exception', repr(ex), traceback.format_exc()[-1500:])], 0))
        q.close()
        q.join_thread()
        os._exit(1)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_distributed_data_parallel_wrapper_gives_mean_gradients():
    """The reference's only parallel mode: DDP(net, find_unused_parameters=True) (main_dist.py:330) with the per-phase
    requires_grad flips.  Every per-node parameter is an input of the flat gather, so DDP's reducer must see and average
    exactly the gradients a single rank computes."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29950 + (os.getpid() % 300)
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=200) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.kill()
    for rank, bad, n_checked in res:
        assert not bad, (rank, bad)
        assert n_checked >= 25 + 50 + 25          # pretrain: kernels; train: kernels + classifiers; finetune: classifiers
