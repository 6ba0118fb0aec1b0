#!/usr/bin/env python
"""Benchmark of the HComP-Net prototype head (forward + fused losses + backward) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores (oracle port)

One "step" = one pass of the hot path over one synthetic batch: `PIPNet.forward` on backbone features
(projection + softmax + max-pool + classifier) + `calculate_loss` (align_pf, tanh, kernel_orth, class) +
backward (dZ recompute, dX, dW, classifier / loss gradients) [+ NCCL mean all-reduce of the head gradients
when N > 1].  The backbone is outside the path (identity here); the workload is BASELINE.json configs[1]:
cub27 tree, 20 prototypes per node (P = 500), batch 64 (=> 128 views of 26x26x768 bf16 features) per GPU.

Prints ONE JSON line (rank 0).  `value` = images/s with the batch resident in HBM; `e2e` = the same step
driven from pinned HOST buffers (H2D of the feature batch + labels and D2H of the loss inside the timed
region); `roofline` = the fused projection+softmax+pool kernel against the measured bf16 tensor peak;
`cpu_baseline` = the oracle port of the reference algorithm on this box's host cores (bounded sample).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch

WORKLOADS = {
    # name: tree, protos/node, per-GPU batch, C, H
    'cub27': dict(tree='cub27', num_features=20, batch=64, C=768, H=26),
    'cub08': dict(tree='cub08', num_features=20, batch=8, C=768, H=26),
    'cub190': dict(tree='synth190', num_features=20, batch=32, C=768, H=26),
    # recipe B of SURVEY.md 8(d): 20 prototypes per CHILD (P_n = 20 / 40 / 60, P = 1020)
    'cub27b': dict(tree='cub27', num_features=0, per_child=20, batch=64, C=768, H=26),
    # BASELINE.json config 4: ResNet-50 features (C = 2048, 28 x 28, NCHW-contiguous as torchvision emits them),
    # 30 prototypes per node on a 38-leaf tree, batch 128
    'fish38': dict(tree='synth38', num_features=30, batch=128, C=2048, H=28, nchw=True),
    # BASELINE.json config 5: large-tree INFERENCE sweep (single view, batch 1024/GPU; 1486-leaf synthetic tree: SURVEY 8d)
    'inat': dict(tree='synth1486', num_features=20, batch=1024, C=768, H=26, mode='inference'),
}
METRIC = 'train images/sec (prototype head fwd+bwd)'
UNIT = 'images/s'


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(bf16_burst=d['bf16_tflops'], bf16_sustained=d.get('bf16_tflops_sustained', d['bf16_tflops']),
                    hbm=d['hbm_gbs'], source='measured (MEASURED_PEAKS.json)')
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source='fallback (B200_PROFILING.md)')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '20',
                                          '-i', str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for l in self.lines:
            f = [x.strip() for x in l.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm)}


# --------------------------------------------------------------------------- optional image-level number
def image_level_throughput(a, wl, dev, steps=5):
    """ConvNeXt-tiny-26 (torchvision, the reference's stride relaxation, random init, bf16 autocast, channels-last) -> fused
    head -> loss -> backward through the backbone, on synthetic 224x224 images copied from pinned host memory each step.
    The backbone is library code (cuDNN / cuBLAS via PyTorch) and dominates the time; reported for context only."""
    from pipnet_b200.fixtures import make_args, make_tree
    from pipnet_b200 import pipnet as pp, train as tr
    args = make_args(net='convnext_tiny_26', num_features=wl['num_features'], num_protos_per_child=wl.get('per_child', 0))
    root = make_tree(wl['tree'], num_features=wl['num_features'], per_child=wl.get('per_child', 0))
    torch.manual_seed(1)
    feats, add_on, pool, cls_layers, k = pp.get_network(len(root.leaf_descendents), args, root=root)
    net = pp.PIPNet(len(root.leaf_descendents), k, feats, args, add_on, pool, cls_layers, len(root.nodes_with_children()), root)
    net = net.to(dev).to(memory_format=torch.channels_last)
    net.train()
    B = wl['batch']
    g = torch.Generator().manual_seed(3)
    host_x = torch.rand(2 * B, 3, 224, 224, generator=g).pin_memory()
    y = torch.randint(0, net.layout.L, (B,), generator=g)
    host_y = torch.cat([y, y]).pin_memory()
    w = tr._phase_weights(False, 1, 10, args)
    host_loss = torch.empty((), dtype=torch.float32).pin_memory()

    def one():
        x = host_x.to(dev, non_blocking=True).contiguous(memory_format=torch.channels_last)
        ys = host_y.to(dev, non_blocking=True)
        for p in net.parameters():
            p.grad = None
        labels = tr.make_labels(net, ys)
        with torch.autocast('cuda', dtype=torch.bfloat16):
            feats_ = net._net(x)
        features, pf, pooled, out = net.forward_from_features(feats_, labels=labels) if hasattr(net, 'forward_from_features') \
            else _head_only(net, feats_, labels)
        res = tr.calculate_loss(1, net, {}, features, pf, pooled, out, ys, net_normalization_multiplier=net._multiplier,
                                pretrain=False, finetune=False, criterion=None, train_iter=None, print=False, EPS=1e-8,
                                root=root, kernel_orth=True, align=False, uni=False, align_pf=True, tanh=True, args=args,
                                device=dev, labels=labels, **w)
        res[0].backward()
        host_loss.copy_(res[0].detach(), non_blocking=True)

    for _ in range(2):
        one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {'value': B / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms, 'steps': steps,
            'what': 'ConvNeXt-tiny-26 (torchvision, random init, bf16 autocast) + fused head, fwd+bwd, 224x224 synthetic images '
                    'from pinned host memory; backbone = PyTorch library kernels'}


def _head_only(net, features, labels):
    """PIPNet.forward without re-running the backbone (features already computed under autocast)"""
    saved = net._net
    try:
        net._net = torch.nn.Identity()
        return net(features, labels=labels)
    finally:
        net._net = saved


# --------------------------------------------------------------------------- reference arm / cpu baseline
def oracle_cpu_throughput(wl, batch, steps, warmup, budget_s=25.0):
    """The reference algorithm (oracle port, fp32, all host threads) on a bounded sample of the workload:
    same tree / prototypes / feature geometry, `batch` images per step.  Returns images/s and details."""
    from oracle import head_oracle as ho
    from oracle.problems import Problem
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    pb = Problem(wl['tree'], wl['C'], wl['H'], batch, seed=1234, num_features=wl['num_features'],
                 per_child=wl.get('per_child', 0))
    x = pb.x.float()
    aw = {k: v.float() for k, v in pb.w.items()}
    cw = {k: v.float() for k, v in pb.wc.items()}
    times = []
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        ho.full_step(x, aw, cw, pb.root, pb.ys, pb.label2name, pretrain=False, finetune=False)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_begin > budget_s and len(times) >= 1:
            break
    mean = sum(times) / len(times)
    return batch / mean, dict(cores=cores, steps=len(times), s_per_step=mean,
                              sample=f'{wl["tree"]} tree, P={pb.layout.P}, batch {batch} ({2 * batch} views of '
                                     f'{wl["H"]}x{wl["H"]}x{wl["C"]}), fp32 oracle port, fwd+losses+bwd')


def quick_train_record(a, name, dev, rank, world, steps=20):
    """Reduced measurement of another training workload inside the same process: graph-captured step, device timing, max
    over ranks; at N > 1 also the step with the gradient exchange disabled (difference = exposed collective time)."""
    import torch.distributed as dist
    from pipnet_b200 import ops
    from pipnet_b200 import train as tr
    from pipnet_b200.fixtures import build_net, make_args
    from pipnet_b200.graphs import GraphedHeadStep
    wl = WORKLOADS[name]
    B, C, H = wl['batch'], wl['C'], wl['H']
    V = 2 * B
    args = make_args(num_features=wl['num_features'], num_protos_per_child=wl.get('per_child', 0))
    net, root = build_net(wl['tree'], C, args, seed=1)
    net = net.to(dev)
    net.train()
    L = net.layout
    g = torch.Generator().manual_seed(4321 + rank)
    feats, labels = [], []
    for i in range(2):
        x = torch.randn(V, H, H, C, generator=g, dtype=torch.float32).to(torch.bfloat16)
        feats.append(x.to(dev).permute(0, 3, 1, 2))
        y = torch.randint(0, L.L, (B,), generator=torch.Generator().manual_seed(17 + i + 100 * rank))
        labels.append(torch.cat([y, y]).to(dev))
    w = tr._phase_weights(False, 1, 10, args)

    def loss_fn(x, ys):
        lab = tr.make_labels(net, ys)
        features, pf, pooled, out = net(x, labels=lab)
        res = tr.calculate_loss(1, net, {}, features, pf, pooled, out, ys, net_normalization_multiplier=net._multiplier,
                                pretrain=False, finetune=False, criterion=None, train_iter=None, print=False, EPS=1e-8,
                                root=root, kernel_orth=True, tanh_desc=False, align=False, uni=False, align_pf=True,
                                tanh=True, args=args, device=dev, labels=lab, **w)
        return res[0]

    def timed(graphs):
        for i in range(3):
            graphs[i % 2].replay()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            graphs[i % 2].replay()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t) / steps

    params = list(net.parameters())
    graphs = [GraphedHeadStep(loss_fn, params, feats[i], labels[i]) for i in range(2)]
    ms = timed(graphs)
    rec = {'value': world * B / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms, 'steps': steps,
           'config': f'{name}: tree {wl["tree"]} ({L.N} nodes, P={L.P}), batch {B}/GPU, fwd+bwd, one CUDA graph replay per step',
           'grad_bucket_bytes': int(sum(p.numel() for p in params if p.requires_grad) * 4)}
    if world > 1:
        del graphs
        saved, ops.GRAD_ALLREDUCE_GROUP = ops.GRAD_ALLREDUCE_GROUP, None
        try:
            local = [GraphedHeadStep(loss_fn, params, feats[i], labels[i]) for i in range(2)]
            ms_local = timed(local)
        finally:
            ops.GRAD_ALLREDUCE_GROUP = saved
        rec['ms_per_step_without_exchange'] = ms_local
        rec['allreduce_exposed_us'] = (ms - ms_local) * 1e3
    return rec


def run_inference(a):
    """BASELINE.json config 5: inference sweep -- single view, prototype head forward (projection + softmax + max-pool,
    inference threshold) + node classifiers + joint leaf prediction; no collective (independent images per rank)."""
    import torch.distributed as dist
    from pipnet_b200 import _cabi, ops
    from pipnet_b200.fixtures import build_net, make_args
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the prototype head has no CPU path')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    wl = WORKLOADS[a.workload]
    B, C, H = wl['batch'], wl['C'], wl['H']
    V, HW = B, H * H
    args = make_args(num_features=wl['num_features'], num_protos_per_child=wl.get('per_child', 0))
    net, root = build_net(wl['tree'], C, args, seed=1)
    net = net.to(dev)
    net.eval()
    L = net.layout
    g = torch.Generator().manual_seed(99 + rank)
    feats = []
    for i in range(2):
        x = torch.randn(V, H, H, C, generator=g, dtype=torch.float32).to(torch.bfloat16)
        feats.append(x.to(dev).permute(0, 3, 1, 2))

    def fwd(x):
        with torch.no_grad():
            _f, _pf, pooled, out = net(x, inference=True)
            _root_out, joint = net.get_joint_distribution(out, device=dev)
            return joint.argmax(dim=1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(3):
        fwd(feats[i % 2])
    barrier()
    launches0 = _cabi.lib().hcomp_launch_count()
    fwd(feats[0])
    torch.cuda.synchronize()
    launches_per_step = _cabi.lib().hcomp_launch_count() - launches0
    static_x = [f.clone() for f in feats]
    graphs, outs = [], []
    try:
        for i in range(2):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                outs.append(fwd(static_x[i]))
            graphs.append(gr)
    except Exception as ex:
        sys.stderr.write(f'graph capture unavailable, timing the eager forward: {ex!r}\n')
        graphs = None

    def run_step(i):
        if graphs is not None:
            graphs[i % 2].replay()
            return outs[i % 2]
        return fwd(feats[i % 2])

    for i in range(max(a.warmup, 3)):
        run_step(i)
    # K1 alone, CUDA events around the C-ABI call in a saturated loop
    ops.PROFILE.reset()
    ops.PROFILE.enabled = True
    for i in range(5):
        fwd(feats[i % 2])
    torch.cuda.synchronize()
    ops.PROFILE.enabled = False
    prof = ops.PROFILE.totals_ms()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        run_step(i)
    e1.record()
    barrier()
    clocks = sampler.stop()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t) / a.steps
    value = world * B / (ms_per_step * 1e-3)
    # end to end: features from pinned host memory, predictions back to the host
    host_x = [f.permute(0, 2, 3, 1).contiguous().cpu().pin_memory() for f in feats]
    host_pred = torch.empty(V, dtype=torch.int64).pin_memory()
    e2e_steps = max(3, min(a.steps, 10))
    barrier()
    e0.record()
    for i in range(e2e_steps):
        j = i % 2
        (static_x[j] if graphs is not None else feats[j]).permute(0, 2, 3, 1).copy_(host_x[j], non_blocking=True)
        host_pred.copy_(run_step(i), non_blocking=True)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t) / e2e_steps
    if rank == 0:
        peaks = load_peaks()
        k1_ms, k1_n = prof.get('k1_proj_softmax_pool_fwd', (0.0, 0))
        k1_avg = k1_ms / max(k1_n, 1)
        flops = 2.0 * V * HW * C * L.P
        achieved = flops / (k1_avg * 1e-3) / 1e12 if k1_avg > 0 else 0.0
        long_region = ms_per_step * a.steps > 1000.0
        peak = peaks['bf16_sustained'] if long_region else peaks['bf16_burst']
        line = {'metric': 'inference images/sec (prototype head forward + joint leaf prediction)', 'value': value, 'unit': UNIT,
                'n_gpus': world, 'steps': a.steps, 'warmup': max(a.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
                'config': {'workload': f'{a.workload}: tree {wl["tree"]} ({L.N} nodes, P={L.P}), batch {B}/GPU, single view of '
                                       f'{H}x{H}x{C} bf16 features, inference forward + node classifiers + joint leaf argmax',
                           'global_batch': world * B, 'parallelism': f'dp{world} (independent images, no collective)',
                           'launch': 'one CUDA graph replay per step' if graphs is not None else 'eager',
                           'l2_policy': f'two alternating input batches of {feats[0].numel() * 2 / 1e6:.0f} MB each (> 126 MB L2)'},
                'clocks': clocks,
                'e2e': {'value': world * B / (e2e_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': int(host_x[0].numel() * 2),
                        'd2h_bytes_per_step': int(V * 8), 'steps': e2e_steps, 'ms_per_step': e2e_ms},
                'gpu_launches': int(launches_per_step * a.steps),
                'roofline': {'bound': 'tensor', 'kernel': 'head_pair_kernel<fwd, single view> (projection+softmax+maxpool)',
                             'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak if peak else None,
                             'peak_kind': ('sustained' if long_region else 'burst') + ', ' + peaks['source'],
                             'avg_launch_ms': k1_avg, 'algorithmic_flops_per_launch': flops, 'traffic': None,
                             'kernels': {k: {'ms_per_step': v[0] / 5, 'calls': v[1]} for k, v in prof.items()}},
                'cpu_baseline': None}
        print(json.dumps(line), flush=True)
    sys.stdout.flush()
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    wl = WORKLOADS[a.workload]
    batch = min(8, wl['batch'])
    val, det = oracle_cpu_throughput(wl, batch, a.steps, a.warmup, budget_s=150.0)
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': det['steps'],
            'warmup': a.warmup, 'ms_per_step': det['s_per_step'] * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'{a.workload}: {det["sample"]}'},
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': det['cores'], 'kind': 'port', 'sample': det['sample']},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm
def run_ours(a):
    import torch.distributed as dist
    from pipnet_b200 import _cabi, ops
    from pipnet_b200 import train as tr
    from pipnet_b200.fixtures import build_net, make_args      # product-side builders (identity backbone, seeded weights)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the prototype head has no CPU path')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        if os.environ.get('NCCL_DEBUG', '').upper() in ('VERSION', 'WARN'):
            os.environ.pop('NCCL_DEBUG')                 # NCCL prints its version banner on stdout: keep it to the JSON line
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
        ops.GRAD_ALLREDUCE_GROUP = dist.group.WORLD
    wl = WORKLOADS[a.workload]
    B, C, H = wl['batch'], wl['C'], wl['H']
    V, HW = 2 * B, H * H
    args = make_args(num_features=wl['num_features'], num_protos_per_child=wl.get('per_child', 0))
    if a.recipe == 'shipped':        # + the three extra terms of run_pipnet_20protos_multi_runs_seed42.sh:72-94
        args.tanh_desc, args.minimize_contrasting_set, args.mask_prune_overspecific = 'y|0.05', 'y', 'y|0|1.1'
    use_td = 'y' in args.tanh_desc
    net, root = build_net(wl['tree'], C, args, seed=1)
    net = net.to(dev)
    net.train()
    net.head_precision = a.head_precision
    prec = ops.PREC_FP32X3 if a.head_precision == 'fp32' else ops.PREC_BF16
    L = net.layout
    names = L.node_names
    cls_params = [getattr(net, '_' + n + '_classification').weight for n in names]

    # two alternating feature batches (each 133 MB for cub27 > L2 126 MB), channels-last bf16 like ConvNeXt-26 emits
    g = torch.Generator().manual_seed(1234 + rank)
    feats, labels_h = [], []
    for i in range(2):
        if wl.get('nchw'):      # ResNet contract: NCHW-contiguous (the head's transpose kernel is part of every step)
            feats.append(torch.randn(V, C, H, H, generator=g, dtype=torch.float32).to(torch.bfloat16).to(dev))
        else:
            x = torch.randn(V, H, H, C, generator=g, dtype=torch.float32).to(torch.bfloat16)
            feats.append(x.to(dev).permute(0, 3, 1, 2))           # [V,C,H,W] view with NHWC strides
        y = torch.randint(0, L.L, (B,), generator=torch.Generator().manual_seed(7 + i + 100 * rank))
        labels_h.append(torch.cat([y, y]))
    labels_d = [y.to(dev) for y in labels_h]
    w = tr._phase_weights(False, 1, 10, args)

    def step(x, ys):
        x = x.detach().requires_grad_(True)
        for p in net.parameters():
            p.grad = None
        labels = tr.make_labels(net, ys)
        features, pf, pooled, out = net(x, labels=labels)
        res = tr.calculate_loss(1, net, {}, features, pf, pooled, out, ys, net_normalization_multiplier=net._multiplier,
                                pretrain=False, finetune=False, criterion=None, train_iter=None, print=False, EPS=1e-8,
                                root=root, kernel_orth=True, tanh_desc=use_td, align=False, uni=False, align_pf=True,
                                tanh=True, args=args, device=dev, labels=labels, **w)
        loss = res[0]
        loss.backward()        # N > 1: both flat gradient buffers are mean all-reduced inside backward (ops.GRAD_ALLREDUCE_GROUP)
        return loss, x.grad

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(a.warmup, 3)):
        step(feats[i % 2], labels_d[i % 2])
    barrier()

    # ---------------- eager region: launches per step + eager step time (host-bound: ~20 small launches + autograd)
    launches0 = _cabi.lib().hcomp_launch_count()
    eager_steps = max(3, min(a.steps, 10))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(eager_steps):
        step(feats[i % 2], labels_d[i % 2])
    e1.record()
    barrier()
    eager_ms = e0.elapsed_time(e1) / eager_steps
    launches_per_step = (_cabi.lib().hcomp_launch_count() - launches0) / eager_steps

    # ---------------- the reference-facing API path: what `train_pipnet` runs per step when shapes are static -- the head step
    # (forward, losses, head backward) as ONE graph replay behind a differentiable op on the backbone's feature map
    # (pipnet_b200.train.GraphedHeadTrainStep; includes the copy of the feature map into the graph's static input)
    api_ms = None
    try:
        loss_kwargs = dict(epoch=1, net_normalization_multiplier=net._multiplier, pretrain=False, finetune=False, criterion=None,
                           EPS=1e-8, root=root, kernel_orth=True, tanh_desc=use_td, align=False, uni=False, align_pf=True,
                           tanh=True, args=args, device=dev, **w)
        gh = tr.GraphedHeadTrainStep(net, feats[0].detach().requires_grad_(True), labels_d[0], loss_kwargs)

        def api_step(i):
            for p in net.parameters():
                p.grad = None
            x = feats[i % 2].detach().requires_grad_(True)
            gh(x, labels_d[i % 2]).backward()

        for i in range(3):
            api_step(i)
        barrier()
        e0.record()
        for i in range(eager_steps):
            api_step(i)
        e1.record()
        barrier()
        api_ms = e0.elapsed_time(e1) / eager_steps
        del gh
    except Exception as ex:
        sys.stderr.write(f'API-path graph step unavailable: {ex!r}\n')

    # ---------------- per-kernel durations: CUDA events around the C-ABI calls of the four large kernels in a
    # GPU-saturated loop (host enqueues run ahead of the device, so a bracket holds the kernel and nothing else;
    # inside the host-bound eager step the same brackets would also contain the host's enqueue gaps).  Inputs alternate
    # between the two > L2 feature batches; the kernels, arguments and launch geometry are those of the step.
    dl = net.device_layout(dev)
    with torch.no_grad():
        w_flat_k = net.flat_prototype_kernels().detach().contiguous()
        wp_k, wpc_k = ops.pack_weights(w_flat_k, dl, prec)
        lab_k = [tr.make_labels(net, y) for y in labels_d]
        xr_k = [(ops.feature_rows_split3(f) if prec == ops.PREC_FP32X3 else ops.feature_rows(f)) for f in feats]
        gp_k = torch.randn(V, L.P, device=dev)
        ga_k = torch.full((L.N,), 0.2, device=dev)
        kern_iters = max(5, min(a.steps, 20))

        def kernel_pass(i):
            sp = []
            pooled, argmax, _al = ops.proj_softmax_pool_raw(xr_k[i % 2], wp_k, dl, V, B, HW, net.softmax_tau, lab_k[i % 2],
                                                            precision=prec, spill_out=sp)
            ops.head_backward_raw(xr_k[i % 2], wp_k, wpc_k, dl, V, B, HW, net.softmax_tau, argmax, gp_k, lab_k[i % 2], ga_k,
                                  precision=prec, spill=sp)

        for i in range(3):
            kernel_pass(i)
        barrier()
        ops.PROFILE.reset()
        ops.PROFILE.enabled = True
        saved_group, ops.GRAD_ALLREDUCE_GROUP = ops.GRAD_ALLREDUCE_GROUP, None      # kernels only, no collective here
        for i in range(kern_iters):
            kernel_pass(i)
        ops.GRAD_ALLREDUCE_GROUP = saved_group
        barrier()
        ops.PROFILE.enabled = False
    prof = ops.PROFILE.totals_ms()
    prof_steps = kern_iters

    # ---------------- CUDA-graph capture of the step (one graph per input batch: no copies, L2-cold inputs)
    graphs = None
    if a.graph != 'off':
        try:
            from pipnet_b200.graphs import GraphedHeadStep

            def loss_fn(x, ys):
                labels = tr.make_labels(net, ys)
                features, pf, pooled, out = net(x, labels=labels)
                res = tr.calculate_loss(1, net, {}, features, pf, pooled, out, ys, net_normalization_multiplier=net._multiplier,
                                        pretrain=False, finetune=False, criterion=None, train_iter=None, print=False,
                                        EPS=1e-8, root=root, kernel_orth=True, tanh_desc=use_td, align=False, uni=False,
                                        align_pf=True, tanh=True, args=args, device=dev, labels=labels, **w)
                return res[0]

            graphs = [GraphedHeadStep(loss_fn, list(net.parameters()), feats[i], labels_d[i]) for i in range(2)]
        except Exception as ex:                        # capture is an optimisation of the launch path, not of the math
            if a.graph == 'on':
                raise
            sys.stderr.write(f'graph capture unavailable, timing the eager step: {ex!r}\n')
            graphs = None

    def run_step(i):
        if graphs is not None:
            return graphs[i % 2].replay()
        return step(feats[i % 2], labels_d[i % 2])

    for i in range(max(a.warmup, 3)):
        run_step(i)
    barrier()

    if a.sustained > 0:              # size the timed region from a short probe
        barrier()
        e0.record()
        for i in range(10):
            run_step(i)
        e1.record()
        barrier()
        a.steps = max(a.steps, int(a.sustained * 1e3 / (e0.elapsed_time(e1) / 10)) + 1)

    # ---------------- device-resident timing
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    barrier()
    e0.record()
    for i in range(a.steps):
        run_step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = int(round(launches_per_step * a.steps))
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    ms_per_step = ms / a.steps
    value = world * B / (ms_per_step * 1e-3)

    if os.environ.get('HC_TRACE') and graphs is not None:      # debugging aid (never set by the driver): kernel timeline
        from torch.profiler import profile, ProfilerActivity
        import contextlib
        barrier()
        ctxm = profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) if rank == 0 else contextlib.nullcontext()
        with ctxm as tprof:
            for i in range(2):                 # every rank replays: the step contains collectives
                run_step(i)
            torch.cuda.synchronize()
        barrier()
        if rank == 0:
            evs = sorted([e for e in tprof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
            t0 = evs[0].time_range.start if evs else 0
            os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
            with open(os.path.join(ROOT, 'gpurun_out', 'trace_rank0.txt'), 'w') as f:
                for e in evs:
                    f.write(f'{e.time_range.start - t0:10.1f} {e.time_range.end - e.time_range.start:8.1f}  {e.name[:90]}\n')

    # ---------------- end-to-end from pinned host buffers
    nchw = bool(wl.get('nchw'))
    host_x = [(f.contiguous() if nchw else f.permute(0, 2, 3, 1).contiguous()).cpu().pin_memory() for f in feats]
    host_y = [y.pin_memory() for y in labels_h]
    dev_x = torch.empty((V, C, H, H) if nchw else (V, H, H, C), device=dev, dtype=torch.bfloat16)
    dev_y = torch.empty(V, device=dev, dtype=torch.int64)
    host_loss = torch.empty((), dtype=torch.float32).pin_memory()
    e2e_steps = max(3, min(a.steps, 20))

    # Double-buffered like a prefetching loader (pin_memory + non_blocking): the H2D copy of step i+1 runs on a copy stream
    # into the OTHER graph's static inputs while step i computes; every step's inputs still cross PCIe inside the timed
    # region and every step's loss is read back to the host.
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]

    def e2e_prefetch(i):
        j = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done[j])                  # the previous step that read this buffer has finished
            if graphs is not None:
                (graphs[j].static_x.detach() if nchw else graphs[j].static_x.detach().permute(0, 2, 3, 1)).copy_(host_x[j], non_blocking=True)
                graphs[j].static_y.copy_(host_y[j], non_blocking=True)
            else:
                dev_xs[j].copy_(host_x[j], non_blocking=True)
                dev_ys[j].copy_(host_y[j], non_blocking=True)
            ready[j].record(copy_stream)

    def e2e_step(i, last):
        j = i % 2
        if not last:
            e2e_prefetch(i + 1)
        main_stream.wait_event(ready[j])
        if graphs is not None:
            loss, _ = graphs[j].replay()
        else:
            loss, _ = step(dev_xs[j] if nchw else dev_xs[j].permute(0, 3, 1, 2), dev_ys[j])
        host_loss.copy_(loss.detach(), non_blocking=True)
        done[j].record(main_stream)

    dev_xs = [dev_x, torch.empty_like(dev_x)] if graphs is None else None
    dev_ys = [dev_y, torch.empty_like(dev_y)] if graphs is None else None
    for j in range(2):
        done[j].record(main_stream)
    e2e_prefetch(0)
    for i in range(2):
        e2e_step(i, last=False)
    barrier()
    copy_stream.synchronize()
    for j in range(2):
        done[j].record(main_stream)
    e0.record()
    e2e_prefetch(0)                                          # step 0's copy is inside the timed region too
    for i in range(e2e_steps):
        e2e_step(i, last=(i == e2e_steps - 1))
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t) / e2e_steps
    e2e = {'value': world * B / (e2e_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': int(host_x[0].numel() * 2 + V * 8),
           'd2h_bytes_per_step': 4, 'steps': e2e_steps, 'ms_per_step': e2e_ms,
           'pipeline': 'H2D of step i+1 on a copy stream overlaps the compute of step i (two device input buffers)'}

    # ---------------- N > 1: are the exchanged head gradients the mean of the per-rank gradients?  (one un-timed step)
    dp_check = None
    if world > 1:
        def flat_grads():
            return torch.cat([p.grad.detach().reshape(-1).float() for p in net.parameters() if p.grad is not None])
        step(feats[0], labels_d[0])
        torch.cuda.synchronize()
        got = flat_grads().clone()
        saved_group, ops.GRAD_ALLREDUCE_GROUP = ops.GRAD_ALLREDUCE_GROUP, None
        step(feats[0], labels_d[0])
        ops.GRAD_ALLREDUCE_GROUP = saved_group
        torch.cuda.synchronize()
        want = flat_grads().clone()
        dist.all_reduce(want, op=dist.ReduceOp.SUM)
        want /= world
        err = (got - want).abs().max() / want.abs().max().clamp_min(1e-30)
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        # dW is a split-K fp32 atomic accumulation (order varies run to run) on top of bf16 dZ: 1e-3 of the largest entry
        dp_check = {'max_rel_err': float(err), 'ok': bool(float(err) <= 1e-3), 'n_values': int(got.numel()),
                    'what': 'head gradients after the in-backward exchange vs the all-reduced mean of the per-rank gradients '
                            '(same inputs, exchange disabled), max |diff| / max |mean|, max over ranks',
                    'exchange': 'symmetric-memory multimem all-reduce' if getattr(ops, 'GRAD_EXCHANGE', '') == 'symm' else 'nccl'}

    # ---------------- every line also carries BASELINE.json config 3 (cub190, 32 images per GPU) as a sub-record
    sub = {}
    if a.workload == 'cub27' and not a.no_subrecords:
        try:
            sub['cub190'] = quick_train_record(a, 'cub190', dev, rank, world, steps=max(5, min(a.steps, 20)))
        except Exception as ex:
            sub['cub190'] = {'error': repr(ex)}

    # ---------------- optional: the same step behind a real backbone on synthetic 224x224 images (N = 1 only)
    image_level = None
    if a.with_backbone and world == 1:
        image_level = image_level_throughput(a, wl, dev)

    if rank == 0:
        peaks = load_peaks()
        # roofline of the dominant fused kernel (K1): algorithmic flops = 2 * M * C * P per launch
        M = V * HW
        k1_ms, k1_n = prof.get('k1_proj_softmax_pool_fwd', (0.0, 0))
        k1_avg = k1_ms / max(k1_n, 1)
        flops = 2.0 * M * C * L.P
        achieved = flops / (k1_avg * 1e-3) / 1e12 if k1_avg > 0 else 0.0
        long_region = ms > 1000.0
        traffic = None
        tp = os.path.join(ROOT, 'profiles', 'traffic.json')      # dram bytes of this kernel from the committed ncu --set full capture
        if os.path.isfile(tp):
            traffic = json.load(open(tp)).get(a.workload, {}).get('k1_dram_bytes_per_launch')
        peak = peaks['bf16_sustained'] if long_region else peaks['bf16_burst']
        kernels = {k: {'ms_per_step': v[0] / prof_steps, 'calls': v[1]} for k, v in prof.items()}
        step_flops = 4.0 * flops          # fwd + recompute + dX + dW ; algorithmic = 3 GEMMs
        seg_classes = sorted(set(int(t[0]) for t in L.tiles))
        roofline = {'bound': 'tensor', 'kernel': f'head_pair_kernel<{"/".join(map(str, seg_classes))},fwd> (projection+softmax+maxpool+align)',
                    'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak if peak else None,
                    'peak_kind': ('sustained' if long_region else 'burst') + ', ' + peaks['source'],
                    'avg_launch_ms': k1_avg, 'algorithmic_flops_per_launch': flops, 'traffic': traffic,
                    'method': 'CUDA events around the C-ABI call (the kernel launch alone) in a GPU-saturated loop of the step\'s four '
                              'large kernels, inputs alternating between two >L2 batches',
                    'step_algorithmic_tflops': 3.0 * flops / (ms_per_step * 1e-3) / 1e12,      # dense-equivalent (reference's work)
                    'step_executed_tflops': step_flops / (ms_per_step * 1e-3) / 1e12, 'kernels': kernels}
        cpu = None
        if world == 1 and not a.no_cpu_baseline:
            v, det = oracle_cpu_throughput(wl, min(8, B), steps=3, warmup=1, budget_s=25.0)
            cpu = {'value': v, 'unit': UNIT, 'cores': det['cores'], 'kind': 'port', 'sample': det['sample']}
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': max(a.warmup, 3),
                'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'bf16' if a.head_precision == 'bf16' else 'bf16x3 (fp32-accurate projection: six bf16 cross terms, fp32 accumulate)',
                'data': 'synthetic',
                'config': {'workload': f'{a.workload}: tree {wl["tree"]} ({L.N} nodes), '
                                       + (f'{wl["per_child"]} protos/child' if wl.get('per_child') else f'{wl["num_features"]} protos/node')
                                       + f' (P={L.P}), batch {B}/GPU = {V} views of {H}x{H}x{C} bf16 features, full-training '
                                       f'phase losses (align_pf+tanh+kernel_orth+class'
                                       + ('+tanh_desc+contrasting_set+mask_prune' if a.recipe == 'shipped' else '') + '), fwd+bwd',
                           'global_batch': world * B, 'parallelism': f'dp{world}',
                           'launch': ('one CUDA graph replay per step' if graphs is not None else 'eager'),
                           'backward': ('block-sparse: the (image, node) blocks of dZ without upstream gradient -- exact zeros for '
                                        'hierarchical labels, 68 % of them on cub27 -- are skipped in K5 and the dX / dW GEMMs '
                                        '(data-dependent, same results; HC_SPARSE_BWD=0 runs them densely)'
                                        if ops.SPARSE_BWD else 'dense'),
                           'eager_ms_per_step': eager_ms,
                           'api_path': 'train_pipnet replays the head step as one CUDA graph between backbone forward and backward '
                                       '(GraphedHeadTrainStep, eager fallback on shape / phase change)',
                           'api_ms_per_step': api_ms,
                           'l2_policy': f'two alternating input batches of {feats[0].numel() * 2 / 1e6:.0f} MB each (> 126 MB L2)'},
                'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roofline, 'cpu_baseline': cpu}
        if image_level is not None:
            line['image_level'] = image_level
        if dp_check is not None:
            line['dp_check'] = dp_check
        if sub:
            line['workloads'] = sub
        print(json.dumps(line), flush=True)
    sys.stdout.flush()
    if world > 1:
        # normal teardown (the step's graphs hold the head's own exchange kernel, or NCCL work on the fallback path, so they
        # go first); a watchdog ends the process if the process-group destructor still hangs behind captured NCCL work
        graphs = None
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        watchdog = threading.Timer(20.0, lambda: os._exit(0))
        watchdog.daemon = True
        watchdog.start()
        dist.destroy_process_group()
        watchdog.cancel()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cub27', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-subrecords', action='store_true', help='skip the cub190 sub-record of the default line')
    ap.add_argument('--sustained', type=float, default=0.0,
                    help='run the timed region for at least this many seconds (overrides --steps): clocks under sustained load, '
                         'roofline against the sustained peak')
    ap.add_argument('--with-backbone', action='store_true',
                    help='also report images/s of ConvNeXt-tiny-26 (torchvision, random init, bf16 autocast) + head on 224x224 images')
    ap.add_argument('--graph', default='auto', choices=['auto', 'on', 'off'])
    ap.add_argument('--head-precision', default='bf16', choices=['bf16', 'fp32'],
                    help="fp32: the fp32-accurate projection (3-way bf16 split operands, 6x the MMA work of K1 / K5)")
    ap.add_argument('--recipe', default='core', choices=['core', 'shipped'],
                    help="core: align_pf+tanh+kernel_orth+class (BASELINE.json); shipped: + tanh_desc, contrasting set, mask pruning")
    a = ap.parse_args()
    if a.impl == 'reference':
        run_reference(a)
    elif WORKLOADS[a.workload].get('mode') == 'inference':
        run_inference(a)
    else:
        run_ours(a)


if __name__ == '__main__':
    main()
