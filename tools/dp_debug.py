import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from oracle.problems import bf16_round, build_net, make_args
from pipnet_b200 import dist as hd, ops, train as tr
rank = int(os.environ['RANK']); torch.cuda.set_device(rank); dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
args = make_args(num_features=20, tanh_desc='y|0.05', minimize_contrasting_set='y', mask_prune_overspecific='y|0|1.1')
net, root = build_net('cub27', 64, args, seed=3); net = net.to(dev)
B, H = 6, 6
g = torch.Generator().manual_seed(100 + rank)
x = bf16_round(torch.randn(2 * B, 64, H, H, generator=g)).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
ys = torch.randint(0, net.layout.L, (B,), generator=g); ys = torch.cat([ys, ys]).to(dev)
gum = (-torch.empty(net.layout.n_welems, 2).exponential_(generator=g).log()).to(dev)
w = tr._phase_weights(False, 1, 10, args)
params = dict(net.named_parameters())
def loss_fn(xs, y):
    labels = tr.make_labels(net, y)
    f, pf, pooled, out = net(xs, labels=labels)
    return tr.calculate_loss(1, net, {}, f, pf, pooled, out, y, net_normalization_multiplier=net._multiplier, pretrain=False, finetune=False,
                             criterion=None, train_iter=None, print=False, EPS=1e-8, root=root, kernel_orth=True, tanh_desc=True, align=False,
                             uni=False, align_pf=True, tanh=True, args=args, device=dev, labels=labels, gumbel_noise=gum, **w)[0]
def run():
    for p in params.values(): p.grad = None
    loss_fn(x.detach().requires_grad_(True), ys).backward()
    torch.cuda.synchronize()
    return {k: p.grad.detach().clone() for k, p in params.items() if p.grad is not None}
local = run(); local2 = run()
want = {}
for k, v in local.items():
    t = v.clone(); dist.all_reduce(t, op=dist.ReduceOp.AVG); want[k] = t
rep = max(float((local[k] - local2[k]).abs().max()) for k in local)
for fresh in (True, False):
    hd.enable_overlapped_allreduce(fresh_grads=fresh)
    got = run()
    nbad = [(k, float((got[k] - want[k]).abs().max()), float((got[k] - local[k]).abs().max()), float(want[k].abs().max())) for k in want
            if float((got[k] - want[k]).abs().max()) > 1e-5 * float(want[k].abs().max()) + 1e-8]
    if rank == 0:
        print(f'fresh={fresh}: repeatability {rep:.2e}; {len(nbad)} of {len(want)} off', flush=True)
        for r in nbad[:6]: print('   key %s  |got-want| %.3e  |got-local| %.3e  max|want| %.3e' % r, flush=True)
hd.disable_overlapped_allreduce()
torch.cuda.synchronize(); dist.barrier(); os._exit(0)
