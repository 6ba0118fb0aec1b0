"""NCCL all-reduce latency for the head's gradient buffers (run under torchrun on N GPUs)."""
import os, torch, torch.distributed as dist
rank = int(os.environ['RANK']); torch.cuda.set_device(rank); dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
def bench(n, graph):
    t = torch.zeros(n, device=dev)
    for _ in range(5): dist.all_reduce(t, op=dist.ReduceOp.AVG)
    torch.cuda.synchronize(); dist.barrier()
    iters = 50
    if graph:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            dist.all_reduce(t, op=dist.ReduceOp.AVG)
        torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
        with torch.cuda.graph(g):
            dist.all_reduce(t, op=dist.ReduceOp.AVG)
        for _ in range(3): g.replay()
        torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay() if graph else dist.all_reduce(t, op=dist.ReduceOp.AVG)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for n in (1000, 384000, 386000, 393216, 1 << 20):
    a, b = bench(n, False), bench(n, True)
    if rank == 0: print(f'{n:8d} floats ({n*4/1e6:.2f} MB): eager {a:7.1f} us   graph {b:7.1f} us', flush=True)
torch.cuda.synchronize(); dist.barrier(); os._exit(0)
