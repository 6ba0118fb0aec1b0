import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from pipnet_b200 import ops
import test_gpu_fullsize as T
ops_, L, dl, x, w, labels, (V, C, H) = T._setup("cub27-b64", seed=5)
HW = H * H
g = torch.Generator(device='cuda').manual_seed(9)
G1 = torch.randn(V, L.P, generator=g, device='cuda')
ni = L.N // 3
p0, p1 = int(L.proto_off[ni]), int(L.proto_off[ni + 1])
Gonly = torch.zeros_like(G1); Gonly[:, p0:p1] = G1[:, p0:p1]
res = {}
for sparse in (True, False):
    ops.SPARSE_BWD = sparse
    xr = x.detach().clone().requires_grad_(True); wr = w.detach().clone().requires_grad_(True)
    pooled, align, argmax, _o = ops.HeadProjPool.apply(xr, wr, dl, V // 2, 1.0, labels, 0.0)
    slot = pooled._hc_prep
    ((pooled * Gonly).sum() + 0.0 * align.sum()).backward()
    torch.cuda.synchronize()
    res[sparse] = (xr.grad.float().permute(0, 2, 3, 1).reshape(V * HW, C).clone(), wr.grad.clone(), slot.blocks)
dxs, dws, blk = res[True]; dxd, dwd, _ = res[False]
print('pcol of node', dl.pcol[p0:p1].tolist())
print('dX max diff', float((dxs - dxd).abs().max()), 'ref max', float(dxd.abs().max()), 'dW rel diff', float((dws - dwd).abs().max() / dwd.abs().max()))
M = V * HW
rowdiff = (dxs - dxd).abs().amax(dim=1)
bad = torch.nonzero(rowdiff > 0).flatten()
print('bad rows', len(bad), bad[:12].tolist(), 'tiles', sorted(set((bad // 256).tolist()))[:20])
coldiff = (dxs - dxd).abs().amax(dim=0)
print('bad channel tiles', sorted(set((torch.nonzero(coldiff > 0).flatten() // 256).tolist())))
n_r256 = (M + 255) // 256
t1 = blk.buf[:n_r256 * blk.struct.ld1].view(n_r256, blk.struct.ld1)
print('t1 col sums', t1.sum(0).tolist(), 'rows', n_r256)
bt = sorted(set((bad // 256).tolist()))
print('marks of bad tiles', [t1[t].tolist() for t in bt[:6]])
nz = torch.nonzero(dxd.abs().amax(dim=1) > 0).flatten()
print('nonzero dense rows', len(nz), 'dense nonzero tiles', len(set((nz // 256).tolist())))
