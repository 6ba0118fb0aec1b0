"""Timing experiments on K1 (fused projection+softmax+pool): python tools/k1_exp.py [path/to/libhcomp_head.so]
Times the forward kernel alone in a GPU-saturated loop for the 1-CTA and the cta_group::2 variants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pipnet_b200 import _cabi
if len(sys.argv) > 1:
    _cabi.LIB_PATH = os.path.abspath(sys.argv[1])
from pipnet_b200 import ops, trees, layout, train as tr
from oracle.problems import Problem

dev = torch.device('cuda:0')
pb = Problem('cub27', 768, 26, 64, seed=1, num_features=20)
dl = ops.DeviceLayout(pb.layout, dev)
V, B, HW = pb.V, pb.V_first, 26 * 26
xr = ops.feature_rows(pb.features(dev))
wp, _wpc = ops.pack_weights(pb.w_flat(dev).contiguous(), dl, ops.PREC_BF16)
lab = ops.LabelTables(pb.ys.to(dev), dl, B)
for mode in (0, 1):
    _cabi.lib().hcomp_set_cta_pair(mode)
    for use_lab in (lab, None):
        for _ in range(3):
            ops.proj_softmax_pool_raw(xr, wp, dl, V, B, HW, 1.0, use_lab)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 30
        e0.record()
        for _ in range(n):
            ops.proj_softmax_pool_raw(xr, wp, dl, V, B, HW, 1.0, use_lab)
        e1.record(); torch.cuda.synchronize()
        print(f'cta_pair={mode} align={"on" if use_lab is not None else "off"}: K1 call {e0.elapsed_time(e1) / n * 1e3:.1f} us (incl. memsets + unpack)')
