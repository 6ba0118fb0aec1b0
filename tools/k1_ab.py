"""A/B timing of the fused forward (K1) and backward-recompute (K5) kernels for several builds of the library:
    python tools/k1_ab.py lib_a.so lib_b.so ...        (each build runs in its own process)
Prints kernel-only times (CUDA events around the C-ABI call in a back-to-back loop, inputs alternating between two
> L2 batches) and a checksum of pooled / argmax / dZ so that variants can be compared bit for bit."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(path, workload):
    import ctypes as C
    import hashlib
    import torch
    from pipnet_b200 import _cabi
    if path != 'default':
        _cabi.LIB_PATH = os.path.abspath(path)
    from pipnet_b200 import ops
    from oracle.problems import Problem
    if os.environ.get('HC_AB_NOFOLD'):
        _cabi.lib().hcomp_set_rider_fold(0)
    tree, nf, batch = {'cub27': ('cub27', 20, 64), 'cub190': ('synth190', 20, 32)}[workload]
    dev = torch.device('cuda:0')
    pb = Problem(tree, 768, 26, batch, seed=1, num_features=nf)
    dl = ops.DeviceLayout(pb.layout, dev)
    V, B, HW = pb.V, pb.V_first, 26 * 26
    xs = [ops.feature_rows(pb.features(dev)), None]
    xs[1] = (xs[0].float().roll(1, 0) * 0.97).to(torch.bfloat16)
    wp, wpc = ops.pack_weights(pb.w_flat(dev).contiguous(), dl, ops.PREC_BF16)
    lab = ops.LabelTables(pb.ys.to(dev), dl, B)
    gp = torch.randn(V, dl.P, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
    ga = torch.full((dl.N,), 0.2, device=dev)
    skip_k5 = bool(os.environ.get('HC_AB_SKIP_K5'))      # ablation builds whose backward kernel is not runnable
    sp0 = []
    pooled, argmax, align = ops.proj_softmax_pool_raw(xs[0], wp, dl, V, B, HW, 1.0, lab, spill_out=sp0)
    if skip_k5:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(3):
            ops.proj_softmax_pool_raw(xs[it % 2], wp, dl, V, B, HW, 1.0, lab)
        ops.PROFILE.enabled = True
        ops.PROFILE.reset()
        for it in range(20):
            ops.proj_softmax_pool_raw(xs[it % 2], wp, dl, V, B, HW, 1.0, lab)
        torch.cuda.synchronize()
        k1 = ops.PROFILE.totals_ms()['k1_proj_softmax_pool_fwd'][0] / 20 * 1e3
        print(f'{os.path.basename(path):14s} {workload:7s} K1 {k1:7.1f} us   (K1 only)', flush=True)
        return
    _, _, dz = ops.head_backward_raw(xs[0], wp, wpc, dl, V, B, HW, 1.0, argmax, gp, lab, ga, pooled=pooled, need_dx=False,
                                     need_dw=False, spill=sp0)
    torch.cuda.synchronize()
    h = hashlib.sha1()
    for t in (pooled, argmax, align, dz):
        h.update(t.detach().cpu().contiguous().view(torch.uint8).numpy().tobytes())
    ops.PROFILE.enabled = True
    n = 20
    for it in range(3 + n):
        if it == 3:
            torch.cuda.synchronize()
            ops.PROFILE.reset()
        sp_ = []
        p_, a_, _ = ops.proj_softmax_pool_raw(xs[it % 2], wp, dl, V, B, HW, 1.0, lab, spill_out=sp_)
        ops.head_backward_raw(xs[it % 2], wp, wpc, dl, V, B, HW, 1.0, a_, gp, lab, ga, pooled=p_, need_dx=False, need_dw=False,
                              spill=sp_)
    torch.cuda.synchronize()
    tot = ops.PROFILE.totals_ms()
    k1 = tot['k1_proj_softmax_pool_fwd'][0] / n * 1e3
    k5 = tot['k5_bwd_dz'][0] / n * 1e3
    print(f'{os.path.basename(path):14s} {workload:7s} K1 {k1:7.1f} us   K5(+scat) {k5:7.1f} us   sha1 {h.hexdigest()[:12]}', flush=True)
    L = _cabi.lib()
    if hasattr(L, 'hcomp_debug_pair_counters'):          # library built with -DHC_EXP_TIMING: per-role cycle counters
        ops.PROFILE.enabled = False
        buf = (C.c_ulonglong * 16)()

        def report(tag):
            L.hcomp_debug_pair_counters(buf)
            c = list(buf)
            pk, mk, it = max(c[2], 1), max(c[6], 1), max(c[9], 1)
            print(f'   {tag}: producer/k-block: wait(empty) {c[0] / pk:7.0f}  issue {c[1] / pk:5.0f} | MMA/k-block: wait(full) {c[3] / mk:7.0f}  '
                  f'issue {c[5] / mk:5.0f}  wait(tmem_empty)/item {c[4] / max(c[6] // 12, 1):7.0f} | epilogue warp/item: wait(tmem_full) '
                  f'{c[7] / it:7.0f}  work {c[8] / it:7.0f} | SM clock {c[10] / max(c[11], 1):.3f} GHz', flush=True)

        def stamps(tag, nblk):
            sb = (C.c_ulonglong * 640)()
            L.hcomp_debug_pair_stamps(sb)
            import numpy as np
            a = np.array(list(sb), dtype=np.int64).reshape(160, 4)[:nblk]
            t0 = a[:, 0].min()
            a = (a - t0) / 1e3
            print(f'   {tag} stamps (us, {nblk} CTAs): entry {a[:,0].min():.1f}..{a[:,0].max():.1f} | setup done {a[:,1].min():.1f}..{a[:,1].max():.1f} '
                  f'| epilogue loop end {a[:,2].min():.1f}..{a[:,2].max():.1f} (median {np.median(a[:,2]):.1f}) | exit {a[:,3].min():.1f}..{a[:,3].max():.1f}',
                  flush=True)

        def trace(tag):
            if not hasattr(L, 'hcomp_debug_pair_trace'):
                return
            tb = (C.c_ulonglong * 512)()
            L.hcomp_debug_pair_trace(tb)
            import numpy as np
            a = np.array(list(tb), dtype=np.int64).reshape(4, 16, 8)
            names = ['mma:wait_te', 'mma:te_ok', 'mma:k0_full', 'mma:issued', 'epi:wait_tf', 'epi:tf_ok', 'epi:release', 'epi:done']
            for slot, cta in enumerate((0, 1, 72, 147)):
                t = a[slot]
                nz = t[t > 0]
                if nz.size == 0:
                    continue
                t0 = nz.min()
                print(f'   {tag} trace CTA {cta} (us since its first event; ' + ' '.join(names) + ')')
                for i in range(16):
                    if (t[i] > 0).any():
                        print('      item %2d: ' % i + ' '.join('%7.2f' % ((x - t0) / 1e3) if x > 0 else '      -' for x in t[i]))

        def wtrace(tag):
            if not hasattr(L, 'hcomp_debug_pair_wtrace'):
                return
            tb = (C.c_ulonglong * 1536)()
            L.hcomp_debug_pair_wtrace(tb)
            import numpy as np
            a8 = np.array(list(tb), dtype=np.int64).reshape(12, 16, 8)
            a = a8[:, :, :3]
            nz = a[a > 0]
            if nz.size == 0:
                return
            t0 = nz.min()
            print(f'   {tag} CTA 0, all epilogue warps (us): rows = items, per warp [accumulators seen / released / done]; warp w: quadrant w%4, part w//4')
            for i in range(16):
                if (a[:, i] > 0).any():
                    print('      item %2d: ' % i + ' | '.join('%5.1f %5.1f %5.1f' % tuple((x - t0) / 1e3 for x in a[w, i]) for w in range(12)))
            print(f'   {tag} CTA 0, first segment of warps 2 and 3 (us after the accumulators were seen): loaded / softmax / pooled view 1 / pooled view 2')
            for i in range(16):
                if (a8[2, i] > 0).any():
                    print('      item %2d: ' % i + ' | '.join(' '.join('%5.2f' % ((a8[w, i, e] - a8[w, i, 0]) / 1e3) for e in (3, 4, 5, 6)) for w in (2, 3)))

        L.hcomp_debug_pair_counters(buf)
        trace('warm')
        wtrace('warm')
        for it in range(6):
            ops.proj_softmax_pool_raw(xs[it % 2], wp, dl, V, B, HW, 1.0, lab)
        report('K1')
        stamps('K1', 148)
        trace('K1')
        wtrace('K1')
        for it in range(6):
            ops.head_backward_raw(xs[it % 2], wp, wpc, dl, V, B, HW, 1.0, argmax, gp, lab, ga, pooled=pooled, need_dx=False,
                                  need_dw=False, spill=sp0)
        report('K5')
        stamps('K5', 148)
        trace('K5')


if __name__ == '__main__':
    if len(sys.argv) >= 3 and sys.argv[1] == '--one':
        one(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else 'cub27')
    else:
        wls = os.environ.get('HC_AB_WORKLOADS', 'cub27').split(',')
        for path in sys.argv[1:] or ['default']:
            for wl in wls:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), '--one', path, wl], capture_output=True, text=True)
                sys.stdout.write(r.stdout)
                if r.returncode != 0:
                    sys.stdout.write(f'{path} {wl}: FAILED\n{r.stderr[-1500:]}\n')
                sys.stdout.flush()
