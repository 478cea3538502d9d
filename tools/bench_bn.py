"""Micro-benchmark of the HBM-bound BatchNorm kernels (CUDA events, rotating buffer sets > L2 so every launch streams
from HBM): achieved GB/s of algorithmic bytes against the measured copy bandwidth (MEASURED_PEAKS.json)."""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, '.')
from ever_b200._lib import check, lib, ptr, stream  # noqa: E402

c_int, c_ll, c_float = ctypes.c_int, ctypes.c_longlong, ctypes.c_float
SHAPES = [(131072, 256), (131072, 64), (32768, 512), (32768, 128), (8192, 1024), (2048, 2048)]


def timeit(fn, nset, iters=40, warm=5):
    for i in range(warm):
        fn(i % nset)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for i in range(iters):
        fn(i % nset)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


def main():
    L = lib()
    peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'] if os.path.exists('MEASURED_PEAKS.json') else 6550.0
    rows = []
    for m, c in SHAPES:
        one = m * c * 2
        nset = max(2, int(400e6 // (4 * one)) + 1)
        mk = lambda: [torch.randn(m, c, device='cuda').bfloat16() for _ in range(nset)]
        x, dy, y, dx, res = mk(), mk(), mk(), mk(), mk()
        st = torch.rand(4, c, device='cuda') + 0.5
        gamma, beta = torch.ones(c, device='cuda'), torch.zeros(c, device='cuda')
        rm, rv = torch.zeros(c, device='cuda'), torch.ones(c, device='cuda')
        dg, db = torch.empty(c, device='cuda'), torch.empty(c, device='cuda')
        ws = torch.empty(L.evb_bn_workspace(c_ll(m), c_int(c)) // 4, device='cuda')
        s = stream()
        r = dict(M=m, C=c, MB=one / 1e6)

        def apply(i):
            check(L.evb_bn_apply(ptr(x[i]), ptr(st[2]), ptr(st[3]), None, ptr(y[i]), c_ll(m), c_int(c), c_int(1), s), 'a')

        def apply_res(i):
            check(L.evb_bn_apply(ptr(x[i]), ptr(st[2]), ptr(st[3]), ptr(res[i]), ptr(y[i]), c_ll(m), c_int(c), c_int(1), s), 'a')

        def stats(i):
            check(L.evb_bn_stats(ptr(x[i]), c_ll(m), c_int(c), ptr(gamma), ptr(beta), ptr(rm), ptr(rv), c_float(0.1),
                                 c_float(1e-5), ptr(st[0]), ptr(st[1]), ptr(st[2]), ptr(st[3]), ptr(ws), s), 's')

        def bwd2(i):   # mask recomputed from x: reads dy, x twice; writes dx
            check(L.evb_bn_bwd(ptr(dy[i]), ptr(x[i]), None, ptr(st[0]), ptr(st[1]), ptr(st[2]), ptr(st[3]), c_int(2), c_int(0),
                               ptr(dx[i]), None, c_int(0), ptr(dg), ptr(db), c_int(0), c_ll(m), c_int(c), ptr(ws), s), 'b')

        def bwd1(i):   # residual block: mask from y, dres written
            check(L.evb_bn_bwd(ptr(dy[i]), ptr(x[i]), ptr(y[i]), ptr(st[0]), ptr(st[1]), ptr(st[2]), ptr(st[3]), c_int(1),
                               c_int(0), ptr(dx[i]), ptr(res[i]), c_int(0), ptr(dg), ptr(db), c_int(0), c_ll(m), c_int(c),
                               ptr(ws), s), 'b')

        def copy(i):
            y[i].copy_(x[i])
        for vec, bps in ((8, 2), (4, 2), (0, 3), (0, 2)):
            L.evb_set_bn_variant(c_int(2 if vec == 0 else 1))
            L.evb_set_bn_vec(c_int(vec))
            L.evb_set_bn_reduce_blocks(c_int(bps))
            for name, fn, units in (('copy', copy, 2), ('apply', apply, 2), ('apply_res', apply_res, 3), ('stats', stats, 1),
                                    ('bwd_mask2', bwd2, 5), ('bwd_mask1_dres', bwd1, 8)):
                if name in ('copy', 'stats') and (vec, bps) != (8, 2):
                    continue
                if bps == 2 and vec == 0:
                    continue
                us = timeit(fn, nset)
                name = name + ('_v%db%d' % (vec, bps) if name not in ('copy', 'stats') else '')
                r[name + '_us'] = round(us, 2)
                r[name + '_gbs'] = round(units * one / us / 1e3, 1)
        L.evb_set_bn_vec(c_int(0))
        L.evb_set_bn_variant(c_int(2))
        L.evb_set_bn_reduce_blocks(c_int(4))
        rows.append(r)
        print(json.dumps(r))
        del x, dy, y, dx, res
        torch.cuda.empty_cache()
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(rows, open('gpurun_out/bench_bn.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
