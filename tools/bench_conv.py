"""Micro-benchmark of the conv kernels on the full FarSeg-R50 problem list (N=8, 512^2), CUDA-graph
replay of back-to-back launches (no host launch overhead; small problems stay L2-warm as they are inside the step).
Prints per-problem fwd / dgrad / wgrad time, the roofline time max(bytes/HBM, flop/tensor peak) and the step-weighted
totals (calls x time), sorted by where the time goes."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, '.')
from ever_b200 import ops  # noqa: E402

# n, h, w, cin, cout, k, s, calls per forward
SHAPES = [
    (8, 128, 128, 256, 256, 3, 1, 2), (8, 64, 64, 256, 256, 3, 1, 4), (8, 32, 32, 256, 256, 3, 1, 8),
    (8, 128, 128, 256, 256, 1, 1, 3), (8, 128, 128, 64, 64, 3, 1, 3), (8, 64, 64, 128, 128, 3, 1, 3),
    (8, 32, 32, 256, 1024, 1, 1, 6), (8, 32, 32, 1024, 256, 1, 1, 6), (8, 16, 16, 512, 512, 3, 1, 2),
    (8, 128, 128, 64, 256, 1, 1, 4), (8, 64, 64, 128, 512, 1, 1, 4), (8, 64, 64, 512, 256, 1, 1, 2),
    (8, 64, 64, 512, 128, 1, 1, 3), (8, 16, 16, 512, 2048, 1, 1, 3), (8, 128, 128, 128, 128, 3, 2, 1),
    (8, 64, 64, 256, 256, 3, 2, 1), (8, 32, 32, 512, 512, 3, 2, 1), (8, 128, 128, 256, 64, 1, 1, 2),
    (8, 128, 128, 256, 128, 1, 1, 1), (8, 128, 128, 256, 512, 1, 2, 1), (8, 64, 64, 512, 1024, 1, 2, 1),
    (8, 32, 32, 1024, 512, 1, 1, 1), (8, 32, 32, 1024, 2048, 1, 2, 1), (8, 16, 16, 2048, 512, 1, 1, 2),
    (8, 64, 64, 256, 256, 1, 1, 2), (8, 16, 16, 256, 256, 3, 1, 2), (8, 16, 16, 2048, 256, 1, 1, 1),
    (8, 32, 32, 256, 256, 1, 1, 2), (8, 128, 128, 64, 64, 1, 1, 1), (8, 128, 128, 256, 64, 1, 1, 0),
    (8, 256, 256, 192, 64, 1, 1, 1),   # stem as a GEMM over im2col rows
]
HBM, TF = 6547e9, 1618e12


def timeit(fn, iters=10):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
        with torch.cuda.graph(g):
            for _ in range(iters):
                fn()
    torch.cuda.current_stream().wait_stream(s)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * iters)


def main():
    rows = []
    for (n, h, w, cin, cout, k, s, calls) in SHAPES:
        if calls == 0:
            continue
        x = torch.randn(n, h, w, cin, device='cuda').bfloat16()
        wt = torch.randn(cout, cin, k, k, device='cuda') * 0.05
        wf, wb = ops.pack_conv_weight_torch(wt)
        ho, wo = h // s, w // s
        dy = torch.randn(n, ho, wo, cout, device='cuda').bfloat16()
        y = torch.empty(n, ho, wo, cout, device='cuda', dtype=torch.bfloat16)
        dx = torch.empty(n, h, w, cin, device='cuda', dtype=torch.bfloat16)
        dw = torch.empty(cout, cin, k, k, device='cuda')
        ws = torch.empty(64 << 20, device='cuda')
        flop = 2.0 * n * ho * wo * cout * cin * k * k
        by = (n * h * w * cin + n * ho * wo * cout) * 2 + cin * cout * k * k * 2
        roof = max(by / HBM, flop / TF) * 1e3
        t_f = timeit(lambda: ops.conv2d_fwd(x, wf, k, s, cout, out=y))
        t_d = timeit(lambda: ops.conv2d_dgrad(dy, wb, k, s, cin, out=dx))
        t_w = timeit(lambda: ops.conv2d_wgrad(x, dy, k, s, dw=dw, ws=ws))
        row = dict(shape=(n, h, w, cin, cout, k, s), calls=calls, gflop=round(flop / 1e9, 1), roof_ms=round(roof, 4),
                   fwd_ms=round(t_f, 4), dgrad_ms=round(t_d, 4), wgrad_ms=round(t_w, 4),
                   step_ms=round(calls * (t_f + t_d + t_w), 4), step_roof_ms=round(calls * 3 * roof, 4))
        rows.append(row)
    rows.sort(key=lambda r: -r['step_ms'])
    for r in rows:
        print(json.dumps(r))
    tot = {k: round(sum(r['calls'] * r[k] for r in rows), 3) for k in ('fwd_ms', 'dgrad_ms', 'wgrad_ms', 'roof_ms')}
    print(json.dumps(dict(total_per_step=tot)))
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(dict(rows=rows, total=tot), open('gpurun_out/bench_conv.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
