"""Micro-benchmark of the conv kernels vs torch/cuDNN bf16 (channels_last) on the FarSeg-R50 problem list."""
import json
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, '.')
from ever_b200 import ops  # noqa: E402

SHAPES = [  # n, h, w, cin, cout, k, s
    (8, 128, 128, 256, 256, 3, 1),
    (8, 64, 64, 256, 256, 3, 1),
    (8, 32, 32, 256, 256, 3, 1),
    (8, 16, 16, 512, 512, 3, 1),
    (8, 128, 128, 64, 64, 3, 1),
    (8, 128, 128, 256, 256, 1, 1),
    (8, 128, 128, 64, 256, 1, 1),
    (8, 128, 128, 256, 64, 1, 1),
    (8, 32, 32, 1024, 256, 1, 1),
    (8, 32, 32, 256, 1024, 1, 1),
    (8, 16, 16, 2048, 512, 1, 1),
    (8, 64, 64, 256, 256, 3, 2),
]


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    rows = []
    for (n, h, w, cin, cout, k, s) in SHAPES:
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
        t_f = timeit(lambda: ops.conv2d_fwd(x, wf, k, s, cout, out=y))
        t_d = timeit(lambda: ops.conv2d_dgrad(dy, wb, k, s, cin, out=dx))
        t_w = timeit(lambda: ops.conv2d_wgrad(x, dy, k, s, dw=dw, ws=ws))
        xc = x.permute(0, 3, 1, 2)  # channels_last view
        wc = wt.bfloat16().contiguous(memory_format=torch.channels_last)
        t_c = timeit(lambda: F.conv2d(xc, wc, None, s, k // 2))
        xn = xc.contiguous()
        wn = wt.bfloat16()
        t_n = timeit(lambda: F.conv2d(xn, wn, None, s, k // 2))
        row = dict(shape=(n, h, w, cin, cout, k, s), gflop=flop / 1e9, fwd_ms=t_f, dgrad_ms=t_d, wgrad_ms=t_w,
                   cudnn_cl_ms=t_c, cudnn_nchw_ms=t_n, fwd_tflops=flop / t_f / 1e9, dgrad_tflops=flop / t_d / 1e9,
                   wgrad_tflops=flop / t_w / 1e9, cudnn_cl_tflops=flop / t_c / 1e9)
        rows.append(row)
        print(json.dumps(row))
    json.dump(rows, open('gpurun_out/bench_conv.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
