"""Run a few launches of one conv problem (for `ncu --set full`)."""
import sys
sys.path.insert(0, '.')
import torch
from ever_b200 import ops
n, h, w, cin, cout, k, s = [int(v) for v in sys.argv[1:8]]
mode = sys.argv[8] if len(sys.argv) > 8 else 'fwd'
x = torch.randn(n, h, w, cin, device='cuda').bfloat16()
wt = torch.randn(cout, cin, k, k, device='cuda') * 0.05
wf, wb = ops.pack_conv_weight_torch(wt)
dy = torch.randn(n, h // s, w // s, cout, device='cuda').bfloat16()
for _ in range(3):
    if mode == 'fwd':
        ops.conv2d_fwd(x, wf, k, s, cout)
    elif mode == 'dgrad':
        ops.conv2d_dgrad(dy, wb, k, s, cin)
    else:
        ops.conv2d_wgrad(x, dy, k, s)
torch.cuda.synchronize()
