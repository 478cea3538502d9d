"""A few launches of the BatchNorm backward / apply kernels on one layer shape (for `ncu --set full`)."""
import ctypes
import sys
sys.path.insert(0, '.')
import torch
from ever_b200._lib import check, lib, ptr, stream
c_int, c_ll, c_float = ctypes.c_int, ctypes.c_longlong, ctypes.c_float
m, c = int(sys.argv[1]), int(sys.argv[2])
L = lib()
x, dy, y, dx = [torch.randn(m, c, device='cuda').bfloat16() for _ in range(4)]
st = torch.rand(4, c, device='cuda') + 0.5
dg, db = torch.empty(c, device='cuda'), torch.empty(c, device='cuda')
ws = torch.empty(L.evb_bn_workspace(c_ll(m), c_int(c)) // 4, device='cuda')
for _ in range(2):
    check(L.evb_bn_apply(ptr(x), ptr(st[2]), ptr(st[3]), None, ptr(y), c_ll(m), c_int(c), c_int(1), stream()), 'a')
    check(L.evb_bn_bwd(ptr(dy), ptr(x), None, ptr(st[0]), ptr(st[1]), ptr(st[2]), ptr(st[3]), c_int(2), c_int(0), ptr(dx), None,
                       c_int(0), ptr(dg), ptr(db), c_int(0), c_ll(m), c_int(c), ptr(ws), stream()), 'b')
torch.cuda.synchronize()
