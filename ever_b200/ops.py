"""Thin Python wrappers over the C ABI (include/evb200.h).  Tensors are torch CUDA tensors used purely as
device memory; all arithmetic happens in libevb200.so on torch's current stream."""
import ctypes

import torch

from ._lib import check, lib, ptr, stream

c_int = ctypes.c_int


def pack_conv_weight_torch(w):
    """OIHW fp32 -> ([kh*kw][Cout][Cin], [kh*kw][Cin][Cout]) bf16 packs (test helper; the engine uses the
    evb_pack_weights kernel)."""
    co, ci, kh, kw = w.shape
    f = w.permute(2, 3, 0, 1).reshape(kh * kw, co, ci).to(torch.bfloat16).contiguous()
    b = w.permute(2, 3, 1, 0).reshape(kh * kw, ci, co).to(torch.bfloat16).contiguous()
    return f, b


def conv2d_fwd(x, wpk, ksize, stride, cout, bias=None, add=None, add_mode=0, out=None, force_nt=0):
    n, h, w, cin = x.shape
    ho, wo = h // stride, w // stride
    if out is None:
        out = torch.empty((n, ho, wo, cout), dtype=torch.bfloat16, device=x.device)
    check(lib().evb_conv2d_fwd(ptr(x), c_int(n), c_int(h), c_int(w), c_int(cin), ptr(wpk), c_int(wpk.shape[1]),
                               c_int(ksize), c_int(stride), ptr(out), c_int(cout), ptr(bias), ptr(add),
                               c_int(add_mode), c_int(force_nt), stream()), 'evb_conv2d_fwd')
    return out


def conv2d_dgrad(dy, wpk_t, ksize, stride, cin, out=None, accumulate=False, force_nt=0):
    n, ho, wo, cout = dy.shape
    h, w = ho * stride, wo * stride
    if out is None:
        assert not accumulate
        out = torch.empty((n, h, w, cin), dtype=torch.bfloat16, device=dy.device)
    check(lib().evb_conv2d_dgrad(ptr(dy), c_int(n), c_int(ho), c_int(wo), c_int(cout), ptr(wpk_t),
                                 c_int(wpk_t.shape[1]), c_int(ksize), c_int(stride), ptr(out), c_int(h), c_int(w),
                                 c_int(cin), c_int(1 if accumulate else 0), c_int(force_nt), stream()),
          'evb_conv2d_dgrad')
    return out


def conv2d_wgrad(x, dy, ksize, stride, dw=None, accumulate=False, ws=None, force_nt=0, force_split=0):
    n, h, w, cin = x.shape
    cout = dy.shape[3]
    need = lib().evb_conv2d_wgrad_workspace
    need.restype = ctypes.c_longlong
    nbytes = need(c_int(n), c_int(h // stride), c_int(w // stride), c_int(cin), c_int(cout), c_int(ksize),
                  c_int(force_nt), c_int(force_split))
    if ws is None or ws.numel() * ws.element_size() < nbytes:
        ws = torch.empty(nbytes // 4, dtype=torch.float32, device=x.device)
    if dw is None:
        assert not accumulate
        dw = torch.empty((cout, cin, ksize, ksize), dtype=torch.float32, device=x.device)
    check(lib().evb_conv2d_wgrad(ptr(x), c_int(n), c_int(h), c_int(w), c_int(cin), ptr(dy), c_int(cout), c_int(ksize),
                                 c_int(stride), ptr(dw), c_int(1 if accumulate else 0), ptr(ws),
                                 ctypes.c_longlong(ws.numel() * ws.element_size()), c_int(force_nt),
                                 c_int(force_split), stream()), 'evb_conv2d_wgrad')
    return dw
