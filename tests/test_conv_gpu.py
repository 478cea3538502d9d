"""GPU parity of the tcgen05 implicit-GEMM convolution against torch's conv (the reference's arithmetic:
nn.Conv2d -> cuDNN), fp32 math on the same bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref_conv(x_nhwc, w, stride, bias=None):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    x = x_nhwc.float().permute(0, 3, 1, 2)
    y = F.conv2d(x, w.to(torch.bfloat16).float(), bias, stride=stride, padding=w.shape[-1] // 2)
    return y.permute(0, 2, 3, 1).contiguous()


def _close(got, ref, tol=1.2e-2):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err <= tol * scale, (err, scale)


CASES = [
    # n, h, w, cin, cout, k, stride
    (2, 16, 16, 64, 64, 1, 1),
    (2, 16, 16, 64, 64, 3, 1),
    (1, 32, 32, 128, 256, 3, 1),
    (2, 16, 16, 256, 128, 1, 1),
    (2, 32, 32, 64, 128, 3, 2),
    (2, 32, 32, 128, 256, 1, 2),
    (1, 8, 8, 512, 512, 3, 1),
    (3, 24, 40, 64, 192, 3, 1),   # ragged tiles
    (1, 128, 128, 256, 256, 3, 1),
]


@pytest.mark.parametrize('case', CASES)
@pytest.mark.parametrize('nt', [0, 64, 128, 256])
def test_conv_fwd(case, nt):
    from ever_b200 import ops
    n, h, w, cin, cout, k, s = case
    if nt > max(cout, 64):
        pytest.skip('tile wider than Cout')
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.randn(n, h, w, cin, device='cuda', generator=g).to(torch.bfloat16)
    wt = torch.randn(cout, cin, k, k, device='cuda', generator=g) * (2.0 / (cin * k * k)) ** 0.5
    bias = torch.randn(cout, device='cuda', generator=g)
    wf, _ = ops.pack_conv_weight_torch(wt)
    y = ops.conv2d_fwd(x, wf, k, s, cout, bias=bias, force_nt=nt)
    torch.cuda.synchronize()
    _close(y, _ref_conv(x, wt, s, bias))


@pytest.mark.parametrize('case', CASES)
def test_conv_dgrad(case):
    from ever_b200 import ops
    n, h, w, cin, cout, k, s = case
    g = torch.Generator(device='cuda').manual_seed(2)
    ho, wo = h // s, w // s
    dy = torch.randn(n, ho, wo, cout, device='cuda', generator=g).to(torch.bfloat16)
    wt = torch.randn(cout, cin, k, k, device='cuda', generator=g) * (2.0 / (cout * k * k)) ** 0.5
    _, wb = ops.pack_conv_weight_torch(wt)
    dx = ops.conv2d_dgrad(dy, wb, k, s, cin)
    torch.cuda.synchronize()
    torch.backends.cudnn.allow_tf32 = False
    ref = torch.nn.grad.conv2d_input((n, cin, h, w), wt.to(torch.bfloat16).float(), dy.float().permute(0, 3, 1, 2),
                                     stride=s, padding=k // 2).permute(0, 2, 3, 1)
    _close(dx, ref)
    # accumulate mode adds into an existing bf16 tensor
    base = torch.randn(n, h, w, cin, device='cuda', generator=g).to(torch.bfloat16)
    acc = base.clone()
    ops.conv2d_dgrad(dy, wb, k, s, cin, out=acc, accumulate=True)
    torch.cuda.synchronize()
    _close(acc, ref + base.float(), tol=2e-2)


def test_conv_fwd_fpn_add():
    from ever_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(3)
    x = torch.randn(2, 32, 32, 128, device='cuda', generator=g).to(torch.bfloat16)
    top = torch.randn(2, 16, 16, 256, device='cuda', generator=g).to(torch.bfloat16)
    wt = torch.randn(256, 128, 1, 1, device='cuda', generator=g) * 0.1
    wf, _ = ops.pack_conv_weight_torch(wt)
    y = ops.conv2d_fwd(x, wf, 1, 1, 256, add=top, add_mode=2)
    torch.cuda.synchronize()
    lat = _ref_conv(x, wt, 1).to(torch.bfloat16)
    up = F.interpolate(top.float().permute(0, 3, 1, 2), scale_factor=2, mode='nearest').permute(0, 2, 3, 1).to(torch.bfloat16)
    _close(y, (lat + up).float(), tol=1e-2)


@pytest.mark.parametrize('case', CASES + [(8, 16, 16, 192, 64, 1, 1)])
@pytest.mark.parametrize('split', [0, 1, 3])
def test_conv_wgrad(case, split):
    from ever_b200 import ops
    n, h, w, cin, cout, k, s = case
    g = torch.Generator(device='cuda').manual_seed(4)
    ho, wo = h // s, w // s
    x = torch.randn(n, h, w, cin, device='cuda', generator=g).to(torch.bfloat16)
    dy = torch.randn(n, ho, wo, cout, device='cuda', generator=g).to(torch.bfloat16)
    dw = ops.conv2d_wgrad(x, dy, k, s, force_split=split)
    torch.cuda.synchronize()
    torch.backends.cudnn.allow_tf32 = False
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (cout, cin, k, k), dy.float().permute(0, 3, 1, 2),
                                      stride=s, padding=k // 2)
    _close(dw, ref, tol=2e-3)
    dw2 = dw.clone()
    ops.conv2d_wgrad(x, dy, k, s, dw=dw2, accumulate=True, force_split=split)
    torch.cuda.synchronize()
    _close(dw2, 2 * ref, tol=2e-3)


@pytest.mark.parametrize('case', [(2, 16, 16, 64, 64, 1, 1), (1, 128, 128, 256, 256, 3, 1), (8, 16, 16, 256, 2048, 1, 1),
                                  (3, 24, 40, 64, 192, 3, 1), (2, 32, 32, 128, 512, 1, 2), (8, 64, 64, 128, 128, 3, 1)])
def test_conv_fwd_fused_bn_stats(case):
    """evb_conv2d_fwd_stats + evb_bn_finalize == conv followed by BatchNorm batch statistics of its bf16 output."""
    import ctypes
    from ever_b200 import ops
    from ever_b200._lib import check, lib, ptr, stream
    n, h, w, cin, cout, k, s = case
    L = lib()
    c_int, c_ll, c_float = ctypes.c_int, ctypes.c_longlong, ctypes.c_float
    g = torch.Generator(device='cuda').manual_seed(6)
    x = (torch.randn(n, h, w, cin, device='cuda', generator=g) + 0.3).to(torch.bfloat16)
    wt = torch.randn(cout, cin, k, k, device='cuda', generator=g) * (2.0 / (cin * k * k)) ** 0.5
    wf, _ = ops.pack_conv_weight_torch(wt)
    ho, wo = h // s, w // s
    y = torch.empty(n, ho, wo, cout, device='cuda', dtype=torch.bfloat16)
    partial = torch.empty(2 * cout * 320, device='cuda')
    nblk = c_int(0)
    check(L.evb_conv2d_fwd_stats(ptr(x), c_int(n), c_int(h), c_int(w), c_int(cin), ptr(wf), c_int(cout), c_int(k), c_int(s),
                                 ptr(y), c_int(cout), ptr(partial), ctypes.byref(nblk), stream()), 'fwd_stats')
    gamma = 1 + 0.1 * torch.randn(cout, device='cuda', generator=g)
    beta = 0.1 * torch.randn(cout, device='cuda', generator=g)
    rm, rv = torch.zeros(cout, device='cuda'), torch.ones(cout, device='cuda')
    st = torch.empty(4, cout, device='cuda')
    m_rows = n * ho * wo
    check(L.evb_bn_finalize(ptr(partial), nblk, c_ll(m_rows), c_int(cout), ptr(gamma), ptr(beta), ptr(rm), ptr(rv),
                            c_float(0.1), c_float(1e-5), ptr(st[0]), ptr(st[1]), ptr(st[2]), ptr(st[3]), stream()), 'fin')
    torch.cuda.synchronize()
    _close(y, _ref_conv(x, wt, s))
    yf = y.float().reshape(-1, cout)
    mean, var = yf.mean(0), yf.var(0, unbiased=False)
    assert float((st[0] - mean).abs().max()) <= 1e-4 * float(yf.abs().max())
    rstd = (var + 1e-5).rsqrt()
    assert float(((st[1] - rstd) / rstd).abs().max()) < 1e-3
    torch.testing.assert_close(rm, 0.1 * mean, rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(rv, 0.9 + 0.1 * yf.var(0, unbiased=True), rtol=1e-3, atol=1e-5)
