"""Op-level GPU parity: every HBM-bound kernel of libevb200.so against the torch op the reference calls
(fp32 torch math on the same bf16-rounded inputs).  Called through the C ABI via ctypes."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
c_int, c_ll, c_float = ctypes.c_int, ctypes.c_longlong, ctypes.c_float


def _L():
    from ever_b200._lib import check, lib, ptr, stream
    return lib(), check, ptr, stream


def _rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-12))


def _gen(seed=0):
    return torch.Generator(device='cuda').manual_seed(seed)


def nhwc(t):  # NCHW fp32 -> NHWC
    return t.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize('shape', [(2, 16, 16, 64), (3, 8, 12, 256), (2, 4, 4, 2048), (1, 32, 32, 128)])
@pytest.mark.parametrize('mode', ['relu', 'plain', 'res_relu'])
def test_bn_train_fwd_bwd(shape, mode):
    L, check, ptr, stream = _L()
    n, h, w, c = shape
    g = _gen(1)
    x = (torch.randn(n, h, w, c, device='cuda', generator=g) * 2 + 0.5).bfloat16()
    res = torch.randn(n, h, w, c, device='cuda', generator=g).bfloat16() if mode == 'res_relu' else None
    gamma = 1 + 0.1 * torch.randn(c, device='cuda', generator=g)
    beta = 0.1 * torch.randn(c, device='cuda', generator=g)
    rm, rv = torch.zeros(c, device='cuda'), torch.ones(c, device='cuda')
    dy = torch.randn(n, h, w, c, device='cuda', generator=g).bfloat16()
    relu = mode != 'plain'
    m_rows = n * h * w
    ws = torch.empty(L.evb_bn_workspace(c_ll(m_rows), c_int(c)) // 4, device='cuda')
    st = torch.empty(4, c, device='cuda')
    y = torch.empty_like(x)
    check(L.evb_bn_stats(ptr(x), c_ll(m_rows), c_int(c), ptr(gamma), ptr(beta), ptr(rm), ptr(rv), c_float(0.1),
                         c_float(1e-5), ptr(st[0]), ptr(st[1]), ptr(st[2]), ptr(st[3]), ptr(ws), stream()), 'stats')
    check(L.evb_bn_apply(ptr(x), ptr(st[2]), ptr(st[3]), ptr(res), ptr(y), c_ll(m_rows), c_int(c), c_int(int(relu)),
                         stream()), 'apply')
    dx = torch.empty_like(x)
    dres = torch.empty_like(x) if res is not None else None
    dgamma, dbeta = torch.empty(c, device='cuda'), torch.empty(c, device='cuda')
    check(L.evb_bn_bwd(ptr(dy), ptr(x), ptr(y if relu else None), ptr(st[0]), ptr(st[1]), ptr(st[2]), ptr(st[3]),
                       c_int(1 if relu else 0), c_int(0), ptr(dx), ptr(dres), c_int(0), ptr(dgamma), ptr(dbeta), c_int(0),
                       c_ll(m_rows), c_int(c), ptr(ws), stream()), 'bwd')
    torch.cuda.synchronize()
    # torch reference (fp32)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm2, rv2 = torch.zeros(c, device='cuda'), torch.ones(c, device='cuda')
    yr = F.batch_norm(xr, rm2, rv2, gr, br, True, 0.1, 1e-5)
    if res is not None:
        # the reference adds two bf16 tensors: BN output is rounded before the add (straight-through here)
        yr = yr + (yr.bfloat16().float() - yr).detach()
        rr = res.float().permute(0, 3, 1, 2).requires_grad_(True)
        yr = yr + rr
    if relu:
        yr = F.relu(yr)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    assert _rel(y.float(), nhwc(yr.detach())) < 6e-3
    assert _rel(rm, rm2) < 1e-4 and _rel(rv, rv2) < 1e-4
    assert _rel(dx.float(), nhwc(xr.grad)) < 1.5e-2
    assert _rel(dgamma, gr.grad) < 1.5e-2 and _rel(dbeta, br.grad) < 1.5e-2
    if res is not None:
        assert _rel(dres.float(), nhwc(rr.grad)) < 1.5e-2
        # bit-mask variant (block tail): same output, mask bits == (y > 0), and a backward that reads the bits instead of y
        # gives the same dx / dres / dgamma / dbeta bit for bit
        y2, bits = torch.empty_like(x), torch.empty(m_rows * c // 8, dtype=torch.int32, device='cuda')
        check(L.evb_bn_apply_mask(ptr(x), ptr(st[2]), ptr(st[3]), ptr(res), ptr(y2), ptr(bits), c_ll(m_rows), c_int(c),
                                  stream()), 'apply_mask')
        dx2, dres2 = torch.empty_like(x), torch.empty_like(x)
        dgamma2, dbeta2 = torch.empty(c, device='cuda'), torch.empty(c, device='cuda')
        check(L.evb_bn_bwd(ptr(dy), ptr(x), ptr(bits), ptr(st[0]), ptr(st[1]), ptr(st[2]), ptr(st[3]), c_int(3), c_int(0),
                           ptr(dx2), ptr(dres2), c_int(0), ptr(dgamma2), ptr(dbeta2), c_int(0), c_ll(m_rows), c_int(c),
                           ptr(ws), stream()), 'bwd_bits')
        torch.cuda.synchronize()
        assert torch.equal(y2, y)
        want = ((y.view(-1, 8) > 0).int() << torch.arange(8, device='cuda', dtype=torch.int32)).sum(dim=1, dtype=torch.int32)
        assert torch.equal(bits, want)
        assert torch.equal(dx2, dx) and torch.equal(dres2, dres)
        assert torch.equal(dgamma2, dgamma) and torch.equal(dbeta2, dbeta)


def test_maxpool():
    L, check, ptr, stream = _L()
    g = _gen(2)
    x = F.relu(torch.randn(2, 32, 48, 64, device='cuda', generator=g)).bfloat16()
    y = torch.empty(2, 16, 24, 64, device='cuda', dtype=torch.bfloat16)
    idx = torch.empty(2, 16, 24, 64, device='cuda', dtype=torch.uint8)
    check(L.evb_maxpool3x3s2_fwd(ptr(x), ptr(y), ptr(idx), c_int(2), c_int(32), c_int(48), c_int(64), stream()), 'mp')
    dy = torch.randn(2, 16, 24, 64, device='cuda', generator=g).bfloat16()
    dx = torch.empty_like(x)
    check(L.evb_maxpool3x3s2_bwd(ptr(dy), ptr(idx), ptr(dx), c_int(2), c_int(32), c_int(48), c_int(64), stream()), 'mpb')
    torch.cuda.synchronize()
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    yr = F.max_pool2d(xr, 3, 2, 1)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    assert torch.equal(y.float(), nhwc(yr.detach()))
    # ties (zeros after ReLU) may route to a different element; compare where the window max is unique via sums
    assert _rel(dx.float().sum(dim=(1, 2)), nhwc(xr.grad).sum(dim=(1, 2))) < 1e-2
    assert _rel(dx.float(), nhwc(xr.grad)) < 0.05


@pytest.mark.parametrize('hw', [(12, 20), (128, 128), (1, 3)])
@pytest.mark.parametrize('f,c,ldx,ldy,bn', [(2, 128, 128, 128, True), (2, 64, 64, 64, False), (4, 16, 64, 16, False)])
def test_bilinear(f, c, ldx, ldy, bn, hw):
    L, check, ptr, stream = _L()
    g = _gen(3)
    n, (h, w) = 2, hw
    x = torch.randn(n, h, w, ldx, device='cuda', generator=g).bfloat16()
    scale = (1 + 0.1 * torch.randn(c, device='cuda', generator=g)) if bn else None
    shift = (0.1 * torch.randn(c, device='cuda', generator=g)) if bn else None
    y = torch.zeros(n, h * f, w * f, ldy, device='cuda', dtype=torch.bfloat16)
    check(L.evb_bilinear_up(ptr(x), ptr(scale), ptr(shift), ptr(y), c_int(n), c_int(h), c_int(w), c_int(c), c_int(ldx),
                            c_int(ldy), c_int(f), stream()), 'up')
    dy = torch.randn(n, h * f, w * f, ldy, device='cuda', generator=g).bfloat16()
    dx = torch.zeros(n, h, w, ldx, device='cuda', dtype=torch.bfloat16)
    check(L.evb_bilinear_up_bwd(ptr(dy), ptr(dx), c_int(n), c_int(h), c_int(w), c_int(c), c_int(ldy), c_int(ldx), c_int(f),
                                stream()), 'upb')
    torch.cuda.synchronize()
    xin = x[..., :c].float()
    if bn:
        xin = F.relu((xin * scale + shift).bfloat16().float())
    xr = xin.permute(0, 3, 1, 2).clone().requires_grad_(True)
    yr = F.interpolate(xr, scale_factor=f, mode='bilinear', align_corners=True)
    yr.backward(dy[..., :c].float().permute(0, 3, 1, 2))
    assert _rel(y[..., :c].float(), nhwc(yr.detach())) < 4e-3
    assert _rel(dx[..., :c].float(), nhwc(xr.grad)) < 4e-3
    # separable two-pass backward
    dx2 = torch.zeros_like(dx)
    ws = torch.empty(L.evb_bilinear_up_bwd_workspace(c_int(n), c_int(h), c_int(w), c_int(c), c_int(f)) // 4, device='cuda')
    check(L.evb_bilinear_up_bwd_sep(ptr(dy), ptr(dx2), c_int(n), c_int(h), c_int(w), c_int(c), c_int(ldy), c_int(ldx), c_int(f),
                                    ptr(ws), c_ll(ws.numel() * 4), stream()), 'upb_sep')
    torch.cuda.synchronize()
    assert _rel(dx2[..., :c].float(), nhwc(xr.grad)) < 4e-3


def test_sumpool_merge_gap():
    L, check, ptr, stream = _L()
    g = _gen(4)
    fine = torch.randn(2, 16, 24, 64, device='cuda', generator=g).bfloat16()
    coarse = torch.randn(2, 8, 12, 64, device='cuda', generator=g).bfloat16()
    ref = coarse.float() + F.avg_pool2d(fine.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1) * 4
    check(L.evb_sumpool2(ptr(fine), ptr(coarse), c_int(2), c_int(8), c_int(12), c_int(64), c_int(1), stream()), 'sp')
    torch.cuda.synchronize()
    assert _rel(coarse.float(), ref) < 4e-3
    ts = [torch.randn(2, 8, 8, 128, device='cuda', generator=g).bfloat16() for _ in range(4)]
    out = torch.empty_like(ts[0])
    check(L.evb_merge4(ptr(ts[0]), ptr(ts[1]), ptr(ts[2]), ptr(ts[3]), ptr(out), c_ll(out.numel()), stream()), 'merge')
    torch.cuda.synchronize()
    assert torch.equal(out, sum(ts) / 4)  # same bf16 rounding sequence as python sum() of bf16 tensors
    x = torch.randn(3, 4, 4, 512, device='cuda', generator=g).bfloat16()
    sc = torch.empty(3, 512, device='cuda')
    check(L.evb_gap_fwd(ptr(x), ptr(sc), c_int(3), c_int(16), c_int(512), stream()), 'gap')
    torch.cuda.synchronize()
    assert _rel(sc, x.float().mean(dim=(1, 2))) < 4e-3
    dsc = torch.randn(3, 512, device='cuda', generator=g)
    dx = torch.zeros_like(x)
    check(L.evb_gap_bwd(ptr(dsc), ptr(dx), c_int(3), c_int(16), c_int(512), stream()), 'gapb')
    torch.cuda.synchronize()
    assert _rel(dx.float(), (dsc / 16)[:, None, None, :].expand(3, 4, 4, 512)) < 4e-3


def test_relation_fwd_bwd():
    L, check, ptr, stream = _L()
    g = _gen(5)
    n, h, w, c = 2, 8, 8, 256
    m_rows = n * h * w
    u1 = torch.randn(n, h, w, c, device='cuda', generator=g).bfloat16()
    u2 = torch.randn(n, h, w, c, device='cuda', generator=g).bfloat16()
    s1, b1, s2, b2 = [(1 + 0.1 * torch.randn(c, device='cuda', generator=g)) if i % 2 == 0 else
                      0.2 * torch.randn(c, device='cuda', generator=g) for i in range(4)]
    sf = (0.1 * torch.randn(n, c, device='cuda', generator=g)).bfloat16().float()
    z = torch.empty(n, h, w, c, device='cuda', dtype=torch.bfloat16)
    rel = torch.empty(m_rows, device='cuda')
    check(L.evb_relation_fwd(ptr(u1), ptr(u2), ptr(s1), ptr(b1), ptr(s2), ptr(b2), ptr(sf), ptr(z), ptr(rel),
                             c_ll(m_rows), c_int(h * w), c_int(c), stream()), 'rel')
    dz = torch.randn(n, h, w, c, device='cuda', generator=g).bfloat16()
    g1, g2 = torch.empty_like(u1), torch.empty_like(u2)
    dsf = torch.full((n, c), 7.0, device='cuda')   # overwritten, not accumulated
    ws = torch.empty(L.evb_relation_bwd_workspace(c_ll(m_rows), c_int(h * w), c_int(c)) // 4, device='cuda')
    check(L.evb_relation_bwd(ptr(dz), ptr(u1), ptr(u2), ptr(s1), ptr(b1), ptr(s2), ptr(b2), ptr(sf), ptr(rel), ptr(g1),
                             ptr(g2), ptr(dsf), c_ll(m_rows), c_int(h * w), c_int(c), ptr(ws), stream()), 'relb')
    torch.cuda.synchronize()
    # reference: fs_relation.py:57-73 on BN-folded inputs
    a1 = (u1.float() * s1 + b1).requires_grad_(True)
    a2 = (u2.float() * s2 + b2).requires_grad_(True)
    sfr = sf.clone().requires_grad_(True)
    cf, pf = F.relu(a1), F.relu(a2)
    r = torch.sigmoid((sfr[:, None, None, :] * cf).sum(dim=3, keepdim=True))
    zr = r * pf
    zr.backward(dz.float())
    assert _rel(z.float(), zr.detach()) < 6e-3
    assert _rel(g1.float(), a1.grad) < 1.5e-2
    assert _rel(g2.float(), a2.grad) < 1.5e-2
    assert _rel(dsf, sfr.grad) < 1.5e-2


def test_linear():
    L, check, ptr, stream = _L()
    g = _gen(6)
    n, i, o = 4, 512, 256
    x = torch.randn(n, i, device='cuda', generator=g).bfloat16().float()
    W = (0.05 * torch.randn(o, i, device='cuda', generator=g))
    b = 0.1 * torch.randn(o, device='cuda', generator=g)
    y = torch.empty(n, o, device='cuda')
    check(L.evb_linear_fwd(ptr(x), ptr(W), ptr(b), ptr(y), c_int(n), c_int(i), c_int(o), c_int(1), stream()), 'lin')
    dy = torch.randn(n, o, device='cuda', generator=g)
    dW, db, dx = torch.empty_like(W), torch.empty_like(b), torch.empty_like(x)
    check(L.evb_linear_bwd(ptr(dy), ptr(y), ptr(x), ptr(W), ptr(dW), ptr(db), ptr(dx), c_int(n), c_int(i), c_int(o), c_int(1),
                           c_int(0), c_int(0), stream()), 'linb')
    torch.cuda.synchronize()
    Wr = W.bfloat16().float().requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    yr = F.relu(xr @ Wr.t() + br)
    yr.backward(dy)
    assert _rel(y, yr.detach()) < 6e-3
    assert _rel(dW, Wr.grad) < 1e-2 and _rel(db, br.grad) < 1e-2 and _rel(dx, xr.grad) < 1e-2


@pytest.mark.parametrize('k', [5, 15])
def test_loss(k):
    L, check, ptr, stream = _L()
    g = _gen(7)
    n, h, w = 2, 32, 48
    npx = n * h * w
    logits = torch.zeros(n, h, w, 16, device='cuda', dtype=torch.bfloat16)
    logits[..., :k] = torch.randn(n, h, w, k, device='cuda', generator=g).bfloat16()
    labels = torch.randint(0, k, (n, h, w), device='cuda', generator=g)
    labels[torch.rand(n, h, w, device='cuda', generator=g) < 0.1] = 255
    stats = torch.empty(2 + 3 * k, device='cuda')
    ws = torch.empty(L.evb_loss_workspace(c_ll(npx), c_int(k)) // 4, device='cuda')
    losses, coef = torch.empty(2, device='cuda'), torch.empty(1 + 2 * k, device='cuda')
    dl = torch.empty_like(logits)
    check(L.evb_loss_stats(ptr(logits), ptr(labels), c_ll(npx), c_int(k), c_int(16), c_int(255), ptr(stats), ptr(ws),
                           stream()), 'ls')
    check(L.evb_loss_finalize(ptr(stats), None, c_int(k), c_float(1.0), c_float(1.0), c_float(1.0), c_float(1.0),
                              ptr(losses), ptr(coef), stream()), 'lf')
    check(L.evb_loss_grad(ptr(logits), ptr(labels), c_ll(npx), c_int(k), c_int(16), c_int(255), ptr(coef), ptr(dl),
                          stream()), 'lg')
    torch.cuda.synchronize()
    from oracle.farseg_oracle import dice_loss_oracle
    lr = logits[..., :k].float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    ce = F.cross_entropy(lr, labels, ignore_index=255)
    dice = dice_loss_oracle(lr, labels)
    (ce + dice).backward()
    assert abs(float(losses[0]) - float(ce)) < 1e-4 * abs(float(ce))
    assert abs(float(losses[1]) - float(dice)) < 1e-4 * abs(float(dice))
    assert _rel(dl[..., :k].float(), nhwc(lr.grad)) < 8e-3
    assert float(dl[..., k:].abs().max()) == 0.0


@pytest.mark.parametrize('mode', ['none_ignored', 'all_ignored', 'class_absent', 'ragged'])
def test_loss_edge_cases(mode):
    """edge cases of CE(ignore_index=255) + Dice (ever/module/loss.py:26-75) with the reference's own behaviour as the
    expected value: no ignored pixel; EVERY pixel ignored (F.cross_entropy is 0/0 = nan there while its gradient is zero, the
    Dice term sees empty selections and is 0); a class that never occurs; a pixel count that is not a multiple of the
    kernel's block"""
    L, check, ptr, stream = _L()
    g = _gen(17)
    k = 5
    n, h, w = (1, 37, 53) if mode == 'ragged' else (2, 32, 48)
    npx = n * h * w
    logits = torch.zeros(n, h, w, 16, device='cuda', dtype=torch.bfloat16)
    logits[..., :k] = torch.randn(n, h, w, k, device='cuda', generator=g).bfloat16()
    labels = torch.randint(0, k - 1 if mode == 'class_absent' else k, (n, h, w), device='cuda', generator=g)
    if mode == 'all_ignored':
        labels.fill_(255)
    elif mode != 'none_ignored':
        labels[torch.rand(n, h, w, device='cuda', generator=g) < 0.2] = 255
    stats = torch.empty(2 + 3 * k, device='cuda')
    ws = torch.empty(L.evb_loss_workspace(c_ll(npx), c_int(k)) // 4, device='cuda')
    losses, coef = torch.empty(2, device='cuda'), torch.empty(1 + 2 * k, device='cuda')
    dl = torch.empty_like(logits)
    check(L.evb_loss_stats(ptr(logits), ptr(labels), c_ll(npx), c_int(k), c_int(16), c_int(255), ptr(stats), ptr(ws),
                           stream()), 'ls')
    check(L.evb_loss_finalize(ptr(stats), None, c_int(k), c_float(1.0), c_float(1.0), c_float(1.0), c_float(1.0),
                              ptr(losses), ptr(coef), stream()), 'lf')
    check(L.evb_loss_grad(ptr(logits), ptr(labels), c_ll(npx), c_int(k), c_int(16), c_int(255), ptr(coef), ptr(dl),
                          stream()), 'lg')
    torch.cuda.synchronize()
    from oracle.farseg_oracle import dice_loss_oracle
    lr = logits[..., :k].float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    ce = F.cross_entropy(lr, labels, ignore_index=255)
    dice = dice_loss_oracle(lr, labels)
    (ce + dice).backward()
    if mode == 'all_ignored':
        assert torch.isnan(ce) and torch.isnan(losses[0])           # the reference's 0 / 0
        assert float(dice) == 0.0 and float(losses[1]) == 0.0
        assert float(lr.grad.abs().max()) == 0.0 and float(dl.float().abs().max()) == 0.0   # and a ZERO gradient
        return
    assert abs(float(losses[0]) - float(ce)) < 1e-4 * abs(float(ce))
    assert abs(float(losses[1]) - float(dice)) < 1e-4 * abs(float(dice))
    assert _rel(dl[..., :k].float(), nhwc(lr.grad)) < 8e-3
    assert float(dl[..., k:].abs().max()) == 0.0


def test_pack_im2col_sgd():
    L, check, ptr, stream = _L()
    g = _gen(8)
    w = torch.randn(96, 64, 3, 3, device='cuda', generator=g)
    wf = torch.empty(9, 128, 64, device='cuda', dtype=torch.bfloat16)
    wb = torch.empty(9, 64, 128, device='cuda', dtype=torch.bfloat16)
    check(L.evb_pack_weight(ptr(w), c_int(96), c_int(64), c_int(9), ptr(wf), c_int(128), c_int(64), ptr(wb), c_int(64),
                            c_int(128), stream()), 'pack')
    x = torch.randn(2, 3, 32, 48, device='cuda', generator=g)
    a = torch.empty(2, 16, 24, 192, device='cuda', dtype=torch.bfloat16)
    check(L.evb_stem_im2col(ptr(x), ptr(a), c_int(2), c_int(3), c_int(32), c_int(48), c_int(192), stream()), 'im2col')
    torch.cuda.synchronize()
    ref_f = w.permute(2, 3, 0, 1).reshape(9, 96, 64).bfloat16()
    assert torch.equal(wf[:, :96], ref_f) and float(wf[:, 96:].float().abs().max()) == 0
    assert torch.equal(wb[:, :, :96], w.permute(2, 3, 1, 0).reshape(9, 64, 96).bfloat16())
    unf = F.unfold(x, 7, padding=3, stride=2)  # [N, 147, L], k = c*49 + r*7 + s
    assert torch.equal(a[..., :147].reshape(2, -1, 147), unf.permute(0, 2, 1).bfloat16())
    assert float(a[..., 147:].float().abs().max()) == 0
    # fused clip + SGD == clip_grad_norm_ + torch.optim.SGD (ever/interface/module.py:83-108)
    nparam = 100003
    wgt = torch.randn(nparam, device='cuda', generator=g)
    grad = torch.randn(nparam, device='cuda', generator=g)
    p = torch.nn.Parameter(wgt.clone())
    opt = torch.optim.SGD([p], lr=0.01, momentum=0.9, weight_decay=1e-4)
    mom = torch.zeros(nparam, device='cuda')
    mine_w, lr_t = wgt.clone(), torch.tensor([0.01], device='cuda')
    ws = torch.empty(L.evb_sgd_workspace(c_ll(nparam)) // 4 + 4, device='cuda')
    norm = torch.empty(2, device='cuda')
    for step in range(3):
        gstep = grad * (step + 1)
        p.grad = gstep.clone()
        tn = torch.nn.utils.clip_grad_norm_([p], max_norm=35, norm_type=2)
        opt.step()
        gm = gstep.clone()
        check(L.evb_grad_norm(ptr(gm), c_ll(nparam), c_float(35.0), ptr(norm), ptr(ws), stream()), 'norm')
        check(L.evb_sgd_step(ptr(mine_w), ptr(gm), ptr(mom), c_ll(nparam), ptr(lr_t), c_float(0.9), c_float(1e-4),
                             ptr(norm), c_int(1 if step == 0 else 0), c_int(1), stream()), 'sgd')
        torch.cuda.synchronize()
        assert abs(float(norm[0]) - float(tn)) < 1e-4 * float(tn)
        assert _rel(mine_w, p.data) < 1e-6
        assert float(gm.abs().max()) == 0


def test_loss_binary():
    """K == 1: masked BCE-with-logits + sigmoid Dice (ever/module/loss.py:66-68, 229-235) and their logit gradient."""
    L, check, ptr, stream = _L()
    g = _gen(9)
    n, h, w = 2, 32, 48
    npx = n * h * w
    logits = torch.zeros(n, h, w, 16, device='cuda', dtype=torch.bfloat16)
    logits[..., 0] = torch.randn(n, h, w, device='cuda', generator=g).bfloat16()
    labels = torch.randint(0, 2, (n, h, w), device='cuda', generator=g)
    labels[torch.rand(n, h, w, device='cuda', generator=g) < 0.1] = 255
    stats = torch.empty(5, device='cuda')
    ws = torch.empty(L.evb_loss_workspace(c_ll(npx), c_int(1)) // 4 + 64, device='cuda')
    losses, coef = torch.empty(2, device='cuda'), torch.empty(3, device='cuda')
    dl = torch.empty_like(logits)
    check(L.evb_loss_stats(ptr(logits), ptr(labels), c_ll(npx), c_int(1), c_int(16), c_int(255), ptr(stats), ptr(ws),
                           stream()), 'ls')
    check(L.evb_loss_finalize(ptr(stats), None, c_int(1), c_float(1.0), c_float(1.0), c_float(1.0), c_float(1.0),
                              ptr(losses), ptr(coef), stream()), 'lf')
    check(L.evb_loss_grad(ptr(logits), ptr(labels), c_ll(npx), c_int(1), c_int(16), c_int(255), ptr(coef), ptr(dl),
                          stream()), 'lg')
    torch.cuda.synchronize()
    from oracle.farseg_oracle import bce_loss_oracle, dice_loss_oracle
    lr = logits[..., :1].float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    bce = bce_loss_oracle(lr, labels)
    dice = dice_loss_oracle(lr, labels)
    (bce + dice).backward()
    assert abs(float(losses[0]) - float(bce)) < 1e-4 * abs(float(bce))
    assert abs(float(losses[1]) - float(dice)) < 1e-4 * abs(float(dice))
    assert _rel(dl[..., :1].float(), nhwc(lr.grad)) < 8e-3
    assert float(dl[..., 1:].abs().max()) == 0.0


@pytest.mark.parametrize('cin,ks,stride,pad,hw', [(3, 7, 2, 3, (96, 304)), (3, 3, 2, 1, (64, 260)), (8, 7, 2, 3, (32, 130)),
                                                   (3, 7, 2, 3, (512, 512)), (16, 7, 2, 3, (32, 64))])
def test_im2col_windows_bit_exact(cin, ks, stride, pad, hw):
    """im2col of the stem windows (7x7 s2 p3, deep-stem 3x3 s2 p1) from float NCHW and from uint8 HWC pixels == F.unfold of
    the (normalised) image, bit for bit: ragged widths, full-size tiles, Cin = 3 / 8 / 16"""
    L, check, ptr, stream = _L()
    g = _gen(31)
    n, (h, w) = 2, hw
    k = cin * ks * ks
    kp = (k + 63) // 64 * 64
    x = torch.randn(n, cin, h, w, device='cuda', generator=g)
    a = torch.full((n, h // stride, w // stride, kp), 7.0, device='cuda', dtype=torch.bfloat16)
    check(L.evb_im2col_nchw(ptr(x), ptr(a), c_int(n), c_int(cin), c_int(h), c_int(w), c_int(kp), c_int(ks), c_int(stride),
                            c_int(pad), stream()), 'im2col')
    torch.cuda.synchronize()
    unf = F.unfold(x, ks, padding=pad, stride=stride).permute(0, 2, 1).bfloat16()
    assert torch.equal(a[..., :k].reshape(n, -1, k), unf)
    assert float(a[..., k:].abs().max()) == 0.0 if kp > k else True
    img = torch.randint(0, 256, (n, h, w, cin), device='cuda', generator=g, dtype=torch.uint8)
    mean = torch.linspace(90.0, 130.0, cin, device='cuda')
    std = torch.linspace(50.0, 60.0, cin, device='cuda')
    a.fill_(7.0)
    check(L.evb_im2col_u8(ptr(img), ptr(mean), ptr(std), ptr(a), c_int(n), c_int(cin), c_int(h), c_int(w), c_int(kp),
                          c_int(ks), c_int(stride), c_int(pad), stream()), 'im2col_u8')
    torch.cuda.synchronize()
    xf = img.permute(0, 3, 1, 2).float().sub(mean.view(1, cin, 1, 1)).div(std.view(1, cin, 1, 1))
    unf = F.unfold(xf, ks, padding=pad, stride=stride).permute(0, 2, 1).bfloat16()
    assert torch.equal(a[..., :k].reshape(n, -1, k), unf)


def test_u8_input_pipeline_and_confusion_matrix():
    L, check, ptr, stream = _L()
    g = _gen(10)
    n, h, w = 2, 32, 48
    img = torch.randint(0, 256, (n, h, w, 3), device='cuda', generator=g, dtype=torch.uint8)
    mean = torch.tensor([123.675, 116.28, 103.53], device='cuda')
    std = torch.tensor([58.395, 57.12, 57.375], device='cuda')
    a = torch.empty(n, h // 2, w // 2, 192, device='cuda', dtype=torch.bfloat16)
    check(L.evb_stem_im2col_u8(ptr(img), ptr(mean), ptr(std), ptr(a), c_int(n), c_int(3), c_int(h), c_int(w), c_int(192),
                               stream()), 'im2col_u8')
    # reference pipeline: HWC uint8 -> CHW float -> th_mean_std_normalize -> (autocast) bf16 conv input
    xf = img.permute(0, 3, 1, 2).float().sub(mean.view(1, 3, 1, 1)).div(std.view(1, 3, 1, 1))
    unf = F.unfold(xf, 7, padding=3, stride=2)
    torch.cuda.synchronize()
    assert torch.equal(a[..., :147].reshape(n, -1, 147), unf.permute(0, 2, 1).bfloat16())
    k = 7
    pred = torch.randint(0, k, (n, h, w), device='cuda', generator=g, dtype=torch.uint8)
    lab = torch.randint(0, k, (n, h, w), device='cuda', generator=g)
    lab[torch.rand(n, h, w, device='cuda', generator=g) < 0.1] = 255
    cm = torch.zeros(k, k, device='cuda', dtype=torch.int64)
    for _ in range(2):
        check(L.evb_confusion_matrix(ptr(pred), ptr(lab), c_ll(n * h * w), c_int(k), ptr(cm), stream()), 'cm')
    torch.cuda.synchronize()
    valid = lab != 255
    ref = torch.bincount(lab[valid] * k + pred[valid].long(), minlength=k * k).view(k, k)
    assert torch.equal(cm, 2 * ref)


@pytest.mark.parametrize('entry,tile', [('evb_pack_weights_tiled', 64)])
def test_batched_weight_pack_bit_exact(entry, tile):
    """fp32 OIHW master -> bf16 packs [tap][CoP][CiP] and [tap][CiP][CoP] for a table of convolutions in one launch: 3x3 and
    1x1, channel counts that are not multiples of the tile (15, 147), zero padding, a Cin sub-range of a wider weight
    (w_off / w_ld), and a conv without a dgrad pack.  Pure rounding + movement: bit-exact."""
    L, check, ptr, stream = _L()
    g = _gen(21)
    cases = [  # (Co, Ci_total, k, ci_lo, ci_n, CoP, CiP, need_wb)
        (256, 256, 3, 0, 256, 256, 256, True), (15, 256, 1, 0, 256, 64, 256, True), (64, 147, 1, 0, 147, 64, 192, False),
        (16, 512, 3, 256, 256, 64, 256, True), (96, 40, 3, 0, 40, 128, 64, True)]
    rows, bmap, nblk, keep = [], [], 0, []
    for i, (co, cit, k, lo, cn, cop, cip, need_wb) in enumerate(cases):
        kk = k * k
        w = torch.randn(co, cit, k, k, device='cuda', generator=g)
        wf = torch.zeros(kk, cop, cip, device='cuda', dtype=torch.bfloat16)
        wb = torch.zeros(kk, cip, cop, device='cuda', dtype=torch.bfloat16) if need_wb else None
        nb = ((co + tile - 1) // tile) * ((cn + tile - 1) // tile)
        rows.append([w.data_ptr() + 4 * lo * kk, wf.data_ptr(), wb.data_ptr() if need_wb else 0, co, cn, kk, cop, cip, cip, cop,
                     nblk, cit * kk if lo or cn != cit else 0])
        bmap += [i] * nb
        nblk += nb
        keep.append((w, wf, wb))
    desc = torch.tensor(rows, dtype=torch.int64, device='cuda')
    bm = torch.tensor(bmap, dtype=torch.int32, device='cuda')
    half = nblk // 2   # two range launches cover the table
    check(getattr(L, entry)(ptr(desc), ptr(bm), c_int(0), c_int(half), stream()), entry)
    check(getattr(L, entry)(ptr(desc), ptr(bm), c_int(half), c_int(nblk - half), stream()), entry)
    torch.cuda.synchronize()
    for (co, cit, k, lo, cn, cop, cip, need_wb), (w, wf, wb) in zip(cases, keep):
        kk = k * k
        sub = w[:, lo:lo + cn].reshape(co, cn, kk).bfloat16()
        ref_f = torch.zeros(kk, cop, cip, device='cuda', dtype=torch.bfloat16)
        ref_f[:, :co, :cn] = sub.permute(2, 0, 1)
        assert torch.equal(wf, ref_f)
        if need_wb:
            ref_b = torch.zeros(kk, cip, cop, device='cuda', dtype=torch.bfloat16)
            ref_b[:, :cn, :co] = sub.permute(2, 1, 0)
            assert torch.equal(wb, ref_b)
