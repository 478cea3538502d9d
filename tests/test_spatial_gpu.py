"""GPU parity of the index-map kernels (csrc/spatial.cu) and of the inference / augmentation host code built on them.
Everything here is byte / fp32 movement plus fixed-order fp32 sums, so the gate is bit-exact (torch.equal) against the
torch ops the reference calls (oracle/spatial_oracle.py restates the reference pipelines)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


from _helpers import eval_r18_model as _model  # noqa: E402


@pytest.mark.parametrize('dtype,tail,fill', [(torch.uint8, (3,), 0), (torch.uint8, (), 7), (torch.int64, (), 255),
                                             (torch.float32, (), -1.5), (torch.bfloat16, (8,), 0.0), (torch.uint8, (200,), 0)])
def test_pixel_gather_bit_exact(dtype, tail, fill):
    from ever_b200.spatial import PixelMap, gather
    g = torch.Generator().manual_seed(1)
    n, h, w = 3, 37, 53
    if dtype.is_floating_point:
        src = torch.randn((n, h, w) + tail, generator=g).to(dtype)
    else:
        src = torch.randint(0, 200, (n, h, w) + tail, generator=g).to(dtype)
    p = PixelMap(h, w)
    chains = [p.rot90(1).hflip().crop(3, 5, 32, 24).pad_to(40, 32), p.vflip().transpose().crop(0, 0, 40, 32).pad_to(40, 32),
              p.rot90(3).crop(10, 2, 40, 32), p.rot90(2).pad_to(64, 64).crop(20, 30, 40, 32)]
    rows = [(chains[i % len(chains)], i % n) for i in range(7)]
    out = gather(src.cuda(), rows, (40, 32), fill=fill)
    torch.cuda.synchronize()
    for i, (m, s) in enumerate(rows):
        want = m.apply_reference(src[s], fill=fill)
        assert torch.equal(out[i].cpu(), want), (i, dtype)


@pytest.mark.parametrize('cfg', [dict(), dict(crop_size=(96, 64)), dict(crop_size=(160, 144), size_divisor=32),
                                 dict(rotate90k=False, hflip_p=1.0, vflip_p=None, crop_size=(100, 100), size_divisor=32)])
def test_batch_augment_equals_reference_pipeline(cfg):
    from ever_b200.augment import BatchAugment
    from oracle.spatial_oracle import augment_oracle
    g = torch.Generator().manual_seed(2)
    n, h, w = 8, 128, 128
    imgs = torch.randint(0, 256, (n, h, w, 3), generator=g, dtype=torch.uint8)
    masks = torch.randint(0, 7, (n, h, w), generator=g)
    np.random.seed(5)
    gi, gm = BatchAugment(**cfg)(imgs.cuda(), masks.cuda())
    torch.cuda.synchronize()
    np.random.seed(5)
    for i in range(n):
        wi, wm = augment_oracle(imgs[i], masks[i], **cfg)
        assert torch.equal(gi[i].cpu(), wi) and torch.equal(gm[i].cpu(), wm), i
    assert gm.dtype == torch.int64 and gi.dtype == torch.uint8


def test_canvas_accumulate_and_finalize_bit_exact():
    """overlapping tiles, dihedral inverse maps, two canvases, partial bounding box"""
    from ever_b200.infer import _canvas_accumulate, _canvas_finalize
    from ever_b200.spatial import PixelMap
    g = torch.Generator().manual_seed(4)
    k, h, w, hc, wc = 5, 24, 16, 64, 48
    prob = torch.rand(6, k, h, w, generator=g)
    canvas = torch.rand(2, k, hc, wc, generator=g)
    count = torch.zeros(2, hc, wc)
    place = [(0, 0, 0, None), (1, 10, 8, 'hflip'), (2, 30, 30, 'rot2'), (3, 10, 8, None), (4, 0, 0, 'vflip'), (5, 40, 32, None)]
    want_c, want_n = canvas.clone(), count.clone()
    rows = []
    for s, y0, x0, t in place:
        cb = s % 2
        m = PixelMap(h, w)
        m = dict(hflip=m.hflip(), vflip=m.vflip(), rot2=m.rot90(2)).get(t, m)
        inv = m.inverse().shifted(y0, x0)
        rows.append(inv.row(s, cb))
        tile = prob[s]
        tile = dict(hflip=torch.flip(tile, [2]), vflip=torch.flip(tile, [1]), rot2=torch.rot90(tile, 2, [1, 2])).get(t, tile)
        want_c[cb, :, y0:y0 + h, x0:x0 + w] += tile
        want_n[cb, y0:y0 + h, x0:x0 + w] += 1
    c_gpu, n_gpu = canvas.cuda(), count.cuda()
    _canvas_accumulate(prob.cuda(), rows, c_gpu, n_gpu, (0, hc, 0, wc))
    torch.cuda.synchronize()
    assert torch.equal(c_gpu.cpu(), want_c) and torch.equal(n_gpu.cpu(), want_n)
    pr, mk = _canvas_finalize(c_gpu, n_gpu, 0.0, True)
    torch.cuda.synchronize()
    want_p = torch.where(want_n[:, None] > 0, want_c / want_n[:, None].clamp(min=1), torch.zeros(()))
    assert torch.equal(pr.cpu(), want_p)
    assert torch.equal(mk.cpu().long(), want_p.argmax(dim=1))


@pytest.mark.parametrize('u8', [False, True])
def test_tta_equals_reference_formula(u8):
    """tta(): transformed copies -> model -> inverse -> sum / len, bit-equal to the reference formula evaluated with torch
    ops around the same model (ever/magic/transform/tta.py:11-23)"""
    from ever_b200.infer import HorizontalFlip, Identity, Rotate90k, Transpose, VerticalFlip, tta
    from oracle.spatial_oracle import tta_oracle
    model = _model()
    g = torch.Generator().manual_seed(6)
    if u8:
        x = torch.randint(0, 256, (2, 96, 128, 3), generator=g, dtype=torch.uint8).cuda()
    else:
        x = torch.randn(2, 3, 96, 128, generator=g).cuda()
    cfg = [Identity(), Rotate90k(1), Rotate90k(2), Rotate90k(3), HorizontalFlip(), VerticalFlip(), Transpose()]
    names = ['Identity', 'Rotate90k1', 'Rotate90k2', 'Rotate90k3', 'HorizontalFlip', 'VerticalFlip', 'Transpose']
    got, mask = tta(model, x, cfg, return_mask=True)
    if u8:   # the oracle transforms NCHW tensors: run the model on NHWC uint8 copies of them
        f = lambda t: model(t.permute(0, 2, 3, 1).contiguous())
        want = tta_oracle(f, x.permute(0, 3, 1, 2).contiguous(), names)
    else:
        want = tta_oracle(model, x, names)
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    assert torch.equal(mask.long(), want.argmax(dim=1))


@pytest.mark.parametrize('shape,size', [((3, 5, 96, 128), (48, 64)), ((2, 3, 64, 64), (128, 160)), ((1, 4, 37, 53), (61, 20)),
                                        ((2, 2, 32, 32), (32, 32)), ((1, 1, 9, 7), (1, 1))])
def test_resize_bilinear_equals_interpolate(shape, size):
    """evb_resize_bilinear_ac == F.interpolate(mode='bilinear', align_corners=True) (the Scale transform, segm.py:71-88):
    same source coordinates and weights; torch's kernel may contract its products into FMAs, hence 1e-6 rather than equal"""
    from ever_b200.infer import resize_bilinear
    g = torch.Generator().manual_seed(12)
    x = torch.randn(*shape, generator=g).cuda()
    want = F.interpolate(x, size=size, mode='bilinear', align_corners=True)
    got = resize_bilinear(x, size)
    torch.cuda.synchronize()
    assert got.shape == want.shape
    assert (got - want).abs().max().item() <= 1e-6 * max(1.0, want.abs().max().item())
    base = torch.rand_like(want)
    acc = resize_bilinear(x, size, out=base.clone(), accumulate=True)
    assert torch.equal(acc, base + got)


def test_tta_with_scale():
    """Scale in a TTA config: resize -> model -> resize back, summed with the index-map transforms in config order
    (tta.py:11-23).  The bf16 engine amplifies a last-bit difference of its input, so the expected value feeds the model the
    product's resized image (checked against F.interpolate to 1e-6 above) and uses torch for everything else: inverse
    resize, flips, the sum and the division."""
    from ever_b200.infer import HorizontalFlip, Identity, Scale, resize_bilinear, tta
    model = _model()
    g = torch.Generator().manual_seed(16)
    x = torch.randn(2, 3, 128, 128, generator=g).cuda()
    cfg = [Identity(), Scale(scale_factor=0.5), HorizontalFlip(), Scale(size=(160, 192)), Scale(scale_factor=1.25)]
    got = tta(model, x, cfg)
    back = lambda p: F.interpolate(p, size=(128, 128), mode='bilinear', align_corners=True)
    outs = [model(x), back(model(resize_bilinear(x, (64, 64)))), torch.flip(model(torch.flip(x, [3]).contiguous()), [3]),
            back(model(resize_bilinear(x, (160, 192)))), back(model(resize_bilinear(x, (160, 160))))]
    want = (sum(outs).cpu() / len(outs)).cuda()
    torch.cuda.synchronize()
    assert (got - want).abs().max().item() <= 2e-6
    with pytest.raises(NotImplementedError):
        tta(model, torch.zeros(1, 64, 64, 3, dtype=torch.uint8, device='cuda'), [Scale(scale_factor=0.5)])


@pytest.mark.parametrize('hw,tile,stride,batch,u8', [((300, 416), 256, 128, 4, True), ((200, 700), 256, 192, 3, False)])
def test_sliding_window_predictor_equals_host_accumulation(hw, tile, stride, batch, u8):
    """the GPU canvas path == the host loop a user of the reference's sliding_window writes
    (crop -> pad to /32 -> model -> out[..., window] += prob ; count += 1 ; out / count), same batches, bit-exact"""
    from ever_b200.infer import SlidingWindowPredictor
    model = _model()
    h, w = hw
    g = torch.Generator().manual_seed(8)
    if u8:
        img = torch.randint(0, 256, (h, w, 3), generator=g, dtype=torch.uint8).cuda()
        chw = img.permute(2, 0, 1)
    else:
        img = torch.randn(3, h, w, generator=g).cuda()
        chw = img
    sw = SlidingWindowPredictor(model, tile=tile, stride=stride, batch=batch)
    prob, mask = sw(img)
    boxes = sw.boxes(h, w)
    canvas = torch.zeros(5, h, w, device='cuda')
    count = torch.zeros(h, w, device='cuda')
    for s in range(0, len(boxes), batch):
        chunk = boxes[s:s + batch].tolist()
        crops = []
        for x0, y0, x1, y1 in chunk:
            c = chw[:, y0:y1, x0:x1]
            ph, pw = -(-(y1 - y0) // 32) * 32, -(-(x1 - x0) // 32) * 32
            crops.append(F.pad(c, [0, pw - (x1 - x0), 0, ph - (y1 - y0)]))
        xb = torch.stack(crops)
        out = model(xb.permute(0, 2, 3, 1).contiguous() if u8 else xb.contiguous())
        for i, (x0, y0, x1, y1) in enumerate(chunk):
            canvas[:, y0:y1, x0:x1] += out[i, :, :y1 - y0, :x1 - x0]
            count[y0:y1, x0:x1] += 1
    want = canvas / count
    torch.cuda.synchronize()
    assert float(count.min()) >= 1
    assert torch.equal(prob, want)
    assert torch.equal(mask.long(), want.argmax(dim=0))
