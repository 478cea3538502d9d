"""Host-side logic of the spatial path (no GPU): index-map arithmetic against the torch ops the reference calls, the
sliding-window boxes against the reference's golden boxes, the augmentation draws against the restated reference pipeline
(and against the real THRandom* classes when /root/reference is present)."""
import json
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from ever_b200.augment import BatchAugment
from ever_b200.infer import SlidingWindowPredictor, sliding_window
from ever_b200.spatial import PixelMap
from oracle.spatial_oracle import augment_oracle, sliding_window_oracle

HAVE_REF = os.path.isdir('/root/reference/ever')


def _x(h=5, w=7, c=3):
    return torch.arange(h * w * c).reshape(h, w, c)


def _same(m, ref, x, fill=-1):
    got = m.apply_reference(x, fill=fill)
    assert got.shape == ref.shape and torch.equal(got, ref)


def test_dihedral_maps_equal_torch_ops():
    x = _x()
    p = PixelMap(5, 7)
    _same(p.hflip(), torch.flip(x, [1]), x)
    _same(p.vflip(), torch.flip(x, [0]), x)
    _same(p.transpose(), x.transpose(0, 1), x)
    for k in range(4):
        _same(p.rot90(k), torch.rot90(x, k, [0, 1]), x)


def test_composed_chain_crop_pad():
    x = _x()
    p = PixelMap(5, 7)
    _same(p.rot90(1).hflip().vflip().crop(1, 2, 4, 3),
          torch.flip(torch.flip(torch.rot90(x, 1, [0, 1]), [1]), [0])[1:5, 2:5], x)
    _same(p.rot90(3).crop(2, 1, 4, 3).pad_to(8, 8), F.pad(torch.rot90(x, 3, [0, 1])[2:6, 1:4], [0, 0, 0, 5, 0, 4], value=-1), x)
    _same(p.pad_to(10, 10).crop(3, 4, 7, 6), F.pad(x, [0, 0, 0, 3, 0, 5], value=-1)[3:10, 4:10], x)
    _same(p.crop(1, 1, 3, 3).divisible_pad(4), F.pad(x[1:4, 1:4], [0, 0, 0, 1, 0, 1], value=-1), x)
    with pytest.raises(ValueError):
        p.crop(0, 0, 6, 7)


def test_inverse_restores_source():
    x = _x()
    p = PixelMap(5, 7)
    for m in (p.hflip(), p.vflip(), p.transpose(), p.rot90(1), p.rot90(2), p.rot90(3)):
        assert torch.equal(m.inverse().apply_reference(m.apply_reference(x)), x)


def test_sliding_window_matches_reference_golden(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, 'sliding_window_boxes.json')))
    assert len(cases) >= 6
    for c in cases:
        size = tuple(c['input_size'])
        k = tuple(c['kernel_size']) if isinstance(c['kernel_size'], list) else c['kernel_size']
        s = tuple(c['stride']) if isinstance(c['stride'], list) else c['stride']
        want = np.asarray(c['boxes'])
        assert np.array_equal(sliding_window(size, k, s), want)
        assert np.array_equal(sliding_window_oracle(size, k, s), want)


@pytest.mark.skipif(not HAVE_REF, reason='reference tree only exists in the build container')
def test_sliding_window_matches_real_reference_randomised():
    for p_ in ('/root/reference', os.path.join(os.path.dirname(__file__), 'golden', '_stubs')):
        if p_ not in sys.path:
            sys.path.insert(0, p_)
    from ever.magic.bigimage.sliding_window import sliding_window as ref
    rng = np.random.RandomState(0)
    for _ in range(200):
        size = (int(rng.randint(1, 3000)), int(rng.randint(1, 3000)))
        k = (int(rng.randint(1, 1100)), int(rng.randint(1, 1100)))
        s = (int(rng.randint(1, 900)), int(rng.randint(1, 900)))
        assert np.array_equal(sliding_window(size, k, s), ref(size, k, s)), (size, k, s)


def test_predictor_boxes_cover_image_once_deduplicated():
    sw = SlidingWindowPredictor(model=None, tile=512, stride=256)
    b = sw.boxes(1000, 1500)
    assert len(np.unique(b, axis=0)) == len(b)
    cover = np.zeros((1000, 1500), dtype=np.int32)
    for x0, y0, x1, y1 in b:
        cover[y0:y1, x0:x1] += 1
    assert cover.min() >= 1


@pytest.mark.parametrize('cfg', [dict(), dict(crop_size=(24, 20)), dict(crop_size=(40, 48), size_divisor=32),
                                 dict(rotate90k=False, hflip_p=1.0, vflip_p=None, size_divisor=32)])
def test_augment_draws_equal_restated_reference_pipeline(cfg):
    g = torch.Generator().manual_seed(3)
    imgs = torch.randint(0, 256, (6, 32, 32, 3), generator=g, dtype=torch.uint8)
    masks = torch.randint(0, 7, (6, 32, 32), generator=g)
    aug = BatchAugment(**cfg)
    np.random.seed(11)
    maps = [aug.draw(32, 32) for _ in range(6)]
    np.random.seed(11)
    for i, m in enumerate(maps):
        want_i, want_m = augment_oracle(imgs[i], masks[i], **{k: v for k, v in cfg.items() if k != 'size_divisor'})
        assert torch.equal(m.apply_reference(imgs[i], fill=0), want_i)
        assert torch.equal(m.apply_reference(masks[i], fill=0), want_m)


@pytest.mark.skipif(not HAVE_REF, reason='reference tree only exists in the build container')
def test_augment_oracle_equals_real_reference_transforms():
    for p_ in ('/root/reference', os.path.join(os.path.dirname(__file__), 'golden', '_stubs')):
        if p_ not in sys.path:
            sys.path.insert(0, p_)
    from ever.preprocess.thcomm import THDivisiblePad
    from ever.preprocess.thsegm import THRandomCrop, THRandomHorizontalFlip, THRandomRotate90k, THRandomVerticalFlip
    g = torch.Generator().manual_seed(5)
    img = torch.randint(0, 256, (48, 40, 3), generator=g, dtype=torch.uint8)
    msk = torch.randint(0, 7, (48, 40), generator=g)
    for seed in range(8):
        np.random.seed(seed)
        a, b = THRandomRotate90k()(img, msk)
        a, b = THRandomHorizontalFlip()(a, b)
        a, b = THRandomVerticalFlip()(a, b)
        a, b = THRandomCrop((36, 44))(a, b)
        # THDivisiblePad works on channel-first images (function.py:35-60)
        a2, b2 = THDivisiblePad(32)(a.permute(2, 0, 1), b)
        np.random.seed(seed)
        oa, ob = augment_oracle(img, msk, crop_size=(36, 44), size_divisor=32)
        assert torch.equal(a2.permute(1, 2, 0), oa) and torch.equal(b2, ob)


def test_random_op_chains_equal_torch_ops_property():
    """property test (hypothesis): any chain of rot90 / flips / transpose / crop / bottom-right pad composes into ONE
    PixelMap that reproduces the torch ops applied one after the other, including fill in padded regions that a later
    crop / flip / rotation moves around"""
    from hypothesis import given, settings, strategies as st

    op = st.one_of(st.tuples(st.just('rot'), st.integers(0, 3)), st.tuples(st.just('hflip'), st.just(0)),
                   st.tuples(st.just('vflip'), st.just(0)), st.tuples(st.just('T'), st.just(0)),
                   st.tuples(st.just('crop'), st.tuples(st.floats(0, 1), st.floats(0, 1), st.floats(0.2, 1), st.floats(0.2, 1))),
                   st.tuples(st.just('pad'), st.tuples(st.integers(0, 5), st.integers(0, 5))))

    @settings(max_examples=150, deadline=None)
    @given(h=st.integers(1, 9), w=st.integers(1, 9), ops=st.lists(op, min_size=1, max_size=6))
    def run(h, w, ops):
        x = torch.arange(1, h * w * 2 + 1).reshape(h, w, 2)
        m, ref = PixelMap(h, w), x
        for name, arg in ops:
            ch, cw = ref.shape[:2]
            if name == 'rot':
                m, ref = m.rot90(arg), torch.rot90(ref, arg, [0, 1])
            elif name == 'hflip':
                m, ref = m.hflip(), torch.flip(ref, [1])
            elif name == 'vflip':
                m, ref = m.vflip(), torch.flip(ref, [0])
            elif name == 'T':
                m, ref = m.transpose(), ref.transpose(0, 1)
            elif name == 'crop':
                fy, fx, fh, fw = arg
                nh, nw = max(1, int(round(fh * ch))), max(1, int(round(fw * cw)))
                y0, x0 = int(fy * (ch - nh)), int(fx * (cw - nw))
                m, ref = m.crop(y0, x0, nh, nw), ref[y0:y0 + nh, x0:x0 + nw]
            else:
                ph, pw = arg
                m, ref = m.pad_to(ch + ph, cw + pw), F.pad(ref, [0, 0, 0, pw, 0, ph], value=-7)
        got = m.apply_reference(x, fill=-7)
        assert got.shape == ref.shape and torch.equal(got, ref)
    run()
