"""Inference around the eval path of the engine: sliding-window tiling of big images and test-time
augmentation, with the crops / transforms / accumulation done on the GPU by the index-map kernels (``csrc/spatial.cu``).

Reference pieces mirrored:
  * ``sliding_window(input_size, kernel_size, stride)``  ever/magic/bigimage/sliding_window.py:8-33 -- same boxes, same
    order, including its duplicated border boxes (the meshgrid runs one step past the last row / column);
  * ``tta(model, image, tta_config)`` / ``TestTimeAugmentation``  ever/magic/transform/tta.py:11-42 with the transforms of
    ever/magic/transform/segm.py:9-88 (Identity, Rotate90k, HorizontalFlip, VerticalFlip, Transpose, Scale): transformed copies
    -> model -> inverse transform -> ``sum(outs) / len(outs)`` (summed in transform order, then one true division --
    the same fp32 operation order as the reference, so equal per-transform outputs give bit-equal means).
The reference does the crops / flips / accumulation with host-side torch ops, one image at a time.
"""
import ctypes
import math

import numpy as np
import torch

from ._lib import check, lib, ptr, stream
from .spatial import PixelMap, gather, map_table

c_int, c_ll, c_float = ctypes.c_int, ctypes.c_longlong, ctypes.c_float


def sliding_window(input_size, kernel_size, stride):
    """boxes [n, 4] = (xmin, ymin, xmax, ymax) int64; restates ever/magic/bigimage/sliding_window.py:8-33"""
    ih, iw = input_size
    kh, kw = (kernel_size, kernel_size) if isinstance(kernel_size, int) else kernel_size
    sh, sw = (stride, stride) if isinstance(stride, int) else stride
    if min(ih, iw, kh, kw, sh, sw) <= 0:
        raise AssertionError('sizes must be positive')
    kh, kw = min(kh, ih), min(kw, iw)

    def steps(extent, k, s):
        n = math.ceil((extent - k) / s)
        return n if n * s + k >= extent else n + 1
    nrow, ncol = steps(ih, kh, sh), steps(iw, kw, sw)
    ys = np.repeat(np.arange(nrow + 1) * sh, ncol + 1)      # row-major: x runs fastest
    xs = np.tile(np.arange(ncol + 1) * sw, nrow + 1)
    x0 = np.where(xs + kw > iw, iw - kw, xs)                # windows past the border are shifted back inside
    y0 = np.where(ys + kh > ih, ih - kh, ys)
    return np.stack([x0, y0, np.minimum(xs + kw, iw), np.minimum(ys + kh, ih)], axis=1)


# ---------------------------------------------------------------------------- TTA transforms (segm.py:9-72)
class Identity:
    def map(self, h, w):
        return PixelMap(h, w)


class Rotate90k:
    def __init__(self, k=1):
        assert k in (1, 2, 3)
        self.k = k

    def map(self, h, w):
        return PixelMap(h, w).rot90(self.k)


class HorizontalFlip:
    def map(self, h, w):
        return PixelMap(h, w).hflip()


class VerticalFlip:
    def map(self, h, w):
        return PixelMap(h, w).vflip()


class Transpose:
    def map(self, h, w):
        return PixelMap(h, w).transpose()


class Scale:
    """bilinear resize with align_corners=True and back (segm.py:71-88); not an index map: runs ``evb_resize_bilinear_ac``"""

    def __init__(self, size=None, scale_factor=None):
        if (size is None) == (scale_factor is None):
            raise ValueError('only one of size or scale_factor should be defined')   # F.interpolate's own rule
        self.size, self.scale_factor = size, scale_factor


def _scaled_size(t, h, w):
    """output size of F.interpolate(size=..., scale_factor=...): floor(in * factor) computed in double"""
    if t.size is not None:
        return (t.size, t.size) if isinstance(t.size, int) else tuple(t.size)
    f = t.scale_factor
    fh, fw = (f, f) if isinstance(f, (int, float)) else f
    return int(math.floor(float(h) * fh)), int(math.floor(float(w) * fw))


def resize_bilinear(x, size, out=None, accumulate=False):
    """F.interpolate(x, size, mode='bilinear', align_corners=True) for fp32 NCHW on the GPU; ``out`` (+)= the result"""
    n, c, h, w = x.shape
    x = x.float().contiguous()
    if out is None:
        out = torch.empty((n, c, size[0], size[1]), dtype=torch.float32, device=x.device)
    check(lib().evb_resize_bilinear_ac(ptr(x), c_ll(n * c), c_int(h), c_int(w), ptr(out), c_int(size[0]), c_int(size[1]),
                                       c_int(1 if accumulate else 0), stream()), 'evb_resize_bilinear_ac')
    return out


def _as_map(t, h, w):
    """accepts the classes above or the reference's own transform objects (matched by class name)"""
    if hasattr(t, 'map'):
        return t.map(h, w)
    name = type(t).__name__
    if name == 'Identity':
        return PixelMap(h, w)
    if name == 'Rotate90k':
        return PixelMap(h, w).rot90(t.k)
    if name == 'HorizontalFlip':
        return PixelMap(h, w).hflip()
    if name == 'VerticalFlip':
        return PixelMap(h, w).vflip()
    if name == 'Transpose':
        return PixelMap(h, w).transpose()
    raise NotImplementedError('TTA transform %s (only the dihedral transforms of segm.py:9-72 run on the GPU path)' % name)


def _canvas_accumulate(prob, rows, canvas, count, bbox):
    n_t, k, h, w = prob.shape
    b, _, hc, wc = canvas.shape
    table = map_table(rows, prob.device)
    check(lib().evb_canvas_accumulate(ptr(prob), c_int(len(rows)), c_int(k), c_int(h), c_int(w), ptr(table), ptr(canvas),
                                      ptr(count), c_int(b), c_int(hc), c_int(wc), c_int(bbox[0]), c_int(bbox[1]),
                                      c_int(bbox[2]), c_int(bbox[3]), stream()), 'evb_canvas_accumulate')


def _canvas_finalize(canvas, count, uniform, want_mask):
    b, k, hc, wc = canvas.shape
    mask = torch.empty((b, hc, wc), dtype=torch.uint8, device=canvas.device) if want_mask else None
    check(lib().evb_canvas_finalize(ptr(canvas), ptr(count), c_float(uniform), c_int(b), c_int(k), c_ll(hc * wc),
                                    ptr(canvas), ptr(mask), stream()), 'evb_canvas_finalize')
    return canvas, mask


def _transformed_batch(x, maps):
    """x: float NCHW or uint8 NHWC batch on the GPU; maps[n]: PixelMap for image n (all with the same output size)"""
    ho, wo = maps[0].size
    if x.dtype == torch.uint8:                       # pixel-interleaved: one map row per image
        return gather(x, [(m, i) for i, m in enumerate(maps)], (ho, wo))
    n, c, h, w = x.shape                             # planar: one row per (image, channel) plane of 4-byte pixels
    planes = gather(x.float().reshape(n * c, h, w), [(m, i * c + ch) for i, m in enumerate(maps) for ch in range(c)],
                    (ho, wo))
    return planes.view(n, c, ho, wo)


@torch.no_grad()
def tta(model, image, tta_config, return_mask=False):
    """mean over the TTA transforms of the model's probabilities (ever/magic/transform/tta.py:11-23).
    image: float [N, C, H, W] or uint8 [N, H, W, C] on the GPU; model: an eval-mode FarSegB200."""
    if model.training:
        raise RuntimeError('tta() needs model.eval()')
    if not image.is_cuda:
        raise RuntimeError('tta() runs on a CUDA device only')
    u8 = image.dtype == torch.uint8
    n = image.shape[0]
    h, w = (image.shape[1], image.shape[2]) if u8 else (image.shape[2], image.shape[3])
    canvas = None
    for t in tta_config:
        scale = type(t).__name__ == 'Scale'
        if scale:
            if u8:
                raise NotImplementedError('Scale interpolates: it needs the float NCHW image, not uint8 pixels')
            prob = model(resize_bilinear(image, _scaled_size(t, h, w)))
        else:
            m = _as_map(t, h, w)
            prob = model(_transformed_batch(image, [m] * n))
        if isinstance(prob, dict):
            raise NotImplementedError('tta() handles single-output models')
        if canvas is None:
            canvas = torch.zeros((n, prob.shape[1], h, w), dtype=torch.float32, device=image.device)
        if scale:
            resize_bilinear(prob, (h, w), out=canvas, accumulate=True)
            continue
        inv = m.inverse().shifted(0, 0)   # canvas rows clip in canvas coordinates: the whole image
        _canvas_accumulate(prob.contiguous(), [inv.row(i, i) for i in range(n)], canvas, None, (0, h, 0, w))
    return _finalize_out(_canvas_finalize(canvas, None, float(len(tta_config)), return_mask), return_mask)


def _finalize_out(pm, return_mask):
    return pm if return_mask else pm[0]


class SlidingWindowPredictor:
    """Tile a big image with ``sliding_window`` boxes, run the eval path on batches of tiles and average the overlapping
    probabilities on a GPU canvas.  Windows (optionally each under several TTA transforms) are cut, transformed and padded to
    a multiple of ``divisor`` by one gather launch per batch; every batch's probabilities are added to the canvas in window
    order (deterministic), then ``canvas / count`` and the argmax mask are produced by one kernel."""

    def __init__(self, model, tile=512, stride=256, batch=8, tta_config=None, divisor=32, dedup=True):
        self.model, self.tile, self.stride, self.batch = model, tile, stride, int(batch)
        self.tta_config = list(tta_config) if tta_config else [Identity()]
        self.divisor, self.dedup = int(divisor), dedup

    def boxes(self, h, w):
        b = sliding_window((h, w), self.tile, self.stride)
        if self.dedup:   # the reference emits duplicated border boxes; averaging is unaffected by dropping the copies
            _, first = np.unique(b, axis=0, return_index=True)
            b = b[np.sort(first)]
        return b

    @torch.no_grad()
    def __call__(self, image, return_mask=True):
        """image: uint8 [H, W, C] (raw, normalised inside the engine) or float [C, H, W] (already normalised), host or GPU.
        Returns (prob [K, H, W] fp32, mask [H, W] uint8) on the GPU."""
        model = self.model
        if model.training:
            raise RuntimeError('SlidingWindowPredictor needs model.eval()')
        dev = next(model.parameters()).device
        image = image.to(dev, non_blocking=True)
        u8 = image.dtype == torch.uint8
        h, w = (image.shape[0], image.shape[1]) if u8 else (image.shape[1], image.shape[2])
        batch_src = image.unsqueeze(0)
        boxes = self.boxes(h, w)
        canvas = count = None
        d = self.divisor
        for t in self.tta_config:
            for s in range(0, len(boxes), self.batch):
                chunk = boxes[s:s + self.batch]
                maps, rows = [], []
                for i, (x0, y0, x1, y1) in enumerate(chunk.tolist()):
                    wh, ww = y1 - y0, x1 - x0
                    ph, pw = -(-wh // d) * d, -(-ww // d) * d
                    win = PixelMap(h, w).crop(y0, x0, wh, ww).pad_to(ph, pw)      # window, zero-padded to the divisor
                    tm = _as_map(t, ph, pw)
                    maps.append(_compose(win, tm))
                    inv = tm.inverse().shifted(y0, x0)
                    inv.clip = (y0, y1, x0, x1)                                   # only the un-padded window lands
                    rows.append(inv.row(i, 0))
                prob = model(_window_batch(batch_src, maps))
                if canvas is None:
                    canvas = torch.zeros((1, prob.shape[1], h, w), dtype=torch.float32, device=dev)
                    count = torch.zeros((1, h, w), dtype=torch.float32, device=dev)
                bbox = (int(chunk[:, 1].min()), int(chunk[:, 3].max()), int(chunk[:, 0].min()), int(chunk[:, 2].max()))
                _canvas_accumulate(prob.contiguous(), rows, canvas, count, bbox)
        prob, mask = _canvas_finalize(canvas, count, 0.0, return_mask)
        return (prob[0], mask[0]) if return_mask else prob[0]


def _compose(first, then):
    """map of `then` applied to the output of `first` (both PixelMap; `then` is a pure dihedral map of first's output)"""
    return first._then(then.a, then.b, then.size)


def _window_batch(src, maps):
    """src: [1, H, W, C] uint8 or [1, C, H, W] float; every map reads image 0"""
    ho, wo = maps[0].size
    if src.dtype == torch.uint8:
        return gather(src, [(m, 0) for m in maps], (ho, wo))
    _, c, h, w = src.shape
    planes = gather(src.float().reshape(c, h, w), [(m, ch) for m in maps for ch in range(c)], (ho, wo))
    return planes.view(len(maps), c, ho, wo)
