"""Index maps for the gather / canvas kernels of libevb200.so (``csrc/spatial.cu``).

A ``PixelMap`` describes how every output pixel (i, j) of a dihedral (flip / rot90 / transpose) + crop + constant-pad
chain reads one source pixel: ``src[A @ (i, j) + b]`` if that lies inside the clip rectangle, else the fill value.
The chain of torch ops the reference applies one after the other on the host (``torch.rot90`` / ``torch.flip`` /
slicing / ``F.pad``: ever/preprocess/thsegm.py:7-147, ever/preprocess/function.py:35-83,
ever/magic/transform/segm.py:9-72) composes into ONE map, i.e. one kernel launch per batch.

Everything here is integer host arithmetic (testable without a GPU); ``gather`` launches ``evb_pixel_gather``.
"""
import ctypes

import torch

from ._lib import check, lib, ptr, stream

c_int, c_ll = ctypes.c_int, ctypes.c_longlong


class PixelMap:
    """out[i, j] = src[a00*i + a01*j + b0, a10*i + a11*j + b1] if inside clip=(y0, y1, x0, x1) (source coordinates)
    else fill.  ``size`` = (H, W) of the output."""
    __slots__ = ('a', 'b', 'size', 'clip')

    def __init__(self, h, w):
        self.a, self.b, self.size, self.clip = (1, 0, 0, 1), (0, 0), (int(h), int(w)), (0, int(h), 0, int(w))

    def _then(self, a2, b2, size):
        """compose with a further op out2[p] = out1[a2 @ p + b2] whose output has ``size``"""
        a00, a01, a10, a11 = self.a
        c00, c01, c10, c11 = a2
        m = PixelMap(*size)
        m.a = (a00 * c00 + a01 * c10, a00 * c01 + a01 * c11, a10 * c00 + a11 * c10, a10 * c01 + a11 * c11)
        m.b = (a00 * b2[0] + a01 * b2[1] + self.b[0], a10 * b2[0] + a11 * b2[1] + self.b[1])
        m.clip = self.clip
        return m

    def _src(self, i, j):
        a00, a01, a10, a11 = self.a
        return a00 * i + a01 * j + self.b[0], a10 * i + a11 * j + self.b[1]

    # ---- the torch ops of the reference, on the two spatial axes
    def hflip(self):        # torch.flip(x, [W axis])
        h, w = self.size
        return self._then((1, 0, 0, -1), (0, w - 1), (h, w))

    def vflip(self):        # torch.flip(x, [H axis])
        h, w = self.size
        return self._then((-1, 0, 0, 1), (h - 1, 0), (h, w))

    def transpose(self):    # torch.transpose(x, H axis, W axis)
        h, w = self.size
        return self._then((0, 1, 1, 0), (0, 0), (w, h))

    def rot90(self, k=1):   # torch.rot90(x, k, [H axis, W axis])
        k = k % 4
        h, w = self.size
        if k == 0:
            return self._then((1, 0, 0, 1), (0, 0), (h, w))
        if k == 1:          # out[i, j] = x[j, W-1-i]
            return self._then((0, 1, -1, 0), (0, w - 1), (w, h))
        if k == 2:          # out[i, j] = x[H-1-i, W-1-j]
            return self._then((-1, 0, 0, -1), (h - 1, w - 1), (h, w))
        return self._then((0, -1, 1, 0), (h - 1, 0), (w, h))   # out[i, j] = x[H-1-j, i]

    def crop(self, ymin, xmin, ch, cw):
        """x[ymin:ymin+ch, xmin:xmin+cw]; pixels of the window that were never part of the image read ``fill``"""
        h, w = self.size
        if ymin < 0 or xmin < 0 or ymin + ch > h or xmin + cw > w:
            raise ValueError('crop window outside the image')
        m = self._then((1, 0, 0, 1), (ymin, xmin), (ch, cw))
        # shrink the clip rectangle to the window (in source coordinates) so that a later pad reads fill, not source
        p0, p1 = m._src(0, 0), m._src(ch - 1, cw - 1)
        y0, y1 = min(p0[0], p1[0]), max(p0[0], p1[0]) + 1
        x0, x1 = min(p0[1], p1[1]), max(p0[1], p1[1]) + 1
        c = self.clip
        m.clip = (max(c[0], y0), min(c[1], y1), max(c[2], x0), min(c[3], x1))
        return m

    def pad_to(self, nh, nw):
        """F.pad(x, [0, nw - W, 0, nh - H]) (bottom / right constant pad: th_pad_to_size, th_divisible_pad)"""
        h, w = self.size
        if nh < h or nw < w:
            raise ValueError('pad_to target smaller than the image')
        # the clip rectangle keeps describing the valid source region; the new border maps outside it only if the
        # current image is exactly the clip window -> enforce by cropping to the full current image first
        m = self.crop(0, 0, h, w)
        m.size = (int(nh), int(nw))
        return m

    def divisible_pad(self, d):
        h, w = self.size
        return self.pad_to(-(-h // d) * d, -(-w // d) * d)

    def inverse(self):
        """the map of the inverse dihedral op (no crop / pad in the chain): applying it to this map's output restores the
        source"""
        a00, a01, a10, a11 = self.a
        det = a00 * a11 - a01 * a10
        assert det in (1, -1)
        i00, i01, i10, i11 = a11 * det, -a01 * det, -a10 * det, a00 * det
        hs, ws = self.clip[1], self.clip[3]
        m = PixelMap(*self.size)
        m.a = (i00, i01, i10, i11)
        m.b = (-(i00 * self.b[0] + i01 * self.b[1]), -(i10 * self.b[0] + i11 * self.b[1]))
        m.size = (hs, ws)
        m.clip = (0, self.size[0], 0, self.size[1])
        return m

    def row(self, src_index=0, canvas_index=0):
        a00, a01, a10, a11 = self.a
        y0, y1, x0, x1 = self.clip
        return [a00, a01, self.b[0], a10, a11, self.b[1], int(src_index), y0, y1, x0, x1, int(canvas_index)]

    def shifted(self, dy, dx):
        """canvas -> tile map of a tile placed at (dy, dx) on a canvas: the same map read at (y - dy, x - dx); the clip
        rectangle (in canvas coordinates) becomes the tile's footprint [dy, dy + H) x [dx, dx + W)"""
        a00, a01, a10, a11 = self.a
        m = PixelMap(*self.size)
        m.a = self.a
        m.b = (self.b[0] - a00 * dy - a01 * dx, self.b[1] - a10 * dy - a11 * dx)
        m.clip = (dy, dy + self.size[0], dx, dx + self.size[1])
        return m

    def apply_reference(self, x, fill=0):
        """evaluate the map with torch indexing on a [H, W, ...] tensor (host-side check of the map arithmetic; the
        product path is ``gather``)"""
        h, w = self.size
        ii, jj = torch.meshgrid(torch.arange(h), torch.arange(w), indexing='ij')
        a00, a01, a10, a11 = self.a
        si = a00 * ii + a01 * jj + self.b[0]
        sj = a10 * ii + a11 * jj + self.b[1]
        y0, y1, x0, x1 = self.clip
        ok = (si >= y0) & (si < y1) & (sj >= x0) & (sj < x1)
        out = x[si.clamp(0, x.shape[0] - 1), sj.clamp(0, x.shape[1] - 1)].clone()
        out[~ok] = fill
        return out


def map_table(rows, device):
    return torch.tensor(rows, dtype=torch.int32, device=device).contiguous()


def gather(src, maps, out_hw, fill=0, out=None):
    """src: [Ns, Hs, Ws, ...] (any dtype, pixel-interleaved, contiguous, CUDA); maps: list of (PixelMap, src_index) or a
    prebuilt device table; returns [N, Ho, Wo, ...] with N = number of map rows.  One launch of evb_pixel_gather."""
    if not src.is_cuda:
        raise RuntimeError('gather runs on a CUDA device only (no CPU path)')
    src = src.contiguous()
    table = maps if torch.is_tensor(maps) else map_table([m.row(s) for m, s in maps], src.device)
    n = table.shape[0]
    ns, hs, ws = src.shape[:3]
    tail = tuple(src.shape[3:])
    elem = src.element_size()
    pix = elem
    for t in tail:
        pix *= t
    ho, wo = out_hw
    if out is None:
        out = torch.empty((n, ho, wo) + tail, dtype=src.dtype, device=src.device)
    if src.dtype.is_floating_point:
        fill_bits = torch.tensor([fill], dtype=src.dtype).view({2: torch.int16, 4: torch.int32, 8: torch.int64}[elem]).item()
    else:
        fill_bits = int(fill)
    check(lib().evb_pixel_gather(ptr(src), c_int(ns), c_int(hs), c_int(ws), c_int(pix), c_int(elem), c_ll(fill_bits),
                                 ptr(table), ptr(out), c_int(n), c_int(ho), c_int(wo), stream()), 'evb_pixel_gather')
    return out
