"""GPU batch augmentation for the training input pipeline: the reference's per-sample host
transforms -- THRandomRotate90k, THRandomHorizontalFlip, THRandomVerticalFlip, THRandomCrop
(ever/preprocess/thsegm.py:7-147) and THDivisiblePad (ever/preprocess/thcomm.py:67-88, function.py:35-83) -- applied to a
whole batch of raw uint8 HWC tiles and their label maps by ONE gather launch each (``evb_pixel_gather``): the chain of
rot90 / flip / crop / pad composes into a single index map per sample (``ever_b200.spatial.PixelMap``).

The random decisions are drawn on the host exactly the way the reference draws them (same numpy calls in the same order per
sample), so seeding numpy identically reproduces the reference's augmented batch bit for bit.  The output feeds the
engine's uint8 stem (normalisation fused into the im2col: THMeanStdNormalize, thcomm.py:47-64).
"""
import numpy as np
import torch

from .spatial import PixelMap, gather


class BatchAugment:
    """rotate90k=True: k ~ choice([0,1,2,3]); hflip_p / vflip_p: flip probabilities (None = off); crop_size=(h, w);
    size_divisor: bottom/right pad (image fill 0, label fill ``mask_pad_value``).  Order = the usual reference pipeline:
    rotate -> hflip -> vflip -> crop -> divisible pad."""

    def __init__(self, rotate90k=True, hflip_p=0.5, vflip_p=0.5, crop_size=None, size_divisor=None, mask_pad_value=255):
        self.rotate90k, self.hflip_p, self.vflip_p = rotate90k, hflip_p, vflip_p
        self.crop_size, self.size_divisor, self.mask_pad_value = crop_size, size_divisor, mask_pad_value

    def draw(self, h, w):
        """one sample's map (and the label-pad stage), consuming numpy's global RNG like the reference transforms do"""
        m = PixelMap(h, w)
        if self.rotate90k:
            k = int(np.random.choice([0, 1, 2, 3], 1)[0])               # thsegm.py:25
            if k:
                m = m.rot90(k)
        if self.hflip_p is not None and not (self.hflip_p < np.random.uniform()):   # thsegm.py:58
            m = m.hflip()
        if self.vflip_p is not None and not (self.vflip_p < np.random.uniform()):   # thsegm.py:91
            m = m.vflip()
        if self.crop_size is not None:
            ch, cw = self.crop_size
            ih, iw = m.size
            if ch > ih or cw > iw:                                      # thsegm.py:125-129: zero pad (image AND mask)
                m = m.pad_to(max(ih, ch), max(iw, cw))
                ih, iw = m.size
            ymin = int(np.random.randint(0, ih - ch + 1, 1)[0])         # thsegm.py:132-135
            xmin = int(np.random.randint(0, iw - cw + 1, 1)[0])
            m = m.crop(ymin, xmin, ch, cw)
        return m

    def __call__(self, images, masks=None):
        """images: uint8 [N, H, W, C] on the GPU; masks: [N, H, W] (uint8 / int64) or None.  Returns the augmented batch
        (and int64 labels, the dtype the loss kernels read)."""
        n, h, w = images.shape[:3]
        maps = [self.draw(h, w) for _ in range(n)]
        sizes = {m.size for m in maps}
        if len(sizes) != 1:
            raise ValueError('samples of one batch must come out with one size (non-square tiles need crop_size)')
        oh, ow = maps[0].size
        rows = [(m, i) for i, m in enumerate(maps)]
        out_i = gather(images, rows, (oh, ow), fill=0)
        out_m = None
        if masks is not None:
            out_m = gather(masks.long(), rows, (oh, ow), fill=0)        # RandomCrop pads masks with 0 too (thsegm.py:129)
        d = self.size_divisor
        if d and (oh % d or ow % d):                                    # THDivisiblePad: image 0, mask mask_pad_value
            pm = PixelMap(oh, ow).divisible_pad(d)
            prow = [(pm, i) for i in range(n)]
            out_i = gather(out_i, prow, pm.size, fill=0)
            if out_m is not None:
                out_m = gather(out_m, prow, pm.size, fill=self.mask_pad_value)
        return (out_i, out_m) if masks is not None else out_i
