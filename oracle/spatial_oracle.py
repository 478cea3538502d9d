"""ORACLE (test infrastructure, NOT product code): the reference's host-side spatial ops restated with the same torch /
numpy calls, for checking ever_b200.spatial / augment / infer.

  * sliding_window_oracle  -- ever/magic/bigimage/sliding_window.py:8-33 (line-by-line restatement; pinned against the
    real function in the build container and against tests/golden/sliding_window_boxes.json made from it);
  * augment_oracle         -- THRandomRotate90k / THRandomHorizontalFlip / THRandomVerticalFlip / THRandomCrop
    (ever/preprocess/thsegm.py:7-147) + THDivisiblePad (ever/preprocess/thcomm.py:67-88) applied per sample in that order,
    drawing from numpy's global RNG exactly as those classes do;
  * tta_oracle             -- ever/magic/transform/tta.py:11-23 over the transforms of ever/magic/transform/segm.py:9-72.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def sliding_window_oracle(input_size, kernel_size, stride):
    ih, iw = input_size
    kh, kw = (kernel_size, kernel_size) if isinstance(kernel_size, int) else kernel_size
    sh, sw = (stride, stride) if isinstance(stride, int) else stride
    assert ih > 0 and iw > 0 and kh > 0 and kw > 0 and sh > 0 and sw > 0
    kh = ih if kh > ih else kh
    kw = iw if kw > iw else kw
    num_rows = math.ceil((ih - kh) / sh) if math.ceil((ih - kh) / sh) * sh + kh >= ih else math.ceil((ih - kh) / sh) + 1
    num_cols = math.ceil((iw - kw) / sw) if math.ceil((iw - kw) / sw) * sw + kw >= iw else math.ceil((iw - kw) / sw) + 1
    x, y = np.meshgrid(np.arange(num_cols + 1), np.arange(num_rows + 1))
    xmin, ymin = (x * sw).ravel(), (y * sh).ravel()
    xoff = np.where(xmin + kw > iw, iw - xmin - kw, np.zeros_like(xmin))
    yoff = np.where(ymin + kh > ih, ih - ymin - kh, np.zeros_like(ymin))
    return np.stack([xmin + xoff, ymin + yoff, np.minimum(xmin + kw, iw), np.minimum(ymin + kh, ih)], axis=1)


def augment_oracle(image, mask, rotate90k=True, hflip_p=0.5, vflip_p=0.5, crop_size=None, size_divisor=None,
                   mask_pad_value=255):
    """image [H, W, C], mask [H, W] (CPU tensors); returns the augmented pair"""
    if rotate90k:
        k = int(np.random.choice([0, 1, 2, 3], 1)[0])
        if k:
            image, mask = torch.rot90(image, k, [0, 1]), torch.rot90(mask, k, [0, 1])
    if hflip_p is not None and not (hflip_p < np.random.uniform()):
        image, mask = torch.flip(image, [1]), torch.flip(mask, [1])
    if vflip_p is not None and not (vflip_p < np.random.uniform()):
        image, mask = torch.flip(image, [0]), torch.flip(mask, [0])
    if crop_size is not None:
        im_h, im_w, _ = image.shape
        c_h, c_w = crop_size
        pad_h, pad_w = c_h - im_h, c_w - im_w
        if pad_h > 0 or pad_w > 0:
            image = F.pad(image, [0, 0, 0, max(pad_w, 0), 0, max(pad_h, 0)], mode='constant', value=0)
            mask = F.pad(mask, [0, max(pad_w, 0), 0, max(pad_h, 0)], mode='constant', value=0)
        im_h, im_w, _ = image.shape
        ymin = int(np.random.randint(0, im_h - c_h + 1, 1)[0])
        xmin = int(np.random.randint(0, im_w - c_w + 1, 1)[0])
        image, mask = image[ymin:ymin + c_h, xmin:xmin + c_w, :], mask[ymin:ymin + c_h, xmin:xmin + c_w]
    if size_divisor:
        h, w = image.shape[:2]
        nh, nw = math.ceil(h / size_divisor) * size_divisor, math.ceil(w / size_divisor) * size_divisor
        image = F.pad(image, [0, 0, 0, nw - w, 0, nh - h], value=0)
        mask = F.pad(mask, [0, nw - w, 0, nh - h], value=mask_pad_value)
    return image.contiguous(), mask.contiguous()


TTA_OPS = {
    'Identity': (lambda x: x, lambda x: x),
    'Rotate90k1': (lambda x: torch.rot90(x, 1, [2, 3]), lambda x: torch.rot90(x, 3, [2, 3])),
    'Rotate90k2': (lambda x: torch.rot90(x, 2, [2, 3]), lambda x: torch.rot90(x, 2, [2, 3])),
    'Rotate90k3': (lambda x: torch.rot90(x, 3, [2, 3]), lambda x: torch.rot90(x, 1, [2, 3])),
    'HorizontalFlip': (lambda x: torch.flip(x, [3]), lambda x: torch.flip(x, [3])),
    'VerticalFlip': (lambda x: torch.flip(x, [2]), lambda x: torch.flip(x, [2])),
    'Transpose': (lambda x: torch.transpose(x, 2, 3), lambda x: torch.transpose(x, 2, 3)),
}


def scale_ops(size=None, scale_factor=None):
    """Scale, ever/magic/transform/segm.py:71-88: bilinear (align_corners=True) to size / by scale_factor, and back to the
    recorded input shape"""
    seen = {}

    def fwd(x):
        seen['hw'] = (x.shape[2], x.shape[3])
        return F.interpolate(x, size=size, scale_factor=scale_factor, mode='bilinear', align_corners=True)

    def inv(y):
        return F.interpolate(y, size=seen['hw'], mode='bilinear', align_corners=True)
    return fwd, inv


def _tta_ops(name):
    return scale_ops(**name[1]) if isinstance(name, tuple) else TTA_OPS[name]


def tta_oracle(model, image, names):
    """tta.py:11-23: outs = [model(t(image))]; outs = inverse transforms; sum(outs) / len(outs).
    The final division is evaluated on the host: it is an IEEE fp32 division there, while torch's CUDA kernel for
    tensor / python-scalar multiplies by the rounded reciprocal (1 ulp off for some elements); the product kernel divides."""
    ops = [_tta_ops(n) for n in names]   # names: keys of TTA_OPS, or ('Scale', dict(size=... | scale_factor=...))
    outs = [inv(model(fwd(image).contiguous())) for fwd, inv in ops]
    return (sum(outs).cpu() / len(outs)).to(outs[0].device)
