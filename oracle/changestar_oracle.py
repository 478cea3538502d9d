"""ORACLE (test infrastructure, NOT product code) -- ChangeStar / ChangeMixin.

PARITY UNPINNED: ChangeStar is NOT in /root/reference (only linked from its README.md:43-44).  This
is a restatement, from the published description of Z-Zheng/ChangeStar (ICCV 2021, "Change is Everywhere"), built only
from reference primitives (ever.module ConvBlock semantics ever/module/ops.py:45-60, Bf16compatible ops.py:152-166,
losses ever/module/loss.py:54-75,229-235) on top of the pinned FarSeg oracle.  It pins the B200 engine's ChangeStar
path to a plain-PyTorch statement of the same arithmetic, not to upstream code.

  features  = FarSeg decoder output (before the classifier) of both temporal images, run as one batch of 2N
  semantic  = FarSegHead classifier on the t1 features            -> CE + Dice (K >= 2) / BCE + Dice (K == 1)
  ChangeMixin.convs = [3x3 conv(2C->16, no bias) + BN + ReLU] + 3 x [3x3 conv(16->16) + BN + ReLU] + 3x3 conv(16->1)
                      + bilinear x4 (align_corners);  c12 = convs(cat(f1, f2)),  c21 = convs(cat(f2, f1))
  change    = BCE + sigmoid Dice of c12 and of c21 against the same binary change label (temporal symmetry)
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .farseg_oracle import FarSegOracle, _Fp32Around, bce_loss_oracle, dice_loss_oracle


class ChangeMixinOracle(nn.Module):
    def __init__(self, in_channels, inner_channels=16, num_convs=4, scale_factor=4.0):
        super().__init__()
        layers = [nn.Sequential(nn.Conv2d(in_channels, inner_channels, 3, 1, 1, bias=False),
                                nn.BatchNorm2d(inner_channels), nn.ReLU(True))]
        layers += [nn.Sequential(nn.Conv2d(inner_channels, inner_channels, 3, 1, 1, bias=False),
                                 nn.BatchNorm2d(inner_channels), nn.ReLU(True)) for _ in range(num_convs - 1)]
        layers += [nn.Conv2d(inner_channels, 1, 3, 1, 1), _Fp32Around(nn.UpsamplingBilinear2d(scale_factor=scale_factor))]
        self.convs = nn.Sequential(*layers)

    def forward(self, f1, f2):
        return self.convs(torch.cat([f1, f2], dim=1)), self.convs(torch.cat([f2, f1], dim=1))


class ChangeStarOracle(nn.Module):
    def __init__(self, resnet_type='resnet50', num_classes=1, decoder_channels=256, inner_channels=16):
        super().__init__()
        fs = FarSegOracle(resnet_type, num_classes, decoder_channels)
        self.en, self.head = fs.en, fs.head
        self.changemixin = ChangeMixinOracle(2 * decoder_channels, inner_channels, 4, 4.0)
        self.num_classes = num_classes

    def features(self, x):
        dec = self.head.fpn_decoder
        feats = self.en(x)
        ps = self.head.fpn(feats)
        scene = F.adaptive_avg_pool2d(feats[-1], 1)
        refined = self.head.fs_relation(scene, ps)
        inner = [blk(f) for blk, f in zip(dec.blocks, refined)]
        return sum(inner) / len(inner)

    def forward(self, x, y=None):
        n = x.shape[0]
        xs = x.view(n, 2, x.shape[1] // 2, x.shape[2], x.shape[3]).transpose(0, 1).reshape(2 * n, -1, x.shape[2], x.shape[3])
        f = self.features(xs)
        f1, f2 = f[:n], f[n:]
        seg = self.head.fpn_decoder.classifier(f1)
        c12, c21 = self.changemixin(f1, f2)
        if self.training:
            cls, chg = y['cls'], y['change']
            out = {}
            if self.num_classes == 1:
                out['bce_loss'] = bce_loss_oracle(seg, cls)
            else:
                out['ce_loss'] = F.cross_entropy(seg, cls.long(), ignore_index=255)
            out['dice_loss'] = dice_loss_oracle(seg, cls)
            for name, c in (('c12', c12), ('c21', c21)):
                out[name + '_bce_loss'] = bce_loss_oracle(c, chg)
                out[name + '_dice_loss'] = dice_loss_oracle(c, chg)
            return out
        return dict(seg=seg.sigmoid() if self.num_classes == 1 else seg.softmax(dim=1), change=c12.sigmoid())
